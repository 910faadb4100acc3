// Data-parallel exchange over peer-mapped ("symmetric") memory: every buffer below is allocated by all ranks of one
// NVSwitch domain and mapped into every process, so a kernel simply dereferences `peer[r]` to read rank r's copy over
// NVLink.  The reference has no distributed code at all; this is the new capability the north
// star names -- dense-gradient all-reduce + row-sparse exchange of the touched embedding rows -- built so that a whole
// training step, exchange included, is ONE CUDA graph with no host-side collective call:
//   * peer_barrier_kernel      : flag barrier (release/acquire at system scope), epoch kept on the device so a
//                                captured graph can be replayed;
//   * allreduce_peers_kernel   : one-shot all-reduce of the flat dense-gradient bucket (5 MB): every rank sums all
//                                ranks' buckets in rank order -> identical bits everywhere;
//   * owner_keys_kernel        : the id half of the row-gradient exchange.  Entity tables are OWNED row-range-wise by
//                                the ranks; a rank combines only the (row id, gradient row) pairs of the rows it owns.
//                                This kernel reads every rank's emitted row ids in place and clamps the ids this rank
//                                does not own to the sort sentinel (they sort last and are dropped), after which the
//                                ordinary plan (stable radix sort + segmentation, layout_sort.cu) runs unchanged and
//                                segment_sum_peers_kernel sums the owned rows straight out of the peers' buffers:
//                                (N-1)/N of one rank's rows cross NVLink per GPU per step, instead of (N-1) x.
#include "common.cuh"

namespace mpqe {
namespace {

struct PeerPtrs {
  void* p[MPQE_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(int* addr, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* addr) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

// flags[r] is rank r's int32[MPQE_MAX_PEERS] flag array (peer-mapped); rank `rank` writes its epoch into slot `rank`
// of every rank's array and waits until every slot of its own array has reached the epoch.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerPtrs flags, int rank, int world,
                                                          int* __restrict__ epoch) {
  __shared__ int s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = *epoch + 1;
    *epoch = s_epoch;
  }
  __syncthreads();
  const int e = s_epoch;
  __threadfence_system();   // everything earlier kernels of this stream wrote is visible to the peers before the flag
  const int t = threadIdx.x;
  if (t < world) {
    st_release_sys(reinterpret_cast<int*>(flags.p[t]) + rank, e);
    const int* mine = reinterpret_cast<const int*>(flags.p[rank]) + t;
    // bounded spin: a protocol bug or a dead peer must surface as a trapped launch, never as a hung GPU
    unsigned long long spins = 0;
    while (ld_acquire_sys(mine) - e < 0) {
      if (++spins > (1ull << 24)) asm volatile("trap;");   // ~20 s
      __nanosleep(64);
    }
  }
  __syncthreads();
  __threadfence_system();
}

// out[i] = scale * (buf[0][i] + buf[1][i] + ... )   (float4 grid-stride; rank order => same bits on every rank)
__global__ void __launch_bounds__(256) allreduce_peers_kernel(const __grid_constant__ PeerPtrs bufs, int world,
                                                              int64_t n4, float scale, float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 v[MPQE_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < MPQE_MAX_PEERS; ++r)
      if (r < world) v[r] = __ldcg(reinterpret_cast<const float4*>(bufs.p[r]) + i);   // all loads in flight first
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < MPQE_MAX_PEERS; ++r)
      if (r < world) {
        acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w;
      }
    out[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
}

// Two-shot form for larger groups (every rank reads 2 (N-1)/N of a bucket instead of N-1 buckets):
// reduce-scatter: rank r sums float4 slice r of all ranks' buffers (rank order) IN PLACE into its own buffer's slice r
// (the peers read other slices of it, and only their owner writes a slice); after a barrier, all-gather: slice p is
// read from rank p's buffer.
__global__ void __launch_bounds__(256) reduce_slice_peers_kernel(const __grid_constant__ PeerPtrs bufs, int world, int rank,
                                                                 int64_t begin4, int64_t end4, float scale) {
  float4* mine = reinterpret_cast<float4*>(bufs.p[rank]);
  for (int64_t i = begin4 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < end4; i += (int64_t)gridDim.x * 256) {
    float4 v[MPQE_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < MPQE_MAX_PEERS; ++r)
      if (r < world) v[r] = __ldcg(reinterpret_cast<const float4*>(bufs.p[r]) + i);
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < MPQE_MAX_PEERS; ++r)
      if (r < world) {
        acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w;
      }
    mine[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
  }
}

__global__ void __launch_bounds__(256) gather_slices_peers_kernel(const __grid_constant__ PeerPtrs bufs, int world,
                                                                  int64_t n4, int64_t chunk4, float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    int owner = (int)(i / chunk4);
    if (owner >= world) owner = world - 1;
    out[i] = __ldcg(reinterpret_cast<const float4*>(bufs.p[owner]) + i);
  }
}

struct Ownership {
  int num_tables, rank, world;
  int64_t begin[MPQE_MAX_TABLES];   // first global row id of table t
  int64_t rows[MPQE_MAX_TABLES];
};

// owner of a table row: rows are cut into `world` equal ranges per table
__device__ __forceinline__ int owner_of(int64_t row, int64_t rows, int world) {
  const int64_t chunk = (rows + world - 1) / world;
  return (int)(row / chunk);
}

// keys[r * per_rank + i] = id if this rank owns it, else the sentinel `limit` (ids read in place from rank r's buffer)
__global__ void __launch_bounds__(256) owner_keys_kernel(const __grid_constant__ PeerPtrs ids,
                                                         const __grid_constant__ Ownership O, int64_t per_rank,
                                                         uint32_t* __restrict__ keys, int64_t limit) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= per_rank * O.world) return;
  const int r = (int)(i / per_rank);
  const int64_t id = reinterpret_cast<const int64_t*>(ids.p[r])[i - (int64_t)r * per_rank];
  uint32_t key = (uint32_t)limit;
  if (id >= 0 && id < limit) {
    for (int t = 0; t < O.num_tables; ++t)
      if (id >= O.begin[t] && id < O.begin[t] + O.rows[t]) {
        if (owner_of(id - O.begin[t], O.rows[t], O.world) == O.rank) key = (uint32_t)id;
        break;
      }
  }
  keys[i] = key;
}

}  // namespace

// defined in layout_sort.cu: the plan (sort + segmentation) over keys already narrowed into the sort's first buffer
int sparse_rows_plan_prepared(int64_t count, int64_t table_rows, int64_t* num_unique, void* workspace,
                              size_t workspace_bytes, void* stream);
uint32_t* sparse_rows_key_buffer(void* workspace, int64_t count);

}  // namespace mpqe

using namespace mpqe;

static int fill_peers(PeerPtrs& P, const void* const* host, int world, const char* what) {
  MPQE_CHECK_ARG(host != nullptr && world >= 1 && world <= MPQE_MAX_PEERS, "%s: world must be in [1,%d]", what,
                 MPQE_MAX_PEERS);
  for (int r = 0; r < MPQE_MAX_PEERS; ++r) P.p[r] = r < world ? const_cast<void*>(host[r]) : nullptr;
  for (int r = 0; r < world; ++r) MPQE_CHECK_ARG(P.p[r] != nullptr, "%s: null buffer of rank %d", what, r);
  return 0;
}

extern "C" int mpqe_peer_barrier(const void* const* peer_flags_host, int32_t rank, int32_t world, int32_t* epoch,
                                 void* stream) {
  PeerPtrs P;
  if (int rc = fill_peers(P, peer_flags_host, world, "mpqe_peer_barrier")) return rc;
  MPQE_CHECK_ARG(rank >= 0 && rank < world && epoch != nullptr, "mpqe_peer_barrier: bad argument");
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, rank, world, epoch);
  MPQE_CHECK_LAUNCH("peer_barrier_kernel");
  return 0;
}

extern "C" int mpqe_allreduce_peers(const void* const* peer_bufs_host, int32_t world, int64_t numel, float scale,
                                    float* out, void* stream) {
  PeerPtrs P;
  if (int rc = fill_peers(P, peer_bufs_host, world, "mpqe_allreduce_peers")) return rc;
  MPQE_CHECK_ARG(out != nullptr && numel >= 0 && numel % 4 == 0, "mpqe_allreduce_peers: numel must be a multiple of 4");
  if (numel == 0) return 0;
  int64_t blocks = (numel / 4 + 255) / 256;
  if (blocks > 4 * 148) blocks = 4 * 148;
  allreduce_peers_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, world, numel / 4, scale,
                                                                          reinterpret_cast<float4*>(out));
  MPQE_CHECK_LAUNCH("allreduce_peers_kernel");
  return 0;
}

static int64_t slice_chunk4(int64_t n4, int world) { return (n4 + world - 1) / world; }

extern "C" int mpqe_reduce_scatter_peers(const void* const* peer_bufs_host, int32_t world, int32_t rank, int64_t numel,
                                         float scale, void* stream) {
  PeerPtrs P;
  if (int rc = fill_peers(P, peer_bufs_host, world, "mpqe_reduce_scatter_peers")) return rc;
  MPQE_CHECK_ARG(rank >= 0 && rank < world && numel >= 0 && numel % 4 == 0, "mpqe_reduce_scatter_peers: bad argument");
  const int64_t n4 = numel / 4, chunk4 = slice_chunk4(n4, world);
  const int64_t begin4 = chunk4 * rank < n4 ? chunk4 * rank : n4;
  const int64_t end4 = rank == world - 1 ? n4 : (chunk4 * (rank + 1) < n4 ? chunk4 * (rank + 1) : n4);
  if (end4 <= begin4) return 0;
  int64_t blocks = (end4 - begin4 + 255) / 256;
  if (blocks > 2 * 148) blocks = 2 * 148;
  reduce_slice_peers_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, world, rank, begin4, end4, scale);
  MPQE_CHECK_LAUNCH("reduce_slice_peers_kernel");
  return 0;
}

extern "C" int mpqe_all_gather_peers(const void* const* peer_bufs_host, int32_t world, int64_t numel, float* out,
                                     void* stream) {
  PeerPtrs P;
  if (int rc = fill_peers(P, peer_bufs_host, world, "mpqe_all_gather_peers")) return rc;
  MPQE_CHECK_ARG(out != nullptr && numel >= 0 && numel % 4 == 0, "mpqe_all_gather_peers: bad argument");
  if (numel == 0) return 0;
  const int64_t n4 = numel / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 4 * 148) blocks = 4 * 148;
  gather_slices_peers_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P, world, n4, slice_chunk4(n4, world),
                                                                              reinterpret_cast<float4*>(out));
  MPQE_CHECK_LAUNCH("gather_slices_peers_kernel");
  return 0;
}

extern "C" int mpqe_sparse_rows_plan_owner(const void* const* peer_ids_host, int32_t world, int32_t rank,
                                           int64_t per_rank_count, const int64_t* table_begin_host,
                                           const int64_t* table_rows_host, int32_t num_tables, int64_t total_rows,
                                           int64_t* num_unique, void* workspace, size_t workspace_bytes, void* stream) {
  PeerPtrs P;
  if (int rc = fill_peers(P, peer_ids_host, world, "mpqe_sparse_rows_plan_owner")) return rc;
  MPQE_CHECK_ARG(rank >= 0 && rank < world && per_rank_count >= 1 && num_tables >= 1 && num_tables <= MPQE_MAX_TABLES &&
                     table_begin_host && table_rows_host && num_unique,
                 "mpqe_sparse_rows_plan_owner: bad argument");
  const int64_t count = (int64_t)world * per_rank_count;
  MPQE_CHECK_ARG(count < (1ll << 31) && total_rows >= 1 && total_rows < (1ll << 32),
                 "mpqe_sparse_rows_plan_owner: too many pairs / rows");
  MPQE_CHECK_ARG(workspace && workspace_bytes >= mpqe_sparse_rows_workspace_bytes(count),
                 "mpqe_sparse_rows_plan_owner: workspace too small");
  Ownership O;
  O.num_tables = num_tables;
  O.rank = rank;
  O.world = world;
  for (int t = 0; t < MPQE_MAX_TABLES; ++t) {
    O.begin[t] = t < num_tables ? table_begin_host[t] : 0;
    O.rows[t] = t < num_tables ? table_rows_host[t] : 0;
  }
  owner_keys_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      P, O, per_rank_count, sparse_rows_key_buffer(workspace, count), total_rows);
  MPQE_CHECK_LAUNCH("owner_keys_kernel");
  return sparse_rows_plan_prepared(count, total_rows, num_unique, workspace, workspace_bytes, stream);
}

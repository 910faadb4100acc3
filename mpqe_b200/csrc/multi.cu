// Multi-item versions of the HBM-bound row kernels: one launch serves every formula group of a training step.
//
// A step touches up to 7 query types x (<=3 anchor slots + <=3 variable slots + positive + negative rows); launching
// one small kernel per (group, slot) made the step launch-bound (166 launches / 3.6 ms in the first B200 profile,
// profiles/r01_launches_simt_v1.csv).  Here the per-(group, slot) work items travel in the kernel parameter block
// and every warp finds its item from a prefix over the item counts.  Arithmetic per row is identical to the
// single-item kernels in gather_score.cu (shared device code in rowops.cuh); reference call sites as listed there.
#include "rowops.cuh"

namespace mpqe {
namespace {

struct GatherLaunch {
  int n;
  mpqe_gather_item_t it[MPQE_MAX_GATHER_ITEMS];
};
struct MarginLaunch {
  int n;
  float margin;
  mpqe_margin_item_t it[MPQE_MAX_MARGIN_ITEMS];
};
struct ColsumLaunch {
  int n;
  float* partials;  // [total blocks][D]
  float* totals;    // [n][D]
  int* counter;     // [n] finished second-stage CTAs per destination (zeroed by the first stage)
  uint8_t leader[MPQE_MAX_COLSUM_ITEMS];   // first item with the same destination
  uint8_t group[MPQE_MAX_COLSUM_ITEMS];    // for a leader: number of items of its destination
  mpqe_colsum_item_t it[MPQE_MAX_COLSUM_ITEMS];
};

template <class Launch>
__device__ __forceinline__ bool find_item(const Launch& L, int64_t w, int& item, int64_t& local) {
  for (int i = 0; i < L.n; ++i) {
    const int64_t c = L.it[i].count;
    if (w < c) {
      item = i;
      local = w;
      return true;
    }
    w -= c;
  }
  return false;
}

__global__ void __launch_bounds__(ROW_THREADS) gather_fwd_multi_kernel(const __grid_constant__ GatherLaunch L) {
  const int lane = threadIdx.x & 31;
  int item;
  int64_t i;
  if (!find_item(L, (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5), item, i)) return;
  const mpqe_gather_item_t& T = L.it[item];
  const int64_t row = resolve_row(T.id2row, T.ids, T.ids_stride, i);
  float4 y;
  if (row < 0 || row >= T.table_rows) {
    y = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
  } else if (T.normalize) {
    const float nrm = normalize_row(table_of_row(T.table, T.peer_tables, T.peer_chunk, row), row, lane, y);
    if (T.norm != nullptr && lane == 0) T.norm[i] = nrm;
  } else {
    y = *reinterpret_cast<const float4*>(T.table + row * D + lane * 4);
  }
  *reinterpret_cast<float4*>(T.out + i * T.out_stride + lane * 4) = y;
}

__global__ void __launch_bounds__(ROW_THREADS) gather_bwd_multi_kernel(const __grid_constant__ GatherLaunch L) {
  const int lane = threadIdx.x & 31;
  int item;
  int64_t i;
  if (!find_item(L, (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5), item, i)) return;
  const mpqe_gather_item_t& T = L.it[item];
  const int64_t row = resolve_row(T.id2row, T.ids, T.ids_stride, i);
  float4 y;
  float nrm;
  if (T.norm != nullptr) {      // the forward kept y (its output) and ||row||: no second gather
    y = *reinterpret_cast<const float4*>(T.out + i * T.out_stride + lane * 4);
    nrm = T.norm[i];
  } else {
    nrm = normalize_row(table_of_row(T.table, T.peer_tables, T.peer_chunk, row), row, lane, y);
  }
  const float4 g = *reinterpret_cast<const float4*>(T.grad + i * T.grad_stride + lane * 4);
  *reinterpret_cast<float4*>(T.rows_out + i * D + lane * 4) = normalize_bwd(g, y, nrm);
  if (lane == 0) T.rows_id[i] = row + T.id_offset;
}

// ids only (one thread per entry): rows_id[i] = row + id_offset, exactly what the backward kernels emit -- lets the
// row-gradient sort (mpqe_sparse_rows_plan) start before any gradient exists
__global__ void __launch_bounds__(256) gather_ids_multi_kernel(const __grid_constant__ GatherLaunch L) {
  int item;
  int64_t i;
  if (!find_item(L, (int64_t)blockIdx.x * 256 + threadIdx.x, item, i)) return;
  const mpqe_gather_item_t& T = L.it[item];
  T.rows_id[i] = resolve_row(T.id2row, T.ids, T.ids_stride, i) + T.id_offset;
}

template <class Launch>
__device__ __forceinline__ bool find_query(const Launch& L, int64_t w, int& item, int64_t& local) {
  for (int i = 0; i < L.n; ++i) {
    const int64_t c = L.it[i].B;
    if (w < c) {
      item = i;
      local = w;
      return true;
    }
    w -= c;
  }
  return false;
}

__global__ void __launch_bounds__(ROW_THREADS) margin_fwd_multi_kernel(const __grid_constant__ MarginLaunch L) {
  const int lane = threadIdx.x & 31;
  int item;
  int64_t b;
  if (!find_query(L, (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5), item, b)) return;
  const mpqe_margin_item_t& T = L.it[item];
  const float4 qv = *reinterpret_cast<const float4*>(T.q + b * D + lane * 4);
  float4 yp, yn;
  const int64_t rp = resolve_row(T.id2row, T.ids_pos, 1, b), rn = resolve_row(T.id2row, T.ids_neg, 1, b);
  normalize_row(table_of_row(T.table, T.peer_tables, T.peer_chunk, rp), rp, lane, yp);
  normalize_row(table_of_row(T.table, T.peer_tables, T.peer_chunk, rn), rn, lane, yn);
  const float sp = cosine(qv, yp).score, sn = cosine(qv, yn).score;
  if (lane == 0) {
    if (T.score_pos != nullptr) T.score_pos[b] = sp;
    if (T.score_neg != nullptr) T.score_neg[b] = sn;
    T.hinge[b] = fmaxf(L.margin - (sp - sn), 0.f);
  }
}

// deterministic mean per item: fixed strided partials per thread + fixed tree (same order as the single-item kernel)
__global__ void __launch_bounds__(1024) margin_mean_multi_kernel(const __grid_constant__ MarginLaunch L) {
  __shared__ float s[1024];
  const mpqe_margin_item_t& T = L.it[blockIdx.x];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < T.B; i += 1024) acc += T.hinge[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) T.loss[0] = s[0] / (float)T.B;
}

__global__ void __launch_bounds__(ROW_THREADS, 5) margin_bwd_multi_kernel(const __grid_constant__ MarginLaunch L) {
  const int lane = threadIdx.x & 31;
  int item;
  int64_t b;
  if (!find_query(L, (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5), item, b)) return;
  const mpqe_margin_item_t& T = L.it[item];
  const int64_t B = T.B;
  const float4 qv = *reinterpret_cast<const float4*>(T.q + b * D + lane * 4);
  const int64_t rp = resolve_row(T.id2row, T.ids_pos, 1, b), rn = resolve_row(T.id2row, T.ids_neg, 1, b);
  float4 yp, yn;
  const float np_ = normalize_row(table_of_row(T.table, T.peer_tables, T.peer_chunk, rp), rp, lane, yp);
  const float nn_ = normalize_row(table_of_row(T.table, T.peer_tables, T.peer_chunk, rn), rn, lane, yn);
  const Cos cp = cosine(qv, yp), cn = cosine(qv, yn);
  const float active = (L.margin - (cp.score - cn.score)) >= 0.f ? 1.f : 0.f;
  const float g = active * T.grad_loss[0] / (float)B;
  float4 dqp, dyp, dqn, dyn;
  cosine_bwd(qv, yp, cp, -g, dqp, dyp);
  cosine_bwd(qv, yn, cn, g, dqn, dyn);
  *reinterpret_cast<float4*>(T.dq + b * D + lane * 4) =
      make_float4(dqp.x + dqn.x, dqp.y + dqn.y, dqp.z + dqn.z, dqp.w + dqn.w);
  *reinterpret_cast<float4*>(T.rows_out + b * D + lane * 4) = normalize_bwd(dyp, yp, np_);
  *reinterpret_cast<float4*>(T.rows_out + (B + b) * D + lane * 4) = normalize_bwd(dyn, yn, nn_);
  if (lane == 0) {
    T.rows_id[b] = rp + T.id_offset;
    T.rows_id[B + b] = rn + T.id_offset;
    if (T.hinge != nullptr) {   // fused forward + backward (mode 2): the forward outputs, same arithmetic
      if (T.score_pos != nullptr) T.score_pos[b] = cp.score;
      if (T.score_neg != nullptr) T.score_neg[b] = cn.score;
      T.hinge[b] = fmaxf(L.margin - (cp.score - cn.score), 0.f);
    }
  }
}

// ---- column sums of many sources in two launches ------------------------------------------------------------
constexpr int CS_ROWS = 128;  // rows per CTA (8 warps x 16 rows)

__device__ __forceinline__ int64_t cs_blocks(int64_t rows) { return (rows + CS_ROWS - 1) / CS_ROWS; }

__global__ void __launch_bounds__(256) colsum_partial_multi_kernel(const __grid_constant__ ColsumLaunch L) {
  __shared__ float4 red[8][32];
  if (blockIdx.x == 0 && threadIdx.x < L.n) L.counter[threadIdx.x] = 0;   // second stage's "last CTA" elections
  int64_t blk = blockIdx.x;
  int item = 0;
  for (; item < L.n - 1; ++item) {
    const int64_t nb = cs_blocks(L.it[item].rows);
    if (blk < nb) break;
    blk -= nb;
  }
  const mpqe_colsum_item_t& T = L.it[item];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = blk * CS_ROWS;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k = 0; k < CS_ROWS / 8; ++k) {  // fixed row order per warp, fixed warp order below: bit-reproducible
    const int64_t r = r0 + warp + 8 * k;
    if (r < T.rows) {
      const float4 v = *reinterpret_cast<const float4*>(T.src + r * T.stride + lane * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float4 s = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = red[w][lane];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(L.partials + (int64_t)blockIdx.x * D + lane * 4) = s;
  }
}

// Second stage: one CTA per item sums the item's block partials (in block order) into totals[item]; per destination,
// the CTA that finishes last folds the totals of that destination's items, in item order:
// dst = ((dst + t_i) + t_k) + ...  (leader[i] = first item with the same destination, group[i] = number of items of
// leader i's destination; both computed on the host) -- bit-reproducible, one launch, destinations fold in parallel.
__global__ void __launch_bounds__(128) colsum_finish_multi_kernel(const __grid_constant__ ColsumLaunch L) {
  __shared__ int s_last;
  const int item = blockIdx.x, tid = threadIdx.x;
  int64_t base = 0;
  for (int i = 0; i < item; ++i) base += cs_blocks(L.it[i].rows);
  const int64_t nb = cs_blocks(L.it[item].rows);
  const float* p = L.partials + base * D + tid;
  float s = 0.f;
  int64_t b = 0;
  for (; b + 8 <= nb; b += 8) {   // eight loads in flight, summed in block order
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[(b + k) * D];
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; b < nb; ++b) s += p[b * D];
  const int lead = L.leader[item], group = L.group[lead];
  float* dst = L.it[item].dst;
  if (group == 1) {               // the only item of its destination: no hand-over needed
    dst[tid] += s * L.it[item].scale;
    return;
  }
  L.totals[(int64_t)item * D + tid] = s * L.it[item].scale;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(L.counter + lead, 1) == group - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float t[MPQE_MAX_COLSUM_ITEMS];
  int m = 0;
  for (int k = lead; k < L.n; ++k)
    if (L.leader[k] == lead) t[m++] = __ldcg(L.totals + (int64_t)k * D + tid);   // independent loads first
  float acc = dst[tid];
  for (int k = 0; k < m; ++k) acc += t[k];
  dst[tid] = acc;
}

}  // namespace
}  // namespace mpqe

using namespace mpqe;

extern "C" int mpqe_gather_multi(const mpqe_gather_item_t* items_host, int32_t n, int32_t backward, void* stream) {
  MPQE_CHECK_ARG(items_host != nullptr && n >= 1 && n <= MPQE_MAX_GATHER_ITEMS, "mpqe_gather_multi: n must be in [1,%d]",
                 MPQE_MAX_GATHER_ITEMS);
  static thread_local GatherLaunch L;
  L.n = n;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    const mpqe_gather_item_t& T = items_host[i];
    MPQE_CHECK_ARG(T.table && T.ids && T.count >= 0 && T.ids_stride >= 0, "mpqe_gather_multi: item %d: bad argument", i);
    if (backward == 2)
      MPQE_CHECK_ARG(T.rows_id != nullptr, "mpqe_gather_multi: item %d: ids mode without rows_id", i);
    else if (backward)
      MPQE_CHECK_ARG(T.grad && T.rows_out && T.rows_id && T.grad_stride >= D && (T.norm == nullptr || T.out != nullptr),
                     "mpqe_gather_multi: item %d: bad bwd argument", i);
    else
      MPQE_CHECK_ARG(T.out && T.out_stride >= D, "mpqe_gather_multi: item %d: bad fwd argument", i);
    L.it[i] = T;
    total += T.count;
  }
  if (total == 0) return 0;
  if (backward == 2)
    gather_ids_multi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(L);
  else if (backward)
    gather_bwd_multi_kernel<<<row_blocks(total), ROW_THREADS, 0, (cudaStream_t)stream>>>(L);
  else
    gather_fwd_multi_kernel<<<row_blocks(total), ROW_THREADS, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("gather_multi_kernel");
  return 0;
}

extern "C" int mpqe_cosine_margin_multi(const mpqe_margin_item_t* items_host, int32_t n, float margin, int32_t backward,
                                        void* stream) {
  MPQE_CHECK_ARG(items_host != nullptr && n >= 1 && n <= MPQE_MAX_MARGIN_ITEMS,
                 "mpqe_cosine_margin_multi: n must be in [1,%d]", MPQE_MAX_MARGIN_ITEMS);
  static thread_local MarginLaunch L;
  L.n = n;
  L.margin = margin;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    const mpqe_margin_item_t& T = items_host[i];
    MPQE_CHECK_ARG(T.q && T.table && T.ids_pos && T.ids_neg && T.B >= 1, "mpqe_cosine_margin_multi: item %d: bad argument", i);
    if (backward)
      MPQE_CHECK_ARG(T.grad_loss && T.dq && T.rows_out && T.rows_id, "mpqe_cosine_margin_multi: item %d: bad bwd argument", i);
    if (backward != 1)
      MPQE_CHECK_ARG(T.hinge && T.loss, "mpqe_cosine_margin_multi: item %d: bad fwd argument", i);
    L.it[i] = T;
    total += T.B;
  }
  if (backward) {
    margin_bwd_multi_kernel<<<row_blocks(total), ROW_THREADS, 0, (cudaStream_t)stream>>>(L);
    MPQE_CHECK_LAUNCH("margin_bwd_multi_kernel");
    if (backward == 2) {   // fused: the gradient of the total wrt each loss is known up front (grad_loss)
      margin_mean_multi_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>(L);
      MPQE_CHECK_LAUNCH("margin_mean_multi_kernel");
    }
  } else {
    margin_fwd_multi_kernel<<<row_blocks(total), ROW_THREADS, 0, (cudaStream_t)stream>>>(L);
    MPQE_CHECK_LAUNCH("margin_fwd_multi_kernel");
    margin_mean_multi_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>(L);
    MPQE_CHECK_LAUNCH("margin_mean_multi_kernel");
  }
  return 0;
}

extern "C" size_t mpqe_colsum_multi_workspace_bytes(const mpqe_colsum_item_t* items_host, int32_t n) {
  size_t blocks = 0;
  for (int i = 0; i < n; ++i) blocks += (size_t)((items_host[i].rows + CS_ROWS - 1) / CS_ROWS);
  return (blocks + (size_t)n + 1) * D * sizeof(float);
}

extern "C" int mpqe_colsum_multi(const mpqe_colsum_item_t* items_host, int32_t n, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  MPQE_CHECK_ARG(items_host != nullptr && n >= 1 && n <= MPQE_MAX_COLSUM_ITEMS, "mpqe_colsum_multi: n must be in [1,%d]",
                 MPQE_MAX_COLSUM_ITEMS);
  MPQE_CHECK_ARG(workspace && workspace_bytes >= mpqe_colsum_multi_workspace_bytes(items_host, n),
                 "mpqe_colsum_multi: workspace too small");
  static thread_local ColsumLaunch L;
  L.n = n;
  int64_t blocks = 0;
  for (int i = 0; i < n; ++i) {
    const mpqe_colsum_item_t& T = items_host[i];
    MPQE_CHECK_ARG(T.src && T.dst && T.rows >= 1 && T.stride >= D && T.stride % 4 == 0,
                   "mpqe_colsum_multi: item %d: bad argument", i);
    L.it[i] = T;
    blocks += (T.rows + CS_ROWS - 1) / CS_ROWS;
  }
  L.partials = (float*)workspace;
  L.totals = L.partials + blocks * D;
  L.counter = reinterpret_cast<int*>(L.totals + (int64_t)n * D);   // the spare row of the workspace
  for (int i = 0; i < n; ++i) {
    int lead = i;
    for (int k = 0; k < i; ++k)
      if (items_host[k].dst == items_host[i].dst) {
        lead = k;
        break;
      }
    L.leader[i] = (uint8_t)lead;
  }
  for (int i = 0; i < n; ++i) L.group[i] = 0;
  for (int i = 0; i < n; ++i) ++L.group[L.leader[i]];
  colsum_partial_multi_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("colsum_partial_multi_kernel");
  colsum_finish_multi_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("colsum_finish_multi_kernel");
  return 0;
}

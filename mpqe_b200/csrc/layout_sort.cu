// Integer kernels: batch layout, relation-sorted edge layout, and the row-sparse gradient combine.
// All results are bit-exact and bit-reproducible (no floating-point atomics, stable sorts, ordered sums).
//
// Reference call sites (under /root/reference/mpqe/):
//   data_utils.py:394-405 + PyG Batch.from_data_list     -> build_query_graph
//   (new, north star) stable sort of edge_type           -> relation_sort
//   autograd embedding_dense_backward    -> sparse_rows_combine / scatter_rows (row-sparse)
#include "common.cuh"

namespace mpqe {
namespace {

__global__ void build_query_graph_kernel(int n, int E, int64_t B, int4 src, int4 dst, longlong4 rel,
                                         int64_t* __restrict__ edge_index, int64_t* __restrict__ edge_type,
                                         int64_t* __restrict__ batch) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nE = B * E;
  if (i < nE) {
    const int64_t b = i / E;
    const int e = (int)(i - b * E);
    const int s = e == 0 ? src.x : e == 1 ? src.y : e == 2 ? src.z : src.w;
    const int t = e == 0 ? dst.x : e == 1 ? dst.y : e == 2 ? dst.z : dst.w;
    const int64_t r = e == 0 ? rel.x : e == 1 ? rel.y : e == 2 ? rel.z : rel.w;
    edge_index[i] = b * n + s;
    edge_index[nE + i] = b * n + t;
    edge_type[i] = r;
  }
  if (i < B * n) batch[i] = i / n;
}

// ---- stable LSD radix sort: digits of up to 10 bits, one WARP per chunk of 256 keys --------------------------
// Per pass: (1) per-chunk digit histogram, (2) exclusive scan of the [bin][chunk] table, done as one warp per bin
// row + one small block over the bin totals, (3) stable scatter (lane order inside a round via match_any, rounds and
// chunks in order).  Small chunks keep >= 500 warps busy at the ~1e5 keys of a training step.
constexpr int SORT_CHUNK = 256;
constexpr int SORT_WARPS = 8;
constexpr int MAX_BINS = 1024;

__device__ __forceinline__ unsigned digit_of(uint32_t key, int shift, unsigned mask) { return (key >> shift) & mask; }

__global__ void __launch_bounds__(SORT_WARPS * 32) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                     int shift, int bins, int nchunks,
                                                                     int32_t* __restrict__ hist) {
  extern __shared__ int sort_smem[];   // [SORT_WARPS][bins]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* cnt = sort_smem + w * bins;
  const int chunk = blockIdx.x * SORT_WARPS + w;
  for (int b = lane; b < bins; b += 32) cnt[b] = 0;
  __syncwarp();
  if (chunk < nchunks) {
    const int64_t i0 = (int64_t)chunk * SORT_CHUNK;
#pragma unroll
    for (int r = 0; r < SORT_CHUNK; r += 32) {
      const int64_t i = i0 + r + lane;
      if (i < n) atomicAdd(&cnt[digit_of(keys[i], shift, bins - 1)], 1);
    }
    __syncwarp();
    for (int b = lane; b < bins; b += 32) hist[(int64_t)b * nchunks + chunk] = cnt[b];
  }
}

// one warp per bin: exclusive scan of the bin's per-chunk counts, total to bin_total[bin]
__global__ void __launch_bounds__(256) radix_row_scan_kernel(int32_t* __restrict__ hist, int bins, int nchunks,
                                                             int32_t* __restrict__ bin_total) {
  const int lane = threadIdx.x & 31;
  const int bin = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (bin >= bins) return;
  int32_t* row = hist + (int64_t)bin * nchunks;
  int run = 0;
  for (int c0 = 0; c0 < nchunks; c0 += 32) {
    const int c = c0 + lane;
    const int v = c < nchunks ? row[c] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (c < nchunks) row[c] = run + incl - v;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) bin_total[bin] = run;
}

// single block: exclusive scan of the (<= 1024) bin totals, in place
__global__ void __launch_bounds__(MAX_BINS) radix_bin_scan_kernel(int32_t* __restrict__ bin_total, int bins) {
  __shared__ int s[MAX_BINS];
  const int t = threadIdx.x;
  const int v = t < bins ? bin_total[t] : 0;
  s[t] = v;
  __syncthreads();
  for (int o = 1; o < MAX_BINS; o <<= 1) {
    const int a = t >= o ? s[t - o] : 0;
    __syncthreads();
    s[t] += a;
    __syncthreads();
  }
  if (t < bins) bin_total[t] = s[t] - v;
}

__global__ void __launch_bounds__(SORT_WARPS * 32) radix_scatter_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int64_t n, int shift, int bins,
    int nchunks, const int32_t* __restrict__ hist, const int32_t* __restrict__ bin_base,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  extern __shared__ int sort_smem[];   // [SORT_WARPS][bins]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* run = sort_smem + w * bins;
  const int chunk = blockIdx.x * SORT_WARPS + w;
  if (chunk >= nchunks) return;
  for (int b = lane; b < bins; b += 32) run[b] = bin_base[b] + hist[(int64_t)b * nchunks + chunk];
  __syncwarp();
  const int64_t i0 = (int64_t)chunk * SORT_CHUNK;
#pragma unroll 1
  for (int r = 0; r < SORT_CHUNK; r += 32) {
    const int64_t i = i0 + r + lane;
    const bool ok = i < n;
    const uint32_t key = ok ? keys_in[i] : 0u;
    const unsigned dg = ok ? digit_of(key, shift, bins - 1) : 0x10000u + lane;  // inactive lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, dg);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int pos = 0;
    if (ok) pos = run[dg] + rank;
    __syncwarp();
    if (ok && rank == 0) run[dg] += __popc(peers);
    __syncwarp();
    if (ok) {
      keys_out[pos] = key;
      vals_out[pos] = vals_in != nullptr ? vals_in[i] : (uint32_t)i;
    }
  }
}

// keys >= limit (padding entries of an earlier combine) are clamped to the sentinel `limit`
__global__ void narrow_keys_kernel(const int64_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out,
                                   int64_t limit) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)(in[i] < limit ? in[i] : limit);
}

// the sentinel keys sort last and form one segment that must not count as a unique row
__global__ void drop_sentinel_kernel(const uint32_t* __restrict__ sorted, int64_t n, uint32_t sentinel,
                                     int64_t* __restrict__ num_unique) {
  if (sorted[n - 1] >= sentinel) *num_unique -= 1;
}

__global__ void widen_vals_kernel(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i];
}

// seg_offsets[r] = first position in the sorted key array with key >= r  (binary search; r in [0, R])
__global__ void segment_offsets_kernel(const uint32_t* __restrict__ sorted, int64_t n, int R,
                                       int64_t* __restrict__ offsets) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < (uint32_t)r) lo = mid + 1; else hi = mid;
  }
  offsets[r] = lo;
}

struct SortBuffers {
  uint32_t *k0, *v0, *k1, *v1;
  int32_t *hist, *bin_total;
  int nchunks;
  size_t bytes;
};

SortBuffers carve_sort(void* ws, int64_t n) {
  SortBuffers s;
  s.nchunks = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
  if (s.nchunks < 1) s.nchunks = 1;
  size_t off = 0;
  char* p = (char*)ws;
  const size_t arr = align_up((size_t)(n > 0 ? n : 1) * sizeof(uint32_t), 256);
  s.k0 = (uint32_t*)(p + off); off += arr;
  s.v0 = (uint32_t*)(p + off); off += arr;
  s.k1 = (uint32_t*)(p + off); off += arr;
  s.v1 = (uint32_t*)(p + off); off += arr;
  s.hist = (int32_t*)(p + off); off += align_up((size_t)MAX_BINS * s.nchunks * sizeof(int32_t), 256);
  s.bin_total = (int32_t*)(p + off); off += align_up((size_t)MAX_BINS * sizeof(int32_t), 256);
  s.bytes = off;
  return s;
}

// number of passes for keys of `key_bits` bits with digits of at most `max_digit_bits` bits
int sort_passes(int key_bits, int max_digit_bits) {
  if (key_bits < 1) key_bits = 1;
  return (key_bits + max_digit_bits - 1) / max_digit_bits;
}

// sorts k0 by the low `key_bits` bits; result ends in (*rk, *rv).  max_digit_bits = 10 minimises the passes;
// 7 keeps the per-CTA shared memory at 4 KB so that the sort can share an SM with a resident tcgen05 CTA when it
// runs on a second stream (mpqe_sparse_rows_plan).
int radix_sort(SortBuffers& s, int64_t n, int key_bits, cudaStream_t st, uint32_t** rk, uint32_t** rv,
               int max_digit_bits = 10) {
  if (key_bits < 1) key_bits = 1;
  const int passes = sort_passes(key_bits, max_digit_bits);
  const size_t smem = (size_t)SORT_WARPS * (1u << ((key_bits + passes - 1) / passes)) * sizeof(int);
  const int digit_bits = (key_bits + passes - 1) / passes;
  const int bins = 1 << digit_bits;
  uint32_t *ki = s.k0, *vi = nullptr, *ko = s.k1, *vo = s.v1;
  const int blocks = (s.nchunks + SORT_WARPS - 1) / SORT_WARPS;
  for (int p = 0; p < passes; ++p) {
    radix_hist_kernel<<<blocks, SORT_WARPS * 32, smem, st>>>(ki, n, digit_bits * p, bins, s.nchunks, s.hist);
    MPQE_CHECK_LAUNCH("radix_hist_kernel");
    radix_row_scan_kernel<<<(bins + 7) / 8, 256, 0, st>>>(s.hist, bins, s.nchunks, s.bin_total);
    MPQE_CHECK_LAUNCH("radix_row_scan_kernel");
    radix_bin_scan_kernel<<<1, MAX_BINS, 0, st>>>(s.bin_total, bins);
    MPQE_CHECK_LAUNCH("radix_bin_scan_kernel");
    radix_scatter_kernel<<<blocks, SORT_WARPS * 32, smem, st>>>(ki, vi, n, digit_bits * p, bins, s.nchunks, s.hist,
                                                            s.bin_total, ko, vo);
    MPQE_CHECK_LAUNCH("radix_scatter_kernel");
    uint32_t* tk = ki; ki = ko; ko = tk;
    uint32_t* tv = (vi == nullptr) ? s.v0 : vi; vi = vo; vo = tv;
  }
  *rk = ki;
  *rv = vi;
  return 0;
}

int bits_for(int64_t max_key_exclusive) {
  int bits = 1;
  while (bits < 32 && (1ll << bits) < max_key_exclusive) ++bits;
  return bits;
}

// ---- row-sparse combine -----------------------------------------------------------------------------------
// Three-kernel exclusive scan (int32) of n entries in blocks of SCAN_BLOCK; total written to *total.
constexpr int SCAN_BLOCK = 2048;   // 256 threads x 8 entries
constexpr int SCAN_MAX_BLOCKS = 1024;

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* total) {
  __shared__ int warp_sum[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sum[w] = incl;
  __syncthreads();
  int base = 0, all = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int s = warp_sum[k];
    if (k < w) base += s;
    all += s;
  }
  __syncthreads();
  *total = all;
  return base + incl - v;
}

__global__ void __launch_bounds__(256) scan_block_sums_kernel(const int32_t* __restrict__ data, int64_t n,
                                                              int32_t* __restrict__ block_sum) {
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 8;
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (i0 + k < n) s += data[i0 + k];
  int total;
  block_exclusive_scan_256(s, &total);
  if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_MAX_BLOCKS) scan_sums_kernel(int32_t* __restrict__ block_sum, int nblk,
                                                                   int64_t* __restrict__ total) {
  __shared__ int s[SCAN_MAX_BLOCKS];
  const int t = threadIdx.x;
  const int v = t < nblk ? block_sum[t] : 0;
  s[t] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_MAX_BLOCKS; o <<= 1) {
    const int a = t >= o ? s[t - o] : 0;
    __syncthreads();
    s[t] += a;
    __syncthreads();
  }
  if (t < nblk) block_sum[t] = s[t] - v;
  if (total != nullptr && t == SCAN_MAX_BLOCKS - 1) *total = s[t];
}

__global__ void __launch_bounds__(256) scan_apply_kernel(int32_t* __restrict__ data, int64_t n,
                                                         const int32_t* __restrict__ block_off) {
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 8;
  int v[8];
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = i0 + k < n ? data[i0 + k] : 0;
    s += v[k];
  }
  int total;
  int run = block_off[blockIdx.x] + block_exclusive_scan_256(s, &total);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (i0 + k < n) data[i0 + k] = run;
    run += v[k];
  }
}

// exclusive scan of data[0..n) in place; block_sum is scratch for ceil(n / SCAN_BLOCK) ints
int exclusive_scan(int32_t* data, int64_t n, int32_t* block_sum, int64_t* total, cudaStream_t st) {
  const int64_t nblk = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  MPQE_CHECK_ARG(nblk <= SCAN_MAX_BLOCKS, "scan of %lld entries exceeds the supported %d", (long long)n,
                 SCAN_BLOCK * SCAN_MAX_BLOCKS);
  scan_block_sums_kernel<<<(unsigned)nblk, 256, 0, st>>>(data, n, block_sum);
  MPQE_CHECK_LAUNCH("scan_block_sums_kernel");
  scan_sums_kernel<<<1, SCAN_MAX_BLOCKS, 0, st>>>(block_sum, (int)nblk, total);
  MPQE_CHECK_LAUNCH("scan_sums_kernel");
  scan_apply_kernel<<<(unsigned)nblk, 256, 0, st>>>(data, n, block_sum);
  MPQE_CHECK_LAUNCH("scan_apply_kernel");
  return 0;
}

__global__ void head_flags_kernel(const uint32_t* __restrict__ sorted, int64_t n, int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
}

// after the exclusive scan, uid[i] = (#heads before i); a head at i has unique index uid[i]
__global__ void segment_starts_kernel(const uint32_t* __restrict__ sorted, const int32_t* __restrict__ uid, int64_t n,
                                      int32_t* __restrict__ seg_start) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || sorted[i] != sorted[i - 1]) seg_start[uid[i]] = (int32_t)i;
}

// Ordered segmented sum: unique row u = sum of the gradient rows of its pairs, in ascending pair index (the sort is
// stable): a fixed summation order.  A warp serves SEG_PW consecutive unique rows: their segment bounds come from one
// load, their first pair indices from one load, and the SEG_PW first rows are in flight together -- most rows have one
// or two pairs, and one row per warp meant four dependent loads (count, bounds, pair index, row) for every 512 bytes.
// Further pairs of a segment are requested four at a time (a remote row is a ~2 us round trip over NVLink).
// PEERS: the pairs of `world` ranks are read IN PLACE from their buffers (peer-mapped memory): pair i lives in rank
// i / per_rank, row i % per_rank -- gather and sum are one kernel, every remote row crosses NVLink once and is never
// staged in local memory.
struct PeerRows {
  const float* rows[MPQE_MAX_PEERS];
};
constexpr int SEG_PW = 4;

template <bool PEERS>
__device__ __forceinline__ float4 pair_row(const float* rows, const PeerRows& P, uint32_t per_rank, uint32_t idx, int lane) {
  if (PEERS) {
    const uint32_t r = idx / per_rank, local = idx - r * per_rank;
    return __ldcg(reinterpret_cast<const float4*>(P.rows[r] + (int64_t)local * D + lane * 4));
  }
  return *reinterpret_cast<const float4*>(rows + (int64_t)idx * D + lane * 4);
}

template <bool PEERS>
__global__ void __launch_bounds__(256) segment_sum_kernel(const uint32_t* __restrict__ sorted_key,
                                                          const uint32_t* __restrict__ sorted_val,
                                                          const int32_t* __restrict__ seg_start,
                                                          const int64_t* __restrict__ num_unique, int64_t n,
                                                          const float* __restrict__ rows,
                                                          const __grid_constant__ PeerRows P, uint32_t per_rank,
                                                          int64_t* __restrict__ unique_ids,
                                                          float* __restrict__ unique_rows, int64_t pad_id,
                                                          uint32_t sentinel, float scale, int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t u0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * SEG_PW;
  const int64_t limit = capacity < n ? capacity : n;
  if (u0 >= limit) return;
  const int64_t nu = *num_unique;
  // the sentinel (padding) keys sort last as one more segment: the last real row must stop where it starts
  const int64_t nu_all = nu + (sorted_key[n - 1] >= sentinel ? 1 : 0);
  int64_t bound = n;       // lane l: start of segment u0 + l (= end of segment u0 + l - 1)
  if (lane <= SEG_PW && u0 + lane < nu_all) bound = seg_start[u0 + lane];
  uint32_t first = 0, key = 0;
  if (lane < SEG_PW && u0 + lane < nu) {
    first = sorted_val[bound];
    key = sorted_key[bound];
    unique_ids[u0 + lane] = (int64_t)key;
  } else if (lane < SEG_PW && u0 + lane < limit) {
    unique_ids[u0 + lane] = pad_id;   // padding entries: a zero row with id `pad_id` (fixed-size consumers, no host sync)
  }
  float4 acc[SEG_PW];
#pragma unroll
  for (int k = 0; k < SEG_PW; ++k) {
    const uint32_t idx = __shfl_sync(0xffffffffu, first, k);
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u0 + k < nu) {
      const float4 v = pair_row<PEERS>(rows, P, per_rank, idx, lane);
      acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
    }
  }
#pragma unroll
  for (int k = 0; k < SEG_PW; ++k) {
    const int64_t i0 = __shfl_sync(0xffffffffu, bound, k), i1 = __shfl_sync(0xffffffffu, bound, k + 1);
    if (u0 + k >= limit) break;
    if (u0 + k < nu) {
      for (int64_t i = i0 + 1; i < i1; i += 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i + j < i1) v[j] = pair_row<PEERS>(rows, P, per_rank, sorted_val[i + j], lane);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i + j < i1) {
            acc[k].x += v[j].x; acc[k].y += v[j].y; acc[k].z += v[j].z; acc[k].w += v[j].w;
          }
      }
    }
    *reinterpret_cast<float4*>(unique_rows + (u0 + k) * D + lane * 4) =
        make_float4(acc[k].x * scale, acc[k].y * scale, acc[k].z * scale, acc[k].w * scale);
  }
}

__global__ void __launch_bounds__(256) scatter_rows_kernel(const int64_t* __restrict__ ids,
                                                           const float* __restrict__ rows,
                                                           const int64_t* __restrict__ num, int64_t max_count,
                                                           float* __restrict__ dense, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t u = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nu = num != nullptr ? *num : max_count;
  if (u >= nu || u >= max_count) return;
  float4 v = *reinterpret_cast<const float4*>(rows + u * D + lane * 4);
  float4* dst = reinterpret_cast<float4*>(dense + ids[u] * D + lane * 4);
  if (accumulate) {
    const float4 o = *dst;
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  *dst = v;
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace
}  // namespace mpqe

using namespace mpqe;

extern "C" int mpqe_build_query_graph(int32_t n, int32_t E, const int32_t* tmpl_src_host, const int32_t* tmpl_dst_host,
                                      const int64_t* tmpl_rel_host, int64_t B, int64_t* edge_index,
                                      int64_t* edge_type, int64_t* batch, void* stream) {
  MPQE_CHECK_ARG(n >= 1 && n <= MPQE_MAX_SLOTS && E >= 1 && E <= 4 && B >= 1, "mpqe_build_query_graph: bad shape");
  MPQE_CHECK_ARG(tmpl_src_host && tmpl_dst_host && tmpl_rel_host && edge_index && edge_type && batch,
                 "mpqe_build_query_graph: null argument");
  int s[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
  long long r[4] = {0, 0, 0, 0};
  for (int e = 0; e < E; ++e) {
    s[e] = tmpl_src_host[e];
    t[e] = tmpl_dst_host[e];
    r[e] = tmpl_rel_host[e];
    MPQE_CHECK_ARG(s[e] >= 0 && s[e] < n && t[e] >= 0 && t[e] < n, "mpqe_build_query_graph: edge %d out of range", e);
  }
  const int64_t work = B * (E > n ? E : n);
  build_query_graph_kernel<<<blocks_for(work, 256), 256, 0, (cudaStream_t)stream>>>(
      n, E, B, make_int4(s[0], s[1], s[2], s[3]), make_int4(t[0], t[1], t[2], t[3]),
      make_longlong4(r[0], r[1], r[2], r[3]), edge_index, edge_type, batch);
  MPQE_CHECK_LAUNCH("build_query_graph_kernel");
  return 0;
}

extern "C" size_t mpqe_relation_sort_workspace_bytes(int64_t num_edges, int32_t num_relations) {
  (void)num_relations;
  return carve_sort(nullptr, num_edges).bytes;
}

extern "C" int mpqe_relation_sort(const int64_t* edge_type, int64_t num_edges, int32_t num_relations, int64_t* perm,
                                  int64_t* seg_offsets, void* workspace, size_t workspace_bytes, void* stream) {
  MPQE_CHECK_ARG(edge_type && perm && seg_offsets && num_edges >= 1 && num_relations >= 1 && num_edges < (1ll << 31),
                 "mpqe_relation_sort: bad argument");
  SortBuffers s = carve_sort(workspace, num_edges);
  MPQE_CHECK_ARG(workspace && workspace_bytes >= s.bytes, "mpqe_relation_sort: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  narrow_keys_kernel<<<blocks_for(num_edges, 256), 256, 0, st>>>(edge_type, num_edges, s.k0, (int64_t)1 << 31);
  MPQE_CHECK_LAUNCH("narrow_keys_kernel");
  uint32_t *rk, *rv;
  if (radix_sort(s, num_edges, bits_for(num_relations), st, &rk, &rv)) return 2;
  widen_vals_kernel<<<blocks_for(num_edges, 256), 256, 0, st>>>(rv, num_edges, perm);
  MPQE_CHECK_LAUNCH("widen_vals_kernel");
  segment_offsets_kernel<<<blocks_for(num_relations + 1, 256), 256, 0, st>>>(rk, num_edges, num_relations, seg_offsets);
  MPQE_CHECK_LAUNCH("segment_offsets_kernel");
  return 0;
}

extern "C" size_t mpqe_sparse_rows_workspace_bytes(int64_t count) {
  const size_t n = (size_t)(count > 0 ? count : 1);
  return carve_sort(nullptr, count).bytes + 2 * align_up(n * sizeof(int32_t), 256) + SCAN_MAX_BLOCKS * sizeof(int32_t);
}

namespace {
constexpr int PLAN_DIGIT_BITS = 7;
struct CombineBuffers {
  SortBuffers s;
  int32_t *uid, *seg_start, *block_sum;
  uint32_t *rk, *rv;   // where the sorted keys / source positions end up (depends on the pass count only)
};
CombineBuffers carve_combine(void* workspace, int64_t count, int64_t table_rows, int digit_bits) {
  CombineBuffers c;
  c.s = carve_sort(workspace, count);
  c.uid = (int32_t*)((char*)workspace + c.s.bytes);
  c.seg_start = (int32_t*)((char*)c.uid + align_up((size_t)count * sizeof(int32_t), 256));
  c.block_sum = (int32_t*)((char*)c.seg_start + align_up((size_t)count * sizeof(int32_t), 256));
  const int passes = sort_passes(bits_for(table_rows + 1), digit_bits);
  c.rk = (passes & 1) ? c.s.k1 : c.s.k0;
  c.rv = (passes & 1) ? c.s.v1 : c.s.v0;
  return c;
}
}  // namespace

// The part of the combine that needs only the row ids: stable sort, segment heads, number of distinct rows.  It can
// run (on another stream) while the gradient rows are still being computed; mpqe_sparse_rows_apply then sums them.
// sort + segmentation of `count` keys already narrowed (uint32, sentinel = table_rows) into the sort's first buffer
static int plan_sorted(int64_t count, int64_t table_rows, int64_t* num_unique, void* workspace, void* stream,
                       int digit_bits) {
  cudaStream_t st = (cudaStream_t)stream;
  CombineBuffers c = carve_combine(workspace, count, table_rows, digit_bits);
  uint32_t *rk, *rv;
  if (radix_sort(c.s, count, bits_for(table_rows + 1), st, &rk, &rv, digit_bits)) return 2;
  MPQE_CHECK_ARG(rk == c.rk && rv == c.rv, "mpqe_sparse_rows_plan: internal buffer parity mismatch");
  head_flags_kernel<<<blocks_for(count, 256), 256, 0, st>>>(rk, count, c.uid);
  MPQE_CHECK_LAUNCH("head_flags_kernel");
  if (exclusive_scan(c.uid, count, c.block_sum, num_unique, st)) return 2;
  drop_sentinel_kernel<<<1, 1, 0, st>>>(rk, count, (uint32_t)table_rows, num_unique);
  MPQE_CHECK_LAUNCH("drop_sentinel_kernel");
  segment_starts_kernel<<<blocks_for(count, 256), 256, 0, st>>>(rk, c.uid, count, c.seg_start);
  MPQE_CHECK_LAUNCH("segment_starts_kernel");
  return 0;
}

static int sparse_rows_plan(const int64_t* rows_id, int64_t count, int64_t table_rows, int64_t* num_unique,
                            void* workspace, size_t workspace_bytes, void* stream, int digit_bits) {
  MPQE_CHECK_ARG(rows_id && num_unique && count >= 1 && count < (1ll << 31) && table_rows >= 1 &&
                     table_rows < (1ll << 32),
                 "mpqe_sparse_rows_plan: bad argument");
  MPQE_CHECK_ARG(workspace && workspace_bytes >= mpqe_sparse_rows_workspace_bytes(count),
                 "mpqe_sparse_rows_plan: workspace too small");
  CombineBuffers c = carve_combine(workspace, count, table_rows, digit_bits);
  narrow_keys_kernel<<<blocks_for(count, 256), 256, 0, (cudaStream_t)stream>>>(rows_id, count, c.s.k0, table_rows);
  MPQE_CHECK_LAUNCH("narrow_keys_kernel");
  return plan_sorted(count, table_rows, num_unique, workspace, stream, digit_bits);
}

// entry points for peer.cu (owner-filtered keys are written straight into the sort's first key buffer)
namespace mpqe {
uint32_t* sparse_rows_key_buffer(void* workspace, int64_t count) { return carve_sort(workspace, count).k0; }
int sparse_rows_plan_prepared(int64_t count, int64_t table_rows, int64_t* num_unique, void* workspace,
                              size_t workspace_bytes, void* stream) {
  (void)workspace_bytes;
  return plan_sorted(count, table_rows, num_unique, workspace, stream, PLAN_DIGIT_BITS);
}
}  // namespace mpqe

static int sparse_rows_apply(const float* rows, int64_t count, int64_t table_rows, int64_t pad_id, float scale,
                             int64_t* unique_ids, float* unique_rows, const int64_t* num_unique, void* workspace,
                             size_t workspace_bytes, void* stream, int digit_bits) {
  MPQE_CHECK_ARG(rows && unique_ids && unique_rows && num_unique && count >= 1 && count < (1ll << 31) &&
                     table_rows >= 1 && table_rows < (1ll << 32),
                 "mpqe_sparse_rows_apply: bad argument");
  MPQE_CHECK_ARG(workspace && workspace_bytes >= mpqe_sparse_rows_workspace_bytes(count),
                 "mpqe_sparse_rows_apply: workspace too small");
  CombineBuffers c = carve_combine(workspace, count, table_rows, digit_bits);
  segment_sum_kernel<false><<<blocks_for((count + SEG_PW - 1) / SEG_PW, 8), 256, 0, (cudaStream_t)stream>>>(
      c.rk, c.rv, c.seg_start, num_unique, count, rows, PeerRows(), 1u, unique_ids, unique_rows, pad_id,
      (uint32_t)table_rows, scale, count);
  MPQE_CHECK_LAUNCH("segment_sum_kernel");
  return 0;
}

// two-phase form: 7-bit digits (4 KB of shared memory per sort CTA) so that the plan can share SMs with the resident
// tcgen05 CTAs of the stream it overlaps; the one-call form uses 10-bit digits (fewest passes)
extern "C" int mpqe_sparse_rows_plan(const int64_t* rows_id, int64_t count, int64_t table_rows, int64_t* num_unique,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  return sparse_rows_plan(rows_id, count, table_rows, num_unique, workspace, workspace_bytes, stream, PLAN_DIGIT_BITS);
}

extern "C" int mpqe_sparse_rows_apply(const float* rows, int64_t count, int64_t table_rows, int64_t pad_id, float scale,
                                      int64_t* unique_ids, float* unique_rows, const int64_t* num_unique,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  return sparse_rows_apply(rows, count, table_rows, pad_id, scale, unique_ids, unique_rows, num_unique, workspace,
                           workspace_bytes, stream, PLAN_DIGIT_BITS);
}

extern "C" int mpqe_sparse_rows_apply_peers(const float* const* peer_rows_host, int32_t world, int64_t per_rank_count,
                                            int64_t table_rows, int64_t pad_id, float scale, int64_t* unique_ids,
                                            float* unique_rows, int64_t out_capacity, const int64_t* num_unique,
                                            void* workspace, size_t workspace_bytes, void* stream) {
  MPQE_CHECK_ARG(peer_rows_host && world >= 1 && world <= MPQE_MAX_PEERS && per_rank_count >= 1 && out_capacity >= 1,
                 "mpqe_sparse_rows_apply_peers: world must be in [1,%d]", MPQE_MAX_PEERS);
  const int64_t count = (int64_t)world * per_rank_count;
  MPQE_CHECK_ARG(unique_ids && unique_rows && num_unique && count < (1ll << 31) && table_rows >= 1 &&
                     table_rows < (1ll << 32),
                 "mpqe_sparse_rows_apply_peers: bad argument");
  MPQE_CHECK_ARG(workspace && workspace_bytes >= mpqe_sparse_rows_workspace_bytes(count),
                 "mpqe_sparse_rows_apply_peers: workspace too small");
  PeerRows P;
  for (int r = 0; r < MPQE_MAX_PEERS; ++r) P.rows[r] = r < world ? peer_rows_host[r] : nullptr;
  for (int r = 0; r < world; ++r)
    MPQE_CHECK_ARG(P.rows[r] != nullptr, "mpqe_sparse_rows_apply_peers: null buffer of rank %d", r);
  CombineBuffers c = carve_combine(workspace, count, table_rows, PLAN_DIGIT_BITS);
  const int64_t cap = out_capacity < count ? out_capacity : count;
  segment_sum_kernel<true><<<blocks_for((cap + SEG_PW - 1) / SEG_PW, 8), 256, 0, (cudaStream_t)stream>>>(
      c.rk, c.rv, c.seg_start, num_unique, count, nullptr, P, (uint32_t)per_rank_count, unique_ids, unique_rows, pad_id,
      (uint32_t)table_rows, scale, cap);
  MPQE_CHECK_LAUNCH("segment_sum_kernel<peers>");
  return 0;
}

extern "C" int mpqe_sparse_rows_combine(const int64_t* rows_id, const float* rows, int64_t count, int64_t table_rows,
                                        int64_t pad_id, int64_t* unique_ids, float* unique_rows, int64_t* num_unique,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  if (sparse_rows_plan(rows_id, count, table_rows, num_unique, workspace, workspace_bytes, stream, 10)) return 1;
  return sparse_rows_apply(rows, count, table_rows, pad_id, 1.0f, unique_ids, unique_rows, num_unique, workspace,
                           workspace_bytes, stream, 10);
}

extern "C" int mpqe_scatter_rows(const int64_t* ids, const float* rows, const int64_t* num, int64_t max_count,
                                 float* dense, int32_t accumulate, void* stream) {
  MPQE_CHECK_ARG(ids && rows && dense && max_count >= 0, "mpqe_scatter_rows: bad argument");
  if (max_count == 0) return 0;
  scatter_rows_kernel<<<blocks_for(max_count, 8), 256, 0, (cudaStream_t)stream>>>(ids, rows, num, max_count, dense,
                                                                               accumulate);
  MPQE_CHECK_LAUNCH("scatter_rows_kernel");
  return 0;
}

// HBM-bound row kernels: embedding gather + L2 normalise, cosine scoring, margin loss, max readout, rank counts.
// One warp owns one 128-float row (one 16-byte load per lane, fully coalesced 512 B).
//
// Reference call sites (under /root/reference/mpqe/):
//   data_utils.py:35 + utils.py:22 + encoders.py:41-43  -> gather_normalize_{fwd,bwd}
//   model.py:421                                        -> broadcast_rows
//   model.py:383-385 (torch_scatter.scatter_max)        -> max_readout_{fwd,bwd}
//   model.py:451-460, 483-485                           -> cosine_scores*, cosine_margin_{fwd,bwd}
//   utils.py:25-32 (scipy percentileofscore 'rank')     -> rank_counts_ragged
#include "rowops.cuh"

namespace mpqe {
namespace {

__global__ void __launch_bounds__(ROW_THREADS) gather_normalize_fwd_kernel(
    const float* __restrict__ table, int64_t table_rows, const int64_t* __restrict__ id2row,
    const int64_t* __restrict__ ids, int64_t ids_stride, int64_t count, float* __restrict__ out, int64_t out_stride,
    float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (i >= count) return;
  const int64_t row = resolve_row(id2row, ids, ids_stride, i);
  float4 y;
  float nrm;
  if (row < 0 || row >= table_rows) {  // id without a row in this mode's table: poison, do not fault
    y = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
    nrm = CUDART_NAN_F;
  } else {
    nrm = normalize_row(table, row, lane, y);
  }
  *reinterpret_cast<float4*>(out + i * out_stride + lane * 4) = y;
  if (inv_norm != nullptr && lane == 0) inv_norm[i] = 1.f / nrm;
}

__global__ void __launch_bounds__(ROW_THREADS) gather_normalize_bwd_kernel(
    const float* __restrict__ table, const int64_t* __restrict__ id2row, const int64_t* __restrict__ ids,
    int64_t ids_stride, int64_t count, const float* __restrict__ grad, int64_t grad_stride,
    float* __restrict__ rows_out, int64_t* __restrict__ rows_id) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (i >= count) return;
  const int64_t row = resolve_row(id2row, ids, ids_stride, i);
  float4 y;
  const float nrm = normalize_row(table, row, lane, y);
  const float4 g = *reinterpret_cast<const float4*>(grad + i * grad_stride + lane * 4);
  *reinterpret_cast<float4*>(rows_out + i * D + lane * 4) = normalize_bwd(g, y, nrm);
  if (lane == 0) rows_id[i] = row;
}

__global__ void __launch_bounds__(ROW_THREADS) broadcast_rows_kernel(const float* __restrict__ src,
                                                                    const int64_t* __restrict__ src_rows, int num_rows,
                                                                    float* __restrict__ out, int64_t out_stride,
                                                                    int64_t count) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (w >= count * num_rows) return;
  const int64_t b = w / num_rows;
  const int j = (int)(w % num_rows);
  const float4 v = *reinterpret_cast<const float4*>(src + src_rows[j] * D + lane * 4);
  *reinterpret_cast<float4*>(out + b * out_stride + (int64_t)j * D + lane * 4) = v;
}

// ---- max readout ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) max_readout_fwd_kernel(const float* __restrict__ z, int64_t B, int n,
                                                                     float* __restrict__ q,
                                                                     int64_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* base = z + b * n * (int64_t)D + lane * 4;
  float4 best = *reinterpret_cast<const float4*>(base);
  int ax = 0, ay = 0, az = 0, aw = 0;
  for (int i = 1; i < n; ++i) {  // strict '>' keeps the smallest node index among equal maxima
    const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)i * D);
    if (v.x > best.x) { best.x = v.x; ax = i; }
    if (v.y > best.y) { best.y = v.y; ay = i; }
    if (v.z > best.z) { best.z = v.z; az = i; }
    if (v.w > best.w) { best.w = v.w; aw = i; }
  }
  *reinterpret_cast<float4*>(q + b * D + lane * 4) = best;
  if (argmax != nullptr) {
    int64_t* a = argmax + b * D + lane * 4;
    a[0] = b * n + ax; a[1] = b * n + ay; a[2] = b * n + az; a[3] = b * n + aw;
  }
}

__global__ void __launch_bounds__(ROW_THREADS) max_readout_bwd_kernel(const float* __restrict__ dq,
                                                                     const int64_t* __restrict__ argmax, int64_t B,
                                                                     int n, float* __restrict__ g) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float4 v = *reinterpret_cast<const float4*>(dq + b * D + lane * 4);
  const int64_t* a = argmax + b * D + lane * 4;
  const int64_t a0 = a[0] - b * n, a1 = a[1] - b * n, a2 = a[2] - b * n, a3 = a[3] - b * n;
  for (int i = 0; i < n; ++i) {
    const float4 o = make_float4(a0 == i ? v.x : 0.f, a1 == i ? v.y : 0.f, a2 == i ? v.z : 0.f, a3 == i ? v.w : 0.f);
    *reinterpret_cast<float4*>(g + (b * n + i) * (int64_t)D + lane * 4) = o;
  }
}

__global__ void __launch_bounds__(ROW_THREADS) cosine_margin_fwd_kernel(
    const float* __restrict__ q, int64_t B, const float* __restrict__ table, const int64_t* __restrict__ id2row,
    const int64_t* __restrict__ ids_pos, const int64_t* __restrict__ ids_neg, float margin,
    float* __restrict__ score_pos, float* __restrict__ score_neg, float* __restrict__ hinge) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float4 qv = *reinterpret_cast<const float4*>(q + b * D + lane * 4);
  float4 yp, yn;
  normalize_row(table, resolve_row(id2row, ids_pos, 1, b), lane, yp);
  normalize_row(table, resolve_row(id2row, ids_neg, 1, b), lane, yn);
  const float sp = cosine(qv, yp).score, sn = cosine(qv, yn).score;
  if (lane == 0) {
    score_pos[b] = sp;
    score_neg[b] = sn;
    hinge[b] = fmaxf(margin - (sp - sn), 0.f);
  }
}

// deterministic mean of hinge[B]: fixed strided partials per thread + fixed tree
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
  __shared__ float s[1024];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += v[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0] / (float)n;
}

__global__ void __launch_bounds__(ROW_THREADS) cosine_margin_bwd_kernel(
    const float* __restrict__ q, int64_t B, const float* __restrict__ table, const int64_t* __restrict__ id2row,
    const int64_t* __restrict__ ids_pos, const int64_t* __restrict__ ids_neg, float margin,
    const float* __restrict__ grad_loss, float* __restrict__ dq, float* __restrict__ rows_out,
    int64_t* __restrict__ rows_id) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float4 qv = *reinterpret_cast<const float4*>(q + b * D + lane * 4);
  const int64_t rp = resolve_row(id2row, ids_pos, 1, b), rn = resolve_row(id2row, ids_neg, 1, b);
  float4 yp, yn;
  const float np_ = normalize_row(table, rp, lane, yp);
  const float nn_ = normalize_row(table, rn, lane, yn);
  const Cos cp = cosine(qv, yp), cn = cosine(qv, yn);
  // clamp(min=0) passes the gradient where the argument is >= 0 (torch clamp backward)
  const float active = (margin - (cp.score - cn.score)) >= 0.f ? 1.f : 0.f;
  const float g = active * grad_loss[0] / (float)B;
  float4 dqp, dyp, dqn, dyn;
  cosine_bwd(qv, yp, cp, -g, dqp, dyp);
  cosine_bwd(qv, yn, cn, g, dqn, dyn);
  *reinterpret_cast<float4*>(dq + b * D + lane * 4) =
      make_float4(dqp.x + dqn.x, dqp.y + dqn.y, dqp.z + dqn.z, dqp.w + dqn.w);
  *reinterpret_cast<float4*>(rows_out + b * D + lane * 4) = normalize_bwd(dyp, yp, np_);
  *reinterpret_cast<float4*>(rows_out + (B + b) * D + lane * 4) = normalize_bwd(dyn, yn, nn_);
  if (lane == 0) {
    rows_id[b] = rp;
    rows_id[B + b] = rn;
  }
}

// owner of candidate i: largest b with offsets[b] <= i
__device__ __forceinline__ int64_t find_owner(const int64_t* offsets, int64_t B, int64_t i) {
  int64_t lo = 0, hi = B;  // invariant: offsets[lo] <= i < offsets[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(ROW_THREADS) cosine_scores_kernel(
    const float* __restrict__ q, int64_t B, const int64_t* __restrict__ offsets, const float* __restrict__ table,
    const int64_t* __restrict__ id2row, const int64_t* __restrict__ ids, int64_t count, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (i >= count) return;
  const int64_t b = offsets != nullptr ? find_owner(offsets, B, i) : i;
  const float4 qv = *reinterpret_cast<const float4*>(q + b * D + lane * 4);
  float4 y;
  normalize_row(table, resolve_row(id2row, ids, 1, i), lane, y);
  const float s = cosine(qv, y).score;
  if (lane == 0) scores[i] = s;
}

__global__ void __launch_bounds__(ROW_THREADS) cosine_scores_bwd_kernel(
    const float* __restrict__ q, int64_t B, const int64_t* __restrict__ offsets, const float* __restrict__ table,
    const int64_t* __restrict__ id2row, const int64_t* __restrict__ ids, const float* __restrict__ grad_scores,
    float* __restrict__ dq, int accumulate, float* __restrict__ rows_out, int64_t* __restrict__ rows_id) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int64_t i0 = offsets != nullptr ? offsets[b] : b;
  const int64_t i1 = offsets != nullptr ? offsets[b + 1] : b + 1;
  const float4 qv = *reinterpret_cast<const float4*>(q + b * D + lane * 4);
  float4 acc = accumulate ? *reinterpret_cast<const float4*>(dq + b * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = i0; i < i1; ++i) {  // ascending candidate order: bit-reproducible
    const int64_t row = resolve_row(id2row, ids, 1, i);
    float4 y, dqi, dyi;
    const float nrm = normalize_row(table, row, lane, y);
    const Cos c = cosine(qv, y);
    cosine_bwd(qv, y, c, grad_scores[i], dqi, dyi);
    acc.x += dqi.x; acc.y += dqi.y; acc.z += dqi.z; acc.w += dqi.w;
    *reinterpret_cast<float4*>(rows_out + i * D + lane * 4) = normalize_bwd(dyi, y, nrm);
    if (lane == 0) rows_id[i] = row;
  }
  *reinterpret_cast<float4*>(dq + b * D + lane * 4) = acc;
}

__global__ void __launch_bounds__(ROW_THREADS) rank_counts_ragged_kernel(const float* __restrict__ pos,
                                                                        const float* __restrict__ neg,
                                                                        const int64_t* __restrict__ offsets, int64_t B,
                                                                        int64_t* __restrict__ left,
                                                                        int64_t* __restrict__ right) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float p = pos[b];
  int lt = 0, le = 0;
  for (int64_t i = offsets[b] + lane; i < offsets[b + 1]; i += 32) {
    const float v = neg[i];
    lt += v < p;
    le += v <= p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
    le += __shfl_xor_sync(0xffffffffu, le, o);
  }
  if (lane == 0) {
    left[b] = lt;
    right[b] = le;
  }
}

// AUC pair counts: counts[0] += #(neg < pos), counts[1] += #(neg == pos) over ALL (positive, negative) pairs after
// nan_to_num (utils.py:34-36 -> sklearn roc_auc_score = the Mann-Whitney statistic (lt + eq/2) / (P*N)).  A CTA owns 256
// positives (one per thread) and a slice of the negatives staged through shared memory; integer atomics only.
__device__ __forceinline__ float nan_to_num(float v) {
  if (v != v) return 0.f;
  if (v > 3.402823466e38f) return 3.402823466e38f;
  if (v < -3.402823466e38f) return -3.402823466e38f;
  return v;
}
constexpr int AUC_NEG_TILE = 2048;
__global__ void __launch_bounds__(256) auc_counts_kernel(const float* __restrict__ pos, int64_t P,
                                                         const float* __restrict__ neg, int64_t N,
                                                         unsigned long long* __restrict__ counts) {
  __shared__ float tile[AUC_NEG_TILE];
  __shared__ unsigned long long red[2][8];
  const int64_t pi = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const float p = pi < P ? nan_to_num(pos[pi]) : 0.f;
  const int64_t n0 = (int64_t)blockIdx.y * AUC_NEG_TILE;
  const int cnt = (int)(N - n0 < AUC_NEG_TILE ? N - n0 : AUC_NEG_TILE);
  for (int i = threadIdx.x; i < cnt; i += 256) tile[i] = nan_to_num(neg[n0 + i]);
  __syncthreads();
  unsigned lt = 0, eq = 0;
  if (pi < P) {
#pragma unroll 8
    for (int i = 0; i < cnt; ++i) {
      const float v = tile[i];
      lt += v < p;
      eq += v == p;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
    eq += __shfl_xor_sync(0xffffffffu, eq, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = lt;
    red[1][warp] = eq;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s) atomicAdd(counts + threadIdx.x, s);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

}  // namespace
}  // namespace mpqe

using namespace mpqe;

extern "C" int mpqe_gather_normalize_fwd(const float* table, int64_t table_rows, const int64_t* id2row,
                                         const int64_t* ids, int64_t ids_stride, int64_t count, float* out,
                                         int64_t out_stride, float* inv_norm, void* stream) {
  MPQE_CHECK_ARG(table && ids && out && count >= 0 && ids_stride >= 1 && out_stride >= D && out_stride % 4 == 0,
                 "mpqe_gather_normalize_fwd: bad argument");
  if (count == 0) return 0;
  gather_normalize_fwd_kernel<<<row_blocks(count), ROW_THREADS, 0, (cudaStream_t)stream>>>(
      table, table_rows, id2row, ids, ids_stride, count, out, out_stride, inv_norm);
  MPQE_CHECK_LAUNCH("gather_normalize_fwd_kernel");
  return 0;
}

extern "C" int mpqe_gather_normalize_bwd(const float* table, const int64_t* id2row, const int64_t* ids,
                                         int64_t ids_stride, int64_t count, const float* grad, int64_t grad_stride,
                                         float* rows_out, int64_t* rows_id, void* stream) {
  MPQE_CHECK_ARG(table && ids && grad && rows_out && rows_id && count >= 0 && ids_stride >= 1 && grad_stride >= D,
                 "mpqe_gather_normalize_bwd: bad argument");
  if (count == 0) return 0;
  gather_normalize_bwd_kernel<<<row_blocks(count), ROW_THREADS, 0, (cudaStream_t)stream>>>(
      table, id2row, ids, ids_stride, count, grad, grad_stride, rows_out, rows_id);
  MPQE_CHECK_LAUNCH("gather_normalize_bwd_kernel");
  return 0;
}

extern "C" int mpqe_broadcast_rows(const float* src, const int64_t* src_rows, int32_t num_rows, float* out,
                                   int64_t out_stride, int64_t count, void* stream) {
  MPQE_CHECK_ARG(src && src_rows && out && num_rows >= 0 && count >= 0, "mpqe_broadcast_rows: bad argument");
  if (count == 0 || num_rows == 0) return 0;
  broadcast_rows_kernel<<<row_blocks(count * num_rows), ROW_THREADS, 0, (cudaStream_t)stream>>>(
      src, src_rows, num_rows, out, out_stride, count);
  MPQE_CHECK_LAUNCH("broadcast_rows_kernel");
  return 0;
}

extern "C" int mpqe_max_readout_fwd(const float* z, int64_t B, int32_t n, float* q, int64_t* argmax, void* stream) {
  MPQE_CHECK_ARG(z && q && B >= 1 && n >= 1, "mpqe_max_readout_fwd: bad argument");
  max_readout_fwd_kernel<<<row_blocks(B), ROW_THREADS, 0, (cudaStream_t)stream>>>(z, B, n, q, argmax);
  MPQE_CHECK_LAUNCH("max_readout_fwd_kernel");
  return 0;
}

extern "C" int mpqe_max_readout_bwd(const float* dq, const int64_t* argmax, int64_t B, int32_t n, float* g,
                                    void* stream) {
  MPQE_CHECK_ARG(dq && argmax && g && B >= 1 && n >= 1, "mpqe_max_readout_bwd: bad argument");
  max_readout_bwd_kernel<<<row_blocks(B), ROW_THREADS, 0, (cudaStream_t)stream>>>(dq, argmax, B, n, g);
  MPQE_CHECK_LAUNCH("max_readout_bwd_kernel");
  return 0;
}

extern "C" size_t mpqe_margin_loss_workspace_bytes(int64_t B) { return (size_t)B * sizeof(float); }

extern "C" int mpqe_cosine_margin_fwd(const float* q, int64_t B, const float* table, const int64_t* id2row,
                                      const int64_t* ids_pos, const int64_t* ids_neg, float margin, float* score_pos,
                                      float* score_neg, float* loss, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  MPQE_CHECK_ARG(q && table && ids_pos && ids_neg && score_pos && score_neg && loss && B >= 1,
                 "mpqe_cosine_margin_fwd: bad argument");
  MPQE_CHECK_ARG(workspace && workspace_bytes >= (size_t)B * sizeof(float), "mpqe_cosine_margin_fwd: workspace too small");
  cosine_margin_fwd_kernel<<<row_blocks(B), ROW_THREADS, 0, (cudaStream_t)stream>>>(
      q, B, table, id2row, ids_pos, ids_neg, margin, score_pos, score_neg, (float*)workspace);
  MPQE_CHECK_LAUNCH("cosine_margin_fwd_kernel");
  mean_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const float*)workspace, B, loss);
  MPQE_CHECK_LAUNCH("mean_kernel");
  return 0;
}

extern "C" int mpqe_cosine_margin_bwd(const float* q, int64_t B, const float* table, const int64_t* id2row,
                                      const int64_t* ids_pos, const int64_t* ids_neg, float margin,
                                      const float* grad_loss, float* dq, float* rows_out, int64_t* rows_id,
                                      void* stream) {
  MPQE_CHECK_ARG(q && table && ids_pos && ids_neg && grad_loss && dq && rows_out && rows_id && B >= 1,
                 "mpqe_cosine_margin_bwd: bad argument");
  cosine_margin_bwd_kernel<<<row_blocks(B), ROW_THREADS, 0, (cudaStream_t)stream>>>(
      q, B, table, id2row, ids_pos, ids_neg, margin, grad_loss, dq, rows_out, rows_id);
  MPQE_CHECK_LAUNCH("cosine_margin_bwd_kernel");
  return 0;
}

extern "C" int mpqe_cosine_scores(const float* q, int64_t B, const int64_t* offsets, const float* table,
                                  const int64_t* id2row, const int64_t* ids, int64_t count, float* scores,
                                  void* stream) {
  MPQE_CHECK_ARG(q && table && ids && scores && B >= 1 && count >= 0, "mpqe_cosine_scores: bad argument");
  MPQE_CHECK_ARG(offsets != nullptr || count == B, "mpqe_cosine_scores: count must equal B without offsets");
  if (count == 0) return 0;
  cosine_scores_kernel<<<row_blocks(count), ROW_THREADS, 0, (cudaStream_t)stream>>>(q, B, offsets, table, id2row, ids,
                                                                                 count, scores);
  MPQE_CHECK_LAUNCH("cosine_scores_kernel");
  return 0;
}

extern "C" int mpqe_cosine_scores_bwd(const float* q, int64_t B, const int64_t* offsets, const float* table,
                                      const int64_t* id2row, const int64_t* ids, int64_t count,
                                      const float* grad_scores, float* dq, int32_t accumulate, float* rows_out,
                                      int64_t* rows_id, void* stream) {
  MPQE_CHECK_ARG(q && table && ids && grad_scores && dq && rows_out && rows_id && B >= 1 && count >= 0,
                 "mpqe_cosine_scores_bwd: bad argument");
  MPQE_CHECK_ARG(offsets != nullptr || count == B, "mpqe_cosine_scores_bwd: count must equal B without offsets");
  cosine_scores_bwd_kernel<<<row_blocks(B), ROW_THREADS, 0, (cudaStream_t)stream>>>(
      q, B, offsets, table, id2row, ids, grad_scores, dq, accumulate, rows_out, rows_id);
  MPQE_CHECK_LAUNCH("cosine_scores_bwd_kernel");
  return 0;
}

extern "C" int mpqe_auc_counts(const float* pos, int64_t num_pos, const float* neg, int64_t num_neg,
                               unsigned long long* counts, void* stream) {
  MPQE_CHECK_ARG(pos && neg && counts && num_pos >= 1 && num_neg >= 1 && num_neg < (1ll << 40),
                 "mpqe_auc_counts: bad argument");
  const int64_t by = (num_neg + AUC_NEG_TILE - 1) / AUC_NEG_TILE;
  MPQE_CHECK_ARG(by <= 65535, "mpqe_auc_counts: more than %lld negatives", (long long)65535 * AUC_NEG_TILE);
  auc_counts_kernel<<<dim3((unsigned)((num_pos + 255) / 256), (unsigned)by), 256, 0, (cudaStream_t)stream>>>(
      pos, num_pos, neg, num_neg, counts);
  MPQE_CHECK_LAUNCH("auc_counts_kernel");
  return 0;
}

extern "C" int mpqe_rank_counts_ragged(const float* pos, const float* neg, const int64_t* offsets, int64_t B,
                                       int64_t* left, int64_t* right, void* stream) {
  MPQE_CHECK_ARG(pos && offsets && left && right && B >= 1, "mpqe_rank_counts_ragged: bad argument");
  rank_counts_ragged_kernel<<<row_blocks(B), ROW_THREADS, 0, (cudaStream_t)stream>>>(pos, neg, offsets, B, left, right);
  MPQE_CHECK_LAUNCH("rank_counts_ragged_kernel");
  return 0;
}

extern "C" int mpqe_adam_dense(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                               float lr, float beta1, float beta2, float eps, int32_t step, void* stream) {
  MPQE_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && numel >= 0 && step >= 1, "mpqe_adam_dense: bad argument");
  if (numel == 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, numel,
                                                                               lr, beta1, beta2, eps, bc1, bc2_sqrt);
  MPQE_CHECK_LAUNCH("adam_kernel");
  return 0;
}

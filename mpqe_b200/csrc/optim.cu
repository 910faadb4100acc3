// Callers of the hot path that live on the device in the fused training step:
//   * the L2 term of margin_loss over the readout-MLP parameters (/root/reference/mpqe/model.py:487-492),
//   * torch.optim.Adam (train.py:86-88, default hyper-parameters) over the dense parameters in one launch, and
//   * the same Adam over the entity tables driven by the ROW-SPARSE gradients of the step.  The reference's dense Adam
//     keeps moving a row through its momentum on the steps that do not touch it; those zero-gradient steps are
//     applied lazily -- right before the next step that reads the row (`catch-up`) -- so that every value the
//     forward pass ever reads, and the table after a final flush, is what dense Adam would have produced, while a
//     step only touches the rows of its batch (HBM traffic ~ touched rows, not the 190 MB table x 3).
// Also: device-side negative sampling (model.py:470-476), a counter-based draw from each query's stored negatives.
#include "common.cuh"

namespace mpqe {
namespace {

struct L2Launch {
  int n;
  int num_losses;
  float weight_decay, grad_scale;
  float* losses;
  float* norms;
  mpqe_l2_item_t it[MPQE_MAX_L2_ITEMS];
};

// One CTA walks the items in order: ||p|| by a fixed-order block reduction (bit-reproducible), then
// grad += grad_scale * weight_decay * p / ||p||; finally every loss gets weight_decay * sum_i ||p_i||.
__global__ void __launch_bounds__(1024) l2_reg_kernel(const __grid_constant__ L2Launch L) {
  __shared__ float warp_part[32];
  __shared__ float s_norm;
  float total = 0.f;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = 0; i < L.n; ++i) {
    const float* p = L.it[i].param;
    const int64_t n = L.it[i].numel;
    float acc = 0.f;
    for (int64_t k = tid; k < n; k += 1024) acc += p[k] * p[k];
    acc = warp_sum(acc);
    if (lane == 0) warp_part[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      float v = warp_part[lane];
      v = warp_sum(v);
      if (lane == 0) s_norm = sqrtf(v);
    }
    __syncthreads();
    const float nrm = s_norm;
    total += nrm;
    if (L.it[i].grad != nullptr && nrm > 0.f) {
      const float c = L.grad_scale * L.weight_decay / nrm;
      float* g = L.it[i].grad;
      for (int64_t k = tid; k < n; k += 1024) g[k] += c * p[k];
    }
    if (tid == 0 && L.norms != nullptr) L.norms[i] = nrm;
    __syncthreads();
  }
  if (L.losses != nullptr)
    for (int j = tid; j < L.num_losses; j += 1024) L.losses[j] += L.weight_decay * total;
}

// ---- Adam ---------------------------------------------------------------------------------------------------------
struct AdamHyper {
  float lr, b1, b2, eps;
};

// one torch.optim.Adam step on one element (single-tensor formula of torch/optim/adam.py)
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamHyper& H, float step_size,
                                            float bc2_sqrt) {
  m = m + (g - m) * (1.f - H.b1);                    // exp_avg.lerp_(grad, 1 - beta1)
  v = v * H.b2 + (1.f - H.b2) * g * g;               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + H.eps;
  p = p - step_size * (m / denom);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
}

struct AdamLaunch {
  int n;
  AdamHyper H;
  float step_size, bc2_sqrt;
  const mpqe_adam_state_t* state;   // device-resident step / bias corrections (CUDA-graph friendly), or nullptr
  mpqe_adam_item_t it[MPQE_MAX_ADAM_ITEMS];
};

__global__ void adam_tick_kernel(mpqe_adam_state_t* st, float lr, float b1, float b2) {
  const int step = st->step + 1;
  st->step = step;
  st->step_size = (float)((double)lr / (1.0 - pow((double)b1, (double)step)));
  st->bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)step));
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamLaunch L) {
  // blockIdx.y = item; grid-stride over its elements
  const mpqe_adam_item_t& T = L.it[blockIdx.y];
  const float step_size = L.state != nullptr ? L.state->step_size : L.step_size;
  const float bc2_sqrt = L.state != nullptr ? L.state->bc2_sqrt : L.bc2_sqrt;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < T.numel; i += (int64_t)gridDim.x * 256) {
    float p = T.param[i], m = T.exp_avg[i], v = T.exp_avg_sq[i];
    adam_update(p, m, v, T.grad[i], L.H, step_size, bc2_sqrt);
    T.param[i] = p;
    T.exp_avg[i] = m;
    T.exp_avg_sq[i] = v;
  }
}

struct TablesLaunch {
  int n;
  AdamHyper H;
  mpqe_adam_table_t t[MPQE_MAX_TABLES];
};

__device__ __forceinline__ bool find_table(const TablesLaunch& L, int64_t gid, int& ti, int64_t& row) {
  for (int i = 0; i < L.n; ++i)
    if (gid >= L.t[i].row_begin && gid < L.t[i].row_begin + L.t[i].rows) {
      ti = i;
      row = gid - L.t[i].row_begin;
      return true;
    }
  return false;
}

__device__ __forceinline__ void bias_corrections(const AdamHyper& H, int step, float& step_size, float& bc2_sqrt) {
  // as torch: python doubles, rounded to float where they meet the tensors
  const double bc1 = 1.0 - pow((double)H.b1, (double)step);
  const double bc2 = 1.0 - pow((double)H.b2, (double)step);
  step_size = (float)((double)H.lr / bc1);
  bc2_sqrt = (float)sqrt(bc2);
}

// Zero-gradient Adam steps last[row]+1 .. upto for every listed row (warp per id; duplicates are claimed once through
// atomicMax, so the result does not depend on which duplicate wins).  ids == nullptr: rows [0, count) of the id space.
__global__ void __launch_bounds__(256) adam_rows_catchup_kernel(const __grid_constant__ TablesLaunch L,
                                                                const int64_t* __restrict__ ids, int64_t count, int upto,
                                                                const mpqe_adam_state_t* __restrict__ state,
                                                                int* __restrict__ last) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= count) return;
  if (state != nullptr) upto = state->step;      // steps completed so far
  if (upto <= 0) return;
  const int64_t gid = ids != nullptr ? ids[w] : w;
  int ti;
  int64_t row;
  if (!find_table(L, gid, ti, row)) return;
  int old = 0;
  if (lane == 0) old = atomicMax(last + gid, upto);
  old = __shfl_sync(0xffffffffu, old, 0);
  if (old >= upto) return;
  const mpqe_adam_table_t& T = L.t[ti];
  float4 p = *reinterpret_cast<float4*>(T.table + row * D + lane * 4);
  float4 m = *reinterpret_cast<float4*>(T.exp_avg + row * D + lane * 4);
  float4 v = *reinterpret_cast<float4*>(T.exp_avg_sq + row * D + lane * 4);
  if (old == 0) {   // never touched: m = v = 0, the zero-gradient steps leave the row where it is
    return;
  }
  for (int s = old + 1; s <= upto; ++s) {
    float step_size, bc2_sqrt;
    bias_corrections(L.H, s, step_size, bc2_sqrt);
    adam_update(p.x, m.x, v.x, 0.f, L.H, step_size, bc2_sqrt);
    adam_update(p.y, m.y, v.y, 0.f, L.H, step_size, bc2_sqrt);
    adam_update(p.z, m.z, v.z, 0.f, L.H, step_size, bc2_sqrt);
    adam_update(p.w, m.w, v.w, 0.f, L.H, step_size, bc2_sqrt);
  }
  *reinterpret_cast<float4*>(T.table + row * D + lane * 4) = p;
  *reinterpret_cast<float4*>(T.exp_avg + row * D + lane * 4) = m;
  *reinterpret_cast<float4*>(T.exp_avg_sq + row * D + lane * 4) = v;
}

// Adam step `step` with the combined gradient rows (unique ids): warp per row
__global__ void __launch_bounds__(256) adam_rows_apply_kernel(const __grid_constant__ TablesLaunch L,
                                                              const int64_t* __restrict__ ids,
                                                              const float* __restrict__ rows,
                                                              const int64_t* __restrict__ num, int64_t max_count, int step,
                                                              float step_size, float bc2_sqrt,
                                                              const mpqe_adam_state_t* __restrict__ state,
                                                              int* __restrict__ last) {
  if (state != nullptr) {
    step = state->step;
    step_size = state->step_size;
    bc2_sqrt = state->bc2_sqrt;
  }
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t n = num != nullptr ? *num : max_count;
  if (w >= n || w >= max_count) return;
  const int64_t gid = ids[w];
  int ti;
  int64_t row;
  if (!find_table(L, gid, ti, row)) return;
  const mpqe_adam_table_t& T = L.t[ti];
  float4 p = *reinterpret_cast<float4*>(T.table + row * D + lane * 4);
  float4 m = *reinterpret_cast<float4*>(T.exp_avg + row * D + lane * 4);
  float4 v = *reinterpret_cast<float4*>(T.exp_avg_sq + row * D + lane * 4);
  const float4 g = *reinterpret_cast<const float4*>(rows + w * D + lane * 4);
  adam_update(p.x, m.x, v.x, g.x, L.H, step_size, bc2_sqrt);
  adam_update(p.y, m.y, v.y, g.y, L.H, step_size, bc2_sqrt);
  adam_update(p.z, m.z, v.z, g.z, L.H, step_size, bc2_sqrt);
  adam_update(p.w, m.w, v.w, g.w, L.H, step_size, bc2_sqrt);
  *reinterpret_cast<float4*>(T.table + row * D + lane * 4) = p;
  *reinterpret_cast<float4*>(T.exp_avg + row * D + lane * 4) = m;
  *reinterpret_cast<float4*>(T.exp_avg_sq + row * D + lane * 4) = v;
  if (lane == 0) last[gid] = step;
}

// ---- negative sampling -------------------------------------------------------------------------------------------
// splitmix64 of (seed, draw index): a counter-based generator, so a draw depends only on (seed, step, query) and not on
// launch geometry
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) sample_negatives_kernel(const int64_t* __restrict__ cand,
                                                               const int64_t* __restrict__ offsets,
                                                               const int64_t* __restrict__ query_index, int64_t first,
                                                               int64_t num_queries_total, int64_t shared_count,
                                                               int64_t count, uint64_t seed,
                                                               uint64_t step, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  // query of the stored set behind batch position i: an explicit index, or the reference's contiguous slice with
  // wrap-around (data_utils.py:300-308)
  const int64_t q = query_index != nullptr ? query_index[i] : (first + i) % num_queries_total;
  const int64_t b = offsets != nullptr ? offsets[q] : 0;
  const int64_t len = offsets != nullptr ? offsets[q + 1] - b : shared_count;
  const uint64_t r = mix64(mix64(seed ^ (step * 0xd1342543de82ef95ull)) + (uint64_t)i);
  out[i] = len > 0 ? cand[b + (int64_t)(r % (uint64_t)len)] : -1;
}

}  // namespace
}  // namespace mpqe

using namespace mpqe;

extern "C" int mpqe_l2_reg_multi(const mpqe_l2_item_t* items_host, int32_t n, float weight_decay, float grad_scale,
                                 float* losses, int32_t num_losses, float* norms, void* stream) {
  MPQE_CHECK_ARG(items_host != nullptr && n >= 1 && n <= MPQE_MAX_L2_ITEMS, "mpqe_l2_reg_multi: 1..%d items",
                 MPQE_MAX_L2_ITEMS);
  static thread_local L2Launch L;
  L.n = n;
  L.num_losses = losses != nullptr ? num_losses : 0;
  L.weight_decay = weight_decay;
  L.grad_scale = grad_scale;
  L.losses = losses;
  L.norms = norms;
  for (int i = 0; i < n; ++i) {
    MPQE_CHECK_ARG(items_host[i].param != nullptr && items_host[i].numel >= 0, "mpqe_l2_reg_multi: bad item %d", i);
    L.it[i] = items_host[i];
  }
  l2_reg_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("l2_reg_kernel");
  return 0;
}

static void host_bias_corrections(float lr, float b1, float b2, int step, float& step_size, float& bc2_sqrt) {
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  step_size = (float)((double)lr / bc1);
  bc2_sqrt = (float)sqrt(bc2);
}

extern "C" int mpqe_adam_tick(mpqe_adam_state_t* state, float lr, float beta1, float beta2, void* stream) {
  MPQE_CHECK_ARG(state != nullptr, "mpqe_adam_tick: state is null");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, lr, beta1, beta2);
  MPQE_CHECK_LAUNCH("adam_tick_kernel");
  return 0;
}

extern "C" int mpqe_adam_multi(const mpqe_adam_item_t* items_host, int32_t n, float lr, float beta1, float beta2,
                               float eps, int32_t step, const mpqe_adam_state_t* state, void* stream) {
  MPQE_CHECK_ARG(items_host != nullptr && n >= 1 && n <= MPQE_MAX_ADAM_ITEMS && (step >= 1 || state != nullptr),
                 "mpqe_adam_multi: 1..%d items, step >= 1 or a device state", MPQE_MAX_ADAM_ITEMS);
  static thread_local AdamLaunch L;
  L.n = n;
  L.H = AdamHyper{lr, beta1, beta2, eps};
  L.state = state;
  host_bias_corrections(lr, beta1, beta2, step >= 1 ? step : 1, L.step_size, L.bc2_sqrt);
  int64_t largest = 0;
  for (int i = 0; i < n; ++i) {
    const mpqe_adam_item_t& t = items_host[i];
    MPQE_CHECK_ARG(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.numel >= 0, "mpqe_adam_multi: bad item %d", i);
    L.it[i] = t;
    if (t.numel > largest) largest = t.numel;
  }
  if (largest == 0) return 0;
  int64_t bx = (largest + 255) / 256;
  if (bx > 1184) bx = 1184;     // 8 x 148 CTAs per item, grid-stride beyond
  adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("adam_multi_kernel");
  return 0;
}

static int fill_tables(TablesLaunch& L, const mpqe_adam_table_t* tables_host, int32_t n, float lr, float b1, float b2,
                       float eps) {
  MPQE_CHECK_ARG(tables_host != nullptr && n >= 1 && n <= MPQE_MAX_TABLES, "adam rows: 1..%d tables", MPQE_MAX_TABLES);
  L.n = n;
  L.H = AdamHyper{lr, b1, b2, eps};
  for (int i = 0; i < n; ++i) {
    const mpqe_adam_table_t& t = tables_host[i];
    MPQE_CHECK_ARG(t.table && t.exp_avg && t.exp_avg_sq && t.rows >= 0 && t.row_begin >= 0, "adam rows: bad table %d", i);
    L.t[i] = t;
  }
  return 0;
}

extern "C" int mpqe_adam_rows_catchup(const mpqe_adam_table_t* tables_host, int32_t num_tables, const int64_t* ids,
                                      int64_t count, int32_t upto_step, float lr, float beta1, float beta2, float eps,
                                      const mpqe_adam_state_t* state, int32_t* last_step, void* stream) {
  static thread_local TablesLaunch L;
  if (int rc = fill_tables(L, tables_host, num_tables, lr, beta1, beta2, eps)) return rc;
  MPQE_CHECK_ARG(last_step != nullptr && count >= 0 && upto_step >= 0, "mpqe_adam_rows_catchup: bad argument");
  if (count == 0 || (state == nullptr && upto_step == 0)) return 0;
  adam_rows_catchup_kernel<<<(unsigned)((count + 7) / 8), 256, 0, (cudaStream_t)stream>>>(L, ids, count, upto_step,
                                                                                         state, last_step);
  MPQE_CHECK_LAUNCH("adam_rows_catchup_kernel");
  return 0;
}

extern "C" int mpqe_adam_rows_apply(const mpqe_adam_table_t* tables_host, int32_t num_tables, const int64_t* ids,
                                    const float* rows, const int64_t* num, int64_t max_count, int32_t step, float lr,
                                    float beta1, float beta2, float eps, const mpqe_adam_state_t* state,
                                    int32_t* last_step, void* stream) {
  static thread_local TablesLaunch L;
  if (int rc = fill_tables(L, tables_host, num_tables, lr, beta1, beta2, eps)) return rc;
  MPQE_CHECK_ARG(ids && rows && last_step && max_count >= 0 && (step >= 1 || state != nullptr),
                 "mpqe_adam_rows_apply: bad argument");
  if (max_count == 0) return 0;
  float step_size, bc2_sqrt;
  host_bias_corrections(lr, beta1, beta2, step >= 1 ? step : 1, step_size, bc2_sqrt);
  adam_rows_apply_kernel<<<(unsigned)((max_count + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      L, ids, rows, num, max_count, step, step_size, bc2_sqrt, state, last_step);
  MPQE_CHECK_LAUNCH("adam_rows_apply_kernel");
  return 0;
}

extern "C" int mpqe_sample_negatives(const int64_t* candidates, const int64_t* offsets, const int64_t* query_index,
                                     int64_t first_query, int64_t num_queries_total, int64_t shared_count,
                                     int64_t count, uint64_t seed, uint64_t step, int64_t* out, void* stream) {
  MPQE_CHECK_ARG(candidates && out && count >= 0 && num_queries_total >= 1 && first_query >= 0 &&
                     (offsets != nullptr || shared_count >= 1),
                 "mpqe_sample_negatives: bad argument");
  if (count == 0) return 0;
  sample_negatives_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      candidates, offsets, query_index, first_query, num_queries_total, shared_count, count, seed, step, out);
  MPQE_CHECK_LAUNCH("sample_negatives_kernel");
  return 0;
}

// Fused R-GCN layer as a term-list GEMM, fp32 FFMA version (strict-fp32 path; the tcgen05 version is layer_tc.cu).
//
// Replaces, per layer pass (reference /root/reference/mpqe/model.py):
//   :292-294  index_select(w, edge_type) + bmm      -> per-template-edge dense [rows,128]x[128,128] terms
//   PyG propagate gather + scatter_add (:277)        -> terms of one output slot accumulate in registers
//   :301-304  x @ root + bias                        -> one more term + epilogue
//   :437      relu                                   -> epilogue
// and, with transposed matrices / swapped slots, the input-gradient of the same (autograd of the above).
#include <algorithm>

#include "common.cuh"

namespace mpqe {

namespace {

constexpr int BM = 64;        // queries per CTA tile
constexpr int KC = 32;        // k-chunk
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int A_PITCH = KC + 4;                 // floats; keeps 16-B alignment, spreads banks
constexpr int A_STAGE = BM * A_PITCH;           // floats
constexpr int B_STAGE = KC * D;                 // floats
constexpr int STAGE_FLOATS = A_STAGE + B_STAGE;
constexpr size_t LAYER_SMEM = size_t(STAGES) * STAGE_FLOATS * sizeof(float);

__device__ __forceinline__ const float* term_row(const mpqe_term_t& t, int64_t q) {
  return t.a + (q * t.a_slots + t.a_slot) * (int64_t)D;
}

__global__ void __launch_bounds__(THREADS, 2) layer_simt_kernel(const __grid_constant__ LayerLaunch L) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_term[MPQE_MAX_TERMS];
  __shared__ int s_nterms;

  // ---- which (group, tile, out slot) is this CTA? ----
  int unit = blockIdx.x;
  int gi = 0;
  for (; gi < L.num_groups - 1; ++gi) {
    const int tiles = int((L.g[gi].num_queries + BM - 1) / BM);
    const int units = tiles * L.g[gi].num_out_slots;
    if (unit < units) break;
    unit -= units;
  }
  const mpqe_layer_group_t& G = L.g[gi];
  const int slot = unit % G.num_out_slots;
  const int64_t q0 = int64_t(unit / G.num_out_slots) * BM;
  const int64_t B = G.num_queries;

  const int tid = threadIdx.x;
  if (tid == 0) {
    int n = 0;
    for (int t = 0; t < G.num_terms; ++t)
      if (G.terms[t].out_slot == slot) s_term[n++] = t;
    s_nterms = n;
  }
  __syncthreads();
  const int nsteps = s_nterms * (D / KC);

  const int tx = tid & 15;   // column group
  const int ty = tid >> 4;   // row group: rows ty*4 .. ty*4+3

  auto load_stage = [&](int step) {
    float* As = smem + (step % STAGES) * STAGE_FLOATS;
    float* Bs = As + A_STAGE;
    const mpqe_term_t& T = G.terms[s_term[step / (D / KC)]];
    const int kc = (step % (D / KC)) * KC;
    {  // A: 64 rows x 32 floats; 8 threads cover one row chunk (128 B)
      const int f4 = tid & 7;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 3) + 32 * i;
        int64_t q = q0 + r;
        if (q >= B) q = B - 1;
        cp_async16(As + r * A_PITCH + f4 * 4, term_row(T, q) + kc + f4 * 4);
      }
    }
    {  // B: 32 k-rows x 128 floats
      const int f4 = tid & 31;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = (tid >> 5) + 8 * i;
        cp_async16(Bs + k * D + f4 * 4, T.m + (int64_t)(kc + k) * D + f4 * 4);
      }
    }
  };

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nsteps) load_stage(s);
    cp_async_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (step + STAGES - 1 < nsteps) load_stage(step + STAGES - 1);
    cp_async_commit();
    const float* As = smem + (step % STAGES) * STAGE_FLOATS;
    const float* Bs = As + A_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty * 4 + i) * A_PITCH + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * D + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * D + 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, b0.x, acc[i][0]);
          acc[i][1] = fmaf(av, b0.y, acc[i][1]);
          acc[i][2] = fmaf(av, b0.z, acc[i][2]);
          acc[i][3] = fmaf(av, b0.w, acc[i][3]);
          acc[i][4] = fmaf(av, b1.x, acc[i][4]);
          acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]);
          acc[i][7] = fmaf(av, b1.w, acc[i][7]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: bias, activation / mask, store ----
  const int c0 = tx * 4, c1 = 64 + tx * 4;
  float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f), bv1 = bv0;
  if (G.bias != nullptr) {
    const float s = G.bias_scale[slot];
    const float* bj = G.bias + (int64_t)slot * G.bias_slot_stride;
    const float4 t0 = *reinterpret_cast<const float4*>(bj + c0);
    const float4 t1 = *reinterpret_cast<const float4*>(bj + c1);
    bv0 = make_float4(s * t0.x, s * t0.y, s * t0.z, s * t0.w);
    bv1 = make_float4(s * t1.x, s * t1.y, s * t1.z, s * t1.w);
  }
  const int oslot = G.out_slot_map[slot];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t q = q0 + ty * 4 + i;
    if (q >= B) continue;
    float4 v0 = make_float4(acc[i][0] + bv0.x, acc[i][1] + bv0.y, acc[i][2] + bv0.z, acc[i][3] + bv0.w);
    float4 v1 = make_float4(acc[i][4] + bv1.x, acc[i][5] + bv1.y, acc[i][6] + bv1.z, acc[i][7] + bv1.w);
    if (G.epilogue == MPQE_EPI_RELU) {
      v0 = make_float4(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f), fmaxf(v0.z, 0.f), fmaxf(v0.w, 0.f));
      v1 = make_float4(fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f), fmaxf(v1.z, 0.f), fmaxf(v1.w, 0.f));
    } else if (G.epilogue == MPQE_EPI_MASK) {
      const float* mrow = G.mask + (q * G.mask_slots + oslot) * (int64_t)D;
      const float4 m0 = *reinterpret_cast<const float4*>(mrow + c0);
      const float4 m1 = *reinterpret_cast<const float4*>(mrow + c1);
      v0 = make_float4(m0.x > 0.f ? v0.x : 0.f, m0.y > 0.f ? v0.y : 0.f, m0.z > 0.f ? v0.z : 0.f,
                       m0.w > 0.f ? v0.w : 0.f);
      v1 = make_float4(m1.x > 0.f ? v1.x : 0.f, m1.y > 0.f ? v1.y : 0.f, m1.z > 0.f ? v1.z : 0.f,
                       m1.w > 0.f ? v1.w : 0.f);
    }
    float* orow = G.out + (q * G.out_slots + oslot) * (int64_t)D;
    *reinterpret_cast<float4*>(orow + c0) = v0;
    *reinterpret_cast<float4*>(orow + c1) = v1;
  }
}

// ------------------------------------------------------------------------------------------------------------
// One-row groups (the batch-constant terms of pass 0: the variable-type embedding row is the same for every query
// of a batch, reference model.py:421).  out[slot] = bias + sum_t a_t @ M_t is a [1,128]x[128,128] product per term:
// one CTA per (group, out slot), its 8 warps each take 16 k's of every term, partial sums combined in a fixed order.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layer_row_kernel(const __grid_constant__ LayerLaunch L) {
  __shared__ float4 part[8][32];
  int unit = blockIdx.x;
  int gi = 0;
  for (; gi < L.num_groups - 1; ++gi) {
    if (unit < L.g[gi].num_out_slots) break;
    unit -= L.g[gi].num_out_slots;
  }
  const mpqe_layer_group_t& G = L.g[gi];
  const int slot = unit;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < G.num_terms; ++t) {
    const mpqe_term_t& T = G.terms[t];
    if (T.out_slot != slot) continue;
    const float* a = T.a + (int64_t)T.a_slot * D + w * 16;
    const float* m = T.m + (int64_t)(w * 16) * D + lane * 4;
    float av[16];
    float4 bv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      av[k] = __ldg(a + k);
      bv[k] = __ldg(reinterpret_cast<const float4*>(m + (int64_t)k * D));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      acc.x = fmaf(av[k], bv[k].x, acc.x);
      acc.y = fmaf(av[k], bv[k].y, acc.y);
      acc.z = fmaf(av[k], bv[k].z, acc.z);
      acc.w = fmaf(av[k], bv[k].w, acc.w);
    }
  }
  part[w][lane] = acc;
  __syncthreads();
  if (w != 0) return;
  float4 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = part[i][lane];
  float4 v = make_float4(((p[0].x + p[1].x) + (p[2].x + p[3].x)) + ((p[4].x + p[5].x) + (p[6].x + p[7].x)),
                         ((p[0].y + p[1].y) + (p[2].y + p[3].y)) + ((p[4].y + p[5].y) + (p[6].y + p[7].y)),
                         ((p[0].z + p[1].z) + (p[2].z + p[3].z)) + ((p[4].z + p[5].z) + (p[6].z + p[7].z)),
                         ((p[0].w + p[1].w) + (p[2].w + p[3].w)) + ((p[4].w + p[5].w) + (p[6].w + p[7].w)));
  if (G.bias != nullptr) {
    const float s = G.bias_scale[slot];
    const float4 b = *reinterpret_cast<const float4*>(G.bias + (int64_t)slot * G.bias_slot_stride + lane * 4);
    v = make_float4(fmaf(s, b.x, v.x), fmaf(s, b.y, v.y), fmaf(s, b.z, v.z), fmaf(s, b.w, v.w));
  }
  const int oslot = G.out_slot_map[slot];
  if (G.epilogue == MPQE_EPI_RELU) {
    v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
  } else if (G.epilogue == MPQE_EPI_MASK) {
    const float4 m = *reinterpret_cast<const float4*>(G.mask + (int64_t)oslot * D + lane * 4);
    v = make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f, m.w > 0.f ? v.w : 0.f);
  }
  *reinterpret_cast<float4*>(G.out + (int64_t)oslot * D + lane * 4) = v;
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient: dM_j = sum over matching terms, over queries, of A[q]^T G[q]  (a [128,B]x[B,128] reduction).
// Each CTA owns (destination j, chunk c) and a fixed query range per group; partials are reduced in order.
// ------------------------------------------------------------------------------------------------------------
constexpr int KQ = 16;  // queries per smem tile
constexpr int W_STAGE = 2 * KQ * D;
constexpr int W_STAGES = 3;
constexpr size_t WGRAD_SMEM = size_t(W_STAGES) * W_STAGE * sizeof(float);

// Iterates the (group, term, query-tile) work of one (dest, chunk) in a fixed order.
struct WgradIter {
  int g, t;
  int64_t q, qe;
};

__device__ __forceinline__ void chunk_range(int64_t B, int chunks, int c, int64_t& qb, int64_t& qe) {
  int64_t per = (B + chunks - 1) / chunks;
  per = (per + KQ - 1) / KQ * KQ;
  qb = per * c;
  qe = qb + per;
  if (qb > B) qb = B;
  if (qe > B) qe = B;
}

__device__ __forceinline__ bool wgrad_seek(const WgradLaunch& L, const float* m_fwd, int chunks, int c, WgradIter& it) {
  // advance (g, t) to the next matching term with a non-empty range, starting AT (it.g, it.t)
  for (; it.g < L.num_groups; ++it.g, it.t = 0) {
    const mpqe_layer_group_t& G = L.g[it.g];
    for (; it.t < G.num_terms; ++it.t) {
      if (G.terms[it.t].m != m_fwd) continue;
      chunk_range(G.num_queries, chunks, c, it.q, it.qe);
      if (it.q < it.qe) return true;
    }
  }
  return false;
}

__global__ void __launch_bounds__(THREADS, 2) wgrad_simt_kernel(const __grid_constant__ WgradLaunch L) {
  extern __shared__ __align__(16) float smem[];
  int unit = blockIdx.x;
  int j = 0;
  for (; j < L.num_dests - 1; ++j) {
    if (unit < L.chunks[j]) break;
    unit -= L.chunks[j];
  }
  const int c = unit;
  const int chunks = L.chunks[j];
  const float* m_fwd = L.d[j].m_fwd;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;

  WgradIter ld{0, 0, 0, 0};
  bool ld_ok = wgrad_seek(L, m_fwd, chunks, c, ld);

  auto issue = [&](int stage) {
    // loads the tile the iterator points at and advances it
    float* As = smem + stage * W_STAGE;
    float* Gs = As + KQ * D;
    const mpqe_layer_group_t& G = L.g[ld.g];
    const mpqe_term_t& T = G.terms[ld.t];
    const mpqe_wgrad_operand_t& O = L.go[ld.g];
    const int f4 = tid & 31;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (tid >> 5) + 8 * i;
      const int64_t q = ld.q + r;
      const bool ok = q < ld.qe;
      const int64_t qq = ok ? q : ld.q;
      const float* asrc = T.a + (qq * T.a_slots + T.a_slot) * (int64_t)D + f4 * 4;
      const float* gsrc = O.g + (qq * O.g_slots + O.slot_map[T.out_slot]) * (int64_t)D + f4 * 4;
      const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(As + r * D + f4 * 4));
      const unsigned sg = static_cast<unsigned>(__cvta_generic_to_shared(Gs + r * D + f4 * 4));
      const int nbytes = ok ? 16 : 0;  // src-size 0 -> zero fill
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(asrc), "r"(nbytes));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sg), "l"(gsrc), "r"(nbytes));
    }
    ld.q += KQ;
    if (ld.q >= ld.qe) {
      ++ld.t;
      ld_ok = wgrad_seek(L, m_fwd, chunks, c, ld);
    }
  };

  int issued = 0, consumed = 0;
#pragma unroll
  for (int s = 0; s < W_STAGES - 1; ++s) {
    if (ld_ok) { issue(issued % W_STAGES); ++issued; }
    cp_async_commit();
  }
  while (consumed < issued) {
    cp_async_wait<W_STAGES - 2>();
    __syncthreads();
    if (ld_ok) { issue(issued % W_STAGES); ++issued; }
    cp_async_commit();
    const float* As = smem + (consumed % W_STAGES) * W_STAGE;
    const float* Gs = As + KQ * D;
#pragma unroll 4
    for (int r = 0; r < KQ; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + r * D + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(As + r * D + 64 + ty * 4);
      const float4 g0 = *reinterpret_cast<const float4*>(Gs + r * D + tx * 4);
      const float4 g1 = *reinterpret_cast<const float4*>(Gs + r * D + 64 + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(av[i], gv[k], acc[i][k]);
    }
    ++consumed;
  }
  cp_async_wait<0>();

  int pbase = 0;
  for (int jj = 0; jj < j; ++jj) pbase += L.chunks[jj];
  float* P = L.partials + (int64_t)(pbase + c) * D * D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    *reinterpret_cast<float4*>(P + row * D + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(P + row * D + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradLaunch L) {
  const int j = blockIdx.y;
  int pbase = 0;
  for (int jj = 0; jj < j; ++jj) pbase += L.chunks[jj];
  const int e4 = blockIdx.x * 256 + threadIdx.x;  // float4 index into the [D,D] matrix
  if (e4 >= D * D / 4) return;
  const int n = L.chunks[j];
  const float* base = L.partials + (int64_t)pbase * D * D + e4 * 4;
  // four independent accumulation chains (chunks c = k mod 4), combined in a fixed order: bit-reproducible, and four
  // loads in flight per thread instead of one dependent chain
  float4 s[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int c = 0;
  for (; c + 4 <= n; c += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)(c + k) * D * D);
      s[k].x += v.x; s[k].y += v.y; s[k].z += v.z; s[k].w += v.w;
    }
  }
  for (int k = 0; c < n; ++c, ++k) {
    const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)c * D * D);
    s[k].x += v.x; s[k].y += v.y; s[k].z += v.z; s[k].w += v.w;
  }
  float4 r = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                         (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
  float4* dst = reinterpret_cast<float4*>(L.d[j].dm) + e4;
  if (L.d[j].accumulate) {
    const float4 o = *dst;
    r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
  }
  *dst = r;
}

// One-row groups: dM_j (+)= sum over matching terms of a_t^T g[slot] -- rank-1 updates, summed in term order by the
// thread that owns the element; no partials, no second stage.
__global__ void __launch_bounds__(256) wgrad_row_kernel(const __grid_constant__ WgradLaunch L) {
  const int j = blockIdx.y;
  const int e4 = blockIdx.x * 256 + threadIdx.x;  // float4 index into the [D,D] matrix
  if (e4 >= D * D / 4) return;
  const int row = e4 / (D / 4), c4 = e4 % (D / 4);
  const float* m_fwd = L.d[j].m_fwd;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g = 0; g < L.num_groups; ++g) {
    const mpqe_layer_group_t& G = L.g[g];
    const mpqe_wgrad_operand_t& O = L.go[g];
    for (int t = 0; t < G.num_terms; ++t) {
      const mpqe_term_t& T = G.terms[t];
      if (T.m != m_fwd) continue;
      const float a = __ldg(T.a + (int64_t)T.a_slot * D + row);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(O.g + (int64_t)O.slot_map[T.out_slot] * D) + c4);
      r.x = fmaf(a, gv.x, r.x);
      r.y = fmaf(a, gv.y, r.y);
      r.z = fmaf(a, gv.z, r.z);
      r.w = fmaf(a, gv.w, r.w);
    }
  }
  float4* dst = reinterpret_cast<float4*>(L.d[j].dm) + e4;
  if (L.d[j].accumulate) {
    const float4 o = *dst;
    r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
  }
  *dst = r;
}

// ------------------------------------------------------------------------------------------------------------
// Column sums (bias and mode-embedding gradients), two deterministic stages.
// ------------------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 256;  // rows per CTA

__global__ void __launch_bounds__(128) colsum_partial_kernel(const float* __restrict__ src, int64_t rows,
                                                             int64_t stride, float* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * CS_ROWS;
  int64_t r1 = r0 + CS_ROWS;
  if (r1 > rows) r1 = rows;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += src[r * stride + threadIdx.x];
  partial[(int64_t)blockIdx.x * D + threadIdx.x] = s;
}

__global__ void __launch_bounds__(128) colsum_final_kernel(const float* __restrict__ partial, int nblk, float scale,
                                                           float* __restrict__ out, int accumulate) {
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * D + threadIdx.x];
  s *= scale;
  out[threadIdx.x] = accumulate ? out[threadIdx.x] + s : s;
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const float* s = src + (int64_t)blockIdx.z * rows * cols;
  float* d = dst + (int64_t)blockIdx.z * rows * cols;
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    if (r < rows && c < cols) tile[i][threadIdx.x] = s[(int64_t)r * cols + c];
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;  // becomes the column of dst
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;          // becomes the row of dst
    if (r2 < rows && c2 < cols) d[(int64_t)c2 * rows + r2] = tile[threadIdx.x][i];
  }
}

int validate_groups(const mpqe_layer_group_t* g, int n, const char* who) {
  MPQE_CHECK_ARG(g != nullptr && n >= 1 && n <= MPQE_MAX_GROUPS, "%s: num_groups must be in [1,%d], got %d", who,
                 MPQE_MAX_GROUPS, n);
  for (int i = 0; i < n; ++i) {
    MPQE_CHECK_ARG(g[i].num_queries >= 1, "%s: group %d has no queries", who, i);
    MPQE_CHECK_ARG(g[i].num_terms >= 0 && g[i].num_terms <= MPQE_MAX_TERMS, "%s: group %d: bad num_terms %d", who, i,
                   g[i].num_terms);
    MPQE_CHECK_ARG(g[i].num_out_slots >= 1 && g[i].num_out_slots <= MPQE_MAX_SLOTS, "%s: group %d: bad num_out_slots",
                   who, i);
    for (int t = 0; t < g[i].num_terms; ++t) {
      const mpqe_term_t& T = g[i].terms[t];
      MPQE_CHECK_ARG(T.a != nullptr && T.m != nullptr, "%s: group %d term %d: null operand", who, i, t);
      MPQE_CHECK_ARG(T.out_slot >= 0 && T.out_slot < g[i].num_out_slots, "%s: group %d term %d: out_slot %d", who, i,
                     t, (int)T.out_slot);
      MPQE_CHECK_ARG(T.a_slots >= 0 && T.a_slot >= 0 && (T.a_slots == 0 || T.a_slot < T.a_slots),
                     "%s: group %d term %d: bad a_slot %d/%d", who, i, t, (int)T.a_slot, T.a_slots);
    }
  }
  return 0;
}

}  // namespace

int layer_forward_simt(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(layer_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAYER_SMEM));
    configured = true;
  }
  LayerLaunch L;
  memset(&L, 0, sizeof(L));
  L.num_groups = num_groups;
  int64_t units = 0;
  for (int i = 0; i < num_groups; ++i) {
    L.g[i] = groups[i];
    units += (groups[i].num_queries + BM - 1) / BM * groups[i].num_out_slots;
  }
  MPQE_CHECK_ARG(units < (1ll << 31), "mpqe_layer_forward: too many tiles");
  layer_simt_kernel<<<(unsigned)units, THREADS, LAYER_SMEM, stream>>>(L);
  MPQE_CHECK_LAUNCH("layer_simt_kernel");
  return 0;
}

// every group has one row: one CTA per (group, out slot)
static int layer_forward_rows(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream) {
  LayerLaunch L;
  memset(&L, 0, sizeof(L));
  L.num_groups = num_groups;
  int units = 0;
  for (int i = 0; i < num_groups; ++i) {
    L.g[i] = groups[i];
    units += groups[i].num_out_slots;
  }
  layer_row_kernel<<<units, 256, 0, stream>>>(L);
  MPQE_CHECK_LAUNCH("layer_row_kernel");
  return 0;
}

}  // namespace mpqe

using namespace mpqe;

extern "C" int mpqe_layer_forward(const mpqe_layer_group_t* groups_host, int32_t num_groups, int32_t use_tensor_cores,
                                  void* stream) {
  if (validate_groups(groups_host, num_groups, "mpqe_layer_forward")) return 1;
  for (int i = 0; i < num_groups; ++i) {
    MPQE_CHECK_ARG(groups_host[i].out != nullptr, "mpqe_layer_forward: group %d: null out", i);
    MPQE_CHECK_ARG(groups_host[i].epilogue != MPQE_EPI_MASK || groups_host[i].mask != nullptr,
                   "mpqe_layer_forward: group %d: EPI_MASK without mask", i);
  }
  bool rows_only = true;
  for (int i = 0; i < num_groups; ++i) rows_only = rows_only && groups_host[i].num_queries == 1;
  if (rows_only) return layer_forward_rows(groups_host, num_groups, (cudaStream_t)stream);
  if (use_tensor_cores)
    return tc_generation() == 2 ? layer_forward_tc2(groups_host, num_groups, (cudaStream_t)stream)
                                : layer_forward_tc(groups_host, num_groups, (cudaStream_t)stream);
  return layer_forward_simt(groups_host, num_groups, (cudaStream_t)stream);
}

extern "C" size_t mpqe_layer_wgrad_workspace_bytes(int32_t num_dests, int32_t num_ctas_hint) {
  if (num_ctas_hint < 1) num_ctas_hint = 296;
  return (size_t)(num_ctas_hint + num_dests) * D * D * sizeof(float);
}

extern "C" int mpqe_layer_wgrad(const mpqe_layer_group_t* groups_host, const mpqe_wgrad_operand_t* grads_host,
                                int32_t num_groups, const mpqe_wgrad_dest_t* dests_host, int32_t num_dests,
                                int32_t use_tensor_cores, void* workspace, size_t workspace_bytes, void* stream) {
  if (validate_groups(groups_host, num_groups, "mpqe_layer_wgrad")) return 1;
  MPQE_CHECK_ARG(num_dests >= 1 && num_dests <= MPQE_MAX_DESTS, "mpqe_layer_wgrad: num_dests must be in [1,%d]",
                 MPQE_MAX_DESTS);
  MPQE_CHECK_ARG(grads_host != nullptr && dests_host != nullptr && workspace != nullptr,
                 "mpqe_layer_wgrad: null argument");
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(wgrad_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGRAD_SMEM));
    configured = true;
  }
  static thread_local WgradLaunch L;
  memset(&L, 0, sizeof(L));
  L.num_groups = num_groups;
  L.num_dests = num_dests;
  L.partials = (float*)workspace;
  for (int i = 0; i < num_groups; ++i) {
    L.g[i] = groups_host[i];
    L.go[i] = grads_host[i];
    MPQE_CHECK_ARG(grads_host[i].g != nullptr, "mpqe_layer_wgrad: group %d: null gradient operand", i);
  }
  bool rows_only = true;
  for (int i = 0; i < num_groups; ++i) rows_only = rows_only && groups_host[i].num_queries == 1;
  if (rows_only) {   // rank-1 updates: no partials
    for (int j = 0; j < num_dests; ++j) {
      L.d[j] = dests_host[j];
      MPQE_CHECK_ARG(dests_host[j].dm != nullptr, "mpqe_layer_wgrad: dest %d: null dm", j);
    }
    wgrad_row_kernel<<<dim3(D * D / 4 / 256, num_dests), 256, 0, (cudaStream_t)stream>>>(L);
    MPQE_CHECK_LAUNCH("wgrad_row_kernel");
    return 0;
  }
  // chunks per destination proportional to its query-rows of work; total bounded by the workspace
  const int64_t max_parts = (int64_t)(workspace_bytes / (D * D * sizeof(float)));
  MPQE_CHECK_ARG(max_parts >= num_dests, "mpqe_layer_wgrad: workspace too small (%zu bytes)", workspace_bytes);
  double weight[MPQE_MAX_DESTS];
  double total = 0;
  for (int j = 0; j < num_dests; ++j) {
    L.d[j] = dests_host[j];
    MPQE_CHECK_ARG(dests_host[j].dm != nullptr, "mpqe_layer_wgrad: dest %d: null dm", j);
    weight[j] = 0;
    for (int i = 0; i < num_groups; ++i)
      for (int t = 0; t < groups_host[i].num_terms; ++t)
        if (groups_host[i].terms[t].m == dests_host[j].m_fwd) weight[j] += (double)groups_host[i].num_queries;
    total += weight[j];
  }
  // Apportion the max_parts (destination, chunk) units so that the LARGEST unit is as small as possible, counted in
  // the stages the kernels actually run: a chunk of destination j costs sum over groups of (matching terms) x (query
  // tiles of the group's chunk range, chunk_range() on the device), and chunk ranges are whole tiles -- 4096 queries in
  // 18 chunks are 16 chunks of 8 tiles and two empty ones.  Smallest feasible maximum by bisection; per destination the
  // fewest chunks that reach it.  (The first version gave each destination 1 + floor(share * budget) chunks: with 18
  // destinations the largest unit had 28 stages against a mean of 20.8 per CTA, and 8 of the 148 CTAs had none.)
  const int64_t gran = use_tensor_cores ? 32 : KQ;
  int terms_of[MPQE_MAX_DESTS][MPQE_MAX_GROUPS];
  int64_t cap[MPQE_MAX_DESTS];
  for (int j = 0; j < num_dests; ++j) {
    cap[j] = (int64_t)(weight[j] / (4 * KQ)) + 1;  // at least ~64 query rows per chunk
    if (cap[j] > 256) cap[j] = 256;
    for (int i = 0; i < num_groups; ++i) {
      terms_of[j][i] = 0;
      for (int t = 0; t < groups_host[i].num_terms; ++t)
        if (groups_host[i].terms[t].m == dests_host[j].m_fwd) ++terms_of[j][i];
    }
  }
  auto unit_cost = [&](int j, int64_t c) {      // stages of the largest chunk of destination j split c ways
    int64_t stages = 0;
    for (int i = 0; i < num_groups; ++i) {
      if (terms_of[j][i] == 0) continue;
      const int64_t B = groups_host[i].num_queries;
      int64_t per = (B + c - 1) / c;
      per = (per + gran - 1) / gran * gran;
      if (per > B) per = B;
      stages += terms_of[j][i] * ((per + gran - 1) / gran);
    }
    return stages;
  };
  auto chunks_for = [&](int j, int64_t limit) {  // fewest chunks with unit_cost <= limit (cap[j] if none reaches it)
    int64_t lo = 1, hi = cap[j];
    if (unit_cost(j, hi) > limit) return hi;
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (unit_cost(j, mid) <= limit) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  int64_t lo = 1, hi = 1;
  for (int j = 0; j < num_dests; ++j) hi = std::max(hi, unit_cost(j, 1));
  while (lo < hi) {
    const int64_t mid = (lo + hi) / 2;
    int64_t need = 0;
    for (int j = 0; j < num_dests; ++j) need += chunks_for(j, mid);
    if (need <= max_parts) hi = mid; else lo = mid + 1;
  }
  int total_chunks = 0;
  for (int j = 0; j < num_dests; ++j) {
    L.chunks[j] = (int)chunks_for(j, lo);
    total_chunks += L.chunks[j];
  }
  if (total_chunks > max_parts) {    // (caps made the bound unreachable: shrink the largest counts)
    for (int j = 0; j < num_dests && total_chunks > max_parts; ++j)
      while (L.chunks[j] > 1 && total_chunks > max_parts) --L.chunks[j], --total_chunks;
  }
  if (use_tensor_cores) {
    if (layer_wgrad_tc_launch(L, total_chunks, (cudaStream_t)stream)) return 2;
  } else {
    wgrad_simt_kernel<<<total_chunks, THREADS, WGRAD_SMEM, (cudaStream_t)stream>>>(L);
    MPQE_CHECK_LAUNCH("wgrad_simt_kernel");
  }
  wgrad_reduce_kernel<<<dim3(D * D / 4 / 256, num_dests), 256, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("wgrad_reduce_kernel");
  return 0;
}

extern "C" size_t mpqe_colsum_workspace_bytes(int64_t rows) {
  return (size_t)((rows + CS_ROWS - 1) / CS_ROWS + 1) * D * sizeof(float);
}

extern "C" int mpqe_colsum(const float* src, int64_t rows, int64_t stride, float scale, float* out,
                           int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  MPQE_CHECK_ARG(src != nullptr && out != nullptr && rows >= 0, "mpqe_colsum: bad argument");
  const int64_t nblk = (rows + CS_ROWS - 1) / CS_ROWS;
  MPQE_CHECK_ARG(workspace_bytes >= (size_t)nblk * D * sizeof(float) && (nblk == 0 || workspace != nullptr),
                 "mpqe_colsum: workspace too small");
  if (nblk > 0) {
    colsum_partial_kernel<<<(unsigned)nblk, 128, 0, (cudaStream_t)stream>>>(src, rows, stride, (float*)workspace);
    MPQE_CHECK_LAUNCH("colsum_partial_kernel");
  }
  colsum_final_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const float*)workspace, (int)nblk, scale, out, accumulate);
  MPQE_CHECK_LAUNCH("colsum_final_kernel");
  return 0;
}

namespace mpqe {
namespace {
struct MatsumLaunch {
  mpqe_matsum_item_t it[MPQE_MAX_MATSUM_ITEMS];
};
// one float4 of one destination per thread; summands in item order (fixed: bit-reproducible)
__global__ void __launch_bounds__(256) matrix_sum_multi_kernel(const __grid_constant__ MatsumLaunch L) {
  const mpqe_matsum_item_t& T = L.it[blockIdx.y];
  const int e4 = blockIdx.x * 256 + threadIdx.x;
  if (e4 >= D * D / 4) return;
  float4* dst = reinterpret_cast<float4*>(T.dst) + e4;
  float4 s = T.accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < T.num_src; ++k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(T.src[k]) + e4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *dst = s;
}

// ---- basis decomposition (RGCNConv num_bases > 0, model.py:281-284): W_r = sum_b att[r, b] * basis[b] --------------
// out[m, e] = sum_k a[m*a_row + k*a_col] * b[k, e]   (K small: <= 64 bases / relations per pass of the loop).
// One float4 of one output row per thread, k ascending: fixed order.  With (a_row, a_col) = (K, 1) this is
// W = att @ basis, with (1, M) and b = dW it is d basis = att^T @ dW.
__global__ void __launch_bounds__(256) small_k_matmul_kernel(const float* __restrict__ a, int64_t a_row, int64_t a_col,
                                                             const float* __restrict__ b, int M, int K, int64_t E4,
                                                             float* __restrict__ out) {
  const int64_t e4 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int m = blockIdx.y;
  if (e4 >= E4) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(a + m * a_row + k * a_col);
    const float4 v = __ldg(reinterpret_cast<const float4*>(b) + (int64_t)k * E4 + e4);
    s.x += w * v.x; s.y += w * v.y; s.z += w * v.z; s.w += w * v.w;
  }
  reinterpret_cast<float4*>(out)[(int64_t)m * E4 + e4] = s;
}

// out[m, k] = sum_e x[m, e] * y[k, e]   (d att = dW . basis): one CTA per (m, k), fixed strided partials + fixed tree
__global__ void __launch_bounds__(256) rows_dot_kernel(const float* __restrict__ x, const float* __restrict__ y, int K,
                                                       int64_t E4, float* __restrict__ out) {
  __shared__ float red[256];
  const int m = blockIdx.y, k = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x) + (int64_t)m * E4;
  const float4* yr = reinterpret_cast<const float4*>(y) + (int64_t)k * E4;
  float acc = 0.f;
  for (int64_t e = threadIdx.x; e < E4; e += 256) acc += dot4(__ldg(xr + e), __ldg(yr + e));
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(int64_t)m * K + k] = red[0];
}
}  // namespace
}  // namespace mpqe

extern "C" int mpqe_small_k_matmul(const float* a, int64_t a_row_stride, int64_t a_col_stride, const float* b,
                                   int32_t M, int32_t K, int64_t E, float* out, void* stream) {
  MPQE_CHECK_ARG(a != nullptr && b != nullptr && out != nullptr && M >= 1 && M < 65536 && K >= 1 && E >= 4 && E % 4 == 0,
                 "mpqe_small_k_matmul: bad argument");
  const int64_t E4 = E / 4;
  small_k_matmul_kernel<<<dim3((unsigned)((E4 + 255) / 256), (unsigned)M), 256, 0, (cudaStream_t)stream>>>(
      a, a_row_stride, a_col_stride, b, M, K, E4, out);
  MPQE_CHECK_LAUNCH("small_k_matmul_kernel");
  return 0;
}

extern "C" int mpqe_rows_dot(const float* x, const float* y, int32_t M, int32_t K, int64_t E, float* out, void* stream) {
  MPQE_CHECK_ARG(x != nullptr && y != nullptr && out != nullptr && M >= 1 && M < 65536 && K >= 1 && E >= 4 && E % 4 == 0,
                 "mpqe_rows_dot: bad argument");
  rows_dot_kernel<<<dim3((unsigned)K, (unsigned)M), 256, 0, (cudaStream_t)stream>>>(x, y, K, E / 4, out);
  MPQE_CHECK_LAUNCH("rows_dot_kernel");
  return 0;
}

extern "C" int mpqe_matrix_sum_multi(const mpqe_matsum_item_t* items_host, int32_t n, void* stream) {
  MPQE_CHECK_ARG(items_host != nullptr && n >= 1 && n <= MPQE_MAX_MATSUM_ITEMS,
                 "mpqe_matrix_sum_multi: n must be in [1,%d]", MPQE_MAX_MATSUM_ITEMS);
  static thread_local MatsumLaunch L;
  for (int i = 0; i < n; ++i) {
    const mpqe_matsum_item_t& T = items_host[i];
    MPQE_CHECK_ARG(T.dst != nullptr && T.num_src >= 1 && T.num_src <= MPQE_MAX_MATSUM_SRCS,
                   "mpqe_matrix_sum_multi: item %d: bad argument", i);
    for (int k = 0; k < T.num_src; ++k)
      MPQE_CHECK_ARG(T.src[k] != nullptr, "mpqe_matrix_sum_multi: item %d: null summand %d", i, k);
    for (int j = 0; j < i; ++j)
      MPQE_CHECK_ARG(items_host[j].dst != T.dst, "mpqe_matrix_sum_multi: items %d and %d share a destination", j, i);
    L.it[i] = T;
  }
  matrix_sum_multi_kernel<<<dim3(D * D / 4 / 256, n), 256, 0, (cudaStream_t)stream>>>(L);
  MPQE_CHECK_LAUNCH("matrix_sum_multi_kernel");
  return 0;
}

extern "C" int mpqe_transpose(const float* src, float* dst, int64_t count, int32_t rows, int32_t cols, void* stream) {
  MPQE_CHECK_ARG(src != nullptr && dst != nullptr && count >= 1 && rows >= 1 && cols >= 1 && count < 65536,
                 "mpqe_transpose: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, (unsigned)count);
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, rows, cols);
  MPQE_CHECK_LAUNCH("transpose_kernel");
  return 0;
}

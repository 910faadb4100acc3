// Fused R-GCN layer on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), fp32-accurate through
// a 3xTF32 split:  A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  with  x_hi = tf32(x), x_lo = x - x_hi.
//
// Same contract as layer_simt.cu (term-list GEMM / term-list outer product, see include/mpqe_b200.h); the reference
// lines replaced are /root/reference/mpqe/model.py:292-294 (index_select + bmm), :277 (gather + scatter_add),
// :301-304 (root, bias), :437 (relu) and their autograd.
//
// Data path (per CTA, 256 threads, 1 CTA / SM):
//   global fp32 --ld.global.v4--> registers --split hi/lo--> st.shared.v4 in the UMMA canonical NO-SWIZZLE layout
//   (8x16-byte core matrices, written 512 contiguous bytes per warp instruction: no bank conflicts)
//   --fence.proxy.async + barrier--> one thread issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8) x 3 products
//   --tcgen05.commit--> mbarrier frees the smem stage; after the last term: tcgen05.ld 32x32b -> epilogue -> global.
// The operands are staged by the threads rather than by TMA because every element has to pass through registers
// once anyway to be split into its hi/lo TF32 parts.
#include "common.cuh"

namespace mpqe {

namespace {

constexpr int THREADS = 256;
constexpr int BM = 128;              // rows (queries) per CTA = UMMA M
constexpr int KC = 32;               // k per pipeline stage = 4 UMMA k-steps of 8
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * KC * 4;          // 16 KB: one operand tile (hi or lo)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int TMEM_COLS = 128;
constexpr size_t TC_SMEM = size_t(STAGES) * STAGE_BYTES + 1024;  // + alignment slack

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded spin: a protocol bug must surface as a trapped launch, never as a hung GPU
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 26)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE ("interleave"), sm_100 version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor: c=f32, a=b=tf32 (both K-major: bits 15/16 clear), N=128, M=128
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
  uint32_t h0, h1, h2, h3;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(x.x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(x.y));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h2) : "f"(x.z));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h3) : "f"(x.w));
  hi = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// Every operand tile is K-major [128 rows][32 k]: element (row, k) at (row/8)*1024 + (k/4)*128 + (row%8)*16 + (k%4)*4
// bytes; descriptor for k-step j (k = 8j..8j+7): start + j*256, LBO = 128 (next 4 k), SBO = 1024 (next 8 rows).
// (Probed on B200 with tests/tc_probe.cu: K-major no-swizzle tf32 operands are exact, MN-major ones yield zeros.)
// Sources whose contiguous dimension is the tile's ROW dimension (weight matrices [k][n], and both operands of the
// weight gradient) are transposed on the way into shared memory with 4-byte stores whose component order is rotated
// per lane so that each warp instruction hits 32 distinct banks.

struct Frag {  // one pipeline stage worth of one operand, per thread
  float4 v[4];
};

// rows = queries (K-major A): idx = warp*4+i -> row group idx/2, k half idx%2
__device__ __forceinline__ void load_kmajor(Frag& f, const mpqe_term_t& T, int64_t q0, int64_t B, int kc, int warp,
                                            int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = warp * 4 + i;
    const int row = (idx >> 1) * 8 + (lane & 7);
    const int kq = (idx & 1) * 4 + (lane >> 3);
    int64_t q = q0 + row;
    if (q >= B) q = B - 1;
    f.v[i] = *reinterpret_cast<const float4*>(T.a + (q * T.a_slots + T.a_slot) * (int64_t)D + kc + kq * 4);
  }
}
__device__ __forceinline__ void store_kmajor(const Frag& f, uint8_t* hi_tile, uint8_t* lo_tile, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = warp * 4 + i;
    const int off = (idx >> 1) * 1024 + ((idx & 1) * 4 + (lane >> 3)) * 128 + (lane & 7) * 16;
    float4 hi, lo;
    split_tf32(f.v[i], hi, lo);
    *reinterpret_cast<float4*>(hi_tile + off) = hi;
    *reinterpret_cast<float4*>(lo_tile + off) = lo;
  }
}
// Transposing stage: source is row-major [32 k][128 mn] (pitch floats between k rows); the tile wants mn as rows.
//   instr t = warp*4+i: k half = t&1, mn4 pair = t>>1;  lane: k = 16*(t&1) + 4*r + (lane&3), r = (lane>>2)&3,
//   mn4 = 2*(t>>1) + (lane>>4).  Loads: 32 contiguous bytes per k row (full sectors).
__device__ __forceinline__ void load_transposed(Frag& f, const float* src, int64_t pitch, int k_valid, int warp,
                                                int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = warp * 4 + i;
    const int k = 16 * (t & 1) + 4 * ((lane >> 2) & 3) + (lane & 3);
    const int mn4 = 2 * (t >> 1) + (lane >> 4);
    f.v[i] = k < k_valid ? *reinterpret_cast<const float4*>(src + k * pitch + mn4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void store_transposed(const Frag& f, uint8_t* hi_tile, uint8_t* lo_tile, int warp, int lane) {
  const int r = (lane >> 2) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = warp * 4 + i;
    const int k = 16 * (t & 1) + 4 * r + (lane & 3);
    const int mn4 = 2 * (t >> 1) + (lane >> 4);
    float4 hi, lo;
    split_tf32(f.v[i], hi, lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j + r) & 3;  // rotated component order: bank = ((mn%8)*4 + k%4) is distinct across the warp
      const int mn = mn4 * 4 + c;
      const int off = (mn >> 3) * 1024 + (k >> 2) * 128 + (mn & 7) * 16 + (k & 3) * 4;
      const float h = c == 0 ? hi.x : c == 1 ? hi.y : c == 2 ? hi.z : hi.w;
      const float l = c == 0 ? lo.x : c == 1 ? lo.y : c == 2 ? lo.z : lo.w;
      *reinterpret_cast<float*>(hi_tile + off) = h;
      *reinterpret_cast<float*>(lo_tile + off) = l;
    }
  }
}

struct TcShared {
  uint64_t empty[STAGES];
  uint64_t done;
  uint32_t tmem_base;
  int term[MPQE_MAX_TERMS];
  int nterms;
};

__device__ __forceinline__ uint8_t* align_1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// issue the 3xTF32 products of one stage: 4 k-steps x (lo*hi, hi*lo, hi*hi), small products first
__device__ __forceinline__ void issue_stage(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            bool first) {
#pragma unroll
  for (int j = 0; j < KC / 8; ++j) {
    const uint64_t dah = make_desc(a_hi + j * 256, 128, 1024), dal = make_desc(a_lo + j * 256, 128, 1024);
    const uint64_t dbh = make_desc(b_hi + j * 256, 128, 1024), dbl = make_desc(b_lo + j * 256, 128, 1024);
    umma_tf32(tmem_d, dal, dbh, IDESC_TF32, (first && j == 0) ? 0u : 1u);
    umma_tf32(tmem_d, dah, dbl, IDESC_TF32, 1u);
    umma_tf32(tmem_d, dah, dbh, IDESC_TF32, 1u);
  }
}

__global__ void __launch_bounds__(THREADS, 1) layer_tc_kernel(const __grid_constant__ LayerLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  uint8_t* smem = align_1024(smem_raw);

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
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    int n = 0;
    for (int t = 0; t < G.num_terms; ++t)
      if (G.terms[t].out_slot == slot) sh.term[n++] = t;
    sh.nterms = n;
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&sh.empty[s]), 1);
    mbar_init(smem_u32(&sh.done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&sh.tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  const int nsteps = sh.nterms * (D / KC);

  Frag fa, fb;
  auto load_step = [&](int step) {
    const mpqe_term_t& T = G.terms[sh.term[step / (D / KC)]];
    const int kc = (step % (D / KC)) * KC;
    load_kmajor(fa, T, q0, B, kc, warp, lane);
    load_transposed(fb, T.m + (int64_t)kc * D, D, KC, warp, lane);
  };
  if (nsteps > 0) load_step(0);
  for (int step = 0; step < nsteps; ++step) {
    const int s = step % STAGES, u = step / STAGES;
    uint8_t* st = smem + s * STAGE_BYTES;
    if (u > 0) mbar_wait(smem_u32(&sh.empty[s]), (u - 1) & 1);  // MMAs that read this stage have completed
    store_kmajor(fa, st, st + TILE_BYTES, warp, lane);
    store_transposed(fb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, warp, lane);
    if (step + 1 < nsteps) load_step(step + 1);  // global loads of the next stage fly while this one is multiplied
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a = smem_u32(st);
      issue_stage(tmem, a, a + TILE_BYTES, a + 2 * TILE_BYTES, a + 3 * TILE_BYTES, step == 0);
      umma_commit(smem_u32(&sh.empty[s]));
      if (step == nsteps - 1) umma_commit(smem_u32(&sh.done));
    }
  }

  // ---- epilogue: TMEM -> registers -> bias / relu / mask -> global -------------------------------------------
  if (nsteps > 0) {
    mbar_wait(smem_u32(&sh.done), 0);
    tc_fence_after();
  }
  const int row = (warp & 3) * 32 + lane;          // TMEM lane = tile row; a warp may only touch its own 32 lanes
  const int64_t q = q0 + row;
  const int oslot = G.out_slot_map[slot];
  const float bscale = G.bias != nullptr ? G.bias_scale[slot] : 0.f;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int c0 = (warp >> 2) * 64 + half * 32;
    uint32_t v[32];
    if (nsteps > 0) {
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
    if (q < B) {
      float* orow = G.out + (q * G.out_slots + oslot) * (int64_t)D + c0;
      const float* mrow = G.epilogue == MPQE_EPI_MASK ? G.mask + (q * G.mask_slots + oslot) * (int64_t)D + c0 : nullptr;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 o = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                               __uint_as_float(v[i + 3]));
        if (G.bias != nullptr) {
          const float4 b = *reinterpret_cast<const float4*>(G.bias + c0 + i);
          o.x += bscale * b.x; o.y += bscale * b.y; o.z += bscale * b.z; o.w += bscale * b.w;
        }
        if (G.epilogue == MPQE_EPI_RELU) {
          o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
        } else if (G.epilogue == MPQE_EPI_MASK) {
          const float4 m = *reinterpret_cast<const float4*>(mrow + i);
          o = make_float4(m.x > 0.f ? o.x : 0.f, m.y > 0.f ? o.y : 0.f, m.z > 0.f ? o.z : 0.f, m.w > 0.f ? o.w : 0.f);
        }
        *reinterpret_cast<float4*>(orow + i) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient on tensor cores: dM = sum_q A[q]^T G[q]; both operands are transposed into K-major tiles.
// ------------------------------------------------------------------------------------------------------------
struct WgradIter {
  int g, t;
  int64_t q, qe;
};

__device__ __forceinline__ void chunk_range(int64_t B, int chunks, int c, int64_t& qb, int64_t& qe) {
  int64_t per = (B + chunks - 1) / chunks;
  per = (per + KC - 1) / KC * KC;
  qb = per * c;
  qe = qb + per;
  if (qb > B) qb = B;
  if (qe > B) qe = B;
}

__device__ __forceinline__ bool wgrad_seek(const WgradLaunch& L, const float* m_fwd, int chunks, int c, WgradIter& it) {
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

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  uint8_t* smem = align_1024(smem_raw);
  int unit = blockIdx.x;
  int j = 0;
  for (; j < L.num_dests - 1; ++j) {
    if (unit < L.chunks[j]) break;
    unit -= L.chunks[j];
  }
  const int c = unit, chunks = L.chunks[j];
  const float* m_fwd = L.d[j].m_fwd;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&sh.empty[s]), 1);
    mbar_init(smem_u32(&sh.done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&sh.tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  WgradIter it{0, 0, 0, 0};
  bool more = wgrad_seek(L, m_fwd, chunks, c, it);
  Frag fa, fb;
  auto load_next = [&]() {  // loads the tile the iterator points at, then advances it
    const mpqe_layer_group_t& G = L.g[it.g];
    const mpqe_term_t& T = G.terms[it.t];
    const mpqe_wgrad_operand_t& O = L.go[it.g];
    const int valid = (int)(it.qe - it.q < KC ? it.qe - it.q : KC);
    load_transposed(fa, T.a + (it.q * T.a_slots + T.a_slot) * (int64_t)D, (int64_t)T.a_slots * D, valid, warp, lane);
    const int gs = O.slot_map[T.out_slot];
    load_transposed(fb, O.g + (it.q * O.g_slots + gs) * (int64_t)D, (int64_t)O.g_slots * D, valid, warp, lane);
    it.q += KC;
    if (it.q >= it.qe) {
      ++it.t;
      more = wgrad_seek(L, m_fwd, chunks, c, it);
    }
  };
  bool have = more;
  if (have) load_next();
  int step = 0;
  while (have) {
    const int s = step % STAGES, u = step / STAGES;
    uint8_t* st = smem + s * STAGE_BYTES;
    if (u > 0) mbar_wait(smem_u32(&sh.empty[s]), (u - 1) & 1);
    store_transposed(fa, st, st + TILE_BYTES, warp, lane);
    store_transposed(fb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, warp, lane);
    have = more;
    if (have) load_next();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a = smem_u32(st);
      issue_stage(tmem, a, a + TILE_BYTES, a + 2 * TILE_BYTES, a + 3 * TILE_BYTES, step == 0);
      umma_commit(smem_u32(&sh.empty[s]));
      if (!have) umma_commit(smem_u32(&sh.done));
    }
    ++step;
  }
  if (step > 0) {
    mbar_wait(smem_u32(&sh.done), 0);
    tc_fence_after();
  }
  int pbase = 0;
  for (int jj = 0; jj < j; ++jj) pbase += L.chunks[jj];
  float* P = L.partials + (int64_t)(pbase + c) * D * D;
  const int row = (warp & 3) * 32 + lane;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int c0 = (warp >> 2) * 64 + half * 32;
    uint32_t v[32];
    if (step > 0) {
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      *reinterpret_cast<float4*>(P + row * D + c0 + i) =
          make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

int layer_forward_tc(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
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
  layer_tc_kernel<<<(unsigned)units, THREADS, TC_SMEM, stream>>>(L);
  MPQE_CHECK_LAUNCH("layer_tc_kernel");
  return 0;
}

// `launch` arrives fully prepared (groups, operands, dests, chunks, partials) from the host code in layer_simt.cu
int layer_wgrad_tc_launch(const WgradLaunch& launch, int total_chunks, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    configured = true;
  }
  wgrad_tc_kernel<<<total_chunks, THREADS, TC_SMEM, stream>>>(launch);
  MPQE_CHECK_LAUNCH("wgrad_tc_kernel");
  return 0;
}

}  // namespace mpqe

extern "C" int mpqe_b200_has_tcgen05(void) { return 1; }

// Fused R-GCN layer on tcgen05, second generation: the QUERIES are the N side of the MMA (N = 256) and the weight
// matrix the M side (M = 128 output features), i.e. the accumulator in tensor memory is D[feature, query].
//
// Why (measured on B200 with tools/tc_rate_probe.cu, profiles/r02_tc_rate_probe.txt): one tcgen05.mma.kind::tf32
// costs ~130-160 cycles whether N is 128 or 256, so a 128 x 128 x 8 instruction (the first-generation kernel,
// layer_tc.cu) runs the tensor pipe at 40 % of its rate and a 128 x 256 x 8 one at 80-90 %.  The contraction width of
// this model is fixed (d = 128 features in, 128 out), so the only dimension that can fill N = 256 is the batch:
//   D[n, q] (+)= sum_k Wt[n, k] * X[q, k],    Wt[n, k] = M[k, n]   (out[q, :] = X[q, :] @ M)
// Same contract as layer_simt.cu / layer_tc.cu (term lists, see include/mpqe_b200.h); the reference lines replaced
// are /root/reference/mpqe/model.py:292-294 (index_select + bmm), :277 (gather + scatter_add), :301-304 (root,
// bias), :437 (relu) and their autograd.  fp32 accuracy through the 3xTF32 split, as before.
//
// Persistent, warp-specialised CTA per SM (14 warps), unit = (group, out slot, 256-query tile):
//   warps 4-11  producers : activation rows global --ld.global.v4 (3 stages of loads in flight per thread)--> split
//                           hi/lo --> st.shared in the canonical K-major no-swizzle layout
//   warp  13    one thread stages the stage's weight chunk (pre-split image, mpqe_pack_weights): a 16 KB bulk copy
//   warp  12    MMA issuer: per stage (16 k) 2 k-steps x 3 products (lo*hi, hi*lo, hi*hi) of 128 x 256 x 8,
//                           tcgen05.commit -> empty[stage]; after the unit's last stage -> acc_full[buffer]
//   warps 0-3   epilogue  : thread = output feature (TMEM lane), 32 queries per tcgen05.ld; a warp's store of one
//                           query is 128 contiguous bytes of the output row, the four warps complete its 512 bytes --
//                           no shared-memory transpose.  Two 256-column accumulators ping-pong (all 512 TMEM columns).
#include <stdlib.h>

#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mpqe {

namespace {

using namespace tc;

constexpr int EPI_WARPS = 4;
constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS + PROD_WARPS;
constexpr int TMA_WARP = MMA_WARP + 1;             // one thread of it stages the weight chunks (bulk copies)
constexpr int XLOAD_WARP = TMA_WARP + 1;           // TMAX kernel: one thread of it loads the activation tiles (TMA)
constexpr int THREADS = (TMA_WARP + 1) * 32;     // 448
constexpr int THREADS_X = (XLOAD_WARP + 1) * 32; // 480
constexpr int BQ = 256;                          // queries per unit = UMMA N
constexpr int KC = 16;                           // k per pipeline stage = 2 UMMA k-steps of 8
constexpr int STAGES = 3;
constexpr int LOADS_AHEAD = 3;                   // stages of activation rows in flight (cp.async) ahead of the stores
constexpr int RING = LOADS_AHEAD + 1;            // raw-row ring slots
constexpr int RING_SLOT = 256 * 4 * 16;          // 16 KB: one stage of raw rows, [i][thread] 16-byte pieces
constexpr int W_TILE = 128 * KC * 4;             // 8 KB: weight chunk, hi or lo      [128 n][16 k]
constexpr int X_TILE = BQ * KC * 4;              // 16 KB: activation chunk, hi or lo [256 q][16 k]
constexpr int STAGE_BYTES = 2 * W_TILE + 2 * X_TILE;   // W_hi | W_lo | X_hi | X_lo = 48 KB
constexpr int TMEM_COLS = 512;                   // two 256-column fp32 accumulators
constexpr size_t TC2_SMEM = size_t(STAGES) * STAGE_BYTES + size_t(RING) * RING_SLOT + 1024;
// both operands K-major: element (row, k) of a [rows][16 k] chunk at (row/8)*512 + (k/4)*128 + (row%8)*16 + (k%4)*4;
// descriptor of k-step j (k = 8j .. 8j+7): start + j*256, LBO = 128 (next 4 k), SBO = 512 (next 8 rows)
constexpr uint32_t LBO = 128, SBO = 512;
// instruction descriptor: c = f32, a = b = tf32, both K-major, N = 256, M = 128
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct Shared2 {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t raw_full[RING];    // TMAX kernel: a raw activation tile has landed in its ring slot
  uint64_t raw_empty[RING];   //              ... and has been read by every producer warp
  uint32_t tmem_base;
};

// TMAX kernel: the activation operands as 3-D tensor maps {128 k, slots, queries} (fp32, box {16 k, 1 slot, 256 queries},
// 64-byte swizzle); of[group][term] = map of that term's operand
constexpr int MAX_XMAPS = 24;
struct alignas(64) XMaps {
  CUtensorMap map[MAX_XMAPS];
  uint8_t of[MPQE_MAX_GROUPS][MPQE_MAX_TERMS];
};


__device__ __forceinline__ uint8_t* align_1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

struct Unit2 {
  int gi, slot;
  int64_t q0;
};

// k-th unit of this CTA from the host-built schedule (see tc_common.cuh): tile = 256-query tile index
__device__ __forceinline__ bool next_unit(const Schedule& S, int k, Unit2& U, uint32_t& mask) {
  const int i = S.start[blockIdx.x] + k;
  if (i >= S.start[blockIdx.x + 1]) return false;
  const uint32_t g = S.gsm[i];
  U.gi = (int)(g >> 24);
  U.slot = (int)((g >> 16) & 0xffu);
  U.q0 = (int64_t)S.tile[i] * BQ;
  mask = g & 0xffffu;
  return true;
}

// x = hi + lo with hi = x rounded to tf32 (nearest, ties away from zero -- cvt.rna.tf32.f32 for finite x) in two integer
// instructions instead of the five of the cvt (which also handles NaN / infinity): activations are finite
__device__ __forceinline__ void split_tf32_fast(const float4& x, float4& hi, float4& lo) {
  hi.x = __uint_as_float((__float_as_uint(x.x) + 0x1000u) & 0xffffe000u);
  hi.y = __uint_as_float((__float_as_uint(x.y) + 0x1000u) & 0xffffe000u);
  hi.z = __uint_as_float((__float_as_uint(x.z) + 0x1000u) & 0xffffe000u);
  hi.w = __uint_as_float((__float_as_uint(x.w) + 0x1000u) & 0xffffe000u);
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// timing experiments (debug builds only; results are wrong when non-zero): 4 = no producer stores, 8 = no MMAs,
// 16 = no weight bulk copies, 32 = no epilogue stores
#ifndef MPQE_TC2_ABLATE
#define MPQE_TC2_ABLATE 0
#endif

// ---- optional per-role cycle accounting (debug builds: -DMPQE_TC_STATS + mpqe_debug_set_stats2) ---------------------
#ifdef MPQE_TC_STATS
__device__ long long* g_stats2 = nullptr;   // [CTA][16] cycle totals
__device__ int g_dbg2 = 0;                  // timing experiments (wrong results): 4 = no producer stores, 8 = no MMAs,
                                            // 16 = no weight bulk copies
#define ST_DECL long long st_t0 = 0, st_acc[4] = {0, 0, 0, 0}
#define ST_BEGIN() st_t0 = clock64()
#define ST_END(i) st_acc[i] += clock64() - st_t0
#define ST_FLUSH(base, n)                                                             \
  if (g_stats2 != nullptr)                                                            \
    for (int i_ = 0; i_ < (n); ++i_) g_stats2[(long long)blockIdx.x * 16 + (base) + i_] = st_acc[i_]
#else
#define ST_DECL
#define ST_BEGIN()
#define ST_END(i)
#define ST_FLUSH(base, n)
#endif

// One 32-query block of the epilogue for one output feature: v[i] = accumulator of query i.  MODE 0: + bias;
// 1: + bias, ReLU (BITS: also returns the sign bits); 2: + bias, keep where bit i of `keep` is set (ReLU backward).
// Kept to ~5 instructions per element: there is one epilogue warp per scheduler, so its instruction count is the
// epilogue's run time (the first version spent ~30 instructions per element on 64-bit address products and
// per-element branches: 18 k cycles per unit, more than the unit's MMAs).
template <int MODE, bool BITS>
__device__ __forceinline__ uint32_t epi_block(const uint32_t (&v)[32], float bv, uint32_t keep, float* o, int out_step,
                                              int valid) {
  uint32_t pos = 0u;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float t = __uint_as_float(v[i]) + bv;
    if (MODE == 1) t = fmaxf(t, 0.f);
    if (MODE == 2) t = ((keep >> i) & 1u) ? t : 0.f;
    if (BITS) pos |= (t > 0.f ? 1u : 0u) << i;
    if (i < valid) o[i * out_step] = t;
  }
  return pos;
}

// TMAX: the raw activation tiles are brought into the ring by TMA tensor loads issued by one thread (warp XLOAD_WARP)
// instead of by cp.async from the eight producer warps.
template <bool TMAX>
__global__ void __launch_bounds__(THREADS_X, 1) layer_tc2_kernel(const __grid_constant__ LayerLaunch L,
                                                                 const __grid_constant__ Schedule S,
                                                                 const __grid_constant__ XMaps XM) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Shared2 sh;
  uint8_t* smem = align_1024(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef MPQE_TC_STATS
  const long long kernel_t0 = clock64();
#endif

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&sh.full[s]), PROD_THREADS + 1);   // producers + the weight stager
      mbar_init(smem_u32(&sh.empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&sh.acc_full[b]), 1);
      mbar_init(smem_u32(&sh.acc_empty[b]), EPI_WARPS * 32);
    }
    for (int r = 0; r < RING; ++r) {
      mbar_init(smem_u32(&sh.raw_full[r]), 1);
      mbar_init(smem_u32(&sh.raw_empty[r]), PROD_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&sh.tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp >= EPI_WARPS && warp < MMA_WARP) {
    // ===== producers ==================================================================================================
    // The activation rows come from HBM with ~1000 cycles of latency (measured: with everything else switched off the
    // wait for the rows alone set the stage time).  They are therefore fetched LOADS_AHEAD stages ahead with cp.async
    // (global -> shared without registers) into a thread-private ring -- a thread later reads back exactly the 16-byte
    // pieces it copied, so cp.async.wait_group is the only synchronisation the ring needs -- and one copy of the stage
    // code (load back, split hi/lo, store into the operand tiles) serves all stages: with 14 warps in four roles the
    // kernel is sensitive to instruction-cache footprint, a software pipeline over register sets tripled it.
    // Thread (pw, lane) owns rows  8*(4 pw + i) + (lane & 7), i = 0..3,  and the k-quad  lane >> 3 of every stage:
    // a warp instruction reads 8 rows x 64 contiguous bytes and writes 512 contiguous bytes of the tile.
    const int pw = warp - EPI_WARPS;
    const int kq = lane >> 3;
    const int ptid = tid - EPI_WARPS * 32;                                // 0..255
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t my_off = pw * 4 * 512 + kq * 128 + (lane & 7) * 16;   // this thread's 16 bytes of row group 4 pw
    if constexpr (TMAX) {
      // the tiles arrive by TMA as [256 rows][64 bytes] with the 64-byte swizzle (16-byte chunk c of row r at chunk
      // c ^ ((r >> 1) & 3)): with lane & 7 = row and lane >> 3 = chunk a quarter-warp's 16-byte reads fall into eight
      // different bank groups, and the stores into the operand tiles are the same as in the cp.async version
      uint32_t nst = 0;
      {
        Unit2 U0{0, 0, 0};
        uint32_t m0;
        for (int k = 0; next_unit(S, k, U0, m0); ++k) nst += __popc(m0) * (D / KC);
      }
      const uint32_t ring0 = smem_base + STAGES * STAGE_BYTES;
      const uint32_t src_off = (uint32_t)(pw * 4 * 8 + (lane & 7)) * 64 + (uint32_t)((kq ^ ((lane >> 1) & 3)) * 16);
      ST_DECL;
#pragma unroll 1
      for (uint32_t it = 0; it < nst; ++it) {
        const int slot = it % RING;
        ST_BEGIN();
        mbar_wait(smem_u32(&sh.raw_full[slot]), (it / RING) & 1);
        ST_END(3);   // waiting for the rows
        const int s = it % STAGES;
        const uint32_t use = it / STAGES;
        ST_BEGIN();
        if (use > 0) mbar_wait(smem_u32(&sh.empty[s]), (use - 1) & 1);
        ST_END(0);   // waiting for a free stage
        ST_BEGIN();
        const uint32_t src = ring0 + slot * RING_SLOT + src_off;
        const uint32_t xh = smem_base + s * STAGE_BYTES + 2 * W_TILE + my_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 x, hi, lo;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                       : "r"(src + i * 512));
          split_tf32_fast(x, hi, lo);
          sts128(xh + i * 512, hi);
          sts128(xh + i * 512 + X_TILE, lo);
        }
        ST_END(1);   // load back + split + stores
        ST_BEGIN();
        fence_proxy_async();
        mbar_arrive(smem_u32(&sh.full[s]));
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sh.raw_empty[slot]));   // this warp has read its rows of the slot
        ST_END(2);   // fence + arrive
      }
      if (pw == 0 && lane == 0) { ST_FLUSH(0, 4); }
    } else {
    const uint32_t ring = smem_base + STAGES * STAGE_BYTES + ptid * 16;  // raw ring: [slot][i][thread] 16-byte pieces
    // load-side iterator over (unit, term, k chunk), LOADS_AHEAD stages ahead of the stores
    int uk = 0;
    uint32_t mask = 0u;
    int kc = D - KC;
    bool alive = true;
    Unit2 U{0, 0, 0};
    uint32_t row_off[4];          // element offset of this thread's four rows (k-quad included) in the term's operand
    const float* a = nullptr;
    uint32_t issued = 0;
    auto fetch_next = [&]() {     // cp.async of the next stage's rows into ring slot issued % RING; always commits
      if (alive) {
        kc += KC;
        if (kc >= D) {
          kc = 0;
          mask &= mask - 1;
          while (mask == 0 && alive) alive = next_unit(S, uk++, U, mask);
          if (alive) {
            const mpqe_layer_group_t& G = L.g[U.gi];
            const mpqe_term_t& T = G.terms[__ffs(mask) - 1];
            a = T.a;
            const int64_t a_slots = T.a_slots, a_slot = T.a_slot, nq = G.num_queries;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              int64_t q = U.q0 + (pw * 4 + i) * 8 + (lane & 7);
              if (q >= nq) q = nq - 1;                    // rows past the end repeat the last row (never stored)
              row_off[i] = (uint32_t)((q * a_slots + a_slot) * (int64_t)D + kq * 4);
            }
          }
        }
      }
      if (alive) {
        const uint32_t dst = ring + (issued % RING) * RING_SLOT;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * (PROD_THREADS * 16)),
                       "l"(a + row_off[i] + kc)
                       : "memory");
        ++issued;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int j = 0; j < LOADS_AHEAD; ++j) fetch_next();
    ST_DECL;
#pragma unroll 1
    for (uint32_t it = 0;; ++it) {
#ifdef MPQE_TC_STATS_FETCH                                                  // (debug: slot 3 = time in fetch_next)
      ST_BEGIN();
      fetch_next();
      ST_END(3);
      if (it >= issued) break;
      asm volatile("cp.async.wait_group %0;" ::"n"(LOADS_AHEAD) : "memory");
#else
      fetch_next();                                                        // stage it + LOADS_AHEAD
      if (it >= issued) break;                                             // nothing left: all issued stages stored
      ST_BEGIN();
      asm volatile("cp.async.wait_group %0;" ::"n"(LOADS_AHEAD) : "memory");   // stage `it` has landed (own pieces)
      ST_END(3);   // waiting for the rows
#endif
      const int s = it % STAGES;
      const uint32_t use = it / STAGES;
      ST_BEGIN();
      if (use > 0) mbar_wait(smem_u32(&sh.empty[s]), (use - 1) & 1);
      ST_END(0);   // waiting for a free stage
      ST_BEGIN();
      const uint32_t src = ring + (it % RING) * RING_SLOT;
      const uint32_t xh = smem_base + s * STAGE_BYTES + 2 * W_TILE + my_off;
#if !(MPQE_TC2_ABLATE & 4)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 x, hi, lo;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                     : "r"(src + i * (PROD_THREADS * 16)));
        split_tf32_fast(x, hi, lo);
        sts128(xh + i * 512, hi);
        sts128(xh + i * 512 + X_TILE, lo);
      }
#else
      (void)src; (void)xh;
#endif
      ST_END(1);   // load back + split + stores
      ST_BEGIN();
      fence_proxy_async();
      mbar_arrive(smem_u32(&sh.full[s]));
      ST_END(2);   // fence + arrive
    }
    if (pw == 0 && lane == 0) { ST_FLUSH(0, 4); }
    }   // !TMAX
  } else if (TMAX && warp == XLOAD_WARP) {
    // ===== activation loader (one thread): one 16 KB tensor load per stage, RING stages ahead =========================
    if (lane == 0) {
      uint32_t it = 0;
      int uc = 0;
      Unit2 U{0, 0, 0};
      const uint32_t ring0 = smem_u32(smem) + STAGES * STAGE_BYTES;
      for (uint32_t umask; next_unit(S, uc, U, umask); ++uc) {
        const mpqe_layer_group_t& G = L.g[U.gi];
        for (uint32_t m = umask; m != 0; m &= m - 1) {
          const int t = __ffs(m) - 1;
          const CUtensorMap* map = &XM.map[XM.of[U.gi][t]];
          const int a_slot = G.terms[t].a_slot;
          for (int c = 0; c < D / KC; ++c, ++it) {
            const int slot = it % RING;
            const uint32_t use = it / RING;
            if (use > 0) mbar_wait(smem_u32(&sh.raw_empty[slot]), (use - 1) & 1);
            const uint32_t bar = smem_u32(&sh.raw_full[slot]);
            mbar_arrive_expect_tx(bar, X_TILE);
            tma_load_3d(ring0 + slot * RING_SLOT, map, c * KC, a_slot, (int)U.q0, bar);
          }
        }
      }
    }
  } else if (warp == TMA_WARP) {
    // ===== weight stager (one thread): one 16 KB bulk copy of the pre-split chunk per stage ===========================
    // (a thread of its own: issued from a producer warp, the copy instruction held that warp -- and with it every
    // stage -- for ~1000 cycles)
    if (lane == 0) {
      uint32_t it = 0;
      int uc = 0;
      Unit2 U{0, 0, 0};
      const uint32_t base = smem_u32(smem);
      for (uint32_t umask; next_unit(S, uc, U, umask); ++uc) {
        const mpqe_layer_group_t& G = L.g[U.gi];
        for (uint32_t m = umask; m != 0; m &= m - 1) {
          const float* packed = G.terms[__ffs(m) - 1].m_packed;
          for (int c = 0; c < D / KC; ++c, ++it) {
            const int s = it % STAGES;
            const uint32_t use = it / STAGES;
            if (use > 0) mbar_wait(smem_u32(&sh.empty[s]), (use - 1) & 1);
            const uint32_t bar = smem_u32(&sh.full[s]);
#if MPQE_TC2_ABLATE & 16
            mbar_arrive(bar);
            (void)packed;
            (void)base;
#else
            mbar_arrive_expect_tx(bar, 2 * W_TILE);
            bulk_copy_g2s(base + s * STAGE_BYTES, packed + c * (2 * W_TILE / 4), 2 * W_TILE, bar);
#endif
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer (one thread) ====================================================================================
    if (lane == 0) {
      uint32_t it = 0;
      int uc = 0;
      Unit2 U{0, 0, 0};
      const uint32_t base = smem_u32(smem);
      ST_DECL;
      for (uint32_t umask; next_unit(S, uc, U, umask); ++uc) {
        const int nsteps = __popc(umask) * (D / KC);
        const int ab = uc & 1, use = uc >> 1;
        ST_BEGIN();
        if (use > 0) mbar_wait(smem_u32(&sh.acc_empty[ab]), (use - 1) & 1);   // epilogue drained this accumulator
        ST_END(0);
        tc_fence_after();
        const uint32_t acc = tmem + ab * BQ;
        for (int step = 0; step < nsteps; ++step, ++it) {
          const int s = it % STAGES;
          ST_BEGIN();
          mbar_wait(smem_u32(&sh.full[s]), (it / STAGES) & 1);
          ST_END(1);
          ST_BEGIN();
          tc_fence_after();
          const uint32_t w_hi = base + s * STAGE_BYTES, w_lo = w_hi + W_TILE, x_hi = w_hi + 2 * W_TILE,
                         x_lo = x_hi + X_TILE;
#if !(MPQE_TC2_ABLATE & 8)
#pragma unroll
          for (int j = 0; j < KC / 8; ++j) {
            const uint64_t dwh = make_desc(w_hi + j * 256, LBO, SBO), dwl = make_desc(w_lo + j * 256, LBO, SBO);
            const uint64_t dxh = make_desc(x_hi + j * 256, LBO, SBO), dxl = make_desc(x_lo + j * 256, LBO, SBO);
            umma_tf32(acc, dwl, dxh, IDESC2, (step == 0 && j == 0) ? 0u : 1u);   // small products first
            umma_tf32(acc, dwh, dxl, IDESC2, 1u);
            umma_tf32(acc, dwh, dxh, IDESC2, 1u);
          }
#else
          (void)w_lo; (void)x_lo; (void)acc;
#endif
          umma_commit(smem_u32(&sh.empty[s]));      // frees the stage when the MMAs have read it
          ST_END(2);
        }
        umma_commit(smem_u32(&sh.acc_full[ab]));    // accumulator complete
      }
#ifdef MPQE_TC_STATS
      st_acc[3] = it;
      ST_FLUSH(4, 4);
      if (g_stats2 != nullptr) g_stats2[(long long)blockIdx.x * 16 + 8] = uc;
#endif
    }
  } else {
    // ===== epilogue: thread = output feature ==========================================================================
    int uc = 0;
    Unit2 U{0, 0, 0};
    const int n = warp * 32 + lane;                 // this thread's feature = its TMEM lane
    ST_DECL;
    for (uint32_t umask; next_unit(S, uc, U, umask); ++uc) {
      const mpqe_layer_group_t& G = L.g[U.gi];
      const int nsteps = __popc(umask) * (D / KC);
      const int ab = uc & 1;
      const int oslot = G.out_slot_map[U.slot];
      const int epi = G.epilogue;
      const bool masked = epi == MPQE_EPI_MASK;
      const int64_t left64 = G.num_queries - U.q0;                // query c of the tile exists iff c < rows_left
      const int rows_left = left64 < BQ ? (int)left64 : BQ;
      // Strides as 32-bit element counts: with 64-bit products the address chain of every store (three IMADs, two LEAs
      // behind a branch) ran at ~70 cycles per store on the one epilogue warp of each scheduler -- 18 k cycles per
      // unit, more than the unit's MMAs
      const int out_step = G.out_slots * D, mask_step = G.mask_slots * D;
      float* outp = G.out + (U.q0 * (int64_t)G.out_slots + oslot) * (int64_t)D + n;
      const float* maskp = masked ? G.mask + (U.q0 * (int64_t)G.mask_slots + oslot) * (int64_t)D + n : nullptr;
      float bv = 0.f;
      if (G.bias != nullptr) bv = G.bias_scale[U.slot] * __ldg(G.bias + (int64_t)U.slot * G.bias_slot_stride + n);
      // ReLU-backward mask.  Preferred form: the sign bits written by the forward launch (one word per 32 queries and
      // feature: 8 loads per unit, all issued up front).  Otherwise the fp32 activations, a 32-query block fetched one
      // block ahead through the read-only path (inside the store loop the loads could not be hoisted above the stores).
      const uint32_t* bitsp = masked && G.mask_bits != nullptr
                                  ? G.mask_bits + ((U.q0 / 32) * (int64_t)G.mask_slots + oslot) * (int64_t)D + n : nullptr;
      uint32_t* bits_out = epi == MPQE_EPI_RELU && G.relu_bits_out != nullptr
                               ? G.relu_bits_out + ((U.q0 / 32) * (int64_t)G.out_slots + oslot) * (int64_t)D + n : nullptr;
      uint32_t kb = bitsp != nullptr ? __ldg(bitsp) : 0u;       // sign bits of the current block (next one prefetched)
      float mk[32];
      auto fetch = [&](int c0) {
        if (masked && bitsp == nullptr) {
          const float* mp = maskp + c0 * mask_step;
          const int valid = rows_left - c0;
#pragma unroll
          for (int i = 0; i < 32; ++i) mk[i] = i < valid ? __ldg(mp + i * mask_step) : 0.f;
        }
      };
      fetch(0);
      ST_BEGIN();
      mbar_wait(smem_u32(&sh.acc_full[ab]), (uc >> 1) & 1);
      ST_END(0);
      ST_BEGIN();
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < BQ / 32; ++j) {
        const int c0 = 32 * j;
        if (c0 >= rows_left) break;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ab * BQ + c0, v);
        if (nsteps == 0) {                       // a slot no term writes to: bias only
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        uint32_t keep = 0xffffffffu;
        if (bitsp != nullptr) {
          keep = kb;
          if (c0 + 32 < rows_left) kb = __ldg(bitsp + (j + 1) * mask_step);
        } else if (masked) {
          keep = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) keep |= (mk[i] > 0.f ? 1u : 0u) << i;
          if (c0 + 32 < rows_left) fetch(c0 + 32);
        }
        float* o = outp + c0 * out_step;
        const int valid = (MPQE_TC2_ABLATE & 32) ? 0 : rows_left - c0;     // >= 32 for whole blocks
        if (epi == MPQE_EPI_RELU) {
          if (bits_out != nullptr) {
            uint32_t pos = epi_block<1, true>(v, bv, keep, o, out_step, valid);
            if (valid < 32) pos &= (1u << valid) - 1u;
            bits_out[j * out_step] = pos;
          } else {
            epi_block<1, false>(v, bv, keep, o, out_step, valid);
          }
        } else if (masked) {
          epi_block<2, false>(v, bv, keep, o, out_step, valid);
        } else {
          epi_block<0, false>(v, bv, keep, o, out_step, valid);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.acc_empty[ab]));
      ST_END(1);
    }
    if (tid == 0) { ST_FLUSH(9, 2); }
#ifdef MPQE_TC_STATS
    if (tid == 0 && g_stats2 != nullptr) g_stats2[(long long)blockIdx.x * 16 + 12] = st_acc[2];
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
#ifdef MPQE_TC_STATS
  if (tid == 0 && g_stats2 != nullptr) g_stats2[(long long)blockIdx.x * 16 + 11] = clock64() - kernel_t0;
#endif
}

// Pre-split weights for this kernel: out[m][16-k chunk c] = [hi 8 KB | lo 8 KB] of Wt[n][k] = M[k][n], the byte image
// of the shared-memory operand chunk (element (n, kk) at (n/8)*512 + (kk/4)*128 + (n%8)*16 + (kk%4)*4).
constexpr int PACK_MAX = 256;
struct PackLaunch2 {
  const float* m[PACK_MAX];
};

__global__ void __launch_bounds__(256) pack_weights2_kernel(const __grid_constant__ PackLaunch2 P,
                                                            float* __restrict__ out) {
  // bit 0 of the pointer: pack the image of M^T (what the input-gradient launches multiply by) straight from M, so the
  // per-step weight preparation does not have to wait for a transposed copy
  const uintptr_t tagged = reinterpret_cast<uintptr_t>(P.m[blockIdx.y]);
  const bool transposed = (tagged & 1) != 0;
  const float* M = reinterpret_cast<const float*>(tagged & ~(uintptr_t)1);
  const int c = blockIdx.x;                          // k chunk (16 k)
  float* hi_tile = out + ((int64_t)blockIdx.y * (D / KC) + c) * (2 * W_TILE / 4);
  float* lo_tile = hi_tile + W_TILE / 4;
  for (int e = threadIdx.x; e < 128 * 4; e += 256) {   // e -> (n, k quad): lanes run over n (coalesced reads of M rows)
    const int n = e & 127, q = e >> 7;
    float4 x;
    if (transposed) {                                // operand element (n, k) = M^T[k][n] = M[n][k]: 16 contiguous bytes
      x = *reinterpret_cast<const float4*>(M + (int64_t)n * D + c * KC + q * 4);
    } else {
      x.x = M[(int64_t)(c * KC + q * 4 + 0) * D + n];
      x.y = M[(int64_t)(c * KC + q * 4 + 1) * D + n];
      x.z = M[(int64_t)(c * KC + q * 4 + 2) * D + n];
      x.w = M[(int64_t)(c * KC + q * 4 + 3) * D + n];
    }
    float4 hi, lo;
    split_tf32(x, hi, lo);
    const int off = ((n >> 3) * 512 + q * 128 + (n & 7) * 16) / 4;
    *reinterpret_cast<float4*>(hi_tile + off) = hi;
    *reinterpret_cast<float4*>(lo_tile + off) = lo;
  }
}

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace

// Tensor maps of the activation operands of one launch.  Returns false when the launch has to use the cp.async loader:
// broadcast operands (a_slots == 0: a tensor map has no zero stride), more distinct operands than MAX_XMAPS, a driver
// without cuTensorMapEncodeTiled, or MPQE_LAYER_LOADS=cpasync (A/B measurements).
static bool build_xmaps(const mpqe_layer_group_t* groups, int num_groups, XMaps& XM) {
  static int mode = -1;   // 0: cp.async, 1: TMA
  if (mode < 0) {
    const char* e = getenv("MPQE_LAYER_LOADS");
    mode = (e != nullptr && strcmp(e, "cpasync") == 0) || tensor_map_encoder() == nullptr ? 0 : 1;
  }
  if (mode == 0) return false;
  struct Key {
    const float* a;
    int32_t slots;
    int64_t nq;
  };
  Key keys[MAX_XMAPS];
  int n = 0;
  for (int i = 0; i < num_groups; ++i)
    for (int t = 0; t < groups[i].num_terms; ++t) {
      const mpqe_term_t& T = groups[i].terms[t];
      if (T.a_slots <= 0 || (reinterpret_cast<uintptr_t>(T.a) & 15) != 0) return false;
      int k = 0;
      while (k < n && !(keys[k].a == T.a && keys[k].slots == T.a_slots && keys[k].nq == groups[i].num_queries)) ++k;
      if (k == n) {
        if (n == MAX_XMAPS) return false;
        keys[n++] = Key{T.a, T.a_slots, groups[i].num_queries};
        if (!encode_rows_map(&XM.map[k], T.a, T.a_slots, groups[i].num_queries, KC, BQ, CU_TENSOR_MAP_SWIZZLE_64B))
          return false;
      }
      XM.of[i][t] = (uint8_t)k;
    }
  return true;
}

static int launch_tc2(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream) {
  static thread_local LayerLaunch L;
  memset(&L, 0, sizeof(L));
  L.num_groups = num_groups;
  int64_t units = 0;
  for (int i = 0; i < num_groups; ++i) {
    L.g[i] = groups[i];
    units += (groups[i].num_queries + BQ - 1) / BQ * groups[i].num_out_slots;
  }
  const int grid = units < sm_count() ? (int)units : sm_count();
  MPQE_CHECK_ARG(grid <= SCHED_MAX_CTAS, "mpqe_layer_forward (tcgen05): %d SMs", grid);
  // longest-processing-time-first assignment of units to the persistent CTAs (cost = terms of the unit's slot)
  static thread_local Schedule S;
  static thread_local int cost[SCHED_MAX_UNITS];
  static thread_local uint16_t tile_of[SCHED_MAX_UNITS];
  static thread_local uint32_t gsm_of[SCHED_MAX_UNITS];
  int u = 0;
  for (int i = 0; i < num_groups; ++i) {
    const int tiles = (int)((groups[i].num_queries + BQ - 1) / BQ);
    for (int slot = 0; slot < groups[i].num_out_slots; ++slot) {
      int nt = 0;
      uint32_t mask = 0;
      for (int t = 0; t < groups[i].num_terms; ++t)
        if (groups[i].terms[t].out_slot == slot) ++nt, mask |= 1u << t;
      for (int k = 0; k < tiles; ++k) {
        // in stages: 8 per term, and the unit's epilogue (256 x 512 bytes of stores) costs about as much as one term
        cost[u] = nt * (D / KC) + D / KC;
        tile_of[u] = (uint16_t)k;
        gsm_of[u] = ((uint32_t)i << 24) | ((uint32_t)slot << 16) | mask;
        ++u;
      }
    }
  }
  build_lpt(S, cost, (int)units, grid);
  for (int pos = 0; pos < (int)units; ++pos) {
    S.tile[pos] = tile_of[S.unit[pos]];
    S.gsm[pos] = gsm_of[S.unit[pos]];
  }
  static thread_local XMaps XM;
  if (build_xmaps(groups, num_groups, XM)) {
    layer_tc2_kernel<true><<<grid, THREADS_X, TC2_SMEM, stream>>>(L, S, XM);
  } else {     // (operands a tensor map cannot describe, or MPQE_LAYER_LOADS=cpasync)
    layer_tc2_kernel<false><<<grid, THREADS, TC2_SMEM, stream>>>(L, S, XM);
  }
  MPQE_CHECK_LAUNCH("layer_tc2_kernel");
  return 0;
}

int layer_forward_tc2(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(layer_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM));
    MPQE_CUDA(cudaFuncSetAttribute(layer_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM));
    configured = true;
  }
  int slots = 0;
  int64_t longest = 0;
  for (int i = 0; i < num_groups; ++i) {
    slots += groups[i].num_out_slots;
    if (groups[i].num_queries > longest) longest = groups[i].num_queries;
    for (int t = 0; t < groups[i].num_terms; ++t)
      MPQE_CHECK_ARG(groups[i].terms[t].m_packed != nullptr,
                     "mpqe_layer_forward (tcgen05): group %d term %d has no packed weight image (mpqe_pack_weights)", i, t);
  }
  if (slots == 0 || longest == 0) return 0;
  // the per-launch schedule holds SCHED_MAX_UNITS units: longer batches go through in slices of whole 256-query tiles
  const int64_t tiles_per_launch = SCHED_MAX_UNITS / slots;
  MPQE_CHECK_ARG(tiles_per_launch >= 1, "mpqe_layer_forward (tcgen05): too many output slots");
  const int64_t slice = tiles_per_launch * BQ;
  for (int64_t q0 = 0; q0 < longest; q0 += slice) {
    mpqe_layer_group_t part[MPQE_MAX_GROUPS];
    int n = 0;
    for (int i = 0; i < num_groups; ++i) {
      if (groups[i].num_queries <= q0) continue;
      mpqe_layer_group_t g = groups[i];
      g.num_queries = groups[i].num_queries - q0 < slice ? groups[i].num_queries - q0 : slice;
      for (int t = 0; t < g.num_terms; ++t) g.terms[t].a += q0 * g.terms[t].a_slots * (int64_t)D;
      g.out += q0 * g.out_slots * (int64_t)D;
      if (g.mask != nullptr) g.mask += q0 * g.mask_slots * (int64_t)D;
      part[n++] = g;
    }
    if (int rc = launch_tc2(part, n, stream)) return rc;
  }
  return 0;
}

int pack_weights_tc2(const float* const* mats_host, const uint8_t* transposed_host, int32_t count, float* packed,
                     cudaStream_t stream) {
  for (int base = 0; base < count; base += PACK_MAX) {
    static thread_local PackLaunch2 P;
    const int n = count - base < PACK_MAX ? count - base : PACK_MAX;
    for (int i = 0; i < n; ++i) {
      MPQE_CHECK_ARG(mats_host[base + i] != nullptr, "mpqe_pack_weights: matrix %d is null", base + i);
      const uintptr_t tag = transposed_host != nullptr && transposed_host[base + i] ? 1 : 0;
      MPQE_CHECK_ARG((reinterpret_cast<uintptr_t>(mats_host[base + i]) & 15) == 0,
                     "mpqe_pack_weights: matrix %d is not 16-byte aligned", base + i);
      P.m[i] = reinterpret_cast<const float*>(reinterpret_cast<uintptr_t>(mats_host[base + i]) | tag);
    }
    pack_weights2_kernel<<<dim3(D / KC, n), 256, 0, stream>>>(P, packed + (int64_t)base * MPQE_PACKED_FLOATS);
    MPQE_CHECK_LAUNCH("pack_weights2_kernel");
  }
  return 0;
}

}  // namespace mpqe

#ifdef MPQE_TC_STATS
// debug only (not part of the public header): [148][16] int64 device buffer receiving per-CTA cycle totals
extern "C" __attribute__((visibility("default"))) int mpqe_debug_set_stats2(void* buf) {
  return (int)cudaMemcpyToSymbol(mpqe::g_stats2, &buf, sizeof(buf));
}
extern "C" __attribute__((visibility("default"))) int mpqe_debug_set_dbg2(int bits) {
  return (int)cudaMemcpyToSymbol(mpqe::g_dbg2, &bits, sizeof(bits));
}
#endif

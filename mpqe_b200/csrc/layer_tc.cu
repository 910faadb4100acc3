// Fused R-GCN layer on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), fp32-accurate through
// a 3xTF32 split:  A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  with  x_hi = tf32(x), x_lo = x - x_hi.
//
// Same contract as layer_simt.cu (term-list GEMM / term-list outer product, see include/mpqe_b200.h); the reference
// lines replaced are /root/reference/mpqe/model.py:292-294 (index_select + bmm), :277 (gather + scatter_add),
// :301-304 (root, bias), :437 (relu) and their autograd.
//
// Persistent, warp-specialised CTAs (one per SM, 13 warps):
//   warps 4-11  producers : global fp32 --ld.global.v4--> registers --split hi/lo--> st.shared in the UMMA canonical
//                           K-major no-swizzle layout --fence.proxy.async--> mbarrier full[stage]
//   warp  12    MMA issuer: waits full[stage], issues 12 x tcgen05.mma.kind::tf32 (M=128, N=128, K=8),
//                           tcgen05.commit -> empty[stage]; after a unit's last stage commit -> acc_full[buffer]
//   warps 0-3   epilogue  : waits acc_full, tcgen05.ld 32x32b (one accumulator row per thread), bias / relu / mask,
//                           global stores, arrives acc_empty  (two 128-column TMEM accumulators ping-pong, so the
//                           epilogue of unit u overlaps the main loop of unit u+1)
// The operands are staged by threads rather than by TMA because every element has to pass through registers once
// anyway to be split into its hi/lo TF32 parts (and half of them need a transpose on the way).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace mpqe {

namespace {

constexpr int EPI_WARPS = 4;
constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS + PROD_WARPS;
constexpr int THREADS = (MMA_WARP + 1) * 32;   // 416
constexpr int BM = 128;              // rows per unit = UMMA M
constexpr int KC = 32;               // k per pipeline stage = 4 UMMA k-steps of 8
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * KC * 4;          // 16 KB: one operand tile (hi or lo)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int TMEM_COLS = 256;                   // two 128-column fp32 accumulators
constexpr size_t TC_SMEM = size_t(STAGES) * STAGE_BYTES + 1024;  // + alignment slack

// ---- optional event trace of CTA 0 (debug tool: mpqe_debug_set_trace) ----------------------------------------
// Compiled in only with -DMPQE_TC_TRACE: every event costs a global atomic (~800 cycles).
__device__ long long* g_trace = nullptr;  // [0] = event count, then (tag, value, clock64) triples
__device__ __forceinline__ void trace(int tag, int value) {
#ifdef MPQE_TC_TRACE
  if (g_trace != nullptr && blockIdx.x == 0) {
    const unsigned long long i = atomicAdd(reinterpret_cast<unsigned long long*>(g_trace), 1ull);
    if (i < 8000) {
      g_trace[1 + 3 * i] = tag;
      g_trace[2 + 3 * i] = value;
      g_trace[3 + 3 * i] = clock64();
    }
  }
#else
  (void)tag;
  (void)value;
#endif
}

// ---- optional per-role cycle accounting (debug tool, -DMPQE_TC_STATS + mpqe_debug_set_stats) ------------------
__device__ long long* g_stats = nullptr;  // [CTA][16] cycle totals
#ifdef MPQE_TC_STATS
#define STAT_DECL long long st_t0 = 0, st_acc[6] = {0, 0, 0, 0, 0, 0}
#define STAT_BEGIN() st_t0 = clock64()
#define STAT_END(i) st_acc[i] += clock64() - st_t0
#define STAT_FLUSH(base)                                                             \
  if (g_stats != nullptr)                                                            \
    for (int i_ = 0; i_ < 6; ++i_) g_stats[(long long)blockIdx.x * 16 + (base) + i_] = st_acc[i_]
#else
#define STAT_DECL
#define STAT_BEGIN()
#define STAT_END(i)
#define STAT_FLUSH(base)
#endif

using namespace tc;

// Every operand tile is K-major [128 rows][32 k]: element (row, k) at (row/8)*1024 + (k/4)*128 + (row%8)*16 + (k%4)*4
// bytes; descriptor for k-step j (k = 8j..8j+7): start + j*256, LBO = 128 (next 4 k), SBO = 1024 (next 8 rows).
// (Probed on B200 with tools/tc_probe.cu: K-major tf32 operands are exact with and without the 128-byte swizzle and
// run at the same ~160 cycles per 128x128x8 MMA; with either MN-major bit of the instruction descriptor set, every
// canonical MN-major layout tried -- no swizzle and 128-byte swizzle, LBO/SBO in both orders -- accumulates exact
// zeros, even on an all-ones operand: kind::tf32 takes K-major operands only here.)
// Sources whose contiguous dimension is the tile's ROW dimension (weight matrices [k][n], and both operands of the
// weight gradient) are read column-wise (lane <-> row of the tile, 4 scalar loads per 16-byte store), so that every
// shared-memory store is a conflict-free 16-byte st.shared.v4 (a first version transposed with 4-byte stores and
// spent 3x longer in the store pipe than in the tensor pipe).

struct Frag {  // one pipeline stage worth of one operand, per producer thread
  float4 v[4];
};

// rows = queries (K-major source): idx = pw*4+i -> row group idx/2, k half idx%2   (pw = producer warp 0..7)
__device__ __forceinline__ void store_kmajor(const Frag& f, uint32_t hi_tile, uint32_t lo_tile, int pw, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = pw * 4 + i;
    const int off = (idx >> 1) * 1024 + ((idx & 1) * 4 + (lane >> 3)) * 128 + (lane & 7) * 16;
    float4 hi, lo;
    split_tf32(f.v[i], hi, lo);
    sts128(hi_tile + off, hi);     // (explicit shared-window stores: see sts128)
    sts128(lo_tile + off, lo);
  }
}
// Column-gather stage: the source is row-major [32 k][128 mn] (pitch floats between k rows) but the tile wants mn as
// its rows.  Lane <-> mn (a warp instruction reads 128 contiguous bytes of one k row), each thread gathers 4
// consecutive k of its column into a float4, which is exactly one 16-byte row of a core matrix -> one st.shared.v4.
//   producer warp pw: mn = 32*(pw&3) + lane, k-quads kq = 4*(pw>>2) + i, i = 0..3
// `col` = address of this thread's column in k row 0 (src + 32*(pw&3) + lane)
__device__ __forceinline__ void load_columns_at(Frag& f, const float* col, int64_t pitch, int k_valid, int pw) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k0 = 16 * (pw >> 2) + 4 * i;
    f.v[i].x = k0 + 0 < k_valid ? col[(k0 + 0) * pitch] : 0.f;
    f.v[i].y = k0 + 1 < k_valid ? col[(k0 + 1) * pitch] : 0.f;
    f.v[i].z = k0 + 2 < k_valid ? col[(k0 + 2) * pitch] : 0.f;
    f.v[i].w = k0 + 3 < k_valid ? col[(k0 + 3) * pitch] : 0.f;
  }
}
__device__ __forceinline__ void load_columns(Frag& f, const float* src, int64_t pitch, int k_valid, int pw, int lane) {
  load_columns_at(f, src + 32 * (pw & 3) + lane, pitch, k_valid, pw);
}
__device__ __forceinline__ void store_columns(const Frag& f, uint32_t hi_tile, uint32_t lo_tile, int pw, int lane) {
  const int mn = 32 * (pw & 3) + lane;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kq = 4 * (pw >> 2) + i;
    const int off = (mn >> 3) * 1024 + kq * 128 + (mn & 7) * 16;
    float4 hi, lo;
    split_tf32(f.v[i], hi, lo);
    sts128(hi_tile + off, hi);
    sts128(lo_tile + off, lo);
  }
}

// Block-transposing stage (both weight-gradient operands): same source orientation as above, but read with 16-byte
// loads.  Thread (producer warp pw, lane) owns the 4x4 block {k = 4 pw .. 4 pw + 3} x {mn = 4 lane .. 4 lane + 3}:
// four ld.global.v4 (a warp instruction reads one whole 512-byte source row), transposed in registers into the four
// 16-byte K-major rows mn = 4 lane + j.  Lane l stores row j = (i + l/2) mod 4 in its i-th store, so that every
// quarter-warp writes eight different 16-byte bank groups (rows 4 lane + j of consecutive lanes would otherwise hit
// only two): 4 wavefronts per 512-byte store instruction, the minimum.  4x fewer load instructions than the
// scalar column gather.
//   `base` = address of (k row 0, mn = 4 lane)
__device__ __forceinline__ void load_block4(Frag& f, const float* base, int64_t pitch, int k_valid, int pw) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = 4 * pw + r;
    f.v[r] = k < k_valid ? __ldg(reinterpret_cast<const float4*>(base + k * pitch)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ float pick4(const float4& v, int j) {
  const float lo = (j & 1) ? v.y : v.x, hi = (j & 1) ? v.w : v.z;
  return (j & 2) ? hi : lo;
}
__device__ __forceinline__ void store_block4(const Frag& f, uint32_t hi_tile, uint32_t lo_tile, int pw, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + (lane >> 1)) & 3;
    const int mn = 4 * lane + j;
    const int off = (mn >> 3) * 1024 + pw * 128 + (mn & 7) * 16;
    const float4 x = make_float4(pick4(f.v[0], j), pick4(f.v[1], j), pick4(f.v[2], j), pick4(f.v[3], j));
    float4 hi, lo;
    split_tf32(x, hi, lo);
    sts128(hi_tile + off, hi);
    sts128(lo_tile + off, lo);
  }
}

constexpr int EPI_PITCH = 36;  // floats per staged row: 16-byte aligned, rows 4 banks apart

struct TcShared {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint8_t* align_1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

__device__ __forceinline__ void setup(TcShared& sh, int warp, int tid, int full_count) {
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&sh.full[s]), full_count);
      mbar_init(smem_u32(&sh.empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&sh.acc_full[b]), 1);
      mbar_init(smem_u32(&sh.acc_empty[b]), EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&sh.tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// issue the 3xTF32 products of one stage: 4 k-steps x (lo*hi, hi*lo, hi*hi), small products first
__device__ __forceinline__ void issue_stage(uint32_t tmem_d, uint32_t stage_addr, bool first) {
  const uint32_t a_hi = stage_addr, a_lo = stage_addr + TILE_BYTES, b_hi = stage_addr + 2 * TILE_BYTES,
                 b_lo = stage_addr + 3 * TILE_BYTES;
#pragma unroll
  for (int j = 0; j < KC / 8; ++j) {
    const uint64_t dah = make_desc(a_hi + j * 256, 128, 1024), dal = make_desc(a_lo + j * 256, 128, 1024);
    const uint64_t dbh = make_desc(b_hi + j * 256, 128, 1024), dbl = make_desc(b_lo + j * 256, 128, 1024);
    umma_tf32(tmem_d, dal, dbh, IDESC_TF32, (first && j == 0) ? 0u : 1u);
    umma_tf32(tmem_d, dah, dbl, IDESC_TF32, 1u);
    umma_tf32(tmem_d, dah, dbh, IDESC_TF32, 1u);
  }
}

// MMA-issuer role, shared by both kernels: consumes `nsteps` stages for the unit with per-CTA index `uc`
__device__ __forceinline__ void mma_unit(TcShared& sh, uint32_t smem_base, uint32_t tmem, int uc, int nsteps,
                                         uint32_t& it, bool skip_mma = false, long long* stat = nullptr) {
  const int ab = uc & 1, use = uc >> 1;
  long long t0 = 0;
  (void)t0;
#ifdef MPQE_TC_STATS
  t0 = clock64();
#endif
  if (use > 0) mbar_wait(smem_u32(&sh.acc_empty[ab]), (use - 1) & 1);   // epilogue drained this accumulator
#ifdef MPQE_TC_STATS
  if (stat) stat[0] += clock64() - t0;
#endif
  tc_fence_after();
  for (int step = 0; step < nsteps; ++step, ++it) {
    const int s = it % STAGES;
    trace(20, it);
#ifdef MPQE_TC_STATS
    t0 = clock64();
#endif
    mbar_wait(smem_u32(&sh.full[s]), (it / STAGES) & 1);
#ifdef MPQE_TC_STATS
    if (stat) stat[1] += clock64() - t0;
    t0 = clock64();
#endif
    trace(21, it);
    tc_fence_after();
    if (!skip_mma) issue_stage(tmem + ab * 128, smem_base + s * STAGE_BYTES, step == 0);
    umma_commit(smem_u32(&sh.empty[s]));
#ifdef MPQE_TC_STATS
    if (stat) stat[2] += clock64() - t0;
#endif
    trace(22, it);                                // frees the stage when the MMAs have read it
  }
  umma_commit(smem_u32(&sh.acc_full[ab]));                              // accumulator complete
}

// producer-side stage acquisition: stage index + wait until the MMAs that read it last have completed
__device__ __forceinline__ uint32_t acquire_stage(TcShared& sh, uint8_t* smem, uint32_t it) {
  const int s = it % STAGES;
  const uint32_t use = it / STAGES;
  if (use > 0) mbar_wait(smem_u32(&sh.empty[s]), (use - 1) & 1);
  return smem_u32(smem) + s * STAGE_BYTES;
}
__device__ __forceinline__ void publish_stage(TcShared& sh, uint32_t it) {
  fence_proxy_async();                                                  // generic-proxy stores -> async proxy (UMMA)
  mbar_arrive(smem_u32(&sh.full[it % STAGES]));
}

__device__ __forceinline__ int sched_unit(const Schedule& S, int k, int total_units) {
  if (S.count == 0) {
    const int u = blockIdx.x + k * gridDim.x;
    return u < total_units ? u : -1;
  }
  const int i = S.start[blockIdx.x] + k;
  return i < S.start[blockIdx.x + 1] ? (int)S.unit[i] : -1;
}

struct UnitInfo {
  int gi, slot;
  int64_t q0;
};

__device__ __forceinline__ UnitInfo decode_unit(const LayerLaunch& L, int unit) {
  int gi = 0;
  for (; gi < L.num_groups - 1; ++gi) {
    const int tiles = int((L.g[gi].num_queries + BM - 1) / BM);
    const int units = tiles * L.g[gi].num_out_slots;
    if (unit < units) break;
    unit -= units;
  }
  // slot-major numbering inside a group: a persistent CTA's units (blockIdx, +grid, +2*grid, ...) then cycle through
  // slots of different cost instead of always landing on the same (possibly heaviest) slot
  const int tiles = int((L.g[gi].num_queries + BM - 1) / BM);
  UnitInfo u;
  u.gi = gi;
  u.slot = unit / tiles;
  u.q0 = int64_t(unit % tiles) * BM;
  return u;
}

__device__ __forceinline__ uint32_t term_mask(const mpqe_layer_group_t& G, int slot) {
  uint32_t m = 0;
  for (int t = 0; t < G.num_terms; ++t)
    if (G.terms[t].out_slot == slot) m |= 1u << t;
  return m;
}

// k-th unit of this CTA: decoded from the schedule, or (round-robin mode) from the unit number
__device__ __forceinline__ bool sched_next(const LayerLaunch& L, const Schedule& S, int k, int total_units, UnitInfo& U,
                                           uint32_t& mask) {
  if (S.count == 0) {
    const int u = blockIdx.x + k * gridDim.x;
    if (u >= total_units) return false;
    U = decode_unit(L, u);
    mask = term_mask(L.g[U.gi], U.slot);
    return true;
  }
  const int i = S.start[blockIdx.x] + k;
  if (i >= S.start[blockIdx.x + 1]) return false;
  const uint32_t g = S.gsm[i];
  U.gi = (int)(g >> 24);
  U.slot = (int)((g >> 16) & 0xffu);
  U.q0 = (int64_t)S.tile[i] * BM;
  mask = g & 0xffffu;
  return true;
}

__global__ void __launch_bounds__(THREADS, 1) layer_tc_kernel(const __grid_constant__ LayerLaunch L,
                                                              const __grid_constant__ Schedule S, int total_units,
                                                              int dbg) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  __shared__ __align__(16) float epi_stage[EPI_WARPS][32][EPI_PITCH];   // per epilogue warp: a 32x32 accumulator block being transposed
  uint8_t* smem = align_1024(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef MPQE_TC_STATS
  const long long kernel_t0 = clock64();
  unsigned long long gt0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
  setup(sh, warp, tid, PROD_THREADS + 1);   // + the thread that arms / stands in for the bulk copy of the B tiles
  const uint32_t tmem = sh.tmem_base;
#ifdef MPQE_TC_STATS
  if (tid == 0 && g_stats != nullptr) g_stats[(long long)blockIdx.x * 16 + 13] = clock64() - kernel_t0;  // setup
#endif

  if (warp >= EPI_WARPS && warp < MMA_WARP) {
    // ===== producers =====
    // Two register sets: the global loads of stage k+2 are issued right after stage k has been stored, so a full
    // stage of store / fence / handshake work hides their latency.
    const int pw = warp - EPI_WARPS;
    uint32_t it = 0;
    // iterator over (unit, term, k chunk).  Everything a stage needs from the launch descriptor is fetched ONCE per
    // term into registers (four row pointers, the weight pointers): the descriptor lives in the kernel-parameter
    // constant bank and is indexed dynamically, and such loads (LDC with a register index) cost a dependent
    // constant-cache round trip each -- per stage they added up to a third of the stage time.
    int uk = 0;
    uint32_t mask = 0u;
    int kc = D - KC;          // "previous term finished": the first call moves to the first term of the first unit
    bool alive = true;
    UnitInfo U{0, 0, 0};
    const float* rowp[4] = {nullptr, nullptr, nullptr, nullptr};   // this thread's four source rows, k = 0
    const float* packed_base = nullptr;
    const float* plain_base = nullptr;
    // advance(): bookkeeping of the next (term, k chunk) stage -- no loads; false when all of this CTA's work has
    // been issued.  It runs BEFORE the warp waits for the stage it is about to store, so that its descriptor
    // look-ups overlap the latency of the loads already in flight; issue() then fires the stage's loads.
    auto advance = [&](const float*& packed) -> bool {
      if (!alive) return false;
      kc += KC;
      if (kc >= D) {          // next term
        kc = 0;
        mask &= mask - 1;
        while (mask == 0) {   // next unit (a unit without terms contributes no stages)
          if (!sched_next(L, S, uk++, total_units, U, mask)) {
            alive = false;
            return false;
          }
        }
        const mpqe_layer_group_t& G = L.g[U.gi];
        const mpqe_term_t& T = G.terms[__ffs(mask) - 1];
        const float* a = T.a;
        const int64_t a_slots = T.a_slots, a_slot = T.a_slot, nq = G.num_queries;
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // same row / k-quad assignment as store_kmajor
          const int idx = pw * 4 + i;
          int64_t q = U.q0 + (idx >> 1) * 8 + (lane & 7);
          if (q >= nq) q = nq - 1;
          rowp[i] = a + (q * a_slots + a_slot) * (int64_t)D + ((idx & 1) * 4 + (lane >> 3)) * 4;
        }
        packed_base = T.m_packed;   // pre-split weight tiles (mpqe_pack_weights): [k chunk][hi 16 KB | lo 16 KB]
        plain_base = T.m;
      }
      packed = packed_base != nullptr ? packed_base + (kc / KC) * (2 * TILE_BYTES / 4) : nullptr;
      return true;
    };
    auto issue = [&](Frag& fa, Frag& fb, const float* packed) {
      if (!((dbg & 2) && it > 1)) {   // (dbg bit 1: timing experiment without global loads after the first stages)
#pragma unroll
        for (int i = 0; i < 4; ++i) fa.v[i] = *reinterpret_cast<const float4*>(rowp[i] + kc);
        if (packed == nullptr) load_columns(fb, plain_base + (int64_t)kc * D, D, KC, pw, lane);
      }
    };
    STAT_DECL;
    auto put = [&](const Frag& fa, const Frag& fb, const float* packed) {
      STAT_BEGIN();
      const uint32_t st = acquire_stage(sh, smem, it);
      STAT_END(0);   // waiting for a free stage
      STAT_BEGIN();
      if (pw == 0 && lane == 0) {    // the B tiles: one 32 KB bulk copy, or (unpacked weights) a plain arrival
        const uint32_t bar = smem_u32(&sh.full[it % STAGES]);
        if (packed != nullptr) {
          mbar_arrive_expect_tx(bar, 2 * TILE_BYTES);
          bulk_copy_g2s(st + 2 * TILE_BYTES, packed, 2 * TILE_BYTES, bar);
        } else {
          mbar_arrive(bar);
        }
      }
      if (!(dbg & 1)) {              // (dbg bit 0: timing experiment without the shared-memory stores)
        store_kmajor(fa, st, st + TILE_BYTES, pw, lane);
        if (packed == nullptr) store_columns(fb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, pw, lane);
      }
      STAT_END(1);   // waiting for the loaded data + split + stores
    };
    Frag a0, b0, a1, b1;
    const float *p0 = nullptr, *p1 = nullptr, *pn = nullptr;
    bool h0 = advance(p0);
    if (h0) issue(a0, b0, p0);
    bool h1 = h0 && advance(p1);
    if (h1) issue(a1, b1, p1);
    while (h0) {
      STAT_BEGIN();
      const bool n0 = h1 && advance(pn);
      STAT_END(2);   // bookkeeping of the next stage
      put(a0, b0, p0);
      if (n0) {
        p0 = pn;
        issue(a0, b0, p0);
      }
      STAT_BEGIN();
      publish_stage(sh, it);
      STAT_END(3);   // fence + arrive
      ++it;
      if (!h1) break;
      STAT_BEGIN();
      const bool n1 = n0 && advance(pn);
      STAT_END(2);
      put(a1, b1, p1);
      if (n1) {
        p1 = pn;
        issue(a1, b1, p1);
      }
      STAT_BEGIN();
      publish_stage(sh, it);
      STAT_END(3);
      ++it;
      h0 = n0;
      h1 = n1;
    }
    if (pw == 0 && lane == 0) { STAT_FLUSH(0); }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;
      int uc = 0;
      long long mstat[3] = {0, 0, 0};
      UnitInfo U{0, 0, 0};
      for (uint32_t umask; sched_next(L, S, uc, total_units, U, umask); ++uc) {
        const int nsteps = __popc(umask) * (D / KC);
        mma_unit(sh, smem_u32(smem), tmem, uc, nsteps, it, (dbg & 4) != 0, mstat);
      }
#ifdef MPQE_TC_STATS
      if (g_stats != nullptr) {
        for (int i = 0; i < 3; ++i) g_stats[(long long)blockIdx.x * 16 + 6 + i] = mstat[i];
        g_stats[(long long)blockIdx.x * 16 + 9] = it;
        g_stats[(long long)blockIdx.x * 16 + 10] = uc;
      }
#endif
    }
  } else {
    // ===== epilogue: one accumulator row per thread =====
    int uc = 0;
    long long estat[4] = {0, 0, 0, 0}, et0 = 0;
    (void)estat;
    (void)et0;
    UnitInfo U{0, 0, 0};
    for (uint32_t umask; sched_next(L, S, uc, total_units, U, umask); ++uc) {
      const mpqe_layer_group_t& G = L.g[U.gi];
      const int nsteps = __popc(umask) * (D / KC);
      const int ab = uc & 1;
      if (tid == 0) trace(30, uc);
      // Everything the store loop needs is pulled out of the launch descriptor into registers here, once per unit
      // (dynamically indexed kernel-parameter loads inside the loop cost a dependent constant-cache round trip per
      // store).  The ReLU mask and the bias of a 32-column block are fetched one block ahead through the read-only
      // path, all eight rows at once: loaded inside the store loop they cannot be hoisted above the stores
      // (possible aliasing) and every row pays a full memory round trip.
      const int oslot = G.out_slot_map[U.slot];
      const int cq = (lane & 7) * 4;          // this lane's 4 columns inside the 32-column block
      const int epi = G.epilogue;
      const bool masked = epi == MPQE_EPI_MASK;
      const int row0 = warp * 32 + (lane >> 3);                  // this lane's rows: row0 + 4 i
      const int64_t rows_left = G.num_queries - U.q0 - row0;     // row 4 i exists iff 4 i < rows_left
      float* outp = G.out + ((U.q0 + row0) * (int64_t)G.out_slots + oslot) * (int64_t)D + cq;
      const int64_t out_step = 4 * (int64_t)G.out_slots * D;
      const float* maskp = masked ? G.mask + ((U.q0 + row0) * (int64_t)G.mask_slots + oslot) * (int64_t)D + cq : nullptr;
      const int64_t mask_step = 4 * (int64_t)G.mask_slots * D;
      const float* biasp = G.bias != nullptr ? G.bias + (int64_t)U.slot * G.bias_slot_stride + cq : nullptr;
      const float bscale = biasp != nullptr ? G.bias_scale[U.slot] : 0.f;
      const bool store = !(dbg & 8);
      float4 mk[8], bnext = make_float4(0.f, 0.f, 0.f, 0.f);
      auto fetch = [&](int c0) {
        if (masked) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            mk[i] = 4 * i < rows_left ? __ldg(reinterpret_cast<const float4*>(maskp + i * mask_step + c0))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (biasp != nullptr) bnext = __ldg(reinterpret_cast<const float4*>(biasp + c0));
      };
      fetch(0);
#ifdef MPQE_TC_STATS
      et0 = clock64();
#endif
      mbar_wait(smem_u32(&sh.acc_full[ab]), (uc >> 1) & 1);
#ifdef MPQE_TC_STATS
      estat[0] += clock64() - et0;
      et0 = clock64();
#endif
      if (tid == 0) trace(31, uc);
      tc_fence_after();
      // TMEM gives each thread one accumulator ROW; a 32x32 block per warp is transposed through shared memory so
      // that every global store instruction writes 4 full 128-byte row segments (instead of 32 scattered 16-byte
      // pieces, which kept the load/store pipe busier than the tensor pipe).
      float* stage = &epi_stage[warp][0][0];
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        // the fetched block: bias scaled, mask compressed to one bit per element; then the next block's loads go out
        const float4 bv = make_float4(bscale * bnext.x, bscale * bnext.y, bscale * bnext.z, bscale * bnext.w);
        uint32_t mbits = 0;
        if (masked) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            mbits |= ((mk[i].x > 0.f ? 1u : 0u) | (mk[i].y > 0.f ? 2u : 0u) | (mk[i].z > 0.f ? 4u : 0u) |
                      (mk[i].w > 0.f ? 8u : 0u)) << (4 * i);
        }
        if (c0 + 32 < D) fetch(c0 + 32);
        uint32_t v[32];
#ifdef MPQE_TC_STATS
        const long long tl0 = clock64();
#endif
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ab * 128 + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(stage + lane * EPI_PITCH + i) =
              make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                          __uint_as_float(v[i + 3]));
        __syncwarp();
#ifdef MPQE_TC_STATS
        estat[2] += clock64() - tl0;
        const long long tl1 = clock64();
#endif
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = *reinterpret_cast<const float4*>(stage + (4 * i + (lane >> 3)) * EPI_PITCH + cq);
          if (nsteps == 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
          o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          if (epi == MPQE_EPI_RELU) {
            o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
          } else if (masked) {
            const uint32_t mb = mbits >> (4 * i);
            o = make_float4((mb & 1u) ? o.x : 0.f, (mb & 2u) ? o.y : 0.f, (mb & 4u) ? o.z : 0.f, (mb & 8u) ? o.w : 0.f);
          }
          if (4 * i < rows_left && store) *reinterpret_cast<float4*>(outp + i * out_step + c0) = o;
        }
        __syncwarp();
#ifdef MPQE_TC_STATS
        estat[3] += clock64() - tl1;
#endif
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.acc_empty[ab]));
#ifdef MPQE_TC_STATS
      estat[1] += clock64() - et0;
#endif
      if (tid == 0) trace(32, uc);
    }
#ifdef MPQE_TC_STATS
    if (tid == 0 && g_stats != nullptr) {
      g_stats[(long long)blockIdx.x * 16 + 11] = estat[0];
      g_stats[(long long)blockIdx.x * 16 + 12] = estat[1];
      g_stats[(long long)blockIdx.x * 16 + 5] = estat[2];
      g_stats[(long long)blockIdx.x * 16 + 13] = estat[3];
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
#ifdef MPQE_TC_STATS
  if (tid == 0 && g_stats != nullptr) {
    g_stats[(long long)blockIdx.x * 16 + 14] = clock64() - kernel_t0;  // whole CTA
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    g_stats[(long long)blockIdx.x * 16 + 15] = (long long)gt0;          // absolute start (ns)
    g_stats[(long long)blockIdx.x * 16 + 4] = (long long)(gt1 - gt0);   // CTA lifetime (ns)
  }
#endif
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient on tensor cores: dM = sum_q A[q]^T G[q]; both operands are transposed into K-major tiles.
// unit = (destination, chunk); its stages are the 32-query tiles of every matching (group, term).
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

__device__ __forceinline__ void decode_wgrad_unit(const WgradLaunch& L, int unit, int& j, int& c) {
  j = 0;
  for (; j < L.num_dests - 1; ++j) {
    if (unit < L.chunks[j]) break;
    unit -= L.chunks[j];
  }
  c = unit;
}

__device__ __forceinline__ int wgrad_unit_steps(const WgradLaunch& L, int j, int c) {
  int steps = 0;
  for (int g = 0; g < L.num_groups; ++g) {
    int64_t qb, qe;
    chunk_range(L.g[g].num_queries, L.chunks[j], c, qb, qe);
    const int tiles = (int)((qe - qb + KC - 1) / KC);
    for (int t = 0; t < L.g[g].num_terms; ++t)
      if (L.g[g].terms[t].m == L.d[j].m_fwd) steps += tiles;
  }
  return steps;
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradLaunch L,
                                                              const __grid_constant__ Schedule S, int total_units) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  __shared__ __align__(16) float epi_stage[EPI_WARPS][32][EPI_PITCH];
  uint8_t* smem = align_1024(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  setup(sh, warp, tid, PROD_THREADS);
  const uint32_t tmem = sh.tmem_base;

  if (warp >= EPI_WARPS && warp < MMA_WARP) {
    const int pw = warp - EPI_WARPS;
    uint32_t it = 0;
    // iterator over (unit, matching (group, term), 32-query tile), crossing units; two register sets as above
    int uk = 0, j = 0, c = 0;
    WgradIter wi{0, 0, 0, 0};
    bool in_unit = false, alive = true;
    // per-term state in registers (see layer_tc_kernel: no descriptor loads on the per-stage path)
    const float *a_col = nullptr, *g_col = nullptr;
    int64_t a_pitch = 0, g_pitch = 0;
    auto enter_term = [&]() {
      const mpqe_layer_group_t& G = L.g[wi.g];
      const mpqe_term_t& T = G.terms[wi.t];
      const mpqe_wgrad_operand_t& O = L.go[wi.g];
      const int col = 4 * lane;          // this thread's 4x4 block: columns 4 lane .. 4 lane + 3 (load_block4)
      a_pitch = (int64_t)T.a_slots * D;
      a_col = T.a + (int64_t)T.a_slot * D + col;
      g_pitch = (int64_t)O.g_slots * D;
      g_col = O.g + (int64_t)O.slot_map[T.out_slot] * D + col;
    };
    auto load_next = [&](Frag& fa, Frag& fb) -> bool {
      if (!alive) return false;
      while (!in_unit) {
        const int unit = sched_unit(S, uk++, total_units);
        if (unit < 0) {
          alive = false;
          return false;
        }
        decode_wgrad_unit(L, unit, j, c);
        wi = WgradIter{0, 0, 0, 0};
        in_unit = wgrad_seek(L, L.d[j].m_fwd, L.chunks[j], c, wi);
        if (in_unit) enter_term();
      }
      const int valid = (int)(wi.qe - wi.q < KC ? wi.qe - wi.q : KC);
      load_block4(fa, a_col + wi.q * a_pitch, a_pitch, valid, pw);
      load_block4(fb, g_col + wi.q * g_pitch, g_pitch, valid, pw);
      wi.q += KC;
      if (wi.q >= wi.qe) {
        ++wi.t;
        in_unit = wgrad_seek(L, L.d[j].m_fwd, L.chunks[j], c, wi);
        if (in_unit) enter_term();
      }
      return true;
    };
    auto put = [&](const Frag& fa, const Frag& fb) {
      const uint32_t st = acquire_stage(sh, smem, it);
      store_block4(fa, st, st + TILE_BYTES, pw, lane);
      store_block4(fb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, pw, lane);
    };
    Frag a0, b0, a1, b1;
    bool h0 = load_next(a0, b0);
    bool h1 = h0 && load_next(a1, b1);
    while (h0) {
      put(a0, b0);
      const bool n0 = h1 && load_next(a0, b0);
      publish_stage(sh, it);
      ++it;
      if (!h1) break;
      put(a1, b1);
      const bool n1 = n0 && load_next(a1, b1);
      publish_stage(sh, it);
      ++it;
      h0 = n0;
      h1 = n1;
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      uint32_t it = 0;
      int uc = 0;
      for (int unit; (unit = sched_unit(S, uc, total_units)) >= 0; ++uc) {
        int j, c;
        decode_wgrad_unit(L, unit, j, c);
        mma_unit(sh, smem_u32(smem), tmem, uc, wgrad_unit_steps(L, j, c), it);
      }
    }
  } else {
    int uc = 0;
    for (int unit; (unit = sched_unit(S, uc, total_units)) >= 0; ++uc) {
      int j, c;
      decode_wgrad_unit(L, unit, j, c);
      const int nsteps = wgrad_unit_steps(L, j, c);
      const int ab = uc & 1;
      mbar_wait(smem_u32(&sh.acc_full[ab]), (uc >> 1) & 1);
      tc_fence_after();
      float* P = L.partials + (int64_t)unit * D * D;   // units are numbered destination-major, chunk-minor
      float* stage = &epi_stage[warp][0][0];
      const int cq = (lane & 7) * 4;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ab * 128 + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(stage + lane * EPI_PITCH + i) =
              make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                          __uint_as_float(v[i + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3);
          float4 o = *reinterpret_cast<const float4*>(stage + rr * EPI_PITCH + cq);
          if (nsteps == 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(P + (warp * 32 + rr) * D + c0 + cq) = o;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.acc_empty[ab]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// The same weight gradient with a TMA loader (default; MPQE_WGRAD_LOADS=ldg selects the kernel above).
// The kernel above has ONE stage of loads in flight per producer thread (two register sets; a third does not fit
// into 128 registers) and issues 2048 16-byte requests per stage: its stage period was 2-3x the ~1700 cycles the
// twelve N=128 MMAs of a stage need.  Here one thread issues two tensor loads per stage (box {128 features, 1 slot,
// 32 queries} of each operand) into a two-slot raw ring, two stages ahead; the producers read the raw tiles back
// (a warp reads one 512-byte row per instruction), transpose 4x4 in registers, split and store as before.
// Shared memory: 2 operand stages x 64 KB + 2 raw slots x 32 KB.
// ------------------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 2;
constexpr int WG_RING = 2;
constexpr int WG_RAW = 2 * TILE_BYTES;                 // A rows | G rows, [32 queries][128 features] each
constexpr int WG_LOAD_WARP = MMA_WARP + 1;
constexpr int WG_THREADS = (WG_LOAD_WARP + 1) * 32;    // 448
constexpr size_t WG_SMEM = size_t(WG_STAGES) * STAGE_BYTES + size_t(WG_RING) * WG_RAW + 1024;
constexpr int WG_MAX_MAPS = 24;
struct alignas(64) WgradMaps {
  CUtensorMap map[WG_MAX_MAPS];
  uint8_t a_of[MPQE_MAX_GROUPS][MPQE_MAX_TERMS];   // map of term t's forward operand
  uint8_t g_of[MPQE_MAX_GROUPS];                   // map of the group's gradient operand
};
struct WgShared {
  uint64_t full[WG_STAGES];
  uint64_t empty[WG_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t raw_full[WG_RING];
  uint64_t raw_empty[WG_RING];
  uint32_t tmem_base;
};

// iterator over the 32-query tiles of a CTA's units: (unit, matching (group, term), tile), crossing units
struct WgWalk {
  int uk = 0, j = 0, c = 0;
  WgradIter wi{0, 0, 0, 0};
  bool in_unit = false, alive = true;
  // positions on the next tile; returns false when the CTA has no more work.  (g, t, q, valid) describe the tile.
  __device__ __forceinline__ bool next(const WgradLaunch& L, const Schedule& S, int total_units, int& g, int& t,
                                       int64_t& q, int& valid) {
    if (!alive) return false;
    while (!in_unit) {
      const int unit = sched_unit(S, uk++, total_units);
      if (unit < 0) {
        alive = false;
        return false;
      }
      decode_wgrad_unit(L, unit, j, c);
      wi = WgradIter{0, 0, 0, 0};
      in_unit = wgrad_seek(L, L.d[j].m_fwd, L.chunks[j], c, wi);
    }
    g = wi.g;
    t = wi.t;
    q = wi.q;
    valid = (int)(wi.qe - wi.q < KC ? wi.qe - wi.q : KC);
    wi.q += KC;
    if (wi.q >= wi.qe) {
      ++wi.t;
      in_unit = wgrad_seek(L, L.d[j].m_fwd, L.chunks[j], c, wi);
    }
    return true;
  }
};

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc2_kernel(const __grid_constant__ WgradLaunch L,
                                                                  const __grid_constant__ Schedule S, int total_units,
                                                                  const __grid_constant__ WgradMaps WM) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ WgShared sh;
  __shared__ __align__(16) float epi_stage[EPI_WARPS][32][EPI_PITCH];
  uint8_t* smem = align_1024(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(smem_u32(&sh.full[s]), PROD_THREADS);
      mbar_init(smem_u32(&sh.empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&sh.acc_full[b]), 1);
      mbar_init(smem_u32(&sh.acc_empty[b]), EPI_WARPS * 32);
    }
    for (int r = 0; r < WG_RING; ++r) {
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
  const uint32_t base = smem_u32(smem);
  const uint32_t ring0 = base + WG_STAGES * STAGE_BYTES;

  if (warp >= EPI_WARPS && warp < MMA_WARP) {
    // ===== producers: raw rows (ring) -> 4x4 register transpose -> hi / lo K-major tiles ============================
    const int pw = warp - EPI_WARPS;
    WgWalk walk;
    int g, t, valid;
    int64_t q;
#pragma unroll 1
    for (uint32_t it = 0; walk.next(L, S, total_units, g, t, q, valid); ++it) {
      const int slot = it % WG_RING, s = it % WG_STAGES;
      mbar_wait(smem_u32(&sh.raw_full[slot]), (it / WG_RING) & 1);
      if (it >= WG_STAGES) mbar_wait(smem_u32(&sh.empty[s]), (it / WG_STAGES - 1) & 1);
      const uint32_t raw = ring0 + slot * WG_RAW + lane * 16;
      const uint32_t st = base + s * STAGE_BYTES;
      Frag fa, fb;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = 4 * pw + r;                   // query row of the tile; rows >= valid belong to another chunk
        fa.v[r] = fb.v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < valid) {
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(fa.v[r].x), "=f"(fa.v[r].y), "=f"(fa.v[r].z), "=f"(fa.v[r].w)
                       : "r"(raw + k * 512));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(fb.v[r].x), "=f"(fb.v[r].y), "=f"(fb.v[r].z), "=f"(fb.v[r].w)
                       : "r"(raw + TILE_BYTES + k * 512));
        }
      }
      store_block4(fa, st, st + TILE_BYTES, pw, lane);
      store_block4(fb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, pw, lane);
      fence_proxy_async();
      mbar_arrive(smem_u32(&sh.full[s]));
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sh.raw_empty[slot]));
    }
  } else if (warp == WG_LOAD_WARP) {
    // ===== loader (one thread): two tensor loads per stage ===========================================================
    if (lane == 0) {
      WgWalk walk;
      int g, t, valid;
      int64_t q;
      for (uint32_t it = 0; walk.next(L, S, total_units, g, t, q, valid); ++it) {
        const int slot = it % WG_RING;
        if (it >= WG_RING) mbar_wait(smem_u32(&sh.raw_empty[slot]), (it / WG_RING - 1) & 1);
        const uint32_t bar = smem_u32(&sh.raw_full[slot]);
        const mpqe_term_t& T = L.g[g].terms[t];
        mbar_arrive_expect_tx(bar, WG_RAW);
        tma_load_3d(ring0 + slot * WG_RAW, &WM.map[WM.a_of[g][t]], 0, T.a_slot, (int)q, bar);
        tma_load_3d(ring0 + slot * WG_RAW + TILE_BYTES, &WM.map[WM.g_of[g]], 0, L.go[g].slot_map[T.out_slot], (int)q, bar);
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      uint32_t it = 0;
      int uc = 0;
      for (int unit; (unit = sched_unit(S, uc, total_units)) >= 0; ++uc) {
        int j, c;
        decode_wgrad_unit(L, unit, j, c);
        const int nsteps = wgrad_unit_steps(L, j, c);
        const int ab = uc & 1, use = uc >> 1;
        if (use > 0) mbar_wait(smem_u32(&sh.acc_empty[ab]), (use - 1) & 1);
        tc_fence_after();
        for (int step = 0; step < nsteps; ++step, ++it) {
          const int s = it % WG_STAGES;
          mbar_wait(smem_u32(&sh.full[s]), (it / WG_STAGES) & 1);
          tc_fence_after();
          issue_stage(tmem + ab * 128, base + s * STAGE_BYTES, step == 0);
          umma_commit(smem_u32(&sh.empty[s]));
        }
        umma_commit(smem_u32(&sh.acc_full[ab]));
      }
    }
  } else if (warp < EPI_WARPS) {
    int uc = 0;
    for (int unit; (unit = sched_unit(S, uc, total_units)) >= 0; ++uc) {
      int j, c;
      decode_wgrad_unit(L, unit, j, c);
      const int nsteps = wgrad_unit_steps(L, j, c);
      const int ab = uc & 1;
      mbar_wait(smem_u32(&sh.acc_full[ab]), (uc >> 1) & 1);
      tc_fence_after();
      float* P = L.partials + (int64_t)unit * D * D;   // units are numbered destination-major, chunk-minor
      float* stage = &epi_stage[warp][0][0];
      const int cq = (lane & 7) * 4;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ab * 128 + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(stage + lane * EPI_PITCH + i) =
              make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                          __uint_as_float(v[i + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3);
          float4 o = *reinterpret_cast<const float4*>(stage + rr * EPI_PITCH + cq);
          if (nsteps == 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(P + (warp * 32 + rr) * D + c0 + cq) = o;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.acc_empty[ab]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// Pre-split weights for the layer kernel: out[m][k chunk] = [hi tile 16 KB | lo tile 16 KB] of B[n][k] = M[k][n],
// i.e. byte-exact images of the shared-memory operand tiles, so that staging them is one bulk copy.
constexpr int PACK_MAX = 256;
struct PackLaunch {
  const float* m[PACK_MAX];
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackLaunch P, float* __restrict__ out) {
  const float* M = P.m[blockIdx.y];
  const int kcb = blockIdx.x;                       // k chunk (32 k)
  float* hi_tile = out + ((int64_t)blockIdx.y * (D / KC) + kcb) * (2 * TILE_BYTES / 4);
  float* lo_tile = hi_tile + TILE_BYTES / 4;
  for (int e = threadIdx.x; e < 128 * 8; e += 256) {  // e -> (n, k quad): lanes run over n (coalesced reads of M rows)
    const int n = e & 127, kq = e >> 7;
    float4 x;
    x.x = M[(int64_t)(kcb * KC + kq * 4 + 0) * D + n];
    x.y = M[(int64_t)(kcb * KC + kq * 4 + 1) * D + n];
    x.z = M[(int64_t)(kcb * KC + kq * 4 + 2) * D + n];
    x.w = M[(int64_t)(kcb * KC + kq * 4 + 3) * D + n];
    float4 hi, lo;
    split_tf32(x, hi, lo);
    const int off = ((n >> 3) * 1024 + kq * 128 + (n & 7) * 16) / 4;
    *reinterpret_cast<float4*>(hi_tile + off) = hi;
    *reinterpret_cast<float4*>(lo_tile + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Full-entity rank counts on tensor cores (mpqe_rank_counts_table, use_tensor_cores = 1).
// D[candidate, query] = table_row . q  for a 128-candidate x 128-query tile, K = 128; both operands are K-major as
// stored.  A CTA owns one (query tile, slice of the candidate range): the query tile is split into tf32 hi/lo ONCE
// and stays resident in shared memory (128 KB); the candidate rows are pre-split once per call into tile images
// (pack_rows_kernel) and stream through 3 x 32 KB stages as bulk copies issued by one thread -- no per-tile
// conversion work (re-splitting every candidate tile for each of the B/128 query tiles kept the tensor pipe at
// ~40 %).  The epilogue scales
// by 1/||row|| and 1/max(||q||,eps), compares with the positive score and counts with warp ballots; only integer
// atomics touch global memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int RANK_STAGES = 3;
constexpr int RANK_A_STAGE = 2 * TILE_BYTES;                       // A_hi | A_lo
constexpr size_t RANK_SMEM = size_t(D / KC) * 2 * TILE_BYTES + size_t(RANK_STAGES) * RANK_A_STAGE + 1024;

struct RankLaunch {
  const float* table;      // rows [row_begin, row_begin + rows)
  const float* packed;     // [candidate tile][k chunk][hi 16 KB | lo 16 KB] images of those rows
  int64_t row_begin, rows;
  const float* inv_norm;   // [rows]
  const float* q;          // [B, D]
  const float* qinv;       // [B]
  const float* pos;        // [B]
  int64_t B;
  unsigned long long* left;
  unsigned long long* right;
  int splits;              // candidate-range slices per query tile
};

// rows [rows][D] (row-major, K contiguous) -> per 128-row tile and 32-k chunk the [hi | lo] K-major tile images.
// Thread e of a (tile, chunk) block writes the e-th 16-byte unit of the image (coalesced stores); rows past the end
// are zero (the epilogue masks them).
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ rows_base, int64_t rows,
                                                        float* __restrict__ out) {
  const int kcb = blockIdx.x;
  const int64_t tile = blockIdx.y;
  float* hi_tile = out + (tile * (D / KC) + kcb) * (2 * TILE_BYTES / 4);
  float* lo_tile = hi_tile + TILE_BYTES / 4;
  for (int e = threadIdx.x; e < 128 * 8; e += 256) {
    const int r = (e >> 6) * 8 + (e & 7), kq = (e >> 3) & 7;
    const int64_t row = tile * BM + r;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) x = *reinterpret_cast<const float4*>(rows_base + row * D + kcb * KC + kq * 4);
    float4 hi, lo;
    split_tf32(x, hi, lo);
    *reinterpret_cast<float4*>(hi_tile + e * 4) = hi;   // e == ((r/8)*1024 + kq*128 + (r%8)*16) / 16
    *reinterpret_cast<float4*>(lo_tile + e * 4) = lo;
  }
}

__device__ __forceinline__ void load_rows_kmajor(Frag& f, const float* base, int64_t first_row, int64_t num_rows,
                                                 int kc, int pw, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = pw * 4 + i;
    int64_t r = first_row + (idx >> 1) * 8 + (lane & 7);
    if (r >= num_rows) r = num_rows - 1;
    const int kq = (idx & 1) * 4 + (lane >> 3);
    f.v[i] = *reinterpret_cast<const float4*>(base + r * D + kc + kq * 4);
  }
}

__global__ void __launch_bounds__(THREADS, 1) rank_tc_kernel(const __grid_constant__ RankLaunch R) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  __shared__ uint64_t bres_full;
  __shared__ float s_qinv[BM], s_pos[BM];
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* bres = smem;                                             // [k chunk][hi | lo]
  uint8_t* astage = smem + (D / KC) * 2 * TILE_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) mbar_init(smem_u32(&bres_full), PROD_THREADS);
  setup(sh, warp, tid, 1);                                          // a stage is filled by one bulk copy
  const uint32_t tmem = sh.tmem_base;

  const int qt = blockIdx.x / R.splits, split = blockIdx.x % R.splits;
  const int64_t b0 = (int64_t)qt * BM;
  const int64_t tiles = (R.rows + BM - 1) / BM;
  const int64_t per = (tiles + R.splits - 1) / R.splits;
  const int64_t t0 = per * split, t1 = t0 + per < tiles ? t0 + per : tiles;   // candidate tiles of this CTA
  const int num_units = t1 > t0 ? (int)(t1 - t0) : 0;

  if (warp >= EPI_WARPS && warp < MMA_WARP) {
    const int pw = warp - EPI_WARPS;
    Frag f;
    for (int kcb = 0; kcb < D / KC; ++kcb) {                        // resident query tile
      load_rows_kmajor(f, R.q, b0, R.B, kcb * KC, pw, lane);
      store_kmajor(f, smem_u32(bres) + kcb * 2 * TILE_BYTES, smem_u32(bres) + kcb * 2 * TILE_BYTES + TILE_BYTES, pw, lane);
    }
    fence_proxy_async();
    mbar_arrive(smem_u32(&bres_full));
    if (pw == 0 && lane == 0) {                                     // candidate tiles: one 32 KB bulk copy per stage
      const int total = num_units * (D / KC);
      const float* src = R.packed + t0 * (int64_t)(D / KC) * (2 * TILE_BYTES / 4);
      for (int it = 0; it < total; ++it) {
        const int s = it % RANK_STAGES;
        const uint32_t use = it / RANK_STAGES;
        if (use > 0) mbar_wait(smem_u32(&sh.empty[s]), (use - 1) & 1);
        const uint32_t bar = smem_u32(&sh.full[s]);
        mbar_arrive_expect_tx(bar, 2 * TILE_BYTES);
        bulk_copy_g2s(smem_u32(astage + s * RANK_A_STAGE), src + (int64_t)it * (2 * TILE_BYTES / 4), 2 * TILE_BYTES, bar);
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      mbar_wait(smem_u32(&bres_full), 0);
      tc_fence_after();
      uint32_t it = 0;
      const uint32_t bres_a = smem_u32(bres), ast_a = smem_u32(astage);
      for (int uc = 0; uc < num_units; ++uc) {
        const int ab = uc & 1, use = uc >> 1;
        if (use > 0) mbar_wait(smem_u32(&sh.acc_empty[ab]), (use - 1) & 1);
        tc_fence_after();
        for (int kcb = 0; kcb < D / KC; ++kcb, ++it) {
          const int s = it % RANK_STAGES;
          mbar_wait(smem_u32(&sh.full[s]), (it / RANK_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_hi = ast_a + s * RANK_A_STAGE, a_lo = a_hi + TILE_BYTES;
          const uint32_t b_hi = bres_a + kcb * 2 * TILE_BYTES, b_lo = b_hi + TILE_BYTES;
#pragma unroll
          for (int j = 0; j < KC / 8; ++j) {
            const uint64_t dah = make_desc(a_hi + j * 256, 128, 1024), dal = make_desc(a_lo + j * 256, 128, 1024);
            const uint64_t dbh = make_desc(b_hi + j * 256, 128, 1024), dbl = make_desc(b_lo + j * 256, 128, 1024);
            umma_tf32(tmem + ab * 128, dal, dbh, IDESC_TF32, (kcb == 0 && j == 0) ? 0u : 1u);
            umma_tf32(tmem + ab * 128, dah, dbl, IDESC_TF32, 1u);
            umma_tf32(tmem + ab * 128, dah, dbh, IDESC_TF32, 1u);
          }
          umma_commit(smem_u32(&sh.empty[s]));
        }
        umma_commit(smem_u32(&sh.acc_full[ab]));
      }
    }
  } else {
    // epilogue: thread = candidate row, columns = the 128 queries of the tile
    for (int c = tid; c < BM; c += EPI_WARPS * 32) {
      const int64_t b = b0 + c;
      s_qinv[c] = b < R.B ? R.qinv[b] : 0.f;
      s_pos[c] = b < R.B ? R.pos[b] : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32));          // only the epilogue warps
    unsigned lt[4] = {0, 0, 0, 0}, le[4] = {0, 0, 0, 0};             // lane c holds the counts of columns c0 + c
    for (int uc = 0; uc < num_units; ++uc) {
      const int ab = uc & 1;
      mbar_wait(smem_u32(&sh.acc_full[ab]), (uc >> 1) & 1);
      tc_fence_after();
      const int64_t row = (t0 + uc) * BM + warp * 32 + lane;
      const bool valid = row < R.rows;
      const float inr = valid ? R.inv_norm[row] : 0.f;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ab * 128 + cb * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sc = __uint_as_float(v[i]) * inr * s_qinv[cb * 32 + i];
          const float p = s_pos[cb * 32 + i];
          const unsigned m_lt = __ballot_sync(0xffffffffu, valid && sc < p);
          const unsigned m_le = __ballot_sync(0xffffffffu, valid && sc <= p);
          if (lane == i) {
            lt[cb] += __popc(m_lt);
            le[cb] += __popc(m_le);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.acc_empty[ab]));
    }
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      const int64_t b = b0 + cb * 32 + lane;
      if (b < R.B) {
        if (lt[cb]) atomicAdd(R.left + b, (unsigned long long)lt[cb]);
        if (le[cb]) atomicAdd(R.right + b, (unsigned long long)le[cb]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

int num_sms() {
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
  const int grid = units < num_sms() ? (int)units : num_sms();
  // longest-processing-time-first assignment of units to the persistent CTAs (cost = terms of the unit's slot)
  static thread_local Schedule S;
  S.count = 0;
  if (units > grid && units <= SCHED_MAX_UNITS && grid <= SCHED_MAX_CTAS) {
    static thread_local int cost[SCHED_MAX_UNITS];
    static thread_local uint16_t tile_of[SCHED_MAX_UNITS];
    static thread_local uint32_t gsm_of[SCHED_MAX_UNITS];
    int u = 0;
    for (int i = 0; i < num_groups; ++i) {
      const int tiles = (int)((groups[i].num_queries + BM - 1) / BM);
      for (int slot = 0; slot < groups[i].num_out_slots; ++slot) {
        int nt = 0;
        uint32_t mask = 0;
        for (int t = 0; t < groups[i].num_terms; ++t)
          if (groups[i].terms[t].out_slot == slot) ++nt, mask |= 1u << t;
        for (int k = 0; k < tiles; ++k) {     // slot-major numbering, as decode_unit
          cost[u] = nt;
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
  }
  // MPQE_TC_DEBUG: timing experiments of debug builds only (-DMPQE_TC_STATS; bit0 no smem stores, bit1 no global loads,
  // bit2 no MMAs, bit3 no epilogue stores: results are wrong when non-zero).  Release builds ignore the variable.
#ifdef MPQE_TC_STATS
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("MPQE_TC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
#else
  const int dbg = 0;
#endif
  layer_tc_kernel<<<grid, THREADS, TC_SMEM, stream>>>(L, S, (int)units, dbg);
  MPQE_CHECK_LAUNCH("layer_tc_kernel");
  return 0;
}

// Tensor maps of the operands of a weight-gradient launch.  false: use the register-prefetch kernel (broadcast
// operands, more distinct operands than WG_MAX_MAPS, no cuTensorMapEncodeTiled, or MPQE_WGRAD_LOADS=ldg).
static bool build_wgrad_maps(const WgradLaunch& launch, WgradMaps& WM) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("MPQE_WGRAD_LOADS");
    mode = (e != nullptr && strcmp(e, "ldg") == 0) || tensor_map_encoder() == nullptr ? 0 : 1;
  }
  if (mode == 0) return false;
  struct Key {
    const float* p;
    int32_t slots;
    int64_t nq;
  };
  Key keys[WG_MAX_MAPS];
  int n = 0;
  auto find = [&](const float* p, int32_t slots, int64_t nq) -> int {
    if (slots <= 0) return -1;
    for (int k = 0; k < n; ++k)
      if (keys[k].p == p && keys[k].slots == slots && keys[k].nq == nq) return k;
    if (n == WG_MAX_MAPS) return -1;
    if (!encode_rows_map(&WM.map[n], p, slots, nq, D, KC, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    keys[n] = Key{p, slots, nq};
    return n++;
  };
  for (int g = 0; g < launch.num_groups; ++g) {
    const mpqe_layer_group_t& G = launch.g[g];
    const int kg = find(launch.go[g].g, launch.go[g].g_slots, G.num_queries);
    if (kg < 0) return false;
    WM.g_of[g] = (uint8_t)kg;
    for (int t = 0; t < G.num_terms; ++t) {
      const int ka = find(G.terms[t].a, G.terms[t].a_slots, G.num_queries);
      if (ka < 0) return false;
      WM.a_of[g][t] = (uint8_t)ka;
    }
  }
  return true;
}

// `launch` arrives fully prepared (groups, operands, dests, chunks, partials) from the host code in layer_simt.cu
int layer_wgrad_tc_launch(const WgradLaunch& launch, int total_chunks, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    MPQE_CUDA(cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    configured = true;
  }
  const int grid = total_chunks < num_sms() ? total_chunks : num_sms();
  static thread_local Schedule S;
  S.count = 0;
  if (total_chunks > grid && total_chunks <= SCHED_MAX_UNITS && grid <= SCHED_MAX_CTAS) {
    // stages of every (destination, chunk) unit, as wgrad_unit_steps computes them on the device
    static thread_local int cost[SCHED_MAX_UNITS];
    int u = 0;
    for (int j = 0; j < launch.num_dests; ++j)
      for (int c = 0; c < launch.chunks[j]; ++c) {
        int steps = 0;
        for (int g = 0; g < launch.num_groups; ++g) {
          const int64_t B = launch.g[g].num_queries;
          int64_t per = (B + launch.chunks[j] - 1) / launch.chunks[j];
          per = (per + KC - 1) / KC * KC;
          int64_t qb = per * c, qe = qb + per;
          if (qb > B) qb = B;
          if (qe > B) qe = B;
          const int tiles = (int)((qe - qb + KC - 1) / KC);
          for (int t = 0; t < launch.g[g].num_terms; ++t)
            if (launch.g[g].terms[t].m == launch.d[j].m_fwd) steps += tiles;
        }
        cost[u++] = steps;
      }
    build_lpt(S, cost, total_chunks, grid);
  }
  static thread_local WgradMaps WM;
  if (build_wgrad_maps(launch, WM)) {
    wgrad_tc2_kernel<<<grid, WG_THREADS, WG_SMEM, stream>>>(launch, S, total_chunks, WM);
    MPQE_CHECK_LAUNCH("wgrad_tc2_kernel");
    return 0;
  }
  wgrad_tc_kernel<<<grid, THREADS, TC_SMEM, stream>>>(launch, S, total_chunks);
  MPQE_CHECK_LAUNCH("wgrad_tc_kernel");
  return 0;
}

size_t rank_packed_bytes(int64_t rows) { return (size_t)((rows + BM - 1) / BM) * (D / KC) * 2 * TILE_BYTES; }

// candidate rows [row_begin, row_begin + rows) of `table` -> tf32 hi/lo tile images (once per table shard)
int rank_pack_rows(const float* table, int64_t row_begin, int64_t rows, float* packed, cudaStream_t stream) {
  const int64_t ctiles = (rows + BM - 1) / BM;
  MPQE_CHECK_ARG(ctiles <= 65535, "mpqe_rank_counts_table: too many candidate rows per call (%lld)", (long long)rows);
  pack_rows_kernel<<<dim3(D / KC, (unsigned)ctiles), 256, 0, stream>>>(table + row_begin * D, rows, packed);
  MPQE_CHECK_LAUNCH("pack_rows_kernel");
  return 0;
}

int rank_counts_packed_tc(int64_t rows, const float* inv_norm, const float* q, const float* qinv, const float* pos,
                          int64_t B, unsigned long long* left, unsigned long long* right, const float* packed,
                          cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(rank_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RANK_SMEM));
    configured = true;
  }
  RankLaunch R;
  R.table = nullptr; R.row_begin = 0; R.rows = rows; R.inv_norm = inv_norm; R.q = q; R.qinv = qinv; R.pos = pos;
  R.B = B; R.left = left; R.right = right; R.packed = packed;
  const int64_t qtiles = (B + BM - 1) / BM;
  const int64_t ctiles = (rows + BM - 1) / BM;
  // candidate-range slices per query tile: minimise waves x (tiles per CTA + start-up), one CTA per SM at a time
  // (rounding the CTA count UP to the SM count, e.g. 160 CTAs on 148 SMs, doubles the run time)
  int64_t splits = 1, best = -1;
  for (int64_t sp = 1; sp <= ctiles && sp <= 4 * num_sms(); ++sp) {
    const int64_t waves = (qtiles * sp + num_sms() - 1) / num_sms();
    const int64_t cost = waves * ((ctiles + sp - 1) / sp + 6);      // + ~6 tile times to stage the query tile
    if (best < 0 || cost < best) best = cost, splits = sp;
  }
  R.splits = (int)splits;
  MPQE_CHECK_ARG(qtiles * splits < (1ll << 31), "mpqe_rank_counts_table: too many tiles");
  rank_tc_kernel<<<(unsigned)(qtiles * splits), THREADS, RANK_SMEM, stream>>>(R);
  MPQE_CHECK_LAUNCH("rank_tc_kernel");
  return 0;
}

int rank_counts_table_tc(const float* table, int64_t row_begin, int64_t rows, const float* inv_norm, const float* q,
                         const float* qinv, const float* pos, int64_t B, unsigned long long* left,
                         unsigned long long* right, float* packed, cudaStream_t stream) {
  if (int rc = rank_pack_rows(table, row_begin, rows, packed, stream)) return rc;
  return rank_counts_packed_tc(rows, inv_norm, q, qinv, pos, B, left, right, packed, stream);
}

}  // namespace mpqe

extern "C" int mpqe_b200_has_tcgen05(void) { return 1; }

using namespace mpqe;

int mpqe::tc_generation() {
  static int gen = 0;
  if (gen == 0) {
    const char* e = getenv("MPQE_LAYER_KERNEL");
    gen = (e != nullptr && atoi(e) == 1) ? 1 : 2;
  }
  return gen;
}

extern "C" int mpqe_pack_weights(const float* const* mats_host, int32_t count, float* packed, void* stream) {
  return mpqe_pack_weights_ex(mats_host, nullptr, count, packed, stream);
}

extern "C" int mpqe_pack_weights_ex(const float* const* mats_host, const uint8_t* transposed_host, int32_t count,
                                    float* packed, void* stream) {
  MPQE_CHECK_ARG(mats_host != nullptr && packed != nullptr && count >= 1, "mpqe_pack_weights: bad argument");
  if (tc_generation() == 2) return pack_weights_tc2(mats_host, transposed_host, count, packed, (cudaStream_t)stream);
  if (transposed_host != nullptr)
    for (int i = 0; i < count; ++i)
      MPQE_CHECK_ARG(!transposed_host[i], "mpqe_pack_weights_ex: the first-generation kernel's image has no transposed "
                     "form (matrix %d); pass a transposed copy", i);
  for (int base = 0; base < count; base += PACK_MAX) {
    static thread_local PackLaunch P;
    const int n = count - base < PACK_MAX ? count - base : PACK_MAX;
    for (int i = 0; i < n; ++i) {
      MPQE_CHECK_ARG(mats_host[base + i] != nullptr, "mpqe_pack_weights: matrix %d is null", base + i);
      P.m[i] = mats_host[base + i];
    }
    pack_weights_kernel<<<dim3(MPQE_D / 32, n), 256, 0, (cudaStream_t)stream>>>(
        P, packed + (int64_t)base * MPQE_PACKED_FLOATS);
    MPQE_CHECK_LAUNCH("pack_weights_kernel");
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int mpqe_debug_set_stats(void* buf) {
  return (int)cudaMemcpyToSymbol(mpqe::g_stats, &buf, sizeof(buf));
}

// debug only (not part of the public header): device buffer of >= 1 + 3*8000 int64 receiving CTA 0's event trace
extern "C" __attribute__((visibility("default"))) int mpqe_debug_set_trace(void* buf) {
  return (int)cudaMemcpyToSymbol(mpqe::g_trace, &buf, sizeof(buf));
}

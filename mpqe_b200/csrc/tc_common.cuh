// PTX wrappers shared by the tcgen05 kernels (layer_tc.cu, layer_tc2.cu): mbarriers, bulk copies, tensor-memory
// allocation, tcgen05.mma / .ld / .commit, shared-memory matrix descriptors and the tf32 hi/lo split.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace mpqe {
namespace tc {

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 16-byte store to a SHARED-window address.  The dynamic shared buffer is re-aligned through an integer round trip
// (align_1024), after which the compiler no longer knows the pointer's address space and emits generic ST.E.128 --
// measured on the producers of the layer kernels as ~100 cycles per store instead of a pipelined STS.128.
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the async proxy (TMA engine, no tensor map); completes on `bar`
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
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


// Static unit -> CTA assignment computed on the host (longest-processing-time first): with a few units per CTA
// (B = 4096: 224..768 units on 148 CTAs) round-robin leaves some CTAs with twice the work of others.
constexpr int SCHED_MAX_UNITS = 2048;
constexpr int SCHED_MAX_CTAS = 160;
// ---- TMA tensor maps (host) -------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda); nullptr when
// the driver does not have it.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}
// An activation-like operand [queries][slots][128] fp32 as a 3-D map {128, slots, queries}; box {box_k, 1, box_q}.
inline bool encode_rows_map(CUtensorMap* map, const float* base, int slots, int64_t queries, int box_k, int box_q,
                            CUtensorMapSwizzle swizzle) {
  EncodeTiledFn encode = tensor_map_encoder();
  if (encode == nullptr || slots <= 0 || queries <= 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)MPQE_D, (cuuint64_t)slots, (cuuint64_t)queries};
  const cuuint64_t strides[2] = {(cuuint64_t)MPQE_D * 4, (cuuint64_t)slots * MPQE_D * 4};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)box_k, 1u, (cuuint32_t)box_q};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

struct Schedule {
  int count;                                 // 0: round-robin (unit = blockIdx + k * gridDim)
  uint16_t start[SCHED_MAX_CTAS + 1];
  uint16_t unit[SCHED_MAX_UNITS];
  // layer kernel only: the decoded unit at the same position (so that a role needs two parameter loads per unit
  // instead of a dependent chain of ~25): query tile, and group << 24 | out slot << 16 | term bit mask
  uint16_t tile[SCHED_MAX_UNITS];
  uint32_t gsm[SCHED_MAX_UNITS];
};

// Longest-processing-time-first assignment of `units` work units (cost[u] stages each, + 1 for the per-unit epilogue /
// pipeline refill) to `grid` persistent CTAs; every CTA runs its units heaviest first.
inline void build_lpt(Schedule& S, const int* cost, int units, int grid) {
  static thread_local int order[SCHED_MAX_UNITS], owner[SCHED_MAX_UNITS];
  for (int i = 0; i < units; ++i) order[i] = i;
  std::stable_sort(order, order + units, [&](int a, int b) { return cost[a] > cost[b]; });
  long long load[SCHED_MAX_CTAS];
  int cnt[SCHED_MAX_CTAS];
  for (int c = 0; c < grid; ++c) load[c] = 0, cnt[c] = 0;
  for (int i = 0; i < units; ++i) {
    int best = 0;
    for (int c = 1; c < grid; ++c)
      if (load[c] < load[best]) best = c;
    owner[order[i]] = best;
    load[best] += cost[order[i]] + 1;
    ++cnt[best];
  }
  S.start[0] = 0;
  for (int c = 0; c < grid; ++c) S.start[c + 1] = (uint16_t)(S.start[c] + cnt[c]);
  int fill[SCHED_MAX_CTAS];
  for (int c = 0; c < grid; ++c) fill[c] = S.start[c];
  for (int i = 0; i < units; ++i) S.unit[fill[owner[order[i]]]++] = (uint16_t)order[i];
  S.count = units;
}


}  // namespace tc
}  // namespace mpqe

// Stand-alone timing probe of tcgen05.mma.kind::tf32 issue rates on sm_100a (debug tool; decides the shape of the
// layer kernel).  Every variant issues `reps` x 4 k-steps of MMAs back to back on resident operands, one commit at the
// end; thread 0 reports cycles per MMA.  Variants: both operands in shared memory ("SS") or A in tensor memory
// ("TS"), N = 128 / 256, no swizzle / 128-byte swizzle, and the same with `bg` background warps hammering shared
// memory with 16-byte stores (the producers of the real kernel).  TS correctness (A rows = TMEM lanes, k = columns)
// is checked against the exact product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/tc_rate_probe tools/tc_rate_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

struct Params {
  int n;          // 128 or 256
  int ts;         // 1: A operand from tensor memory
  int sw128;      // 1: 128-byte swizzle K-major operands (timing only)
  int reps;
  int bg;         // background warps storing to shared memory while the MMAs run
  int check;      // dump the accumulator (n == 128 only)
  int split3;     // 1: issue the 3xTF32 pattern (al*bh, ah*bl, ah*bh) instead of one MMA per k-step
  int fixed;      // 1: every MMA uses the k-step-0 descriptors (loop-invariant operands: pure issue rate)
  int alt;        // 1: consecutive MMAs alternate between two accumulators (columns 0.. and 256.. / 128..)
};

// A image: K-major [128][32] no swizzle: (r/8)*1024 + (k/4)*128 + (r%8)*16 + (k%4)*4   (16 KB)
// B image: same with n rows (n/8 row groups: 16 or 32 KB)
__global__ void __launch_bounds__(384) probe_kernel(const float* a_plain, const uint8_t* a_img, const uint8_t* b_img,
                                                    Params P, float* out, uint32_t* info) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sa = smem;               // 2 x 16 KB (hi, "lo" copy)
  uint8_t* sb = smem + 32768;       // 2 x 32 KB
  uint8_t* sbg = smem + 98304;      // 64 KB scratch for the background stores
  for (int i = tid * 16; i < 32768; i += blockDim.x * 16) *(uint4*)(sa + i) = *(const uint4*)(a_img + (i & 16383));
  for (int i = tid * 16; i < 65536; i += blockDim.x * 16) *(uint4*)(sb + i) = *(const uint4*)(b_img + (i & 32767));
  if (tid == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  // TMEM map: accumulator columns [0, 256), A (TS mode) columns [256, 256 + 64): hi 32 columns, "lo" 32 columns
  if (P.ts && warp < 4) {
    // thread = lane = row of A; its 32 k values go to 32 consecutive columns (32x32b shape)
    const int row = warp * 32 + lane;
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(a_plain[row * 32 + k]);
    for (int half = 0; half < 2; ++half) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256 + half * 32;
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
          "%29,%30,%31,%32};" ::"r"(taddr),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
          "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
          "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)P.n >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    const uint32_t lt = P.sw128 ? 2u : 0u;
    const uint32_t lbo = P.sw128 ? 16u : 128u, step = P.sw128 ? 32u : 256u;
    const int per = P.split3 ? 3 : 1;
    int cnt = 0;
    for (int rep = 0; rep < P.reps; ++rep)
      for (int j = 0; j < 4; ++j)
        for (int s = 0; s < per; ++s) {
          // 3xTF32 pattern: s=0 -> (a_lo, b_hi), s=1 -> (a_hi, b_lo), s=2 -> (a_hi, b_hi); the "lo" tiles are copies
          const int a_sel = (P.split3 && s == 0) ? 1 : 0, b_sel = (P.split3 && s == 1) ? 1 : 0;
          const int je = P.fixed ? 0 : j;
          const uint64_t da = make_desc(smem_u32(sa) + a_sel * 16384 + je * step, lbo, 1024, lt);
          const uint64_t db = make_desc(smem_u32(sb) + b_sel * 32768 + je * step, lbo, 1024, lt);
          const uint32_t acc = (j > 0 || rep > 0 || s > 0) ? 1u : 0u;
          ++cnt;
          if (P.ts) {
            const uint32_t ta = tmem + 256 + a_sel * 32 + je * 8;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                "r"(ta), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
          } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + ((P.alt && (cnt & 1)) ? (uint32_t)P.n : 0u)),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
          }
        }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  if (warp >= 4 && warp < 4 + P.bg) {
    // background: conflict-free 16-byte stores (512 B per warp instruction) until the MMAs are done
    uint8_t* dst = sbg + ((warp - 4) * 8192) + lane * 16;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    unsigned long long n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(dst + i * 512)), "f"(v.x), "f"(v.y), "f"(v.z),
                     "f"(v.w)
                     : "memory");
      ++n;
    }
    if (lane == 0 && warp == 4) info[3] = (uint32_t)n;   // 16 x 512 B per iteration per warp
  }
  if (warp == 0) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 26) && !ok; ++spin) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0)
          : "memory");
    }
    if (tid == 0) {
      const long long t1 = clock64();
      info[1] = ok;
      info[2] = (uint32_t)(t1 - t0);
      stop = 1;
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (P.check && warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
          "%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) out[row * 128 + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static int off_kmajor(int r, int k) { return (r / 8) * 1024 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4; }

int main() {
  const int K = 32;
  std::vector<float> A(128 * K), Bm(K * 256);
  srand(1);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& v : Bm) v = (float)((rand() % 17) - 8) / 8.0f;
  std::vector<uint8_t> ia(16384, 0), ib(32768, 0);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < K; ++k) memcpy(&ia[off_kmajor(m, k)], &A[m * K + k], 4);
  for (int n = 0; n < 256; ++n)
    for (int k = 0; k < K; ++k) memcpy(&ib[off_kmajor(n, k)], &Bm[k * 256 + n], 4);
  float *dap, *dout;
  uint8_t *da, *db;
  uint32_t* dinfo;
  CK(cudaMalloc(&dap, 128 * K * 4));
  CK(cudaMalloc(&da, 16384));
  CK(cudaMalloc(&db, 32768));
  CK(cudaMalloc(&dout, 128 * 128 * 4));
  CK(cudaMalloc(&dinfo, 64));
  CK(cudaMemcpy(dap, A.data(), 128 * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(da, ia.data(), 16384, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, ib.data(), 32768, cudaMemcpyHostToDevice));
  const int SMEM = 98304 + 65536 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));

  // ---- correctness of the TS form (A in tensor memory), N = 128, one pass over K = 32
  for (int ts = 0; ts < 2; ++ts) {
    Params P = {128, ts, 0, 1, 0, 1, 0, 0, 0};
    CK(cudaMemset(dout, 0xff, 128 * 128 * 4));
    probe_kernel<<<1, 384, SMEM>>>(dap, da, db, P, dout, dinfo);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * 128);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * Bm[k * 256 + n];
        const double err = fabs((double)out[m * 128 + n] - s);
        if (!(err <= 1e-3)) ++bad;
        if (err > maxerr) maxerr = err;
      }
    printf("correctness %s N=128: mismatches=%d maxerr=%.4f\n", ts ? "TS (A in TMEM)" : "SS", bad, maxerr);
  }

  // ---- what does kind::tf32 do with the low 13 mantissa bits of a raw fp32 operand: truncate or round?
  {
    std::vector<float> A2(128 * K);
    for (auto& v : A2) v = 0.5f + (float)rand() / (float)RAND_MAX;
    std::vector<uint8_t> ia2(16384, 0);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) memcpy(&ia2[off_kmajor(m, k)], &A2[m * K + k], 4);
    CK(cudaMemcpy(da, ia2.data(), 16384, cudaMemcpyHostToDevice));
    Params P = {128, 0, 0, 1, 0, 1, 0, 0, 0};
    probe_kernel<<<1, 384, SMEM>>>(dap, da, db, P, dout, dinfo);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * 128);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double e_trunc = 0, e_rna = 0, e_exact = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double st = 0, sr = 0, se = 0;
        for (int k = 0; k < K; ++k) {
          uint32_t u;
          memcpy(&u, &A2[m * K + k], 4);
          uint32_t ut = u & 0xffffe000u, ur = (u + 0x1000u) & 0xffffe000u;
          float ft, fr;
          memcpy(&ft, &ut, 4);
          memcpy(&fr, &ur, 4);
          const double b = Bm[k * 256 + n];
          st += (double)ft * b;
          sr += (double)fr * b;
          se += (double)A2[m * K + k] * b;
        }
        e_trunc = fmax(e_trunc, fabs(out[m * 128 + n] - st));
        e_rna = fmax(e_rna, fabs(out[m * 128 + n] - sr));
        e_exact = fmax(e_exact, fabs(out[m * 128 + n] - se));
      }
    printf("raw fp32 operand A: max |D - ref| with A truncated %.3e, A rounded (rna) %.3e, A exact %.3e\n", e_trunc, e_rna,
           e_exact);
    CK(cudaMemcpy(da, ia.data(), 16384, cudaMemcpyHostToDevice));
  }

  // ---- rates
  struct V {
    const char* name;
    Params p;
  };
  const int R = 128;
  V vs[] = {
      {"SS N=128 nosw", {128, 0, 0, R, 0, 0, 0, 0, 0}},
      {"SS N=256 nosw", {256, 0, 0, R, 0, 0, 0, 0, 0}},
      {"SS N=128 sw128", {128, 0, 1, R, 0, 0, 0, 0, 0}},
      {"SS N=256 sw128", {256, 0, 1, R, 0, 0, 0, 0, 0}},
      {"TS N=128 nosw", {128, 1, 0, R, 0, 0, 0, 0, 0}},
      {"TS N=256 nosw", {256, 1, 0, R, 0, 0, 0, 0, 0}},
      {"TS N=256 sw128", {256, 1, 1, R, 0, 0, 0, 0, 0}},
      {"SS N=128 nosw 3x", {128, 0, 0, R, 0, 0, 1, 0, 0}},
      {"SS N=256 nosw 3x", {256, 0, 0, R, 0, 0, 1, 0, 0}},
      {"TS N=128 nosw 3x", {128, 1, 0, R, 0, 0, 1, 0, 0}},
      {"TS N=256 nosw 3x", {256, 1, 0, R, 0, 0, 1, 0, 0}},
      {"SS N=128 nosw 3x + 8 store warps", {128, 0, 0, R, 8, 0, 1, 0, 0}},
      {"SS N=256 nosw 3x + 8 store warps", {256, 0, 0, R, 8, 0, 1, 0, 0}},
      {"TS N=128 nosw 3x + 8 store warps", {128, 1, 0, R, 8, 0, 1, 0, 0}},
      {"TS N=256 nosw 3x + 8 store warps", {256, 1, 0, R, 8, 0, 1, 0, 0}},
      {"TS N=256 nosw 3x + 4 store warps", {256, 1, 0, R, 4, 0, 1, 0, 0}},
      {"SS N=128 nosw fixed operands", {128, 0, 0, R, 0, 0, 0, 1, 0}},
      {"SS N=256 nosw fixed operands", {256, 0, 0, R, 0, 0, 0, 1, 0}},
      {"SS N=64 nosw", {64, 0, 0, R, 0, 0, 0, 0, 0}},
      {"SS N=128 nosw alt accumulators", {128, 0, 0, R, 0, 0, 0, 0, 1}},
      {"SS N=128 nosw 3x alt accumulators", {128, 0, 0, R, 0, 0, 1, 0, 1}},
      {"SS N=256 nosw alt accumulators", {256, 0, 0, R, 0, 0, 0, 0, 1}},
      {"SS N=256 nosw 3x alt accumulators", {256, 0, 0, R, 0, 0, 1, 0, 1}},
  };
  for (int grid : {1, 148}) {
    for (const V& v : vs) {
      CK(cudaMemset(dinfo, 0, 64));
      probe_kernel<<<grid, 384, SMEM>>>(dap, da, db, v.p, dout, dinfo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%-40s : error %s\n", v.name, cudaGetErrorString(e));
        return 1;
      }
      uint32_t inf[4];
      CK(cudaMemcpy(inf, dinfo, 16, cudaMemcpyDeviceToHost));
      const int mmas = v.p.reps * 4 * (v.p.split3 ? 3 : 1);
      const double cyc = (double)inf[2] / mmas;
      const double flops_per_cyc = 2.0 * 128 * v.p.n * 8 / cyc;
      printf("grid=%3d %-36s: %5d MMAs %8u cyc -> %6.1f cyc/MMA  %6.0f flop/cyc/SM  (ok=%u", grid, v.name, mmas, inf[2],
             cyc, flops_per_cyc, inf[1]);
      if (v.p.bg) printf(", bg stores %.1f B/cyc", (double)inf[3] * 16 * 512 * v.p.bg / inf[2]);
      printf(")\n");
    }
  }
  return 0;
}

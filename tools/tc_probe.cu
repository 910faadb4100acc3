// Stand-alone probe of tcgen05.mma.kind::tf32 shared-memory layouts / descriptors on sm_100a (debug tool).
// The HOST builds the byte image of the A and B tiles under a hypothesised canonical layout; the kernel copies
// the images to shared memory verbatim, issues the MMAs with the given descriptor parameters and dumps the 128x128
// accumulator.  The host compares with the exact product and reports which hypotheses hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/tc_probe tools/tc_probe.cu
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

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t version,
                                              uint32_t layout_type = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)(version & 3) << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

struct Params {
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_step, b_step;  // byte advance of the descriptor start per k-step
  int ksteps;
  int use_mask_form;        // CUTLASS 4-register disable_output_lane form
  int version;
  int layout_type;          // 0 none, 2 = SWIZZLE_128B (operand A)
  int b_layout_type;        // same for operand B
  int repeat;               // timing: issue the k-step sequence this many times
};

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, int a_bytes, int b_bytes,
                                                    Params P, float* out, uint32_t* info) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 65536;
  for (int i = tid * 16; i < a_bytes; i += 128 * 16) *(uint4*)(sa + i) = *(const uint4*)(a_img + i);
  for (int i = tid * 16; i < b_bytes; i += 128 * 16) *(uint4*)(sb + i) = *(const uint4*)(b_img + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  long long t0 = 0;
  if (tid == 0) {
    info[0] = tmem;
    t0 = clock64();
    for (int rep = 0; rep < P.repeat; ++rep)
    for (int j = 0; j < P.ksteps; ++j) {
      const uint64_t da = make_desc(smem_u32(sa) + j * P.a_step, P.a_lbo, P.a_sbo, P.version, P.layout_type);
      const uint64_t db = make_desc(smem_u32(sb) + j * P.b_step, P.b_lbo, P.b_sbo, P.version, P.b_layout_type);
      const uint32_t acc = (j > 0 || rep > 0) ? 1u : 0u;
      if (P.use_mask_form) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem),
            "l"(da), "l"(db), "r"(P.idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(da), "l"(db), "r"(P.idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  // wait (bounded)
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < (1u << 24) && !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(&bar)), "r"(0)
        : "memory");
  }
  if (tid == 0) {
    info[1] = ok;
    const long long t1 = clock64();
    info[2] = (uint32_t)(t1 - t0);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

// ---- host: layouts -------------------------------------------------------------------------------------------
// K-major tile [rows=128][k], element (r,k) -> byte offset
static int off_kmajor(int r, int k, int lbo, int sbo) { return (r / 8) * sbo + (k / 4) * lbo + (r % 8) * 16 + (k % 4) * 4; }
// MN-major tile [k][mn=128], element (k,mn) -> byte offset
static int off_mnmajor(int k, int mn, int lbo, int sbo) { return (mn / 4) * sbo + (k / 8) * lbo + (k % 8) * 16 + (mn % 4) * 4; }

// K-major SWIZZLE_128B tile [rows][32 k] (one 128-byte row per matrix row, 16-byte chunks XOR-ed with row%8)
static int off_kmajor_sw128(int r, int k) { return (r / 8) * 1024 + (r % 8) * 128 + (((k / 4) ^ (r % 8)) * 16) + (k % 4) * 4; }
// MN-major SWIZZLE_128B (CUTLASS canonical ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) with Swizzle<3,4,3>): a 1024-byte atom
// is 8 k-rows of 128 bytes (32 mn), chunks XOR-ed with k%8; mn blocks of 32 are LBO apart, k blocks of 8 SBO apart
static int off_mnmajor_sw128(int k, int mn, int lbo, int sbo) {
  return (mn / 32) * lbo + (k / 8) * sbo + (k % 8) * 128 + ((((mn % 32) / 4) ^ (k % 8)) * 16) + (mn % 4) * 4;
}

static const uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct Operand {
  int mn;        // 1: MN-major
  int sw128;     // 1: SWIZZLE_128B
  int lbo, sbo;  // descriptor fields (bytes); for the image: K-major (lbo = k-quad stride, sbo = 8-row stride),
                 // MN-major no swizzle (lbo = k-block stride, sbo = mn-block(4) stride), MN-major SW128 (lbo = mn-block(32)
                 // stride, sbo = k-block stride)
  int swap;      // write lbo into the SBO field and vice versa
  int step;      // descriptor start advance per k-step (bytes)
};
struct Hyp {
  const char* name;
  Operand a, b;
};

static int off_of(const Operand& o, int mn, int k) {
  if (!o.mn) return o.sw128 ? off_kmajor_sw128(mn, k) : off_kmajor(mn, k, o.lbo, o.sbo);
  return o.sw128 ? off_mnmajor_sw128(k, mn, o.lbo, o.sbo) : off_mnmajor(k, mn, o.lbo, o.sbo);
}

int main() {
  const int K = 32;  // 4 k-steps
  std::vector<float> A(128 * K), Bm(K * 128);   // A[m][k], B[k][n]  ->  D[m][n] = sum_k A[m][k] B[k][n]
  srand(1);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;   // exactly representable in tf32
  for (auto& v : Bm) v = (float)((rand() % 17) - 8) / 8.0f;
  std::vector<double> ref(128 * 128, 0.0);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * Bm[k * 128 + n];
      ref[m * 128 + n] = s;
    }
  uint8_t *da, *db;
  float* dout;
  uint32_t* dinfo;
  CK(cudaMalloc(&da, 65536));
  CK(cudaMalloc(&db, 65536));
  CK(cudaMalloc(&dout, 128 * 128 * 4));
  CK(cudaMalloc(&dinfo, 64));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));

  const Operand KM = {0, 0, 128, 1024, 0, 256};            // K-major, no swizzle (known good)
  const Operand KSW = {0, 1, 16, 1024, 0, 32};             // K-major SW128 (known good)
  Hyp hyps[] = {
      {"A:K  B:K  no-swizzle (control)", KM, KM},
      {"A:MN nosw lbo=128 sbo=512", {1, 0, 128, 512, 0, 128}, KM},
      {"A:MN nosw lbo=128 sbo=512 swapped", {1, 0, 128, 512, 1, 128}, KM},
      {"A:MN nosw lbo=4096 sbo=128", {1, 0, 4096, 128, 0, 4096}, KM},
      {"A:MN nosw lbo=4096 sbo=128 swapped", {1, 0, 4096, 128, 1, 4096}, KM},
      {"A:MN nosw lbo=128 sbo=528 (padded)", {1, 0, 128, 528, 0, 128}, KM},
      {"B:MN nosw lbo=128 sbo=512", KM, {1, 0, 128, 512, 0, 128}},
      {"A:MN B:MN nosw lbo=128 sbo=512", {1, 0, 128, 512, 0, 128}, {1, 0, 128, 512, 0, 128}},
      {"A:MN SW128 lbo=4096 sbo=1024", {1, 1, 4096, 1024, 0, 1024}, KM},
      {"A:MN SW128 lbo=4096 sbo=1024 swapped", {1, 1, 4096, 1024, 1, 1024}, KM},
      {"A:MN SW128 lbo=1024 sbo=4096", {1, 1, 1024, 4096, 0, 4096}, KM},
      {"A:MN SW128 lbo=1024 sbo=4096 swapped", {1, 1, 1024, 4096, 1, 4096}, KM},
      {"A:MN B:MN SW128 lbo=4096 sbo=1024", {1, 1, 4096, 1024, 0, 1024}, {1, 1, 4096, 1024, 0, 1024}},
      {"A:MN SW128, B:K SW128", {1, 1, 4096, 1024, 0, 1024}, KSW},
  };
  for (const Hyp& h : hyps) {
    std::vector<uint8_t> ia(65536, 0), ib(65536, 0);
    if (getenv("PROBE_FILL")) {   // diagnostic: operand A reads 1.0 wherever the hardware looks
      const float one = 1.0f;
      for (int i = 0; i < 65536; i += 4) memcpy(&ia[i], &one, 4);
    } else
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) memcpy(&ia[off_of(h.a, m, k)], &A[m * K + k], 4);
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < K; ++k) memcpy(&ib[off_of(h.b, n, k)], &Bm[k * 128 + n], 4);
    CK(cudaMemcpy(da, ia.data(), 65536, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, ib.data(), 65536, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xff, 128 * 128 * 4));
    Params P;
    P.idesc = IDESC_BASE | (h.a.mn ? (1u << 15) : 0) | (h.b.mn ? (1u << 16) : 0);
    P.a_lbo = h.a.swap ? h.a.sbo : h.a.lbo;
    P.a_sbo = h.a.swap ? h.a.lbo : h.a.sbo;
    P.b_lbo = h.b.swap ? h.b.sbo : h.b.lbo;
    P.b_sbo = h.b.swap ? h.b.lbo : h.b.sbo;
    P.a_step = h.a.step;
    P.b_step = h.b.step;
    P.ksteps = K / 8;
    P.use_mask_form = 0;
    P.version = 1;
    P.layout_type = h.a.sw128 ? 2 : 0;
    P.b_layout_type = h.b.sw128 ? 2 : 0;
    P.repeat = 1;
    probe_kernel<<<1, 128, 140 * 1024>>>(da, db, 65536, 65536, P, dout, dinfo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-40s : launch error %s\n", h.name, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> out(128 * 128);
    uint32_t info[2];
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(info, dinfo, 8, cudaMemcpyDeviceToHost));
    int bad = 0, zeros = 0;
    double maxerr = 0;
    for (int i = 0; i < 128 * 128; ++i) {
      const double err = fabs((double)out[i] - ref[i]);
      if (err > 1e-3) ++bad;
      if (out[i] == 0.f) ++zeros;
      if (err > maxerr) maxerr = err;
    }
    printf("%-40s : waited=%u mismatches=%5d zeros=%5d maxerr=%.3f  D[0][0..3]=%.3f %.3f %.3f %.3f (ref %.3f %.3f %.3f %.3f)\n",
           h.name, info[1], bad, zeros, maxerr, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    if (bad == 0) {   // timing: many MMAs back to back on the same tiles, all SMs busy
      P.repeat = 64;
      probe_kernel<<<148, 128, 140 * 1024>>>(da, db, 65536, 65536, P, dout, dinfo);
      CK(cudaDeviceSynchronize());
      uint32_t inf[3];
      CK(cudaMemcpy(inf, dinfo, 12, cudaMemcpyDeviceToHost));
      printf("    timing grid=148: %4d MMAs (128x128x8 tf32) in %8u cycles -> %.1f cycles/MMA\n", 64 * P.ksteps, inf[2],
             (double)inf[2] / (64 * P.ksteps));
    }
  }
  return 0;
}

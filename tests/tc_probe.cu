// Stand-alone probe of tcgen05.mma.kind::tf32 shared-memory layouts / descriptors on sm_100a (debug tool).
// The HOST builds the byte image of the A and B tiles under a hypothesised canonical layout; the kernel copies
// the images to shared memory verbatim, issues the MMAs with the given descriptor parameters and dumps the 128x128
// accumulator.  The host compares with the exact product and reports which hypotheses hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_probe/tc_probe tests/tc_probe.cu
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
  int layout_type;          // 0 none, 2 = SWIZZLE_128B
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
      const uint64_t db = make_desc(smem_u32(sb) + j * P.b_step, P.b_lbo, P.b_sbo, P.version, P.layout_type);
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

static const uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct Hyp {
  const char* name;
  int a_mn, b_mn;            // operand majors
  int a_lbo, a_sbo, b_lbo, b_sbo;
  int swap_desc;             // put LBO value in the SBO field and vice versa
  int mask_form;
  int version;
  int sw128;                 // both operands K-major SWIZZLE_128B
};

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

  Hyp hyps[] = {
      // name                       a_mn b_mn  a_lbo a_sbo b_lbo b_sbo swap mask ver sw128
      {"A:K B:K no-swizzle",           0, 0,   128, 1024, 128, 1024,  0, 0, 1, 0},
      {"A:K B:K SW128 lbo=16",         0, 0,    16, 1024,  16, 1024,  0, 0, 1, 1},
      {"A:K B:K SW128 lbo=0",          0, 0,     0, 1024,   0, 1024,  0, 0, 1, 1},
  };
  for (const Hyp& h : hyps) {
    std::vector<uint8_t> ia(65536, 0), ib(65536, 0);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) {
        const int off = h.sw128 ? off_kmajor_sw128(m, k)
                                : (h.a_mn ? off_mnmajor(k, m, h.a_lbo, h.a_sbo) : off_kmajor(m, k, h.a_lbo, h.a_sbo));
        memcpy(&ia[off], &A[m * K + k], 4);
      }
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < K; ++k) {
        const int off = h.sw128 ? off_kmajor_sw128(n, k)
                                : (h.b_mn ? off_mnmajor(k, n, h.b_lbo, h.b_sbo) : off_kmajor(n, k, h.b_lbo, h.b_sbo));
        memcpy(&ib[off], &Bm[k * 128 + n], 4);
      }
    CK(cudaMemcpy(da, ia.data(), 65536, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, ib.data(), 65536, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xff, 128 * 128 * 4));
    Params P;
    P.idesc = IDESC_BASE | (h.a_mn ? (1u << 15) : 0) | (h.b_mn ? (1u << 16) : 0);
    P.a_lbo = h.swap_desc ? h.a_sbo : h.a_lbo;
    P.a_sbo = h.swap_desc ? h.a_lbo : h.a_sbo;
    P.b_lbo = h.swap_desc ? h.b_sbo : h.b_lbo;
    P.b_sbo = h.swap_desc ? h.b_lbo : h.b_sbo;
    // per k-step (8 k) advance: K-major -> 2 k-quads = 2*lbo ; MN-major -> one 8-k group = lbo
    P.a_step = h.sw128 ? 32 : (h.a_mn ? h.a_lbo : 2 * h.a_lbo);
    P.b_step = h.sw128 ? 32 : (h.b_mn ? h.b_lbo : 2 * h.b_lbo);
    P.ksteps = K / 8;
    P.use_mask_form = h.mask_form;
    P.version = h.version;
    P.layout_type = h.sw128 ? 2 : 0;
    P.repeat = 1;
    probe_kernel<<<1, 128, 140 * 1024>>>(da, db, 65536, 65536, P, dout, dinfo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-34s : launch error %s\n", h.name, cudaGetErrorString(e));
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
    printf("%-34s : tmem=0x%08x waited=%u mismatches=%5d zeros=%5d maxerr=%.3f  D[0][0..3]=%.3f %.3f %.3f %.3f (ref %.3f %.3f %.3f %.3f)\n",
           h.name, info[0], info[1], bad, zeros, maxerr, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    // timing: many MMAs back to back on the same tiles (1 CTA, then all SMs busy)
    for (int grid : {1, 148}) {
      for (int rep : {16, 64}) {
        P.repeat = rep;
        probe_kernel<<<grid, 128, 140 * 1024>>>(da, db, 65536, 65536, P, dout, dinfo);
        CK(cudaDeviceSynchronize());
        uint32_t inf[3];
        CK(cudaMemcpy(inf, dinfo, 12, cudaMemcpyDeviceToHost));
        printf("    timing grid=%3d: %4d MMAs (128x128x8 tf32) in %8u cycles -> %.1f cycles/MMA\n", grid, rep * P.ksteps,
               inf[2], (double)inf[2] / (rep * P.ksteps));
      }
    }
  }
  return 0;
}

// Stand-alone probe of the global store / load patterns of the layer kernel's epilogue and producers (debug tool).
// A "unit" is one node slot of 256 consecutive queries: 256 rows of 512 B.  The activation buffer is either
// query-major [B, n, 128] (a unit's rows lie n*512 B apart) or slot-major [n, B, 128] (a unit is one contiguous
// 128 KB region).  Persistent grid of 148 CTAs; every CTA walks units round-robin.  L2 is filled with dirty lines
// before every launch (the bench does the same), so writes have to evict.
//   W4 : warp instruction = 32 lanes x 4 B  (lane = feature; the epilogue of layer_tc2_kernel), 4 warps = 4 row quarters
//   W16: warp instruction = 32 lanes x 16 B (one full 512-B row)
//   R16: 16-byte loads of full rows (the producers' pattern), summed so the loads are consumed
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/store_rate_probe tools/store_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

constexpr int D = 128;
constexpr int TILE = 256;

struct P {
  float* buf;
  int64_t B;
  int n;            // slots per query
  int slot_major;   // 0: [B,n,D]  1: [n,B,D]
  int units;        // tiles * n
  int mode;         // 0 W4, 1 W16, 2 R16, 3 W4 with all warps of a CTA on different units
  float* sink;
};

__device__ __forceinline__ float* row_ptr(const P& p, int64_t q, int slot) {
  return p.slot_major ? p.buf + ((int64_t)slot * p.B + q) * D : p.buf + (q * p.n + slot) * D;
}

__global__ void __launch_bounds__(512) probe_kernel(P p) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, warps = blockDim.x / 32;
  const int tiles = (int)(p.B / TILE);
  float acc = 0.f;
  for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
    const int slot = u / tiles;
    const int64_t q0 = (int64_t)(u % tiles) * TILE;
    if (p.mode == 0) {
      // warp w writes features 32*(w%4) .. +31 of queries (w/4)*TILE/(warps/4) ...
      const int groups = warps / 4, per = TILE / groups, g = warp / 4, f = (warp % 4) * 32 + lane;
      const float v = (float)(u + f);
#pragma unroll 8
      for (int j = 0; j < per; ++j) row_ptr(p, q0 + g * per + j, slot)[f] = v;
    } else if (p.mode == 1) {
      const float4 v = make_float4(u, lane, 0.f, 1.f);
#pragma unroll 8
      for (int j = warp; j < TILE; j += warps) reinterpret_cast<float4*>(row_ptr(p, q0 + j, slot))[lane] = v;
    } else if (p.mode == 2) {
#pragma unroll 8
      for (int j = warp; j < TILE; j += warps) {
        const float4 v = reinterpret_cast<const float4*>(row_ptr(p, q0 + j, slot))[lane];
        acc += v.x + v.y + v.z + v.w;
      }
    }
  }
  if (p.mode == 2 && acc == 123.456f) p.sink[threadIdx.x] = acc;
}

int main() {
  const int64_t B = 28672;
  const int n = 3;
  float *buf, *sink, *flush;
  const size_t bytes = (size_t)B * n * D * sizeof(float);
  const size_t flush_bytes = (size_t)256 << 20;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 4096));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemset(buf, 0, bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char* names[] = {"W4 ", "W16", "R16"};
  printf("buffer %.1f MB, %d units of 128 KB, 148 CTAs\n", bytes / 1e6, (int)(B / TILE) * n);
  for (int mode = 0; mode < 3; ++mode)
    for (int slot_major = 0; slot_major < 2; ++slot_major)
      for (int threads : {128, 256, 512})
        for (int ctas : {148, 296}) {
          P p{buf, B, n, slot_major, (int)(B / TILE) * n, mode, sink};
          float best = 1e9f;
          for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(flush, rep, flush_bytes));
            CK(cudaEventRecord(e0));
            probe_kernel<<<ctas, threads>>>(p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
          }
          CK(cudaGetLastError());
          printf("%s %s threads %3d ctas %3d : %7.1f us  %6.0f GB/s  (%.0f cycles per unit per SM at 1.965 GHz)\n",
                 names[mode], slot_major ? "slot-major " : "query-major", threads, ctas, best * 1e3,
                 bytes / (best * 1e-3) / 1e9, best * 1e-3 * 1.965e9 / (p.units / 148.0));
        }
  return 0;
}

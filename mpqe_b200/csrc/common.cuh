// Shared helpers for the mpqe_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mpqe_b200.h"

namespace mpqe {

constexpr int D = MPQE_D;  // embedding width (floats per row) = 512 B = one float4 per lane of a warp

void set_error(const char* fmt, ...);

#define MPQE_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      ::mpqe::set_error(__VA_ARGS__);  \
      return 1;                        \
    }                                  \
  } while (0)

#define MPQE_CHECK_LAUNCH(name)                                                        \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ::mpqe::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define MPQE_CUDA(call)                                                                \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::mpqe::set_error("%s failed: %s", #call, cudaGetErrorString(e__));              \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

// kernel-parameter blocks shared by the FFMA (layer_simt.cu) and tcgen05 (layer_tc.cu) versions
struct LayerLaunch {
  int num_groups;
  mpqe_layer_group_t g[MPQE_MAX_GROUPS];
};

struct WgradLaunch {
  int num_groups;
  int num_dests;
  float* partials;  // [sum chunks][D][D]
  int chunks[MPQE_MAX_DESTS];
  mpqe_wgrad_dest_t d[MPQE_MAX_DESTS];
  mpqe_layer_group_t g[MPQE_MAX_GROUPS];
  mpqe_wgrad_operand_t go[MPQE_MAX_GROUPS];
};

int layer_forward_simt(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream);
int layer_forward_tc(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream);
// second-generation tcgen05 layer kernel (layer_tc2.cu: queries on the N = 256 side) and its weight image
int layer_forward_tc2(const mpqe_layer_group_t* groups, int num_groups, cudaStream_t stream);
int pack_weights_tc2(const float* const* mats_host, const uint8_t* transposed_host, int32_t count, float* packed,
                     cudaStream_t stream);
// 1: layer_tc.cu (128-query units), 2: layer_tc2.cu (default); MPQE_LAYER_KERNEL selects, read once
int tc_generation();
int layer_wgrad_tc_launch(const WgradLaunch& launch, int total_chunks, cudaStream_t stream);
size_t rank_packed_bytes(int64_t rows);   // bytes of the pre-split candidate tile images (tcgen05 rank kernel)
int rank_counts_table_tc(const float* table, int64_t row_begin, int64_t rows, const float* inv_norm, const float* q,
                         const float* qinv, const float* pos, int64_t B, unsigned long long* left,
                         unsigned long long* right, float* packed, cudaStream_t stream);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

// streaming 16-byte load that does not allocate in L1 (rows touched once)
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

}  // namespace mpqe

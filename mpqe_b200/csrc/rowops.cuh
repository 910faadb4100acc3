// Warp-per-row device helpers shared by the single-item (gather_score.cu) and multi-item (multi.cu) row kernels.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace mpqe {

constexpr int ROW_THREADS = 256;           // 8 warps = 8 rows per CTA
constexpr float COS_EPS = 1e-8f;           // F.cosine_similarity default eps

__device__ __forceinline__ int64_t resolve_row(const int64_t* id2row, const int64_t* ids, int64_t stride, int64_t i) {
  const int64_t id = ids[i * stride];
  return id2row != nullptr ? id2row[id] : id;
}

// Base address of the table that holds `row`.  With peer tables (data-parallel training, entity tables owned
// row-range-wise: rank r owns rows [r*chunk, (r+1)*chunk)) the row is read from its OWNER's copy -- peers[r] is rank r's
// full-size table as mapped into this process -- so a rank always sees the owner's latest update; otherwise `table`.
__device__ __forceinline__ const float* table_of_row(const float* table, const float* const* peers, int64_t chunk,
                                                     int64_t row) {
  return peers != nullptr ? peers[row / chunk] : table;
}

__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ float4 div4(const float4& v, float s) { return make_float4(v.x / s, v.y / s, v.z / s, v.w / s); }

// y = row / ||row||  (division, like Tensor.div in encoders.py:43); returns ||row||
__device__ __forceinline__ float normalize_row(const float* table, int64_t row, int lane, float4& y) {
  const float4 v = *reinterpret_cast<const float4*>(table + row * D + lane * 4);
  const float nrm = sqrtf(warp_sum(dot4(v, v)));
  y = div4(v, nrm);
  return nrm;
}

// d(row) = (g - (g.y) y) / ||row||
__device__ __forceinline__ float4 normalize_bwd(const float4& g, const float4& y, float nrm) {
  const float gy = warp_sum(dot4(g, y));
  return make_float4((g.x - gy * y.x) / nrm, (g.y - gy * y.y) / nrm, (g.z - gy * y.z) / nrm, (g.w - gy * y.w) / nrm);
}

// ---- cosine scoring ---------------------------------------------------------------------------------------------
struct Cos {
  float score, nq, nqc, ny, nyc;  // norms and their eps-clamped versions
};

__device__ __forceinline__ Cos cosine(const float4& q, const float4& y) {
  Cos c;
  c.nq = sqrtf(warp_sum(dot4(q, q)));
  c.ny = sqrtf(warp_sum(dot4(y, y)));
  c.nqc = fmaxf(c.nq, COS_EPS);
  c.nyc = fmaxf(c.ny, COS_EPS);
  // torch: sum((x1 / clamp_min(|x1|, eps)) * (x2 / clamp_min(|x2|, eps)))
  const float4 a = div4(q, c.nqc), b = div4(y, c.nyc);
  c.score = warp_sum(dot4(a, b));
  return c;
}

// gradient of score wrt q (dsq) and wrt y (dsy), times upstream g
__device__ __forceinline__ void cosine_bwd(const float4& q, const float4& y, const Cos& c, float g, float4& dq,
                                           float4& dy) {
  const float kq = c.nq > COS_EPS ? c.score / (c.nqc * c.nq) : 0.f;
  const float ky = c.ny > COS_EPS ? c.score / (c.nyc * c.ny) : 0.f;
  const float iq = 1.f / c.nqc, iy = 1.f / c.nyc;
  dq = make_float4(g * (y.x * iy * iq - kq * q.x), g * (y.y * iy * iq - kq * q.y), g * (y.z * iy * iq - kq * q.z),
                   g * (y.w * iy * iq - kq * q.w));
  dy = make_float4(g * (q.x * iq * iy - ky * y.x), g * (q.y * iq * iy - ky * y.y), g * (q.z * iq * iy - ky * y.z),
                   g * (q.w * iq * iy - ky * y.w));
}


inline unsigned row_blocks(int64_t rows) { return (unsigned)((rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32)); }

}  // namespace mpqe

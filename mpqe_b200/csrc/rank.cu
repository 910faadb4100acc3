// Full-entity ranking: (count_lt, count_le) of every candidate row of a table shard against each query's positive
// score, without materialising the [B, N] score matrix.  fp32 FFMA version (the tcgen05 version is rank_tc.cu).
//
// New capability named by the north star ("full-entity ranking eval ... per-shard rank counts merged with NCCL");
// the reference only ranks against <=1000 stored negatives (utils.py:72-95, model.py:454-460).  Score semantics
// are those of model.py:451-452: cos(q_b, normalised row).
#include "common.cuh"

namespace mpqe {
namespace {

constexpr int BM = 64;    // candidate rows per CTA
constexpr int BN = 128;   // queries per CTA
constexpr int KC = 32;
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int A_PITCH = KC + 4;
constexpr int A_STAGE = BM * A_PITCH;
constexpr int B_STAGE = KC * BN;
constexpr int STAGE_FLOATS = A_STAGE + B_STAGE;
constexpr size_t RANK_SMEM = size_t(STAGES) * STAGE_FLOATS * sizeof(float);
constexpr float COS_EPS = 1e-8f;

// qt[k * Bp + b] = q[b, k] (zero for b >= B); qinv[b] = 1 / max(||q_b||, eps)
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ q, int64_t B, int64_t Bp,
                                                           float* __restrict__ qt, float* __restrict__ qinv) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= Bp) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b < B) v = *reinterpret_cast<const float4*>(q + b * D + lane * 4);
  const float nrm = sqrtf(warp_sum(dot4(v, v)));
  qt[(int64_t)(lane * 4 + 0) * Bp + b] = v.x;
  qt[(int64_t)(lane * 4 + 1) * Bp + b] = v.y;
  qt[(int64_t)(lane * 4 + 2) * Bp + b] = v.z;
  qt[(int64_t)(lane * 4 + 3) * Bp + b] = v.w;
  if (lane == 0 && b < B) qinv[b] = 1.f / fmaxf(nrm, COS_EPS);
}

__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ table, int64_t row_begin,
                                                           int64_t rows, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4 v = *reinterpret_cast<const float4*>(table + (row_begin + r) * D + lane * 4);
  const float nrm = sqrtf(warp_sum(dot4(v, v)));
  if (lane == 0) inv[r] = 1.f / nrm;
}

__global__ void __launch_bounds__(THREADS, 2) rank_counts_table_kernel(
    const float* __restrict__ table, int64_t row_begin, int64_t rows, const float* __restrict__ inv_norm,
    const float* __restrict__ qt, int64_t Bp, const float* __restrict__ qinv, const float* __restrict__ pos, int64_t B,
    unsigned long long* __restrict__ left, unsigned long long* __restrict__ right) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * BM;
  const int64_t b0 = (int64_t)blockIdx.y * BN;

  auto load_stage = [&](int step) {
    float* As = smem + (step % STAGES) * STAGE_FLOATS;
    float* Bs = As + A_STAGE;
    const int kc = step * KC;
    {
      const int f4 = tid & 7;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 3) + 32 * i;
        int64_t row = r0 + r;
        if (row >= rows) row = rows - 1;
        cp_async16(As + r * A_PITCH + f4 * 4, table + (row_begin + row) * D + kc + f4 * 4);
      }
    }
    {
      const int f4 = tid & 31;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = (tid >> 5) + 8 * i;
        cp_async16(Bs + k * BN + f4 * 4, qt + (int64_t)(kc + k) * Bp + b0 + f4 * 4);
      }
    }
  };

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  constexpr int NSTEPS = D / KC;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    load_stage(s);
    cp_async_commit();
  }
#pragma unroll
  for (int step = 0; step < NSTEPS; ++step) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (step + STAGES - 1 < NSTEPS) load_stage(step + STAGES - 1);
    cp_async_commit();
    const float* As = smem + (step % STAGES) * STAGE_FLOATS;
    const float* Bs = As + A_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty * 4 + i) * A_PITCH + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b0v = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * BN + tx * 4);
        const float4 b1v = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * BN + 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, b0v.x, acc[i][0]);
          acc[i][1] = fmaf(av, b0v.y, acc[i][1]);
          acc[i][2] = fmaf(av, b0v.z, acc[i][2]);
          acc[i][3] = fmaf(av, b0v.w, acc[i][3]);
          acc[i][4] = fmaf(av, b1v.x, acc[i][4]);
          acc[i][5] = fmaf(av, b1v.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1v.z, acc[i][6]);
          acc[i][7] = fmaf(av, b1v.w, acc[i][7]);
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: scale, compare with the positive score, count per query column ----
  int* cnt = reinterpret_cast<int*>(smem);  // [2][16][BN]
  int lt[8], le[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) lt[j] = le[j] = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = r0 + ty * 4 + i;
    if (row >= rows) continue;
    const float inr = inv_norm[row];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t b = b0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (b >= B) continue;
      const float s = acc[i][j] * inr * qinv[b];
      const float p = pos[b];
      lt[j] += s < p;
      le[j] += s <= p;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
    cnt[ty * BN + col] = lt[j];
    cnt[16 * BN + ty * BN + col] = le[j];
  }
  __syncthreads();
  if (tid < BN) {
    int s_lt = 0, s_le = 0;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      s_lt += cnt[g * BN + tid];
      s_le += cnt[16 * BN + g * BN + tid];
    }
    const int64_t b = b0 + tid;
    if (b < B) {  // integer atomics: order-independent, hence bit-exact
      if (s_lt) atomicAdd(left + b, (unsigned long long)s_lt);
      if (s_le) atomicAdd(right + b, (unsigned long long)s_le);
    }
  }
}

struct RankWs {
  float *qt, *qinv, *inv_norm, *packed;
  int64_t Bp;
  size_t bytes;
};

RankWs carve_rank(void* ws, int64_t B, int64_t rows) {
  RankWs w;
  w.Bp = (B + BN - 1) / BN * BN;
  char* p = (char*)ws;
  size_t off = 0;
  w.qt = (float*)(p + off); off += align_up((size_t)D * w.Bp * sizeof(float), 256);
  w.qinv = (float*)(p + off); off += align_up((size_t)w.Bp * sizeof(float), 256);
  w.inv_norm = (float*)(p + off); off += align_up((size_t)(rows > 0 ? rows : 1) * sizeof(float), 256);
  off = align_up(off, 1024);
  w.packed = (float*)(p + off); off += rank_packed_bytes(rows);   // tf32 hi/lo tile images of the candidate rows
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace mpqe

using namespace mpqe;

// ---- two-phase form: the table-dependent half (1/||row||, tf32 hi/lo tile images of the shard) is prepared ONCE and
// reused by every batch of an evaluation; only the query-dependent half runs per batch --------------------------------
namespace {
struct TableWs {
  float* inv_norm;
  float* packed;
  size_t bytes;
};
TableWs carve_table(void* ws, int64_t rows, int use_tc) {
  TableWs w;
  char* p = (char*)ws;
  size_t off = 0;
  w.inv_norm = (float*)(p + off); off += align_up((size_t)(rows > 0 ? rows : 1) * sizeof(float), 1024);
  w.packed = (float*)(p + off); off += use_tc ? rank_packed_bytes(rows) : 0;
  w.bytes = off;
  return w;
}
struct QueryWs {
  float *qt, *qinv;
  int64_t Bp;
  size_t bytes;
};
QueryWs carve_query(void* ws, int64_t B) {
  QueryWs w;
  w.Bp = (B + BN - 1) / BN * BN;
  char* p = (char*)ws;
  size_t off = 0;
  w.qt = (float*)(p + off); off += align_up((size_t)D * w.Bp * sizeof(float), 256);
  w.qinv = (float*)(p + off); off += align_up((size_t)w.Bp * sizeof(float), 256);
  w.bytes = off;
  return w;
}
}  // namespace

namespace mpqe {
int rank_pack_rows(const float* table, int64_t row_begin, int64_t rows, float* packed, cudaStream_t stream);
int rank_counts_packed_tc(int64_t rows, const float* inv_norm, const float* q, const float* qinv, const float* pos,
                          int64_t B, unsigned long long* left, unsigned long long* right, const float* packed,
                          cudaStream_t stream);
}

extern "C" size_t mpqe_rank_table_workspace_bytes(int64_t rows, int32_t use_tensor_cores) {
  return carve_table(nullptr, rows, use_tensor_cores).bytes;
}

extern "C" int mpqe_rank_table_prepare(const float* table, int64_t row_begin, int64_t row_end, void* table_ws,
                                       size_t table_ws_bytes, int32_t use_tensor_cores, void* stream) {
  MPQE_CHECK_ARG(table && row_begin >= 0 && row_end >= row_begin, "mpqe_rank_table_prepare: bad argument");
  const int64_t rows = row_end - row_begin;
  if (rows == 0) return 0;
  TableWs w = carve_table(table_ws, rows, use_tensor_cores);
  MPQE_CHECK_ARG(table_ws && table_ws_bytes >= w.bytes, "mpqe_rank_table_prepare: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  row_inv_norm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(table, row_begin, rows, w.inv_norm);
  MPQE_CHECK_LAUNCH("row_inv_norm_kernel");
  if (use_tensor_cores) return rank_pack_rows(table, row_begin, rows, w.packed, st);
  return 0;
}

extern "C" size_t mpqe_rank_query_workspace_bytes(int64_t B) { return carve_query(nullptr, B).bytes; }

extern "C" int mpqe_rank_counts_prepared(const float* q, int64_t B, const float* pos, const float* table,
                                         int64_t row_begin, int64_t row_end, const void* table_ws, int64_t* left,
                                         int64_t* right, void* query_ws, size_t query_ws_bytes,
                                         int32_t use_tensor_cores, void* stream) {
  MPQE_CHECK_ARG(q && pos && table && table_ws && left && right && B >= 1 && row_begin >= 0 && row_end >= row_begin,
                 "mpqe_rank_counts_prepared: bad argument");
  const int64_t rows = row_end - row_begin;
  if (rows == 0) return 0;
  TableWs tw = carve_table(const_cast<void*>(table_ws), rows, use_tensor_cores);
  QueryWs qw = carve_query(query_ws, B);
  MPQE_CHECK_ARG(query_ws && query_ws_bytes >= qw.bytes, "mpqe_rank_counts_prepared: query workspace too small");
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(rank_counts_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)RANK_SMEM));
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  prep_queries_kernel<<<(unsigned)((qw.Bp + 7) / 8), 256, 0, st>>>(q, B, qw.Bp, qw.qt, qw.qinv);
  MPQE_CHECK_LAUNCH("prep_queries_kernel");
  if (use_tensor_cores)
    return rank_counts_packed_tc(rows, tw.inv_norm, q, qw.qinv, pos, B, (unsigned long long*)left,
                                 (unsigned long long*)right, tw.packed, st);
  dim3 grid((unsigned)((rows + BM - 1) / BM), (unsigned)(qw.Bp / BN));
  MPQE_CHECK_ARG(grid.y <= 65535, "mpqe_rank_counts_prepared: too many queries (%lld)", (long long)B);
  rank_counts_table_kernel<<<grid, THREADS, RANK_SMEM, st>>>(table, row_begin, rows, tw.inv_norm, qw.qt, qw.Bp, qw.qinv,
                                                            pos, B, (unsigned long long*)left,
                                                            (unsigned long long*)right);
  MPQE_CHECK_LAUNCH("rank_counts_table_kernel");
  return 0;
}

extern "C" size_t mpqe_rank_counts_table_workspace_bytes(int64_t B, int64_t rows) {
  return carve_rank(nullptr, B, rows).bytes;
}

extern "C" int mpqe_rank_counts_table(const float* q, int64_t B, const float* pos, const float* table,
                                      int64_t row_begin, int64_t row_end, int64_t* left, int64_t* right,
                                      void* workspace, size_t workspace_bytes, int32_t use_tensor_cores,
                                      void* stream) {
  MPQE_CHECK_ARG(q && pos && table && left && right && B >= 1 && row_begin >= 0 && row_end >= row_begin,
                 "mpqe_rank_counts_table: bad argument");
  const int64_t rows = row_end - row_begin;
  if (rows == 0) return 0;
  RankWs w = carve_rank(workspace, B, rows);
  MPQE_CHECK_ARG(workspace && workspace_bytes >= w.bytes, "mpqe_rank_counts_table: workspace too small");
  static bool configured = false;
  if (!configured) {
    MPQE_CUDA(cudaFuncSetAttribute(rank_counts_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)RANK_SMEM));
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  prep_queries_kernel<<<(unsigned)((w.Bp + 7) / 8), 256, 0, st>>>(q, B, w.Bp, w.qt, w.qinv);
  MPQE_CHECK_LAUNCH("prep_queries_kernel");
  row_inv_norm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(table, row_begin, rows, w.inv_norm);
  MPQE_CHECK_LAUNCH("row_inv_norm_kernel");
  if (use_tensor_cores)
    return rank_counts_table_tc(table, row_begin, rows, w.inv_norm, q, w.qinv, pos, B, (unsigned long long*)left,
                                (unsigned long long*)right, w.packed, st);
  dim3 grid((unsigned)((rows + BM - 1) / BM), (unsigned)(w.Bp / BN));
  MPQE_CHECK_ARG(grid.y <= 65535, "mpqe_rank_counts_table: too many queries (%lld)", (long long)B);
  rank_counts_table_kernel<<<grid, THREADS, RANK_SMEM, st>>>(table, row_begin, rows, w.inv_norm, w.qt, w.Bp, w.qinv,
                                                            pos, B, (unsigned long long*)left,
                                                            (unsigned long long*)right);
  MPQE_CHECK_LAUNCH("rank_counts_table_kernel");
  return 0;
}

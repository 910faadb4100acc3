/*
 * mpqe_b200 -- C ABI of the B200-native MPQE query-encoding hot path.
 *
 * The reference (dfdazac/mpqe) is pure Python: its "plugin boundary" for this path is the set of
 * PyTorch / torch_geometric / torch_scatter library calls made from
 *     mpqe/data_utils.py:377-409   (batch layout),
 *     mpqe/encoders.py:29-45       (embedding gather + L2 normalise),
 *     mpqe/model.py:269-305        (RGCNConv: per-edge transform, scatter-add, root, bias),
 *     mpqe/model.py:380-398,497-515 (readouts),
 *     mpqe/model.py:451-460,483-485 (cosine scoring, margin loss),
 *     mpqe/utils.py:25-32          (percentile rank counts)
 * and their autograd backward.  Every entry point below replaces one such group of calls and cites it.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes, a `cudaStream_t` passed as `void*`; no torch types.
 *   - all floating point data is fp32, ids are int64 (as in the reference), d (embedding width) must be 128.
 *   - activations are row-major [B, slots, d] ("query-major": row b*slots+i, the reference's `x.reshape(-1, d)`).
 *   - weight matrices are row-major [d_in, d_out] (the layout of `RGCNConv.basis[r]` / `.root`, model.py:244-249).
 *   - functions never allocate, never synchronise the device and are re-entrant per stream; scratch memory is a
 *     caller-provided workspace whose size comes from the matching *_workspace_bytes query.
 *   - return value: 0 = ok; non-zero = error, text via mpqe_b200_last_error() (thread-local).
 */
#ifndef MPQE_B200_H_
#define MPQE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MPQE_API __attribute__((visibility("default")))
#else
#define MPQE_API
#endif

#define MPQE_D 128            /* embedding width the kernels are specialised for (train.py:19 default) */
#define MPQE_MAX_GROUPS 8     /* formula groups fused into one launch */
#define MPQE_MAX_TERMS 16     /* (edge + self-loop) terms per group */
#define MPQE_MAX_SLOTS 8      /* output node slots per query */
#define MPQE_MAX_DESTS 64     /* distinct weight matrices receiving a gradient in one launch */

/* epilogues of the layer kernel */
#define MPQE_EPI_NONE 0
#define MPQE_EPI_RELU 1       /* F.relu between passes, model.py:437 */
#define MPQE_EPI_MASK 2       /* multiply by (mask > 0): the ReLU backward */

/* One term of a layer:  out[q, out_slot, :] += A[q, a_slot, :] @ M   with A = a + (q*a_slots + a_slot)*d.
 * a_slots == 0 broadcasts one row to every query (variable-type embeddings, model.py:421).
 * An edge e of the query template is the term (a_slot = src[e], M = basis[rel[e]], out_slot = dst[e]);
 * the self-loop of node i is (a_slot = i, M = root, out_slot = i)  -- model.py:292-294, 301. */
typedef struct {
  const float* a;
  const float* m;
  const float* m_packed; /* optional: mpqe_pack_weights image of m (tcgen05 path stages it with one bulk copy) */
  int32_t a_slots;
  int16_t a_slot;
  int16_t out_slot;
} mpqe_term_t;

/* One formula group (all queries share one template, data_utils.py:293-311, 394-405). */
typedef struct {
  int64_t num_queries;
  int32_t num_terms;
  int32_t num_out_slots;                 /* output slots computed for this group */
  mpqe_term_t terms[MPQE_MAX_TERMS];
  float* out;                            /* [num_queries, out_slots, d] */
  int32_t out_slots;                     /* slots per query in `out` (row stride / d) */
  int32_t epilogue;                      /* MPQE_EPI_* */
  const float* bias;                     /* [d] or NULL (model.py:303-304); out slot j adds bias_scale[j] *
                                            bias[j*bias_slot_stride .. +d] */
  float bias_scale[MPQE_MAX_SLOTS];      /* bias multiplier per out slot (n for a fused sum readout) */
  int16_t out_slot_map[MPQE_MAX_SLOTS];  /* out slot j is stored at out[q, out_slot_map[j], :] */
  const float* mask;                     /* EPI_MASK: [num_queries, mask_slots, d], slot = out_slot_map[j] */
  int32_t mask_slots;
  int32_t bias_slot_stride;              /* 0: one bias vector for all slots; d: a vector per out slot (used for the
                                            contribution of batch-constant input rows, computed once per group) */
  /* ReLU sign bits, an optional 32x smaller form of the ReLU-backward mask.  Word ((q/32)*S + slot)*d + c holds, for
   * feature c of node slot `slot`, bit (q%32) = (value of query q > 0); S = out_slots (written) / mask_slots (read).
   * relu_bits_out (EPI_RELU launches): the kernel also writes the words of the slots it stores, [ceil(B/32), out_slots,
   * d]; mask_bits (EPI_MASK launches): used instead of reading `mask` when non-NULL (`mask` must still be given: kernels
   * without a bit path read it).  Only the tcgen05 kernel produces / consumes the bits. */
  uint32_t* relu_bits_out;
  const uint32_t* mask_bits;
} mpqe_layer_group_t;

/* One weight-gradient destination: dM = sum over every term (of every group) whose `m` equals `m_fwd` of
 * A[:, a_slot]^T @ G[:, out_slot]   (autograd of bmm/index_select/matmul, model.py:292-294, 301). */
typedef struct {
  const float* m_fwd;
  float* dm;              /* [d, d] */
  int32_t accumulate;     /* 0: overwrite, 1: add to existing contents */
  int32_t reserved;
} mpqe_wgrad_dest_t;

/* Per-group gradient operand for the weight-gradient kernel: G = g + (q*g_slots + slot_map[out_slot])*d. */
typedef struct {
  const float* g;
  int32_t g_slots;                       /* 1 with slot_map all 0 broadcasts dq to all slots (sum readout) */
  int16_t slot_map[MPQE_MAX_SLOTS];
} mpqe_wgrad_operand_t;

MPQE_API const char* mpqe_b200_last_error(void);
MPQE_API int mpqe_b200_version(void);
/* sizeof the ABI structs (0: term, 1: layer group, 2: wgrad dest, 3: wgrad operand, 4: gather item, 5: margin item,
 * 6: colsum item, 7: matsum item, 8: l2 item, 9: adam item, 10: adam table) so bindings can self-check */
MPQE_API int mpqe_b200_sizeof(int which);
/* 1 if the library was built with the tcgen05 (sm_100a tensor core) layer kernels */
MPQE_API int mpqe_b200_has_tcgen05(void);

/* ---- a1: batch layout (data_utils.py:394-405 + PyG Batch.from_data_list) ---------------------------------
 * edge_index[2, B*E], edge_type[B*E], batch[B*n] (all int64), bit-exact with the reference. */
MPQE_API int mpqe_build_query_graph(int32_t n, int32_t E, const int32_t* tmpl_src_host, const int32_t* tmpl_dst_host,
                           const int64_t* tmpl_rel_host, int64_t B,
                           int64_t* edge_index, int64_t* edge_type, int64_t* batch, void* stream);

/* Relation-sorted edge layout: perm = stable argsort of edge_type, seg_offsets[R+1] = exclusive histogram.
 * (new artefact of the north star; CPU definition: torch.sort(edge_type, stable=True)). */
MPQE_API size_t mpqe_relation_sort_workspace_bytes(int64_t num_edges, int32_t num_relations);
MPQE_API int mpqe_relation_sort(const int64_t* edge_type, int64_t num_edges, int32_t num_relations,
                       int64_t* perm, int64_t* seg_offsets, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a3/a4: embedding gather + L2 normalise (data_utils.py:35, utils.py:22, encoders.py:41-43) -------------
 * row = id2row ? id2row[ids[i]] : ids[i];  out[i*out_stride .. +d] = table[row] / ||table[row]||  (no eps).
 * inv_norm[i] (optional) receives 1/||row|| for the backward. */
MPQE_API int mpqe_gather_normalize_fwd(const float* table, int64_t table_rows, const int64_t* id2row,
                              const int64_t* ids, int64_t ids_stride, int64_t count,
                              float* out, int64_t out_stride, float* inv_norm, void* stream);
/* d(raw row) = (g - (g.y) y) * inv_norm, y = normalised row re-derived from the table;
 * rows_out[i, :] and rows_id[i] (= table row) feed mpqe_sparse_rows_combine. */
MPQE_API int mpqe_gather_normalize_bwd(const float* table, const int64_t* id2row, const int64_t* ids, int64_t ids_stride,
                              int64_t count, const float* grad, int64_t grad_stride,
                              float* rows_out, int64_t* rows_id, void* stream);
/* a5 (model.py:418-422): x[b, i<a] = normalised anchor rows, x[b, i>=a] = mode_embeddings[var_ids[i-a]].
 * tables/anchor ids are given per anchor slot. */
MPQE_API int mpqe_broadcast_rows(const float* src, const int64_t* src_rows, int32_t num_rows,
                        float* out, int64_t out_stride, int64_t count, void* stream);

/* ---- a6-a9, a12: fused R-GCN layer / MLP layer (model.py:269-305, 435-441, 507-509) -----------------------
 * For every group: out[q, j] = epilogue( sum_{terms t: out_slot==j} A_t[q] @ M_t + bias_scale[j]*bias ).
 * With transposed matrices and swapped slots the same entry point computes the input gradient. */
MPQE_API int mpqe_layer_forward(const mpqe_layer_group_t* groups_host, int32_t num_groups, int32_t use_tensor_cores,
                       void* stream);

/* Pre-split ("packed") weights for the tensor-core layer kernel: for each [d,d] matrix, MPQE_PACKED_FLOATS floats
 * holding the tf32 hi / lo parts laid out as the kernel's shared-memory operand tiles.  mats_host is a HOST array of
 * `count` device pointers; packed receives count * MPQE_PACKED_FLOATS floats.  Repack whenever the weights change. */
#define MPQE_PACKED_FLOATS (2 * MPQE_D * MPQE_D)
MPQE_API int mpqe_pack_weights(const float* const* mats_host, int32_t count, float* packed, void* stream);
/* The same with an optional HOST array of per-matrix flags: transposed_host[i] != 0 packs the image of mats[i]^T (the
 * matrix the input-gradient launches multiply by) straight from mats[i] -- no transposed copy has to exist first. */
MPQE_API int mpqe_pack_weights_ex(const float* const* mats_host, const uint8_t* transposed_host, int32_t count,
                                  float* packed, void* stream);

/* Weight gradients of the same term lists (deterministic: fixed split over queries, ordered reduction). */
MPQE_API size_t mpqe_layer_wgrad_workspace_bytes(int32_t num_dests, int32_t num_ctas_hint);
MPQE_API int mpqe_layer_wgrad(const mpqe_layer_group_t* groups_host, const mpqe_wgrad_operand_t* grads_host,
                     int32_t num_groups, const mpqe_wgrad_dest_t* dests_host, int32_t num_dests,
                     int32_t use_tensor_cores, void* workspace, size_t workspace_bytes, void* stream);

/* Column sums: out[d] (+)= scale * sum over rows of src[r*stride .. +d], r in [0, rows): bias / mode-embedding grads. */
MPQE_API size_t mpqe_colsum_workspace_bytes(int64_t rows);
MPQE_API int mpqe_colsum(const float* src, int64_t rows, int64_t stride, float scale, float* out, int32_t accumulate,
                void* workspace, size_t workspace_bytes, void* stream);

/* batched transpose of [count, d, d] (or one [rows, cols]) matrices */
MPQE_API int mpqe_transpose(const float* src, float* dst, int64_t count, int32_t rows, int32_t cols, void* stream);

/* Sums of [d, d] matrices: dst (+)= src[0] + src[1] + ... (in this order).  With a sum readout every term of the last
 * pass leaves a node slot through the same output row, so the terms of one source slot collapse into ONE term with
 * the matrix  root + sum_{edges e out of the slot} basis[rel[e]]  (model.py:292-304 are linear in the weights); the
 * backward spreads the gradient of such a sum back over its summands with the same entry point. */
#define MPQE_MAX_MATSUM_ITEMS 64
#define MPQE_MAX_MATSUM_SRCS 32
typedef struct {
  float* dst;
  const float* src[MPQE_MAX_MATSUM_SRCS];
  int32_t num_src;
  int32_t accumulate;   /* 0: overwrite dst, 1: add to it */
} mpqe_matsum_item_t;
MPQE_API int mpqe_matrix_sum_multi(const mpqe_matsum_item_t* items_host, int32_t n, void* stream);

/* Basis decomposition of the relation weights (RGCNConv with num_bases > 0, reference model.py:281-284:
 * w = torch.matmul(att, basis.view(num_bases, -1))) and its backward.
 * mpqe_small_k_matmul: out[m, e] = sum_k a[m*a_row_stride + k*a_col_stride] * b[k, e], b [K, E], out [M, E], E % 4 == 0;
 *   W = att @ basis is (a = att, strides (K, 1)); d basis = att^T @ dW is (a = att, strides (1, num_bases), b = dW).
 * mpqe_rows_dot: out[m, k] = sum_e x[m, e] * y[k, e]  (d att = dW . basis).  Fixed summation orders. */
MPQE_API int mpqe_small_k_matmul(const float* a, int64_t a_row_stride, int64_t a_col_stride, const float* b, int32_t M,
                                 int32_t K, int64_t E, float* out, void* stream);
MPQE_API int mpqe_rows_dot(const float* x, const float* y, int32_t M, int32_t K, int64_t E, float* out, void* stream);

/* ---- a11: max readout (model.py:383-385; torch_scatter.scatter_max) -------------------------------------
 * q[b, c] = max_i z[b, i, c]; argmax[b, c] = smallest i attaining it (int64 node ROW b*n+i, like scatter_max). */
MPQE_API int mpqe_max_readout_fwd(const float* z, int64_t B, int32_t n, float* q, int64_t* argmax, void* stream);
/* g[b, i, c] = (argmax[b,c] == b*n+i) ? dq[b,c] : 0 */
MPQE_API int mpqe_max_readout_bwd(const float* dq, const int64_t* argmax, int64_t B, int32_t n, float* g, void* stream);

/* ---- a14/a15: cosine scoring + margin loss (model.py:451-452, 483-485) -----------------------------------
 * y = normalised table row of each id; score = q.y / (max(||q||,1e-8) * max(||y||,1e-8)).
 * ids_pos/ids_neg: [B]; loss[0] = mean(relu(margin - (pos - neg))). */
MPQE_API size_t mpqe_margin_loss_workspace_bytes(int64_t B);
MPQE_API int mpqe_cosine_margin_fwd(const float* q, int64_t B, const float* table, const int64_t* id2row,
                           const int64_t* ids_pos, const int64_t* ids_neg, float margin,
                           float* score_pos, float* score_neg, float* loss,
                           void* workspace, size_t workspace_bytes, void* stream);
/* backward of the above for upstream d(loss) = grad_loss[0] (device scalar):
 * dq[B,d]; rows_out[2B,d] / rows_id[2B] = raw-row gradients of the positive (first B) and negative rows. */
MPQE_API int mpqe_cosine_margin_bwd(const float* q, int64_t B, const float* table, const int64_t* id2row,
                           const int64_t* ids_pos, const int64_t* ids_neg, float margin,
                           const float* grad_loss, float* dq, float* rows_out, int64_t* rows_id, void* stream);
/* eval scoring (model.py:451-460): candidate i belongs to query b with offsets[b] <= i < offsets[b+1]
 * (the reference's repeat_interleave by neg_lengths); offsets == NULL means one candidate per query (count == B).
 * scores[i] = cos(q[b], y(ids[i])). */
MPQE_API int mpqe_cosine_scores(const float* q, int64_t B, const int64_t* offsets, const float* table, const int64_t* id2row,
                       const int64_t* ids, int64_t count, float* scores, void* stream);
/* backward of mpqe_cosine_scores for given d(scores): dq[B,d] (overwritten, or added to when accumulate != 0) sums
 * each query's candidates in ascending order (bit-reproducible); rows_out[count,d] / rows_id[count] as above. */
MPQE_API int mpqe_cosine_scores_bwd(const float* q, int64_t B, const int64_t* offsets, const float* table,
                           const int64_t* id2row, const int64_t* ids, int64_t count, const float* grad_scores,
                           float* dq, int32_t accumulate, float* rows_out, int64_t* rows_id, void* stream);

/* ---- multi-item launches: every formula group of a training step in one kernel ------------------------------
 * Same arithmetic per row as the single-item entry points above; the work items travel in the kernel parameters. */
#define MPQE_MAX_GATHER_ITEMS 32
#define MPQE_MAX_MARGIN_ITEMS 8
#define MPQE_MAX_COLSUM_ITEMS 64

/* One (group, node slot) of the input build (model.py:418-421) or of its backward.
 * forward : out[i*out_stride .. +d] = normalize ? table[row]/||table[row]|| : table[row],
 *           row = id2row ? id2row[ids[i*ids_stride]] : ids[i*ids_stride]   (ids_stride 0 broadcasts one row)
 * backward: rows_out[i] = normalise-backward of grad[i*grad_stride .. +d], rows_id[i] = row + id_offset
 * backward == 2 ("ids"): only rows_id[i] = row + id_offset (input of mpqe_sparse_rows_plan) */
typedef struct {
  const float* table;
  int64_t table_rows;
  const int64_t* id2row;
  const int64_t* ids;
  int64_t ids_stride;
  int64_t count;
  float* out;
  int64_t out_stride;
  const float* grad;
  int64_t grad_stride;
  float* rows_out;
  int64_t* rows_id;
  int64_t id_offset;
  int32_t normalize;
  int32_t reserved;
  /* data-parallel training with row-range-owned tables: peer_tables (DEVICE array of world pointers, entry r = rank r's
   * full-size copy of `table` as mapped into this process) and peer_chunk = ceil(table_rows / world); a normalised row
   * is then read from its owner's copy (peer_tables[row / peer_chunk]).  NULL: read `table`. */
  const float* const* peer_tables;
  int64_t peer_chunk;
  /* optional [count] row norms.  forward: written (||table[row]||).  backward: when non-NULL the table is NOT read
   * again -- the normalised row y is taken from the forward output (`out`, `out_stride`) and the norm from here (same
   * arithmetic, one gather of 512 bytes per row less; with owner-read tables one NVLink round trip less). */
  float* norm;
} mpqe_gather_item_t;
MPQE_API int mpqe_gather_multi(const mpqe_gather_item_t* items_host, int32_t n, int32_t backward, void* stream);

/* One formula group of the margin loss (model.py:451-452, 483-485); fields as mpqe_cosine_margin_{fwd,bwd}.
 * hinge is a [B] scratch array; id_offset is added to the emitted table row ids.
 * backward: 0 = forward (hinge, loss), 1 = backward (dq, rows), 2 = both in one pass over q and the table rows --
 * usable when d total / d loss (grad_loss) is known before the forward, as in a training step. */
typedef struct {
  const float* q;
  int64_t B;
  const float* table;
  const int64_t* id2row;
  const int64_t* ids_pos;
  const int64_t* ids_neg;
  float* score_pos; /* optional */
  float* score_neg; /* optional */
  float* hinge;
  float* loss;
  const float* grad_loss;
  float* dq;
  float* rows_out;
  int64_t* rows_id;
  int64_t id_offset;
  const float* const* peer_tables;   /* as in mpqe_gather_item_t */
  int64_t peer_chunk;
} mpqe_margin_item_t;
MPQE_API int mpqe_cosine_margin_multi(const mpqe_margin_item_t* items_host, int32_t n, float margin, int32_t backward,
                                      void* stream);

/* dst[d] += scale * sum over rows of src[r*stride .. +d], items applied in order (bit-reproducible). */
typedef struct {
  const float* src;
  int64_t rows;
  int64_t stride;
  float* dst;
  float scale;
  int32_t reserved;
} mpqe_colsum_item_t;
MPQE_API size_t mpqe_colsum_multi_workspace_bytes(const mpqe_colsum_item_t* items_host, int32_t n);
MPQE_API int mpqe_colsum_multi(const mpqe_colsum_item_t* items_host, int32_t n, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ---- a17: rank counts (utils.py:25-32; scipy percentileofscore kind='rank') ------------------------------
 * ragged negatives: left[b] = #(neg < pos[b]), right[b] = #(neg <= pos[b]) over neg[offsets[b]:offsets[b+1]]. */
MPQE_API int mpqe_rank_counts_ragged(const float* pos, const float* neg, const int64_t* offsets, int64_t B,
                            int64_t* left, int64_t* right, void* stream);
/* ROC AUC pair counts (utils.py:34-36: roc_auc_score(labels, nan_to_num(predictions))): counts[0] += #(neg < pos),
 * counts[1] += #(neg == pos) over all (positive, negative) score pairs after nan_to_num; counts is ACCUMULATED (zero it
 * first).  AUC = (counts[0] + counts[1] / 2) / (num_pos * num_neg), the Mann-Whitney form of the same statistic. */
MPQE_API int mpqe_auc_counts(const float* pos, int64_t num_pos, const float* neg, int64_t num_neg,
                             unsigned long long* counts, void* stream);
/* full-entity ranking against one table shard (new; north star "full-entity ranking eval"):
 * candidates = rows [row_begin, row_end) of `table`, score = cos(q[b], y(row)) as above;
 * left/right are ACCUMULATED (zero them first) so shards can be chained or merged with an integer allreduce. */
MPQE_API size_t mpqe_rank_counts_table_workspace_bytes(int64_t B, int64_t rows);
MPQE_API int mpqe_rank_counts_table(const float* q, int64_t B, const float* pos, const float* table,
                           int64_t row_begin, int64_t row_end, int64_t* left, int64_t* right,
                           void* workspace, size_t workspace_bytes, int32_t use_tensor_cores, void* stream);
/* The same in two phases, for an evaluation that ranks many batches against one table shard: `prepare` computes the
 * table-dependent half once (1/||row|| and, for the tensor-core kernel, the pre-split tile images of the shard) into
 * `table_ws`; `mpqe_rank_counts_prepared` then does only the per-batch work.  Prepare again when the table changes. */
MPQE_API size_t mpqe_rank_table_workspace_bytes(int64_t rows, int32_t use_tensor_cores);
MPQE_API int mpqe_rank_table_prepare(const float* table, int64_t row_begin, int64_t row_end, void* table_ws,
                            size_t table_ws_bytes, int32_t use_tensor_cores, void* stream);
MPQE_API size_t mpqe_rank_query_workspace_bytes(int64_t B);
MPQE_API int mpqe_rank_counts_prepared(const float* q, int64_t B, const float* pos, const float* table,
                              int64_t row_begin, int64_t row_end, const void* table_ws, int64_t* left, int64_t* right,
                              void* query_ws, size_t query_ws_bytes, int32_t use_tensor_cores, void* stream);

/* ---- a16: row-sparse embedding gradients ----------------------------------------------------------------
 * Combine `count` (row id, gradient row) pairs: unique_ids ascending, rows summed in ascending pair order
 * (bit-reproducible).  num_unique is a device int64 scalar; outputs are sized for `count` rows and entries past
 * num_unique are zero rows with id `pad_id`.  Input ids >= table_rows are padding of an earlier combine (use
 * pad_id = table_rows when results are combined again, e.g. after an all-gather) and are dropped. */
MPQE_API size_t mpqe_sparse_rows_workspace_bytes(int64_t count);
MPQE_API int mpqe_sparse_rows_combine(const int64_t* rows_id, const float* rows, int64_t count, int64_t table_rows,
                             int64_t pad_id, int64_t* unique_ids, float* unique_rows, int64_t* num_unique,
                             void* workspace, size_t workspace_bytes, void* stream);
/* The same combine in two phases sharing one workspace.  `plan` needs only the row ids (stable sort, segment heads,
 * num_unique) -- the ids of a training step are known before its backward has produced any gradient row, so the
 * plan can run on another stream meanwhile; `apply` then sums the rows (one kernel) and multiplies every sum by
 * `scale` (1/world_size of a data-parallel average).  plan + apply(scale = 1) == combine. */
MPQE_API int mpqe_sparse_rows_plan(const int64_t* rows_id, int64_t count, int64_t table_rows, int64_t* num_unique,
                          void* workspace, size_t workspace_bytes, void* stream);
MPQE_API int mpqe_sparse_rows_apply(const float* rows, int64_t count, int64_t table_rows, int64_t pad_id, float scale,
                           int64_t* unique_ids, float* unique_rows, const int64_t* num_unique,
                           void* workspace, size_t workspace_bytes, void* stream);
/* Data-parallel form of `apply`: the pairs of `world` ranks (plan built over the rank-major concatenation of their
 * ids, `per_rank_count` each) are summed straight out of the ranks' own row buffers -- peer_rows_host[r] is rank r's
 * buffer as mapped into THIS process (peer memory over NVLink / NVSwitch).  Gather and combine are one kernel: a
 * remote row crosses NVLink once and is never staged locally (replaces an NCCL all-gather of the rows + apply).
 * unique_ids / unique_rows hold `out_capacity` entries (<= world * per_rank_count; with an owner plan the number of
 * rows this rank owns bounds the distinct rows it can receive): entries in [*num_unique, out_capacity) are padding. */
#define MPQE_MAX_PEERS 16
MPQE_API int mpqe_sparse_rows_apply_peers(const float* const* peer_rows_host, int32_t world, int64_t per_rank_count,
                                 int64_t table_rows, int64_t pad_id, float scale, int64_t* unique_ids,
                                 float* unique_rows, int64_t out_capacity, const int64_t* num_unique, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* ---- data-parallel exchange over peer-mapped memory (new capability; the reference is single-process) ------------
 * `peer_*_host[r]` is rank r's buffer as mapped into THIS process (e.g. torch symmetric memory); all ranks call the
 * same sequence of these entry points on their streams. */
/* flag barrier: peer_flags_host[r] = rank r's int32[MPQE_MAX_PEERS] flag array (zero-initialised), `epoch` a
 * zero-initialised device int32 private to the rank.  Everything the stream did before the barrier is visible to the
 * peers after it.  Graph-capturable (the epoch lives on the device). */
MPQE_API int mpqe_peer_barrier(const void* const* peer_flags_host, int32_t rank, int32_t world, int32_t* epoch,
                      void* stream);
/* out[i] = scale * sum_r peer_bufs[r][i] in rank order (identical bits on every rank); numel % 4 == 0 */
MPQE_API int mpqe_allreduce_peers(const void* const* peer_bufs_host, int32_t world, int64_t numel, float scale, float* out,
                         void* stream);
/* The same all-reduce in two shots for larger groups: `reduce_scatter` sums float4 slice `rank` (slices of
 * ceil(numel/4/world) float4) of all ranks' buffers, in rank order, IN PLACE into this rank's buffer; after a barrier
 * `all_gather` reads every slice from its owner into `out`.  Per rank 2 (N-1)/N buckets cross NVLink instead of N-1. */
MPQE_API int mpqe_reduce_scatter_peers(const void* const* peer_bufs_host, int32_t world, int32_t rank, int64_t numel,
                              float scale, void* stream);
MPQE_API int mpqe_all_gather_peers(const void* const* peer_bufs_host, int32_t world, int64_t numel, float* out,
                          void* stream);
/* Owner-side plan of the row-gradient exchange.  The entity tables are owned row-range-wise: rank r owns rows
 * [r*ceil(rows_t/world), (r+1)*ceil(rows_t/world)) of every table t (global row id = table_begin[t] + row).  The ids
 * every rank emitted for its step (per_rank_count each, mpqe_gather_multi mode 2) are read in place; ids this rank
 * does not own are dropped; the rest is planned exactly like mpqe_sparse_rows_plan over the rank-major concatenation,
 * so that mpqe_sparse_rows_apply_peers then sums the owned rows out of the peers' row buffers. */
MPQE_API int mpqe_sparse_rows_plan_owner(const void* const* peer_ids_host, int32_t world, int32_t rank,
                                int64_t per_rank_count, const int64_t* table_begin_host,
                                const int64_t* table_rows_host, int32_t num_tables, int64_t total_rows,
                                int64_t* num_unique, void* workspace, size_t workspace_bytes, void* stream);
/* dense[ids[i], :] (+)= rows[i, :] for i < *num (ids unique) */
MPQE_API int mpqe_scatter_rows(const int64_t* ids, const float* rows, const int64_t* num, int64_t max_count,
                      float* dense, int32_t accumulate, void* stream);

/* ---- optimiser (train.py:86-88, torch.optim.Adam defaults; caller of the hot path, row (f)) -------------- */
MPQE_API int mpqe_adam_dense(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                    float lr, float beta1, float beta2, float eps, int32_t step, void* stream);

/* ---- L2 term of margin_loss over the readout-MLP parameters (model.py:487-492) -----------------------------------
 * loss += weight_decay * sum_i ||param_i||_2  (un-squared norms, one term per margin_loss call).  For every item
 * grad_i += grad_scale * weight_decay * param_i / ||param_i|| (grad_scale = sum of d total / d loss_j over the step's
 * batches; `grad` is laid out like `param`), every losses[j < num_losses] += weight_decay * sum_i ||param_i||, and
 * norms[i] (optional) receives ||param_i||.  One CTA, fixed summation order: bit-reproducible. */
#define MPQE_MAX_L2_ITEMS 8
typedef struct {
  const float* param;
  float* grad;     /* may be NULL (loss only) */
  int64_t numel;
} mpqe_l2_item_t;
MPQE_API int mpqe_l2_reg_multi(const mpqe_l2_item_t* items_host, int32_t n, float weight_decay, float grad_scale,
                      float* losses, int32_t num_losses, float* norms, void* stream);

/* Adam over up to MPQE_MAX_ADAM_ITEMS dense tensors in one launch (torch/optim/adam.py single-tensor formula:
 * exp_avg.lerp_, exp_avg_sq.mul_.addcmul_, bias corrections in double, param.addcdiv_). */
#define MPQE_MAX_ADAM_ITEMS 32
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} mpqe_adam_item_t;
/* Device-resident optimiser clock, so that a captured CUDA graph can be replayed step after step: `mpqe_adam_tick`
 * (one thread) increments `step` and recomputes the bias corrections; the Adam entry points below read them from
 * `state` when it is non-NULL (their `step` argument is then ignored). */
typedef struct {
  int32_t step;        /* optimiser steps completed (after a tick: the step being applied) */
  float step_size;     /* lr / (1 - beta1^step) */
  float bc2_sqrt;      /* sqrt(1 - beta2^step) */
  int32_t reserved;
} mpqe_adam_state_t;
MPQE_API int mpqe_adam_tick(mpqe_adam_state_t* state, float lr, float beta1, float beta2, void* stream);
MPQE_API int mpqe_adam_multi(const mpqe_adam_item_t* items_host, int32_t n, float lr, float beta1, float beta2, float eps,
                    int32_t step, const mpqe_adam_state_t* state, void* stream);

/* Adam over the entity tables from the row-sparse gradient of a step, trajectory-equivalent to the reference's DENSE
 * Adam (train.py:86-88): a row that a step does not touch still moves through its momentum under dense Adam; those
 * zero-gradient steps are applied lazily by `catchup` -- called with the ids a step is about to read, BEFORE its
 * forward pass, with upto_step = step - 1 -- and `apply` then performs step `step` on the combined (unique) gradient
 * rows.  `last_step` is an int32 array over the global row id space (table t owns ids [row_begin, row_begin+rows)),
 * zero-initialised.  catchup with ids == NULL brings rows [0, count) of the id space up to date (before an
 * evaluation, a checkpoint or an export).  Duplicate ids are fine in catchup (claimed once).  With a device `state`
 * catchup runs up to state->step (call it BEFORE the step's tick) and apply performs step state->step (after it). */
#define MPQE_MAX_TABLES 16
typedef struct {
  float* table;       /* [rows, d] */
  float* exp_avg;     /* [rows, d] */
  float* exp_avg_sq;  /* [rows, d] */
  int64_t row_begin;  /* first global row id of this table */
  int64_t rows;
} mpqe_adam_table_t;
MPQE_API int mpqe_adam_rows_catchup(const mpqe_adam_table_t* tables_host, int32_t num_tables, const int64_t* ids,
                           int64_t count, int32_t upto_step, float lr, float beta1, float beta2, float eps,
                           const mpqe_adam_state_t* state, int32_t* last_step, void* stream);
MPQE_API int mpqe_adam_rows_apply(const mpqe_adam_table_t* tables_host, int32_t num_tables, const int64_t* ids,
                         const float* rows, const int64_t* num, int64_t max_count, int32_t step, float lr, float beta1,
                         float beta2, float eps, const mpqe_adam_state_t* state, int32_t* last_step, void* stream);

/* ---- negative sampling on the device (model.py:470-476: random.choice over each query's stored negatives) --------
 * out[i] = candidates[offsets[q] + r % (offsets[q+1] - offsets[q])] with q = query_index ? query_index[i]
 * : (first_query + i) % num_queries_total (the reference's contiguous batch slice with wrap-around,
 * data_utils.py:300-308) and r = splitmix64(seed, step, i): a counter-based draw, reproducible for a given
 * (seed, step) whatever the launch geometry.  offsets == NULL: all queries share candidates[0 .. shared_count)
 * (1-chain negatives are drawn from the whole target mode, model.py:472-473).  A query without candidates gets -1. */
MPQE_API int mpqe_sample_negatives(const int64_t* candidates, const int64_t* offsets, const int64_t* query_index,
                          int64_t first_query, int64_t num_queries_total, int64_t shared_count, int64_t count,
                          uint64_t seed, uint64_t step, int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPQE_B200_H_ */

"""ctypes binding of the C-ABI library (`include/mpqe_b200.h`).  No CPU fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MPQE_LIB_PATH') or os.path.join(HERE, '_C', 'libmpqe_b200.so')

MAX_GROUPS, MAX_TERMS, MAX_SLOTS, MAX_DESTS, D = 8, 16, 8, 64, 128
EPI_NONE, EPI_RELU, EPI_MASK = 0, 1, 2


class Term(C.Structure):
    _fields_ = [('a', C.c_void_p), ('m', C.c_void_p), ('m_packed', C.c_void_p), ('a_slots', C.c_int32), ('a_slot', C.c_int16),
                ('out_slot', C.c_int16)]


class LayerGroup(C.Structure):
    _fields_ = [('num_queries', C.c_int64), ('num_terms', C.c_int32), ('num_out_slots', C.c_int32),
                ('terms', Term * MAX_TERMS), ('out', C.c_void_p), ('out_slots', C.c_int32),
                ('epilogue', C.c_int32), ('bias', C.c_void_p), ('bias_scale', C.c_float * MAX_SLOTS),
                ('out_slot_map', C.c_int16 * MAX_SLOTS), ('mask', C.c_void_p), ('mask_slots', C.c_int32),
                ('bias_slot_stride', C.c_int32), ('relu_bits_out', C.c_void_p), ('mask_bits', C.c_void_p)]


class WgradDest(C.Structure):
    _fields_ = [('m_fwd', C.c_void_p), ('dm', C.c_void_p), ('accumulate', C.c_int32), ('reserved', C.c_int32)]


class WgradOperand(C.Structure):
    _fields_ = [('g', C.c_void_p), ('g_slots', C.c_int32), ('slot_map', C.c_int16 * MAX_SLOTS)]


class GatherItem(C.Structure):
    _fields_ = [('table', C.c_void_p), ('table_rows', C.c_int64), ('id2row', C.c_void_p), ('ids', C.c_void_p),
                ('ids_stride', C.c_int64), ('count', C.c_int64), ('out', C.c_void_p), ('out_stride', C.c_int64),
                ('grad', C.c_void_p), ('grad_stride', C.c_int64), ('rows_out', C.c_void_p), ('rows_id', C.c_void_p),
                ('id_offset', C.c_int64), ('normalize', C.c_int32), ('reserved', C.c_int32),
                ('peer_tables', C.c_void_p), ('peer_chunk', C.c_int64), ('norm', C.c_void_p)]


class MarginItem(C.Structure):
    _fields_ = [('q', C.c_void_p), ('B', C.c_int64), ('table', C.c_void_p), ('id2row', C.c_void_p),
                ('ids_pos', C.c_void_p), ('ids_neg', C.c_void_p), ('score_pos', C.c_void_p), ('score_neg', C.c_void_p),
                ('hinge', C.c_void_p), ('loss', C.c_void_p), ('grad_loss', C.c_void_p), ('dq', C.c_void_p),
                ('rows_out', C.c_void_p), ('rows_id', C.c_void_p), ('id_offset', C.c_int64),
                ('peer_tables', C.c_void_p), ('peer_chunk', C.c_int64)]


class ColsumItem(C.Structure):
    _fields_ = [('src', C.c_void_p), ('rows', C.c_int64), ('stride', C.c_int64), ('dst', C.c_void_p),
                ('scale', C.c_float), ('reserved', C.c_int32)]


MAX_GATHER_ITEMS, MAX_MARGIN_ITEMS, MAX_COLSUM_ITEMS = 32, 8, 64
MAX_MATSUM_ITEMS, MAX_MATSUM_SRCS = 64, 32


class MatsumItem(C.Structure):
    _fields_ = [('dst', C.c_void_p), ('src', C.c_void_p * MAX_MATSUM_SRCS), ('num_src', C.c_int32),
                ('accumulate', C.c_int32)]


MAX_L2_ITEMS, MAX_ADAM_ITEMS, MAX_TABLES, MAX_PEERS = 8, 32, 16, 16


class L2Item(C.Structure):
    _fields_ = [('param', C.c_void_p), ('grad', C.c_void_p), ('numel', C.c_int64)]


class AdamItem(C.Structure):
    _fields_ = [('param', C.c_void_p), ('grad', C.c_void_p), ('exp_avg', C.c_void_p), ('exp_avg_sq', C.c_void_p),
                ('numel', C.c_int64)]


class AdamTable(C.Structure):
    _fields_ = [('table', C.c_void_p), ('exp_avg', C.c_void_p), ('exp_avg_sq', C.c_void_p), ('row_begin', C.c_int64),
                ('rows', C.c_int64)]


ABI_STRUCTS = (Term, LayerGroup, WgradDest, WgradOperand, GatherItem, MarginItem, ColsumItem, MatsumItem, L2Item,
               AdamItem, AdamTable)

P, I32, I64, F32, SZ, U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); mirrors include/mpqe_b200.h one to one (tests/test_abi.py checks both directions)
SIGNATURES = {
    'mpqe_b200_last_error': (C.c_char_p, []),
    'mpqe_b200_version': (I32, []),
    'mpqe_b200_has_tcgen05': (I32, []),
    'mpqe_b200_sizeof': (I32, [I32]),
    'mpqe_build_query_graph': (I32, [I32, I32, P, P, P, I64, P, P, P, P]),
    'mpqe_relation_sort_workspace_bytes': (SZ, [I64, I32]),
    'mpqe_relation_sort': (I32, [P, I64, I32, P, P, P, SZ, P]),
    'mpqe_gather_normalize_fwd': (I32, [P, I64, P, P, I64, I64, P, I64, P, P]),
    'mpqe_gather_normalize_bwd': (I32, [P, P, P, I64, I64, P, I64, P, P, P]),
    'mpqe_broadcast_rows': (I32, [P, P, I32, P, I64, I64, P]),
    'mpqe_layer_forward': (I32, [P, I32, I32, P]),
    'mpqe_layer_wgrad_workspace_bytes': (SZ, [I32, I32]),
    'mpqe_layer_wgrad': (I32, [P, P, I32, P, I32, I32, P, SZ, P]),
    'mpqe_colsum_workspace_bytes': (SZ, [I64]),
    'mpqe_colsum': (I32, [P, I64, I64, F32, P, I32, P, SZ, P]),
    'mpqe_transpose': (I32, [P, P, I64, I32, I32, P]),
    'mpqe_matrix_sum_multi': (I32, [P, I32, P]),
    'mpqe_small_k_matmul': (I32, [P, I64, I64, P, I32, I32, I64, P, P]),
    'mpqe_rows_dot': (I32, [P, P, I32, I32, I64, P, P]),
    'mpqe_max_readout_fwd': (I32, [P, I64, I32, P, P, P]),
    'mpqe_max_readout_bwd': (I32, [P, P, I64, I32, P, P]),
    'mpqe_margin_loss_workspace_bytes': (SZ, [I64]),
    'mpqe_cosine_margin_fwd': (I32, [P, I64, P, P, P, P, F32, P, P, P, P, SZ, P]),
    'mpqe_cosine_margin_bwd': (I32, [P, I64, P, P, P, P, F32, P, P, P, P, P]),
    'mpqe_cosine_scores': (I32, [P, I64, P, P, P, P, I64, P, P]),
    'mpqe_cosine_scores_bwd': (I32, [P, I64, P, P, P, P, I64, P, P, I32, P, P, P]),
    'mpqe_rank_counts_ragged': (I32, [P, P, P, I64, P, P, P]),
    'mpqe_auc_counts': (I32, [P, I64, P, I64, P, P]),
    'mpqe_rank_counts_table_workspace_bytes': (SZ, [I64, I64]),
    'mpqe_rank_counts_table': (I32, [P, I64, P, P, I64, I64, P, P, P, SZ, I32, P]),
    'mpqe_rank_table_workspace_bytes': (SZ, [I64, I32]),
    'mpqe_rank_table_prepare': (I32, [P, I64, I64, P, SZ, I32, P]),
    'mpqe_rank_query_workspace_bytes': (SZ, [I64]),
    'mpqe_rank_counts_prepared': (I32, [P, I64, P, P, I64, I64, P, P, P, P, SZ, I32, P]),
    'mpqe_sparse_rows_workspace_bytes': (SZ, [I64]),
    'mpqe_sparse_rows_combine': (I32, [P, P, I64, I64, I64, P, P, P, P, SZ, P]),
    'mpqe_sparse_rows_plan': (I32, [P, I64, I64, P, P, SZ, P]),
    'mpqe_sparse_rows_apply': (I32, [P, I64, I64, I64, F32, P, P, P, P, SZ, P]),
    'mpqe_sparse_rows_apply_peers': (I32, [P, I32, I64, I64, I64, F32, P, P, I64, P, P, SZ, P]),
    'mpqe_peer_barrier': (I32, [P, I32, I32, P, P]),
    'mpqe_allreduce_peers': (I32, [P, I32, I64, F32, P, P]),
    'mpqe_reduce_scatter_peers': (I32, [P, I32, I32, I64, F32, P]),
    'mpqe_all_gather_peers': (I32, [P, I32, I64, P, P]),
    'mpqe_sparse_rows_plan_owner': (I32, [P, I32, I32, I64, P, P, I32, I64, P, P, SZ, P]),
    'mpqe_scatter_rows': (I32, [P, P, P, I64, P, I32, P]),
    'mpqe_adam_dense': (I32, [P, P, P, P, I64, F32, F32, F32, F32, I32, P]),
    'mpqe_pack_weights': (I32, [P, I32, P, P]),
    'mpqe_pack_weights_ex': (I32, [P, P, I32, P, P]),
    'mpqe_gather_multi': (I32, [P, I32, I32, P]),
    'mpqe_cosine_margin_multi': (I32, [P, I32, F32, I32, P]),
    'mpqe_colsum_multi_workspace_bytes': (SZ, [P, I32]),
    'mpqe_colsum_multi': (I32, [P, I32, P, SZ, P]),
    'mpqe_l2_reg_multi': (I32, [P, I32, F32, F32, P, I32, P, P]),
    'mpqe_adam_tick': (I32, [P, F32, F32, F32, P]),
    'mpqe_adam_multi': (I32, [P, I32, F32, F32, F32, F32, I32, P, P]),
    'mpqe_adam_rows_catchup': (I32, [P, I32, P, I64, I32, F32, F32, F32, F32, P, P, P]),
    'mpqe_adam_rows_apply': (I32, [P, I32, P, P, P, I64, I32, F32, F32, F32, F32, P, P, P]),
    'mpqe_sample_negatives': (I32, [P, P, P, I64, I64, I64, I64, U64, U64, P, P]),
}

_lib = None


class MpqeError(RuntimeError):
    pass


def load():
    """Load `libmpqe_b200.so` (built in-tree by `python -m mpqe_b200.build`).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MpqeError('CUDA extension %s is missing: run `python -m mpqe_b200.build` (there is no CPU fallback)'
                        % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, struct in enumerate(ABI_STRUCTS):
        if lib.mpqe_b200_sizeof(which) != C.sizeof(struct):
            raise MpqeError('ABI mismatch: %s is %d bytes in python, %d in the library'
                            % (struct.__name__, C.sizeof(struct), lib.mpqe_b200_sizeof(which)))
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        raise MpqeError('%s failed (%d): %s' % (what, status, load().mpqe_b200_last_error().decode()))

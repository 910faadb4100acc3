"""Tensor-level wrappers over the C ABI (`include/mpqe_b200.h`).

PyTorch is used here only for device memory (caching allocator) and stream identity; every computation is a
hand-written CUDA kernel behind `libmpqe_b200.so`.  There is no CPU or eager fallback: tensors must live on a
CUDA device and the library must load, otherwise an exception is raised.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import D, EPI_MASK, EPI_NONE, EPI_RELU, MAX_DESTS, MAX_GROUPS, MAX_SLOTS, MAX_TERMS  # noqa: F401

# use_tensor_cores: None -> tcgen05 path when the library has it and MPQE_TENSOR_CORES != 0
_tc_default = None
launch_count = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n=1):
    global launch_count
    launch_count += n


def tensor_cores_default():
    global _tc_default
    if _tc_default is None:
        import os
        want = os.environ.get('MPQE_TENSOR_CORES', '1') != '0'
        _tc_default = bool(want and _lib.load().mpqe_b200_has_tcgen05())
    return _tc_default


def set_tensor_cores(flag):
    global _tc_default
    if flag and not _lib.load().mpqe_b200_has_tcgen05():
        raise _lib.MpqeError('library was built without the tcgen05 kernels')
    _tc_default = bool(flag)


def device_guard(device):
    """Context manager selecting `device`; the single place that enforces "CUDA only, no CPU fallback"."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise _lib.MpqeError('mpqe_b200 runs on a CUDA device only (there is no CPU fallback); got %s' % device)
    return torch.cuda.device(device)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise _lib.MpqeError('%s must be a CUDA tensor (no CPU fallback)' % name)
    if t.dtype != dtype:
        raise _lib.MpqeError('%s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.MpqeError('%s must be contiguous' % name)
    return t


_ws = {}


def workspace(nbytes, device, tag='ws'):
    """Grow-only scratch buffer per (device, tag); kernels on one stream are ordered so reuse is safe."""
    key = (device.index, tag, torch.cuda.current_stream().cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


# ---------------------------------------------------------------------------------------------------------------
# Layer term lists
# ---------------------------------------------------------------------------------------------------------------
class Term(object):
    """out[q, out_slot] += A[q, a_slot] @ M.  `a` is a [B, a_slots, D] tensor (a_slots = 0: one broadcast row set
    [rows, D], a_slot selects the row), `m` a contiguous [D, D] tensor (view)."""
    __slots__ = ('a', 'a_slots', 'a_slot', 'm', 'out_slot', 'mp')

    def __init__(self, a, a_slots, a_slot, m, out_slot, mp=None):
        self.a, self.a_slots, self.a_slot, self.m, self.out_slot = a, int(a_slots), int(a_slot), m, int(out_slot)
        self.mp = mp   # optional pack_weights() image of m


class Group(object):
    """One formula group of a layer launch (see mpqe_layer_group_t)."""

    def __init__(self, num_queries, terms, num_out_slots, out, out_slots, out_slot_map=None, epilogue=EPI_NONE,
                 bias=None, bias_scale=None, mask=None, mask_slots=0, bias_slot_stride=0, bits_out=None, mask_bits=None):
        self.num_queries, self.terms, self.num_out_slots = int(num_queries), list(terms), int(num_out_slots)
        self.out, self.out_slots = out, int(out_slots)
        self.out_slot_map = list(out_slot_map) if out_slot_map is not None else list(range(num_out_slots))
        self.epilogue, self.bias = epilogue, bias
        self.bias_scale = list(bias_scale) if bias_scale is not None else [1.0] * num_out_slots
        self.mask, self.mask_slots = mask, int(mask_slots)
        self.bias_slot_stride = int(bias_slot_stride)   # 0: one bias vector; D: bias is [num_out_slots, D]
        # ReLU sign bits (int32 [ceil(B/32), slots, D]): written by an EPI_RELU launch / read by an EPI_MASK launch
        self.bits_out, self.mask_bits = bits_out, mask_bits

    def to_c(self):
        if len(self.terms) > MAX_TERMS or self.num_out_slots > MAX_SLOTS:
            raise _lib.MpqeError('group exceeds MPQE_MAX_TERMS / MPQE_MAX_SLOTS')
        g = _lib.LayerGroup()
        g.num_queries, g.num_terms, g.num_out_slots = self.num_queries, len(self.terms), self.num_out_slots
        for i, t in enumerate(self.terms):
            _chk(t.a, torch.float32, 'term.a')
            _chk(t.m, torch.float32, 'term.m')
            if t.m.numel() != D * D:
                raise _lib.MpqeError('term matrix must be [%d,%d]' % (D, D))
            g.terms[i].a, g.terms[i].m = t.a.data_ptr(), t.m.data_ptr()
            g.terms[i].m_packed = t.mp.data_ptr() if t.mp is not None else 0
            g.terms[i].a_slots, g.terms[i].a_slot, g.terms[i].out_slot = t.a_slots, t.a_slot, t.out_slot
        g.out = _chk(self.out, torch.float32, 'out').data_ptr() if self.out is not None else 0
        g.out_slots, g.epilogue = self.out_slots, self.epilogue
        g.bias = _chk(self.bias, torch.float32, 'bias').data_ptr() if self.bias is not None else 0
        for j in range(self.num_out_slots):
            g.bias_scale[j] = float(self.bias_scale[j])
            g.out_slot_map[j] = int(self.out_slot_map[j])
        g.mask = _chk(self.mask, torch.float32, 'mask').data_ptr() if self.mask is not None else 0
        g.mask_slots = self.mask_slots
        g.bias_slot_stride = self.bias_slot_stride
        g.relu_bits_out = _chk(self.bits_out, torch.int32, 'bits_out').data_ptr() if self.bits_out is not None else 0
        g.mask_bits = _chk(self.mask_bits, torch.int32, 'mask_bits').data_ptr() if self.mask_bits is not None else 0
        return g


profile = None  # bench.py sets this to a list: (kind, start event, end event, algorithmic bytes, flops) per launch


def _algorithmic(groups, grad_operands=None):
    """Algorithmic bytes / flops of a term-list launch: every distinct input row and every output row once."""
    nbytes = flops = 0
    for i, g in enumerate(groups):
        rows = len({(t.a.data_ptr(), t.a_slot) for t in g.terms if t.a_slots > 0})
        if grad_operands is None:
            rows += g.num_out_slots * (2 if g.epilogue == EPI_MASK else 1)
        else:
            gt, g_slots, smap = grad_operands[i]
            rows += len({smap[t.out_slot] for t in g.terms})
        nbytes += g.num_queries * rows * D * 4
        flops += g.num_queries * len(g.terms) * 2 * D * D
    return nbytes, flops


class _Profiled(object):
    def __init__(self, kind, groups, grad_operands=None):
        self.on = profile is not None
        if self.on:
            self.kind = kind
            self.cost = _algorithmic(groups, grad_operands)
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        if self.on:
            # keep the GPU busy (~150 us spin) while the host prepares and enqueues [start, kernel, end]; otherwise
            # an idle GPU stamps `start` long before the launch arrives and the interval includes host latency
            torch.cuda._sleep(300000)
            self.s.record()

    def __exit__(self, *exc):
        if self.on:
            self.e.record()
            profile.append((self.kind, self.s, self.e, self.cost[0], self.cost[1]))


def layer_forward(groups, use_tensor_cores=None):
    """mpqe_layer_forward over up to MPQE_MAX_GROUPS groups per launch."""
    lib = _lib.load()
    tc = tensor_cores_default() if use_tensor_cores is None else bool(use_tensor_cores)
    for i in range(0, len(groups), MAX_GROUPS):
        chunk = groups[i:i + MAX_GROUPS]
        rows_only = all(g.num_queries == 1 for g in chunk)     # batch-constant rows: the one-row kernel
        if tc and not rows_only:
            _ensure_packed(chunk)
        arr = (_lib.LayerGroup * len(chunk))(*[g.to_c() for g in chunk])
        with _Profiled('layer_rows' if rows_only else 'layer', chunk):
            _lib.check(lib.mpqe_layer_forward(arr, len(chunk), int(tc), _stream()), 'mpqe_layer_forward')
        _count()


def _ensure_packed(groups):
    """The tcgen05 layer kernel stages every weight matrix as a pre-split tile image: matrices that reach a launch
    without one (stand-alone RGCNConv calls; the per-step weights arrive packed from `model.Weights`) are packed here,
    all in one launch."""
    missing = {}
    for g in groups:
        for t in g.terms:
            if t.mp is None:
                missing.setdefault(t.m.data_ptr(), t.m)
    if not missing:
        return
    keys = list(missing)
    packed = pack_weights([missing[k] for k in keys])
    image = {k: packed[j] for j, k in enumerate(keys)}
    for g in groups:
        for t in g.terms:
            if t.mp is None:
                t.mp = image[t.m.data_ptr()]


def relu_bits(num_queries, slots, device):
    """Buffer for the ReLU sign bits of a [num_queries, slots, D] activation (see mpqe_layer_group_t.relu_bits_out)."""
    return torch.empty((num_queries + 31) // 32, slots, D, dtype=torch.int32, device=device)


def layer_wgrad(groups, grad_operands, dests, ctas_hint=296, use_tensor_cores=None):
    """mpqe_layer_wgrad.  grad_operands[i] = (g tensor, g_slots, slot_map list) for groups[i];
    dests = [(m_fwd tensor view, dm tensor view, accumulate)]."""
    lib = _lib.load()
    if len(groups) > MAX_GROUPS or len(dests) > MAX_DESTS:
        raise _lib.MpqeError('too many groups / destinations for one wgrad launch')
    garr = (_lib.LayerGroup * len(groups))(*[g.to_c() for g in groups])
    oarr = (_lib.WgradOperand * len(groups))()
    for i, (g, g_slots, slot_map) in enumerate(grad_operands):
        oarr[i].g, oarr[i].g_slots = _chk(g, torch.float32, 'grad operand').data_ptr(), int(g_slots)
        for j, s in enumerate(slot_map):
            oarr[i].slot_map[j] = int(s)
    darr = (_lib.WgradDest * len(dests))()
    for j, (m_fwd, dm, acc) in enumerate(dests):
        darr[j].m_fwd, darr[j].dm, darr[j].accumulate = m_fwd.data_ptr(), _chk(dm, torch.float32, 'dm').data_ptr(), int(acc)
    dev = groups[0].terms[0].a.device
    tc = tensor_cores_default() if use_tensor_cores is None else bool(use_tensor_cores)
    if tc:
        import os
        # persistent tcgen05 kernel: the workspace holds ctas_hint + destinations partial sums = the number of
        # (destination, chunk) units the launch is split into: one per SM (with 148 + destinations units a few CTAs
        # ran two, twice as long as the rest: SMs active 56 % of the kernel's duration)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        ctas_hint = int(os.environ.get('MPQE_WGRAD_UNITS', str(max(sms - len(dests), sms // 2))))
    nbytes = lib.mpqe_layer_wgrad_workspace_bytes(len(dests), ctas_hint)
    ws = workspace(nbytes, dev, 'wgrad')
    rows_only = all(g.num_queries == 1 for g in groups)
    with _Profiled('wgrad_rows' if rows_only else 'wgrad', groups, grad_operands):
        # the chunking (hence the summation order) follows the workspace size: pass the requested size, not the
        # size of the cached buffer, so that results do not depend on what ran before
        _lib.check(lib.mpqe_layer_wgrad(garr, oarr, len(groups), darr, len(dests), int(tc), _ptr(ws), nbytes,
                                        _stream()), 'mpqe_layer_wgrad')
    _count(1 if rows_only else 2)


def small_k_matmul(a, b, transpose_a=False):
    """out = a @ b (or a^T @ b) for a small inner dimension: a [M, K] ([K, M] with transpose_a), b [K, ...] -> [M, ...]
    (mpqe_small_k_matmul: the basis decomposition of the relation weights and its d basis)."""
    lib = _lib.load()
    a, b = _chk(a, torch.float32, 'a'), _chk(b, torch.float32, 'b')
    M, K = (a.shape[1], a.shape[0]) if transpose_a else (a.shape[0], a.shape[1])
    if b.shape[0] != K:
        raise _lib.MpqeError('small_k_matmul: inner dimensions differ (%d vs %d)' % (K, b.shape[0]))
    E = b.numel() // K
    out = torch.empty((M,) + tuple(b.shape[1:]), dtype=torch.float32, device=b.device)
    rs, cs = (1, a.shape[1]) if transpose_a else (a.shape[1], 1)
    _lib.check(lib.mpqe_small_k_matmul(_ptr(a), rs, cs, _ptr(b), M, K, E, _ptr(out), _stream()), 'mpqe_small_k_matmul')
    _count()
    return out


def rows_dot(x, y):
    """out[m, k] = <x[m], y[k]> over all trailing elements (mpqe_rows_dot: d att of the basis decomposition)."""
    lib = _lib.load()
    x, y = _chk(x, torch.float32, 'x'), _chk(y, torch.float32, 'y')
    M, K = x.shape[0], y.shape[0]
    E = x.numel() // M
    if y.numel() // K != E:
        raise _lib.MpqeError('rows_dot: row lengths differ')
    out = torch.empty(M, K, dtype=torch.float32, device=x.device)
    _lib.check(lib.mpqe_rows_dot(_ptr(x), _ptr(y), M, K, E, _ptr(out), _stream()), 'mpqe_rows_dot')
    _count()
    return out


def matrix_sum_multi(items):
    """items: [(dst [D, D] view, [src [D, D] views], accumulate)]; dst (+)= sum(src) in order (mpqe_matrix_sum_multi).
    Destinations must be distinct within one call; a summand list longer than MPQE_MAX_MATSUM_SRCS is continued by
    accumulating follow-up launches."""
    lib = _lib.load()
    pending = [(dst, list(srcs), bool(acc)) for dst, srcs, acc in items]
    while pending:
        later = []
        for i in range(0, len(pending), _lib.MAX_MATSUM_ITEMS):
            chunk = pending[i:i + _lib.MAX_MATSUM_ITEMS]
            arr = (_lib.MatsumItem * len(chunk))()
            for j, (dst, srcs, acc) in enumerate(chunk):
                head, tail = srcs[:_lib.MAX_MATSUM_SRCS], srcs[_lib.MAX_MATSUM_SRCS:]
                arr[j].dst = _chk(dst, torch.float32, 'dst').data_ptr()
                for k, m in enumerate(head):
                    if m.numel() != D * D:
                        raise _lib.MpqeError('matrix_sum_multi: summands must be [%d,%d]' % (D, D))
                    arr[j].src[k] = _chk(m, torch.float32, 'summand').data_ptr()
                arr[j].num_src, arr[j].accumulate = len(head), int(acc)
                if tail:
                    later.append((dst, tail, True))
            _lib.check(lib.mpqe_matrix_sum_multi(arr, len(chunk), _stream()), 'mpqe_matrix_sum_multi')
            _count()
        pending = later


PACKED_FLOATS = 2 * D * D


def layer_generation():
    """Which tcgen05 layer kernel the library dispatches to (MPQE_LAYER_KERNEL=1: the first-generation kernel)."""
    import os
    return 1 if os.environ.get('MPQE_LAYER_KERNEL') == '1' else 2


def pack_weights(mats, transposed=None):
    """mats: [count, D, D] tensor (or list of [D, D] views) -> [count, PACKED_FLOATS] tf32 hi/lo tile images for the
    tensor-core layer kernel (see mpqe_pack_weights).  transposed[i]: pack the image of mats[i]^T instead."""
    lib = _lib.load()
    views = [mats[i] for i in range(mats.shape[0])] if torch.is_tensor(mats) else list(mats)
    for v in views:
        _chk(v, torch.float32, 'matrix')
    out = torch.empty(len(views), PACKED_FLOATS, dtype=torch.float32, device=views[0].device)
    ptrs = (C.c_void_p * len(views))(*[v.data_ptr() for v in views])
    flags = None
    if transposed is not None and any(transposed):
        flags = (C.c_uint8 * len(views))(*[1 if f else 0 for f in transposed])
    _lib.check(lib.mpqe_pack_weights_ex(ptrs, flags, len(views), _ptr(out), _stream()), 'mpqe_pack_weights')
    _count((len(views) + 255) // 256)
    return out


def colsum(src, rows, stride, out, scale=1.0, accumulate=False):
    """out[D] (+)= scale * sum_r src.flat[r*stride : r*stride+D]; `src` may be an offset view into a larger buffer."""
    lib = _lib.load()
    nbytes = lib.mpqe_colsum_workspace_bytes(rows)
    ws = workspace(nbytes, out.device, 'colsum')
    _lib.check(lib.mpqe_colsum(C.c_void_p(src.data_ptr()), rows, stride, scale, _ptr(_chk(out, torch.float32, 'out')),
                               int(accumulate), _ptr(ws), ws.numel(), _stream()), 'mpqe_colsum')
    _count(2)


def transpose(src, dst=None):
    """[count, r, c] (or [r, c]) -> per-matrix transpose."""
    lib = _lib.load()
    _chk(src, torch.float32, 'src')
    count = src.shape[0] if src.dim() == 3 else 1
    r, c = src.shape[-2], src.shape[-1]
    if dst is None:
        dst = torch.empty(src.shape[:-2] + (c, r), dtype=torch.float32, device=src.device)
    _lib.check(lib.mpqe_transpose(_ptr(src), _ptr(_chk(dst, torch.float32, 'dst')), count, r, c, _stream()),
               'mpqe_transpose')
    _count()
    return dst


# ---------------------------------------------------------------------------------------------------------------
# Row kernels
# ---------------------------------------------------------------------------------------------------------------
def gather_normalize(table, id2row, ids, out=None, out_offset=0, out_stride=D, ids_offset=0, ids_stride=1,
                     count=None, inv_norm=None):
    """out.flat[out_offset + i*out_stride : +D] = normalised table[id2row[ids.flat[ids_offset + i*ids_stride]]]."""
    lib = _lib.load()
    _chk(table, torch.float32, 'table')
    _chk(ids, torch.int64, 'ids')
    if count is None:
        count = ids.numel()
    if out is None:
        out = torch.empty(count, D, dtype=torch.float32, device=table.device)
    _chk(out, torch.float32, 'out')
    _lib.check(lib.mpqe_gather_normalize_fwd(
        _ptr(table), table.shape[0], _ptr(id2row), C.c_void_p(ids.data_ptr() + 8 * ids_offset), ids_stride, count,
        C.c_void_p(out.data_ptr() + 4 * out_offset), out_stride, _ptr(inv_norm), _stream()),
        'mpqe_gather_normalize_fwd')
    _count()
    return out


def gather_normalize_bwd(table, id2row, ids, grad, rows_out, rows_id, grad_offset=0, grad_stride=D, ids_offset=0,
                         ids_stride=1, count=None, rows_offset=0):
    lib = _lib.load()
    if count is None:
        count = ids.numel()
    _lib.check(lib.mpqe_gather_normalize_bwd(
        _ptr(table), _ptr(id2row), C.c_void_p(ids.data_ptr() + 8 * ids_offset), ids_stride, count,
        C.c_void_p(grad.data_ptr() + 4 * grad_offset), grad_stride,
        C.c_void_p(rows_out.data_ptr() + 4 * D * rows_offset), C.c_void_p(rows_id.data_ptr() + 8 * rows_offset),
        _stream()), 'mpqe_gather_normalize_bwd')
    _count()


def broadcast_rows(src, src_rows, out, out_offset, out_stride, count):
    lib = _lib.load()
    _lib.check(lib.mpqe_broadcast_rows(_ptr(_chk(src, torch.float32, 'src')), _ptr(_chk(src_rows, torch.int64, 'rows')),
                                       src_rows.numel(), C.c_void_p(out.data_ptr() + 4 * out_offset), out_stride,
                                       count, _stream()), 'mpqe_broadcast_rows')
    _count()


def max_readout(z, B, n):
    lib = _lib.load()
    q = torch.empty(B, D, dtype=torch.float32, device=z.device)
    arg = torch.empty(B, D, dtype=torch.int64, device=z.device)
    _lib.check(lib.mpqe_max_readout_fwd(_ptr(_chk(z, torch.float32, 'z')), B, n, _ptr(q), _ptr(arg), _stream()),
               'mpqe_max_readout_fwd')
    _count()
    return q, arg


def max_readout_bwd(dq, argmax, B, n):
    lib = _lib.load()
    g = torch.empty(B, n, D, dtype=torch.float32, device=dq.device)
    _lib.check(lib.mpqe_max_readout_bwd(_ptr(_chk(dq, torch.float32, 'dq')), _ptr(argmax), B, n, _ptr(g), _stream()),
               'mpqe_max_readout_bwd')
    _count()
    return g


def cosine_margin(q, table, id2row, ids_pos, ids_neg, margin, loss_out=None):
    lib = _lib.load()
    B = q.shape[0]
    dev = q.device
    sp = torch.empty(B, dtype=torch.float32, device=dev)
    sn = torch.empty(B, dtype=torch.float32, device=dev)
    loss = loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=dev)
    ws = workspace(lib.mpqe_margin_loss_workspace_bytes(B), dev, 'loss')
    _lib.check(lib.mpqe_cosine_margin_fwd(_ptr(_chk(q, torch.float32, 'q')), B, _ptr(_chk(table, torch.float32, 'table')),
                                          _ptr(id2row), _ptr(_chk(ids_pos, torch.int64, 'ids_pos')),
                                          _ptr(_chk(ids_neg, torch.int64, 'ids_neg')), margin, _ptr(sp), _ptr(sn),
                                          _ptr(loss), _ptr(ws), ws.numel(), _stream()), 'mpqe_cosine_margin_fwd')
    _count(2)
    return sp, sn, loss


def cosine_margin_bwd(q, table, id2row, ids_pos, ids_neg, margin, grad_loss, rows_out, rows_id, rows_offset=0):
    """rows_out/rows_id receive 2B entries starting at rows_offset."""
    lib = _lib.load()
    B = q.shape[0]
    dq = torch.empty(B, D, dtype=torch.float32, device=q.device)
    _lib.check(lib.mpqe_cosine_margin_bwd(_ptr(q), B, _ptr(table), _ptr(id2row), _ptr(ids_pos), _ptr(ids_neg), margin,
                                          _ptr(_chk(grad_loss, torch.float32, 'grad_loss')), _ptr(dq),
                                          C.c_void_p(rows_out.data_ptr() + 4 * D * rows_offset),
                                          C.c_void_p(rows_id.data_ptr() + 8 * rows_offset), _stream()),
               'mpqe_cosine_margin_bwd')
    _count()
    return dq


def cosine_scores(q, table, id2row, ids, offsets=None, out=None, out_offset=0):
    lib = _lib.load()
    count = ids.numel()
    if out is None:
        out = torch.empty(count, dtype=torch.float32, device=q.device)
    _lib.check(lib.mpqe_cosine_scores(_ptr(_chk(q, torch.float32, 'q')), q.shape[0], _ptr(offsets),
                                      _ptr(_chk(table, torch.float32, 'table')), _ptr(id2row),
                                      _ptr(_chk(ids, torch.int64, 'ids')), count,
                                      C.c_void_p(out.data_ptr() + 4 * out_offset), _stream()), 'mpqe_cosine_scores')
    _count()
    return out


def cosine_scores_bwd(q, table, id2row, ids, offsets, grad_scores, grad_offset, dq, accumulate, rows_out, rows_id,
                      rows_offset=0):
    lib = _lib.load()
    _lib.check(lib.mpqe_cosine_scores_bwd(_ptr(q), q.shape[0], _ptr(offsets), _ptr(table), _ptr(id2row), _ptr(ids),
                                          ids.numel(), C.c_void_p(grad_scores.data_ptr() + 4 * grad_offset), _ptr(dq),
                                          int(accumulate), C.c_void_p(rows_out.data_ptr() + 4 * D * rows_offset),
                                          C.c_void_p(rows_id.data_ptr() + 8 * rows_offset), _stream()),
               'mpqe_cosine_scores_bwd')
    _count()


def rank_counts_ragged(pos, neg, offsets):
    lib = _lib.load()
    B = pos.numel()
    left = torch.empty(B, dtype=torch.int64, device=pos.device)
    right = torch.empty(B, dtype=torch.int64, device=pos.device)
    _lib.check(lib.mpqe_rank_counts_ragged(_ptr(_chk(pos, torch.float32, 'pos')), _ptr(_chk(neg, torch.float32, 'neg')),
                                           _ptr(_chk(offsets, torch.int64, 'offsets')), B, _ptr(left), _ptr(right),
                                           _stream()), 'mpqe_rank_counts_ragged')
    _count()
    return left, right


def auc_counts(pos, neg, counts=None):
    """counts[0] += #(neg < pos), counts[1] += #(neg == pos) over all pairs (int64 [2] on the device, mpqe_auc_counts)."""
    lib = _lib.load()
    if counts is None:
        counts = torch.zeros(2, dtype=torch.int64, device=pos.device)
    _lib.check(lib.mpqe_auc_counts(_ptr(_chk(pos, torch.float32, 'pos')), pos.numel(),
                                   _ptr(_chk(neg, torch.float32, 'neg')), neg.numel(), _ptr(counts), _stream()),
               'mpqe_auc_counts')
    _count()
    return counts


def rank_counts_table(q, pos, table, row_begin, row_end, left, right, use_tensor_cores=False):
    """Accumulates into left/right (int64 [B])."""
    lib = _lib.load()
    B = q.shape[0]
    nbytes = lib.mpqe_rank_counts_table_workspace_bytes(B, row_end - row_begin)
    ws = workspace(nbytes, q.device, 'rank')
    _lib.check(lib.mpqe_rank_counts_table(_ptr(_chk(q, torch.float32, 'q')), B, _ptr(_chk(pos, torch.float32, 'pos')),
                                          _ptr(_chk(table, torch.float32, 'table')), row_begin, row_end,
                                          _ptr(_chk(left, torch.int64, 'left')), _ptr(_chk(right, torch.int64, 'right')),
                                          _ptr(ws), ws.numel(), int(use_tensor_cores), _stream()),
               'mpqe_rank_counts_table')
    _count(3)


class RankTable(object):
    """The table-dependent half of the full-entity ranking, prepared once per table shard (mpqe_rank_table_prepare:
    1/||row|| and the pre-split tile images of rows [row_begin, row_end)) and reused by every batch."""

    def __init__(self, table, row_begin, row_end, use_tensor_cores):
        lib = _lib.load()
        self.table, self.row_begin, self.row_end = _chk(table, torch.float32, 'table'), int(row_begin), int(row_end)
        self.tc = bool(use_tensor_cores)
        self.version = table._version
        nbytes = lib.mpqe_rank_table_workspace_bytes(self.row_end - self.row_begin, int(self.tc))
        self.ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=table.device)
        _lib.check(lib.mpqe_rank_table_prepare(_ptr(table), self.row_begin, self.row_end, _ptr(self.ws), self.ws.numel(),
                                               int(self.tc), _stream()), 'mpqe_rank_table_prepare')
        _count(2 if self.tc else 1)

    def counts(self, q, pos, left, right):
        """Accumulates (count_lt, count_le) of the shard's rows against each query's positive into left/right."""
        lib = _lib.load()
        B = q.shape[0]
        ws = workspace(lib.mpqe_rank_query_workspace_bytes(B), q.device, 'rankq')
        _lib.check(lib.mpqe_rank_counts_prepared(_ptr(_chk(q, torch.float32, 'q')), B, _ptr(_chk(pos, torch.float32, 'pos')),
                                                 _ptr(self.table), self.row_begin, self.row_end, _ptr(self.ws),
                                                 _ptr(_chk(left, torch.int64, 'left')),
                                                 _ptr(_chk(right, torch.int64, 'right')), _ptr(ws), ws.numel(),
                                                 int(self.tc), _stream()), 'mpqe_rank_counts_prepared')
        _count(2)


# ---------------------------------------------------------------------------------------------------------------
# Integer layout kernels
# ---------------------------------------------------------------------------------------------------------------
def build_query_graph(n, src, dst, rel, B, device):
    lib = _lib.load()
    E = len(src)
    edge_index = torch.empty(2, B * E, dtype=torch.int64, device=device)
    edge_type = torch.empty(B * E, dtype=torch.int64, device=device)
    batch = torch.empty(B * n, dtype=torch.int64, device=device)
    s = (C.c_int32 * E)(*src)
    t = (C.c_int32 * E)(*dst)
    r = (C.c_int64 * E)(*rel)
    _lib.check(lib.mpqe_build_query_graph(n, E, s, t, r, B, _ptr(edge_index), _ptr(edge_type), _ptr(batch), _stream()),
               'mpqe_build_query_graph')
    _count()
    return edge_index, edge_type, batch


def relation_sort(edge_type, num_relations):
    lib = _lib.load()
    nE = edge_type.numel()
    perm = torch.empty(nE, dtype=torch.int64, device=edge_type.device)
    offsets = torch.empty(num_relations + 1, dtype=torch.int64, device=edge_type.device)
    ws = workspace(lib.mpqe_relation_sort_workspace_bytes(nE, num_relations), edge_type.device, 'sort')
    _lib.check(lib.mpqe_relation_sort(_ptr(_chk(edge_type, torch.int64, 'edge_type')), nE, num_relations, _ptr(perm),
                                      _ptr(offsets), _ptr(ws), ws.numel(), _stream()), 'mpqe_relation_sort')
    _count(6)
    return perm, offsets


def sparse_rows_combine(rows_id, rows, table_rows, pad_id=0):
    """(unique_ids[count], unique_rows[count, D], num_unique[1]); entries past num_unique are (pad_id, zero row).
    Input ids >= table_rows are treated as padding and dropped (pass pad_id=table_rows when the result is combined
    again after an all-gather)."""
    lib = _lib.load()
    count = rows_id.numel()
    dev = rows.device
    uid = torch.empty(count, dtype=torch.int64, device=dev)
    urows = torch.empty(count, D, dtype=torch.float32, device=dev)
    num = torch.empty(1, dtype=torch.int64, device=dev)
    ws = workspace(lib.mpqe_sparse_rows_workspace_bytes(count), dev, 'sparse')
    _lib.check(lib.mpqe_sparse_rows_combine(_ptr(_chk(rows_id, torch.int64, 'rows_id')),
                                            _ptr(_chk(rows, torch.float32, 'rows')), count, table_rows, pad_id, _ptr(uid),
                                            _ptr(urows), _ptr(num), _ptr(ws), ws.numel(), _stream()),
               'mpqe_sparse_rows_combine')
    _count(10)
    return uid, urows, num


class SparseRowsPlan(object):
    """The id-only half of `sparse_rows_combine` (mpqe_sparse_rows_plan): sort + segmentation of `rows_id`, held in a
    private workspace until `apply(rows)` sums the gradient rows.  The plan may be built on another stream."""

    def __init__(self, rows_id, table_rows):
        lib = _lib.load()
        self.count, self.table_rows = rows_id.numel(), int(table_rows)
        dev = rows_id.device
        self.num = torch.empty(1, dtype=torch.int64, device=dev)
        nbytes = lib.mpqe_sparse_rows_workspace_bytes(self.count)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)   # private: lives across streams
        self.nbytes = nbytes
        _lib.check(lib.mpqe_sparse_rows_plan(_ptr(_chk(rows_id, torch.int64, 'rows_id')), self.count, self.table_rows,
                                             _ptr(self.num), _ptr(self.ws), nbytes, _stream()), 'mpqe_sparse_rows_plan')
        passes = (max(1, self.table_rows.bit_length()) + 6) // 7      # 7-bit digits, see mpqe_sparse_rows_plan
        _count(7 + 4 * passes)

    def apply(self, rows, pad_id=0, scale=1.0):
        lib = _lib.load()
        uid = torch.empty(self.count, dtype=torch.int64, device=rows.device)
        urows = torch.empty(self.count, D, dtype=torch.float32, device=rows.device)
        _lib.check(lib.mpqe_sparse_rows_apply(_ptr(_chk(rows, torch.float32, 'rows')), self.count, self.table_rows,
                                              pad_id, float(scale), _ptr(uid), _ptr(urows), _ptr(self.num), _ptr(self.ws),
                                              self.nbytes, _stream()), 'mpqe_sparse_rows_apply')
        _count()
        return uid, urows, self.num

    def apply_peers(self, peer_ptrs, per_rank_count, pad_id=0, scale=1.0, capacity=None):
        """`apply` over the row buffers of all ranks read in place through peer-mapped memory: `peer_ptrs[r]` is the
        device address of rank r's rows in this process (torch symmetric memory); the plan must have been built over
        the rank-major concatenation of the ranks' ids (mpqe_sparse_rows_apply_peers)."""
        lib = _lib.load()
        peer_ptrs = [p.data_ptr() if torch.is_tensor(p) else p for p in peer_ptrs]     # (tensors: gathered local copies)
        world = len(peer_ptrs)
        if world * per_rank_count != self.count:
            raise _lib.MpqeError('apply_peers: plan covers %d pairs, peers provide %d' % (self.count, world * per_rank_count))
        dev = self.num.device
        # `capacity`: an upper bound of the distinct rows the plan can hold (the rows this rank owns, for an owner
        # plan): outputs and padding writes are sized by it instead of by the total number of pairs
        cap = self.count if capacity is None else max(1, min(int(capacity), self.count))
        uid = torch.empty(cap, dtype=torch.int64, device=dev)
        urows = torch.empty(cap, D, dtype=torch.float32, device=dev)
        ptrs = (C.c_void_p * world)(*[int(p) for p in peer_ptrs])
        _lib.check(lib.mpqe_sparse_rows_apply_peers(ptrs, world, int(per_rank_count), self.table_rows, pad_id,
                                                    float(scale), _ptr(uid), _ptr(urows), cap, _ptr(self.num),
                                                    _ptr(self.ws), self.nbytes, _stream()), 'mpqe_sparse_rows_apply_peers')
        _count()
        return uid, urows, self.num


def scatter_rows(ids, rows, num, dense, accumulate=False):
    lib = _lib.load()
    _lib.check(lib.mpqe_scatter_rows(_ptr(ids), _ptr(rows), _ptr(num), ids.numel(), _ptr(_chk(dense, torch.float32, 'dense')),
                                     int(accumulate), _stream()), 'mpqe_scatter_rows')
    _count()


def adam_dense(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step):
    lib = _lib.load()
    _lib.check(lib.mpqe_adam_dense(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), lr, beta1,
                                   beta2, eps, step, _stream()), 'mpqe_adam_dense')
    _count()


# ---------------------------------------------------------------------------------------------------------------
# Multi-item launches (one kernel for every formula group of a step)
# ---------------------------------------------------------------------------------------------------------------
def _addr(t, elem_offset=0):
    return t.data_ptr() + elem_offset * t.element_size() if t is not None else 0


PEER_TABLES = {}   # table data_ptr -> (device int64 tensor of the ranks' table addresses, rows per owner)


def _peer_tables(table):
    """(device pointer to the peers' addresses of this table, rows per owner) when the table is owned row-range-wise
    across a data-parallel group (see train_step.TrainStep.shard_tables), else (0, 0): read the local table."""
    ent = PEER_TABLES.get(table.data_ptr())
    if ent is None:
        return 0, 0
    return ent[0].data_ptr(), ent[1]


class GatherItem(object):
    """One (group, node slot) work item of `gather_multi` (see mpqe_gather_item_t); tensors + element offsets."""

    def __init__(self, table, id2row, ids, count, ids_offset=0, ids_stride=1, out=None, out_offset=0, out_stride=D,
                 grad=None, grad_offset=0, grad_stride=D, rows_out=None, rows_id=None, rows_offset=0, id_offset=0,
                 normalize=True, norm=None, norm_offset=0):
        self.table, self.id2row, self.ids, self.count = table, id2row, ids, int(count)
        self.ids_offset, self.ids_stride = int(ids_offset), int(ids_stride)
        self.out, self.out_offset, self.out_stride = out, int(out_offset), int(out_stride)
        self.grad, self.grad_offset, self.grad_stride = grad, int(grad_offset), int(grad_stride)
        self.rows_out, self.rows_id, self.rows_offset = rows_out, rows_id, int(rows_offset)
        self.id_offset, self.normalize = int(id_offset), bool(normalize)
        self.norm, self.norm_offset = norm, int(norm_offset)   # [count] row norms: written forward, read backward

    def to_c(self):
        it = _lib.GatherItem()
        it.table, it.table_rows, it.id2row = _addr(self.table), self.table.shape[0], _addr(self.id2row)
        it.ids, it.ids_stride, it.count = _addr(self.ids, self.ids_offset), self.ids_stride, self.count
        it.out, it.out_stride = _addr(self.out, self.out_offset), self.out_stride
        it.grad, it.grad_stride = _addr(self.grad, self.grad_offset), self.grad_stride
        it.rows_out, it.rows_id = _addr(self.rows_out, self.rows_offset * D), _addr(self.rows_id, self.rows_offset)
        it.id_offset, it.normalize = self.id_offset, int(self.normalize)
        it.peer_tables, it.peer_chunk = _peer_tables(self.table)
        it.norm = _addr(self.norm, self.norm_offset)
        return it


def gather_multi(items, backward=False):
    """backward: False = forward gather, True = backward (rows + ids), 'ids' = only the row ids of the backward."""
    lib = _lib.load()
    backward = 2 if backward == 'ids' else int(bool(backward))
    for i in range(0, len(items), _lib.MAX_GATHER_ITEMS):
        chunk = items[i:i + _lib.MAX_GATHER_ITEMS]
        arr = (_lib.GatherItem * len(chunk))(*[it.to_c() for it in chunk])
        _lib.check(lib.mpqe_gather_multi(arr, len(chunk), int(backward), _stream()), 'mpqe_gather_multi')
        _count()


class GatherLaunch(object):
    """A pre-marshalled `gather_multi` call for items whose tensors stay put (the static buffers of graph mode): the
    per-step host cost is one ctypes call per chunk instead of rebuilding the item structs."""

    def __init__(self, items, backward=False):
        self.mode = 2 if backward == 'ids' else int(bool(backward))
        self.keep = list(items)
        self.chunks = []
        for i in range(0, len(self.keep), _lib.MAX_GATHER_ITEMS):
            chunk = self.keep[i:i + _lib.MAX_GATHER_ITEMS]
            self.chunks.append(((_lib.GatherItem * len(chunk))(*[it.to_c() for it in chunk]), len(chunk)))

    def launch(self):
        lib = _lib.load()
        for arr, n in self.chunks:
            _lib.check(lib.mpqe_gather_multi(arr, n, self.mode, _stream()), 'mpqe_gather_multi')
            _count()


class MarginItem(object):
    """One formula group of `cosine_margin_multi` (see mpqe_margin_item_t)."""

    def __init__(self, q, table, id2row, ids_pos, ids_neg, hinge=None, loss=None, grad_loss=None, dq=None,
                 rows_out=None, rows_id=None, rows_offset=0, id_offset=0):
        self.q, self.table, self.id2row, self.ids_pos, self.ids_neg = q, table, id2row, ids_pos, ids_neg
        self.hinge, self.loss, self.grad_loss, self.dq = hinge, loss, grad_loss, dq
        self.rows_out, self.rows_id, self.rows_offset, self.id_offset = rows_out, rows_id, int(rows_offset), int(id_offset)

    def to_c(self):
        it = _lib.MarginItem()
        it.q, it.B, it.table, it.id2row = _addr(self.q), self.q.shape[0], _addr(self.table), _addr(self.id2row)
        it.ids_pos, it.ids_neg = _addr(self.ids_pos), _addr(self.ids_neg)
        it.score_pos = it.score_neg = 0
        it.hinge, it.loss, it.grad_loss, it.dq = _addr(self.hinge), _addr(self.loss), _addr(self.grad_loss), _addr(self.dq)
        it.rows_out, it.rows_id = _addr(self.rows_out, self.rows_offset * D), _addr(self.rows_id, self.rows_offset)
        it.id_offset = self.id_offset
        it.peer_tables, it.peer_chunk = _peer_tables(self.table)
        return it


def cosine_margin_multi(items, margin, backward=False):
    """backward: False = forward (losses), True = backward (dq, table rows), 'both' = one pass doing both (needs
    grad_loss before the forward: a training step knows d total / d loss_i up front)."""
    lib = _lib.load()
    mode = 2 if backward == 'both' else int(bool(backward))
    for i in range(0, len(items), _lib.MAX_MARGIN_ITEMS):
        chunk = items[i:i + _lib.MAX_MARGIN_ITEMS]
        arr = (_lib.MarginItem * len(chunk))(*[it.to_c() for it in chunk])
        _lib.check(lib.mpqe_cosine_margin_multi(arr, len(chunk), margin, mode, _stream()),
                   'mpqe_cosine_margin_multi')
        _count(1 if mode == 1 else 2)


class ColsumItem(object):
    """dst[D] += scale * sum_r src.flat[r*stride : r*stride+D]  (src may be an offset view into a larger buffer)."""

    def __init__(self, src, rows, stride, dst, scale=1.0):
        self.src, self.rows, self.stride, self.dst, self.scale = src, int(rows), int(stride), dst, float(scale)

    def to_c(self):
        it = _lib.ColsumItem()
        it.src, it.rows, it.stride, it.dst, it.scale = self.src.data_ptr(), self.rows, self.stride, self.dst.data_ptr(), self.scale
        return it


def colsum_multi(items, device):
    """Applies every item in order (2 launches per <=64 items)."""
    lib = _lib.load()
    for i in range(0, len(items), _lib.MAX_COLSUM_ITEMS):
        chunk = items[i:i + _lib.MAX_COLSUM_ITEMS]
        arr = (_lib.ColsumItem * len(chunk))(*[it.to_c() for it in chunk])
        ws = workspace(lib.mpqe_colsum_multi_workspace_bytes(arr, len(chunk)), device, 'colsum')
        _lib.check(lib.mpqe_colsum_multi(arr, len(chunk), _ptr(ws), ws.numel(), _stream()), 'mpqe_colsum_multi')
        _count(2)


# ---------------------------------------------------------------------------------------------------------------
# Regulariser, optimiser and negative sampling of the fused training step (csrc/optim.cu)
# ---------------------------------------------------------------------------------------------------------------
def l2_reg(params, grads, weight_decay, grad_scale, losses=None, norms=None):
    """model.py:487-492 on the device: losses[:] += weight_decay * sum_i ||params[i]||, and
    grads[i] += grad_scale * weight_decay * params[i] / ||params[i]|| (grads[i] laid out like params[i], or None)."""
    lib = _lib.load()
    arr = (_lib.L2Item * len(params))()
    for i, (p, g) in enumerate(zip(params, grads)):
        arr[i].param = _chk(p, torch.float32, 'param').data_ptr()
        arr[i].grad = _chk(g, torch.float32, 'grad').data_ptr() if g is not None else 0
        if g is not None and g.numel() != p.numel():
            raise _lib.MpqeError('l2_reg: gradient and parameter sizes differ')
        arr[i].numel = p.numel()
    _lib.check(lib.mpqe_l2_reg_multi(arr, len(params), float(weight_decay), float(grad_scale), _ptr(losses),
                                     losses.numel() if losses is not None else 0, _ptr(norms), _stream()),
               'mpqe_l2_reg_multi')
    _count()


def adam_state(device):
    """Device-resident optimiser clock (mpqe_adam_state_t: step, lr / bias correction 1, sqrt(bias correction 2))."""
    return torch.zeros(4, dtype=torch.int32, device=device)


def adam_tick(state, lr, beta1, beta2):
    lib = _lib.load()
    _lib.check(lib.mpqe_adam_tick(_ptr(state), lr, beta1, beta2, _stream()), 'mpqe_adam_tick')
    _count()


def adam_multi(items, lr, beta1, beta2, eps, step=0, state=None):
    """One torch.optim.Adam step over [(param, grad, exp_avg, exp_avg_sq)] dense tensors, <= 32 per launch.
    `state`: device clock from `adam_state` (graph-capturable) instead of the host `step`."""
    lib = _lib.load()
    for i in range(0, len(items), _lib.MAX_ADAM_ITEMS):
        chunk = items[i:i + _lib.MAX_ADAM_ITEMS]
        arr = (_lib.AdamItem * len(chunk))()
        for j, (p, g, m, v) in enumerate(chunk):
            for t, name in ((p, 'param'), (g, 'grad'), (m, 'exp_avg'), (v, 'exp_avg_sq')):
                _chk(t, torch.float32, name)
                if t.numel() != p.numel():
                    raise _lib.MpqeError('adam_multi: %s has %d elements, param %d' % (name, t.numel(), p.numel()))
            arr[j].param, arr[j].grad, arr[j].exp_avg, arr[j].exp_avg_sq = (p.data_ptr(), g.data_ptr(), m.data_ptr(),
                                                                            v.data_ptr())
            arr[j].numel = p.numel()
        _lib.check(lib.mpqe_adam_multi(arr, len(chunk), lr, beta1, beta2, eps, int(step), _ptr(state), _stream()),
                   'mpqe_adam_multi')
        _count()


class RowAdam(object):
    """torch.optim.Adam over the entity tables driven by row-sparse gradients, trajectory-equivalent to the dense
    optimiser of the reference (train.py:86-88): see mpqe_adam_rows_catchup / mpqe_adam_rows_apply.
    `tables` = [(table tensor [rows, D], first global row id)]."""

    def __init__(self, tables, lr=0.01, betas=(0.9, 0.999), eps=1e-8):
        self.hyper = (float(lr), float(betas[0]), float(betas[1]), float(eps))
        self.tables = [(t, int(begin)) for t, begin in tables]
        if len(self.tables) > _lib.MAX_TABLES:
            raise _lib.MpqeError('RowAdam: at most %d tables' % _lib.MAX_TABLES)
        dev = self.tables[0][0].device
        self.exp_avg = [torch.zeros_like(t) for t, _ in self.tables]
        self.exp_avg_sq = [torch.zeros_like(t) for t, _ in self.tables]
        self.total_rows = max(begin + t.shape[0] for t, begin in self.tables)
        self.last = torch.zeros(self.total_rows, dtype=torch.int32, device=dev)
        self.arr = (_lib.AdamTable * len(self.tables))()
        for i, (t, begin) in enumerate(self.tables):
            _chk(t, torch.float32, 'table')
            self.arr[i].table, self.arr[i].exp_avg = t.data_ptr(), self.exp_avg[i].data_ptr()
            self.arr[i].exp_avg_sq, self.arr[i].row_begin, self.arr[i].rows = self.exp_avg_sq[i].data_ptr(), begin, t.shape[0]

    def catchup(self, ids, upto_step=0, state=None):
        """Zero-gradient steps up to `upto_step` (or the device clock `state`) for the rows `ids` (global row ids,
        duplicates allowed) -- call with the ids a step is about to read, before its forward and before the step's
        tick.  ids=None: every row (before eval / export)."""
        lib = _lib.load()
        lr, b1, b2, eps = self.hyper
        count = ids.numel() if ids is not None else self.total_rows
        _lib.check(lib.mpqe_adam_rows_catchup(self.arr, len(self.tables), _ptr(ids), count, int(upto_step), lr, b1, b2,
                                              eps, _ptr(state), _ptr(self.last), _stream()), 'mpqe_adam_rows_catchup')
        _count()

    def apply(self, ids, rows, num, step=0, state=None):
        """Adam step `step` (or the device clock) on the combined (unique ids, summed rows, device count) gradient."""
        lib = _lib.load()
        lr, b1, b2, eps = self.hyper
        _lib.check(lib.mpqe_adam_rows_apply(self.arr, len(self.tables), _ptr(_chk(ids, torch.int64, 'ids')),
                                            _ptr(_chk(rows, torch.float32, 'rows')), _ptr(num), ids.numel(), int(step),
                                            lr, b1, b2, eps, _ptr(state), _ptr(self.last), _stream()),
                   'mpqe_adam_rows_apply')
        _count()


def sample_negatives(candidates, offsets, count, seed, step, first_query=0, num_queries_total=None, query_index=None,
                     out=None):
    """Device-side negative draw (model.py:470-476): one candidate per batch position from the CSR (candidates,
    offsets) of the stored query set, or from the shared list `candidates` when offsets is None (1-chain)."""
    lib = _lib.load()
    if out is None:
        out = torch.empty(count, dtype=torch.int64, device=candidates.device)
    if num_queries_total is None:
        num_queries_total = offsets.numel() - 1 if offsets is not None else count
    shared = candidates.numel() if offsets is None else 0
    _lib.check(lib.mpqe_sample_negatives(_ptr(_chk(candidates, torch.int64, 'candidates')), _ptr(offsets),
                                         _ptr(query_index), int(first_query), int(num_queries_total), int(shared),
                                         int(count), int(seed), int(step), _ptr(_chk(out, torch.int64, 'out')),
                                         _stream()), 'mpqe_sample_negatives')
    _count()
    return out


# ---------------------------------------------------------------------------------------------------------------
# Data-parallel exchange over peer-mapped memory (csrc/peer.cu)
# ---------------------------------------------------------------------------------------------------------------
def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def peer_barrier(flag_ptrs, rank, epoch):
    """Flag barrier across the ranks of a peer group, on the current stream (graph-capturable)."""
    lib = _lib.load()
    _lib.check(lib.mpqe_peer_barrier(_ptr_array(flag_ptrs), int(rank), len(flag_ptrs), _ptr(epoch), _stream()),
               'mpqe_peer_barrier')
    _count()


def allreduce_peers(buf_ptrs, numel, scale, out):
    """out = scale * sum over ranks of their buffers (read in place over NVLink, rank order)."""
    lib = _lib.load()
    _lib.check(lib.mpqe_allreduce_peers(_ptr_array(buf_ptrs), len(buf_ptrs), int(numel), float(scale),
                                        _ptr(_chk(out, torch.float32, 'out')), _stream()), 'mpqe_allreduce_peers')
    _count()


def reduce_scatter_peers(buf_ptrs, rank, numel, scale):
    """First half of the two-shot all-reduce: this rank's slice of every rank's buffer, summed in place into its own."""
    lib = _lib.load()
    _lib.check(lib.mpqe_reduce_scatter_peers(_ptr_array(buf_ptrs), len(buf_ptrs), int(rank), int(numel), float(scale),
                                             _stream()), 'mpqe_reduce_scatter_peers')
    _count()


def all_gather_peers(buf_ptrs, numel, out):
    """Second half: every slice read from its owner's buffer into `out` (after a barrier)."""
    lib = _lib.load()
    _lib.check(lib.mpqe_all_gather_peers(_ptr_array(buf_ptrs), len(buf_ptrs), int(numel),
                                         _ptr(_chk(out, torch.float32, 'out')), _stream()), 'mpqe_all_gather_peers')
    _count()


def owner_plan(id_ptrs, rank, per_rank_count, table_begin, table_rows, total_rows, device):
    """SparseRowsPlan over the ids of all ranks (read in place from `id_ptrs`), restricted to the rows `rank` owns."""
    lib = _lib.load()
    id_ptrs = [_chk(p, torch.int64, 'ids').data_ptr() if torch.is_tensor(p) else p for p in id_ptrs]
    world = len(id_ptrs)
    plan = object.__new__(SparseRowsPlan)
    plan.count, plan.table_rows = world * int(per_rank_count), int(total_rows)
    plan.num = torch.empty(1, dtype=torch.int64, device=device)
    plan.nbytes = lib.mpqe_sparse_rows_workspace_bytes(plan.count)
    plan.ws = torch.empty(plan.nbytes, dtype=torch.uint8, device=device)
    nt = len(table_begin)
    _lib.check(lib.mpqe_sparse_rows_plan_owner(_ptr_array(id_ptrs), world, int(rank), int(per_rank_count),
                                               (C.c_int64 * nt)(*[int(x) for x in table_begin]),
                                               (C.c_int64 * nt)(*[int(x) for x in table_rows]), nt, int(total_rows),
                                               _ptr(plan.num), _ptr(plan.ws), plan.nbytes, _stream()),
               'mpqe_sparse_rows_plan_owner')
    passes = (max(1, plan.table_rows.bit_length()) + 6) // 7
    _count(7 + 4 * passes)
    return plan

"""Debug tool: per-role cycle accounting of the second-generation tcgen05 layer kernel (layer_tc2.cu) on launches shaped
like the four layer launches of the bench step.  Needs the debug build:
    MPQE_BUILD_DIR=$PWD/mpqe_b200/_C_stats MPQE_NVCC_FLAGS=-DMPQE_TC_STATS python -m mpqe_b200.build
    MPQE_LIB_PATH=$PWD/mpqe_b200/_C_stats/libmpqe_b200.so python tools/debug_tc2_stats.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mpqe_b200 import _lib, ops

lib = _lib.load()
has_stats = hasattr(lib, 'mpqe_debug_set_stats2')
if has_stats:
    lib.mpqe_debug_set_stats2.argtypes = [ctypes.c_void_p]
dev = 'cuda:0'
torch.manual_seed(0)
w = torch.randn(40, 128, 128, device=dev) * 0.05
bias = torch.randn(8, 128, device=dev)
wp = ops.pack_weights(w)
B = int(os.environ.get('B', '4096'))
# (a, n, edges src->dst); anchors 0..a-1, target = a
T = [(1, 2, [(0, 1)]), (1, 3, [(0, 2), (2, 1)]), (1, 4, [(0, 3), (3, 2), (2, 1)]), (2, 3, [(0, 2), (1, 2)]),
     (3, 4, [(0, 3), (1, 3), (2, 3)]), (2, 4, [(0, 2), (1, 3), (3, 2)]), (2, 4, [(0, 3), (1, 3), (3, 2)])]


def launches():
    L1, L2, L3, L4 = [], [], [], []
    for gi, (a, n, edges) in enumerate(T):
        x = torch.randn(B, n, 128, device=dev)
        h = torch.empty(B, n, 128, device=dev)
        q = torch.empty(B, 128, device=dev)
        dq = torch.randn(B, 128, device=dev)
        dh = torch.empty(B, n, 128, device=dev)
        dx = torch.empty(B, n, 128, device=dev)
        t1 = [ops.Term(x, n, s, w[3 * gi + e], d, wp[3 * gi + e]) for e, (s, d) in enumerate(edges) if s < a]
        t1 += [ops.Term(x, n, i, w[39], i, wp[39]) for i in range(a)]
        L1.append(ops.Group(B, t1, n, h, n, epilogue=ops.EPI_RELU, bias=bias, bias_slot_stride=128))
        L2.append(ops.Group(B, [ops.Term(h, n, i, w[30 + i], 0, wp[30 + i]) for i in range(n)], 1, q, 1, out_slot_map=[0],
                            bias=bias[0], bias_scale=[float(n)]))
        L3.append(ops.Group(B, [ops.Term(dq, 1, 0, w[30 + i], i, wp[30 + i]) for i in range(n)], n, dh, n,
                            epilogue=ops.EPI_MASK, mask=h, mask_slots=n))
        t4 = [ops.Term(dh, n, d, w[3 * gi + e], s, wp[3 * gi + e]) for e, (s, d) in enumerate(edges) if s < a]
        t4 += [ops.Term(dh, n, i, w[39], i, wp[39]) for i in range(a)]
        L4.append(ops.Group(B, t4, a, dx, n, out_slot_map=list(range(a))))
    return [('fwd pass 0', L1), ('fwd pass 1 (collapsed, fused sum)', L2), ('bwd pass 1 (masked)', L3), ('bwd pass 0', L4)]


names = ['P wait empty', 'P data+stores', 'P fence+arrive', 'P wait data', 'M wait acc_empty', 'M wait full', 'M issue+commit',
         'stages', 'units', 'E wait acc_full', 'E work', 'CTA total', 'E tmem loads']
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dbg = int(os.environ.get('DBG', '0'))
if has_stats and dbg:
    lib.mpqe_debug_set_dbg2(dbg)
    print('#### debug bits %d (4 = no producer stores, 8 = no MMAs, 16 = no weight bulk copies)' % dbg)
for name, groups in launches():
    for _ in range(3):
        ops.layer_forward(groups, use_tensor_cores=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        s.record()
        ops.layer_forward(groups, use_tensor_cores=True)
        e.record()
        torch.cuda.synchronize()
        ts.append(1e3 * s.elapsed_time(e))
    nbytes, flops = ops._algorithmic(groups)
    print('== %s: %.1f us (min of 5, L2 flushed; all: %s)  %.0f GB/s algorithmic  %.1f TFLOP/s' % (
        name, min(ts), ' '.join('%.1f' % t for t in ts), nbytes / min(ts) / 1e3, flops / min(ts) / 1e6))
    if not has_stats:
        continue
    stats = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib.mpqe_debug_set_stats2(ctypes.c_void_p(stats.data_ptr()))
    flush.zero_()
    ops.layer_forward(groups, use_tensor_cores=True)
    torch.cuda.synchronize()
    lib.mpqe_debug_set_stats2(ctypes.c_void_p(0))
    st = stats.cpu().view(148, 16).numpy().astype(np.float64)
    for i, nm in enumerate(names):
        if nm != '-':
            print('   %-18s mean %9.0f  min %9.0f  max %9.0f' % (nm, st[:, i].mean(), st[:, i].min(), st[:, i].max()))
    stages = np.maximum(st[:, 7], 1)
    print('   per stage: P wait empty %.0f, P data+stores %.0f (of which waiting for the rows %.0f), P publish %.0f | M wait '
          'full %.0f, M issue %.0f | CTA total / stage %.0f' % tuple((st[:, i] / stages).mean() for i in (0, 1, 3, 2, 5, 6, 11)))
    units = np.maximum(st[:, 8], 1)
    print('   per unit: E wait %.0f, E work %.0f (tmem loads %.0f), M wait acc_empty %.0f' % tuple((st[:, i] / units).mean() for i in (9, 10, 12, 4)))

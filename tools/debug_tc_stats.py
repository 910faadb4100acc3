"""Debug tool: per-role cycle accounting of the tcgen05 layer kernel on the bench workload's first layer launch.
Needs a library built with MPQE_NVCC_FLAGS=-DMPQE_TC_STATS."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mpqe_b200 import _lib, ops

lib = _lib.load()
lib.mpqe_debug_set_stats.argtypes = [ctypes.c_void_p]
dev = 'cuda:0'
torch.manual_seed(0)
w = torch.randn(40, 128, 128, device=dev) * 0.05
bias = torch.randn(128, device=dev)
wp = ops.pack_weights(w)
groups = []
B = 4096
# the 7 templates' first pass: (n, edges)
T = [(2, [(0, 1)]), (3, [(0, 2), (2, 1)]), (4, [(0, 3), (3, 2), (2, 1)]), (3, [(0, 2), (1, 2)]),
     (4, [(0, 3), (1, 3), (2, 3)]), (4, [(0, 2), (1, 3), (3, 2)]), (4, [(0, 3), (1, 3), (3, 2)])]
for gi, (n, edges) in enumerate(T):
    x = torch.randn(B, n, 128, device=dev)
    out = torch.empty(B, n, 128, device=dev)
    terms = [ops.Term(x, n, s, w[3 * gi + e], d, wp[3 * gi + e]) for e, (s, d) in enumerate(edges)] + [ops.Term(x, n, i, w[39], i, wp[39]) for i in range(n)]
    if os.environ.get('STATS_SINGLE'):   # one term per output slot (the shape of the first forward pass / last input-gradient pass)
        terms = [ops.Term(x, n, i, w[39], i, wp[39]) for i in range(n)]
    groups.append(ops.Group(B, terms, n, out, n, epilogue=ops.EPI_RELU, bias=bias))
for _ in range(3):
    ops.layer_forward(groups, use_tensor_cores=True)
torch.cuda.synchronize()
stats = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.mpqe_debug_set_stats(ctypes.c_void_p(stats.data_ptr()))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
ops.layer_forward(groups, use_tensor_cores=True)
e.record()
torch.cuda.synchronize()
lib.mpqe_debug_set_stats(ctypes.c_void_p(0))
print('kernel us %.1f' % (1e3 * s.elapsed_time(e)))
s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s2.record()
ops.layer_forward(groups, use_tensor_cores=True)
e2.record()
torch.cuda.synchronize()
print('kernel us without stats buffer %.1f' % (1e3 * s2.elapsed_time(e2)))
st = stats.cpu().view(148, 16).numpy().astype(np.float64)
names = ['P wait empty', 'P data+stores', 'P issue loads', 'P fence+arrive', '-', 'E tmem ld+sts', 'M wait acc_empty', 'M wait full',
         'M issue+commit', 'stages', 'units', 'E wait acc_full', 'E work', 'E fetch+stores', 'CTA total', '-']
for i, nm in enumerate(names):
    if nm != '-':
        print('%-18s mean %9.0f  min %9.0f  max %9.0f' % (nm, st[:, i].mean(), st[:, i].min(), st[:, i].max()))
g0 = st[:, 15].min()
print('CTA start offsets ns: min %.0f max %.0f ; lifetimes ns mean %.0f max %.0f ; last end %.0f' % ((st[:,15]-g0).min(), (st[:,15]-g0).max(), st[:,4].mean(), st[:,4].max(), (st[:,15]-g0+st[:,4]).max()))
print('per stage: P wait %.0f, P data+stores %.0f, P loads %.0f, P publish %.0f | M wait full %.0f, M issue %.0f | total/stage %.0f' % tuple(
    (st[:, i] / st[:, 9]).mean() for i in (0, 1, 2, 3, 7, 8, 14)))
print('per unit: E wait %.0f, E work %.0f, M wait acc_empty %.0f' % tuple((st[:, i] / st[:, 10]).mean() for i in (11, 12, 6)))

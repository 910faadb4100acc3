"""Debug tool: event trace of CTA 0 of the tcgen05 layer kernel (python tools/debug_tc_trace.py)."""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpqe_b200 import _lib, ops

lib = _lib.load()
lib.mpqe_debug_set_trace.argtypes = [ctypes.c_void_p]
dev = 'cuda:0'
B, n = 128 * 148 * 2, 4
x = torch.randn(B, n, 128, device=dev)
w = torch.randn(5, 128, 128, device=dev) * 0.05
bias = torch.randn(128, device=dev)
out = torch.empty(B, n, 128, device=dev)
terms = [ops.Term(x, n, 0, w[0], 3), ops.Term(x, n, 1, w[1], 3), ops.Term(x, n, 2, w[2], 3), ops.Term(x, n, 3, w[4], 3),
         ops.Term(x, n, 0, w[4], 0), ops.Term(x, n, 1, w[4], 1), ops.Term(x, n, 2, w[4], 2)]
g = ops.Group(B, terms, n, out, n, epilogue=ops.EPI_RELU, bias=bias)
for _ in range(3):
    ops.layer_forward([g], use_tensor_cores=True)
torch.cuda.synchronize()
buf = torch.zeros(1 + 3 * 8000, dtype=torch.int64, device=dev)
lib.mpqe_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
ops.layer_forward([g], use_tensor_cores=True)
e.record()
torch.cuda.synchronize()
lib.mpqe_debug_set_trace(ctypes.c_void_p(0))
print('kernel ms', s.elapsed_time(e), 'units', B // 128 * n, 'stages total', B // 128 * 7 * 4)
t = buf.cpu()
cnt = int(t[0])
ev = t[1:1 + 3 * min(cnt, 8000)].view(-1, 3).tolist()
ev.sort(key=lambda r: r[2])
t0 = ev[0][2]
names = {10: 'P wait-empty', 11: 'P acquired', 12: 'P stored', 13: 'P loads issued', 14: 'P published', 20: 'M wait-full',
         21: 'M got-full', 22: 'M committed', 30: 'E wait-acc', 31: 'E got-acc', 32: 'E done'}
for tag, val, clk in ev[:120]:
    print('%8d  %-16s %d' % (clk - t0, names.get(tag, tag), val))
# summary: mean cycles between consecutive 'M committed'
mc = [r[2] for r in ev if r[0] == 22]
pc = [r[2] for r in ev if r[0] == 14]
if len(mc) > 10:
    print('mean cycles per stage (MMA commits): %.0f over %d stages' % ((mc[-1] - mc[5]) / (len(mc) - 6), len(mc)))
for a, b, name in ((10, 11, 'P wait empty'), (11, 12, 'P stores'), (12, 13, 'P load issue'), (13, 14, 'P publish'),
                   (20, 21, 'M wait full'), (21, 22, 'M issue+commit'), (30, 31, 'E wait acc'), (31, 32, 'E epilogue')):
    A = {r[1]: r[2] for r in ev if r[0] == a}
    Bv = {r[1]: r[2] for r in ev if r[0] == b}
    d = [Bv[k] - A[k] for k in A if k in Bv]
    if d:
        print('%-16s mean %7.0f  max %7d  n=%d' % (name, sum(d) / len(d), max(d), len(d)))

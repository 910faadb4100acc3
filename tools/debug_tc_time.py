"""Debug tool: true device time per layer launch (back-to-back launches, CPU preparation overlapped)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpqe_b200 import ops
dev = 'cuda:0'
torch.manual_seed(0)
w = torch.randn(40, 128, 128, device=dev) * 0.05
wp = ops.pack_weights(w)
bias = torch.randn(128, device=dev)
T = [(2, [(0, 1)]), (3, [(0, 2), (2, 1)]), (4, [(0, 3), (3, 2), (2, 1)]), (3, [(0, 2), (1, 2)]),
     (4, [(0, 3), (1, 3), (2, 3)]), (4, [(0, 2), (1, 3), (3, 2)]), (4, [(0, 3), (1, 3), (3, 2)])]

def make(B, packed, fused_sum=False):
    groups = []
    for gi, (n, edges) in enumerate(T):
        x = torch.randn(B, n, 128, device=dev)
        mk = lambda i: wp[i] if packed else None
        if fused_sum:
            out = torch.empty(B, 1, 128, device=dev)
            terms = [ops.Term(x, n, s, w[3 * gi + e], 0, mk(3 * gi + e)) for e, (s, d) in enumerate(edges)] + \
                    [ops.Term(x, n, i, w[39], 0, mk(39)) for i in range(n)]
            groups.append(ops.Group(B, terms, 1, out, 1, bias=bias, bias_scale=[float(n)]))
        else:
            out = torch.empty(B, n, 128, device=dev)
            terms = [ops.Term(x, n, s, w[3 * gi + e], d, mk(3 * gi + e)) for e, (s, d) in enumerate(edges)] + \
                    [ops.Term(x, n, i, w[39], i, mk(39)) for i in range(n)]
            groups.append(ops.Group(B, terms, n, out, n, epilogue=ops.EPI_RELU, bias=bias))
    return groups

def timeit(name, groups, tc, reps=30):
    for _ in range(3):
        ops.layer_forward(groups, use_tensor_cores=tc)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            ops.layer_forward(groups, use_tensor_cores=tc)
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / reps
    nb, fl = ops._algorithmic(groups)
    print('%-44s %8.1f us  %7.1f GB/s algorithmic  %6.1f TFLOP/s' % (name, us, nb / us / 1e3, fl / us / 1e6), flush=True)

for B in (4096, 32768):
    for fused in (False, True):
        tag = 'B=%d %s' % (B, 'sum-fused last pass' if fused else 'all-slot pass')
        timeit(tag + ' tcgen05 packed', make(B, True, fused), True)
        timeit(tag + ' tcgen05 unpacked', make(B, False, fused), True)
        timeit(tag + ' ffma', make(B, False, fused), False)
timeit('tiny launch (1 group of 128 queries) tcgen05', [make(128, True)[0]], True)
timeit('tiny launch (1 group of 128 queries) ffma', [make(128, False)[0]], False)

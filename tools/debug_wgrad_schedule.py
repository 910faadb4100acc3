"""Debug tool (CPU only): the (destination, chunk) units of the bench step's weight-gradient launches and how their
stages spread over 148 CTAs, replicating the host logic of mpqe_layer_wgrad (layer_simt.cu) / layer_wgrad_tc_launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mpqe_b200 import ops, synthetic
from tests import emulator

KC, SMS = 32, 148


class MP(object):
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


emulator.install(MP())
calls = []
orig = ops.layer_wgrad


def spy(groups, grad_operands, dests, *a, **kw):
    calls.append((groups, dests))
    return orig(groups, grad_operands, dests, *a, **kw)


ops.layer_wgrad = spy
import mpqe_b200.model as M
B = int(os.environ.get('B', '256'))
run = bench.Run('am_sum', B, torch.device('cpu'), 0)
ts = run.ts
res = ts.forward_backward([ts.to_device(hb) for hb in run.host])
scale = 4096 // B
for groups, dests in calls:
    if all(g.num_queries == 1 for g in groups):
        continue
    nd = len(dests)
    hint = int(os.environ.get('HINT', str(max(SMS - nd, SMS // 2))))
    weight = []
    for (m_fwd, dm, acc) in dests:
        w = sum(g.num_queries * scale for g in groups for t in g.terms if t.m.data_ptr() == m_fwd.data_ptr())
        weight.append(w)
    total = sum(weight)
    # min-max apportionment of hint + destinations units in stages (mpqe_layer_wgrad)
    caps = [min(int(w / (4 * 16)) + 1, 256) for w in weight]
    tof = [[sum(1 for t in g.terms if t.m.data_ptr() == m_fwd.data_ptr()) for g in groups] for (m_fwd, _, _) in dests]

    def unit_cost(j, c):
        st = 0
        for g, k in zip(groups, tof[j]):
            if k:
                Bq = g.num_queries * scale
                per = -(-Bq // c)
                per = min(-(-per // KC) * KC, Bq)
                st += k * (-(-per // KC))
        return st

    def chunks_for(j, limit):
        for c in range(1, caps[j] + 1):
            if unit_cost(j, c) <= limit:
                return c
        return caps[j]
    lo, hi = 1, max(unit_cost(j, 1) for j in range(nd))
    while lo < hi:
        mid = (lo + hi) // 2
        if sum(chunks_for(j, mid) for j in range(nd)) <= hint + nd:
            hi = mid
        else:
            lo = mid + 1
    chunks = [chunks_for(j, lo) for j in range(nd)]
    units = []
    for j, (m_fwd, dm, acc) in enumerate(dests):
        for c in range(chunks[j]):
            steps = 0
            for g in groups:
                Bq = g.num_queries * scale
                per = -(-Bq // chunks[j])
                per = -(-per // KC) * KC
                qb, qe = min(per * c, Bq), min(per * c + per, Bq)
                tiles = -(-(qe - qb) // KC)
                steps += tiles * sum(1 for t in g.terms if t.m.data_ptr() == m_fwd.data_ptr())
            units.append(steps)
    units.sort(reverse=True)
    load = [0] * min(SMS, len(units))
    for u in units:
        i = int(np.argmin(load))
        load[i] += u
    print('dests %d units %d total stages %d | per CTA: mean %.1f max %d min %d | unit stages: max %d median %d min %d' % (
        nd, len(units), sum(units), np.mean(load), max(load), min(load), units[0], units[len(units) // 2], units[-1]))

"""Shared plumbing for the parity tests: rebuild a golden case (graph, parameters, ids) from its fixture."""
import glob
import json
import os

import numpy as np
import torch

from mpqe_b200 import synthetic
from oracle import mpqe_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def _tuplify(x):
    return tuple(_tuplify(i) for i in x) if isinstance(x, list) else x


class GoldenCase(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
        self.z = z
        self.name = name
        self.query_type = str(z['query_type'])
        self.rels = _tuplify(json.loads(str(z['rels_json'])))
        self.cfg = O.Config(readout=str(z['readout']), num_layers=int(z['num_layers']),
                            adaptive=bool(z['adaptive']), shared_layers=bool(z['shared_layers']),
                            weight_decay=float(z['weight_decay']),
                            scatter_op=str(z['scatter_op']) if 'scatter_op' in z.files else 'add')
        self.d = int(z['d'])
        self.kg = synthetic.make_kg('tiny', seed=int(z['kg_seed']))
        rels, _, node_maps = self.kg.raw()
        self.schema_rels, self.node_maps = rels, node_maps
        self.mode_ids, self.rel_ids = O.schema_ids(rels)
        self.id2row = O.id_to_row(node_maps)
        self.params = O.init_params(rels, node_maps, self.cfg, d=self.d, seed=int(z['param_seed']))
        self.spec = O.formula_spec(self.query_type, self.rels)
        self.anchor_ids = torch.from_numpy(z['anchor_ids'])
        self.targets = torch.from_numpy(z['targets'])

    def grads(self):
        """{param name: dense numpy gradient} for fixtures that store gradients."""
        out = {}
        for k in self.z.files:
            if k.startswith('grad:'):
                out[k[5:]] = self.z[k]
            elif k.startswith('grad_idx:'):
                name = k[9:]
                dense = np.zeros(tuple(self.params[name].shape), dtype=np.float32)
                dense[self.z[k]] = self.z['grad_rows:' + name]
                out[name] = dense
        return out


def assert_close(actual, expected, rtol, atol, what=''):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    err = np.abs(actual - expected)
    tol = atol + rtol * np.abs(expected)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError('%s: max violation at %s: got %r want %r (|err|=%.3e, tol=%.3e)' % (
            what, i, actual[i], expected[i], err[i], tol[i]))


# ---- stated fp32 tolerances ---------------------------------------------------------------
# Gradients are sums over the batch: a tensor's error is measured against the tensor itself.
#   (i)  relative Frobenius error  ||got - want||_F / ||want||_F :
#            strict-fp32 FFMA kernels ........ <= 1e-5
#            tcgen05 3xTF32 kernels .......... <= 5e-5   (the split drops the lo*lo products)
#   (ii) largest single deviation  max|got - want| / max|want|  <= 50 x the bound of (i).
# Why two tiers: the network has ReLUs.  At B = 4096 there are millions of pre-activations per batch, some within
# fp32 rounding of zero; for those the sign -- hence whether a gradient flows through that one element -- depends on
# the summation order, which legitimately differs from the CPU reference.  Such a flip moves single entries of single
# rows by up to ~2e-4 of the tensor's maximum (observed: 2.1e-4 on MUTAG / concat at B = 4096, FFMA) without moving
# the tensor as a whole; on tensors without such a flip the largest deviation itself stays below the bound of (i)
# (observed 7e-7 ... 6e-6, see OBSERVED -> gpurun_out/parity_observed.json).
GRAD_TOL = {'ffma': 1e-5, 'tcgen05': 5e-5}
MAX_FACTOR = 50.0
OBSERVED = {}     # what -> largest (frobenius, max) errors seen in this session (written out by conftest at exit)


def assert_grad_close(got, want, mode, what, fro_tol=None, max_tol=None, median_tol=None):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.isfinite(got).all(), '%s: non-finite gradient' % what
    diff = got - want
    fro = float(np.sqrt((diff * diff).sum()) / max(np.sqrt((want * want).sum()), 1e-30))
    mx = float(np.abs(diff).max() / max(np.abs(want).max(), 1e-30))
    key = '%s %s' % (mode, what.split(' ')[0])
    old = OBSERVED.get(key, (0.0, 0.0))
    OBSERVED[key] = (max(old[0], fro), max(old[1], mx))
    tol = GRAD_TOL[mode] if fro_tol is None else fro_tol
    mtol = MAX_FACTOR * GRAD_TOL[mode] if max_tol is None else max_tol
    assert fro <= tol, '%s: relative Frobenius error %.3e exceeds the %s bound %.1e' % (what, fro, mode, tol)
    assert mx <= mtol, '%s: max|err| = %.3e * max|g| exceeds the %s bound %.1e' % (what, mx, mode, mtol)
    if median_tol is not None:
        # with relaxed outer bounds (configurations whose gradient is discontinuous in rounding-level events, see the
        # callers) the BULK of the tensor must still meet the tight bound: a discontinuity event perturbs a minority of
        # the entries, a systematic error (scale, a dropped chunk, a wrong operand) moves most of them
        nz = want != 0
        med = float(np.median(np.abs(diff[nz])) / max(np.abs(want).max(), 1e-30)) if nz.any() else 0.0
        assert med <= median_tol, '%s: median|err| = %.3e * max|g| exceeds %.1e' % (what, med, median_tol)

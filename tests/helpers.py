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
# Gradients are sums over the batch: the error of a tensor is measured against that tensor's largest entry.
#   strict-fp32 FFMA kernels ........ max|err| <= 1e-5 * max|g|
#   tcgen05 3xTF32 kernels .......... max|err| <= 5e-5 * max|g|   (the split drops the lo*lo products)
GRAD_TOL = {'ffma': 1e-5, 'tcgen05': 5e-5}
OBSERVED = {}     # what -> largest normalised error seen in this session (written out by conftest at exit)


def assert_grad_close(got, want, mode, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    key = '%s %s' % (mode, what.split(' ')[0])
    OBSERVED[key] = max(OBSERVED.get(key, 0.0), float(err))
    assert np.isfinite(got).all(), '%s: non-finite gradient' % what
    assert err <= GRAD_TOL[mode], '%s: max|err| = %.3e * max|g| exceeds the %s bound %.1e' % (
        what, err, mode, GRAD_TOL[mode])

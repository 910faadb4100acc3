import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpqe_b200 import synthetic
from mpqe_b200.graph import Formula
from mpqe_b200.train_step import HostBatch, TrainStep
from oracle import mpqe_oracle as O
from tests.model_utils import build_model
DEV = 'cuda:0'
kg = synthetic.make_kg('aifb', seed=3)
rels, _, node_maps = kg.raw()
cfg = O.Config(readout='sum', num_layers=2)
params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
frng = np.random.RandomState(0)
formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in synthetic.QUERY_TYPES]
rng = np.random.RandomState(1)
h1 = [HostBatch(f, *[torch.from_numpy(x) for x in synthetic.sample_id_batch(kg, f, 500, rng)]) for f in formulas]
ts = TrainStep(model)
def views(G):
    out = {}
    for i, v in enumerate(G.dw): out['dw%d' % i] = v.clone()
    for i, v in enumerate(G.droot): out['droot%d' % i] = v.clone()
    for i, v in enumerate(G.dbias): out['dbias%d' % i] = v.clone()
    out['dmode'] = G.dmode.clone()
    return out
runs = []
for k in range(3):
    r = ts.forward_backward([ts.to_device(hb) for hb in h1])
    torch.cuda.synchronize()
    runs.append((views(r.dense), r.losses.clone()))
ts.capture(h1)
for k in range(2):
    r = ts.replay(h1); torch.cuda.synchronize()
    runs.append((views(r.dense), r.losses.clone()))
names = ['eager0', 'eager1', 'eager2', 'graph0', 'graph1']
for i in range(1, len(runs)):
    diffs = {k: float((runs[i][0][k] - runs[0][0][k]).abs().max()) for k in runs[0][0]}
    bad = {k: v for k, v in diffs.items() if v != 0}
    print(names[i], 'vs eager0:', 'identical' if not bad else bad, 'loss equal', bool(torch.equal(runs[i][1], runs[0][1])))

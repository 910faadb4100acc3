"""Debug tool: is a 40-step fused training run (concat readout, fused Adam) independent of kernel timing?  Prints a hash
of the loss curve and of the final parameters; run it twice, and once with CUDA_LAUNCH_BLOCKING=1 (every launch
serialised, as under compute-sanitizer): all hashes must agree."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mpqe_b200 import synthetic
from mpqe_b200.graph import Formula
from mpqe_b200.train_step import HostBatch, TrainStep
from oracle import mpqe_oracle as O
from tests.model_utils import build_model

DEV = 'cuda:0'
readout = os.environ.get('READOUT', 'concat')
kg = synthetic.make_kg('tiny', seed=5)
rels, _, node_maps = kg.raw()
cfg = O.Config(readout=readout, num_layers=2, weight_decay=1e-3)
params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
ts = TrainStep(model)
frng, rng = np.random.RandomState(0), np.random.RandomState(1)
formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in synthetic.QUERY_TYPES]
pool = [[synthetic.sample_id_batch(kg, f, 24, rng) for f in formulas] for _ in range(3)]
curve = []
for step in range(40):
    data = pool[step % 3]
    batches = [ts.to_device(HostBatch(f, *[torch.from_numpy(x) for x in d])) for f, d in zip(formulas, data)]
    ts.catchup_rows(batches)
    res = ts.forward_backward(batches)
    curve.append(res.total.item())
    ts.adam_step(res, lr=0.01)
ts.catchup_rows(None)
torch.cuda.synchronize()
h = hashlib.sha1(np.array(curve, dtype=np.float32).tobytes()).hexdigest()[:12]
hp = hashlib.sha1(b''.join(v.detach().cpu().numpy().tobytes() for k, v in sorted(model.state_dict().items()))).hexdigest()[:12]
print('blocking=%s readout=%s curve %s params %s  first %.6f last %.6f step36 %.6f' % (
    os.environ.get('CUDA_LAUNCH_BLOCKING', '0'), readout, h, hp, curve[0], curve[-1], curve[36]))

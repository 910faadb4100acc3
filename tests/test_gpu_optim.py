"""Regulariser, optimiser and negative sampling of the fused training step (csrc/optim.cu) against torch on the CPU:
the L2 term of margin_loss (model.py:487-492), torch.optim.Adam (train.py:86-88) on dense tensors and -- through the
lazy zero-gradient catch-up -- on the entity tables from row-sparse gradients, and the counter-based negative draw
(model.py:470-476)."""
import numpy as np
import pytest
import torch

from mpqe_b200 import ops, synthetic
from oracle import mpqe_oracle as O
from tests.helpers import assert_close
from tests.model_utils import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_l2_reg_matches_autograd():
    torch.manual_seed(0)
    params = [torch.randn(256, 128), torch.randn(128), torch.randn(128, 128), torch.zeros(128)]
    ref = [p.clone().requires_grad_(True) for p in params]
    wd, scale = 1e-3, 2.5
    reg = wd * sum(torch.norm(p) for p in ref)
    (scale * reg).backward()
    grads = [torch.full_like(p, 0.25).to(DEV) for p in params]
    losses = torch.tensor([1.0, 2.0, 3.0], device=DEV)
    norms = torch.empty(4, device=DEV)
    with torch.cuda.device(0):
        ops.l2_reg([p.to(DEV) for p in params], grads, wd, scale, losses=losses, norms=norms)
    assert_close(losses.cpu().numpy(), np.array([1.0, 2.0, 3.0]) + float(reg), 1e-6, 1e-7, 'losses')
    assert_close(norms.cpu().numpy(), np.array([float(torch.norm(p)) for p in params]), 1e-6, 0, 'norms')
    for g, r in zip(grads, ref):
        want = 0.25 + (r.grad if r.grad is not None else 0)
        assert_close(g.cpu().numpy(), np.broadcast_to(np.asarray(want), g.shape), 1e-6, 1e-9, 'grad')
    assert bool(torch.isfinite(grads[3]).all())      # ||0|| has the zero sub-gradient, not NaN


def test_adam_multi_matches_torch_adam():
    torch.manual_seed(1)
    shapes = [(38, 128, 128), (128, 128), (128,), (5, 128)]
    ref = [torch.nn.Parameter(torch.randn(s) * 0.1) for s in shapes]
    opt = torch.optim.Adam(ref, lr=0.01)
    mine = [p.detach().clone().to(DEV) for p in ref]
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in mine]
    clock = ops.adam_state(torch.device(DEV))
    for step in range(1, 8):
        grads = [torch.randn(s) * (0.5 if step % 2 else 1e-3) for s in shapes]
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        opt.step()
        with torch.cuda.device(0):
            ops.adam_tick(clock, 0.01, 0.9, 0.999)
            ops.adam_multi([(p, g.to(DEV), m, v) for p, g, (m, v) in zip(mine, grads, state)], 0.01, 0.9, 0.999, 1e-8,
                           state=clock)
    assert int(clock[0]) == 7
    for p, r in zip(mine, ref):
        assert_close(p.cpu().numpy(), r.detach().numpy(), 2e-6, 1e-7, 'adam param')


def test_row_adam_equals_dense_adam_trajectory():
    """Row-sparse Adam with lazy catch-up == dense torch.optim.Adam fed the same gradients as dense tensors with zero
    rows: (i) every row, at the moment a step reads it, (ii) the whole table after the final flush."""
    torch.manual_seed(2)
    rng = np.random.RandomState(0)
    sizes = [300, 513]
    offs = [0, 300]
    ref = [torch.nn.Parameter(torch.randn(n, 128) * 0.1) for n in sizes]
    opt = torch.optim.Adam(ref, lr=0.01)
    tabs = [p.detach().clone().to(DEV) for p in ref]
    ra = ops.RowAdam([(t, o) for t, o in zip(tabs, offs)], lr=0.01)
    clock = ops.adam_state(torch.device(DEV))
    total = sum(sizes)
    for step in range(1, 41):
        k = int(rng.randint(5, 60))
        touched = np.sort(rng.choice(total, size=k, replace=False))        # unique, ascending (what the combine emits)
        reads = np.concatenate([touched, rng.choice(touched, size=7)])      # the ids a step reads: with duplicates
        rows = torch.randn(k, 128)
        with torch.cuda.device(0):
            ra.catchup(torch.from_numpy(reads).to(DEV), state=clock)
        got = torch.cat(tabs)[torch.from_numpy(touched).to(DEV)].cpu()
        want = torch.cat([p.detach() for p in ref])[torch.from_numpy(touched)]
        assert_close(got.numpy(), want.numpy(), 3e-6, 1e-7, 'rows read at step %d' % step)
        dense = torch.zeros(total, 128)
        dense[torch.from_numpy(touched)] = rows
        for p, o, n in zip(ref, offs, sizes):
            p.grad = dense[o:o + n].clone()
        opt.step()
        with torch.cuda.device(0):
            ops.adam_tick(clock, 0.01, 0.9, 0.999)
            ra.apply(torch.from_numpy(touched).to(DEV), rows.to(DEV), torch.tensor([k], device=DEV), state=clock)
    with torch.cuda.device(0):
        ra.catchup(None, state=clock)
    for t, p in zip(tabs, ref):
        assert_close(t.cpu().numpy(), p.detach().numpy(), 3e-6, 1e-7, 'table after flush')


def test_sample_negatives_counter_based():
    rng = np.random.RandomState(3)
    lengths = rng.randint(0, 9, size=500)
    lengths[:5] = [0, 1, 1, 8, 3]
    offsets = np.zeros(501, dtype=np.int64)
    offsets[1:] = np.cumsum(lengths)
    cand = rng.randint(0, 10 ** 6, size=int(offsets[-1])).astype(np.int64)
    c, o = torch.from_numpy(cand).to(DEV), torch.from_numpy(offsets).to(DEV)
    with torch.cuda.device(0):
        a = ops.sample_negatives(c, o, 700, seed=11, step=4, first_query=450)     # wraps around the 500 queries
        b = ops.sample_negatives(c, o, 700, seed=11, step=4, first_query=450)
        d = ops.sample_negatives(c, o, 700, seed=11, step=5, first_query=450)
    a, b, d = a.cpu().numpy(), b.cpu().numpy(), d.cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, d)
    for i in range(700):
        q = (450 + i) % 500
        seg = cand[offsets[q]:offsets[q + 1]]
        assert (a[i] == -1 and len(seg) == 0) or a[i] in seg
    # shared candidate list (1-chain: the whole target mode), roughly uniform
    pool = torch.arange(1000, 1016, device=DEV)
    with torch.cuda.device(0):
        s = ops.sample_negatives(pool, None, 64000, seed=1, step=0).cpu().numpy()
    counts = np.bincount(s - 1000, minlength=16)
    assert counts.min() > 3600 and counts.max() < 4400


def test_train_step_adam_follows_oracle_training():
    """40 optimiser steps of the fused step + fused Adam (row-sparse over the entity tables, lazily caught up) against
    the oracle's margin_loss + dense torch.optim.Adam on the CPU, same batches: the loss curves coincide."""
    from mpqe_b200.graph import Formula
    from mpqe_b200.train_step import HostBatch, TrainStep
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='concat', num_layers=2, weight_decay=1e-3)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    mode_ids, rel_ids = O.schema_ids(rels)
    id2row = O.id_to_row(node_maps)
    model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
    ts = TrainStep(model)
    frng, rng = np.random.RandomState(0), np.random.RandomState(1)
    formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in synthetic.QUERY_TYPES]
    ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(ref.values()), lr=0.01)
    curve_ref, curve = [], []
    # three batch sets taken in turn: the model can fit them (the loss falls), and every entity row sits out two of
    # three steps, which is what the lazy catch-up has to get right
    pool = [[synthetic.sample_id_batch(kg, f, 24, rng) for f in formulas] for _ in range(3)]
    for step in range(40):
        data = pool[step % 3]
        opt.zero_grad()
        total = 0
        for f, (a, t, n) in zip(formulas, data):
            spec = O.formula_spec(f.query_type, f.rels)
            a_ids, var_ids, ei, et, batch = O.query_graph(spec, a.tolist(), rel_ids, mode_ids)
            total = total + O.margin_loss(ref, cfg, spec, a_ids, var_ids, ei, et, batch, id2row, torch.from_numpy(t),
                                          torch.from_numpy(n))
        total.backward()
        opt.step()
        curve_ref.append(float(total))
        batches = [ts.to_device(HostBatch(f, *[torch.from_numpy(x) for x in d])) for f, d in zip(formulas, data)]
        ts.catchup_rows(batches)
        res = ts.forward_backward(batches)
        curve.append(float(res.total))
        ts.adam_step(res, lr=0.01)
    ts.catchup_rows(None)
    assert curve_ref[-1] < 0.9 * curve_ref[0], 'the reference run should be learning: %r' % (curve_ref[::8],)
    # The GPU run is bit-reproducible and independent of kernel timing (tools/debug_train_determinism.py); the CPU
    # reference is not (threaded torch reductions), and Adam amplifies rounding-level differences step by step: the
    # curves are held to 1e-3 over the first 20 steps and to 1e-2 over all 40 (observed under compute-sanitizer, which
    # changes the host's threading: 1.6e-3 at step 36).
    assert_close(np.array(curve[:20]), np.array(curve_ref[:20]), 1e-3, 1e-3, 'loss curve, first 20 steps')
    assert_close(np.array(curve), np.array(curve_ref), 1e-2, 1e-3, 'loss curve')
    # Adam divides by sqrt(v): an element whose gradient is at the level of fp32 rounding moves by lr per step in a
    # direction set by rounding noise, so parameters are compared as a whole (the loss curve above is the sharp check):
    # all but a few per cent of the entries of every tensor stay within 1 % of the tensor's range.
    sd = model.state_dict()
    for k, v in ref.items():
        want = v.detach().numpy()
        err = np.abs(sd[k].cpu().numpy() - want)
        frac = float(np.mean(err > 1e-2 * np.abs(want).max()))
        assert frac < 0.05, 'param %s: %.4f of the entries differ' % (k, frac)

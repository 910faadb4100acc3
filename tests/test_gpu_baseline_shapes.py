"""Parity of the CUDA path with the CPU oracle ON THE CONFIGURATIONS BASELINE.json NAMES, at the batch sizes it names:

  config 2  MPQE-TM  (--readout mp --adaptive, 3 layers) on the AIFB-shaped graph
  config 3  MPQE-max and MPQE-concat (2 layers) on the MUTAG-shaped graph
  config 4  MPQE-sum (2 layers) on the AM-shaped graph (the bench workload)

each with all 7 query types at B = 512 and B = 4096 queries per type, on the strict-fp32 FFMA kernels and on the
tcgen05 3xTF32 kernels.  Compared: the loss of every formula batch, EVERY parameter gradient (entity tables
included, made dense) and the set of touched entity rows.  The oracle (reference arithmetic: per-edge weight gather
+ bmm + scatter, two encoder passes) is evaluated once per case and shared by both kernel modes.
Tolerances: tests/helpers.py (GRAD_TOL) and DESIGN.md section 5.
"""
import functools

import numpy as np
import pytest
import torch

from mpqe_b200 import synthetic
from mpqe_b200.graph import Formula
from oracle import mpqe_oracle as O
from tests.helpers import GRAD_TOL, assert_close, assert_grad_close
from tests.model_utils import build_model, model_grads, oracle_loss_and_grads, queries_from_ids, train_step_grads

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

# name -> (graph shape, readout, layers, adaptive)
CASES = {
    'aifb_tm': ('aifb', 'mp', 3, True),
    'mutag_max': ('mutag', 'max', 2, False),
    'mutag_concat': ('mutag', 'concat', 2, False),
    'am_sum': ('am', 'sum', 2, False),
}


@functools.lru_cache(maxsize=None)
def world(case):
    shape, readout, layers, adaptive = CASES[case]
    kg = synthetic.make_kg(shape, seed=0)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout=readout, num_layers=layers, adaptive=adaptive, weight_decay=1e-3)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    mode_ids, rel_ids = O.schema_ids(rels)
    frng = np.random.RandomState(0)
    formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in synthetic.QUERY_TYPES]
    return kg, cfg, params, mode_ids, rel_ids, O.id_to_row(node_maps), formulas


@functools.lru_cache(maxsize=None)
def oracle_case(case, B):
    """[(formula, anchors, targets, negatives, oracle loss, oracle gradients)] for the 7 query types."""
    kg, cfg, params, mode_ids, rel_ids, id2row, formulas = world(case)
    rng = np.random.RandomState(B)
    out = []
    for f in formulas:
        a, t, n = synthetic.sample_id_batch(kg, f, B, rng)
        spec = O.formula_spec(f.query_type, f.rels)
        loss, grads = oracle_loss_and_grads(params, cfg, spec, torch.from_numpy(a), rel_ids, mode_ids, id2row,
                                            torch.from_numpy(t), torch.from_numpy(n))
        out.append((f, a, t, n, loss, grads))
    return out


@pytest.fixture(params=['ffma', 'tcgen05'])
def mode(request):
    from mpqe_b200 import _lib, ops
    tc = request.param == 'tcgen05'
    if tc and not _lib.load().mpqe_b200_has_tcgen05():
        pytest.skip('library built without tcgen05 kernels')
    ops.set_tensor_cores(tc)
    yield request.param
    ops.set_tensor_cores(False)


def touched_rows(model, ts, formula, anchors, targets, negatives):
    """Global table rows (table offset + row) a formula batch touches: anchors, targets, negatives."""
    id2row = model.enc.node_maps.cpu().numpy()
    rows = [ts.table_offsets[m] + id2row[anchors[:, i]] for i, m in enumerate(formula.anchor_modes)]
    rows += [ts.table_offsets[formula.target_mode] + id2row[targets], ts.table_offsets[formula.target_mode] + id2row[negatives]]
    return np.concatenate(rows)


def discontinuity_events(case, B, batches, mode):
    """Number of ReLU signs / arg-maxes of the CUDA step that differ from the oracle's, after asserting that each of
    them is a near-tie in the oracle's own values and that they are a vanishing fraction of all decisions."""
    import torch.nn.functional as F
    kg, cfg, params, mode_ids, rel_ids, id2row, formulas = world(case)
    rel = 2e-5 if mode == 'tcgen05' else 4e-6
    events = total = 0
    for b_, (f, a, t, n, loss, grads) in zip(batches, oracle_case(case, B)):
        job = b_.job
        spec = O.formula_spec(f.query_type, f.rels)
        a_ids, var_ids, ei, et, batch = O.query_graph(spec, a.tolist(), rel_ids, mode_ids)
        conv_out, hidden = [], []
        conv, mlp = O.rgcn_conv, O.mlp

        def spy_conv(*args, **kw):
            conv_out.append(conv(*args, **kw))
            return conv_out[-1]

        def spy_mlp(x, p, prefix='readout.layers.'):
            hidden.append(F.linear(x, p[prefix + '0.weight'], p[prefix + '0.bias']))
            return mlp(x, p, prefix)

        O.rgcn_conv, O.mlp = spy_conv, spy_mlp
        try:
            with torch.no_grad():
                _, arg_want = O.encode_queries(params, cfg, spec, a_ids, var_ids, ei, et, batch, id2row, want_argmax=True)
        finally:
            O.rgcn_conv, O.mlp = conv, mlp
        # ReLU between the passes: job.acts[k + 1] = relu(pass k), every pass but the last
        # (the target-message readout prunes the slots that cannot reach the target: nothing to compare there)
        signs = [(job.acts[k + 1], conv_out[k]) for k in range(len(conv_out) - 1)] if cfg.readout != 'mp' else []
        if hidden:
            signs.append((job.u, hidden[0]))
        for got_act, pre in signs:
            got_pos = (got_act.cpu().reshape(pre.shape) > 0)
            differ = got_pos != (pre > 0)
            total += pre.numel()
            events += int(differ.sum())
            if differ.any():
                worst = float(pre[differ].abs().max())
                assert worst <= rel * float(pre.abs().max()), ('ReLU sign differs on a clearly non-zero value', worst)
        if cfg.readout == 'max':
            got_arg, z = job.argmax.cpu(), job.z.cpu().reshape(-1, 128)
            differ = (got_arg != arg_want).nonzero()
            total += got_arg.numel()
            events += differ.shape[0]
            for q_, c_ in differ.tolist():
                v_got, v_want = float(z[got_arg[q_, c_], c_]), float(z[arg_want[q_, c_], c_])
                assert abs(v_got - v_want) <= 4e-6 * abs(v_want) + 2e-7, ('argmax differs on a clear maximum', q_, c_)
    assert events <= 1e-4 * total, 'too many discontinuity events: %d of %d' % (events, total)
    return events


@pytest.mark.parametrize('B', [512, 4096])
@pytest.mark.parametrize('case', sorted(CASES))
def test_fused_step_vs_oracle(case, B, mode):
    """The fused training step (all 7 formula batches in one pass: the bench path) against the oracle."""
    from mpqe_b200.train_step import HostBatch, TrainStep
    kg, cfg, params, mode_ids, rel_ids, id2row, formulas = world(case)
    model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
    ts = TrainStep(model)
    want_losses, want, host, touched = [], {}, [], []
    for f, a, t, n, loss, grads in oracle_case(case, B):
        host.append(HostBatch(f, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n)))
        want_losses.append(loss)
        for k, g in grads.items():
            want[k] = want.get(k, 0) + g
        touched.append(touched_rows(model, ts, f, a, t, n))
    batches = [ts.to_device(hb) for hb in host]
    res = ts.forward_backward(batches)
    torch.cuda.synchronize()
    atol = 1e-5 if mode == 'tcgen05' else 2e-6        # cosine scores live in [-1, 1]; the loss is their batch mean
    assert_close(res.losses.cpu().numpy(), np.array(want_losses, dtype=np.float32), 1e-5, atol, 'losses')
    got, uid = train_step_grads(ts, model, res, DEV)
    # integer artefact: the combined row set is exactly the set of rows the batches touch, ascending, no duplicates
    assert np.array_equal(uid.cpu().numpy(), np.unique(np.concatenate(touched)))
    # Two discontinuities sit on this path.  (1) ReLU (between the passes, and inside the readout MLP of `concat`): a
    # pre-activation within fp32 rounding of zero lets the gradient through or not depending on the summation order.
    # (2) The max readout routes each feature's gradient to the node attaining the maximum: where two nodes' values
    # agree to rounding, a different (equally valid) choice moves O(1) of that feature's gradient to another node.
    # At B = 4096 a handful of the ~15 M pre-activations / 3.7 M arg-maxes are such events.  They are FOUND here, not
    # tolerated blindly: every ReLU sign and every arg-max that differs from the oracle's must be a near-tie (checked
    # element by element), they must be a vanishing fraction, and only a batch set that has one is compared with the
    # relaxed outer bound; without any, the tight bound applies.
    events = 0
    if B >= 4096:
        events = discontinuity_events(case, B, batches, mode)
    fro_tol, max_tol = (2e-3, 2e-2) if events else (None, None)
    for name, g in want.items():
        assert got.get(name) is not None, name
        assert_grad_close(got[name].detach().cpu().numpy(), g, mode, '%s:B%d grad %s' % (case, B, name), fro_tol, max_tol)


@pytest.mark.parametrize('case', sorted(CASES))
def test_margin_loss_api_vs_oracle(case, mode):
    """The reference-shaped API (`margin_loss` + autograd `backward`, one formula batch per call) at B = 512."""
    kg, cfg, params, mode_ids, rel_ids, id2row, formulas = world(case)
    model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
    atol = 1e-5 if mode == 'tcgen05' else 2e-6
    for f, a, t, n, want_loss, want in oracle_case(case, 512):
        queries = queries_from_ids(f.query_type, f.rels, a, t)
        model.zero_grad()
        loss = model.margin_loss_ids(queries[0].formula, queries, torch.from_numpy(t), torch.from_numpy(n))
        assert_close(loss.item(), want_loss, 1e-5, atol, '%s %s loss' % (case, f.query_type))
        loss.backward()
        got = model_grads(model)
        for k, g in want.items():
            assert_grad_close(got[k], g, mode, '%s:api grad %s (%s)' % (case, k, f.query_type))

"""Callers of the hot path (SURVEY section 8 row f): eval metrics and the training loop, on the GPU, against the oracle."""
import logging

import numpy as np
import pytest
import torch

from mpqe_b200 import data_utils, eval as mp_eval, synthetic, train_helpers, utils
from mpqe_b200.graph import Query
from oracle import mpqe_oracle as O
from tests.helpers import assert_close
from tests.model_utils import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def setup(readout='sum', seed=5, per_formula=40):
    kg = synthetic.make_kg('tiny', seed=seed)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout=readout, num_layers=2)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    model = build_model(kg.raw(), cfg, params, DEV)
    qsets = synthetic.make_query_sets(kg, queries_per_formula=per_formula, formulas_per_type=1, seed=2, num_neg=6)
    return kg, cfg, params, model, qsets


def test_auc_and_percentile_match_oracle_scores():
    kg, cfg, params, model, qsets = setup()
    rels, _, node_maps = kg.raw()
    mode_ids, rel_ids = O.schema_ids(rels)
    id2row = O.id_to_row(node_maps)
    frm_rels, raw = qsets['3-inter_chain'][0]
    queries = [Query.deserialize(r) for r in raw]
    formula = queries[0].formula
    test_queries = {formula: queries}
    perc = utils.eval_perc_queries(test_queries, model, batch_size=16)
    # oracle: same scores on CPU, same percentile definition
    spec = O.formula_spec('3-inter_chain', frm_rels)
    want = []
    for off in range(0, len(queries), 16):
        batch = queries[off:off + 16]
        anchors = [q.anchor_nodes for q in batch]
        a_ids, var_ids, ei, et, b = O.query_graph(spec, anchors, rel_ids, mode_ids)
        lengths = [len(q.neg_samples) for q in batch]
        negs = [n for q in batch for n in q.neg_samples]
        with torch.no_grad():
            s = O.forward_scores(params, cfg, spec, a_ids, var_ids, ei, et, b, id2row,
                                 torch.tensor([q.target_node for q in batch]), torch.tensor(negs), lengths).numpy()
        l, r = O.rank_counts(s[:len(batch)], s[len(batch):], lengths)
        want.extend(O.percentile_from_counts(l, r, lengths))
    assert abs(perc - float(np.mean(want))) < 1.0   # percentiles move in steps of 100/len; scores agree to 1e-5
    auc, per_formula = utils.eval_auc_queries(test_queries, model, batch_size=16)
    assert 0.0 <= auc <= 1.0 and formula in per_formula
    assert utils.auc_from_scores([1, 1, 0, 0], [0.9, 0.4, 0.5, 0.1]) == O.auc([1, 1, 0, 0], [0.9, 0.4, 0.5, 0.1]) == 0.75


@pytest.mark.parametrize('tc', [False, True])
def test_full_rank_eval_counts_and_metrics(tc):
    from mpqe_b200 import _lib
    if tc and not _lib.load().mpqe_b200_has_tcgen05():
        pytest.skip('library built without tcgen05 kernels')
    kg, cfg, params, model, qsets = setup(per_formula=33)
    rels, _, node_maps = kg.raw()
    mode_ids, rel_ids = O.schema_ids(rels)
    id2row = O.id_to_row(node_maps)
    frm_rels, raw = qsets['2-inter'][0]
    queries = [Query.deserialize(r) for r in raw]
    formula = queries[0].formula
    left, right, pos, n = mp_eval.full_rank_counts(model, formula, queries, [q.target_node for q in queries],
                                                   use_tensor_cores=tc)
    spec = O.formula_spec('2-inter', frm_rels)
    a_ids, var_ids, ei, et, b = O.query_graph(spec, [q.anchor_nodes for q in queries], rel_ids, mode_ids)
    with torch.no_grad():
        q = O.encode_queries(params, cfg, spec, a_ids, var_ids, ei, et, b, id2row).double()
    table = params['enc.feat-%s.weight' % spec['target_mode']][:-1].double()
    s = (q / q.norm(dim=1, keepdim=True)) @ (table / table.norm(dim=1, keepdim=True)).t()
    p = pos.cpu().double().unsqueeze(1)
    tol = 6e-6 if tc else 3e-6
    l, r = left.cpu(), right.cpu()
    assert n == table.shape[0]
    assert bool(((l >= (s < p - tol).sum(1)) & (l <= (s < p + tol).sum(1))).all())
    assert bool(((r >= (s <= p - tol).sum(1)) & (r <= (s <= p + tol).sum(1))).all())
    m = mp_eval.ranking_metrics(left, right, n)
    assert 0 < m['MRR'] <= 1 and 0 <= m['APR'] <= 100 and m['queries'] == 33
    # two "shards" chained on one device give the same integers as one
    from mpqe_b200 import ops
    l2 = torch.zeros_like(left)
    r2 = torch.zeros_like(right)
    job = model.make_job(formula, queries)
    from mpqe_b200.model import Weights
    with torch.cuda.device(0):
        model._engine.encode([job], Weights(model, False))
        tab = model.enc.table(formula.target_mode)
        for rank in range(3):
            b0, b1 = mp_eval.shard_rows(n, rank, 3)
            ops.rank_counts_table(job.q, pos, tab, b0, b1, l2, r2, use_tensor_cores=tc)
    assert torch.equal(l2, left) and torch.equal(r2, right)


def test_graphed_rank_counts_equal_eager_counts():
    """A ranking batch replayed as one CUDA graph gives the integers of the launch-by-launch path, also after new ids
    were copied into its buffers."""
    kg, cfg, params, model, qsets = setup(per_formula=48)
    frm_rels, raw = qsets['3-inter'][0]
    queries = [Query.deserialize(r) for r in raw]
    formula = queries[0].formula
    a_ids, var_ids, q_graphs = data_utils.RGCNQueryDataset.get_query_graph(formula, queries, model.rel_ids, model.mode_ids)
    targets = torch.tensor([q.target_node for q in queries])
    index = mp_eval.RankIndex(model, distributed=False)
    first = slice(0, 24)
    graphed = mp_eval.GraphedCounts(index, formula, a_ids[first], targets[first], var_ids,
                                    data_utils.QueryGraphBatch(q_graphs.template, q_graphs.edge_rel_ids, 24))
    for sl in (first, slice(24, 48), first):
        l, r, pos, n = graphed(a_ids[sl], targets[sl])
        want_l, want_r, want_pos, want_n = index.counts(formula, queries[sl], targets[sl])
        torch.cuda.synchronize()
        assert n == want_n and torch.equal(l, want_l) and torch.equal(r, want_r) and torch.equal(pos, want_pos)


def test_training_loop_runs_and_loss_decreases():
    kg, cfg, params, model, qsets = setup(per_formula=64)
    train = {qt: {Query.deserialize(raw[0]).formula: [Query.deserialize(r) for r in raw]}
             for qt, groups in qsets.items() for (_, raw) in groups}
    evalq = {'one_neg': {qt: {f: qs[:8] for f, qs in d.items()} for qt, d in train.items()},
             'full_neg': {qt: {f: qs[:8] for f, qs in d.items()} for qt, d in train.items()}}
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    log = logging.getLogger('mpqe_test')
    frm = next(iter(train['1-chain']))
    qs = train['1-chain'][frm]
    with torch.no_grad():
        before = float(model.margin_loss_ids(frm, qs, [q.target_node for q in qs], [q.neg_samples[0] for q in qs]))
    train_helpers.run_train(model, opt, train, evalq, evalq, log, max_burn_in=3, batch_size=32, log_every=1000,
                            val_every=1000, max_iter=12)
    with torch.no_grad():
        after = float(model.margin_loss_ids(frm, qs, [q.target_node for q in qs], [q.neg_samples[0] for q in qs]))
    assert np.isfinite(after) and after < before


def test_fused_training_loop_runs_and_loss_decreases():
    """`run_train_fused`: the reference's loop structure on the fused step -- pre-tensorised query sets on the device,
    negatives drawn on the device, fused Adam, no per-step loss read-back."""
    from mpqe_b200.tensor_queries import TensorQuerySet
    kg, cfg, params, model, qsets = setup(per_formula=64)
    train = {qt: {Query.deserialize(raw[0]).formula: [Query.deserialize(r) for r in raw]}
             for qt, groups in qsets.items() for (_, raw) in groups}
    evalq = {'one_neg': {qt: {f: qs[:8] for f, qs in d.items()} for qt, d in train.items()},
             'full_neg': {qt: {f: qs[:8] for f, qs in d.items()} for qt, d in train.items()}}
    sets = {qt: TensorQuerySet.from_queries(d) for qt, d in train.items()}
    log = logging.getLogger('mpqe_test')
    frm = next(iter(train['1-chain']))
    qs = train['1-chain'][frm]
    with torch.no_grad():
        before = float(model.margin_loss_ids(frm, qs, [q.target_node for q in qs], [q.neg_samples[0] for q in qs]))
    train_helpers.run_train_fused(model, sets, evalq, evalq, log, full_lists=model.graph.full_lists, max_burn_in=3,
                                  batch_size=32, log_every=5, val_every=1000, max_iter=14)
    with torch.no_grad():
        after = float(model.margin_loss_ids(frm, qs, [q.target_node for q in qs], [q.neg_samples[0] for q in qs]))
    assert np.isfinite(after) and after < before


def test_train_step_graph_replay_equals_eager():
    """The CUDA-graph step (row (f)1) gives the same bits as the eager step, and follows new ids copied into the
    static buffers."""
    from mpqe_b200.graph import Formula
    from mpqe_b200.train_step import HostBatch, TrainStep
    kg = synthetic.make_kg('aifb', seed=3)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
    frng = np.random.RandomState(0)
    formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in synthetic.QUERY_TYPES]

    def host(seed):
        rng = np.random.RandomState(seed)
        return [HostBatch(f, *[torch.from_numpy(x) for x in synthetic.sample_id_batch(kg, f, 500, rng)]) for f in formulas]

    ts = TrainStep(model)
    h1, h2 = host(1), host(2)
    eager1 = ts.forward_backward([ts.to_device(hb) for hb in h1])
    e1 = (eager1.losses.clone(), eager1.dense.flat.clone(), [x.clone() for x in eager1.sparse])
    eager2 = ts.forward_backward([ts.to_device(hb) for hb in h2])
    e2 = (eager2.losses.clone(), eager2.dense.flat.clone(), [x.clone() for x in eager2.sparse])
    ts.capture(h1)
    for hb, want in ((h1, e1), (h2, e2), (h1, e1)):
        res, losses_host = ts.run_host(hb)
        torch.cuda.synchronize()
        assert torch.equal(res.losses, want[0]) and torch.equal(losses_host, want[0].cpu())
        assert torch.equal(res.dense.flat, want[1]), 'dense gradients differ between graph replay and eager step'
        k = int(want[2][2])
        assert int(res.sparse[2]) == k and torch.equal(res.sparse[0][:k], want[2][0][:k])
        assert torch.equal(res.sparse[1][:k], want[2][1][:k])
    # a loader writing the next batch in place into the step's single pinned buffer: one H2D copy, same result
    staging = ts.staging()
    for st, hb in zip(staging, h2):
        st.anchor_ids.copy_(hb.anchor_ids)
        st.targets.copy_(hb.targets)
        st.negatives.copy_(hb.negatives)
    res, losses_host = ts.run_host(staging)
    torch.cuda.synchronize()
    assert torch.equal(losses_host, e2[0].cpu()) and torch.equal(res.dense.flat, e2[1])


@pytest.mark.parametrize('readout,num_layers,adaptive,shared,scatter_op',
                         [('sum', 2, False, False, 'add'), ('sum', 3, False, True, 'add'), ('mp', 3, True, False, 'add'),
                          ('max', 2, False, False, 'add'), ('concat', 2, False, False, 'add'),
                          ('targetmlp', 2, False, False, 'max'), ('mlp', 2, False, False, 'mean')])
def test_train_step_gradients_match_oracle(readout, num_layers, adaptive, shared, scatter_op):
    """The fused multi-batch step (planned row slots, margin backward inside the forward, collapsed last pass,
    batch-constant rows, second stream) against the oracle: losses per batch, every dense gradient summed over the
    seven formula batches, and the combined row-sparse entity gradients scattered back into dense tables."""
    from mpqe_b200.graph import Formula
    from mpqe_b200.train_step import HostBatch, TrainStep
    from tests.model_utils import oracle_loss_and_grads
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout=readout, num_layers=num_layers, adaptive=adaptive, shared_layers=shared,
                   scatter_op=scatter_op, weight_decay=1e-3)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    mode_ids, rel_ids = O.schema_ids(rels)
    id2row = O.id_to_row(node_maps)
    model = build_model(kg.raw(), cfg, params, DEV, sparse_grad=True)
    frng, rng = np.random.RandomState(0), np.random.RandomState(1)
    want_losses, want = [], {}
    host = []
    for qt in synthetic.QUERY_TYPES:
        frm_rels = kg.sample_formula(qt, frng)
        formula = Formula(qt, frm_rels)
        a, t, n = synthetic.sample_id_batch(kg, formula, 70, rng)
        host.append(HostBatch(formula, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n)))
        spec = O.formula_spec(qt, frm_rels)
        loss, grads = oracle_loss_and_grads(params, cfg, spec, torch.from_numpy(a), rel_ids, mode_ids, id2row,
                                            torch.from_numpy(t), torch.from_numpy(n))
        want_losses.append(loss)
        for k, g in grads.items():
            want[k] = want.get(k, 0) + g
    ts = TrainStep(model)
    res = ts.forward_backward([ts.to_device(hb) for hb in host])
    torch.cuda.synchronize()
    assert_close(res.losses.cpu().numpy(), np.array(want_losses, dtype=np.float32), 1e-5, 1e-5, 'losses')
    assert_close(float(res.total), float(np.sum(want_losses)), 1e-5, 1e-5, 'total')
    from tests.model_utils import train_step_grads
    got, _ = train_step_grads(ts, model, res, DEV)
    for name, g in want.items():
        if shared and name.startswith('layers.') and not name.startswith('layers.0.'):
            continue      # shared layers: the oracle reports the one parameter set under layers.0
        assert got.get(name) is not None, name
        g = np.asarray(g)
        from tests.helpers import assert_grad_close
        from mpqe_b200 import ops
        assert_grad_close(got[name].detach().cpu().numpy(), g, 'tcgen05' if ops.tensor_cores_default() else 'ffma',
                          'tiny-step:%s grad %s' % (readout, name))

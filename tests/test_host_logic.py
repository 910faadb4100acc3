"""Host-side orchestration (term lists, slot maps, gradient routing, autograd wiring) checked on CPU: the CUDA
entry points are replaced by `tests/emulator.py`, the result is compared with the golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from mpqe_b200 import data_utils, ops
from oracle import mpqe_oracle as O
from tests import emulator
from tests.helpers import GoldenCase, assert_close, golden_names
from tests.model_utils import build_model, model_grads, oracle_loss_and_grads, queries_from_ids


@pytest.fixture
def emu(monkeypatch):
    emulator.install(monkeypatch)


@pytest.mark.parametrize('name', golden_names())
def test_golden_through_host_logic(emu, name):
    c = GoldenCase(name)
    model = build_model(c.kg.raw(), c.cfg, c.params, 'cpu')
    assert model.mode_ids == dict(c.mode_ids) and model.rel_ids == dict(c.rel_ids)
    queries = queries_from_ids(c.query_type, c.rels, c.z['anchor_ids'], c.z['targets'])
    formula = queries[0].formula
    # integer layout
    a_ids, var_ids, qg = data_utils.RGCNQueryDataset.get_query_graph(formula, queries, model.rel_ids, model.mode_ids)
    assert np.array_equal(a_ids.numpy(), c.z['anchor_ids']) and np.array_equal(var_ids.numpy(), c.z['var_ids'])
    assert np.array_equal(qg.edge_index.numpy(), c.z['edge_index'])
    assert np.array_equal(qg.edge_type.numpy(), c.z['edge_type'])
    assert np.array_equal(qg.batch.numpy(), c.z['batch'])
    # eval-style scores
    with torch.no_grad():
        s = model.forward(formula, queries, c.z['targets'].tolist(), neg_nodes=c.z['eval_neg_nodes'].tolist(),
                          neg_lengths=c.z['eval_neg_lengths'].tolist())
    assert_close(s.numpy(), c.z['eval_scores'], 1e-4, 2e-6, 'scores')
    # loss + gradients
    model.zero_grad()
    loss = model.margin_loss_ids(formula, queries, c.z['targets'].tolist(), c.z['train_neg_nodes'].tolist())
    assert_close(loss.item(), c.z['loss'], 1e-5, 1e-6, 'loss')
    loss.backward()
    want = c.grads()
    got = model_grads(model)
    for k, g in want.items():
        assert_close(got[k], g, 1e-3, 2e-5 * max(np.abs(g).max(), 1e-12), 'grad ' + k)


@pytest.mark.parametrize('readout,num_layers,adaptive,shared', [
    ('sum', 2, False, False), ('sum', 3, False, True), ('max', 2, False, False), ('mp', 3, True, False),
    ('mp', 3, True, True), ('mp', 2, False, False), ('concat', 2, False, False), ('concat', 3, False, False),
    ('mlp', 2, False, False), ('targetmlp', 2, False, False), ('sum', 1, False, False)])
def test_all_types_vs_oracle(emu, readout, num_layers, adaptive, shared):
    from mpqe_b200 import synthetic
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout=readout, num_layers=num_layers, adaptive=adaptive, shared_layers=shared, weight_decay=1e-3)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    mode_ids, rel_ids = O.schema_ids(rels)
    id2row = O.id_to_row(node_maps)
    qsets = synthetic.make_query_sets(kg, queries_per_formula=7, formulas_per_type=1, seed=2)
    rng = np.random.RandomState(0)
    for sparse in (False, True):
        model = build_model(kg.raw(), cfg, params, 'cpu', sparse_grad=sparse)
        for qt in synthetic.QUERY_TYPES:
            frm_rels, raw = qsets[qt][0]
            spec = O.formula_spec(qt, frm_rels)
            parsed = [O.query_anchors_target(r[0]) for r in raw]
            anchors = torch.tensor([p[2] for p in parsed])
            targets = torch.tensor([p[3] for p in parsed])
            negs = torch.tensor(node_maps[spec['target_mode']])[rng.randint(len(node_maps[spec['target_mode']]),
                                                                            size=len(raw))]
            want_loss, want = oracle_loss_and_grads(params, cfg, spec, anchors, rel_ids, mode_ids, id2row, targets,
                                                    negs)
            queries = queries_from_ids(qt, frm_rels, anchors, targets)
            model.zero_grad()
            loss = model.margin_loss_ids(queries[0].formula, queries, targets, negs)
            assert_close(loss.item(), want_loss, 1e-5, 1e-6, qt + ' loss')
            loss.backward()
            got = model_grads(model)
            for k, g in want.items():
                assert_close(got[k], g, 1e-3, 2e-5 * max(np.abs(g).max(), 1e-12), '%s grad %s' % (qt, k))


def test_adaptive_needs_enough_layers(emu):
    from mpqe_b200 import synthetic
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2, adaptive=True)
    model = build_model(kg.raw(), cfg, O.init_params(rels, node_maps, cfg, seed=1), 'cpu')
    frm_rels, raw = synthetic.make_query_sets(kg, 3, 1, seed=2, query_types=('3-chain',))['3-chain'][0]
    parsed = [O.query_anchors_target(r[0]) for r in raw]
    queries = queries_from_ids('3-chain', frm_rels, [p[2] for p in parsed], [p[3] for p in parsed])
    with pytest.raises(ValueError, match='adaptive with 2 layers'):
        model.margin_loss_ids(queries[0].formula, queries, [p[3] for p in parsed], [p[3] for p in parsed])
    with pytest.raises(Exception, match='Hard negative'):
        model.margin_loss(queries[0].formula, queries, hard_negatives=True)


def test_export_embeddings_matches_reference_loop(emu, tmp_path):
    """`embeddings.npy` export (reference train.py:147-163): [id, unit-norm embedding] per entity, zero row for ids in
    no mode; checked against the reference's one-id-at-a-time loop restated on the oracle's tables."""
    from mpqe_b200 import synthetic
    from mpqe_b200.utils import export_embeddings
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2)
    params = O.init_params(rels, node_maps, cfg, seed=1)
    model = build_model(kg.raw(), cfg, params, 'cpu')
    ids = sorted({int(i) for m in node_maps for i in node_maps[m]})
    entity_ids = {('e%d' % k): i for k, i in enumerate(ids[::3] + [10 ** 6])}      # one id that no mode contains
    path = str(tmp_path / 'embeddings.npy')
    got = export_embeddings(model, entity_ids, path, batch_size=7)
    assert got.shape == (len(entity_ids), 129) and np.array_equal(np.load(path), got)
    id2row = O.id_to_row(node_maps)
    for k, ent in enumerate(entity_ids.values()):
        modes = [m for m in model.graph.full_sets if ent in model.graph.full_sets[m]]
        if not modes:
            assert not got[k].any()
            continue
        row = params['enc.feat-%s.weight' % modes[-1]][id2row[ent]]
        want = (row / row.norm()).numpy()
        assert got[k, 0] == ent
        np.testing.assert_allclose(got[k, 1:], want, rtol=1e-6, atol=1e-7)


def test_rgcn_conv_reference_signature_infers_the_template(emu):
    """`RGCNConv.forward(x, edge_index, edge_type)` exactly as the reference calls it (model.py:269): the template of
    the batched query graphs is recovered from the edge list; anything that is not such a batch is refused."""
    from mpqe_b200.model import RGCNConv, infer_template_batch
    from mpqe_b200.data_utils import QueryGraphBatch, template_of
    torch.manual_seed(0)
    conv = RGCNConv(128, 128, 6, 0)
    for qt, rels, B in (('3-chain_inter', [4, 1, 2], 13), ('1-chain', [3], 5), ('3-inter', [0, 0, 5], 1)):
        t = template_of(qt)
        g = QueryGraphBatch(t, rels, B)
        inferred = infer_template_batch(B * t.num_nodes, g.edge_index, g.edge_type)
        assert inferred.num_graphs == B and inferred.template.num_nodes == t.num_nodes
        assert inferred.template.src == list(t.src) and inferred.template.dst == list(t.dst)
        assert inferred.edge_rel_ids == rels
        x = torch.randn(B * t.num_nodes, 128, requires_grad=True)
        out = conv(x, g.edge_index, g.edge_type)
        pc = {k: v.detach().clone().requires_grad_(True) for k, v in conv.named_parameters()}
        xc = x.detach().clone().requires_grad_(True)
        want = O.rgcn_conv(xc, g.edge_index, g.edge_type, pc['basis'], pc['root'], pc['bias'])
        assert_close(out.detach().numpy(), want.detach().numpy(), 1e-5, 1e-5, qt + ' conv out')
        w = torch.randn_like(want)
        conv.zero_grad()
        (out * w).sum().backward()
        (want * w).sum().backward()
        assert_close(x.grad.numpy(), xc.grad.numpy(), 1e-4, 1e-5, qt + ' dx')
        for k, prm in conv.named_parameters():
            assert_close(prm.grad.numpy(), pc[k].grad.numpy(), 1e-3, 1e-4, qt + ' d' + k)
    with pytest.raises(NotImplementedError, match='not a batch of identical query templates'):
        infer_template_batch(40, torch.randint(0, 40, (2, 30)), torch.randint(0, 6, (30,)))


def test_rgcn_conv_basis_decomposition(emu):
    """`num_bases > 0` (reference model.py:243-248, 281-284): W_r = sum_b att[r, b] basis[b]; output and the gradients
    of `att` and `basis` against the oracle."""
    from mpqe_b200.model import RGCNConv
    from mpqe_b200.data_utils import QueryGraphBatch, template_of
    torch.manual_seed(1)
    conv = RGCNConv(128, 128, 7, 3)
    assert conv.att.shape == (7, 3) and conv.basis.shape == (3, 128, 128)
    t = template_of('3-inter_chain')
    g = QueryGraphBatch(t, [6, 2, 0], 11)
    x = torch.randn(11 * t.num_nodes, 128, requires_grad=True)
    out = conv(x, g.edge_index, g.edge_type, graph=g)
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in conv.named_parameters()}
    xc = x.detach().clone().requires_grad_(True)
    want = O.rgcn_conv(xc, g.edge_index, g.edge_type, pc['basis'], pc['root'], pc['bias'], att=pc['att'])
    assert_close(out.detach().numpy(), want.detach().numpy(), 1e-5, 1e-5, 'conv out (bases)')
    w = torch.randn_like(want)
    (out * w).sum().backward()
    (want * w).sum().backward()
    assert_close(x.grad.numpy(), xc.grad.numpy(), 1e-4, 1e-5, 'dx')
    for k, prm in conv.named_parameters():
        assert_close(prm.grad.numpy(), pc[k].grad.numpy(), 1e-3, 1e-4, 'd' + k)


@pytest.mark.parametrize('readout,num_layers,adaptive,shared,scatter_op',
                         [('sum', 2, False, False, 'add'), ('sum', 3, False, True, 'add'), ('mp', 3, True, False, 'add'),
                          ('max', 2, False, False, 'add'), ('concat', 2, False, False, 'mean'),
                          ('targetmlp', 2, False, False, 'max')])
def test_fused_train_step_host_logic_vs_oracle(emu, readout, num_layers, adaptive, shared, scatter_op):
    """The host side of the fused multi-batch step (planned row slots, margin backward inside the forward, collapsed
    last pass, batch-constant rows) with emulated kernels: losses, dense gradients summed over the seven formula
    batches and the combined row-sparse entity gradients against the oracle."""
    from mpqe_b200 import synthetic
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
    model = build_model(kg.raw(), cfg, params, 'cpu', sparse_grad=True)
    frng, rng = np.random.RandomState(0), np.random.RandomState(1)
    want_losses, want, host = [], {}, []
    for qt in synthetic.QUERY_TYPES:
        frm_rels = kg.sample_formula(qt, frng)
        formula = Formula(qt, frm_rels)
        a, t, n = synthetic.sample_id_batch(kg, formula, 9, rng)
        host.append(HostBatch(formula, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n)))
        loss, grads = oracle_loss_and_grads(params, cfg, O.formula_spec(qt, frm_rels), torch.from_numpy(a), rel_ids,
                                            mode_ids, id2row, torch.from_numpy(t), torch.from_numpy(n))
        want_losses.append(loss)
        for k, g in grads.items():
            want[k] = want.get(k, 0) + g
    ts = TrainStep(model)
    res = ts.forward_backward([ts.to_device(hb) for hb in host])
    assert_close(res.losses.numpy(), np.array(want_losses, dtype=np.float32), 1e-5, 1e-5, 'losses')
    G = res.dense
    got = {}
    for name, prm in model.named_parameters():
        for li, layer in enumerate(model.distinct_layers()):
            if prm is layer.basis:
                got[name] = G.dw[li]
            elif prm is layer.root:
                got[name] = G.droot[li]
            elif prm is layer.bias:
                got[name] = G.dbias[li]
        if prm is model.mode_embeddings.weight:
            got[name] = G.dmode
    if isinstance(model.readout, torch.nn.Module):
        lin1, lin2 = model.readout.layers[0], model.readout.layers[2]
        for name, prm in model.named_parameters():
            if prm is lin1.weight:
                got[name] = G.dw1t.t()
            elif prm is lin2.weight:
                got[name] = G.dw2t.t()
            elif prm is lin1.bias:
                got[name] = G.db1
            elif prm is lin2.bias:
                got[name] = G.db2
    uid, urows, num = res.sparse
    k = int(num)
    tables = torch.zeros(ts.total_rows, 128)
    tables[uid[:k]] = urows[:k]
    for mode, off in ts.table_offsets.items():
        got['enc.feat-%s.weight' % mode] = tables[off:off + model.enc.table(mode).shape[0]]
    for name, g in want.items():
        assert got.get(name) is not None, name
        g = np.asarray(g)
        assert_close(got[name].detach().numpy(), g, 1e-3, 3e-5 * max(np.abs(g).max(), 1e-12), 'grad ' + name)


def test_unknown_readout_and_scatter():
    from mpqe_b200 import synthetic
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    with pytest.raises(ValueError, match='Unknown readout'):
        build_model(kg.raw(), O.Config(readout='nope'), {}, 'cpu')
    with pytest.raises(ValueError, match='Unknown scatter op'):
        build_model(kg.raw(), O.Config(readout='sum', scatter_op='prod'), {}, 'cpu')


def test_no_cpu_fallback():
    """Without the emulator the product refuses to run on CPU tensors."""
    from mpqe_b200 import synthetic, _lib
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2)
    model = build_model(kg.raw(), cfg, O.init_params(rels, node_maps, cfg, seed=1), 'cpu')
    frm_rels, raw = synthetic.make_query_sets(kg, 3, 1, seed=2, query_types=('2-inter',))['2-inter'][0]
    parsed = [O.query_anchors_target(r[0]) for r in raw]
    queries = queries_from_ids('2-inter', frm_rels, [p[2] for p in parsed], [p[3] for p in parsed])
    with pytest.raises(_lib.MpqeError, match='no CPU fallback'):
        model.margin_loss_ids(queries[0].formula, queries, [p[3] for p in parsed], [p[3] for p in parsed])


@pytest.mark.parametrize('kind,op', [('mlp', 'add'), ('concat', 'add'), ('concat', 'max'), ('targetmlp', 'mean'),
                                     ('targetmlp', 'add')])
def test_readout_modules_forward_with_the_reference_signature(kind, op, monkeypatch):
    """`MLPReadout` / `TargetMLPReadout` called on their own, `readout(embs=, batch_idx=, batch_size=, num_nodes=,
    num_anchors=)` (reference model.py:447-449, 506-515, 531-553), against the oracle's restatement: value and every
    gradient."""
    import torch
    from mpqe_b200 import model as M
    emulator.install(monkeypatch)
    torch.manual_seed(0)
    B, n, a, d = 7, 4, 2, 128
    blocks = {'mlp': 1, 'concat': 2, 'targetmlp': 1}[kind]
    mod = M.TargetMLPReadout(d, op) if kind == 'targetmlp' else M.MLPReadout(d * blocks, d, op)
    embs = torch.randn(B * n, d * blocks, requires_grad=True)
    batch_idx = torch.arange(B).repeat_interleave(n)
    out = mod(embs=embs, batch_idx=batch_idx, batch_size=B, num_nodes=n, num_anchors=a)
    p = {'readout.layers.0.weight': mod.layers[0].weight.detach().clone().requires_grad_(True),
         'readout.layers.0.bias': mod.layers[0].bias.detach().clone().requires_grad_(True),
         'readout.layers.2.weight': mod.layers[2].weight.detach().clone().requires_grad_(True),
         'readout.layers.2.bias': mod.layers[2].bias.detach().clone().requires_grad_(True)}
    e2 = embs.detach().clone().requires_grad_(True)
    want, _ = O.readout(kind, e2, batch_idx, B, n, a, p, op)
    assert out.shape == (B, d)
    np.testing.assert_allclose(out.detach().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-5)
    w = torch.randn(B, d)
    (out * w).sum().backward()
    (want * w).sum().backward()
    np.testing.assert_allclose(embs.grad.numpy(), e2.grad.numpy(), rtol=1e-4, atol=1e-5)
    for name, prm in (('readout.layers.0.weight', mod.layers[0].weight), ('readout.layers.0.bias', mod.layers[0].bias),
                      ('readout.layers.2.weight', mod.layers[2].weight), ('readout.layers.2.bias', mod.layers[2].bias)):
        np.testing.assert_allclose(prm.grad.numpy(), p[name].grad.numpy(), rtol=1e-4, atol=1e-5)


def test_train_step_check_ids_rejects_unknown_entities():
    """The kernels index the tables without validating ids (the reference's nn.Embedding raises IndexError); the step's
    host-side check does, and `capture` runs it on its batches."""
    from mpqe_b200 import synthetic
    from mpqe_b200.graph import Formula
    from mpqe_b200.train_step import HostBatch, TrainStep
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    model = build_model(kg.raw(), cfg, params, 'cpu', sparse_grad=True)
    ts = TrainStep(model)
    f = Formula('2-inter', kg.sample_formula('2-inter', np.random.RandomState(0)))
    a, t, n = synthetic.sample_id_batch(kg, f, 8, np.random.RandomState(1))
    ts.check_ids(HostBatch(f, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n)))
    bad = a.copy()
    bad[3, 1] = 10 ** 9
    with pytest.raises(IndexError):
        ts.check_ids(HostBatch(f, torch.from_numpy(bad), torch.from_numpy(t), torch.from_numpy(n)))
    bad_t = t.copy()
    bad_t[0] = -1
    with pytest.raises(IndexError):
        ts.check_ids(HostBatch(f, torch.from_numpy(a), torch.from_numpy(bad_t), torch.from_numpy(n)))

"""Live pin of the oracle against the unmodified reference (only where /root/reference exists: the build container)."""
import random

import numpy as np
import pytest
import torch

from mpqe_b200 import synthetic
from oracle import mpqe_oracle as O
from oracle import ref_loader
from oracle.make_golden import load_params_into
from tests.helpers import assert_close

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')

CONFIGS = [('sum', 2, False, False), ('max', 2, False, False), ('concat', 2, False, False),
           ('mp', 3, True, False), ('mp', 3, True, True), ('mlp', 2, False, False), ('targetmlp', 2, False, False)]


@pytest.mark.parametrize('readout,num_layers,adaptive,shared', CONFIGS)
def test_all_query_types_match_reference(readout, num_layers, adaptive, shared):
    ref = ref_loader.load()
    kg = synthetic.make_kg('tiny', seed=21)
    rels, _, node_maps = kg.raw()
    qsets = synthetic.make_query_sets(kg, queries_per_formula=9, formulas_per_type=1, seed=4, num_neg=4)
    cfg = O.Config(readout=readout, num_layers=num_layers, adaptive=adaptive, shared_layers=shared,
                   weight_decay=1e-3)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=2)
    model, graph, id2row = ref_loader.build_reference_model(kg.raw(), 128, readout, num_layers, adaptive,
                                                            shared_layers=shared, weight_decay=1e-3)
    load_params_into(model, params, cfg)
    mode_ids, rel_ids = O.schema_ids(rels)
    assert dict(mode_ids) == model.mode_ids and dict(rel_ids) == model.rel_ids
    assert torch.equal(id2row, O.id_to_row(node_maps))
    for qt in synthetic.QUERY_TYPES:
        frm_rels, raw = qsets[qt][0]
        queries = ref_loader.deserialize_queries(raw)
        formula = queries[0].formula
        spec = O.formula_spec(qt, frm_rels)
        assert spec['target_mode'] == formula.target_mode and spec['anchor_modes'] == formula.anchor_modes
        a_ref, v_ref, qg = ref['data_utils'].RGCNQueryDataset.get_query_graph(formula, queries, model.rel_ids,
                                                                             model.mode_ids)
        anchors = [O.query_anchors_target(r[0])[2] for r in raw]
        targets = [O.query_anchors_target(r[0])[3] for r in raw]
        assert anchors == [q.anchor_nodes for q in queries] and targets == [q.target_node for q in queries]
        a_ids, var_ids, ei, et, batch = O.query_graph(spec, anchors, rel_ids, mode_ids)
        assert torch.equal(a_ids, a_ref) and torch.equal(var_ids, v_ref)
        assert torch.equal(ei, qg.edge_index) and torch.equal(et, qg.edge_type) and torch.equal(batch, qg.batch)

        random.seed(9)
        negs = [random.choice(graph.full_lists[formula.target_mode]) if qt == '1-chain'
                else random.choice(q.neg_samples) for q in queries]
        random.seed(9)
        model.zero_grad()
        loss_ref = model.margin_loss(formula, queries, a_ref, v_ref, qg)
        loss_ref.backward()
        p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        loss = O.margin_loss(p, cfg, spec, a_ids, var_ids, ei, et, batch, id2row, torch.tensor(targets),
                             torch.tensor(negs))
        loss.backward()
        assert_close(loss.item(), loss_ref.item(), 1e-6, 1e-7, qt + ' loss')
        for k, prm in model.named_parameters():
            g_ref = np.zeros(tuple(prm.shape), np.float32) if prm.grad is None else prm.grad.numpy()
            g = p[k].grad
            g = np.zeros_like(g_ref) if g is None else g.numpy()
            assert_close(g, g_ref, 1e-5, 1e-7 * max(np.abs(g_ref).max(), 1e-30) + 1e-12, qt + ' grad ' + k)

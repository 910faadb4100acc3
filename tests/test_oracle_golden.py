"""The CPU oracle against the committed outputs of the reference itself (tests/golden, made by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import mpqe_oracle as O
from tests.helpers import GoldenCase, assert_close, golden_names

# fp32 tolerance between two CPU fp32 evaluations of the same expression tree
RTOL, ATOL = 1e-5, 1e-6


@pytest.mark.parametrize('name', golden_names())
def test_layout_bit_exact(name):
    c = GoldenCase(name)
    anchors = c.anchor_ids.tolist()
    a_ids, var_ids, ei, et, batch = O.query_graph(c.spec, anchors, c.rel_ids, c.mode_ids)
    for got, key in ((a_ids, 'anchor_ids'), (var_ids, 'var_ids'), (ei, 'edge_index'), (et, 'edge_type'),
                     (batch, 'batch')):
        assert got.dtype == torch.int64
        assert np.array_equal(got.numpy(), c.z[key]), key


@pytest.mark.parametrize('name', golden_names())
def test_eval_scores_and_percentiles(name):
    c = GoldenCase(name)
    a_ids, var_ids, ei, et, batch = O.query_graph(c.spec, c.anchor_ids.tolist(), c.rel_ids, c.mode_ids)
    lengths = c.z['eval_neg_lengths'].tolist()
    with torch.no_grad():
        s = O.forward_scores(c.params, c.cfg, c.spec, a_ids, var_ids, ei, et, batch, c.id2row, c.targets,
                             c.z['eval_neg_nodes'], lengths)
    assert_close(s.numpy(), c.z['eval_scores'], RTOL, ATOL, 'scores')
    B = len(lengths)
    left, right = O.rank_counts(c.z['eval_scores'][:B], c.z['eval_scores'][B:], lengths)
    assert np.array_equal(O.percentile_from_counts(left, right, lengths), c.z['eval_perc'])


@pytest.mark.parametrize('name', golden_names())
def test_margin_loss_and_grads(name):
    c = GoldenCase(name)
    p = {k: v.clone().requires_grad_(True) for k, v in c.params.items()}
    a_ids, var_ids, ei, et, batch = O.query_graph(c.spec, c.anchor_ids.tolist(), c.rel_ids, c.mode_ids)
    loss = O.margin_loss(p, c.cfg, c.spec, a_ids, var_ids, ei, et, batch, c.id2row, c.targets,
                         torch.from_numpy(c.z['train_neg_nodes']))
    assert_close(loss.item(), c.z['loss'], RTOL, ATOL, 'loss')
    want = c.grads()
    if not want:
        return
    loss.backward()
    for k, g in want.items():
        got = p[k].grad
        got = np.zeros_like(g) if got is None else got.numpy()
        scale = max(np.abs(g).max(), 1e-12)
        assert_close(got, g, 1e-4, 1e-5 * scale, 'grad ' + k)


def test_percentile_known_answer():
    # scipy.stats.percentileofscore([1,2,3,4], 3) == 75 ; ties: ([1,2,3,3,4], 3) == 70
    l, r = O.rank_counts([3.0, 3.0], [1, 2, 3, 4, 1, 2, 3, 3, 4], [4, 5])
    assert l.tolist() == [2, 2] and r.tolist() == [3, 4]
    assert O.percentile_from_counts(l, r, [4, 5]).tolist() == [75.0, 70.0]


def test_auc_known_answer():
    assert O.auc([1, 1, 0, 0], [0.9, 0.4, 0.5, 0.1]) == 0.75
    assert O.auc([1, 0], [0.5, 0.5]) == 0.5


def test_relation_sorted_layout_small():
    et = torch.tensor([2, 0, 1, 2, 0, 1, 2, 0, 1])
    perm, off = O.relation_sorted_layout(et, 4)
    assert perm.tolist() == [1, 4, 7, 2, 5, 8, 0, 3, 6]
    assert off.tolist() == [0, 3, 6, 9, 9]


def test_rgcn_layer_hand_computed():
    # 2-node graph, one edge 0->1 of relation 1; W[1] = 2*I, root = I, bias = 1
    d = 4
    x = torch.tensor([[1., 2., 3., 4.], [10., 20., 30., 40.]])
    basis = torch.stack([torch.zeros(d, d), 2 * torch.eye(d)])
    out = O.rgcn_conv(x, torch.tensor([[0], [1]]), torch.tensor([1]), basis, torch.eye(d), torch.ones(d))
    assert out.tolist() == [[2., 3., 4., 5.], [13., 25., 37., 49.]]

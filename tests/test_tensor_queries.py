"""Pre-tensorised query sets: arrays == the Query objects they were made from, storage
round trip, the reference's batch-slicing rule, negatives drawn from the right lists."""
import numpy as np
import pytest
import torch

from mpqe_b200 import data_utils, synthetic
from mpqe_b200.graph import Query
from mpqe_b200.tensor_queries import TensorQuerySet


@pytest.fixture(scope='module')
def query_sets():
    kg = synthetic.make_kg('tiny', seed=7)
    qsets = synthetic.make_query_sets(kg, queries_per_formula=23, formulas_per_type=2, seed=3, num_neg=6, num_hard_neg=3)
    by_type = {}
    for qt, groups in qsets.items():
        by_type[qt] = {}
        for _, raw in groups:
            qs = [Query.deserialize(r) for r in raw]
            by_type[qt][qs[0].formula] = qs
    return kg, by_type


@pytest.mark.parametrize('qt', synthetic.QUERY_TYPES)
def test_arrays_match_query_objects_and_round_trip(query_sets, qt, tmp_path):
    kg, by_type = query_sets
    queries = by_type[qt]
    ts = TensorQuerySet.from_queries(queries)
    assert ts.num_queries == sum(len(v) for v in queries.values())
    path = str(tmp_path / 'q.npz')
    ts.save(path)
    back = TensorQuerySet.load(path)
    assert list(back.by_formula) == list(ts.by_formula)      # formulas (type + nested relation tuples) survive
    for f, qs in queries.items():
        for fq in (ts.by_formula[f], back.by_formula[f]):
            assert len(fq) == len(qs)
            # the anchor ids the reference's collation builds for the same queries (data_utils.py:377-393)
            ref_anchor_ids, _, _ = data_utils.RGCNQueryDataset.get_query_graph(f, qs, *_ids(kg))
            assert torch.equal(torch.from_numpy(fq.anchors), ref_anchor_ids)
            assert fq.targets.tolist() == [q.target_node for q in qs]
            for i, q in enumerate(qs):
                assert fq.negatives_of(i).tolist() == list(q.neg_samples)
                assert fq.negatives_of(i, hard=True).tolist() == list(q.hard_neg_samples or [])
    again = back.to_queries()
    for f, qs in queries.items():
        for q, r in zip(qs, again[f]):
            assert r.formula == f and tuple(r.anchor_nodes) == tuple(q.anchor_nodes) and r.target_node == q.target_node
            assert list(r.neg_samples) == list(q.neg_samples)


def _ids(kg):
    from oracle import mpqe_oracle as O
    mode_ids, rel_ids = O.schema_ids(kg.raw()[0])
    return rel_ids, mode_ids


def test_pick_follows_the_reference_slicing_rule(query_sets):
    kg, by_type = query_sets
    queries = by_type['2-inter']
    ts = TensorQuerySet.from_queries(queries)
    ds = data_utils.QueryDataset(queries)
    for window in ([0, 1, 2, 3, 4], [20, 21, 22], [18, 19, 20, 21, 22], [5]):
        np.random.seed(11)
        f_ref, qs_ref = ds.collate_fn(window)
        rng = np.random.RandomState(11)
        f, start, end = ts.pick(window, rng)
        assert f == f_ref
        assert ts.by_formula[f].targets[start:end].tolist() == [q.target_node for q in qs_ref]


def test_negatives_come_from_each_querys_list(query_sets):
    kg, by_type = query_sets
    rng = np.random.RandomState(0)
    full_lists = {m: sorted(ids) for m, ids in kg.raw()[2].items()}
    for qt in synthetic.QUERY_TYPES:
        ts = TensorQuerySet.from_queries(by_type[qt])
        for f, fq in ts.by_formula.items():
            hb = ts.host_batch(f, 2, 19, rng, full_lists=full_lists)
            assert hb.anchor_ids.shape == (17, len(f.anchor_modes)) and hb.targets.shape == (17,)
            neg = hb.negatives.tolist()
            for k, i in enumerate(range(2, 19)):
                if qt == '1-chain':
                    assert neg[k] in full_lists[f.target_mode]
                else:
                    assert neg[k] in fq.negatives_of(i).tolist()
            if 'inter' in qt:      # hard negatives exist for the intersection types only (reference model.py:466-468)
                hard = fq.sample_negatives(0, len(fq), rng, hard=True)
                for i in range(len(fq)):
                    assert hard[i] in fq.negatives_of(i, hard=True).tolist()
    it = ts.batches(8, rng, full_lists=full_lists)
    sizes = [next(it).targets.numel() for _ in range(6)]
    assert all(1 <= s <= 8 for s in sizes)


def test_template_inference_reproduces_any_batched_edge_list():
    """Property: for a batch of B copies of ANY small template, `infer_template_batch` returns a template (possibly a
    finer one, when the template itself repeats) whose B' copies are exactly the given edge list."""
    from hypothesis import given, settings, strategies as st
    from mpqe_b200.model import infer_template_batch

    @settings(max_examples=120, deadline=None)
    @given(st.integers(1, 6), st.integers(1, 6), st.integers(1, 9), st.randoms(use_true_random=False))
    def check(n, E, B, rnd):
        src = [rnd.randrange(n) for _ in range(E)]
        dst = [rnd.randrange(n) for _ in range(E)]
        rel = [rnd.randrange(5) for _ in range(E)]
        ei = torch.tensor([[s + b * n for b in range(B) for s in src], [d + b * n for b in range(B) for d in dst]])
        et = torch.tensor(rel * B)
        got = infer_template_batch(B * n, ei, et)
        t = got.template
        assert got.num_graphs * t.num_nodes == B * n and got.num_graphs * t.num_edges == B * E
        assert got.num_graphs >= B      # the largest decomposition is preferred
        re_src = [s + b * t.num_nodes for b in range(got.num_graphs) for s in t.src]
        re_dst = [d + b * t.num_nodes for b in range(got.num_graphs) for d in t.dst]
        assert re_src == ei[0].tolist() and re_dst == ei[1].tolist()
        assert got.edge_rel_ids * got.num_graphs == et.tolist()

    check()

"""Pre-tensorised query sets.

The reference keeps a query set as `{formula: [Query]}` -- Python objects unpickled from lists of tuples
(`data_utils.py:155-186`, `graph.py:60-123`) -- and rebuilds id arrays from them for every batch
(`data_utils.py:377-393`: a list comprehension over `q.anchor_nodes`).  Here the same information is converted ONCE
into flat int64 arrays per formula:

    anchors [N, a]     targets [N]     negatives / hard negatives as CSR (offsets [N+1], ids [sum])

so that collating a batch is a slice (the contiguous-slice-with-wrap rule of `QueryDataset.collate_fn`,
`data_utils.py:293-311`), drawing the training negative is one vectorised index computation (`model.py:470-476`
does `random.choice` per query), and the result feeds `TrainStep` as `HostBatch` objects without touching a `Query`.
`save` / `load` store the arrays in one `.npz`; the pickle readers in `data_utils` stay for the published files.
"""
import json
from collections import OrderedDict

import numpy as np
import torch

from .graph import Formula, Query


def _tuplify(x):
    return tuple(_tuplify(v) for v in x) if isinstance(x, (list, tuple)) else x


def _query_from_ids(formula, anchors, target):
    """A `Query` with the given anchors / target; variable nodes, which the R-GCN path never reads, are set to -1."""
    qt, rels = formula.query_type, formula.rels
    if qt.endswith('-chain'):
        nodes = [target] + [-1] * (len(rels) - 1) + [anchors[0]]
        qg = (qt,) + tuple((nodes[i], rels[i], nodes[i + 1]) for i in range(len(rels)))
    elif qt in ('2-inter', '3-inter'):
        qg = (qt,) + tuple((target, r, a) for r, a in zip(rels, anchors))
    elif qt == '3-inter_chain':
        r1, (r2, r3) = rels
        qg = (qt, (target, r1, anchors[0]), ((target, r2, -1), (-1, r3, anchors[1])))
    else:                                   # 3-chain_inter
        r1, (r2, r3) = rels
        qg = (qt, (target, r1, -1), ((-1, r2, anchors[0]), (-1, r3, anchors[1])))
    return Query(qg, None, None)


def _csr(lists):
    offsets = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        offsets[i + 1] = offsets[i] + len(l)
    ids = np.fromiter((v for l in lists for v in l), dtype=np.int64, count=int(offsets[-1]))
    return offsets, ids


class FormulaQueries(object):
    """All queries of one formula as arrays."""

    def __init__(self, formula, anchors, targets, neg_offsets, neg_ids, hard_offsets, hard_ids):
        self.formula = formula
        self.anchors, self.targets = anchors, targets
        self.neg_offsets, self.neg_ids = neg_offsets, neg_ids
        self.hard_offsets, self.hard_ids = hard_offsets, hard_ids

    def __len__(self):
        return int(self.targets.shape[0])

    @classmethod
    def from_queries(cls, formula, queries):
        a = len(formula.anchor_modes)
        anchors = np.asarray([q.anchor_nodes for q in queries], dtype=np.int64).reshape(len(queries), a)
        targets = np.asarray([q.target_node for q in queries], dtype=np.int64)
        no, ni = _csr([list(q.neg_samples) if q.neg_samples is not None else [] for q in queries])
        ho, hi = _csr([list(q.hard_neg_samples) if q.hard_neg_samples is not None else [] for q in queries])
        return cls(formula, anchors, targets, no, ni, ho, hi)

    def negatives_of(self, i, hard=False):
        off, ids = (self.hard_offsets, self.hard_ids) if hard else (self.neg_offsets, self.neg_ids)
        return ids[off[i]:off[i + 1]]

    def sample_negatives(self, start, end, rng, hard=False, full_list=None):
        """One negative per query of [start, end): uniform over the query's (hard) negative list, or -- 1-chain
        queries, as `model.py:473-474` -- over `full_list`, all entities of the target mode.  `rng` is a
        `numpy.random.RandomState`; the draw is vectorised, so it is not the reference's `random.choice` stream."""
        n = end - start
        if self.formula.query_type == '1-chain' and not hard:
            if full_list is None:
                raise ValueError('1-chain negatives are drawn from all entities of the target mode: pass full_list')
            pool = np.asarray(full_list, dtype=np.int64)
            return pool[rng.randint(len(pool), size=n)]
        off, ids = (self.hard_offsets, self.hard_ids) if hard else (self.neg_offsets, self.neg_ids)
        lens = off[start + 1:end + 1] - off[start:end]
        if (lens <= 0).any():
            raise ValueError('a query of %s has no %snegative samples' % (self.formula.query_type, 'hard ' if hard else ''))
        pick = (rng.random_sample(n) * lens).astype(np.int64)
        return ids[off[start:end] + np.minimum(pick, lens - 1)]


class TensorQuerySet(object):
    """`{formula: FormulaQueries}` of one query type (the tensorised `{formula: [Query]}`)."""

    def __init__(self, by_formula):
        self.by_formula = OrderedDict(by_formula)
        self.counts = OrderedDict((f, len(fq)) for f, fq in self.by_formula.items())
        self.num_queries = sum(self.counts.values())
        self.max_num_queries = max(self.counts.values()) if self.counts else 0

    @classmethod
    def from_queries(cls, queries_by_formula):
        return cls((f, FormulaQueries.from_queries(f, qs)) for f, qs in queries_by_formula.items())

    def to_queries(self):
        """Back to `{formula: [Query]}` (anchors / target / negative lists; the query graph is rebuilt from them)."""
        out = OrderedDict()
        for f, fq in self.by_formula.items():
            qs = []
            for i in range(len(fq)):
                q = _query_from_ids(f, tuple(int(v) for v in fq.anchors[i]), int(fq.targets[i]))
                q.neg_samples = [int(v) for v in fq.negatives_of(i)]
                q.hard_neg_samples = [int(v) for v in fq.negatives_of(i, hard=True)]
                qs.append(q)
            out[f] = qs
        return out

    # ---- storage -------------------------------------------------------------------------------------------
    def save(self, path):
        arrays = {'formulas': np.array(json.dumps([[f.query_type, f.rels] for f in self.by_formula]))}
        for k, fq in enumerate(self.by_formula.values()):
            for name in ('anchors', 'targets', 'neg_offsets', 'neg_ids', 'hard_offsets', 'hard_ids'):
                arrays['%d/%s' % (k, name)] = getattr(fq, name)
        np.savez_compressed(path, **arrays)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        formulas = [Formula(qt, _tuplify(rels)) for qt, rels in json.loads(str(z['formulas']))]
        return cls((f, FormulaQueries(f, *[z['%d/%s' % (k, name)] for name in
                                           ('anchors', 'targets', 'neg_offsets', 'neg_ids', 'hard_offsets', 'hard_ids')]))
                   for k, f in enumerate(formulas))

    # ---- batches -------------------------------------------------------------------------------------------
    def pick(self, index_list, rng):
        """(formula, start, end) of the batch the reference's `QueryDataset.collate_fn` would cut for `index_list`:
        formula drawn with probability proportional to its query count, contiguous slice with wrap."""
        counts = np.fromiter(self.counts.values(), dtype=np.float64)
        formula = list(self.counts)[int(np.argmax(rng.multinomial(1, counts / float(self.num_queries))))]
        n = self.counts[formula]
        start = index_list[0] % n
        end = min((index_list[-1] + 1) % n, n)
        if end <= start:
            end = n
        return formula, start, end

    def host_batch(self, formula, start, end, rng, hard=False, full_lists=None, weight=1.0):
        """`train_step.HostBatch` (pinned id tensors) for queries [start, end) of `formula`."""
        from .train_step import HostBatch
        fq = self.by_formula[formula]
        full = full_lists.get(formula.target_mode) if full_lists is not None else None
        neg = fq.sample_negatives(start, end, rng, hard=hard, full_list=full)
        return HostBatch(formula, torch.from_numpy(fq.anchors[start:end]), torch.from_numpy(fq.targets[start:end]),
                         torch.from_numpy(neg), weight)

    def batches(self, batch_size, rng, hard=False, full_lists=None):
        """Endless iterator of HostBatch objects in the reference's order: consecutive index windows of
        `batch_size` over `max_num_queries`, one formula pick per window (`get_queries_iterator`)."""
        while True:
            for first in range(0, self.max_num_queries, batch_size):
                idx = list(range(first, min(first + batch_size, self.max_num_queries)))
                formula, start, end = self.pick(idx, rng)
                yield self.host_batch(formula, start, end, rng, hard=hard, full_lists=full_lists)

"""Seeded synthetic knowledge graphs and query sets with the SHAPE of the reference's datasets.

The reference ships no data generator (datasets are remote tarballs, /root/reference/README.md:28-38), so this
module is new.  It emits exactly the raw, picklable structures the reference's loaders consume, so the same bytes
feed the reference (through `oracle/ref_loader.py`), the oracle and the CUDA path:

* graph:   `(rels, adj_lists, node_maps)` as unpickled by `load_graph` (/root/reference/mpqe/data_utils.py:18-37)
           rels      = {mode: [(to_mode, rel_name), ...]}
           adj_lists = {(mode, rel_name, to_mode): {node: set(neighbours)}}
           node_maps = {mode: [global ids]}      (row i of the mode's table is node_maps[mode][i])
* queries: `(query_graph, neg_samples, hard_neg_samples)` as consumed by `Query.deserialize`
           (/root/reference/mpqe/graph.py:60-90, 120-123); relation triples are oriented target -> anchor.

Entity / mode / relation counts follow SURVEY.md section 8(d): they are shape targets, not dataset claims.
"""
from collections import OrderedDict

import numpy as np

QUERY_TYPES = ('1-chain', '2-chain', '3-chain', '2-inter', '3-inter', '3-inter_chain', '3-chain_inter')

# name -> (entities, modes, typed relations incl. inverses)
SHAPES = {
    'tiny': (90, 3, 8),
    'aifb': (2601, 6, 78),
    'mutag': (22372, 4, 16),
    'am': (372584, 5, 38),
}


class SyntheticKG(object):
    """Typed-relation schema + id assignment.  `raw()` gives the tuple `load_graph` unpickles."""

    def __init__(self, num_entities, num_modes, num_relations, seed=0, adj_nodes_per_rel=64):
        if num_relations % 2:
            raise ValueError('typed relations come in (relation, inverse) pairs; need an even count')
        rng = np.random.RandomState(seed)
        self.modes = ['m%d' % i for i in range(num_modes)]
        # Split the global id space between the modes after a permutation, so that the
        # global-id -> table-row map is not the identity.
        perm = rng.permutation(num_entities)
        cuts = np.linspace(0, num_entities, num_modes + 1).astype(np.int64)
        self.node_maps = OrderedDict(
            (m, perm[cuts[i]:cuts[i + 1]].astype(np.int64)) for i, m in enumerate(self.modes))
        self.num_entities = int(num_entities)

        # Predicates between distinct modes; a ring first (every mode gets an in- and an out-relation,
        # so every query template is type-consistent from every target mode), then random pairs.
        pairs = [(i, (i + 1) % num_modes) for i in range(num_modes)] if num_modes > 1 else []
        while len(pairs) < num_relations // 2:
            i, j = rng.randint(num_modes), rng.randint(num_modes)
            if i != j:
                pairs.append((i, j))
        pairs = pairs[:num_relations // 2]
        self.rels = OrderedDict((m, []) for m in self.modes)
        for p, (i, j) in enumerate(pairs):
            name = 'p%d' % p
            self.rels[self.modes[i]].append((self.modes[j], name))
            self.rels[self.modes[j]].append((self.modes[i], name))
        self.typed_relations = [(m, name, to) for m in self.modes for (to, name) in self.rels[m]]

        # A thin adjacency: enough for the reference `Graph` bookkeeping (edge counts -> rel/mode ordering,
        # full_lists for 1-chain negatives, graph.py:131-170); the hot path never walks it.
        self.adj_lists = {}
        for (m, name, to) in self.typed_relations:
            src = self.node_maps[m]
            dst = self.node_maps[to]
            k = min(adj_nodes_per_rel, len(src))
            chosen = src[rng.choice(len(src), size=k, replace=False)]
            self.adj_lists[(m, name, to)] = {
                int(u): {int(dst[rng.randint(len(dst))])} for u in chosen}

    def raw(self):
        return (OrderedDict((m, list(v)) for m, v in self.rels.items()), self.adj_lists,
                OrderedDict((m, [int(x) for x in ids]) for m, ids in self.node_maps.items()))

    # -- schema walks -------------------------------------------------------------------------
    def out_rels(self, mode):
        return [(mode, name, to) for (to, name) in self.rels[mode]]

    def sample_formula(self, query_type, rng, target_mode=None):
        """Random type-consistent relation structure `rels` for `Formula(query_type, rels)`."""
        t = target_mode if target_mode is not None else self.modes[rng.randint(len(self.modes))]

        def pick(mode):
            outs = self.out_rels(mode)
            return outs[rng.randint(len(outs))]

        if query_type.endswith('-chain'):
            hops, cur, out = int(query_type[0]), t, []
            for _ in range(hops):
                r = pick(cur)
                out.append(r)
                cur = r[2]
            return tuple(out)
        if query_type in ('2-inter', '3-inter'):
            return tuple(pick(t) for _ in range(int(query_type[0])))
        if query_type == '3-inter_chain':
            r1, r2 = pick(t), pick(t)
            return (r1, (r2, pick(r2[2])))
        if query_type == '3-chain_inter':
            r1 = pick(t)
            return (r1, (pick(r1[2]), pick(r1[2])))
        raise ValueError('unknown query type %r' % (query_type,))

    def _ent(self, mode, rng, size=None):
        ids = self.node_maps[mode]
        return ids[rng.randint(len(ids), size=size)]

    def sample_queries(self, query_type, rels, count, rng, num_neg=4, num_hard_neg=2):
        """`count` raw queries of one formula with uniform anchors / targets / negatives of the proper modes."""
        out = []
        t_mode = rels[0][0]
        for _ in range(count):
            t = int(self._ent(t_mode, rng))
            if query_type.endswith('-chain'):
                nodes = [t] + [int(self._ent(r[2], rng)) for r in rels]
                qg = (query_type,) + tuple((nodes[i], rels[i], nodes[i + 1]) for i in range(len(rels)))
            elif query_type in ('2-inter', '3-inter'):
                qg = (query_type,) + tuple((t, r, int(self._ent(r[2], rng))) for r in rels)
            elif query_type == '3-inter_chain':
                r1, (r2, r3) = rels
                v = int(self._ent(r2[2], rng))
                qg = (query_type, (t, r1, int(self._ent(r1[2], rng))),
                      ((t, r2, v), (v, r3, int(self._ent(r3[2], rng)))))
            else:  # 3-chain_inter
                r1, (r2, r3) = rels
                v = int(self._ent(r1[2], rng))
                qg = (query_type, (t, r1, v),
                      ((v, r2, int(self._ent(r2[2], rng))), (v, r3, int(self._ent(r3[2], rng)))))
            neg = [int(x) for x in self._ent(t_mode, rng, size=num_neg)]
            hard = [int(x) for x in self._ent(t_mode, rng, size=num_hard_neg)] \
                if 'inter' in query_type else None
            out.append((qg, neg, hard))
        return out


def make_kg(shape='tiny', seed=0, **kw):
    n, m, r = SHAPES[shape]
    return SyntheticKG(n, m, r, seed=seed, **kw)


def make_query_sets(kg, queries_per_formula=16, formulas_per_type=2, seed=0, num_neg=4, num_hard_neg=2,
                    query_types=QUERY_TYPES):
    """{query_type: [(rels, [raw queries])]} -- deterministic in `seed`."""
    rng = np.random.RandomState(seed + 7919)
    out = OrderedDict()
    for qt in query_types:
        groups, seen = [], set()
        while len(groups) < formulas_per_type:
            rels = kg.sample_formula(qt, rng)
            if rels in seen:
                continue
            seen.add(rels)
            groups.append((rels, kg.sample_queries(qt, rels, queries_per_formula, rng, num_neg, num_hard_neg)))
        out[qt] = groups
    return out


def sample_id_batch(kg, formula, count, rng):
    """Vectorised ids for one formula batch: (anchor_ids [count, a], targets [count], negatives [count]) int64,
    uniform over the proper modes (the tensorised form of `sample_queries`, for benchmark-sized batches)."""
    anchors = np.stack([kg._ent(m, rng, size=count) for m in formula.anchor_modes], axis=1).astype(np.int64)
    targets = kg._ent(formula.target_mode, rng, size=count).astype(np.int64)
    negatives = kg._ent(formula.target_mode, rng, size=count).astype(np.int64)
    return anchors, targets, negatives

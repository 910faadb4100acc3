"""Query data model: formulas, query instances and the typed-graph container.

Same public names and attributes as the reference's `mpqe/graph.py` (Formula :11-58, Query :60-123,
Graph.__init__/_cache_edge_counts :131-170) so pickled datasets and calling code carry over; the samplers
(:211-592) are offline CPU tooling and out of scope.  Written table-driven rather than as the
reference's if-chains: a query type is described once, by where its anchors and variables sit in the nested
relation structure.
"""
from collections import OrderedDict, defaultdict
import random

CHAIN_TYPES = ('1-chain', '2-chain', '3-chain')
INTER_TYPES = ('2-inter', '3-inter')


def _reverse_relation(relation):
    """(from_mode, name, to_mode) -> (to_mode, name, from_mode)  (reference graph.py:4-5)."""
    return (relation[-1], relation[1], relation[0])


def _reverse_edge(edge):
    return (edge[-1], _reverse_relation(edge[1]), edge[0])


def _flat_triples(rels):
    """Relation triples of a possibly nested `rels` structure, in depth-first order."""
    if len(rels) == 3 and not isinstance(rels[0], tuple) and not isinstance(rels[2], tuple):
        return [tuple(rels)]
    out = []
    for r in rels:
        out.extend(_flat_triples(r))
    return out


# where the anchor modes live inside `rels`, per query type: paths of tuple indices ending at a relation triple
_ANCHOR_PATHS = {
    '3-inter_chain': ((0,), (1, 1)),
    '3-chain_inter': ((1, 0), (1, 1)),
}


class Formula(object):
    """A query type plus its typed relation structure (`rels` oriented target -> anchor)."""

    def __init__(self, query_type, rels):
        self.query_type = query_type
        self.rels = rels
        self.target_mode = rels[0][0]
        if query_type in CHAIN_TYPES:
            self.anchor_modes = (rels[-1][-1],)
        elif query_type in INTER_TYPES:
            self.anchor_modes = tuple(r[-1] for r in rels)
        elif query_type in _ANCHOR_PATHS:
            modes = []
            for path in _ANCHOR_PATHS[query_type]:
                r = rels
                for i in path:
                    r = r[i]
                modes.append(r[-1])
            self.anchor_modes = tuple(modes)
        else:
            raise ValueError('unknown query type %r' % (query_type,))

    def get_rels(self):
        return _flat_triples(self.rels)

    def get_nodes(self):
        nodes = []
        for t in _flat_triples(self.rels):
            nodes.extend((t[0], t[2]))
        return nodes

    def _key(self):
        return (self.query_type, self.rels)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return self._key() == other._key()

    def __ne__(self, other):
        return self._key() != other._key()

    def __str__(self):
        return self.query_type + ': ' + str(self.rels)

    __repr__ = __str__


def _cap(samples, limit, inclusive):
    """Keep a negative-sample collection as a list, subsampled to `limit` (random.sample, as the reference)."""
    if samples is None:
        return None
    small = len(samples) <= limit if inclusive else len(samples) < limit
    return list(samples) if small else random.sample(list(samples), limit)


class Query(object):
    """One query instance: `query_graph = (type, edge, ...)`, edge = (node, relation triple, node)."""

    def __init__(self, query_graph, neg_samples, hard_neg_samples, neg_sample_max=100, keep_graph=False):
        qt = query_graph[0]
        edges = query_graph[1:]
        if qt in CHAIN_TYPES:
            rels = tuple(e[1] for e in edges)
            self.anchor_nodes = (edges[-1][-1],)
        elif qt in INTER_TYPES:
            rels = tuple(e[1] for e in edges)
            self.anchor_nodes = tuple(e[-1] for e in edges)
        elif qt in _ANCHOR_PATHS:
            first, (second, third) = edges
            rels = (first[1], (second[1], third[1]))
            self.anchor_nodes = (first[-1], third[-1]) if qt == '3-inter_chain' else (second[-1], third[-1])
        else:
            raise ValueError('unknown query type %r' % (qt,))
        self.formula = Formula(qt, rels)
        self.target_node = edges[0][0]
        self.query_graph = query_graph if keep_graph else None
        self.neg_samples = _cap(neg_samples, neg_sample_max, inclusive=False)
        self.hard_neg_samples = _cap(hard_neg_samples, neg_sample_max, inclusive=True)

    def _edges(self):
        if self.query_graph is None:
            raise Exception('Can only test edge contain if graph is kept. Reinit with keep_graph=True')
        edges = self.query_graph[1:]
        if self.query_graph[0] in _ANCHOR_PATHS:
            edges = (edges[0], edges[1][0], edges[1][1])
        return edges

    def contains_edge(self, edge):
        edges = self._edges()
        return edge in edges or (edge[1], _reverse_relation(edge[1]), edge[0]) in edges

    def get_edges(self):
        edges = self._edges()
        return set(edges).union(_reverse_edge(e) for e in edges)

    def _key(self):
        return (self.formula, self.target_node, self.anchor_nodes)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return self._key() == other._key()

    def __ne__(self, other):
        return hash(self) != hash(other)

    def serialize(self):
        if self.query_graph is None:
            raise Exception('Cannot serialize query loaded with query graph!')
        return (self.query_graph, self.neg_samples, self.hard_neg_samples)

    @staticmethod
    def deserialize(serial_info, keep_graph=False):
        limit = None if serial_info[1] is None else len(serial_info[1])
        return Query(serial_info[0], serial_info[1], serial_info[2], limit, keep_graph=keep_graph)


class Graph(object):
    """Typed multigraph container.  Only what the query-encoding path reads is kept:
    `relations`, `adj_lists`, `feature_dims`, `features`, `full_sets/full_lists` (1-chain negatives, model.py:474),
    `rel_edges` / `mode_weights` (their ORDER defines rel_ids / mode_ids, model.py:326-338)."""

    def __init__(self, features, feature_dims, relations, adj_lists):
        self.features = features
        self.feature_dims = feature_dims
        self.relations = relations
        self.adj_lists = adj_lists
        self.full_sets = defaultdict(set)
        for rel, adjs in adj_lists.items():
            self.full_sets[rel[0]].update(adjs.keys())
        self.full_lists = {mode: list(nodes) for mode, nodes in self.full_sets.items()}
        self._cache_edge_counts()

    def _cache_edge_counts(self):
        self.edges = 0.
        self.rel_edges = OrderedDict()
        for mode, outgoing in self.relations.items():
            for to_mode, name in outgoing:
                rel = (mode, name, to_mode)
                neighbours = self.adj_lists[rel].values()
                self.rel_edges[rel] = float(sum(len(v) for v in neighbours))
                self.edges += float(len(neighbours))
        self.rel_weights = OrderedDict()
        self.mode_edges = OrderedDict()
        for rel, count in self.rel_edges.items():
            self.rel_weights[rel] = count / self.edges
            self.mode_edges[rel[0]] = self.mode_edges.get(rel[0], 0.) + count
        self.mode_weights = OrderedDict((m, c / self.edges) for m, c in self.mode_edges.items())

    def remove_edges(self, edge_list):
        for node, rel, other in edge_list:
            for r, u, v in ((rel, node, other), (_reverse_relation(rel), other, node)):
                try:
                    self.adj_lists[r][u].remove(v)
                except Exception:
                    break
        self._cache_edge_counts()

"""Batch collation and query-graph layout for the R-GCN query encoder.

Public surface mirrors the reference's `mpqe/data_utils.py`: `load_graph` (:18-37), the `load_*queries*` readers
(:155-186), `QueryDataset` (:268-311), `RGCNQueryDataset` (:314-409) and `get_queries_iterator` (:422-426).
What differs is where the work happens: a batch always holds queries of ONE formula, so its graph is one <=4-node
template replicated B times.  Instead of building B `Data` objects and concatenating them on the host, the batch
carries the template and the device materialises `edge_index / edge_type / batch` (bit-exact with PyG's
`Batch.from_data_list`) and the relation-sorted edge layout on demand; the fused layer kernels consume the
template directly and never read the edge list.
"""
from collections import OrderedDict, defaultdict
import pickle

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .graph import Graph, Query, _reverse_relation

# --- query templates (node numbering: anchors 0..a-1, then variables; the target is node a) -------------------
# type -> (num anchors, num variables, edges (src, dst), index into formula.get_rels() per edge,
#          index into formula.get_nodes() per variable, diameter)
TEMPLATES = OrderedDict([
    ('1-chain',       (1, 1, ((0, 1),),                 (0,),      (0,),      1)),
    ('2-chain',       (1, 2, ((0, 2), (2, 1)),          (1, 0),    (0, 2),    2)),
    ('3-chain',       (1, 3, ((0, 3), (3, 2), (2, 1)),  (2, 1, 0), (0, 2, 4), 3)),
    ('2-inter',       (2, 1, ((0, 2), (1, 2)),          (0, 1),    (0,),      1)),
    ('3-inter',       (3, 1, ((0, 3), (1, 3), (2, 3)),  (0, 1, 2), (0,),      1)),
    ('3-inter_chain', (2, 2, ((0, 2), (1, 3), (3, 2)),  (0, 2, 1), (0, 3),    2)),
    ('3-chain_inter', (2, 2, ((0, 3), (1, 3), (3, 2)),  (1, 2, 0), (0, 2),    2)),
])


class QueryTemplate(object):
    """Static shape of one query type."""

    def __init__(self, query_type):
        if query_type not in TEMPLATES:
            raise ValueError('unknown query type %r' % (query_type,))
        a, v, edges, rel_idx, var_idx, diameter = TEMPLATES[query_type]
        self.query_type = query_type
        self.num_anchors, self.num_vars, self.num_nodes = a, v, a + v
        self.src = tuple(e[0] for e in edges)
        self.dst = tuple(e[1] for e in edges)
        self.num_edges = len(edges)
        self.rel_idx, self.var_idx, self.diameter = rel_idx, var_idx, diameter
        self.target_slot = a


_template_cache = {}


def template_of(query_type):
    t = _template_cache.get(query_type)
    if t is None:
        t = _template_cache[query_type] = QueryTemplate(query_type)
    return t


class QueryGraphBatch(object):
    """Stand-in for the PyG `Batch` the reference builds (data_utils.py:402-405).

    Holds the template, the per-edge relation ids and the batch size.  `.edge_index [2,B*E]`, `.edge_type [B*E]`
    and `.batch [B*n]` (int64) are produced by the `mpqe_build_query_graph` kernel the first time they are read on
    a CUDA device; `.relation_sorted()` gives the stable relation-sorted permutation and segment offsets."""

    def __init__(self, template, edge_rel_ids, batch_size, device=None):
        self.template = template
        self.edge_rel_ids = tuple(int(r) for r in edge_rel_ids)
        self.num_graphs = int(batch_size)
        self.num_nodes = self.num_graphs * template.num_nodes
        self.device = torch.device(device) if device is not None else torch.device('cpu')
        self.x = None
        self._arrays = None

    def to(self, device):
        device = torch.device(device)
        if device != self.device:
            self.device = device
            self._arrays = None
            if self.x is not None:
                self.x = self.x.to(device)
        return self

    def _materialise(self):
        if self._arrays is None:
            from . import ops
            t = self.template
            with ops.device_guard(self.device):
                self._arrays = ops.build_query_graph(t.num_nodes, t.src, t.dst, self.edge_rel_ids, self.num_graphs,
                                                     self.device)
        return self._arrays

    @property
    def edge_index(self):
        return self._materialise()[0]

    @property
    def edge_type(self):
        return self._materialise()[1]

    @property
    def batch(self):
        return self._materialise()[2]

    def relation_sorted(self, num_relations):
        """(perm, seg_offsets): stable sort of the edge list by relation id, offsets of each relation's segment."""
        from . import ops
        with ops.device_guard(self.device):
            return ops.relation_sort(self.edge_type, num_relations)


def load_graph(data_dir, embed_dim):
    """`graph_data.pkl` = (rels, adj_lists, node_maps) -> (Graph, {mode: nn.Embedding}, id->row tensor).
    Tables have one spare row and N(0, 1/d) init like the reference (data_utils.py:30-33)."""
    with open(data_dir + '/graph_data.pkl', 'rb') as f:
        rels, adj_lists, node_maps = pickle.load(f)
    return build_graph(rels, adj_lists, node_maps, embed_dim)


def build_graph(rels, adj_lists, node_maps, embed_dim):
    total = sum(len(ids) for ids in node_maps.values())
    id2row = torch.full((total + 1,), -1, dtype=torch.long)
    for ids in node_maps.values():
        idx = torch.as_tensor(np.asarray(ids, dtype=np.int64))
        assert bool((id2row[idx] == -1).all()), 'node id assigned to two modes'
        id2row[idx] = torch.arange(len(ids))
    feature_dims = {m: embed_dim for m in rels}
    feature_modules = {m: torch.nn.Embedding(len(node_maps[m]) + 1, embed_dim) for m in rels}
    for m in rels:
        feature_modules[m].weight.data.normal_(0, 1. / embed_dim)

    def features(nodes, mode):
        return feature_modules[mode](id2row[nodes])

    features.node_maps = id2row
    return Graph(features, feature_dims, rels, adj_lists), feature_modules, id2row


def load_queries(data_file, keep_graph=False):
    with open(data_file, 'rb') as f:
        return [Query.deserialize(info, keep_graph=keep_graph) for info in pickle.load(f)]


def load_queries_by_formula(data_file):
    with open(data_file, 'rb') as f:
        raw = pickle.load(f)
    return queries_by_formula(raw)


def queries_by_formula(raw_queries):
    out = defaultdict(lambda: defaultdict(list))
    for info in raw_queries:
        q = Query.deserialize(info)
        out[q.formula.query_type][q.formula].append(q)
    return out


def load_queries_by_type(data_file, keep_graph=True):
    out = defaultdict(list)
    for q in load_queries(data_file, keep_graph=keep_graph):
        out[q.formula.query_type].append(q)
    return out


def load_test_queries_by_formula(data_file):
    with open(data_file, 'rb') as f:
        raw = pickle.load(f)
    out = {'full_neg': defaultdict(lambda: defaultdict(list)), 'one_neg': defaultdict(lambda: defaultdict(list))}
    for info in raw:
        q = Query.deserialize(info)
        out['full_neg' if len(info[1]) > 1 else 'one_neg'][q.formula.query_type][q.formula].append(q)
    return out


class QueryDataset(Dataset):
    """{formula: [queries]} of one query type; a batch is a contiguous slice of ONE formula's list, the formula
    drawn with probability proportional to its query count (reference data_utils.py:293-311)."""

    def __init__(self, queries, *args, **kwargs):
        self.queries = queries
        self.num_formula_queries = OrderedDict((f, len(qs)) for f, qs in queries.items())
        self.num_queries = sum(self.num_formula_queries.values())
        self.max_num_queries = max(self.num_formula_queries.values())

    def __len__(self):
        return self.max_num_queries

    def __getitem__(self, index):
        return index

    def collate_fn(self, idx_list):
        counts = np.fromiter(self.num_formula_queries.values(), dtype=np.float64)
        pick = int(np.argmax(np.random.multinomial(1, counts / float(self.num_queries))))
        formula = list(self.num_formula_queries)[pick]
        n = self.num_formula_queries[formula]
        start = idx_list[0] % n
        end = min((idx_list[-1] + 1) % n, n)
        if end <= start:
            end = n
        return formula, self.queries[formula][start:end]


class RGCNQueryDataset(QueryDataset):
    """Adds the query-graph layout to each batch (reference data_utils.py:314-409)."""
    query_edge_indices = {qt: [list(template_of(qt).src), list(template_of(qt).dst)] for qt in TEMPLATES}
    query_diameters = {qt: template_of(qt).diameter for qt in TEMPLATES}
    query_edge_label_idx = {qt: list(template_of(qt).rel_idx) for qt in TEMPLATES}
    variable_node_idx = {qt: list(template_of(qt).var_idx) for qt in TEMPLATES}

    def __init__(self, queries, enc_dec):
        super(RGCNQueryDataset, self).__init__(queries)
        self.mode_ids = enc_dec.mode_ids
        self.rel_ids = enc_dec.rel_ids

    def collate_fn(self, idx_list):
        formula, queries = super(RGCNQueryDataset, self).collate_fn(idx_list)
        anchor_ids, var_ids, graph = RGCNQueryDataset.get_query_graph(formula, queries, self.rel_ids, self.mode_ids)
        return formula, queries, anchor_ids, var_ids, graph

    @staticmethod
    def formula_layout(formula, rel_ids, mode_ids):
        """(template, var mode ids, relation id per template edge) of a formula."""
        t = template_of(formula.query_type)
        nodes = formula.get_nodes()
        rels = formula.get_rels()
        var_ids = [mode_ids[nodes[i]] for i in t.var_idx]
        edge_rel = [rel_ids[_reverse_relation(rels[i])] for i in t.rel_idx]
        return t, var_ids, edge_rel

    @staticmethod
    def get_query_graph(formula, queries, rel_ids, mode_ids):
        t, var_ids, edge_rel = RGCNQueryDataset.formula_layout(formula, rel_ids, mode_ids)
        anchor_ids = torch.as_tensor(np.asarray([q.anchor_nodes for q in queries], dtype=np.int64)
                                     .reshape(len(queries), t.num_anchors))
        return anchor_ids, torch.tensor(var_ids, dtype=torch.long), QueryGraphBatch(t, edge_rel, len(queries))


def make_data_iterator(data_loader):
    while True:
        for batch in data_loader:
            yield batch


def get_queries_iterator(queries, batch_size, enc_dec=None):
    dataset = RGCNQueryDataset(queries, enc_dec)
    loader = DataLoader(dataset, batch_size, shuffle=False, collate_fn=dataset.collate_fn)
    return make_data_iterator(loader)

"""Import the UNMODIFIED reference (`/root/reference/mpqe`) in this container, behind `oracle/shim`.

TEST INFRASTRUCTURE ONLY: used by `oracle/make_golden.py` and by the `not gpu` tests that pin
`oracle/mpqe_oracle.py`.  `/root/reference` does not exist on the GPU box, so nothing that runs there imports
this module (`available()` is False there).  What is restated rather than executed is only the third-party
layer the reference leaves un-vendored and unpinned (torch_scatter / torch_geometric, see the shim headers) and
`numpy.int` (removed from numpy>=1.24; used at /root/reference/mpqe/data_utils.py:382, 392).
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get('MPQE_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')
_cache = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'mpqe', 'model.py'))


def load():
    """Returns the reference's modules as a dict: graph, data_utils, encoders, model, utils."""
    if _cache:
        return _cache
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int  # noqa: NPY001 - reference relies on the removed alias
    sys.dont_write_bytecode = True
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ('graph', 'data_utils', 'encoders', 'model', 'utils'):
        _cache[name] = importlib.import_module('mpqe.' + name)
    return _cache


def build_reference_model(raw_graph, embed_dim, readout, num_layers, adaptive, shared_layers=False,
                          scatter_op='add', weight_decay=0.0, seed=0):
    """Reference `load_graph` (data_utils.py:18-37, minus the unpickle) + `DirectEncoder` + `RGCNEncoderDecoder`."""
    import torch
    ref = load()
    rels, adj_lists, node_maps = raw_graph
    torch.manual_seed(seed)
    counts = {m: len(node_maps[m]) for m in node_maps}
    total = sum(counts.values())
    id2row = torch.ones(total + 1, dtype=torch.long).fill_(-1)
    for m, ids in node_maps.items():
        id2row[torch.tensor(ids, dtype=torch.long)] = torch.arange(len(ids))
    feature_dims = {m: embed_dim for m in rels}
    feature_modules = {m: torch.nn.Embedding(counts[m] + 1, embed_dim) for m in rels}
    for m in rels:
        feature_modules[m].weight.data.normal_(0, 1. / embed_dim)
    features = lambda nodes, mode: feature_modules[mode](id2row[nodes])  # noqa: E731
    graph = ref['graph'].Graph(features, feature_dims, rels, adj_lists)
    enc = ref['encoders'].DirectEncoder(graph.features, feature_modules)
    model = ref['model'].RGCNEncoderDecoder(graph, enc, readout=readout, scatter_op=scatter_op, dropout=0,
                                            weight_decay=weight_decay, num_layers=num_layers,
                                            shared_layers=shared_layers, adaptive=adaptive)
    return model, graph, id2row


def deserialize_queries(raw_queries):
    ref = load()
    return [ref['graph'].Query.deserialize(r) for r in raw_queries]

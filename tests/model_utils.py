"""Build the product model from oracle-style parameters, and run it side by side with the oracle."""
import numpy as np
import torch

from mpqe_b200 import data_utils, encoders, model as M
from mpqe_b200.graph import Query
from oracle import mpqe_oracle as O


def build_model(raw_graph, cfg, params, device, d=128, sparse_grad=False):
    rels, adj_lists, node_maps = raw_graph
    graph, feature_modules, id2row = data_utils.build_graph(rels, adj_lists, node_maps, d)
    enc = encoders.DirectEncoder(graph.features, feature_modules, sparse_grad=sparse_grad)
    model = M.RGCNEncoderDecoder(graph, enc, readout=cfg.readout, scatter_op=cfg.scatter_op, dropout=0,
                                 weight_decay=cfg.weight_decay, num_layers=cfg.num_layers,
                                 shared_layers=cfg.shared_layers, adaptive=cfg.adaptive)
    sd = model.state_dict()
    for k in sd:
        src = k
        if cfg.shared_layers and k.startswith('layers.'):
            parts = k.split('.')
            parts[1] = '0'
            src = '.'.join(parts)
        sd[k] = params[src].clone()
    model.load_state_dict(sd)
    return model.to(device)


def queries_from_ids(query_type, rels, anchor_ids, targets):
    """Minimal Query objects (formula + anchors + target) for a batch given as id arrays."""
    out = []
    for anchors, t in zip(np.asarray(anchor_ids).tolist(), np.asarray(targets).tolist()):
        out.append(_make_query(query_type, rels, anchors, t))
    return out


def _make_query(qt, rels, anchors, t):
    if qt.endswith('-chain'):
        nodes = [t] + [-1] * (len(rels) - 1) + [anchors[0]]
        qg = (qt,) + tuple((nodes[i], rels[i], nodes[i + 1]) for i in range(len(rels)))
    elif qt in ('2-inter', '3-inter'):
        qg = (qt,) + tuple((t, r, a) for r, a in zip(rels, anchors))
    elif qt == '3-inter_chain':
        r1, (r2, r3) = rels
        qg = (qt, (t, r1, anchors[0]), ((t, r2, -1), (-1, r3, anchors[1])))
    else:
        r1, (r2, r3) = rels
        qg = (qt, (t, r1, -1), ((-1, r2, anchors[0]), (-1, r3, anchors[1])))
    return Query(qg, None, None)


def oracle_loss_and_grads(params, cfg, spec, anchor_ids, rel_ids, mode_ids, id2row, targets, negs, margin=1.0):
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    a_ids, var_ids, ei, et, batch = O.query_graph(spec, anchor_ids.tolist(), rel_ids, mode_ids)
    loss = O.margin_loss(p, cfg, spec, a_ids, var_ids, ei, et, batch, id2row, targets, negs, margin)
    loss.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32)) for k, v in p.items()}
    return loss.item(), grads


def model_grads(model):
    out = {}
    for k, prm in model.named_parameters():
        g = prm.grad
        if g is None:
            out[k] = np.zeros(tuple(prm.shape), np.float32)
        else:
            out[k] = (g.to_dense() if g.is_sparse else g).detach().cpu().numpy()
    return out


def train_step_grads(ts, model, res, device):
    """{state_dict name: gradient tensor} of a TrainStep result: views of the flat dense bucket plus the combined
    row-sparse entity gradients scattered back into dense tables (global row = table offset + row)."""
    G = res.dense
    got = {}
    layers = model.distinct_layers()
    for name, prm in model.named_parameters():
        for li, layer in enumerate(layers):
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
    dense_tables = torch.zeros(ts.total_rows, urows.shape[1], device=device)
    dense_tables[uid[:k]] = urows[:k]
    for mode, off in ts.table_offsets.items():
        rows = model.enc.table(mode).shape[0]
        got['enc.feat-%s.weight' % mode] = dense_tables[off:off + rows]
    return got, uid[:k]

"""CPU oracle for the MPQE query-encoding hot path: a restatement of the reference's algorithm on plain tensors.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this file; nothing under `mpqe_b200/` does.  It follows the reference line by
line (same operators in the same order, including the per-edge weight gather + bmm that dominates the
reference's run time), so it doubles as the timed CPU baseline (`cpu_baseline.kind = "port"`).

Floating point: torch fp32 on CPU, gradients by autograd (a torch reference is kept because the path is a
floating-point kernel).  Integer artefacts (edge layout, relation sort, arg-max, rank counts) are exact.

PARITY PINNING.  The reference has no tests or golden vectors for this path, and its
third-party arithmetic (torch_geometric ~1.4 `MessagePassing`, torch_scatter 2.x `scatter_*`) is un-vendored and
unpinned (/root/reference/README.md:22).  This oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF
run in the build container behind `oracle/shim` (which restates only that third-party layer):
`tests/golden/*.npz`, written by `oracle/make_golden.py`, and a live comparison in `tests/test_oracle_vs_reference.py`
when `/root/reference` is present.  The scatter_max tie-break ("smallest node row among the maxima") is this
project's stated contract; torch_scatter's own behaviour on ties could not be checked offline.

Reference lines followed by each function are given in its docstring (paths relative to /root/reference/mpqe/).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------------------
# Query templates (data_utils.py:325-362).  Node numbering: anchors 0..a-1, then variables; target = node a.
# ---------------------------------------------------------------------------------------------------------------
EDGE_INDEX = {
    '1-chain': ((0,), (1,)),
    '2-chain': ((0, 2), (2, 1)),
    '3-chain': ((0, 3, 2), (3, 2, 1)),
    '2-inter': ((0, 1), (2, 2)),
    '3-inter': ((0, 1, 2), (3, 3, 3)),
    '3-inter_chain': ((0, 1, 3), (2, 3, 2)),
    '3-chain_inter': ((0, 1, 3), (3, 3, 2)),
}
DIAMETER = {'1-chain': 1, '2-chain': 2, '3-chain': 3, '2-inter': 1, '3-inter': 1,
            '3-inter_chain': 2, '3-chain_inter': 2}
EDGE_LABEL = {'1-chain': (0,), '2-chain': (1, 0), '3-chain': (2, 1, 0), '2-inter': (0, 1),
              '3-inter': (0, 1, 2), '3-inter_chain': (0, 2, 1), '3-chain_inter': (1, 2, 0)}
VARIABLE_NODE = {'1-chain': (0,), '2-chain': (0, 2), '3-chain': (0, 2, 4), '2-inter': (0,),
                 '3-inter': (0,), '3-chain_inter': (0, 2), '3-inter_chain': (0, 3)}


def reverse_relation(rel):
    """graph.py:4-5."""
    return (rel[-1], rel[1], rel[0])


def _flatten(nested):
    out = []
    for item in nested:
        if isinstance(item, tuple):
            out.extend(_flatten(item))
        else:
            out.append(item)
    return out


def formula_spec(query_type, rels):
    """Formula algebra, graph.py:11-46: target mode, anchor modes, flat relation triples, node-mode list."""
    flat = _flatten(rels)
    triples = [tuple(flat[i:i + 3]) for i in range(0, len(flat), 3)]
    nodes = []
    for t in triples:
        nodes.extend([t[0], t[2]])
    if query_type in ('1-chain', '2-chain', '3-chain'):
        anchor_modes = (rels[-1][-1],)
    elif query_type in ('2-inter', '3-inter'):
        anchor_modes = tuple(r[-1] for r in rels)
    elif query_type == '3-inter_chain':
        anchor_modes = (rels[0][-1], rels[1][-1][-1])
    elif query_type == '3-chain_inter':
        anchor_modes = (rels[1][0][-1], rels[1][1][-1])
    else:
        raise ValueError(query_type)
    return dict(query_type=query_type, rels=rels, target_mode=rels[0][0], anchor_modes=anchor_modes,
                triples=triples, nodes=nodes)


def query_anchors_target(raw_query_graph):
    """Anchor / target extraction, graph.py:60-78."""
    qt = raw_query_graph[0]
    g = raw_query_graph
    if qt in ('1-chain', '2-chain', '3-chain'):
        rels = tuple(g[i][1] for i in range(1, len(g)))
        anchors = (g[-1][-1],)
    elif qt in ('2-inter', '3-inter'):
        rels = tuple(g[i][1] for i in range(1, len(g)))
        anchors = tuple(g[i][-1] for i in range(1, len(g)))
    elif qt == '3-inter_chain':
        rels = (g[1][1], (g[2][0][1], g[2][1][1]))
        anchors = (g[1][-1], g[2][-1][-1])
    elif qt == '3-chain_inter':
        rels = (g[1][1], (g[2][0][1], g[2][1][1]))
        anchors = (g[2][0][-1], g[2][1][-1])
    else:
        raise ValueError(qt)
    return qt, rels, anchors, g[1][0]


def schema_ids(rels):
    """mode_ids / rel_ids as the model enumerates them (model.py:326-338 over graph.py:153-170):
    relations in nested `rels` order; modes in order of first appearance as a relation's source."""
    rel_ids, mode_ids = OrderedDict(), OrderedDict()
    for m in rels:
        for (to, name) in rels[m]:
            rel_ids[(m, name, to)] = len(rel_ids)
            if m not in mode_ids:
                mode_ids[m] = len(mode_ids)
    return mode_ids, rel_ids


def id_to_row(node_maps):
    """Global id -> per-mode table row, -1 where absent (data_utils.py:23-28)."""
    total = sum(len(v) for v in node_maps.values())
    out = torch.full((total + 1,), -1, dtype=torch.long)
    for ids in node_maps.values():
        out[torch.as_tensor(np.asarray(ids), dtype=torch.long)] = torch.arange(len(ids))
    return out


def query_graph(spec, anchor_nodes, rel_ids, mode_ids):
    """Batch layout, data_utils.py:377-409 (+ PyG Batch.from_data_list): returns int64
    anchor_ids[B,a], var_ids[v], edge_index[2,B*E], edge_type[B*E], batch[B*n]."""
    qt = spec['query_type']
    anchor_ids = torch.as_tensor(np.asarray(anchor_nodes, dtype=np.int64).reshape(len(anchor_nodes), -1))
    B, a = anchor_ids.shape
    var_ids = torch.tensor([mode_ids[spec['nodes'][i]] for i in VARIABLE_NODE[qt]], dtype=torch.long)
    n = a + var_ids.numel()
    tmpl = torch.tensor(EDGE_INDEX[qt], dtype=torch.long)
    etype = torch.tensor([rel_ids[reverse_relation(spec['triples'][i])] for i in EDGE_LABEL[qt]],
                         dtype=torch.long)
    shift = (torch.arange(B, dtype=torch.long) * n).view(B, 1, 1)
    edge_index = (tmpl.unsqueeze(0) + shift).permute(1, 0, 2).reshape(2, -1)
    edge_type = etype.repeat(B)
    batch = torch.arange(B, dtype=torch.long).repeat_interleave(n)
    return anchor_ids, var_ids, edge_index, edge_type, batch


def relation_sorted_layout(edge_type, num_relations):
    """Relation-sorted edge layout (new artefact named by the north star; defined as the STABLE sort of the
    PyG-ordered edge list by relation id): permutation[int64, nE] and segment offsets[int64, R+1]."""
    perm = torch.sort(edge_type, stable=True)[1]
    counts = torch.bincount(edge_type, minlength=num_relations)
    offsets = torch.zeros(num_relations + 1, dtype=torch.long)
    offsets[1:] = torch.cumsum(counts, 0)
    return perm, offsets


# ---------------------------------------------------------------------------------------------------------------
# Floating-point path
# ---------------------------------------------------------------------------------------------------------------
def direct_encode(table, rows):
    """DirectEncoder.forward, encoders.py:41-43: gather, transpose to [d,B], divide by the column L2 norm (no eps)."""
    embeds = F.embedding(rows, table).t()
    norm = embeds.norm(p=2, dim=0, keepdim=True)
    return embeds.div(norm.expand_as(embeds))


def rgcn_conv(x, edge_index, edge_type, basis, root, bias, att=None):
    """RGCNConv.forward/message/update, model.py:269-305, through PyG propagate (aggr='add', source->target):
    per-edge weight gather, bmm, scatter-add at edge_index[1] with dim_size = x.size(0), + x@root + bias."""
    d_in, d_out = root.shape
    if att is None:
        w = basis.view(basis.size(0), -1)
        num_rel = basis.size(0)
    else:
        w = torch.matmul(att, basis.view(basis.size(0), -1))
        num_rel = att.size(0)
    w = w.view(num_rel, d_in, d_out)
    x_j = x.index_select(0, edge_index[0])
    w_e = torch.index_select(w, 0, edge_type)
    msg = torch.bmm(x_j.unsqueeze(1), w_e).squeeze(-2)
    aggr = torch.zeros(x.size(0), d_out, dtype=x.dtype).index_add(0, edge_index[1], msg)
    out = aggr + torch.matmul(x, root)
    if bias is not None:
        out = out + bias
    return out


def scatter_max_first(src, index, dim_size):
    """scatter_max with this project's tie-break: arg = smallest source row among the maxima (model.py:384)."""
    idx = index.view(-1, 1).expand_as(src)
    val = torch.zeros(dim_size, src.size(1), dtype=src.dtype).scatter_reduce(
        0, idx, src, 'amax', include_self=False)
    rows = torch.arange(src.size(0)).view(-1, 1).expand_as(src)
    cand = torch.where(src == val.index_select(0, index), rows, torch.full_like(rows, src.size(0)))
    arg = torch.full((dim_size, src.size(1)), src.size(0), dtype=torch.long).scatter_reduce(
        0, idx, cand, 'amin', include_self=True)
    return val, arg


def _scatter(op, x, index, dim_size):
    if op == 'add':
        return torch.zeros(dim_size, x.size(1), dtype=x.dtype).index_add(0, index, x)
    if op == 'mean':
        tot = torch.zeros(dim_size, x.size(1), dtype=x.dtype).index_add(0, index, x)
        cnt = torch.zeros(dim_size, dtype=x.dtype).index_add(0, index, torch.ones(index.numel(), dtype=x.dtype))
        return tot / cnt.clamp(min=1).view(-1, 1)
    if op == 'max':
        return scatter_max_first(x, index, dim_size)[0]
    raise ValueError('Unknown scatter op %s' % op)


def mlp(x, p, prefix='readout.layers.'):
    """nn.Sequential(Linear, ReLU, Linear), model.py:500-504."""
    h = F.relu(F.linear(x, p[prefix + '0.weight'], p[prefix + '0.bias']))
    return F.linear(h, p[prefix + '2.weight'], p[prefix + '2.bias'])


def readout(name, h, batch, B, n, a, p, scatter_op='add'):
    """sum / max / mp readouts (model.py:380-398), MLPReadout for mlp+concat (:497-515), TargetMLPReadout (:518-553).
    Returns (query embedding [B,d], argmax or None)."""
    if name == 'sum':
        return torch.zeros(B, h.size(1), dtype=h.dtype).index_add(0, batch, h), None
    if name == 'max':
        return scatter_max_first(h, batch, B)
    if name == 'mp':
        return h.reshape(B, n, -1)[:, a], None
    if name in ('mlp', 'concat'):
        return _scatter(scatter_op, mlp(h, p), batch, B), None
    if name == 'targetmlp':
        keep = [i for i in range(n) if i != a]
        hb = h.reshape(B, n, -1)
        non_targets = hb[:, keep]
        targets = hb[:, a:a + 1].expand_as(non_targets)
        x = torch.cat((targets, non_targets), dim=-1).reshape(B * (n - 1), -1)
        bidx = batch.reshape(B, n)[:, keep].reshape(-1)
        return _scatter(scatter_op, mlp(x, p), bidx, B), None
    raise ValueError('Unknown readout function %s' % name)


class Config(object):
    """The constructor arguments of RGCNEncoderDecoder that change the arithmetic (model.py:314-316)."""

    def __init__(self, readout='sum', num_layers=2, adaptive=False, shared_layers=False, scatter_op='add',
                 weight_decay=0.0):
        self.readout, self.num_layers, self.adaptive = readout, num_layers, adaptive
        self.shared_layers, self.scatter_op, self.weight_decay = shared_layers, scatter_op, weight_decay


def layer_params(p, i):
    return p['layers.%d.basis' % i], p['layers.%d.root' % i], p['layers.%d.bias' % i], p.get('layers.%d.att' % i)


def encode_queries(p, cfg, spec, anchor_ids, var_ids, edge_index, edge_type, batch, id2row, want_argmax=False):
    """RGCNEncoderDecoder.forward up to the readout, model.py:414-449.  `p` maps state_dict names to tensors."""
    B, a = anchor_ids.shape
    n = a + var_ids.numel()
    d = p['mode_embeddings.weight'].size(1)
    x = torch.empty(B, n, d)
    for i, mode in enumerate(spec['anchor_modes']):
        x[:, i] = direct_encode(p['enc.feat-%s.weight' % mode], id2row[anchor_ids[:, i]]).t()
    x[:, a:] = F.embedding(var_ids, p['mode_embeddings.weight'])
    h = x.reshape(-1, d)
    if cfg.adaptive:
        passes = DIAMETER[spec['query_type']]
        if passes > cfg.num_layers:
            raise ValueError('RGCN is adaptive with %d layers, but query requires %d.' % (cfg.num_layers, passes))
    else:
        passes = cfg.num_layers
    last = cfg.num_layers - 1
    kept = []
    for i in range(passes - 1):
        li = 0 if cfg.shared_layers else i
        basis, root, bias, att = layer_params(p, li)
        h = F.relu(rgcn_conv(h, edge_index, edge_type, basis, root, bias, att))
        if cfg.readout == 'concat':
            kept.append(h)
    basis, root, bias, att = layer_params(p, 0 if cfg.shared_layers else last)
    h = rgcn_conv(h, edge_index, edge_type, basis, root, bias, att)
    if cfg.readout == 'concat':
        kept.append(h)
        h = torch.cat(kept, dim=1)
    out, arg = readout(cfg.readout, h, batch, B, n, a, p, cfg.scatter_op)
    return (out, arg) if want_argmax else out


def forward_scores(p, cfg, spec, anchor_ids, var_ids, edge_index, edge_type, batch, id2row, target_nodes,
                   neg_nodes=None, neg_lengths=None):
    """RGCNEncoderDecoder.forward, model.py:400-462: cosine of the query embedding against targets and, in eval,
    against the ragged negatives (`repeat_interleave` by neg_lengths); scores = cat(pos, neg)."""
    out = encode_queries(p, cfg, spec, anchor_ids, var_ids, edge_index, edge_type, batch, id2row)
    table = p['enc.feat-%s.weight' % spec['target_mode']]
    tgt = direct_encode(table, id2row[torch.as_tensor(target_nodes, dtype=torch.long)]).t()
    scores = F.cosine_similarity(out, tgt, dim=1)
    if neg_nodes is not None:
        neg = direct_encode(table, id2row[torch.as_tensor(neg_nodes, dtype=torch.long)]).t()
        rep = out.repeat_interleave(torch.as_tensor(neg_lengths, dtype=torch.long), dim=0)
        scores = torch.cat((scores, F.cosine_similarity(rep, neg)), dim=0)
    return scores


def margin_loss(p, cfg, spec, anchor_ids, var_ids, edge_index, edge_type, batch, id2row, target_nodes, neg_nodes,
                margin=1.0):
    """RGCNEncoderDecoder.margin_loss, model.py:478-492, with the negatives passed in (the reference draws them
    with random.choice, :470-476).  Two full forwards, exactly like the reference."""
    affs = forward_scores(p, cfg, spec, anchor_ids, var_ids, edge_index, edge_type, batch, id2row, target_nodes)
    neg_affs = forward_scores(p, cfg, spec, anchor_ids, var_ids, edge_index, edge_type, batch, id2row, neg_nodes)
    loss = torch.clamp(margin - (affs - neg_affs), min=0).mean()
    if cfg.readout in ('mlp', 'targetmlp', 'concat') and cfg.weight_decay > 0:
        reg = 0
        for k in ('readout.layers.0.weight', 'readout.layers.0.bias', 'readout.layers.2.weight',
                  'readout.layers.2.bias'):
            reg = reg + torch.norm(p[k])
        loss = loss + cfg.weight_decay * reg
    return loss


# ---------------------------------------------------------------------------------------------------------------
# Metrics (utils.py:25-95)
# ---------------------------------------------------------------------------------------------------------------
def rank_counts(pos, neg, lengths):
    """Integer core of scipy.stats.percentileofscore(kind='rank') as used by utils.py:25-32:
    left = #(neg < pos), right = #(neg <= pos) per query.  pos [B], neg [sum(lengths)] (numpy or tensor)."""
    pos = np.asarray(pos)
    neg = np.asarray(neg)
    left = np.zeros(len(lengths), dtype=np.int64)
    right = np.zeros(len(lengths), dtype=np.int64)
    off = 0
    for i, ln in enumerate(lengths):
        seg = neg[off:off + ln]
        left[i] = np.count_nonzero(seg < pos[i])
        right[i] = np.count_nonzero(seg <= pos[i])
        off += ln
    return left, right


def percentile_from_counts(left, right, lengths):
    """(left + right + (left < right)) * 50 / n  -- scipy's 'rank' percentile."""
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    n = np.asarray(lengths, dtype=np.float64)
    return (left + right + (left < right)) * (50.0 / n)


def reciprocal_rank_from_counts(num_le_or_ge):
    """MRR is NOT in the reference; defined here from the bit-exact counts: rank = 1 + #(candidates scoring
    strictly higher than the positive) -> callers pass that count."""
    return 1.0 / (1.0 + np.asarray(num_le_or_ge, dtype=np.float64))


def auc(labels, scores):
    """roc_auc_score(labels, nan_to_num(scores)) (utils.py:65-68) via the rank-sum (Mann-Whitney) identity with
    average ranks for ties; avoids importing sklearn on the GPU box."""
    labels = np.asarray(labels).astype(bool)
    scores = np.nan_to_num(np.asarray(scores, dtype=np.float64))
    order = np.argsort(scores, kind='mergesort')
    s = scores[order]
    ranks = np.empty(len(s), dtype=np.float64)
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r = np.empty_like(ranks)
    r[order] = ranks
    npos = labels.sum()
    nneg = len(labels) - npos
    return (r[labels].sum() - npos * (npos + 1) / 2.0) / (npos * nneg)


# ---------------------------------------------------------------------------------------------------------------
# Parameter construction shared by tests / golden generation (deterministic, numpy RandomState only)
# ---------------------------------------------------------------------------------------------------------------
def init_params(rels, node_maps, cfg, d=128, seed=0):
    """Parameters with the reference's shapes and init DISTRIBUTIONS (data_utils.py:31-33 N(0,1/d) tables;
    model.py:321 N(0,1) mode embeddings; model.py:258-267 U(+-1/sqrt(R*d)) basis/root/bias; nn.Linear default
    for the readout MLP), drawn from numpy RandomState so fixtures do not depend on torch's RNG stream."""
    rng = np.random.RandomState(seed)
    mode_ids, rel_ids = schema_ids(rels)
    R = len(rel_ids)
    p = OrderedDict()

    def t(a):
        return torch.tensor(np.asarray(a, dtype=np.float32))

    for m in rels:
        p['enc.feat-%s.weight' % m] = t(rng.normal(0, 1.0 / d, size=(len(node_maps[m]) + 1, d)))
    p['mode_embeddings.weight'] = t(rng.normal(0, 1.0, size=(len(mode_ids), d)))
    bound = 1.0 / np.sqrt(R * d)
    for i in range(1 if cfg.shared_layers else cfg.num_layers):
        p['layers.%d.basis' % i] = t(rng.uniform(-bound, bound, size=(R, d, d)))
        p['layers.%d.root' % i] = t(rng.uniform(-bound, bound, size=(d, d)))
        p['layers.%d.bias' % i] = t(rng.uniform(-bound, bound, size=(d,)))
    if cfg.readout in ('mlp', 'concat', 'targetmlp'):
        d_in = {'mlp': d, 'concat': d * cfg.num_layers, 'targetmlp': 2 * d}[cfg.readout]
        for name, (o, i_) in (('0', (d, d_in)), ('2', (d, d))):
            b = 1.0 / np.sqrt(i_)
            p['readout.layers.%s.weight' % name] = t(rng.uniform(-b, b, size=(o, i_)))
            p['readout.layers.%s.bias' % name] = t(rng.uniform(-b, b, size=(o,)))
    return p

"""MPQE query encoder + scorer on B200: R-GCN over batched query graphs, readout, cosine scoring, margin loss.

Mirror of the reference's `mpqe/model.py` R-GCN path (`RGCNConv` :206-310, `RGCNEncoderDecoder` :313-494,
`MLPReadout` :497-515, `TargetMLPReadout` :518-553): same constructor arguments, attributes, state_dict names,
`forward(...) -> scores` and `margin_loss(...) -> loss`, same exceptions.

How it runs is different.  Every query of a batch shares one <=4-node template, so a layer pass is a short list
of dense terms  out[:, dst] += h[:, src] @ W[rel]  (+ the self-loop term, bias, ReLU); `engine` turns a batch
into such term lists and drives the fused CUDA layer kernel (`mpqe_layer_forward`) once per pass for ALL groups,
with the readout folded into the last pass (sum: every term accumulates into one output row; target-message:
only the target's terms are computed at all).  The backward is the same kernel on transposed matrices plus a
deterministic weight-gradient kernel; entity-table gradients are row-sparse.  The reference instead gathers a
d x d matrix per EDGE and runs bmm (:292-294), and runs the whole encoder twice per loss (:478-482); here the
query embedding is computed once and scored against positives and negatives.
"""
import math
import random

import torch
import torch.nn as nn

from . import ops
from .data_utils import QueryGraphBatch, RGCNQueryDataset, template_of
from .encoders import DirectEncoder, table_gradient
from .ops import D, EPI_MASK, EPI_NONE, EPI_RELU, Group, Term

MLP_READOUTS = ('mlp', 'targetmlp', 'concat')


def _uniform(size, tensor):
    """PyG inits.uniform: U(+-1/sqrt(size)) (reference model.py:263-267)."""
    if tensor is not None:
        bound = 1.0 / math.sqrt(size)
        tensor.data.uniform_(-bound, bound)


class _BasisFn(torch.autograd.Function):
    """W = att @ basis over the flattened [in, out] matrices (reference model.py:281-284) and its backward
    (d att = dW . basis, d basis = att^T @ dW) on the library's small-inner-dimension kernels."""

    @staticmethod
    def forward(ctx, att, basis):
        ctx.save_for_backward(att, basis)
        with ops.device_guard(basis.device):
            return ops.small_k_matmul(att.detach().contiguous(), basis.detach().contiguous())

    @staticmethod
    def backward(ctx, dw):
        att, basis = ctx.saved_tensors
        with ops.device_guard(basis.device):
            dw = dw.contiguous()
            return (ops.rows_dot(dw, basis.detach().contiguous()),
                    ops.small_k_matmul(att.detach().contiguous(), dw, transpose_a=True))


class RGCNConv(nn.Module):
    """Relational graph convolution  x'_i = x_i @ root + sum_{j ->r i} x_j @ W_r + bias  (aggr='add').

    Parameters as in the reference: `basis [R | num_bases, in, out]`, `att [R, num_bases] | None`, `root [in, out]`,
    `bias [out]`.  `forward(x, edge_index, edge_type)` takes either the reference's tensors plus `graph=` (the
    `QueryGraphBatch` that produced them) or just a `QueryGraphBatch` as `edge_index`."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases, bias=True):
        super(RGCNConv, self).__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases = num_relations, num_bases
        if num_bases == 0:
            self.basis = nn.Parameter(torch.Tensor(num_relations, in_channels, out_channels))
            self.att = None
        else:
            self.basis = nn.Parameter(torch.Tensor(num_bases, in_channels, out_channels))
            self.att = nn.Parameter(torch.Tensor(num_relations, num_bases))
        self.root = nn.Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        size = (self.num_relations if self.att is None else self.num_bases) * self.in_channels
        _uniform(size, self.att)
        _uniform(size, self.basis)
        _uniform(size, self.root)
        _uniform(size, self.bias)

    def relation_weights(self):
        """[R, in, out] weights; with a basis decomposition W_r = sum_b att[r, b] * basis[b] (model.py:281-284)."""
        if self.att is None:
            return self.basis
        return _BasisFn.apply(self.att, self.basis)

    def forward(self, x, edge_index, edge_type=None, edge_norm=None, graph=None):
        if isinstance(edge_index, QueryGraphBatch):
            graph = edge_index
        if edge_norm is not None:
            raise NotImplementedError('edge_norm is never passed by MPQE (model.py:436, 441)')
        if graph is None:
            # the reference's call signature (model.py:269): recover the template from the edge list itself.  Costs a
            # device -> host copy of the indices; callers on the hot path pass the QueryGraphBatch instead.
            graph = infer_template_batch(x.shape[0], edge_index, edge_type)
        if self.in_channels != D or self.out_channels != D:
            raise ValueError('kernels are specialised for %d channels' % D)
        return _ConvFn.apply(x, self.relation_weights(), self.root, self.bias, graph)

    def __repr__(self):
        return '{}({}, {}, num_relations={})'.format(self.__class__.__name__, self.in_channels, self.out_channels,
                                                     self.num_relations)


class _InferredTemplate(object):
    def __init__(self, num_nodes, src, dst):
        self.num_nodes, self.num_edges, self.src, self.dst = int(num_nodes), len(src), list(src), list(dst)


class _InferredBatch(object):
    """What `_ConvFn` needs of a QueryGraphBatch, recovered from (edge_index, edge_type)."""

    def __init__(self, template, edge_rel_ids, num_graphs):
        self.template, self.edge_rel_ids, self.num_graphs = template, list(edge_rel_ids), int(num_graphs)


def infer_template_batch(num_rows, edge_index, edge_type):
    """A PyG-style batch of B identical graphs (`Batch.from_data_list`, reference data_utils.py:402-405) is B copies of
    one template with node offsets b*n: find the largest B for which (edge_index, edge_type) has that form and return
    the template.  Raises NotImplementedError for edge lists that are not such a batch (general graphs are outside
    the MPQE path) or whose template exceeds the kernels' limits."""
    ei = edge_index.detach().cpu().numpy()
    et = edge_type.detach().cpu().numpy()
    total_e = int(ei.shape[1])
    if ei.ndim != 2 or ei.shape[0] != 2 or et.shape[0] != total_e or total_e == 0 or num_rows == 0:
        raise NotImplementedError('RGCNConv: expected edge_index [2, E] and edge_type [E] of a non-empty batch')
    g = math.gcd(int(num_rows), total_e)
    for B in sorted((b for b in range(1, g + 1) if g % b == 0), reverse=True):
        n, E = num_rows // B, total_e // B
        if n > ops.MAX_SLOTS or E + n > ops.MAX_TERMS:
            break      # smaller B only makes the template larger
        src, dst, rel = ei[0, :E], ei[1, :E], et[:E]
        if src.max() >= n or dst.max() >= n or src.min() < 0 or dst.min() < 0:
            continue
        offs = (torch.arange(B).repeat_interleave(E) * n).numpy()
        if ((ei[0] == (offs + list(src) * B)).all() and (ei[1] == (offs + list(dst) * B)).all() and
                (et == list(rel) * B).all()):
            return _InferredBatch(_InferredTemplate(n, src.tolist(), dst.tolist()), rel.tolist(), B)
    raise NotImplementedError('RGCNConv: the edge list is not a batch of identical query templates with at most %d '
                              'nodes and %d edge + node terms (general graphs are outside the MPQE hot path)'
                              % (ops.MAX_SLOTS, ops.MAX_TERMS))


def _conv_terms(t, rels, x, n, w, root):
    terms = [Term(x, n, t.src[e], w[rels[e]], t.dst[e]) for e in range(t.num_edges)]
    terms += [Term(x, n, i, root, i) for i in range(n)]
    return terms


class _ConvFn(torch.autograd.Function):
    """One stand-alone layer pass over a template batch (the per-layer API of the reference's RGCNConv)."""

    @staticmethod
    def forward(ctx, x, w, root, bias, graph):
        t, rels, B = graph.template, graph.edge_rel_ids, graph.num_graphs
        n = t.num_nodes
        with ops.device_guard(x.device):
            x = x.contiguous()
            out = torch.empty(B * n, D, dtype=torch.float32, device=x.device)
            ops.layer_forward([Group(B, _conv_terms(t, rels, x, n, w, root), n, out, n, bias=bias)])
        ctx.save_for_backward(x, w, root)
        ctx.graph, ctx.has_bias = graph, bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, root = ctx.saved_tensors
        t, rels, B = ctx.graph.template, ctx.graph.edge_rel_ids, ctx.graph.num_graphs
        n = t.num_nodes
        with ops.device_guard(x.device):
            g = g.contiguous()
            wt, roott = ops.transpose(w), ops.transpose(root)
            dx = torch.empty_like(x)
            back = [Term(g, n, t.dst[e], wt[rels[e]], t.src[e]) for e in range(t.num_edges)]
            back += [Term(g, n, i, roott, i) for i in range(n)]
            ops.layer_forward([Group(B, back, n, dx, n)])
            dw, droot = torch.zeros_like(w), torch.zeros_like(root)
            fwd = Group(B, _conv_terms(t, rels, x, n, w, root), n, None, n)
            dests = [(w[r], dw[r], 1) for r in sorted(set(rels))] + [(root, droot, 1)]
            ops.layer_wgrad([fwd], [(g, n, list(range(n)))], dests)
            dbias = None
            if ctx.has_bias:
                dbias = torch.empty(D, dtype=torch.float32, device=x.device)
                ops.colsum(g, B * n, D, dbias)
        return dx, dw, droot, dbias, None


class MLPReadout(nn.Module):
    """Linear-ReLU-Linear on every node, then scatter over each query's nodes (reference model.py:497-515).
    Inside RGCNEncoderDecoder the arithmetic runs in the fused step; called on its own -- with the reference's
    signature, `readout(embs=, batch_idx=, batch_size=, num_nodes=, num_anchors=)` -- it runs the same CUDA kernels
    (term-list launches for both linear layers, their input / weight gradients and the scatter) behind autograd."""

    def __init__(self, input_dim, output_dim, scatter_fn):
        super(MLPReadout, self).__init__()
        self.layers = nn.Sequential(nn.Linear(in_features=input_dim, out_features=output_dim), nn.ReLU(),
                                    nn.Linear(in_features=output_dim, out_features=output_dim))
        self.scatter_fn = scatter_fn

    def forward(self, embs, batch_idx=None, batch_size=None, num_nodes=None, num_anchors=None, **kwargs):
        return _readout_forward(self, 'concat' if self.layers[0].in_features > D else 'mlp', embs, batch_size,
                                num_nodes, num_anchors)


class TargetMLPReadout(nn.Module):
    """MLP on cat(target node, other node) for every non-target node, then scatter (reference model.py:518-553)."""

    def __init__(self, dim, scatter_fn):
        super(TargetMLPReadout, self).__init__()
        self.layers = nn.Sequential(nn.Linear(in_features=2 * dim, out_features=dim), nn.ReLU(),
                                    nn.Linear(in_features=dim, out_features=dim))
        self.scatter_fn = scatter_fn

    def forward(self, embs, batch_idx=None, batch_size=None, num_nodes=None, num_anchors=None, **kwargs):
        return _readout_forward(self, 'targetmlp', embs, batch_size, num_nodes, num_anchors)


def _scatter_name(fn):
    name = fn if isinstance(fn, str) else getattr(fn, '__name__', str(fn))
    for op in ('add', 'max', 'mean'):
        if op in name:
            return op
    raise ValueError('Unknown scatter op %r' % (fn,))


def _readout_forward(module, kind, embs, batch_size, num_nodes, num_anchors):
    if batch_size is None or num_nodes is None:
        raise ValueError('the readout needs batch_size and num_nodes (queries of a batch share one template)')
    lin1, lin2 = module.layers[0], module.layers[2]
    if lin1.out_features != D or lin1.in_features % D != 0:
        raise ValueError('kernels are specialised for %d channels' % D)
    return _ReadoutFn.apply(embs, lin1.weight, lin1.bias, lin2.weight, lin2.bias, kind, _scatter_name(module.scatter_fn),
                            int(batch_size), int(num_nodes), int(num_anchors if num_anchors is not None else 0))


class _ReadoutFn(torch.autograd.Function):
    """A stand-alone MLP readout: embs [B*n, blocks*D] -> [B, D] (the per-module API of the reference's readouts)."""

    @staticmethod
    def _units(kind, n, target, blocks):
        """[(unit k, [(node slot j, W1 block b)])]: which node rows feed MLP unit k, through which block of W1."""
        if kind == 'targetmlp':
            return [[(target, 0), (j, 1)] for j in range(n) if j != target]
        return [[(j, b) for b in range(blocks)] for j in range(n)]

    @staticmethod
    def forward(ctx, embs, w1, b1, w2, b2, kind, op, B, n, a):
        with ops.device_guard(embs.device):
            blocks = w1.shape[1] // D
            nb = blocks if kind != 'targetmlp' else 1            # feature blocks per node row of `embs`
            x = embs.contiguous().view(B, n * nb, D)              # node j, block b -> slot j*nb + b
            w1t = ops.transpose(w1.contiguous())                  # [blocks*D, D]
            w2t = ops.transpose(w2.contiguous())
            units = _ReadoutFn._units(kind, n, a, blocks)
            nu = len(units)
            slot = (lambda j, b: j * nb + b) if kind != 'targetmlp' else (lambda j, b: j)
            u = torch.empty(B, nu, D, dtype=torch.float32, device=embs.device)
            t1 = [Term(x, n * nb, slot(j, b), w1t[b * D:(b + 1) * D], k) for k, srcs in enumerate(units) for (j, b) in srcs]
            g1 = Group(B, t1, nu, u, nu, epilogue=EPI_RELU, bias=b1)
            ops.layer_forward([g1])
            argmax = None
            if op == 'max':
                z2 = torch.empty(B, nu, D, dtype=torch.float32, device=embs.device)
                g2 = Group(B, [Term(u, nu, k, w2t, k) for k in range(nu)], nu, z2, nu, bias=b2)
                ops.layer_forward([g2])
                q, argmax = ops.max_readout(z2, B, nu)
            else:
                q = torch.empty(B, D, dtype=torch.float32, device=embs.device)
                g2 = Group(B, [Term(u, nu, k, w2t, 0) for k in range(nu)], 1, q, 1, out_slot_map=[0], bias=b2,
                           bias_scale=[float(nu)])
                ops.layer_forward([g2])
                if op == 'mean':
                    q = q * (1.0 / nu)
        ctx.save_for_backward(x, u, w1, w2, w1t, w2t)
        ctx.meta = (kind, op, B, n, a, nb, blocks, units, argmax, g1, g2, embs.shape)
        return q

    @staticmethod
    def backward(ctx, dq):
        x, u, w1, w2, w1t, w2t = ctx.saved_tensors
        kind, op, B, n, a, nb, blocks, units, argmax, g1, g2, shape = ctx.meta
        nu = len(units)
        dev = dq.device
        with ops.device_guard(dev):
            dq = dq.contiguous()
            if op == 'max':
                gt, gs, smap = ops.max_readout_bwd(dq, argmax, B, nu), nu, list(range(nu))
            else:
                gt, gs, smap = (dq * (1.0 / nu) if op == 'mean' else dq), 1, [0] * nu
            du = torch.empty_like(u)       # dU[:, k] = (dZ2[:, k] @ W2) * (u > 0)
            ops.layer_forward([Group(B, [Term(gt, gs, smap[k], w2.contiguous(), k) for k in range(nu)], nu, du, nu,
                                     epilogue=EPI_MASK, mask=u, mask_slots=nu)])
            dw1t, dw2t = torch.zeros_like(w1t), torch.zeros_like(w2t)
            ops.layer_wgrad([g2], [(gt, gs, smap)], [(w2t, dw2t, 1)])
            ops.layer_wgrad([g1], [(du, nu, list(range(nu)))],
                            [(w1t[b * D:(b + 1) * D], dw1t[b * D:(b + 1) * D], 1) for b in range(blocks)])
            db1 = torch.empty(D, dtype=torch.float32, device=dev)
            db2 = torch.empty(D, dtype=torch.float32, device=dev)
            ops.colsum(du, B * nu, D, db1)
            ops.colsum(gt, B * gs, D, db2, scale=float(nu) if gs == 1 else 1.0)
            # d embs: node slot (j, b) collects dU of the units it feeds, through block b of W1
            w1b = ops.transpose(w1t.view(blocks, D, D))
            dx = torch.zeros_like(x)
            slot = (lambda j, b: j * nb + b) if kind != 'targetmlp' else (lambda j, b: j)
            feeds = {}
            for k, srcs in enumerate(units):
                for (j, b) in srcs:
                    feeds.setdefault(slot(j, b), []).append((k, b))
            slots = sorted(feeds)
            for lo in range(0, len(slots), ops.MAX_SLOTS):
                part = slots[lo:lo + ops.MAX_SLOTS]
                terms = [Term(du, nu, k, w1b[b], i) for i, sl in enumerate(part) for (k, b) in feeds[sl]]
                for t0 in range(0, len(terms), ops.MAX_TERMS):     # accumulate over term chunks via separate passes
                    chunk = terms[t0:t0 + ops.MAX_TERMS]
                    if t0 == 0:
                        ops.layer_forward([Group(B, chunk, len(part), dx, x.shape[1], out_slot_map=part)])
                    else:
                        extra = torch.zeros_like(dx)
                        ops.layer_forward([Group(B, chunk, len(part), extra, x.shape[1], out_slot_map=part)])
                        dx += extra
        return (dx.view(shape), ops.transpose(dw1t), db1, ops.transpose(dw2t), db2, None, None, None, None, None)


# ===============================================================================================================
# Engine: one formula group = one Job; passes of all jobs are launched together (<= MPQE_MAX_GROUPS per launch)
# ===============================================================================================================
class Job(object):
    """One batch of queries of a single formula, ids already on the device."""

    def __init__(self, template, edge_rels, var_rows, anchor_modes, target_mode, anchor_ids, num_passes):
        self.t, self.rels, self.var_rows = template, tuple(edge_rels), var_rows
        self.anchor_modes, self.target_mode = tuple(anchor_modes), target_mode
        self.anchor_ids = anchor_ids                      # int64 [B, a] on device
        self.B = int(anchor_ids.shape[0])
        self.P = int(num_passes)
        # filled by the engine
        self.acts = None       # acts[p] = input of pass p, [B, n, D]
        self.outs = None       # outs[p] = node slots computed by pass p
        self.z = self.u = self.q = self.argmax = None
        self.fwd_groups = None


class Weights(object):
    """Per-step views of the dense parameters in the layouts the kernels want."""

    def __init__(self, model, need_grad, defer_raw=False):
        """defer_raw: with the tensor-core kernels the transposed matrices are only READ late in the backward (by the
        one-row kernels of the batch-constant tail) -- their tile images are packed straight from the untransposed
        parameters -- so a caller may run `raw_transposes()` itself, off the chain the first layer launch waits for."""
        self.w, self.root, self.bias = [], [], []
        self.raw_pending, self.raw_event = False, None
        for layer in model.distinct_layers():
            self.w.append(layer.relation_weights().contiguous())
            self.root.append(layer.root)
            self.bias.append(layer.bias)
        self.mode_emb = model.mode_embeddings.weight
        self.need_grad = need_grad
        self.wt = self.roott = None
        self.w1t = self.w1b = self.w2t = None
        ro = model.readout if isinstance(model.readout, nn.Module) else None
        self.ro = ro
        if ro is not None:
            lin1, lin2 = ro.layers[0], ro.layers[2]
            self.w1, self.b1, self.w2, self.b2 = lin1.weight, lin1.bias, lin2.weight, lin2.bias
            self.blocks = self.w1.shape[1] // D
            self.w1t = ops.transpose(self.w1.contiguous())                     # [blocks*D, D]: block p = W1[:, pD:(p+1)D]^T
            self.w2t = ops.transpose(self.w2.contiguous())
        # second-generation tensor-core kernel: the images of the transposed matrices come from the pack kernel
        self.pack_transposed = need_grad and ops.tensor_cores_default() and ops.layer_generation() == 2
        if need_grad:
            if self.pack_transposed:
                self.wt = [torch.empty_like(w) for w in self.w]
                self.roott = [torch.empty_like(r) for r in self.root]
                self.raw_pending = True
            else:
                self.wt = [ops.transpose(w) for w in self.w]
                self.roott = [ops.transpose(r) for r in self.root]
            if ro is not None:
                self.w1b = ops.transpose(self.w1t.view(self.blocks, D, D))     # [blocks, D(u), D(h)] contiguous
        # tcgen05 path: tf32 hi/lo tile images of every matrix, staged by the kernel with one bulk copy per tile;
        # all matrices of the step are packed by ONE launch
        self.wp = self.rootp = self.wtp = self.roottp = None
        self.w1tp = self.w2tp = self.w2p = self.w1bp = None
        if ops.tensor_cores_default():
            mats, flags = [], []
            for li in range(len(self.w)):
                fwd = [self.w[li][r] for r in range(self.w[li].shape[0])] + [self.root[li]]
                mats += fwd
                flags += [0] * len(fwd)
                if need_grad and self.pack_transposed:
                    mats += fwd
                    flags += [1] * len(fwd)
                elif need_grad:
                    mats += [self.wt[li][r] for r in range(self.wt[li].shape[0])] + [self.roott[li]]
                    flags += [0] * len(fwd)
            packed = ops.pack_weights(mats, flags)
            self.wp, self.rootp, self.wtp, self.roottp = [], [], [], []
            off = 0
            for li in range(len(self.w)):
                R = self.w[li].shape[0]
                self.wp.append(packed[off:off + R])
                self.rootp.append(packed[off + R])
                off += R + 1
                if need_grad:
                    self.wtp.append(packed[off:off + R])
                    self.roottp.append(packed[off + R])
                    off += R + 1
            if not need_grad:
                self.wtp = self.roottp = None
            if ro is not None:
                # readout MLP: W1^T blocks and W2^T (forward), W2 and the W1 blocks (their input gradients)
                mats = [self.w1t[b * D:(b + 1) * D] for b in range(self.blocks)] + [self.w2t]
                if need_grad:
                    mats += [self.w2] + [self.w1b[b] for b in range(self.blocks)]
                packed = ops.pack_weights([m.contiguous() for m in mats])
                self.w1tp, self.w2tp = packed[:self.blocks], packed[self.blocks]
                if need_grad:
                    self.w2p, self.w1bp = packed[self.blocks + 1], packed[self.blocks + 2:]
        self.msum = self.msumt = None
        if not defer_raw:
            self.raw_transposes()

    def raw_transposes(self):
        """Fills the transposed copies whose tile images were packed from the untransposed matrices."""
        if not self.raw_pending:
            return
        for w, wt in zip(self.w, self.wt):
            ops.transpose(w, wt)
        for r, rt in zip(self.root, self.roott):
            ops.transpose(r, rt)
        if self.msum is not None and self.msumt is not None:
            ops.transpose(self.msum, self.msumt)
        self.raw_pending = False


def _needed_slots(job, readout):
    """outs[p]: node slots whose output pass p must produce (all, except for the target-message readout)."""
    t, P = job.t, job.P
    if readout != 'mp':
        return [list(range(t.num_nodes)) for _ in range(P)]
    need = {t.target_slot}
    outs = [None] * P
    for p in range(P - 1, -1, -1):
        outs[p] = sorted(need)
        need = need | {t.src[e] for e in range(t.num_edges) if t.dst[e] in need}
    return outs


class Engine(object):
    def __init__(self, model):
        self.m = model

    # ---- forward ------------------------------------------------------------------------------------------
    def build_inputs(self, jobs, W):
        """x[b, i<a] = normalised anchor rows (reference model.py:418-420), one multi-item launch for every (job, anchor
        slot).  The variable slots x[b, i>=a] = mode_embeddings[var_ids] (model.py:421) are the same row for every query
        of the batch, so they are never materialised: their contribution to the first pass is a per-slot constant
        (see `pass_terms`)."""
        enc = self.m.enc
        items = []
        for job in jobs:
            t, B, n = job.t, job.B, job.t.num_nodes
            x = torch.empty(B, n, D, dtype=torch.float32, device=job.anchor_ids.device)
            # ||row|| of every gathered anchor row: with it (and x itself) the backward needs no second gather
            job.anchor_norm = torch.empty(t.num_anchors, B, dtype=torch.float32, device=job.anchor_ids.device) \
                if W.need_grad else None
            for i, mode in enumerate(job.anchor_modes):
                items.append(ops.GatherItem(enc.table(mode), enc.node_maps, job.anchor_ids, B, ids_offset=i,
                                            ids_stride=t.num_anchors, out=x, out_offset=i * D, out_stride=n * D,
                                            norm=job.anchor_norm, norm_offset=i * B))
            job.acts = [x]
            job.act_bits = {}
        ops.gather_multi(items)

    def layer_index(self, p, P):
        if self.m.shared_layers:
            return 0
        return p if p < P - 1 else self.m.num_layers - 1

    def pass_terms(self, job, p, W, outs, const_only=False):
        """Forward terms of pass p restricted to the output slots `outs` (term.out_slot = index into outs), split into
        (per-query terms, batch-constant terms, layer index).  In pass 0 every term whose source is a variable slot
        multiplies the same variable-type embedding row for all queries: it is evaluated once per group (a 1-row
        launch) and enters the big launch as a per-slot bias; its weight gradient is a rank-1 update and its input
        gradient a column sum (see `backward`).  `const_only`: just the batch-constant terms (no activations needed)."""
        t, n, a = job.t, job.t.num_nodes, job.t.num_anchors
        li = self.layer_index(p, job.P)
        x = job.acts[p] if not const_only else None
        if p > 0 and p == job.P - 1 and job.collapsed is not None:
            # fused readout: one term per source slot with the summed matrix (see collapse_last_pass)
            return [Term(x, n, c[0], self._cmat(W, c, 'm'), 0, self._cmat(W, c, 'mp')) for c in job.collapsed], [], li
        pos = {s: k for k, s in enumerate(outs)}
        wp = W.wp[li] if W.wp is not None else None
        rp = W.rootp[li] if W.rootp is not None else None
        main, const = [], []

        def add(src, m, mp, out):
            if p == 0 and src >= a:
                const.append(Term(W.mode_emb, 0, job.var_rows_host[src - a], m, out))
            elif not const_only:
                main.append(Term(x, n, src, m, out, mp))

        for e in range(t.num_edges):
            if t.dst[e] in pos:
                add(t.src[e], W.w[li][job.rels[e]], wp[job.rels[e]] if wp is not None else None, pos[t.dst[e]])
        for s in outs:
            add(s, W.root[li], rp, pos[s])
        return main, const, li

    # ---- last pass under a sum / target-message readout: terms collapse per source slot --------------------
    @staticmethod
    def _pmat(W, li, key, kind):
        """Parameter matrix `key` (('root',) or ('w', r)) of layer li as: 'm' forward, 'mp' its packed image,
        't' transposed, 'tp' packed transposed."""
        if key[0] == 'root':
            src = {'m': W.root, 'mp': W.rootp, 't': W.roott, 'tp': W.roottp}[kind]
            return src[li] if src is not None else None
        src = {'m': W.w, 'mp': W.wp, 't': W.wt, 'tp': W.wtp}[kind]
        return src[li][key[1]] if src is not None else None

    def _cmat(self, W, c, kind):
        """Matrix of a collapsed term c = (slot, li, keys, k): the summed matrix k, or the single parameter matrix."""
        s, li, keys, k = c
        if k is None:
            return self._pmat(W, li, keys[0], kind)
        src = {'m': W.msum, 'mp': W.msump, 't': W.msumt, 'tp': W.msumtp}[kind]
        return src[k] if src is not None else None

    def collapse_last_pass(self, jobs, W):
        """With a sum (or target-message) readout the last pass writes ONE row per query: all terms that leave a source
        slot read the same input row and add into the same output row, so they collapse into one term with the matrix
        root + sum of basis[rel] over the slot's edges (the layer is linear in its weights, model.py:292-304) --
        41 -> 24 term tiles per query on the bench workload, in the forward, the input gradient and the weight
        gradient.  job.collapsed = [(slot, layer, keys of the summed parameter matrices, index k into the step's
        summed-matrix buffers or None for a single matrix)]."""
        for job in jobs:
            job.collapsed = None
        W.msum = W.msump = W.msumt = W.msumtp = W.dmsum = None
        W.msum_keys = []
        if self.m.readout_str not in ('sum', 'mp'):
            return
        for job in jobs:
            if job.P < 2:
                continue
            t, li, outs = job.t, self.layer_index(job.P - 1, job.P), job.outs[job.P - 1]
            by_src = {}
            for s in outs:
                by_src.setdefault(s, []).append(('root',))
            for e in range(t.num_edges):
                if t.dst[e] in outs:
                    by_src.setdefault(t.src[e], []).append(('w', job.rels[e]))
            job.collapsed = []
            for s in sorted(by_src):
                keys, k = by_src[s], None
                if len(keys) > 1:
                    k = len(W.msum_keys)
                    W.msum_keys.append((li, keys))
                job.collapsed.append((s, li, keys, k))
        if not W.msum_keys:
            return
        dev = jobs[0].anchor_ids.device
        W.msum = torch.empty(len(W.msum_keys), D, D, dtype=torch.float32, device=dev)
        ops.matrix_sum_multi([(W.msum[k], [self._pmat(W, li, key, 'm') for key in keys], False)
                              for k, (li, keys) in enumerate(W.msum_keys)])
        tc = ops.tensor_cores_default()
        K = len(W.msum_keys)
        deferred = W.need_grad and getattr(W, 'raw_pending', False)
        if W.need_grad:
            W.msumt = torch.empty_like(W.msum) if deferred else ops.transpose(W.msum)   # (filled by raw_transposes)
            W.dmsum = torch.empty_like(W.msum)
        if tc:
            fwd = [W.msum[k] for k in range(K)]
            if deferred:
                packed = ops.pack_weights(fwd + fwd, [0] * K + [1] * K)
            else:
                packed = ops.pack_weights(fwd + ([W.msumt[k] for k in range(K)] if W.need_grad else []))
            W.msump = packed[:K]
            W.msumtp = packed[K:] if W.need_grad else None

    @staticmethod
    def _pgrad(G, li, key):
        return G.droot[li] if key[0] == 'root' else G.dw[li][key[1]]

    def wgrad_dests(self, job, p, li, W, G, spread):
        """{matrix address: (forward matrix, gradient destination, accumulate)} of job's pass p.  Summed matrices of a
        collapsed last pass get their own (overwritten) gradient; `spread` collects where those go afterwards."""
        dests = {}
        if p > 0 and p == job.P - 1 and job.collapsed is not None:
            for c in job.collapsed:
                s, cli, keys, k = c
                if k is None:
                    m = self._pmat(W, cli, keys[0], 'm')
                    dests[m.data_ptr()] = (m, self._pgrad(G, cli, keys[0]), 1)
                else:
                    if W.dmsum is None:
                        W.dmsum = torch.empty_like(W.msum)
                    dests[W.msum[k].data_ptr()] = (W.msum[k], W.dmsum[k], 0)
                    for key in keys:
                        spread.setdefault((cli, key), []).append(W.dmsum[k])
            return dests
        used = {term.m.data_ptr() for term in job.fwd_groups[p].terms}   # (pass 0: without the batch-constant terms)
        for r in sorted(set(job.rels)):
            if W.w[li][r].data_ptr() in used:
                dests[W.w[li][r].data_ptr()] = (W.w[li][r], G.dw[li][r], 1)
        if W.root[li].data_ptr() in used:
            dests[W.root[li].data_ptr()] = (W.root[li], G.droot[li], 1)
        return dests

    def prepare(self, jobs, W):
        """Everything of a step that depends only on the weights and the formulas (needed slots, summed matrices of
        a collapsed last pass).  May be called ahead of `encode`, e.g. on a second stream (see TrainStep)."""
        readout = self.m.readout_str
        for job in jobs:
            job.outs = _needed_slots(job, readout)
            job.fwd_groups = [None] * job.P
        self.collapse_last_pass(jobs, W)
        # batch-constant part of pass 0 (+ bias), evaluated on one row per group: job.const_fwd.out enters the big
        # launch as a per-slot bias
        const_groups = []
        for job in jobs:
            job.const_fwd = None
            outs = job.outs[0]
            _, const, li = self.pass_terms(job, 0, W, outs, const_only=True)
            if not const:
                continue
            fused = job.P == 1 and readout in ('sum', 'mp')
            nb = 1 if fused else len(outs)
            if fused:
                for term in const:
                    term.out_slot = 0
            cb = torch.empty(1, nb, D, dtype=torch.float32, device=job.anchor_ids.device)
            job.const_fwd = Group(1, const, nb, cb, nb, bias=W.bias[li],
                                  bias_scale=[float(len(outs))] if fused else [1.0] * nb)
            const_groups.append(job.const_fwd)
        if const_groups:
            ops.layer_forward(const_groups, use_tensor_cores=False)   # 1-row groups: the one-row kernel
        W.prepared = jobs

    def encode(self, jobs, W):
        """Runs every pass of every job; leaves job.q [B, D] (query embeddings)."""
        readout = self.m.readout_str
        self.build_inputs(jobs, W)
        prepared = getattr(W, 'prepared', None)
        if prepared is not jobs and (prepared is None or not all(any(job is pj for pj in prepared) for job in jobs)):
            self.prepare(jobs, W)      # (a subset of the prepared jobs -- one lane of a step -- is prepared)
        if getattr(W, 'ready_event', None) is not None:    # weights prepared on another stream
            torch.cuda.current_stream().wait_event(W.ready_event)
        max_p = max(job.P for job in jobs)
        for p in range(max_p):
            groups = []
            for job in jobs:
                if p >= job.P:
                    continue
                B, n = job.B, job.t.num_nodes
                outs = job.outs[p]
                terms, _, li = self.pass_terms(job, p, W, outs)
                dev = job.anchor_ids.device
                fused = p == job.P - 1 and readout in ('sum', 'mp')
                nb = 1 if fused else len(outs)
                if fused:      # readout folded into the last pass: every term accumulates into the single output row
                    for term in terms:
                        term.out_slot = 0
                bias, bias_scale, stride = W.bias[li], ([float(len(outs))] if fused else [1.0] * nb), 0
                job.fused_bias_count = float(len(outs)) if fused else 1.0
                if p == 0 and job.const_fwd is not None:   # constant part + bias, computed by `prepare`
                    bias, bias_scale, stride = job.const_fwd.out, [1.0] * nb, D
                if p < job.P - 1:
                    h = torch.empty(B, n, D, dtype=torch.float32, device=dev)
                    # the tcgen05 kernel also leaves the ReLU sign bits (1/32 of h) for the backward's mask
                    bits = ops.relu_bits(B, n, dev) if ops.tensor_cores_default() and W.need_grad else None
                    g = Group(B, terms, nb, h, n, out_slot_map=outs, epilogue=EPI_RELU, bias=bias, bias_scale=bias_scale,
                              bias_slot_stride=stride, bits_out=bits)
                    job.acts.append(h)
                    job.act_bits[p + 1] = bits
                elif fused:
                    job.q = torch.empty(B, D, dtype=torch.float32, device=dev)
                    g = Group(B, terms, 1, job.q, 1, out_slot_map=[0], bias=bias, bias_scale=bias_scale)
                else:
                    job.z = torch.empty(B, n, D, dtype=torch.float32, device=dev)
                    g = Group(B, terms, nb, job.z, n, out_slot_map=outs, bias=bias, bias_scale=bias_scale,
                              bias_slot_stride=stride)
                job.fwd_groups[p] = g
                groups.append(g)
            ops.layer_forward(groups)
        if readout == 'max':
            for job in jobs:
                job.q, job.argmax = ops.max_readout(job.z, job.B, job.t.num_nodes)
        elif readout in MLP_READOUTS:
            self.mlp_readout(jobs, W)

    def mlp_inputs(self, job):
        """[(u slot k, [(tensor, slots, slot, block)])]: which rows feed MLP unit k, and through which W1 block."""
        t, n = job.t, job.t.num_nodes
        ro = self.m.readout_str
        if ro == 'mlp':
            return [[(job.z, n, j, 0)] for j in range(n)]
        if ro == 'concat':
            feats = job.acts[1:] + [job.z]
            return [[(f, n, j, b) for b, f in enumerate(feats)] for j in range(n)]
        others = [j for j in range(n) if j != t.target_slot]
        return [[(job.z, n, t.target_slot, 0), (job.z, n, j, 1)] for j in others]

    def mlp_readout(self, jobs, W):
        g1, g2 = [], []
        for job in jobs:
            units = self.mlp_inputs(job)
            nu = len(units)
            dev = job.anchor_ids.device
            terms = [Term(a, s, j, W.w1t[b * D:(b + 1) * D], k, W.w1tp[b] if W.w1tp is not None else None)
                     for k, srcs in enumerate(units) for (a, s, j, b) in srcs]
            job.u = torch.empty(job.B, nu, D, dtype=torch.float32, device=dev)
            job.u_bits = ops.relu_bits(job.B, nu, dev) if ops.tensor_cores_default() and W.need_grad else None
            job.mlp1 = Group(job.B, terms, nu, job.u, nu, epilogue=EPI_RELU, bias=W.b1, bits_out=job.u_bits)
            if self.m.scatter_op == 'max':     # per-unit outputs, then the per-feature max over the query's units
                job.z2 = torch.empty(job.B, nu, D, dtype=torch.float32, device=dev)
                job.mlp2 = Group(job.B, [Term(job.u, nu, k, W.w2t, k, W.w2tp) for k in range(nu)], nu, job.z2, nu, bias=W.b2)
            else:                              # add / mean: the second linear layer sums over the units directly
                job.q = torch.empty(job.B, D, dtype=torch.float32, device=dev)
                job.mlp2 = Group(job.B, [Term(job.u, nu, k, W.w2t, 0, W.w2tp) for k in range(nu)], 1, job.q, 1,
                                 out_slot_map=[0], bias=W.b2, bias_scale=[float(nu)])
            g1.append(job.mlp1)
            g2.append(job.mlp2)
        ops.layer_forward(g1)
        ops.layer_forward(g2)
        for job in jobs:
            nu = job.u.shape[1]
            if self.m.scatter_op == 'max':
                job.q, job.argmax2 = ops.max_readout(job.z2, job.B, nu)
            elif self.m.scatter_op == 'mean':  # scatter_mean = scatter_add / units per query (exact for 2 and 4 units)
                job.q.mul_(1.0 / nu)

    # ---- backward -----------------------------------------------------------------------------------------
    def backward(self, jobs, W, dqs, G, defer_constant=False, side=None):
        """dqs[i] = d loss / d job.q.  Accumulates dense gradients into `G` (a Grads) and row gradients into G.rows.
        `side` = (run, join): run(fn) executes fn() on another stream forked from the current one, join() makes the
        current stream wait for it.  The weight-gradient launches of a pass then run next to its input-gradient
        launch, which does not depend on them: one kernel's write-bound tail overlaps the other's main loop
        (0.439 -> 0.400 ms per bench step).  Joined here, before the batch-constant tail."""
        readout = self.m.readout_str
        # gradient wrt the last pass output, as (tensor, slots, slot_map over outs[P-1])
        last = {}
        if readout in ('sum', 'mp'):
            for job, dq in zip(jobs, dqs):
                last[job] = (dq, 1, [0] * len(job.outs[job.P - 1]))
        elif readout == 'max':
            for job, dq in zip(jobs, dqs):
                n = job.t.num_nodes
                last[job] = (ops.max_readout_bwd(dq, job.argmax, job.B, n), n, list(range(n)))
        else:
            self.mlp_backward(jobs, W, dqs, G, last)

        cur = dict(last)  # job -> gradient operand of the pass being processed
        max_p = max(job.P for job in jobs)
        for p in range(max_p - 1, -1, -1):
            active = [job for job in jobs if p < job.P]
            # ---- weight / bias gradients of pass p (grouped by layer so destinations are unique per launch)
            by_layer = {}
            for job in active:
                by_layer.setdefault(self.layer_index(p, job.P), []).append(job)
            spread = {}    # (layer, parameter key) -> gradients of the summed matrices it is a summand of
            launches = []  # (forward groups, gradient operands, destinations) per weight-gradient launch
            for li, ljobs in by_layer.items():
                chunk, dests = [], {}
                for job in ljobs + [None]:
                    jd = self.wgrad_dests(job, p, li, W, G, spread) if job is not None else {}
                    if chunk and (job is None or len(chunk) == ops.MAX_GROUPS or
                                  len(set(dests) | set(jd)) > ops.MAX_DESTS):
                        if dests:
                            launches.append(([j.fwd_groups[p] for j in chunk], [cur[j] for j in chunk],
                                             list(dests.values())))
                        chunk, dests = [], {}
                    if job is not None:
                        chunk.append(job)
                        dests.update(jd)
            G.keep.append(launches)   # the operands stay alive until the caller has joined the second stream

            def weight_gradients(launches=launches, spread=spread):
                for fwd, operands, dests in launches:
                    ops.layer_wgrad(fwd, operands, dests)
                if spread:     # d(root + sum basis[rel]) goes to every summand, in a fixed order
                    ops.matrix_sum_multi([(self._pgrad(G, li, key), srcs, True) for (li, key), srcs in spread.items()])
            if side is not None:
                side[0](weight_gradients)
            else:
                weight_gradients()
            for li, ljobs in by_layer.items():
                for job in ljobs:
                    if p == 0 and job.const_fwd is not None:
                        continue   # bias gradient = sum of the per-slot sums taken below (see constant_backward)
                    g, g_slots, smap = cur[job]
                    if g_slots == 1:
                        # sum readout: the bias was added once per node; target-message: once
                        G.colsum(g, job.B, D, G.dbias[li], job.fused_bias_count)
                    elif len(smap) == g_slots:
                        G.colsum(g, job.B * g_slots, D, G.dbias[li])
                    else:
                        for s in smap:
                            G.colsum(g[:, s], job.B, g_slots * D, G.dbias[li])
            # ---- batch-constant terms of pass 0: per-slot column sums of the output gradient, consumed after flush
            if p == 0:
                cjobs = [job for job in active if job.const_fwd is not None]
                if cjobs:      # one zero-filled buffer for every job's sums (one memset, not one per job)
                    widths = [1 if cur[job][1] == 1 else job.t.num_nodes for job in cjobs]
                    cs_all = torch.zeros(sum(widths), D, dtype=torch.float32, device=cur[cjobs[0]][0].device)
                    off = 0
                for job, width in zip(cjobs, widths if cjobs else []):
                    g, g_slots, smap = cur[job]
                    job.cs = cs_all[off:off + width].view(1, width, D)
                    off += width
                    if g_slots == 1:
                        G.colsum(g, job.B, D, job.cs[0, 0])
                        job.cs_operand = (job.cs, 1, [0] * len(smap))
                    else:
                        for s_ in smap:
                            G.colsum(g[:, s_], job.B, g_slots * D, job.cs[0, s_])
                        job.cs_operand = (job.cs, width, list(smap))
            # ---- input gradients of pass p
            groups, nxt = [], {}
            for job in active:
                t, n = job.t, job.t.num_nodes
                g, g_slots, smap = cur[job]
                li = self.layer_index(p, job.P)
                outs = job.outs[p]
                if p > 0:
                    ins = job.outs[p - 1]
                else:   # only the anchors receive a per-query gradient; variable slots get a column sum (below)
                    ins = self.grad_anchor_slots(job)
                if not ins:
                    continue
                pos = {s: k for k, s in enumerate(ins)}
                okey = {s: k for k, s in enumerate(outs)}
                wtp = W.wtp[li] if W.wtp is not None else None
                rtp = W.roottp[li] if W.roottp is not None else None
                if p > 0 and p == job.P - 1 and job.collapsed is not None:
                    # collapsed last pass: d h[:, s] = dq @ (root + sum basis[rel])^T, one term per source slot
                    terms = [Term(g, g_slots, 0, self._cmat(W, c, 't'), pos[c[0]], self._cmat(W, c, 'tp'))
                             for c in job.collapsed if c[0] in pos]
                else:
                    terms = [Term(g, g_slots, smap[okey[t.dst[e]]], W.wt[li][job.rels[e]], pos[t.src[e]],
                                  wtp[job.rels[e]] if wtp is not None else None)
                             for e in range(t.num_edges) if t.dst[e] in okey and t.src[e] in pos]
                    terms += [Term(g, g_slots, smap[okey[s]], W.roott[li], pos[s], rtp) for s in outs if s in pos]
                if readout == 'concat' and p > 0:
                    # h_p also feeds block p-1 of the concat MLP
                    terms += [Term(job.du, n, j, W.w1b[p - 1], pos[j], W.w1bp[p - 1] if W.w1bp is not None else None)
                              for j in range(n)]
                dx = torch.empty(job.B, n, D, dtype=torch.float32, device=g.device)
                if p > 0:
                    groups.append(Group(job.B, terms, len(ins), dx, n, out_slot_map=ins, epilogue=EPI_MASK,
                                        mask=job.acts[p], mask_slots=n, mask_bits=job.act_bits.get(p)))
                else:
                    groups.append(Group(job.B, terms, len(ins), dx, n, out_slot_map=ins))
                nxt[job] = (dx, n, ins)
            if groups:
                ops.layer_forward(groups)
            for job in active:
                if job not in nxt:
                    continue
                dx, n, ins = nxt[job]
                if p > 0:
                    cur[job] = (dx, n, ins)
                else:
                    self.input_backward(job, dx, ins, G)
        G.flush()
        if side is not None:    # the batch-constant tail accumulates into the same matrices as the weight gradients
            side[1]()
        if defer_constant:      # the caller runs it (e.g. on another stream, under independent work)
            G.finish = lambda: self.constant_backward(jobs, W, G)
        else:
            self.constant_backward(jobs, W, G)

    def constant_backward(self, jobs, W, G):
        """Gradients of the batch-constant pass-0 terms from the per-slot column sums cs[s] = sum_q dZ0[q, s]:
        dW += v^T cs[dst] (a 1-row weight-gradient launch) and d mode_embedding[v] += cs[dst] @ W^T (a 1-row layer
        launch followed by 1-row column sums)."""
        cjobs = [job for job in jobs if getattr(job, 'const_fwd', None) is not None]
        if not cjobs:
            return
        if getattr(W, 'raw_event', None) is not None:     # transposed copies filled on another stream (TrainStep)
            torch.cuda.current_stream().wait_event(W.raw_event)
        by_layer = {}
        for job in cjobs:
            by_layer.setdefault(self.layer_index(0, job.P), []).append(job)
        dgroups = []
        for li, ljobs in by_layer.items():
            for i in range(0, len(ljobs), ops.MAX_GROUPS):
                chunk = ljobs[i:i + ops.MAX_GROUPS]
                mats = {}
                for job in chunk:
                    for term in job.const_fwd.terms:
                        mats[term.m.data_ptr()] = term.m
                dests = []
                for r in range(W.w[li].shape[0]):
                    if W.w[li][r].data_ptr() in mats:
                        dests.append((W.w[li][r], G.dw[li][r], 1))
                if W.root[li].data_ptr() in mats:
                    dests.append((W.root[li], G.droot[li], 1))
                ops.layer_wgrad([job.const_fwd for job in chunk], [job.cs_operand for job in chunk], dests,
                                use_tensor_cores=False)
            for job in ljobs:
                t, n, a = job.t, job.t.num_nodes, job.t.num_anchors
                cs, cs_slots, smap = job.cs_operand
                # d bias of pass 0 from the same sums: one row per slot (zero rows for slots outside `smap`)
                G.colsum(cs.view(-1, D), cs_slots, D, G.dbias[li], job.fused_bias_count if cs_slots == 1 else 1.0)
                outs = job.outs[0]
                okey = {s: k for k, s in enumerate(outs)}
                var_slots = sorted({t.src[e] for e in range(t.num_edges) if t.dst[e] in okey and t.src[e] >= a} |
                                   {s for s in outs if s >= a})
                pos = {s: k for k, s in enumerate(var_slots)}
                terms = [Term(cs, cs_slots, smap[okey[t.dst[e]]], W.wt[li][job.rels[e]], pos[t.src[e]])
                         for e in range(t.num_edges) if t.dst[e] in okey and t.src[e] in pos]
                terms += [Term(cs, cs_slots, smap[okey[s]], W.roott[li], pos[s]) for s in outs if s in pos]
                job.dconst = torch.empty(1, n, D, dtype=torch.float32, device=cs.device)
                dgroups.append(Group(1, terms, len(var_slots), job.dconst, n, out_slot_map=var_slots))
                for s in var_slots:
                    G.colsum(job.dconst[:, s], 1, n * D, G.dmode[job.var_rows_host[s - a]])
        ops.layer_forward(dgroups, use_tensor_cores=False)
        G.flush()

    def grad_anchor_slots(self, job):
        """Anchor slots that receive a gradient: the sources (and self-loops) of the first pass's output slots."""
        t = job.t
        outs = _needed_slots(job, self.m.readout_str)[0]
        return sorted(s for s in ({t.src[e] for e in range(t.num_edges) if t.dst[e] in outs} | set(outs))
                      if s < t.num_anchors)

    def input_backward(self, job, dx, ins, G):
        """d loss / d x (anchor slots) -> entity-table rows through the normalisation."""
        t, n, B = job.t, job.t.num_nodes, job.B
        enc = self.m.enc
        planned = getattr(job, 'anchor_res', None) if G.rows.planned else None
        for i, mode in enumerate(job.anchor_modes):
            if i not in ins:
                continue
            rows, rows_id, off, id_off = planned[i] if planned is not None else G.rows.reserve(mode, B)
            norm = getattr(job, 'anchor_norm', None)
            G.gathers.append(ops.GatherItem(enc.table(mode), enc.node_maps, job.anchor_ids, B, ids_offset=i,
                                            ids_stride=t.num_anchors, grad=dx, grad_offset=i * D, grad_stride=n * D,
                                            rows_out=rows, rows_id=rows_id, rows_offset=off, id_offset=id_off,
                                            out=job.acts[0] if norm is not None else None, out_offset=i * D,
                                            out_stride=n * D, norm=norm, norm_offset=i * B))

    def mlp_backward(self, jobs, W, dqs, G, last):
        """Backward of the MLP readouts; fills last[job] (gradient wrt the last R-GCN pass output) and job.du."""
        ro, op = self.m.readout_str, self.m.scatter_op
        g_du, g2 = [], []      # g2[i] = gradient operand of the second linear layer: (tensor, slots, slot map)
        for job, dq in zip(jobs, dqs):
            nu = job.u.shape[1]
            job.du = torch.empty_like(job.u)
            if op == 'max':    # the gradient reaches only the unit that attained the maximum, feature by feature
                dz2 = ops.max_readout_bwd(dq, job.argmax2, job.B, nu)
                g2.append((dz2, nu, list(range(nu))))
            else:
                if op == 'mean':
                    dq = dq * (1.0 / nu)
                g2.append((dq, 1, [0] * nu))
            g, gs, smap = g2[-1]
            # dU[:, k] = (dZ2[:, k] @ W2) * (u > 0)       (W2 is stored [out, in] = the matrix this product needs)
            g_du.append(Group(job.B, [Term(g, gs, smap[k], W.w2, k, W.w2p) for k in range(nu)], nu, job.du, nu,
                              epilogue=EPI_MASK, mask=job.u, mask_slots=nu, mask_bits=getattr(job, 'u_bits', None)))
        ops.layer_forward(g_du)
        for i in range(0, len(jobs), ops.MAX_GROUPS):
            chunk = jobs[i:i + ops.MAX_GROUPS]
            ops.layer_wgrad([job.mlp2 for job in chunk], g2[i:i + ops.MAX_GROUPS], [(W.w2t, G.dw2t, 1)])
            dests = [(W.w1t[b * D:(b + 1) * D], G.dw1t[b * D:(b + 1) * D], 1) for b in range(W.blocks)]
            ops.layer_wgrad([job.mlp1 for job in chunk],
                            [(job.du, job.u.shape[1], list(range(job.u.shape[1]))) for job in chunk], dests)
        g_last = []
        for job, (g, gs, smap) in zip(jobs, g2):
            nu, n, t = job.u.shape[1], job.t.num_nodes, job.t
            if gs == 1:
                G.colsum(g, job.B, D, G.db2, float(nu))
            else:
                G.colsum(g, job.B * nu, D, G.db2)
            G.colsum(job.du, job.B * nu, D, G.db1)
            gz = torch.empty(job.B, n, D, dtype=torch.float32, device=g.device)
            if ro == 'targetmlp':
                others = [j for j in range(n) if j != t.target_slot]
                bp = W.w1bp if W.w1bp is not None else [None] * W.blocks
                terms = [Term(job.du, nu, k, W.w1b[0], t.target_slot, bp[0]) for k in range(nu)]
                terms += [Term(job.du, nu, k, W.w1b[1], j, bp[1]) for k, j in enumerate(others)]
            else:
                bp = W.w1bp if W.w1bp is not None else [None] * W.blocks
                terms = [Term(job.du, nu, j, W.w1b[W.blocks - 1], j, bp[W.blocks - 1]) for j in range(n)]
            g_last.append(Group(job.B, terms, n, gz, n))
            last[job] = (gz, n, list(range(n)))
        ops.layer_forward(g_last)


class RowGrads(object):
    """Collects (table row, gradient row) pairs; capacities are known on the host before the backward.
    Per mode by default; with `table_offsets` ({mode: first global row}) all modes share ONE buffer and the kernels
    emit global row ids (mode offset + row), so a step needs a single sort/combine and a single all-gather."""

    def __init__(self, capacities, device, table_offsets=None, rows_buffer=None, ids_buffer=None):
        self.buf = {}
        self.table_offsets = table_offsets
        self.planned = False     # True: every slot was reserved (and its row ids emitted) before the backward
        if table_offsets is not None:
            cap = sum(capacities.values())
            # rows_buffer(cap) -> [>= cap, D] tensor, ids_buffer(cap) -> [>= cap] int64: let the caller place the pairs
            # in peer-visible memory
            rows = rows_buffer(cap) if rows_buffer is not None else torch.empty(cap, D, dtype=torch.float32, device=device)
            ids = ids_buffer(cap) if ids_buffer is not None else torch.empty(cap, dtype=torch.int64, device=device)
            self.shared = [rows, ids, 0]
            return
        for mode, cap in capacities.items():
            if cap > 0:
                self.buf[mode] = [torch.empty(cap, D, dtype=torch.float32, device=device),
                                  torch.empty(cap, dtype=torch.int64, device=device), 0]

    def reserve(self, mode, count):
        """(rows buffer, ids buffer, first entry, offset the kernel adds to the row ids)."""
        entry = self.shared if self.table_offsets is not None else self.buf[mode]
        rows, ids, used = entry
        assert used + count <= ids.numel()
        entry[2] = used + count
        return rows, ids, used, (self.table_offsets[mode] if self.table_offsets is not None else 0)


class Grads(object):
    """Dense gradient buffers (zero-initialised, kernels accumulate) + the row-gradient collector."""

    def __init__(self, model, W, row_capacities, device, table_offsets=None, rows=None, flat=None):
        self.device = device
        self.colsums, self.gathers, self.keep = [], [], []
        # one flat zeroed bucket: a single memset here, a single NCCL all-reduce in data-parallel training
        shapes = [tuple(w.shape) for w in W.w] + [tuple(r.shape) for r in W.root] + [(D,)] * len(W.root)
        shapes.append(tuple(W.mode_emb.shape))
        if W.ro is not None:
            shapes += [tuple(W.w1t.shape), tuple(W.w2t.shape), (D,), (D,)]
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        if flat is not None:       # caller-provided bucket (peer-visible memory in data-parallel training)
            if flat.numel() != sum(sizes):
                raise ops._lib.MpqeError('dense gradient bucket has %d elements, need %d' % (flat.numel(), sum(sizes)))
            self.flat = flat.zero_()
        else:
            self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self._layout = (shapes, sizes, len(W.w), W.ro is not None)
        self._bind(self.flat)
        self.rows = rows if rows is not None else RowGrads(row_capacities, device, table_offsets)

    def _bind(self, flat):
        shapes, sizes, L, has_ro = self._layout
        views, off = [], 0
        for s, k in zip(shapes, sizes):
            views.append(flat[off:off + k].view(s))
            off += k
        self.dw, self.droot, self.dbias = views[:L], views[L:2 * L], views[2 * L:3 * L]
        self.dmode = views[3 * L]
        if has_ro:
            self.dw1t, self.dw2t, self.db1, self.db2 = views[3 * L + 1:3 * L + 5]

    def over(self, flat):
        """The same views over another bucket of the same layout (the all-reduced copy of a data-parallel step)."""
        other = object.__new__(Grads)
        other.device, other.colsums, other.gathers, other.keep = self.device, [], [], []
        other._layout, other.flat, other.rows = self._layout, flat, self.rows
        other._bind(flat)
        return other

    def colsum(self, src, rows, stride, dst, scale=1.0):
        """Deferred dst += scale * column-sum(src): all of a backward's reductions run in one multi-item launch."""
        self.colsums.append(ops.ColsumItem(src, rows, stride, dst, scale))

    def flush_gathers(self):
        if self.gathers:
            ops.gather_multi(self.gathers, backward=True)
        self.gathers = []

    def flush_colsums(self):
        if self.colsums:
            self.keep.append(self.colsums)
            ops.colsum_multi(self.colsums, self.device)
        self.colsums = []

    def flush(self):
        self.flush_gathers()
        self.flush_colsums()


# ===============================================================================================================
class RGCNEncoderDecoder(nn.Module):
    def __init__(self, graph, enc, readout='mp', scatter_op='add', dropout=0, weight_decay=1e-3, num_layers=3,
                 shared_layers=True, adaptive=True):
        super(RGCNEncoderDecoder, self).__init__()
        self.enc = enc
        self.graph = graph
        self.emb_dim = graph.feature_dims[next(iter(graph.feature_dims))]
        self.mode_embeddings = nn.Embedding(len(graph.mode_weights), self.emb_dim)
        self.num_layers = num_layers
        self.adaptive = adaptive
        self.shared_layers = shared_layers
        self.mode_ids = {mode: i for i, mode in enumerate(graph.mode_weights)}
        self.rel_ids = {}
        for r1 in graph.relations:
            for r2 in graph.relations[r1]:
                self.rel_ids[(r1, r2[1], r2[0])] = len(self.rel_ids)

        self.layers = nn.ModuleList()
        for i in range(num_layers):
            if len(self.layers) == 0 or not shared_layers:
                rgcn = RGCNConv(in_channels=self.emb_dim, out_channels=self.emb_dim,
                                num_relations=len(graph.rel_edges), num_bases=0)
            self.layers.append(rgcn)

        if scatter_op not in ('add', 'max', 'mean'):
            raise ValueError(f'Unknown scatter op {scatter_op}')
        self.scatter_op = scatter_op
        self.readout_str = readout
        if readout == 'sum':
            self.readout = self.sum_readout
        elif readout == 'max':
            self.readout = self.max_readout
        elif readout == 'mlp':
            self.readout = MLPReadout(self.emb_dim, self.emb_dim, scatter_op)
        elif readout == 'targetmlp':
            self.readout = TargetMLPReadout(self.emb_dim, scatter_op)
        elif readout == 'concat':
            self.readout = MLPReadout(self.emb_dim * num_layers, self.emb_dim, scatter_op)
        elif readout == 'mp':
            self.readout = self.target_message_readout
        else:
            raise ValueError(f'Unknown readout function {readout}')

        self.dropout = nn.Dropout(dropout)  # constructed but never applied, as in the reference (model.py:377)
        self.weight_decay = weight_decay
        self.sparse_embedding_grad = getattr(enc, 'sparse_grad', False)
        self._engine = Engine(self)
        self._device_cache = {}

    # The three named readouts stay callable with the reference's signature (model.py:380-398).
    def sum_readout(self, embs, batch_idx=None, batch_size=None, num_nodes=None, **kwargs):
        return embs.reshape(batch_size, num_nodes, -1).sum(dim=1) if batch_size else embs

    def max_readout(self, embs, batch_idx=None, batch_size=None, num_nodes=None, **kwargs):
        with ops.device_guard(embs.device):
            return ops.max_readout(embs.contiguous(), batch_size, num_nodes)[0]

    def target_message_readout(self, embs, batch_size, num_nodes, num_anchors, **kwargs):
        return embs.reshape(batch_size, num_nodes, -1)[:, num_anchors]

    def distinct_layers(self):
        return [self.layers[0]] if self.shared_layers else list(self.layers)

    # ---- job construction -----------------------------------------------------------------------------
    def num_passes(self, formula):
        if self.adaptive:
            passes = RGCNQueryDataset.query_diameters[formula.query_type]
            if passes > len(self.layers):
                raise ValueError(f'RGCN is adaptive with {len(self.layers)} layers, but query requires {passes}.')
            return passes
        return self.num_layers

    def make_job(self, formula, queries, anchor_ids=None, var_ids=None, q_graphs=None):
        if self.emb_dim != D:
            raise ValueError('kernels are specialised for embed_dim=%d' % D)
        if self.readout_str == 'concat' and self.num_passes(formula) != self.num_layers:
            raise ValueError('concat readout needs num_passes == num_layers (reference model.py:369-371 vs 443-445)')
        device = self.mode_embeddings.weight.device
        ops.device_guard(device)
        if anchor_ids is None or var_ids is None or q_graphs is None:
            anchor_ids, var_ids, q_graphs = RGCNQueryDataset.get_query_graph(formula, queries, self.rel_ids,
                                                                             self.mode_ids)
        q_graphs = q_graphs.to(device)
        t = q_graphs.template if isinstance(q_graphs, QueryGraphBatch) else template_of(formula.query_type)
        rels = q_graphs.edge_rel_ids
        key = ('var', tuple(int(v) for v in var_ids.tolist()))
        var_dev = self._device_cache.get(key)
        if var_dev is None or var_dev.device != device:
            var_dev = self._device_cache[key] = var_ids.to(device)
        a_dev = anchor_ids.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
        job = Job(t, rels, var_dev, formula.anchor_modes, formula.target_mode, a_dev, self.num_passes(formula))
        job.var_rows_host = key[1]
        return job

    def _ids(self, nodes, device):
        return self.enc.ids_on_device(nodes, device).reshape(-1)

    def _params(self):
        return [p for p in self.parameters()]

    # ---- public API (reference model.py:400-494) ------------------------------------------------------
    def forward(self, formula, queries, target_nodes, anchor_ids=None, var_ids=None, q_graphs=None,
                neg_nodes=None, neg_lengths=None):
        job = self.make_job(formula, queries, anchor_ids, var_ids, q_graphs)
        device = job.anchor_ids.device
        tgt = self._ids(target_nodes, device)
        neg = offsets = None
        if neg_nodes is not None:
            neg = self._ids(neg_nodes, device)
            lengths = torch.as_tensor(neg_lengths, dtype=torch.int64)
            offsets = torch.zeros(lengths.numel() + 1, dtype=torch.int64)
            offsets[1:] = torch.cumsum(lengths, 0)
            offsets = offsets.to(device, non_blocking=True)
        return _ScoresFn.apply(self, job, tgt, neg, offsets, *self._params())

    def margin_loss(self, formula, queries, anchor_ids=None, var_ids=None, q_graphs=None, hard_negatives=False,
                    margin=1):
        if 'inter' not in formula.query_type and hard_negatives:
            raise Exception('Hard negative examples can only be used with intersection queries')
        elif hard_negatives:
            neg_nodes = [random.choice(query.hard_neg_samples) for query in queries]
        elif formula.query_type == '1-chain':
            neg_nodes = [random.choice(self.graph.full_lists[formula.target_mode]) for _ in queries]
        else:
            neg_nodes = [random.choice(query.neg_samples) for query in queries]
        return self.margin_loss_ids(formula, queries, [query.target_node for query in queries], neg_nodes,
                                    anchor_ids, var_ids, q_graphs, margin)

    def margin_loss_ids(self, formula, queries, target_nodes, neg_nodes, anchor_ids=None, var_ids=None,
                        q_graphs=None, margin=1):
        """margin_loss with the negatives chosen by the caller (tensors or lists of global node ids)."""
        job = self.make_job(formula, queries, anchor_ids, var_ids, q_graphs)
        device = job.anchor_ids.device
        loss = _MarginLossFn.apply(self, [job], [self._ids(target_nodes, device)], [self._ids(neg_nodes, device)],
                                   float(margin), *self._params())
        if isinstance(self.readout, nn.Module) and self.weight_decay > 0:
            l2_reg = 0
            for param in self.readout.parameters():
                l2_reg = l2_reg + torch.norm(param)
            loss = loss + self.weight_decay * l2_reg
        return loss

    # ---- gradient plumbing shared by the autograd functions -------------------------------------------
    def _row_capacities(self, jobs, extra):
        cap = {}
        for job in jobs:
            for mode in job.anchor_modes:
                cap[mode] = cap.get(mode, 0) + job.B
        for mode, count in extra:
            cap[mode] = cap.get(mode, 0) + count
        return cap

    def _collect_grads(self, W, G):
        """Gradients in `self.parameters()` order."""
        by_param = {}
        for li, layer in enumerate(self.distinct_layers()):
            if layer.att is None:
                by_param[id(layer.basis)] = G.dw[li]
            else:  # W_r = sum_b att[r,b] basis[b]
                dw = G.dw[li].view(layer.num_relations, -1)
                by_param[id(layer.att)] = ops.rows_dot(dw, layer.basis.detach())
                by_param[id(layer.basis)] = ops.small_k_matmul(layer.att.detach(), dw, transpose_a=True).view_as(layer.basis)
            by_param[id(layer.root)] = G.droot[li]
            if layer.bias is not None:
                by_param[id(layer.bias)] = G.dbias[li]
        by_param[id(self.mode_embeddings.weight)] = G.dmode
        if W.ro is not None:
            by_param[id(W.w1)] = ops.transpose(G.dw1t)
            by_param[id(W.w2)] = ops.transpose(G.dw2t)
            by_param[id(W.b1)], by_param[id(W.b2)] = G.db1, G.db2
        for mode, (rows, ids, used) in G.rows.buf.items():
            table = self.enc.table(mode)
            if used > 0:
                by_param[id(table)] = table_gradient(table, ids[:used], rows[:used], self.sparse_embedding_grad)
        return tuple(by_param.get(id(p)) for p in self.parameters())


def _one(x, i):
    """Element i of a per-job device vector as a 1-element view; `x` may also be a list of such views."""
    return x[i] if isinstance(x, (list, tuple)) else x[i:i + 1]


def loss_forward(model, jobs, targets, negatives, margin, need_grad, grad_losses=None, W=None, losses_out=None):
    """Encodes every job once and scores it against its positives and negatives.
    Returns (per-job losses [len(jobs)] on the device, Weights).  With `grad_losses` (d total / d loss_i, known up
    front in a training step) and row slots reserved by `plan_rows`, the margin backward (job.dq and the target /
    negative row gradients) is produced by the same pass over q and the table rows; `loss_backward` then skips it."""
    device = jobs[0].anchor_ids.device
    if W is None:
        W = Weights(model, need_grad)
    model._engine.encode(jobs, W)
    # losses_out: per-job 1-element views to write the losses to (a lane of a step writes into the step's vector)
    losses = torch.empty(len(jobs), dtype=torch.float32, device=device) if losses_out is None else losses_out
    hinge = torch.empty(sum(job.B for job in jobs), dtype=torch.float32, device=device)
    items, off = [], 0
    for i, (job, tgt, neg) in enumerate(zip(jobs, targets, negatives)):
        it = ops.MarginItem(job.q, model.enc.table(job.target_mode), model.enc.node_maps, tgt, neg,
                            hinge=hinge[off:off + job.B], loss=_one(losses, i))
        job.dq = None
        if grad_losses is not None:
            rbuf, ids, roff, id_off = job.margin_res
            job.dq = torch.empty(job.B, D, dtype=torch.float32, device=device)
            it.grad_loss, it.dq = _one(grad_losses, i), job.dq
            it.rows_out, it.rows_id, it.rows_offset, it.id_offset = rbuf, ids, roff, id_off
        items.append(it)
        off += job.B
    ops.cosine_margin_multi(items, margin, backward='both' if grad_losses is not None else False)
    return losses, W


def plan_rows(model, jobs, targets, negatives, table_offsets, rows_buffer=None, launch=True, ids_buffer=None):
    """Reserves every row-gradient slot of a step in the shared (row id, row) buffer BEFORE the forward and emits the
    row ids with one launch, so that the id-only half of the combine (`ops.SparseRowsPlan`) can overlap the step.
    The reservations are left on the jobs (`margin_res`, `anchor_res`) for `loss_backward(..., rows=...)`.
    `rows_buffer(capacity)` may supply the gradient-row buffer (peer-visible memory in data-parallel training)."""
    device = jobs[0].anchor_ids.device
    enc = model.enc
    cap = model._row_capacities(jobs, [(job.target_mode, 2 * job.B) for job in jobs])
    R = RowGrads(cap, device, table_offsets, rows_buffer, ids_buffer)
    items = []
    for job, tgt, neg in zip(jobs, targets, negatives):
        res = job.margin_res = R.reserve(job.target_mode, 2 * job.B)
        table = enc.table(job.target_mode)
        items.append(ops.GatherItem(table, enc.node_maps, tgt, job.B, rows_id=res[1], rows_offset=res[2], id_offset=res[3]))
        items.append(ops.GatherItem(table, enc.node_maps, neg, job.B, rows_id=res[1], rows_offset=res[2] + job.B,
                                    id_offset=res[3]))
    for job in jobs:
        job.anchor_res = {}
        for i in model._engine.grad_anchor_slots(job):
            mode = job.anchor_modes[i]
            res = job.anchor_res[i] = R.reserve(mode, job.B)
            items.append(ops.GatherItem(enc.table(mode), enc.node_maps, job.anchor_ids, job.B, ids_offset=i,
                                        ids_stride=job.t.num_anchors, rows_id=res[1], rows_offset=res[2], id_offset=res[3]))
    R.id_items = items
    if launch:
        ops.gather_multi(items, backward='ids')
    R.planned = True
    return R


def loss_backward(model, jobs, W, targets, negatives, margin, grad_losses, table_offsets=None, rows=None,
                  defer_constant=False, flat=None, side=None):
    """Backward of `loss_forward` for d(total)/d(loss_i) = grad_losses[i] (device tensor [len(jobs)]).
    Returns the filled `Grads` (dense bucket + (row id, gradient row) pairs, not yet combined).  `rows`: a RowGrads
    from `plan_rows` (slots reserved and ids already emitted).  `defer_constant`: leave the batch-constant tail of the
    backward to the caller as `G.finish()` (the dense gradients are complete only after it)."""
    device = jobs[0].anchor_ids.device
    cap = model._row_capacities(jobs, [(job.target_mode, 2 * job.B) for job in jobs]) if rows is None else None
    G = Grads(model, W, cap, device, table_offsets, rows, flat)
    dqs, items = [], []
    for i, (job, tgt, neg) in enumerate(zip(jobs, targets, negatives)):
        if G.rows.planned and getattr(job, 'dq', None) is not None:
            dqs.append(job.dq)      # margin backward already done by loss_forward(grad_losses=...)
            continue
        rbuf, ids, off, id_off = job.margin_res if G.rows.planned else G.rows.reserve(job.target_mode, 2 * job.B)
        dq = torch.empty(job.B, D, dtype=torch.float32, device=device)
        items.append(ops.MarginItem(job.q, model.enc.table(job.target_mode), model.enc.node_maps, tgt, neg,
                                    grad_loss=_one(grad_losses, i), dq=dq, rows_out=rbuf, rows_id=ids, rows_offset=off,
                                    id_offset=id_off))
        dqs.append(dq)
    if items:
        ops.cosine_margin_multi(items, margin, backward=True)
    model._engine.backward(jobs, W, dqs, G, defer_constant=defer_constant, side=side)
    return G


class _MarginLossFn(torch.autograd.Function):
    """sum over formula groups of mean(relu(margin - (cos(q, target) - cos(q, negative))))."""

    @staticmethod
    def forward(ctx, model, jobs, targets, negatives, margin, *params):
        device = jobs[0].anchor_ids.device
        with ops.device_guard(device):
            losses, W = loss_forward(model, jobs, targets, negatives, margin, any(ctx.needs_input_grad))
            total = losses[0].clone() if len(jobs) == 1 else losses.sum()
        ctx.model, ctx.jobs, ctx.W, ctx.margin = model, jobs, W, margin
        ctx.targets, ctx.negatives = targets, negatives
        return total

    @staticmethod
    def backward(ctx, grad_loss):
        model, jobs, W = ctx.model, ctx.jobs, ctx.W
        with ops.device_guard(jobs[0].anchor_ids.device):
            gl = grad_loss.contiguous().reshape(1).expand(len(jobs)).contiguous()
            G = loss_backward(model, jobs, W, ctx.targets, ctx.negatives, ctx.margin, gl)
            grads = model._collect_grads(W, G)
        return (None, None, None, None, None) + grads


class _ScoresFn(torch.autograd.Function):
    """scores = cat(cos(q, targets), cos(q[owner], negatives))  (reference model.py:451-460)."""

    @staticmethod
    def forward(ctx, model, job, tgt, neg, offsets, *params):
        device = job.anchor_ids.device
        need_grad = any(ctx.needs_input_grad)
        table = model.enc.table(job.target_mode)
        with ops.device_guard(device):
            W = Weights(model, need_grad)
            model._engine.encode([job], W)
            total = job.B + (neg.numel() if neg is not None else 0)
            scores = torch.empty(total, dtype=torch.float32, device=device)
            ops.cosine_scores(job.q, table, model.enc.node_maps, tgt, out=scores)
            if neg is not None and neg.numel() > 0:
                ops.cosine_scores(job.q, table, model.enc.node_maps, neg, offsets=offsets, out=scores, out_offset=job.B)
        ctx.model, ctx.job, ctx.W = model, job, W
        ctx.tgt, ctx.neg, ctx.offsets = tgt, neg, offsets
        return scores

    @staticmethod
    def backward(ctx, grad_scores):
        model, job, W = ctx.model, ctx.job, ctx.W
        device = job.anchor_ids.device
        table = model.enc.table(job.target_mode)
        nneg = ctx.neg.numel() if ctx.neg is not None else 0
        with ops.device_guard(device):
            G = Grads(model, W, model._row_capacities([job], [(job.target_mode, job.B + nneg)]), device)
            gs = grad_scores.contiguous()
            dq = torch.empty(job.B, D, dtype=torch.float32, device=device)
            rows, ids, off, _ = G.rows.reserve(job.target_mode, job.B)
            ops.cosine_scores_bwd(job.q, table, model.enc.node_maps, ctx.tgt, None, gs, 0, dq, False, rows, ids, off)
            if nneg > 0:
                rows, ids, off, _ = G.rows.reserve(job.target_mode, nneg)
                ops.cosine_scores_bwd(job.q, table, model.enc.node_maps, ctx.neg, ctx.offsets, gs, job.B, dq, True,
                                      rows, ids, off)
            model._engine.backward([job], W, [dq], G)
            grads = model._collect_grads(W, G)
        return (None, None, None, None, None) + grads

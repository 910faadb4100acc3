"""CPU emulation of the C-ABI kernel semantics, for testing the HOST-SIDE orchestration without a GPU.

TEST INFRASTRUCTURE ONLY.  `install(monkeypatch)` swaps the functions of `mpqe_b200.ops` for torch-CPU
re-statements of what each CUDA entry point is documented to compute (include/mpqe_b200.h), so that the term
lists, slot maps, gradient routing and autograd wiring in `mpqe_b200/model.py` can be checked against the oracle
in the `not gpu` suite.  The product never selects this code: outside these tests `ops` always calls the CUDA
library and raises when it is unavailable.
"""
import contextlib

import torch

from mpqe_b200 import ops

D = ops.D


def _rows(t, slots, slot, B):
    """[B, D] view of A[q] = t.flat[(q*slots + slot)*D : +D]."""
    flat = t.reshape(-1)
    if slots == 0:
        return flat[slot * D:(slot + 1) * D].unsqueeze(0).expand(B, D)
    return flat[:B * slots * D].view(B, slots, D)[:, slot]


def layer_forward(groups, use_tensor_cores=None):
    for g in groups:
        B = g.num_queries
        acc = [torch.zeros(B, D) for _ in range(g.num_out_slots)]
        for t in g.terms:
            acc[t.out_slot] = acc[t.out_slot] + _rows(t.a, t.a_slots, t.a_slot, B) @ t.m.reshape(D, D)
        out = g.out.view(B, g.out_slots, D)
        for j in range(g.num_out_slots):
            v = acc[j]
            if g.bias is not None:
                stride = getattr(g, 'bias_slot_stride', 0)
                v = v + g.bias_scale[j] * g.bias.reshape(-1)[j * stride:j * stride + D]
            s = g.out_slot_map[j]
            if g.epilogue == ops.EPI_RELU:
                v = torch.relu(v)
            elif g.epilogue == ops.EPI_MASK:
                v = v * (g.mask.view(B, g.mask_slots, D)[:, s] > 0).float()
            out[:, s] = v


def layer_wgrad(groups, grad_operands, dests, ctas_hint=296, use_tensor_cores=None):
    for (m_fwd, dm, acc) in dests:
        total = torch.zeros(D, D)
        for g, (gt, g_slots, smap) in zip(groups, grad_operands):
            for t in g.terms:
                if t.m.data_ptr() != m_fwd.data_ptr():
                    continue
                a = _rows(t.a, t.a_slots, t.a_slot, g.num_queries)
                gr = _rows(gt, g_slots, smap[t.out_slot], g.num_queries)
                total = total + a.t() @ gr
        if acc:
            dm += total
        else:
            dm.copy_(total)


def colsum(src, rows, stride, out, scale=1.0, accumulate=False):
    v = torch.as_strided(src, (rows, D), (stride, 1), src.storage_offset()).sum(0) * scale
    if accumulate:
        out += v
    else:
        out.copy_(v)


def transpose(src, dst=None):
    res = src.transpose(-1, -2).contiguous()
    if dst is not None:
        dst.copy_(res)
        return dst
    return res


def _resolve(id2row, ids):
    return id2row[ids] if id2row is not None else ids


def _norm_rows(table, rows):
    v = table[rows]
    nrm = v.norm(dim=1, keepdim=True)
    return v / nrm, nrm


def gather_normalize(table, id2row, ids, out=None, out_offset=0, out_stride=D, ids_offset=0, ids_stride=1,
                     count=None, inv_norm=None):
    if count is None:
        count = ids.numel()
    sel = ids.reshape(-1)[ids_offset::ids_stride][:count]
    y, _ = _norm_rows(table, _resolve(id2row, sel))
    if out is None:
        return y
    torch.as_strided(out, (count, D), (out_stride, 1), out.storage_offset() + out_offset).copy_(y)
    return out


def _normalize_bwd(g, y, nrm):
    return (g - (g * y).sum(1, keepdim=True) * y) / nrm


def gather_normalize_bwd(table, id2row, ids, grad, rows_out, rows_id, grad_offset=0, grad_stride=D, ids_offset=0,
                         ids_stride=1, count=None, rows_offset=0):
    if count is None:
        count = ids.numel()
    sel = ids.reshape(-1)[ids_offset::ids_stride][:count]
    rows = _resolve(id2row, sel)
    y, nrm = _norm_rows(table, rows)
    g = torch.as_strided(grad, (count, D), (grad_stride, 1), grad.storage_offset() + grad_offset)
    rows_out[rows_offset:rows_offset + count] = _normalize_bwd(g, y, nrm)
    rows_id[rows_offset:rows_offset + count] = rows


def broadcast_rows(src, src_rows, out, out_offset, out_stride, count):
    k = src_rows.numel()
    view = torch.as_strided(out, (count, k, D), (out_stride, D, 1), out.storage_offset() + out_offset)
    view.copy_(src[src_rows].unsqueeze(0).expand(count, k, D))


def max_readout(z, B, n):
    zz = z.view(B, n, D)
    q = zz.max(dim=1).values
    first = (zz == q.unsqueeze(1)).float().argmax(dim=1)  # first index attaining the max
    return q, first + torch.arange(B).view(B, 1) * n


def max_readout_bwd(dq, argmax, B, n):
    local = argmax - torch.arange(B).view(B, 1) * n
    return (local.unsqueeze(1) == torch.arange(n).view(1, n, 1)).float() * dq.unsqueeze(1)


EPS = 1e-8


def _cos(q, y):
    nq, ny = q.norm(dim=1), y.norm(dim=1)
    return ((q / nq.clamp_min(EPS).unsqueeze(1)) * (y / ny.clamp_min(EPS).unsqueeze(1))).sum(1)


def cosine_margin(q, table, id2row, ids_pos, ids_neg, margin, loss_out=None):
    yp, _ = _norm_rows(table, _resolve(id2row, ids_pos))
    yn, _ = _norm_rows(table, _resolve(id2row, ids_neg))
    sp, sn = _cos(q, yp), _cos(q, yn)
    loss = torch.clamp(margin - (sp - sn), min=0).mean()
    if loss_out is not None:
        loss_out.copy_(loss.reshape(loss_out.shape))
        loss = loss_out
    return sp, sn, loss


def _pair_bwd(q, table, rows, gscore):
    """gradients of sum(gscore * cos(q, normalise(table[rows]))) wrt q and wrt the raw table rows."""
    with torch.enable_grad():
        qq = q.detach().clone().requires_grad_(True)
        raw = table[rows].detach().clone().requires_grad_(True)
        y = raw / raw.norm(dim=1, keepdim=True)
        (gscore.detach() * _cos(qq, y)).sum().backward()
    return qq.grad, raw.grad


def cosine_margin_bwd(q, table, id2row, ids_pos, ids_neg, margin, grad_loss, rows_out, rows_id, rows_offset=0):
    B = q.shape[0]
    rp, rn = _resolve(id2row, ids_pos), _resolve(id2row, ids_neg)
    yp, _ = _norm_rows(table, rp)
    yn, _ = _norm_rows(table, rn)
    active = ((margin - (_cos(q, yp) - _cos(q, yn))) >= 0).float()
    g = active * grad_loss.reshape(()) / B
    dqp, drp = _pair_bwd(q, table, rp, -g)
    dqn, drn = _pair_bwd(q, table, rn, g)
    rows_out[rows_offset:rows_offset + B] = drp
    rows_out[rows_offset + B:rows_offset + 2 * B] = drn
    rows_id[rows_offset:rows_offset + B] = rp
    rows_id[rows_offset + B:rows_offset + 2 * B] = rn
    return dqp + dqn


def _owner(offsets, count, B):
    if offsets is None:
        return torch.arange(count)
    return torch.repeat_interleave(torch.arange(B), offsets[1:] - offsets[:-1])


def cosine_scores(q, table, id2row, ids, offsets=None, out=None, out_offset=0):
    y, _ = _norm_rows(table, _resolve(id2row, ids))
    s = _cos(q[_owner(offsets, ids.numel(), q.shape[0])], y)
    if out is None:
        return s
    out[out_offset:out_offset + s.numel()] = s
    return out


def cosine_scores_bwd(q, table, id2row, ids, offsets, grad_scores, grad_offset, dq, accumulate, rows_out, rows_id,
                      rows_offset=0):
    count, B = ids.numel(), q.shape[0]
    owner = _owner(offsets, count, B)
    rows = _resolve(id2row, ids)
    g = grad_scores.reshape(-1)[grad_offset:grad_offset + count]
    dqi, dr = _pair_bwd(q[owner], table, rows, g)
    tot = torch.zeros(B, D).index_add(0, owner, dqi)
    if accumulate:
        dq += tot
    else:
        dq.copy_(tot)
    rows_out[rows_offset:rows_offset + count] = dr
    rows_id[rows_offset:rows_offset + count] = rows


def rank_counts_ragged(pos, neg, offsets):
    B = pos.numel()
    owner = _owner(offsets, neg.numel(), B)
    lt = torch.zeros(B, dtype=torch.int64).index_add(0, owner, (neg < pos[owner]).long())
    le = torch.zeros(B, dtype=torch.int64).index_add(0, owner, (neg <= pos[owner]).long())
    return lt, le


def rank_counts_table(q, pos, table, row_begin, row_end, left, right, use_tensor_cores=False):
    rows = table[row_begin:row_end]
    s = (q @ rows.t()) / rows.norm(dim=1).unsqueeze(0) / q.norm(dim=1).clamp_min(EPS).unsqueeze(1)
    left += (s < pos.unsqueeze(1)).sum(1)
    right += (s <= pos.unsqueeze(1)).sum(1)


def build_query_graph(n, src, dst, rel, B, device):
    E = len(src)
    shift = (torch.arange(B) * n).view(B, 1)
    ei = torch.stack([(torch.tensor(src).view(1, E) + shift).reshape(-1),
                      (torch.tensor(dst).view(1, E) + shift).reshape(-1)])
    return ei, torch.tensor(rel, dtype=torch.int64).repeat(B), torch.arange(B).repeat_interleave(n)


def relation_sort(edge_type, num_relations):
    perm = torch.sort(edge_type, stable=True)[1]
    off = torch.zeros(num_relations + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.bincount(edge_type, minlength=num_relations), 0)
    return perm, off


def sparse_rows_combine(rows_id, rows, table_rows, pad_id=0):
    count = rows_id.numel()
    keep = rows_id < table_rows
    uniq, inv = torch.unique(rows_id[keep], sorted=True, return_inverse=True)
    uid = torch.full((count,), pad_id, dtype=torch.int64)
    urows = torch.zeros(count, D)
    uid[:uniq.numel()] = uniq
    urows.index_add_(0, inv, rows[keep])
    return uid, urows, torch.tensor([uniq.numel()], dtype=torch.int64)


class SparseRowsPlan(object):
    """plan(ids) + apply(rows) == sparse_rows_combine(ids, rows)."""

    def __init__(self, rows_id, table_rows):
        self.rows_id, self.table_rows = rows_id.clone(), table_rows

    def apply(self, rows, pad_id=0, scale=1.0):
        uid, urows, num = sparse_rows_combine(self.rows_id, rows, self.table_rows, pad_id)
        return uid, urows * scale, num

    def apply_peers(self, peer_rows, per_rank_count, pad_id=0, scale=1.0, capacity=None):
        uid, urows, num = self.apply(torch.cat([r[:per_rank_count] for r in peer_rows]), pad_id, scale)
        cap = uid.numel() if capacity is None else max(1, min(int(capacity), uid.numel()))
        assert int(num) <= cap
        return uid[:cap], urows[:cap], num


def owner_plan(id_sources, rank, per_rank_count, table_begin, table_rows, total_rows, device):
    """mpqe_sparse_rows_plan_owner: ids of all ranks, those `rank` does not own clamped to the sentinel."""
    world = len(id_sources)
    ids = torch.cat([t[:per_rank_count] for t in id_sources]).clone()
    owned = torch.zeros_like(ids, dtype=torch.bool)
    for begin, rows in zip(table_begin, table_rows):
        chunk = (rows + world - 1) // world
        inside = (ids >= begin) & (ids < begin + rows)
        owned |= inside & (((ids - begin) // chunk) == rank)
    ids[~owned] = total_rows
    return SparseRowsPlan(ids, total_rows)


def scatter_rows(ids, rows, num, dense, accumulate=False):
    k = int(num.item()) if num is not None else ids.numel()
    if accumulate:
        dense[ids[:k]] += rows[:k]
    else:
        dense[ids[:k]] = rows[:k]


def install(monkeypatch):
    for name in ('layer_forward', 'layer_wgrad', 'colsum', 'transpose', 'gather_normalize', 'gather_normalize_bwd',
                 'broadcast_rows', 'max_readout', 'max_readout_bwd', 'cosine_margin', 'cosine_margin_bwd',
                 'cosine_scores', 'cosine_scores_bwd', 'rank_counts_ragged', 'rank_counts_table',
                 'build_query_graph', 'relation_sort', 'sparse_rows_combine', 'SparseRowsPlan', 'scatter_rows',
                 'gather_multi', 'matrix_sum_multi', 'small_k_matmul', 'rows_dot',
                 'cosine_margin_multi', 'colsum_multi', 'l2_reg', 'owner_plan'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'device_guard', lambda device: contextlib.nullcontext())
    monkeypatch.setattr(ops, 'tensor_cores_default', lambda: False)
    # the fused training step orders its work over two CUDA streams; on the CPU everything simply runs in order
    from mpqe_b200 import model as M
    from mpqe_b200.train_step import TrainStep

    def weights_now(self, jobs, dev):
        W = M.Weights(self.model, True)
        self.model._engine.prepare(jobs, W)
        return W

    monkeypatch.setattr(TrainStep, '_plan_on_side_stream', lambda self, make_plan, dev, keep=(): make_plan())
    monkeypatch.setattr(TrainStep, '_weights_on_side_stream', weights_now)
    monkeypatch.setattr(TrainStep, '_join_side', lambda self, dev: None)
    monkeypatch.setattr(TrainStep, '_wait_plan', lambda self, dev: None)
    monkeypatch.setattr(TrainStep, '_on_side_stream', lambda self, dev, fn: fn())
    monkeypatch.setattr(TrainStep, '_on_wgrad_stream', lambda self, dev, fn: fn())
    monkeypatch.setattr(TrainStep, '_join_wgrad', lambda self, dev: None)


def l2_reg(params, grads, weight_decay, grad_scale, losses=None, norms=None):
    total = 0.0
    for i, (p, g) in enumerate(zip(params, grads)):
        nrm = torch.sqrt((p * p).sum())
        total = total + nrm
        if g is not None and float(nrm) > 0:
            g += (grad_scale * weight_decay / nrm) * p.reshape(g.shape)
        if norms is not None:
            norms[i] = nrm
    if losses is not None:
        losses += weight_decay * total


def gather_multi(items, backward=False):
    for it in items:
        if backward == 'ids':
            idx = it.ids.reshape(-1)[it.ids_offset::max(it.ids_stride, 1)][:it.count]
            it.rows_id[it.rows_offset:it.rows_offset + it.count] = _resolve(it.id2row, idx) + it.id_offset
        elif backward:
            gather_normalize_bwd(it.table, it.id2row, it.ids, it.grad, it.rows_out, it.rows_id, it.grad_offset,
                                 it.grad_stride, it.ids_offset, max(it.ids_stride, 1), it.count, it.rows_offset)
            it.rows_id[it.rows_offset:it.rows_offset + it.count] += it.id_offset
        elif it.normalize:
            gather_normalize(it.table, it.id2row, it.ids, it.out, it.out_offset, it.out_stride, it.ids_offset,
                             it.ids_stride, it.count)
            if getattr(it, 'norm', None) is not None:
                idx = it.ids.reshape(-1)[it.ids_offset::max(it.ids_stride, 1)][:it.count]
                it.norm.reshape(-1)[it.norm_offset:it.norm_offset + it.count] = it.table[_resolve(it.id2row, idx)].norm(dim=1)
        else:  # plain row copy; ids_stride 0 broadcasts one row
            row = it.ids.reshape(-1)[it.ids_offset]
            view = torch.as_strided(it.out, (it.count, D), (it.out_stride, 1), it.out.storage_offset() + it.out_offset)
            view.copy_(it.table[row].unsqueeze(0).expand(it.count, D))


def cosine_margin_multi(items, margin, backward=False):
    for it in items:
        if backward == 'both':
            cosine_margin(it.q, it.table, it.id2row, it.ids_pos, it.ids_neg, margin, loss_out=it.loss)
        if backward:
            dq = cosine_margin_bwd(it.q, it.table, it.id2row, it.ids_pos, it.ids_neg, margin, it.grad_loss, it.rows_out,
                                   it.rows_id, it.rows_offset)
            it.dq.copy_(dq)
            B = it.q.shape[0]
            it.rows_id[it.rows_offset:it.rows_offset + 2 * B] += it.id_offset
        else:
            cosine_margin(it.q, it.table, it.id2row, it.ids_pos, it.ids_neg, margin, loss_out=it.loss)


def small_k_matmul(a, b, transpose_a=False):
    a2 = a.t() if transpose_a else a
    return (a2 @ b.reshape(b.shape[0], -1)).reshape((a2.shape[0],) + tuple(b.shape[1:]))


def rows_dot(x, y):
    return x.reshape(x.shape[0], -1) @ y.reshape(y.shape[0], -1).t()


def matrix_sum_multi(items):
    for dst, srcs, acc in items:
        total = dst.clone() if acc else torch.zeros_like(dst)
        for m in srcs:
            total = total + m.reshape(dst.shape)
        dst.copy_(total)


def colsum_multi(items, device):
    for it in items:
        colsum(it.src, it.rows, it.stride, it.dst, it.scale, True)

"""Each CUDA entry point against a CPU restatement on seeded inputs (run with -m gpu on the B200 box).
Integer results must be bit-exact; fp32 results within the tolerance stated per test."""
import numpy as np
import pytest
import torch

from mpqe_b200 import ops
from tests import emulator as E
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
D = ops.D
DEV = 'cuda:0'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def run_layer(groups_cpu, use_tc=False):
    """groups given with CPU tensors -> run the same on GPU, return outputs."""
    moved = {}

    def mv(t):
        if t is None:
            return None
        key = (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape))
        if key not in moved:
            base_key = t.untyped_storage().data_ptr()
            if base_key not in moved:
                base = torch.empty(0)
                base.set_(t.untyped_storage())
                moved[base_key] = base.to(DEV)
            moved[key] = moved[base_key].as_strided(t.shape, t.stride(), t.storage_offset())
        return moved[key]

    gg = []
    for g in groups_cpu:
        terms = [ops.Term(mv(t.a), t.a_slots, t.a_slot, mv(t.m), t.out_slot) for t in g.terms]
        gg.append(ops.Group(g.num_queries, terms, g.num_out_slots, mv(g.out), g.out_slots, g.out_slot_map, g.epilogue,
                            mv(g.bias), g.bias_scale, mv(g.mask), g.mask_slots))
    return gg


def need_tc(tc):
    from mpqe_b200 import _lib
    if tc and not _lib.load().mpqe_b200_has_tcgen05():
        pytest.skip('library built without tcgen05 kernels')


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('B', [1, 63, 64, 65, 255, 257, 300, 4096])
@pytest.mark.parametrize('epi', [ops.EPI_NONE, ops.EPI_RELU, ops.EPI_MASK])
def test_layer_forward_matches_cpu(B, epi, tc):
    need_tc(tc)
    n = 4
    x = rnd(B, n, D, seed=1)
    w = rnd(5, D, D, seed=2, scale=0.05)
    bias = rnd(D, seed=3)
    mask = rnd(B, n, D, seed=4)
    out = torch.full((B, n, D), float('nan'))
    terms = [ops.Term(x, n, 0, w[0], 3), ops.Term(x, n, 1, w[1], 3), ops.Term(x, n, 2, w[2], 3),
             ops.Term(x, n, 3, w[4], 3), ops.Term(x, n, 0, w[4], 0), ops.Term(x, n, 1, w[4], 1),
             ops.Term(x, n, 2, w[4], 2)]
    g = ops.Group(B, terms, n, out, n, epilogue=epi, bias=bias, bias_scale=[1, 1, 1, 2.0],
                  mask=mask if epi == ops.EPI_MASK else None, mask_slots=n)
    E.layer_forward([g])
    gg = run_layer([g])
    gg[0].out.fill_(float('nan'))
    ops.layer_forward(gg, use_tensor_cores=tc)
    got = gg[0].out.cpu()
    print('tc=%s max abs err %.3e (max |out| %.3f)' % (tc, float((got - out).abs().max()), float(out.abs().max())))
    # fp32 tolerance: K=128..512 products of O(1)*O(0.05) -> abs error ~1e-6 (FFMA and 3xTF32 alike)
    assert_close(got.numpy(), out.numpy(), 1e-5, 2e-5, 'layer out')


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('B', [40, 513])
def test_layer_forward_slot_without_terms_and_per_slot_bias(B, tc):
    """An output slot no term writes to receives only its bias (batch-constant contributions arrive that way), and
    `bias_slot_stride` selects a bias vector per slot."""
    need_tc(tc)
    x = rnd(B, 3, D, seed=1)
    w = rnd(2, D, D, seed=2, scale=0.05)
    bias = rnd(3, D, seed=3)
    out = torch.full((B, 3, D), float('nan'))
    g = ops.Group(B, [ops.Term(x, 3, 0, w[0], 2), ops.Term(x, 3, 1, w[1], 2), ops.Term(x, 3, 2, w[1], 0)], 3, out, 3,
                  epilogue=ops.EPI_RELU, bias=bias, bias_scale=[1.0, -2.0, 0.5], bias_slot_stride=D)
    E.layer_forward([g])
    gg = run_layer([g])
    gg[0].bias_slot_stride = D
    gg[0].out.fill_(float('nan'))
    ops.layer_forward(gg, use_tensor_cores=tc)
    assert_close(gg[0].out.cpu().numpy(), out.numpy(), 1e-5, 2e-5, 'layer out')


@pytest.mark.parametrize('B', [70, 256, 1000])
def test_relu_sign_bits_replace_the_fp32_mask(B):
    """tcgen05 kernel: an EPI_RELU launch also writes the ReLU sign bits (1 bit per element, word per 32 queries); an
    EPI_MASK launch that reads them gives the same bits as one that reads the fp32 activations."""
    need_tc(True)
    n = 3
    x = rnd(B, n, D, seed=1).to(DEV)
    w = (rnd(4, D, D, seed=2, scale=0.05)).to(DEV)
    bias = rnd(D, seed=3).to(DEV)
    h = torch.empty(B, n, D, device=DEV)
    bits = ops.relu_bits(B, n, torch.device(DEV))
    bits.fill_(-1)
    fwd = ops.Group(B, [ops.Term(x, n, 0, w[0], 1), ops.Term(x, n, 1, w[1], 1), ops.Term(x, n, 2, w[2], 0)], 2, h, n,
                    out_slot_map=[0, 2], epilogue=ops.EPI_RELU, bias=bias, bits_out=bits)
    ops.layer_forward([fwd], use_tensor_cores=True)
    torch.cuda.synchronize()
    # the words of the written slots spell (h > 0); queries past B read as 0
    hb = torch.zeros((B + 31) // 32 * 32, n, D, dtype=torch.bool, device=DEV)
    hb[:B] = h > 0
    want = (hb.view(-1, 32, n, D).to(torch.int64) << torch.arange(32, device=DEV).view(1, 32, 1, 1)).sum(1)
    got = bits.to(torch.int64) & 0xffffffff
    for slot in (0, 2):
        assert torch.equal(got[:, slot], want[:, slot])
    g = rnd(B, n, D, seed=5).to(DEV)
    outs = []
    for mb in (None, bits):
        dx = torch.full((B, n, D), float('nan'), device=DEV)
        back = ops.Group(B, [ops.Term(g, n, 2, w[3], 0), ops.Term(g, n, 0, w[1], 1)], 2, dx, n, out_slot_map=[0, 2],
                         epilogue=ops.EPI_MASK, mask=h, mask_slots=n, mask_bits=mb)
        ops.layer_forward([back], use_tensor_cores=True)
        outs.append(dx)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][:, 0], outs[1][:, 0]) and torch.equal(outs[0][:, 2], outs[1][:, 2])


@pytest.mark.parametrize('tc', [False, True])
def test_layer_forward_multi_group_slot_maps_and_broadcast(tc):
    need_tc(tc)
    B1, B2 = 130, 70
    x1, x2 = rnd(B1, 3, D, seed=1), rnd(B2, 2, D, seed=2)
    vrow = rnd(4, D, seed=3)
    w = rnd(3, D, D, seed=4, scale=0.05)
    bias = rnd(D, seed=5)
    out1 = torch.zeros(B1, 1, D)
    out2 = torch.full((B2, 4, D), 7.0)
    g1 = ops.Group(B1, [ops.Term(x1, 3, j, w[j], 0) for j in range(3)] + [ops.Term(vrow, 0, 2, w[0], 0)], 1, out1, 1,
                   bias=bias, bias_scale=[3.0])
    g2 = ops.Group(B2, [ops.Term(x2, 2, 0, w[1], 0), ops.Term(x2, 2, 1, w[2], 1)], 2, out2, 4, out_slot_map=[3, 1],
                   epilogue=ops.EPI_RELU)
    E.layer_forward([g1, g2])
    gg = run_layer([g1, g2])
    gg[0].out.zero_()
    gg[1].out.fill_(7.0)
    ops.layer_forward(gg, use_tensor_cores=tc)
    assert_close(gg[0].out.cpu().numpy(), out1.numpy(), 1e-5, 2e-5, 'group 1')
    assert_close(gg[1].out.cpu().numpy(), out2.numpy(), 1e-5, 2e-5, 'group 2 (untouched slots keep 7.0)')


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('B', [5, 64, 1000, 5000])
def test_layer_wgrad_matches_cpu_and_is_deterministic(B, tc):
    need_tc(tc)
    n = 3
    x = rnd(B, n, D, seed=1)
    g = rnd(B, n, D, seed=2)
    dq = rnd(B, D, seed=5)
    w = rnd(3, D, D, seed=3)
    grp = ops.Group(B, [ops.Term(x, n, 0, w[0], 2), ops.Term(x, n, 1, w[0], 2), ops.Term(x, n, 2, w[2], 2),
                        ops.Term(x, n, 0, w[2], 0), ops.Term(x, n, 1, w[2], 1)], n, None, n)
    grp2 = ops.Group(B, [ops.Term(x, n, 1, w[1], 0), ops.Term(x, n, 2, w[2], 0)], 1, None, 1)
    dm = torch.zeros(3, D, D)
    dm[1] = 1.0
    E.layer_wgrad([grp, grp2], [(g, n, [0, 1, 2]), (dq, 1, [0])], [(w[0], dm[0], 0), (w[1], dm[1], 1), (w[2], dm[2], 0)])
    xs, gs, dqs, ws = x.to(DEV), g.to(DEV), dq.to(DEV), w.to(DEV)
    gg = ops.Group(B, [ops.Term(xs, n, 0, ws[0], 2), ops.Term(xs, n, 1, ws[0], 2), ops.Term(xs, n, 2, ws[2], 2),
                       ops.Term(xs, n, 0, ws[2], 0), ops.Term(xs, n, 1, ws[2], 1)], n, None, n)
    gg2 = ops.Group(B, [ops.Term(xs, n, 1, ws[1], 0), ops.Term(xs, n, 2, ws[2], 0)], 1, None, 1)
    outs = []
    for _ in range(2):
        dmd = torch.zeros(3, D, D, device=DEV)
        dmd[1] = 1.0
        ops.layer_wgrad([gg, gg2], [(gs, n, [0, 1, 2]), (dqs, 1, [0])],
                        [(ws[0], dmd[0], 0), (ws[1], dmd[1], 1), (ws[2], dmd[2], 0)], use_tensor_cores=tc)
        outs.append(dmd.cpu())
    scale = float(dm.abs().max())
    print('tc=%s max abs err %.3e (scale %.3f)' % (tc, float((outs[0] - dm).abs().max()), scale))
    assert_close(outs[0].numpy(), dm.numpy(), 1e-4, 1e-5 * scale, 'dM')
    assert torch.equal(outs[0], outs[1]), 'weight gradient must be bit-reproducible'


def test_colsum_and_transpose():
    src = rnd(1000, 3, D, seed=1)
    out = torch.ones(D)
    E.colsum(src[:, 1], 1000, 3 * D, out, scale=2.0, accumulate=True)
    s = src.to(DEV)
    o = torch.ones(D, device=DEV)
    ops.colsum(s[:, 1], 1000, 3 * D, o, scale=2.0, accumulate=True)
    assert_close(o.cpu().numpy(), out.numpy(), 1e-5, 1e-4, 'colsum')
    m = rnd(5, D, 2 * D, seed=2)
    assert torch.equal(ops.transpose(m.to(DEV)).cpu(), m.transpose(1, 2).contiguous())
    m2 = rnd(37, 91, seed=3)
    assert torch.equal(ops.transpose(m2.to(DEV)).cpu(), m2.t().contiguous())


def test_gather_normalize_fwd_bwd():
    table = rnd(50, D, seed=1, scale=1.0 / D)
    id2row = torch.randperm(200, generator=torch.Generator().manual_seed(0)) % 50
    ids = torch.randint(0, 200, (33, 2), generator=torch.Generator().manual_seed(1))
    out = torch.zeros(33, 3, D)
    E.gather_normalize(table, id2row, ids, out=out, out_offset=D, out_stride=3 * D, ids_offset=1, ids_stride=2, count=33)
    od = torch.zeros(33, 3, D, device=DEV)
    ops.gather_normalize(table.to(DEV), id2row.to(DEV), ids.to(DEV), out=od, out_offset=D, out_stride=3 * D,
                         ids_offset=1, ids_stride=2, count=33)
    assert_close(od.cpu().numpy(), out.numpy(), 2e-6, 1e-7, 'normalised rows')
    assert torch.equal(od[:, 0].cpu(), torch.zeros(33, D))
    grad = rnd(33, 3, D, seed=2)
    rows, rid = torch.zeros(40, D), torch.zeros(40, dtype=torch.int64)
    E.gather_normalize_bwd(table, id2row, ids, grad, rows, rid, grad_offset=D, grad_stride=3 * D, ids_offset=1,
                           ids_stride=2, count=33, rows_offset=7)
    rows_d, rid_d = torch.zeros(40, D, device=DEV), torch.zeros(40, dtype=torch.int64, device=DEV)
    ops.gather_normalize_bwd(table.to(DEV), id2row.to(DEV), ids.to(DEV), grad.to(DEV), rows_d, rid_d, grad_offset=D,
                             grad_stride=3 * D, ids_offset=1, ids_stride=2, count=33, rows_offset=7)
    assert torch.equal(rid_d.cpu(), rid)
    assert_close(rows_d.cpu().numpy(), rows.numpy(), 1e-4, 1e-3, 'row grads')  # rows are O(100): abs tol scaled


def test_max_readout_ties_pick_smallest_node():
    B, n = 70, 4
    z = torch.relu(rnd(B, n, D, seed=1))  # exact zeros -> real ties
    z[:, 2] = z[:, 1]                     # and duplicated rows
    q, arg = E.max_readout(z, B, n)
    qd, argd = ops.max_readout(z.to(DEV), B, n)
    assert torch.equal(qd.cpu(), q)
    assert torch.equal(argd.cpu(), arg), 'argmax must be bit-exact (smallest node row among maxima)'
    dq = rnd(B, D, seed=2)
    assert torch.equal(ops.max_readout_bwd(dq.to(DEV), argd, B, n).cpu(), E.max_readout_bwd(dq, arg, B, n))


def test_cosine_margin_fwd_bwd():
    B = 77
    table = rnd(60, D, seed=1, scale=1.0 / D)
    q = rnd(B, D, seed=2)
    q[3] = 0  # zero query embedding -> eps clamp path
    ip = torch.randint(0, 60, (B,), generator=torch.Generator().manual_seed(3))
    ineg = torch.randint(0, 60, (B,), generator=torch.Generator().manual_seed(4))
    sp, sn, loss = E.cosine_margin(q, table, None, ip, ineg, 1.0)
    spd, snd, lossd = ops.cosine_margin(q.to(DEV), table.to(DEV), None, ip.to(DEV), ineg.to(DEV), 1.0)
    assert_close(spd.cpu().numpy(), sp.numpy(), 1e-5, 1e-6, 'pos scores')
    assert_close(snd.cpu().numpy(), sn.numpy(), 1e-5, 1e-6, 'neg scores')
    assert_close(lossd.item(), loss.item(), 1e-6, 1e-6, 'loss')
    gl = torch.tensor([0.7])
    rows, rid = torch.zeros(2 * B, D), torch.zeros(2 * B, dtype=torch.int64)
    dq = E.cosine_margin_bwd(q, table, None, ip, ineg, 1.0, gl, rows, rid)
    rows_d, rid_d = torch.zeros(2 * B, D, device=DEV), torch.zeros(2 * B, dtype=torch.int64, device=DEV)
    dqd = ops.cosine_margin_bwd(q.to(DEV), table.to(DEV), None, ip.to(DEV), ineg.to(DEV), 1.0, gl.to(DEV), rows_d, rid_d)
    assert torch.equal(rid_d.cpu(), rid)
    assert_close(dqd.cpu().numpy(), dq.numpy(), 1e-4, 1e-7, 'dq')
    assert_close(rows_d.cpu().numpy(), rows.numpy(), 1e-4, 1e-5, 'row grads')


def test_cosine_scores_ragged_and_rank_counts():
    B = 41
    table = rnd(80, D, seed=1, scale=1.0 / D)
    q = rnd(B, D, seed=2)
    lengths = torch.randint(0, 9, (B,), generator=torch.Generator().manual_seed(5))
    offsets = torch.zeros(B + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(lengths, 0)
    ids = torch.randint(0, 80, (int(offsets[-1]),), generator=torch.Generator().manual_seed(6))
    s = E.cosine_scores(q, table, None, ids, offsets)
    sd = ops.cosine_scores(q.to(DEV), table.to(DEV), None, ids.to(DEV), offsets.to(DEV))
    assert_close(sd.cpu().numpy(), s.numpy(), 1e-5, 1e-6, 'ragged scores')
    pos = rnd(B, seed=7, scale=0.1)
    neg = sd.cpu()
    neg[::5] = pos[E._owner(offsets, neg.numel(), B)][::5]  # exact ties
    lt, le = E.rank_counts_ragged(pos, neg, offsets)
    ltd, led = ops.rank_counts_ragged(pos.to(DEV), neg.to(DEV), offsets.to(DEV))
    assert torch.equal(ltd.cpu(), lt) and torch.equal(led.cpu(), le), 'rank counts must be bit-exact'
    # backward
    gs = rnd(B + ids.numel(), seed=8)
    dq = torch.zeros(B, D)
    rows, rid = torch.zeros(ids.numel(), D), torch.zeros(ids.numel(), dtype=torch.int64)
    E.cosine_scores_bwd(q, table, None, ids, offsets, gs, B, dq, False, rows, rid)
    dqd = torch.zeros(B, D, device=DEV)
    rows_d = torch.zeros(ids.numel(), D, device=DEV)
    rid_d = torch.zeros(ids.numel(), dtype=torch.int64, device=DEV)
    ops.cosine_scores_bwd(q.to(DEV), table.to(DEV), None, ids.to(DEV), offsets.to(DEV), gs.to(DEV), B, dqd, False,
                          rows_d, rid_d)
    assert torch.equal(rid_d.cpu(), rid)
    assert_close(dqd.cpu().numpy(), dq.numpy(), 1e-4, 1e-6, 'dq')
    assert_close(rows_d.cpu().numpy(), rows.numpy(), 1e-4, 1e-5, 'rows')


@pytest.mark.parametrize('qt', ['1-chain', '2-chain', '3-chain', '2-inter', '3-inter', '3-inter_chain',
                                '3-chain_inter'])
def test_query_graph_layout_bit_exact(qt):
    from mpqe_b200.data_utils import template_of
    t = template_of(qt)
    rel = [11, 3, 7][:t.num_edges]
    for B in (1, 5, 513):
        ei, et, b = E.build_query_graph(t.num_nodes, t.src, t.dst, rel, B, None)
        eid, etd, bd = ops.build_query_graph(t.num_nodes, t.src, t.dst, rel, B, torch.device(DEV))
        assert torch.equal(eid.cpu(), ei) and torch.equal(etd.cpu(), et) and torch.equal(bd.cpu(), b)


@pytest.mark.parametrize('n,R', [(1, 3), (100, 7), (5000, 78), (300000, 300), (70000, 70000)])
def test_relation_sort_bit_exact(n, R):
    et = torch.randint(0, R, (n,), generator=torch.Generator().manual_seed(n))
    perm, off = E.relation_sort(et, R)
    pd, od = ops.relation_sort(et.to(DEV), R)
    assert torch.equal(pd.cpu(), perm), 'stable permutation must be bit-exact'
    assert torch.equal(od.cpu(), off)


@pytest.mark.parametrize('count,table_rows', [(1, 10), (64, 5), (3000, 400), (100000, 372584), (50000, 3)])
def test_sparse_rows_combine(count, table_rows):
    ids = torch.randint(0, table_rows, (count,), generator=torch.Generator().manual_seed(count))
    rows = rnd(count, D, seed=9)
    uid, urows, num = E.sparse_rows_combine(ids, rows, table_rows)
    outs = []
    for _ in range(2):
        uidd, urowsd, numd = ops.sparse_rows_combine(ids.to(DEV), rows.to(DEV), table_rows)
        outs.append(urowsd.cpu())
    k = int(num)
    assert int(numd) == k
    assert torch.equal(uidd.cpu()[:k], uid[:k]), 'unique ids (sorted) must be bit-exact'
    assert torch.equal(uidd.cpu()[k:], torch.zeros(count - k, dtype=torch.int64))
    assert torch.equal(outs[0][k:], torch.zeros(count - k, D))
    assert_close(outs[0][:k].numpy(), urows[:k].numpy(), 1e-5, 1e-4 * max(1.0, count / table_rows / 10), 'summed rows')
    assert torch.equal(outs[0], outs[1]), 'combine must be bit-reproducible'
    dense = torch.zeros(table_rows, D, device=DEV)
    ops.scatter_rows(uidd, urowsd, numd, dense)
    ref = torch.zeros(table_rows, D).index_add(0, ids, rows)
    assert_close(dense.cpu().numpy(), ref.numpy(), 1e-5, 1e-4 * max(1.0, count / table_rows / 10), 'dense scatter')


@pytest.mark.parametrize('tc', [False, True])
@pytest.mark.parametrize('B,N', [(5, 100), (130, 1000), (64, 20000), (300, 777)])
def test_rank_counts_table_sandwich(B, N, tc):
    need_tc(tc)
    table = rnd(N, D, seed=1, scale=1.0 / D)
    q = rnd(B, D, seed=2)
    tgt = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(3))
    qd, td = q.to(DEV), table.to(DEV)
    pos = ops.cosine_scores(qd, td, None, tgt.to(DEV))
    left = torch.zeros(B, dtype=torch.int64, device=DEV)
    right = torch.zeros(B, dtype=torch.int64, device=DEV)
    half = N // 3
    ops.rank_counts_table(qd, pos, td, 0, half, left, right, use_tensor_cores=tc)      # two shards chained = one table
    ops.rank_counts_table(qd, pos, td, half, N, left, right, use_tensor_cores=tc)
    # float64 scores on the CPU bracket the fp32 ones: counts must lie between the counts at pos -/+ tol
    y = table.double() / table.double().norm(dim=1, keepdim=True)
    s = (q.double() / q.double().norm(dim=1, keepdim=True)) @ y.t()
    p = pos.cpu().double().unsqueeze(1)
    tol = 4e-6 if tc else 2e-6
    lo_lt, hi_lt = (s < p - tol).sum(1), (s < p + tol).sum(1)
    lo_le, hi_le = (s <= p - tol).sum(1), (s <= p + tol).sum(1)
    l, r = left.cpu(), right.cpu()
    assert bool(((l >= lo_lt) & (l <= hi_lt)).all()) and bool(((r >= lo_le) & (r <= hi_le)).all())
    assert bool((r >= l).all()) and bool((r <= N).all())


def test_adam_matches_torch():
    p = rnd(1000, seed=1)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=0.01)
    pd = p.to(DEV)
    m, v = torch.zeros_like(pd), torch.zeros_like(pd)
    for step in range(1, 4):
        g = rnd(1000, seed=10 + step)
        ref.grad = g.clone()
        opt.step()
        ops.adam_dense(pd, g.to(DEV), m, v, 0.01, 0.9, 0.999, 1e-8, step)
    assert_close(pd.cpu().numpy(), ref.detach().numpy(), 1e-5, 1e-6, 'adam')


def test_sparse_rows_recombine_after_gather():
    """Two ranks' padded combines, concatenated and combined again (what the data-parallel exchange does), equal the
    combine of all pairs; the sentinel padding (id = table_rows) is dropped and does not form a long segment."""
    table_rows = 1000
    parts = []
    all_ids, all_rows = [], []
    for r in range(2):
        ids = torch.randint(0, table_rows, (5000,), generator=torch.Generator().manual_seed(r))
        rows = rnd(5000, D, seed=20 + r)
        all_ids.append(ids)
        all_rows.append(rows)
        u, ur, k = ops.sparse_rows_combine(ids.to(DEV), rows.to(DEV), table_rows, pad_id=table_rows)
        assert bool((u[int(k):] == table_rows).all()) and bool((ur[int(k):] == 0).all())
        parts.append((u, ur))
    u2, ur2, k2 = ops.sparse_rows_combine(torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), table_rows,
                                          pad_id=table_rows)
    want_u, want_r, want_k = E.sparse_rows_combine(torch.cat(all_ids), torch.cat(all_rows), table_rows)
    k2 = int(k2)
    assert k2 == int(want_k) and torch.equal(u2[:k2].cpu(), want_u[:k2])
    assert_close(ur2[:k2].cpu().numpy(), want_r[:k2].numpy(), 1e-5, 1e-4, 'recombined rows')


@pytest.mark.parametrize('count,table_rows', [(7, 10), (3000, 400), (100000, 372584)])
def test_sparse_rows_plan_apply_equals_combine(count, table_rows):
    """The two-phase combine (id-only plan on a second stream, then the row summation) is bit-identical to the
    one-call combine; the ids come from the ids-only mode of the multi-item gather."""
    ids = torch.randint(0, table_rows, (count,), generator=torch.Generator().manual_seed(count))
    rows = rnd(count, D, seed=4).to(DEV)
    table = torch.zeros(table_rows, D, device=DEV)
    ids_d = ids.to(DEV)
    rid = torch.empty(count, dtype=torch.int64, device=DEV)
    half = count // 2
    ops.gather_multi([ops.GatherItem(table, None, ids_d, half, rows_id=rid, id_offset=0),
                      ops.GatherItem(table, None, ids_d, count - half, ids_offset=half, rows_id=rid, rows_offset=half)],
                     backward='ids')
    assert torch.equal(rid.cpu(), ids), 'ids-only gather must reproduce the row ids'
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan = ops.SparseRowsPlan(rid, table_rows)
    torch.cuda.current_stream().wait_stream(side)
    u1, r1, k1 = plan.apply(rows, pad_id=table_rows)
    u2, r2, k2 = ops.sparse_rows_combine(ids_d, rows, table_rows, pad_id=table_rows)
    assert int(k1) == int(k2)
    assert torch.equal(u1, u2) and torch.equal(r1, r2)


@pytest.mark.parametrize('B', [100, 4096])
def test_layer_forward_packed_weights(B):
    """tcgen05 path with pre-split weight tiles (bulk-copy staging) == column-gather staging == CPU."""
    need_tc(True)
    n = 3
    x = rnd(B, n, D, seed=1)
    w = rnd(4, D, D, seed=2, scale=0.05)
    bias = rnd(D, seed=3)
    out = torch.zeros(B, n, D)
    terms = [ops.Term(x, n, 0, w[0], 2), ops.Term(x, n, 1, w[1], 2), ops.Term(x, n, 2, w[3], 2),
             ops.Term(x, n, 0, w[3], 0), ops.Term(x, n, 1, w[2], 1)]
    E.layer_forward([ops.Group(B, terms, n, out, n, epilogue=ops.EPI_RELU, bias=bias)])
    xd, wd, bd = x.to(DEV), w.to(DEV), bias.to(DEV)
    wp = ops.pack_weights(wd)
    outs = []
    for packed in (False, True):
        od = torch.zeros(B, n, D, device=DEV)
        # mixed: some terms packed, some not, in the same launch
        td = [ops.Term(xd, n, 0, wd[0], 2, wp[0] if packed else None), ops.Term(xd, n, 1, wd[1], 2),
              ops.Term(xd, n, 2, wd[3], 2, wp[3] if packed else None), ops.Term(xd, n, 0, wd[3], 0, wp[3] if packed else None),
              ops.Term(xd, n, 1, wd[2], 1, wp[2] if packed else None)]
        ops.layer_forward([ops.Group(B, td, n, od, n, epilogue=ops.EPI_RELU, bias=bd)], use_tensor_cores=True)
        outs.append(od.cpu())
        assert_close(outs[-1].numpy(), out.numpy(), 1e-5, 2e-5, 'packed=%s' % packed)
    assert torch.equal(outs[0], outs[1]), 'packed and unpacked staging must give identical bits'


def test_matrix_sum_multi():
    """dst (+)= sum of [D, D] matrices in item order; long summand lists continue with accumulating launches."""
    mats = rnd(40, D, D, seed=5)
    md = mats.to(DEV)
    dst = rnd(3, D, D, seed=6)
    dd = dst.to(DEV)
    items_cpu = [(dst[0], [mats[i] for i in (0, 3, 3, 7)], False), (dst[1], [mats[i] for i in range(40)], True),
                 (dst[2], [mats[9]], True)]
    items_gpu = [(dd[0], [md[i] for i in (0, 3, 3, 7)], False), (dd[1], [md[i] for i in range(40)], True),
                 (dd[2], [md[9]], True)]
    E.matrix_sum_multi(items_cpu)
    ops.matrix_sum_multi(items_gpu)
    assert_close(dd.cpu().numpy(), dst.numpy(), 1e-6, 1e-5, 'matrix sums')
    with pytest.raises(Exception):
        ops.matrix_sum_multi([(dd[0], [md[0]], False), (dd[0], [md[1]], False)])   # shared destination in one call


def test_cosine_margin_multi_fused_equals_forward_then_backward():
    """Mode 'both' (margin backward inside the forward pass) gives the same bits as the two separate launches."""
    B, N = 333, 500
    table = rnd(N, D, seed=1, scale=1.0 / D).to(DEV)
    gl = torch.tensor([0.3, 1.7], device=DEV)
    res = []
    for mode in ('split', 'both'):
        items, outs = [], []
        for j in range(2):
            q = rnd(B, D, seed=10 + j).to(DEV)
            ip = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(20 + j)).to(DEV)
            ineg = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(30 + j)).to(DEV)
            hinge, loss = torch.empty(B, device=DEV), torch.empty(1, device=DEV)
            dq = torch.empty(B, D, device=DEV)
            rows, rid = torch.empty(2 * B, D, device=DEV), torch.empty(2 * B, dtype=torch.int64, device=DEV)
            items.append(ops.MarginItem(q, table, None, ip, ineg, hinge=hinge, loss=loss, grad_loss=gl[j:j + 1], dq=dq,
                                        rows_out=rows, rows_id=rid, id_offset=7 * j))
            outs.append((loss, dq, rows, rid))
        if mode == 'split':
            ops.cosine_margin_multi(items, 1.0)
            ops.cosine_margin_multi(items, 1.0, backward=True)
        else:
            ops.cosine_margin_multi(items, 1.0, backward='both')
        res.append([[t.clone() for t in o] for o in outs])
    for a, b in zip(res[0], res[1]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_colsum_multi_shared_destinations():
    """Items sharing a destination are folded in item order (bit-reproducible); single-item destinations too."""
    srcs = [rnd(r, D, seed=40 + i) for i, r in enumerate((1, 130, 4096, 7, 1000, 257))]
    dst = rnd(3, D, seed=50)
    dd = dst.to(DEV)
    want = dst.clone()
    plan = [(0, 0, 1.0), (1, 1, 2.0), (2, 0, 1.0), (3, 2, 0.5), (4, 0, 1.0), (5, 1, 1.0)]
    items = []
    for i, d, scale in plan:
        want[d] += scale * srcs[i].sum(0)
        sd = srcs[i].to(DEV)
        items.append(ops.ColsumItem(sd, sd.shape[0], D, dd[d], scale))
    outs = []
    for _ in range(2):
        cur = dst.to(DEV)
        its = [ops.ColsumItem(it.src, it.rows, it.stride, cur[plan[k][1]], it.scale) for k, it in enumerate(items)]
        ops.colsum_multi(its, torch.device(DEV))
        outs.append(cur.cpu())
    assert_close(outs[0].numpy(), want.numpy(), 1e-5, 1e-4, 'column sums')
    assert torch.equal(outs[0], outs[1])


def test_basis_decomposition_kernels():
    """W = att @ basis, d basis = att^T @ dW, d att = dW . basis (RGCNConv num_bases > 0, reference model.py:281-284)
    against float64 (plain fp32 sums of 5 / 38 / 16384 products in a fixed order; tolerances per result below)."""
    att, basis, dw = rnd(38, 5, seed=1), rnd(5, D, D, seed=2), rnd(38, D, D, seed=3)
    with torch.cuda.device(0):
        w = ops.small_k_matmul(att.to(DEV), basis.to(DEV))
        dbasis = ops.small_k_matmul(att.to(DEV), dw.to(DEV), transpose_a=True)
        datt = ops.rows_dot(dw.to(DEV), basis.to(DEV))
    want_w = (att.double() @ basis.double().view(5, -1)).view(38, D, D)
    want_db = (att.double().t() @ dw.double().view(38, -1)).view(5, D, D)
    want_da = dw.double().view(38, -1) @ basis.double().view(5, -1).t()
    assert w.shape == (38, D, D) and dbasis.shape == (5, D, D) and datt.shape == (38, 5)
    assert_close(w.cpu().numpy(), want_w.numpy(), 1e-6, 1e-5, 'att @ basis')
    assert_close(dbasis.cpu().numpy(), want_db.numpy(), 1e-6, 1e-5, 'att^T @ dW')
    assert_close(datt.cpu().numpy(), want_da.numpy(), 1e-5, 1e-3, 'dW . basis')       # sums of 16384 products of O(1)


def test_rgcn_conv_basis_decomposition_on_device():
    """RGCNConv with num_bases > 0 on the GPU: output and the gradients of att / basis / root / bias / x against the
    oracle's rgcn_conv."""
    from mpqe_b200.data_utils import QueryGraphBatch, template_of
    from mpqe_b200.model import RGCNConv
    from oracle import mpqe_oracle as O
    torch.manual_seed(1)
    conv = RGCNConv(128, 128, 7, 3)
    t = template_of('3-inter_chain')
    g = QueryGraphBatch(t, [6, 2, 0], 37)
    x = torch.randn(37 * t.num_nodes, 128)
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in conv.named_parameters()}
    xc = x.clone().requires_grad_(True)
    gd = g.to(DEV)       # (the edge list is materialised by a kernel)
    want = O.rgcn_conv(xc, gd.edge_index.cpu(), gd.edge_type.cpu(), pc['basis'], pc['root'], pc['bias'], att=pc['att'])
    wgt = torch.randn_like(want)
    (want * wgt).sum().backward()
    conv = conv.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    out = conv(xd, gd)
    assert_close(out.detach().cpu().numpy(), want.detach().numpy(), 1e-5, 1e-5, 'conv out (bases)')
    (out * wgt.to(DEV)).sum().backward()
    assert_close(xd.grad.cpu().numpy(), xc.grad.numpy(), 1e-4, 1e-5, 'dx')
    for k, prm in conv.named_parameters():
        assert_close(prm.grad.cpu().numpy(), pc[k].grad.numpy(), 1e-4, 1e-4, 'd' + k)


@pytest.mark.parametrize('P,N', [(1, 1), (300, 5000), (4097, 2049)])
def test_auc_counts_match_sklearn(P, N):
    """Exact integer pair counts -> the AUC of sklearn.metrics.roc_auc_score on the same scores (ties, NaN and inf
    included: the reference applies nan_to_num first, utils.py:34-36)."""
    from sklearn.metrics import roc_auc_score
    from mpqe_b200 import utils
    g = np.random.RandomState(P + N)
    pos = np.round(g.randn(P).astype(np.float32), 2)        # rounding makes ties
    neg = np.round(g.randn(N).astype(np.float32) - 0.3, 2)
    if P > 10:
        pos[3], pos[4], neg[5], neg[6] = np.nan, np.inf, np.nan, -np.inf
    counts = ops.auc_counts(torch.from_numpy(pos).to(DEV), torch.from_numpy(neg).to(DEV)).cpu().numpy()
    p64, n64 = np.nan_to_num(pos).astype(np.float64), np.nan_to_num(neg).astype(np.float64)
    lt = int((n64[None, :] < p64[:, None]).sum())
    eq = int((n64[None, :] == p64[:, None]).sum())
    assert counts.tolist() == [lt, eq]
    got = utils.auc_from_device_scores(torch.from_numpy(pos).to(DEV), torch.from_numpy(neg).to(DEV))
    if lt + eq not in (0, P * N) or P * N > 1:
        labels = np.concatenate([np.ones(P), np.zeros(N)])
        want = roc_auc_score(labels, np.nan_to_num(np.concatenate([pos, neg])))
        assert abs(got - want) < 1e-12


def test_pack_weights_transposed_flag_equals_packing_a_transposed_copy():
    """The tile image of M^T packed straight from M is bit-identical to the image packed from a transposed copy."""
    m = rnd(5, D, D, seed=9).to(DEV)
    with torch.cuda.device(0):
        mt = ops.transpose(m)
        want = ops.pack_weights([mt[i] for i in range(5)] + [m[i] for i in range(5)])
        got = ops.pack_weights([m[i] for i in range(5)] * 2, [1] * 5 + [0] * 5)
    assert torch.equal(got, want)

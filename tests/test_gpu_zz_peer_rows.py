"""The data-parallel form of the row-gradient combine on ONE GPU: several local buffers stand in for the ranks'
peer-mapped row buffers (the kernel only sees device addresses).  Collected last on purpose."""
import pytest
import torch

from mpqe_b200 import ops

pytestmark = pytest.mark.gpu
D = ops.D
DEV = 'cuda:0'


@pytest.mark.parametrize('world,per_rank,table_rows', [(2, 1000, 300), (3, 3000, 5000), (8, 517, 100000)])
def test_apply_peers_equals_apply_on_the_concatenation(world, per_rank, table_rows):
    """Gather + sum in one kernel over `world` separate buffers == the plain apply over their concatenation, bit for
    bit (same ascending (rank, pair) summation order, same scale)."""
    gen = torch.Generator().manual_seed(world * 7 + per_rank)
    ids = torch.randint(0, table_rows, (world * per_rank,), generator=gen)
    rows = torch.randn(world * per_rank, D, generator=gen)
    bufs = [rows[r * per_rank:(r + 1) * per_rank].contiguous().to(DEV) for r in range(world)]
    plan = ops.SparseRowsPlan(ids.to(DEV), table_rows)
    u1, r1, k1 = plan.apply(torch.cat(bufs), pad_id=table_rows, scale=0.25)
    u2, r2, k2 = plan.apply_peers([b.data_ptr() for b in bufs], per_rank, pad_id=table_rows, scale=0.25)
    torch.cuda.synchronize()
    k = int(k1)
    assert k == int(k2) == int(torch.unique(ids).numel())
    assert torch.equal(u1, u2)
    assert torch.equal(r1, r2)
    want = torch.zeros(table_rows, D).index_add(0, ids, rows) * 0.25
    got = torch.zeros(table_rows, D, device=DEV)
    got[u2[:k]] = r2[:k]
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5 * max(1.0, world * per_rank / table_rows))


@pytest.mark.parametrize('world', [2, 4, 8])
def test_owner_plan_partitions_the_concatenated_combine(world):
    """Every "rank" plans only the ids it owns (read in place from all ranks' id buffers) and sums their rows out of
    the ranks' row buffers; the partitions are disjoint, each inside its owner's row ranges, and their union equals
    the plain combine of the concatenation (rows bit for bit: same ascending (rank, pair) summation order)."""
    gen = torch.Generator().manual_seed(world)
    tables = [(0, 1000), (1000, 37), (1037, 50001)]      # (first global row, rows)
    total = 51038
    per_rank = 4096
    ids = torch.randint(0, total, (world, per_rank), generator=gen)
    ids[0, :5] = torch.tensor([0, 999, 1000, 1036, 51037])          # range edges
    rows = torch.randn(world, per_rank, D, generator=gen)
    id_bufs = [ids[r].contiguous().to(DEV) for r in range(world)]
    row_bufs = [rows[r].contiguous().to(DEV) for r in range(world)]
    full = ops.SparseRowsPlan(torch.cat(id_bufs), total)
    u, rr, k = full.apply(torch.cat(row_bufs), pad_id=total, scale=0.5)
    k = int(k)
    got_ids, got_rows = [], []
    for rank in range(world):
        plan = ops.owner_plan([b.data_ptr() for b in id_bufs], rank, per_rank, [t[0] for t in tables],
                              [t[1] for t in tables], total, torch.device(DEV))
        ui, ri, ki = plan.apply_peers([b.data_ptr() for b in row_bufs], per_rank, pad_id=total, scale=0.5)
        ki = int(ki)
        inside = torch.zeros(ki, dtype=torch.bool, device=DEV)
        for begin, n in tables:
            chunk = (n + world - 1) // world
            lo, hi = begin + min(rank * chunk, n), begin + min((rank + 1) * chunk, n)
            inside |= (ui[:ki] >= lo) & (ui[:ki] < hi)
        assert bool(inside.all())
        assert bool((ui[ki:] == total).all()) and float(ri[ki:].abs().max() if ki < ui.numel() else 0.0) == 0.0
        got_ids.append(ui[:ki])
        got_rows.append(ri[:ki])
    gi, gr = torch.cat(got_ids), torch.cat(got_rows)
    order = torch.argsort(gi)
    assert torch.equal(gi[order], u[:k]), 'union of the owners\' partitions == unique ids of the concatenation'
    assert torch.equal(gr[order], rr[:k])


def test_allreduce_peers_and_barrier():
    """One-shot all-reduce over four buffers (rank order, scaled) and the flag barrier run by four "ranks" on four
    streams of one device (all four kernels must be resident at once for the barrier to complete)."""
    world, n = 4, 4 * 3001
    gen = torch.Generator().manual_seed(0)
    bufs = [torch.randn(n, generator=gen).to(DEV) for _ in range(world)]
    out = torch.empty(n, device=DEV)
    ops.allreduce_peers([b.data_ptr() for b in bufs], n, 0.25, out)
    want = ((bufs[0] + bufs[1]) + bufs[2] + bufs[3]) * 0.25
    assert torch.equal(out, want)
    # two shots (every "rank" reduces its slice in place, then the slices are gathered) give the same bits
    two = [b.clone() for b in bufs]
    for r in range(world):
        ops.reduce_scatter_peers([b.data_ptr() for b in two], r, n, 0.25)
    out2 = torch.empty(n, device=DEV)
    ops.all_gather_peers([b.data_ptr() for b in two], n, out2)
    assert torch.equal(out2, want)
    flags = [torch.zeros(16, dtype=torch.int32, device=DEV) for _ in range(world)]
    epochs = [torch.zeros(1, dtype=torch.int32, device=DEV) for _ in range(world)]
    data = [torch.zeros(1024, device=DEV) for _ in range(world)]
    seen = [torch.zeros(1024, device=DEV) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for rounds in range(3):
        for r, st in enumerate(streams):
            with torch.cuda.stream(st):
                if r == 0 and rounds == 1:
                    torch.cuda._sleep(2000000)          # a late rank: the others must wait for it
                data[r].fill_(float(rounds + 1))        # written before the barrier ...
                ops.peer_barrier([f.data_ptr() for f in flags], r, epochs[r])
                seen[r].copy_(data[(r + 1) % world])    # ... and read by another "rank" after it
        torch.cuda.synchronize()
        assert [int(e) for e in epochs] == [rounds + 1] * world
        for r in range(world):
            assert float(seen[r].min()) == float(seen[r].max()) == float(rounds + 1)

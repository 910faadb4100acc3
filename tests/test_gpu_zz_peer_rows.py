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

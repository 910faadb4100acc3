"""Multi-GPU parity check, launched with torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_multi_gpu_check.py

(1) data-parallel training step: every rank's synchronised gradients == the single-GPU gradients of the concatenated
    batch (dense bucket within fp32 tolerance, touched-row id set bit-exact, identical bits on all ranks);
(2) entity-sharded full-rank eval: merged (count_lt, count_le) == the unsharded counts, bit-exact.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mpqe_b200 import eval as mp_eval, synthetic  # noqa: E402
from mpqe_b200.graph import Formula, Query  # noqa: E402
from mpqe_b200.train_step import HostBatch, TrainStep  # noqa: E402
from oracle import mpqe_oracle as O  # noqa: E402
from tests.model_utils import build_model  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    kg = synthetic.make_kg('aifb', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    model = build_model(kg.raw(), cfg, params, dev, sparse_grad=True)
    frng = np.random.RandomState(0)
    formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in synthetic.QUERY_TYPES]
    B = 300
    per_rank = [[synthetic.sample_id_batch(kg, f, B, np.random.RandomState(100 * r + i)) for i, f in enumerate(formulas)]
                for r in range(world)]

    def host(ids_list):
        return [HostBatch(f, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n))
                for f, (a, t, n) in zip(formulas, ids_list)]

    # ---- (1) data-parallel step vs single-GPU step on the concatenated batch
    ts = TrainStep(model)
    res = ts.forward_backward([ts.to_device(hb) for hb in host(per_rank[rank])])
    flat = res.dense.flat.clone()
    uid, urows, num = [x.clone() for x in res.sparse]
    single = TrainStep(model)
    single.world = 1
    cat = [tuple(np.concatenate([per_rank[r][i][k] for r in range(world)]) for k in range(3)) for i in range(len(formulas))]
    ref = single.forward_backward([single.to_device(hb) for hb in host(cat)])
    scale = float(ref.dense.flat.abs().max())
    err = float((flat - ref.dense.flat).abs().max())
    assert err <= 1e-4 * scale, ('dense gradients differ', err, scale)
    k, kr = int(num), int(ref.sparse[2])
    nz = urows[:k].abs().sum(1) > 0
    nzr = ref.sparse[1][:kr].abs().sum(1) > 0
    assert torch.equal(uid[:k][nz], ref.sparse[0][:kr][nzr]), 'touched-row id sets differ'
    rerr = float((urows[:k][nz] - ref.sparse[1][:kr][nzr]).abs().max())
    assert rerr <= 1e-4 * float(ref.sparse[1][:kr].abs().max()), ('row gradients differ', rerr)
    # identical bits on every rank
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered), 'ranks hold different dense gradients'
    rows_g = [torch.empty_like(urows) for _ in range(world)]
    dist.all_gather(rows_g, urows)
    assert all(torch.equal(g, rows_g[0]) for g in rows_g), 'ranks hold different row gradients'

    # ---- (2) sharded full-rank eval vs unsharded
    qsets = synthetic.make_query_sets(kg, queries_per_formula=200, formulas_per_type=1, seed=2, query_types=('3-inter',))
    queries = [Query.deserialize(r) for r in qsets['3-inter'][0][1]]
    f = queries[0].formula
    tg = [q.target_node for q in queries]
    left, right, pos, n = mp_eval.full_rank_counts(model, f, queries, tg)            # sharded over the process group
    dist.barrier()
    initialized = dist.is_initialized
    dist.is_initialized = lambda: False                                               # unsharded on this rank alone
    try:
        l1, r1, _, n1 = mp_eval.full_rank_counts(model, f, queries, tg)
    finally:
        dist.is_initialized = initialized
    assert n == n1 and torch.equal(left, l1) and torch.equal(right, r1), 'sharded rank counts differ from unsharded'
    m = mp_eval.ranking_metrics(left, right, n)
    if rank == 0:
        print('multi-gpu check ok: world=%d dense err %.2e (scale %.2e), %d touched rows, full-rank MRR %.4f APR %.2f'
              % (world, err, scale, int(nz.sum()), m['MRR'], m['APR']), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from mpqe_b200 import data_utils, encoders, model as M, ops, synthetic
from mpqe_b200.graph import Formula
from mpqe_b200.train_step import HostBatch, TrainStep
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
kg = synthetic.make_kg('am', seed=0)
rng0 = np.random.RandomState(0)
formulas = [Formula(qt, kg.sample_formula(qt, rng0)) for qt in synthetic.QUERY_TYPES]
rels, adj, node_maps = kg.raw()
graph, fm, id2row = data_utils.build_graph(rels, adj, node_maps, 128)
enc = encoders.DirectEncoder(graph.features, fm, sparse_grad=True)
model = M.RGCNEncoderDecoder(graph, enc, readout='sum', scatter_op='add', dropout=0, weight_decay=0.0, num_layers=2, shared_layers=False, adaptive=False).to(dev)
ts = TrainStep(model)
rng = np.random.RandomState(1000 + rank)
host = [HostBatch(f, *[torch.from_numpy(x) for x in synthetic.sample_id_batch(kg, f, 4096, rng)]) for f in formulas]
res = ts.capture(host)
def timeit(name, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    if rank == 0: print('%-28s %.3f ms' % (name, (time.perf_counter() - t0) / n * 1e3), flush=True)
uid, urows = res.sparse[0], res.sparse[1]
cap = uid.numel()
if rank == 0: print('cap rows', cap, 'flat MB', res.dense.flat.numel() * 4 / 1e6, 'rows MB', urows.numel() * 4 / 1e6)
all_ids = torch.empty(world * cap, dtype=torch.int64, device=dev)
all_rows = torch.empty(world * cap, 128, dtype=torch.float32, device=dev)
timeit('graph replay only', lambda: ts._graph.replay())
timeit('allreduce flat', lambda: dist.all_reduce(res.dense.flat))
timeit('allgather ids', lambda: dist.all_gather_into_tensor(all_ids, uid))
timeit('allgather rows', lambda: dist.all_gather_into_tensor(all_rows, urows))
timeit('torch.empty rows', lambda: torch.empty(world * cap, 128, dtype=torch.float32, device=dev))
timeit('combine gathered', lambda: ops.sparse_rows_combine(all_ids, all_rows, ts.total_rows))
def plan_apply():
    plan = ops.SparseRowsPlan(all_ids, ts.total_rows)
    return plan.apply(all_rows, pad_id=ts.total_rows, scale=0.5)
timeit('plan+apply same stream', plan_apply)
timeit('plan only', lambda: ops.SparseRowsPlan(all_ids, ts.total_rows))
timeit('full sync()', lambda: ts.sync(res.dense, res.sparse))
timeit('replay()+sync', lambda: ts.replay())
dist.destroy_process_group()

"""Benchmark of the MPQE query-encoding hot path (BASELINE.json metric: train query-graphs/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape am|mutag|aifb] [--batch B]

One STEP = margin-loss forward + backward (all parameter gradients, row-sparse entity-table gradients, and for
N > 1 the data-parallel gradient exchange) over 7 batches of B queries, one per query type, on a synthetic
AM-shaped graph (the configuration BASELINE.json's scaling target is quoted on).  Per-GPU work is fixed (weak
scaling).  Prints ONE JSON line (see the task contract): `value` is device-timed with ids resident in HBM, `e2e`
is the same step driven from pinned host id buffers with the H2D copies and the D2H loss read inside the timed
region, `roofline` is the fused layer kernel's achieved algorithmic bandwidth against the measured HBM peak, and
`cpu_baseline` is the CPU oracle port (oracle/mpqe_oracle.py) timed on this box's host cores on a bounded sample.

`--impl reference` times only that CPU port (the reference is pure Python and /root/reference does not travel to
the GPU box; the port follows it operator by operator and is pinned to it by tests/golden).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'train query-graphs/s (fwd+bwd)'
UNIT = 'query-graphs/s'
D = 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--shape', default='am', choices=['am', 'mutag', 'aifb', 'tiny'])
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--readout', default='sum')
    ap.add_argument('--cpu-batch', type=int, default=256, help='queries per type in the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--tensor-cores', type=int, default=-1, help='-1: library default, 0: fp32 FFMA, 1: tcgen05')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of one CUDA graph')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def workload_config(args, kg):
    return {'workload': '%s-shaped synthetic KG (%d entities, %d modes, R=%d), MPQE-%s 2-layer RGCN d=128; step = 7 '
                        'query types x %d queries fwd+bwd per GPU' % (
                            args.shape.upper(), kg.num_entities, len(kg.modes), len(kg.typed_relations), args.readout,
                            args.batch),
            'queries_per_step_per_gpu': 7 * args.batch, 'batch_per_type': args.batch, 'embed_dim': D,
            'num_layers': 2, 'readout': args.readout, 'parallelism': 'dp%d' % args.gpus,
            'l2': 'flushed between timed steps (256 MiB memset outside the per-step events)',
            'launch': 'eager' if args.no_graph else 'one CUDA graph per step (+ eager NCCL exchange when N > 1)'}


def make_formulas(kg, seed=0):
    from mpqe_b200 import synthetic
    from mpqe_b200.graph import Formula
    rng = np.random.RandomState(seed)
    return [Formula(qt, kg.sample_formula(qt, rng)) for qt in synthetic.QUERY_TYPES]


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline and the reference arm)
# ---------------------------------------------------------------------------------------------------------------
def oracle_setup(kg, params, readout):
    from oracle import mpqe_oracle as O
    rels, _, node_maps = kg.raw()
    mode_ids, rel_ids = O.schema_ids(rels)
    return O, O.Config(readout=readout, num_layers=2), mode_ids, rel_ids, O.id_to_row(node_maps)


def oracle_step(O, cfg, params, mode_ids, rel_ids, id2row, formulas, id_batches):
    """One bounded CPU step: margin_loss fwd+bwd for each of the 7 formula batches (dense grads, as the reference)."""
    p = {k: v.detach().requires_grad_(True) for k, v in params.items()}
    for f, (anchors, targets, negs) in zip(formulas, id_batches):
        spec = O.formula_spec(f.query_type, f.rels)
        a_ids, var_ids, ei, et, batch = O.query_graph(spec, anchors, rel_ids, mode_ids)
        loss = O.margin_loss(p, cfg, spec, a_ids, var_ids, ei, et, batch, id2row, torch.from_numpy(targets),
                             torch.from_numpy(negs))
        loss.backward()
    return float(loss.detach())


def time_oracle(kg, params, formulas, args, steps, warmup):
    from mpqe_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, cfg, mode_ids, rel_ids, id2row = oracle_setup(kg, params, args.readout)
    rng = np.random.RandomState(123)
    batches = [synthetic.sample_id_batch(kg, f, args.cpu_batch, rng) for f in formulas]
    for _ in range(warmup):
        oracle_step(O, cfg, params, mode_ids, rel_ids, id2row, formulas, batches)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(O, cfg, params, mode_ids, rel_ids, id2row, formulas, batches)
    dt = time.perf_counter() - t0
    qps = 7 * args.cpu_batch * steps / dt
    return qps, dt / steps, cores, ('%d step(s) of 7 query types x %d queries, oracle port of mpqe.model margin_loss '
                                    'fwd+bwd, %d torch threads' % (steps, args.cpu_batch, cores))


def reference_arm(args, kg, formulas, params, rank):
    if rank != 0:
        return
    qps, sec, cores, sample = time_oracle(kg, params, formulas, args, args.steps, args.warmup)
    cfg = workload_config(args, kg)
    cfg['reference_sample'] = sample
    line = {'impl': 'reference', 'metric': METRIC, 'value': qps, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': qps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': qps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.FIELDS,
                                       '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    from mpqe_b200 import synthetic
    kg = synthetic.make_kg(args.shape, seed=0)
    formulas = make_formulas(kg)

    if args.impl == 'reference':
        if rank != 0:
            return
        from oracle import mpqe_oracle as O
        rels, _, node_maps = kg.raw()
        params = O.init_params(rels, node_maps, O.Config(readout=args.readout, num_layers=2), d=D, seed=0)
        reference_arm(args, kg, formulas, params, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    import __graft_entry__
    __graft_entry__.build()
    from mpqe_b200 import data_utils, encoders, model as M, ops
    from mpqe_b200.train_step import HostBatch, TrainStep

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
    if args.tensor_cores >= 0:
        ops.set_tensor_cores(bool(args.tensor_cores))

    torch.manual_seed(0)
    rels, adj, node_maps = kg.raw()
    graph, feature_modules, id2row = data_utils.build_graph(rels, adj, node_maps, D)
    enc = encoders.DirectEncoder(graph.features, feature_modules, sparse_grad=True)
    model = M.RGCNEncoderDecoder(graph, enc, readout=args.readout, scatter_op='add', dropout=0, weight_decay=0.0,
                                 num_layers=2, shared_layers=False, adaptive=False).to(dev)
    ts = TrainStep(model)

    rng = np.random.RandomState(1000 + rank)  # every rank draws its own queries (data parallel)
    host = []
    for f in formulas:
        a, t, n = synthetic.sample_id_batch(kg, f, args.batch, rng)
        host.append(HostBatch(f, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n)))
    resident = [ts.to_device(hb) for hb in host]
    units = 7 * args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)   # samples clocks / throttle reasons from warm-up to the end of the e2e loop
    # ---- warm-up (and CUDA-graph capture of the local step: forward, backward, row-gradient combine) ---------------
    for _ in range(max(args.warmup, 3)):
        ts.forward_backward(resident)
    launches_per_step = None
    if not args.no_graph:
        l0 = ops.launch_count
        ts.capture(host)
        host = ts.staging()      # the step's ids in ONE pinned buffer (what a loader fills in place): one H2D copy
        launches_per_step = (ops.launch_count - l0) // 3      # capture() runs the step 2x eagerly + 1x captured
        for _ in range(3):
            ts.replay()
    barrier()

    def one_step():
        return ts.replay() if not args.no_graph else ts.forward_backward(resident)

    # ---- timed region: K steps, device time per step, L2 flushed between steps --------------------------------
    launches0 = ops.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()
        s.record()
        res = one_step()
        e.record()
    barrier()
    launches = ops.launch_count - launches0
    if launches_per_step is not None:
        launches += launches_per_step * args.steps            # kernels inside the replayed graphs
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = units * world / (ms_per_step * 1e-3)

    # ---- per-kernel timing for the roofline: the same step launched eagerly with CUDA events around the layer and
    # weight-gradient launches (events cannot be recorded inside a replayed graph) ---------------------------------
    ops.profile = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        ts.forward_backward(resident)
    barrier()
    prof = ops.profile
    ops.profile = None

    # ---- end to end: pinned host ids -> H2D -> step -> D2H losses, wall clock --------------------------------
    for _ in range(3):
        ts.run_host(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, losses_host = ts.run_host(host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_value = units * world * args.steps / float(t.item())
    clocks = sampler.stop()
    h2d = sum(hb.nbytes() for hb in host)

    # ---- roofline of the dominant kernel (the fused layer kernel: forward and input-gradient launches) ----------
    peak, peak_src = peaks()
    roofline = None
    kernels = {}
    if prof:
        for kind in sorted({p[0] for p in prof}):
            sel = [p for p in prof if p[0] == kind]
            ms = sum(s.elapsed_time(e) for _, s, e, _, _ in sel)
            kernels[kind] = {'launches_per_step': len(sel) / args.steps, 'ms_per_step': ms / args.steps,
                             'avg_us': 1e3 * ms / len(sel), 'algorithmic_GBps': sum(p[3] for p in sel) / ms / 1e6,
                             'TFLOPs': sum(p[4] for p in sel) / ms / 1e9}
        lk = kernels.get('layer')
        if lk:
            # DRAM bytes per launch of the same kernel on the same workload, from the committed `ncu --set full`
            # capture (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the step's launches)
            traffic = None
            tpath = os.path.join(ROOT, 'profiles', 'layer_kernel_traffic.json')
            if os.path.isfile(tpath):
                with open(tpath) as f:
                    tinfo = json.load(f)
                traffic = tinfo.get('layer_tc_kernel' if ops.tensor_cores_default() else 'layer_simt_kernel')
            roofline = {'kernel': 'layer_tc_kernel' if ops.tensor_cores_default() else 'layer_simt_kernel',
                        'bound': 'hbm', 'achieved': lk['algorithmic_GBps'], 'peak': peak, 'unit': 'GB/s',
                        'frac': lk['algorithmic_GBps'] / peak, 'traffic': traffic, 'peak_source': peak_src,
                        'avg_launch_us': lk['avg_us'], 'achieved_TFLOPs': lk['TFLOPs'],
                        'share_of_step': lk['ms_per_step'] / ms_per_step}

    # ---- second headline: full-entity ranking eval (queries/s), entity table sharded over the ranks ----------------
    from mpqe_b200 import eval as mp_eval
    from mpqe_b200.graph import Query
    eval_info = None
    try:
        ef = formulas[4]                                       # 3-inter
        erng = np.random.RandomState(7)
        ea, et_, _ = synthetic.sample_id_batch(kg, ef, args.batch, erng)   # same queries on every rank (replicated)
        ea_d = torch.from_numpy(ea)
        et_d = torch.from_numpy(et_).to(dev)
        eq = [None] * args.batch
        t, var_ids, rels_e = data_utils.RGCNQueryDataset.formula_layout(ef, model.rel_ids, model.mode_ids)
        qg = data_utils.QueryGraphBatch(t, rels_e, args.batch)
        var_t = torch.tensor(var_ids, dtype=torch.int64)

        def eval_step():
            return mp_eval.full_rank_counts(model, ef, eq, et_d, anchor_ids=ea_d, var_ids=var_t, q_graphs=qg)

        for _ in range(3):
            l_, r_, _, n_ent = eval_step()
        barrier()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es.record()
        for _ in range(5):
            l_, r_, _, n_ent = eval_step()
        ee.record()
        barrier()
        ems = torch.tensor([es.elapsed_time(ee) / 5], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ems, op=torch.distributed.ReduceOp.MAX)
        m = mp_eval.ranking_metrics(l_, r_, n_ent)
        eval_info = {'metric': 'full-rank eval queries/s', 'value': args.batch / (float(ems.item()) * 1e-3),
                     'unit': 'queries/s', 'ms_per_batch': float(ems.item()), 'queries': args.batch,
                     'candidates_per_query': int(n_ent), 'table_sharded_over': world, 'MRR': m['MRR'], 'APR': m['APR']}
    except Exception as exc:  # the train metric is the contract line; never lose it to the extra one
        eval_info = {'error': repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        qps, sec, cores, sample = time_oracle(kg, params, formulas, args, 2, 1)
        cpu = {'value': qps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': workload_config(args, kg), 'roofline': roofline, 'cpu_baseline': cpu,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': 4 * len(host)},
                'gpu_launches': launches, 'clocks': clocks, 'kernels': kernels,
                'tensor_cores': bool(ops.tensor_cores_default()), 'loss': [float(x) for x in losses_host],
                'eval': eval_info}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()

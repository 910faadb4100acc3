"""Benchmark of the MPQE query-encoding hot path (BASELINE.json metric: train query-graphs/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape am|mutag|aifb] [--batch B]

One STEP = margin-loss forward + backward (all parameter gradients, row-sparse entity-table gradients, and for
N > 1 the data-parallel gradient exchange) over 7 batches of B queries, one per query type, on a synthetic
AM-shaped graph (the configuration BASELINE.json's scaling target is quoted on).  Per-GPU work is fixed (weak
scaling).  Prints ONE JSON line (see the task contract): `value` is device-timed with ids resident in HBM, `e2e`
is the same step driven from pinned host id buffers with the H2D copies and the D2H loss read inside the timed
region, `roofline` is the fused layer kernel's achieved algorithmic bandwidth against the measured HBM peak, and
`cpu_baseline` is the CPU oracle port (oracle/mpqe_oracle.py) timed on this box's host cores on a bounded sample of
the SAME configuration (7 query types x B queries).  `configs` holds the other BASELINE.json configurations
(MPQE-TM on the AIFB shape, MPQE-max / MPQE-concat on the MUTAG shape) measured the same way, `eval` the
full-entity ranking evaluation with the entity table sharded over the ranks, and for N > 1 `dp_check` compares a small
N-rank step with the single-rank step on the concatenated batch before anything is timed (the run fails if it differs).

`--impl reference` times only that CPU port, on the same configuration (the reference is pure Python and
/root/reference does not travel to the GPU box; the port follows it operator by operator and is pinned to it by
tests/golden).
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
# name -> (graph shape, readout, layers, adaptive): BASELINE.json configs 2-4
CONFIGS = {
    'am_sum': ('am', 'sum', 2, False),
    'aifb_tm': ('aifb', 'mp', 3, True),
    'mutag_max': ('mutag', 'max', 2, False),
    'mutag_concat': ('mutag', 'concat', 2, False),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='am_sum', choices=sorted(CONFIGS), help='the headline configuration')
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra-configs', action='store_true', help='skip the `configs` sub-records')
    ap.add_argument('--no-eval', action='store_true')
    ap.add_argument('--eval-queries', type=int, default=16384)
    ap.add_argument('--tensor-cores', type=int, default=-1, help='-1: library default, 0: fp32 FFMA, 1: tcgen05')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of one CUDA graph')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def workload_config(name, batch, kg, gpus, graph=True):
    shape, readout, layers, adaptive = CONFIGS[name]
    return {'workload': '%s-shaped synthetic KG (%d entities, %d modes, R=%d), MPQE-%s%s %d-layer RGCN d=128; step = 7 '
                        'query types x %d queries fwd+bwd per GPU' % (
                            shape.upper(), kg.num_entities, len(kg.modes), len(kg.typed_relations), readout,
                            ' adaptive' if adaptive else '', layers, batch),
            'queries_per_step_per_gpu': 7 * batch, 'batch_per_type': batch, 'embed_dim': D,
            'num_layers': layers, 'readout': readout, 'adaptive': adaptive, 'parallelism': 'dp%d' % gpus,
            'l2': 'flushed between timed steps (256 MiB memset outside the per-step events)',
            'launch': 'one CUDA graph per step, data-parallel exchange included' if graph else 'eager'}


def make_formulas(kg, seed=0):
    from mpqe_b200 import synthetic
    from mpqe_b200.graph import Formula
    rng = np.random.RandomState(seed)
    return [Formula(qt, kg.sample_formula(qt, rng)) for qt in synthetic.QUERY_TYPES]


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline and the reference arm)
# ---------------------------------------------------------------------------------------------------------------
def oracle_step(O, cfg, params, mode_ids, rel_ids, id2row, formulas, id_batches):
    """One CPU step: margin_loss fwd+bwd for each of the 7 formula batches (dense grads, as the reference)."""
    p = {k: v.detach().requires_grad_(True) for k, v in params.items()}
    for f, (anchors, targets, negs) in zip(formulas, id_batches):
        spec = O.formula_spec(f.query_type, f.rels)
        a_ids, var_ids, ei, et, batch = O.query_graph(spec, anchors, rel_ids, mode_ids)
        loss = O.margin_loss(p, cfg, spec, a_ids, var_ids, ei, et, batch, id2row, torch.from_numpy(targets),
                             torch.from_numpy(negs))
        loss.backward()
    return float(loss.detach())


def time_oracle(name, kg, params, formulas, batch, steps, warmup):
    from mpqe_b200 import synthetic
    from oracle import mpqe_oracle as O
    shape, readout, layers, adaptive = CONFIGS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rels, _, node_maps = kg.raw()
    mode_ids, rel_ids = O.schema_ids(rels)
    cfg, id2row = O.Config(readout=readout, num_layers=layers, adaptive=adaptive), O.id_to_row(node_maps)
    rng = np.random.RandomState(123)
    batches = [synthetic.sample_id_batch(kg, f, batch, rng) for f in formulas]
    for _ in range(warmup):
        oracle_step(O, cfg, params, mode_ids, rel_ids, id2row, formulas, batches)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(O, cfg, params, mode_ids, rel_ids, id2row, formulas, batches)
    dt = time.perf_counter() - t0
    qps = 7 * batch * steps / dt
    return qps, dt / steps, cores, ('%d step(s) of the full workload (7 query types x %d queries), oracle port of '
                                    'mpqe.model margin_loss fwd+bwd, %d torch threads' % (steps, batch, cores))


def reference_arm(args, kg, formulas):
    from oracle import mpqe_oracle as O
    shape, readout, layers, adaptive = CONFIGS[args.config]
    rels, _, node_maps = kg.raw()
    params = O.init_params(rels, node_maps, O.Config(readout=readout, num_layers=layers, adaptive=adaptive), d=D, seed=0)
    # the whole run (warm-up + timed steps) is bounded to a few minutes: one step of this configuration takes several
    # seconds on the host cores, so at most 2 warm-up and 8 timed steps are run whatever was asked for
    steps, warmup = min(args.steps, 8), min(args.warmup, 2)
    qps, sec, cores, sample = time_oracle(args.config, kg, params, formulas, args.batch, steps, warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': qps, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.config, args.batch, kg, args.gpus, not args.no_graph),
            'cpu_baseline': {'value': qps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                             'steps_run': steps, 'warmup_run': warmup},
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


class Run(object):
    """One configuration on this rank: model, fused step, host batches."""

    def __init__(self, name, batch, dev, rank, seed_base=1000):
        from mpqe_b200 import data_utils, encoders, model as M, synthetic
        from mpqe_b200.train_step import HostBatch, TrainStep
        shape, readout, layers, adaptive = CONFIGS[name]
        self.name, self.batch = name, batch
        self.kg = synthetic.make_kg(shape, seed=0)
        self.formulas = make_formulas(self.kg)
        torch.manual_seed(0)
        rels, adj, node_maps = self.kg.raw()
        graph, feature_modules, _ = data_utils.build_graph(rels, adj, node_maps, D)
        enc = encoders.DirectEncoder(graph.features, feature_modules, sparse_grad=True)
        self.model = M.RGCNEncoderDecoder(graph, enc, readout=readout, scatter_op='add', dropout=0, weight_decay=0.0,
                                          num_layers=layers, shared_layers=False, adaptive=adaptive).to(dev)
        self.ts = TrainStep(self.model)
        rng = np.random.RandomState(seed_base + rank)      # every rank draws its own queries (data parallel)
        self.host = [HostBatch(f, *[torch.from_numpy(x) for x in synthetic.sample_id_batch(self.kg, f, batch, rng)])
                     for f in self.formulas]
        self.units = 7 * batch


def timed_steps(run, args, dev, world, barrier, flush):
    """Warm-up, graph capture and the timed region of one configuration.  Returns (ms per step (max over ranks),
    launches in the timed region, per-kernel eager timings)."""
    from mpqe_b200 import ops
    ts = run.ts
    resident = [ts.to_device(hb) for hb in run.host]
    for _ in range(max(args.warmup, 3)):
        ts.forward_backward(resident)
    if os.environ.get('MPQE_NCU_RANGE'):      # evidence capture: one eager step inside a profiler range
        torch.cuda.synchronize()              # (ncu --profile-from-start off ...; profiles/capture_r02.sh)
        torch.cuda.cudart().cudaProfilerStart()
        flush.zero_()
        ts.forward_backward(resident)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    launches_per_step = None
    if not args.no_graph:
        l0 = ops.launch_count
        ts.capture(run.host)
        run.host = ts.staging()      # the step's ids in ONE pinned buffer (what a loader fills in place): one H2D copy
        launches_per_step = (ops.launch_count - l0) // 3      # capture() runs the step 2x eagerly + 1x captured
        for _ in range(3):
            ts.replay()
    barrier()

    def one_step():
        return ts.replay() if not args.no_graph else ts.forward_backward(resident)

    launches0 = ops.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()
        s.record()
        one_step()
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

    # per-kernel timing for the roofline: the same step launched eagerly with CUDA events around the layer and
    # weight-gradient launches (events cannot be recorded inside a replayed graph)
    ops.profile = []
    overlap, ts.overlap_wgrad = ts.overlap_wgrad, False      # every kernel timed alone, not next to another stream's
    barrier()
    for _ in range(min(args.steps, 10)):
        flush.zero_()
        ts.forward_backward(resident)
    barrier()
    ts.overlap_wgrad = overlap
    prof, ops.profile = ops.profile, None
    if os.environ.get('MPQE_DP_TRACE'):      # phase times of the eager data-parallel step (diagnostic)
        ts.trace = []
        for _ in range(5):
            ts.forward_backward(resident)
        rep = ts.trace_report()
        ts.trace = None
        sys.stderr.write('[rank %d] phases ms: %s\n' % (int(os.environ.get('RANK', '0')), json.dumps({k: round(v, 4) for k, v in rep.items()})))
    kernels = {}
    steps_prof = min(args.steps, 10)
    for kind in sorted({p[0] for p in prof}):
        sel = [p for p in prof if p[0] == kind]
        ms = sum(s.elapsed_time(e) for _, s, e, _, _ in sel)
        kernels[kind] = {'launches_per_step': len(sel) / steps_prof, 'ms_per_step': ms / steps_prof,
                         'avg_us': 1e3 * ms / len(sel), 'algorithmic_GBps': sum(p[3] for p in sel) / ms / 1e6,
                         'TFLOPs': sum(p[4] for p in sel) / ms / 1e9}
    return ms_per_step, launches, launches_per_step, kernels


def layer_roofline(kernels, ms_per_step):
    from mpqe_b200 import ops
    peak, peak_src = peaks()
    lk = kernels.get('layer')
    if not lk:
        return None
    kname = 'layer_simt_kernel'
    if ops.tensor_cores_default():
        kname = 'layer_tc_kernel' if os.environ.get('MPQE_LAYER_KERNEL') == '1' else 'layer_tc2_kernel'
    # DRAM bytes per launch of the same kernel on the same workload, from the committed `ncu --set full` capture
    # (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the step's launches)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'layer_kernel_traffic.json')
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(kname)
    return {'kernel': kname, 'bound': 'hbm', 'achieved': lk['algorithmic_GBps'], 'peak': peak, 'unit': 'GB/s',
            'frac': lk['algorithmic_GBps'] / peak, 'traffic': traffic, 'peak_source': peak_src,
            'avg_launch_us': lk['avg_us'], 'achieved_TFLOPs': lk['TFLOPs'],
            'share_of_step': lk['ms_per_step'] / ms_per_step}


# ---------------------------------------------------------------------------------------------------------------
def dp_check(run, dev, rank, world):
    """Before anything is timed (N > 1): one small data-parallel step (300 queries per type and rank) against the
    single-rank step on the concatenated batch, on the real exchange path (peer memory, barriers, owner combine), and
    the sharded full-entity rank counts against the unsharded ones."""
    from mpqe_b200 import eval as mp_eval, synthetic
    from mpqe_b200.train_step import HostBatch, TrainStep
    dist = torch.distributed
    B = 300
    per_rank = []
    for r in range(world):
        rng = np.random.RandomState(5000 + r)
        per_rank.append([synthetic.sample_id_batch(run.kg, f, B, rng) for f in run.formulas])

    def host(ids_per_formula):
        return [HostBatch(f, *[torch.from_numpy(x) for x in ids]) for f, ids in zip(run.formulas, ids_per_formula)]

    ts_n = TrainStep(run.model)
    res = ts_n.forward_backward([ts_n.to_device(hb) for hb in host(per_rank[rank])])
    flat_n = res.dense.flat.clone()
    uid, urows, num = res.sparse
    k = int(num)
    # identical bits on every rank: compare an integer checksum of the reduced dense bucket
    chk = flat_n.view(torch.int32).to(torch.int64).sum().reshape(1)
    chks = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(chks, chk)
    same_bits = all(int(c) == int(chks[0]) for c in chks)
    # the owners' partitions, gathered (padded to the common capacity)
    capt = torch.tensor([uid.numel()], dtype=torch.int64, device=dev)
    dist.all_reduce(capt, op=dist.ReduceOp.MAX)          # the owners' capacities differ by a few rows
    cap = int(capt)
    ids_pad = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    ids_pad[:k] = uid[:k]
    rows_pad = torch.zeros(cap, D, device=dev)
    rows_pad[:k] = urows[:k]
    all_ids = torch.empty(world * cap, dtype=torch.int64, device=dev)
    all_rows = torch.empty(world * cap, D, device=dev)
    dist.all_gather_into_tensor(all_ids, ids_pad)
    dist.all_gather_into_tensor(all_rows, rows_pad)
    keep = all_ids >= 0
    got_ids, got_rows = all_ids[keep], all_rows[keep]
    order = torch.argsort(got_ids)
    got_ids, got_rows = got_ids[order], got_rows[order]
    # single-rank step on the concatenation of all ranks' batches (mean over N*B == average of the ranks' means)
    cat = [tuple(np.concatenate([per_rank[r][i][j] for r in range(world)]) for j in range(3))
           for i in range(len(run.formulas))]
    ts_1 = TrainStep(run.model, data_parallel=False)
    res1 = ts_1.forward_backward([ts_1.to_device(hb) for hb in host(cat)])
    flat_1 = res1.dense.flat
    u1, r1, n1 = res1.sparse
    k1 = int(n1)
    dense_err = float((flat_n - flat_1).abs().max() / flat_1.abs().max().clamp_min(1e-30))
    ids_equal = got_ids.numel() == k1 and bool(torch.equal(got_ids, u1[:k1]))
    rows_err = float((got_rows - r1[:k1]).abs().max() / r1[:k1].abs().max().clamp_min(1e-30)) if ids_equal else None
    # sharded vs unsharded rank counts (same queries on every rank)
    from mpqe_b200 import data_utils
    ef = run.formulas[4]
    ea, et_, _ = synthetic.sample_id_batch(run.kg, ef, 512, np.random.RandomState(9))
    t, var_ids, rels_e = data_utils.RGCNQueryDataset.formula_layout(ef, run.model.rel_ids, run.model.mode_ids)
    args_ = dict(anchor_ids=torch.from_numpy(ea), var_ids=torch.tensor(var_ids, dtype=torch.int64),
                 q_graphs=data_utils.QueryGraphBatch(t, rels_e, 512))
    ts_n.gather_tables()
    sharded = mp_eval.RankIndex(run.model).counts(ef, [None] * 512, torch.from_numpy(et_).to(dev), **args_)
    single = mp_eval.RankIndex(run.model, distributed=False).counts(ef, [None] * 512, torch.from_numpy(et_).to(dev),
                                                                    **args_)
    counts_equal = bool(torch.equal(sharded[0], single[0]) and torch.equal(sharded[1], single[1]))
    tol = 2e-5
    ok = bool(same_bits and ids_equal and counts_equal and dense_err <= tol and rows_err is not None and rows_err <= tol)
    out = {'ok': ok, 'dense_err': dense_err, 'rows_equal': bool(ids_equal and rows_err is not None and rows_err <= tol),
           'rows_err': rows_err, 'touched_ids_equal': ids_equal, 'identical_bits_across_ranks': same_bits,
           'counts_equal': counts_equal, 'queries_per_type_per_rank': B, 'tolerance': tol,
           'exchange': 'peer memory' if ts_n.peers is not None else 'torch.distributed fallback'}
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out['ok'] = bool(int(flags))
    return out


def eval_bench(run, args, dev, world, barrier):
    """Second headline: full-entity ranking eval (queries/s): every rank encodes the (replicated) queries and ranks them
    against the rows of the target mode's table it owns; integer counts merged with one all-reduce."""
    from mpqe_b200 import data_utils, eval as mp_eval, synthetic
    Q = args.eval_queries
    ef = run.formulas[4]                                       # 3-inter
    ea, et_, _ = synthetic.sample_id_batch(run.kg, ef, Q, np.random.RandomState(7))   # same queries on every rank
    ea_d = torch.from_numpy(ea)
    et_d = torch.from_numpy(et_).to(dev)
    t, var_ids, rels_e = data_utils.RGCNQueryDataset.formula_layout(ef, run.model.rel_ids, run.model.mode_ids)
    qg = data_utils.QueryGraphBatch(t, rels_e, Q)
    var_t = torch.tensor(var_ids, dtype=torch.int64)
    index = mp_eval.RankIndex(run.model)
    eq = [None] * Q

    if os.environ.get('MPQE_NCU_RANGE'):
        index.counts(ef, eq, et_d, anchor_ids=ea_d, var_ids=var_t, q_graphs=qg)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        index.counts(ef, eq, et_d, anchor_ids=ea_d, var_ids=var_t, q_graphs=qg)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    if args.no_graph:
        def step():
            return index.counts(ef, eq, et_d, anchor_ids=ea_d, var_ids=var_t, q_graphs=qg)
    else:      # the batch as one CUDA graph (collectives included); ids stay in its device buffers
        step = mp_eval.GraphedCounts(index, ef, ea_d, et_d, var_t, qg)

    for _ in range(3):
        l_, r_, _, n_ent = step()
    barrier()
    es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es.record()
    for _ in range(5):
        l_, r_, _, n_ent = step()
    ee.record()
    barrier()
    ems = torch.tensor([es.elapsed_time(ee) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(ems, op=torch.distributed.ReduceOp.MAX)
    m = mp_eval.ranking_metrics(l_, r_, n_ent)
    ms = float(ems.item())
    flops = 2.0 * Q * n_ent * D
    return {'metric': 'full-rank eval queries/s', 'value': Q / (ms * 1e-3), 'unit': 'queries/s', 'ms_per_batch': ms,
            'queries': Q, 'candidates_per_query': int(n_ent), 'table_sharded_over': world, 'MRR': m['MRR'],
            'APR': m['APR'], 'score_gemm': {'shape': '[%d,128] x [128,%d]' % (Q, n_ent), 'flops': flops,
                                            'TFLOPs_fp32_equivalent': flops / (ms * 1e-3) / 1e12,
                                            'note': 'whole batch: encode + positive scores + ranking + count merge'}}


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        if rank != 0:
            return
        from mpqe_b200 import synthetic
        kg = synthetic.make_kg(CONFIGS[args.config][0], seed=0)
        reference_arm(args, kg, make_formulas(kg))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    import __graft_entry__
    __graft_entry__.build()
    from mpqe_b200 import ops

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
    if args.tensor_cores >= 0:
        ops.set_tensor_cores(bool(args.tensor_cores))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    run = Run(args.config, args.batch, dev, rank)

    check = None
    if world > 1:
        check = dp_check(run, dev, rank, world)
        if not check['ok']:
            if rank == 0:
                print(json.dumps({'metric': METRIC, 'error': 'dp_check failed', 'dp_check': check}), flush=True)
            torch.distributed.destroy_process_group()
            raise SystemExit(3)

    sampler = ClockSampler(local_rank)   # samples clocks / throttle reasons from warm-up to the end of the e2e loop
    ms_per_step, launches, launches_per_step, kernels = timed_steps(run, args, dev, world, barrier, flush)
    value = run.units * world / (ms_per_step * 1e-3)

    # ---- end to end: pinned host ids -> H2D -> step -> D2H losses, wall clock --------------------------------
    ts, host = run.ts, run.host
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
    e2e_value = run.units * world * args.steps / float(t.item())
    clocks = sampler.stop()
    h2d = sum(hb.nbytes() for hb in host)
    roofline = layer_roofline(kernels, ms_per_step)

    eval_info = None
    if not args.no_eval:
        try:
            if world > 1:
                run.ts.gather_tables()
            eval_info = eval_bench(run, args, dev, world, barrier)
        except Exception as exc:  # the train metric is the contract line; never lose it to the extra one
            eval_info = {'error': repr(exc)}

    # ---- the other BASELINE.json configurations, measured the same way ----------------------------------------------
    extra = {}
    if not args.no_extra_configs:
        for name in sorted(CONFIGS):
            if name == args.config:
                continue
            try:
                r2 = Run(name, args.batch, dev, rank)
                ms2, _, lps2, k2 = timed_steps(r2, args, dev, world, barrier, flush)
                rf = layer_roofline(k2, ms2)
                extra[name] = {'value': r2.units * world / (ms2 * 1e-3), 'unit': UNIT, 'ms_per_step': ms2,
                               'launches_per_step': lps2, 'layer_frac_of_hbm_peak': rf['frac'] if rf else None,
                               'layer_avg_us': rf['avg_launch_us'] if rf else None,
                               'config': workload_config(name, args.batch, r2.kg, world, not args.no_graph)}
                del r2
            except Exception as exc:
                extra[name] = {'error': repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        params = {k: v.detach().cpu() for k, v in run.model.state_dict().items()}
        qps, sec, cores, sample = time_oracle(args.config, run.kg, params, run.formulas, args.batch, 4, 1)     # ~10-15 s of host work on the GPU boxes
        cpu = {'value': qps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': workload_config(args.config, args.batch, run.kg, world, not args.no_graph),
                'roofline': roofline, 'cpu_baseline': cpu,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': 4 * len(host)},
                'gpu_launches': launches, 'launches_per_step': launches_per_step, 'clocks': clocks, 'kernels': kernels,
                'tensor_cores': bool(ops.tensor_cores_default()), 'loss': [float(x) for x in losses_host],
                'eval': eval_info, 'configs': extra, 'dp_check': check}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()

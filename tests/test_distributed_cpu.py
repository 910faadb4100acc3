"""Data-parallel gradient exchange on CPU: 2 gloo ranks (kernels emulated) must reproduce the 1-rank gradients of
the concatenated batch: dense bucket within fp32 tolerance and identical bits on both ranks; the row-sparse entity
gradient arrives partitioned by row OWNER -- each rank holds exactly the rows it owns, the partitions are disjoint and
their union is the single-process result (touched-row id set bit-exact)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(rank_seed, B):
    from mpqe_b200 import synthetic
    from mpqe_b200.graph import Formula
    from mpqe_b200.train_step import HostBatch, TrainStep
    from oracle import mpqe_oracle as O
    from tests.model_utils import build_model
    kg = synthetic.make_kg('tiny', seed=5)
    rels, _, node_maps = kg.raw()
    cfg = O.Config(readout='sum', num_layers=2)
    params = O.init_params(rels, node_maps, cfg, d=128, seed=1)
    model = build_model(kg.raw(), cfg, params, 'cpu', sparse_grad=True)
    frng = np.random.RandomState(0)
    formulas = [Formula(qt, kg.sample_formula(qt, frng)) for qt in ('2-chain', '3-inter', '3-inter_chain')]
    rng = np.random.RandomState(rank_seed)
    ids = [synthetic.sample_id_batch(kg, f, B, rng) for f in formulas]
    return model, formulas, ids, TrainStep, HostBatch


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import contextlib
    from tests import emulator
    from mpqe_b200 import ops

    class MP(object):
        def setattr(self, obj, name, value):
            setattr(obj, name, value)
    emulator.install(MP())
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    model, formulas, ids, TrainStep, HostBatch = _build(100 + rank, 12)
    ts = TrainStep(model)
    host = [HostBatch(f, torch.from_numpy(a), torch.from_numpy(t), torch.from_numpy(n)) for f, (a, t, n) in
            zip(formulas, ids)]
    res = ts.forward_backward([ts.to_device(hb) for hb in host])
    u, r, k = res.sparse
    sparse = {'all': (u[:int(k)].clone(), r[:int(k)].clone())}
    owned = [(ts.table_offsets[mode] + lo, ts.table_offsets[mode] + hi) for mode, lo, hi in ts.owned_rows()]
    torch.save({'flat': res.dense.flat.clone(), 'sparse': sparse, 'losses': res.losses.clone(), 'owned': owned},
               out % rank)
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_rank(tmp_path, monkeypatch):
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / 'rank%d.pt')
    mp.start_processes(_worker, args=(2, port, out), nprocs=2, join=True, start_method='spawn')
    r0, r1 = torch.load(out % 0), torch.load(out % 1)
    assert torch.equal(r0['flat'], r1['flat']), 'all ranks must hold identical dense gradients'
    # every rank holds exactly the rows it owns
    for r in (r0, r1):
        ids = r['sparse']['all'][0]
        inside = torch.zeros_like(ids, dtype=torch.bool)
        for lo, hi in r['owned']:
            inside |= (ids >= lo) & (ids < hi)
        assert bool(inside.all()) and ids.numel() > 0
    # the union of the partitions, in id order (ownership ranges interleave over the tables)
    ids_all = torch.cat([r0['sparse']['all'][0], r1['sparse']['all'][0]])
    rows_all = torch.cat([r0['sparse']['all'][1], r1['sparse']['all'][1]])
    assert ids_all.unique().numel() == ids_all.numel(), 'the owners' + "'" + ' partitions must be disjoint'
    order = torch.argsort(ids_all)
    r0['sparse']['all'] = (ids_all[order], rows_all[order])

    # single process on the concatenated batch (mean over 2B == average of the two ranks' means)
    from tests import emulator
    emulator.install(monkeypatch)
    m0, formulas, ids0, TrainStep, HostBatch = _build(100, 12)
    _, _, ids1, _, _ = _build(101, 12)
    ts = TrainStep(m0)
    host = [HostBatch(f, torch.from_numpy(np.concatenate([a0, a1])), torch.from_numpy(np.concatenate([t0, t1])),
                      torch.from_numpy(np.concatenate([n0, n1])))
            for f, (a0, t0, n0), (a1, t1, n1) in zip(formulas, ids0, ids1)]
    res = ts.forward_backward([ts.to_device(hb) for hb in host])
    np.testing.assert_allclose(r0['flat'].numpy(), res.dense.flat.numpy(), rtol=1e-4,
                               atol=1e-6 * float(res.dense.flat.abs().max()))
    for m, (u, r, k) in {'all': res.sparse}.items():
        k = int(k)
        ids_dp, rows_dp = r0['sparse'][m]
        nz = rows_dp.abs().sum(1) > 0
        want_nz = r[:k].abs().sum(1) > 0
        assert torch.equal(ids_dp[nz], u[:k][want_nz]), 'touched-row id set must be bit-exact'
        np.testing.assert_allclose(rows_dp[nz].numpy(), r[:k][want_nz].numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(0.5 * (r0['losses'] + r1['losses']).numpy(), res.losses.numpy(), rtol=1e-5)

"""Full-entity ranking evaluation with the entity table sharded across GPUs (north star; new, not in the reference,
which ranks against <= 1000 stored negatives only: mpqe/utils.py:72-95).

Every rank encodes the (replicated) queries, scores them against ITS row range of the target mode's table with
`mpqe_rank_counts_table` (the [B, N] score matrix is never materialised) and the integer (count_lt, count_le) pairs are
merged with one all-reduce -- bit-exact for any number of ranks.  From the merged counts:
    APR  = scipy 'rank' percentile of the positive among all entities of its mode (utils.percentile_from_counts)
    MRR  = mean of 1 / (1 + #candidates scoring strictly higher than the positive) = 1 / (1 + N - count_le)
"""
import numpy as np
import torch

from . import ops
from .model import Weights
from .utils import percentile_from_counts


def shard_rows(num_rows, rank, world):
    """Row range [begin, end) of rank `rank` (contiguous, balanced)."""
    base, rem = divmod(num_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@torch.no_grad()
def full_rank_counts(model, formula, queries, target_nodes, anchor_ids=None, var_ids=None, q_graphs=None,
                     process_group=None, use_tensor_cores=None):
    """(count_lt, count_le, positive scores, N): counts of entities of the target mode scoring below / not above each
    query's positive.  The spare last table row (data_utils.py:31) is not a candidate."""
    dist = torch.distributed
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(process_group) if distributed else 0
    world = dist.get_world_size(process_group) if distributed else 1
    job = model.make_job(formula, queries, anchor_ids, var_ids, q_graphs)
    device = job.anchor_ids.device
    table = model.enc.table(formula.target_mode)
    num_entities = table.shape[0] - 1
    with ops.device_guard(device):
        model._engine.encode([job], Weights(model, False))
        tgt = model.enc.ids_on_device(target_nodes, device).reshape(-1)
        pos = ops.cosine_scores(job.q, table, model.enc.node_maps, tgt)
        left = torch.zeros(job.B, dtype=torch.int64, device=device)
        right = torch.zeros(job.B, dtype=torch.int64, device=device)
        begin, end = shard_rows(num_entities, rank, world)
        tc = ops.tensor_cores_default() if use_tensor_cores is None else bool(use_tensor_cores)
        ops.rank_counts_table(job.q, pos, table, begin, end, left, right, use_tensor_cores=tc)
        if world > 1:
            both = torch.stack((left, right))
            dist.all_reduce(both, group=process_group)     # integer sum: exact, order independent
            left, right = both[0], both[1]
    return left, right, pos, num_entities


def ranking_metrics(left, right, num_entities):
    left = left.cpu().numpy()
    right = right.cpu().numpy()
    n = np.full(left.shape, num_entities, dtype=np.float64)
    apr = percentile_from_counts(left, right, n)
    rr = 1.0 / (1.0 + (num_entities - right))
    return {'APR': float(apr.mean()), 'MRR': float(rr.mean()), 'queries': int(left.size)}


@torch.no_grad()
def eval_full_rank(model, test_queries, batch_size=4096, process_group=None, use_tensor_cores=None):
    """{formula: [queries]} -> overall APR / MRR against all entities of each target mode."""
    lefts, rights, ns = [], [], []
    for formula, formula_queries in test_queries.items():
        for off in range(0, len(formula_queries), batch_size):
            batch = formula_queries[off:off + batch_size]
            l, r, _, n = full_rank_counts(model, formula, batch, [q.target_node for q in batch],
                                          process_group=process_group, use_tensor_cores=use_tensor_cores)
            lefts.append(l.cpu().numpy())
            rights.append(r.cpu().numpy())
            ns.append(np.full(len(batch), n, dtype=np.float64))
    left, right, n = np.concatenate(lefts), np.concatenate(rights), np.concatenate(ns)
    apr = percentile_from_counts(left, right, n)
    rr = 1.0 / (1.0 + (n - right))
    return {'APR': float(apr.mean()), 'MRR': float(rr.mean()), 'queries': int(left.size)}

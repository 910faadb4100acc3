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


def shard_rows(num_rows, rank, world, table_rows=None):
    """Row range [begin, end) of rank `rank`: the rows of the table the rank OWNS in data-parallel training (ranges of
    ceil(table_rows / world) rows, see train_step.TrainStep), clipped to the `num_rows` candidate rows -- so that a
    rank always ranks against rows whose latest values it holds."""
    table_rows = num_rows if table_rows is None else table_rows
    chunk = (table_rows + world - 1) // world
    return min(rank * chunk, num_rows), min((rank + 1) * chunk, num_rows)


class RankIndex(object):
    """Everything of a full-entity ranking evaluation that does not depend on the batch, cached across batches: the
    per-step weight layouts of the encoder (`Weights`: transposes, tf32 tile images) and, per target mode, the prepared
    table shard of this rank (1/||row||, pre-split tile images).  Parameters are frozen during an evaluation
    (`torch.no_grad`); entries are rebuilt when a parameter's version counter has moved."""

    def __init__(self, model, process_group=None, use_tensor_cores=None, distributed=None):
        dist = torch.distributed
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        self.model, self.pg, self.distributed = model, process_group, bool(distributed)
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        self.tc = ops.tensor_cores_default() if use_tensor_cores is None else bool(use_tensor_cores)
        self._weights = None
        self._tables = {}

    def _param_version(self):
        return tuple(p._version for p in self.model.parameters())

    def weights(self):
        v = self._param_version()
        if self._weights is None or self._weights[0] != v:
            self._weights = (v, Weights(self.model, False))
        return self._weights[1]

    def table(self, mode):
        table = self.model.enc.table(mode)
        ent = self._tables.get(mode)
        if ent is None or ent.version != table._version or ent.table.data_ptr() != table.data_ptr():
            num_entities = table.shape[0] - 1          # the spare last row (data_utils.py:31) is not a candidate
            begin, end = shard_rows(num_entities, self.rank, self.world, table.shape[0])
            ent = self._tables[mode] = ops.RankTable(table, begin, end, self.tc)
        return ent

    @torch.no_grad()
    def counts(self, formula, queries, target_nodes, anchor_ids=None, var_ids=None, q_graphs=None):
        """(count_lt, count_le, positive scores, N) of one formula batch, merged over the ranks.
        With several ranks the batch is also SPLIT for the encoder: every rank encodes B / world of the queries and
        scores them against their positives, the query embeddings and positive scores are all-gathered (B x 516
        bytes), and every rank then ranks all B queries against the rows of the table it owns."""
        from .data_utils import QueryGraphBatch
        model = self.model
        B = len(queries)
        lo, hi = 0, B
        split = self.world > 1 and B % self.world == 0 and B >= self.world
        if split:
            per = B // self.world
            lo, hi = self.rank * per, (self.rank + 1) * per
            queries = queries[lo:hi]
            if anchor_ids is not None:
                anchor_ids = anchor_ids[lo:hi]
            if isinstance(q_graphs, QueryGraphBatch):
                q_graphs = QueryGraphBatch(q_graphs.template, q_graphs.edge_rel_ids, hi - lo)
            elif q_graphs is not None:
                anchor_ids = var_ids = q_graphs = None      # rebuild the layout of the slice from the queries
        job = model.make_job(formula, queries, anchor_ids, var_ids, q_graphs)
        device = job.anchor_ids.device
        table = model.enc.table(formula.target_mode)
        with ops.device_guard(device):
            W = self.weights()
            W.prepared = None
            model._engine.encode([job], W)
            tgt = model.enc.ids_on_device(target_nodes, device).reshape(-1)[lo:hi].contiguous()
            pos = ops.cosine_scores(job.q, table, model.enc.node_maps, tgt)
            q = job.q
            if split:
                q_all = torch.empty(B, q.shape[1], dtype=q.dtype, device=device)
                pos_all = torch.empty(B, dtype=pos.dtype, device=device)
                torch.distributed.all_gather_into_tensor(q_all, q.contiguous(), group=self.pg)
                torch.distributed.all_gather_into_tensor(pos_all, pos.contiguous(), group=self.pg)
                q, pos = q_all, pos_all
            both = torch.zeros(2, B, dtype=torch.int64, device=device)
            self.table(formula.target_mode).counts(q, pos, both[0], both[1])
            if self.world > 1:
                torch.distributed.all_reduce(both, group=self.pg)     # integer sum: exact, order independent
        return both[0], both[1], pos, table.shape[0] - 1


class GraphedCounts(object):
    """`RankIndex.counts` for batches of one formula and one size as ONE CUDA graph: encoder, positive scores, rank
    counts over this rank's table shard and -- with several ranks -- the all-gather of the embeddings and the integer
    all-reduce of the counts (NCCL, captured).  A ranking batch on 1/N of the table is ~0.25 ms of kernels behind
    ~30 launches and three collectives: issued one by one it is bound by the host.  Every rank must construct and call
    it in step.  `anchor_ids` [B, anchors] / `target_nodes` [B]: the first batch (the graph is warmed up on it);
    later batches are copied into the same device buffers."""

    def __init__(self, index, formula, anchor_ids, target_nodes, var_ids, q_graphs):
        model = index.model
        dev = model.mode_embeddings.weight.device
        self.index, self.formula = index, formula
        self.anchors = anchor_ids.to(device=dev, dtype=torch.int64).contiguous().clone()
        self.targets = model.enc.ids_on_device(target_nodes, dev).reshape(-1).contiguous().clone()
        self._queries = [None] * self.anchors.shape[0]
        self._args = dict(anchor_ids=self.anchors, var_ids=var_ids, q_graphs=q_graphs)
        cur = torch.cuda.current_stream(dev)
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            for _ in range(2):
                index.counts(formula, self._queries, self.targets, **self._args)
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = index.counts(formula, self._queries, self.targets, **self._args)

    def __call__(self, anchor_ids=None, target_nodes=None):
        """(count_lt, count_le, positive scores, N) -- views of the graph's output buffers, valid until the next call."""
        if anchor_ids is not None:
            self.anchors.copy_(anchor_ids, non_blocking=True)
        if target_nodes is not None:
            self.targets.copy_(self.index.model.enc.ids_on_device(target_nodes, self.targets.device).reshape(-1),
                               non_blocking=True)
        self.graph.replay()
        return self.out


@torch.no_grad()
def full_rank_counts(model, formula, queries, target_nodes, anchor_ids=None, var_ids=None, q_graphs=None,
                     process_group=None, use_tensor_cores=None, index=None):
    """(count_lt, count_le, positive scores, N): counts of entities of the target mode scoring below / not above each
    query's positive.  Pass a `RankIndex` as `index` to reuse the prepared weights and table shards across batches."""
    if index is None:
        index = RankIndex(model, process_group, use_tensor_cores)
    return index.counts(formula, queries, target_nodes, anchor_ids, var_ids, q_graphs)


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
    index = RankIndex(model, process_group, use_tensor_cores)
    for formula, formula_queries in test_queries.items():
        for off in range(0, len(formula_queries), batch_size):
            batch = formula_queries[off:off + batch_size]
            l, r, _, n = index.counts(formula, batch, [q.target_node for q in batch])
            lefts.append(l.cpu().numpy())
            rights.append(r.cpu().numpy())
            ns.append(np.full(len(batch), n, dtype=np.float64))
    left, right, n = np.concatenate(lefts), np.concatenate(rights), np.concatenate(ns)
    apr = percentile_from_counts(left, right, n)
    rr = 1.0 / (1.0 + (n - right))
    return {'APR': float(apr.mean()), 'MRR': float(rr.mean()), 'queries': int(left.size)}

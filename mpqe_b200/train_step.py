"""One optimiser step's worth of the hot path without Python autograd in the loop: margin-loss forward + backward
over a list of formula batches, gradient synchronisation across ranks, optional fused Adam.

This is the caller-side row (f)1 of SURVEY.md section 8 (`run_train` / `run_batch_v2`, reference
train_helpers.py:76-120): the reference issues up to 11 `margin_loss` calls per step, one per query type, each
a separate pair of encoder forwards; here all batches of a step go through the same per-pass launches.

Multi-GPU (one process per GPU, torch.distributed for the rendezvous): query batches are data-parallel, dense
parameters are replicated, and the entity tables are OWNED row-range-wise: rank r owns rows [r*ceil(rows/N),
(r+1)*ceil(rows/N)) of every table.  All buffers the ranks exchange live in peer-mapped (symmetric) memory, and the whole
step -- exchange included -- is one stream of kernels (one CUDA graph), with no host-side collective call:
  forward   the gather / margin kernels read every entity row from its owner's copy of the table over NVLink;
  B1        flag barrier: every rank has emitted its row ids and finished the previous step;
  side      the owner plan: all ranks' ids are read in place, the ones this rank owns are sorted and segmented;
  backward  gradient rows and the dense bucket land in this rank's peer-visible buffers;
  B2        flag barrier: all rows and buckets are final;
  exchange  one-shot all-reduce of the dense bucket (rank order: identical bits everywhere) and ONE kernel that sums
            the owned rows straight out of the peers' buffers.  Per GPU and step (N-1)/N of one rank's gradient rows
            cross NVLink once, plus the entity rows its batch reads.
The result of a step on rank r is the dense gradient (replicated) and the combined row gradient OF THE ROWS r OWNS;
the ranks' partitions are disjoint and their union is the single-process result on the concatenated batch.
Without peer-mapped memory (CPU / gloo tests, or when symmetric memory cannot be set up) the same kernels run on
buffers assembled with torch.distributed all_gather / all_reduce.
"""
import os

import numpy as np
import torch

from . import ops
from .model import Job, Weights, loss_backward, loss_forward, plan_rows
from .ops import D


class Batch(object):
    """Device-resident ids of one formula batch."""

    def __init__(self, job, targets, negatives, weight=1.0):
        self.job, self.targets, self.negatives, self.weight = job, targets, negatives, float(weight)


class HostBatch(object):
    """Pinned host ids of one formula batch (what a data loader hands over)."""

    def __init__(self, formula, anchor_ids, targets, negatives, weight=1.0, pin=True):
        self.formula = formula
        pin = pin and torch.cuda.is_available()      # pin=False: the tensors already are views of pinned memory
        self.anchor_ids = anchor_ids.contiguous().pin_memory() if pin else anchor_ids
        self.targets = targets.contiguous().pin_memory() if pin else targets
        self.negatives = negatives.contiguous().pin_memory() if pin else negatives
        self.weight = float(weight)

    def nbytes(self):
        return 8 * (self.anchor_ids.numel() + self.targets.numel() + self.negatives.numel())


class StepResult(object):
    """losses [batches] (device), dense = Grads (flat bucket + views), sparse = (unique ids, rows, count);
    `total` = weighted sum of the losses, computed on demand (two small launches kept out of the step)."""

    def __init__(self, losses, weights, dense, sparse):
        self.losses, self.weights, self.dense, self.sparse = losses, weights, dense, sparse

    @property
    def total(self):
        return (self.losses * self.weights).sum()


class PeerGroup(object):
    """Peer-mapped buffers of one data-parallel group (torch symmetric memory: cuMem allocations exchanged at a
    rendezvous and mapped into every process) and the flag barrier over them."""

    def __init__(self, pg, device):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        self.pg = pg if pg is not None else torch.distributed.group.WORLD
        self.rank = torch.distributed.get_rank(self.pg)
        self.world = torch.distributed.get_world_size(self.pg)
        self.device = device
        self._keep = []
        self.flags, self.flag_ptrs = self.alloc((ops._lib.MAX_PEERS,), torch.int32, zero=True)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        torch.distributed.barrier(group=self.pg)     # every rank's flags are zero before the first device barrier

    def alloc(self, shape, dtype, zero=False):
        """(local tensor, [address of every rank's tensor in this process]).  A collective: never during capture."""
        if torch.cuda.is_current_stream_capturing():
            raise ops._lib.MpqeError('peer-visible buffers must be allocated before graph capture')
        buf = self._symm.empty(*shape, dtype=dtype, device=self.device)
        if zero:
            buf.zero_()
        hdl = self._symm.rendezvous(buf, self.pg)
        self._keep.append((buf, hdl))
        return buf, [int(p) for p in hdl.buffer_ptrs]

    def barrier(self):
        ops.peer_barrier(self.flag_ptrs, self.rank, self.epoch)


class TrainStep(object):
    def __init__(self, model, margin=1.0, process_group=None, average=True, data_parallel=True):
        """data_parallel=False: a single-rank step even inside an initialised process group (reference runs, checks)."""
        self.model = model
        self.margin = float(margin)
        self.pg = process_group
        self._dp = bool(data_parallel)
        self.world = torch.distributed.get_world_size(process_group) if self._dist() else 1
        self.average = average
        self._layouts = {}
        self.adam_state = None
        self._side = None   # second stream for the id-only half of the row-gradient combine
        self._wgrad_stream = None   # third stream: weight gradients next to the input-gradient launches
        self.overlap_wgrad = os.environ.get('MPQE_OVERLAP_WGRAD', '1') != '0'
        self.rank = torch.distributed.get_rank(process_group) if self._dist() else 0
        self.peers = None        # PeerGroup (world > 1, peer-mapped memory available)
        self._xcap = None        # pairs per rank the exchange buffers were set up for
        self._xrows = self._xids = self._xflat = self._dense_out = None
        self.steps = 0
        # all entity tables as one id space: global row = table_offsets[mode] + row
        self.table_offsets, off = {}, 0
        for mode, module in model.enc.feature_modules.items():
            self.table_offsets[mode] = off
            off += module.weight.shape[0]
        self.total_rows = off

    def _dist(self):
        return self._dp and torch.distributed.is_available() and torch.distributed.is_initialized()

    # ---- batch construction ---------------------------------------------------------------------------
    def layout(self, formula):
        key = formula
        lay = self._layouts.get(key)
        if lay is None:
            from .data_utils import RGCNQueryDataset
            m = self.model
            t, var_ids, rels = RGCNQueryDataset.formula_layout(formula, m.rel_ids, m.mode_ids)
            dev = m.mode_embeddings.weight.device
            lay = self._layouts[key] = (t, tuple(rels), tuple(var_ids),
                                        torch.tensor(var_ids, dtype=torch.int64, device=dev), m.num_passes(formula))
        return lay

    def to_device(self, hb, device_ids=None):
        """H2D copy of one host batch (async from pinned memory) -> Batch.  `device_ids` = (anchor ids, targets,
        negatives) already on the device (views of the step's id buffer in graph mode)."""
        dev = self.model.mode_embeddings.weight.device
        t, rels, var_host, var_dev, passes = self.layout(hb.formula)
        if device_ids is None:
            device_ids = tuple(x.to(dev, non_blocking=True) for x in (hb.anchor_ids, hb.targets, hb.negatives))
        job = Job(t, rels, var_dev, hb.formula.anchor_modes, hb.formula.target_mode, device_ids[0], passes)
        job.var_rows_host = var_host
        return Batch(job, device_ids[1], device_ids[2], hb.weight)

    def check_ids(self, hb):
        """Raises IndexError (as the reference's nn.Embedding lookup would, data_utils.py:35) when a host batch holds an
        entity id that is unknown or not of the mode its slot has.  The kernels do not validate ids -- an unknown id
        would index its table at row -1 -- so loaders check the ids they are built from (`capture` does it for its
        batches; a `replay` / `forward_backward` on ids from elsewhere should be preceded by this check)."""
        maps = getattr(self, '_node_maps_host', None)
        if maps is None:
            maps = self._node_maps_host = self.model.enc.node_maps.cpu().numpy()
        rows = {m: mod.weight.shape[0] - 1 for m, mod in self.model.enc.feature_modules.items()}
        f = hb.formula
        cols = [(hb.anchor_ids[:, i], m) for i, m in enumerate(f.anchor_modes)]
        cols += [(hb.targets, f.target_mode), (hb.negatives, f.target_mode)]
        for ids, mode in cols:
            ids = np.asarray(ids).reshape(-1)
            if ids.size == 0:
                continue
            if ids.min() < 0 or ids.max() >= maps.shape[0]:
                raise IndexError('entity id out of range in a %s batch (mode %s)' % (f.query_type, mode))
            r = maps[ids]
            if (r < 0).any() or (r >= rows[mode]).any():
                raise IndexError('entity id without a row in the %s table (batch of %s)' % (mode, f.query_type))

    def refresh(self, batch):
        """A Batch can be re-run: drop the activations of the previous step."""
        j = batch.job
        j.acts = j.outs = j.z = j.u = j.q = j.argmax = j.fwd_groups = None
        return batch

    # ---- the step -------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, batches):
        """Returns StepResult: per-batch losses (device scalars), the flat dense gradient bucket (`.dense.flat`, views
        per parameter in `.dense`; averaged over the ranks) and the row-sparse entity gradient `.sparse = (unique global
        row ids, summed rows, num_unique)` with global row = table_offsets[mode] + row.  With several ranks `.sparse`
        covers the rows THIS rank owns (see the module docstring)."""
        dev = self.model.mode_embeddings.weight.device
        with ops.device_guard(dev):
            return self._local_step(batches)

    def _plan_on_side_stream(self, make_plan, dev, keep=()):
        """make_plan() -> ops.SparseRowsPlan, run on the second stream (joined by `_join_side`).  `keep`: tensors the
        plan kernels read, allocated on the current stream."""
        cur = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            plan = make_plan()
            self._plan_done = torch.cuda.Event()
            self._plan_done.record(self._side)
        for t in keep:
            # read by the sort on the second stream: keep the caching allocator from handing the block to the main
            # stream's next allocation while those kernels are still pending
            t.record_stream(self._side)
        plan.ws.record_stream(cur)
        plan.num.record_stream(cur)
        return plan

    def _weights_on_side_stream(self, jobs, dev):
        """`Weights` + `Engine.prepare` on the second stream; `W.ready_event` orders the first layer launch after it."""
        cur = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            W = Weights(self.model, True, defer_raw=True)
            self.model._engine.prepare(jobs, W)
            W.ready_event = torch.cuda.Event()
            W.ready_event.record(self._side)
            # the transposed copies themselves are read only by the one-row kernels at the end of the backward
            if W.raw_pending:
                W.raw_transposes()
                W.raw_event = torch.cuda.Event()
                W.raw_event.record(self._side)
        return W

    def _join_side(self, dev):
        torch.cuda.current_stream(dev).wait_stream(self._side)

    def _wait_plan(self, dev):
        """The current stream waits for the plan only, not for what was put on the second stream after it."""
        torch.cuda.current_stream(dev).wait_event(self._plan_done)

    def _on_side_stream(self, dev, fn):
        """Runs fn() on the second stream, ordered after everything enqueued on the current stream so far."""
        self._side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._side):
            fn()

    def _on_wgrad_stream(self, dev, fn):
        if self._wgrad_stream is None:
            self._wgrad_stream = torch.cuda.Stream(device=dev)
        self._wgrad_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._wgrad_stream):
            fn()

    def _join_wgrad(self, dev):
        if self._wgrad_stream is not None:
            torch.cuda.current_stream(dev).wait_stream(self._wgrad_stream)

    # ---- data-parallel exchange: buffers ---------------------------------------------------------------
    def _table_ranges(self):
        """[(mode, first global row, rows)] in table_offsets order."""
        m = self.model
        return [(mode, self.table_offsets[mode], m.enc.feature_modules[mode].weight.shape[0]) for mode in
                m.enc.feature_modules]

    def owned_rows(self, rank=None):
        """[(mode, first owned row, one past the last)] of `rank` (default: this rank)."""
        rank = self.rank if rank is None else rank
        out = []
        for mode, _, rows in self._table_ranges():
            chunk = (rows + self.world - 1) // self.world
            out.append((mode, min(rank * chunk, rows), min((rank + 1) * chunk, rows)))
        return out

    def _use_peer_memory(self, dev):
        import os
        return (self.world > 1 and dev.type == 'cuda' and not getattr(self, '_peers_off', False) and
                os.environ.get('MPQE_PEER_ROWS', '1') != '0')

    def setup_exchange(self, batches):
        """Allocates the peer-visible buffers of the data-parallel exchange for steps shaped like `batches` (row
        gradients, their ids, the dense bucket), moves the entity tables into peer-visible memory and registers them as
        owner-read tables.  A collective over the process group (rendezvous): called automatically by the first step,
        and before graph capture.  All ranks must run steps with the same number of (row id, row) pairs -- batches of
        the same formulas and sizes -- which is checked here."""
        m = self.model
        dev = m.mode_embeddings.weight.device
        jobs = [b.job for b in batches]
        cap = sum(m._row_capacities(jobs, [(job.target_mode, 2 * job.B) for job in jobs]).values())
        dist = torch.distributed
        caps = [None] * self.world
        dist.all_gather_object(caps, int(cap), group=self.pg)
        if len(set(caps)) != 1:
            raise ops._lib.MpqeError('data-parallel step: ranks differ in their (row id, row) pair counts %r; every rank '
                                     'must run batches of the same formulas and sizes' % (caps,))
        dense_numel = sum(int(p.numel()) for p in self._dense_shapes())
        self._xcap, self._xdense = cap, dense_numel
        self._dense_out = None
        if not self._use_peer_memory(dev):
            self.peers = None
            self._xrows = self._xids = self._xflat = None
            return
        try:
            if self.peers is None:
                self.peers = PeerGroup(self.pg, dev)
            self._xrows, self._row_ptrs = self.peers.alloc((cap, D), torch.float32)
            self._xids, self._id_ptrs = self.peers.alloc((cap,), torch.int64)
            self._xflat, self._flat_ptrs = self.peers.alloc((dense_numel,), torch.float32)
            self.shard_tables()
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.pg)
        except Exception as exc:      # no peer mapping on this system: torch.distributed collectives from now on
            import warnings
            warnings.warn('mpqe_b200: symmetric (peer-mapped) memory unavailable (%r); the gradient exchange falls '
                          'back to torch.distributed all_gather / all_reduce' % (exc,))
            self._peers_off = True
            self.peers = None
            self._xrows = self._xids = self._xflat = None

    def _dense_shapes(self):
        m = self.model
        ps = []
        for layer in m.distinct_layers():
            ps += [layer.relation_weights(), layer.root, layer.bias]
        ps.append(m.mode_embeddings.weight)
        if isinstance(m.readout, torch.nn.Module):
            ps += [m.readout.layers[0].weight, m.readout.layers[2].weight, m.readout.layers[0].bias,
                   m.readout.layers[2].bias]
        return ps

    def shard_tables(self):
        """Moves every entity table into peer-visible memory and registers it with the row kernels as owner-read: row r
        of a table is read from rank (r // ceil(rows / world))'s copy.  Replicas start identical; after that only the
        owner's range of a copy is authoritative (an optimiser updates the rows it owns), until `gather_tables()`."""
        if getattr(self, '_sharded', False):
            return
        dev = self.model.mode_embeddings.weight.device
        for mode, module in self.model.enc.feature_modules.items():
            w = module.weight
            if w.data_ptr() in ops.PEER_TABLES:      # already owner-read (another TrainStep of the same model)
                continue
            buf, ptrs = self.peers.alloc(tuple(w.shape), torch.float32)
            buf.copy_(w.data)
            w.data = buf
            chunk = (w.shape[0] + self.world - 1) // self.world
            ops.PEER_TABLES[buf.data_ptr()] = (torch.tensor(ptrs, dtype=torch.int64, device=dev), chunk)
        self._sharded = True

    @torch.no_grad()
    def gather_tables(self):
        """Every rank's copy of every entity table receives the owners' rows (before a checkpoint, an export or an
        evaluation that reads tables locally)."""
        if self.world == 1:
            return
        for (mode, _, rows), (_, lo, hi) in zip(self._table_ranges(), self.owned_rows()):
            w = self.model.enc.feature_modules[mode].weight.data
            chunk = (rows + self.world - 1) // self.world
            mine = torch.zeros(chunk, w.shape[1], dtype=w.dtype, device=w.device)
            mine[:hi - lo] = w[lo:hi]
            full = torch.empty(self.world * chunk, w.shape[1], dtype=w.dtype, device=w.device)
            torch.distributed.all_gather_into_tensor(full, mine, group=self.pg)
            w.copy_(full[:rows])

    # ---- the step proper -------------------------------------------------------------------------------
    def _local_step(self, batches):
        """forward + backward + row-gradient combine (+ the cross-rank exchange): everything here is enqueued on the
        current stream and one helper stream, with no host synchronisation -- the part that is graph-captured."""
        m = self.model
        dev = m.mode_embeddings.weight.device
        jobs = [self.refresh(b).job for b in batches]
        tg = [b.targets for b in batches]
        ng = [b.negatives for b in batches]
        mark = self._mark
        mark('start')
        multi = self.world > 1
        if multi and getattr(self, '_xcap', None) is None:
            self.setup_exchange(batches)
        peer = multi and self.peers is not None
        # The row ids of the step's entity gradients depend only on the batch ids: emit them first and run the id-only
        # half of the combine (stable sort + segmentation, ~20 small latency-bound launches) on a second stream, under
        # the forward and backward; only the final row summation waits for the gradient rows.
        # The per-step weight preparation (transposes, tf32 tile images, summed matrices) goes to that stream too, ahead
        # of the sort: it overlaps the input gather; the first layer launch waits for its event.
        # (the ids-only launch stays on this stream: moved in front of the sort it lets the input gather start at once,
        # which then crowds out the small weight-preparation kernels the first layer launch waits for: +9 us per step)
        R = plan_rows(m, jobs, tg, ng, self.table_offsets,
                      rows_buffer=(lambda cap: self._xrows) if peer else None,
                      ids_buffer=(lambda cap: self._xids) if peer else None)
        rows, ids, used = R.shared
        if multi and used != self._xcap:
            raise ops._lib.MpqeError('data-parallel step: %d (row id, row) pairs, the exchange was set up for %d; call '
                                     'setup_exchange(batches) when the batch shapes change' % (used, self._xcap))
        ranges = self._table_ranges()
        if peer:
            mark('ids emitted')
            self.peers.barrier()      # B1: all ranks' ids are in place and every rank is done with the previous step
            mark('B1')
        W = self._weights_on_side_stream(jobs, dev)
        all_ids = None
        if not multi:
            plan = self._plan_on_side_stream(lambda: ops.SparseRowsPlan(ids[:used], self.total_rows), dev, keep=(ids,))
        else:
            if peer:
                id_src = self._id_ptrs
            else:     # ids through torch.distributed; the same owner plan then reads the gathered copy
                all_ids = torch.empty(self.world * used, dtype=torch.int64, device=ids.device)
                torch.distributed.all_gather_into_tensor(all_ids, ids[:used].contiguous(), group=self.pg)
                id_src = [all_ids[r * used:(r + 1) * used] for r in range(self.world)]
            plan = self._plan_on_side_stream(
                lambda: ops.owner_plan(id_src, self.rank, used, [r[1] for r in ranges], [r[2] for r in ranges],
                                       self.total_rows, ids.device), dev, keep=() if all_ids is None else (all_ids,))
        key = tuple(b.weight for b in batches)
        wts = getattr(self, '_wts', None)
        if wts is None or wts[0] != key:
            wts = self._wts = (key, torch.tensor(key, dtype=torch.float32, device=dev))
        # d total / d loss_i = the batch weights, known now: the margin backward rides on the margin forward
        losses, W = loss_forward(m, jobs, tg, ng, self.margin, True, grad_losses=wts[1], W=W)
        mark('forward')
        side = (lambda fn: self._on_wgrad_stream(dev, fn), lambda: self._join_wgrad(dev)) if self.overlap_wgrad else None
        G = loss_backward(m, jobs, W, tg, ng, self.margin, wts[1], self.table_offsets, rows=R,
                          defer_constant=not multi, flat=self._xflat if peer else None, side=side)
        if not multi:
            self._wait_plan(dev)
            # the batch-constant tail of the backward (five small latency-bound launches) runs on another stream
            # under the row summation, which does not depend on it
            self._on_side_stream(dev, G.finish)
            sparse = plan.apply(rows[:used], pad_id=self.total_rows)
            self._join_side(dev)
            self._join_wgrad(dev)
            self._weight_decay(W, G, losses, sum(key))
            return StepResult(losses, wts[1], G, sparse)
        self._join_wgrad(dev)
        self._weight_decay(W, G, losses, sum(key))
        mark('backward')
        scale = 1.0 / self.world if self.average else 1.0
        owned = sum(hi - lo for _, lo, hi in self.owned_rows())     # bound of the distinct rows this rank can receive
        if peer:
            # the owner plan has finished reading the peers' ids before this rank signals B2: a rank that has passed B2
            # may start its next step and overwrite its ids
            self._wait_plan(dev)
            mark('owner plan joined')
            self.peers.barrier()      # B2: every rank's gradient rows and dense bucket are final
            mark('B2')
            if self._dense_out is None:
                self._dense_out = torch.empty_like(self._xflat)
            # the dense bucket is all-reduced on the second stream, next to the row combine on this one (both read over
            # NVLink; the bucket is 1/10 of the rows).  Fewer than four ranks: one shot, every rank sums all buckets
            # (N-1 buckets in, one kernel); else two shots: reduce this rank's slice, B3 (every slice is reduced),
            # gather the slices
            def one_shot():
                ops.allreduce_peers(self._flat_ptrs, self._xdense, scale, self._dense_out)

            def two_shot():
                ops.reduce_scatter_peers(self._flat_ptrs, self.rank, self._xdense, scale)
                self.peers.barrier()
                ops.all_gather_peers(self._flat_ptrs, self._xdense, self._dense_out)
            self._on_side_stream(dev, one_shot if self.world < 4 else two_shot)
            sparse = plan.apply_peers(self._row_ptrs, used, pad_id=self.total_rows, scale=scale, capacity=owned)
            mark('row combine')
            self._join_side(dev)
            mark('exchange done')
            return StepResult(losses, wts[1], G.over(self._dense_out), sparse)
        torch.distributed.all_reduce(G.flat, group=self.pg)
        if scale != 1.0:
            G.flat.mul_(scale)
        all_rows = torch.empty(self.world * used, D, dtype=torch.float32, device=rows.device)
        torch.distributed.all_gather_into_tensor(all_rows, rows[:used].contiguous(), group=self.pg)
        self._join_side(dev)
        sparse = plan.apply_peers([all_rows[r * used:(r + 1) * used] for r in range(self.world)], used,
                                  pad_id=self.total_rows, scale=scale, capacity=owned)
        return StepResult(losses, wts[1], G, sparse)

    def _mark(self, name):
        """Phase tracing of eager steps (`self.trace = []` switches it on): CUDA events on the current stream."""
        tr = getattr(self, 'trace', None)
        if tr is not None and not torch.cuda.is_current_stream_capturing():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            tr.append((name, ev))

    def trace_report(self):
        """{phase: mean ms since the previous mark} over the traced steps."""
        tr, out, cnt = self.trace, {}, {}
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(tr[:-1], tr[1:]):
            if n1 == 'start':
                continue
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        return {k: out[k] / cnt[k] for k in out}

    def _weight_decay(self, W, G, losses, weight_sum):
        """The L2 term of every margin_loss call of the step (reference model.py:487-492: weight_decay * sum of the
        un-squared norms of the readout-MLP parameters, once per formula batch): each loss gets the term, the readout
        gradients get (sum of the batch weights) * weight_decay * p / ||p||.  The transposed copies W.w1t / W.w2t have
        the layout of the gradient buffers and the same norms as the parameters."""
        wd = float(self.model.weight_decay)
        if W.ro is None or wd <= 0:
            return
        ops.l2_reg([W.w1t, W.b1, W.w2t, W.b2], [G.dw1t, G.db1, G.dw2t, G.db2], wd, float(weight_sum), losses=losses)

    # ---- CUDA-graph mode: the whole step (exchange included) becomes one graph launch --------------------------
    @torch.no_grad()
    def capture(self, host_batches):
        """Stages `host_batches` into static device buffers, warms up and captures the step into a CUDA graph.
        Later steps with batches of the same formulas and sizes call `replay(host_batches)`.  With several ranks the
        graph contains the flag barriers and the peer-memory exchange: every rank must capture and replay in step."""
        m = self.model
        dev = m.mode_embeddings.weight.device
        with ops.device_guard(dev):
            # all ids of a step live in ONE device buffer mirrored by ONE pinned host buffer: a step's input is a
            # single H2D copy instead of three small copies per formula batch
            for hb in host_batches:
                self.check_ids(hb)
            total = sum(hb.anchor_ids.numel() + hb.targets.numel() + hb.negatives.numel() for hb in host_batches)
            self._host_ids = torch.empty(total, dtype=torch.int64).pin_memory()
            self._dev_ids = torch.empty(total, dtype=torch.int64, device=dev)
            self._static, self._staging, off = [], [], 0
            for hb in host_batches:
                views = []
                for src in (hb.anchor_ids, hb.targets, hb.negatives):
                    n = src.numel()
                    hv = self._host_ids[off:off + n].view(src.shape)
                    hv.copy_(src)
                    views.append((hv, self._dev_ids[off:off + n].view(src.shape)))
                    off += n
                self._staging.append(HostBatch(hb.formula, views[0][0], views[1][0], views[2][0], hb.weight, pin=False))
                self._static.append(self.to_device(hb, tuple(v[1] for v in views)))
            self._dev_ids.copy_(self._host_ids, non_blocking=True)
            self._wts = None
            if self.world > 1:
                self.setup_exchange(self._static)     # collective: peer-visible buffers exist before the capture
                if self.peers is None:
                    # torch.distributed fallback: the collectives stay eager, the step is not captured
                    self._graph = None
                    self._graph_res = self._local_step(self._static)
                    return self._graph_res
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._local_step(self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            profile, ops.profile = ops.profile, None      # CUDA events cannot be recorded during capture
            self._graph = torch.cuda.CUDAGraph()
            # captured on a high-priority stream: the kernels of the critical chain are scheduled ahead of the ones on
            # the helper streams (sort, weight preparation, weight gradients) when both are pending -- 0.359 -> 0.345
            # ms per bench step, and the same from run to run (MPQE_MAIN_PRIORITY=0: default priority)
            prio = os.environ.get('MPQE_MAIN_PRIORITY', '1') == '1'
            cap_stream = torch.cuda.Stream(device=dev, priority=-1) if prio else None
            with torch.cuda.graph(self._graph, stream=cap_stream):
                self._graph_res = self._local_step(self._static)
            ops.profile = profile
        return self._graph_res

    @torch.no_grad()
    def replay(self, host_batches=None):
        """One step through the captured graph; with `host_batches` their ids are first copied (async, pinned) into
        the static buffers.  Returns the same StepResult object every time (its tensors are overwritten)."""
        if host_batches is not None:
            if host_batches is not self._staging:    # not written in place (see `staging`): pack on the host first
                for st, hb in zip(self._staging, host_batches):
                    st.anchor_ids.copy_(hb.anchor_ids)
                    st.targets.copy_(hb.targets)
                    st.negatives.copy_(hb.negatives)
            self._dev_ids.copy_(self._host_ids, non_blocking=True)
        if self._graph is None:
            with ops.device_guard(self._dev_ids.device):
                self._graph_res = self._local_step(self._static)
        else:
            self._graph.replay()
        return self._graph_res

    def staging(self):
        """Graph mode: HostBatch objects whose id tensors are views of the step's single pinned host buffer.  A data
        loader that writes the next batch into them in place and passes this very list to `replay` / `run_host` gets
        the whole step's input across in one H2D copy, with no packing on the host."""
        return self._staging

    @torch.no_grad()
    def run_host(self, host_batches):
        """End-to-end step from pinned host ids: H2D copies, forward+backward(+sync), D2H of the losses."""
        if getattr(self, '_static', None) is not None:
            res = self.replay(host_batches)
        else:
            res = self.forward_backward([self.to_device(hb) for hb in host_batches])
        return res, res.losses.cpu()

    # ---- fused optimiser (torch.optim.Adam defaults, reference train.py:86-88) ------------------------------
    def _adam_setup(self, lr, betas, eps):
        m = self.model
        for layer in m.distinct_layers():
            if layer.att is not None:
                raise ops._lib.MpqeError('TrainStep.adam_step does not handle a basis decomposition (num_bases > 0); '
                                         'use torch.optim with margin_loss instead')
        dev = m.mode_embeddings.weight.device
        params = [l.basis for l in m.distinct_layers()] + [l.root for l in m.distinct_layers()]
        params += [l.bias for l in m.distinct_layers()] + [m.mode_embeddings.weight]
        if isinstance(m.readout, torch.nn.Module):
            lin1, lin2 = m.readout.layers[0], m.readout.layers[2]
            params += [lin1.weight, lin2.weight, lin1.bias, lin2.bias]
        self.adam_params = [p.data for p in params]
        self.adam_state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in self.adam_params]
        self.adam_hyper = (float(lr), float(betas[0]), float(betas[1]), float(eps))
        self.adam_clock = ops.adam_state(dev)
        self.row_adam = ops.RowAdam([(module.weight.data, self.table_offsets[mode])
                                     for mode, module in m.enc.feature_modules.items()], lr, betas, eps)

    def _dense_grads(self, G):
        grads = list(G.dw) + list(G.droot) + list(G.dbias) + [G.dmode]
        if isinstance(self.model.readout, torch.nn.Module):
            # the readout gradients are kept transposed (the layout the kernels produce): bring them to the parameters'
            grads += [ops.transpose(G.dw1t), ops.transpose(G.dw2t), G.db1, G.db2]
        return grads

    @torch.no_grad()
    def catchup_rows(self, batches=None):
        """Row-sparse Adam bookkeeping BEFORE a step's forward: the rows this step reads receive the zero-gradient
        Adam steps they skipped since they were last touched, so the forward sees exactly what dense Adam would have
        stored (see mpqe_adam_rows_catchup).  batches=None: every row (before evaluation / checkpoint / export)."""
        if self.adam_state is None:
            return
        dev = self.model.mode_embeddings.weight.device
        with ops.device_guard(dev):
            if batches is None:
                self.row_adam.catchup(None, state=self.adam_clock)
                return
            from .model import plan_rows
            R0 = plan_rows(self.model, [b.job for b in batches], [b.targets for b in batches],
                           [b.negatives for b in batches], self.table_offsets)
            _, ids, used = R0.shared
            self.row_adam.catchup(ids[:used], state=self.adam_clock)

    @torch.no_grad()
    def adam_step(self, res, lr=0.01, betas=(0.9, 0.999), eps=1e-8):
        """One torch.optim.Adam step on EVERY parameter from a StepResult: the dense parameters (layers, mode
        embeddings, readout MLP) in one launch from the flat gradient bucket, the entity tables from the combined
        row-sparse gradient.  Call `catchup_rows(batches)` before the step's forward: with it the trajectory equals
        the reference's dense Adam over the tables (rows a step does not touch are brought up to date lazily)."""
        if self.adam_state is None:
            self._adam_setup(lr, betas, eps)
        lr, b1, b2, eps = self.adam_hyper
        G = res.dense
        with ops.device_guard(G.flat.device):
            ops.adam_tick(self.adam_clock, lr, b1, b2)
            items = [(p, g.contiguous(), m1, m2) for p, g, (m1, m2) in
                     zip(self.adam_params, self._dense_grads(G), self.adam_state)]
            ops.adam_multi(items, lr, b1, b2, eps, state=self.adam_clock)
            uid, urows, num = res.sparse
            self.row_adam.apply(uid, urows, num, state=self.adam_clock)
        self.steps += 1

    adam_step_dense = adam_step     # round-1 name

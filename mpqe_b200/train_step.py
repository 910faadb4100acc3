"""One optimiser step's worth of the hot path without Python autograd in the loop: margin-loss forward + backward
over a list of formula batches, gradient synchronisation across ranks, optional fused Adam.

This is the caller-side row (f)1 of SURVEY.md section 8 (`run_train` / `run_batch_v2`, reference
train_helpers.py:76-120): the reference issues up to 11 `margin_loss` calls per step, one per query type, each
a separate pair of encoder forwards; here all batches of a step go through the same per-pass launches.

Multi-GPU (one process per GPU, torch.distributed): query batches are data-parallel.  Per step there is exactly one
exchange: an all-reduce of the flat dense-gradient bucket and an all-gather of the per-mode (row id, gradient row)
pairs, followed by the same deterministic combine on every rank, so that all ranks apply identical updates.
"""
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


class TrainStep(object):
    def __init__(self, model, margin=1.0, process_group=None, average=True):
        self.model = model
        self.margin = float(margin)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if self._dist() else 1
        self.average = average
        self._layouts = {}
        self.adam_state = None
        self._side = None   # second stream for the id-only half of the row-gradient combine
        self._sym_rows = self._sym_hdl = self._peer_ptrs = None   # peer-visible gradient-row buffer (world > 1)
        self._early_cache = None                                  # marshalled id-emitting launch of the static batches
        self.steps = 0
        # all entity tables as one id space: global row = table_offsets[mode] + row
        self.table_offsets, off = {}, 0
        for mode, module in model.enc.feature_modules.items():
            self.table_offsets[mode] = off
            off += module.weight.shape[0]
        self.total_rows = off

    def _dist(self):
        return torch.distributed.is_available() and torch.distributed.is_initialized()

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

    def refresh(self, batch):
        """A Batch can be re-run: drop the activations of the previous step."""
        j = batch.job
        j.acts = j.outs = j.z = j.u = j.q = j.argmax = j.fwd_groups = None
        return batch

    # ---- the step -------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, batches):
        """Returns StepResult: per-batch losses, weighted total (device scalars), the flat dense gradient bucket
        (`.dense.flat`, views per parameter in `.dense`) and the row-sparse entity gradient
        `.sparse = (unique global row ids, summed rows, num_unique)` with global row = table_offsets[mode] + row."""
        dev = self.model.mode_embeddings.weight.device
        with ops.device_guard(dev):
            early = self._early_plan(batches) if self._use_peer_rows(dev) else None
            res = self._local_step(batches)
            if self.world > 1:
                res = StepResult(res.losses, res.weights, res.dense, self.sync(res.dense, res.sparse, early))
        return res

    def _early_plan(self, batches, defer=False):
        """Data-parallel, peer-memory path: the row ids of EVERY rank's step are known before any rank computes --
        emit them (one small launch), all-gather them (0.8 MB per rank) and build the global combine plan on the second
        stream, under the local step.  The all-gather is also the step's opening barrier: it completes only when every
        rank has finished reading the others' gradient rows of the previous step, after which they may be overwritten."""
        m = self.model
        cache = self._early_cache
        if cache is None or cache[0] is not batches:
            # the id-emitting launch is marshalled once per batch list: in graph mode the same static batches come
            # back every step and the per-step host cost is a single ctypes call
            jobs = [b.job for b in batches]
            R0 = plan_rows(m, jobs, [b.targets for b in batches], [b.negatives for b in batches], self.table_offsets,
                           rows_buffer=lambda cap: None, launch=False)
            cache = (batches, ops.GatherLaunch(R0.id_items, 'ids'), R0)
            self._early_cache = cache if batches is getattr(self, '_static', None) else None
        cache[1].launch()
        _, ids0, used = cache[2].shared
        dev = ids0.device
        all_ids = torch.empty(self.world * used, dtype=torch.int64, device=dev)
        torch.distributed.all_gather_into_tensor(all_ids, ids0[:used], group=self.pg)
        if defer:      # graph mode: the caller enqueues the graph first, then the plan (ordered after this event only)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            return (lambda: self._plan_on_side_stream(all_ids, dev, after=ev)), used
        return self._plan_on_side_stream(all_ids, dev), used

    def _plan_on_side_stream(self, ids, dev, after=None):
        """ops.SparseRowsPlan(ids) on the second stream (joined by `_join_side`).
        `after`: an event to order the plan after, instead of everything enqueued on the current stream so far."""
        cur = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        if after is not None:
            self._side.wait_event(after)
        else:
            self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            plan = ops.SparseRowsPlan(ids, self.total_rows)
        # the ids are read by the sort on the second stream: keep the caching allocator from handing their block to
        # the main stream's next allocation while those kernels are still pending
        ids.record_stream(self._side)
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
            W = Weights(self.model, True)
            self.model._engine.prepare(jobs, W)
            W.ready_event = torch.cuda.Event()
            W.ready_event.record(self._side)
        return W

    def _join_side(self, dev):
        torch.cuda.current_stream(dev).wait_stream(self._side)

    def _on_side_stream(self, dev, fn):
        """Runs fn() on the second stream, ordered after everything enqueued on the current stream so far."""
        self._side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._side):
            fn()

    def _peer_rows_buffer(self, cap):
        """[>= cap, D] gradient-row buffer in memory that every rank of the group can address (torch symmetric memory:
        cuMem allocations exchanged at a rendezvous).  Allocated once (a collective; never during graph capture)."""
        if self._sym_rows is None or self._sym_rows.shape[0] < cap:
            if torch.cuda.is_current_stream_capturing():
                raise ops._lib.MpqeError('the peer-visible row buffer must be allocated before graph capture')
            dev = self.model.mode_embeddings.weight.device
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(int(cap), D, dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(buf, self.pg if self.pg is not None else torch.distributed.group.WORLD)
                self._sym_rows, self._sym_hdl, self._peer_ptrs = buf, hdl, [int(p) for p in hdl.buffer_ptrs]
            except Exception as exc:      # no peer mapping on this system: NCCL all-gather of the rows from now on
                import warnings
                warnings.warn('mpqe_b200: symmetric (peer-mapped) memory unavailable (%r); the row-gradient exchange '
                              'falls back to an NCCL all-gather' % (exc,))
                self._peers_off = True
                self._sym_rows = self._sym_hdl = self._peer_ptrs = None
                return torch.empty(int(cap), D, dtype=torch.float32, device=dev)
        return self._sym_rows

    def _use_peer_rows(self, dev):
        import os
        return (self.world > 1 and dev.type == 'cuda' and not getattr(self, '_peers_off', False) and
                os.environ.get('MPQE_PEER_ROWS', '1') != '0')

    def sync(self, G, sparse, early=None):
        """Data-parallel exchange: all-reduce(dense bucket), all-gather of the ranks' raw (row id, gradient row) pairs
        and ONE combine of all of them, identical on every rank (rank order + stable sort => same bits).  The ids
        travel first (0.8 MB per rank) so that their sort runs on the second stream under the all-gather of the rows
        (54 MB per rank at the bench shape); the 1/world averaging is folded into the row summation.
        `early` = (plan, pairs per rank) from `_early_plan`: the global plan was built under the local step."""
        dist = torch.distributed
        scale = 1.0 / self.world if self.average else 1.0
        ids, rows, _ = sparse
        dev = ids.device
        cap = ids.numel()
        if (early is not None and early[1] == cap and self._peer_ptrs is not None and
                rows.data_ptr() == self._sym_rows.data_ptr()):
            # Peer-memory path.  The dense all-reduce completes only when every rank has finished its local step (its
            # gradient rows are final) -- the barrier the gather needs; then ONE kernel gathers and sums all ranks'
            # rows in place over NVLink (no NCCL all-gather of 54 MB per rank, no local staging).  The next step's id
            # all-gather (`_early_plan`) keeps any rank from overwriting its rows before all have read them.
            dist.all_reduce(G.flat, group=self.pg)
            if scale != 1.0:
                G.flat.mul_(scale)
            self._join_side(dev)
            return early[0].apply_peers(self._peer_ptrs, cap, pad_id=self.total_rows, scale=scale)
        all_ids = torch.empty(self.world * cap, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_ids, ids, group=self.pg)
        plan = self._plan_on_side_stream(all_ids, dev)
        if self._peer_ptrs is not None and rows.data_ptr() == self._sym_rows.data_ptr():
            # Peer-memory path: every rank's rows sit in a buffer mapped into all processes (torch symmetric memory).
            # The id all-gather above completes only after every rank has finished its local step, so the peers' rows
            # are final; ONE kernel then gathers and sums them in place over NVLink (no NCCL all-gather of 54 MB per
            # rank, no local staging).  The dense all-reduce comes last and doubles as the closing barrier: no rank
            # starts overwriting its rows (next step) before every rank has finished reading them.
            self._join_side(dev)
            out = plan.apply_peers(self._peer_ptrs, cap, pad_id=self.total_rows, scale=scale)
            dist.all_reduce(G.flat, group=self.pg)
            if scale != 1.0:
                G.flat.mul_(scale)
            return out
        all_rows = torch.empty(self.world * cap, D, dtype=torch.float32, device=dev)
        dist.all_reduce(G.flat, group=self.pg)
        if scale != 1.0:
            G.flat.mul_(scale)
        dist.all_gather_into_tensor(all_rows, rows, group=self.pg)
        self._join_side(dev)
        return plan.apply(all_rows, pad_id=self.total_rows, scale=scale)

    def _local_step(self, batches):
        """forward + backward + local row-gradient combine (no cross-rank exchange): the part that is graph-captured."""
        m = self.model
        dev = m.mode_embeddings.weight.device
        jobs = [self.refresh(b).job for b in batches]
        tg = [b.targets for b in batches]
        ng = [b.negatives for b in batches]
        # The row ids of the step's entity gradients depend only on the batch ids: emit them first and run the id-only
        # half of the combine (stable sort + segmentation, ~10 small latency-bound launches) on a second stream, under
        # the forward and backward; only the final row summation waits for the gradient rows.
        # With several ranks the pairs are exchanged raw and combined once, after the all-gather (see `sync`).
        # The per-step weight preparation (transposes, tf32 tile images, summed matrices) goes to that stream too, ahead
        # of the sort: it overlaps the input gather; the first layer launch waits for its event.
        R = plan_rows(m, jobs, tg, ng, self.table_offsets,
                      rows_buffer=self._peer_rows_buffer if self._use_peer_rows(dev) else None)
        rows, ids, used = R.shared
        W = self._weights_on_side_stream(jobs, dev)
        plan = self._plan_on_side_stream(ids[:used], dev) if self.world == 1 else None
        key = tuple(b.weight for b in batches)
        wts = getattr(self, '_wts', None)
        if wts is None or wts[0] != key:
            wts = self._wts = (key, torch.tensor(key, dtype=torch.float32, device=dev))
        # d total / d loss_i = the batch weights, known now: the margin backward rides on the margin forward
        losses, W = loss_forward(m, jobs, tg, ng, self.margin, True, grad_losses=wts[1], W=W)
        overlap_tail = plan is not None
        G = loss_backward(m, jobs, W, tg, ng, self.margin, wts[1], self.table_offsets, rows=R,
                          defer_constant=overlap_tail)
        self._weight_decay(W, G, losses, sum(key))
        if plan is not None:
            self._join_side(dev)
            # the batch-constant tail of the backward (five small latency-bound launches) runs on the second
            # stream under the row summation, which does not depend on it
            self._on_side_stream(dev, G.finish)
            sparse = plan.apply(rows[:used], pad_id=self.total_rows)
            self._join_side(dev)
        else:
            sparse = (ids[:used], rows[:used], None)
        return StepResult(losses, wts[1], G, sparse)

    def _weight_decay(self, W, G, losses, weight_sum):
        """The L2 term of every margin_loss call of the step (reference model.py:487-492: weight_decay * sum of the
        un-squared norms of the readout-MLP parameters, once per formula batch): each loss gets the term, the readout
        gradients get (sum of the batch weights) * weight_decay * p / ||p||.  The transposed copies W.w1t / W.w2t have
        the layout of the gradient buffers and the same norms as the parameters."""
        wd = float(self.model.weight_decay)
        if W.ro is None or wd <= 0:
            return
        ops.l2_reg([W.w1t, W.b1, W.w2t, W.b2], [G.dw1t, G.db1, G.dw2t, G.db2], wd, float(weight_sum), losses=losses)

    # ---- CUDA-graph mode: the whole local step becomes one graph launch --------------------------------------
    @torch.no_grad()
    def capture(self, host_batches):
        """Stages `host_batches` into static device buffers, warms up and captures the local step into a CUDA graph.
        Later steps with batches of the same formulas and sizes call `replay(host_batches)`."""
        m = self.model
        dev = m.mode_embeddings.weight.device
        if self._use_peer_rows(dev):
            # the warm-up steps below overwrite this rank's peer-visible gradient rows without the opening barrier of a
            # regular step: first let every rank finish reading them (previous step's gather)
            torch.distributed.barrier(group=self.pg)
        with ops.device_guard(dev):
            # all ids of a step live in ONE device buffer mirrored by ONE pinned host buffer: a step's input is a
            # single H2D copy instead of three small copies per formula batch
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
            self._early_cache = None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._local_step(self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            profile, ops.profile = ops.profile, None      # CUDA events cannot be recorded during capture
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
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
        dev = self._dev_ids.device
        early = None
        if self._use_peer_rows(dev):
            with ops.device_guard(dev):
                # (defer=True would enqueue the plan after the graph launch: less host latency in front of the graph,
                # +12 % end-to-end at N=2, but the plan then overlaps the graph's kernels worse: -12 % device-timed)
                early = self._early_plan(self._static, defer=False)
        self._graph.replay()
        if early is not None and callable(early[0]):     # deferred plan: enqueued after the graph launch, ordered only
            with ops.device_guard(dev):                  # after the id all-gather
                early = (early[0](), early[1])
        res = self._graph_res
        if self.world > 1:
            with ops.device_guard(res.dense.flat.device):
                res = StepResult(res.losses, res.weights, res.dense, self.sync(res.dense, res.sparse, early))
        return res

    def staging(self):
        """Graph mode: HostBatch objects whose id tensors are views of the step's single pinned host buffer.  A data
        loader that writes the next batch into them in place and passes this very list to `replay` / `run_host` gets
        the whole step's input across in one H2D copy, with no packing on the host."""
        return self._staging

    @torch.no_grad()
    def run_host(self, host_batches):
        """End-to-end step from pinned host ids: H2D copies, forward+backward(+sync), D2H of the losses."""
        if getattr(self, '_graph', None) is not None:
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

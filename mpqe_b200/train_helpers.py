"""Training / evaluation loop of the R-GCN path with the reference's structure (mpqe/train_helpers.py:11-162):
burn-in on 1-chain queries, then all query types with `path_weight` / `inter_weight`, hard negatives for the
intersection types, EMA loss, periodic validation and the convergence test.  sacred is not a dependency: metrics
go to the logger and to an optional `log_scalar(name, value, step)` callback.
"""
import zlib

import numpy as np
import torch

from .data_utils import get_queries_iterator
from .utils import eval_auc_queries, eval_perc_queries


def check_conv(vals, window=2, tol=1e-6):
    if len(vals) < 2 * window:
        return False
    return np.mean(vals[-window:]) - np.mean(vals[-2 * window:-window]) < tol


def update_loss(loss, losses, ema_loss, ema_alpha=0.01):
    losses.append(loss)
    ema_loss = loss if ema_loss is None else (1 - ema_alpha) * ema_loss + ema_alpha * loss
    return losses, ema_loss


@torch.no_grad()
def run_eval(model, queries, iteration, logger, batch_size=128, by_type=False, log_scalar=None):
    model.eval()
    vals = {}
    for query_type in queries['one_neg']:
        variants = [('', False)] + ([('hard_', True)] if 'inter' in query_type else [])
        for prefix, hard in variants:
            auc, rel_aucs = eval_auc_queries(queries['one_neg'][query_type], model, hard_negatives=hard)
            perc = eval_perc_queries(queries['full_neg'][query_type], model, batch_size, hard_negatives=hard)
            vals[query_type + ('hard' if hard else '')] = auc
            logger.info('{:s}{:s} val AUC: {:f} val perc {:f}; iteration: {:d}'.format(
                'Hard-' if hard else '', query_type, auc, perc, iteration))
            if log_scalar is not None:
                log_scalar('%s%s_val_auc' % (prefix, query_type), auc, iteration)
                log_scalar('%s%s_val_perc' % (prefix, query_type), perc, iteration)
            if by_type:
                for rels, a in rel_aucs.items():
                    logger.info(str(rels) + '\t' + str(a))
    return vals


def run_batch_v2(queries_iterator, enc_dec, hard_negatives=False):
    enc_dec.train()
    batch = next(queries_iterator)
    return enc_dec.margin_loss(*batch, hard_negatives=hard_negatives)


def run_train(model, optimizer, train_queries, val_queries, test_queries, logger, max_burn_in=100000, batch_size=512,
              log_every=500, val_every=1000, tol=1e-6, max_iter=int(10e7), inter_weight=0.005, path_weight=0.01,
              model_file=None, log_scalar=None):
    edge_conv, ema_loss, vals, losses, conv_test = False, None, [], [], None
    iterators = {qt: get_queries_iterator(qs, batch_size, model) for qt, qs in train_queries.items()}
    i = -1
    for i in range(max_iter):
        optimizer.zero_grad()
        loss = run_batch_v2(iterators['1-chain'], model)
        if not edge_conv and (check_conv(vals) or len(losses) >= max_burn_in):
            logger.info('Edge converged at iteration {:d}'.format(i - 1))
            conv_test = float(np.mean(list(run_eval(model, test_queries, i, logger, log_scalar=log_scalar).values())))
            edge_conv, losses, ema_loss, vals = True, [], None, []
            if model_file is not None:
                torch.save(model.state_dict(), model_file + '-edge_conv')
        if edge_conv:
            for query_type in train_queries:
                if query_type == '1-chain' and max_burn_in > 0:
                    continue
                if 'inter' in query_type:
                    loss = loss + inter_weight * run_batch_v2(iterators[query_type], model)
                    loss = loss + inter_weight * run_batch_v2(iterators[query_type], model, hard_negatives=True)
                else:
                    loss = loss + path_weight * run_batch_v2(iterators[query_type], model)
            if check_conv(vals):
                logger.info('Fully converged at iteration {:d}'.format(i))
                break
        losses, ema_loss = update_loss(loss.item(), losses, ema_loss)
        loss.backward()
        optimizer.step()
        if i % log_every == 0:
            logger.info('Iter: {:d}; ema_loss: {:f}'.format(i, ema_loss))
            if log_scalar is not None:
                log_scalar('ema_loss', ema_loss, i)
        if i >= val_every and i % val_every == 0:
            v = run_eval(model, val_queries, i, logger, log_scalar=log_scalar)
            vals.append(np.mean(list(v.values())) if edge_conv else v['1-chain'])
    v = run_eval(model, test_queries, i, logger, log_scalar=log_scalar)
    test_avg = float(np.mean(list(v.values())))
    logger.info('Test macro-averaged val: {:f}'.format(test_avg))
    if log_scalar is not None:
        log_scalar('test_auc', test_avg, i)
    if conv_test:
        logger.info('Improvement from edge conv: {:f}'.format((test_avg - conv_test) / conv_test))
    if model_file is not None:
        torch.save(model.state_dict(), model_file)
    return test_avg


# ---------------------------------------------------------------------------------------------------------------
# The same loop on the fused step
# ---------------------------------------------------------------------------------------------------------------
class DeviceQuerySets(object):
    """Pre-tensorised query sets (`tensor_queries.TensorQuerySet` per query type) resident on the device: collating a
    batch is a slice of device arrays, and the training negative of every query is drawn on the device by a
    counter-based generator (`ops.sample_negatives`; the reference does `random.choice` per query on the host,
    model.py:470-476) -- a step's host work is the formula pick and a handful of launches."""

    def __init__(self, sets, device, full_lists=None, seed=0):
        self.device, self.seed = device, int(seed)
        self.sets = sets
        self.dev = {}
        for qt, tq in sets.items():
            for f, fq in tq.by_formula.items():
                t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
                self.dev[(qt, f)] = dict(anchors=t(fq.anchors), targets=t(fq.targets), neg_off=t(fq.neg_offsets),
                                         neg_ids=t(fq.neg_ids), hard_off=t(fq.hard_offsets), hard_ids=t(fq.hard_ids))
        self.full = {m: torch.as_tensor(np.asarray(ids, dtype=np.int64)).to(device) for m, ids in
                     (full_lists or {}).items()}
        self.cursor = {qt: 0 for qt in sets}
        self.rng = np.random.RandomState(seed)

    def next_batch(self, qt, batch_size, step, hard=False):
        """(formula, anchors [b, a], targets [b], negatives [b]) -- device tensors -- for the next index window of
        query type `qt`, cut as the reference's loader does (`QueryDataset.collate_fn`, data_utils.py:293-311)."""
        from . import ops
        tq = self.sets[qt]
        first = self.cursor[qt]
        if first >= tq.max_num_queries:
            first = 0
        last = min(first + batch_size, tq.max_num_queries)
        self.cursor[qt] = last
        formula, start, end = tq.pick(list(range(first, last)), self.rng)
        d = self.dev[(qt, formula)]
        n = end - start
        draw = (zlib.crc32(qt.encode()) & 0xffff) * 2000003 + (1000003 if hard else 0) + self.seed   # stream per draw site
        with ops.device_guard(self.device):
            if qt == '1-chain' and not hard:
                neg = ops.sample_negatives(self.full[formula.target_mode], None, n, draw, step)
            else:
                off, ids = (d['hard_off'], d['hard_ids']) if hard else (d['neg_off'], d['neg_ids'])
                neg = ops.sample_negatives(ids, off, n, draw, step, first_query=start)
        return formula, d['anchors'][start:end], d['targets'][start:end], neg


def run_train_fused(model, train_sets, val_queries, test_queries, logger, full_lists=None, max_burn_in=100000,
                    batch_size=512, log_every=500, val_every=1000, tol=1e-6, max_iter=int(10e7), inter_weight=0.005,
                    path_weight=0.01, lr=0.01, model_file=None, log_scalar=None, seed=0):
    """`run_train` (reference train_helpers.py:58-139) on the fused step: all batches of an iteration -- the 1-chain
    batch, after burn-in also the chain batches (x path_weight) and the intersection batches with plain and with hard
    negatives (x inter_weight) -- go through ONE forward + backward (`TrainStep`), the optimiser is the fused Adam
    (dense parameters in one launch, entity tables row-sparse with lazy catch-up: the trajectory of the reference's
    dense Adam), negatives are drawn on the device and the loss is read back only every `log_every` iterations
    (the reference synchronises on `loss.item()` every step, :118).  `train_sets`: {query type: TensorQuerySet}."""
    from .train_step import Batch, TrainStep
    ts = TrainStep(model)
    dev = model.mode_embeddings.weight.device
    data = DeviceQuerySets(train_sets, dev, full_lists, seed)
    edge_conv, ema_loss, vals, conv_test, burn = False, None, [], None, 0
    pending = []          # device totals since the last read-back
    i = -1

    def make(qt, weight, hard=False):
        formula, a, t, n = data.next_batch(qt, batch_size, i, hard)
        tpl, rels, var_host, var_dev, passes = ts.layout(formula)
        from .model import Job
        job = Job(tpl, rels, var_dev, formula.anchor_modes, formula.target_mode, a.contiguous(), passes)
        job.var_rows_host = var_host
        return Batch(job, t.contiguous(), n, weight)

    for i in range(max_iter):
        batches = [make('1-chain', 1.0)]
        if not edge_conv and (check_conv(vals) or burn >= max_burn_in):
            logger.info('Edge converged at iteration {:d}'.format(i - 1))
            ts.catchup_rows(None)
            conv_test = float(np.mean(list(run_eval(model, test_queries, i, logger, log_scalar=log_scalar).values())))
            edge_conv, ema_loss, vals, burn = True, None, [], 0
            if model_file is not None:
                torch.save(model.state_dict(), model_file + '-edge_conv')
        if edge_conv:
            for qt in train_sets:
                if qt == '1-chain' and max_burn_in > 0:
                    continue
                if 'inter' in qt:
                    batches += [make(qt, inter_weight), make(qt, inter_weight, hard=True)]
                else:
                    batches.append(make(qt, path_weight))
            if check_conv(vals):
                logger.info('Fully converged at iteration {:d}'.format(i))
                break
        ts.catchup_rows(batches)
        res = ts.forward_backward(batches)
        pending.append(res.total)
        ts.adam_step(res, lr=lr)
        burn += 1
        if i % log_every == 0:
            for v in torch.stack(pending).cpu().tolist():      # one read-back per log interval
                ema_loss = v if ema_loss is None else 0.99 * ema_loss + 0.01 * v
            pending = []
            logger.info('Iter: {:d}; ema_loss: {:f}'.format(i, ema_loss))
            if log_scalar is not None:
                log_scalar('ema_loss', ema_loss, i)
        if i >= val_every and i % val_every == 0:
            ts.catchup_rows(None)
            v = run_eval(model, val_queries, i, logger, log_scalar=log_scalar)
            vals.append(np.mean(list(v.values())) if edge_conv else v['1-chain'])
    ts.catchup_rows(None)
    v = run_eval(model, test_queries, i, logger, log_scalar=log_scalar)
    test_avg = float(np.mean(list(v.values())))
    logger.info('Test macro-averaged val: {:f}'.format(test_avg))
    if log_scalar is not None:
        log_scalar('test_auc', test_avg, i)
    if conv_test:
        logger.info('Improvement from edge conv: {:f}'.format((test_avg - conv_test) / conv_test))
    if model_file is not None:
        torch.save(model.state_dict(), model_file)
    return test_avg

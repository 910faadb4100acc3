"""Training / evaluation loop of the R-GCN path with the reference's structure (mpqe/train_helpers.py:11-162):
burn-in on 1-chain queries, then all query types with `path_weight` / `inter_weight`, hard negatives for the
intersection types, EMA loss, periodic validation and the convergence test.  sacred is not a dependency: metrics
go to the logger and to an optional `log_scalar(name, value, step)` callback.
"""
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

"""Evaluation metrics of the R-GCN path with the reference's definitions (mpqe/utils.py:25-95).

`eval_auc_queries` / `eval_perc_queries` keep the reference's signatures and batching (128 queries, negatives drawn
with `random.choice` under `random.seed(seed)`), but the ragged negative scoring and the percentile counts run in
the CUDA kernels: scores never go through `.tolist()` + scipy per query.  AUC is the rank statistic of
`sklearn.metrics.roc_auc_score` (average ranks for ties), computed on the host from the device scores.
"""
import random

import numpy as np
import torch

from . import ops


def export_embeddings(enc_dec, entity_ids, path=None, batch_size=65536):
    """The `embeddings.npy` array the reference writes after training for its node-classification script
    (train.py:147-163): one row `[entity id, unit-norm embedding]` per value of `entity_ids` (dict or iterable, in
    iteration order); an id found in no mode's `graph.full_sets` keeps a zero row, the last matching mode wins, as in
    the reference loop.  Rows are gathered and normalised on the device, `batch_size` ids per launch."""
    ids = list(entity_ids.values()) if hasattr(entity_ids, 'values') else list(entity_ids)
    dim = enc_dec.emb_dim
    out = np.zeros((len(ids), 1 + dim))
    device = enc_dec.mode_embeddings.weight.device
    for mode, members in enc_dec.graph.full_sets.items():
        pos = [i for i, e in enumerate(ids) if e in members]
        for lo in range(0, len(pos), batch_size):
            sel = pos[lo:lo + batch_size]
            nodes = torch.tensor([ids[i] for i in sel], dtype=torch.int64, device=device)
            emb = enc_dec.enc(nodes, mode).detach().t().cpu().numpy()      # enc returns [d, n]
            out[sel, 0] = [ids[i] for i in sel]
            out[sel, 1:] = emb
    if path is not None:
        np.save(path, out)
    return out


def auc_from_scores(labels, scores):
    """roc_auc_score(labels, nan_to_num(scores)) via the Mann-Whitney identity (average ranks for ties)."""
    labels = np.asarray(labels).astype(bool)
    scores = np.nan_to_num(np.asarray(scores, dtype=np.float64))
    order = np.argsort(scores, kind='mergesort')
    s = scores[order]
    boundaries = np.flatnonzero(np.concatenate(([True], s[1:] != s[:-1], [True])))
    ranks = np.empty(len(s), dtype=np.float64)
    for lo, hi in zip(boundaries[:-1], boundaries[1:]):
        ranks[lo:hi] = 0.5 * (lo + hi - 1) + 1.0
    r = np.empty_like(ranks)
    r[order] = ranks
    npos = int(labels.sum())
    nneg = len(labels) - npos
    return (r[labels].sum() - npos * (npos + 1) / 2.0) / (npos * nneg)


def auc_from_device_scores(pos, neg):
    """The same statistic from device-resident scores of the positives and of the negatives, without sorting or copying
    them: exact integer pair counts on the device (`ops.auc_counts`), one 16-byte read-back."""
    with ops.device_guard(pos.device):
        lt, eq = ops.auc_counts(pos.contiguous(), neg.contiguous()).tolist()
    return (lt + 0.5 * eq) / (float(pos.numel()) * float(neg.numel()))


def percentile_from_counts(left, right, lengths):
    """scipy.stats.percentileofscore(kind='rank'): (left + right + (left < right)) * 50 / n."""
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    return (left + right + (left < right)) * (50.0 / np.asarray(lengths, dtype=np.float64))


def _get_perc_scores(scores, lengths):
    """Reference signature (utils.py:25-32): scores = [pos..., ragged neg...] -> percentile rank per query."""
    scores = torch.as_tensor(scores, dtype=torch.float32)
    B = len(lengths)
    lengths_t = torch.as_tensor(lengths, dtype=torch.int64)
    offsets = torch.zeros(B + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(lengths_t, 0)
    if scores.is_cuda:
        with ops.device_guard(scores.device):
            left, right = ops.rank_counts_ragged(scores[:B].contiguous(), scores[B:].contiguous(),
                                                 offsets.to(scores.device))
        left, right = left.cpu().numpy(), right.cpu().numpy()
    else:
        raise RuntimeError('percentile counts run on the device: pass CUDA scores (there is no CPU fallback)')
    return percentile_from_counts(left, right, lengths).tolist()


def _batches(formula_queries, batch_size):
    for offset in range(0, len(formula_queries), batch_size):
        yield formula_queries[offset:offset + batch_size]


@torch.no_grad()
def eval_auc_queries(test_queries, enc_dec, batch_size=128, hard_negatives=False, seed=0):
    all_pos, all_neg, formula_aucs = [], [], {}
    random.seed(seed)
    for formula, formula_queries in test_queries.items():
        f_pos, f_neg = [], []
        for batch in _batches(formula_queries, batch_size):
            pool = (lambda q: q.hard_neg_samples) if hard_negatives else (lambda q: q.neg_samples)
            negatives = [random.choice(pool(q)) for q in batch]
            scores = enc_dec.forward(formula, batch, [q.target_node for q in batch], neg_nodes=negatives,
                                     neg_lengths=[1] * len(batch))
            f_pos.append(scores[:len(batch)])       # label 1 (utils.py:49-50)
            f_neg.append(scores[len(batch):])       # label 0
        f_pos, f_neg = torch.cat(f_pos), torch.cat(f_neg)
        formula_aucs[formula] = _auc(f_pos, f_neg)
        all_pos.append(f_pos)
        all_neg.append(f_neg)
    return _auc(torch.cat(all_pos), torch.cat(all_neg)), formula_aucs


def _auc(pos, neg):
    if not pos.is_cuda:
        raise RuntimeError('AUC pair counts run on the device: pass CUDA scores (there is no CPU fallback)')
    return auc_from_device_scores(pos, neg)


@torch.no_grad()
def eval_perc_queries(test_queries, enc_dec, batch_size=128, hard_negatives=False):
    perc = []
    for formula, formula_queries in test_queries.items():
        for batch in _batches(formula_queries, batch_size):
            pool = (lambda q: q.hard_neg_samples) if hard_negatives else (lambda q: q.neg_samples)
            lengths = [len(pool(q)) for q in batch]
            negatives = [n for q in batch for n in pool(q)]
            scores = enc_dec.forward(formula, batch, [q.target_node for q in batch], neg_nodes=negatives,
                                     neg_lengths=lengths)
            perc.extend(_get_perc_scores(scores, lengths))
    return float(np.mean(perc))

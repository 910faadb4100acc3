"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference (behind `oracle/shim`) in this container.

    python -m oracle.make_golden            # rewrites tests/golden/

Each fixture holds the inputs (ids; the graph and the parameters are regenerated deterministically from numpy
RandomState seeds by `mpqe_b200.synthetic.make_kg` and `oracle.mpqe_oracle.init_params`) and the reference's
outputs: the integer batch layout of `RGCNQueryDataset.get_query_graph` (data_utils.py:377-409), eval-style
scores of `RGCNEncoderDecoder.forward` with ragged negatives (model.py:400-462), the percentile scores of
`utils._get_perc_scores` (utils.py:25-32), the `margin_loss` value (model.py:464-494) and every parameter gradient
(non-zero relation slices / table rows only, to keep the files small).
"""
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mpqe_b200 import synthetic  # noqa: E402
from oracle import mpqe_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# (name, query_type, readout, num_layers, adaptive, shared_layers, weight_decay, store_grads)
CASES = [
    ('sum_1chain', '1-chain', 'sum', 2, False, False, 0.0, False),
    ('sum_2chain', '2-chain', 'sum', 2, False, False, 0.0, False),
    ('sum_3chain', '3-chain', 'sum', 2, False, False, 0.0, True),
    ('sum_2inter', '2-inter', 'sum', 2, False, False, 0.0, False),
    ('sum_3inter', '3-inter', 'sum', 2, False, False, 0.0, False),
    ('sum_3interchain', '3-inter_chain', 'sum', 2, False, False, 0.0, True),
    ('sum_3chaininter', '3-chain_inter', 'sum', 2, False, False, 0.0, False),
    ('max_2inter', '2-inter', 'max', 2, False, False, 0.0, True),
    ('max_3chaininter', '3-chain_inter', 'max', 2, False, False, 0.0, False),
    ('concat_3inter', '3-inter', 'concat', 2, False, False, 1e-3, True),
    ('concat_2chain', '2-chain', 'concat', 2, False, False, 1e-3, False),
    ('tm_3chain', '3-chain', 'mp', 3, True, False, 0.0, True),
    ('tm_3inter', '3-inter', 'mp', 3, True, False, 0.0, False),
    ('tm_shared_3interchain', '3-inter_chain', 'mp', 3, True, True, 0.0, False),
    ('mlp_2inter', '2-inter', 'mlp', 2, False, False, 1e-3, False),
    ('targetmlp_3inter', '3-inter', 'targetmlp', 2, False, False, 1e-3, False),
    # scatter_op variants of the MLP readouts (model.py:350-355, 507-513): 9th field
    ('concat_max_3inter', '3-inter', 'concat', 2, False, False, 1e-3, True, 'max'),
    ('mlp_mean_3chain', '3-chain', 'mlp', 2, False, False, 1e-3, True, 'mean'),
    ('targetmlp_max_3interchain', '3-inter_chain', 'targetmlp', 2, False, False, 1e-3, False, 'max'),
    ('concat_mean_2inter', '2-inter', 'concat', 2, False, False, 1e-3, False, 'mean'),
]
D = 128
B = 6
KG_SEED = 3


def load_params_into(model, params, cfg):
    sd = model.state_dict()
    for k in sd:
        src = k
        if cfg.shared_layers and k.startswith('layers.'):
            parts = k.split('.')
            parts[1] = '0'
            src = '.'.join(parts)
        sd[k] = params[src].clone()
    model.load_state_dict(sd)


def run_case(case, kg, qsets):
    name, qt, ro, nl, adaptive, shared, wd, store_grads = case[:8]
    scatter_op = case[8] if len(case) > 8 else 'add'
    ref = ref_loader.load()
    cfg = O.Config(readout=ro, num_layers=nl, adaptive=adaptive, shared_layers=shared, weight_decay=wd,
                   scatter_op=scatter_op)
    rels, raw_queries = qsets[qt][0]
    raw_queries = raw_queries[:B]
    model, graph, id2row = ref_loader.build_reference_model(
        kg.raw(), D, ro, nl, adaptive, shared_layers=shared, weight_decay=wd, scatter_op=scatter_op)
    params = O.init_params(kg.raw()[0], kg.raw()[2], cfg, d=D, seed=11)
    load_params_into(model, params, cfg)
    queries = ref_loader.deserialize_queries(raw_queries)
    formula = queries[0].formula

    anchor_ids, var_ids, qg = ref['data_utils'].RGCNQueryDataset.get_query_graph(
        formula, queries, model.rel_ids, model.mode_ids)
    out = dict(query_type=qt, rels_json=json.dumps(rels), readout=ro, num_layers=nl, adaptive=adaptive,
               shared_layers=shared, weight_decay=wd, scatter_op=scatter_op, kg_seed=KG_SEED, param_seed=11, d=D,
               anchor_ids=anchor_ids.numpy(), var_ids=var_ids.numpy(), edge_index=qg.edge_index.numpy(),
               edge_type=qg.edge_type.numpy(), batch=qg.batch.numpy(),
               targets=np.array([q.target_node for q in queries], dtype=np.int64))

    # eval-style forward with ragged negatives (utils.py:72-95)
    lengths = [len(q.neg_samples) - (i % 3) for i, q in enumerate(queries)]
    negs = [n for i, q in enumerate(queries) for n in q.neg_samples[:lengths[i]]]
    with torch.no_grad():
        scores = model.forward(formula, queries, [q.target_node for q in queries],
                               neg_nodes=negs, neg_lengths=lengths)
    out['eval_neg_nodes'] = np.array(negs, dtype=np.int64)
    out['eval_neg_lengths'] = np.array(lengths, dtype=np.int64)
    out['eval_scores'] = scores.numpy()
    out['eval_perc'] = np.array(ref['utils']._get_perc_scores(scores.tolist(), lengths), dtype=np.float64)

    # training loss: replay the reference's random.choice stream to learn which negatives it drew
    random.seed(5)
    if qt == '1-chain':
        drawn = [random.choice(graph.full_lists[formula.target_mode]) for _ in queries]
    else:
        drawn = [random.choice(q.neg_samples) for q in queries]
    random.seed(5)
    model.zero_grad()
    loss = model.margin_loss(formula, queries, anchor_ids, var_ids, qg)
    loss.backward()
    out['train_neg_nodes'] = np.array(drawn, dtype=np.int64)
    out['loss'] = np.array(loss.item(), dtype=np.float32)
    if store_grads:
        for k, prm in model.named_parameters():
            g = prm.grad
            if g is None:
                g = torch.zeros_like(prm)
            g = g.numpy()
            if g.ndim >= 2 and (k.startswith('enc.') or k.endswith('.basis')):
                nz = np.nonzero(np.abs(g).reshape(g.shape[0], -1).sum(1))[0]
                out['grad_idx:' + k] = nz.astype(np.int64)
                out['grad_rows:' + k] = g[nz]
            else:
                out['grad:' + k] = g
    return name, out


def main():
    if not ref_loader.available():
        raise SystemExit('reference tree missing; goldens can only be generated in the build container')
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)
    kg = synthetic.make_kg('tiny', seed=KG_SEED)
    qsets = synthetic.make_query_sets(kg, queries_per_formula=B, formulas_per_type=1, seed=KG_SEED,
                                      num_neg=5, num_hard_neg=2)
    only = set(sys.argv[1].split(',')) if len(sys.argv) > 1 else None    # regenerate just these cases
    for case in CASES:
        if only is not None and case[0] not in only:
            continue
        name, out = run_case(case, kg, qsets)
        path = os.path.join(GOLDEN_DIR, name + '.npz')
        np.savez_compressed(path, **out)
        print('%-28s loss=%.6f  %6.1f KB' % (name, float(out['loss']), os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    main()

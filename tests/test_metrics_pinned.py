"""The metric definitions are pinned to the libraries the reference calls (mpqe/utils.py:25-32, 65-68):
`sklearn.metrics.roc_auc_score` for the AUC and `scipy.stats.percentileofscore` (default kind='rank') for the
percentile rank -- both on scores with many ties, which is where a re-implementation goes wrong."""
import numpy as np
import pytest

from mpqe_b200 import utils
from oracle import mpqe_oracle as O


def _cases():
    rng = np.random.RandomState(0)
    for n, levels in ((40, 3), (500, 7), (2000, 50), (3000, 0)):
        labels = rng.randint(0, 2, size=n)
        labels[:2] = [0, 1]
        scores = rng.rand(n) if levels == 0 else rng.randint(0, levels, size=n) / float(levels)
        yield labels, scores.astype(np.float32)


def test_auc_equals_sklearn_roc_auc_score():
    sk = pytest.importorskip('sklearn.metrics')
    for labels, scores in _cases():
        want = sk.roc_auc_score(labels, np.nan_to_num(scores))
        assert abs(utils.auc_from_scores(labels, scores) - want) < 1e-12
        assert abs(O.auc(labels, scores) - want) < 1e-12
    with_nan = np.array([0.3, np.nan, 0.7, 0.1], dtype=np.float32)
    assert utils.auc_from_scores([1, 0, 1, 0], with_nan) == sk.roc_auc_score([1, 0, 1, 0], np.nan_to_num(with_nan))


def test_percentile_equals_scipy_percentileofscore():
    st = pytest.importorskip('scipy.stats')
    rng = np.random.RandomState(1)
    for _ in range(200):
        n = int(rng.randint(1, 40))
        neg = (rng.randint(0, 6, size=n) / 5.0).astype(np.float32)
        pos = np.float32(rng.randint(0, 6) / 5.0)
        left, right = int((neg < pos).sum()), int((neg <= pos).sum())
        want = st.percentileofscore(neg, pos)          # the reference's call: default kind='rank'
        assert abs(utils.percentile_from_counts([left], [right], [n])[0] - want) < 1e-9
        assert abs(O.percentile_from_counts([left], [right], [n])[0] - want) < 1e-9
        l2, r2 = O.rank_counts([pos], neg, [n])
        assert (int(l2[0]), int(r2[0])) == (left, right)

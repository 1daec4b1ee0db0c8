"""GPU: downstream parity (north_star: link-prediction AUC and node-classification F1 of the
embeddings must match the reference's within +-0.005).  The full reference configuration is run
on the GPU through the reference-shaped API (`Algorithms.AlgoForce2Vec*`, 1200 epochs on cora);
the comparison embedding is the reference's own shipped golden .embd (option 5) or the oracle's
run of the same configuration (option 6).  After 1200 free-running epochs the two embeddings
differ by chaotic fp32 drift (two builds of the reference differ by the same amount, SURVEY
section 4), so the meaningful statement is about the scores, computed with identical seeded
splits by tools/evalscores.py (the reference's protocols restated)."""
import os
import sys
import numpy as np
import pytest

import force2vec_b200 as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu
TOL = 0.005


def _scores(rp, ci, X, labels):
    import evalscores as E
    lp = E.link_prediction(rp, ci, X, seeds=(1, 2))
    nc = E.node_classification(X, labels, seeds=tuple(range(10)))
    return lp, nc


def _compare(rp, ci, got, want):
    import evalscores as E
    labels = E.read_labels(os.path.join(GOLDEN, "cora.nodes.labels"), len(rp) - 1)
    lp_a, nc_a = _scores(rp, ci, got, labels)
    lp_b, nc_b = _scores(rp, ci, want, labels)
    for k in ("accuracy", "f1_macro", "f1_micro", "auc"):
        assert abs(lp_a[k] - lp_b[k]) <= TOL, ("link prediction", k, lp_a, lp_b)
    for tf in nc_a:
        for k in ("f1_macro", "f1_micro"):
            assert abs(nc_a[tf][k] - nc_b[tf][k]) <= TOL, ("node classification", tf, k, nc_a[tf], nc_b[tf])
    return lp_a, nc_a


def test_option5_scores_match_shipped_golden(cora):
    rp, ci = cora
    alg = F.Algorithms(rp, ci, "cora.mtx", "/tmp/", 128)
    alg.AlgoForce2VecNS(1200, 0, 384, 5, 0.02, write=False)
    want = np.load(os.path.join(GOLDEN, "shipped_cora_F2VNS384D128IT1200NS5.npz"))["X"]
    rel = np.linalg.norm(alg.nCoordinates - want) / np.linalg.norm(want)
    assert rel < 3e-2, rel          # free-running 1200 epochs: the spread two reference builds show (1.2e-2)
    lp, nc = _compare(rp, ci, alg.nCoordinates, want)
    assert lp["accuracy"] > 0.8 and nc[0.25]["f1_micro"] > 0.7     # and the embedding is a useful one


def test_option6_scores_match_oracle(cora, oracle):
    rp, ci = cora
    alg = F.Algorithms(rp, ci, "cora.mtx", "/tmp/", 128)
    alg.AlgoForce2VecNSRW(1200, 0, 256, 5, 0.02, write=False)
    want = oracle.run(6, 0, rp, ci, 128, 1200, 256, 5, 0.02, threads=os.cpu_count() or 1)["X"]
    lp, nc = _compare(rp, ci, alg.nCoordinates, want)
    assert lp["accuracy"] > 0.95

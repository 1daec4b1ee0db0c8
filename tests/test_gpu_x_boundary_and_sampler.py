"""GPU: the drop-in boundary proven with the reference's own driver (INTEGRATION.md section B compiled), and the
statistical parity claim of the device walk sampler.  (Kept in their own file, after the parity suites.)"""
import os
import subprocess
import sys
import numpy as np
import pytest
from conftest import GOLDEN, ROOT

import force2vec_b200 as F

sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu
TOL = 0.005


def test_reference_cli_bound_to_libf2v_equals_drop_in_cli(tmp_path):
    """The drop-in boundary, proven: oracle/_ref/Force2Vec_f2v is the REFERENCE's own driver, loaders
    and class (Test/Force2Vec.cpp, IO.h/CSC.h/CSR.h, algorithms.h incl. its writeToFile) compiled with
    the five hot-path method bodies replaced by INTEGRATION.md section B's binding (oracle/ref_binding.cpp
    -> f2v_train in libf2v.so).  Its .embd must be byte-identical to bin/Force2Vec's (our loader, our
    writer, the same engine) for options 5/6/7, bs 0/1 -- and its Results.txt row has the same shape."""
    ref = os.path.join(ROOT, "oracle", "_ref", "Force2Vec_f2v")
    exe = os.path.join(ROOT, "bin", "Force2Vec")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/Force2Vec_f2v not built (needs /root/reference at build time)")
    cases = [("cora.mtx", 5, 0, 128, 256, 3), ("cora.mtx", 5, 1, 64, 384, 2), ("cora.mtx", 6, 0, 128, 256, 3),
             ("cora.mtx", 6, 1, 64, 100, 2), ("cora.mtx", 7, 0, 64, 256, 3), ("karate.mtx", 5, 0, 16, 8, 5)]
    for k, (g, opt, bs, dim, batch, it) in enumerate(cases):
        outs = []
        for who, binary in (("ref", ref), ("ours", exe)):
            d = tmp_path / ("%s%d" % (who, k))
            d.mkdir()
            r = subprocess.run([binary, "-input", os.path.join(GOLDEN, g), "-output", str(d) + "/", "-iter", str(it),
                                "-batch", str(batch), "-dim", str(dim), "-nsamples", "5", "-option", str(opt), "-bs", str(bs)],
                               capture_output=True, cwd=str(d))
            assert r.returncode == 0, (who, r.stdout[-500:], r.stderr[-500:])
            embd = [f for f in os.listdir(d) if f.endswith(".embd")]
            assert len(embd) == 1, embd
            outs.append((embd[0], (d / embd[0]).read_bytes(), (d / "Results.txt").read_text()))
        assert outs[0][0] == outs[1][0]                       # same output file name
        assert outs[0][1] == outs[1][1], (g, opt, bs)         # same bytes
        strip = lambda row: row.split("\tTime(sec.):")[0]
        assert strip(outs[0][2]) == strip(outs[1][2])



def test_option7_device_walk_sampler_scores_match_libc_walks(cora):
    """`-walk 1` (the device sampler: counter-based draws, parallel) cannot consume the serial libc
    stream, so its embeddings are a different random sample, not the reference's bits.  Its parity claim is
    downstream and statistical: after the full reference configuration (cora, option 7, 1200 epochs) every
    link-prediction score (accuracy / F1 / AUC) and node-classification F1 lies inside the band that runs
    walking off the libc-compatible stream (`-walk 0`, the path the reference-golden tests pin) span when only
    their srand() seed changes (1, 2, 3, 4), widened by the north_star's 0.005 -- same seeded splits everywhere.
    (Measured once: against a single libc seed every metric was within 0.0053; a +-0.005 comparison between
    two different random streams tests the seeds as much as the sampler, hence the band.)"""
    import evalscores as E
    rp, ci = cora
    labels = E.read_labels(os.path.join(GOLDEN, "cora.nodes.labels"), len(rp) - 1)

    def run(walk, seed):
        alg = F.Algorithms(rp, ci, "cora.mtx", "/tmp/", 64)
        alg.walk_sampler, alg.seed = walk, seed
        alg.AlgoForce2VecNSRWEFF(1200, 0, 256, 5, 0.02, write=False)
        lp_ = E.link_prediction(rp, ci, alg.nCoordinates, seeds=(1, 2))
        nc_ = E.node_classification(alg.nCoordinates, labels, seeds=tuple(range(10)))
        return lp_, nc_

    libc = [run(0, seed) for seed in (1, 2, 3, 4)]
    lp, nc = run(1, 1)
    for k in ("accuracy", "f1_macro", "f1_micro", "auc"):
        lo, hi = min(r[0][k] for r in libc), max(r[0][k] for r in libc)
        assert lo - TOL <= lp[k] <= hi + TOL, ("link prediction", k, lp[k], lo, hi)
    for tf in nc:
        for k in ("f1_macro", "f1_micro"):
            lo, hi = min(r[1][tf][k] for r in libc), max(r[1][tf][k] for r in libc)
            assert lo - TOL <= nc[tf][k] <= hi + TOL, ("node classification", tf, k, nc[tf][k], lo, hi)

#!/usr/bin/env python
"""tools/tune.py -- sweep engine knobs on one resident workload (development aid, not the bench).
   python tools/tune.py --scale 20 --model 6 --dim 128 --batches 16384,65536 --variants 3,8,11 --chunks 64,128"""
import argparse
import itertools
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import force2vec_b200 as F  # noqa: E402
from force2vec_b200 import host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--model", type=int, default=6)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--bs", type=int, default=0)
    ap.add_argument("--batches", default="16384")
    ap.add_argument("--variants", default="-1")
    ap.add_argument("--chunks", default="64")
    ap.add_argument("--modes", default="0")
    ap.add_argument("--negsmem", default="1")
    ap.add_argument("--pars", default="9472")
    ap.add_argument("--pdl", default="2")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rp, ci = host.rmat_csr(a.scale, 16, 1)
    n, nnz = len(rp) - 1, len(ci)
    g = host.RandStream(1)
    X0 = g.init_embeddings(a.model, n, a.dim)
    pairs = n * 10 if a.model == 7 else nnz + 5 * n
    byts = pairs * a.dim * 4 + n * a.dim * 4
    eng = F.Engine(rp, ci, a.dim)
    if a.model != 5:
        eng.set_lut()
    eng.set_embeddings(X0)
    if a.model == 7:
        eng.sample_walks(1, 0)
    rows = []
    for B, var, ch, mode, ns, par in itertools.product([int(x) for x in a.batches.split(",")], [int(x) for x in a.variants.split(",")],
                                                       [int(x) for x in a.chunks.split(",")], [int(x) for x in a.modes.split(",")],
                                                       [int(x) for x in a.negsmem.split(",")], [int(x) for x in a.pars.split(",")]):
      for pdl in [int(x) for x in a.pdl.split(",")]:
            neg = g.epoch_negatives(a.model, n, B, 5, a.bs).copy()
            eng.set_negatives(neg)
            eng.set_option("variant", var)
            eng.set_option("neg_smem", ns)
            eng.set_option("pdl", pdl)
            eng.set_option("par", par)
            try:
                eng.set_epoch_mode(mode)
            except F.F2VError as ex:
                print("mode", mode, "unavailable:", ex)
                continue
            eng.run_epoch(a.model, B, 5, a.bs, 0.02, ch)   # warm (plan build)
            eng.sync()
            ms = []
            for _ in range(a.reps):
                eng.run_epoch(a.model, B, 5, a.bs, 0.02, ch)
                ms.append(eng.last_epoch_ms())
            best = min(ms)
            row = {"B": B, "variant": var, "chunk": ch, "mode": mode, "neg_smem": ns, "par": par, "pdl": pdl, "ms": best, "ms_all": ms,
                   "Gpairs_s": pairs / best / 1e6, "GBs": byts / best / 1e6, "frac": byts / best / 1e6 / 6553.3}
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()

"""Dataflow epoch (mode 2) against per-minibatch launches (mode 0) on one resident workload.
  python tools/r2_tune_flow.py SCALE MODEL DIM BS "B1,B2,..." """
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
    scale, model, dim, bs = (int(x) for x in sys.argv[1:5])
    batches = [int(x) for x in sys.argv[5].split(",")]
    rp, ci = host.rmat_csr_cached(scale, 16, 1)
    n, nnz = len(rp) - 1, len(ci)
    pairs = n * 10 if model == 7 else nnz + 5 * n
    g = host.RandStream(1)
    X0 = g.init_embeddings(model, n, dim)
    e = F.Engine(rp, ci, dim)
    if model != 5:
        e.set_lut()
    if model == 7:
        e.sample_walks(1, 0)
    for B in batches:
        neg = g.epoch_negatives(model, n, B, 5, bs).copy()
        e.set_negatives(neg)
        sums = {}
        for mode, variant, par in ((0, -1, 9472), (2, -1, 9472), (2, 11, 9472), (2, 8, 9472), (2, -1, 0), (0, -1, 0)):
            e.set_epoch_mode(mode)
            e.set_option("variant", variant)
            e.set_option("par", par)
            e.set_embeddings(X0)
            e.set_negative_offset(0)
            t = time.time()
            e.run_epoch(model, B, 5, bs, 0.02)
            e.sync()
            first = time.time() - t
            sums[(mode, variant, par)] = e.checksum()
            ms = []
            for k in range(3):
                e.set_negative_offset(0)
                e.run_epoch(model, B, 5, bs, 0.02)
                ms.append(e.last_epoch_ms())
            f, tot = e.device_memory()
            print(json.dumps({"scale": scale, "model": model, "dim": dim, "bs": bs, "B": B, "mode": mode, "variant": variant, "par": par,
                              "epoch_ms": [round(x, 3) for x in ms], "best_ms": min(ms), "Gpairs_s": pairs / min(ms) / 1e6,
                              "first_call_s": round(first, 2), "mem_used_GiB": round((tot - f) / 2**30, 2),
                              "checksum": "%016x" % sums[(mode, variant, par)]}), flush=True)
        a = {k: v for k, v in sums.items() if k[2] == 9472}
        b = {k: v for k, v in sums.items() if k[2] == 0}
        print("B", B, "adaptive-chunk checksums equal:", len(set(a.values())) == 1, "fixed-chunk equal:", len(set(b.values())) == 1, flush=True)
    e.close()


if __name__ == "__main__":
    main()

"""A/B of the force-kernel layouts on one resident workload: epoch time per variant and a device
checksum after one epoch from the same state (all variants must agree bit for bit).
  python tools/r2_tune_ring.py SCALE MODEL DIM BS BATCH "v1,v2,..." """
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
    scale, model, dim, bs, B = (int(x) for x in sys.argv[1:6])
    variants = [int(x) for x in sys.argv[6].split(",")]
    rp, ci = host.rmat_csr_cached(scale, 16, 1)
    n, nnz = len(rp) - 1, len(ci)
    pairs = n * 10 if model == 7 else nnz + 5 * n
    g = host.RandStream(1)
    X0 = g.init_embeddings(model, n, dim)
    neg = g.epoch_negatives(model, n, B, 5, bs).copy()
    e = F.Engine(rp, ci, dim)
    if model != 5:
        e.set_lut()
    e.set_negatives(neg)
    if model == 7:
        e.sample_walks(1, 0)
    sums = {}
    for v in variants:
        e.set_option("variant", v)
        e.set_embeddings(X0)
        e.set_negative_offset(0)
        e.run_epoch(model, B, 5, bs, 0.02)
        sums[v] = e.checksum()
        ms = []
        for k in range(4):
            e.set_negative_offset(0)
            e.run_epoch(model, B, 5, bs, 0.02)
            ms.append(e.last_epoch_ms())
        print(json.dumps({"scale": scale, "model": model, "dim": dim, "bs": bs, "B": B, "variant": v,
                          "epoch_ms": [round(x, 3) for x in ms], "best_ms": min(ms), "Gpairs_s": pairs / min(ms) / 1e6,
                          "checksum": "%016x" % sums[v]}), flush=True)
    print("CHECKSUMS_EQUAL" if len(set(sums.values())) == 1 else "CHECKSUMS_DIFFER", sums, flush=True)
    e.close()


if __name__ == "__main__":
    main()

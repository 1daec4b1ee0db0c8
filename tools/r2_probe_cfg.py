"""Single-GPU probe of a BASELINE config: epoch time per batch size, per-minibatch trace summary.
  python tools/r2_probe_cfg.py SCALE MODEL DIM BS "B1,B2,..." [epochs]
Env: VARIANT, CHUNK, ONLY_EPOCHS=1 (no trace; for ncu)."""
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
    epochs = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    t = time.time()
    rp, ci = host.rmat_csr_cached(scale, 16, 1)
    n, nnz = len(rp) - 1, len(ci)
    print("graph", n, nnz, "%.1fs" % (time.time() - t), flush=True)
    pairs = n * 10 if model == 7 else nnz + 5 * n
    alg = pairs * dim * 4 + n * dim * 4
    g = host.RandStream(1)
    t = time.time()
    X0 = g.init_embeddings(model, n, dim)
    print("init %.1fs" % (time.time() - t), flush=True)
    e = F.Engine(rp, ci, dim)
    if model != 5:
        e.set_lut()
    e.set_embeddings(X0)
    del X0
    if "VARIANT" in os.environ:
        e.set_option("variant", int(os.environ["VARIANT"]))
    chunk = int(os.environ.get("CHUNK", "0"))
    deg = np.diff(np.asarray(rp).astype(np.int64))
    for B in batches:
        neg = g.epoch_negatives(model, n, B, 5, bs).copy()
        e.set_negatives(neg)
        ms = []
        for k in range(epochs):
            e.set_negative_offset(0)
            if model == 7:
                e.sample_walks(1, k)
            e.run_epoch(model, B, 5, bs, 0.02, chunk)
            ms.append(e.last_epoch_ms())
        best = min(ms[1:]) if len(ms) > 1 else ms[0]
        out = {"scale": scale, "model": model, "dim": dim, "bs": bs, "B": B, "epoch_ms": [round(x, 3) for x in ms],
               "Gpairs_s": pairs / best / 1e6, "alg_TBs": alg / best / 1e9}
        if os.environ.get("ONLY_EPOCHS") != "1":
            e.set_option("trace", 1)
            e.set_negative_offset(0)
            e.run_epoch(model, B, 5, bs, 0.02, chunk)
            tr = e.trace_ms() * 1e3
            e.set_option("trace", 0)
            nb = len(tr)
            ed = np.add.reduceat(deg, np.arange(0, n, B))[:nb]
            out["trace_us"] = {"nb": nb, "sum_ms": float(tr.sum() / 1e3), "first8": [round(float(x), 1) for x in tr[:8]],
                               "last4": [round(float(x), 1) for x in tr[-4:]],
                               "quartile_sums_ms": [round(float(tr[q * nb // 4:(q + 1) * nb // 4].sum() / 1e3), 3) for q in range(4)],
                               "quartile_edges_M": [round(float(ed[q * nb // 4:(q + 1) * nb // 4].sum() / 1e6), 1) for q in range(4)]}
        print(json.dumps(out), flush=True)
    e.close()


if __name__ == "__main__":
    main()

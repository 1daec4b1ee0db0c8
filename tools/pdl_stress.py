"""Stress for launch chaining: many short runs on tiny graphs with the sample streams re-uploaded
every epoch; all results must be bit-identical to the unchained (pdl 0) run."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import force2vec_b200 as F
from force2vec_b200 import host
bad = runs = 0
for scale, ef, dim, batches in ((8, 4, 32, (37, 306)), (8, 4, 20, (16, 37)), (13, 16, 20, (16,)), (10, 16, 128, (64, 256)), (12, 16, 64, (256,))):
    rp, ci = host.rmat_csr(scale, ef, 2)
    n = len(rp) - 1
    for model, bs in ((5, 0), (5, 1), (6, 0), (6, 1), (7, 0)):
        for batch in batches:
            g = host.RandStream(1)
            X0 = g.init_embeddings(model, n, dim)
            st = []
            for it in range(3):
                w = g.walks(rp, ci).copy() if model == 7 else None
                st.append((w, g.epoch_negatives(model, n, batch, 5, bs).copy()))
            ref = None
            for pdl in (0,) + (2,) * 10 + (1,) * 4:
                with F.Engine(rp, ci, dim) as e:
                    e.set_option("pdl", pdl)
                    e.set_embeddings(X0)
                    if model != 5: e.set_lut()
                    for w, neg in st:
                        if w is not None: e.set_walks(w)
                        e.set_negatives(neg)
                        e.run_epoch(model, batch, 5, bs, 0.02)
                    X = e.get_embeddings()
                runs += 1
                if ref is None: ref = X
                elif not np.array_equal(ref, X):
                    bad += 1
                    print("MISMATCH scale", scale, "dim", dim, "model", model, bs, "batch", batch, "pdl", pdl, np.abs(ref - X).max(), flush=True)
print("stress done, runs", runs, "mismatches:", bad)

"""Development probe: host enqueue time vs device time of one epoch at small batch sizes
(is the launch loop CPU-bound?)."""
import sys, time; sys.path.insert(0, ".")
import numpy as np
import force2vec_b200 as F
from force2vec_b200 import host
rp, ci = host.rmat_csr(20, 16, 1); n = len(rp) - 1
g = host.RandStream(1); X0 = g.init_embeddings(6, n, 128)
e = F.Engine(rp, ci, 128); e.set_lut(); e.set_embeddings(X0)
for B in (256, 1024, 4096):
    neg = g.epoch_negatives(6, n, B, 5, 0).copy(); e.set_negatives(neg)
    for pdl in (2, 0):
        e.set_option("pdl", pdl)
        for it in range(3):
            e.set_negative_offset(0); e.sync()
            t0 = time.perf_counter(); e.run_epoch(6, B, 5, 0, 0.02); t1 = time.perf_counter()
            ms = e.last_epoch_ms()
        print("B", B, "pdl", pdl, "enqueue ms %.2f" % ((t1 - t0) * 1e3), "device ms %.2f" % ms, flush=True)

"""Development probe: epoch time at small batch sizes for layout variants x minimum chunk length."""
import sys; sys.path.insert(0,".")
import numpy as np
import force2vec_b200 as F
from force2vec_b200 import host
rp,ci=host.rmat_csr(20,16,1); n=len(rp)-1
g=host.RandStream(1); X0=g.init_embeddings(6,n,128)
e=F.Engine(rp,ci,128); e.set_lut(); e.set_embeddings(X0)
for B in (256, 1024, 4096, 16384):
    neg=g.epoch_negatives(6,n,B,5,0).copy(); e.set_negatives(neg)
    for variant in (3, 11):
        for mc in (8, 16, 32, 64):
            e.set_option("variant", variant); e.set_option("min_chunk", mc)
            ms=[]
            for it in range(4):
                e.set_negative_offset(0); e.run_epoch(6,B,5,0,0.02); ms.append(e.last_epoch_ms())
            print("B",B,"variant",variant,"min_chunk",mc,"ms %.3f"%min(ms[1:]), "us/minibatch %.2f"%(min(ms[1:])*1e3/((n+B-1)//B)), flush=True)

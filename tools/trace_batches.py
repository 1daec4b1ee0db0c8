import sys; sys.path.insert(0,".")
import numpy as np
import force2vec_b200 as F
from force2vec_b200 import host
rp,ci=host.rmat_csr(20,16,1); n=len(rp)-1
deg=np.diff(rp.astype(np.int64))
g=host.RandStream(1); X0=g.init_embeddings(6,n,128)
e=F.Engine(rp,ci,128); e.set_lut(); e.set_embeddings(X0)
for B in (256, 4096):
    neg=g.epoch_negatives(6,n,B,5,0).copy(); e.set_negatives(neg)
    e.run_epoch(6,B,5,0,0.02); e.set_option("trace",1); e.set_negative_offset(0); e.run_epoch(6,B,5,0,0.02)
    t=e.trace_ms()*1e3; e.set_option("trace",0)
    nb=len(t); md=np.array([deg[b*B:(b+1)*B].max() for b in range(nb)]); ed=np.array([deg[b*B:(b+1)*B].sum() for b in range(nb)])
    print("B",B,"nb",nb,"sum ms",t.sum()/1e3,"mean us",t.mean(),"median",np.median(t),"p10",np.percentile(t,10),"p90",np.percentile(t,90),"p99",np.percentile(t,99),"max",t.max())
    order=np.argsort(-t)[:8]
    print(" top:",[(int(b),round(float(t[b]),1),int(md[b]),int(ed[b])) for b in order])
    # correlation with max degree: bucket by max degree
    for lo,hi in ((0,16),(16,64),(64,256),(256,1024),(1024,4096),(4096,1<<30)):
        m=(md>=lo)&(md<hi)
        if m.any(): print("  maxdeg [%d,%d): %d minibatches, mean %.1f us, mean edges %.0f"%(lo,hi,m.sum(),t[m].mean(),ed[m].mean()))

"""Host side of BASELINE config 5 at FULL scale, no GPU needed: R-MAT scale 26 from the product's generator
(cached under /dev/shm like bench.py does), the work plan of a single-GPU engine and of every one of 8 ranks
(degree-balanced ownership, the orders and chunk lengths bench.py's cfg5 runs use), 64-bit edge offsets.
Output kept in profiles/r2_cfg5_host_side.log.   python tools/cfg5_host_check.py [scale]"""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from force2vec_b200 import host  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 26
t = time.time()
rp, ci = host.rmat_csr_cached(scale, 16, 1)
rp = np.ascontiguousarray(rp)
n, nnz = len(rp) - 1, len(ci)
print("R-MAT scale %d: n %d  nnz %d  largest row %d  n*128 = 2^%d  (%.1f s incl. page-cache copy; %d host threads)" %
      (scale, n, nnz, int(np.diff(rp).max()), (n * 128 - 1).bit_length(), time.time() - t, len(os.sched_getaffinity(0))), flush=True)
batch = 262144
for world, chunk, assign in ((1, 64, 0), (8, 64, 3), (8, 256, 3)):
    shares, items = [], []
    t = time.time()
    for rank in range(world):
        pl = host.plan_build(rp, batch, chunk=chunk, rank=rank, world=world, par=9472, assign=assign)
        it = pl["items"]
        shares.append(int((it["len"] & 0x7fffffff).astype(np.uint64).sum()))
        items.append(len(it))
        assert int(it["e0"].max()) <= nnz and it["e0"].dtype == np.uint64
    assert sum(shares) == nnz                                   # every CSR entry planned exactly once
    dev = max(abs(x - nnz / world) for x in shares) / (nnz / world)
    print("plan: world %d  hub chunk %d  minibatches %d  items/rank %d..%d  edges/rank within %.4f %% of nnz/world  "
          "(%.1f s per rank)" % (world, chunk, pl["nb"], min(items), max(items), 100 * dev, (time.time() - t) / world), flush=True)
print("largest edge offset %d (%d bits; kept in 64-bit fields); largest table element index n*128 - 1 needs %d bits" %
      (nnz, nnz.bit_length(), (n * 128 - 1).bit_length()))

"""Worker for tests/test_multi_gpu.py (one rank per GPU under torch.distributed.run)."""
import os
import sys
# torch.distributed.run exports OMP_NUM_THREADS=1: the host side of the library is OpenMP code
os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import force2vec_b200 as F  # noqa: E402
from force2vec_b200 import host  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rp, ci = host.rmat_csr(13, 16, 2)
    n = len(rp) - 1
    ok = True
    for model, bs, dim, batch in ((5, 0, 128, 512), (5, 1, 128, 512), (6, 0, 128, 1000), (7, 0, 64, 512), (6, 1, 64, 256),
                                  (6, 0, 128, 8192), (5, 0, 20, 16)):
        if batch % world:
            batch += world - batch % world
        g = host.RandStream(1)
        X0 = g.init_embeddings(model, n, dim)
        streams = []
        for it in range(2):
            w = g.walks(rp, ci).copy() if model == 7 else None
            streams.append((w, g.epoch_negatives(model, n, batch, 5, bs).copy()))

        def run(eng):
            eng.set_embeddings(X0)
            if model != 5:
                eng.set_lut()
            for w, neg in streams:
                if w is not None:
                    eng.set_walks(w)
                eng.set_negatives(neg)
                eng.run_epoch(model, batch, 5, bs, 0.02, chunk=64)   # same hub chunking on every engine
            return eng.get_embeddings()

        single = F.Engine(rp, ci, dim, device=local)
        b = run(single)
        single.close()
        # peer: NVLink multicast stores where the box supports them; peer_unicast: one store per peer
        # peer_fallback: rank 0 reports that the multicast object cannot be created -> all ranks go unicast
        # peer_sharded: row-sharded tables (each GPU stores 1/world of the rows, gathers cross NVLink)
        for comm in ("peer", "peer_multicast", "peer_unicast", "peer_fallback", "peer_nopdl", "peer_sharded", "nccl"):
            if comm == "peer_sharded" and world & (world - 1):
                continue
            multi = F.Engine(rp, ci, dim, device=local)
            if comm == "peer_nopdl":
                multi.set_option("pdl", 0)       # ordinary launches: every launch waits for the peers at kernel entry
            if comm == "peer_multicast":
                multi.set_option("multicast", 3)     # NVLink multicast stores also at N=2 (default there: one store per peer)
            if comm == "peer_unicast":
                multi.set_option("multicast", 0)
            if comm == "peer_fallback":
                multi.set_option("multicast", 2)
            if comm == "peer_sharded":
                multi.set_option("sharded", 1)
            if comm == "nccl":
                ids = [F.Engine.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                multi.comm_init(ids[0], rank, world)
            else:
                blobs = [None] * world
                dist.all_gather_object(blobs, multi.comm_peer_export())
                multi.comm_peer_init(blobs, rank, world)
            a = run(multi)
            if comm in ("peer", "peer_multicast", "peer_unicast"):
                # host-buffer epochs: every rank moves only its 1/world share of the table over PCIe
                # (the rest travels over NVLink) and gets its share of the result back
                per = -(-n // world)
                lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
                Xin, Xout = X0.copy(), np.zeros_like(X0)
                for w, neg in streams:
                    multi.run_epoch_host(model, batch, 5, bs, 0.02, X_in=Xin, neg=neg, walks=w, X_out=Xout, chunk=64)
                    parts = [None] * world
                    dist.all_gather_object(parts, Xout[lo:hi].copy())
                    Xin = np.concatenate(parts)              # the job's result = the ranks' shares
                if not np.array_equal(Xin, b):
                    print("rank", rank, comm, "host-buffer epochs differ, model", model, bs, flush=True)
                    ok = False
            dist.barrier()           # nobody unmaps a table a peer may still be storing into
            multi.close()
            same = np.array_equal(a, b)
            if not same:
                print("rank", rank, comm, "model", model, "bs", bs, "max diff", np.abs(a - b).max(), flush=True)
            ok &= same
    # a rank that never shows up must be reported, not waited for: rank 0 runs one single-minibatch
    # epoch alone with a 1.5 s exchange time-out and has to get an error at the synchronisation
    rp2, ci2 = host.rmat_csr(9, 8, 3)
    n2 = len(rp2) - 1
    lonely = F.Engine(rp2, ci2, 32, device=local)
    lonely.set_option("exchange_timeout_ms", 1500)
    blobs = [None] * world
    dist.all_gather_object(blobs, lonely.comm_peer_export())
    lonely.comm_peer_init(blobs, rank, world)
    if rank == 0:
        g2 = host.RandStream(1)
        lonely.set_embeddings(g2.init_embeddings(5, n2, 32))
        lonely.set_negatives(g2.epoch_negatives(5, n2, n2, 5, 0).copy())
        lonely.run_epoch(5, n2, 5, 0, 0.02)
        try:
            lonely.sync()
            print("rank 0: the missing peer was not reported", flush=True)
            ok = False
        except F.F2VError as ex:
            if "timed out" not in str(ex):
                print("rank 0: unexpected error", ex, flush=True)
                ok = False
    dist.barrier()
    lonely.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK" if int(t.item()) == 1 else "MGPU_FAIL", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()

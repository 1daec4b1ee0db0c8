"""Timing probe for the multi-GPU exchange (run under torch.distributed.run, one rank per GPU).
Prints per-variant epoch times: full peer-store exchange, without the row stores, without the
flag barrier, without both (= this rank's share of the compute alone)."""
import json
import os
import sys
# torch.distributed.run exports OMP_NUM_THREADS=1: the host side of the library is OpenMP code
os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import force2vec_b200 as F  # noqa: E402
from force2vec_b200 import host  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    scale = int(os.environ.get("SCALE", "20"))
    model, dim, s = int(os.environ.get("MODEL", "6")), int(os.environ.get("DIM", "128")), 5
    bs = int(os.environ.get("BS", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank != 0:
        dist.barrier()
    rp, ci = host.rmat_csr_cached(scale, 16, 1)
    if rank == 0:
        dist.barrier()
    n = len(rp) - 1
    g = host.RandStream(1)
    X0 = g.init_embeddings(model, n, dim)
    eng = F.Engine(rp, ci, dim, device=local)
    if world > 1:
        eng.set_option("multicast", int(os.environ.get("MC", "1")))
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.comm_peer_export())
        eng.comm_peer_init(blobs, rank, world)
    if model != 5:
        eng.set_lut()
    eng.set_embeddings(X0)
    batches = [int(x) for x in os.environ.get("BATCHES", "16384,65536").split(",")]
    chunks = [int(x) for x in os.environ.get("CHUNKS", "0").split(",")]
    for batch in batches:
        neg = g.epoch_negatives(model, n, batch, s, bs).copy()
        eng.set_negatives(neg)
        # variants with the flag barrier first: once it is off the ranks' step counters run free
        orders = [int(x) for x in os.environ.get("ORDERS", "0").split(",")]
        sigs = [int(x) for x in os.environ.get("SIGS", "2").split(",")]
        variants = [(0, sg, o, 0, c) for c in chunks for o in orders for sg in sigs]
        if batch == batches[-1] and os.environ.get("FREE", "1") == "1":
            variants += [(1, 2, orders[0], 0, chunks[-1]), (2, 2, orders[0], 0, chunks[-1]), (3, 2, orders[0], 0, chunks[-1])]
        for dbg, sig, persist, mode, chunk in variants:
            eng.set_option("peer_debug", dbg)
            eng.set_option("peer_sig", sig)
            eng.set_option("order", persist)
            eng.set_epoch_mode(mode)
            ms = []
            for it in range(6):
                eng.set_negative_offset(0)
                dist.barrier()
                torch.cuda.synchronize()
                eng.run_epoch(model, batch, s, bs, 0.02, chunk)
                ms.append(eng.last_epoch_ms())
                if dbg & 2:
                    dist.barrier()
            if os.environ.get("TRACE") == "1":
                eng.set_option("trace", 1)
                eng.set_negative_offset(0)
                dist.barrier()
                torch.cuda.synchronize()
                eng.run_epoch(model, batch, s, bs, 0.02, chunk)
                tr = eng.trace_ms()
                eng.set_option("trace", 0)
                allt = [None] * world
                dist.all_gather_object(allt, [round(float(x) * 1e3, 1) for x in tr])
                if rank == 0:
                    for r_, t_ in enumerate(allt):
                        nb_ = len(t_)
                        q_ = [round(sum(t_[k * nb_ // 4:(k + 1) * nb_ // 4]) / 1e3, 3) for k in range(4)]
                        print(json.dumps({"trace_rank": r_, "dbg": dbg, "order": persist, "chunk": chunk, "quartile_ms": q_,
                                          "first8_us": t_[:8], "last4_us": t_[-4:]}), flush=True)
            t = torch.tensor([min(ms[2:])], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"world": world, "B": batch, "peer_debug": dbg, "peer_sig": sig, "mode": mode, "chunk": chunk, "order": persist,
                                  "ms": float(t.item()), "rank0_ms": [round(x, 3) for x in ms]}), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

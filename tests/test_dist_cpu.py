"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- each rank plans only
the rows of every minibatch it owns (f2v_plan_build with rank/world, the same code the engine
uploads: contiguous slices for the NCCL all-gather exchange, the degree-balanced partition for
the peer-store exchange), computes them, and the rows are exchanged before the next minibatch.  The per-slice compute is stood in for by the oracle so the test runs without a GPU;
the result must equal the single-rank epoch bit for bit."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, model, bs, batch, ret):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from force2vec_b200 import host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rp, ci = host.rmat_csr(9, 8, 3)
    n = len(rp) - 1
    dim, s, lr = 16, 5, 0.02
    g = host.RandStream(1)                                     # every rank regenerates the same stream
    X = g.init_embeddings(model, n, dim)
    walks = g.walks(rp, ci) if model == 7 else None
    W = (batch + s - 1) if (bs and model != 7) else s
    neg = g.epoch_negatives(model, n, batch, s, bs).reshape(-1, W)
    plan = host.plan_build(rp, batch, 8, walk=(model == 7), rank=rank, world=world)
    nb = plan["nb"]
    slice_rows = batch // world
    pad = np.zeros((nb * batch, dim), np.float32)              # tables padded to whole minibatches
    pad[:n] = X
    for b in range(nb):
        items = plan["items"][int(plan["item_ptr"][b]):int(plan["item_ptr"][b + 1])]
        rows = np.unique(items["v"])
        lo, hi = b * batch, min(n, (b + 1) * batch)
        mine_lo = min(lo + rank * slice_rows, hi)
        mine_hi = min(mine_lo + slice_rows, hi)
        assert rows.tolist() == list(range(mine_lo, mine_hi))  # the plan is exactly the rank's slice
        full_idx = np.zeros(max(s * batch, 1) + s, np.uint32)
        full_idx[:W] = neg[b]
        Xb = pad[:n].copy()
        # vertex k of the minibatch uses idx[k..k+s-1]: shift the window to the slice origin
        off = (mine_lo - lo) if (bs and model != 7) else 0
        if mine_hi > mine_lo:
            O.step(model, bs, rp, ci, Xb, mine_lo, mine_hi, full_idx[off:], s, lr, walks=walks)
        send = torch.from_numpy(np.ascontiguousarray(
            np.vstack([Xb[mine_lo:mine_hi], np.zeros((slice_rows - (mine_hi - mine_lo), dim), np.float32)])))
        out = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(out, send)
        pad[lo:lo + batch] = torch.cat(out).numpy()
        pad[n:] = 0
    if rank == 0:
        ret["X"] = pad[:n].copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("model,bs", [(5, 0), (5, 1), (6, 0), (7, 0)])
def test_two_ranks_equal_one_rank(model, bs):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from force2vec_b200 import host
    world, batch = 2, 96
    port = 29500 + (os.getpid() + model * 7 + bs) % 2000
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, model, bs, batch, ret), nprocs=world, join=True)
    rp, ci = host.rmat_csr(9, 8, 3)
    want = O.run(model, bs, rp, ci, 16, 1, batch, 5, 0.02)["X"]
    assert np.array_equal(ret["X"], want)


def _worker_balanced(rank, world, port, model, batch, ret):
    """Peer-store exchange protocol: degree-balanced ownership, every rank ends up with every row."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from force2vec_b200 import host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rp, ci = host.rmat_csr(9, 8, 3)
    n = len(rp) - 1
    dim, s, lr = 16, 5, 0.02
    g = host.RandStream(1)
    X = g.init_embeddings(model, n, dim)
    walks = g.walks(rp, ci) if model == 7 else None
    neg = g.epoch_negatives(model, n, batch, s, 0).reshape(-1, s)
    plan = host.plan_build(rp, batch, 8, walk=(model == 7), rank=rank, world=world, assign=1)
    deg = np.diff(rp.astype(np.int64))
    loads = []
    for b in range(plan["nb"]):
        items = plan["items"][int(plan["item_ptr"][b]):int(plan["item_ptr"][b + 1])]
        rows = np.unique(items["v"]).astype(np.int64)
        lo, hi = b * batch, min(n, (b + 1) * batch)
        assert rows.size == 0 or (rows.min() >= lo and rows.max() < hi)
        mine = np.empty((len(rows), dim), np.float32)
        for k, v in enumerate(rows):                            # Jacobi: every row sees the pre-minibatch table
            T = X.copy()                                        # (bs=0: one-row steps share the minibatch's negatives)
            O.step(model, 0, rp, ci, T, int(v), int(v) + 1, neg[b], s, lr, walks=walks)
            mine[k] = T[v]
        got = [None] * world
        dist.all_gather_object(got, (rows, mine))
        seen = np.concatenate([r for r, _ in got])
        assert sorted(seen.tolist()) == list(range(lo, hi))     # a partition of the minibatch
        for r, vals in got:
            X[r] = vals
        loads.append([int(deg[r].sum() + 6 * len(r)) for r, _ in got])
    if rank == 0:
        ret["X"] = X.copy()
        ret["loads"] = loads
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("model", [5, 6, 7])
def test_balanced_partition_two_ranks(model):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from force2vec_b200 import host
    world, batch = 2, 96
    port = 31500 + (os.getpid() + model * 11) % 2000
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker_balanced, args=(world, port, model, batch, ret), nprocs=world, join=True)
    rp, ci = host.rmat_csr(9, 8, 3)
    want = O.run(model, 0, rp, ci, 16, 1, batch, 5, 0.02)["X"]
    assert np.array_equal(ret["X"], want)
    if model != 7:
        rp64 = rp.astype(np.int64)
        for b, l in enumerate(ret["loads"]):                    # greedy LPT: loads differ by at most one row's cost
            lo, hi = b * batch, min(len(rp) - 1, (b + 1) * batch)
            biggest = int(np.diff(rp64[lo:hi + 1]).max()) + 6
            assert abs(l[0] - l[1]) <= biggest, (b, l, biggest)


def test_bench_children_run_in_their_own_process_groups_and_never_cost_the_parent():
    """bench.py's N = 8 run measures BASELINE config 5 (R-MAT 26) in CHILD processes, one per rank, started by
    the ranks of the torch.distributed.run job after the headline measurement (bench.cfg5_extras / run_child).
    Under a real torch.distributed.run (gloo, world 2): a child forms its own process group on a fresh port (the
    agent store of the parent job must not leak into it), its line comes back, every child's exit code is reported;
    a crashing child, a child that fails on one rank, a child that hangs past its share of the wall-clock budget
    (killed) and a run that no longer fits the budget (not started) leave the parent job alive and exiting 0."""
    import json
    import subprocess
    port = 29400 + os.getpid() % 150
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "_child_spawn_worker.py")], capture_output=True, timeout=600)
    out = r.stdout.decode()
    assert r.returncode == 0, (out + r.stderr.decode())[-3000:]
    res = json.loads([ln for ln in out.splitlines() if ln.startswith("PARENT_RESULT ")][-1][len("PARENT_RESULT "):])
    assert "not_needed" not in res                               # a fall-back runs only after the run it stands in for failed
    for key in ("ok", "after_crash"):
        assert res[key]["pairs_per_s"] == 3.0 and res[key]["n_gpus"] == 2 and res[key]["child_rcs"] == [0, 0]
        assert res[key]["parity"] == {"bit_exact": True}
    assert "error" in res["crash"] and res["crash"]["child_rcs"] == [-6, -6]
    assert res["rank1_fails"]["child_rcs"] == [0, 7]
    assert res["hang"]["timed_out"] is True and res["hang"]["rc"] == -9
    assert "skipped" in res["late"]

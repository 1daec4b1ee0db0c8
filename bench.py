#!/usr/bin/env python
"""bench.py -- the reference's headline metric (force-pair updates / s, epoch time, fraction of
the HBM roofline) on BASELINE.json configs[1]: synthetic R-MAT scale-20, sForce2Vec (option 6),
d=128, one B200.  A "step" is one epoch = one pass of the hot path over every minibatch.

  python bench.py --gpus N --steps K --warmup W            our engine (N ranks under torchrun)
  python bench.py --impl reference ...                      the reference's own CPU code
                                                            (oracle/_ref), same config

value    = pair updates / s, inputs resident in HBM, CUDA events around K epochs, max over ranks
           (N > 1: replicated tables, every minibatch's rows dealt to the ranks, the exchange fused
           into the force kernel -- NVLink multicast or peer stores; --comm nccl = all-gather baseline,
           --sharded 1 = row-sharded tables)
e2e      = the same through f2v_run_epoch_host: pinned HOST table + sample stream in, HOST table
           out, every step (PCIe copies inside the timed region; N > 1: each rank moves its 1/N
           share of the table over its own PCIe link, the rest travels over NVLink)
roofline = algorithmic bytes per force-kernel launch / its average duration over the timed region
           (bytes per epoch = (nnz + n*s)*d*4 read + n*d*4 written)
cpu_baseline = the unmodified reference (oracle/_ref) on this box's host cores, one epoch sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

MODEL_NAMES = {5: "tForce2Vec", 6: "sForce2Vec", 7: "rForce2Vec"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--model", type=int, default=6, choices=[5, 6, 7])
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--nsamples", type=int, default=5)
    ap.add_argument("--bs", type=int, default=0)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0, help="engine epoch mode (0 per-minibatch launches, 1 persistent)")
    ap.add_argument("--variant", type=int, default=-1, help="d=128 kernel lane layout (f2v_set_option; -1 = auto)")
    ap.add_argument("--neg-smem", type=int, default=1)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N>1 exchange: peer = stores into the peers' replicas fused into the force kernel; "
                         "nccl = all-gather per minibatch (baseline)")
    ap.add_argument("--multicast", type=int, default=1, help="peer exchange through NVLink multicast (NVLS) stores")
    ap.add_argument("--sharded", type=int, default=0, help="N>1: row-sharded tables instead of replicas (capacity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extra-batches", default="256,4096,16384", help="comma list of additional batch sizes to report")
    return ap.parse_args()


def workload_name(a):
    return "rmat%d_ef%d_seed1 option%d(%s) d%d s%d bs%d B%d lr%g" % (
        a.scale, a.edge_factor, a.model, MODEL_NAMES[a.model], a.dim, a.nsamples, a.bs, a.batch, a.lr)


def pairs_per_epoch(a, n, nnz):
    return n * (5 + a.nsamples) if a.model == 7 else nnz + n * a.nsamples


def bytes_per_epoch(a, n, nnz):
    pairs = pairs_per_epoch(a, n, nnz)
    b = pairs * a.dim * 4 + n * a.dim * 4
    if a.model == 7:
        b += n * 5 * (8 + 4)
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=10)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic(a):
    """dram bytes per force-kernel launch from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    return t.get(workload_name(a))


def make_graph(a):
    from force2vec_b200 import host
    t = time.time()
    rp, ci = host.rmat_csr(a.scale, a.edge_factor, 1)
    return rp, ci, time.time() - t


# ------------------------------------------------------------------------------------------
class StdoutToStderr:
    """The reference prints progress on C stdout; keep our stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference(a, rp, ci, epochs, warm):
    """The unmodified reference (oracle/_ref) on this box's host cores.  Returns dict."""
    with StdoutToStderr():
        return _cpu_reference(a, rp, ci, epochs, warm)


def _cpu_reference(a, rp, ci, epochs, warm):
    from oracle import oracle as O
    n, nnz = len(rp) - 1, len(ci)
    cores = os.cpu_count() or 1
    if not O.ref_available():
        # the reference could not be compiled: time the oracle port instead
        t0 = time.time()
        O.run(a.model, a.bs, rp, ci, a.dim, 0, a.batch, a.nsamples, a.lr, threads=cores)
        t_init = time.time() - t0
        t0 = time.time()
        O.run(a.model, a.bs, rp, ci, a.dim, epochs, a.batch, a.nsamples, a.lr, threads=cores)
        sec = (time.time() - t0 - t_init) / epochs
        kind, what = "port", "oracle/f2v_oracle.c option %d" % a.model
    else:
        kind = "reference"
        # best CPU path for this model: the AVX-512 variants (options 8/9/10/11) where the host has
        # avx512f+dq and the dimension is one they implement, else the OpenMP scalar option itself
        avx_opt = {5: 11, 6: 9, 7: 10}[a.model]
        use_avx = O.host_has_avx512() and O.ref_available(avx512=True) and a.dim in (64, 128) and not a.bs
        opt = avx_opt if use_avx else a.model
        if warm:
            O.ref_run(opt, a.bs, rp, ci, a.dim, 0, a.batch, a.nsamples, a.lr, threads=cores, avx512=use_avx, want_X=False)
        # the reference's own timer spans init + epochs (algorithms.cpp:557,647): difference it out
        _, t_init = O.ref_run(opt, a.bs, rp, ci, a.dim, 0, a.batch, a.nsamples, a.lr, threads=cores, avx512=use_avx, want_X=False)
        _, t_run = O.ref_run(opt, a.bs, rp, ci, a.dim, epochs, a.batch, a.nsamples, a.lr, threads=cores, avx512=use_avx, want_X=False)
        sec = max(t_run - t_init, 1e-9) / epochs
        what = "reference option %d%s" % (opt, " (AVX-512)" if use_avx else " (OpenMP scalar)")
    pairs = pairs_per_epoch(a, n, nnz)
    return {"value": pairs / sec, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": "%d full epoch(s) of the same workload, %s, %d threads, init time differenced out"
                      % (epochs, what, cores),
            "epoch_s": sec}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rp, ci, _ = make_graph(a)
    n, nnz = len(rp) - 1, len(ci)
    epochs = max(1, min(a.steps, 3))         # bounded sample: at most 3 timed epochs
    r = cpu_reference(a, rp, ci, epochs, warm=a.warmup > 0)
    line = {"impl": "reference", "metric": "force_pair_updates_per_sec", "value": r["value"], "unit": "pairs/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["epoch_s"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "n": n, "nnz": nnz, "pairs_per_epoch": pairs_per_epoch(a, n, nnz),
                       "timed_epochs": epochs},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def timed_epochs(torch, dist, eng, a, K, neg_all, stride, first_epoch, world):
    """K epochs, device-timed on the engine's (= torch's current) stream; returns seconds (max over ranks)."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for k in range(K):
        eng.set_negative_offset((first_epoch + k) * stride)
        if a.model == 7:
            eng.sample_walks(1, first_epoch + k)
        eng.run_epoch(a.model, a.batch, a.nsamples, a.bs, a.lr, a.chunk)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sec = ev0.elapsed_time(ev1) / 1e3
    if world > 1:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


def run_ours(a):
    import torch
    import torch.distributed as dist
    import force2vec_b200 as F
    from force2vec_b200 import host
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rp, ci, t_graph = make_graph(a)
    n, nnz = len(rp) - 1, len(ci)
    pairs = pairs_per_epoch(a, n, nnz)
    K, W = a.steps, a.warmup
    total_epochs = W + K + (0 if a.no_e2e else W + K)

    g = host.RandStream(1)
    X0 = torch.empty((n, a.dim), dtype=torch.float32, pin_memory=True)
    g.init_embeddings(a.model, n, a.dim, out=X0.numpy())
    stride = host.neg_stream_len(a.model, n, a.batch, a.nsamples, a.bs)
    neg_all = torch.empty(max(total_epochs * stride, 1), dtype=torch.int32, pin_memory=True)
    neg_np = neg_all.numpy().view(np.uint32)
    for k in range(total_epochs):
        g.epoch_negatives(a.model, n, a.batch, a.nsamples, a.bs, out=neg_np[k * stride:(k + 1) * stride])

    eng = F.Engine(rp, ci, a.dim, device=local)
    # the engine launches on torch's current stream so that torch.cuda.Event brackets its kernels;
    # that must be a real (non-default) stream: handle 0 means "engine's own stream" in the C ABI
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    if a.mode:
        eng.set_epoch_mode(a.mode)
    eng.set_option("variant", a.variant)
    eng.set_option("neg_smem", a.neg_smem)
    if world > 1 and a.comm == "nccl":
        ids = [F.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)
    elif world > 1:
        eng.set_option("multicast", a.multicast)
        eng.set_option("sharded", a.sharded)
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.comm_peer_export())
        eng.comm_peer_init(blobs, rank, world)
    if a.model != 5:
        eng.set_lut()
    eng.set_embeddings(X0.numpy())
    eng.set_negatives(neg_np[:(W + K) * stride])
    eng.sync()

    # ---- resident-input throughput ("value")
    timed_epochs(torch, dist, eng, a, W, neg_all, stride, 0, world)          # warm-up (plan build, clocks)
    l0 = eng.launch_count()
    cs = ClockSampler(local)
    cs.__enter__()                                   # sampled across the value AND the e2e timed regions
    sec = timed_epochs(torch, dist, eng, a, K, neg_all, stride, W, world)
    launches = eng.launch_count() - l0
    epoch_s = sec / K
    value = pairs / epoch_s

    # ---- roofline of the force kernel (the only kernel of an option 5/6 epoch)
    peak, peak_src = measured_peak()
    nb = (n + a.batch - 1) // a.batch
    alg_bytes_epoch = bytes_per_epoch(a, n, nnz)
    achieved = alg_bytes_epoch / epoch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": committed_traffic(a), "kernel": "f2v::force_batch_kernel",
                "algorithmic_bytes_per_launch": alg_bytes_epoch / nb / max(world, 1),
                "avg_launch_us": epoch_s / nb * 1e6, "peak_source": peak_src}

    # ---- end to end through the host-buffer call
    e2e = None
    if not a.no_e2e:
        Xin, Xout = X0, torch.empty(X0.shape, dtype=torch.float32, pin_memory=True)   # empty_like would not pin
        assert Xin.is_pinned() and Xout.is_pinned()
        base = W + K

        def e2e_step(k):
            eng.run_epoch_host(a.model, a.batch, a.nsamples, a.bs, a.lr, X_in=Xin.numpy(),
                               neg=neg_np[(base + k) * stride:(base + k + 1) * stride], X_out=Xout.numpy(), chunk=a.chunk)
        if a.model == 7:
            eng.sample_walks(1, 0)
        for k in range(W):
            e2e_step(k)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            e2e_step(W + k)
            Xin, Xout = Xout, Xin                   # next epoch starts from this epoch's host result
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_sec = t1 - t0
        if world > 1:
            t = torch.tensor([e2e_sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_sec = float(t.item())
        # bytes over PCIe per step, whole job: with the peer exchange every rank moves 1/world of the
        # table each way (NCCL mode: every rank moves the whole table)
        tbl = n * a.dim * 4
        copies = 1 if (world == 1 or (a.comm == "peer" and not a.sharded)) else world
        e2e = {"value": pairs / (e2e_sec / K), "unit": "pairs/s",
               "h2d_bytes_per_step": int(tbl * copies + stride * 4 * world), "d2h_bytes_per_step": int(tbl * copies),
               "ms_per_step": e2e_sec / K * 1e3,
               "call": "f2v_run_epoch_host (pinned host table in/out" +
                       ("" if world == 1 else "; each rank moves its 1/%d share over PCIe, the rest over NVLink" % world
                        if a.comm == "peer" else "; every rank moves the whole table") + ")"}

    cs.__exit__()
    clocks = cs.summary()

    # ---- extra batch sizes (reported, not the headline)
    extra = {}
    for bsz in [int(x) for x in a.extra_batches.split(",") if x]:
        b = argparse.Namespace(**vars(a))
        b.batch = bsz
        st = host.neg_stream_len(b.model, n, bsz, b.nsamples, b.bs)
        nn = np.empty(max(4 * st, 1), np.uint32)
        for k in range(4):
            g.epoch_negatives(b.model, n, bsz, b.nsamples, b.bs, out=nn[k * st:(k + 1) * st])
        eng.set_negatives(nn)
        timed_epochs(torch, dist, eng, b, 2, None, st, 0, world)
        s2 = timed_epochs(torch, dist, eng, b, 2, None, st, 2, world) / 2
        extra["B%d" % bsz] = {"epoch_ms": s2 * 1e3, "pairs_per_s": pairs / s2,
                              "roofline_frac": bytes_per_epoch(b, n, nnz) / s2 / 1e9 / peak}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(a, rp, ci, 1, warm=False)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "force_pair_updates_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": epoch_s * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "n": n, "nnz": nnz, "pairs_per_epoch": pairs,
                           "minibatches_per_epoch": nb, "epoch_mode": a.mode,
                           "l2": "inputs larger than L2: 2 x %.0f MiB tables + %.0f MiB CSR vs 126 MB L2; no flush needed"
                                 % (n * a.dim * 4 / 2**20, nnz * 4 / 2**20),
                           "init": "glibc-compatible srand(1) stream (reference order)",
                           "parallelism": "replicated table, minibatch split over %d rank(s)%s" %
                                          (world, "" if world == 1 else (", NCCL all-gather per minibatch" if a.comm == "nccl" else
                                                      (", row-sharded tables (1/%d of the rows per GPU, remote gathers over NVLink"
                                                       " + flag barrier per minibatch)" % world) if a.sharded else
                                                      ", rows stored into the peers' replicas from the force kernel "
                                                      "(NVLink multicast / peer stores + flag barrier per minibatch)"))},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "epoch_s": epoch_s, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()                  # nobody unmaps a table a peer may still be storing into
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    # stdout carries exactly one JSON line: anything libraries print on fd 1 (NCCL's version banner,
    # the reference's progress lines) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

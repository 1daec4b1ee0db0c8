#!/usr/bin/env python
"""bench.py -- the reference's headline metric (force-pair updates / s, epoch time, fraction of the
HBM roofline) on the configuration BASELINE.json quotes "1/2/4/8 B200" on: configs[3], synthetic
R-MAT scale-24, tForce2Vec (option 5) with per-vertex negatives (bs=1), d=128.  It fits one GPU
(2 x 8 GiB tables + 2 GiB CSR).  A "step" is one epoch = one pass of the hot path over every minibatch.

  python bench.py --gpus N --steps K --warmup W            our engine (N ranks under torchrun)
  python bench.py --impl reference ...                      the reference's own CPU code (oracle/_ref)
  python bench.py --workload cfg2|cfg3|cfg4|cfg5            the other BASELINE configs (cfg5: 8 GPUs)

value    = pair updates / s, inputs resident in HBM, CUDA events around K epochs, max over ranks
           (N > 1: replicated tables, every minibatch's rows dealt to the ranks, the exchange fused into
           the force kernel -- NVLink multicast or peer stores; --comm nccl = all-gather baseline,
           --sharded 1 = row-sharded tables).  At N > 1 one epoch is first CHECKED: every rank compares
           its replica (device checksum + probe rows) with a single-GPU engine run from the same state
           on its own device -- bit for bit -- and the line carries "parity"; a mismatch exits non-zero.
e2e      = the same through f2v_run_epoch_host: pinned HOST table + sample stream in, HOST table out,
           every step (PCIe copies inside the timed region; N > 1: each rank moves its 1/N share of the
           table over its own PCIe link, the rest travels over NVLink).  The e2e epochs upload the same
           W + K negative streams the resident epochs used (one per step, from pinned host memory).
roofline = algorithmic bytes per force-kernel launch / its average duration over the timed region
           (bytes per epoch = (nnz + n*s)*d*4 read + n*d*4 written), PER GPU at N > 1 (a
           rank's launch processes 1/N of the minibatch's pairs; the peak is one GPU's); frac_dram = DRAM
           bytes measured by ncu for this workload and N (profiles/traffic.json) / epoch time / peak
cpu_baseline = the unmodified reference (oracle/_ref) on this box's host cores, bounded sample
extra    = N = 1: the other BASELINE configs that fit one GPU (cfg2, cfg2 at batch 256, cfg3, cfg1 through
           f2v_train).  N = 8: BASELINE configs[4] -- R-MAT scale 26, which the reference cannot address --
           measured in CHILD processes (one `bench.py --workload cfg5` per rank, own process group, killed
           with their session at the end of a wall-clock budget: cfg5_extras), replicated and row-sharded,
           each with the checked epoch; a child's failure is an {"error": ...} entry, never a lost line.
"""
import argparse
import json
import os
import signal
import subprocess
import sys
import threading
import time

T_START = time.time()            # the wall-clock budget of the N = 8 extras is counted from here
_exit = os._exit                 # (a name the dry-run test can replace)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

MODEL_NAMES = {5: "tForce2Vec", 6: "sForce2Vec", 7: "rForce2Vec"}
# BASELINE.json configs[1..4]; batch = the throughput-optimal minibatch (the CLI's -batch is free;
# the reference's README batch 256 is reported under "extra")
WORKLOADS = {
    "cfg2": dict(scale=20, model=6, dim=128, bs=0, batch=65536),
    "cfg3": dict(scale=22, model=7, dim=64, bs=0, batch=65536),
    "cfg4": dict(scale=24, model=5, dim=128, bs=1, batch=262144),
    "cfg5": dict(scale=26, model=5, dim=128, bs=0, batch=262144),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=int)
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--model", type=int, choices=[5, 6, 7])
    ap.add_argument("--dim", type=int)
    ap.add_argument("--batch", type=int)
    ap.add_argument("--nsamples", type=int, default=5)
    ap.add_argument("--bs", type=int)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0, help="engine epoch mode (0 per-minibatch launches, 2 one dataflow launch per epoch)")
    ap.add_argument("--variant", type=int, default=-1, help="kernel lane layout (f2v_set_option; -1 = auto)")
    ap.add_argument("--neg-smem", type=int, default=1)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N>1 exchange: peer = stores into the peers' replicas fused into the force kernel; "
                         "nccl = all-gather per minibatch (baseline)")
    ap.add_argument("--multicast", type=int, default=1,
                    help="peer exchange through NVLink multicast (NVLS) stores: 1 = from three ranks on (two ranks: one store "
                         "per peer is as cheap and skips the loop through the switch), 0 = never, 3 = always")
    ap.add_argument("--sharded", type=int, default=0, help="N>1: row-sharded tables instead of replicas (capacity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra lines (other configs / batch sizes)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the checked epoch (timing probes only)")
    a = ap.parse_args()
    for k, v in WORKLOADS[a.workload].items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    return a


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


def config_of(a, n, nnz):
    """The same dict from both arms (the driver compares them)."""
    return {"workload": workload_name(a), "n": int(n), "nnz": int(nnz), "pairs_per_epoch": int(pairs_per_epoch(a, n, nnz)),
            "minibatches_per_epoch": int((n + a.batch - 1) // a.batch)}


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


def committed_capture(a, world):
    """The committed ncu capture for exactly this workload and world size (profiles/traffic.json):
    DRAM bytes per epoch and per launch, the kernel layout it was taken on, the gather-only ceiling."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    for e in json.load(open(p)).get("captures", []):
        if e.get("workload") == workload_name(a) and int(e.get("n_gpus", 1)) == world:
            return e
    return None


def make_graph(a, dist=None, rank=0):
    """R-MAT CSR, built once per box (page-cache copy under /dev/shm shared by the ranks)."""
    from force2vec_b200 import host
    t = time.time()
    if dist is not None and rank != 0:
        dist.barrier()
    rp, ci = host.rmat_csr_cached(a.scale, a.edge_factor, 1)
    if dist is not None and rank == 0:
        dist.barrier()
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
    rp, ci = np.ascontiguousarray(rp), np.ascontiguousarray(ci)
    n, nnz = len(rp) - 1, len(ci)
    cores = usable_cores()
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
        # (bs=1 has no AVX-512 variant in the reference)
        avx_opt = {5: 11, 6: 9, 7: 10}[a.model]
        use_avx = O.host_has_avx512() and O.ref_available(avx512=True) and a.dim in (64, 128) and not a.bs
        opt = avx_opt if use_avx else a.model
        if warm:
            O.ref_run(opt, a.bs, rp, ci, a.dim, 0, a.batch, a.nsamples, a.lr, threads=cores, avx512=use_avx, want_X=False)
        # the reference's own timer spans init + epochs (algorithms.cpp:557,647): difference it out
        _, t_init = O.ref_run(opt, a.bs, rp, ci, a.dim, 0, a.batch, a.nsamples, a.lr, threads=cores, avx512=use_avx, want_X=False)
        _, t_run = O.ref_run(opt, a.bs, rp, ci, a.dim, epochs, a.batch, a.nsamples, a.lr, threads=cores, avx512=use_avx, want_X=False)
        sec = max(t_run - t_init, 1e-9) / epochs
        what = "reference option %d%s%s" % (opt, " bs=1" if a.bs else "", " (AVX-512)" if use_avx else " (OpenMP scalar)")
    pairs = pairs_per_epoch(a, n, nnz)
    return {"value": pairs / sec, "unit": "pairs/s", "cores": cores, "kind": kind, "what": what, "epoch_s": sec,
            "epochs": epochs, "n": n}


def run_reference(a):
    """The reference's own CPU implementation of the path, all host threads, on OUR arm's config.
    One timed epoch (the reference's timer spans init + epochs, so a run with 0 epochs is
    differenced out): at R-MAT 24 that is minutes of CPU work -- the bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if (1 << a.scale) * a.dim >= 2**32:
        # sample/algorithms.h:40,68 and algorithms.cpp:593: 32-bit `i*DIM` -- the reference cannot address this table
        print(json.dumps({"impl": "reference", "unavailable": "the reference indexes the embedding table with 32-bit i*DIM: "
                          "n*dim = 2^%d is out of its range (sample/algorithms.h:40,68)" % (a.scale + (a.dim - 1).bit_length())}), flush=True)
        return
    rp, ci, _ = make_graph(a)
    n, nnz = len(rp) - 1, len(ci)
    epochs = 1 if (a.scale >= 22 or a.steps < 2) else min(a.steps, 3)
    r = cpu_reference(a, rp, ci, epochs, warm=False)
    sample = ("%d full epoch(s) of this workload, %s, %d threads; the reference times init + epochs "
              "(algorithms.cpp:557,647), a 0-epoch run is differenced out" % (epochs, r["what"], r["cores"]))
    line = {"impl": "reference", "metric": "force_pair_updates_per_sec", "value": r["value"], "unit": "pairs/s",
            "n_gpus": a.gpus, "steps": epochs, "warmup": 0, "requested": {"steps": a.steps, "warmup": a.warmup},
            "ms_per_step": r["epoch_s"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(a, n, nnz),
            "cpu_baseline": {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def timed_epochs(torch, dist, eng, a, K, stride, first_epoch, world):
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


def shared_init(a, n, host, dist, rank):
    """N > 1: the n x dim initial table off the glibc-compatible stream (reference order).  Rank 0 draws
    it (all host threads, jump-ahead) into a page-cache file, the others map the same pages.  Returns
    the table and, on rank 0, the generator positioned after the initial draws (the negatives follow)."""
    path = "/dev/shm/f2v_X0_rmat%d_m%d_d%d.npy" % (a.scale, a.model, a.dim)
    g = None
    if rank == 0:
        g = host.RandStream(1)
        X = np.lib.format.open_memmap(path + ".tmp.npy", mode="w+", dtype=np.float32, shape=(n, a.dim))
        g.init_embeddings(a.model, n, a.dim, out=X)
        X.flush()
        del X
        os.replace(path + ".tmp.npy", path)
    dist.barrier()
    return np.load(path, mmap_mode="r"), g


def make_engine(F, a, rp, ci, local, stream=None):
    eng = F.Engine(rp, ci, a.dim, device=local)
    if stream is not None:
        eng.set_stream(stream)
    if a.mode:
        eng.set_epoch_mode(a.mode)
    eng.set_option("variant", a.variant)
    eng.set_option("neg_smem", a.neg_smem)
    if a.model != 5:
        eng.set_lut()
    return eng


def checked_epoch_single(F, a, rp, ci, local, X0, neg0):
    """N > 1, first half of the checked epoch: one epoch from the initial state on a SINGLE-GPU engine on
    this rank's device (created, run and destroyed before the N-rank engine exists: at scale 26 the two
    would not fit one GPU together).  Returns what the N-rank result is compared with."""
    chunk = a.chunk or 64
    n = len(rp) - 1
    probe = sorted(set(int(x) for x in np.linspace(0, n - 1, 64)))
    single = make_engine(F, a, rp, ci, local)
    single.set_embeddings(X0)
    single.set_negatives(neg0)
    if a.model == 7:
        single.sample_walks(1, 0)
    single.run_epoch(a.model, a.batch, a.nsamples, a.bs, a.lr, chunk)
    h1 = single.checksum()
    rows1 = np.stack([single.get_rows(v, 1)[0] for v in probe])
    ms = single.last_epoch_ms()
    single.close()
    return {"chunk": chunk, "probe": probe, "checksum": h1, "rows": rows1, "epoch_ms": ms}


def checked_epoch_multi(a, ref, X0, neg0, multi, dist, torch):
    """Second half: the same epoch on the N-rank engine, same hub chunk length: every replica (or the
    row-sharded table) must equal the single-GPU table bit for bit -- device checksum over the whole
    table + probe rows compared value by value, on every rank."""
    multi.set_embeddings(X0)
    multi.set_negatives(neg0)
    if a.model == 7:
        multi.sample_walks(1, 0)
    multi.sync()
    dist.barrier()                       # the ranks' uploads take different times (32 GiB each at scale 26): start the
                                         # epoch together instead of spending the exchange time-out on the skew
    multi.run_epoch(a.model, a.batch, a.nsamples, a.bs, a.lr, ref["chunk"])
    hN = multi.checksum()
    rowsN = np.stack([multi.get_rows(v, 1)[0] for v in ref["probe"]])
    same = (ref["checksum"] == hN) and np.array_equal(ref["rows"], rowsN)
    t = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"vs": "single_gpu", "bit_exact": bool(int(t.item()) == 1), "checksum": "%016x" % ref["checksum"],
            "checksum_this_rank": "%016x" % hN, "epochs": 1, "chunk": ref["chunk"], "probe_rows": len(ref["probe"]),
            "single_gpu_epoch_ms": ref["epoch_ms"],
            "checked_on": "every rank's table against a single-GPU engine run on the same device"}


def parallelism_of(a, world):
    if world == 1:
        return "one GPU"
    if a.comm == "nccl":
        return "minibatch cut into %d contiguous slices, replicated tables, NCCL all-gather per minibatch (baseline)" % world
    head = "minibatch split over %d ranks (degree-balanced row ownership), " % world
    if a.sharded:
        return head + ("row-sharded tables: 1/%d of the rows of both tables per GPU in one flat VMM range, remote rows "
                       "gathered over NVLink, finished rows stored into their shard, flag barrier per minibatch" % world)
    return head + ("replicated tables, finished rows stored into every replica from the force kernel (NVLink multicast / "
                   "peer stores) + flag barrier per minibatch")


def free_port():
    import socket
    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    return port


def run_child(argv, port, timeout_s, script=None):
    """A fresh `python bench.py <argv>` on this rank's GPU with this rank's RANK / LOCAL_RANK / WORLD_SIZE and a
    rendezvous of its own (the children's rank 0 hosts a new store on `port`; torch.distributed.run's agent store
    already holds this job's keys, so the TORCHELASTIC_* variables must not reach the child).  Own session: a
    child that is still running at `timeout_s` is killed with everything it started.  Whatever the child does --
    a crash, a hang, a device fault -- stays in the child.  Returns dict(rc, timed_out, wall_s, line, stderr_tail)."""
    env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC_")}
    env["MASTER_PORT"] = str(port)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    # the child's NCCL only carries barriers and one-element reductions: keep its communicator off the NVLink
    # multicast (NVLS) resources, which the parent's communicator and the engine's own multicast object use
    env.setdefault("NCCL_NVLS_ENABLE", "0")
    t0 = time.time()
    res = {"rc": None, "timed_out": False, "line": None}
    try:
        proc = subprocess.Popen([sys.executable, script or os.path.join(ROOT, "bench.py")] + list(argv), env=env,
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, start_new_session=True)
        try:
            out, err = proc.communicate(timeout=max(1.0, timeout_s))
        except subprocess.TimeoutExpired:
            res["timed_out"] = True
            try:
                os.killpg(proc.pid, signal.SIGKILL)
            except OSError:
                pass
            out, err = proc.communicate()
        res["rc"] = proc.returncode
        for ln in out.decode(errors="replace").splitlines():
            if ln.startswith("{"):
                try:
                    res["line"] = json.loads(ln)
                except ValueError:
                    pass
        res["stderr_tail"] = err.decode(errors="replace")[-600:]
    except Exception as ex:            # the child machinery must never cost the parent its line
        res["stderr_tail"] = repr(ex)
    res["wall_s"] = round(time.time() - t0, 1)
    return res


# BASELINE configs[4] (R-MAT scale 26, 67 M vertices, d=128, 8 x B200: n*d = 2^33, which the reference's 32-bit
# i*DIM cannot address) measured by the driver's own N = 8 run: after the headline workload every rank starts a
# child bench.py for it on its GPU -- replicated tables with the fused exchange, then row-sharded tables -- each
# with bench.py's checked epoch (8-GPU table == single-GPU table, device checksum of all 32 GiB + probe rows).
# (key, arguments, seconds the run needs, key of the run whose FAILURE this one is the fall-back for)
CFG5_RUNS = (("cfg5_rmat26_replicated", ["--workload", "cfg5", "--steps", "3", "--warmup", "3", "--no-e2e"], 270, None),
             ("cfg5_rmat26_replicated_unicast", ["--workload", "cfg5", "--steps", "3", "--warmup", "3", "--no-e2e", "--multicast", "0"],
              200, "cfg5_rmat26_replicated"),
             ("cfg5_rmat26_row_sharded", ["--workload", "cfg5", "--sharded", "1", "--steps", "2", "--warmup", "3"], 220, None))


def cfg5_extras(dist, rank, world, runs=CFG5_RUNS, budget_s=None, script=None):
    """Returns {key: summary} on every rank (the line is read from rank 0's child).  Every child gets what is left
    of the wall-clock budget (counted from this process's start; the driver allows 870 s per N) minus a margin,
    and is not started at all if less than its `need` seconds are left: the extras can never cost the headline."""
    budget_s = float(os.environ.get("F2V_BENCH_BUDGET_S", "720")) if budget_s is None else budget_s
    out = {}
    for key, argv, need, fallback_for in runs:
        msg = [None]
        if rank == 0:
            left = budget_s - (time.time() - T_START)
            wanted = fallback_for is None or "error" in out.get(fallback_for, {})
            msg[0] = {"wanted": wanted, "go": left >= need, "timeout": min(left - 30.0, 600.0), "port": free_port(), "left": round(left, 1)}
        dist.broadcast_object_list(msg, src=0)
        m = msg[0]
        if not m["wanted"]:
            continue
        if not m["go"]:
            out[key] = {"skipped": "%.0f s of the wall-clock budget left, this run needs about %d s" % (m["left"], need)}
            continue
        r = run_child(list(argv) + ["--gpus", str(world), "--no-extra", "--no-cpu-baseline"], m["port"], m["timeout"], script)
        rcs = [None] * world
        dist.all_gather_object(rcs, r["rc"])             # the line comes from rank 0's child: report every child's exit code
        ln = r["line"]
        if ln and "error" not in ln and ln.get("value"):
            rf = ln.get("roofline") or {}
            st = ln.get("setup") or {}
            out[key] = {"workload": ln["config"]["workload"], "n": ln["config"]["n"], "nnz": ln["config"]["nnz"],
                        "n_gpus": ln["n_gpus"], "steps": ln["steps"], "warmup": ln["warmup"],
                        "epoch_ms": ln["ms_per_step"], "pairs_per_s": ln["value"], "parity": ln.get("parity"),
                        "device_memory_used_GiB_per_gpu": st.get("device_memory_used_GiB"),
                        "graph_build_s": st.get("graph_build_s"), "parallelism": st.get("parallelism"),
                        "frac_algorithmic_per_gpu": rf.get("frac_algorithmic"), "clocks": ln.get("clocks"),
                        "gpu_launches": ln.get("gpu_launches"), "e2e": ln.get("e2e"), "child_wall_s": r["wall_s"], "child_rcs": rcs,
                        "how": "child `bench.py %s --gpus %d` on every rank's GPU, started by the N = %d run" %
                               (" ".join(argv), world, world)}
        else:
            out[key] = {"error": "child did not produce a line", "rc": r["rc"], "child_rcs": rcs, "timed_out": r["timed_out"],
                        "child_wall_s": r["wall_s"], "line": ln, "stderr_tail": r.get("stderr_tail")}
    return out


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
    rp, ci, t_graph = make_graph(a, dist if world > 1 else None, rank)
    n, nnz = len(rp) - 1, len(ci)
    pairs = pairs_per_epoch(a, n, nnz)
    K, W = a.steps, a.warmup
    if world > 1 and a.sharded:
        a.no_e2e = True          # the host-buffer call of a row-sharded engine takes the FULL table on every rank
                                 # (rows are placed by hash): world x table bytes of host memory -- not benchmarked
    total_epochs = W + K         # the end-to-end epochs re-use these streams (each still uploads its own, every step)

    # ---- inputs: the reference's own stream order (init, then per epoch the negatives)
    per = -(-n // world)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)      # this rank's share of the host table (e2e)
    if world == 1:
        g = host.RandStream(1)
        X0t = torch.empty((n, a.dim), dtype=torch.float32, pin_memory=True)
        X0 = X0t.numpy()
        g.init_embeddings(a.model, n, a.dim, out=X0)
    else:
        X0, g = shared_init(a, n, host, dist, rank)
    stride = host.neg_stream_len(a.model, n, a.batch, a.nsamples, a.bs)
    neg_all = torch.empty(max(total_epochs * stride, 1), dtype=torch.int32, pin_memory=True)
    neg_np = neg_all.numpy().view(np.uint32)
    if rank == 0:
        # rank 0 owns the serial stream (as thread 0 does in f2v_train_gpus); the others receive the draws
        for k in range(total_epochs):
            g.epoch_negatives(a.model, n, a.batch, a.nsamples, a.bs, out=neg_np[k * stride:(k + 1) * stride])
    if world > 1:
        path = "/dev/shm/f2v_neg_rmat%d_m%d_B%d_bs%d_e%d.npy" % (a.scale, a.model, a.batch, a.bs, total_epochs)
        if rank == 0:
            np.save(path + ".tmp.npy", neg_np)
            os.replace(path + ".tmp.npy", path)
        dist.barrier()
        if rank != 0:
            neg_np[:] = np.load(path, mmap_mode="r")

    parity_ref = None
    if world > 1 and not a.no_parity:
        parity_ref = checked_epoch_single(F, a, rp, ci, local, X0, neg_np[:stride])

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    # the engine launches on torch's current stream so that torch.cuda.Event brackets its kernels;
    # that must be a real (non-default) stream: handle 0 means "engine's own stream" in the C ABI
    eng = make_engine(F, a, rp, ci, local, stream.cuda_stream)
    if world > 1 and a.comm == "nccl":
        ids = [F.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)
    elif world > 1:
        eng.set_option("multicast", a.multicast)
        eng.set_option("sharded", a.sharded)
        eng.set_option("exchange_timeout_ms", 120000)       # (default 30 s) host-side skew between ranks is not a failure
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.comm_peer_export())
        eng.comm_peer_init(blobs, rank, world)

    parity = None
    if parity_ref is not None:
        parity = checked_epoch_multi(a, parity_ref, X0, neg_np[:stride], eng, dist, torch)
        if not parity["bit_exact"]:
            if rank == 0:
                print(json.dumps({"error": "multi-GPU table differs from the single-GPU run", "parity": parity}), flush=True)
            dist.barrier()
            sys.exit(3)
    eng.set_embeddings(X0)
    eng.set_negatives(neg_np[:(W + K) * stride])
    eng.sync()
    free_b, total_b = eng.device_memory()

    # ---- resident-input throughput ("value")
    timed_epochs(torch, dist, eng, a, W, stride, 0, world)          # warm-up (plan build, clocks)
    l0 = eng.launch_count()
    cs = ClockSampler(local)
    cs.__enter__()                                   # sampled across the value AND the e2e timed regions
    sec = timed_epochs(torch, dist, eng, a, K, stride, W, world)
    launches = eng.launch_count() - l0
    epoch_s = sec / K
    value = pairs / epoch_s

    # ---- roofline of the force kernel (the only kernel of an option 5/6 epoch)
    peak, peak_src = measured_peak()
    nb = (n + a.batch - 1) // a.batch
    alg_bytes_epoch = bytes_per_epoch(a, n, nnz)
    # per GPU: a rank's launch processes 1/world of the minibatch's pairs (degree-balanced ownership) against ONE
    # GPU's HBM peak -- the whole-job figure divided by one GPU's peak would not be a fraction of anything
    achieved = alg_bytes_epoch / max(world, 1) / epoch_s / 1e9
    cap = committed_capture(a, world)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "f2v::force_batch_kernel",
                "algorithmic_bytes_per_launch": alg_bytes_epoch / nb / max(world, 1),
                "avg_launch_us": epoch_s / nb * 1e6, "peak_source": peak_src,
                "frac_algorithmic": achieved / peak, "frac_dram": None, "frac_of_gather_ceiling": None,
                "per": "GPU (each rank's launch: 1/%d of the minibatch's pairs; peak = one GPU's HBM)" % world if world > 1 else "GPU",
                "achieved_whole_job": alg_bytes_epoch / epoch_s / 1e9,
                "note": "frac = algorithmic bytes (every gathered row billed to HBM, SURVEY 8(d)) / time / peak: it exceeds "
                        "1 where gathered hub rows hit in the 126 MB L2; frac_dram = DRAM bytes ncu measured for this "
                        "workload and N / time / peak is the utilisation"}
    if cap:
        roofline["traffic"] = cap["dram_bytes_per_epoch"] / cap["launches_per_epoch"]
        roofline["frac_dram"] = cap["dram_bytes_per_epoch"] / epoch_s / 1e9 / peak
        roofline["traffic_source"] = cap.get("source")
        roofline["traffic_layout"] = cap.get("layout")
        if cap.get("gather_ceiling_ms"):
            roofline["frac_of_gather_ceiling"] = cap["gather_ceiling_ms"] / (epoch_s * 1e3)

    # ---- end to end through the host-buffer call
    e2e = None
    if not a.no_e2e:
        if world == 1:
            Xin = X0
            Xout = torch.empty((n, a.dim), dtype=torch.float32, pin_memory=True).numpy()   # empty_like would not pin
            regs = []
        else:
            # full-size (virtual) buffers; only this rank's row range is touched and page-locked
            Xin, Xout = np.empty((n, a.dim), np.float32), np.empty((n, a.dim), np.float32)
            Xin[lo:hi] = X0[lo:hi]
            Xout[lo:hi] = 0
            regs = [Xin[lo:hi], Xout[lo:hi]] if hi > lo else []
            for r in regs:
                F.capi.check(F.lib().f2v_host_register(r.ctypes.data, r.nbytes), "f2v_host_register")
        def e2e_step(k):
            eng.run_epoch_host(a.model, a.batch, a.nsamples, a.bs, a.lr, X_in=Xin,
                               neg=neg_np[k * stride:(k + 1) * stride], X_out=Xout, chunk=a.chunk)
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
        for r in regs:
            F.lib().f2v_host_unregister(r.ctypes.data)
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
    if world > 1:
        dist.barrier()                  # nobody unmaps a table a peer may still be storing into
    eng.close()
    del eng

    def make_line(extra, cpu):
        return {"metric": "force_pair_updates_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": epoch_s * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(a, n, nnz),
                "setup": {"epoch_mode": a.mode, "graph_build_s": round(t_graph, 1),
                          "l2": "inputs larger than L2: 2 x %.0f MiB tables + %.0f MiB CSR vs 126 MB L2; no flush needed"
                                % (n * a.dim * 4 / 2**20, nnz * 4 / 2**20),
                          "init": "glibc-compatible srand(1) stream (reference order)",
                          "device_memory_used_GiB": round((total_b - free_b) / 2**30, 1),
                          "parallelism": parallelism_of(a, world)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "epoch_s": epoch_s, "parity": parity, "extra": extra}
    emitted = threading.Lock()               # whoever takes it first prints the ONE line

    def emit(extra, cpu=None):
        if emitted.acquire(blocking=False) and rank == 0:
            print(json.dumps(make_line(extra, cpu)), flush=True)

    # ---- extra lines (N = 1 only; reported, not the headline): the other BASELINE configs that fit
    # one GPU, the reference's README batch, and the reference's README command through f2v_train
    extra = {}
    if world == 1 and not a.no_extra:
        try:
            extra = extra_lines(torch, F, host, a, peak)
        except Exception as ex:            # an extra must never cost the headline
            extra = {"error": repr(ex)}
    if world == 8 and not a.no_extra and a.workload != "cfg5":
        # the 8-GPU configuration of BASELINE (R-MAT 26), in child processes, inside what is left of the time budget
        torch.cuda.empty_cache()
        # last line of defence: should the extras' own machinery hang (a collective between the parents after a rank
        # died, say), every rank leaves on its own shortly after the budget -- rank 0 with the headline line printed
        budget = float(os.environ.get("F2V_BENCH_BUDGET_S", "720"))
        stop_watchdog = threading.Event()

        def watchdog():
            while not stop_watchdog.wait(1.0):
                if time.time() - T_START > budget + 45.0:
                    emit({"error": "the scale-26 extras did not return inside the wall-clock budget (watchdog)"})
                    sys.stdout.flush()
                    _exit(0)
        threading.Thread(target=watchdog, daemon=True).start()
        try:
            extra = cfg5_extras(dist, rank, world, budget_s=budget)
        except Exception as ex:
            extra = {"error": repr(ex)}
        stop_watchdog.set()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = bounded_cpu_baseline(a, host)

    emit(extra, cpu)
    if world > 1:
        try:                               # the line is out: nothing after it may turn the run into a failure
            dist.barrier()
            dist.destroy_process_group()
        except Exception as ex:
            print("bench.py: shutdown of the process group failed: %r" % (ex,), file=sys.stderr)


def quick_epochs(torch, eng, a, host, g, n, epochs=3):
    """best-of epoch time (ms) of a resident workload (engine's own events)."""
    st = host.neg_stream_len(a.model, n, a.batch, a.nsamples, a.bs)
    neg = g.epoch_negatives(a.model, n, a.batch, a.nsamples, a.bs).copy()
    eng.set_negatives(neg)
    ms = []
    for k in range(epochs):
        eng.set_negative_offset(0)
        if a.model == 7:
            eng.sample_walks(1, k)
        eng.run_epoch(a.model, a.batch, a.nsamples, a.bs, a.lr, a.chunk)
        ms.append(eng.last_epoch_ms())
    return min(ms[1:]) if len(ms) > 1 else ms[0], st


def extra_lines(torch, F, host, a, peak):
    out = {}
    for name in ("cfg2", "cfg3"):
        if name == a.workload:
            continue
        b = argparse.Namespace(**vars(a))
        for k, v in WORKLOADS[name].items():
            setattr(b, k, v)
        rp, ci = host.rmat_csr_cached(b.scale, b.edge_factor, 1)
        n, nnz = len(rp) - 1, len(ci)
        g = host.RandStream(1)
        X0 = g.init_embeddings(b.model, n, b.dim)
        with make_engine(F, b, rp, ci, torch.cuda.current_device()) as eng:
            eng.set_embeddings(X0)
            batches = [b.batch] + ([256] if name == "cfg2" else [])
            for bsz in batches:
                b.batch = bsz
                ms, _ = quick_epochs(torch, eng, b, host, g, n, 4 if bsz >= 4096 else 2)
                key = name if bsz == WORKLOADS[name]["batch"] else "%s_B%d" % (name, bsz)
                out[key] = {"workload": workload_name(b), "epoch_ms": ms, "pairs_per_s": pairs_per_epoch(b, n, nnz) / ms * 1e3,
                            "frac_algorithmic": bytes_per_epoch(b, n, nnz) / ms / 1e6 / peak}
                if b.model == 7:
                    # the epoch above uses the device walk sampler (`-walk 1`); the product default `-walk 0` draws the
                    # walks off the libc-compatible stream on the host, serially by construction (algorithms.cpp:1097-1118:
                    # the draws a walk consumes depend on its path) -- that, not the GPU, bounds an epoch of `-walk 0`
                    t0 = time.time()
                    g.walks(rp, ci)
                    out[key]["host_walks_ms_walk0"] = (time.time() - t0) * 1e3
    # cfg1: the reference's README command (cora, option 5, d=128, batch 256, 1200 iterations) through the
    # whole-run driver f2v_train (init + epochs + download, the span the reference itself times)
    mtx = os.path.join(ROOT, "tests", "golden", "cora.mtx")
    if os.path.exists(mtx):
        rp, ci = host.load_mtx(mtx)
        alg = F.Algorithms(rp, ci, "cora.mtx", "/tmp/", 128)
        alg.AlgoForce2VecNS(50, 0, 256, 5, 0.02, write=False)            # warm-up (context, plan)
        sec = alg.AlgoForce2VecNS(1200, 0, 256, 5, 0.02, write=False)[0]
        n, nnz = len(rp) - 1, len(ci)
        out["cfg1_cora_B256_it1200"] = {"workload": "cora.mtx option5 d128 B256 it1200 s5 (README.md:44) via f2v_train",
                                        "wall_s": sec, "pairs_per_s": (nnz + 5 * n) * 1200 / sec,
                                        "reference_wall_s_survey_8vcpu": 3.81}
    return out


def bounded_cpu_baseline(a, host):
    """cpu_baseline: the unmodified reference on this box's cores, on a BOUNDED sample of the workload --
    the same option / bs / dim / batch on the R-MAT graph of the same family with at most 2^20 vertices
    (one epoch; ~10-30 s of CPU work), in pair updates / s."""
    b = argparse.Namespace(**vars(a))
    b.scale = min(a.scale, 20)
    b.batch = min(a.batch, 1 << b.scale)
    rp, ci = host.rmat_csr_cached(b.scale, b.edge_factor, 1)
    r = cpu_reference(b, rp, ci, 1, warm=False)
    return {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
            "sample": "one epoch of %s (%s, %d threads, init differenced out)%s" %
                      (workload_name(b), r["what"], r["cores"],
                       "" if b.scale == a.scale else "; the R-MAT-%d graph of the same generator stands in for R-MAT-%d: "
                       "a full epoch of the headline workload on the CPU is what `--impl reference` times" % (b.scale, a.scale))}


def usable_cores():
    """Cores this process may run on (the affinity mask / cpuset, which a container lease can set below the
    machine's core count), not the machine's."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank unless it is set; the host side of the
    library (R-MAT generator, jump-ahead init draws, plan builder) is OpenMP code and the serial phases run on
    ONE rank while the others wait at a barrier (R-MAT 24 took 140 s instead of 11 s under torchrun).  Rank 0
    gets every core, the other ranks their share.  Must run before libgomp is loaded (torch, libf2v.so)."""
    if os.environ.get("F2V_KEEP_OMP_NUM_THREADS") == "1":
        return
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
    cores = usable_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores if rank == 0 else max(1, cores // max(world, 1)))


def main():
    host_threads()
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

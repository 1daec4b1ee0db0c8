"""CPU: bench.py's own control flow, dry.  The GPU arm of bench.py cannot run here, but everything around the
engine calls can: argument handling, the reference-ordered input streams (real host library), the JSON line the
driver parses (keys, units, per-GPU roofline at N > 1, parity field, extras), the N = 8 branch that starts the
R-MAT-26 children.  `torch`, `torch.distributed` and the engine are replaced by recording stand-ins that do no
arithmetic -- this checks the harness, not the product (the product path has no CPU fallback; the parity and
timing claims come from the GPU tests and the GPU bench)."""
import importlib
import json
import os
import sys
import types
import numpy as np
import pytest
from conftest import ROOT


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self):
        pass

    def elapsed_time(self, other):
        return 50.0


class _Tensor:
    def __init__(self, v):
        self.v = list(v)

    def item(self):
        return self.v[0]

    def numpy(self):
        return self._np


class _Stream:
    cuda_stream = 0x1234


def _fake_torch(world):
    t = types.ModuleType("torch")
    t.float32, t.float64, t.int32 = np.float32, np.float64, np.int32

    def empty(shape, dtype=None, pin_memory=False):
        x = _Tensor([0])
        x._np = np.empty(shape, dtype)
        return x
    t.empty = empty
    t.tensor = lambda v, device=None, dtype=None: _Tensor(v)
    t.device = lambda *a: ("cuda",) + a
    cuda = types.SimpleNamespace(set_device=lambda d: None, Stream=_Stream, set_stream=lambda s: None, Event=_Event,
                                 synchronize=lambda: None, empty_cache=lambda: None, current_device=lambda: 0)
    t.cuda = cuda
    calls = {"collectives": 0}

    def bump(*a, **k):
        calls["collectives"] += 1
    d = types.ModuleType("torch.distributed")
    d.ReduceOp = types.SimpleNamespace(MAX="max", MIN="min")
    d.init_process_group = bump
    d.barrier = bump
    d.all_reduce = bump                              # one process plays rank 0: its value is the reduction
    d.destroy_process_group = bump
    d.broadcast_object_list = bump                   # rank 0's object stays in place

    def all_gather_object(out, obj):
        for k in range(len(out)):
            out[k] = obj
    d.all_gather_object = all_gather_object
    t.distributed = d
    return t, d, calls


class _Engine:
    """Records what bench.py asks of the engine; returns fixed numbers."""
    log = []

    def __init__(self, rp, ci, dim, device=0):
        self.n, self.dim, self.nl = len(rp) - 1, dim, 0
        _Engine.log.append(("create", self.n, dim, device))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        _Engine.log.append(("close",))

    def _rec(name):
        def f(self, *a, **k):
            _Engine.log.append((name,) + tuple(x if isinstance(x, (int, float, str)) else type(x).__name__ for x in a))
        return f
    set_stream = _rec("set_stream"); set_epoch_mode = _rec("set_epoch_mode"); set_option = _rec("set_option")
    set_lut = _rec("set_lut"); set_embeddings = _rec("set_embeddings"); sync = _rec("sync")
    set_negative_offset = _rec("set_negative_offset"); sample_walks = _rec("sample_walks")
    comm_peer_init = _rec("comm_peer_init"); comm_init = _rec("comm_init")

    def set_negatives(self, idx):
        _Engine.log.append(("set_negatives", int(np.asarray(idx).size)))

    def run_epoch(self, model, batch, s, bs, lr, chunk=0):
        self.nl += -(-self.n // batch)
        _Engine.log.append(("run_epoch", model, batch, s, bs, chunk))

    def run_epoch_host(self, model, batch, s, bs, lr, X_in=None, neg=None, walks=None, X_out=None, chunk=0):
        assert X_in.shape == (self.n, self.dim) and X_out.shape == (self.n, self.dim) and neg is not None and neg.size > 0
        _Engine.log.append(("run_epoch_host", int(neg.size), int(neg[0])))

    def checksum(self):
        return 0xabcdef

    def get_rows(self, v, k):
        return np.full((k, self.dim), float(v), np.float32)

    def last_epoch_ms(self):
        return 2.0

    def device_memory(self):
        return 100 << 30, 180 << 30

    def launch_count(self):
        return self.nl

    def comm_peer_export(self):
        return b"blob"

    @staticmethod
    def comm_unique_id():
        return b"id"


class _Algorithms:
    def __init__(self, *a, **k):
        pass

    def AlgoForce2VecNS(self, *a, **k):
        return [0.125]


def _fake_pkg():
    import force2vec_b200 as real
    F = types.ModuleType("force2vec_b200")
    F.Engine, F.Algorithms, F.host = _Engine, _Algorithms, real.host
    fake_lib = types.SimpleNamespace(f2v_host_register=lambda p, b: 0, f2v_host_unregister=lambda p: 0)
    F.lib = lambda: fake_lib
    F.capi = types.SimpleNamespace(check=lambda rc, what: None)
    return F, real.host


def _run(monkeypatch, capsys, argv, world=1, rank=0, children=None):
    bench = importlib.import_module("bench")
    torch, dist, calls = _fake_torch(world)
    F, host = _fake_pkg()
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.distributed", dist)
    monkeypatch.setitem(sys.modules, "force2vec_b200", F)
    monkeypatch.setitem(sys.modules, "force2vec_b200.host", host)
    monkeypatch.setenv("RANK", str(rank)); monkeypatch.setenv("LOCAL_RANK", str(rank)); monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    small = {k: dict(v, scale=9, batch=128) for k, v in bench.WORKLOADS.items()}
    monkeypatch.setattr(bench, "WORKLOADS", small)
    monkeypatch.setattr(bench, "ClockSampler", lambda idx: types.SimpleNamespace(
        __enter__=lambda: None, __exit__=lambda *a: None,
        summary=lambda: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}))
    started = []
    if children is not None:
        def run_child(argv, port, timeout_s, script=None):
            started.append((list(argv), timeout_s))
            return children(argv)
        monkeypatch.setattr(bench, "run_child", run_child)
    _Engine.log = []
    a = bench.parse()
    bench.run_ours(a)
    out = capsys.readouterr().out
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out                    # exactly ONE JSON line on stdout
    return json.loads(lines[0]), list(_Engine.log), started, calls


def test_single_gpu_line_has_every_key_the_driver_reads(monkeypatch, capsys):
    line, log, _, _ = _run(monkeypatch, capsys, ["--steps", "4", "--warmup", "3"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "extra"):
        assert k in line, k
    assert line["metric"] == "force_pair_updates_per_sec" and line["unit"] == "pairs/s" and line["dtype"] == "f32"
    assert (line["n_gpus"], line["steps"], line["warmup"]) == (1, 4, 3) and line["vs_baseline"] is None
    n, nnz = line["config"]["n"], line["config"]["nnz"]
    assert n == 512 and line["config"]["pairs_per_epoch"] == nnz + 5 * n          # option 5: nnz + n*s
    assert line["ms_per_step"] == pytest.approx(50.0 / 4) and line["value"] == pytest.approx((nnz + 5 * n) / (0.050 / 4))
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"])
    assert rf["achieved"] == pytest.approx(((nnz + 5 * n) * 128 * 4 + n * 128 * 4) / (0.050 / 4) / 1e9)
    e = line["e2e"]
    assert e["unit"] == "pairs/s" and e["h2d_bytes_per_step"] > n * 128 * 4 and e["d2h_bytes_per_step"] == n * 128 * 4
    assert line["gpu_launches"] == 4 * 4                                           # 4 minibatches x 4 timed epochs
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    assert line["parity"] is None and set(line["extra"]) >= {"cfg2", "cfg3", "cfg1_cora_B256_it1200"}
    # resident part: 3 + 4 epochs, each at its own offset of the resident stream; end to end: 3 + 4 host-buffer
    # epochs, every one with ITS OWN stream (re-using the 7 drawn for the resident part) and host tables
    offs = [x[1] for x in log if x[0] == "set_negative_offset"][:7]
    stride = [x for x in log if x[0] == "run_epoch_host"][0][1]
    assert offs == [k * stride for k in range(7)]
    hostcalls = [x for x in log if x[0] == "run_epoch_host"]
    assert len(hostcalls) == 7 and len({x[2] for x in hostcalls}) > 1


def test_eight_gpu_line_is_per_gpu_checked_and_carries_the_scale26_children(monkeypatch, capsys):
    canned = {"value": 9e10, "ms_per_step": 80.0, "n_gpus": 8, "steps": 3, "warmup": 3, "gpu_launches": 768,
              "config": {"workload": "rmat26 child", "n": 1 << 26, "nnz": 2103827042},
              "parity": {"vs": "single_gpu", "bit_exact": True}, "setup": {"device_memory_used_GiB": 75.2, "graph_build_s": 50.0},
              "roofline": {"frac_algorithmic": 0.9}, "clocks": {"sm_mhz": 1965.0}}

    def children(argv):
        if "--sharded" in argv:
            return {"rc": 3, "timed_out": False, "line": {"error": "multi-GPU table differs"}, "stderr_tail": "x", "wall_s": 9.0}
        return {"rc": 0, "timed_out": False, "line": canned, "stderr_tail": "", "wall_s": 170.0}
    monkeypatch.setenv("F2V_BENCH_BUDGET_S", "100000")
    monkeypatch.setattr(np, "save", lambda *a, **k: None)            # (the shared negative file of the real run)
    monkeypatch.setattr(os, "replace", lambda *a, **k: None)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "shared_init", lambda a, n, host, dist, rank: (
        host.RandStream(1).init_embeddings(a.model, n, a.dim), host.RandStream(1)))
    line, log, started, calls = _run(monkeypatch, capsys, ["--gpus", "8", "--steps", "4", "--warmup", "3"], world=8, children=children)
    assert line["n_gpus"] == 8 and line["parity"]["bit_exact"] is True and line["parity"]["vs"] == "single_gpu"
    n, nnz = line["config"]["n"], line["config"]["nnz"]
    rf = line["roofline"]
    whole = ((nnz + 5 * n) * 128 * 4 + n * 128 * 4) / (0.050 / 4) / 1e9
    assert rf["achieved_whole_job"] == pytest.approx(whole) and rf["achieved"] == pytest.approx(whole / 8)
    assert rf["frac"] == pytest.approx(whole / 8 / rf["peak"]) and rf["traffic"] is None       # no N=8 capture: never a constant
    assert line["cpu_baseline"] is None                                          # rank 0 at N = 1 only
    # the checked epoch: a single-GPU engine is created, run once and closed BEFORE the 8-rank engine exists
    kinds = [x[0] for x in log]
    assert kinds.index("close") < [i for i, x in enumerate(kinds) if x == "create"][1]
    assert ("comm_peer_init", "list", 0, 8) in log
    # both children were started with the remaining budget, with the headline's arguments stripped down
    assert [s[0][:2] for s in started] == [["--workload", "cfg5"], ["--workload", "cfg5"]]
    assert all("--no-extra" in s[0] and s[0][s[0].index("--gpus") + 1] == "8" and 0 < s[1] <= 600 for s in started)
    ex = line["extra"]
    assert ex["cfg5_rmat26_replicated"]["pairs_per_s"] == 9e10 and ex["cfg5_rmat26_replicated"]["n"] == 1 << 26
    assert ex["cfg5_rmat26_replicated"]["parity"]["bit_exact"] is True
    assert ex["cfg5_rmat26_replicated"]["device_memory_used_GiB_per_gpu"] == 75.2
    assert "error" in ex["cfg5_rmat26_row_sharded"] and ex["cfg5_rmat26_row_sharded"]["rc"] == 3


def test_eight_gpu_children_are_skipped_when_the_budget_is_spent(monkeypatch, capsys):
    monkeypatch.setenv("F2V_BENCH_BUDGET_S", "1")                    # already over
    monkeypatch.setattr(np, "save", lambda *a, **k: None)
    monkeypatch.setattr(os, "replace", lambda *a, **k: None)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "shared_init", lambda a, n, host, dist, rank: (
        host.RandStream(1).init_embeddings(a.model, n, a.dim), host.RandStream(1)))
    line, _, started, _ = _run(monkeypatch, capsys, ["--gpus", "8", "--no-e2e"], world=8, children=lambda argv: None)
    assert started == [] and all("skipped" in v for v in line["extra"].values()) and len(line["extra"]) == 2   # (no fall-back either)
    assert line["e2e"] is None and line["value"] > 0


def test_eight_gpu_watchdog_prints_the_headline_if_the_extras_never_return(monkeypatch, capsys):
    """Should the machinery around the scale-26 children hang, every rank leaves on its own shortly after the
    budget and rank 0 has printed the headline line by then -- once."""
    import time
    monkeypatch.setenv("F2V_BENCH_BUDGET_S", "-1000")                # the deadline has passed before the extras start
    monkeypatch.setattr(np, "save", lambda *a, **k: None)
    monkeypatch.setattr(os, "replace", lambda *a, **k: None)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "shared_init", lambda a, n, host, dist, rank: (
        host.RandStream(1).init_embeddings(a.model, n, a.dim), host.RandStream(1)))
    exits = []
    monkeypatch.setattr(bench, "_exit", lambda code: exits.append(code))
    monkeypatch.setattr(bench, "cfg5_extras", lambda *a, **k: (time.sleep(2.5), {"late": 1})[1])     # "hangs" past the deadline
    line, _, _, _ = _run(monkeypatch, capsys, ["--gpus", "8", "--no-e2e"], world=8)
    assert exits and set(exits) == {0}
    assert "watchdog" in line["extra"]["error"] and line["value"] > 0 and line["n_gpus"] == 8


@pytest.mark.parametrize("argv", [["--workload", "cfg5", "--steps", "3", "--warmup", "3", "--no-e2e"],
                                  ["--workload", "cfg5", "--steps", "3", "--warmup", "3", "--no-e2e", "--multicast", "0"],
                                  ["--workload", "cfg5", "--sharded", "1", "--steps", "2", "--warmup", "3"]])
def test_the_scale26_child_command_lines_run_through(monkeypatch, capsys, argv):
    """The three command lines cfg5_extras gives its children (bench.CFG5_RUNS + the arguments it appends), dry:
    they parse, take the N = 8 path with the checked epoch, start no children of their own, skip the host-buffer
    epoch where asked (a row-sharded engine always does) and print one line with the parity field."""
    bench = importlib.import_module("bench")
    assert list(argv) in [list(r[1]) for r in bench.CFG5_RUNS]
    monkeypatch.setattr(np, "save", lambda *a, **k: None)
    monkeypatch.setattr(os, "replace", lambda *a, **k: None)
    monkeypatch.setattr(bench, "shared_init", lambda a, n, host, dist, rank: (
        host.RandStream(1).init_embeddings(a.model, n, a.dim), host.RandStream(1)))
    called = []
    monkeypatch.setattr(bench, "cfg5_extras", lambda *a, **k: called.append(1) or {})
    line, log, _, _ = _run(monkeypatch, capsys, argv + ["--gpus", "8", "--no-extra", "--no-cpu-baseline"], world=8)
    assert not called and line["extra"] == {} and line["e2e"] is None
    assert line["parity"]["bit_exact"] is True and line["n_gpus"] == 8 and line["steps"] == int(argv[argv.index("--steps") + 1])
    assert ("sharded" in line["setup"]["parallelism"]) == ("--sharded" in argv)
    opts = {x[1]: x[2] for x in log if x[0] == "set_option"}
    assert opts["sharded"] == (1 if "--sharded" in argv else 0) and opts["multicast"] == (0 if "--multicast" in argv else 1)
    assert not any(x[0] == "run_epoch_host" for x in log)

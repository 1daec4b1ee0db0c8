"""GPU (needs >= 2 devices): the multi-GPU paths (fused peer-store / multicast exchange, row-sharded
tables, NCCL all-gather baseline) must reproduce the single-GPU epoch bit for bit on every rank (same
per-vertex accumulation order); the scale-26 configuration the reference cannot run at all
(sample/algorithms.h:40,68: 32-bit n*DIM) is checked by 1-GPU-vs-8-GPU self-consistency."""
import json
import os
import subprocess
import sys
import pytest

import force2vec_b200 as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_equals_single_gpu(world):
    ndev = F.lib().f2v_device_count()
    if ndev < world:
        pytest.skip("needs %d GPUs, have %d" % (world, ndev))
    port = 29600 + (os.getpid() + world) % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "_mgpu_worker.py")], capture_output=True, timeout=900)
    out = r.stdout.decode() + r.stderr.decode()
    assert r.returncode == 0 and "MGPU_OK" in out, out[-3000:]


def test_driver_on_two_gpus_in_one_process(cora):
    """`f2v_train_gpus` (what `bin/Force2Vec -gpus 2` runs): one process, one host thread and one
    engine per GPU, plain peer access between the engines.  Same result as the single-GPU driver
    bit for bit for equal chunk."""
    import numpy as np
    if F.lib().f2v_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rp, ci = cora
    out = []
    for gpus in (1, 2):
        alg = F.Algorithms(rp, ci, "cora.mtx", "/tmp/", 128)
        alg.gpus, alg.chunk = gpus, 64
        for run in (alg.AlgoForce2VecNS, alg.AlgoForce2VecNSRW, alg.AlgoForce2VecNSRWEFF):
            run(5, 0, 256, 5, 0.02, write=False)
            out.append(alg.nCoordinates.copy())
    for a, b in zip(out[:3], out[3:]):
        assert np.array_equal(a, b)


def _bench(world, extra, timeout=3000):
    port = 29900 + (os.getpid() + world) % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
           "--gpus", str(world), "--steps", "1", "--warmup", "1", "--no-extra", "--no-cpu-baseline"] + extra
    r = subprocess.run(cmd, capture_output=True, timeout=timeout)
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout.decode() + r.stderr.decode())[-3000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("sharded", [0, 1])
def test_scale26_one_gpu_vs_eight_gpus(sharded):
    """BASELINE config 5: R-MAT scale 26 (67 M vertices, ~2 G CSR entries), option 5, d=128 -- n*d = 2^33,
    beyond the reference's 32-bit indexing.  No oracle exists, so every rank's table after one epoch on
    8 GPUs (replicated with the fused exchange, and row-sharded) must equal -- device checksum of the
    whole 32 GiB table plus 64 probe rows -- a single-GPU engine's from the same state.  The check is
    bench.py's own checked epoch; F2V_BIG_SCALE lowers the scale on smaller boxes."""
    world = int(os.environ.get("F2V_BIG_WORLD", "8"))
    if F.lib().f2v_device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    scale = os.environ.get("F2V_BIG_SCALE", "26")
    line = _bench(world, ["--workload", "cfg5", "--scale", scale, "--sharded", str(sharded), "--no-e2e"])
    assert line["parity"]["bit_exact"] is True and line["parity"]["vs"] == "single_gpu"
    assert line["config"]["n"] == 1 << int(scale)
    print(json.dumps(line))


def test_bench_checks_parity_at_two_gpus():
    """bench.py at N > 1 runs one checked epoch (N-rank table == single-GPU table, bit for bit) before it
    times anything and reports it in the line."""
    if F.lib().f2v_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    line = _bench(2, ["--workload", "cfg2", "--scale", "16", "--batch", "4096"], timeout=900)
    assert line["parity"]["bit_exact"] is True
    assert line["e2e"]["value"] > 0 and line["n_gpus"] == 2


def test_driver_reports_a_failure_on_one_rank_instead_of_hanging(cora):
    """f2v_train_gpus: a failure on ONE device (injected at epoch 1's upload on rank 1) must come back as
    an error from every thread's barrier, not deadlock the others (threads that read a shared
    flag at different points ran different numbers of barriers)."""
    if F.lib().f2v_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rp, ci = cora
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, force2vec_b200 as F; from force2vec_b200 import host; "
            "rp, ci = host.load_mtx(%r); alg = F.Algorithms(rp, ci, 'cora.mtx', '/tmp/', 64); alg.gpus = 2\n"
            "try:\n    alg.AlgoForce2VecNS(4, 0, 256, 5, 0.02, write=False); print('NO_ERROR')\n"
            "except F.F2VError as ex:\n    print('GOT_ERROR', ex)") % (ROOT, os.path.join(ROOT, "tests", "golden", "cora.mtx"))
    for rank, epoch in ((1, 1), (0, 2), (1, 0)):
        env = dict(os.environ, F2V_TEST_FAIL_RANK=str(rank), F2V_TEST_FAIL_EPOCH=str(epoch))
        r = subprocess.run([sys.executable, "-c", code.replace("\\n", "\n")], capture_output=True, timeout=120, env=env)
        out = r.stdout.decode() + r.stderr.decode()
        assert "GOT_ERROR" in out and "injected failure" in out, out[-2000:]

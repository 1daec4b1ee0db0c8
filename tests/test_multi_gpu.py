"""GPU (needs >= 2 devices): the minibatch-split + NCCL all-gather path must reproduce the
single-GPU epoch bit for bit on every rank (same per-vertex accumulation order)."""
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

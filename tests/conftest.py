import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import force2vec_b200 as F
        return F.lib().f2v_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a machine without a GPU must fail loudly, not skip: only auto-skip when the
    # user did not ask for gpu tests explicitly.
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def cora(oracle):
    return oracle.load_mtx(os.path.join(GOLDEN, "cora.mtx"))


@pytest.fixture(scope="session")
def karate(oracle):
    return oracle.load_mtx(os.path.join(GOLDEN, "karate.mtx"))


@pytest.fixture(scope="session")
def ref_outputs():
    return np.load(os.path.join(GOLDEN, "ref_outputs.npz"))


def gkey(graph, opt, bs, dim, B, it):
    return "%s_opt%d_bs%d_d%d_B%d_it%d" % (graph, opt, bs, dim, B, it)

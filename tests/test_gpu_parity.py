"""GPU: parity of the CUDA force step (through the C ABI, include/f2v.h) with the oracle
(oracle/f2v_oracle.c) on identical inputs and identical injected sample streams, with the
committed reference outputs (tests/golden/ref_outputs.npz), and size-independent properties at
the benchmark's full size.  Tolerances (fp32): teacher-forced single minibatch rtol 1e-5;
free-running <= 5 epochs rtol 1e-4 / atol 1e-5; 50 epochs rtol 1e-4 / atol 1e-4 (chaotic drift +
truncating-LUT bin flips, see tests/test_oracle.py)."""
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, gkey

import force2vec_b200 as F
from force2vec_b200 import host

pytestmark = pytest.mark.gpu
LR = 0.02


def _streams(oracle, model, bs, rp, ci, dim, iters, batch, s):
    """Oracle run that also logs the init and every draw, re-laid-out for the engine."""
    n = len(rp) - 1
    full = oracle.run(model, bs, rp, ci, dim, iters, batch, s, LR, want_init=True, want_logs=True)
    W = (batch + s - 1) if (bs and model != 7) else s
    neg = np.ascontiguousarray(full["neg"][:, :, :W])       # engine layout: first batch+s-1 draws
    return full, neg


def _engine(rp, ci, dim, X0, model):
    e = F.Engine(rp, ci, dim)
    e.set_embeddings(X0)
    if model != 5:
        e.set_lut()
    return e


GRAPH_CASES = [("karate", 8), ("cora", 256)]
MODEL_CASES = [(5, 0), (5, 1), (6, 0), (6, 1), (7, 0)]


@pytest.mark.parametrize("gname,batch", GRAPH_CASES)
@pytest.mark.parametrize("model,bs", MODEL_CASES)
@pytest.mark.parametrize("dim", [128, 64, 32, 256, 20, 100, 300])
def test_step_teacher_forced(oracle, request, gname, batch, model, bs, dim):
    """Every minibatch of one epoch, each from the SAME pre-step state in both engines."""
    if gname == "cora" and dim not in (128, 64, 100):
        pytest.skip("dim sweep runs on karate")
    rp, ci = request.getfixturevalue(gname)
    n = len(rp) - 1
    s = 5
    full, neg = _streams(oracle, model, bs, rp, ci, dim, 1, batch, s)
    X = full["X0"].copy()
    walks = full["walks"][0] if model == 7 else None
    with _engine(rp, ci, dim, X, model) as e:
        if model == 7:
            e.set_walks(walks)
        for b in range((n + batch - 1) // batch):
            lo, hi = b * batch, min(n, (b + 1) * batch)
            e.set_embeddings(X)                               # teacher forcing: oracle state in
            W = (hi - lo + s - 1) if (bs and model != 7) else s
            e.step(model, lo, hi - lo, neg[0, b, :W], s, bs, LR)
            oracle.step(model, bs, rp, ci, X, lo, hi, full["neg"][0, b], s, LR, walks=walks)
            got = e.get_rows(lo, hi - lo)
            np.testing.assert_allclose(got, X[lo:hi], rtol=1e-5, atol=1e-6)
            if lo > 0:
                assert np.array_equal(e.get_rows(0, lo), X[:lo])       # untouched rows stay bit-identical
    assert np.array_equal(X, full["X"])


@pytest.mark.parametrize("model,bs", MODEL_CASES)
@pytest.mark.parametrize("dim,batch", [(128, 256), (64, 256), (128, 384), (128, 5000), (48, 100)])
def test_epochs_free_running_vs_oracle(oracle, cora, model, bs, dim, batch):
    rp, ci = cora
    s, iters = 5, 3
    full, neg = _streams(oracle, model, bs, rp, ci, dim, iters, batch, s)
    with _engine(rp, ci, dim, full["X0"], model) as e:
        for it in range(iters):
            if model == 7:
                e.set_walks(full["walks"][it])
            e.set_negatives(neg[it])
            e.run_epoch(model, batch, s, bs, LR)
        got = e.get_embeddings()
    np.testing.assert_allclose(got, full["X"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("model,bs", MODEL_CASES)
def test_epoch_equals_loop_of_steps(oracle, cora, model, bs):
    """f2v_run_epoch (ping-pong tables) == a loop of f2v_step (stage + apply), bit for bit."""
    rp, ci = cora
    n = len(rp) - 1
    dim, batch, s = 128, 256, 5
    full, neg = _streams(oracle, model, bs, rp, ci, dim, 1, batch, s)
    with _engine(rp, ci, dim, full["X0"], model) as e:
        if model == 7:
            e.set_walks(full["walks"][0])
        e.set_negatives(neg[0])
        e.run_epoch(model, batch, s, bs, LR)
        a = e.get_embeddings()
        e.set_embeddings(full["X0"])
        for b in range((n + batch - 1) // batch):
            lo, hi = b * batch, min(n, (b + 1) * batch)
            W = (hi - lo + s - 1) if (bs and model != 7) else s
            e.step(model, lo, hi - lo, neg[0, b, :W], s, bs, LR)
        c = e.get_embeddings()
    assert np.array_equal(a, c)


CORA_GOLD = [(5, 0, 128, 256, 1), (5, 0, 128, 256, 5), (5, 0, 128, 256, 50), (5, 1, 128, 256, 2),
             (6, 0, 128, 256, 1), (6, 0, 128, 256, 5), (6, 0, 128, 256, 50), (6, 1, 128, 256, 2),
             (7, 0, 64, 256, 1), (7, 0, 64, 256, 5), (7, 0, 64, 256, 50), (7, 0, 128, 384, 2)]


@pytest.mark.parametrize("opt,bs,dim,B,it", CORA_GOLD)
def test_driver_vs_reference_golden(cora, ref_outputs, opt, bs, dim, B, it):
    """The C++ host driver (f2v_train: own rand() stream, own samplers, GPU epochs) against the
    outputs of the unmodified reference on cora -- nothing from oracle/ is involved."""
    rp, ci = cora
    a = F.Algorithms(rp, ci, "cora.mtx", "/tmp/", dim)
    fn = {(5, 0): a.AlgoForce2VecNS, (5, 1): a.AlgoForce2VecNSBS, (6, 0): a.AlgoForce2VecNSRW,
          (6, 1): a.AlgoForce2VecNSRWBS, (7, 0): a.AlgoForce2VecNSRWEFF}[(opt, bs)]
    sec = fn(it, 1, B, 5, LR, write=False)
    assert len(sec) == 1 and sec[0] > 0
    k = gkey("cora", opt, bs, dim, B, it)
    X = a.nCoordinates
    np.testing.assert_allclose(X[::4], ref_outputs[k], rtol=1e-4, atol=1e-5 if it <= 5 else 1e-4)
    assert abs(X.astype(np.float64).sum() - float(ref_outputs[k + "_sum"])) < 5e-2
    assert abs(np.linalg.norm(X.astype(np.float64)) - float(ref_outputs[k + "_fro"])) < 5e-3


@pytest.mark.parametrize("opt,bs,dim,it", [(o, b, d, i) for o in (5, 6, 7) for b in ((0, 1) if o != 7 else (0,))
                                            for d in (128, 64, 20) for i in (1, 3)])
def test_driver_vs_reference_golden_karate(karate, ref_outputs, opt, bs, dim, it):
    rp, ci = karate
    a = F.Algorithms(rp, ci, "karate.mtx", "/tmp/", dim)
    fn = {(5, 0): a.AlgoForce2VecNS, (5, 1): a.AlgoForce2VecNSBS, (6, 0): a.AlgoForce2VecNSRW,
          (6, 1): a.AlgoForce2VecNSRWBS, (7, 0): a.AlgoForce2VecNSRWEFF}[(opt, bs)]
    fn(it, 1, 8, 5, LR, write=False)
    np.testing.assert_allclose(a.nCoordinates, ref_outputs[gkey("karate", opt, bs, dim, 8, it)], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("model,bs", MODEL_CASES)
@pytest.mark.parametrize("chunk", [8, 64, 1 << 30])
def test_hub_rows_split_across_warps(oracle, model, bs, chunk):
    """R-MAT scale 12 (max degree in the hundreds): rows longer than `chunk` are cut into chunks
    whose partial sums are folded in a fixed order; any chunking agrees with the oracle."""
    rp, ci = host.rmat_csr(12, 16, 1)
    dim, batch, s = 128, 512, 5
    full, neg = _streams(oracle, model, bs, rp, ci, dim, 2, batch, s)
    outs = []
    for rep in range(2):
        with _engine(rp, ci, dim, full["X0"], model) as e:
            for it in range(2):
                if model == 7:
                    e.set_walks(full["walks"][it])
                e.set_negatives(neg[it])
                e.run_epoch(model, batch, s, bs, LR, chunk=chunk)
            outs.append(e.get_embeddings())
    assert np.array_equal(outs[0], outs[1])                  # deterministic (no float atomics)
    np.testing.assert_allclose(outs[0], full["X"], rtol=1e-4, atol=1e-5)


def test_self_negative_quirk_on_device(oracle, karate):
    """vertex == its own negative -> every component gets lr*(-5) (option 5)."""
    rp, ci = karate
    rng = np.random.default_rng(0)
    X0 = rng.uniform(-1, 1, (34, 128)).astype(np.float32)
    idx = np.array([20, 21, 22, 23, 3], np.uint32)
    with F.Engine(rp, ci, 128) as e:
        e.set_embeddings(X0)
        e.step(5, 0, 8, idx, 5, 0, LR)
        b = e.get_rows(0, 8)
        e.set_embeddings(X0)
        e.step(5, 0, 8, idx[:4], 4, 0, LR)
        c = e.get_rows(0, 8)
    np.testing.assert_allclose(b[3] - c[3], np.full(128, -0.1, np.float32), rtol=0, atol=1e-6)
    assert np.isfinite(b).all()
    Xo = X0.copy()
    oracle.step(5, 0, rp, ci, Xo, 0, 8, idx, 5, LR)
    np.testing.assert_allclose(b, Xo[:8], rtol=1e-5, atol=1e-6)


def test_edge_cases(oracle):
    # empty graph (nnz = 0), isolated vertices, batch > n, s = 0, s > 32, partial last minibatch
    n, dim = 70, 64
    rp0 = np.zeros(n + 1, np.uint64)
    ci0 = np.zeros(0, np.uint32)
    rng = np.random.default_rng(1)
    X0 = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    for model in (5, 6):
        for s in (0, 3, 40):
            idx = rng.integers(0, n - 1, size=max(s, 1)).astype(np.uint32)[:s]
            with _engine(rp0, ci0, dim, X0, model) as e:
                e.step(model, 0, n, idx, s, 0, LR)
                got = e.get_embeddings()
            Xo = X0.copy()
            oracle.step(model, 0, rp0, ci0, Xo, 0, n, idx, s, LR)
            np.testing.assert_allclose(got, Xo, rtol=1e-5, atol=1e-6)
            if s == 0:
                assert np.array_equal(got, X0) or model == 5     # no pairs: x + 0
    rp, ci = host.rmat_csr(8, 4, 2)
    n = len(rp) - 1
    for model, bs in MODEL_CASES:
        for batch in (n + 50, 37):
            full, neg = _streams(oracle, model, bs, rp, ci, 32, 2, batch, 5)
            with _engine(rp, ci, 32, full["X0"], model) as e:
                for it in range(2):
                    if model == 7:
                        e.set_walks(full["walks"][it])
                    e.set_negatives(neg[it])
                    e.run_epoch(model, batch, 5, bs, LR)
                np.testing.assert_allclose(e.get_embeddings(), full["X"], rtol=1e-4, atol=1e-5)


def test_argument_errors(karate):
    rp, ci = karate
    with F.Engine(rp, ci, 16) as e:
        with pytest.raises(F.F2VError):
            e.run_epoch(6, 8, 5, 0, LR)            # sigmoid table not set
        with pytest.raises(F.F2VError):
            e.run_epoch(5, 8, 5, 0, LR)            # no negative stream resident
        with pytest.raises(F.F2VError):
            e.run_epoch(4, 8, 5, 0, LR)            # unknown model
        with pytest.raises(F.F2VError):
            e.step(5, 30, 10, np.zeros(5, np.uint32), 5, 0, LR)   # rows out of range
        e.set_lut()
        with pytest.raises(F.F2VError):
            e.run_epoch(7, 8, 5, 0, LR)            # walks not set
    bad = ci.copy()
    bad[0] = 1000
    with pytest.raises(F.F2VError):
        F.Engine(rp, bad, 16)


def test_device_walk_sampler_matches_host_mirror(oracle, cora):
    rp, ci = cora
    with F.Engine(rp, ci, 16) as e:
        for seed, epoch in ((1, 0), (1, 1), (99, 7)):
            e.sample_walks(seed, epoch)
            assert np.array_equal(e.get_walks(), oracle.walks_counter(seed, epoch, rp, ci))
    rp, ci = host.rmat_csr(13, 16, 4)
    with F.Engine(rp, ci, 16) as e:
        e.sample_walks(5, 2)
        assert np.array_equal(e.get_walks(), oracle.walks_counter(5, 2, rp, ci))


def test_option7_with_device_sampler(oracle, cora):
    """Option 7 end to end with the device sampler: identical to the oracle fed the same walks."""
    rp, ci = cora
    n = len(rp) - 1
    dim, batch, s = 64, 256, 5
    g = oracle.Rng(1)
    X = oracle.init_embeddings(g, 7, n, dim)
    with _engine(rp, ci, dim, X, 7) as e:
        for it in range(3):
            e.sample_walks(1, it)
            walks = oracle.walks_counter(1, it, rp, ci)
            nb = (n + batch - 1) // batch
            neg = np.stack([oracle.draw_negatives(g, 7, 0, n, batch, s, b) for b in range(nb)])
            e.set_negatives(neg)
            e.run_epoch(7, batch, s, 0, LR)
            for b in range(nb):
                oracle.step(7, 0, rp, ci, X, b * batch, min(n, (b + 1) * batch), neg[b], s, LR, walks=walks)
        np.testing.assert_allclose(e.get_embeddings(), X, rtol=1e-4, atol=1e-5)


def test_host_buffer_epoch_equals_resident_epoch(oracle, cora):
    rp, ci = cora
    full, neg = _streams(oracle, 6, 0, rp, ci, 128, 1, 256, 5)
    out = np.empty_like(full["X0"])
    with F.Engine(rp, ci, 128) as e:
        e.set_lut()
        e.run_epoch_host(6, 256, 5, 0, LR, X_in=full["X0"], neg=neg[0], X_out=out)
        e.set_embeddings(full["X0"])
        e.set_negatives(neg[0])
        e.run_epoch(6, 256, 5, 0, LR)
        assert np.array_equal(out, e.get_embeddings())
        assert e.last_epoch_ms() > 0 and e.launch_count() >= 22
    np.testing.assert_allclose(out, full["X"], rtol=1e-4, atol=1e-5)


def test_cli_drop_in(oracle, tmp_path):
    """bin/Force2Vec: reference flags in, reference-format .embd + Results.txt out."""
    exe = os.path.join(ROOT, "bin", "Force2Vec")
    assert os.path.exists(exe), "build bin/Force2Vec (make cli)"
    for opt, tag in ((5, "F2VNS"), (6, "F2VWNS"), (7, "F2VWNSF")):
        r = subprocess.run([exe, "-input", os.path.join(GOLDEN, "karate.mtx"), "-output", str(tmp_path) + "/",
                            "-iter", "3", "-batch", "8", "-dim", "16", "-nsamples", "5", "-option", str(opt)],
                           capture_output=True, cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr
        out = tmp_path / ("karate.mtx%s8D16IT3NS5.embd" % tag)
        assert out.exists()
        got_txt = out.read_text()
        want_txt = open(os.path.join(GOLDEN, "karate_opt%d.embd" % opt)).read()
        gl, wl = got_txt.split("\n"), want_txt.split("\n")
        assert gl[0] == wl[0] == "34 16" and len(gl) == len(wl)
        assert all(l.endswith(" ") for l in gl[1:-1])
        np.testing.assert_allclose(oracle.read_embd(str(out)), oracle.read_embd(os.path.join(GOLDEN, "karate_opt%d.embd" % opt)),
                                   rtol=1e-4, atol=1e-5)
        assert b"Creating output file in following directory:" in r.stdout
    res = (tmp_path / "Results.txt").read_text().strip().split("\n")
    assert len(res) == 3 and res[0].startswith("Algo:Force2Vec:t-distribution with negative sampling\tInit:RAND\tIteration:3\t")


def test_full_size_rmat20_properties(oracle):
    """BASELINE config 2 at full size (R-MAT scale 20, option 6, d=128): the whole epoch against
    the oracle on the box's host cores, determinism, and the zero-degree closed form."""
    rp, ci = host.rmat_csr(20, 16, 1)
    n = len(rp) - 1
    dim, batch, s = 128, 16384, 5
    g = host.RandStream(1)
    X0 = g.init_embeddings(6, n, dim)
    neg = g.epoch_negatives(6, n, batch, s, 0).copy()
    lut = host.build_lut()
    with F.Engine(rp, ci, dim) as e:
        e.set_lut(lut)
        e.set_embeddings(X0)
        e.set_negatives(neg)
        e.run_epoch(6, batch, s, 0, LR)
        a = e.get_embeddings()
        e.set_embeddings(X0)
        e.set_negatives(neg)
        e.run_epoch(6, batch, s, 0, LR, chunk=256)
        b = e.get_embeddings()
        e.set_embeddings(X0)
        e.set_negatives(neg)
        e.run_epoch(6, batch, s, 0, LR)
        c = e.get_embeddings()
    assert np.array_equal(a, c)                                         # deterministic
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)              # chunking only reorders sums
    assert np.isfinite(a).all()
    Xo = X0.copy()
    nb = (n + batch - 1) // batch
    lut_o = oracle.build_lut()
    for bb in range(nb):
        oracle.step(6, 0, rp, ci, Xo, bb * batch, min(n, (bb + 1) * batch), neg[bb * s:(bb + 1) * s], s, LR, lut=lut_o)
    np.testing.assert_allclose(a, Xo, rtol=1e-4, atol=1e-5)


def test_full_size_rmat22_option7(oracle):
    """BASELINE config 3 at full size (R-MAT scale 22, option 7 semi-random walks, d=64): a whole
    epoch with the reference's serial libc-stream walks against the oracle, and the device
    sampler against its host mirror at 21 M draws."""
    rp, ci = host.rmat_csr(22, 16, 1)
    n = len(rp) - 1
    dim, batch, s = 64, 65536, 5
    g = host.RandStream(1)
    X0 = g.init_embeddings(7, n, dim)
    walks = g.walks(rp, ci).copy()
    neg = g.epoch_negatives(7, n, batch, s, 0).copy()
    lut = host.build_lut()
    with F.Engine(rp, ci, dim) as e:
        e.set_lut(lut)
        e.set_embeddings(X0)
        e.set_walks(walks)
        e.set_negatives(neg)
        e.run_epoch(7, batch, s, 0, LR)
        a = e.get_embeddings()
        e.sample_walks(7, 3)
        dev = e.get_walks()
    assert np.array_equal(dev, oracle.walks_counter(7, 3, rp, ci))
    Xo = X0.copy()
    nb = (n + batch - 1) // batch
    lut_o = oracle.build_lut()
    for bb in range(nb):
        oracle.step(7, 0, rp, ci, Xo, bb * batch, min(n, (bb + 1) * batch), neg[bb * s:(bb + 1) * s], s, LR,
                    lut=lut_o, walks=walks)
    np.testing.assert_allclose(a, Xo, rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(os.environ.get("F2V_SKIP_BIG") == "1", reason="F2V_SKIP_BIG=1")
def test_full_size_rmat24_option5_bs1(oracle):
    """BASELINE config 4 at full size (R-MAT scale 24: 16.8 M vertices, ~0.5 G CSR entries, option 5
    with per-vertex negatives, d=128; 2 x 8 GiB tables).  The oracle cannot finish an epoch of this
    in seconds, so: teacher-forced minibatches (the hub-heavy first one, one from the middle, the
    last) against the oracle on the full table; a whole epoch twice (bit-reproducible); and the
    size-independent property that the first minibatch of an epoch equals the teacher-forced step
    from the same table bit for bit."""
    scale = int(os.environ.get("F2V_BIG_SCALE", "24"))
    rp, ci = host.rmat_csr(scale, 16, 1)
    n = len(rp) - 1
    dim, batch, s = 128, 65536, 5
    W = batch + s - 1
    rng = np.random.default_rng(24)
    X0 = rng.random((n, dim), dtype=np.float32) * 2.0 - 1.0       # U[-1,1) like randInitF; stream parity is covered elsewhere
    nb = (n + batch - 1) // batch
    neg = rng.integers(0, n - 1, size=nb * W, dtype=np.uint32)    # range of randIndex(n-1)
    with F.Engine(rp, ci, dim) as e:
        e.set_embeddings(X0)
        steps = {}
        for bb in (0, nb // 2, nb - 1):
            lo, hi = bb * batch, min(n, (bb + 1) * batch)
            e.step(5, lo, hi - lo, neg[bb * W:bb * W + (hi - lo) + s - 1], s, 1, LR)
            steps[bb] = e.get_rows(lo, hi - lo)
            saved = X0[lo:hi].copy()
            idx = np.zeros(s * batch + s, np.uint32)
            idx[:(hi - lo) + s - 1] = neg[bb * W:bb * W + (hi - lo) + s - 1]
            oracle.step(5, 1, rp, ci, X0, lo, hi, idx, s, LR, threads=os.cpu_count() or 1)
            np.testing.assert_allclose(steps[bb], X0[lo:hi], rtol=1e-4, atol=1e-5)
            X0[lo:hi] = saved
            e.set_embeddings(X0)                                  # teacher-forced: back to the same table
        e.set_negatives(neg)
        # (hub rows cut at 128 edges, as f2v_step cuts them: results are bit-identical for EQUAL chunk; the epoch's
        # own default at this batch size and dimension is 256)
        e.run_epoch(5, batch, s, 1, LR, chunk=128)
        first = e.get_rows(0, batch)
        probe = [(0, 1 << 18), (n // 2, 1 << 18), (n - (1 << 18), 1 << 18)]
        a = [e.get_rows(lo, cnt) for lo, cnt in probe]
        e.set_embeddings(X0)
        e.set_negatives(neg)
        e.run_epoch(5, batch, s, 1, LR, chunk=128)
        b = [e.get_rows(lo, cnt) for lo, cnt in probe]
    assert np.array_equal(first, steps[0])
    for x, y in zip(a, b):
        assert np.array_equal(x, y) and np.isfinite(x).all()


@pytest.mark.parametrize("model,bs,dim,batch", [(6, 0, 32, 37), (5, 1, 128, 64), (7, 0, 64, 256), (6, 1, 128, 1000)])
def test_dependent_launch_chaining_is_exact(model, bs, dim, batch):
    """Programmatic dependent launch lets minibatch b+1 start while b drains; whatever it reads
    before its dependency wait must be data b cannot have written.  Three epochs with the chaining
    off (pdl 0), on (pdl 1: wait before the own row) and on with the early own-row read (pdl 2),
    several times each: all bit-identical."""
    rp, ci = host.rmat_csr(10, 16, 2)
    n = len(rp) - 1
    g = host.RandStream(1)
    X0 = g.init_embeddings(model, n, dim)
    streams = []
    for it in range(3):
        w = g.walks(rp, ci).copy() if model == 7 else None
        streams.append((w, g.epoch_negatives(model, n, batch, 5, bs).copy()))
    ref = None
    for pdl in (0, 2, 2, 2, 1, 1, 2, 2):
        with _engine(rp, ci, dim, X0, model) as e:
            e.set_option("pdl", pdl)
            for w, neg in streams:
                if w is not None:
                    e.set_walks(w)
                e.set_negatives(neg)
                e.run_epoch(model, batch, 5, bs, LR)
            X = e.get_embeddings()
        if ref is None:
            ref = X
        assert np.array_equal(ref, X), pdl


@pytest.mark.parametrize("model,bs", MODEL_CASES)
@pytest.mark.parametrize("dim,batch,scale,chunk", [(128, 512, 12, 0), (64, 1000, 12, 16), (128, 4096, 13, 8), (64, 64, 10, 0)])
def test_async_ring_layouts_are_bit_identical(oracle, model, bs, dim, batch, scale, chunk):
    """The asynchronous shared-memory-ring gather (RingL layouts: cp.async stages, neighbours and
    per-vertex negatives in one stream, split rows taking their negatives after the fold) performs the
    same operations in the same order as the register layouts: every variant gives the same bits, with
    PDL chaining on, and agrees with the oracle."""
    rp, ci = host.rmat_csr(scale, 16, 3)
    s = 5
    full, neg = _streams(oracle, model, bs, rp, ci, dim, 2, batch, s)
    ref = None
    for variant in (3 if dim == 128 else 1, 21, 22, -1):
        with _engine(rp, ci, dim, full["X0"], model) as e:
            e.set_option("variant", variant)
            for it in range(2):
                if model == 7:
                    e.set_walks(full["walks"][it])
                e.set_negatives(neg[it])
                e.run_epoch(model, batch, s, bs, LR, chunk=chunk)
            X = e.get_embeddings()
            h = e.checksum()
        if ref is None:
            ref, href = X, h
            np.testing.assert_allclose(X, full["X"], rtol=1e-4, atol=1e-5)
        assert np.array_equal(ref, X), variant
        assert h == href                                         # the device checksum sees the same table


def test_checksum_detects_a_single_bit(cora):
    rp, ci = cora
    n = len(rp) - 1
    X = host.RandStream(1).init_embeddings(5, n, 128)
    with F.Engine(rp, ci, 128) as e:
        e.set_embeddings(X)
        a = e.checksum()
        assert a == e.checksum()
        Y = X.copy()
        Y.view(np.uint32)[n // 2, 77] ^= 1
        e.set_embeddings(Y)
        b = e.checksum()
        Y = X.copy()
        Y[[3, 4]] = Y[[4, 3]]                                     # position matters, not only the multiset of values
        e.set_embeddings(Y)
        c = e.checksum()
    assert len({a, b, c}) == 3


@pytest.mark.parametrize("model,bs", MODEL_CASES)
@pytest.mark.parametrize("dim,batch,scale", [(128, 256, 12), (64, 37, 9), (128, 16, 10), (20, 100, 9), (256, 64, 10), (64, 1024, 13)])
def test_dataflow_epoch_equals_launch_per_minibatch(oracle, model, bs, dim, batch, scale):
    """Epoch mode 2 (ONE ordinary launch per epoch, no barrier: items handed out in order, a warp waits
    only for the minibatches that wrote the rows it reads) is the same computation as mode 0 (one
    launch per minibatch): bit-identical tables after 3 epochs, repeated to give races a chance, with
    split hub rows, a partial last minibatch and batches small enough that dozens of dependent
    minibatches are in flight at once -- and it agrees with the oracle."""
    rp, ci = host.rmat_csr(scale, 16, 5)
    n = len(rp) - 1
    s = 5
    full, neg = _streams(oracle, model, bs, rp, ci, dim, 3, batch, s)
    out = []
    for mode in (0, 2, 2, 2, 2):
        with _engine(rp, ci, dim, full["X0"], model) as e:
            e.set_epoch_mode(mode)
            for it in range(3):
                if model == 7:
                    e.set_walks(full["walks"][it])
                e.set_negatives(neg[it])
                e.run_epoch(model, batch, s, bs, LR, chunk=32)
            out.append(e.get_embeddings())
            launches = e.launch_count()
        if mode == 2:
            assert launches == 3                 # one launch per epoch
    for X in out[1:]:
        assert np.array_equal(out[0], X)
    np.testing.assert_allclose(out[0], full["X"], rtol=2e-4, atol=2e-5)

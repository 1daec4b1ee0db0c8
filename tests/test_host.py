"""CPU: host side of the drop-in (libf2v.so, include/f2v_host.h) against the oracle and the
golden fixtures: rand() stream and samplers, MatrixMarket loader, .embd writer, R-MAT generator,
work plan; and that the C-ABI library loads, exports every declared symbol and fails loudly
without a GPU."""
import ctypes
import os
import re
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT

import force2vec_b200 as F
from force2vec_b200 import host, capi


def test_library_exports_every_declared_symbol():
    L = F.lib()
    declared = set()
    for h in ("f2v.h", "f2v_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(f2v_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 40
    assert set(capi.ENGINE_SYMBOLS + capi.HOST_SYMBOLS) == declared
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.f2v_abi_version() == 1


def test_no_gpu_fails_loudly(karate):
    """No CPU fallback: on a box without a usable device, engine creation raises."""
    if F.lib().f2v_device_count() > 0:
        pytest.skip("a GPU is present")
    rp, ci = karate
    with pytest.raises(F.F2VError):
        F.Engine(rp, ci, 16)
    a = F.Algorithms(rp, ci, "karate.mtx", "/tmp/", 16)
    with pytest.raises(F.F2VError):
        a.AlgoForce2VecNS(1, 1, 8, 5, 0.02)


def test_rand_stream(oracle):
    want = np.load(os.path.join(GOLDEN, "rand_srand1.npy"))
    g = host.RandStream(1)
    assert [g.rand() for _ in range(len(want))] == want.tolist()


def test_lut(oracle):
    assert np.array_equal(host.build_lut(), oracle.build_lut()[:2048])
    np.testing.assert_allclose(host.build_lut(), np.load(os.path.join(GOLDEN, "ref_lut.npy")), rtol=0, atol=2.4e-7)


@pytest.mark.parametrize("name", ["cora.mtx", "karate.mtx"])
def test_loader_matches_reference_semantics(oracle, name):
    path = os.path.join(GOLDEN, name)
    rp, ci = host.load_mtx(path)
    rp2, ci2 = oracle.load_mtx(path)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    if name == "cora.mtx":
        assert len(rp) - 1 == 2708 and len(ci) == 10858


def _reference_cli_csr(path, workdir):
    """CSR the REFERENCE's own driver and loaders (Test/Force2Vec.cpp:121-127 with IO.h / CSC.h / CSR.h,
    compiled from the reference tree: oracle/ref_csr_dump.cpp -> oracle/_ref/Force2Vec_csrdump) hand to the
    hot-path methods for this file."""
    exe = os.path.join(ROOT, "oracle", "_ref", "Force2Vec_csrdump")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/Force2Vec_csrdump not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "-input", path, "-output", str(workdir) + "/", "-iter", "1", "-option", "5"],
                       cwd=str(workdir), capture_output=True)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
    raw = open(os.path.join(str(workdir), "csr_dump.bin"), "rb").read()
    rows, nnz = (int(x) for x in np.frombuffer(raw, np.uint64, 2))
    return np.frombuffer(raw, np.uint64, rows + 1, 16), np.frombuffer(raw, np.uint32, nnz, 16 + 8 * (rows + 1))


def test_loader_equals_the_reference_cli_loader(tmp_path):
    """f2v_load_mtx builds exactly the CSR the reference's own loader chain builds (same row order, same
    neighbour order, duplicates / self-loops / mirroring treated alike): the byte-equality of the two CLIs'
    .embd files (tests/test_gpu_x_boundary_and_sampler.py) rests on it."""
    rng = np.random.default_rng(7)
    files = [os.path.join(GOLDEN, "cora.mtx"), os.path.join(GOLDEN, "karate.mtx")]
    g = tmp_path / "general.mtx"              # duplicates, self-loops, values, comment lines, unsorted entries
    ent = rng.integers(1, 61, size=(700, 2))
    g.write_text("%%MatrixMarket matrix coordinate real general\n% made by the test\n60 60 700\n" +
                 "".join("%d %d %g\n" % (a, b, rng.random()) for a, b in ent))
    s = tmp_path / "symmetric.mtx"            # lower triangle incl. diagonal entries (dropped by both loaders)
    low = sorted({(int(max(a, b)), int(min(a, b))) for a, b in rng.integers(1, 81, size=(500, 2))})
    s.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n80 80 %d\n" % len(low) +
                 "".join("%d %d\n" % e for e in low))
    files += [str(g), str(s)]
    for k, path in enumerate(files):
        d = tmp_path / ("run%d" % k)
        d.mkdir()
        rp_ref, ci_ref = _reference_cli_csr(path, d)
        rp, ci = host.load_mtx(path)
        assert np.array_equal(rp, rp_ref) and np.array_equal(ci, ci_ref), path


def test_loader_general_and_symmetric(tmp_path, oracle):
    # general: entries kept as they are, including self-loops and duplicates, values ignored
    p = tmp_path / "g.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% c\n4 4 6\n1 2 0.5\n2 1 3\n3 3 1\n1 2 7\n4 1 1\n2 4 1\n")
    rp, ci = host.load_mtx(str(p))
    assert rp.tolist() == [0, 2, 4, 5, 6] and ci.tolist() == [1, 1, 0, 3, 2, 0]
    rp2, ci2 = oracle.load_mtx(str(p))
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    # symmetric: mirrored, self-loops dropped
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n4 4 4\n2 1\n3 3\n4 1\n4 2\n")
    rp, ci = host.load_mtx(str(p))
    assert rp.tolist() == [0, 2, 4, 4, 6] and ci.tolist() == [1, 3, 0, 3, 0, 1]
    with pytest.raises(F.F2VError):
        host.load_mtx(str(tmp_path / "missing.mtx"))
    p = tmp_path / "bad.mtx"
    p.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n4 4 3\n2 1\n")
    with pytest.raises(F.F2VError):
        host.load_mtx(str(p))


@pytest.mark.parametrize("model,bs", [(5, 0), (5, 1), (6, 0), (6, 1), (7, 0)])
def test_streams_match_oracle(oracle, cora, model, bs):
    rp, ci = cora
    n = len(rp) - 1
    g, o = host.RandStream(1), oracle.Rng(1)
    assert np.array_equal(g.init_embeddings(model, n, 24), oracle.init_embeddings(o, model, n, 24))
    B, s = 200, 5
    nb = (n + B - 1) // B
    W = (B + s - 1) if (bs and model != 7) else s
    for _ in range(2):
        if model == 7:
            assert np.array_equal(g.walks(rp, ci), oracle.walks(o, rp, ci))
        neg = g.epoch_negatives(model, n, B, s, bs).reshape(nb, W)
        for b in range(nb):
            assert np.array_equal(neg[b], oracle.draw_negatives(o, model, bs, n, B, s, b)[:W])
    assert g.rand() == o.rand()          # both consumed exactly the same number of draws


def test_embd_writer_matches_reference_text(oracle):
    """Byte-for-byte: re-emitting the values of a reference-written .embd reproduces the file."""
    for opt in (5, 6, 7):
        src = os.path.join(GOLDEN, "karate_opt%d.embd" % opt)
        X = oracle.read_embd(src)
        out = "/tmp/f2v_test_writer_%d.embd" % os.getpid()
        host.write_embd(out, X)
        assert open(out).read() == open(src).read()
        os.remove(out)


def test_embd_writer_format():
    X = np.array([[1e-5, 123456.789, -0.5, 1.0], [0.1, 2.5e-7, 3.0, -1e10]], np.float32)
    out = "/tmp/f2v_test_fmt_%d.embd" % os.getpid()
    host.write_embd(out, X)
    lines = open(out).read().split("\n")
    os.remove(out)
    assert lines[0] == "2 4"
    assert lines[1] == "1 1e-05 123457 -0.5 1 "          # ostream default precision, trailing space
    assert lines[2] == "2 0.1 2.5e-07 3 -1e+10 "
    assert lines[3] == ""


def test_rmat_generator():
    rp, ci = host.rmat_csr(12, 16, 1)
    n = len(rp) - 1
    assert n == 4096 and rp[-1] == len(ci)
    deg = np.diff(rp.astype(np.int64))
    rows = np.repeat(np.arange(n), deg)
    assert (rows != ci).all()                                            # no self-loops
    key = rows.astype(np.int64) * n + ci
    assert (np.diff(key) > 0).all()                                      # sorted rows, no duplicates
    assert np.array_equal(np.sort(ci.astype(np.int64) * n + rows), key)  # symmetric
    rp2, ci2 = host.rmat_csr(12, 16, 1)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)           # deterministic
    rp3, ci3 = host.rmat_csr(12, 16, 2)
    assert not np.array_equal(ci, ci3)
    assert deg.max() > 20 * deg.mean()                                   # skewed


def test_rmat_thread_count_independent():
    code = ("import sys; sys.path.insert(0, %r); from force2vec_b200 import host; import numpy as np;"
            "rp, ci = host.rmat_csr(11, 8, 5); print(int(rp.sum()), int(ci.astype(np.int64).sum()), len(ci))" % ROOT)
    outs = []
    for t in ("1", "4"):
        env = dict(os.environ, OMP_NUM_THREADS=t)
        outs.append(subprocess.check_output(["python", "-c", code], env=env).decode())
    assert outs[0] == outs[1]


def test_mtx_roundtrip(tmp_path, oracle):
    rp, ci = host.rmat_csr(10, 8, 3)
    p = str(tmp_path / "r.mtx")
    host.write_mtx(p, rp, ci)
    a = host.load_mtx(p)
    b = oracle.load_mtx(p)
    assert np.array_equal(a[0], rp) and np.array_equal(a[1], ci)
    assert np.array_equal(b[0], rp) and np.array_equal(b[1], ci)


def _hub_slots(nc):
    # one partial-sum slot per chunk, plus one per fold block of 32 chunks (two-level fold)
    nblk = (nc + 31) // 32
    return nc + (nblk if nblk > 1 else 1)


def _check_plan(rp, batch, chunk, world, par=0):
    n = len(rp) - 1
    deg = np.diff(rp.astype(np.int64))
    nb = (n + batch - 1) // batch
    seen_rows = np.zeros(n, np.int64)
    seen_edges = np.zeros(n, np.int64)
    for rank in range(world):
        pl = host.plan_build(rp, batch, chunk, rank=rank, world=world, par=par)
        assert pl["nb"] == nb
        it, hb = pl["items"], pl["hub"]
        for b in range(nb):
            lo, hi = int(pl["item_ptr"][b]), int(pl["item_ptr"][b + 1])
            items, hubs = it[lo:hi], hb[lo:hi]
            nh = int(pl["n_hub"][b])
            flag = (items["len"] & host.CHUNK_FLAG) != 0
            assert flag[:nh].all() and not flag[nh:].any()              # hub chunks lead the minibatch
            ln = (items["len"] & 0x7fffffff).astype(np.int64)
            assert (ln[nh:] <= chunk).all() and (ln[:nh] <= chunk).all()
            # rank's slice of the minibatch
            blo, bhi = b * batch, min(n, (b + 1) * batch)
            sl = batch // world
            slo = min(blo + rank * sl, bhi) if world > 1 else blo
            shi = min(slo + sl, bhi) if world > 1 else bhi
            assert ((items["v"] >= slo) & (items["v"] < shi)).all()
            # non-hub rows appear once, sorted by descending degree class
            cls = np.where(ln[nh:] == 0, 0, np.floor(np.log2(np.maximum(ln[nh:], 1))).astype(int) + 1)
            assert (np.diff(cls) <= 0).all()
            np.add.at(seen_rows, items["v"][nh:], 1)
            np.add.at(seen_edges, items["v"], ln)
            # hub rows: chunks 0..nchunks-1 in order, contiguous edge ranges, distinct slots
            if nh:
                hv = items["v"][:nh]
                ranges = []
                for v in np.unique(hv):
                    m = hv == v
                    h = hubs[:nh][m]
                    assert h["chunk"].tolist() == list(range(len(h))) and (h["nchunks"] == len(h)).all()
                    assert (h["deg"] == deg[v]).all()
                    e0 = items["e0"][:nh][m].astype(np.int64)
                    assert e0[0] == rp[v] and np.array_equal(e0[1:], e0[:-1] + ln[:nh][m][:-1])
                    seen_rows[v] += 1
                    # partial-sum slots: chunk c at slot0 + c; slot ranges of different rows are disjoint
                    slot0 = int(h["slot"][0])
                    assert h["slot"].tolist() == list(range(slot0, slot0 + len(h)))
                    ranges.append((slot0, slot0 + _hub_slots(len(h))))
                ranges.sort()
                assert ranges[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert (seen_rows == 1).all()
    assert np.array_equal(seen_edges, deg)


@pytest.mark.parametrize("batch,chunk,world,par", [(256, 64, 1, 0), (100, 8, 1, 0), (5000, 16, 1, 0), (256, 64, 2, 0),
                                                    (96, 8, 4, 0), (2048, 4, 1, 0), (512, 64, 1, 100), (2048, 64, 2, 4000)])
def test_plan_covers_every_row_once(batch, chunk, world, par):
    rp, ci = host.rmat_csr(11, 16, 1)
    _check_plan(rp, batch, chunk, world, par)


def test_plan_adaptive_chunk():
    """par > 0: a minibatch with few edges is cut finer (down to 8), one with many keeps `chunk`."""
    rp, ci = host.rmat_csr(12, 16, 1)
    pl = host.plan_build(rp, 512, 64, par=256)
    for b in range(pl["nb"]):
        lo, hi = int(pl["item_ptr"][b]), int(pl["item_ptr"][b + 1])
        ln = (pl["items"]["len"][lo:hi] & 0x7fffffff).astype(np.int64)
        edges = int(rp[min(len(rp) - 1, (b + 1) * 512)] - rp[b * 512])
        want = min(64, max(16, -(-edges // 256)))      # batch 512 <= 8192: lower bound 16
        assert ln.max() <= want


def test_cli_without_gpu_exits_nonzero():
    if F.lib().f2v_device_count() > 0:
        pytest.skip("a GPU is present")
    exe = os.path.join(ROOT, "bin", "Force2Vec")
    if not os.path.exists(exe):
        pytest.skip("bin/Force2Vec not built")
    r = subprocess.run([exe, "-input", os.path.join(GOLDEN, "karate.mtx"), "-output", "/tmp/", "-iter", "1"],
                       capture_output=True, cwd="/tmp")
    assert r.returncode == 1 and b"engine error" in r.stderr
    r = subprocess.run([exe], capture_output=True, cwd="/tmp")
    assert r.returncode == 1 and b"Valid input file needed" in r.stdout


@pytest.mark.parametrize("lg", [0, 1, 2, 3])
def test_shard_row_is_a_dense_bijection_that_spreads_hubs(lg):
    """Placement of the row-sharded mode: vertex -> (shard, local row) must be one-to-one with
    local row = vertex >> lg (dense), the identity for one GPU, and must not put R-MAT's hubs
    (ids with few one-bits) on one GPU the way `vertex mod world` does."""
    L = F.lib()
    world, n = 1 << lg, 1 << 16
    rows = n >> lg
    r = np.array([L.f2v_shard_row(j, lg, rows) for j in range(n)], np.int64)
    assert len(np.unique(r)) == n and r.min() == 0 and r.max() == n - 1          # bijection onto [0, n)
    assert np.array_equal(r % rows, np.arange(n) >> lg)                           # local row = j >> lg
    if lg == 0:
        assert np.array_equal(r, np.arange(n))
        return
    shard = r // rows
    assert np.bincount(shard, minlength=world).tolist() == [rows] * world
    hubs = np.array([0] + [1 << k for k in range(16)] + [(1 << a) | (1 << b) for a in range(16) for b in range(a)])
    cnt = np.bincount(shard[hubs], minlength=world)
    assert cnt.max() <= 2.0 * len(hubs) / world, cnt                              # vs j mod world: almost all on shard 0


def test_shard_row_at_scale26_on_eight_gpus_stays_inside_32_bits():
    """BASELINE config 5 row-sharded: 2^26 vertices over 8 GPUs.  Combined row ids (both tables in one flat range:
    row of table 1 = rows_alloc + row) must fit 32 bits -- the kernels keep them in uint32 and widen only when they
    multiply by the row length -- and the placement must still be a bijection.  The C function (the kernels' own
    __host__ __device__ shard_row) is checked against a vectorised restatement on all 67 M ids."""
    L = F.lib()
    lg, n = 3, 1 << 26
    rows = n >> lg
    j = np.arange(n, dtype=np.uint64)
    h = j >> np.uint64(lg)
    mix = (h * np.uint64(0x9E3779B1)) >> np.uint64(32)
    r = ((j ^ mix) & np.uint64((1 << lg) - 1)) * np.uint64(rows) + h
    assert int(r.max()) == n - 1 and 2 * n - 1 < 2**32                            # table 1's last row: n + (n - 1)
    seen = np.zeros(n, bool)
    seen[r.astype(np.int64)] = True
    assert seen.all()                                                             # onto [0, n): a bijection
    assert np.bincount((r // np.uint64(rows)).astype(np.int64), minlength=8).tolist() == [rows] * 8
    probe = np.concatenate([np.arange(0, 4096), np.arange(n - 4096, n), np.random.default_rng(1).integers(0, n, 20000)])
    assert all(L.f2v_shard_row(int(v), lg, rows) == int(r[v]) for v in probe)


def test_multi_gpu_driver_without_gpus_fails_loudly():
    """f2v_train_gpus asks for more devices than exist (none here): an error, never a CPU fallback."""
    if F.lib().f2v_device_count() >= 2:
        pytest.skip("two GPUs are present")
    from force2vec_b200 import capi
    rp, ci = host.rmat_csr(6, 4, 1)
    alg = F.Algorithms(rp, ci, "g.mtx", "/tmp/", 16)
    alg.gpus = 2
    with pytest.raises(capi.F2VError):
        alg.AlgoForce2VecNS(1, 0, 16, 5, 0.02, write=False)


@pytest.mark.parametrize("assign", [2, 4, 3, 5])
def test_plan_item_orders_cover_every_row_once(assign):
    """Item-order flags (2 = light rows right after the hub chunks, 4 = light rows interleaved; | 1 =
    balanced ownership): a permutation of the default plan's items, minibatch by minibatch."""
    rp, ci = host.rmat_csr(11, 16, 1)
    world = 2 if assign & 1 else 1
    for rank in range(world):
        base = host.plan_build(rp, 512, 32, rank=rank, world=world, assign=assign & 1)
        alt = host.plan_build(rp, 512, 32, rank=rank, world=world, assign=assign)
        assert np.array_equal(base["item_ptr"], alt["item_ptr"]) and np.array_equal(base["n_hub"], alt["n_hub"])
        for b in range(base["nb"]):
            lo, hi = int(base["item_ptr"][b]), int(base["item_ptr"][b + 1])
            nh = int(base["n_hub"][b])
            assert np.array_equal(base["items"][lo:lo + nh], alt["items"][lo:lo + nh])      # hub chunks unchanged
            a = np.sort(base["items"][lo + nh:hi], order=["v", "e0"])
            c = np.sort(alt["items"][lo + nh:hi], order=["v", "e0"])
            assert np.array_equal(a, c)
            assert np.array_equal(alt["hub"]["deg"][lo + nh:hi], alt["items"]["len"][lo + nh:hi])


@pytest.mark.parametrize("model", [5, 6])
def test_parallel_init_equals_serial_stream(oracle, model):
    """Initial embeddings longer than one chunk (2^20 draws) are generated by independent threads
    from jump-ahead states of the glibc-compatible generator: the same numbers in the same places as
    the serial reference loop, and the stream continues where the serial loop would have left it."""
    n, d = 33000, 100                       # 3.3 M draws: four chunks, the last one partial
    g, o = host.RandStream(1), oracle.Rng(1)
    for _ in range(7):                      # start from a state that is not the seed state
        assert g.rand() == o.rand()
    assert np.array_equal(g.init_embeddings(model, n, d), oracle.init_embeddings(o, model, n, d))
    assert [g.rand() for _ in range(64)] == [o.rand() for _ in range(64)]


@pytest.mark.parametrize("model,B", [(5, 1024), (6, 2048), (5, 5000)])
def test_bs1_negatives_with_jump_ahead_match_serial_draws(oracle, cora, model, B):
    """bs=1: the reference consumes s*batch draws per minibatch and reads the first batch+s-1.  For
    large batches the unread draws are skipped by jump-ahead and the minibatches are drawn in
    parallel: same kept values as the oracle's serial loop, same stream position afterwards."""
    rp, ci = cora
    n, s = len(rp) - 1, 5
    nb, W = (n + B - 1) // B, B + s - 1
    g, o = host.RandStream(1), oracle.Rng(1)
    for _ in range(2):
        neg = g.epoch_negatives(model, n, B, s, 1).reshape(nb, W)
        for b in range(nb):
            assert np.array_equal(neg[b], oracle.draw_negatives(o, model, 1, n, B, s, b)[:W])
    assert [g.rand() for _ in range(8)] == [o.rand() for _ in range(8)]


def test_fast_g6_formatter_equals_printf():
    """The .embd writer formats values itself ("%.6g" = ostream's default, algorithms.h:126-134);
    compare with the C library's formatting on random values of every magnitude, ties, zeros, inf."""
    import ctypes as C
    L = F.lib()
    buf = C.create_string_buffer(40)
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.random(60000, dtype=np.float32) * 2 - 1,
        rng.standard_normal(30000).astype(np.float32) * 1e-5,
        (rng.random(30000) * 1e7).astype(np.float32),
        rng.integers(0, 2 ** 32, 60000, dtype=np.uint64).astype(np.uint32).view(np.float32),     # any bit pattern
        np.array([0.0, -0.0, 1, -1, 0.5, 1e-5, 9.99999e-5, 1e-4, 123456, 1234567, 999999.5, 999999.4, 0.1, 1e6, 1e5,
                  99999.95, 1e-10, 3.4e38, 1e-40, np.inf, -np.inf, 0.02, -0.1, 5, -5, 100000, 999999, 1000000,
                  2.5e-5, 0.000123456789, 0.15625, 2.5, 1.5e-7], np.float32)])
    vals = vals[~np.isnan(vals)]            # printf prints the sign of a NaN, Python does not
    for v in vals:
        L.f2v_format_g6(C.c_float(float(v)), buf)
        assert buf.value.decode() == "%.6g" % float(v), repr(float(v))


def test_binary_csr_cache_round_trip(tmp_path):
    rp, ci = host.rmat_csr(10, 8, 4)
    p = str(tmp_path / "g.f2vcsr")
    host.write_csr(p, rp, ci)
    rp2, ci2 = host.load_csr(p)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    raw = bytearray(open(p, "rb").read())
    raw[0] = ord("X")                                   # wrong magic
    open(p, "wb").write(raw)
    with pytest.raises(F.F2VError):
        host.load_csr(p)
    raw[0] = ord("F")
    raw[-4:] = (1 << 31).to_bytes(4, "little")          # a column id out of range
    open(p, "wb").write(raw)
    with pytest.raises(F.F2VError):
        host.load_csr(p)
    with pytest.raises(F.F2VError):
        host.load_csr(str(tmp_path / "missing.f2vcsr"))


def test_threaded_loader_on_a_file_larger_than_1MiB(tmp_path, oracle):
    """The entry lines of a file above 1 MiB are parsed by all host threads (pieces cut at line
    boundaries, the first `cnt` entries of the file kept): symmetric file, trailing blank lines, more
    entry lines than the size line declares, several OMP thread counts -- all equal to the serial
    oracle loader and to the graph written."""
    import subprocess
    import sys
    rp, ci = host.rmat_csr(14, 8, 7)
    n = len(rp) - 1
    p = str(tmp_path / "big.mtx")
    host.write_mtx(p, rp, ci)
    raw = open(p).read()
    assert len(raw) > (1 << 20)
    # surplus entries after the declared count (ignored) and trailing blank lines
    raw += "1 2\n3 1\n\n\n\r\n"
    open(p, "w").write(raw)
    a = host.load_mtx(p)
    assert np.array_equal(a[0], rp) and np.array_equal(a[1], ci)
    b = oracle.load_mtx(p)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; from force2vec_b200 import host; "
            "rp, ci = host.load_mtx(%r); np.save(%r, rp); np.save(%r, ci)")
    for nt in (1, 3, 8):
        o1, o2 = str(tmp_path / ("rp%d.npy" % nt)), str(tmp_path / ("ci%d.npy" % nt))
        subprocess.check_call([sys.executable, "-c", code % (ROOT, p, o1, o2)], env=dict(os.environ, OMP_NUM_THREADS=str(nt)))
        assert np.array_equal(np.load(o1), rp) and np.array_equal(np.load(o2), ci), nt
    # an entry out of range among the first `cnt` is an error with a message; a directory and a FIFO are refused
    bad = raw.replace("\n", "\n%d 1\n" % (n + 5), 3)
    open(p, "w").write(bad)
    with pytest.raises(F.F2VError) as ei:
        host.load_mtx(p)
    assert "load_mtx" in str(ei.value)
    with pytest.raises(F.F2VError) as ei:
        host.load_mtx(str(tmp_path))
    assert "regular file" in str(ei.value)
    fifo = str(tmp_path / "fifo.mtx")
    os.mkfifo(fifo)
    with pytest.raises(F.F2VError):
        host.load_mtx(fifo)


def test_host_errors_carry_a_message(tmp_path):
    with pytest.raises(F.F2VError) as ei:
        host.load_csr(str(tmp_path / "nope.f2vcsr"))
    assert "f2v_load_csr" in str(ei.value)


def test_bench_arms_share_config_and_thread_setup(monkeypatch):
    """bench.py: both arms build `config` with the same function (the driver compares the dicts), the default
    workload is the configuration BASELINE quotes "1/2/4/8 B200" on, and the host thread count is taken back from
    torch.distributed.run's OMP_NUM_THREADS=1 before libgomp is loaded."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.scale, a.model, a.dim, a.bs) == (24, 5, 128, 1)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "4"])
    b = bench.parse()
    assert bench.config_of(a, 1 << 24, 520757804) == bench.config_of(b, 1 << 24, 520757804)
    assert bench.workload_name(a).startswith("rmat24_ef16_seed1 option5(tForce2Vec) d128 s5 bs1 B")
    cap = bench.committed_capture(a, 1)
    assert cap and cap["dram_bytes_per_epoch"] < cap["algorithmic_bytes_per_epoch"]      # traffic <= algorithmic bytes
    assert bench.committed_capture(a, 8) is None                                          # never a constant across N
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "4")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.delenv("F2V_KEEP_OMP_NUM_THREADS", raising=False)
    bench.host_threads()
    assert int(os.environ["OMP_NUM_THREADS"]) == bench.usable_cores() == len(os.sched_getaffinity(0))
    monkeypatch.setenv("LOCAL_RANK", "3")
    bench.host_threads()
    assert int(os.environ["OMP_NUM_THREADS"]) == max(1, bench.usable_cores() // 4)


def test_plan_at_scale26_keeps_64bit_edge_offsets():
    """BASELINE config 5 on the host side: n = 2^26 vertices and more than 2^32 CSR entries -- past the reference's
    32-bit `INDEXTYPE` (sample/algorithms.h:40, utility.h:128 `my_malloc(unsigned int)`).  The work plan of one of
    8 ranks (degree-balanced ownership, hub rows cut into chunks) must keep every edge offset as a 64-bit value,
    cover this rank's rows exactly once and give the 8 ranks equal shares of the edges.  (A synthetic power-law
    degree sequence stands in for the R-MAT graph: the plan is a function of rowptr only.)"""
    n, batch, world, chunk = 1 << 26, 262144, 8, 256
    rng = np.random.default_rng(5)
    deg = np.minimum((rng.pareto(1.2, n) * 20).astype(np.uint64), 3_000_000)
    deg[:4096] += rng.integers(100_000, 900_000, 4096).astype(np.uint64)      # hubs at the low ids, like un-permuted R-MAT
    rp = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=rp[1:])
    assert int(rp[-1]) > (1 << 32)
    shares = []
    for rank in (0, 5):
        pl = host.plan_build(rp, batch, chunk, rank=rank, world=world, par=9472, assign=1 | 2)
        it = pl["items"]
        ln = (it["len"] & 0x7fffffff).astype(np.int64)
        assert pl["nb"] == n // batch
        assert int(it["e0"].max()) > (1 << 32)                                  # offsets beyond 32 bits survive
        plain = (it["len"] & host.CHUNK_FLAG) == 0
        assert np.array_equal(it["e0"][plain], rp[it["v"][plain]])              # a row's item starts at its rowptr
        assert np.array_equal(ln[plain], deg[it["v"][plain]].astype(np.int64))
        rows = np.unique(it["v"])
        assert len(rows) == plain.sum() + len(np.unique(it["v"][~plain]))       # every owned row: one item or one chunk set
        shares.append(int(ln.sum()))
        assert ln[~plain].max() <= chunk
    total = int(rp[-1])
    for sh in shares:
        assert abs(sh - total / world) < 0.02 * total / world                   # LPT partition: equal edge shares


def test_create_validates_the_csr_before_touching_a_device():
    """f2v_create checks rowptr / colids on the host (all threads) and names the first offending entry -- on a
    box without a GPU too, because the check comes before any CUDA call."""
    rp = np.array([0, 2, 3, 5], np.uint64)
    with pytest.raises(F.F2VError) as ei:
        F.Engine(rp, np.array([1, 2, 0, 7, 9], np.uint32), 8)
    assert "colids[3] = 7 out of range" in str(ei.value)
    with pytest.raises(F.F2VError) as ei:
        F.Engine(np.array([0, 3, 2, 5], np.uint64), np.array([1, 2, 0, 0, 1], np.uint32), 8)
    assert "rowptr not monotone at row 1" in str(ei.value)


def _csr_from_edges(n, edges):
    adj = [set() for _ in range(n)]
    for a, b in edges:
        if a != b:
            adj[a].add(b)
            adj[b].add(a)
    rp = np.zeros(n + 1, np.uint64)
    ci = []
    for v in range(n):
        ci += sorted(adj[v])
        rp[v + 1] = len(ci)
    return rp, np.asarray(ci, np.uint32)


def test_interleaved_walk_sampler_is_the_serial_loop_bit_for_bit(oracle, cora):
    """f2v_draw_walks keeps 24 walks in flight (round-robin half-steps with prefetches) and lets a younger walk
    start at the stream position its elders are predicted to leave; a wrong prediction moves / restarts the younger
    walks.  Whatever the graph does to the predictions -- every degree <= 2 (all wrong), isolated start vertices
    whose 'edge index' is their own id, fewer CSR entries than vertices, no edges at all, R-MAT
    skew -- the walks and the position of the stream afterwards equal the reference's serial loop
    (algorithms.cpp:1097-1118 as restated in oracle/f2v_oracle.c), over several epochs and seeds."""
    rng = np.random.default_rng(3)
    graphs = {
        "path": _csr_from_edges(50, [(i, i + 1) for i in range(49)]),
        "star": _csr_from_edges(40, [(0, i) for i in range(1, 40)]),
        "ring_of_degree_4": _csr_from_edges(30, [(i, (i + 1) % 30) for i in range(30)] + [(i, (i + 2) % 30) for i in range(30)]),
        "clique_and_56_isolated": _csr_from_edges(64, [(i, j) for i in range(8) for j in range(i + 1, 8)]),
        "no_edges": (np.zeros(11, np.uint64), np.zeros(0, np.uint32)),
        "sparse_random": _csr_from_edges(3000, [tuple(int(x) for x in rng.integers(0, 3000, 2)) for _ in range(2500)]),
        "rmat10": host.rmat_csr(10, 2, 5), "rmat13": host.rmat_csr(13, 16, 5), "rmat15": host.rmat_csr(15, 4, 5),
        "cora": cora,
    }
    for name, (rp, ci) in graphs.items():
        for seed in (1, 7):
            g, o = host.RandStream(seed), oracle.Rng(seed)
            for _ in range(seed % 5):                       # start somewhere inside the stream
                assert g.rand() == o.rand()
            for epoch in range(3):
                got, want = g.walks(rp, ci).copy(), oracle.walks(o, rp, ci)
                assert np.array_equal(got, want), (name, seed, epoch)
                assert [g.rand() for _ in range(40)] == [o.rand() for _ in range(40)], (name, seed, epoch)

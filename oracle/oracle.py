"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to (a) oracle/liboracle.so, the plain-C restatement of the reference's
option 5/6/7 path (oracle/f2v_oracle.c), and (b) oracle/_ref/libf2vref*.so, the UNMODIFIED
reference compiled from /root/reference with our extern "C" shim (oracle/ref_shim.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under force2vec_b200/ does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TDIST, SIGMOID, WALK = 5, 6, 7
WALKLEN = 5
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the checkers (make -C oracle).  Building the checker is not using it."""
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []),
                          stdout=subprocess.DEVNULL)


class _Rng(C.Structure):
    _fields_ = [("r", C.c_int32 * 31), ("f", C.c_int), ("b", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.f2vo_srand.argtypes = [C.POINTER(_Rng), C.c_uint32]
        L.f2vo_rand.argtypes = [C.POINTER(_Rng)]
        L.f2vo_rand.restype = C.c_int32
        L.f2vo_init_embeddings.argtypes = [C.POINTER(_Rng), C.c_int, C.c_uint64, C.c_uint32, _f32p]
        L.f2vo_build_lut.argtypes = [_f32p]
        L.f2vo_fast_sm.argtypes = [_f32p, C.c_float]
        L.f2vo_fast_sm.restype = C.c_float
        L.f2vo_walks.argtypes = [C.POINTER(_Rng), C.c_uint64, C.c_uint64, _u64p, _u32p, _u32p]
        L.f2vo_draws_per_batch.argtypes = [C.c_int, C.c_uint32, C.c_uint32]
        L.f2vo_draws_per_batch.restype = C.c_uint64
        L.f2vo_draw_negatives.argtypes = [C.POINTER(_Rng), C.c_int, C.c_int, C.c_uint64, C.c_uint32,
                                          C.c_uint32, C.c_uint64, _u32p]
        L.f2vo_step.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint32, _u64p, _u32p, _f32p,
                                C.c_uint64, C.c_uint64, _u32p, C.c_uint32, C.c_float, _f32p,
                                C.c_void_p, C.c_int]
        L.f2vo_run.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, _u64p, _u32p, C.c_uint32,
                               C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_int,
                               _f32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.f2vo_run.restype = C.c_int
        L.f2vo_counter_rand.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.f2vo_counter_rand.restype = C.c_uint32
        L.f2vo_walks_counter.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _u64p, _u32p, _u32p]
        _lib = L
    return _lib


class Rng:
    """glibc srand(seed)/rand() stream (restated, no libc state)."""

    def __init__(self, seed=1):
        self.s = _Rng()
        lib().f2vo_srand(C.byref(self.s), seed)

    def rand(self):
        return lib().f2vo_rand(C.byref(self.s))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_lut():
    t = np.zeros(2049, np.float32)
    lib().f2vo_build_lut(t)
    return t


def init_embeddings(rng, model, n, dim):
    X = np.empty((n, dim), np.float32)
    lib().f2vo_init_embeddings(C.byref(rng.s), model, n, dim, X)
    return X


def draws_per_batch(bs, batch, s):
    return int(lib().f2vo_draws_per_batch(bs, batch, s))


def draw_negatives(rng, model, bs, n, batch, s, b):
    cnt = s if model == WALK else draws_per_batch(bs, batch, s)
    idx = np.empty(max(cnt, 1), np.uint32)
    lib().f2vo_draw_negatives(C.byref(rng.s), model, bs, n, batch, s, b, idx)
    return idx[:cnt]


def walks(rng, rowptr, colids):
    n = len(rowptr) - 1
    w = np.empty((n, WALKLEN), np.uint32)
    lib().f2vo_walks(C.byref(rng.s), n, len(colids), rowptr, _pad(colids), w)
    return w


def walks_counter(seed, epoch, rowptr, colids):
    n = len(rowptr) - 1
    w = np.empty((n, WALKLEN), np.uint32)
    lib().f2vo_walks_counter(seed, epoch, n, len(colids), rowptr, _pad(colids), w)
    return w


def _pad(colids):
    return colids if len(colids) else np.zeros(1, np.uint32)


def step(model, bs, rowptr, colids, X, lo, hi, idx, s, lr, lut=None, walks=None, threads=0):
    """One Jacobi minibatch [lo,hi), in place on X (float32 [n,dim], C-contiguous)."""
    n, dim = X.shape
    if lut is None:
        lut = build_lut()
    idx = np.ascontiguousarray(idx, np.uint32)
    if len(idx) == 0:
        idx = np.zeros(1, np.uint32)
    lib().f2vo_step(model, bs, n, dim, rowptr, _pad(colids), X, lo, hi, idx, s, lr, lut,
                    _ptr(walks), threads)
    return X


def run(model, bs, rowptr, colids, dim, iterations, batch, s, lr, seed=1, threads=0,
        want_init=False, want_logs=False):
    """Whole reference run (srand(seed) -> init -> epochs).  Returns dict."""
    n = len(rowptr) - 1
    if model == WALK:
        bs = 0
    X = np.empty((n, dim), np.float32)
    X0 = np.empty((n, dim), np.float32) if want_init else None
    nb = (n + batch - 1) // batch
    dpb = draws_per_batch(bs, batch, s)
    neg = np.empty((iterations, nb, dpb), np.uint32) if want_logs else None
    wl = np.empty((iterations, n, WALKLEN), np.uint32) if (want_logs and model == WALK) else None
    rc = lib().f2vo_run(model, bs, n, len(colids), rowptr, _pad(colids), dim, iterations, batch, s, lr,
                        seed, threads, X, _ptr(X0), _ptr(neg), _ptr(wl))
    if rc != 0:
        raise ValueError("f2vo_run: bad arguments")
    return {"X": X, "X0": X0, "neg": neg, "walks": wl}


# ----------------------------------------------------------------------------------------
# reference-semantics MatrixMarket loader (restates sample/IO.h:59-156 + CSC.h:146-188 +
# CSR.h:154-186) -- for tests of the product's C++ loader.
def load_mtx(path):
    sym = False
    with open(path, "r") as f:
        line = f.readline()
        while line.startswith("%"):
            if "symmetric" in line[1:]:
                sym = True
            line = f.readline()
        m, n, nnz = [int(x) for x in line.split()[:3]]
        rows, cols = [], []
        for _ in range(nnz):
            line = f.readline()
            if not line:
                break
            tok = line.split(" ")
            r, c = int(tok[0]) - 1, int(tok[1]) - 1
            if sym:
                if r == c:
                    continue                      # self-loops dropped (IO.h:130-134)
                rows += [r, c]
                cols += [c, r]                    # mirrored (IO.h:122-129)
            else:
                rows.append(r)
                cols.append(c)
    rows = np.asarray(rows, np.int64)
    cols = np.asarray(cols, np.int64)
    order = np.lexsort((cols, rows))              # ascending colids within each row
    rows, cols = rows[order], cols[order]
    rowptr = np.zeros(m + 1, np.uint64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.uint64)
    return rowptr, cols.astype(np.uint32)


# ----------------------------------------------------------------------------------------
# the unmodified reference (oracle/_ref)
_ref = {}


def ref_available(avx512=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libf2vref_avx512.so" if avx512 else "libf2vref.so"))


def host_has_avx512():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " avx512f" in flags and " avx512dq" in flags


def ref_lib(avx512=False):
    key = bool(avx512)
    if key not in _ref:
        name = "libf2vref_avx512.so" if avx512 else "libf2vref.so"
        L = C.CDLL(os.path.join(HERE, "_ref", name))
        L.f2vref_run.argtypes = [C.c_uint32, C.c_uint32, _u32p, _u32p, C.c_uint32, C.c_int, C.c_int,
                                 C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_char_p,
                                 C.c_void_p]
        L.f2vref_run.restype = C.c_double
        L.f2vref_lut.argtypes = [_f32p]
        _ref[key] = L
    return _ref[key]


def ref_run(option, bs, rowptr, colids, dim, iterations, batch, s, lr, threads=1, avx512=False,
            want_X=True):
    """Run the UNMODIFIED reference in memory.  Returns (X, wall_seconds_reported_by_reference)."""
    n = len(rowptr) - 1
    X = np.empty((n, dim), np.float32) if want_X else None
    rp = np.ascontiguousarray(rowptr, np.uint32)
    sec = ref_lib(avx512).f2vref_run(n, len(colids), rp, _pad(np.ascontiguousarray(colids, np.uint32)),
                                     dim, option, bs, iterations, threads, batch, s, lr,
                                     b"/nonexistent_f2vref/", _ptr(X))
    if sec < 0:
        raise ValueError("f2vref_run failed: %r" % sec)
    return X, sec


def ref_lut():
    t = np.zeros(2048, np.float32)
    ref_lib().f2vref_lut(t)
    return t


def ref_cli(avx512=False):
    return os.path.join(HERE, "_ref", "Force2Vec_avx512" if avx512 else "Force2Vec")


def read_embd(path):
    """Parse the reference's text .embd (algorithms.h:118-136): 'N D' then 'id v1 .. vD '."""
    with open(path) as f:
        n, d = [int(x) for x in f.readline().split()]
        X = np.zeros((n, d), np.float32)
        for line in f:
            tok = line.split()
            if not tok:
                continue
            X[int(tok[0]) - 1] = np.asarray(tok[1:1 + d], np.float32)
    return X

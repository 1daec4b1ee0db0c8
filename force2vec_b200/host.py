"""Host-side helpers of the drop-in (include/f2v_host.h): the glibc-compatible rand() stream
and the samplers that consume it, MatrixMarket loader, .embd writer, R-MAT generator."""
import ctypes as C
import numpy as np
from .capi import lib, check, LUT_SIZE, WALKLEN, WALK


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RandStream:
    """srand(seed)/rand() of glibc, as consumed by the reference (Test/Force2Vec.cpp:126)."""

    def __init__(self, seed=1):
        self._h = lib().f2v_rng_create(seed)
        if not self._h:
            raise MemoryError("f2v_rng_create")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().f2v_rng_destroy(self._h)
            self._h = None

    def rand(self):
        return lib().f2v_rng_next(self._h)

    def init_embeddings(self, model, n, dim, out=None):
        X = np.empty((n, dim), np.float32) if out is None else out
        check(lib().f2v_init_embeddings(self._h, model, n, dim, _p(X)), "f2v_init_embeddings")
        return X

    def epoch_negatives(self, model, n, batch, s, bs_mode, out=None):
        cnt = neg_stream_len(model, n, batch, s, bs_mode)
        idx = np.empty(max(cnt, 1), np.uint32) if out is None else out
        check(lib().f2v_draw_epoch_negatives(self._h, model, n, batch, s, bs_mode, _p(idx)),
              "f2v_draw_epoch_negatives")
        return idx[:cnt]

    def walks(self, rowptr, colids, out=None):
        n = len(rowptr) - 1
        w = np.empty((n, WALKLEN), np.uint32) if out is None else out
        ci = colids if len(colids) else np.zeros(1, np.uint32)
        check(lib().f2v_draw_walks(self._h, n, len(colids), _p(rowptr), _p(ci), _p(w)), "f2v_draw_walks")
        return w


def neg_stream_len(model, n, batch, s, bs_mode):
    return int(lib().f2v_neg_stream_len(model, n, batch, s, 0 if model == WALK else bs_mode))


def build_lut():
    t = np.empty(LUT_SIZE, np.float32)
    check(lib().f2v_build_lut(_p(t)), "f2v_build_lut")
    return t


def _take_csr(n, nnz, rp, ci):
    rowptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint64)), shape=(n.value + 1,)).copy()
    colids = np.ctypeslib.as_array(C.cast(ci, C.POINTER(C.c_uint32)), shape=(max(nnz.value, 1),)).copy()[:nnz.value]
    lib().f2v_free(rp)
    lib().f2v_free(ci)
    return rowptr, colids


def load_mtx(path):
    n, nnz, rp, ci = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p()
    check(lib().f2v_load_mtx(path.encode(), C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci)),
          "f2v_load_mtx(%s)" % path)
    return _take_csr(n, nnz, rp, ci)


def write_csr(path, rowptr, colids):
    ci = colids if len(colids) else np.zeros(1, np.uint32)
    check(lib().f2v_write_csr(path.encode(), len(rowptr) - 1, len(colids), _p(rowptr), _p(ci)), "f2v_write_csr")


def load_csr(path):
    n, nnz, rp, ci = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p()
    check(lib().f2v_load_csr(path.encode(), C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci)),
          "f2v_load_csr(%s)" % path)
    return _take_csr(n, nnz, rp, ci)


def rmat_csr(scale, edge_factor=16, seed=1):
    n, nnz, rp, ci = C.c_uint64(), C.c_uint64(), C.c_void_p(), C.c_void_p()
    check(lib().f2v_rmat_csr(scale, edge_factor, seed, C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci)),
          "f2v_rmat_csr")
    return _take_csr(n, nnz, rp, ci)


def write_embd(path, X):
    X = np.ascontiguousarray(X, np.float32)
    check(lib().f2v_write_embd(path.encode(), _p(X), X.shape[0], X.shape[1]), "f2v_write_embd")


def write_mtx(path, rowptr, colids):
    ci = colids if len(colids) else np.zeros(1, np.uint32)
    check(lib().f2v_write_mtx(path.encode(), len(rowptr) - 1, _p(rowptr), _p(ci)), "f2v_write_mtx")


ITEM_DTYPE = np.dtype([("v", np.uint32), ("len", np.uint32), ("e0", np.uint64)])
HUB_DTYPE = np.dtype([("chunk", np.uint32), ("nchunks", np.uint32), ("slot", np.uint32), ("deg", np.uint32)])
CHUNK_FLAG = 0x80000000


def plan_build(rowptr, batch, chunk=64, walk=False, rank=0, world=1, first_row=0, nrows=None, par=0, assign=0):
    """The engine's per-minibatch work plan (f2v_plan_build).  Returns dict of numpy arrays."""
    n = len(rowptr) - 1
    nrows = n - first_row if nrows is None else nrows
    nb, ip, nh, it, hb = C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    check(lib().f2v_plan_build(_p(rowptr), first_row, nrows, batch, chunk, par, int(walk), rank, world, assign,
                               C.byref(nb), C.byref(ip), C.byref(nh), C.byref(it), C.byref(hb)), "f2v_plan_build")
    nbv = nb.value
    item_ptr = np.ctypeslib.as_array(C.cast(ip, C.POINTER(C.c_uint64)), shape=(nbv + 1,)).copy()
    n_hub = np.ctypeslib.as_array(C.cast(nh, C.POINTER(C.c_uint32)), shape=(max(nbv, 1),)).copy()[:nbv]
    total = int(item_ptr[-1])
    raw_i = np.ctypeslib.as_array(C.cast(it, C.POINTER(C.c_uint8)), shape=(max(total, 1) * 16,)).copy()
    raw_h = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_uint8)), shape=(max(total, 1) * 16,)).copy()
    for q in (ip, nh, it, hb):
        lib().f2v_free(q)
    return {"nb": nbv, "item_ptr": item_ptr, "n_hub": n_hub,
            "items": raw_i.view(ITEM_DTYPE)[:total], "hub": raw_h.view(HUB_DTYPE)[:total]}


def rmat_csr_cached(scale, edge_factor=16, seed=1, cache_dir="/dev/shm"):
    """rmat_csr with a page-cache copy under `cache_dir` (two .npy files, memory-mapped read-only),
    so that the ranks of one node -- and successive commands on one box -- build the graph once
    (multi-rank callers let one rank call this first, then a barrier, then the others)."""
    import os
    base = os.path.join(cache_dir, "f2v_rmat%d_ef%d_seed%d" % (scale, edge_factor, seed))
    rp_path, ci_path = base + ".rowptr.npy", base + ".colids.npy"
    if not (os.path.exists(rp_path) and os.path.exists(ci_path)):
        rp, ci = rmat_csr(scale, edge_factor, seed)
        try:
            tmp = "%s.%d.tmp" % (base, os.getpid())
            np.save(tmp + ".rp.npy", rp)
            np.save(tmp + ".ci.npy", ci)
            os.replace(tmp + ".ci.npy", ci_path)
            os.replace(tmp + ".rp.npy", rp_path)       # rowptr last: its presence marks the pair complete
        except OSError:
            return rp, ci
        del rp, ci
    return np.load(rp_path, mmap_mode="r"), np.load(ci_path, mmap_mode="r")

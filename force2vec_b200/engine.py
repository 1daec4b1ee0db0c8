"""Engine: the GPU force-step engine behind the C ABI (include/f2v.h).
Algorithms: Python mirror of the reference's `class algorithms` for options 5/6/7
(/root/reference/sample/algorithms.h:60-70,86-90,118-136) -- same method names, argument
meaning and return value ({wall seconds}); it drives f2v_train (the C++ host driver)."""
import ctypes as C
import os
import numpy as np
from . import capi
from .capi import lib, check, TDIST, SIGMOID, WALK, WALKLEN
from . import host


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, rowptr, colids, dim, device=0):
        self.rowptr = np.ascontiguousarray(rowptr, np.uint64)
        self.colids = np.ascontiguousarray(colids, np.uint32)
        self.n = len(self.rowptr) - 1
        self.nnz = len(self.colids)
        self.dim = int(dim)
        self._h = C.c_void_p()
        ci = self.colids if self.nnz else np.zeros(1, np.uint32)
        check(lib().f2v_create(C.byref(self._h), device, self.n, self.nnz, _p(self.rowptr), _p(ci), self.dim),
              "f2v_create")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().f2v_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state
    def set_stream(self, cuda_stream):
        check(lib().f2v_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None), "f2v_set_stream")

    def sync(self):
        check(lib().f2v_sync(self._h), "f2v_sync")

    def set_embeddings(self, X):
        X = np.ascontiguousarray(X, np.float32)
        assert X.shape == (self.n, self.dim)
        check(lib().f2v_set_embeddings(self._h, _p(X)), "f2v_set_embeddings")

    def get_embeddings(self, out=None):
        X = np.empty((self.n, self.dim), np.float32) if out is None else out
        check(lib().f2v_get_embeddings(self._h, _p(X)), "f2v_get_embeddings")
        return X

    def get_rows(self, first, nrows):
        X = np.empty((nrows, self.dim), np.float32)
        check(lib().f2v_get_rows(self._h, first, nrows, _p(X)), "f2v_get_rows")
        return X

    def checksum(self):
        """64-bit device-side checksum of the live table (f2v_checksum)."""
        h = C.c_uint64()
        check(lib().f2v_checksum(self._h, C.byref(h)), "f2v_checksum")
        return int(h.value)

    def device_memory(self):
        f, t = C.c_uint64(), C.c_uint64()
        check(lib().f2v_device_memory(self._h, C.byref(f), C.byref(t)), "f2v_device_memory")
        return int(f.value), int(t.value)

    def set_lut(self, table=None):
        t = host.build_lut() if table is None else np.ascontiguousarray(table, np.float32)
        check(lib().f2v_set_lut(self._h, _p(t), len(t)), "f2v_set_lut")

    def set_negatives(self, idx):
        idx = np.ascontiguousarray(idx, np.uint32).ravel()
        self._neg_keep = idx          # async copy: keep the host buffer alive
        check(lib().f2v_set_negatives(self._h, _p(idx) if len(idx) else None, len(idx)), "f2v_set_negatives")

    def set_negative_offset(self, offset):
        check(lib().f2v_set_negative_offset(self._h, offset), "f2v_set_negative_offset")

    def set_walks(self, walks):
        w = np.ascontiguousarray(walks, np.uint32)
        assert w.size == self.n * WALKLEN
        self._walks_keep = w
        check(lib().f2v_set_walks(self._h, _p(w)), "f2v_set_walks")

    def get_walks(self):
        w = np.empty((self.n, WALKLEN), np.uint32)
        check(lib().f2v_get_walks(self._h, _p(w)), "f2v_get_walks")
        return w

    def sample_walks(self, seed, epoch):
        check(lib().f2v_sample_walks(self._h, seed, epoch), "f2v_sample_walks")

    # ---- hot path
    def step(self, model, first_row, nrows, neg_idx, s, bs_mode, lr, walks=None):
        idx = np.ascontiguousarray(neg_idx, np.uint32).ravel()
        w = None if walks is None else np.ascontiguousarray(walks, np.uint32)
        check(lib().f2v_step(self._h, model, first_row, nrows, _p(idx) if len(idx) else None, s, bs_mode, lr, _p(w)),
              "f2v_step")

    def run_epoch(self, model, batch, s, bs_mode, lr, chunk=0):
        check(lib().f2v_run_epoch(self._h, model, batch, s, bs_mode, lr, chunk), "f2v_run_epoch")

    def run_epoch_host(self, model, batch, s, bs_mode, lr, X_in=None, neg=None, walks=None, X_out=None, chunk=0):
        cnt = 0 if neg is None else neg.size
        check(lib().f2v_run_epoch_host(self._h, model, batch, s, bs_mode, lr, chunk, _p(X_in), _p(neg), cnt,
                                       _p(walks), _p(X_out)), "f2v_run_epoch_host")

    def set_epoch_mode(self, mode):
        check(lib().f2v_set_epoch_mode(self._h, mode), "f2v_set_epoch_mode")

    def set_option(self, name, value):
        check(lib().f2v_set_option(self._h, name.encode(), int(value)), "f2v_set_option")

    def launch_count(self):
        return int(lib().f2v_launch_count(self._h))

    def last_epoch_ms(self):
        ms = C.c_float()
        check(lib().f2v_last_epoch_ms(self._h, C.byref(ms)), "f2v_last_epoch_ms")
        return ms.value

    def trace_ms(self, cap=1 << 16):
        """Per-minibatch device times (ms) of the last epoch; needs set_option("trace", 1)."""
        buf = np.zeros(cap, np.float32)
        cnt = C.c_uint32()
        check(lib().f2v_trace_ms(self._h, _p(buf), cap, C.byref(cnt)), "f2v_trace_ms")
        return buf[:cnt.value].copy()

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        check(lib().f2v_comm_unique_id(buf), "f2v_comm_unique_id")
        return bytes(buf)

    def comm_init(self, id128, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(id128)
        check(lib().f2v_comm_init(self._h, buf, rank, world), "f2v_comm_init")


    def comm_peer_export(self):
        buf = (C.c_char * capi.PEER_BLOB)()
        check(lib().f2v_comm_peer_export(self._h, buf), "f2v_comm_peer_export")
        return bytes(buf)

    def comm_peer_init(self, blobs, rank, world):
        """blobs: the ranks' comm_peer_export() results in rank order."""
        raw = b"".join(blobs)
        assert len(raw) == world * capi.PEER_BLOB
        buf = (C.c_char * len(raw)).from_buffer_copy(raw)
        check(lib().f2v_comm_peer_init(self._h, buf, rank, world), "f2v_comm_peer_init")


class Algorithms:
    """Mirror of the reference's `algorithms` (sample/algorithms.h:51-137), options 5/6/7."""

    def __init__(self, rowptr, colids, input_name, outputdir, dim, gamma=1.0, bsize=384, device=0):
        self.rowptr = np.ascontiguousarray(rowptr, np.uint64)
        self.colids = np.ascontiguousarray(colids, np.uint32)
        self.rows = len(self.rowptr) - 1
        self.DIM = int(dim)
        self.GAMMA = gamma
        self.filename = input_name
        self.outputdir = outputdir
        self.device = device
        self.epoch_mode = 0
        self.walk_sampler = 0
        self.seed = 1
        self.chunk = 0
        self.gpus = 1
        self.nCoordinates = np.zeros((self.rows, self.DIM), np.float32)

    def _run(self, option, bs, iterations, batch, ns, lr, tag, write):
        a = capi.TrainArgs()
        ci = self.colids if len(self.colids) else np.zeros(1, np.uint32)
        a.n, a.nnz = self.rows, len(self.colids)
        a.rowptr, a.colids = self.rowptr.ctypes.data, ci.ctypes.data
        a.dim, a.option, a.bs = self.DIM, option, bs
        a.iterations, a.batch, a.nsamples, a.lr = iterations, batch, ns, lr
        a.seed, a.device, a.walk_sampler, a.epoch_mode, a.chunk = self.seed, self.device, self.walk_sampler, self.epoch_mode, self.chunk
        sec = C.c_double()
        if self.gpus > 1:
            check(lib().f2v_train_gpus(C.byref(a), self.gpus, _p(self.nCoordinates), C.byref(sec)), "f2v_train_gpus")
        else:
            check(lib().f2v_train(C.byref(a), _p(self.nCoordinates), C.byref(sec)), "f2v_train")
        if write:
            self.writeToFile("%s%dD%dIT%dNS%d" % (tag, batch, self.DIM, iterations, ns))
        return [sec.value]

    def AlgoForce2VecNS(self, ITERATIONS, NUMOFTHREADS, BATCHSIZE, ns, lr, write=True):
        return self._run(TDIST, 0, ITERATIONS, BATCHSIZE, ns, lr, "F2VNS", write)

    def AlgoForce2VecNSBS(self, ITERATIONS, NUMOFTHREADS, BATCHSIZE, ns, lr, write=True):
        return self._run(TDIST, 1, ITERATIONS, BATCHSIZE, ns, lr, "F2VNS", write)

    def AlgoForce2VecNSRW(self, ITERATIONS, NUMOFTHREADS, BATCHSIZE, ns, lr, write=True):
        return self._run(SIGMOID, 0, ITERATIONS, BATCHSIZE, ns, lr, "F2VWNS", write)

    def AlgoForce2VecNSRWBS(self, ITERATIONS, NUMOFTHREADS, BATCHSIZE, ns, lr, write=True):
        return self._run(SIGMOID, 1, ITERATIONS, BATCHSIZE, ns, lr, "F2VWNS", write)

    def AlgoForce2VecNSRWEFF(self, ITERATIONS, NUMOFTHREADS, BATCHSIZE, ns, lr, write=True):
        return self._run(WALK, 0, ITERATIONS, BATCHSIZE, ns, lr, "F2VWNSF", write)

    def writeToFile(self, f):
        lasttok = self.filename.split("/")[-1]
        self.filename = self.outputdir + lasttok + f + ".embd"
        host.write_embd(self.filename, self.nCoordinates)
        return self.filename

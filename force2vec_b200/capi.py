"""ctypes binding of libf2v.so (include/f2v.h + include/f2v_host.h).  Loud on failure."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
TDIST, SIGMOID, WALK = 5, 6, 7
WALKLEN = 5
LUT_SIZE = 2048
PEER_BLOB = 256

ENGINE_SYMBOLS = [
    "f2v_last_error", "f2v_abi_version", "f2v_device_count", "f2v_create", "f2v_destroy",
    "f2v_set_stream", "f2v_sync", "f2v_host_alloc", "f2v_host_free", "f2v_set_embeddings",
    "f2v_get_embeddings", "f2v_get_rows", "f2v_set_lut", "f2v_set_negatives", "f2v_set_negative_offset",
    "f2v_set_walks",
    "f2v_get_walks", "f2v_sample_walks", "f2v_step", "f2v_run_epoch", "f2v_run_epoch_host",
    "f2v_set_epoch_mode", "f2v_set_option", "f2v_launch_count", "f2v_last_epoch_ms", "f2v_comm_unique_id",
    "f2v_comm_init", "f2v_comm_peer_export", "f2v_comm_peer_init", "f2v_trace_ms", "f2v_shard_row",
    "f2v_checksum", "f2v_host_register", "f2v_host_unregister", "f2v_device_memory",
]
HOST_SYMBOLS = [
    "f2v_rng_create", "f2v_rng_destroy", "f2v_rng_next", "f2v_init_embeddings", "f2v_build_lut",
    "f2v_neg_stream_len", "f2v_draw_epoch_negatives", "f2v_draw_walks", "f2v_load_mtx", "f2v_free",
    "f2v_write_embd", "f2v_format_g6", "f2v_write_csr", "f2v_load_csr", "f2v_write_mtx", "f2v_rmat_csr", "f2v_plan_build", "f2v_train", "f2v_train_gpus",
]


class F2VError(RuntimeError):
    pass


class TrainArgs(C.Structure):
    _fields_ = [("n", C.c_uint64), ("nnz", C.c_uint64), ("rowptr", C.c_void_p), ("colids", C.c_void_p),
                ("dim", C.c_uint32), ("option", C.c_int), ("bs", C.c_int), ("iterations", C.c_uint32),
                ("batch", C.c_uint32), ("nsamples", C.c_uint32), ("lr", C.c_float), ("seed", C.c_uint32),
                ("device", C.c_int), ("walk_sampler", C.c_int), ("epoch_mode", C.c_int), ("chunk", C.c_uint32)]


def lib_path():
    # F2V_LIB lets development tools A/B two builds of the library in one GPU session
    return os.environ.get("F2V_LIB") or os.path.join(HERE, "lib", "libf2v.so")


_lib = None


def lib():
    """Load libf2v.so.  Raises F2VError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise F2VError("%s is missing: build it with `make lib` (or __graft_entry__.build()); "
                       "there is no CPU fallback" % path)
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    L.f2v_last_error.restype = C.c_char_p
    L.f2v_create.argtypes = [C.POINTER(vp), i32, u64, u64, vp, vp, u32]
    L.f2v_destroy.argtypes = [vp]
    L.f2v_set_stream.argtypes = [vp, vp]
    L.f2v_sync.argtypes = [vp]
    L.f2v_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.f2v_host_free.argtypes = [vp]
    L.f2v_host_register.argtypes = [vp, u64]
    L.f2v_host_unregister.argtypes = [vp]
    L.f2v_device_memory.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.f2v_checksum.argtypes = [vp, C.POINTER(u64)]
    L.f2v_set_embeddings.argtypes = [vp, vp]
    L.f2v_get_embeddings.argtypes = [vp, vp]
    L.f2v_get_rows.argtypes = [vp, u64, u64, vp]
    L.f2v_set_lut.argtypes = [vp, vp, u32]
    L.f2v_set_negatives.argtypes = [vp, vp, u64]
    L.f2v_set_negative_offset.argtypes = [vp, u64]
    L.f2v_set_walks.argtypes = [vp, vp]
    L.f2v_get_walks.argtypes = [vp, vp]
    L.f2v_sample_walks.argtypes = [vp, u64, u64]
    L.f2v_step.argtypes = [vp, i32, u64, u32, vp, u32, i32, f32, vp]
    L.f2v_run_epoch.argtypes = [vp, i32, u32, u32, i32, f32, u32]
    L.f2v_run_epoch_host.argtypes = [vp, i32, u32, u32, i32, f32, u32, vp, vp, u64, vp, vp]
    L.f2v_set_epoch_mode.argtypes = [vp, i32]
    L.f2v_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.f2v_launch_count.argtypes = [vp]
    L.f2v_launch_count.restype = u64
    L.f2v_last_epoch_ms.argtypes = [vp, C.POINTER(f32)]
    L.f2v_comm_unique_id.argtypes = [vp]
    L.f2v_comm_init.argtypes = [vp, vp, i32, i32]
    L.f2v_trace_ms.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.f2v_shard_row.argtypes = [u32, u32, u32]
    L.f2v_shard_row.restype = u32
    L.f2v_comm_peer_export.argtypes = [vp, vp]
    L.f2v_comm_peer_init.argtypes = [vp, vp, i32, i32]
    # host side
    L.f2v_rng_create.argtypes = [u32]
    L.f2v_rng_create.restype = vp
    L.f2v_rng_destroy.argtypes = [vp]
    L.f2v_rng_next.argtypes = [vp]
    L.f2v_rng_next.restype = C.c_int32
    L.f2v_init_embeddings.argtypes = [vp, i32, u64, u32, vp]
    L.f2v_build_lut.argtypes = [vp]
    L.f2v_neg_stream_len.argtypes = [i32, u64, u32, u32, i32]
    L.f2v_neg_stream_len.restype = u64
    L.f2v_draw_epoch_negatives.argtypes = [vp, i32, u64, u32, u32, i32, vp]
    L.f2v_draw_walks.argtypes = [vp, u64, u64, vp, vp, vp]
    L.f2v_load_mtx.argtypes = [C.c_char_p, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(vp)]
    L.f2v_free.argtypes = [vp]
    L.f2v_write_embd.argtypes = [C.c_char_p, vp, u64, u32]
    L.f2v_format_g6.argtypes = [f32, C.c_char_p]
    L.f2v_write_csr.argtypes = [C.c_char_p, u64, u64, vp, vp]
    L.f2v_load_csr.argtypes = [C.c_char_p, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(vp)]
    L.f2v_write_mtx.argtypes = [C.c_char_p, u64, vp, vp]
    L.f2v_rmat_csr.argtypes = [i32, i32, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(vp)]
    L.f2v_plan_build.argtypes = [vp, u64, u64, u32, u32, u32, i32, i32, i32, i32, C.POINTER(u64), C.POINTER(vp),
                                 C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.f2v_train.argtypes = [C.POINTER(TrainArgs), vp, C.POINTER(C.c_double)]
    L.f2v_train_gpus.argtypes = [C.POINTER(TrainArgs), i32, vp, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().f2v_last_error()
        raise F2VError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

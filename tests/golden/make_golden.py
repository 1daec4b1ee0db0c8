"""tests/golden/make_golden.py -- regenerates the golden fixtures in this directory.

Run in the build container (needs /root/reference and `make -C oracle`):
    python tests/golden/make_golden.py
It drives the UNMODIFIED reference (oracle/_ref/libf2vref.so = /root/reference/sample/
algorithms.cpp + oracle/ref_shim.cpp, and the oracle/_ref/Force2Vec CLI) and stores
  ref_outputs.npz          nCoordinates after `it` epochs for a grid of (graph, option, bs, dim,
                           batch, it); cora arrays keep every 4th row plus whole-array
                           checksums (sum, Frobenius norm) to stay small
  ref_lut.npy              the reference's sm_table (init_SM_TABLE, algorithms.cpp:757-764)
  rand_srand1.npy          first 4096 outputs of libc rand() after srand(1)
  shipped_cora_F2VNS384D128IT1200NS5.npz   the reference's own shipped golden embedding
                           (datasets/output/cora.mtxF2VNS384D128IT1200NS5.embd), as float32
  karate_opt{5,6,7}.embd   text written by the reference CLI (pins the .embd format)
cora.mtx / karate.mtx are copies of the reference's input DATA files (datasets/input/).
"""
import ctypes
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def key(graph, opt, bs, dim, B, it):
    return "%s_opt%d_bs%d_d%d_B%d_it%d" % (graph, opt, bs, dim, B, it)


def main():
    out = {}
    rp, ci = O.load_mtx(os.path.join(HERE, "karate.mtx"))
    for opt in (5, 6, 7):
        for bs in ((0, 1) if opt != 7 else (0,)):
            for dim in (128, 64, 20):
                for it in (1, 3):
                    X, _ = O.ref_run(opt, bs, rp, ci, dim, it, 8, 5, 0.02, threads=2)
                    out[key("karate", opt, bs, dim, 8, it)] = X
    rp, ci = O.load_mtx(os.path.join(HERE, "cora.mtx"))
    grid = [(5, 0, 128, 256, 1), (5, 0, 128, 256, 5), (5, 0, 128, 256, 50), (5, 1, 128, 256, 2),
            (6, 0, 128, 256, 1), (6, 0, 128, 256, 5), (6, 0, 128, 256, 50), (6, 1, 128, 256, 2),
            (7, 0, 64, 256, 1), (7, 0, 64, 256, 5), (7, 0, 64, 256, 50), (7, 0, 128, 384, 2)]
    for (opt, bs, dim, B, it) in grid:
        X, _ = O.ref_run(opt, bs, rp, ci, dim, it, B, 5, 0.02, threads=4)
        k = key("cora", opt, bs, dim, B, it)
        out[k] = X[::4].copy()
        out[k + "_sum"] = np.float64(X.astype(np.float64).sum())
        out[k + "_fro"] = np.float64(np.linalg.norm(X.astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    np.save(os.path.join(HERE, "ref_lut.npy"), O.ref_lut())
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    np.save(os.path.join(HERE, "rand_srand1.npy"), np.array([libc.rand() for _ in range(4096)], np.int32))
    G = O.read_embd(os.path.join(REF, "datasets/output/cora.mtxF2VNS384D128IT1200NS5.embd"))
    np.savez_compressed(os.path.join(HERE, "shipped_cora_F2VNS384D128IT1200NS5.npz"), X=G)
    with tempfile.TemporaryDirectory() as td:
        for opt, tag in ((5, "F2VNS"), (6, "F2VWNS"), (7, "F2VWNSF")):
            subprocess.check_call([O.ref_cli(), "-input", os.path.join(HERE, "karate.mtx"), "-output", td + "/",
                                   "-iter", "3", "-batch", "8", "-dim", "16", "-nsamples", "5", "-option", str(opt),
                                   "-threads", "2"], stdout=subprocess.DEVNULL, cwd=td)
            src = os.path.join(td, "karate.mtx%s8D16IT3NS5.embd" % tag)
            with open(src) as f, open(os.path.join(HERE, "karate_opt%d.embd" % opt), "w") as g:
                g.write(f.read())
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()

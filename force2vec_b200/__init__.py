"""force2vec_b200 -- B200-native (sm_100a) Force2Vec force-step engine.

Python is only the test/bench harness language here: the product is libf2v.so (hand-written
CUDA kernels + C ABI, include/f2v.h, include/f2v_host.h) and the drop-in bin/Force2Vec CLI.
This package binds the C ABI with ctypes and mirrors the reference's `algorithms` interface
(/root/reference/sample/algorithms.h:51-137) for options 5/6/7.  There is no CPU fallback:
every compute call goes to the CUDA library and raises if it is missing or no GPU is usable.
"""
from .capi import F2VError, lib, lib_path, TDIST, SIGMOID, WALK, WALKLEN  # noqa: F401
from .engine import Engine, Algorithms  # noqa: F401
from . import host  # noqa: F401

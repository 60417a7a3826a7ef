"""TEST INFRASTRUCTURE: CPU emulation of the data-parallel scoring passes (alphadia_b200/csrc/adb_score_dp.cuh).

The pass bodies are ``__host__ __device__`` functions without warp-level cooperation; ``hostsim.cu`` calls them thread by thread.
Only ``tests/`` may import this package — it lets the CUDA source be compared with the oracle on a machine without a GPU.
"""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

from alphadia_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO_PATH = os.path.join(HERE, "libadb_hostsim.so")
SRC = os.path.join(HERE, "hostsim.cu")
DEPS = [SRC, os.path.join(ROOT, "alphadia_b200", "csrc", "adb_score_dp.cuh"), os.path.join(ROOT, "alphadia_b200", "csrc", "adb_common.cuh"),
        os.path.join(ROOT, "include", "alphadia_b200.h")]
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= max(os.path.getmtime(d) for d in DEPS):
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false", "-Xcompiler",
                           "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-cudart", "static", "-o", SO_PATH, SRC])
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO_PATH)
    return _lib


def score_candidates(raw, lib_arrays, cfg_struct, cand_in, *, batch=64, order=None, ks=None, n_buckets=0):
    """Same signature and result dict as oracle.score_candidates / _lib.score_candidates."""
    L = lib()
    rd, k1 = _abi.make_rawfile3d_desc(raw)
    ld, k2 = _abi.make_library_desc(lib_arrays)
    n = int(cand_in.n)
    max_frag = int(np.max(lib_arrays["frag_stop_idx"].astype(np.int64) - lib_arrays["frag_start_idx"].astype(np.int64))) if ld.n_precursors else 1
    K = int(cfg_struct.top_k_fragments)
    KS = int(ks) if ks is not None else max(1, min(K, max_frag))
    od, arrs = _abi.alloc_scores_out(n, K)
    status = C.c_uint32(0)
    order_p = None
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.int32)
        order_p = order.ctypes.data_as(C.POINTER(C.c_int32))
    rc = L.adb_hostsim_score(C.byref(rd), C.byref(ld), C.byref(cfg_struct), C.byref(cand_in), C.byref(od), C.c_int32(K), C.c_int32(KS),
                             C.c_int64(batch), order_p, C.c_int32(n_buckets), C.byref(status))
    if rc != 0:
        raise RuntimeError("hostsim failed")
    arrs["status"] = status.value
    return arrs

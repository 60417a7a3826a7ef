"""CPU oracle (test infrastructure): C restatement of the reference hot path + its ctypes wrapper.

ONLY ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product package ``alphadia_b200`` never does.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from alphadia_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libadb_oracle.so")
SRC = os.path.join(HERE, "adb_oracle.c")

_lib = None


def build(force: bool = False) -> str:
    """gcc-compile oracle/adb_oracle.c -> oracle/libadb_oracle.so (no fast-math, no FMA contraction)."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "alphadia_b200.h")
    if not force and os.path.exists(SO_PATH):
        if os.path.getmtime(SO_PATH) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
            return SO_PATH
    cmd = ["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wno-maybe-uninitialized",
           "-o", SO_PATH, SRC, "-lm"]
    subprocess.check_call(cmd)
    return SO_PATH


class SelectionTap(C.Structure):
    _fields_ = [("precursor_row", C.c_int64), ("dense_precursors", _abi.c_f32p), ("dense_fragments", _abi.c_f32p),
                ("score", _abi.c_f64p), ("capacity", C.c_int64), ("C", C.c_int64), ("F", C.c_int64), ("I", C.c_int64)]


class ScoringTap(C.Structure):
    _fields_ = [("candidate", C.c_int64), ("dense_fragments", _abi.c_f32p), ("dense_precursors", _abi.c_f32p),
                ("template_", _abi.c_f32p), ("capacity", C.c_int64), ("F", C.c_int64), ("nobs", C.c_int64),
                ("C", C.c_int64), ("I", C.c_int64)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < os.path.getmtime(SRC):
            build()
        _lib = C.CDLL(SO_PATH)
        _lib.adbo_num_threads.restype = C.c_int
    return _lib


def select_candidates(raw, lib_arrays, cfg_struct, kernel, n_threads=0, tap_row=None, rows=None):
    """Run oracle selection.  Returns dict of CandidateContainer arrays (and tap dict if requested)."""
    L = lib()
    rd, k1 = _abi.make_rawfile3d_desc(raw)
    ld, k2 = _abi.make_library_desc(lib_arrays)
    n_rows = int(ld.n_precursors * cfg_struct.candidate_count)
    od, arrs = _abi.alloc_candidates_out(n_rows)
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    tap = None
    tap_bufs = None
    if tap_row is not None:
        cap = 1 << 16
        tap_bufs = dict(dp=np.zeros(cap, np.float32), df=np.zeros(cap, np.float32), score=np.zeros(cap, np.float64))
        tap = SelectionTap(int(tap_row), _abi.ptr(tap_bufs["dp"]), _abi.ptr(tap_bufs["df"]), _abi.ptr(tap_bufs["score"]), cap, 0, 0, 0)
    r0, r1 = (0, int(ld.n_precursors)) if rows is None else rows
    rc = L.adbo_select_candidates(C.byref(rd), C.byref(ld), C.byref(cfg_struct), _abi.ptr(kernel),
                                  C.c_int32(kernel.shape[0]), C.c_int32(kernel.shape[1]), C.byref(od),
                                  C.c_int64(r0), C.c_int64(r1), C.c_int32(n_threads),
                                  C.byref(tap) if tap is not None else None)
    if rc != 0:
        raise RuntimeError(f"oracle selection failed rc={rc}")
    if tap is not None:
        Cn, F, I = tap.C, tap.F, tap.I
        return arrs, dict(C=Cn, F=F, I=I, dense_precursors=tap_bufs["dp"][: I * Cn].reshape(I, Cn).copy(),
                          dense_fragments=tap_bufs["df"][: F * Cn].reshape(F, Cn).copy(), score=tap_bufs["score"][:Cn].copy())
    return arrs


def score_candidates(raw, lib_arrays, cfg_struct, cand_in_struct, n_threads=0, tap_candidate=None):
    L = lib()
    rd, k1 = _abi.make_rawfile3d_desc(raw)
    ld, k2 = _abi.make_library_desc(lib_arrays)
    od, arrs = _abi.alloc_scores_out(int(cand_in_struct.n), int(cfg_struct.top_k_fragments))
    tap = None
    bufs = None
    if tap_candidate is not None:
        cap = 1 << 18
        bufs = dict(df=np.zeros(cap, np.float32), dp=np.zeros(cap, np.float32), t=np.zeros(cap, np.float32))
        tap = ScoringTap(int(tap_candidate), _abi.ptr(bufs["df"]), _abi.ptr(bufs["dp"]), _abi.ptr(bufs["t"]), cap, 0, 0, 0, 0)
    rc = L.adbo_score_candidates(C.byref(rd), C.byref(ld), C.byref(cfg_struct), C.byref(cand_in_struct), C.byref(od),
                                 C.c_int32(n_threads), C.byref(tap) if tap is not None else None)
    if rc != 0:
        raise RuntimeError(f"oracle scoring failed rc={rc}")
    if tap is not None:
        F, nobs, Cn, I = tap.F, tap.nobs, tap.C, tap.I
        return arrs, dict(F=F, nobs=nobs, C=Cn, I=I,
                          dense_fragments=bufs["df"][: 2 * F * nobs * Cn].reshape(2, F, nobs, Cn).copy(),
                          dense_precursors=bufs["dp"][: 2 * I * Cn].reshape(2, I, Cn).copy(),
                          template=bufs["t"][: nobs * Cn].reshape(nobs, Cn).copy())
    return arrs


def fragment_competition(window_start, window_stop, rt, frag_start, frag_stop, fragment_mz, rt_tol, ppm_tol,
                         valid=None, n_threads=0):
    L = lib()
    ws = _abi.as_c(window_start, np.int64)
    we = _abi.as_c(window_stop, np.int64)
    fs = _abi.as_c(frag_start, np.int64)
    fe = _abi.as_c(frag_stop, np.int64)
    rt_f64 = np.asarray(rt).dtype == np.float64
    mz_f64 = np.asarray(fragment_mz).dtype == np.float64
    is_f64 = int(rt_f64) | (int(mz_f64) << 1)
    rt_c = _abi.as_c(rt, np.float64 if rt_f64 else np.float32)
    mz_c = _abi.as_c(fragment_mz, np.float64 if mz_f64 else np.float32)
    v = np.ones(len(rt_c), np.uint8) if valid is None else _abi.as_c(valid, np.uint8).copy()
    rc = L.adbo_fragment_competition(C.c_int64(len(ws)), _abi.ptr(ws), _abi.ptr(we), C.c_int64(len(rt_c)),
                                     rt_c.ctypes.data_as(C.c_void_p), _abi.ptr(fs), _abi.ptr(fe), C.c_int64(len(mz_c)),
                                     mz_c.ctypes.data_as(C.c_void_p), C.c_int32(is_f64), C.c_double(rt_tol),
                                     C.c_double(ppm_tol), _abi.ptr(v), C.c_int32(n_threads))
    if rc != 0:
        raise RuntimeError("oracle fragcomp failed")
    return v.astype(bool)


def select_candidates_4d(raw4d, lib_arrays, cfg_struct, kernel, n_threads=0):
    """Oracle selection on a timsTOF-layout raw file."""
    L = lib()
    rd, k1 = _abi.make_rawfile4d_desc(raw4d)
    ld, k2 = _abi.make_library_desc(lib_arrays)
    od, arrs = _abi.alloc_candidates_out(int(ld.n_precursors * cfg_struct.candidate_count))
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    rc = L.adbo_select_candidates_4d(C.byref(rd), C.byref(ld), C.byref(cfg_struct), _abi.ptr(kernel),
                                     C.c_int32(kernel.shape[0]), C.c_int32(kernel.shape[1]), C.byref(od), C.c_int32(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle 4-D selection failed rc={rc}")
    return arrs


def score_candidates_4d(raw4d, lib_arrays, cfg_struct, cand_in_struct, n_threads=0):
    L = lib()
    rd, k1 = _abi.make_rawfile4d_desc(raw4d)
    ld, k2 = _abi.make_library_desc(lib_arrays)
    od, arrs = _abi.alloc_scores_out(int(cand_in_struct.n), int(cfg_struct.top_k_fragments))
    rc = L.adbo_score_candidates_4d(C.byref(rd), C.byref(ld), C.byref(cfg_struct), C.byref(cand_in_struct), C.byref(od),
                                    C.c_int32(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle 4-D scoring failed rc={rc}")
    return arrs


def transpose_csr(tof_indices, push_indptr, n_tof_indices, values):
    """C restatement of ``_transpose`` (alphadia/raw_data/bruker.py:202-274)."""
    L = lib()
    tof = _abi.as_c(tof_indices, np.uint32)
    ptr_ = _abi.as_c(push_indptr, np.int64)
    vals = _abi.as_c(values, np.uint16)
    n, n_push = len(tof), len(ptr_) - 1
    push_out = np.zeros(n, np.uint32)
    indptr_out = np.zeros(int(n_tof_indices) + 1, np.int64)
    vals_out = np.zeros(n, np.uint16)
    rc = L.adbo_transpose_csr(C.c_int64(n), C.c_int64(n_push), C.c_int64(int(n_tof_indices)), _abi.ptr(tof), _abi.ptr(ptr_),
                              _abi.ptr(vals), _abi.ptr(push_out), _abi.ptr(indptr_out), _abi.ptr(vals_out))
    if rc != 0:
        raise RuntimeError("adbo_transpose_csr failed (tof index out of range)")
    return push_out, indptr_out, vals_out


def q_values(score, decoy, extra_key):
    """C restatement of ``get_q_values`` (alphadia/fdr/fdr.py:226-297): ``(order, qval)``."""
    L = lib()
    sc, dc, ek = _abi.as_c(score, np.float64), _abi.as_c(decoy, np.uint8), _abi.as_c(extra_key, np.uint64)
    n = len(sc)
    order, qval = np.zeros(n, np.int64), np.zeros(n, np.float64)
    L.adbo_q_values.restype = None
    L.adbo_q_values(C.c_int64(n), _abi.ptr(sc), _abi.ptr(dc), _abi.ptr(ek), _abi.ptr(order), _abi.ptr(qval))
    return order, qval


def keep_best(score, group_key):
    """C restatement of ``keep_best`` (alphadia/fdr/fdr.py:195-224): the u8 mask of surviving rows."""
    L = lib()
    sc, gk = _abi.as_c(score, np.float64), _abi.as_c(group_key, np.uint64)
    keep = np.zeros(len(sc), np.uint8)
    L.adbo_keep_best.restype = None
    L.adbo_keep_best(C.c_int64(len(sc)), _abi.ptr(sc), _abi.ptr(gk), _abi.ptr(keep))
    return keep


def fdr_to_q_values(fdr_values):
    L = lib()
    f = _abi.as_c(fdr_values, np.float64)
    q = np.zeros(len(f), np.float64)
    L.adbo_fdr_to_q_values.restype = None
    L.adbo_fdr_to_q_values(_abi.ptr(f), C.c_int64(len(f)), _abi.ptr(q))
    return q


def classifier_predict_proba(network_state: dict, x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """TEST ORACLE: FeedForwardNN.forward in eval mode (alphadia/fdr/classifiers.py:473-532) restated in numpy float32:
    BatchNorm1d with running statistics (module 0), [Linear, ReLU, Dropout = identity] per hidden layer (modules 1, 4, 7, ...),
    Linear, softmax over axis 1.  Pinned against the live reference's predict_proba in tests/golden/classifier_small.npz."""
    f32 = np.float32
    g = {k: np.asarray(v) for k, v in network_state.items()}
    a = np.asarray(x, dtype=f32)
    a = (a - g["fc_layers.0.running_mean"].astype(f32)) / np.sqrt(g["fc_layers.0.running_var"].astype(f32) + f32(eps))
    a = a * g["fc_layers.0.weight"].astype(f32) + g["fc_layers.0.bias"].astype(f32)
    idx = sorted(int(k.split(".")[1]) for k in g if k.endswith(".weight") and not k.startswith("fc_layers.0."))
    for n, i in enumerate(idx):
        a = a @ g[f"fc_layers.{i}.weight"].astype(f32).T + g[f"fc_layers.{i}.bias"].astype(f32)
        if n < len(idx) - 1:
            a = np.maximum(a, f32(0))
    a = a - a.max(axis=1, keepdims=True)
    e = np.exp(a)
    return (e / e.sum(axis=1, keepdims=True)).astype(f32)

"""ctypes loader for libalphadia_b200.so + thin device-handle wrappers.

There is NO CPU fallback: if the CUDA extension is missing, cannot be loaded, or no CUDA device is
visible, the operators raise immediately.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from alphadia_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ADB_LIB_PATH") or os.path.join(HERE, "libalphadia_b200.so")  # ADB_LIB_PATH: tuning builds of the same library

# every symbol include/alphadia_b200.h declares
EXPORTED_SYMBOLS = [
    "adb_last_error", "adb_version", "adb_device_count",
    "adb_rawfile3d_create", "adb_rawfile4d_create", "adb_rawfile_destroy", "adb_rawfile_device_bytes", "adb_rawfile_stream",
    "adb_library_create", "adb_library_destroy",
    "adb_select_candidates", "adb_score_candidates", "adb_score_candidates_ragged", "adb_fragment_competition", "adb_transpose_csr",
    "adb_q_values", "adb_keep_best", "adb_classifier_predict_proba",
    "adb_select_candidates_resident", "adb_score_candidates_resident", "adb_select_score_candidates_ragged",
    "adb_fetch_candidates", "adb_fetch_candidate_table", "adb_fetch_scores", "adb_resident_score_table",
    "adb_last_timing", "adb_kernel_launches", "adb_last_main_kernel_ms",
]

_lib = None


class ExtensionError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA extension; raises ExtensionError loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ExtensionError(
            f"{SO_PATH} is missing — build it with `python -m alphadia_b200.build` "
            "(alphadia_b200 has no CPU fallback)"
        )
    lib = C.CDLL(SO_PATH)
    lib.adb_last_error.restype = C.c_char_p
    lib.adb_version.restype = C.c_char_p
    lib.adb_device_count.restype = C.c_int
    lib.adb_rawfile_device_bytes.restype = C.c_int64
    lib.adb_rawfile_device_bytes.argtypes = [C.c_void_p]
    lib.adb_rawfile_stream.restype = C.c_void_p
    lib.adb_rawfile_stream.argtypes = [C.c_void_p]
    lib.adb_kernel_launches.restype = C.c_int64
    lib.adb_kernel_launches.argtypes = [C.c_void_p]
    lib.adb_last_main_kernel_ms.restype = C.c_float
    lib.adb_last_main_kernel_ms.argtypes = [C.c_void_p]
    lib.adb_rawfile_destroy.argtypes = [C.c_void_p]
    lib.adb_rawfile_destroy.restype = None
    lib.adb_library_destroy.argtypes = [C.c_void_p]
    lib.adb_library_destroy.restype = None
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().adb_last_error()
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def require_device() -> int:
    lib = load()
    n = lib.adb_device_count()
    if n < 1:
        raise ExtensionError("alphadia_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    return n


def current_device() -> int:
    """One process per GPU: LOCAL_RANK selects the device (torchrun), default 0."""
    return int(os.environ.get("ADB_DEVICE", os.environ.get("LOCAL_RANK", "0")))


class DeviceRawFile:
    """Raw file resident in HBM (replaces ``dia_data.to_jitclass()``)."""

    def __init__(self, raw, device: int | None = None):
        require_device()
        self._lib = load()
        self.device = current_device() if device is None else device
        h = C.c_void_p()
        self.is_4d = bool(getattr(raw, "is_4d", False)) or hasattr(raw, "tof_indptr")
        if self.is_4d:
            desc, keep = _abi.make_rawfile4d_desc(raw)
            check(self._lib.adb_rawfile4d_create(C.byref(desc), C.c_int(self.device), C.byref(h)), "adb_rawfile4d_create")
            self.cycle_len = int(desc.frames_per_cycle)
            self.n_spectra = int(desc.n_frames)
        else:
            desc, keep = _abi.make_rawfile3d_desc(raw)
            check(self._lib.adb_rawfile3d_create(C.byref(desc), C.c_int(self.device), C.byref(h)), "adb_rawfile3d_create")
            self.cycle_len = int(desc.cycle_len)
            self.n_spectra = int(desc.n_spectra)
        self.handle = h

    @property
    def device_bytes(self) -> int:
        return int(self._lib.adb_rawfile_device_bytes(self.handle))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.adb_kernel_launches(self.handle))

    def last_timing(self) -> dict:
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._lib.adb_last_timing(self.handle, C.byref(a), C.byref(b), C.byref(c))
        return dict(h2d_ms=a.value, kernel_ms=b.value, d2h_ms=c.value,
                    main_kernel_ms=float(self._lib.adb_last_main_kernel_ms(self.handle)))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.adb_rawfile_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceLibrary:
    """Flat spectral library resident in HBM."""

    def __init__(self, lib_arrays: dict, device: int | None = None):
        require_device()
        self._lib = load()
        self.device = current_device() if device is None else device
        desc, keep = _abi.make_library_desc(lib_arrays)
        h = C.c_void_p()
        check(self._lib.adb_library_create(C.byref(desc), C.c_int(self.device), C.byref(h)), "adb_library_create")
        self.handle = h
        self.n_precursors = int(desc.n_precursors)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.adb_library_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_RAW_CACHE: dict = {}


def device_rawfile_for(dia_data, adapted) -> DeviceRawFile:
    """One upload per raw-file object (selection and scoring of the same file share it)."""
    key = id(dia_data)
    ent = _RAW_CACHE.get(key)
    if ent is not None and ent[0] is dia_data and ent[1].handle:
        return ent[1]
    for k in list(_RAW_CACHE):
        _RAW_CACHE.pop(k)[1].close()
    dev = DeviceRawFile(adapted)
    _RAW_CACHE[key] = (dia_data, dev)
    return dev


def select_candidates(dev_raw: DeviceRawFile, dev_lib: DeviceLibrary, cfg_struct, kernel: np.ndarray) -> dict:
    lib = load()
    n_rows = int(dev_lib.n_precursors * cfg_struct.candidate_count)
    od, arrs = _abi.alloc_candidates_out(n_rows)
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    check(lib.adb_select_candidates(dev_raw.handle, dev_lib.handle, C.byref(cfg_struct), _abi.ptr(kernel),
                                    C.c_int32(kernel.shape[0]), C.c_int32(kernel.shape[1]), C.byref(od)),
          "adb_select_candidates")
    return arrs


def score_candidates(dev_raw: DeviceRawFile, dev_lib: DeviceLibrary, cfg_struct, cand_in_struct) -> dict:
    lib = load()
    od, arrs = _abi.alloc_scores_out(int(cand_in_struct.n), int(cfg_struct.top_k_fragments))
    check(lib.adb_score_candidates(dev_raw.handle, dev_lib.handle, C.byref(cfg_struct), C.byref(cand_in_struct), C.byref(od)),
          "adb_score_candidates")
    return arrs


def score_candidates_ragged(dev_raw: DeviceRawFile, dev_lib: DeviceLibrary, cfg_struct, cand_in_struct, bufs: dict | None = None,
                            max_fragments: int | None = None) -> dict:
    """Scoring with the ragged result of ``adb_score_candidates_ragged``: the feature rows of the valid candidates and their
    kept fragment slots, flattened in candidate order.  Returns views of length ``n_rows`` / ``n_fragments`` into ``bufs``
    (allocated on demand: ``n`` rows, ``n * min(top_k_fragments, max_fragments)`` fragment entries)."""
    lib = load()
    n = int(cand_in_struct.n)
    if bufs is None:
        per = int(cfg_struct.top_k_fragments) if max_fragments is None else min(int(cfg_struct.top_k_fragments), int(max_fragments))
        rd, bufs = _abi.alloc_scores_ragged(n, n * max(per, 1))
    else:
        rd = _abi.scores_ragged_struct(bufs)
    check(lib.adb_score_candidates_ragged(dev_raw.handle, dev_lib.handle, C.byref(cfg_struct), C.byref(cand_in_struct), C.byref(rd)),
          "adb_score_candidates_ragged")
    nr, nf = int(rd.n_rows), int(rd.n_fragments)
    out = dict(n_rows=nr, n_fragments=nf, row_index=bufs["row_index"][:nr], features=bufs["features"][:nr],
               frag_offset=bufs["frag_offset"][:nr + 1])
    for k in _abi.FRAG_F32 + _abi.FRAG_U8:
        out[k] = bufs[k][:nf]
    return out


def select_score_candidates_ragged(dev_raw, dev_lib, sel_struct, kernel, score_struct, table: dict, bufs: dict,
                                   score_cutoff: float = float("-inf")) -> dict:
    """Selection + score cutoff + scoring in one call (``adb_select_score_candidates_ragged``): fills the candidate table
    ``table`` (``_abi.alloc_candidate_table(capacity)``) and the ragged result buffers ``bufs``; returns the views of
    ``score_candidates_ragged`` plus ``n_candidates``."""
    lib = load()
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    cap = int(table["lib_row"].shape[0])
    t = _abi.candidate_table_struct(table, cap)
    rd = _abi.scores_ragged_struct(bufs)
    check(lib.adb_select_score_candidates_ragged(dev_raw.handle, dev_lib.handle, C.byref(sel_struct), _abi.ptr(kernel),
                                                 C.c_int32(kernel.shape[0]), C.c_int32(kernel.shape[1]), C.c_float(score_cutoff),
                                                 C.byref(score_struct), C.byref(t), C.byref(rd)),
          "adb_select_score_candidates_ragged")
    nr, nf = int(rd.n_rows), int(rd.n_fragments)
    out = dict(n_candidates=int(t.n), n_rows=nr, n_fragments=nf, row_index=bufs["row_index"][:nr], features=bufs["features"][:nr],
               frag_offset=bufs["frag_offset"][:nr + 1])
    for k in _abi.FRAG_F32 + _abi.FRAG_U8:
        out[k] = bufs[k][:nf]
    return out


def select_candidates_resident(dev_raw, dev_lib, cfg_struct, kernel) -> int:
    lib = load()
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    n = C.c_int64(0)
    check(lib.adb_select_candidates_resident(dev_raw.handle, dev_lib.handle, C.byref(cfg_struct), _abi.ptr(kernel),
                                             C.c_int32(kernel.shape[0]), C.c_int32(kernel.shape[1]), C.byref(n)),
          "adb_select_candidates_resident")
    return int(n.value)


def score_candidates_resident(dev_raw, dev_lib, cfg_struct) -> None:
    check(load().adb_score_candidates_resident(dev_raw.handle, dev_lib.handle, C.byref(cfg_struct)),
          "adb_score_candidates_resident")


def fetch_candidates(dev_raw, n_rows: int) -> dict:
    od, arrs = _abi.alloc_candidates_out(n_rows)
    check(load().adb_fetch_candidates(dev_raw.handle, C.byref(od)), "adb_fetch_candidates")
    return arrs


def fetch_candidate_table(dev_raw, n: int, arrs: dict | None = None) -> dict:
    """Rows with score > 0 of the last resident selection (container order), int64 index columns."""
    if arrs is None:
        arrs = _abi.alloc_candidate_table(n)
    t = _abi.candidate_table_struct(arrs, n)
    check(load().adb_fetch_candidate_table(dev_raw.handle, C.byref(t)), "adb_fetch_candidate_table")
    return arrs


def fetch_scores(dev_raw, n: int, top_k: int) -> dict:
    od, arrs = _abi.alloc_scores_out(n, top_k)
    arrs["lib_row"] = np.zeros(n, np.int64)
    arrs["rank"] = np.zeros(n, np.uint8)
    check(load().adb_fetch_scores(dev_raw.handle, C.byref(od), _abi.ptr(arrs["lib_row"]), _abi.ptr(arrs["rank"])),
          "adb_fetch_scores")
    return arrs


def fragment_competition(window_start, window_stop, rt, frag_start, frag_stop, fragment_mz, rt_tol, ppm_tol,
                         device: int | None = None) -> np.ndarray:
    require_device()
    lib = load()
    ws = _abi.as_c(window_start, np.int64)
    we = _abi.as_c(window_stop, np.int64)
    fs = _abi.as_c(frag_start, np.int64)
    fe = _abi.as_c(frag_stop, np.int64)
    rt_f64 = np.asarray(rt).dtype == np.float64
    mz_f64 = np.asarray(fragment_mz).dtype == np.float64
    is_f64 = int(rt_f64) | (int(mz_f64) << 1)
    rt_c = _abi.as_c(rt, np.float64 if rt_f64 else np.float32)
    mz_c = _abi.as_c(fragment_mz, np.float64 if mz_f64 else np.float32)
    valid = np.ones(len(rt_c), np.uint8)
    dev = current_device() if device is None else device
    check(lib.adb_fragment_competition(C.c_int(dev), C.c_int64(len(ws)), _abi.ptr(ws), _abi.ptr(we), C.c_int64(len(rt_c)),
                                       rt_c.ctypes.data_as(C.c_void_p), _abi.ptr(fs), _abi.ptr(fe), C.c_int64(len(mz_c)),
                                       mz_c.ctypes.data_as(C.c_void_p), C.c_int32(is_f64), C.c_double(rt_tol),
                                       C.c_double(ppm_tol), _abi.ptr(valid)),
          "adb_fragment_competition")
    return valid.astype(bool)


def transpose_csr(tof_indices, push_indptr, n_tof_indices: int, values, device: int | None = None):
    """Device version of ``_transpose(tof_indices, push_indptr, n_tof_indices, values)`` (alphadia/raw_data/bruker.py:202-274):
    returns ``(push_indices u32, tof_indptr i64, new_values)`` of the tof-major CSR."""
    require_device()
    lib = load()
    tof = _abi.as_c(tof_indices, np.uint32)
    ptr_ = _abi.as_c(push_indptr, np.int64)
    vals = _abi.as_c(values, np.uint16)
    if len(ptr_) < 1 or len(tof) != len(vals):
        raise ValueError("tof_indices and values must have the same length and push_indptr at least one entry")
    n, n_push = len(tof), len(ptr_) - 1
    push_out = np.zeros(n, np.uint32)
    indptr_out = np.zeros(int(n_tof_indices) + 1, np.int64)
    vals_out = np.zeros(n, np.uint16)
    dev = current_device() if device is None else device
    check(lib.adb_transpose_csr(C.c_int(dev), C.c_int64(n), C.c_int64(n_push), C.c_int64(int(n_tof_indices)), _abi.ptr(tof),
                                _abi.ptr(ptr_), _abi.ptr(vals), _abi.ptr(push_out), _abi.ptr(indptr_out), _abi.ptr(vals_out)),
          "adb_transpose_csr")
    return push_out, indptr_out, vals_out


def q_values(score, decoy, extra_key, device: int | None = None):
    """``adb_q_values``: ``(order i64, qval f64)`` of get_q_values (alphadia/fdr/fdr.py:226-297) — ``order`` is the row
    permutation of ``sort_values([score, decoy, *extra])``, ``qval`` is in that sorted order."""
    require_device()
    sc, dc, ek = _abi.as_c(score, np.float64), _abi.as_c(decoy, np.uint8), _abi.as_c(extra_key, np.uint64)
    if not len(sc) == len(dc) == len(ek):
        raise ValueError("score, decoy and extra_key must have the same length")
    order, qval = np.zeros(len(sc), np.int64), np.zeros(len(sc), np.float64)
    dev = current_device() if device is None else device
    check(load().adb_q_values(C.c_int(dev), C.c_int64(len(sc)), _abi.ptr(sc), _abi.ptr(dc), _abi.ptr(ek), _abi.ptr(order),
                              _abi.ptr(qval)), "adb_q_values")
    return order, qval


def keep_best(score, group_key, device: int | None = None):
    """``adb_keep_best``: u8 mask of the best (lowest score, earliest) row of every group (alphadia/fdr/fdr.py:195-224)."""
    require_device()
    sc, gk = _abi.as_c(score, np.float64), _abi.as_c(group_key, np.uint64)
    if len(sc) != len(gk):
        raise ValueError("score and group_key must have the same length")
    keep = np.zeros(len(sc), np.uint8)
    dev = current_device() if device is None else device
    check(load().adb_keep_best(C.c_int(dev), C.c_int64(len(sc)), _abi.ptr(sc), _abi.ptr(gk), _abi.ptr(keep)), "adb_keep_best")
    return keep

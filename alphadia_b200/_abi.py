"""ctypes mirror of ``include/alphadia_b200.h`` (struct layouts + marshalling helpers).

Host buffers are plain numpy arrays owned by the caller; the descriptors only carry pointers.
Each ``make_*`` helper returns ``(struct, keepalive)`` — keep ``keepalive`` referenced for the
duration of the C call.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

NUM_FEATURES = 46
MAX_FRAGMENTS = 32
MAX_ISOTOPES = 8

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)
c_u64p = C.POINTER(C.c_uint64)


class RawFile3DDesc(C.Structure):
    _fields_ = [
        ("cycle", c_f64p), ("cycle_len", C.c_int64),
        ("rt_values", c_f32p), ("n_spectra", C.c_int64),
        ("mobility_values", c_f32p), ("n_mobility", C.c_int64),
        ("peak_start_idx", c_i64p), ("peak_stop_idx", c_i64p),
        ("mz_values", c_f32p), ("intensity_values", c_f32p), ("n_peaks", C.c_int64),
        ("zeroth_frame", C.c_int64), ("precursor_cycle_max_index", C.c_int64),
        ("scan_max_index", C.c_int64), ("frame_max_index", C.c_int64),
    ]


c_u16p = C.POINTER(C.c_uint16)


class RawFile4DDesc(C.Structure):
    _fields_ = [
        ("cycle", c_f64p), ("frames_per_cycle", C.c_int64), ("scans", C.c_int64),
        ("dia_precursor_cycle", c_i64p),
        ("rt_values", c_f64p), ("n_frames", C.c_int64),
        ("mobility_values", c_f64p),
        ("mz_values", c_f64p), ("n_tof", C.c_int64),
        ("tof_indptr", c_i64p), ("push_indices", c_u32p), ("intensity_values", c_u16p), ("n_events", C.c_int64),
        ("zeroth_frame", C.c_int64), ("precursor_cycle_max_index", C.c_int64),
        ("scan_max_index", C.c_int64), ("frame_max_index", C.c_int64),
    ]


class LibraryDesc(C.Structure):
    _fields_ = [
        ("n_precursors", C.c_int64),
        ("precursor_idx", c_u32p), ("frag_start_idx", c_u32p), ("frag_stop_idx", c_u32p),
        ("charge", c_u8p), ("rt", c_f32p), ("mobility", c_f32p), ("mz", c_f32p),
        ("isotopes", c_f32p), ("n_isotopes", C.c_int32),
        ("n_fragments", C.c_int64),
        ("frag_mz_library", c_f32p), ("frag_mz", c_f32p), ("frag_intensity", c_f32p),
        ("frag_type", c_u8p), ("frag_loss_type", c_u8p), ("frag_charge", c_u8p),
        ("frag_number", c_u8p), ("frag_position", c_u8p), ("frag_cardinality", c_u8p),
    ]


class SelectionConfig(C.Structure):
    _fields_ = [
        ("rt_tolerance", C.c_double), ("precursor_mz_tolerance", C.c_double),
        ("fragment_mz_tolerance", C.c_double), ("mobility_tolerance", C.c_double),
        ("candidate_count", C.c_int64), ("top_k_precursors", C.c_int64), ("top_k_fragments", C.c_int64),
        ("exclude_shared_ions", C.c_int32), ("kernel_size", C.c_int64),
        ("f_mobility", C.c_double), ("f_rt", C.c_double), ("center_fraction", C.c_double),
        ("min_size_mobility", C.c_int64), ("min_size_rt", C.c_int64),
        ("max_size_mobility", C.c_int64), ("max_size_rt", C.c_int64),
        ("use_weighted_score", C.c_int32), ("join_close_candidates", C.c_int32),
        ("join_close_candidates_scan_threshold", C.c_double),
        ("join_close_candidates_cycle_threshold", C.c_double),
        ("feature_std", C.c_double), ("feature_mean", C.c_double), ("feature_weight", C.c_double),
    ]


class CandidatesOut(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("precursor_idx", c_u32p), ("rank", c_u8p), ("score", c_f32p),
        ("scan_center", c_u32p), ("scan_start", c_u32p), ("scan_stop", c_u32p),
        ("frame_center", c_u32p), ("frame_start", c_u32p), ("frame_stop", c_u32p),
    ]


class ScoringConfig(C.Structure):
    _fields_ = [
        ("collect_fragments", C.c_int32), ("exclude_shared_ions", C.c_int32),
        ("top_k_fragments", C.c_uint32), ("top_k_isotopes", C.c_uint32), ("quant_window", C.c_uint32),
        ("quant_all", C.c_int32), ("precursor_mz_tolerance", C.c_float), ("fragment_mz_tolerance", C.c_float),
        ("experimental_xic", C.c_int32), ("quad_sigma", C.c_double * 2), ("quad_delta_mu", C.c_double * 2),
    ]


class CandidatesIn(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("lib_row", c_i64p), ("rank", c_u8p),
        ("scan_start", c_i64p), ("scan_stop", c_i64p), ("scan_center", c_i64p),
        ("frame_start", c_i64p), ("frame_stop", c_i64p), ("frame_center", c_i64p),
    ]


class CandidateTable(C.Structure):
    """adb_candidate_table: compacted candidate rows (score > 0); the first nine fields are laid out like CandidatesIn."""
    _fields_ = [
        ("n", C.c_int64), ("lib_row", c_i64p), ("rank", c_u8p),
        ("scan_start", c_i64p), ("scan_stop", c_i64p), ("scan_center", c_i64p),
        ("frame_start", c_i64p), ("frame_stop", c_i64p), ("frame_center", c_i64p),
        ("precursor_idx", c_u32p), ("score", c_f32p),
    ]


TABLE_COLUMNS = (("lib_row", np.int64), ("rank", np.uint8), ("scan_start", np.int64), ("scan_stop", np.int64),
                 ("scan_center", np.int64), ("frame_start", np.int64), ("frame_stop", np.int64), ("frame_center", np.int64),
                 ("precursor_idx", np.uint32), ("score", np.float32))


def alloc_candidate_table(n: int, alloc=np.zeros):
    """Host buffers for adb_fetch_candidate_table; ``alloc(shape, dtype)`` may hand out pinned memory."""
    arrs = {k: alloc(max(n, 1), dt) for k, dt in TABLE_COLUMNS}
    return arrs


def candidate_table_struct(arrs: dict, n: int) -> CandidateTable:
    d = CandidateTable()
    d.n = n
    for k, _ in TABLE_COLUMNS:
        setattr(d, k, ptr(arrs[k]))
    return d


def candidates_in_from_table(arrs: dict, n: int) -> CandidatesIn:
    d = CandidatesIn()
    d.n = n
    for k in ("lib_row", "rank", "scan_start", "scan_stop", "scan_center", "frame_start", "frame_stop", "frame_center"):
        setattr(d, k, ptr(arrs[k]))
    return d


class ScoresOut(C.Structure):
    _fields_ = [
        ("features", c_f32p), ("valid", c_u8p),
        ("fragment_mz_library", c_f32p), ("fragment_mz", c_f32p), ("fragment_mz_observed", c_f32p),
        ("fragment_height", c_f32p), ("fragment_intensity", c_f32p), ("fragment_mass_error", c_f32p),
        ("fragment_correlation", c_f32p),
        ("fragment_position", c_u8p), ("fragment_number", c_u8p), ("fragment_type", c_u8p),
        ("fragment_charge", c_u8p), ("fragment_loss_type", c_u8p),
    ]


class ScoresRagged(C.Structure):
    """adb_scores_ragged (include/alphadia_b200.h): valid rows + their kept fragment slots, flattened."""
    _fields_ = [
        ("row_capacity", C.c_int64), ("frag_capacity", C.c_int64), ("n_rows", C.c_int64), ("n_fragments", C.c_int64),
        ("row_index", c_i64p), ("features", c_f32p), ("frag_offset", c_i64p),
        ("fragment_mz_library", c_f32p), ("fragment_mz", c_f32p), ("fragment_mz_observed", c_f32p),
        ("fragment_height", c_f32p), ("fragment_intensity", c_f32p), ("fragment_mass_error", c_f32p),
        ("fragment_correlation", c_f32p),
        ("fragment_position", c_u8p), ("fragment_number", c_u8p), ("fragment_type", c_u8p),
        ("fragment_charge", c_u8p), ("fragment_loss_type", c_u8p),
    ]


_PTR = {
    np.dtype(np.float32): c_f32p, np.dtype(np.float64): c_f64p, np.dtype(np.int64): c_i64p,
    np.dtype(np.uint32): c_u32p, np.dtype(np.uint8): c_u8p, np.dtype(np.bool_): c_u8p,
    np.dtype(np.uint16): c_u16p, np.dtype(np.uint64): c_u64p,
}


def ptr(a: np.ndarray):
    """Typed pointer to a C-contiguous numpy array (no copy)."""
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(_PTR[a.dtype])


def as_c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def make_rawfile3d_desc(raw):
    """Descriptor from a RawFile3D-like object (see alphadia_b200.raw_data.adapt_dia_data)."""
    keep = dict(
        cycle=as_c(raw.cycle, np.float64), rt=as_c(raw.rt_values, np.float32),
        mob=as_c(raw.mobility_values, np.float32),
        ps=as_c(raw.peak_start_idx_list, np.int64), pe=as_c(raw.peak_stop_idx_list, np.int64),
        mz=as_c(raw.mz_values, np.float32), it=as_c(raw.intensity_values, np.float32),
    )
    if keep["cycle"].ndim != 4 or keep["cycle"].shape[0] != 1 or keep["cycle"].shape[2] != 1:
        raise ValueError(f"3-D raw file expects cycle of shape (1, L, 1, 2), got {keep['cycle'].shape}")
    d = RawFile3DDesc()
    d.cycle = ptr(keep["cycle"]); d.cycle_len = keep["cycle"].shape[1]
    d.rt_values = ptr(keep["rt"]); d.n_spectra = len(keep["rt"])
    d.mobility_values = ptr(keep["mob"]); d.n_mobility = len(keep["mob"])
    d.peak_start_idx = ptr(keep["ps"]); d.peak_stop_idx = ptr(keep["pe"])
    d.mz_values = ptr(keep["mz"]); d.intensity_values = ptr(keep["it"]); d.n_peaks = len(keep["mz"])
    d.zeroth_frame = int(raw.zeroth_frame)
    d.precursor_cycle_max_index = int(raw.precursor_cycle_max_index)
    d.scan_max_index = int(raw.scan_max_index)
    d.frame_max_index = int(raw.frame_max_index)
    return d, keep


def make_rawfile4d_desc(raw):
    """Descriptor from a RawFile4D-like object (timsTOF layout, see alphadia_b200.raw_data.adapt_dia_data)."""
    keep = dict(
        cycle=as_c(raw.cycle, np.float64), dpc=as_c(raw.dia_precursor_cycle, np.int64),
        rt=as_c(raw.rt_values, np.float64), mob=as_c(raw.mobility_values, np.float64),
        mz=as_c(raw.mz_values, np.float64), indptr=as_c(raw.tof_indptr, np.int64),
        push=as_c(raw.push_indices, np.uint32), it=as_c(raw.intensity_values, np.uint16),
    )
    cyc = keep["cycle"]
    if cyc.ndim != 4 or cyc.shape[0] != 1 or cyc.shape[3] != 2:
        raise ValueError(f"4-D raw file expects cycle of shape (1, frames, scans, 2), got {cyc.shape}")
    if len(keep["dpc"]) != cyc.shape[1] * cyc.shape[2] or len(keep["mob"]) != cyc.shape[2]:
        raise ValueError("dia_precursor_cycle / mobility_values do not match the cycle shape")
    d = RawFile4DDesc()
    d.cycle = ptr(cyc); d.frames_per_cycle = cyc.shape[1]; d.scans = cyc.shape[2]
    d.dia_precursor_cycle = ptr(keep["dpc"])
    d.rt_values = ptr(keep["rt"]); d.n_frames = len(keep["rt"])
    d.mobility_values = ptr(keep["mob"])
    d.mz_values = ptr(keep["mz"]); d.n_tof = len(keep["mz"])
    d.tof_indptr = ptr(keep["indptr"]); d.push_indices = ptr(keep["push"]); d.intensity_values = ptr(keep["it"])
    d.n_events = len(keep["push"])
    d.zeroth_frame = int(raw.zeroth_frame)
    d.precursor_cycle_max_index = int(raw.precursor_cycle_max_index)
    d.scan_max_index = int(raw.scan_max_index)
    d.frame_max_index = int(raw.frame_max_index)
    return d, keep


def make_library_desc(lib):
    """Descriptor from a dict of SoA arrays (alphadia_b200.library.FlatLibrary.arrays)."""
    keep = lib
    d = LibraryDesc()
    d.n_precursors = len(keep["precursor_idx"])
    for name in ("precursor_idx", "frag_start_idx", "frag_stop_idx", "charge", "rt", "mobility", "mz", "isotopes",
                 "frag_mz_library", "frag_mz", "frag_intensity", "frag_type", "frag_loss_type", "frag_charge",
                 "frag_number", "frag_position", "frag_cardinality"):
        setattr(d, name, ptr(keep[name]))
    d.n_isotopes = keep["isotopes"].shape[1]
    d.n_fragments = len(keep["frag_mz"])
    return d, keep


def alloc_candidates_out(n_rows: int):
    arrs = dict(
        precursor_idx=np.zeros(n_rows, np.uint32), rank=np.zeros(n_rows, np.uint8), score=np.zeros(n_rows, np.float32),
        scan_center=np.zeros(n_rows, np.uint32), scan_start=np.zeros(n_rows, np.uint32), scan_stop=np.zeros(n_rows, np.uint32),
        frame_center=np.zeros(n_rows, np.uint32), frame_start=np.zeros(n_rows, np.uint32), frame_stop=np.zeros(n_rows, np.uint32),
    )
    d = CandidatesOut()
    d.n_rows = n_rows
    for k, v in arrs.items():
        setattr(d, k, ptr(v))
    return d, arrs


FRAG_F32 = ("fragment_mz_library", "fragment_mz", "fragment_mz_observed", "fragment_height",
            "fragment_intensity", "fragment_mass_error", "fragment_correlation")
FRAG_U8 = ("fragment_position", "fragment_number", "fragment_type", "fragment_charge", "fragment_loss_type")


def alloc_scores_out(n: int, top_k: int):
    arrs = dict(features=np.zeros((n, NUM_FEATURES), np.float32), valid=np.zeros(n, np.uint8))
    for k in FRAG_F32:
        arrs[k] = np.zeros((n, top_k), np.float32)
    for k in FRAG_U8:
        arrs[k] = np.zeros((n, top_k), np.uint8)
    d = ScoresOut()
    for k, v in arrs.items():
        setattr(d, k, ptr(v))
    return d, arrs


def alloc_scores_ragged(row_capacity: int, frag_capacity: int, alloc=np.empty):
    """Host buffers of a ragged scoring result; ``alloc(shape, dtype)`` may hand out pinned memory."""
    rc, fc = max(int(row_capacity), 1), max(int(frag_capacity), 1)
    arrs = dict(row_index=alloc(rc, np.int64), features=alloc((rc, NUM_FEATURES), np.float32), frag_offset=alloc(rc + 1, np.int64))
    for k in FRAG_F32:
        arrs[k] = alloc(fc, np.float32)
    for k in FRAG_U8:
        arrs[k] = alloc(fc, np.uint8)
    return scores_ragged_struct(arrs), arrs


def scores_ragged_struct(arrs: dict) -> "ScoresRagged":
    d = ScoresRagged()
    d.row_capacity = int(arrs["row_index"].shape[0])
    d.frag_capacity = int(arrs[FRAG_F32[0]].shape[0])
    for k, v in arrs.items():
        setattr(d, k, ptr(v))
    return d


def make_candidates_in(lib_row, rank, scan_start, scan_stop, scan_center, frame_start, frame_stop, frame_center):
    keep = dict(
        lib_row=as_c(lib_row, np.int64), rank=as_c(rank, np.uint8),
        scan_start=as_c(scan_start, np.int64), scan_stop=as_c(scan_stop, np.int64), scan_center=as_c(scan_center, np.int64),
        frame_start=as_c(frame_start, np.int64), frame_stop=as_c(frame_stop, np.int64), frame_center=as_c(frame_center, np.int64),
    )
    d = CandidatesIn()
    d.n = len(keep["lib_row"])
    for k, v in keep.items():
        setattr(d, k, ptr(v))
    return d, keep

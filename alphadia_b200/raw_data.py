"""Adapters from the reference's DiaData objects to the engine's raw-file descriptor.

The reference hands ``dia_data`` (an ``AlphaRaw`` subclass, alphadia/raw_data/alpharaw_wrapper.py:22-156)
to ``CandidateSelection`` / ``CandidateScoring`` and calls ``dia_data.to_jitclass()``.  The engine
needs the same 15 fields (alphadia/search/jitclasses/alpharaw_jit.py:78-138); they are taken from
whichever of these the caller passes:

* ``alphadia_b200.synthetic.RawFile3D`` (public attribute names),
* a reference ``AlphaRaw`` wrapper (private ``_mz_values`` ... attributes),
* an ``AlphaRawJIT`` jitclass instance (public fields).
"""

from __future__ import annotations

from types import SimpleNamespace

import numpy as np


def _first(obj, *names):
    for n in names:
        if hasattr(obj, n):
            v = getattr(obj, n)
            if v is not None:
                return v
    raise AttributeError(f"{type(obj).__name__} has none of {names}")


def _is_timstof(dia_data) -> bool:
    """timsTOF layout = CSR by tof index (TimsTOFTransposeJIT fields), whatever ``has_mobility`` says."""
    return any(hasattr(dia_data, n) for n in ("tof_indptr", "_tof_indptr")) and any(
        hasattr(dia_data, n) for n in ("push_indices", "_push_indices")
    )


def adapt_dia_data_4d(dia_data) -> SimpleNamespace:
    """Normalise a timsTOF DiaData-like object (alphadia/raw_data/bruker.py:37-152 ``TimsTOFTranspose``, a
    ``TimsTOFTransposeJIT`` instance, or ``alphadia_b200.synthetic.RawFile4D``) to the field names
    `_abi.make_rawfile4d_desc` expects (jitclasses/bruker_jit.py:20-137)."""
    cycle = np.asarray(_first(dia_data, "cycle", "_cycle"), dtype=np.float64)
    rt = np.asarray(_first(dia_data, "rt_values", "_rt_values"), dtype=np.float64)
    fr, sc = cycle.shape[1], cycle.shape[2]
    ns = SimpleNamespace(
        cycle=cycle,
        rt_values=rt,
        mobility_values=np.asarray(_first(dia_data, "mobility_values", "_mobility_values"), dtype=np.float64),
        mz_values=np.asarray(_first(dia_data, "mz_values", "_mz_values"), dtype=np.float64),
        tof_indptr=np.asarray(_first(dia_data, "tof_indptr", "_tof_indptr"), dtype=np.int64),
        push_indices=np.asarray(_first(dia_data, "push_indices", "_push_indices"), dtype=np.uint32),
        intensity_values=np.asarray(_first(dia_data, "intensity_values", "_intensity_values"), dtype=np.uint16),
        dia_precursor_cycle=np.asarray(_first(dia_data, "dia_precursor_cycle", "_dia_precursor_cycle"), dtype=np.int64),
        zeroth_frame=int(_first(dia_data, "zeroth_frame", "_zeroth_frame")),
        scan_max_index=int(_first(dia_data, "scan_max_index", "_scan_max_index")),
        frame_max_index=int(_first(dia_data, "frame_max_index", "_frame_max_index")),
        has_mobility=True,
        is_4d=True,
    )
    try:
        ns.precursor_cycle_max_index = int(_first(dia_data, "precursor_cycle_max_index", "_precursor_cycle_max_index"))
    except AttributeError:
        ns.precursor_cycle_max_index = ns.frame_max_index // fr
    if ns.dia_precursor_cycle.shape[0] != fr * sc:
        raise ValueError("dia_precursor_cycle does not match the cycle shape")
    return ns


def adapt_dia_data(dia_data) -> SimpleNamespace:
    """Normalise a DiaData-like object to the field names `_abi.make_rawfile3d_desc` / `make_rawfile4d_desc` expect."""
    if _is_timstof(dia_data):
        return adapt_dia_data_4d(dia_data)
    if getattr(dia_data, "has_mobility", False):
        raise ValueError("raw file reports has_mobility but does not expose the timsTOF CSR arrays (tof_indptr, push_indices)")
    cycle = np.asarray(_first(dia_data, "cycle"), dtype=np.float64)
    rt = np.asarray(_first(dia_data, "rt_values"), dtype=np.float32)
    ns = SimpleNamespace(
        cycle=cycle,
        rt_values=rt,
        mobility_values=np.asarray(_first(dia_data, "mobility_values"), dtype=np.float32),
        peak_start_idx_list=np.asarray(_first(dia_data, "peak_start_idx_list", "_peak_start_idx_list"), dtype=np.int64),
        peak_stop_idx_list=np.asarray(_first(dia_data, "peak_stop_idx_list", "_peak_stop_idx_list"), dtype=np.int64),
        mz_values=np.asarray(_first(dia_data, "mz_values", "_mz_values"), dtype=np.float32),
        intensity_values=np.asarray(_first(dia_data, "intensity_values", "_intensity_values"), dtype=np.float32),
        zeroth_frame=int(_first(dia_data, "zeroth_frame", "_zeroth_frame")),
        scan_max_index=int(_first(dia_data, "scan_max_index", "_scan_max_index")),
        has_mobility=False,
        is_4d=False,
    )
    L = cycle.shape[1]
    try:
        ns.precursor_cycle_max_index = int(_first(dia_data, "precursor_cycle_max_index", "_precursor_cycle_max_index"))
    except AttributeError:
        ns.precursor_cycle_max_index = len(rt) // L
    try:
        ns.frame_max_index = int(_first(dia_data, "frame_max_index"))
    except AttributeError:
        ns.frame_max_index = len(rt) - 1
    return ns

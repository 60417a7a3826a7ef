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


def adapt_dia_data(dia_data) -> SimpleNamespace:
    """Normalise a 3-D DiaData-like object to the field names `_abi.make_rawfile3d_desc` expects."""
    if getattr(dia_data, "has_mobility", False):
        raise NotImplementedError(
            "ion-mobility (timsTOF, 4-D) raw files are not supported by the B200 engine yet "
            "(SURVEY.md §8 rows a2/a7/a18)"
        )
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

"""``CandidateSelection`` — drop-in for alphadia/search/selection/selection.py:547-737 on the B200 engine.

Same constructor, same ``__call__(thread_count, debug) -> DataFrame`` and the same output table
(``precursor_idx, rank, score, scan_center, scan_start, scan_stop, frame_center, frame_start,
frame_stop, elution_group_idx, decoy``; rows with ``score > 0`` only, config_df.py:258-298).
The per-precursor numba loop (selection.py:78-203) runs as one CUDA launch behind the C ABI
``adb_select_candidates``; ``thread_count`` is accepted and ignored.
"""

from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from alphadia_b200 import _lib
from alphadia_b200.config import CandidateSelectionConfig
from alphadia_b200.kernel import GaussianKernel
from alphadia_b200.library import assemble_library_arrays
from alphadia_b200.raw_data import adapt_dia_data

logger = logging.getLogger()

CANDIDATE_COLUMNS = [
    "precursor_idx", "rank", "score", "scan_center", "scan_start", "scan_stop",
    "frame_center", "frame_start", "frame_stop",
]


class CandidateSelection:
    def __init__(
        self,
        dia_data,
        precursors_flat: pd.DataFrame,
        fragments_flat: pd.DataFrame,
        config: CandidateSelectionConfig,
        rt_column: str,
        mobility_column: str,
        precursor_mz_column: str,
        fragment_mz_column: str,
        fwhm_rt: float = 5.0,
        fwhm_mobility: float = 0.012,
    ) -> None:
        self._dia_data = dia_data
        self._raw = adapt_dia_data(dia_data)
        # selection.py:598-600
        self.precursors_flat = precursors_flat.sort_values("precursor_idx").reset_index(drop=True)
        self.fragments_flat = fragments_flat
        self.config = config
        self.config_struct = config.to_struct()

        self.rt_column = rt_column
        self.precursor_mz_column = precursor_mz_column
        self.fragment_mz_column = fragment_mz_column
        self.mobility_column = mobility_column

        # selection.py:609-620
        gaussian_filter = GaussianKernel(
            self._raw,
            fwhm_rt=fwhm_rt,
            sigma_scale_rt=config.sigma_scale_rt,
            fwhm_mobility=fwhm_mobility,
            sigma_scale_mobility=config.sigma_scale_mobility,
            kernel_width=config.kernel_size,
            kernel_height=min(config.kernel_size, self._raw.scan_max_index + 1),
        )
        self.kernel = gaussian_filter.get_dense_matrix()

    def __call__(self, thread_count: int = 10, debug: bool = False) -> pd.DataFrame:
        logging.info("Starting candidate selection")
        del thread_count  # parallelism is the device's
        precursors = self.precursors_flat
        n_total = len(precursors)
        if debug:  # selection.py:652-654
            precursors = precursors.iloc[: min(10, n_total)]
        lib_arrays = assemble_library_arrays(
            precursors, self.fragments_flat, self.rt_column, self.mobility_column,
            self.precursor_mz_column, self.fragment_mz_column,
        )
        dev_raw = _lib.device_rawfile_for(self._dia_data, self._raw)
        dev_lib = _lib.DeviceLibrary(lib_arrays, device=dev_raw.device)
        try:
            # selection + `score > 0` filter on the device (config_df.py:270-298 candidate_container_to_df);
            # only the surviving rows travel back
            n = _lib.select_candidates_resident(dev_raw, dev_lib, self.config_struct, self.kernel)
            table = _lib.fetch_candidate_table(dev_raw, n)
        finally:
            dev_lib.close()
        self.last_timing = dev_raw.last_timing()
        container_dtypes = {"precursor_idx": np.uint32, "rank": np.uint8, "score": np.float32}
        candidate_df = pd.DataFrame(
            {c: table[c][:n].astype(container_dtypes.get(c, np.uint32), copy=False) for c in CANDIDATE_COLUMNS}
        )
        # selection.py:670-676: left merge with the precursor table on precursor_idx.  Every candidate row carries the
        # library row it came from, so with unique precursor_idx values the merge is a gather of two columns (row order,
        # columns and dtypes of DataFrame.merge(how="left")); duplicated precursor_idx values take the real merge.
        pidx = precursors["precursor_idx"].values
        if len(pidx) < 2 or bool(np.all(pidx[1:] > pidx[:-1])):
            rows = table["lib_row"][:n]
            candidate_df["elution_group_idx"] = precursors["elution_group_idx"].values[rows]
            candidate_df["decoy"] = precursors["decoy"].values[rows]
            return candidate_df
        return candidate_df.merge(
            self.precursors_flat[["precursor_idx", "elution_group_idx", "decoy"]],
            on="precursor_idx",
            how="left",
        )

"""DataFrame schemas at the hot-path boundary — the dtype contract of the drop-in.

Mirrors alphadia/validation/base.py:11-152 (Required/Optional cast in place, missing required
column -> ValueError) and alphadia/validation/schemas.py:11-120 (column sets and dtypes).
"""

from __future__ import annotations

import logging
import os

import numpy as np
import pandas as pd

logger = logging.getLogger()


class Property:
    required = False

    def __init__(self, name, type):
        self.name = name
        self.type = type

    def __call__(self, df: pd.DataFrame, logging: bool = True) -> bool:
        if self.name in df.columns:
            if df[self.name].dtype != self.type:
                df[self.name] = df[self.name].astype(self.type)
            return True
        return not self.required


class Optional(Property):
    required = False


class Required(Property):
    required = True


class Schema:
    def __init__(self, name, properties):
        self.name = name
        self.schema = properties
        for p in self.schema:
            if not isinstance(p, Property):
                raise ValueError("Schema must contain only Property objects")

    def validate(self, df: pd.DataFrame, logging: bool = True, warn_on_critical_values: bool = False) -> None:
        if warn_on_critical_values:
            self._warn_on_critical_values(df)
        for p in self.schema:
            if not p(df, logging=logging):
                raise ValueError(
                    f"Validation of {self.name} failed: Column {p.name} is not present in the dataframe"
                )

    @staticmethod
    def _warn_on_critical_values(df: pd.DataFrame) -> None:
        """validation/base.py:120-152 (NaN / Inf counts per float column).  ``sum`` is finite only if every element is,
        so the common all-finite case costs one read and no temporary.  Columns that are strided views into one shared
        2-D array (a feature matrix wrapped in a DataFrame) are cleared together: one row-major pass over that array
        (row blocks on a few threads) yields every column sum; only the columns whose sum is not finite are counted."""
        floats = []
        for col in df.columns:
            dtype = df[col].dtype
            if isinstance(dtype, np.dtype) and np.issubdtype(dtype, np.floating):  # extension dtypes (str) are skipped
                floats.append((col, df[col].values))
        shared: dict = {}
        for _, v in floats:
            base = v.base
            if (isinstance(base, np.ndarray) and base.ndim == 2 and base.dtype == v.dtype and base.flags.c_contiguous
                    and v.ndim == 1 and v.strides[0] == base.strides[0] and len(v) == base.shape[0]):
                shared.setdefault(id(base), base)
        column_finite: dict = {}
        with np.errstate(over="ignore", invalid="ignore"):
            for key, base in shared.items():
                column_finite[key] = np.isfinite(_column_sums(base))
            suspects = []
            singles = []
            for col, v in floats:
                key = id(v.base)
                if key in column_finite:
                    j = (v.__array_interface__["data"][0] - shared[key].__array_interface__["data"][0]) // v.itemsize
                    if 0 <= j < shared[key].shape[1]:
                        if not column_finite[key][j]:
                            suspects.append((col, v))
                        continue
                singles.append((col, v))
            # stand-alone columns: one sum each, on a few threads when the table is large (numpy releases the GIL)
            if len(singles) > 1 and len(df) >= (1 << 20):
                from concurrent.futures import ThreadPoolExecutor

                with ThreadPoolExecutor(max_workers=min(8, len(singles), os.cpu_count() or 1)) as pool:
                    finite = list(pool.map(lambda cv: bool(np.isfinite(cv[1].sum())), singles))
            else:
                finite = [bool(np.isfinite(v.sum())) for _, v in singles]
            suspects += [cv for cv, ok in zip(singles, finite) if not ok]
            order = {col: i for i, (col, _) in enumerate(floats)}
            for col, v in sorted(suspects, key=lambda cv: order[cv[0]]):
                n_nan = int(np.isnan(v).sum())
                n_inf = int(np.isinf(v).sum())
                if n_nan:
                    logger.warning(f"{col} has {n_nan} NaNs ( {n_nan / len(df) * 100:.2f} % out of {len(df)})")
                if n_inf:
                    logger.warning(f"{col} has {n_inf} Infs ( {n_inf / len(df) * 100:.2f} % out of {len(df)})")


def _column_sums(a: np.ndarray) -> np.ndarray:
    """Column sums of a C-contiguous 2-D array; large arrays are summed in row blocks on a few threads (numpy releases the
    GIL inside the reduction)."""
    n = a.shape[0]
    if a.size < (1 << 22):
        return a.sum(axis=0)
    from concurrent.futures import ThreadPoolExecutor

    workers = min(8, os.cpu_count() or 1)
    bounds = np.linspace(0, n, 4 * workers + 1).astype(np.int64)
    with ThreadPoolExecutor(max_workers=workers) as pool:
        parts = list(pool.map(lambda i: a[bounds[i]:bounds[i + 1]].sum(axis=0), range(len(bounds) - 1)))
    return np.sum(parts, axis=0)


_ISO = [Optional(f"i_{i}", np.float32) for i in range(10)]

precursors_flat_schema = Schema(
    "precursors_flat",
    [
        Required("elution_group_idx", np.uint32), Optional("score_group_idx", np.uint32),
        Required("precursor_idx", np.uint32), Required("channel", np.uint32), Required("decoy", np.uint8),
        Required("flat_frag_start_idx", np.uint32), Required("flat_frag_stop_idx", np.uint32),
        Required("charge", np.uint8),
        Required("rt_library", np.float32), Optional("rt_calibrated", np.float32),
        Required("mobility_library", np.float32), Optional("mobility_calibrated", np.float32),
        Required("mz_library", np.float32), Optional("mz_calibrated", np.float32),
        Required("proteins", object), Required("genes", object), *_ISO,
    ],
)

fragments_flat_schema = Schema(
    "fragments_flat",
    [
        Required("mz_library", np.float32), Optional("mz_calibrated", np.float32),
        Required("intensity", np.float32), Required("cardinality", np.uint8), Required("type", np.uint8),
        Required("loss_type", np.uint8), Required("charge", np.uint8), Required("number", np.uint8),
        Required("position", np.uint8),
    ],
)

candidates_schema = Schema(
    "candidates_df",
    [
        Required("elution_group_idx", np.uint32), Required("precursor_idx", np.uint32), Required("rank", np.uint8),
        Required("scan_start", np.int64), Required("scan_stop", np.int64), Required("scan_center", np.int64),
        Required("frame_start", np.int64), Required("frame_stop", np.int64), Required("frame_center", np.int64),
        Optional("score", np.float32), Optional("score_group_idx", np.uint32), Optional("channel", np.uint8),
        Optional("decoy", np.uint8), Optional("flat_frag_start_idx", np.uint32),
        Optional("flat_frag_stop_idx", np.uint32), Optional("mz_library", np.float32),
        Optional("mz_calibrated", np.float32), *_ISO,
    ],
)

features_schema = Schema(
    "candidate_features_df",
    [
        Required("precursor_idx", np.uint32), Required("elution_group_idx", np.uint32), Required("rank", np.uint8),
        Required("decoy", np.uint8), Required("channel", np.uint8), Required("charge", np.uint8),
        Required("flat_frag_start_idx", np.uint32), Required("flat_frag_stop_idx", np.uint32),
        Required("scan_center", np.int64), Required("scan_start", np.int64), Required("scan_stop", np.int64),
        Required("frame_center", np.int64), Required("frame_start", np.int64), Required("frame_stop", np.int64),
        Required("mz_library", np.float32), Optional("mz_calibrated", np.float32), Required("mz_observed", np.float32),
        Required("rt_library", np.float32), Optional("rt_calibrated", np.float32), Required("rt_observed", np.float32),
        Required("mobility_library", np.float32), Optional("mobility_calibrated", np.float32),
        Required("mobility_observed", np.float32), *_ISO,
    ],
)

fragment_features_schema = Schema(
    "fragment_features_df",
    [
        Required("precursor_idx", np.uint32), Required("rank", np.uint8), Required("elution_group_idx", np.uint32),
        Required("mz_library", np.float32), Required("mz_observed", np.float32), Required("mass_error", np.float32),
        Required("height", np.float32), Required("intensity", np.float32), Required("decoy", np.uint8),
    ],
)


def get_isotope_columns(colnames) -> np.ndarray:
    """alphadia/utils.py:53-74"""
    isotopes = []
    for col in colnames:
        if col[:2] == "i_":
            try:
                isotopes.append(int(col[2:]))
            except Exception:
                logging.warning(f"Column {col} does not seem to be a valid isotope column")
    isotopes = np.array(sorted(isotopes))
    if not np.all(np.diff(isotopes) == 1):
        logging.warning("Isotopes are not consecutive")
    return isotopes

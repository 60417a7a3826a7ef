"""``CandidateScoring`` — drop-in for alphadia/search/scoring/scoring.py:140-661 on the B200 engine.

Same keyword-only constructor, same ``__call__(candidates_df, thread_count, debug,
include_decoy_fragment_features) -> (features_df, fragments_df)``, same static helpers
(``merge_candidate_data``, ``merge_precursor_data``) and the same output tables: the 46
``DEFAULT_FEATURE_COLUMNS`` in order + ids + candidate/precursor columns + ``delta_rt, n_K, n_R, n_P``
(scoring.py:394-467) and the flattened fragment table (scoring.py:520-580).

The per-candidate numba loop (scoring.py:114-137 -> Candidate.process, candidate.py:166-481) is one CUDA
launch behind ``adb_score_candidates``.  The per-candidate jitclass construction of the reference
(``ScoreGroupContainer.build_from_df``, score_group.py:145-229) is replaced by direct SoA marshalling;
its checks (duplicate precursor in a score group, missing reference channel) are kept.
"""

from __future__ import annotations

import logging
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pandas as pd

from alphadia_b200 import _abi, _lib
from alphadia_b200.config import CandidateScoringConfig
from alphadia_b200.library import assemble_library_arrays
from alphadia_b200.raw_data import adapt_dia_data
from alphadia_b200.validation import (
    candidates_schema,
    features_schema,
    fragment_features_schema,
    fragments_flat_schema,
    get_isotope_columns,
    precursors_flat_schema,
)

logger = logging.getLogger()

DEFAULT_FEATURE_COLUMNS = [
    "base_width_mobility", "base_width_rt", "rt_observed", "mobility_observed",
    "mono_ms1_intensity", "top_ms1_intensity", "sum_ms1_intensity", "weighted_ms1_intensity",
    "weighted_mass_deviation", "weighted_mass_error", "mz_observed",
    "mono_ms1_height", "top_ms1_height", "sum_ms1_height", "weighted_ms1_height",
    "isotope_intensity_correlation", "isotope_height_correlation", "n_observations",
    "intensity_correlation", "height_correlation", "intensity_fraction", "height_fraction",
    "intensity_fraction_weighted", "height_fraction_weighted", "mean_observation_score",
    "sum_b_ion_intensity", "sum_y_ion_intensity", "diff_b_y_ion_intensity", "f_masked",
    "fragment_scan_correlation", "template_scan_correlation", "fragment_frame_correlation",
    "top3_frame_correlation", "template_frame_correlation", "top3_b_ion_correlation", "n_b_ions",
    "top3_y_ion_correlation", "n_y_ions", "cycle_fwhm", "mobility_fwhm", "delta_frame_peak",
    "top_3_ms2_mass_error", "mean_ms2_mass_error", "n_overlapping", "mean_overlapping_intensity",
    "mean_overlapping_mass_error",
]

DEFAULT_CANDIDATE_COLUMNS = [
    "elution_group_idx", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop",
]

DEFAULT_PRECURSOR_COLUMNS = [
    "rt_library", "mobility_library", "mz_library", "charge", "decoy", "channel",
    "flat_frag_start_idx", "flat_frag_stop_idx", "proteins", "genes", "sequence", "mods", "mod_sites",
]

FRAGMENT_COLUMNS = [
    "precursor_idx", "rank", "mz_library", "mz", "mz_observed", "height", "intensity", "mass_error",
    "correlation", "position", "number", "type", "charge", "loss_type",
]


def _get_isotope_column_names(colnames):
    return [f"i_{i}" for i in get_isotope_columns(colnames)]


def merge_missing_columns(left_df, right_df, right_columns, on=None, how="left"):
    """alphadia/search/scoring/utils.py:203-266 (missing columns keep the order of ``right_columns``)."""
    if isinstance(on, str):
        on = [on]
    if isinstance(right_columns, str):
        right_columns = [right_columns]
    seen = set()
    missing_from_left = [c for c in right_columns if c not in left_df.columns and not (c in seen or seen.add(c))]
    missing_from_right = [c for c in missing_from_left if c not in right_df.columns]
    if len(missing_from_left) == 0:
        return left_df
    if missing_from_right:
        raise ValueError(f"Columns {missing_from_right} must be present in right_df")
    if on is None:
        raise ValueError("Parameter on must be specified")
    if not all(col in left_df.columns for col in on):
        raise ValueError(f"Columns {on} must be present in left_df")
    if not all(col in right_df.columns for col in on):
        raise ValueError(f"Columns {on} must be present in right_df")
    if how not in ["left", "right", "inner", "outer"]:
        raise ValueError("Parameter how must be one of left, right, inner, outer")
    fast = _merge_by_sorted_key(left_df, right_df, on, missing_from_left) if how == "left" else None
    if fast is not None:
        return fast
    return left_df.merge(right_df[on + missing_from_left], on=on, how=how)


def _composite_key(df, on):
    """int64 key for one integer column or an (id, small non-negative integer) pair such as (precursor_idx, rank)."""
    cols = [df[c].values for c in on]
    if not all(np.issubdtype(c.dtype, np.integer) for c in cols):
        return None
    if len(cols) == 1:
        return cols[0].astype(np.int64, copy=False)
    if len(cols) == 2:
        a, b = cols[0].astype(np.int64, copy=False), cols[1].astype(np.int64, copy=False)
        if len(a) and (a.min() < 0 or a.max() >= (1 << 46) or b.min() < 0 or b.max() >= (1 << 16)):
            return None
        return (a << 16) | b
    return None


def _positions_in_sorted_unique(sorted_keys: np.ndarray, keys: np.ndarray):
    """Index of every key in a strictly ascending key array, or None when a key is missing.  Consecutive keys (the usual
    ``precursor_idx = 0 .. P-1``) are resolved by subtraction instead of a binary search per row."""
    n = len(sorted_keys)
    if n == 0:
        return None if len(keys) else np.zeros(0, dtype=np.int64)
    first, last = int(sorted_keys[0]), int(sorted_keys[-1])
    if last - first == n - 1 and (n < 2 or bool(np.all(sorted_keys[1:] > sorted_keys[:-1]))):  # consecutive values
        pos = keys.astype(np.int64, copy=False) - first
        if len(pos) and (pos.min() < 0 or pos.max() >= n):
            return None
        return pos
    pos = np.minimum(np.searchsorted(sorted_keys, keys), n - 1)
    if not np.array_equal(sorted_keys[pos], keys):
        return None
    return pos


def _merge_by_sorted_key(left_df, right_df, on, columns):
    """Left merge as a gather: valid when the right keys are unique and every left key occurs in the right frame (the
    result of ``DataFrame.merge(how="left")`` is then the left frame, in order, plus the gathered columns).  The pandas
    hash merge costs seconds at millions of candidates; returns None when the preconditions do not hold."""
    if len(left_df) == 0 or len(right_df) == 0:
        return None
    lk, rk = _composite_key(left_df, on), _composite_key(right_df, on)
    if lk is None or rk is None:
        return None
    order = None
    if len(rk) > 1 and not np.all(rk[1:] > rk[:-1]):
        order = np.argsort(rk, kind="stable")
        rk = rk[order]
        if not np.all(rk[1:] > rk[:-1]):
            return None  # duplicate keys on the right: a real merge multiplies rows
    pos = _positions_in_sorted_unique(rk, lk)
    if pos is None:
        return None  # unmatched left keys would become NaN rows
    if order is not None:
        pos = order[pos]
    out = left_df.reset_index(drop=True)
    new_cols = {}
    for c in columns:
        src = right_df[c]
        if isinstance(src.dtype, np.dtype) and src.dtype != object:
            new_cols[c] = src.values[pos]
        else:  # object / extension columns: keep the dtype (pandas would re-infer `str` from an object array)
            new_cols[c] = _series_like(_gather_source(src)[pos], src.dtype, out.index)
    # one concat instead of one block-manager insert per column
    return pd.concat([out, pd.DataFrame(new_cols, index=out.index, copy=False)], axis=1)


def _count_residues_arrow(sequences, res):
    """pyarrow's vectorised ``count_substring`` when the column is Arrow-backed (the ``str`` dtype of pandas >= 3): no
    conversion of the strings to Python objects.  None when not applicable (other dtypes, missing values)."""
    pa_array = getattr(sequences, "_pa_array", None)
    if pa_array is None:
        return None
    try:
        import pyarrow.compute as pc

        if pa_array.null_count:
            return None
        return [np.asarray(pc.count_substring(pa_array, r)).astype(np.int64, copy=False) for r in res]
    except Exception:  # an older pyarrow without the kernel: the numpy path below gives the same counts
        return None


def _gather_source(series: pd.Series):
    """What to index when rows of a column are gathered: the numpy array of a numpy-typed column (object columns too - wrapped
    as an extension array they would be scanned for missing values on every Series construction), else the extension array."""
    return series.values if isinstance(series.dtype, np.dtype) else series.array


def _series_like(values, dtype, index) -> pd.Series:
    """A Series of ``dtype`` from gathered values.  An extension array that already has the dtype (``str`` columns) is wrapped
    as it is: passing ``dtype=`` again makes pandas re-validate every element (a missing-value scan per column)."""
    if not isinstance(dtype, np.dtype) and not isinstance(values, np.ndarray) and getattr(values, "dtype", None) == dtype:
        return pd.Series(values, index=index, copy=False)
    return pd.Series(values, index=index, dtype=dtype)


def gather_missing_columns(left_df, right_df, right_columns, pos):
    """``merge_missing_columns(left_df, right_df, right_columns, on=key, how="left")`` when row ``i`` of ``left_df`` is known
    to match row ``pos[i]`` of ``right_df`` (unique right keys, every left key present): the columns missing on the left are
    gathered by position, in the order of ``right_columns`` - no key search."""
    seen = set()
    missing = [c for c in right_columns if c not in left_df.columns and not (c in seen or seen.add(c))]
    if not missing:
        return left_df
    absent = [c for c in missing if c not in right_df.columns]
    if absent:
        raise ValueError(f"Columns {absent} must be present in right_df")
    out = left_df.reset_index(drop=True)
    numeric = {c for c in missing if isinstance(right_df[c].dtype, np.dtype) and right_df[c].dtype != object}
    # numpy and pyarrow release the GIL inside the gathers: numeric and Arrow-backed string columns are taken side by side
    gathered = _run_column_tasks({c: ((lambda a: a[pos]), _gather_source(right_df[c])) for c in missing}, len(out))
    new_cols = {}
    for c in missing:
        if c in numeric:
            new_cols[c] = gathered[c]
        else:  # object / extension columns: keep the dtype (pandas would re-infer `str` from an object array)
            new_cols[c] = _series_like(gathered[c], right_df[c].dtype, out.index)
    return pd.concat([out, pd.DataFrame(new_cols, index=out.index, copy=False)], axis=1)


def count_residues(sequences, residues) -> list:
    """``Series.str.count(r)`` for single-character patterns (scoring.py:461-463 n_K / n_R / n_P) without a Python call
    per row: the distinct sequences are counted once as fixed-width code points and the counts are gathered.
    Returns one int64 array per residue (a single array when ``residues`` is a single character)."""
    single = isinstance(residues, str) and len(residues) == 1
    res = [residues] if single else list(residues)
    pa_counts = _count_residues_arrow(sequences, res)
    if pa_counts is not None:
        return pa_counts[0] if single else pa_counts
    values = np.asarray(sequences, dtype=object)
    if len(values) == 0:
        out = [np.zeros(0, dtype=np.int64) for _ in res]
        return out[0] if single else out
    codes, uniques = pd.factorize(values, use_na_sentinel=True)
    if (codes < 0).any() or not all(isinstance(u, str) for u in uniques[: min(len(uniques), 64)]):
        out = [pd.Series(values).str.count(r).values for r in res]  # missing or non-string entries: pandas semantics
        return out[0] if single else out
    fixed = np.asarray(uniques, dtype=str)
    width = fixed.dtype.itemsize // 4
    if width == 0:
        out = [np.zeros(len(values), dtype=np.int64) for _ in res]
        return out[0] if single else out
    points = fixed.view(np.uint32).reshape(len(fixed), width)
    out = [(points == ord(r)).sum(axis=1).astype(np.int64)[codes] for r in res]
    return out[0] if single else out


def _run_column_tasks(tasks: dict, n_rows: int) -> dict:
    """Evaluate ``{column: (function, array)}`` and keep the column order; large tables use a small thread pool."""
    if n_rows < (1 << 20) or len(tasks) < 2:
        return {c: f(a) for c, (f, a) in tasks.items()}
    workers = min(8, len(tasks), os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=workers) as pool:
        futures = {c: pool.submit(f, a) for c, (f, a) in tasks.items()}
        return {c: fut.result() for c, fut in futures.items()}


def _sorted_by(df: pd.DataFrame, by) -> pd.DataFrame:
    """``df.sort_values(by=by)`` for numeric key columns (stable lexicographic order), skipping the row gather when the
    frame is already in order — the usual case for a candidate table that comes straight from the selection."""
    if len(df) < 2:
        return df
    keys = [df[c].values for c in by]
    if not all(isinstance(k, np.ndarray) and k.dtype.kind in "iub" for k in keys):
        return df.sort_values(by=by)
    keys = [k.astype(np.int64, copy=False) for k in keys]
    in_order = np.ones(len(df) - 1, dtype=bool)
    tie = np.ones(len(df) - 1, dtype=bool)
    for k in keys:  # lexicographic "previous row <= next row"
        in_order &= ~(tie & (k[1:] < k[:-1]))
        tie &= k[1:] == k[:-1]
        if not in_order.all():
            break
    if in_order.all():
        return df
    return df.take(np.lexsort(tuple(reversed(keys))))


def calculate_score_groups(input_df: pd.DataFrame, group_channels: bool = False) -> pd.DataFrame:
    """alphadia/search/scoring/utils.py:269-410."""
    if "rank" in input_df.columns:
        input_df = _sorted_by(input_df, ["elution_group_idx", "decoy", "rank", "precursor_idx"])
        rank_values = input_df["rank"].values
    else:
        input_df = _sorted_by(input_df, ["elution_group_idx", "decoy", "precursor_idx"])
        rank_values = np.zeros(len(input_df), dtype=np.uint32)
    input_df = input_df.reset_index(drop=True)
    if group_channels:
        eg = input_df["elution_group_idx"].values
        decoy = input_df["decoy"].values
        n = len(eg)
        if n == 0:
            groups = np.zeros(0, dtype=np.uint32)
        else:
            change = np.zeros(n, dtype=bool)
            change[1:] = (eg[1:] != eg[:-1]) | (decoy[1:] != decoy[:-1]) | (rank_values[1:] != rank_values[:-1])
            groups = np.cumsum(change).astype(np.uint32)
        input_df["score_group_idx"] = groups
        return _sorted_by(input_df, ["score_group_idx", "precursor_idx"]).reset_index(drop=True)
    # one group per candidate: score_group_idx ascends with the row, the final sort of the reference is the identity
    input_df["score_group_idx"] = np.arange(len(input_df), dtype=np.uint32)
    return input_df


def logistic_rectangle(mu1, mu2, sigma1, sigma2, x):
    """alphadia/search/scoring/quadrupole.py:12-44: difference of two logistic edges."""
    with np.errstate(over="ignore"):  # exp overflow far outside the window gives the correct limit 0
        return 1 / (1 + np.exp(-((x - mu1) / sigma1))) - 1 / (1 + np.exp(-((x - mu2) / sigma2)))


def expand_cycle(cycle, lower_mz, upper_mz):
    """alphadia/search/scoring/quadrupole.py:339-347."""
    new_cycle = cycle.copy()
    new_cycle[..., 0] -= lower_mz * (new_cycle[..., 0] > 0)
    new_cycle[..., 1] += upper_mz * (new_cycle[..., 1] > 0)
    return new_cycle


def quadrupole_transfer_function_single(quadrupole_calibration_jit, observation_indices, scan_indices, isotope_mz):
    """alphadia/search/scoring/quadrupole.py:261-301: transfer efficiency ``[n_isotopes, n_observations, n_scans]`` (the
    device kernels evaluate the same expression per candidate, adb_score.cu / adb_score4d.cu)."""
    isotope_mz = np.asarray(isotope_mz, dtype=np.float64)
    observation_indices = np.asarray(observation_indices, dtype=np.int64)
    scan_indices = np.asarray(scan_indices, dtype=np.int64)
    n_i, n_o, n_s = len(isotope_mz), len(observation_indices), len(scan_indices)
    mz_column = np.repeat(isotope_mz, n_s * n_o)
    observation_column = np.tile(np.repeat(observation_indices, n_s), n_i)
    scan_column = np.tile(scan_indices, n_i * n_o)
    return quadrupole_calibration_jit.predict(observation_column, scan_column, mz_column).reshape(n_i, n_o, n_s)


class SimpleQuadrupoleJit:
    """State of alphadia/search/scoring/quadrupole.py:46-128 (uncalibrated: sigma 0.2, delta_mu 0); the device kernels take
    ``sigma`` and ``delta_mu`` from here (``adb_scoring_config.quad_sigma / quad_delta_mu``)."""

    def __init__(self, cycle):
        self.cycle = np.ascontiguousarray(cycle, dtype=np.float64)
        self.sigma = np.array([0.2, 0.2])
        self.delta_mu = np.array([0.0, 0.0])
        self._cycle_calibrated = None
        self._calibrated_provider = None  # set by SimpleQuadrupole: the calibrated cycle is computed on first use

    @property
    def cycle_calibrated(self):
        if self._cycle_calibrated is None:
            self._cycle_calibrated = self._calibrated_provider() if self._calibrated_provider else self.cycle
        return self._cycle_calibrated

    @property
    def dia_mz_cycle_calibrated(self):
        c = self.cycle_calibrated
        return np.reshape(c, (c.shape[1] * c.shape[2], 2))

    def predict(self, P, S, X):  # quadrupole.py:80-115
        P, S = np.asarray(P, dtype=np.int64), np.asarray(S, dtype=np.int64)
        mu1 = self.cycle[0, P, S, 0] + self.delta_mu[0]
        mu2 = self.cycle[0, P, S, 1] + self.delta_mu[1]
        return logistic_rectangle(mu1, mu2, self.sigma[0], self.sigma[1], np.asarray(X, dtype=np.float64))

    def set_cycle_calibrated(self, cycle_calibrated):  # quadrupole.py:117-121
        self._cycle_calibrated = cycle_calibrated

    def get_dia_mz_cycle(self, lower_mz, upper_mz):  # quadrupole.py:123-127
        expanded = expand_cycle(self.cycle_calibrated, lower_mz, upper_mz)
        return np.reshape(expanded, (expanded.shape[1] * expanded.shape[2], 2))


class SimpleQuadrupole:
    """alphadia/search/scoring/quadrupole.py:130-258 on the host.  ``get_calibrated_cycle`` evaluates the
    2000-point transfer function of all (frame, scan) windows in one vectorised pass — the reference loops over them in
    Python, which is slow for timsTOF cycles; the result only feeds debug plots and ``get_dia_mz_cycle``."""

    def __init__(self, cycle):
        self.cycle = cycle
        self.jit = SimpleQuadrupoleJit(cycle)
        self.jit._calibrated_provider = self.get_calibrated_cycle  # lazily: features never read it

    def fit(self, P, S, X, y):  # quadrupole.py:166-211
        from scipy.optimize import curve_fit

        mu1 = self.jit.cycle[0, P, S, 0]
        mu2 = self.jit.cycle[0, P, S, 1]
        X_train = np.stack([mu1, mu2, X], axis=1)

        def _wrapper(X, sigma1, sigma2, delta_mu1, delta_mu2):
            return logistic_rectangle(X[:, 0] + delta_mu1, X[:, 1] + delta_mu2, sigma1, sigma2, X[:, 2])

        p0 = np.concatenate([self.jit.sigma, self.jit.delta_mu])
        popt, _ = curve_fit(_wrapper, X_train, y, p0=p0)
        self.jit.sigma = popt[:2]
        self.jit.delta_mu = popt[2:]
        self.jit.set_cycle_calibrated(self.get_calibrated_cycle())
        return self

    def predict(self, P, S, X):
        return self.jit.predict(P, S, X)

    def get_calibrated_cycle(self, treshold=0.01):  # quadrupole.py:227-258
        cycle = self.jit.cycle
        non_zero = cycle[cycle > 0]
        new_cycle = cycle.copy()
        if non_zero.size == 0:
            return new_cycle
        lowest_mz, highest_mz = np.min(non_zero), np.max(non_zero)
        mz_width = highest_mz - lowest_mz
        mz_space = np.linspace(lowest_mz - mz_width * 0.1, highest_mz + mz_width * 0.1, 2000)
        lo, hi = cycle[0, :, :, 0], cycle[0, :, :, 1]
        active = lo > 0
        if not active.any():
            return new_cycle
        mu1 = lo[active][:, None] + self.jit.delta_mu[0]
        mu2 = hi[active][:, None] + self.jit.delta_mu[1]
        new_lo, new_hi = np.empty(mu1.shape[0]), np.empty(mu1.shape[0])
        for a in range(0, mu1.shape[0], 4096):  # bounded temporaries: 4096 windows x 2000 points
            inten = logistic_rectangle(mu1[a:a + 4096], mu2[a:a + 4096], self.jit.sigma[0], self.jit.sigma[1], mz_space[None, :])
            above = inten > treshold
            if not above.any(axis=1).all():
                raise ValueError("zero-size array to reduction operation minimum which has no identity")  # as np.min(q_range)
            first = np.argmax(above, axis=1)
            last = above.shape[1] - 1 - np.argmax(above[:, ::-1], axis=1)
            new_lo[a:a + 4096], new_hi[a:a + 4096] = mz_space[first], mz_space[last]
        out_lo, out_hi = new_cycle[0, :, :, 0], new_cycle[0, :, :, 1]
        out_lo[active], out_hi[active] = new_lo, new_hi
        return new_cycle


class CandidateScoring:
    """Calculate features for each precursor candidate used in scoring."""

    def __init__(
        self,
        *,
        dia_data,
        precursors_flat: pd.DataFrame,
        fragments_flat: pd.DataFrame,
        rt_column: str,
        mobility_column: str,
        precursor_mz_column: str,
        fragment_mz_column: str,
        config: CandidateScoringConfig | None = None,
        quadrupole_calibration=None,
    ):
        self._dia_data = dia_data
        self._raw_arrays = None  # adapted on first use, like the reference's dia_data.to_jitclass() (scoring.py:622)

        precursors_flat_schema.validate(precursors_flat, warn_on_critical_values=True)
        self.precursors_flat_df = precursors_flat

        fragments_flat_schema.validate(fragments_flat, warn_on_critical_values=True)
        self.fragments_flat = fragments_flat

        if quadrupole_calibration is None:
            self.quadrupole_calibration = SimpleQuadrupole(dia_data.cycle)  # scoring.py:210
        else:
            self.quadrupole_calibration = quadrupole_calibration

        self.config = CandidateScoringConfig() if config is None else config

        self.rt_column = rt_column
        self.mobility_column = mobility_column
        self.precursor_mz_column = precursor_mz_column
        self.fragment_mz_column = fragment_mz_column

    # ---- properties mirroring the reference ---------------------------------------------------
    @property
    def dia_data(self):
        return self._dia_data

    @property
    def _raw(self):
        if self._raw_arrays is None:
            self._raw_arrays = adapt_dia_data(self._dia_data)
        return self._raw_arrays

    @property
    def precursors_flat_df(self) -> pd.DataFrame:
        return self._precursors_flat_df

    @precursors_flat_df.setter
    def precursors_flat_df(self, precursors_flat_df) -> None:
        precursors_flat_schema.validate(precursors_flat_df, warn_on_critical_values=True)
        self._precursors_flat_df = precursors_flat_df.sort_values(by="precursor_idx")

    @property
    def fragments_flat_df(self) -> pd.DataFrame:
        return self._fragments_flat

    @fragments_flat_df.setter
    def fragments_flat_df(self, fragments_flat: pd.DataFrame) -> None:
        fragments_flat_schema.validate(fragments_flat, warn_on_critical_values=True)
        self._fragments_flat = fragments_flat

    @property
    def quadrupole_calibration(self):
        return self._quadrupole_calibration

    @quadrupole_calibration.setter
    def quadrupole_calibration(self, quadrupole_calibration) -> None:
        if not hasattr(quadrupole_calibration, "jit"):
            raise AttributeError("quadrupole_calibration must have a jit method")
        self._quadrupole_calibration = quadrupole_calibration

    @property
    def config(self) -> CandidateScoringConfig:
        return self._config

    @config.setter
    def config(self, config: CandidateScoringConfig) -> None:
        config.validate()
        self._config = config

    # ---- marshalling -----------------------------------------------------------------------------
    def assemble_candidates(self, candidates_df: pd.DataFrame, lib_precursor_idx: np.ndarray):
        """Replaces assemble_score_group_container (scoring.py:273-353): returns the sorted candidates
        frame, the ``adb_candidates_in`` struct (+keepalive) and the mask of rows sent to the device."""
        precursor_columns = [
            "channel", "flat_frag_start_idx", "flat_frag_stop_idx", "charge", "decoy", "channel",
            self.precursor_mz_column,
        ] + _get_isotope_column_names(self.precursors_flat_df.columns)
        candidates_df = merge_missing_columns(
            candidates_df, self.precursors_flat_df, precursor_columns, on=["precursor_idx"], how="left"
        )
        if "channel" not in candidates_df.columns:
            candidates_df["channel"] = np.zeros(len(candidates_df), dtype=np.uint8)
        if "i_0" not in candidates_df.columns:
            candidates_df["i_0"] = np.ones(len(candidates_df), dtype=np.float32)
        candidates_df = calculate_score_groups(candidates_df, group_channels=self.config.score_grouped)
        candidates_schema.validate(candidates_df, warn_on_critical_values=True)

        pidx = candidates_df["precursor_idx"].values
        sg = candidates_df["score_group_idx"].values
        # score_group.py:221-226: a precursor may appear once per score group
        if len(pidx) > 1 and np.any((pidx[1:] == pidx[:-1]) & (sg[1:] == sg[:-1])):
            raise ValueError("precursor_idx must be unique within a score group")
        process = np.ones(len(candidates_df), dtype=bool)
        if self.config.reference_channel >= 0 and len(candidates_df):
            # score_group.py:49-63: groups without the reference channel are skipped entirely
            has_ref = candidates_df["channel"].values == self.config.reference_channel
            groups_with_ref = np.unique(sg[has_ref])
            process = np.isin(sg, groups_with_ref)

        lib_row = _positions_in_sorted_unique(lib_precursor_idx.astype(np.int64, copy=False), pidx.astype(np.int64, copy=False))
        if lib_row is None:
            raise ValueError("candidates_df contains precursor_idx values that are not in precursors_flat")
        sel = np.flatnonzero(process)
        cin, keep = _abi.make_candidates_in(
            lib_row[sel], candidates_df["rank"].values[sel],
            candidates_df["scan_start"].values[sel], candidates_df["scan_stop"].values[sel],
            candidates_df["scan_center"].values[sel],
            candidates_df["frame_start"].values[sel], candidates_df["frame_stop"].values[sel],
            candidates_df["frame_center"].values[sel],
        )
        return candidates_df, cin, keep, sel

    # ---- result collection (scoring.py:394-580) ---------------------------------------------------
    def collect_candidates(self, candidates_df, psm, feature_columns=None, candidate_columns=None,
                           precursor_df_columns=None) -> pd.DataFrame:
        if feature_columns is None:
            feature_columns = DEFAULT_FEATURE_COLUMNS.copy()
        if candidate_columns is None:
            candidate_columns = DEFAULT_CANDIDATE_COLUMNS.copy()
        if precursor_df_columns is None:
            precursor_df_columns = DEFAULT_PRECURSOR_COLUMNS.copy()
        if hasattr(psm, "to_precursor_df"):  # an OutputPsmDF-like object (output.py:92-97), as the reference passes
            precursor_idx, rank, features = psm.to_precursor_df()
        elif psm.get("ragged"):  # adb_score_candidates_ragged: the valid rows only, already compacted
            precursor_idx, rank, features = psm["precursor_idx"], psm["rank"], psm["features"]
        else:  # the arrays adb_score_candidates filled
            valid = psm["valid"].astype(bool)
            if valid.all():
                precursor_idx, rank, features = psm["precursor_idx"], psm["rank"], psm["features"]
            else:
                precursor_idx, rank, features = psm["precursor_idx"][valid], psm["rank"][valid], psm["features"][valid]
        candidates_psm_df = pd.DataFrame(features, columns=feature_columns, copy=False)  # one block, no second copy
        candidates_psm_df["precursor_idx"] = precursor_idx
        candidates_psm_df["rank"] = rank
        positions = psm.get("positions") if isinstance(psm, dict) else None
        if positions is not None:
            # the ragged result knows which candidate row and which precursor row every feature row belongs to: the two left
            # merges (scoring.py:436-459) are gathers by position ((precursor_idx, rank) is unique after the score-group check,
            # precursor_idx is unique in precursors_flat)
            cols = candidate_columns + (["score"] if "score" in candidates_df.columns else [])
            candidates_psm_df = gather_missing_columns(candidates_psm_df, positions["candidates_df"], cols, positions["candidate_rows"])
            pcols = precursor_df_columns + _get_isotope_column_names(self.precursors_flat_df.columns)
            for col in [self.rt_column, self.mobility_column, self.precursor_mz_column]:
                if col not in pcols:
                    pcols.append(col)
            candidates_psm_df = gather_missing_columns(candidates_psm_df, self.precursors_flat_df, pcols, positions["precursor_rows"])
        else:
            candidates_psm_df = self.merge_candidate_data(candidates_psm_df, candidates_df, candidate_columns)
            candidates_psm_df = self.merge_precursor_data(
                candidates_psm_df, self.precursors_flat_df, self.rt_column, self.mobility_column,
                self.precursor_mz_column, precursor_df_columns,
            )
        candidates_psm_df["delta_rt"] = candidates_psm_df["rt_observed"] - candidates_psm_df[self.rt_column]
        flat = self.precursors_flat_df
        if (positions is not None and "sequence" in flat.columns and "sequence" not in cols and "sequence" not in feature_columns
                and len(flat) <= len(candidates_psm_df)):
            # the sequences were gathered from precursors_flat by row: count once per precursor and gather the counts
            prow = positions["precursor_rows"]
            n_k, n_r, n_p = (c[prow] for c in count_residues(flat["sequence"].array, ["K", "R", "P"]))
        else:
            n_k, n_r, n_p = count_residues(candidates_psm_df["sequence"].array, ["K", "R", "P"])
        candidates_psm_df["n_K"], candidates_psm_df["n_R"], candidates_psm_df["n_P"] = n_k, n_r, n_p
        return candidates_psm_df

    @staticmethod
    def merge_candidate_data(df, candidates_df, candidate_columns=None):
        if candidate_columns is None:
            candidate_columns = DEFAULT_CANDIDATE_COLUMNS.copy()
        candidate_columns += ["score"] if "score" in candidates_df.columns else []
        return merge_missing_columns(df, candidates_df, candidate_columns, on=["precursor_idx", "rank"], how="left")

    @staticmethod
    def merge_precursor_data(df, precursors_flat_df, rt_column, mobility_column, precursor_mz_column,
                             precursor_df_columns=None):
        if precursor_df_columns is None:
            precursor_df_columns = DEFAULT_PRECURSOR_COLUMNS.copy()
        precursor_df_columns = precursor_df_columns + _get_isotope_column_names(precursors_flat_df.columns)
        for col in [rt_column, mobility_column, precursor_mz_column]:
            if col not in precursor_df_columns:
                precursor_df_columns.append(col)
        return merge_missing_columns(df, precursors_flat_df, precursor_df_columns, on=["precursor_idx"], how="left")

    def collect_fragments(self, candidates_df, psm) -> pd.DataFrame:
        """output.py:72-90 + scoring.py:520-580: one row per fragment slot with ``mz_library > 0``.  The slot mask is
        applied column by column on a few threads (numpy releases the GIL inside the gathers), candidate-level columns
        are expanded with per-candidate slot counts, and ``elution_group_idx`` / ``decoy`` are looked up once per
        candidate instead of once per fragment row."""
        lib_mz = psm["fragment_mz_library"]
        if psm.get("ragged"):  # already flattened on the device: slot arrays are the columns, one count per valid candidate
            counts = psm["fragment_counts"]
            n, n_rows, top_k, mask = len(counts), int(lib_mz.shape[0]), 0, None

            def slots(a):
                return a

            def per_candidate(a):
                return np.repeat(a, counts)
        else:
            n, top_k = lib_mz.shape
            mask = lib_mz.reshape(-1) > 0
            n_rows = int(np.count_nonzero(mask))
        if mask is None:
            pass
        elif n_rows == mask.size:  # every slot filled: the flat arrays are the columns
            def slots(a):
                return a.reshape(-1)

            def per_candidate(a):
                return np.repeat(a, top_k)
        else:
            counts = np.count_nonzero(mask.reshape(n, top_k), axis=1)

            def slots(a):
                return a.reshape(-1)[mask]

            def per_candidate(a):
                return np.repeat(a, counts)

        tasks = {"precursor_idx": (per_candidate, psm["precursor_idx"]), "rank": (per_candidate, psm["rank"])}
        for col in FRAGMENT_COLUMNS[2:]:
            tasks[col] = (slots, psm["fragment_" + col])
        right_columns = ["elution_group_idx", "decoy"]
        positions = psm.get("positions")
        if positions is not None and len(positions["precursor_rows"]) == n:  # row mapping known: gather, no key search
            lookup_per_candidate = True
            for col in right_columns:
                tasks[col] = (per_candidate, self.precursors_flat_df[col].values[positions["precursor_rows"]])
        else:
            per_cand_df = merge_missing_columns(pd.DataFrame({"precursor_idx": psm["precursor_idx"]}), self.precursors_flat_df,
                                                right_columns, on=["precursor_idx"], how="left")
            lookup_per_candidate = len(per_cand_df) == n  # False only if precursors_flat repeats a precursor_idx
            if lookup_per_candidate:
                for col in right_columns:
                    tasks[col] = (per_candidate, per_cand_df[col].values)
        df = pd.DataFrame(_run_column_tasks(tasks, n_rows), copy=False)  # fresh arrays: no consolidation copy
        if lookup_per_candidate:
            return df
        return merge_missing_columns(df, self.precursors_flat_df, right_columns, on=["precursor_idx"], how="left")

    # ---- call ----------------------------------------------------------------------------------
    def __call__(self, candidates_df, thread_count=10, debug=False, include_decoy_fragment_features=False):
        logger.info("Starting candidate scoring")
        del thread_count, include_decoy_fragment_features  # the latter is unused by the reference as well

        if "cardinality" not in self.fragments_flat.columns:  # scoring.py:368-375
            logger.warning("Fragment cardinality column not found in fragment dataframe. Setting cardinality to 1.")
            self.fragments_flat["cardinality"] = np.ones(len(self.fragments_flat), dtype=np.uint8)
        lib_arrays = assemble_library_arrays(
            self.precursors_flat_df, self.fragments_flat, self.rt_column, self.mobility_column,
            self.precursor_mz_column, self.fragment_mz_column,
        )
        candidates_schema.validate(candidates_df, warn_on_critical_values=True)
        sorted_df, cin, keep, sel = self.assemble_candidates(candidates_df, lib_arrays["precursor_idx"])
        n_total = len(sorted_df)
        if debug:  # scoring.py:628-631: only the first 10 score groups
            logger.info("Debug mode enabled. Processing only the first 10 score groups")
            first = np.unique(sorted_df["score_group_idx"].values)[:10]
            in_first = np.isin(sorted_df["score_group_idx"].values[sel], first)
            sel = sel[in_first]
            cin, keep = _abi.make_candidates_in(*(keep[k][in_first] for k in (
                "lib_row", "rank", "scan_start", "scan_stop", "scan_center", "frame_start", "frame_stop", "frame_center")))

        quad = self.quadrupole_calibration.jit
        cfg_struct = self.config.to_struct(quad_sigma=quad.sigma, quad_delta_mu=quad.delta_mu)
        dev_raw = _lib.device_rawfile_for(self._dia_data, self._raw)
        dev_lib = _lib.DeviceLibrary(lib_arrays, device=dev_raw.device)
        max_frag = 1
        if len(lib_arrays["frag_start_idx"]):
            max_frag = max(1, int(np.max(lib_arrays["frag_stop_idx"].astype(np.int64) - lib_arrays["frag_start_idx"].astype(np.int64))))
        try:
            # ragged result: the feature rows of the valid candidates and their fragment slots with mz_library > 0, compacted
            # on the device = exactly what OutputPsmDF.to_precursor_df / to_fragment_df hand to the collectors below
            # (output.py:72-97).  top_k_fragments is not capped: 9999 (transfer-library requantification) keeps every fragment.
            dev_out = _lib.score_candidates_ragged(dev_raw, dev_lib, cfg_struct, cin, max_fragments=max_frag)
        finally:
            dev_lib.close()
        self.last_timing = dev_raw.last_timing()

        rows = sel[dev_out["row_index"]] if len(sel) != n_total else dev_out["row_index"]
        psm = {k: dev_out[k] for k in dev_out if k.startswith("fragment_")}
        psm["ragged"] = True
        psm["features"] = dev_out["features"]
        psm["fragment_counts"] = np.diff(dev_out["frag_offset"])
        psm["precursor_idx"] = sorted_df["precursor_idx"].values.astype(np.uint32)[rows]
        psm["rank"] = sorted_df["rank"].values.astype(np.uint8)[rows]
        lib_pidx = lib_arrays["precursor_idx"]
        if len(lib_pidx) < 2 or bool(np.all(lib_pidx[1:] > lib_pidx[:-1])):  # unique keys: merges become gathers
            psm["positions"] = dict(candidates_df=sorted_df, candidate_rows=rows, precursor_rows=keep["lib_row"][dev_out["row_index"]])

        logger.info("Finished candidate processing")
        logger.info("Collecting candidate features")
        candidate_features_df = self.collect_candidates(candidates_df, psm)
        features_schema.validate(candidate_features_df, warn_on_critical_values=True)

        logger.info("Collecting fragment features")
        fragment_features_df = self.collect_fragments(candidates_df, psm)
        fragment_features_schema.validate(fragment_features_df, warn_on_critical_values=True)

        logger.info("Finished candidate scoring")
        return candidate_features_df, fragment_features_df

"""``FragmentCompetition`` — drop-in for alphadia/fragcomp/fragcomp.py:146-299 on the B200 engine.

The table preparation keeps the reference's helper names, results and in-place side effects on ``psm_df`` / ``frag_df``
(candidate hash, fragment start/stop indices, DIA-window assignment, the ``[window_idx, proba, precursor_idx]`` sort) but is
written on run boundaries and sorted look-ups instead of group-by / merge; the greedy veto itself (``_compete_for_fragments``,
fragcomp.py:51-143) runs on the device behind ``adb_fragment_competition`` (conflict graph over RT-sorted windows).
"""

from __future__ import annotations

import logging
from dataclasses import dataclass

import numpy as np
import pandas as pd

from alphadia_b200 import _lib

logger = logging.getLogger(__name__)


def candidate_hash(precursor_idx: np.ndarray, rank: np.ndarray) -> np.ndarray:
    """alphadia/fragcomp/utils.py:48-58: precursor_idx in the lower 32 bits, rank above."""
    return (precursor_idx.astype(np.int64) + (rank.astype(np.int64) << 32)).astype(np.uint64)


def _key_extents(keys: np.ndarray):
    """``(distinct keys ascending, first row, last row + 1)`` of every key of a column - what ``groupby(key).agg(min, max)`` of
    the row number gives.  Rows of one key are normally one contiguous run; a key that comes back later spans from its first
    run to its last (the rows in between included, as with the reference's min / max)."""
    n = len(keys)
    if n == 0:
        return keys[:0], np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    cut = np.flatnonzero(keys[1:] != keys[:-1]) + 1
    run_start = np.concatenate(([0], cut)).astype(np.int64)
    run_stop = np.concatenate((cut, [n])).astype(np.int64)
    order = np.argsort(keys[run_start], kind="stable")
    run_key, run_start, run_stop = keys[run_start][order], run_start[order], run_stop[order]
    head = np.ones(len(run_key), dtype=bool)
    head[1:] = run_key[1:] != run_key[:-1]
    if head.all():
        return run_key, run_start, run_stop
    at = np.flatnonzero(head)
    return run_key[at], np.minimum.reduceat(run_start, at), np.maximum.reduceat(run_stop, at)


def add_frag_start_stop_idx(psm_df: pd.DataFrame, frag_df: pd.DataFrame) -> pd.DataFrame:
    """alphadia/fragcomp/utils.py:10-45: ``_frag_start_idx`` / ``_frag_stop_idx`` of every PSM = the extent of its
    ``_candidate_idx`` in ``frag_df`` (which gains a ``frag_idx`` column, as in the reference); PSMs without fragments are
    dropped (inner join) and the result has a fresh index."""
    if "_frag_start_idx" in psm_df.columns and "_frag_stop_idx" in psm_df.columns:
        logger.warning("Fragment start and stop indices already present in PSM dataframe. Skipping.")
        return psm_df
    frag_df["frag_idx"] = np.arange(len(frag_df))
    key, first, last = _key_extents(frag_df["_candidate_idx"].values)
    wanted = psm_df["_candidate_idx"].values
    if len(key):
        pos = np.minimum(np.searchsorted(key, wanted), len(key) - 1)
        found = key[pos] == wanted
    else:
        pos, found = np.zeros(len(wanted), dtype=np.int64), np.zeros(len(wanted), dtype=bool)
    out = psm_df[found].reset_index(drop=True)
    out["_frag_start_idx"] = first[pos[found]]
    out["_frag_stop_idx"] = last[pos[found]]
    return out


@dataclass
class FragcompPlan:
    psm_df: pd.DataFrame  # sorted by [window_idx, proba, precursor_idx]
    window_start: np.ndarray
    window_stop: np.ndarray
    rt: np.ndarray
    frag_start: np.ndarray
    frag_stop: np.ndarray
    fragment_mz: np.ndarray


class FragmentCompetition:
    """Remove PSMs that share fragments with better PSMs."""

    def __init__(self, rt_tol_seconds: int = 3, mass_tol_ppm: int = 15, thread_count: int = 8):
        self.rt_tol_seconds = rt_tol_seconds
        self.mass_tol_ppm = mass_tol_ppm
        self.thread_count = thread_count  # accepted for API compatibility; parallelism is the device's

    @staticmethod
    def _add_window_idx(psm_df: pd.DataFrame, cycle: np.ndarray) -> pd.DataFrame:
        """fragcomp.py:170-202: the first cycle position whose quadrupole range [min lower, max upper) over the scans holds
        ``mz_observed``; position 0 when none does.  One pass per position instead of an n x positions matrix."""
        if "window_idx" in psm_df.columns:
            logger.warning("Window index already present in PSM dataframe. Skipping.")
            return psm_df
        lower = cycle[0, :, :, 0].min(axis=1)
        upper = cycle[0, :, :, 1].max(axis=1)
        mz = psm_df["mz_observed"].values
        window = np.zeros(len(mz), dtype=np.int64)
        for position in range(len(lower) - 1, -1, -1):  # earlier positions overwrite later ones: the first match wins
            window[(mz >= lower[position]) & (mz < upper[position])] = position
        psm_df["window_idx"] = window
        return psm_df

    @staticmethod
    def _get_thread_plan_df(psm_df: pd.DataFrame) -> pd.DataFrame:
        """fragcomp.py:204-229: one row per DIA window with the extent ``[start_idx, stop_idx)`` of its PSMs in the (sorted)
        table."""
        window, start, stop = _key_extents(psm_df["window_idx"].values)
        return pd.DataFrame({"window_idx": window, "start_idx": start, "stop_idx": stop})

    def plan(self, psm_df: pd.DataFrame, frag_df: pd.DataFrame, cycle: np.ndarray) -> FragcompPlan:
        """Everything of ``__call__`` up to the kernel launch (fragcomp.py:254-273)."""
        psm_df["_candidate_idx"] = candidate_hash(psm_df["precursor_idx"].values, psm_df["rank"].values)
        frag_df["_candidate_idx"] = candidate_hash(frag_df["precursor_idx"].values, frag_df["rank"].values)
        psm_df = add_frag_start_stop_idx(psm_df, frag_df)
        psm_df = self._add_window_idx(psm_df, cycle)
        psm_df.sort_values(by=["window_idx", "proba", "precursor_idx"], inplace=True)
        thread_plan_df = self._get_thread_plan_df(psm_df)
        return FragcompPlan(
            psm_df=psm_df,
            window_start=thread_plan_df["start_idx"].values.astype(np.int64),
            window_stop=thread_plan_df["stop_idx"].values.astype(np.int64),
            rt=psm_df["rt_observed"].values,
            frag_start=psm_df["_frag_start_idx"].values.astype(np.int64),
            frag_stop=psm_df["_frag_stop_idx"].values.astype(np.int64),
            fragment_mz=frag_df["mz_observed"].values,
        )

    def __call__(self, psm_df: pd.DataFrame, frag_df: pd.DataFrame, cycle: np.ndarray) -> pd.DataFrame:
        plan = self.plan(psm_df, frag_df, cycle)
        valid = _lib.fragment_competition(
            plan.window_start, plan.window_stop, plan.rt, plan.frag_start, plan.frag_stop, plan.fragment_mz,
            float(self.rt_tol_seconds), float(self.mass_tol_ppm),
        )
        psm_df = plan.psm_df
        psm_df["valid"] = valid
        psm_df.drop(columns=["_frag_start_idx", "_frag_stop_idx", "window_idx"], inplace=True)
        return psm_df[psm_df["valid"]]


def plan_fragment_competition(psm_df: pd.DataFrame, frag_df: pd.DataFrame, cycle: np.ndarray) -> FragcompPlan:
    return FragmentCompetition().plan(psm_df, frag_df, cycle)

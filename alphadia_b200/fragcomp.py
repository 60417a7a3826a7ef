"""``FragmentCompetition`` — drop-in for alphadia/fragcomp/fragcomp.py:146-299 on the B200 engine.

The pandas preparation (candidate hash, fragment start/stop indices, DIA-window assignment, the
``[window_idx, proba, precursor_idx]`` sort) follows the reference line by line, including its
in-place side effects on ``psm_df`` / ``frag_df``; the greedy veto itself
(``_compete_for_fragments``, fragcomp.py:51-143) runs as one CUDA launch (one CTA per DIA window)
behind ``adb_fragment_competition``.
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


def add_frag_start_stop_idx(psm_df: pd.DataFrame, frag_df: pd.DataFrame) -> pd.DataFrame:
    """alphadia/fragcomp/utils.py:10-45."""
    if "_frag_start_idx" in psm_df.columns and "_frag_stop_idx" in psm_df.columns:
        logger.warning("Fragment start and stop indices already present in PSM dataframe. Skipping.")
        return psm_df
    frag_df["frag_idx"] = np.arange(len(frag_df))
    index_df = frag_df.groupby("_candidate_idx", as_index=False).agg(
        _frag_start_idx=pd.NamedAgg("frag_idx", "min"),
        _frag_stop_idx=pd.NamedAgg("frag_idx", "max"),
    )
    index_df["_frag_stop_idx"] += 1
    return psm_df.merge(index_df, "inner", on="_candidate_idx")


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
        """fragcomp.py:170-202."""
        if "window_idx" in psm_df.columns:
            logger.warning("Window index already present in PSM dataframe. Skipping.")
            return psm_df
        lower_limit = np.min(cycle[0, :, :, 0], axis=1, keepdims=True).T
        upper_limit = np.max(cycle[0, :, :, 1], axis=1, keepdims=True).T
        mz = np.expand_dims(psm_df["mz_observed"].values, axis=-1)
        idx = (mz >= lower_limit) & (mz < upper_limit)
        psm_df["window_idx"] = np.argmax(idx, axis=1)
        return psm_df

    @staticmethod
    def _get_thread_plan_df(psm_df: pd.DataFrame) -> pd.DataFrame:
        """fragcomp.py:204-229."""
        psm_df["_thread_idx"] = np.arange(len(psm_df))
        index_df = psm_df.groupby("window_idx", as_index=False).agg(
            start_idx=pd.NamedAgg("_thread_idx", "min"),
            stop_idx=pd.NamedAgg("_thread_idx", "max"),
        )
        index_df["stop_idx"] += 1
        psm_df.drop(columns=["_thread_idx"], inplace=True)
        return index_df

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

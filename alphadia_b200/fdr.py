"""FDR bookkeeping around fragment competition — drop-ins for ``get_q_values`` and ``keep_best``
(alphadia/fdr/fdr.py:195-297; both are called by ``perform_fdr``, fdr.py:157-186, once before and once after
``FragmentCompetition``).  SURVEY 8f.2; the classifier itself is not part of this package.

Same signatures, same returned DataFrames (row order, index, columns).  The multi-column stable sorts and the scans run on
the device behind ``adb_q_values`` / ``adb_keep_best``; the host only packs the key columns.  No CPU fallback.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from alphadia_b200 import _lib


def _integer_columns(df: pd.DataFrame, columns) -> list | None:
    cols = []
    for c in columns:
        v = df[c].to_numpy()
        if v.dtype.kind == "b":
            v = v.astype(np.uint8)
        if v.dtype.kind not in "iu":
            return None
        cols.append(v)
    return cols


def _pack_order_preserving(df: pd.DataFrame, columns) -> np.ndarray:
    """uint64 key (< 2^63) whose order is the lexicographic order of the given non-negative integer columns."""
    if len(columns) == 0:
        return np.zeros(len(df), dtype=np.uint64)
    cols = _integer_columns(df, columns)
    if cols is None:
        raise NotImplementedError(f"sort columns {list(columns)} must be integer columns")
    widths = []
    for v in cols:
        if len(v) and int(v.min()) < 0:
            raise NotImplementedError(f"sort columns {list(columns)} must be non-negative")
        widths.append(max(int(v.max()).bit_length() if len(v) else 1, 1))
    if sum(widths) > 63:
        raise NotImplementedError(f"sort columns {list(columns)} need {sum(widths)} key bits, 63 are available")
    key = np.zeros(len(df), dtype=np.uint64)
    for v, w in zip(cols, widths):
        key = (key << np.uint64(w)) | v.astype(np.uint64)
    return key


def _pack_groups(df: pd.DataFrame, columns) -> np.ndarray:
    """uint64 key that is equal exactly for rows of the same group."""
    cols = _integer_columns(df, columns)
    if cols is not None and all(len(v) == 0 or (int(v.min()) >= 0) for v in cols):
        widths = [max(int(v.max()).bit_length() if len(v) else 1, 1) for v in cols]
        if sum(widths) <= 64:
            key = np.zeros(len(df), dtype=np.uint64)
            for v, w in zip(cols, widths):
                key = (key << np.uint64(w % 64)) | v.astype(np.uint64)
            return key
    # any other dtype: dense group numbers (key packing only — the selection itself stays on the device)
    return df.groupby(list(columns), sort=False, dropna=False).ngroup().to_numpy().astype(np.uint64)


def get_q_values(
    df: pd.DataFrame,
    score_column: str = "proba",
    decoy_column: str = "_decoy",
    qval_column: str = "qval",
    extra_sort_columns: list[str] | None = None,
) -> pd.DataFrame:
    """fdr.py:226-297: rows sorted by ``[score, decoy, *extra_sort_columns]`` with the q-value column added."""
    if extra_sort_columns is None:
        extra_sort_columns = ["precursor_idx"]
    decoy = df[decoy_column].to_numpy()
    if len(decoy) and not np.isin(decoy, (0, 1)).all():
        raise ValueError(f"{decoy_column} must hold 0 (target) or 1 (decoy)")
    order, qval = _lib.q_values(
        df[score_column].to_numpy().astype(np.float64, copy=False),
        decoy.astype(np.uint8),
        _pack_order_preserving(df, extra_sort_columns),
    )
    out = df.iloc[order].copy()
    out[qval_column] = qval
    return out


def keep_best(df: pd.DataFrame, score_column: str = "proba", group_columns: list[str] | None = None) -> pd.DataFrame:
    """fdr.py:195-224: the best-scoring (lowest ``score_column``) row of every group, in the original row order."""
    if group_columns is None:
        group_columns = ["channel", "precursor_idx"]
    df = df.reset_index(drop=True)
    keep = _lib.keep_best(df[score_column].to_numpy().astype(np.float64, copy=False), _pack_groups(df, group_columns))
    return df[keep.astype(bool)].reset_index(drop=True)

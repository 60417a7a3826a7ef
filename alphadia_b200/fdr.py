"""FDR bookkeeping around fragment competition — drop-ins for ``get_q_values``, ``keep_best`` and the ``perform_fdr``
sequence that calls them (alphadia/fdr/fdr.py:25-297: classifier -> q-values -> fragment competition -> best per group ->
q-values).  SURVEY 8f.2; the classifier itself is not part of this package: ``perform_fdr`` takes the caller's classifier
object (``fit`` / ``predict_proba``), exactly like the reference.

Same signatures, same returned DataFrames (row order, index, columns).  The multi-column stable sorts and the scans run on
the device behind ``adb_q_values`` / ``adb_keep_best``; the host only packs the key columns.  No CPU fallback.
"""

from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from alphadia_b200 import _lib
from alphadia_b200.fragcomp import FragmentCompetition

logger = logging.getLogger()

max_dia_cycle_shape = 2  # fdr.py:19: fragment competition only for data without ion mobility


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


def _dense_codes(v: np.ndarray) -> np.ndarray:
    """Order-preserving dense codes 0 .. n_distinct-1 of one column (any sortable dtype; NaN / None sort last, as pandas
    ``sort_values`` places them)."""
    if v.dtype.kind in "OUS":
        s = pd.Series(v)
        codes = s.rank(method="dense", na_option="bottom").to_numpy()
        return (codes - 1).astype(np.uint64)
    if v.dtype.kind == "f":
        nan = np.isnan(v)
        uniq, inv = np.unique(v[~nan], return_inverse=True)
        codes = np.full(len(v), len(uniq), dtype=np.uint64)
        codes[~nan] = inv.astype(np.uint64)
        return codes
    _, inv = np.unique(v, return_inverse=True)
    return inv.astype(np.uint64)


def _pack_order_preserving(df: pd.DataFrame, columns) -> np.ndarray:
    """uint64 key (< 2^63) whose order is the lexicographic order of the given columns.  Non-negative integer columns are
    packed as they are; any other column (strings such as the protein group of outputtransform/protein_fdr.py:72, negative
    or float values) or a key that would not fit is first mapped to order-preserving dense codes on the host; if even the
    codes need more than 63 bits the rows are ranked by one host lexsort.  Key packing only: the sort that produces the
    q-values stays on the device."""
    if len(columns) == 0:
        return np.zeros(len(df), dtype=np.uint64)
    cols = []
    for c in columns:
        v = df[c].to_numpy()
        if v.dtype.kind == "b":
            v = v.astype(np.uint8)
        if v.dtype.kind not in "iu" or (len(v) and int(v.min()) < 0):
            v = _dense_codes(v)
        cols.append(v)
    widths = [max(int(v.max()).bit_length() if len(v) else 1, 1) for v in cols]
    if sum(widths) > 63:
        cols = [_dense_codes(v) for v in cols]
        widths = [max(int(v.max()).bit_length() if len(v) else 1, 1) for v in cols]
    if sum(widths) > 63:  # rank of the row's key tuple among the distinct tuples
        order = np.lexsort(tuple(reversed(cols)))
        stacked = np.stack([v[order] for v in cols], axis=1)
        new_group = np.ones(len(order), dtype=bool)
        new_group[1:] = (stacked[1:] != stacked[:-1]).any(axis=1)
        key = np.empty(len(order), dtype=np.uint64)
        key[order] = (np.cumsum(new_group) - 1).astype(np.uint64)
        return key
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
    """fdr.py:226-297: rows sorted by ``[score, decoy, *extra_sort_columns]`` with the q-value column added.

    NaN scores sort last, as ``sort_values`` places them: the device ranks the rows with a score; the NaN rows follow in
    ``[decoy, *extra_sort_columns]`` order, their FDR values continue the two running counts, and the running minimum from the
    back (fdr.py:227-246) carries the smallest FDR of that tail into the q-values in front of it.  A decoy value other than
    0 / 1 raises."""
    if extra_sort_columns is None:
        extra_sort_columns = ["precursor_idx"]
    decoy = df[decoy_column].to_numpy()
    if len(decoy) and not np.isin(decoy, (0, 1)).all():
        raise ValueError(f"{decoy_column} must hold 0 (target) or 1 (decoy)")
    score = df[score_column].to_numpy().astype(np.float64, copy=False)
    decoy8 = decoy.astype(np.uint8)
    key = _pack_order_preserving(df, extra_sort_columns)
    missing = np.isnan(score)
    if not missing.any():
        order, qval = _lib.q_values(score, decoy8, key)
    else:
        have = np.flatnonzero(~missing)
        order_f, q_f = _lib.q_values(score[have], decoy8[have], key[have])
        tail = np.flatnonzero(missing)
        tail = tail[np.lexsort((key[tail], decoy8[tail]))]  # stable: decoy, then the tie-break key, then the row order
        d_tail = decoy8[tail].astype(np.int64)
        with np.errstate(all="ignore"):
            fdr_tail = (int(decoy8[have].sum()) + np.cumsum(d_tail)) / (int((1 - decoy8[have].astype(np.int64)).sum()) + np.cumsum(1 - d_tail))
        q_tail = np.flip(np.minimum.accumulate(np.flip(fdr_tail)))
        order = np.concatenate([have[order_f], tail])
        qval = np.concatenate([np.minimum(q_f, q_tail[0]), q_tail])
    out = df.iloc[order].copy()
    out[qval_column] = qval
    return out


def keep_best(df: pd.DataFrame, score_column: str = "proba", group_columns: list[str] | None = None) -> pd.DataFrame:
    """fdr.py:195-224: the best-scoring (lowest ``score_column``) row of every group, in the original row order.

    Missing values as pandas treats them: rows whose group key holds a NaN / None belong to no group and are dropped
    (``groupby`` default ``dropna=True``); a NaN score sorts after every number, so a group keeps a NaN row - its first -
    only if none of its rows has a score."""
    if group_columns is None:
        group_columns = ["channel", "precursor_idx"]
    df = df.reset_index(drop=True)
    in_group = ~df[list(group_columns)].isna().any(axis=1).to_numpy() if len(group_columns) else np.ones(len(df), dtype=bool)
    if not in_group.all():
        df_in = df[in_group]
    else:
        df_in = df
    score = df_in[score_column].to_numpy().astype(np.float64, copy=False)
    groups = _pack_groups(df_in, group_columns)
    missing = np.isnan(score)
    if not missing.any():
        keep = _lib.keep_best(score, groups).astype(bool)
    else:
        keep = np.zeros(len(df_in), dtype=bool)
        have = np.flatnonzero(~missing)
        keep[have] = _lib.keep_best(score[have], groups[have]).astype(bool)
        rest = np.flatnonzero(missing)
        rest = rest[~np.isin(groups[rest], groups[have])]  # groups without any score: their first row
        _, first = np.unique(groups[rest], return_index=True)
        keep[rest[first]] = True
    return df_in[keep].reset_index(drop=True)


def _drop_incomplete(df: pd.DataFrame, columns, label: str) -> None:
    """In-place removal of rows with a missing classifier feature, as the reference does to its arguments (fdr.py:84-98)."""
    before = len(df)
    df.dropna(subset=columns, inplace=True)
    if len(df) < before:
        logger.warning(f"{before - len(df)} {label} PSMs lack at least one feature and were dropped")


def perform_fdr(
    classifier,
    available_columns: list[str],
    df_target: pd.DataFrame,
    df_decoy: pd.DataFrame,
    *,
    competitive: bool = False,
    group_channels: bool = True,
    figure_path: str | None = None,
    df_fragments: pd.DataFrame | None = None,
    dia_cycle: np.ndarray | None = None,
    fdr_heuristic: float = 0.1,
    random_state: int | None = None,
) -> pd.DataFrame:
    """Signature, flow and result of the reference's ``perform_fdr`` (fdr.py:25-192).  The classifier is the caller's object
    (``fit(X, y)``, ``predict_proba(X)``) trained on the same 80 % split; the two q-value passes, the fragment competition on
    the rows below ``fdr_heuristic`` (data without ion mobility only) and the best-row-per-group selection run on the device."""
    from sklearn.model_selection import train_test_split

    _drop_incomplete(df_target, available_columns, "target")
    _drop_incomplete(df_decoy, available_columns, "decoy")
    n_t, n_d = len(df_target), len(df_decoy)
    if abs(n_t - n_d) / ((n_t + n_d) / 2) > 0.1:
        logger.warning(f"{n_t} target vs {n_d} decoy PSMs differ by more than 10 %: the FDR estimate may be off")

    features = np.concatenate([df_target[available_columns].to_numpy(), df_decoy[available_columns].to_numpy()])
    is_decoy = np.concatenate([np.zeros(n_t), np.ones(n_d)])
    psm_df = pd.concat([df_target, df_decoy])
    try:  # fdr/utils.py:16-52: the split raises ValueError when there are too few rows
        train_x, _, train_y, _ = train_test_split(features, is_decoy, test_size=0.2, random_state=random_state)
    except ValueError:
        logger.warning("not enough PSMs to train the classifier: every PSM gets qval = proba = 1")
        psm_df["qval"] = 1.0
        psm_df["proba"] = 1.0
        return psm_df
    classifier.fit(train_x, train_y)

    psm_df["_decoy"] = is_decoy
    psm_df["proba"] = classifier.predict_proba(features)[:, 1]
    # the reference sorts by (proba, precursor_idx) first; the stable (proba, _decoy, precursor_idx) sort inside
    # get_q_values gives the same order without it
    psm_df = get_q_values(psm_df, "proba", "_decoy")

    if dia_cycle is not None and dia_cycle.shape[2] <= max_dia_cycle_shape:
        n_head = int(psm_df["qval"].searchsorted(fdr_heuristic, side="left")) or len(psm_df)
        if df_fragments is not None:
            psm_df = FragmentCompetition()(psm_df.iloc[:n_head], df_fragments, dia_cycle)

    if not competitive:
        groups = ["precursor_idx"]
    elif group_channels:
        groups = ["elution_group_idx", "channel"]
    else:
        groups = ["elution_group_idx"]
    psm_df = get_q_values(keep_best(psm_df, group_columns=groups), "proba", "_decoy")
    if figure_path is not None:
        logger.info("figure_path is ignored: the diagnostic plots belong to the reference (alphadia/fdr/plotting.py)")
    return psm_df

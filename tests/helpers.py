"""Shared test helpers: named synthetic workloads, default configs, golden loading, comparisons."""

from __future__ import annotations

import functools
import hashlib
import os

import numpy as np

from alphadia_b200 import _abi
from alphadia_b200.config import CandidateScoringConfig, CandidateSelectionConfig
from alphadia_b200.kernel import GaussianKernel
from alphadia_b200.library import assemble_library_arrays
from alphadia_b200.synthetic import CONFIGS_4D, make_config_3d, make_config_4d

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# ClassicExtractionHandler parameters (reference extraction_handler.py:349-409)
SELECTION_BASE = {
    "peak_len_rt": 10.0, "sigma_scale_rt": 0.5, "peak_len_mobility": 0.01, "sigma_scale_mobility": 1.0,
    "top_k_precursors": 3, "kernel_size": 30, "f_mobility": 1.0, "f_rt": 0.99, "center_fraction": 0.5,
    "min_size_mobility": 8, "min_size_rt": 3, "max_size_mobility": 20, "max_size_rt": 15,
    "group_channels": False, "use_weighted_score": True, "join_close_candidates": False,
    "join_close_candidates_scan_threshold": 0.6, "join_close_candidates_cycle_threshold": 0.6,
    "top_k_fragments": 12, "exclude_shared_ions": True,
}
SCORING_BASE = {
    "score_grouped": False, "top_k_isotopes": 3, "reference_channel": -1,
    "precursor_mz_tolerance": 10, "fragment_mz_tolerance": 15, "exclude_shared_ions": True,
    "quant_window": 3, "quant_all": True, "experimental_xic": True, "top_k_fragments": 12,
}


# configuration variants pinned against the live reference in tests/golden/variants.npz (generate_golden.py variants)
SELECTION_VARIANTS = {
    "parity_small": {
        "c1": dict(candidate_count=1),
        "join": dict(candidate_count=5, join_close_candidates=True, join_close_candidates_scan_threshold=0.01),
        "unweighted": dict(use_weighted_score=False),
        "rt8": dict(rt_tolerance=8.0),
        "rt400": dict(rt_tolerance=400.0),
        "k6": dict(top_k_fragments=6, top_k_precursors=2),
    },
    "parity_4d": {
        "c1": dict(candidate_count=1),
        "join": dict(candidate_count=5, join_close_candidates=True, join_close_candidates_scan_threshold=0.01),
        "unweighted": dict(use_weighted_score=False),
        "rt5": dict(rt_tolerance=5.0),
        "mob02": dict(mobility_tolerance=0.2),
        "wide": dict(rt_tolerance=100.0, mobility_tolerance=0.4),
    },
}


def multiplexed_library(precursor_df):
    """Three channels (0 / 4 / 8) per elution group, every fifth group without channel 0 (it carries 12 instead): the
    shape MultiplexingRequantificationHandler scores with score_grouped=True, reference_channel=0."""
    pdf = precursor_df.copy()
    n = len(pdf)
    group = np.arange(n) // 3
    pdf["elution_group_idx"] = group.astype(np.uint32)
    pdf["decoy"] = (group % 2).astype(np.uint8)
    ch = np.array([0, 4, 8], dtype=np.uint32)[np.arange(n) % 3]
    no_ref = group % 5 == 0
    ch[no_ref] = np.array([12, 4, 8], dtype=np.uint32)[np.arange(n) % 3][no_ref]
    pdf["channel"] = ch
    return pdf


def selection_config(rt_tolerance: float, **kw) -> CandidateSelectionConfig:
    c = CandidateSelectionConfig()
    c.update({**SELECTION_BASE, "rt_tolerance": float(rt_tolerance), "mobility_tolerance": 0.1, "candidate_count": 3,
              "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0, **kw})
    return c


def scoring_config(**kw) -> CandidateScoringConfig:
    c = CandidateScoringConfig()
    c.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, **kw})
    return c


@functools.lru_cache(maxsize=4)
def workload(name: str):
    if name in CONFIGS_4D:
        raw, precursor_df, fragment_df, p = make_config_4d(name)
    else:
        raw, precursor_df, fragment_df, p = make_config_3d(name)
    lib = assemble_library_arrays(precursor_df, fragment_df, "rt_library", "mobility_library", "mz_library", "mz_library")
    return raw, precursor_df, fragment_df, lib, p


def input_checksum(raw, precursor_df, fragment_df) -> str:
    h = hashlib.sha256()
    index = raw.tof_indptr if hasattr(raw, "tof_indptr") else raw.peak_start_idx_list
    extra = (raw.push_indices,) if hasattr(raw, "push_indices") else ()
    for a in (raw.mz_values, raw.intensity_values, index, raw.rt_values, *extra,
              precursor_df["mz_library"].values, precursor_df["rt_library"].values,
              fragment_df["mz_library"].values, fragment_df["intensity"].values):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def load_golden(name: str):
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    if not os.path.exists(path):
        return None
    return np.load(path, allow_pickle=False)


def default_kernel(raw, fwhm_rt=5.0, fwhm_mobility=0.01) -> np.ndarray:
    """GaussianKernel as CandidateSelection builds it (selection.py:609-620)."""
    g = GaussianKernel(raw, fwhm_rt=fwhm_rt, sigma_scale_rt=0.5, fwhm_mobility=fwhm_mobility, sigma_scale_mobility=1.0,
                       kernel_width=30, kernel_height=min(30, raw.scan_max_index + 1))
    return g.get_dense_matrix(verbose=False)


def candidates_in_from_arrays(lib, cand: dict):
    """Build adb_candidates_in from candidate arrays (precursor_idx, rank, scan_*, frame_*), in scoring order."""
    order = np.lexsort((cand["rank"], cand["precursor_idx"]))  # == sort by (elution_group=precursor_idx, decoy, rank)
    pidx = np.asarray(cand["precursor_idx"])[order]
    lib_row = np.searchsorted(lib["precursor_idx"], pidx)
    d, keep = _abi.make_candidates_in(
        lib_row, np.asarray(cand["rank"])[order],
        np.asarray(cand["scan_start"])[order], np.asarray(cand["scan_stop"])[order], np.asarray(cand["scan_center"])[order],
        np.asarray(cand["frame_start"])[order], np.asarray(cand["frame_stop"])[order], np.asarray(cand["frame_center"])[order],
    )
    keep["precursor_idx"] = pidx
    keep["order"] = order
    return d, keep


def ragged_from_dense(dense: dict) -> dict:
    """The result dict of ``_lib.score_candidates_ragged`` from dense ``[n, top_k]`` score tables (test stubs: the oracle only
    produces dense tables): rows with ``valid``, slots with ``mz_library > 0`` (output.py:72-97)."""
    v = dense["valid"].astype(bool)
    m = (dense["fragment_mz_library"] > 0) & v[:, None]
    out = dict(n_rows=int(v.sum()), n_fragments=int(m.sum()), row_index=np.nonzero(v)[0].astype(np.int64),
               features=dense["features"][v], frag_offset=np.concatenate([[0], np.cumsum(m.sum(axis=1)[v])]).astype(np.int64))
    for k in dense:
        if k.startswith("fragment_"):
            out[k] = dense[k][m]
    return out


def ragged_scoring_stub(score):
    """``_lib.score_candidates_ragged`` replacement for CPU tests: ``score(raw_arrays, lib_arrays, cfg, cin)`` is an oracle
    scoring function; top_k_fragments is capped at the library's widest precursor (a candidate cannot keep more)."""
    import copy

    def stub(dev_raw, dev_lib, cfg, cin, bufs=None, max_fragments=None):
        cfg2 = copy.copy(cfg)
        if max_fragments is not None:
            cfg2.top_k_fragments = max(1, min(int(cfg.top_k_fragments), int(max_fragments)))
        return ragged_from_dense(score(dev_raw.arrays, dev_lib.arrays, cfg2, cin))

    return stub


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    d = np.where(both_nan, 0.0, d)
    d = np.where(np.isnan(d), np.inf, d)
    return d


def fdr_inputs(seed: int = 11, n: int = 6000):
    """PSM table for the FDR bookkeeping tests (tests/golden/fdr_small.npz): probabilities with many exact ties, several
    ranks per precursor, three channels, a shuffled non-default index, and the best scores held by decoys (so that the
    first FDR values are x / 0)."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    precursor_idx = rng.integers(0, n // 3, n).astype(np.uint32)
    proba = np.round(rng.random(n), 2)  # 101 distinct values: ties everywhere
    decoy = (rng.random(n) < 0.45).astype(np.uint8)
    best = np.argsort(proba, kind="stable")[:3]
    proba[best] = [-1.0, -0.5, -0.0]  # negative scores and a signed zero sort first
    decoy[best] = 1
    decoy[proba == 0.0] = 1
    df = pd.DataFrame({
        "precursor_idx": precursor_idx, "rank": rng.integers(0, 3, n).astype(np.uint8),
        "elution_group_idx": (precursor_idx // 2).astype(np.uint32), "channel": (4 * rng.integers(0, 3, n)).astype(np.uint32),
        "proba": proba, "_decoy": decoy.astype(np.float64), "row": np.arange(n),
    })
    df.index = rng.permutation(n) + 100
    return df


def tied_fragment_library(name: str = "parity_f20", seed: int = 31):
    """``name``'s library with exact ties inside every precursor's fragment list (tests/golden/ties_f20.npz): three pairs of
    fragments share their m/z (different type / position / number), and the library intensities are rounded to one decimal, so
    that top-k by intensity has ties at its boundary.  With more than 15 fragments numba's argsort is a quicksort that is not
    stable - the order of the tied fragments follows its partition steps."""
    raw, pdf, fdf, lib, p = workload(name)
    rng = np.random.default_rng(seed)
    fdf = fdf.copy()
    mz = fdf["mz_library"].to_numpy().copy()
    inten = np.round(fdf["intensity"].to_numpy().astype(np.float64), 1).astype(np.float32)
    inten[inten == 0] = 0.1
    for s, e in zip(pdf["flat_frag_start_idx"].to_numpy(), pdf["flat_frag_stop_idx"].to_numpy()):
        idx = rng.permutation(np.arange(int(s), int(e)))
        for a, b in zip(idx[0:6:2], idx[1:6:2]):
            mz[b] = mz[a]
    fdf["mz_library"] = mz
    fdf["intensity"] = inten
    return raw, pdf.copy(), fdf, p


def fdr_inputs_nan(seed: int = 12, n: int = 3000):
    """``fdr_inputs`` with missing values (tests/golden/fdr_nan.npz): NaN probabilities scattered over the table, every row
    of a few precursors NaN (a group without any finite score), +inf next to NaN inside one group, and a float group
    column ``gnan`` with missing keys (pandas drops such rows from a groupby)."""
    df = fdr_inputs(seed, n)
    rng = np.random.default_rng(seed + 1)
    proba = df["proba"].to_numpy().copy()
    pidx = df["precursor_idx"].to_numpy()
    proba[rng.random(n) < 0.08] = np.nan
    for p in np.unique(pidx)[:5]:  # groups that hold NaN only
        proba[pidx == p] = np.nan
    rows = np.flatnonzero(pidx == np.unique(pidx)[7])
    if len(rows) >= 2:  # NaN first, +inf later in the same group: the +inf row wins (NaN sorts after every number)
        proba[rows[0]], proba[rows[1:]] = np.nan, np.inf
    df["proba"] = proba
    g = (pidx % 50).astype(np.float64)
    g[rng.random(n) < 0.1] = np.nan
    df["gnan"] = g
    return df


class PseudoClassifier:
    """Deterministic stand-in for the FDR classifier (fit is a no-op): probability of being a decoy from two correlation
    features, rounded to two decimals so that ties occur.  Used identically by the golden generator (reference
    perform_fdr) and by the tests (alphadia_b200.fdr.perform_fdr)."""

    def fit(self, X, y):
        self.fitted_rows = len(X)

    def predict_proba(self, X):
        z = np.clip(0.5 * (X[:, 0] + X[:, 1]), 0.0, 1.0)
        p = np.round(1.0 - z, 2)
        return np.stack([1.0 - p, p], axis=1)


FDR_FEATURE_COLUMNS = ["intensity_correlation", "top3_frame_correlation", "mean_observation_score"]


def perform_fdr_inputs(g):
    """(df_target, df_decoy, df_fragments) from the scoring golden of parity_small; elution groups pair precursor 2k (target)
    with 2k + 1 (decoy) so that the competitive mode has something to decide."""
    import pandas as pd

    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS

    psm = pd.DataFrame(g["feat_matrix"], columns=DEFAULT_FEATURE_COLUMNS)
    psm["precursor_idx"] = g["feat_precursor_idx"]
    psm["rank"] = g["feat_rank"]
    psm["decoy"] = g["feat_decoy"]
    psm["elution_group_idx"] = (g["feat_precursor_idx"] // 2).astype(np.uint32)
    psm["channel"] = np.zeros(len(psm), dtype=np.uint32)
    frag = pd.DataFrame({"precursor_idx": g["frag_precursor_idx"], "rank": g["frag_rank"], "mz_observed": g["frag_mz_observed"]})
    return psm[psm["decoy"] == 0].copy(), psm[psm["decoy"] == 1].copy(), frag


PERFORM_FDR_CASES = {
    "competitive_fragments": dict(competitive=True, group_channels=True, fragments=True),
    "competitive_no_channels": dict(competitive=True, group_channels=False, fragments=True),
    "plain": dict(competitive=False, group_channels=True, fragments=False),
}


def ragged_library_frames(precursor_df, fragment_df, rt_max: float, seed: int = 17):
    """Library shapes the reference meets in practice, as DataFrames (tests/golden/ragged.npz): 0-12 fragments for every
    fifth precursor, shared ions (cardinality 2), duplicate and nearly identical fragment m/z, retention times before and
    after the run, charges 1-4."""
    rng = np.random.default_rng(seed)
    pdf, fdf = precursor_df.copy(), fragment_df.copy()
    P = len(pdf)
    start = pdf["flat_frag_start_idx"].values.astype(np.int64)
    stop = pdf["flat_frag_stop_idx"].values.copy()
    for i in range(0, P, 5):
        stop[i] = start[i] + int(rng.integers(0, 13))
    pdf["flat_frag_stop_idx"] = stop
    card = fdf["cardinality"].values.copy()
    card[rng.random(len(card)) < 0.15] = 2
    fdf["cardinality"] = card
    mz = fdf["mz_library"].values.copy()
    for i in range(1, P, 7):
        s = start[i]
        mz[s + 1] = mz[s]
        mz[s + 3] = np.float32(mz[s + 2] * (1 + 6e-6))
    fdf["mz_library"] = mz
    rt = pdf["rt_library"].values.copy()
    rt[::11] = np.float32(-50.0)
    rt[5::13] = np.float32(rt_max + 500.0)
    pdf["rt_library"] = rt
    pdf["charge"] = rng.integers(1, 5, size=P).astype(np.uint8)
    return pdf, fdf


def edge_candidate_frame(name: str, seed: int = 5):
    """Hand-made candidate windows for the scoring edge-case golden (tests/golden/edge.npz): widths from 1 to 241 cycles,
    windows clipped at the first / last cycle of the run, for timsTOF files also 1-scan to full-height scan windows.
    One candidate per precursor, so (precursor_idx, rank) stays unique."""
    import pandas as pd

    raw, pdf, fdf, lib, p = workload(name)
    rng = np.random.default_rng(seed)
    P = len(pdf)
    n = min(400, P)
    rows = np.sort(rng.permutation(P)[:n])
    ncyc = int(raw.precursor_cycle_max_index)
    centers = rng.integers(0, ncyc, n)
    half = rng.choice([0, 1, 2, 3, 7, 14, 40, 120], n)
    c0 = np.clip(centers - half, 0, ncyc - 1)
    c1 = np.clip(centers + half + 1, 1, ncyc)
    if name in CONFIGS_4D:
        Fr, Sc, z = raw.cycle.shape[1], raw.cycle.shape[2], int(raw.zeroth_frame)
        sc_c = rng.integers(0, Sc, n)
        sh = rng.choice([0, 1, 4, 9, 20, Sc], n)
        scan_center, scan_start, scan_stop = sc_c, np.clip(sc_c - sh, 0, Sc - 1), np.clip(sc_c + sh + 1, 1, Sc)
        frame_center = np.minimum(centers * Fr + z, raw.frame_max_index - 1)
        frame_start, frame_stop = c0 * Fr + z, np.minimum(c1 * Fr + z, raw.frame_max_index)
    else:
        L = raw.cycle_len
        scan_center, scan_start, scan_stop = np.zeros(n, np.int64), np.zeros(n, np.int64), np.ones(n, np.int64)
        frame_center = np.minimum(centers * L, raw.frame_max_index)
        frame_start, frame_stop = c0 * L, np.minimum(c1 * L, raw.frame_max_index)
    return pd.DataFrame({
        "precursor_idx": pdf["precursor_idx"].values[rows].astype(np.uint32), "rank": (np.arange(n) % 3).astype(np.uint8),
        "score": np.ones(n, dtype=np.float32),
        "scan_center": np.asarray(scan_center, np.int64), "scan_start": np.asarray(scan_start, np.int64),
        "scan_stop": np.asarray(scan_stop, np.int64), "frame_center": np.asarray(frame_center, np.int64),
        "frame_start": np.asarray(frame_start, np.int64), "frame_stop": np.asarray(frame_stop, np.int64),
        "elution_group_idx": pdf["elution_group_idx"].values[rows].astype(np.uint32),
        "decoy": pdf["decoy"].values[rows].astype(np.uint8),
    })


# scoring under a fitted quadrupole model and other tolerances (tests/golden/scoring_variants.npz); candidates = the golden
# candidate table of the file
SCORING_VARIANTS_EXTRA = {
    "quad": dict(config={}, quad_sigma=(0.45, 0.8), quad_delta_mu=(0.6, -0.4)),
    "tol": dict(config={"precursor_mz_tolerance": 12, "fragment_mz_tolerance": 25, "top_k_isotopes": 2}, quad_sigma=(0.2, 0.2),
                quad_delta_mu=(0.0, 0.0)),
    "shared": dict(config={"exclude_shared_ions": False, "quant_window": 4}, quad_sigma=(0.2, 0.2), quad_delta_mu=(0.0, 0.0)),
}
SCORING_VARIANT_FILES = ("parity_small", "parity_4d_overlap")


def fragcomp_dense_inputs(dtype_rt=np.float64, dtype_mz=np.float32, seed: int = 7, n: int = 6000, nwin: int = 5,
                          rt_span: float = 40.0, p_replace: float = 0.45):
    """Many PSMs that share fragments inside few windows (tests/golden/fragcomp_dense.npz): 300 base spectra reused with a
    4 ppm jitter, 45 % of the fragments replaced at random, 4-12 fragments per PSM, all inside 40 s of retention time."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(200, 1800, size=(300, 12))
    src = rng.integers(0, 300, n)
    mz = base[src] * (1 + rng.normal(0, 4e-6, size=(n, 12)))
    replace = rng.random((n, 12)) < p_replace
    mz = np.where(replace, rng.uniform(200, 1800, size=(n, 12)), mz)
    nfr = rng.integers(4, 13, n)
    frag_start = np.concatenate([[0], np.cumsum(nfr)[:-1]]).astype(np.int64)
    frag_stop = frag_start + nfr
    frag_mz = np.concatenate([np.sort(mz[i, : nfr[i]]) for i in range(n)]).astype(dtype_mz)
    rt = rng.uniform(0, rt_span, n).astype(dtype_rt)
    bounds = np.linspace(0, n, nwin + 1).astype(np.int64)
    return bounds[:-1], bounds[1:], rt, frag_start, frag_stop, frag_mz


FRAGCOMP_DTYPES = {"f32_f32": (np.float32, np.float32), "f64_f32": (np.float64, np.float32), "f64_f64": (np.float64, np.float64)}


# further selection variants (tests/golden/variants2.npz): peak-limit parameters and the smoothing kernel's widths
# ("fwhm_rt" / "fwhm_mobility" are constructor arguments of CandidateSelection, everything else is configuration)
SELECTION_VARIANTS2 = {
    "limits": dict(f_rt=0.8, f_mobility=0.9, center_fraction=0.3, min_size_rt=5, max_size_rt=25, min_size_mobility=4,
                   max_size_mobility=12),
    "kernel": dict(fwhm_rt=9.0, fwhm_mobility=0.025, sigma_scale_rt=0.8, sigma_scale_mobility=0.6),
}


def patch_device_with_oracle(monkeypatch, oracle_lib, is4d: bool = False):
    """CPU tests of the HOST side of the operator classes: the device entry points of ``alphadia_b200._lib`` are replaced by
    the oracle for the duration of one test (there is no GPU and no CPU fallback in the product; the GPU tests run the real
    calls)."""
    from alphadia_b200 import _abi, _lib

    int_cols = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]
    state = {}

    class HostRaw:
        device = 0

        def __init__(self, arrays):
            self.arrays = arrays

        def last_timing(self):
            return {}

    class HostLibrary:
        def __init__(self, arrays, device=0):
            self.arrays = arrays

        def close(self):
            pass

    def select_resident(dev_raw, dev_lib, cfg, kernel):
        select = oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates
        arrs = select(dev_raw.arrays, dev_lib.arrays, cfg, kernel)
        keep = arrs["score"] > 0  # adb_fetch_candidate_table: rows with score > 0 in container order
        state["table"] = {c: arrs[c][keep] for c in int_cols + ["score"]}
        # library row of every container row (the container holds candidate_count rows per precursor, in library order)
        state["table"]["lib_row"] = (np.flatnonzero(keep) // int(cfg.candidate_count)).astype(np.int64)
        return int(keep.sum())

    def fetch_table(dev_raw, n, arrs=None):
        table = _abi.alloc_candidate_table(n)
        for c, v in state["table"].items():
            table[c][:n] = v
        return table

    score = oracle_lib.score_candidates_4d if is4d else oracle_lib.score_candidates
    monkeypatch.setattr(_lib, "device_rawfile_for", lambda dia_data, adapted: HostRaw(adapted))
    monkeypatch.setattr(_lib, "DeviceLibrary", HostLibrary)
    monkeypatch.setattr(_lib, "select_candidates_resident", select_resident)
    monkeypatch.setattr(_lib, "fetch_candidate_table", fetch_table)
    monkeypatch.setattr(_lib, "score_candidates", lambda dev_raw, dev_lib, cfg, cin: score(dev_raw.arrays, dev_lib.arrays, cfg, cin))
    monkeypatch.setattr(_lib, "score_candidates_ragged", ragged_scoring_stub(score))


def classifier_inputs(seed: int = 21, n: int = 12000, n_features: int = 47):
    """Feature matrix and target / decoy labels for the FDR classifier tests (tests/golden/classifier_small.npz): two
    overlapping populations with feature scales spanning six orders of magnitude, as the 46 scoring features do."""
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.5).astype(np.float64)
    scale = 10.0 ** rng.uniform(-2, 4, n_features)
    shift = rng.normal(0, 1, n_features) * (rng.random(n_features) < 0.6)
    x = (rng.normal(0, 1, (n, n_features)) + y[:, None] * shift[None, :]) * scale[None, :]
    x[:, 3] = np.round(x[:, 3])            # an integer-valued feature
    x[:, 5] = 0.0                          # a constant feature (zero variance in the batch norm)
    return x.astype(np.float32), y

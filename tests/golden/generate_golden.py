"""Generate golden vectors by running the UNMODIFIED reference numba path (CPU, this container).

    python tests/golden/generate_golden.py [config1 parity_small parity_4d transpose variants fdr perform_fdr ragged edge scoring_variants fragcomp_dense variants2 iso2 ...]

Needs /root/reference (read-only) and numba; pays ~6 min of JIT once per process.  Writes
``tests/golden/<name>.npz`` — inputs are NOT stored (they are regenerated from the seed by
``alphadia_b200.synthetic``; a checksum of the inputs is stored so the tests can tell).

Parameters mirror ``ClassicExtractionHandler`` (reference
``alphadia/workflow/peptidecentric/extraction_handler.py:349-409,423-431,460-468``).
"""

from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from alphadia_b200.synthetic import CONFIGS_4D, make_config_3d, make_config_4d  # noqa: E402
from oracle import refshim  # noqa: E402

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


def input_checksum(raw, precursor_df, fragment_df) -> str:
    h = hashlib.sha256()
    index = raw.tof_indptr if hasattr(raw, "tof_indptr") else raw.peak_start_idx_list
    extra = (raw.push_indices,) if hasattr(raw, "push_indices") else ()
    for a in (raw.mz_values, raw.intensity_values, index, raw.rt_values, *extra,
              precursor_df["mz_library"].values, precursor_df["rt_library"].values,
              fragment_df["mz_library"].values, fragment_df["intensity"].values):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def prepare_library(precursor_df, fragment_df):
    precursor_df = precursor_df.copy()
    fragment_df = fragment_df.copy()
    return precursor_df, fragment_df


def run(name: str, threads: int, variants: bool):
    is_4d = name in CONFIGS_4D
    if is_4d:
        raw, precursor_df, fragment_df, p = make_config_4d(name)
        dia = refshim.RefDiaData4D(raw)
    else:
        raw, precursor_df, fragment_df, p = make_config_3d(name)
        dia = refshim.RefDiaData(raw)
    chk = input_checksum(raw, precursor_df, fragment_df)

    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    fc_mod = refshim.ref("alphadia.fragcomp.fragcomp")

    out = {"input_checksum": np.array(chk)}

    # ---------------- selection ----------------
    sel_cfg = cfg_mod.CandidateSelectionConfig()
    sel_cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]),
                    "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)),
                    "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0})
    sel = sel_mod.CandidateSelection(
        dia, precursor_df.copy(), fragment_df.copy(), sel_cfg,
        rt_column="rt_library", mobility_column="mobility_library",
        precursor_mz_column="mz_library", fragment_mz_column="mz_library",
        fwhm_rt=5.0, fwhm_mobility=0.01,
    )
    out["sel_kernel"] = np.asarray(sel.kernel)
    t0 = time.perf_counter()
    cand = sel(thread_count=threads)
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    cand = sel(thread_count=threads)
    t_sel = time.perf_counter() - t0
    print(f"[{name}] selection: first {t_first:.1f}s warm {t_sel:.3f}s -> {len(cand)} candidates "
          f"({len(precursor_df) / t_sel:.0f} precursors/s, {threads} threads)", flush=True)
    for c in cand.columns:
        out[f"cand_{c}"] = cand[c].values

    # ---------------- scoring ------------------
    def score(cfg_updates, tag):
        sc_cfg = sccfg_mod.CandidateScoringConfig()
        sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, **cfg_updates})
        scorer = sc_mod.CandidateScoring(
            dia_data=dia, precursors_flat=precursor_df.copy(), fragments_flat=fragment_df.copy(),
            config=sc_cfg, rt_column="rt_library", mobility_column="mobility_library",
            precursor_mz_column="mz_library", fragment_mz_column="mz_library",
        )
        t0 = time.perf_counter()
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
        t1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
        t2 = time.perf_counter() - t0
        print(f"[{name}] scoring{tag}: first {t1:.1f}s warm {t2:.3f}s -> {len(feat)} rows, {len(frag)} fragment rows "
              f"({len(cand) / t2:.0f} candidates/s)", flush=True)
        out[f"feat{tag}_columns"] = np.array(list(feat.columns))
        fcols = sc_mod.DEFAULT_FEATURE_COLUMNS
        out[f"feat{tag}_matrix"] = feat[fcols].values.astype(np.float32)
        out[f"feat{tag}_precursor_idx"] = feat["precursor_idx"].values
        out[f"feat{tag}_rank"] = feat["rank"].values
        for c in ("delta_rt", "n_K", "n_R", "n_P", "score", "frame_start", "frame_stop", "decoy", "charge"):
            if c in feat.columns:
                out[f"feat{tag}_{c}"] = feat[c].values
        for c in frag.columns:
            out[f"frag{tag}_{c}"] = frag[c].values
        return feat, frag

    feat, frag = score({}, "")
    if variants:
        score({"quant_all": False, "experimental_xic": False}, "_legacy")
        score({"top_k_fragments": 6, "top_k_isotopes": 4, "quant_window": 2}, "_k6")

    if is_4d:  # fragment competition is only applied to non-mobility data (alphadia/fdr/fdr.py:19,157)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)
        return
    # ---------------- fragment competition ------------------
    psm = feat.copy()
    # deterministic pseudo-classifier output: lower proba = better (fdr.py sorts ascending)
    rng = np.random.default_rng(12345)
    psm["proba"] = rng.uniform(0, 1, size=len(psm)).astype(np.float64)
    fc = fc_mod.FragmentCompetition(rt_tol_seconds=3, mass_tol_ppm=15, thread_count=threads)
    t0 = time.perf_counter()
    kept = fc(psm.copy(), frag.copy(), raw.cycle)
    t1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    kept = fc(psm.copy(), frag.copy(), raw.cycle)
    t2 = time.perf_counter() - t0
    print(f"[{name}] fragcomp: first {t1:.1f}s warm {t2:.3f}s -> kept {len(kept)} of {len(psm)}", flush=True)
    out["fc_proba"] = psm["proba"].values
    out["fc_kept_precursor_idx"] = kept["precursor_idx"].values
    out["fc_kept_rank"] = kept["rank"].values
    out["fc_kept_candidate_idx"] = kept["_candidate_idx"].values

    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_variants(threads: int):
    """Configuration variants of the selection (3-D and 4-D) and the multiplexed scoring set-up
    (score_grouped / reference_channel, multiplexing_requantification_handler.py:120-149) -> tests/golden/variants.npz."""
    from tests.helpers import SELECTION_VARIANTS, multiplexed_library

    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    out = {}

    def select(dia, precursor_df, fragment_df, p, updates):
        cfg = cfg_mod.CandidateSelectionConfig()
        cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]),
                    "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)),
                    "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0, **updates})
        sel = sel_mod.CandidateSelection(
            dia, precursor_df.copy(), fragment_df.copy(), cfg, rt_column="rt_library", mobility_column="mobility_library",
            precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
        return sel(thread_count=threads)

    for name, variants in SELECTION_VARIANTS.items():
        if name in CONFIGS_4D:
            raw, precursor_df, fragment_df, p = make_config_4d(name)
            dia = refshim.RefDiaData4D(raw)
        else:
            raw, precursor_df, fragment_df, p = make_config_3d(name)
            dia = refshim.RefDiaData(raw)
        out[f"{name}__input_checksum"] = np.array(input_checksum(raw, precursor_df, fragment_df))
        for tag, updates in variants.items():
            t0 = time.perf_counter()
            cand = select(dia, precursor_df, fragment_df, p, updates)
            print(f"[variants] {name}/{tag}: {len(cand)} candidates in {time.perf_counter() - t0:.1f}s", flush=True)
            for c in cand.columns:
                out[f"{name}__{tag}__cand_{c}"] = cand[c].values

    # multiplexed scoring on the 3-D case
    raw, precursor_df, fragment_df, p = make_config_3d("parity_small")
    dia = refshim.RefDiaData(raw)
    mpdf = multiplexed_library(precursor_df)
    cand = select(dia, mpdf, fragment_df, p, {"candidate_count": 1})
    for c in cand.columns:
        out[f"mplex__cand_{c}"] = cand[c].values
    for tag, updates in {"ref0": {"score_grouped": True, "reference_channel": 0},
                         "grouped": {"score_grouped": True, "reference_channel": -1}}.items():
        sc_cfg = sccfg_mod.CandidateScoringConfig()
        sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, **updates})
        scorer = sc_mod.CandidateScoring(
            dia_data=dia, precursors_flat=mpdf.copy(), fragments_flat=fragment_df.copy(), config=sc_cfg,
            rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
            fragment_mz_column="mz_library")
        t0 = time.perf_counter()
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
        print(f"[variants] mplex/{tag}: {len(feat)} rows, {len(frag)} fragment rows in {time.perf_counter() - t0:.1f}s", flush=True)
        out[f"mplex__{tag}__feat_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
        for c in ("precursor_idx", "rank", "elution_group_idx", "channel", "decoy"):
            out[f"mplex__{tag}__feat_{c}"] = feat[c].values
        for c in ("precursor_idx", "rank", "mz_library", "number", "intensity", "elution_group_idx", "decoy"):
            out[f"mplex__{tag}__frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "variants.npz")
    np.savez_compressed(path, **out)
    print(f"[variants] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_fdr():
    """get_q_values / keep_best of the unmodified reference (alphadia/fdr/fdr.py:195-297) -> tests/golden/fdr_small.npz."""
    from tests.helpers import fdr_inputs

    fdr = refshim.ref("alphadia.fdr.fdr")
    df = fdr_inputs()
    out = {"input_checksum": np.array(hashlib.sha256(df.to_numpy().tobytes() + df.index.to_numpy().tobytes()).hexdigest())}
    q = fdr.get_q_values(df.copy(), "proba", "_decoy")
    out["q_row"], out["q_index"], out["q_qval"] = q["row"].values, q.index.values, q["qval"].values
    q2 = fdr.get_q_values(df.copy(), "proba", "_decoy", extra_sort_columns=["precursor_idx", "rank"])
    out["q2_row"], out["q2_qval"] = q2["row"].values, q2["qval"].values
    for tag, cols in {"precursor": ["precursor_idx"], "channel_eg": ["elution_group_idx", "channel"], "eg": ["elution_group_idx"],
                      "default": None}.items():
        kept = fdr.keep_best(df.copy(), group_columns=cols)
        out[f"keep_{tag}_row"] = kept["row"].values
        assert np.array_equal(kept.index.values, np.arange(len(kept)))
    # the sequence perform_fdr runs (fdr.py:157-186): q-values, best per group, q-values again
    final = fdr.get_q_values(fdr.keep_best(q, group_columns=["elution_group_idx", "channel"]), "proba", "_decoy")
    out["final_row"], out["final_qval"] = final["row"].values, final["qval"].values
    path = os.path.join(HERE, "fdr_small.npz")
    np.savez_compressed(path, **out)
    print(f"[fdr] {len(df)} rows -> q-values {len(q)}, final {len(final)}; wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_fdr_nan():
    """get_q_values / keep_best of the unmodified reference on a table with NaN scores and NaN group keys -> fdr_nan.npz."""
    from tests.helpers import fdr_inputs_nan

    fdr = refshim.ref("alphadia.fdr.fdr")
    df = fdr_inputs_nan()
    out = {"input_checksum": np.array(hashlib.sha256(df.to_numpy().tobytes() + df.index.to_numpy().tobytes()).hexdigest())}
    with np.errstate(all="ignore"):
        q = fdr.get_q_values(df.copy(), "proba", "_decoy")
        q2 = fdr.get_q_values(df.copy(), "proba", "_decoy", extra_sort_columns=["rank", "gnan"])
    out["q_row"], out["q_index"], out["q_qval"] = q["row"].values, q.index.values, q["qval"].values
    out["q2_row"], out["q2_qval"] = q2["row"].values, q2["qval"].values
    for tag, cols in {"precursor": ["precursor_idx"], "gnan": ["gnan"], "gnan_channel": ["gnan", "channel"]}.items():
        kept = fdr.keep_best(df.copy(), group_columns=cols)
        out[f"keep_{tag}_row"] = kept["row"].values
    path = os.path.join(HERE, "fdr_nan.npz")
    np.savez_compressed(path, **out)
    print(f"[fdr_nan] {len(df)} rows, {int(df['proba'].isna().sum())} NaN scores, {int(df['gnan'].isna().sum())} NaN group keys; "
          f"kept {[len(out[k]) for k in out if k.startswith('keep_')]}; wrote {path}", flush=True)


def run_perform_fdr():
    """perform_fdr of the unmodified reference (alphadia/fdr/fdr.py:25-192) with a deterministic stand-in classifier on the
    scoring golden of parity_small -> tests/golden/perform_fdr_small.npz."""
    from tests.helpers import FDR_FEATURE_COLUMNS, PERFORM_FDR_CASES, PseudoClassifier, perform_fdr_inputs

    fdr = refshim.ref("alphadia.fdr.fdr")
    g = np.load(os.path.join(HERE, "parity_small.npz"), allow_pickle=False)
    raw = make_config_3d("parity_small")[0]
    out = {"source_checksum": g["input_checksum"]}
    for tag, case in PERFORM_FDR_CASES.items():
        df_target, df_decoy, frag = perform_fdr_inputs(g)
        res = fdr.perform_fdr(PseudoClassifier(), FDR_FEATURE_COLUMNS, df_target, df_decoy, competitive=case["competitive"],
                              group_channels=case["group_channels"], df_fragments=frag if case["fragments"] else None,
                              dia_cycle=raw.cycle, random_state=7)
        print(f"[perform_fdr] {tag}: {len(df_target)} targets + {len(df_decoy)} decoys -> {len(res)} rows, "
              f"{int((res['qval'] <= 0.01).sum())} at 1% FDR", flush=True)
        for c in ("precursor_idx", "rank", "proba", "qval", "_decoy"):
            out[f"{tag}__{c}"] = res[c].values
        out[f"{tag}__index"] = res.index.values
        out[f"{tag}__columns"] = np.array(list(res.columns))
    path = os.path.join(HERE, "perform_fdr_small.npz")
    np.savez_compressed(path, **out)
    print(f"[perform_fdr] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_ragged(threads: int):
    """Selection and scoring of the unmodified reference on ragged libraries (tests.helpers.ragged_library_frames), 3-D and
    4-D -> tests/golden/ragged.npz."""
    from tests.helpers import ragged_library_frames

    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    out = {}
    for name in ("parity_small", "parity_4d"):
        if name in CONFIGS_4D:
            raw, precursor_df, fragment_df, p = make_config_4d(name)
            dia = refshim.RefDiaData4D(raw)
        else:
            raw, precursor_df, fragment_df, p = make_config_3d(name)
            dia = refshim.RefDiaData(raw)
        out[f"{name}__input_checksum"] = np.array(input_checksum(raw, precursor_df, fragment_df))
        pdf, fdf = ragged_library_frames(precursor_df, fragment_df, float(np.max(raw.rt_values)))
        cfg = cfg_mod.CandidateSelectionConfig()
        cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]),
                    "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)),
                    "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0})
        sel = sel_mod.CandidateSelection(
            dia, pdf.copy(), fdf.copy(), cfg, rt_column="rt_library", mobility_column="mobility_library",
            precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
        cand = sel(thread_count=threads)
        print(f"[ragged] {name}: {len(cand)} candidates", flush=True)
        for c in cand.columns:
            out[f"{name}__cand_{c}"] = cand[c].values
        sc_cfg = sccfg_mod.CandidateScoringConfig()
        sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10})
        scorer = sc_mod.CandidateScoring(
            dia_data=dia, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(), config=sc_cfg, rt_column="rt_library",
            mobility_column="mobility_library", precursor_mz_column="mz_library", fragment_mz_column="mz_library")
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
        print(f"[ragged] {name}: {len(feat)} feature rows, {len(frag)} fragment rows", flush=True)
        out[f"{name}__feat_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
        out[f"{name}__feat_precursor_idx"] = feat["precursor_idx"].values
        out[f"{name}__feat_rank"] = feat["rank"].values
        for c in frag.columns:
            out[f"{name}__frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "ragged.npz")
    np.savez_compressed(path, **out)
    print(f"[ragged] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_edge(threads: int):
    """CandidateScoring of the unmodified reference on hand-made candidate windows (tests.helpers.edge_candidate_frame),
    3-D and 4-D -> tests/golden/edge.npz."""
    from tests.helpers import edge_candidate_frame

    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    out = {}
    for name in ("parity_small", "parity_4d"):
        if name in CONFIGS_4D:
            raw, precursor_df, fragment_df, p = make_config_4d(name)
            dia = refshim.RefDiaData4D(raw)
        else:
            raw, precursor_df, fragment_df, p = make_config_3d(name)
            dia = refshim.RefDiaData(raw)
        out[f"{name}__input_checksum"] = np.array(input_checksum(raw, precursor_df, fragment_df))
        cand = edge_candidate_frame(name)
        for c in cand.columns:
            out[f"{name}__cand_{c}"] = cand[c].values
        sc_cfg = sccfg_mod.CandidateScoringConfig()
        sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10})
        scorer = sc_mod.CandidateScoring(
            dia_data=dia, precursors_flat=precursor_df.copy(), fragments_flat=fragment_df.copy(), config=sc_cfg,
            rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
            fragment_mz_column="mz_library")
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
        print(f"[edge] {name}: {len(cand)} candidates -> {len(feat)} feature rows, {len(frag)} fragment rows", flush=True)
        out[f"{name}__feat_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
        out[f"{name}__feat_precursor_idx"] = feat["precursor_idx"].values
        out[f"{name}__feat_rank"] = feat["rank"].values
        for c in frag.columns:
            out[f"{name}__frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "edge.npz")
    np.savez_compressed(path, **out)
    print(f"[edge] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_scoring_variants(threads: int):
    """CandidateScoring of the unmodified reference with a fitted quadrupole model (sigma / delta_mu of SimpleQuadrupoleJit,
    quadrupole.py:46-115) and with other tolerances, on the golden candidates of a 3-D and a two-observation 4-D file
    -> tests/golden/scoring_variants.npz."""
    from tests.helpers import SCORING_VARIANT_FILES, SCORING_VARIANTS_EXTRA

    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    quad_mod = refshim.ref("alphadia.search.scoring.quadrupole")
    out = {}
    for name in SCORING_VARIANT_FILES:
        if name in CONFIGS_4D:
            raw, precursor_df, fragment_df, p = make_config_4d(name)
            dia = refshim.RefDiaData4D(raw)
        else:
            raw, precursor_df, fragment_df, p = make_config_3d(name)
            dia = refshim.RefDiaData(raw)
        g = np.load(os.path.join(HERE, f"{name}.npz"), allow_pickle=False)
        assert str(g["input_checksum"]) == input_checksum(raw, precursor_df, fragment_df)
        out[f"{name}__input_checksum"] = g["input_checksum"]
        cand = pd.DataFrame({k[len("cand_"):]: g[k] for k in g.files if k.startswith("cand_")})
        for tag, var in SCORING_VARIANTS_EXTRA.items():
            sc_cfg = sccfg_mod.CandidateScoringConfig()
            sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, **var["config"]})
            quad = quad_mod.SimpleQuadrupole(raw.cycle)
            quad.jit.sigma[:] = var["quad_sigma"]
            quad.jit.delta_mu[:] = var["quad_delta_mu"]
            scorer = sc_mod.CandidateScoring(
                dia_data=dia, precursors_flat=precursor_df.copy(), fragments_flat=fragment_df.copy(), config=sc_cfg,
                quadrupole_calibration=quad, rt_column="rt_library", mobility_column="mobility_library",
                precursor_mz_column="mz_library", fragment_mz_column="mz_library")
            feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
            print(f"[scoring_variants] {name}/{tag}: {len(feat)} feature rows, {len(frag)} fragment rows", flush=True)
            key = f"{name}__{tag}__"
            out[key + "feat_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
            out[key + "feat_precursor_idx"] = feat["precursor_idx"].values
            out[key + "feat_rank"] = feat["rank"].values
            for c in frag.columns:
                out[key + f"frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "scoring_variants.npz")
    np.savez_compressed(path, **out)
    print(f"[scoring_variants] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_fragcomp_dense(threads: int):
    """_compete_for_fragments of the unmodified reference (fragcomp/fragcomp.py:51-143) on a collision-rich table, for the
    dtype combinations the workflow produces -> tests/golden/fragcomp_dense.npz."""
    from tests.helpers import FRAGCOMP_DTYPES, fragcomp_dense_inputs

    fc_mod = refshim.ref("alphadia.fragcomp.fragcomp")
    refshim.ref("alphatims.utils").set_threads(threads)
    out = {}
    for tag, (dtype_rt, dtype_mz) in FRAGCOMP_DTYPES.items():
        ws, we, rt, fs, fe, mz = fragcomp_dense_inputs(dtype_rt, dtype_mz)
        valid = np.ones(len(rt)).astype(bool)
        fc_mod._compete_for_fragments(np.arange(len(ws)), ws, we, rt, fs, fe, mz, 3, 15, valid)
        print(f"[fragcomp_dense] {tag}: {int(valid.sum())} of {len(valid)} PSMs keep their fragments", flush=True)
        out[f"{tag}__valid"] = valid
        out[f"{tag}__checksum"] = np.array(hashlib.sha256(rt.tobytes() + mz.tobytes()).hexdigest())
    path = os.path.join(HERE, "fragcomp_dense.npz")
    np.savez_compressed(path, **out)
    print(f"[fragcomp_dense] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_variants2(threads: int):
    """More selection variants of the unmodified reference (peak-limit parameters, kernel widths), 3-D and 4-D
    -> tests/golden/variants2.npz."""
    from tests.helpers import SELECTION_VARIANTS2

    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    out = {}
    for name in ("parity_small", "parity_4d"):
        if name in CONFIGS_4D:
            raw, precursor_df, fragment_df, p = make_config_4d(name)
            dia = refshim.RefDiaData4D(raw)
        else:
            raw, precursor_df, fragment_df, p = make_config_3d(name)
            dia = refshim.RefDiaData(raw)
        out[f"{name}__input_checksum"] = np.array(input_checksum(raw, precursor_df, fragment_df))
        for tag, updates in SELECTION_VARIANTS2.items():
            updates = dict(updates)
            fwhm_rt, fwhm_mobility = updates.pop("fwhm_rt", 5.0), updates.pop("fwhm_mobility", 0.01)
            cfg = cfg_mod.CandidateSelectionConfig()
            cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]),
                        "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)),
                        "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0, **updates})
            sel = sel_mod.CandidateSelection(
                dia, precursor_df.copy(), fragment_df.copy(), cfg, rt_column="rt_library", mobility_column="mobility_library",
                precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=fwhm_rt, fwhm_mobility=fwhm_mobility)
            cand = sel(thread_count=threads)
            print(f"[variants2] {name}/{tag}: {len(cand)} candidates", flush=True)
            out[f"{name}__{tag}__kernel"] = np.asarray(sel.kernel)
            for c in cand.columns:
                out[f"{name}__{tag}__cand_{c}"] = cand[c].values
    path = os.path.join(HERE, "variants2.npz")
    np.savez_compressed(path, **out)
    print(f"[variants2] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_iso2(threads: int):
    """Library with only two isotope columns (i_0, i_1) while top_k_precursors = top_k_isotopes = 3: selection and scoring of
    the unmodified reference on parity_small -> tests/golden/iso2.npz."""
    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    name = "parity_small"
    raw, precursor_df, fragment_df, p = make_config_3d(name)
    dia = refshim.RefDiaData(raw)
    out = {"input_checksum": np.array(input_checksum(raw, precursor_df, fragment_df))}
    pdf = precursor_df.drop(columns=["i_2", "i_3"])
    cfg = cfg_mod.CandidateSelectionConfig()
    cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]), "mobility_tolerance": 0.1,
                "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0})
    sel = sel_mod.CandidateSelection(
        dia, pdf.copy(), fragment_df.copy(), cfg, rt_column="rt_library", mobility_column="mobility_library",
        precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
    cand = sel(thread_count=threads)
    print(f"[iso2] {len(cand)} candidates", flush=True)
    for c in cand.columns:
        out[f"cand_{c}"] = cand[c].values
    sc_cfg = sccfg_mod.CandidateScoringConfig()
    sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10})
    scorer = sc_mod.CandidateScoring(
        dia_data=dia, precursors_flat=pdf.copy(), fragments_flat=fragment_df.copy(), config=sc_cfg, rt_column="rt_library",
        mobility_column="mobility_library", precursor_mz_column="mz_library", fragment_mz_column="mz_library")
    feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
    print(f"[iso2] {len(feat)} feature rows, {len(frag)} fragment rows", flush=True)
    out["feat_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
    out["feat_precursor_idx"] = feat["precursor_idx"].values
    out["feat_rank"] = feat["rank"].values
    for c in frag.columns:
        out[f"frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "iso2.npz")
    np.savez_compressed(path, **out)
    print(f"[iso2] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def transpose_inputs(seed: int = 5, n_push: int = 1500, n_tof: int = 257):
    """Push-major CSR of a small synthetic timsTOF file: ragged pushes (20 % empty), ascending tof indices inside a push."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 24, n_push)
    counts[rng.random(n_push) < 0.2] = 0
    push_indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    tof = np.concatenate([np.sort(rng.choice(n_tof, size=c, replace=False)) for c in counts] + [np.zeros(0, np.int64)]).astype(np.uint32)
    values = rng.integers(1, 60000, len(tof)).astype(np.uint16)
    return tof, push_indptr, n_tof, values


def run_transpose():
    """Golden vectors of the load-time CSR transpose: the reference's _transpose (alphadia/raw_data/bruker.py:202-274)."""
    refshim.install()
    bruker = refshim.ref("alphadia.raw_data.bruker")
    tof, push_indptr, n_tof, values = transpose_inputs()
    push_indices, tof_indptr, new_values = bruker._transpose(tof, push_indptr, n_tof, values)
    path = os.path.join(HERE, "transpose_small.npz")
    np.savez_compressed(path, push_indices=push_indices, tof_indptr=tof_indptr, new_values=new_values,
                        input_checksum=hashlib.sha256(tof.tobytes() + push_indptr.tobytes() + values.tobytes()).hexdigest())
    print(f"[transpose] wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB)", flush=True)


def run_k9999_f48(threads: int):
    """The same requantification scoring (top_k_fragments = 9999) on the 48-fragment library `parity_f48` - more fragments per
    candidate than the dense device tables hold (32) - with the candidates the reference itself selects
    -> tests/golden/k9999_f48.npz (candidate table, feature matrix, fragment table)."""
    name = "parity_f48"
    raw, precursor_df, fragment_df, p = make_config_3d(name)
    dia = refshim.RefDiaData(raw)
    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    sel_cfg = cfg_mod.CandidateSelectionConfig()
    sel_cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]), "mobility_tolerance": 0.1, "candidate_count": 3,
                    "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0})
    sel = sel_mod.CandidateSelection(dia, precursor_df.copy(), fragment_df.copy(), sel_cfg, rt_column="rt_library",
                                     mobility_column="mobility_library", precursor_mz_column="mz_library",
                                     fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
    cand = sel(thread_count=threads)
    print(f"[k9999_f48] selection -> {len(cand)} candidates", flush=True)
    sc_cfg = sccfg_mod.CandidateScoringConfig()
    sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, "top_k_fragments": 9999})
    scorer = sc_mod.CandidateScoring(dia_data=dia, precursors_flat=precursor_df.copy(), fragments_flat=fragment_df.copy(), config=sc_cfg,
                                     rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                                     fragment_mz_column="mz_library")
    feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
    print(f"[k9999_f48] scoring -> {len(feat)} rows, {len(frag)} fragment rows", flush=True)
    out = {"input_checksum": np.array(input_checksum(raw, precursor_df, fragment_df))}
    for c in cand.columns:
        out[f"cand_{c}"] = cand[c].values
    out["feat_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
    out["feat_precursor_idx"] = feat["precursor_idx"].values
    out["feat_rank"] = feat["rank"].values
    for c in frag.columns:
        out[f"frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "k9999_f48.npz")
    np.savez_compressed(path, **out)
    print(f"[k9999_f48] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_ties_f20(threads: int):
    """Scoring of the tie-heavy 20-fragment library (tests/helpers.py tied_fragment_library) with top_k_fragments = 20 and 16 by
    the unmodified reference -> tests/golden/ties_f20.npz.  Pins the order numba's (unstable) argsort gives tied fragments."""
    from tests.helpers import tied_fragment_library

    raw, precursor_df, fragment_df, p = tied_fragment_library()
    dia = refshim.RefDiaData(raw)
    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    sel_cfg = cfg_mod.CandidateSelectionConfig()
    sel_cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]), "mobility_tolerance": 0.1, "candidate_count": 3,
                    "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0})
    sel = sel_mod.CandidateSelection(dia, precursor_df.copy(), fragment_df.copy(), sel_cfg, rt_column="rt_library",
                                     mobility_column="mobility_library", precursor_mz_column="mz_library",
                                     fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
    cand = sel(thread_count=threads)
    out = {"input_checksum": np.array(input_checksum(raw, precursor_df, fragment_df))}
    for c in cand.columns:
        out[f"cand_{c}"] = cand[c].values
    for k in (20, 16):
        sc_cfg = sccfg_mod.CandidateScoringConfig()
        sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, "top_k_fragments": k})
        scorer = sc_mod.CandidateScoring(dia_data=dia, precursors_flat=precursor_df.copy(), fragments_flat=fragment_df.copy(), config=sc_cfg,
                                         rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                                         fragment_mz_column="mz_library")
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
        print(f"[ties_f20] top_k_fragments={k}: {len(cand)} candidates -> {len(feat)} rows, {len(frag)} fragment rows", flush=True)
        out[f"feat_k{k}_matrix"] = feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
        out[f"feat_k{k}_precursor_idx"] = feat["precursor_idx"].values
        out[f"feat_k{k}_rank"] = feat["rank"].values
        for c in frag.columns:
            out[f"frag_k{k}_{c}"] = frag[c].values
    path = os.path.join(HERE, "ties_f20.npz")
    np.savez_compressed(path, **out)
    print(f"[ties_f20] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_classifier():
    """BinaryClassifierLegacyNewBatching of the unmodified reference (alphadia/fdr/classifiers.py:145-532), trained here on
    the CPU for three epochs: its state dict, inputs and predict_proba output -> tests/golden/classifier_small.npz."""
    from tests.helpers import classifier_inputs

    cl = refshim.ref("alphadia.fdr.classifiers")
    x, y = classifier_inputs()
    clf = cl.BinaryClassifierLegacyNewBatching(test_size=0.001, batch_size=500, learning_rate=0.001, epochs=3, random_state=3)
    clf.fit(x, y)
    proba = clf.predict_proba(x)
    sd = clf.to_state_dict()
    out = {"input_checksum": np.array(hashlib.sha256(x.tobytes() + y.tobytes()).hexdigest()), "proba": proba.astype(np.float32),
           "predict": clf.predict(x), "layers": np.array(sd["layers"]), "input_dim": np.array(sd["input_dim"])}
    for k, v in sd["network_state_dict"].items():
        out["w__" + k] = v.detach().cpu().numpy()
    acc = float(np.mean(np.argmax(proba, axis=1) == y))
    path = os.path.join(HERE, "classifier_small.npz")
    np.savez_compressed(path, **out)
    print(f"[classifier] accuracy {acc:.3f}, keys {[k for k in out if k.startswith('w__')]}; wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


def run_k9999(threads: int):
    """Transfer-library requantification scoring (transfer_library_requantification_handler.py:102-124): every library
    fragment is quantified, top_k_fragments = 9999, on the 20-fragment library of parity_f20 and the candidates the
    reference selected there (tests/golden/parity_f20.npz) -> tests/golden/k9999.npz."""
    import pandas as pd

    name = "parity_f20"
    raw, precursor_df, fragment_df, p = make_config_3d(name)
    dia = refshim.RefDiaData(raw)
    g = np.load(os.path.join(HERE, f"{name}.npz"), allow_pickle=False)
    assert str(g["input_checksum"]) == input_checksum(raw, precursor_df, fragment_df)
    cand = pd.DataFrame({k[len("cand_"):]: g[k] for k in g.files if k.startswith("cand_")})
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    sc_cfg = sccfg_mod.CandidateScoringConfig()
    sc_cfg.update({**SCORING_BASE, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, "top_k_fragments": 9999})
    scorer = sc_mod.CandidateScoring(
        dia_data=dia, precursors_flat=precursor_df.copy(), fragments_flat=fragment_df.copy(),
        config=sc_cfg, rt_column="rt_library", mobility_column="mobility_library",
        precursor_mz_column="mz_library", fragment_mz_column="mz_library",
    )
    t0 = time.perf_counter()
    feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
    print(f"[k9999] scoring: {time.perf_counter() - t0:.1f}s -> {len(feat)} rows, {len(frag)} fragment rows", flush=True)
    out = {"input_checksum": g["input_checksum"]}
    fcols = sc_mod.DEFAULT_FEATURE_COLUMNS
    out["feat_matrix"] = feat[fcols].values.astype(np.float32)
    out["feat_precursor_idx"] = feat["precursor_idx"].values
    out["feat_rank"] = feat["rank"].values
    for c in frag.columns:
        out[f"frag_{c}"] = frag[c].values
    path = os.path.join(HERE, "k9999.npz")
    np.savez_compressed(path, **out)
    print(f"[k9999] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or ["config1", "parity_small"]
    threads = int(os.environ.get("ADB_THREADS", os.cpu_count() or 1))
    for i, n in enumerate(names):
        if n == "transpose":
            run_transpose()
        elif n == "variants":
            run_variants(threads)
        elif n == "fdr":
            run_fdr()
        elif n == "fdr_nan":
            run_fdr_nan()
        elif n == "ties_f20":
            run_ties_f20(threads)
        elif n == "perform_fdr":
            run_perform_fdr()
        elif n == "ragged":
            run_ragged(threads)
        elif n == "edge":
            run_edge(threads)
        elif n == "scoring_variants":
            run_scoring_variants(threads)
        elif n == "fragcomp_dense":
            run_fragcomp_dense(threads)
        elif n == "variants2":
            run_variants2(threads)
        elif n == "iso2":
            run_iso2(threads)
        elif n == "k9999":
            run_k9999(threads)
        elif n == "classifier":
            run_classifier()
        elif n == "k9999_f48":
            run_k9999_f48(threads)
        else:
            run(n, threads, variants=(n in ("parity_small", "parity_4d", "parity_4d_overlap", "parity_f20", "parity_4d_f20")))

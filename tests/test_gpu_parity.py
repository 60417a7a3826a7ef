"""GPU parity: the CUDA path (through the C ABI) vs the C oracle on identical seeded inputs, and vs the
committed golden vectors of the reference.  Runs on the B200 box (``-m gpu``)."""

import numpy as np
import pandas as pd
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

INT_COLS = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]
FRAG_F32 = ["fragment_mz_library", "fragment_mz", "fragment_mz_observed", "fragment_height", "fragment_intensity",
            "fragment_mass_error", "fragment_correlation"]
FRAG_U8 = ["fragment_position", "fragment_number", "fragment_type", "fragment_charge", "fragment_loss_type"]
# float tolerance of BASELINE.json north_star: 1e-4 relative (absolute floor 1e-6 on the feature scale)
RTOL = 1e-4


@pytest.fixture(scope="module")
def engine():
    from alphadia_b200 import _lib

    _lib.require_device()
    return _lib


def _device_objects(engine, name):
    raw, pdf, fdf, lib, p = H.workload(name)
    return raw, lib, p, engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)


def assert_candidates_equal(a, b):
    for c in INT_COLS:
        assert np.array_equal(a[c], b[c]), c
    assert np.array_equal(a["score"], b["score"]), "score (f32) must be bit-identical"


def feature_scale_floor(F):
    # absolute floor per feature column: 1e-6 of the column's typical magnitude, at least 1e-6
    s = np.nanmedian(np.abs(F), axis=0)
    return np.maximum(1e-6, 1e-6 * np.where(np.isfinite(s), s, 0.0))


def assert_scores_close(a, b, what=""):
    assert np.array_equal(a["valid"], b["valid"]), f"{what} valid mask"
    v = a["valid"].astype(bool)
    Fa, Fb = a["features"][v], b["features"][v]
    nan_a, nan_b = np.isnan(Fa), np.isnan(Fb)
    assert np.array_equal(nan_a, nan_b), f"{what} NaN pattern"
    floor = feature_scale_floor(Fb)
    err = np.abs(Fa - Fb) / np.maximum(np.maximum(np.abs(Fa), np.abs(Fb)), floor[None, :])
    err = np.where(nan_a, 0.0, err)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < RTOL, f"{what} feature {worst[1]} row {worst[0]}: {Fa[worst]} vs {Fb[worst]} (rel {err.max():.3e})"
    for k in FRAG_U8:
        assert np.array_equal(a[k], b[k]), f"{what} {k}"
    for k in FRAG_F32:
        x, y = a[k], b[k]
        assert np.array_equal(x > 0, y > 0) or k not in ("fragment_mz_library", "fragment_mz"), k
        # mass errors are differences of two nearly equal m/z values (ppm): below 1e-3 ppm they are fp64 rounding noise
        # (an observed m/z that equals the library m/z gives +-1e-10 ppm on either side), so the floor is absolute there
        floor = 1e-3 if k == "fragment_mass_error" else 1e-6
        err = H.rel_err(x, y, floor=floor) if x.size else np.zeros(1)
        w = np.unravel_index(np.argmax(err), err.shape)
        assert err.max() < RTOL, f"{what} {k}: rel {err.max():.3e} at {w}: {x[w] if x.size else None!r} vs {y[w] if x.size else None!r}"


@pytest.mark.parametrize("name", ["config1", "parity_small"])
def test_selection_matches_oracle_and_golden(engine, oracle_lib, name):
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    g = H.load_golden(name)
    if g is not None and str(g["input_checksum"]) == H.input_checksum(*H.workload(name)[:3]):
        m = got["score"] > 0
        assert m.sum() == len(g["cand_precursor_idx"])
        for c in INT_COLS:
            assert np.array_equal(got[c][m].astype(np.int64), g["cand_" + c].astype(np.int64)), c
        assert np.array_equal(got["score"][m], g["cand_score"])
    dlib.close(); draw.close()


@pytest.mark.parametrize("kw", [dict(candidate_count=1), dict(candidate_count=5, join_close_candidates=True,
                                                               join_close_candidates_scan_threshold=0.01),
                                dict(use_weighted_score=False), dict(rt_tolerance=8.0), dict(rt_tolerance=400.0),
                                dict(top_k_fragments=6, top_k_precursors=2)])
def test_selection_config_variants(engine, oracle_lib, kw):
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_small")
    args = dict(kw)
    rt_tol = args.pop("rt_tolerance", p["rt_tolerance"])
    cfg = H.selection_config(rt_tol, **args).to_struct()
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    assert (got["score"] > 0).sum() > 0
    dlib.close(); draw.close()


SCORING_VARIANTS = {
    "default": {},
    "legacy": dict(quant_all=False, experimental_xic=False),
    "k6": dict(top_k_fragments=6, top_k_isotopes=4, quant_window=2),
    "qall_legacy_xic": dict(quant_all=True, experimental_xic=False),
    "noqall_xic": dict(quant_all=False, experimental_xic=True),
}


def _candidates(oracle_lib, raw, lib, p):
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    arrs = oracle_lib.select_candidates(raw, lib, cfg, H.default_kernel(raw))
    m = arrs["score"] > 0
    return {c: arrs[c][m] for c in INT_COLS}


@pytest.mark.parametrize("name", ["config1", "parity_small"])
@pytest.mark.parametrize("variant", list(SCORING_VARIANTS))
def test_scoring_matches_oracle(engine, oracle_lib, name, variant):
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    cand = _candidates(oracle_lib, raw, lib, p)
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    cfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
    got = engine.score_candidates(draw, dlib, cfg, cin)
    ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
    assert ref["valid"].sum() > 50
    assert_scores_close(got, ref, what=f"{name}/{variant}")
    dlib.close(); draw.close()


GOLDEN_TAGS = {"": "default", "_legacy": "legacy", "_k6": "k6"}


def _check_scores_against_golden(engine, name, tag):
    g = H.load_golden(name)
    raw, pdf, fdf, lib, p = H.workload(name)
    if g is None or str(g["input_checksum"]) != H.input_checksum(raw, pdf, fdf) or f"feat{tag}_matrix" not in g:
        pytest.skip("golden not applicable")
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    cand = {c: g["cand_" + c] for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    cfg = H.scoring_config(**SCORING_VARIANTS[GOLDEN_TAGS[tag]]).to_struct()
    got = engine.score_candidates(draw, dlib, cfg, cin)
    v = got["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g[f"feat{tag}_precursor_idx"])
    assert np.array_equal(keep["rank"][v], g[f"feat{tag}_rank"])
    F, G = got["features"][v], g[f"feat{tag}_matrix"]
    floor = feature_scale_floor(G)
    err = np.abs(F - G) / np.maximum(np.maximum(np.abs(F), np.abs(G)), floor[None, :])
    err = np.where(np.isnan(F) & np.isnan(G), 0.0, err)
    assert err.max() < RTOL
    m = got["fragment_mz_library"] > 0
    assert m.sum() == len(g[f"frag{tag}_mz_library"])
    assert np.array_equal(got["fragment_mz_library"][m], g[f"frag{tag}_mz_library"])
    assert np.array_equal(got["fragment_number"][m], g[f"frag{tag}_number"])
    dlib.close(); draw.close()


@pytest.mark.parametrize("tag", list(GOLDEN_TAGS))
def test_scoring_matches_reference_golden(engine, tag):
    _check_scores_against_golden(engine, "parity_small", tag)


def test_scoring_edge_cases(engine, oracle_lib):
    """Ragged / degenerate candidates: empty table, tiny and huge windows, clamped limits, too few fragments."""
    raw, pdf, fdf, lib, p = H.workload("parity_small")
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    cfg = H.scoring_config().to_struct()
    L = raw.cycle_len
    # empty
    cin, _ = H.candidates_in_from_arrays(lib, {c: np.zeros(0, np.int64) for c in INT_COLS})
    got = engine.score_candidates(draw, dlib, cfg, cin)
    assert got["features"].shape == (0, 46)
    # hand-made windows of many widths, including the first/last cycles of the run
    rng = np.random.default_rng(5)
    n = 400
    pidx = rng.integers(0, len(pdf), n)
    centers = rng.integers(0, raw.precursor_cycle_max_index, n)
    half = rng.choice([0, 1, 2, 3, 7, 14, 40, 120], n)
    c0 = np.clip(centers - half, 0, raw.precursor_cycle_max_index - 1)
    c1 = np.clip(centers + half + 1, 1, raw.precursor_cycle_max_index)
    cand = dict(precursor_idx=lib["precursor_idx"][pidx].astype(np.int64), rank=np.arange(n) % 7,
                scan_center=np.zeros(n, np.int64), scan_start=np.zeros(n, np.int64), scan_stop=np.ones(n, np.int64),
                frame_center=np.minimum(centers * L, raw.frame_max_index), frame_start=c0 * L,
                frame_stop=np.minimum(c1 * L, raw.frame_max_index))
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    got = engine.score_candidates(draw, dlib, cfg, cin)
    ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
    assert_scores_close(got, ref, what="edge")
    dlib.close(); draw.close()


def test_operator_classes_end_to_end(engine, oracle_lib):
    """The reference-shaped Python API: CandidateSelection -> CandidateScoring -> FragmentCompetition."""
    from alphadia_b200 import CandidateScoring, CandidateSelection, FragmentCompetition
    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS

    raw, pdf, fdf, lib, p = H.workload("parity_small")
    sel = CandidateSelection(raw, pdf.copy(), fdf.copy(), H.selection_config(p["rt_tolerance"]),
                             rt_column="rt_library", mobility_column="mobility_library",
                             precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=5.0,
                             fwhm_mobility=0.01)
    cand_df = sel(thread_count=4)
    assert list(cand_df.columns) == ["precursor_idx", "rank", "score", "scan_center", "scan_start", "scan_stop",
                                     "frame_center", "frame_start", "frame_stop", "elution_group_idx", "decoy"]
    ref = oracle_lib.select_candidates(raw, lib, H.selection_config(p["rt_tolerance"]).to_struct(), sel.kernel)
    m = ref["score"] > 0
    for c in INT_COLS:
        assert np.array_equal(cand_df[c].values.astype(np.int64), ref[c][m].astype(np.int64))

    scorer = CandidateScoring(dia_data=raw, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(),
                              config=H.scoring_config(), rt_column="rt_library", mobility_column="mobility_library",
                              precursor_mz_column="mz_library", fragment_mz_column="mz_library")
    feat_df, frag_df = scorer(cand_df.copy(), thread_count=4, include_decoy_fragment_features=True)
    assert list(feat_df.columns[:46]) == DEFAULT_FEATURE_COLUMNS
    for c in ["precursor_idx", "rank", "elution_group_idx", "decoy", "charge", "delta_rt", "n_K", "n_R", "n_P", "score",
              "rt_library", "mz_library", "i_0", "proteins", "genes", "sequence"]:
        assert c in feat_df.columns, c
    assert list(frag_df.columns) == ["precursor_idx", "rank", "mz_library", "mz", "mz_observed", "height", "intensity",
                                     "mass_error", "correlation", "position", "number", "type", "charge", "loss_type",
                                     "elution_group_idx", "decoy"]
    g = H.load_golden("parity_small")
    if g is not None and str(g["input_checksum"]) == H.input_checksum(raw, pdf, fdf):
        assert np.array_equal(feat_df["precursor_idx"].values, g["feat_precursor_idx"])
        assert len(frag_df) == len(g["frag_mz_library"])
        np.testing.assert_allclose(feat_df["delta_rt"].values, g["feat_delta_rt"], rtol=1e-4, atol=1e-4)

    psm = feat_df.copy()
    psm["proba"] = np.random.default_rng(12345).uniform(0, 1, size=len(psm))
    fc = FragmentCompetition(rt_tol_seconds=3, mass_tol_ppm=15)
    plan = fc.plan(psm.copy(), frag_df.copy(), raw.cycle)
    ref_valid = oracle_lib.fragment_competition(plan.window_start, plan.window_stop, plan.rt, plan.frag_start,
                                                plan.frag_stop, plan.fragment_mz, 3, 15)
    kept = fc(psm.copy(), frag_df.copy(), raw.cycle)
    assert np.array_equal(kept["_candidate_idx"].values, plan.psm_df["_candidate_idx"].values[ref_valid])


@pytest.mark.parametrize("dtype_rt,dtype_mz", [(np.float32, np.float32), (np.float64, np.float32), (np.float64, np.float64)])
def test_fragcomp_dense_competition(engine, oracle_lib, dtype_rt, dtype_mz):
    """Many PSMs sharing fragments inside few windows (the golden runs have no collisions)."""
    rng = np.random.default_rng(7)
    n, nwin = 6000, 5
    base = rng.uniform(200, 1800, size=(300, 12))
    src = rng.integers(0, 300, n)
    mz = base[src] * (1 + rng.normal(0, 4e-6, size=(n, 12)))
    replace = rng.random((n, 12)) < 0.7
    mz = np.where(replace, rng.uniform(200, 1800, size=(n, 12)), mz)
    nfr = rng.integers(4, 13, n)
    frag_start = np.concatenate([[0], np.cumsum(nfr)[:-1]])
    frag_stop = frag_start + nfr
    frag_mz = np.concatenate([np.sort(mz[i, : nfr[i]]) for i in range(n)]).astype(dtype_mz)
    rt = rng.uniform(0, 600, n).astype(dtype_rt)
    bounds = np.linspace(0, n, nwin + 1).astype(np.int64)
    ws, we = bounds[:-1], bounds[1:]
    ref = oracle_lib.fragment_competition(ws, we, rt, frag_start, frag_stop, frag_mz, 3.0, 15.0)
    got = engine.fragment_competition(ws, we, rt, frag_start, frag_stop, frag_mz, 3.0, 15.0, device=0)
    assert ref.sum() < n  # something was vetoed
    assert np.array_equal(got, ref)


def test_fragcomp_reference_known_answers(engine):
    """Known-answer vectors of the reference's own unit tests (tests/unit_tests/fragcomp/test_fragcomp.py:38-58)."""
    rt = np.array([10.0, 10.0, 10.0, 20.0, 20.0, 20.0])
    valid_expected = np.array([True, True, False, True, False, True])
    frag_start = np.array([0, 10, 20, 30, 40, 50])
    frag_stop = frag_start + 10
    fragment_mz = np.hstack([np.arange(100, 110), np.arange(200, 210), np.arange(100, 110),
                             np.arange(100, 110), np.arange(100, 110), np.arange(200, 210)]).astype(np.float64)
    got = engine.fragment_competition(np.array([0, 3]), np.array([3, 6]), rt, frag_start, frag_stop, fragment_mz, 10.0, 10.0, device=0)
    assert np.array_equal(got, valid_expected)


def test_resident_pipeline_matches_host_pipeline(engine):
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_small")
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    host = engine.select_candidates(draw, dlib, cfg, kernel)
    n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
    m = host["score"] > 0
    assert n == m.sum()
    scfg = H.scoring_config().to_struct()
    engine.score_candidates_resident(draw, dlib, scfg)
    res = engine.fetch_scores(draw, n, 12)
    cand = {c: host[c][m] for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    ref = engine.score_candidates(draw, dlib, scfg, cin)
    assert np.array_equal(res["rank"], keep["rank"])
    assert np.array_equal(lib["precursor_idx"][res["lib_row"]], keep["precursor_idx"])
    assert np.array_equal(res["valid"], ref["valid"])
    assert np.array_equal(res["features"], ref["features"], equal_nan=True)
    assert draw.kernel_launches > 0
    dlib.close(); draw.close()


# ---- timsTOF (4-D) -------------------------------------------------------------------------------
def _sel_cfg_4d(p, **kw):
    args = dict(kw)
    rt_tol = args.pop("rt_tolerance", p["rt_tolerance"])
    mob_tol = args.pop("mobility_tolerance", p["mobility_tolerance"])
    return H.selection_config(rt_tol, mobility_tolerance=mob_tol, **args).to_struct()


def _candidates_4d(oracle_lib, raw, lib, p):
    arrs = oracle_lib.select_candidates_4d(raw, lib, _sel_cfg_4d(p), H.default_kernel(raw))
    m = arrs["score"] > 0
    return {c: arrs[c][m] for c in INT_COLS}


def test_selection_4d_matches_oracle_and_golden(engine, oracle_lib):
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_4d")
    assert draw.is_4d
    cfg = _sel_cfg_4d(p)
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates_4d(raw, lib, cfg, kernel)
    assert (ref["score"] > 0).sum() > 100
    assert_candidates_equal(got, ref)
    g = H.load_golden("parity_4d")
    if g is not None and str(g["input_checksum"]) == H.input_checksum(*H.workload("parity_4d")[:3]):
        m = got["score"] > 0
        assert m.sum() == len(g["cand_precursor_idx"])
        for c in INT_COLS:
            assert np.array_equal(got[c][m].astype(np.int64), g["cand_" + c].astype(np.int64)), c
        assert np.array_equal(got["score"][m], g["cand_score"])
    dlib.close(); draw.close()


@pytest.mark.parametrize("kw", [dict(candidate_count=1), dict(candidate_count=5, join_close_candidates=True,
                                                               join_close_candidates_scan_threshold=0.01),
                                dict(use_weighted_score=False), dict(rt_tolerance=5.0), dict(mobility_tolerance=0.2),
                                dict(rt_tolerance=100.0, mobility_tolerance=0.4), dict(top_k_fragments=6, top_k_precursors=2)])
def test_selection_4d_config_variants(engine, oracle_lib, kw):
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_4d")
    cfg = _sel_cfg_4d(p, **kw)
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates_4d(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    assert (got["score"] > 0).sum() > 0
    dlib.close(); draw.close()


@pytest.mark.parametrize("variant", list(SCORING_VARIANTS))
def test_scoring_4d_matches_oracle(engine, oracle_lib, variant):
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_4d")
    cand = _candidates_4d(oracle_lib, raw, lib, p)
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    cfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
    got = engine.score_candidates(draw, dlib, cfg, cin)
    ref = oracle_lib.score_candidates_4d(raw, lib, cfg, cin)
    assert ref["valid"].sum() > 50
    assert np.abs(ref["features"][:, 29]).max() > 0 and np.abs(ref["features"][:, 39]).max() > 0
    assert_scores_close(got, ref, what=f"parity_4d/{variant}")
    dlib.close(); draw.close()


@pytest.mark.parametrize("tag", list(GOLDEN_TAGS))
def test_scoring_4d_matches_reference_golden(engine, tag):
    _check_scores_against_golden(engine, "parity_4d", tag)


def test_scoring_4d_edge_cases(engine, oracle_lib):
    """Hand-made 4-D candidate windows: 1-scan and full-height windows, first/last cycles, huge windows."""
    raw, pdf, fdf, lib, p = H.workload("parity_4d")
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    cfg = H.scoring_config().to_struct()
    Fr, Sc, z = raw.cycle.shape[1], raw.cycle.shape[2], raw.zeroth_frame
    cin, _ = H.candidates_in_from_arrays(lib, {c: np.zeros(0, np.int64) for c in INT_COLS})
    assert engine.score_candidates(draw, dlib, cfg, cin)["features"].shape == (0, 46)
    rng = np.random.default_rng(6)
    n = 300
    pidx = rng.integers(0, len(pdf), n)
    ncyc = raw.precursor_cycle_max_index
    centers = rng.integers(0, ncyc, n)
    half = rng.choice([0, 1, 2, 3, 7, 14, 30], n)
    c0 = np.clip(centers - half, 0, ncyc - 1)
    c1 = np.clip(centers + half + 1, 1, ncyc)
    sc_c = rng.integers(0, Sc, n)
    sh = rng.choice([0, 1, 4, 9, 20, Sc], n)
    s0 = np.clip(sc_c - sh, 0, Sc - 1)
    s1 = np.clip(sc_c + sh + 1, 1, Sc)
    cand = dict(precursor_idx=lib["precursor_idx"][pidx].astype(np.int64), rank=np.arange(n) % 7,
                scan_center=sc_c, scan_start=s0, scan_stop=s1,
                frame_center=np.minimum(centers * Fr + z, raw.frame_max_index - 1), frame_start=c0 * Fr + z,
                frame_stop=np.minimum(c1 * Fr + z, raw.frame_max_index))
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    got = engine.score_candidates(draw, dlib, cfg, cin)
    ref = oracle_lib.score_candidates_4d(raw, lib, cfg, cin)
    assert ref["valid"].sum() > 20
    assert_scores_close(got, ref, what="edge4d")
    dlib.close(); draw.close()


def test_operator_classes_4d(engine, oracle_lib):
    """CandidateSelection -> CandidateScoring on a timsTOF-shaped raw file through the reference-shaped API."""
    from alphadia_b200 import CandidateScoring, CandidateSelection

    raw, pdf, fdf, lib, p = H.workload("parity_4d")
    scfg = H.selection_config(p["rt_tolerance"], mobility_tolerance=p["mobility_tolerance"])
    sel = CandidateSelection(raw, pdf.copy(), fdf.copy(), scfg, rt_column="rt_library", mobility_column="mobility_library",
                             precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
    assert sel.kernel.shape == (30, 30)
    cand_df = sel(thread_count=4)
    g = H.load_golden("parity_4d")
    if g is not None and str(g["input_checksum"]) == H.input_checksum(raw, pdf, fdf):
        for c in INT_COLS:
            assert np.array_equal(cand_df[c].values.astype(np.int64), g["cand_" + c].astype(np.int64)), c
    scorer = CandidateScoring(dia_data=raw, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(), config=H.scoring_config(),
                              rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                              fragment_mz_column="mz_library")
    feat_df, frag_df = scorer(cand_df.copy(), thread_count=4, include_decoy_fragment_features=True)
    if g is not None and str(g["input_checksum"]) == H.input_checksum(raw, pdf, fdf):
        assert np.array_equal(feat_df["precursor_idx"].values, g["feat_precursor_idx"])
        assert len(frag_df) == len(g["frag_mz_library"])
    assert feat_df["mobility_fwhm"].abs().max() > 0


def test_resident_pipeline_4d(engine):
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_4d")
    cfg = _sel_cfg_4d(p)
    kernel = H.default_kernel(raw)
    host = engine.select_candidates(draw, dlib, cfg, kernel)
    n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
    m = host["score"] > 0
    assert n == m.sum()
    scfg = H.scoring_config().to_struct()
    engine.score_candidates_resident(draw, dlib, scfg)
    res = engine.fetch_scores(draw, n, 12)
    cand = {c: host[c][m] for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    ref = engine.score_candidates(draw, dlib, scfg, cin)
    assert np.array_equal(res["valid"], ref["valid"])
    assert np.array_equal(res["features"], ref["features"], equal_nan=True)
    dlib.close(); draw.close()


@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_candidate_table_matches_container(engine, name):
    """adb_fetch_candidate_table == the `score > 0` rows of the candidate container, in container order."""
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    cfg = _sel_cfg_4d(p) if name == "parity_4d" else H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    host = engine.select_candidates(draw, dlib, cfg, kernel)
    n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
    table = engine.fetch_candidate_table(draw, n)
    m = host["score"] > 0
    assert n == m.sum() > 0
    for c in INT_COLS:
        assert np.array_equal(table[c][:n].astype(np.int64), host[c][m].astype(np.int64)), c
    assert np.array_equal(table["score"][:n], host["score"][m])
    assert np.array_equal(lib["precursor_idx"][table["lib_row"][:n]], host["precursor_idx"][m])
    # the table is scoring input as it is
    from alphadia_b200 import _abi
    cin = _abi.candidates_in_from_table(table, n)
    got = engine.score_candidates(draw, dlib, H.scoring_config().to_struct(), cin)
    cin2, keep = H.candidates_in_from_arrays(lib, {c: host[c][m] for c in INT_COLS})
    ref = engine.score_candidates(draw, dlib, H.scoring_config().to_struct(), cin2)
    assert np.array_equal(got["features"], ref["features"], equal_nan=True)
    dlib.close(); draw.close()


def test_chunked_scoring_equals_single_block(engine):
    """>= 200k candidates are scored in 4 row blocks with overlapped D2H: same result as the resident single launch."""
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_small")
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    n = engine.select_candidates_resident(draw, dlib, cfg, H.default_kernel(raw))
    table = engine.fetch_candidate_table(draw, n)
    reps = 200000 // n + 1
    big = {k: np.ascontiguousarray(np.tile(v[:n], reps)) for k, v in table.items()}
    from alphadia_b200 import _abi
    N = n * reps
    assert N >= 200000
    scfg = H.scoring_config().to_struct()
    got = engine.score_candidates(draw, dlib, scfg, _abi.candidates_in_from_table(big, N))
    one = engine.score_candidates(draw, dlib, scfg, _abi.candidates_in_from_table(table, n))
    for k in ("features", "valid", "fragment_mz_observed", "fragment_correlation", "fragment_type"):
        assert np.array_equal(got[k], np.concatenate([one[k]] * reps), equal_nan=True), k
    dlib.close(); draw.close()


@pytest.mark.parametrize("kw", [dict(), dict(rt_tolerance=400.0), dict(use_weighted_score=False, candidate_count=5)])
def test_selection_legacy_pair_matches_oracle(engine, oracle_lib, monkeypatch, kw):
    """Cycle windows too long for the fused kernel use the extract + dense-smoothing kernel pair; force it here."""
    monkeypatch.setenv("ADB_SELECT_LEGACY", "1")
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_small")
    args = dict(kw)
    rt_tol = args.pop("rt_tolerance", p["rt_tolerance"])
    cfg = H.selection_config(rt_tol, **args).to_struct()
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    assert (got["score"] > 0).sum() > 0
    dlib.close(); draw.close()


def _ragged_library(lib, rng, raw_rt_max):
    """Library variants the reference meets in practice: precursors with too few fragments, shared (cardinality > 1)
    fragments, duplicate fragment m/z, overlapping ppm windows, retention times outside the run, charge 1-4."""
    lib = {k: v.copy() for k, v in lib.items()}
    P = len(lib["precursor_idx"])
    # 1. fragment counts 0..12: shorten every 5th precursor
    for i in range(0, P, 5):
        n = int(rng.integers(0, 13))
        lib["frag_stop_idx"][i] = lib["frag_start_idx"][i] + n
    # 2. shared ions
    lib["frag_cardinality"][rng.random(len(lib["frag_cardinality"])) < 0.15] = 2
    # 3. duplicate and nearly identical fragment m/z (overlapping windows exercise the forward-only cursor)
    fs = lib["frag_start_idx"].astype(np.int64)
    for i in range(1, P, 7):
        s = fs[i]
        lib["frag_mz"][s + 1] = lib["frag_mz"][s]
        lib["frag_mz"][s + 3] = np.float32(lib["frag_mz"][s + 2] * (1 + 6e-6))
    # 4. retention times before / after the run, charges 1..4
    lib["rt"][::11] = np.float32(-50.0)
    lib["rt"][5::13] = np.float32(raw_rt_max + 500.0)
    lib["charge"][:] = rng.integers(1, 5, size=P).astype(np.uint8)
    return lib


@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_selection_ragged_library(engine, oracle_lib, name):
    raw, pdf, fdf, lib0, p = H.workload(name)
    lib = _ragged_library(lib0, np.random.default_rng(17), float(np.max(raw.rt_values)))
    is4d = name == "parity_4d"
    cfg = _sel_cfg_4d(p) if is4d else H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = (oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates)(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    assert 0 < (got["score"] > 0).sum() < (H.workload(name)[3]["precursor_idx"].shape[0] * 3)
    # and the ragged candidates score identically
    m = got["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    scfg = H.scoring_config().to_struct()
    s_got = engine.score_candidates(draw, dlib, scfg, cin)
    s_ref = (oracle_lib.score_candidates_4d if is4d else oracle_lib.score_candidates)(raw, lib, scfg, cin)
    assert_scores_close(s_got, s_ref, what=f"ragged/{name}")
    dlib.close(); draw.close()


@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_empty_and_single_precursor_library(engine, oracle_lib, name):
    raw, pdf, fdf, lib0, p = H.workload(name)
    is4d = name == "parity_4d"
    cfg = _sel_cfg_4d(p) if is4d else H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    draw = engine.DeviceRawFile(raw, device=0)
    for n_keep in (0, 1):
        lib = {k: (v[:n_keep].copy() if k in ("precursor_idx", "frag_start_idx", "frag_stop_idx", "charge", "rt", "mobility", "mz", "isotopes") else v)
               for k, v in lib0.items()}
        dlib = engine.DeviceLibrary(lib, device=0)
        got = engine.select_candidates(draw, dlib, cfg, kernel)
        assert len(got["score"]) == 3 * n_keep
        if n_keep:
            ref = (oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates)(raw, lib, cfg, kernel)
            assert_candidates_equal(got, ref)
        n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
        assert n == int((got["score"] > 0).sum())
        engine.score_candidates_resident(draw, dlib, H.scoring_config().to_struct())
        dlib.close()
    draw.close()


def test_unsupported_inputs_fail_loudly(engine):
    """Inputs the device cannot take return an error through adb_last_error (no silent fallback)."""
    raw, pdf, fdf, lib0, p = H.workload("parity_small")
    draw = engine.DeviceRawFile(raw, device=0)
    lib = {k: v.copy() for k, v in lib0.items()}
    lib["frag_stop_idx"][0] = lib["frag_start_idx"][0] + 200  # > 128 library fragments for one precursor
    dlib = engine.DeviceLibrary(lib, device=0)
    with pytest.raises(RuntimeError, match="library fragments"):
        engine.select_candidates(draw, dlib, H.selection_config(p["rt_tolerance"]).to_struct(), H.default_kernel(raw))
    dlib.close()
    dlib = engine.DeviceLibrary(lib0, device=0)
    with pytest.raises(RuntimeError, match="candidate_count"):
        engine.select_candidates(draw, dlib, H.selection_config(p["rt_tolerance"], candidate_count=40).to_struct(), H.default_kernel(raw))
    bad = dict(precursor_idx=np.array([lib0["precursor_idx"][0]]), rank=np.array([0]), scan_start=np.array([0]), scan_stop=np.array([1]),
               scan_center=np.array([0]), frame_start=np.array([0]), frame_stop=np.array([76 * 5]), frame_center=np.array([76 * 2]))
    cin, keep = H.candidates_in_from_arrays(lib0, bad)
    keep["lib_row"][0] = 10 ** 9  # outside the library
    with pytest.raises(RuntimeError, match="outside the library"):
        engine.score_candidates(draw, dlib, H.scoring_config().to_struct(), cin)
    with pytest.raises(RuntimeError, match="top_k_fragments"):
        from alphadia_b200 import _abi
        cfg = H.scoring_config().to_struct()
        cfg.top_k_fragments = 64
        engine.score_candidates(draw, dlib, cfg, H.candidates_in_from_arrays(lib0, bad)[0])
    dlib.close(); draw.close()


def test_full_size_config2_properties(engine, oracle_lib):
    """BASELINE.json config 2 at full size (50 k precursors, 91 200 spectra, 1.4e8 peaks): size-independent properties of the
    candidate table (determinism, shard invariance, structural invariants) plus oracle parity on random subsamples."""
    raw, pdf, fdf, lib, p = H.workload("config2")
    P = len(lib["precursor_idx"])
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    full = engine.select_candidates(draw, dlib, cfg, kernel)
    again = engine.select_candidates(draw, dlib, cfg, kernel)
    for c in INT_COLS + ["score"]:
        assert np.array_equal(full[c], again[c]), f"selection is not deterministic in {c}"
    m = full["score"] > 0
    assert m.sum() > 0.5 * 3 * P
    # structural invariants of the candidate container (selection.py:480-526)
    fs, fc, fe = (full[c][m].astype(np.int64) for c in ("frame_start", "frame_center", "frame_stop"))
    assert (fs <= fc).all() and (fc < fe).all() and (fe <= raw.frame_max_index).all()
    L = raw.cycle_len
    assert ((fe - fs) // L <= 2 * 15 - 1).all() and ((fe - fs) // L >= 3 + 1).all()  # centre +- [3, 14] cycles, clipped at the run edges
    rows = np.flatnonzero(m)
    prec, rank = rows // 3, full["rank"][m].astype(np.int64)
    assert np.array_equal(rank, rows % 3)  # ranks are the positions inside the precursor's block, no holes
    sc = full["score"].reshape(P, 3)
    assert (np.diff(np.where(sc > 0, sc, -np.inf), axis=1) <= 0).all()  # scores descend with rank
    assert np.array_equal(full["precursor_idx"][m], lib["precursor_idx"][prec])
    # shard invariance: two library halves == the whole library (bit-exact)
    half = P // 2
    prec_keys = ("precursor_idx", "frag_start_idx", "frag_stop_idx", "charge", "rt", "mobility", "mz", "isotopes")
    parts = []
    for sl in (slice(0, half), slice(half, P)):
        sub = {k: (np.ascontiguousarray(v[sl]) if k in prec_keys else v) for k, v in lib.items()}
        dsub = engine.DeviceLibrary(sub, device=0)
        parts.append(engine.select_candidates(draw, dsub, cfg, kernel))
        dsub.close()
    for c in INT_COLS + ["score"]:
        assert np.array_equal(np.concatenate([parts[0][c], parts[1][c]]), full[c]), f"shard invariance broken in {c}"
    # oracle parity on a random subsample of precursors
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(P, size=4000, replace=False))
    sub = {k: (np.ascontiguousarray(v[pick]) if k in prec_keys else v) for k, v in lib.items()}
    ref = oracle_lib.select_candidates(raw, sub, cfg, kernel)
    idx = (pick[:, None] * 3 + np.arange(3)[None, :]).ravel()
    for c in INT_COLS + ["score"]:
        assert np.array_equal(full[c][idx], ref[c]), f"oracle parity (subsample) broken in {c}"
    # scoring: the whole table on the device, a random subsample against the oracle, determinism
    n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
    table = engine.fetch_candidate_table(draw, n)
    from alphadia_b200 import _abi
    scfg = H.scoring_config().to_struct()
    got = engine.score_candidates(draw, dlib, scfg, _abi.candidates_in_from_table(table, n))
    got2 = engine.score_candidates(draw, dlib, scfg, _abi.candidates_in_from_table(table, n))
    assert np.array_equal(got["features"], got2["features"], equal_nan=True) and np.array_equal(got["valid"], got2["valid"])
    assert got["valid"].sum() > 0.5 * n
    pick_c = np.sort(rng.choice(n, size=6000, replace=False))
    sub_t = {k: np.ascontiguousarray(v[:n][pick_c]) for k, v in table.items()}
    ref_s = oracle_lib.score_candidates(raw, lib, scfg, _abi.candidates_in_from_table(sub_t, len(pick_c)))
    got_s = {k: v[pick_c] for k, v in got.items()}
    assert_scores_close(got_s, ref_s, what="config2 subsample")
    dlib.close(); draw.close()


def test_mid_size_4d_properties(engine, oracle_lib):
    """A timsTOF-shaped run with 3 000 precursors (about 12x the golden case): determinism, shard invariance, structural
    invariants on the whole candidate table, oracle parity on a random subsample (selection bit-exact, scores <= 1e-4)."""
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.synthetic import make_config_4d

    raw, pdf, fdf, p = make_config_4d("parity_4d", n_precursors=3000, seed=77)
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    P = len(lib["precursor_idx"])
    cfg = _sel_cfg_4d(p)
    kernel = H.default_kernel(raw)
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    full = engine.select_candidates(draw, dlib, cfg, kernel)
    again = engine.select_candidates(draw, dlib, cfg, kernel)
    for c in INT_COLS + ["score"]:
        assert np.array_equal(full[c], again[c]), f"4-D selection is not deterministic in {c}"
    m = full["score"] > 0
    assert m.sum() > P
    s0, sc_, s1 = (full[c][m].astype(np.int64) for c in ("scan_start", "scan_center", "scan_stop"))
    f0, fc, f1 = (full[c][m].astype(np.int64) for c in ("frame_start", "frame_center", "frame_stop"))
    assert (s0 <= sc_).all() and (sc_ < s1).all() and (s1 <= raw.scan_max_index).all()
    assert (f0 <= fc).all() and (fc < f1).all() and (f1 <= raw.frame_max_index).all()
    assert ((s1 - s0) <= 2 * 20 - 1).all()  # max_size_mobility = 20
    prec_keys = ("precursor_idx", "frag_start_idx", "frag_stop_idx", "charge", "rt", "mobility", "mz", "isotopes")
    half = P // 2
    parts = []
    for sl in (slice(0, half), slice(half, P)):
        sub = {k: (np.ascontiguousarray(v[sl]) if k in prec_keys else v) for k, v in lib.items()}
        dsub = engine.DeviceLibrary(sub, device=0)
        parts.append(engine.select_candidates(draw, dsub, cfg, kernel))
        dsub.close()
    for c in INT_COLS + ["score"]:
        assert np.array_equal(np.concatenate([parts[0][c], parts[1][c]]), full[c]), f"4-D shard invariance broken in {c}"
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(P, size=250, replace=False))
    sub = {k: (np.ascontiguousarray(v[pick]) if k in prec_keys else v) for k, v in lib.items()}
    ref = oracle_lib.select_candidates_4d(raw, sub, cfg, kernel)
    idx = (pick[:, None] * 3 + np.arange(3)[None, :]).ravel()
    for c in INT_COLS + ["score"]:
        assert np.array_equal(full[c][idx], ref[c]), f"4-D oracle parity (subsample) broken in {c}"
    n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
    table = engine.fetch_candidate_table(draw, n)
    from alphadia_b200 import _abi
    scfg = H.scoring_config().to_struct()
    got = engine.score_candidates(draw, dlib, scfg, _abi.candidates_in_from_table(table, n))
    got2 = engine.score_candidates(draw, dlib, scfg, _abi.candidates_in_from_table(table, n))
    assert np.array_equal(got["features"], got2["features"], equal_nan=True)
    pick_c = np.sort(rng.choice(n, size=min(n, 1500), replace=False))
    sub_t = {k: np.ascontiguousarray(v[:n][pick_c]) for k, v in table.items()}
    ref_s = oracle_lib.score_candidates_4d(raw, lib, scfg, _abi.candidates_in_from_table(sub_t, len(pick_c)))
    assert_scores_close({k: v[pick_c] for k, v in got.items()}, ref_s, what="4-D mid-size subsample")
    dlib.close(); draw.close()


def test_4d_repeated_observation_ids(engine, oracle_lib):
    """Two frames of the cycle carry the same observation id (e.g. two MS1 frames per cycle): events of one tof row can then
    fall into the same cube cell, the scoring kernel must process each query sequentially (DevRaw4::obs_unique_per_scan = 0)."""
    from types import SimpleNamespace

    raw0, pdf, fdf, lib, p = H.workload("parity_4d")
    fr, sc = raw0.cycle.shape[1], raw0.cycle.shape[2]
    cycle = raw0.cycle.copy()
    cycle[0, 2, :, :] = -1.0                       # frame 2 of the cycle becomes a second MS1 frame ...
    dpc = raw0.dia_precursor_cycle.copy()
    dpc[2 * sc:3 * sc] = 0                         # ... with the observation id of the first one
    dpc[3 * sc:4 * sc] = 1                         # and the last MS2 frame reuses id 1: repeated ids among MS2 frames as well
    cycle[0, 3, :, :] = cycle[0, 1, :, :]
    raw = SimpleNamespace(cycle=cycle, rt_values=raw0.rt_values, mobility_values=raw0.mobility_values, mz_values=raw0.mz_values,
                          tof_indptr=raw0.tof_indptr, push_indices=raw0.push_indices, intensity_values=raw0.intensity_values,
                          dia_precursor_cycle=dpc, zeroth_frame=raw0.zeroth_frame, scan_max_index=raw0.scan_max_index,
                          frame_max_index=raw0.frame_max_index, precursor_cycle_max_index=raw0.precursor_cycle_max_index,
                          has_mobility=True, is_4d=True)
    cfg = _sel_cfg_4d(p)
    kernel = H.default_kernel(raw0)
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates_4d(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    m = got["score"] > 0
    assert m.sum() > 50
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    scfg = H.scoring_config().to_struct()
    s_got = engine.score_candidates(draw, dlib, scfg, cin)
    s_ref = oracle_lib.score_candidates_4d(raw, lib, scfg, cin)
    assert s_ref["valid"].sum() > 20
    assert_scores_close(s_got, s_ref, what="repeated observation ids")
    dlib.close(); draw.close()


def test_selection_windows_longer_than_1024_cycles(engine, oracle_lib):
    """rt_tolerance larger than the run: every precursor's window spans all 1400 cycles, which takes the extract +
    dense-smoothing kernel pair with XIC rows accumulated directly in the HBM buffer."""
    raw, pdf, fdf, lib, p = H.workload("long_run")
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    m = got["score"] > 0
    assert m.sum() > 100
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    scfg = H.scoring_config().to_struct()
    assert_scores_close(engine.score_candidates(draw, dlib, scfg, cin), oracle_lib.score_candidates(raw, lib, scfg, cin), what="long_run")
    dlib.close(); draw.close()


def test_transpose_csr_matches_oracle_and_golden(engine, oracle_lib):
    """adb_transpose_csr == the oracle == the reference's _transpose golden; ragged and degenerate inputs."""
    import os

    from tests.test_oracle_golden import _transpose_inputs

    tof, push_indptr, n_tof, values = _transpose_inputs()
    got = engine.transpose_csr(tof, push_indptr, n_tof, values, device=0)
    g = np.load(os.path.join(H.GOLDEN_DIR, "transpose_small.npz"), allow_pickle=False)
    for a, k in zip(got, ("push_indices", "tof_indptr", "new_values")):
        assert np.array_equal(a, g[k]) and a.dtype == g[k].dtype, k
    # a larger ragged case against the oracle, then the real layout: the synthetic timsTOF file, transposed back and forth
    tof, push_indptr, n_tof, values = _transpose_inputs(seed=9, n_push=200_000, n_tof=40_000)
    ref = oracle_lib.transpose_csr(tof, push_indptr, n_tof, values)
    got = engine.transpose_csr(tof, push_indptr, n_tof, values, device=0)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    raw = H.workload("parity_4d")[0]
    n_push = raw.frame_max_index * raw.scan_max_index
    tof_of = np.repeat(np.arange(len(raw.mz_values), dtype=np.uint32), np.diff(raw.tof_indptr))
    back = np.argsort(raw.push_indices, kind="stable")  # tof-major -> push-major (tof ascending inside a push)
    pm_tof, pm_val = tof_of[back], raw.intensity_values[back]
    pm_indptr = np.concatenate([[0], np.cumsum(np.bincount(raw.push_indices, minlength=n_push))]).astype(np.int64)
    push_indices, tof_indptr, new_values = engine.transpose_csr(pm_tof, pm_indptr, len(raw.mz_values), pm_val, device=0)
    assert np.array_equal(push_indices, raw.push_indices) and np.array_equal(tof_indptr, raw.tof_indptr)
    assert np.array_equal(new_values, raw.intensity_values)
    # degenerate inputs and the error path
    p0, i0, v0 = engine.transpose_csr(np.zeros(0, np.uint32), np.zeros(5, np.int64), 7, np.zeros(0, np.uint16), device=0)
    assert len(p0) == 0 and np.array_equal(i0, np.zeros(8, np.int64))
    with pytest.raises(RuntimeError, match="n_tof_indices"):
        engine.transpose_csr(np.array([1, 9], np.uint32), np.array([0, 2], np.int64), 5, np.array([1, 2], np.uint16), device=0)


def test_multiplexed_score_groups_and_reference_channel(engine, oracle_lib):
    """MultiplexingRequantificationHandler settings (score_grouped=True, reference_channel=0): candidates of the channels
    of one elution group form a score group; groups without the reference channel are skipped entirely
    (score_group.py:49-63), the others are scored candidate by candidate exactly like the ungrouped path."""
    from alphadia_b200 import CandidateScoring, CandidateSelection

    raw, pdf, fdf, lib, p = H.workload("parity_small")
    pdf = pdf.copy()
    n = len(pdf)
    # three channels per elution group: precursors 3k, 3k+1, 3k+2 share elution group k and carry channels 0 / 4 / 8;
    # every fifth group lacks channel 0
    pdf["elution_group_idx"] = (np.arange(n) // 3).astype(np.uint32)
    pdf["decoy"] = ((np.arange(n) // 3) % 2).astype(np.uint8)
    ch = np.array([0, 4, 8], dtype=np.uint32)[np.arange(n) % 3]
    ch[(np.arange(n) // 3) % 5 == 0] = np.array([12, 4, 8], dtype=np.uint32)[np.arange(n) % 3][(np.arange(n) // 3) % 5 == 0]
    pdf["channel"] = ch
    sel = CandidateSelection(raw, pdf.copy(), fdf.copy(), H.selection_config(p["rt_tolerance"], candidate_count=1),
                             rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                             fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
    cand_df = sel()
    cfg = H.scoring_config(score_grouped=True, reference_channel=0)
    scorer = CandidateScoring(dia_data=raw, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(), config=cfg,
                              rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                              fragment_mz_column="mz_library")
    feat_df, frag_df = scorer(cand_df.copy())
    # a score group has the reference channel only if the channel-0 precursor itself produced a candidate
    cand_channel = pdf.set_index("precursor_idx")["channel"].reindex(cand_df["precursor_idx"]).values
    groups_with_ref = set(cand_df["elution_group_idx"].values[cand_channel == 0])
    assert 0 < len(groups_with_ref) < cand_df["elution_group_idx"].nunique()
    assert len(feat_df) > 50
    assert set(feat_df["elution_group_idx"]).issubset(groups_with_ref)
    # the same candidates through the ungrouped configuration: identical feature rows for the processed groups
    scorer0 = CandidateScoring(dia_data=raw, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(), config=H.scoring_config(),
                               rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                               fragment_mz_column="mz_library")
    feat0, _ = scorer0(cand_df.copy())
    keep0 = feat0[feat0["elution_group_idx"].isin(groups_with_ref)].sort_values(["precursor_idx", "rank"]).reset_index(drop=True)
    got = feat_df.sort_values(["precursor_idx", "rank"]).reset_index(drop=True)
    assert np.array_equal(got["precursor_idx"].values, keep0["precursor_idx"].values)
    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS
    assert np.array_equal(got[DEFAULT_FEATURE_COLUMNS].values, keep0[DEFAULT_FEATURE_COLUMNS].values, equal_nan=True)
    # a precursor twice in one score group is rejected like in the reference (score_group.py:221-226)
    dup = pd.concat([cand_df, cand_df.iloc[:1]], ignore_index=True)
    with pytest.raises(ValueError, match="unique within a score group"):
        scorer(dup)


def test_fdr_bookkeeping(engine, oracle_lib):
    """adb_q_values / adb_keep_best == the oracle == the live reference's get_q_values / keep_best
    (tests/golden/fdr_small.npz) == the reference's own unit-test vectors; bit-exact (integer order, float64 q-values)."""
    import hashlib
    import os

    from alphadia_b200 import fdr
    from tests.test_reference_known_answers import FDR_KNOWN_ANSWERS

    for check in FDR_KNOWN_ANSWERS:
        check(fdr)
    df = H.fdr_inputs()
    path = os.path.join(H.GOLDEN_DIR, "fdr_small.npz")
    g = np.load(path, allow_pickle=False) if os.path.exists(path) else None
    if g is not None and str(g["input_checksum"]) == hashlib.sha256(df.to_numpy().tobytes() + df.index.to_numpy().tobytes()).hexdigest():
        q = fdr.get_q_values(df.copy(), "proba", "_decoy")
        assert np.array_equal(q["row"].values, g["q_row"]) and np.array_equal(q.index.values, g["q_index"])
        assert np.array_equal(q["qval"].values, g["q_qval"])
        q2 = fdr.get_q_values(df.copy(), "proba", "_decoy", extra_sort_columns=["precursor_idx", "rank"])
        assert np.array_equal(q2["row"].values, g["q2_row"]) and np.array_equal(q2["qval"].values, g["q2_qval"])
        for tag, cols in {"precursor": ["precursor_idx"], "channel_eg": ["elution_group_idx", "channel"], "eg": ["elution_group_idx"],
                          "default": None}.items():
            assert np.array_equal(fdr.keep_best(df.copy(), group_columns=cols)["row"].values, g[f"keep_{tag}_row"]), tag
        final = fdr.get_q_values(fdr.keep_best(q, group_columns=["elution_group_idx", "channel"]), "proba", "_decoy")
        assert np.array_equal(final["row"].values, g["final_row"]) and np.array_equal(final["qval"].values, g["final_qval"])
    # larger random tables against the oracle: heavy ties, negative / zero / huge scores, all-decoy and all-target runs
    rng = np.random.default_rng(3)
    for n, levels in [(1, 1), (2, 1), (1000, 7), (300_000, 5000), (300_000, 10 ** 9)]:
        score = rng.integers(-levels, levels + 1, n) / max(levels, 1) * rng.choice([1.0, 1e-300, 1e300])
        decoy = (rng.random(n) < rng.choice([0.0, 0.5, 1.0])).astype(np.uint8)
        extra = rng.integers(0, max(n // 4, 1), n).astype(np.uint64)
        o_dev, q_dev = engine.q_values(score, decoy, extra)
        o_ref, q_ref = oracle_lib.q_values(score, decoy, extra)
        assert np.array_equal(o_dev, o_ref), f"q-value order n={n}"
        assert np.array_equal(q_dev, q_ref, equal_nan=True), f"q-values n={n}"
        group = rng.integers(0, max(n // 3, 1), n).astype(np.uint64) << np.uint64(rng.integers(0, 40))
        assert np.array_equal(engine.keep_best(score, group), oracle_lib.keep_best(score, group)), f"keep_best n={n}"
    # loud failures
    with pytest.raises(RuntimeError, match="NaN"):
        engine.q_values(np.array([0.1, np.nan]), np.array([0, 1], np.uint8), np.array([0, 1], np.uint64))
    with pytest.raises(RuntimeError, match="NaN"):
        engine.keep_best(np.array([np.nan]), np.array([0], np.uint64))
    with pytest.raises(ValueError, match="_decoy"):
        fdr.get_q_values(pd.DataFrame({"precursor_idx": [0], "proba": [0.5], "_decoy": [2]}))


def test_fdr_missing_values_device_vs_reference(engine):
    """NaN scores / NaN group keys through the device calls against the live reference's tables (tests/golden/fdr_nan.npz)."""
    from alphadia_b200 import fdr

    g = H.load_golden("fdr_nan")
    df = H.fdr_inputs_nan()
    q = fdr.get_q_values(df.copy(), "proba", "_decoy")
    assert np.array_equal(q["row"].values, g["q_row"]) and np.array_equal(q["qval"].values, g["q_qval"], equal_nan=True)
    q2 = fdr.get_q_values(df.copy(), "proba", "_decoy", extra_sort_columns=["rank", "gnan"])
    assert np.array_equal(q2["row"].values, g["q2_row"]) and np.array_equal(q2["qval"].values, g["q2_qval"], equal_nan=True)
    for tag, cols in {"precursor": ["precursor_idx"], "gnan": ["gnan"], "gnan_channel": ["gnan", "channel"]}.items():
        assert np.array_equal(fdr.keep_best(df.copy(), group_columns=cols)["row"].values, g[f"keep_{tag}_row"]), tag


def test_perform_fdr_matches_reference_golden(engine):
    """The perform_fdr sequence (q-values -> fragment competition -> best per group -> q-values) on the device against the
    reference's perform_fdr run with the same stand-in classifier (tests/golden/perform_fdr_small.npz)."""
    from alphadia_b200 import fdr
    from tests.test_oracle_golden import check_perform_fdr_against_golden

    check_perform_fdr_against_golden(fdr)


def test_4d_two_observations_per_candidate(engine, oracle_lib):
    """timsTOF file whose neighbouring quadrupole windows overlap (parity_4d_overlap): about a fifth of the candidates are
    seen by two frames of the cycle, so the 4-D scoring kernel runs with n_observations = 2 (observation importance,
    collapsed MS1, per-observation profiles)."""
    name = "parity_4d_overlap"
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    cfg = _sel_cfg_4d(p)
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates_4d(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    m = got["score"] > 0
    assert m.sum() > 100
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    for variant in ("default", "legacy", "k6"):
        scfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
        s_got = engine.score_candidates(draw, dlib, scfg, cin)
        s_ref = oracle_lib.score_candidates_4d(raw, lib, scfg, cin)
        v = s_ref["valid"].astype(bool)
        assert (s_ref["features"][v, 17] == 2).sum() >= 10 and (s_ref["features"][v, 17] == 1).sum() >= 10
        assert_scores_close(s_got, s_ref, what=f"{name}/{variant}")
    dlib.close(); draw.close()
    # the oracle itself is pinned against the live reference on this file (tests/golden/parity_4d_overlap.npz,
    # tests/test_oracle_golden.py::test_scoring_4d_vs_reference[parity_4d_overlap-*])


# ---- cases whose oracle is pinned against the live reference by dedicated golden files (variants2, scoring_variants, ragged,
# parity_f20, parity_4d_f20, iso2): device vs oracle ------------------------------------------------------------------------


@pytest.mark.parametrize("tag", list(H.SELECTION_VARIANTS2))
@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_extra_selection_variants2(engine, oracle_lib, name, tag):
    from alphadia_b200.kernel import GaussianKernel

    raw, lib, p, draw, dlib = _device_objects(engine, name)
    kw = dict(H.SELECTION_VARIANTS2[tag])
    fwhm_rt, fwhm_mobility = kw.pop("fwhm_rt", 5.0), kw.pop("fwhm_mobility", 0.01)
    if "mobility_tolerance" in p:
        kw.setdefault("mobility_tolerance", p["mobility_tolerance"])
    config = H.selection_config(p["rt_tolerance"], **kw)
    kernel = GaussianKernel(raw, fwhm_rt=fwhm_rt, sigma_scale_rt=config.sigma_scale_rt, fwhm_mobility=fwhm_mobility,
                            sigma_scale_mobility=config.sigma_scale_mobility, kernel_width=config.kernel_size,
                            kernel_height=min(config.kernel_size, raw.scan_max_index + 1)).get_dense_matrix(verbose=False)
    got = engine.select_candidates(draw, dlib, config.to_struct(), kernel)
    ref = (oracle_lib.select_candidates_4d if "mobility_tolerance" in p else oracle_lib.select_candidates)(raw, lib, config.to_struct(), kernel)
    assert_candidates_equal(got, ref)
    dlib.close(); draw.close()


@pytest.mark.parametrize("tag", list(H.SCORING_VARIANTS_EXTRA))
@pytest.mark.parametrize("name", list(H.SCORING_VARIANT_FILES))
def test_extra_scoring_variants(engine, oracle_lib, name, tag):
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    g = H.load_golden(name)
    if g is None:
        pytest.skip("golden missing")
    var = H.SCORING_VARIANTS_EXTRA[tag]
    cfg = H.scoring_config(**var["config"]).to_struct(quad_sigma=var["quad_sigma"], quad_delta_mu=var["quad_delta_mu"])
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    got = engine.score_candidates(draw, dlib, cfg, cin)
    ref = (oracle_lib.score_candidates_4d if draw.is_4d else oracle_lib.score_candidates)(raw, lib, cfg, cin)
    assert_scores_close(got, ref, what=f"{name}/{tag}")
    dlib.close(); draw.close()


@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_extra_ragged_frames(engine, oracle_lib, name):
    from alphadia_b200.library import assemble_library_arrays

    raw, pdf0, fdf0, _, p = H.workload(name)
    pdf, fdf = H.ragged_library_frames(pdf0, fdf0, float(np.max(raw.rt_values)))
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    is4d = name == "parity_4d"
    cfg = _sel_cfg_4d(p) if is4d else H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = (oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates)(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    m = got["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    scfg = H.scoring_config().to_struct()
    s_got = engine.score_candidates(draw, dlib, scfg, cin)
    s_ref = (oracle_lib.score_candidates_4d if is4d else oracle_lib.score_candidates)(raw, lib, scfg, cin)
    assert_scores_close(s_got, s_ref, what=f"ragged frames/{name}")
    dlib.close(); draw.close()


@pytest.mark.parametrize("name", ["parity_f20", "parity_4d_f20"])
def test_extra_twenty_fragment_library(engine, oracle_lib, name):
    """20 library fragments per precursor (20 selection layers; scoring keeps the top 12 or 6 by library intensity), 3-D and 4-D."""
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    is4d = draw.is_4d
    cfg = _sel_cfg_4d(p) if is4d else H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = (oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates)(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    m = got["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    score = oracle_lib.score_candidates_4d if is4d else oracle_lib.score_candidates
    for variant in ("default", "legacy", "k6"):
        scfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
        assert_scores_close(engine.score_candidates(draw, dlib, scfg, cin), score(raw, lib, scfg, cin), what=f"{name}/{variant}")
    dlib.close(); draw.close()


def test_extra_two_isotope_library(engine, oracle_lib):
    """Library with two isotope columns under top_k_precursors = top_k_isotopes = 3 (oracle pinned in tests/golden/iso2.npz)."""
    from alphadia_b200.library import assemble_library_arrays

    raw, pdf0, fdf, _, p = H.workload("parity_small")
    lib = assemble_library_arrays(pdf0.drop(columns=["i_2", "i_3"]), fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
    cfg, kernel = H.selection_config(p["rt_tolerance"]).to_struct(), H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    assert_candidates_equal(got, oracle_lib.select_candidates(raw, lib, cfg, kernel))
    m = got["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    scfg = H.scoring_config().to_struct()
    assert_scores_close(engine.score_candidates(draw, dlib, scfg, cin), oracle_lib.score_candidates(raw, lib, scfg, cin), what="iso2")
    dlib.close(); draw.close()


def _assert_ragged_equals_dense(rag, dense, K):
    """adb_score_candidates_ragged == the rows / slots OutputPsmDF.to_precursor_df / to_fragment_df keep of the dense tables."""
    v = dense["valid"].astype(bool)
    assert rag["n_rows"] == int(v.sum())
    assert np.array_equal(rag["row_index"], np.nonzero(v)[0])
    assert np.array_equal(rag["features"], dense["features"][v], equal_nan=True)
    m = (dense["fragment_mz_library"] > 0) & v[:, None]
    assert rag["n_fragments"] == int(m.sum())
    counts = m.sum(axis=1)[v]
    assert np.array_equal(rag["frag_offset"], np.concatenate([[0], np.cumsum(counts)]))
    for k in FRAG_F32 + FRAG_U8:
        assert np.array_equal(rag[k], dense[k][m], equal_nan=True), k


@pytest.mark.parametrize("name", ["parity_small", "parity_4d", "parity_f20"])
def test_ragged_scores_equal_dense_tables(engine, oracle_lib, name):
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    is4d = draw.is_4d
    cfg = _sel_cfg_4d(p) if is4d else H.selection_config(p["rt_tolerance"]).to_struct()
    sel = engine.select_candidates(draw, dlib, cfg, H.default_kernel(raw))
    m = sel["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: sel[c][m] for c in INT_COLS})
    for variant in ("default", "k6"):
        scfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
        dense = engine.score_candidates(draw, dlib, scfg, cin)
        rag = engine.score_candidates_ragged(draw, dlib, scfg, cin)
        assert 0 < rag["n_rows"] < int(cin.n)
        _assert_ragged_equals_dense(rag, dense, int(scfg.top_k_fragments))
    # capacity too small: loud failure that names the needed sizes
    from alphadia_b200 import _abi
    rd, bufs = _abi.alloc_scores_ragged(int(cin.n), 8)
    with pytest.raises(RuntimeError, match="capacity"):
        engine.score_candidates_ragged(draw, dlib, scfg, cin, bufs=bufs)
    dlib.close(); draw.close()


def test_ragged_scores_chunked(engine):
    """>= 200k candidates: row blocks are compacted and copied while the next block is scored; same result as the dense call."""
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_small")
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    n = engine.select_candidates_resident(draw, dlib, cfg, H.default_kernel(raw))
    table = engine.fetch_candidate_table(draw, n)
    reps = 200000 // n + 1
    big = {k: np.ascontiguousarray(np.tile(v[:n], reps)) for k, v in table.items()}
    from alphadia_b200 import _abi
    N = n * reps
    scfg = H.scoring_config().to_struct()
    cin = _abi.candidates_in_from_table(big, N)
    dense = engine.score_candidates(draw, dlib, scfg, cin)
    rag = engine.score_candidates_ragged(draw, dlib, scfg, cin)
    _assert_ragged_equals_dense(rag, dense, int(scfg.top_k_fragments))
    dlib.close(); draw.close()


def test_top_k_fragments_9999_transfer_requantification(engine, oracle_lib):
    """top_k_fragments = 9999 (transfer_library_requantification_handler.py:102-124): every library fragment is quantified.
    Ragged device result vs the oracle's dense tables at the library's width (20) and vs the live reference's tables."""
    name = "parity_f20"
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    g = H.load_golden(name)
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    wide = int(np.max(lib["frag_stop_idx"] - lib["frag_start_idx"]))
    assert wide == 20
    ref = oracle_lib.score_candidates(raw, lib, H.scoring_config(top_k_fragments=wide).to_struct(), cin)
    rag = engine.score_candidates_ragged(draw, dlib, H.scoring_config(top_k_fragments=9999).to_struct(), cin, max_fragments=wide)
    v = ref["valid"].astype(bool)
    assert np.array_equal(rag["row_index"], np.nonzero(v)[0])
    F, G = rag["features"], ref["features"][v]
    floor = feature_scale_floor(G)
    err = np.abs(F - G) / np.maximum(np.maximum(np.abs(F), np.abs(G)), floor[None, :])
    assert np.where(np.isnan(F) & np.isnan(G), 0.0, err).max() < RTOL
    m = (ref["fragment_mz_library"] > 0) & v[:, None]
    assert rag["n_fragments"] == int(m.sum()) and m.sum(axis=1).max() > 12
    for k in FRAG_U8:
        assert np.array_equal(rag[k], ref[k][m]), k
    for k in FRAG_F32:
        assert H.rel_err(rag[k], ref[k][m]).max() < RTOL, k
    g9 = H.load_golden("k9999")
    if g9 is not None and str(g9["input_checksum"]) == str(g["input_checksum"]):
        assert np.array_equal(keep["precursor_idx"][rag["row_index"]], g9["feat_precursor_idx"])
        assert np.array_equal(keep["rank"][rag["row_index"]], g9["feat_rank"])
        G = g9["feat_matrix"]
        err = np.abs(F - G) / np.maximum(np.maximum(np.abs(F), np.abs(G)), feature_scale_floor(G)[None, :])
        assert np.where(np.isnan(F) & np.isnan(G), 0.0, err).max() < RTOL
        assert np.array_equal(rag["fragment_mz_library"], g9["frag_mz_library"])
        assert np.array_equal(rag["fragment_number"], g9["frag_number"])
        assert H.rel_err(rag["fragment_intensity"], g9["frag_intensity"]).max() < RTOL
    # the dense tables cannot hold 9999 slots per candidate: loud refusal
    with pytest.raises(RuntimeError, match="ragged"):
        engine.score_candidates(draw, dlib, H.scoring_config(top_k_fragments=64).to_struct(), cin)
    dlib.close(); draw.close()


@pytest.mark.parametrize("dtype_rt,dtype_mz", [(np.float32, np.float32), (np.float64, np.float64)])
def test_fragcomp_conflict_graph_equals_window_serial_kernel_and_oracle(engine, oracle_lib, monkeypatch, dtype_rt, dtype_mz):
    """The conflict-graph formulation (RT-sorted windows, all pairs in parallel, greedy pass over the PSMs with an edge) against
    the window-serial kernel and the oracle: 40 000 PSMs in 75 windows with planted shadows, uncovered PSMs between windows,
    an empty window, NaN retention times and PSMs that start out invalid."""
    import bench

    w = bench.fragcomp_workload(n_psm=40_000, seed=3)
    rt = w["rt"].astype(dtype_rt)
    rt[::997] = np.nan
    mz = w["fragment_mz"].astype(dtype_mz)
    ws, we = w["window_start"].copy(), w["window_stop"].copy()
    we[10] = ws[10]            # empty window
    ws[20] += 5; we[30] -= 7   # PSMs that belong to no window keep their flag
    args = (ws, we, rt, w["frag_start"], w["frag_stop"], mz, 3.0, 15.0)
    ref = oracle_lib.fragment_competition(*args).astype(bool)
    got = engine.fragment_competition(*args)
    monkeypatch.setenv("ADB_FRAGCOMP_SERIAL", "1")
    serial = engine.fragment_competition(*args)
    assert 1000 < (~ref).sum() < 10000
    assert np.array_equal(got, ref) and np.array_equal(serial, ref)


def test_classifier_predict_proba_matches_reference(engine):
    """adb_classifier_predict_proba with the weights the live reference trained (tests/golden/classifier_small.npz): within 1e-4
    of the reference's predict_proba and of the numpy oracle, same classes; then fit on the device and check it learns."""
    import hashlib

    import oracle
    from alphadia_b200.classifier import BinaryClassifierLegacyNewBatching

    g = H.load_golden("classifier_small")
    x, y = H.classifier_inputs()
    if g is None or str(g["input_checksum"]) != hashlib.sha256(x.tobytes() + y.tobytes()).hexdigest():
        pytest.skip("golden not applicable")
    state = {k[3:]: g[k] for k in g.files if k.startswith("w__")}
    clf = BinaryClassifierLegacyNewBatching()
    clf.from_state_dict({"input_dim": int(g["input_dim"]), "output_dim": 2, "layers": [int(v) for v in g["layers"]], "dropout": 0.001,
                         "network_state_dict": state})
    p = clf.predict_proba(x)
    assert p.shape == g["proba"].shape and p.dtype == np.float32
    assert np.abs(p - g["proba"]).max() < 1e-5 and H.rel_err(p[:, 1], g["proba"][:, 1], floor=1e-3).max() < RTOL
    assert np.abs(p - oracle.classifier_predict_proba(state, x)).max() < 1e-5
    assert np.array_equal(clf.predict(x), g["predict"])
    assert clf.predict_proba(x[:0]).shape == (0, 2)
    # a wide input and odd layer widths (padding of the 4-output groups)
    rng = np.random.default_rng(0)
    dims = [128, 37, 6, 3]
    st = {"fc_layers.0.weight": rng.random(dims[0]).astype(np.float32) + 0.5, "fc_layers.0.bias": rng.normal(size=dims[0]).astype(np.float32),
          "fc_layers.0.running_mean": rng.normal(size=dims[0]).astype(np.float32), "fc_layers.0.running_var": rng.random(dims[0]).astype(np.float32) + 0.1}
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        st[f"fc_layers.{1 + 3 * i}.weight"] = (rng.normal(size=(b, a)) / np.sqrt(a)).astype(np.float32)
        st[f"fc_layers.{1 + 3 * i}.bias"] = rng.normal(size=b).astype(np.float32)
    xx = rng.normal(size=(1000, dims[0])).astype(np.float32)
    from alphadia_b200.classifier import network_forward_device
    assert np.abs(network_forward_device(st, xx) - oracle.classifier_predict_proba(st, xx)).max() < 1e-5
    # training on the device: the same recipe learns the planted separation
    fit = BinaryClassifierLegacyNewBatching(test_size=0.001, batch_size=500, learning_rate=0.001, epochs=3, random_state=3)
    fit.fit(x, y)
    assert fit.fitted and np.mean(fit.predict(x) == y) > 0.97
    assert len(fit.metrics["train_loss"]) >= 1


@pytest.mark.parametrize("workload,n_prec", [("config3", 700), ("config4", 600)])
def test_full_size_benchmark_workloads_vs_oracle(engine, oracle_lib, workload, n_prec):
    """The two benchmarked workloads at their FULL size (config 3: 2 M precursors, 3-D; config 4: 200 k precursors, timsTOF,
    80 scans x 112 cycles per search window): one resident step on the device, then the oracle selects and scores a random
    subsample of the library against the same raw file (>= 1 500 candidate rows) - candidate container rows bit for bit, valid
    mask, 46 features and the fragment tables within 1e-4 (the check bench.py attaches to its timed results)."""
    import bench
    from alphadia_b200.engine import HotPath

    raw, pdf, fdf, lib, p, sel, sc, kernel = bench.build_workload(workload, 0, None)
    hp = HotPath(raw, lib, sel, sc, kernel, device=0)
    try:
        stats = hp.resident_step()
        assert stats["n_candidates"] > 0.9 * hp.n_precursors * sel.candidate_count
        cont = engine.fetch_candidates(hp.dev_raw, int(hp.n_precursors * sel.candidate_count))
        r = bench.parity_spot_check(hp, raw, lib, sel, sc, kernel, cont, n_prec=n_prec)
    finally:
        hp.close()
    assert r["n"] >= 1500 and r["valid"] > 0.5 * r["n"], r
    assert r["int_exact"] and r["selection_score_bit_exact"] and r["valid_exact"], r
    assert r["max_rel"] < RTOL and r["fragment_table_max_rel"] < 3 * RTOL, r


@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_fused_select_score_equals_two_calls(engine, oracle_lib, name):
    """adb_select_score_candidates_ragged (selection -> score cutoff -> scoring from the resident table, one call) against the
    two-call flow select_candidates_resident / fetch_candidate_table / score_candidates_ragged, with and without a cutoff."""
    from alphadia_b200 import _abi

    raw, lib, p, draw, dlib = _device_objects(engine, name)
    cfg = _sel_cfg_4d(p) if draw.is_4d else H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    scfg = H.scoring_config().to_struct()
    n = engine.select_candidates_resident(draw, dlib, cfg, kernel)
    table = engine.fetch_candidate_table(draw, n)
    ref = engine.score_candidates_ragged(draw, dlib, scfg, _abi.candidates_in_from_table(table, n))
    cap = int(len(lib["precursor_idx"]) * cfg.candidate_count)
    for cutoff in (float("-inf"), float(np.median(table["score"][:n]))):
        t2 = _abi.alloc_candidate_table(cap)
        _, bufs = _abi.alloc_scores_ragged(cap, cap * 12)
        got = engine.select_score_candidates_ragged(draw, dlib, cfg, kernel, scfg, t2, bufs, score_cutoff=cutoff)
        keep = table["score"][:n] > np.float32(cutoff)
        assert got["n_candidates"] == int(keep.sum()) and 0 < got["n_candidates"] <= n
        for k, v in table.items():
            assert np.array_equal(t2[k][:got["n_candidates"]], v[:n][keep]), k
        # rows of the filtered table that are valid == the valid rows of the full table that pass the cutoff
        old_rows = np.flatnonzero(keep)[got["row_index"]]
        m = keep[ref["row_index"]]
        assert np.array_equal(old_rows, ref["row_index"][m])
        assert np.array_equal(got["features"], ref["features"][m], equal_nan=True)
        counts = np.diff(ref["frag_offset"])
        fm = np.repeat(m, counts)
        for k in FRAG_F32 + FRAG_U8:
            assert np.array_equal(got[k], ref[k][fm], equal_nan=True), k
    dlib.close(); draw.close()


def test_ragged_scores_empty_and_all_invalid(engine):
    """Zero candidates, and candidates that are all rejected (windows of one cycle): the ragged result is empty and well-formed."""
    raw, lib, p, draw, dlib = _device_objects(engine, "parity_small")
    scfg = H.scoring_config().to_struct()
    empty = {c: np.zeros(0, np.int64) for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, empty)
    rag = engine.score_candidates_ragged(draw, dlib, scfg, cin)
    assert rag["n_rows"] == 0 and rag["n_fragments"] == 0 and list(rag["frag_offset"]) == [0]
    L = raw.cycle.shape[1]
    one = dict(precursor_idx=lib["precursor_idx"][:50].astype(np.int64), rank=np.zeros(50, np.int64), scan_center=np.zeros(50, np.int64),
               scan_start=np.zeros(50, np.int64), scan_stop=np.ones(50, np.int64), frame_center=np.full(50, 10 * L, np.int64),
               frame_start=np.full(50, 10 * L, np.int64), frame_stop=np.full(50, 10 * L, np.int64))  # zero cycles wide
    cin, keep = H.candidates_in_from_arrays(lib, one)
    rag = engine.score_candidates_ragged(draw, dlib, scfg, cin)
    dense = engine.score_candidates(draw, dlib, scfg, cin)
    assert not dense["valid"].any() and rag["n_rows"] == 0 and rag["n_fragments"] == 0
    dlib.close(); draw.close()


@pytest.mark.parametrize("name", ["parity_f48", "parity_f96"])
def test_wide_fragment_library_ragged(engine, oracle_lib, name):
    """48 / 96 library fragments per precursor: selection (plan over all of them, top 12 layers) bit-exact, ragged scoring with
    top_k_fragments = 9999 (every fragment kept, more than the dense tables' 32 slots) against the oracle at the same width
    and, for 48, the live reference's tables (tests/golden/k9999_f48.npz)."""
    wide = int(name[len("parity_f"):])
    raw, lib, p, draw, dlib = _device_objects(engine, name)
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    got = engine.select_candidates(draw, dlib, cfg, kernel)
    ref = oracle_lib.select_candidates(raw, lib, cfg, kernel)
    assert_candidates_equal(got, ref)
    m = got["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: got[c][m] for c in INT_COLS})
    for kw in (dict(), dict(quant_all=False, experimental_xic=False)):
        o = oracle_lib.score_candidates(raw, lib, H.scoring_config(top_k_fragments=wide, **kw).to_struct(), cin)
        rag = engine.score_candidates_ragged(draw, dlib, H.scoring_config(top_k_fragments=9999, **kw).to_struct(), cin, max_fragments=wide)
        v = o["valid"].astype(bool)
        assert np.array_equal(rag["row_index"], np.nonzero(v)[0])
        F, G = rag["features"], o["features"][v]
        err = np.abs(F - G) / np.maximum(np.maximum(np.abs(F), np.abs(G)), feature_scale_floor(G)[None, :])
        assert np.where(np.isnan(F) & np.isnan(G), 0.0, err).max() < RTOL
        fm = (o["fragment_mz_library"] > 0) & v[:, None]
        assert rag["n_fragments"] == int(fm.sum()) and fm.sum(axis=1).max() == wide
        for k in FRAG_U8:
            assert np.array_equal(rag[k], o[k][fm]), k
        for k in FRAG_F32:
            # the mass error (ppm) is a difference of two nearly equal m/z values: relative to 0.01 ppm at least
            floor = 1e-2 if k == "fragment_mass_error" else 1e-6
            assert H.rel_err(rag[k], o[k][fm], floor=floor).max() < RTOL, k
    g = H.load_golden("k9999_f48") if wide == 48 else None
    if g is not None and str(g["input_checksum"]) == H.input_checksum(*H.workload("parity_f48")[:3]):
        rag = engine.score_candidates_ragged(draw, dlib, H.scoring_config(top_k_fragments=9999).to_struct(), cin, max_fragments=48)
        assert np.array_equal(keep["precursor_idx"][rag["row_index"]], g["feat_precursor_idx"])
        assert np.array_equal(rag["fragment_mz_library"], g["frag_mz_library"]) and np.array_equal(rag["fragment_number"], g["frag_number"])
        assert H.rel_err(rag["fragment_intensity"], g["frag_intensity"]).max() < RTOL
    dlib.close(); draw.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [201, 202])
def test_randomized_sweep_device_vs_oracle(engine, oracle_lib, seed):
    """The generator of tests/test_hostsim.py::test_hostsim_randomized_sweep on the device: random raw files, libraries (40 % ragged),
    selection and scoring configurations, quadrupole parameters — candidate table bit-exact, scores within the tolerance."""
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.synthetic import make_config_3d, make_config_4d

    rng = np.random.default_rng(seed)
    compared = 0
    for it in range(8):
        is4d = it % 4 == 3
        s = int(rng.integers(1, 10**6))
        if is4d:
            name = str(rng.choice(["parity_4d", "parity_4d_overlap"]))
            raw, pdf, fdf, p = make_config_4d(name, seed=s, n_precursors=int(rng.integers(60, 200)))
        else:
            name = str(rng.choice(["parity_small", "parity_f20", "config1"]))
            raw, pdf, fdf, p = make_config_3d(name, seed=s, n_precursors=int(rng.integers(80, 300)), scale_noise=float(rng.choice([0.3, 1.0, 3.0])))
        if rng.random() < 0.4:
            pdf, fdf = H.ragged_library_frames(pdf, fdf, float(np.max(raw.rt_values)), seed=s)
        lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
        skw = dict(candidate_count=int(rng.integers(1, 6)))
        if is4d:
            skw["mobility_tolerance"] = p["mobility_tolerance"]
        selcfg = H.selection_config(p["rt_tolerance"] * float(rng.choice([0.5, 1.0, 2.0])), **skw).to_struct()
        kernel = H.default_kernel(raw, fwhm_rt=float(rng.choice([2.0, 5.0, 10.0])))
        draw, dlib = engine.DeviceRawFile(raw, device=0), engine.DeviceLibrary(lib, device=0)
        sel = engine.select_candidates(draw, dlib, selcfg, kernel)
        osel = oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates
        oscore = oracle_lib.score_candidates_4d if is4d else oracle_lib.score_candidates
        assert_candidates_equal(sel, osel(raw, lib, selcfg, kernel))
        m = sel["score"] > 0
        if m.sum():
            cin, keep = H.candidates_in_from_arrays(lib, {c: sel[c][m] for c in INT_COLS})
            kw = dict(top_k_fragments=int(rng.choice([3, 6, 12, 20, 32])), top_k_isotopes=int(rng.integers(1, 5)),
                      quant_window=int(rng.integers(1, 5)), quant_all=bool(rng.integers(0, 2)), experimental_xic=bool(rng.integers(0, 2)),
                      precursor_mz_tolerance=float(rng.choice([2, 5, 15, 50])), fragment_mz_tolerance=float(rng.choice([3, 10, 30, 100])))
            cfg = H.scoring_config(**kw).to_struct(quad_sigma=(float(rng.choice([0.2, 0.5, 1.0])), float(rng.choice([0.2, 0.8]))),
                                                  quad_delta_mu=(float(rng.choice([0.0, 0.3, -0.5])), float(rng.choice([0.0, -0.4]))))
            got = engine.score_candidates(draw, dlib, cfg, cin)
            ref = oscore(raw, lib, cfg, cin)
            if ref["valid"].sum() == 0:
                assert got["valid"].sum() == 0
            else:
                assert_scores_close(got, ref, what=f"seed {seed} case {it} {name} {kw}")
                compared += 1
        dlib.close(); draw.close()
    assert compared >= 4

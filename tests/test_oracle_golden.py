"""Pin the C oracle against golden vectors produced by the UNMODIFIED reference numba path
(tests/golden/generate_golden.py, run in the build container).  CPU only."""

import numpy as np
import pytest

from tests import helpers as H

INT_COLS = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]
FRAG_MAP = {
    "mz_library": "fragment_mz_library", "mz": "fragment_mz", "mz_observed": "fragment_mz_observed",
    "height": "fragment_height", "intensity": "fragment_intensity", "mass_error": "fragment_mass_error",
    "correlation": "fragment_correlation", "position": "fragment_position", "number": "fragment_number",
    "type": "fragment_type", "charge": "fragment_charge", "loss_type": "fragment_loss_type",
}
VARIANTS = {
    "": {},
    "_legacy": dict(quant_all=False, experimental_xic=False),
    "_k6": dict(top_k_fragments=6, top_k_isotopes=4, quant_window=2),
}
# features whose reference value goes through BLAS dots (np.dot / np.corrcoef): summation order is
# unspecified there, everything else must match the reference bit for bit.
BLAS_FEATURES = {18, 19, 29, 30, 31, 32, 33, 34, 36}


def _golden(name):
    g = H.load_golden(name)
    if g is None:
        pytest.skip(f"golden {name} missing")
    raw, pdf, fdf, lib, p = H.workload(name)
    if str(g["input_checksum"]) != H.input_checksum(raw, pdf, fdf):
        pytest.skip("synthetic generator output differs from the one the golden file was made with (numpy version?)")
    return g, raw, lib, p


@pytest.mark.parametrize("name", ["config1", "parity_small"])
def test_gaussian_kernel_matches_reference(name):
    g, raw, lib, p = _golden(name)
    k = H.default_kernel(raw)
    assert k.dtype == np.float32 and k.shape == g["sel_kernel"].shape
    np.testing.assert_allclose(k, g["sel_kernel"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["config1", "parity_small", "parity_f20"])
def test_selection_bit_exact_vs_reference(name, oracle_lib):
    g, raw, lib, p = _golden(name)
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    arrs = oracle_lib.select_candidates(raw, lib, cfg, g["sel_kernel"])
    m = arrs["score"] > 0
    assert m.sum() == len(g["cand_precursor_idx"])
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g["cand_" + c].astype(np.int64)), c
    assert np.array_equal(arrs["score"][m], g["cand_score"])  # f32, bit-exact


@pytest.mark.parametrize("name,tag", [("config1", ""), ("parity_small", ""), ("parity_small", "_legacy"), ("parity_small", "_k6"),
                                      ("parity_f20", ""), ("parity_f20", "_legacy"), ("parity_f20", "_k6")])
def test_scoring_vs_reference(name, tag, oracle_lib):
    g, raw, lib, p = _golden(name)
    cand = {c: g["cand_" + c] for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    cfg = H.scoring_config(**VARIANTS[tag]).to_struct()
    arrs = oracle_lib.score_candidates(raw, lib, cfg, cin)
    v = arrs["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g[f"feat{tag}_precursor_idx"])
    assert np.array_equal(keep["rank"][v], g[f"feat{tag}_rank"])
    F, G = arrs["features"][v], g[f"feat{tag}_matrix"]
    for j in range(46):
        same = (F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j  # tolerance of BASELINE.json north_star
        else:
            assert same.all(), f"feature {j} not bit-exact"
    m = arrs["fragment_mz_library"] > 0
    assert m.sum() == len(g[f"frag{tag}_mz_library"])
    for k, v2 in FRAG_MAP.items():
        a, b = arrs[v2][m], g[f"frag{tag}_{k}"]
        if k == "correlation":
            assert H.rel_err(a, b).max() < 1e-4
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("name,golden", [("parity_f20", "k9999"), ("parity_f48", "k9999_f48")])
def test_scoring_top_k_9999_vs_reference(oracle_lib, name, golden):
    """Transfer-library requantification (transfer_library_requantification_handler.py:102-124, top_k_fragments = 9999): the
    live reference quantifies every library fragment (20 / 48 per precursor here - 48 is more than the dense device tables
    hold).  A candidate keeps at most the fragments its precursor has, so the oracle at the library's width must reproduce the
    reference's tables (tests/golden/k9999.npz, k9999_f48.npz)."""
    g9 = H.load_golden(golden)
    raw, pdf, fdf, lib, p = H.workload(name)
    assert g9 is not None and str(g9["input_checksum"]) == H.input_checksum(raw, pdf, fdf)
    g = g9 if "cand_precursor_idx" in g9.files else H.load_golden(name)
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    wide = int(np.max(lib["frag_stop_idx"] - lib["frag_start_idx"]))
    arrs = oracle_lib.score_candidates(raw, lib, H.scoring_config(top_k_fragments=wide).to_struct(), cin)
    v = arrs["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g9["feat_precursor_idx"])
    assert np.array_equal(keep["rank"][v], g9["feat_rank"])
    F, G = arrs["features"][v], g9["feat_matrix"]
    for j in range(46):
        same = (F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert same.all(), f"feature {j} not bit-exact"
    m = arrs["fragment_mz_library"] > 0
    assert m.sum() == len(g9["frag_mz_library"]) and m.sum(axis=1).max() == wide
    for k, v2 in FRAG_MAP.items():
        a, b = arrs[v2][m], g9[f"frag_{k}"]
        if k == "correlation":
            assert H.rel_err(a, b).max() < 1e-4
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("name", ["config1", "parity_small", "parity_f20"])
def test_fragcomp_vs_reference(name, oracle_lib):
    """FragmentCompetition on the golden feature table with the golden pseudo-proba."""
    from alphadia_b200.fragcomp import FragmentCompetition, plan_fragment_competition

    g, raw, lib, p = _golden(name)
    import pandas as pd

    psm = pd.DataFrame({"precursor_idx": g["feat_precursor_idx"], "rank": g["feat_rank"],
                        "rt_observed": g["feat_matrix"][:, 2], "mz_observed": g["feat_matrix"][:, 10],
                        "proba": g["fc_proba"]})
    frag = pd.DataFrame({"precursor_idx": g["frag_precursor_idx"], "rank": g["frag_rank"],
                         "mz_observed": g["frag_mz_observed"]})
    plan = plan_fragment_competition(psm, frag, raw.cycle)
    valid = oracle_lib.fragment_competition(plan.window_start, plan.window_stop, plan.rt, plan.frag_start, plan.frag_stop,
                                            plan.fragment_mz, 3, 15)
    kept = plan.psm_df[valid]
    assert np.array_equal(kept["precursor_idx"].values, g["fc_kept_precursor_idx"])
    assert np.array_equal(kept["rank"].values, g["fc_kept_rank"])
    assert np.array_equal(kept["_candidate_idx"].values, g["fc_kept_candidate_idx"])
    assert FragmentCompetition is not None


@pytest.mark.parametrize("tag", list(H.FRAGCOMP_DTYPES))
def test_fragcomp_dense_vs_reference(tag, oracle_lib):
    """512 of 6000 PSMs lose the competition in this table (the synthetic runs above have almost no collisions): the greedy
    veto in probability order (fragcomp.py:110-143), compared with the reference's mask for every rt / m/z dtype pair."""
    import hashlib
    import os

    path = os.path.join(H.GOLDEN_DIR, "fragcomp_dense.npz")
    if not os.path.exists(path):
        pytest.skip("golden fragcomp_dense.npz missing")
    g = np.load(path, allow_pickle=False)
    ws, we, rt, fs, fe, mz = H.fragcomp_dense_inputs(*H.FRAGCOMP_DTYPES[tag])
    if str(g[f"{tag}__checksum"]) != hashlib.sha256(rt.tobytes() + mz.tobytes()).hexdigest():
        pytest.skip("inputs differ from the ones the golden file was made with (numpy version?)")
    valid = oracle_lib.fragment_competition(ws, we, rt, fs, fe, mz, 3, 15).astype(bool)
    assert 100 < (~g[f"{tag}__valid"]).sum() < 3000
    assert np.array_equal(valid, g[f"{tag}__valid"])


# ---- timsTOF (4-D) -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["parity_4d", "parity_4d_overlap", "parity_4d_f20"])
def test_selection_4d_bit_exact_vs_reference(name, oracle_lib):
    g, raw, lib, p = _golden(name)
    k = H.default_kernel(raw)
    assert k.shape == (30, 30) and np.array_equal(k, g["sel_kernel"])
    cfg = H.selection_config(p["rt_tolerance"], mobility_tolerance=p["mobility_tolerance"]).to_struct()
    arrs = oracle_lib.select_candidates_4d(raw, lib, cfg, g["sel_kernel"])
    m = arrs["score"] > 0
    assert m.sum() == len(g["cand_precursor_idx"]) > 100
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g["cand_" + c].astype(np.int64)), c
    assert np.array_equal(arrs["score"][m], g["cand_score"])
    # the scan window really is two-dimensional here
    assert (g["cand_scan_stop"] - g["cand_scan_start"]).max() > 8


@pytest.mark.parametrize("tag", ["", "_legacy", "_k6"])
@pytest.mark.parametrize("name", ["parity_4d", "parity_4d_overlap", "parity_4d_f20"])
def test_scoring_4d_vs_reference(name, tag, oracle_lib):
    g, raw, lib, p = _golden(name)
    if f"feat{tag}_matrix" not in g:
        pytest.skip(f"golden {name} has no {tag} variant")
    if name == "parity_4d_overlap":  # neighbouring windows overlap: candidates seen by two frames of the cycle
        n_obs = g[f"feat{tag}_matrix"][:, 17]
        assert (n_obs == 2).sum() >= 10 and (n_obs == 1).sum() >= 10
    cand = {c: g["cand_" + c] for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    arrs = oracle_lib.score_candidates_4d(raw, lib, H.scoring_config(**VARIANTS[tag]).to_struct(), cin)
    v = arrs["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g[f"feat{tag}_precursor_idx"])
    assert np.array_equal(keep["rank"][v], g[f"feat{tag}_rank"])
    F, G = arrs["features"][v], g[f"feat{tag}_matrix"]
    assert np.abs(G[:, 29]).max() > 0 and np.abs(G[:, 39]).max() > 0  # mobility features are live
    for j in range(46):
        same = (F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert same.all(), f"feature {j} not bit-exact"
    m = arrs["fragment_mz_library"] > 0
    assert m.sum() == len(g[f"frag{tag}_mz_library"])
    for k, v2 in FRAG_MAP.items():
        a, b = arrs[v2][m], g[f"frag{tag}_{k}"]
        assert np.array_equal(a, b) if k != "correlation" else H.rel_err(a, b).max() < 1e-4, k


# ---- configuration variants and the multiplexed set-up (tests/golden/variants.npz) ---------------------
def _variants(name):
    import os

    path = os.path.join(H.GOLDEN_DIR, "variants.npz")
    if not os.path.exists(path):
        pytest.skip("golden variants.npz missing")
    g = np.load(path, allow_pickle=False)
    raw, pdf, fdf, lib, p = H.workload(name)
    if str(g[f"{name}__input_checksum"]) != H.input_checksum(raw, pdf, fdf):
        pytest.skip("synthetic generator output differs from the one the golden file was made with (numpy version?)")
    return g, raw, pdf, fdf, lib, p


@pytest.mark.parametrize("name,tag", [(n, t) for n, v in H.SELECTION_VARIANTS.items() for t in v])
def test_selection_variants_bit_exact_vs_reference(name, tag, oracle_lib):
    g, raw, pdf, fdf, lib, p = _variants(name)
    kw = dict(H.SELECTION_VARIANTS[name][tag])
    rt_tol = kw.pop("rt_tolerance", p["rt_tolerance"])
    if "mobility_tolerance" in p:
        kw.setdefault("mobility_tolerance", p["mobility_tolerance"])
    cfg = H.selection_config(rt_tol, **kw).to_struct()
    select = oracle_lib.select_candidates_4d if "mobility_tolerance" in p else oracle_lib.select_candidates
    arrs = select(raw, lib, cfg, H.default_kernel(raw))
    m = arrs["score"] > 0
    key = f"{name}__{tag}__cand_"
    assert m.sum() == len(g[key + "precursor_idx"]) > 0
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g[key + c].astype(np.int64)), c
    if tag == "unweighted":
        # amean1 / astd1 (selection/utils.py:113-126) are compiled as callees of the fastmath `_build_candidates`
        # (selection.py:367) and inherit its flag: their f32 sum is vector-reduced in an order that depends on the
        # host's SIMD width (compiled on their own they are sequential and equal to the oracle bit for bit).  The
        # f32 score is therefore compared within the float tolerance; observed difference <= 6e-6.
        assert H.rel_err(arrs["score"][m], g[key + "score"]).max() < 1e-4
    else:
        assert np.array_equal(arrs["score"][m], g[key + "score"])  # f32, bit-exact


@pytest.mark.parametrize("tag", list(H.SELECTION_VARIANTS2))
@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_selection_variants2_bit_exact_vs_reference(name, tag, oracle_lib):
    """Peak-limit parameters (symetric_limits_2d, selection/utils.py:276-312) and other widths of the smoothing kernel
    (selection/kernel.py:98-218): the kernel matrix and the candidate table against the live reference."""
    import os

    from alphadia_b200.kernel import GaussianKernel

    path = os.path.join(H.GOLDEN_DIR, "variants2.npz")
    if not os.path.exists(path):
        pytest.skip("golden variants2.npz missing")
    g = np.load(path, allow_pickle=False)
    raw, pdf, fdf, lib, p = H.workload(name)
    if str(g[f"{name}__input_checksum"]) != H.input_checksum(raw, pdf, fdf):
        pytest.skip("golden not applicable")
    kw = dict(H.SELECTION_VARIANTS2[tag])
    fwhm_rt, fwhm_mobility = kw.pop("fwhm_rt", 5.0), kw.pop("fwhm_mobility", 0.01)
    if "mobility_tolerance" in p:
        kw.setdefault("mobility_tolerance", p["mobility_tolerance"])
    config = H.selection_config(p["rt_tolerance"], **kw)
    kernel = GaussianKernel(raw, fwhm_rt=fwhm_rt, sigma_scale_rt=config.sigma_scale_rt, fwhm_mobility=fwhm_mobility,
                            sigma_scale_mobility=config.sigma_scale_mobility, kernel_width=config.kernel_size,
                            kernel_height=min(config.kernel_size, raw.scan_max_index + 1)).get_dense_matrix(verbose=False)
    key = f"{name}__{tag}__"
    assert kernel.shape == g[key + "kernel"].shape and kernel.dtype == g[key + "kernel"].dtype
    np.testing.assert_allclose(kernel, g[key + "kernel"], rtol=1e-6, atol=0)
    select = oracle_lib.select_candidates_4d if "mobility_tolerance" in p else oracle_lib.select_candidates
    arrs = select(raw, lib, config.to_struct(), g[key + "kernel"])
    m = arrs["score"] > 0
    assert m.sum() == len(g[key + "cand_precursor_idx"]) > 0
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g[key + "cand_" + c].astype(np.int64)), c
    assert np.array_equal(arrs["score"][m], g[key + "cand_score"])


@pytest.mark.parametrize("tag,cfg_kw", [("ref0", dict(score_grouped=True, reference_channel=0)),
                                        ("grouped", dict(score_grouped=True, reference_channel=-1))])
def test_multiplexed_scoring_vs_reference(tag, cfg_kw, oracle_lib, monkeypatch):
    """score_grouped / reference_channel (multiplexing_requantification_handler.py:120-149) through the HOST side of
    CandidateScoring — score groups, the skip of groups without the reference channel, the order and content of both
    result tables — against the live reference.  There is no GPU here, so the one device call is replaced by the oracle
    for this test only; tests/test_gpu_parity.py::test_multiplexed_score_groups_and_reference_channel runs the real one."""
    import pandas as pd

    from alphadia_b200 import _lib
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS, CandidateScoring

    g, raw, pdf, fdf, lib, p = _variants("parity_small")
    mpdf = H.multiplexed_library(pdf)
    mlib = assemble_library_arrays(mpdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    arrs = oracle_lib.select_candidates(raw, mlib, H.selection_config(p["rt_tolerance"], candidate_count=1).to_struct(),
                                        H.default_kernel(raw))
    m = arrs["score"] > 0
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g["mplex__cand_" + c].astype(np.int64)), c
    cand_df = pd.DataFrame({c: g["mplex__cand_" + c] for c in INT_COLS + ["score", "elution_group_idx", "decoy"]})

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

    monkeypatch.setattr(_lib, "device_rawfile_for", lambda dia_data, adapted: HostRaw(adapted))
    monkeypatch.setattr(_lib, "DeviceLibrary", HostLibrary)
    monkeypatch.setattr(_lib, "score_candidates_ragged", H.ragged_scoring_stub(oracle_lib.score_candidates))
    scorer = CandidateScoring(dia_data=raw, precursors_flat=mpdf.copy(), fragments_flat=fdf.copy(), config=H.scoring_config(**cfg_kw),
                              rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                              fragment_mz_column="mz_library")
    feat, frag = scorer(cand_df.copy())
    key = f"mplex__{tag}__"
    assert len(feat) == len(g[key + "feat_precursor_idx"]) > 50
    if tag == "ref0":
        assert len(feat) < len(g["mplex__grouped__feat_precursor_idx"])  # groups without channel 0 were skipped
    for c in ("precursor_idx", "rank", "elution_group_idx", "channel", "decoy"):
        assert np.array_equal(feat[c].values, g[key + "feat_" + c]), c  # same rows in the same order
    F, G = feat[DEFAULT_FEATURE_COLUMNS].values, g[key + "feat_matrix"]
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    for c in ("precursor_idx", "rank", "mz_library", "number", "intensity", "elution_group_idx", "decoy"):
        assert np.array_equal(frag[c].values, g[key + "frag_" + c]), c


# ---- timsTOF load-time CSR transpose --------------------------------------------------
def _transpose_inputs(seed=5, n_push=1500, n_tof=257):
    """Same generator as tests/golden/generate_golden.py::transpose_inputs."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 24, n_push)
    counts[rng.random(n_push) < 0.2] = 0
    push_indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    tof = np.concatenate([np.sort(rng.choice(n_tof, size=c, replace=False)) for c in counts] + [np.zeros(0, np.int64)]).astype(np.uint32)
    values = rng.integers(1, 60000, len(tof)).astype(np.uint16)
    return tof, push_indptr, n_tof, values


def test_transpose_oracle_vs_reference_golden(oracle_lib):
    import hashlib
    import os

    g = np.load(os.path.join(H.GOLDEN_DIR, "transpose_small.npz"), allow_pickle=False)
    tof, push_indptr, n_tof, values = _transpose_inputs()
    assert str(g["input_checksum"]) == hashlib.sha256(tof.tobytes() + push_indptr.tobytes() + values.tobytes()).hexdigest()
    push_indices, tof_indptr, new_values = oracle_lib.transpose_csr(tof, push_indptr, n_tof, values)
    assert np.array_equal(push_indices, g["push_indices"]) and push_indices.dtype == g["push_indices"].dtype
    assert np.array_equal(tof_indptr, g["tof_indptr"]) and tof_indptr.dtype == g["tof_indptr"].dtype
    assert np.array_equal(new_values, g["new_values"]) and new_values.dtype == g["new_values"].dtype
    # independent restatement: stable sort by tof index; and the round trip back to push-major order
    order = np.argsort(tof, kind="stable")
    push_of = np.repeat(np.arange(len(push_indptr) - 1, dtype=np.uint32), np.diff(push_indptr))
    assert np.array_equal(push_of[order], push_indices) and np.array_equal(values[order], new_values)
    tof_of = np.repeat(np.arange(n_tof, dtype=np.uint32), np.diff(tof_indptr))
    back = np.lexsort((tof_of, push_indices))
    assert np.array_equal(tof_of[back], tof) and np.array_equal(new_values[back], values)
    # ragged edges: no events at all, a single push, every event in one tof row
    p0, i0, v0 = oracle_lib.transpose_csr(np.zeros(0, np.uint32), np.zeros(5, np.int64), 7, np.zeros(0, np.uint16))
    assert len(p0) == 0 and np.array_equal(i0, np.zeros(8, np.int64))
    p1, i1, v1 = oracle_lib.transpose_csr(np.full(4, 3, np.uint32), np.array([0, 1, 1, 3, 4], np.int64), 5, np.array([9, 8, 7, 6], np.uint16))
    assert np.array_equal(p1, [0, 2, 2, 3]) and np.array_equal(i1, [0, 0, 0, 0, 4, 4]) and np.array_equal(v1, [9, 8, 7, 6])


# ---- ragged libraries (tests/golden/ragged.npz) ---------------------------------------------------------------
@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_ragged_library_vs_reference(name, oracle_lib):
    """0-12 fragments per precursor, shared ions, duplicate fragment m/z, retention times outside the run, charges 1-4:
    selection bit-exact and scoring at the usual bar against the live reference."""
    import os

    from alphadia_b200.library import assemble_library_arrays

    path = os.path.join(H.GOLDEN_DIR, "ragged.npz")
    if not os.path.exists(path):
        pytest.skip("golden ragged.npz missing")
    g = np.load(path, allow_pickle=False)
    raw, pdf0, fdf0, _, p = H.workload(name)
    if f"{name}__input_checksum" not in g or str(g[f"{name}__input_checksum"]) != H.input_checksum(raw, pdf0, fdf0):
        pytest.skip("golden not applicable")
    pdf, fdf = H.ragged_library_frames(pdf0, fdf0, float(np.max(raw.rt_values)))
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    is4d = "mobility_tolerance" in p
    kw = {"mobility_tolerance": p["mobility_tolerance"]} if is4d else {}
    cfg = H.selection_config(p["rt_tolerance"], **kw).to_struct()
    arrs = (oracle_lib.select_candidates_4d if is4d else oracle_lib.select_candidates)(raw, lib, cfg, H.default_kernel(raw))
    m = arrs["score"] > 0
    key = f"{name}__cand_"
    assert m.sum() == len(g[key + "precursor_idx"]) > 100
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g[key + c].astype(np.int64)), c
    assert np.array_equal(arrs["score"][m], g[key + "score"])
    cin, keep = H.candidates_in_from_arrays(lib, {c: g[key + c] for c in INT_COLS})
    sc = (oracle_lib.score_candidates_4d if is4d else oracle_lib.score_candidates)(raw, lib, H.scoring_config().to_struct(), cin)
    v = sc["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g[f"{name}__feat_precursor_idx"])
    assert np.array_equal(keep["rank"][v], g[f"{name}__feat_rank"])
    F, G = sc["features"][v], g[f"{name}__feat_matrix"]
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    fm = sc["fragment_mz_library"] > 0
    assert fm.sum() == len(g[f"{name}__frag_mz_library"])
    for k, v2 in FRAG_MAP.items():
        a, b = sc[v2][fm], g[f"{name}__frag_{k}"]
        assert np.array_equal(a, b) if k != "correlation" else H.rel_err(a, b).max() < 1e-4, k


# ---- hand-made candidate windows (tests/golden/edge.npz) ---------------------------------------------------------
@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_scoring_edge_windows_vs_reference(name, oracle_lib):
    """Windows of 1 to 241 cycles, clipped at the ends of the run, 1-scan to full-height scan windows, placed at random
    (mostly off any signal): which candidates the reference keeps and what it computes for them."""
    import os

    path = os.path.join(H.GOLDEN_DIR, "edge.npz")
    if not os.path.exists(path):
        pytest.skip("golden edge.npz missing")
    g = np.load(path, allow_pickle=False)
    raw, pdf, fdf, lib, p = H.workload(name)
    if f"{name}__input_checksum" not in g or str(g[f"{name}__input_checksum"]) != H.input_checksum(raw, pdf, fdf):
        pytest.skip("golden not applicable")
    cand = H.edge_candidate_frame(name)
    for c in cand.columns:
        assert np.array_equal(cand[c].values, g[f"{name}__cand_{c}"]), c
    cin, keep = H.candidates_in_from_arrays(lib, {c: cand[c].values.astype(np.int64) for c in INT_COLS})
    score = oracle_lib.score_candidates_4d if name == "parity_4d" else oracle_lib.score_candidates
    sc = score(raw, lib, H.scoring_config().to_struct(), cin)
    v = sc["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g[f"{name}__feat_precursor_idx"])
    assert np.array_equal(keep["rank"][v], g[f"{name}__feat_rank"])
    assert 20 < v.sum() < len(cand)
    F, G = sc["features"][v], g[f"{name}__feat_matrix"]
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    fm = sc["fragment_mz_library"] > 0
    assert fm.sum() == len(g[f"{name}__frag_mz_library"])
    for k, v2 in FRAG_MAP.items():
        a, b = sc[v2][fm], g[f"{name}__frag_{k}"]
        assert np.array_equal(a, b) if k != "correlation" else H.rel_err(a, b).max() < 1e-4, k


# ---- fitted quadrupole model, other tolerances (tests/golden/scoring_variants.npz) ------------------------------
@pytest.mark.parametrize("tag", list(H.SCORING_VARIANTS_EXTRA))
@pytest.mark.parametrize("name", list(H.SCORING_VARIANT_FILES))
def test_scoring_extra_variants_vs_reference(name, tag, oracle_lib):
    import os

    path = os.path.join(H.GOLDEN_DIR, "scoring_variants.npz")
    if not os.path.exists(path):
        pytest.skip("golden scoring_variants.npz missing")
    gv = np.load(path, allow_pickle=False)
    g, raw, lib, p = _golden(name)
    if f"{name}__input_checksum" not in gv or str(gv[f"{name}__input_checksum"]) != str(g["input_checksum"]):
        pytest.skip("golden not applicable")
    var = H.SCORING_VARIANTS_EXTRA[tag]
    cfg = H.scoring_config(**var["config"]).to_struct(quad_sigma=var["quad_sigma"], quad_delta_mu=var["quad_delta_mu"])
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    score = oracle_lib.score_candidates_4d if name in ("parity_4d", "parity_4d_overlap") else oracle_lib.score_candidates
    sc = score(raw, lib, cfg, cin)
    v = sc["valid"].astype(bool)
    key = f"{name}__{tag}__"
    assert np.array_equal(keep["precursor_idx"][v], gv[key + "feat_precursor_idx"])
    assert np.array_equal(keep["rank"][v], gv[key + "feat_rank"])
    F, G = sc["features"][v], gv[key + "feat_matrix"]
    if tag == "quad":  # the fitted model must matter, or the case pins nothing
        assert not np.array_equal(G, g["feat_matrix"]) if G.shape == g["feat_matrix"].shape else True
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    fm = sc["fragment_mz_library"] > 0
    assert fm.sum() == len(gv[key + "frag_mz_library"])
    for k, v2 in FRAG_MAP.items():
        a, b = sc[v2][fm], gv[key + f"frag_{k}"]
        assert np.array_equal(a, b) if k != "correlation" else H.rel_err(a, b).max() < 1e-4, k


# ---- library with fewer isotope columns than top_k_precursors / top_k_isotopes (tests/golden/iso2.npz) -------------
def test_two_isotope_library_vs_reference(oracle_lib):
    import os

    from alphadia_b200.library import assemble_library_arrays

    path = os.path.join(H.GOLDEN_DIR, "iso2.npz")
    if not os.path.exists(path):
        pytest.skip("golden iso2.npz missing")
    g = np.load(path, allow_pickle=False)
    raw, pdf0, fdf, _, p = H.workload("parity_small")
    if str(g["input_checksum"]) != H.input_checksum(raw, pdf0, fdf):
        pytest.skip("golden not applicable")
    lib = assemble_library_arrays(pdf0.drop(columns=["i_2", "i_3"]), fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    assert lib["isotopes"].shape[1] == 2
    arrs = oracle_lib.select_candidates(raw, lib, H.selection_config(p["rt_tolerance"]).to_struct(), H.default_kernel(raw))
    m = arrs["score"] > 0
    assert m.sum() == len(g["cand_precursor_idx"]) > 100
    for c in INT_COLS:
        assert np.array_equal(arrs[c][m].astype(np.int64), g["cand_" + c].astype(np.int64)), c
    assert np.array_equal(arrs["score"][m], g["cand_score"])
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    sc = oracle_lib.score_candidates(raw, lib, H.scoring_config().to_struct(), cin)
    v = sc["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g["feat_precursor_idx"]) and np.array_equal(keep["rank"][v], g["feat_rank"])
    F, G = sc["features"][v], g["feat_matrix"]
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    fm = sc["fragment_mz_library"] > 0
    assert fm.sum() == len(g["frag_mz_library"])
    for k, v2 in FRAG_MAP.items():
        a, b = sc[v2][fm], g[f"frag_{k}"]
        assert np.array_equal(a, b) if k != "correlation" else H.rel_err(a, b).max() < 1e-4, k


# ---- CandidateSelection.__call__ host side: the returned table (columns, order, dtypes) vs the reference's -------------
@pytest.mark.parametrize("name", ["parity_small", "parity_4d"])
def test_candidate_selection_table_vs_reference(name, oracle_lib, monkeypatch):
    """selection.py:622-676 + config_df.py:258-298: same columns in the same order with the same dtypes and values as the
    DataFrame the reference returned.  No GPU here: the two device calls are replaced by the oracle for this test only
    (tests/test_gpu_parity.py::test_operator_classes_end_to_end / _4d run the real ones)."""
    from alphadia_b200 import _abi, _lib
    from alphadia_b200.selection import CandidateSelection

    g, raw, lib, p = _golden(name)
    _, pdf, fdf, _, _ = H.workload(name)
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
        select = oracle_lib.select_candidates_4d if name == "parity_4d" else oracle_lib.select_candidates
        arrs = select(dev_raw.arrays, dev_lib.arrays, cfg, kernel)
        keep = arrs["score"] > 0  # adb_fetch_candidate_table: rows with score > 0 in container order
        state["table"] = {c: arrs[c][keep] for c in INT_COLS + ["score"]}
        # library row of every container row (the container holds candidate_count rows per precursor, in library order)
        state["table"]["lib_row"] = (np.flatnonzero(keep) // int(cfg.candidate_count)).astype(np.int64)
        return int(keep.sum())

    def fetch_table(dev_raw, n, arrs=None):
        table = _abi.alloc_candidate_table(n)
        for c, v in state["table"].items():
            table[c][:n] = v
        return table

    monkeypatch.setattr(_lib, "device_rawfile_for", lambda dia_data, adapted: HostRaw(adapted))
    monkeypatch.setattr(_lib, "DeviceLibrary", HostLibrary)
    monkeypatch.setattr(_lib, "select_candidates_resident", select_resident)
    monkeypatch.setattr(_lib, "fetch_candidate_table", fetch_table)
    kw = {"mobility_tolerance": p["mobility_tolerance"]} if "mobility_tolerance" in p else {}
    sel = CandidateSelection(raw, pdf.copy(), fdf.copy(), H.selection_config(p["rt_tolerance"], **kw), rt_column="rt_library",
                             mobility_column="mobility_library", precursor_mz_column="mz_library", fragment_mz_column="mz_library",
                             fwhm_rt=5.0, fwhm_mobility=0.01)
    assert np.array_equal(sel.kernel, g["sel_kernel"]) or np.allclose(sel.kernel, g["sel_kernel"], rtol=1e-6, atol=0)
    df = sel(thread_count=4)
    expected_columns = ["precursor_idx", "rank", "score", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start",
                        "frame_stop", "elution_group_idx", "decoy"]
    assert list(df.columns) == expected_columns
    for c in expected_columns:
        assert df[c].dtype == g["cand_" + c].dtype, (c, df[c].dtype, g["cand_" + c].dtype)
        assert np.array_equal(df[c].values, g["cand_" + c]), c
    assert np.array_equal(df.index.values, np.arange(len(df)))


@pytest.mark.parametrize("name,tag", [("parity_small", ""), ("parity_small", "_k6"), ("parity_4d", ""), ("parity_4d", "_legacy")])
def test_candidate_scoring_tables_vs_reference(name, tag, oracle_lib, monkeypatch):
    """scoring.py:582-661: both returned DataFrames — column sets, the fixed part of the column order, dtypes, row order and
    values — against the tables the reference returned for the same candidates.  The device call is replaced by the oracle
    for this CPU test only (the GPU operator tests run the real one)."""
    import pandas as pd

    from alphadia_b200 import _lib
    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS, FRAGMENT_COLUMNS, CandidateScoring

    g, raw, lib, p = _golden(name)
    _, pdf, fdf, _, _ = H.workload(name)
    score = oracle_lib.score_candidates_4d if name == "parity_4d" else oracle_lib.score_candidates

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

    monkeypatch.setattr(_lib, "device_rawfile_for", lambda dia_data, adapted: HostRaw(adapted))
    monkeypatch.setattr(_lib, "DeviceLibrary", HostLibrary)
    monkeypatch.setattr(_lib, "score_candidates_ragged", H.ragged_scoring_stub(score))
    cand_df = pd.DataFrame({c: g["cand_" + c] for c in INT_COLS + ["score", "elution_group_idx", "decoy"]})
    scorer = CandidateScoring(dia_data=raw, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(), config=H.scoring_config(**VARIANTS[tag]),
                              rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                              fragment_mz_column="mz_library")
    feat, frag = scorer(cand_df.copy())
    # feature table: the reference builds the merged column lists from sets, so only the fixed parts of the order are pinned
    ref_cols = [str(c) for c in g[f"feat{tag}_columns"]]
    assert set(feat.columns) == set(ref_cols)
    assert list(feat.columns[:48]) == ref_cols[:48] == DEFAULT_FEATURE_COLUMNS + ["precursor_idx", "rank"]
    assert list(feat.columns[-4:]) == ref_cols[-4:] == ["delta_rt", "n_K", "n_R", "n_P"]
    for c in ("precursor_idx", "rank", "n_K", "n_R", "n_P", "score", "frame_start", "frame_stop", "decoy", "charge"):
        assert feat[c].dtype == g[f"feat{tag}_{c}"].dtype, (c, feat[c].dtype)
        assert np.array_equal(feat[c].values, g[f"feat{tag}_{c}"]), c
    assert feat["delta_rt"].dtype == np.float32 and np.array_equal(feat["delta_rt"].values, g[f"feat{tag}_delta_rt"])
    F, G = feat[DEFAULT_FEATURE_COLUMNS].values, g[f"feat{tag}_matrix"]
    assert F.dtype == np.float32
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    # fragment table: exact column order, dtypes, rows
    assert list(frag.columns) == FRAGMENT_COLUMNS + ["elution_group_idx", "decoy"]
    for c in frag.columns:
        assert frag[c].dtype == g[f"frag{tag}_{c}"].dtype, (c, frag[c].dtype)
        if c == "correlation":
            assert H.rel_err(frag[c].values, g[f"frag{tag}_{c}"]).max() < 1e-4
        else:
            assert np.array_equal(frag[c].values, g[f"frag{tag}_{c}"]), c


def test_candidate_scoring_top_k_9999_tables_vs_reference(oracle_lib, monkeypatch):
    """CandidateScoring with top_k_fragments = 9999, the configuration quantify_candidates uses for the transfer library
    (extraction_handler.py:488-508, transfer_library_requantification_handler.py:102-124): the fragment table holds every
    library fragment of every valid candidate, equal to the live reference's table (tests/golden/k9999.npz).  Device call
    replaced by the oracle in this CPU test; tests/test_gpu_parity.py runs the real one."""
    import pandas as pd

    from alphadia_b200 import _lib
    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS, FRAGMENT_COLUMNS, CandidateScoring

    g, raw, lib, p = _golden("parity_f20")
    g9 = H.load_golden("k9999")
    assert g9 is not None and str(g9["input_checksum"]) == str(g["input_checksum"])
    _, pdf, fdf, _, _ = H.workload("parity_f20")

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

    monkeypatch.setattr(_lib, "device_rawfile_for", lambda dia_data, adapted: HostRaw(adapted))
    monkeypatch.setattr(_lib, "DeviceLibrary", HostLibrary)
    monkeypatch.setattr(_lib, "score_candidates_ragged", H.ragged_scoring_stub(oracle_lib.score_candidates))
    cand_df = pd.DataFrame({c: g["cand_" + c] for c in INT_COLS + ["score", "elution_group_idx", "decoy"]})
    scorer = CandidateScoring(dia_data=raw, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(),
                              config=H.scoring_config(top_k_fragments=9999), rt_column="rt_library",
                              mobility_column="mobility_library", precursor_mz_column="mz_library", fragment_mz_column="mz_library")
    feat, frag = scorer(cand_df.copy())
    assert np.array_equal(feat["precursor_idx"].values, g9["feat_precursor_idx"]) and np.array_equal(feat["rank"].values, g9["feat_rank"])
    F, G = feat[DEFAULT_FEATURE_COLUMNS].values, g9["feat_matrix"]
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j]).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    assert list(frag.columns) == FRAGMENT_COLUMNS + ["elution_group_idx", "decoy"]
    assert len(frag) == len(g9["frag_mz_library"]) and frag.groupby(["precursor_idx", "rank"]).size().max() == 20
    for c in frag.columns:
        assert frag[c].dtype == g9[f"frag_{c}"].dtype, (c, frag[c].dtype)
        if c == "correlation":
            assert H.rel_err(frag[c].values, g9[f"frag_{c}"]).max() < 1e-4
        else:
            assert np.array_equal(frag[c].values, g9[f"frag_{c}"]), c


# ---- FDR bookkeeping: get_q_values / keep_best of the live reference ------------------------
def _fdr_golden():
    import hashlib
    import os

    path = os.path.join(H.GOLDEN_DIR, "fdr_small.npz")
    if not os.path.exists(path):
        pytest.skip("golden fdr_small.npz missing")
    g = np.load(path, allow_pickle=False)
    df = H.fdr_inputs()
    if str(g["input_checksum"]) != hashlib.sha256(df.to_numpy().tobytes() + df.index.to_numpy().tobytes()).hexdigest():
        pytest.skip("fdr_inputs differs from the one the golden file was made with (numpy version?)")
    return g, df


def host_fdr_with_oracle(monkeypatch, oracle_lib):
    """alphadia_b200.fdr with its two device calls replaced by the oracle (CPU tests of the host packing only)."""
    from alphadia_b200 import _lib, fdr

    monkeypatch.setattr(_lib, "q_values", lambda score, decoy, extra, device=None: oracle_lib.q_values(score, decoy, extra))
    monkeypatch.setattr(_lib, "keep_best", lambda score, group, device=None: oracle_lib.keep_best(score, group))
    monkeypatch.setattr(_lib, "fragment_competition", lambda ws, we, rt, fs, fe, mz, rt_tol, ppm, device=None:
                        oracle_lib.fragment_competition(ws, we, rt, fs, fe, mz, rt_tol, ppm).astype(bool))
    return fdr


def test_fdr_oracle_and_host_packing_vs_reference(oracle_lib, monkeypatch):
    g, df = _fdr_golden()
    fdr = host_fdr_with_oracle(monkeypatch, oracle_lib)
    q = fdr.get_q_values(df.copy(), "proba", "_decoy")
    assert np.array_equal(q["row"].values, g["q_row"]) and np.array_equal(q.index.values, g["q_index"])
    assert np.array_equal(q["qval"].values, g["q_qval"])  # float64, bit-exact
    # the three best rows are decoys: their FDR is x / 0 = inf, which the running minimum from the back removes
    assert (q["_decoy"].values[:3] == 1).all() and np.isfinite(g["q_qval"]).all()
    q2 = fdr.get_q_values(df.copy(), "proba", "_decoy", extra_sort_columns=["precursor_idx", "rank"])
    assert np.array_equal(q2["row"].values, g["q2_row"]) and np.array_equal(q2["qval"].values, g["q2_qval"])
    for tag, cols in {"precursor": ["precursor_idx"], "channel_eg": ["elution_group_idx", "channel"], "eg": ["elution_group_idx"],
                      "default": None}.items():
        kept = fdr.keep_best(df.copy(), group_columns=cols)
        assert np.array_equal(kept["row"].values, g[f"keep_{tag}_row"]), tag
        assert np.array_equal(kept.index.values, np.arange(len(kept)))
    final = fdr.get_q_values(fdr.keep_best(q, group_columns=["elution_group_idx", "channel"]), "proba", "_decoy")
    assert np.array_equal(final["row"].values, g["final_row"]) and np.array_equal(final["qval"].values, g["final_qval"])


def check_perform_fdr_against_golden(fdr):
    """alphadia_b200.fdr.perform_fdr == the reference's perform_fdr (fdr.py:25-192) with the same stand-in classifier:
    same rows in the same order, same index, same columns, probabilities and q-values bit-identical."""
    import os

    path = os.path.join(H.GOLDEN_DIR, "perform_fdr_small.npz")
    g_scores = H.load_golden("parity_small")
    if not os.path.exists(path) or g_scores is None:
        pytest.skip("golden perform_fdr_small.npz missing")
    g = np.load(path, allow_pickle=False)
    if str(g["source_checksum"]) != str(g_scores["input_checksum"]):
        pytest.skip("perform_fdr golden was made from another scoring golden")
    raw = H.workload("parity_small")[0]
    for tag, case in H.PERFORM_FDR_CASES.items():
        df_target, df_decoy, frag = H.perform_fdr_inputs(g_scores)
        res = fdr.perform_fdr(H.PseudoClassifier(), H.FDR_FEATURE_COLUMNS, df_target, df_decoy, competitive=case["competitive"],
                              group_channels=case["group_channels"], df_fragments=frag if case["fragments"] else None,
                              dia_cycle=raw.cycle, random_state=7)
        assert list(res.columns) == [str(c) for c in g[f"{tag}__columns"]], tag
        assert np.array_equal(res.index.values, g[f"{tag}__index"]), tag
        for c in ("precursor_idx", "rank", "proba", "qval", "_decoy"):
            assert np.array_equal(res[c].values, g[f"{tag}__{c}"]), (tag, c)
        assert 0 < len(res) < len(df_target) + len(df_decoy)


def test_perform_fdr_vs_reference(oracle_lib, monkeypatch):
    check_perform_fdr_against_golden(host_fdr_with_oracle(monkeypatch, oracle_lib))


def test_fdr_key_packing(oracle_lib, monkeypatch):
    """Host packing of tie-break and group columns: lexicographic order of several integer columns, non-integer group
    columns, and the inputs that must be refused."""
    import pandas as pd

    fdr = host_fdr_with_oracle(monkeypatch, oracle_lib)
    rng = np.random.default_rng(5)
    n = 4000
    df = pd.DataFrame({"proba": np.round(rng.random(n), 1), "_decoy": rng.integers(0, 2, n), "a": rng.integers(0, 2 ** 20, n),
                       "b": rng.integers(0, 2 ** 31, n).astype(np.int64), "flag": rng.random(n) < 0.5,
                       "name": rng.choice(np.array(["x", "y", "zz", "w"], dtype=object), n), "f": np.round(rng.random(n), 1)})
    df.loc[::7, "a"] = 5  # ties in the first extra column so that the second decides
    got = fdr.get_q_values(df, extra_sort_columns=["a", "b"])
    ref = df.sort_values(["proba", "_decoy", "a", "b"])
    assert np.array_equal(got.index.values, ref.index.values)
    got = fdr.get_q_values(df, extra_sort_columns=[])
    assert np.array_equal(got.index.values, df.sort_values(["proba", "_decoy"], kind="stable").index.values)
    for cols in (["name"], ["name", "a"], ["f"], ["flag", "a"], ["a", "b"]):
        kept = fdr.keep_best(df, group_columns=cols)
        ref = df.reset_index(drop=True).sort_values(["proba", *cols]).groupby(cols).head(1).sort_index().reset_index(drop=True)
        pd.testing.assert_frame_equal(kept, ref)
    # float32 probabilities (what a torch classifier hands back): same order and q-values as pandas on the float32 column
    df32 = df.assign(proba=rng.random(n).astype(np.float32).round(2), precursor_idx=df["a"])
    got = fdr.get_q_values(df32)
    ref = df32.sort_values(["proba", "_decoy", "precursor_idx"])
    assert got["proba"].dtype == np.float32 and np.array_equal(got.index.values, ref.index.values)
    d = ref["_decoy"].to_numpy()
    with np.errstate(all="ignore"):
        expect = np.flip(np.minimum.accumulate(np.flip(np.cumsum(d) / np.cumsum(1 - d))))
    assert np.array_equal(got["qval"].values, expect)
    # tie-break columns of any dtype (the reference sorts by the protein-group string, outputtransform/protein_fdr.py:72),
    # negative values and keys wider than 63 bits: same rows, order and q-values as the pandas formulation
    wide = df.assign(a=df["a"].values.astype(np.int64) * 2 ** 30, b=df["b"].values * 2 ** 10, c=df["a"].values.astype(np.int64) * 2 ** 31)
    huge = df.assign(a=df["a"].values.astype(np.int64) * 2 ** 40 + rng.integers(0, 2 ** 40, n), b=rng.integers(0, 2 ** 62, n), c=rng.integers(0, 2 ** 62, n))
    for frame, cols in ((df.assign(a=-df["a"]), ["a"]), (df, ["name"]), (df, ["name", "b"]), (df, ["f"]), (wide, ["a", "b", "c"]),
                        (huge, ["a", "b", "c"])):
        got = fdr.get_q_values(frame, extra_sort_columns=cols)
        ref = frame.sort_values(["proba", "_decoy", *cols], kind="stable")
        assert np.array_equal(got.index.values, ref.index.values), cols
        d = ref["_decoy"].to_numpy()
        with np.errstate(all="ignore"):
            expect = np.flip(np.minimum.accumulate(np.flip(np.cumsum(d) / np.cumsum(1 - d))))
        assert np.array_equal(got["qval"].values, expect, equal_nan=True), cols


@pytest.mark.parametrize("n,levels", [(1, 1), (2, 1), (1000, 7), (50_000, 300), (50_000, 10 ** 9)])
def test_fdr_oracle_equals_pandas_formulation(n, levels, oracle_lib):
    """Random tables (heavy ties, negative and huge scores, all-target / all-decoy runs): the oracle against the reference's
    formulation written out with pandas (sort_values, cumsum, minimum.accumulate; sort_values, groupby.head, sort_index)."""
    import pandas as pd

    rng = np.random.default_rng(n + levels % 1000)
    score = rng.integers(-levels, levels + 1, n) / max(levels, 1) * rng.choice([1.0, 1e-300, 1e300])
    decoy = (rng.random(n) < rng.choice([0.0, 0.5, 1.0])).astype(np.uint8)
    extra = rng.integers(0, max(n // 4, 1), n).astype(np.uint64)
    order, q = oracle_lib.q_values(score, decoy, extra)
    df = pd.DataFrame({"s": score, "d": decoy, "e": extra}).sort_values(["s", "d", "e"])
    assert np.array_equal(df.index.values, order)
    with np.errstate(all="ignore"):
        fdr_values = np.cumsum(df["d"].values) / np.cumsum(1 - df["d"].values.astype(np.int64))
    assert np.array_equal(np.flip(np.minimum.accumulate(np.flip(fdr_values))), q, equal_nan=True)
    group = rng.integers(0, max(n // 3, 1), n).astype(np.uint64) << np.uint64(rng.integers(0, 40))
    keep = oracle_lib.keep_best(score, group)
    best = pd.DataFrame({"s": score, "g": group}).sort_values(["s", "g"]).groupby("g").head(1).sort_index()
    assert np.array_equal(np.flatnonzero(keep), best.index.values)


# ---- FDR classifier: BinaryClassifierLegacyNewBatching of the live reference ----------------------------
def _classifier_golden():
    import hashlib

    g = H.load_golden("classifier_small")
    if g is None:
        pytest.skip("golden classifier_small.npz missing")
    x, y = H.classifier_inputs()
    if str(g["input_checksum"]) != hashlib.sha256(x.tobytes() + y.tobytes()).hexdigest():
        pytest.skip("classifier_inputs differs from the one the golden file was made with (numpy version?)")
    state = {k[3:]: g[k] for k in g.files if k.startswith("w__")}
    return g, x, y, state


def test_classifier_oracle_vs_reference():
    """The numpy restatement of FeedForwardNN.forward (eval mode) against predict_proba / predict of the reference classifier
    trained in this container (torch CPU): probabilities within 1e-5 absolute, identical classes."""
    import oracle

    g, x, y, state = _classifier_golden()
    p = oracle.classifier_predict_proba(state, x)
    assert p.shape == g["proba"].shape and p.dtype == np.float32
    assert np.abs(p - g["proba"]).max() < 1e-5
    assert H.rel_err(p[:, 1], g["proba"][:, 1], floor=1e-3).max() < 1e-4
    assert np.array_equal(np.argmax(p, axis=1), g["predict"])


def test_classifier_state_dict_round_trip():
    """from_state_dict accepts the reference's dictionary (classifiers.py:258-309) and to_state_dict hands it back."""
    from alphadia_b200.classifier import BinaryClassifierLegacyNewBatching

    g, x, y, state = _classifier_golden()
    clf = BinaryClassifierLegacyNewBatching()
    assert not clf.fitted
    with pytest.raises(ValueError):
        clf.predict_proba(x)
    clf.from_state_dict({"_fitted": True, "input_dim": int(g["input_dim"]), "output_dim": 2, "layers": [int(v) for v in g["layers"]],
                         "dropout": 0.001, "epochs": 3, "network_state_dict": state}, load_hyperparameters=True)
    assert clf.fitted and clf.input_dim == x.shape[1] and clf.layers == [100, 50, 20, 5] and clf.epochs == 3
    sd = clf.to_state_dict()
    assert set(sd["network_state_dict"]) == set(state)
    for k, v in state.items():
        assert np.array_equal(sd["network_state_dict"][k], v), k


def test_fdr_missing_values_vs_reference(oracle_lib, monkeypatch):
    """NaN scores and NaN group keys (tests/golden/fdr_nan.npz from the live reference): same rows, order and q-values."""
    g = H.load_golden("fdr_nan")
    if g is None:
        pytest.skip("golden fdr_nan.npz missing")
    import hashlib

    df = H.fdr_inputs_nan()
    assert str(g["input_checksum"]) == hashlib.sha256(df.to_numpy().tobytes() + df.index.to_numpy().tobytes()).hexdigest()
    fdr = host_fdr_with_oracle(monkeypatch, oracle_lib)
    q = fdr.get_q_values(df.copy(), "proba", "_decoy")
    assert np.array_equal(q["row"].values, g["q_row"]) and np.array_equal(q.index.values, g["q_index"])
    assert np.array_equal(q["qval"].values, g["q_qval"], equal_nan=True)
    q2 = fdr.get_q_values(df.copy(), "proba", "_decoy", extra_sort_columns=["rank", "gnan"])
    assert np.array_equal(q2["row"].values, g["q2_row"]) and np.array_equal(q2["qval"].values, g["q2_qval"], equal_nan=True)
    for tag, cols in {"precursor": ["precursor_idx"], "gnan": ["gnan"], "gnan_channel": ["gnan", "channel"]}.items():
        kept = fdr.keep_best(df.copy(), group_columns=cols)
        assert np.array_equal(kept["row"].values, g[f"keep_{tag}_row"]), tag
        assert np.array_equal(kept.index.values, np.arange(len(kept)))
    # the NaN-only groups keep exactly one row each, the group with NaN first and +inf later keeps an +inf row
    kept = fdr.keep_best(df.copy(), group_columns=["precursor_idx"])
    pidx = np.unique(df["precursor_idx"].values)
    assert kept[kept["precursor_idx"].isin(pidx[:5])]["proba"].isna().all() and len(kept[kept["precursor_idx"].isin(pidx[:5])]) == 5
    assert np.isposinf(kept[kept["precursor_idx"] == pidx[7]]["proba"].values).all()


@pytest.mark.parametrize("k", [20, 16])
def test_scoring_tied_fragments_vs_reference(oracle_lib, k):
    """More than 15 fragments per precursor with tied m/z and tied library intensities (tests/golden/ties_f20.npz from the live
    reference): numba's argsort is an unstable quicksort there, and the order it gives the tied fragments decides their
    type / position / number columns, the b- and y-ion features and - with top_k_fragments = 16 - which fragments are kept."""
    from alphadia_b200.library import assemble_library_arrays

    g = H.load_golden("ties_f20")
    if g is None:
        pytest.skip("golden ties_f20.npz missing")
    raw, pdf, fdf, p = H.tied_fragment_library()
    assert str(g["input_checksum"]) == H.input_checksum(raw, pdf, fdf)
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    arrs = oracle_lib.score_candidates(raw, lib, H.scoring_config(top_k_fragments=k).to_struct(), cin)
    v = arrs["valid"].astype(bool)
    assert np.array_equal(keep["precursor_idx"][v], g[f"feat_k{k}_precursor_idx"]) and np.array_equal(keep["rank"][v], g[f"feat_k{k}_rank"])
    F, G = arrs["features"][v], g[f"feat_k{k}_matrix"]
    for j in range(46):
        if j in BLAS_FEATURES:
            assert H.rel_err(F[:, j], G[:, j], floor=1e-3).max() < 1e-4, j
        else:
            assert ((F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))).all(), f"feature {j} not bit-exact"
    m = arrs["fragment_mz_library"] > 0
    assert m.sum() == len(g[f"frag_k{k}_mz_library"])
    for name, col in FRAG_MAP.items():
        a, b = arrs[col][m], g[f"frag_k{k}_{name}"]
        if name == "correlation":
            assert H.rel_err(a, b, floor=1e-3).max() < 1e-4
        else:
            assert np.array_equal(a, b), name

"""The CUDA scoring passes (alphadia_b200/csrc/adb_score_dp.cuh) executed thread by thread on the CPU (tests/hostsim) against
the oracle, on machines without a GPU.  Same bars as tests/test_gpu_parity.py: valid mask and integer columns exact, features
within 1e-4 relative."""

import numpy as np
import pytest

from tests import helpers as H
from tests import hostsim

INT_COLS = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]
FRAG_F32 = ["fragment_mz_library", "fragment_mz", "fragment_mz_observed", "fragment_height", "fragment_intensity",
            "fragment_mass_error", "fragment_correlation"]
FRAG_U8 = ["fragment_position", "fragment_number", "fragment_type", "fragment_charge", "fragment_loss_type"]
RTOL = 1e-4

SCORING_VARIANTS = {
    "default": {},
    "legacy": dict(quant_all=False, experimental_xic=False),
    "k6": dict(top_k_fragments=6, top_k_isotopes=4, quant_window=2),
    "qall_legacy_xic": dict(quant_all=True, experimental_xic=False),
    "noqall_xic": dict(quant_all=False, experimental_xic=True),
}


def assert_scores_close(a, b, what=""):
    assert a["status"] == 0, f"{what} status {a['status']}"
    assert np.array_equal(a["valid"], b["valid"]), f"{what} valid mask"
    v = a["valid"].astype(bool)
    assert not a["features"][~v].any(), f"{what}: rows of invalid candidates must stay zero"
    Fa, Fb = a["features"][v], b["features"][v]
    nan_a, nan_b = np.isnan(Fa), np.isnan(Fb)
    assert np.array_equal(nan_a, nan_b), f"{what} NaN pattern"
    s = np.nanmedian(np.abs(Fb), axis=0)
    floor = np.maximum(1e-6, 1e-6 * np.where(np.isfinite(s), s, 0.0))
    err = np.abs(Fa - Fb) / np.maximum(np.maximum(np.abs(Fa), np.abs(Fb)), floor[None, :])
    err = np.where(nan_a, 0.0, err)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < RTOL, f"{what} feature {worst[1]} row {worst[0]}: {Fa[worst]} vs {Fb[worst]} (rel {err.max():.3e})"
    for k in FRAG_U8:
        assert np.array_equal(a[k], b[k]), f"{what} {k}"
    for k in FRAG_F32:
        e = H.rel_err(a[k], b[k])
        assert e.max() < RTOL, f"{what} {k}: rel {e.max():.3e}"
    return float(err.max()), int((Fa != Fb).sum())


def _golden_candidates(name, lib):
    g = H.load_golden(name)
    return H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})


@pytest.mark.parametrize("variant", list(SCORING_VARIANTS))
@pytest.mark.parametrize("name", ["config1", "parity_small", "parity_f20"])
def test_hostsim_matches_oracle(oracle_lib, name, variant):
    raw, pdf, fdf, lib, p = H.workload(name)
    cin, keep = _golden_candidates(name, lib)
    cfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
    ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
    assert ref["valid"].sum() > 50
    # a shuffled processing order and a batch size that splits the table must not change anything
    order = np.random.default_rng(3).permutation(int(cin.n)).astype(np.int32)
    # (nor the bucket width of the time-blocked m/z index: 3 buckets per segment = long searches, 4096 = mostly empty buckets)
    got = hostsim.score_candidates(raw, lib, cfg, cin, batch=257, order=order, n_buckets={'default': 0, 'legacy': 3, 'k6': 4096}.get(variant, 0))
    assert_scores_close(got, ref, what=f"{name}/{variant}")


@pytest.mark.parametrize("tag", list(H.SCORING_VARIANTS_EXTRA))
def test_hostsim_scoring_variants_extra(oracle_lib, tag):
    raw, pdf, fdf, lib, p = H.workload("parity_small")
    cin, keep = _golden_candidates("parity_small", lib)
    var = H.SCORING_VARIANTS_EXTRA[tag]
    cfg = H.scoring_config(**var["config"]).to_struct(quad_sigma=var["quad_sigma"], quad_delta_mu=var["quad_delta_mu"])
    ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
    got = hostsim.score_candidates(raw, lib, cfg, cin, batch=100000)
    assert_scores_close(got, ref, what=f"parity_small/{tag}")


def test_hostsim_ragged_library_and_edge_windows(oracle_lib):
    from alphadia_b200.library import assemble_library_arrays

    raw, pdf0, fdf0, _, p = H.workload("parity_small")
    pdf, fdf = H.ragged_library_frames(pdf0, fdf0, float(np.max(raw.rt_values)))
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    sel = oracle_lib.select_candidates(raw, lib, H.selection_config(p["rt_tolerance"]).to_struct(), H.default_kernel(raw))
    m = sel["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: sel[c][m] for c in INT_COLS})
    for variant in ("default", "legacy"):
        cfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
        assert_scores_close(hostsim.score_candidates(raw, lib, cfg, cin), oracle_lib.score_candidates(raw, lib, cfg, cin), what=f"ragged/{variant}")
    # hand-made windows: 1 to 241 cycles wide, clipped at the ends of the run
    raw, pdf, fdf, lib, p = H.workload("parity_small")
    df = H.edge_candidate_frame("parity_small")
    cin, keep = H.candidates_in_from_arrays(lib, {c: df[c].values for c in INT_COLS})
    for variant in ("default", "legacy"):
        cfg = H.scoring_config(**SCORING_VARIANTS[variant]).to_struct()
        ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
        assert 0 < ref["valid"].sum() < int(cin.n)
        assert_scores_close(hostsim.score_candidates(raw, lib, cfg, cin), ref, what=f"edge/{variant}")


@pytest.mark.parametrize("name,kw", [("parity_f48", dict(top_k_fragments=9999)),
                                     ("parity_f48", dict(top_k_fragments=40, quant_all=False, experimental_xic=False)),
                                     ("parity_f96", dict(top_k_fragments=9999))])
def test_hostsim_wide_fragment_libraries(oracle_lib, name, kw):
    """48 and 96 library fragments per precursor (more than the 32 slots of the dense device tables; the ragged entry keeps up to
    128): the scoring passes with top_k_fragments = 9999 / 40 against the oracle at the same width."""
    raw, pdf, fdf, lib, p = H.workload(name)
    sel = oracle_lib.select_candidates(raw, lib, H.selection_config(p["rt_tolerance"]).to_struct(), H.default_kernel(raw))
    m = sel["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: sel[c][m] for c in INT_COLS})
    wide = int(np.max(lib["frag_stop_idx"] - lib["frag_start_idx"]))
    assert wide == int(name[len("parity_f"):])
    k_eff = min(int(kw["top_k_fragments"]), wide)
    ref = oracle_lib.score_candidates(raw, lib, H.scoring_config(**{**kw, "top_k_fragments": k_eff}).to_struct(), cin)
    assert ref["valid"].sum() > 100 and (ref["fragment_mz_library"] > 0).sum(axis=1).max() == k_eff
    cfg = H.scoring_config(**{**kw, "top_k_fragments": k_eff}).to_struct()
    got = hostsim.score_candidates(raw, lib, cfg, cin, batch=97, ks=k_eff)
    assert_scores_close(got, ref, what=f"{name}/{kw}")


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_hostsim_randomized_sweep(oracle_lib, seed):
    """Random raw files, libraries (40 % ragged), selection and scoring configurations, quadrupole parameters, processing orders,
    batch sizes and index bucket counts: the CUDA passes run thread by thread on the CPU stay bit-identical to the oracle
    (385 further cases of this generator were run once by hand, 0 differences)."""
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.synthetic import make_config_3d

    rng = np.random.default_rng(seed)
    compared = 0
    for it in range(8):
        name = str(rng.choice(["parity_small", "parity_f20", "config1"]))
        s = int(rng.integers(1, 10**6))
        raw, pdf, fdf, p = make_config_3d(name, seed=s, n_precursors=int(rng.integers(80, 300)), scale_noise=float(rng.choice([0.3, 1.0, 3.0])))
        if rng.random() < 0.4:
            pdf, fdf = H.ragged_library_frames(pdf, fdf, float(np.max(raw.rt_values)), seed=s)
        lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
        selcfg = H.selection_config(p["rt_tolerance"] * float(rng.choice([0.5, 1.0, 2.0])), candidate_count=int(rng.integers(1, 6)))
        sel = oracle_lib.select_candidates(raw, lib, selcfg.to_struct(), H.default_kernel(raw, fwhm_rt=float(rng.choice([2.0, 5.0, 10.0]))))
        m = sel["score"] > 0
        if m.sum() == 0:
            continue
        cin, keep = H.candidates_in_from_arrays(lib, {c: sel[c][m] for c in INT_COLS})
        kw = dict(top_k_fragments=int(rng.choice([3, 6, 12, 20, 32])), top_k_isotopes=int(rng.integers(1, 5)),
                  quant_window=int(rng.integers(1, 5)), quant_all=bool(rng.integers(0, 2)), experimental_xic=bool(rng.integers(0, 2)),
                  precursor_mz_tolerance=float(rng.choice([2, 5, 15, 50])), fragment_mz_tolerance=float(rng.choice([3, 10, 30, 100])))
        cfg = H.scoring_config(**kw).to_struct(quad_sigma=(float(rng.choice([0.2, 0.5, 1.0])), float(rng.choice([0.2, 0.8]))),
                                              quad_delta_mu=(float(rng.choice([0.0, 0.3, -0.5])), float(rng.choice([0.0, -0.4]))))
        ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
        got = hostsim.score_candidates(raw, lib, cfg, cin, batch=int(rng.choice([33, 257, 100000])),
                                       order=rng.permutation(int(cin.n)).astype(np.int32), n_buckets=int(rng.choice([0, 3, 64, 4096])))
        if ref["valid"].sum() == 0:
            assert got["status"] == 0 and got["valid"].sum() == 0
            continue
        err, _ = assert_scores_close(got, ref, what=f"seed {seed} case {it} {name} {kw}")
        assert err == 0.0, f"seed {seed} case {it}: features differ (max rel {err:.3e})"
        compared += 1
    assert compared >= 4


@pytest.mark.parametrize("k", [20, 16])
def test_hostsim_tied_fragments(oracle_lib, k):
    """The tie-heavy 20-fragment library of tests/golden/ties_f20.npz: the passes follow numba's quicksort order among tied
    fragment m/z and intensities exactly as the oracle (which the golden pins against the live reference)."""
    from alphadia_b200.library import assemble_library_arrays

    raw, pdf, fdf, p = H.tied_fragment_library()
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    g = H.load_golden("ties_f20")
    cin, keep = H.candidates_in_from_arrays(lib, {c: g["cand_" + c] for c in INT_COLS})
    cfg = H.scoring_config(top_k_fragments=k).to_struct()
    ref = oracle_lib.score_candidates(raw, lib, cfg, cin)
    got = hostsim.score_candidates(raw, lib, cfg, cin, batch=101, ks=k)
    err, _ = assert_scores_close(got, ref, what=f"ties k={k}")
    assert err == 0.0

"""Known-answer vectors taken from the reference's OWN unit tests, replayed against the
C oracle and the host-side restatements.  CPU only."""

import ctypes as C

import numpy as np
import pandas as pd
import pytest

from alphadia_b200 import _abi
from alphadia_b200.fragcomp import FragmentCompetition, candidate_hash
from alphadia_b200.scoring import calculate_score_groups, merge_missing_columns


# ---- tests/unit_tests/fragcomp/test_fragcomp.py:12-35 --------------------------------------------
@pytest.mark.parametrize("a,b,expected", [
    (np.arange(100, 1100, 100), np.arange(100, 1100, 100), 10),
    (np.arange(100, 1100, 100), np.array([100]), 1),
    (np.array([]), np.array([]), 0),
    (np.arange(100, 1100, 100), np.array([]), 0),
    (np.array([]), np.array([100, 200, 300, 400, 500, 600, 700, 801, 901, 1001]), 0),
    (np.arange(100, 1100, 100), np.arange(101, 1101, 100), 0),
])
def test_fragment_overlap(oracle_lib, a, b, expected):
    L = oracle_lib.lib()
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    L.adbo_fragment_overlap_f64.restype = C.c_int
    got = L.adbo_fragment_overlap_f64(a.ctypes.data_as(C.c_void_p), C.c_int(len(a)), b.ctypes.data_as(C.c_void_p),
                                      C.c_int(len(b)), C.c_double(10.0))
    assert got == expected


# ---- test_fragcomp.py:38-58 ------------------------------------------------------------------------
def test_compete_for_fragments(oracle_lib):
    rt = np.array([10.0, 20.0, 20.0, 10.0, 10.0, 20])
    frag_start = np.array([0, 10, 20, 30, 40, 50])
    frag_stop = np.array([10, 20, 30, 40, 50, 60])
    fragment_mz = np.tile(np.arange(100, 110), 6).astype(np.float64)
    valid = oracle_lib.fragment_competition(np.array([0, 3]), np.array([3, 6]), rt, frag_start, frag_stop, fragment_mz, 3, 15)
    assert np.all(valid == np.array([True, True, False, True, False, True]))


# ---- test_fragcomp.py:61-100 (the pandas preparation; the kernel is replaced by the oracle) ---------
def test_fragment_competition_dataframes(oracle_lib):
    psm_df = pd.DataFrame({
        "precursor_idx": np.arange(6), "rank": np.zeros(6, dtype=np.uint8),
        "rt_observed": np.array([10.0, 20.0, 20.0, 10.0, 10.0, 20]), "proba": np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6]),
        "mz_observed": np.array([150.0, 150.0, 150.0, 250.0, 250.0, 250.0]),
    })
    frag_df = pd.DataFrame({
        "precursor_idx": np.repeat(np.arange(6), 10), "rank": np.zeros(60, dtype=np.uint8),
        "mz_observed": np.tile(np.arange(100, 110), 6).astype(np.float64),
    })
    cycle = np.zeros((1, 3, 1, 2))
    cycle[0, :, 0, 0] = [-1, 100, 200]
    cycle[0, :, 0, 1] = [-1, 200, 300]
    fc = FragmentCompetition()
    plan = fc.plan(psm_df, frag_df, cycle)
    assert plan.psm_df["_candidate_idx"].dtype == np.uint64
    valid = oracle_lib.fragment_competition(plan.window_start, plan.window_stop, plan.rt, plan.frag_start, plan.frag_stop,
                                            plan.fragment_mz, 3, 15)
    kept = plan.psm_df[valid]
    assert list(kept["precursor_idx"]) == [0, 1, 3, 5]


# ---- test_fragcomp.py:103-113 ----------------------------------------------------------------------
def test_candidate_hash():
    h = candidate_hash(np.array([1, 2, 1000000], dtype=np.uint32), np.array([0, 1, 2], dtype=np.uint8))
    assert h.dtype == np.uint64
    assert np.all(h == np.array([1, 4294967298, 8590934592], dtype=np.uint64))


# ---- tests/unit_tests/search/selection/test_fft.py:71-87: delta (*) ones == box --------------------
@pytest.mark.parametrize("shape", [(64, 64), (32, 48)])
def test_convolution_delta_box(oracle_lib, shape):
    L = oracle_lib.lib()
    x = np.zeros(shape, np.float32)
    x[shape[0] // 2, shape[1] // 2] = 1.0
    k = np.ones((20, 20), np.float32)
    out = np.zeros(shape, np.float32)
    L.adbo_conv_circular(_abi.ptr(x), C.c_int(shape[0]), C.c_int(shape[1]), _abi.ptr(k), C.c_int(20), C.c_int(20), _abi.ptr(out))
    expect = np.zeros(shape, np.float32)
    expect[shape[0] // 2 - 10: shape[0] // 2 + 10, shape[1] // 2 - 10: shape[1] // 2 + 10] = 1.0
    assert np.allclose(out, expect, atol=1e-6)


def test_convolution_is_circular(oracle_lib):
    L = oracle_lib.lib()
    x = np.zeros((2, 40), np.float32)
    x[:, 1] = 1.0  # next to the left edge: the kernel support wraps around
    k = np.random.default_rng(0).uniform(0.1, 1, (2, 30)).astype(np.float32)
    out = np.zeros_like(x)
    L.adbo_conv_circular(_abi.ptr(x), C.c_int(2), C.c_int(40), _abi.ptr(k), C.c_int(2), C.c_int(30), _abi.ptr(out))
    # FFT definition (alphadia/search/selection/fft.py:158-167) in float64
    F = np.fft.irfft2(np.fft.rfft2(x.astype(np.float64)) * np.fft.rfft2(k.astype(np.float64), x.shape), x.shape)
    ref = np.roll(F, (-1, -15), axis=(0, 1))
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    assert out[0, 39] > 0  # wrapped


# ---- tests/unit_tests/raw_data/test_raw_data.py:26-175 get_frame_indices ----------------------------
@pytest.mark.parametrize("rt,opt,mn,expected", [
    ((10.0, 20.0), 1, 1, (10, 20)),
    ((10.0, 20.0), 4, 1, (10, 30)),
    ((10.0, 20.0), 4, 8, (10, 50)),
    ((90.0, 95.0), 4, 1, (75, 95)),
    ((90.0, 95.0), 4, 8, (55, 95)),
    ((90.0, 95.0), 4, 1000, (5, 95)),
])
def test_get_frame_indices(oracle_lib, rt, opt, mn, expected):
    L = oracle_lib.lib()
    cycle = np.zeros((1, 5, 1, 2))
    cycle[0, :, 0, 0] = [100.0, 200.0, 300.0, 400.0, 500.0]
    cycle[0, :, 0, 1] = [200.0, 300.0, 400.0, 500.0, 600.0]

    class Raw:
        pass

    r = Raw()
    r.cycle = cycle
    r.rt_values = np.arange(0, 100, 1).astype(np.float32)
    r.mobility_values = np.array([0.0, 0.0], np.float32)
    r.zeroth_frame = 0
    r.precursor_cycle_max_index = 19
    r.peak_start_idx_list = np.arange(0, 1000, 10, dtype=np.int64)
    r.peak_stop_idx_list = r.peak_start_idx_list + 1
    r.mz_values = np.linspace(100, 1000, 1000).astype(np.float32)
    r.intensity_values = np.ones(1000, np.float32)
    r.scan_max_index = 0
    r.frame_max_index = 99
    desc, keep = _abi.make_rawfile3d_desc(r)
    out = np.zeros(2, np.int64)
    center, tol = (rt[0] + rt[1]) / 2, (rt[1] - rt[0]) / 2
    L.adbo_frame_indices(C.byref(desc), C.c_float(center), C.c_double(tol), C.c_int64(opt), C.c_int64(mn), _abi.ptr(out))
    assert tuple(out) == expected


# ---- tests/unit_tests/search/scoring/test_features.py:7-79 center_envelope_1d ------------------------
@pytest.mark.parametrize("x,expected", [
    ([1, 1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1]),
    ([100, 10, 1, 1, 1, 10, 100], [1, 1, 1, 1, 1, 1, 1]),
    ([100, 0, 0, 1, 0, 0, 100], [0, 0, 0, 1, 0, 0, 0]),
    ([1, 1, 1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1]),
    ([100, 10, 1, 1, 1, 1, 10, 100], [1, 1, 1, 1, 1, 1, 1, 1]),
    ([100, 0, 0, 1, 1, 0, 0, 100], [0, 0, 0, 1, 1, 0, 0, 0]),
])
def test_center_envelope(oracle_lib, x, expected):
    L = oracle_lib.lib()
    a = np.array([x], dtype=np.float32)
    L.adbo_center_envelope(_abi.ptr(a), C.c_int(1), C.c_int(a.shape[1]))
    np.testing.assert_array_almost_equal(a[0], np.array(expected, np.float32))


# ---- tests/unit_tests/search/selection/test_search_utils.py:9-33 _symetric_limits_1d invariants ------
def test_symetric_limits_invariants(oracle_lib):
    L = oracle_lib.lib()
    rng = np.random.default_rng(0)
    for _ in range(1000):
        n = int(rng.integers(1, 60))
        a = rng.uniform(0, 10, n)
        center = int(rng.integers(0, n))
        f = float(rng.uniform(0.5, 1.0))
        cf = float(rng.uniform(0.01, 0.9))
        mn = int(rng.integers(0, 5))
        mx = int(rng.integers(mn, 20))
        out = np.zeros(2, np.int32)
        L.adbo_symetric_limits_1d(_abi.ptr(a), C.c_int(n), C.c_int(center), C.c_double(f), C.c_double(cf), C.c_int(mn),
                                  C.c_int(mx), out.ctypes.data_as(C.c_void_p))
        assert 0 <= out[0] <= center < out[1] <= n
        assert out[1] - out[0] <= 2 * max(mx, mn) + 1


# ---- tests/unit_tests/search/scoring/test_scoring_utils.py:64-120 calculate_score_groups -------------
def test_score_groups():
    base = dict(precursor_idx=np.arange(10), elution_group_idx=np.array([0, 0, 0, 0, 0, 1, 1, 1, 1, 1]),
                channel=np.array([0, 1, 2, 3, 0, 0, 1, 2, 3, 0]), decoy=np.array([0, 0, 0, 0, 1, 0, 0, 0, 0, 1]))
    assert np.allclose(calculate_score_groups(pd.DataFrame(base))["score_group_idx"].values, np.arange(10))
    got = calculate_score_groups(pd.DataFrame(base), group_channels=True)["score_group_idx"].values
    assert np.allclose(got, np.array([0, 0, 0, 0, 1, 2, 2, 2, 2, 3]))
    df = pd.DataFrame({**base, "rank": np.array([0, 1, 2, 3, 4, 0, 1, 2, 3, 4])})
    assert np.allclose(calculate_score_groups(df, group_channels=True)["score_group_idx"].values, np.arange(10))
    df = pd.DataFrame(dict(precursor_idx=np.arange(10), elution_group_idx=np.array([0, 0, 0, 0, 1, 1, 1, 1, 0, 0]),
                           channel=np.array([0, 0, 1, 1, 0, 0, 1, 1, 0, 0]), decoy=np.array([0, 0, 0, 0, 0, 0, 0, 0, 1, 1]),
                           rank=np.array([0, 1, 0, 1, 0, 1, 0, 1, 0, 1])))
    got = calculate_score_groups(df, group_channels=True)["score_group_idx"].values
    assert np.allclose(got, np.array([0, 0, 1, 1, 2, 3, 4, 4, 5, 5]))


def test_merge_missing_columns():
    left = pd.DataFrame([{"idx": 1, "col_1": 0, "col_2": 0}])
    right = pd.DataFrame([{"idx": 1, "col_3": 0, "col_4": 0}])
    df = merge_missing_columns(left, right, ["col_3"], on="idx")
    assert list(df.columns) == ["idx", "col_1", "col_2", "col_3"]
    with pytest.raises(ValueError):
        merge_missing_columns(left, right, ["col_5"], on="idx")
    with pytest.raises(ValueError):
        merge_missing_columns(left, right, ["col_3"], on=None)
    assert merge_missing_columns(left, right, ["col_1"], on="idx") is left


def test_simple_quadrupole_calibrated_cycle():
    """vectorised get_calibrated_cycle == the reference's per-window loop (restated here verbatim in numpy).
    Checked once against the live reference class through oracle/refshim.py in the build container: cycle_calibrated,
    dia_mz_cycle_calibrated and predict() bit-identical for the 3-D config1 cycle and the 4-D parity_4d cycle."""

    from alphadia_b200.scoring import SimpleQuadrupole, logistic_rectangle
    from alphadia_b200.synthetic import make_config_3d, make_config_4d

    for raw in (make_config_3d("config1")[0], make_config_4d("parity_4d", n_precursors=16)[0]):
        cycle = np.ascontiguousarray(raw.cycle, dtype=np.float64)
        q = SimpleQuadrupole(cycle)
        got = q.jit.cycle_calibrated
        # the reference loop (quadrupole.py:227-258)
        nz = cycle[cycle > 0]
        lo, hi = nz.min(), nz.max()
        space = np.linspace(lo - (hi - lo) * 0.1, hi + (hi - lo) * 0.1, 2000)
        exp = cycle.copy()
        for pr in range(cycle.shape[1]):
            for sc in range(cycle.shape[2]):
                if cycle[0, pr, sc, 0] <= 0:
                    continue
                inten = logistic_rectangle(cycle[0, pr, sc, 0], cycle[0, pr, sc, 1], 0.2, 0.2, space)
                rng_ = space[inten > 0.01]
                exp[0, pr, sc, 0], exp[0, pr, sc, 1] = rng_.min(), rng_.max()
        assert np.array_equal(got, exp)
        assert got.shape == cycle.shape and (got[cycle <= 0] == cycle[cycle <= 0]).all()
        assert (got[..., 0][cycle[..., 0] > 0] < cycle[..., 0][cycle[..., 0] > 0]).all()  # the 1 % threshold widens the window
        P = np.array([1, 1, 2]); S = np.array([0, 0, 0]); X = np.array([cycle[0, 1, 0, 0], cycle[0, 1, 0, :].mean(), 100.0])
        pred = q.predict(P, S, X)
        assert abs(pred[0] - 0.5) < 1e-6 and pred[1] > 0.99 and pred[2] < 1e-6
        assert q.jit.get_dia_mz_cycle(1.0, 2.0).shape == (cycle.shape[1] * cycle.shape[2], 2)
    # fit recovers a shifted, wider window
    rng = np.random.default_rng(0)
    cycle = np.ascontiguousarray(make_config_3d("config1")[0].cycle, dtype=np.float64)
    q = SimpleQuadrupole(cycle)
    P = rng.integers(1, cycle.shape[1], 4000); S = np.zeros(4000, dtype=np.int64)
    X = cycle[0, P, S, 0] + rng.uniform(-3, cycle[0, 1, 0, 1] - cycle[0, 1, 0, 0] + 3, 4000)
    y = logistic_rectangle(cycle[0, P, S, 0] + 0.3, cycle[0, P, S, 1] - 0.2, 0.35, 0.5, X)
    q.fit(P, S, X, y)
    assert np.allclose(q.jit.sigma, [0.35, 0.5], atol=1e-3) and np.allclose(q.jit.delta_mu, [0.3, -0.2], atol=1e-3)


def test_transpose_known_answer(oracle_lib):
    """tests/unit_tests/raw_data/test_raw_data.py:7-23 (exact CSR transpose)."""
    values = np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.uint16)
    tof_indices = np.array([0, 3, 2, 4, 1, 2, 4], dtype=np.uint32)
    push_ptr = np.array([0, 2, 4, 5, 7], dtype=np.int64)
    push_indices, tof_indptr, intensity_values = oracle_lib.transpose_csr(tof_indices, push_ptr, 7, values)
    assert np.array_equal(push_indices, [0, 2, 1, 3, 0, 1, 3])
    assert np.array_equal(tof_indptr, [0, 1, 2, 4, 5, 7, 7, 7])
    assert np.array_equal(intensity_values, [1, 5, 3, 6, 2, 4, 7])


def test_save_corrcoeff_and_fragment_correlation(oracle_lib):
    """tests/unit_tests/search/scoring/test_scoring_utils.py:158-217."""
    import ctypes as C

    L = oracle_lib.lib()
    L.adbo_save_corrcoeff_f32.restype = C.c_double
    f32p = C.POINTER(C.c_float)

    def corr(x, y):
        x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
        return L.adbo_save_corrcoeff_f32(x.ctypes.data_as(f32p), y.ctypes.data_as(f32p), C.c_int(len(x)))

    up = np.arange(1, 11, dtype=np.float32)
    assert np.isclose(corr(up, up[::-1]), -1.0) and np.isclose(corr(up, up), 1.0) and np.isclose(corr(np.zeros(10), np.zeros(10)), 0.0)

    a = np.array([[[1, 2, 3], [1, 2, 3]], [[3, 2, 1], [1, 2, 3]], [[0, 0, 0], [0, 0, 0]]], dtype=np.float32)  # [F=3][nobs=2][n=3]
    out = np.zeros((2, 3, 3), np.float32)
    L.adbo_fragment_correlation(a.ctypes.data_as(f32p), C.c_int(3), C.c_int(2), C.c_int(3), out.ctypes.data_as(f32p))
    expected = np.array([[[1.0, -1.0, 0.0], [-1.0, 1.0, 0.0], [0.0, 0.0, 0.0]], [[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0]]])
    assert np.allclose(out, expected, atol=1e-6)
    z = np.zeros((10, 10, 10), np.float32)
    outz = np.ones((10, 10, 10), np.float32)
    L.adbo_fragment_correlation(z.ctypes.data_as(f32p), C.c_int(10), C.c_int(10), C.c_int(10), outz.ctypes.data_as(f32p))
    assert np.allclose(outz, 0.0)
    # fragment_correlation_different(a, y) with the first fragment's profile as the per-observation template:
    # row o of the result == column 0 of fragment_correlation(a)[o]
    y = np.ascontiguousarray(a[0])  # [nobs][n]
    ct = np.zeros((2, 3), np.float32)
    L.adbo_corr_with_template(a.ctypes.data_as(f32p), y.ctypes.data_as(f32p), C.c_int(3), C.c_int(2), C.c_int(3), ct.ctypes.data_as(f32p))
    assert np.allclose(ct, expected[:, :, 0], atol=1e-6)


def test_search_sorted_left(oracle_lib):
    """tests/unit_tests/search/jitclasses/test_alpharaw_jit.py:6-9 plus the boundaries np.searchsorted(..., 'left') defines."""
    import ctypes as C

    L = oracle_lib.lib()
    L.adbo_search_sorted_left_f32.restype = C.c_int64
    arr = np.arange(100, dtype=np.float32)
    f = lambda v: L.adbo_search_sorted_left_f32(arr.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(len(arr)), C.c_float(v))  # noqa: E731
    assert f(50) == 50
    for v in (-1.0, 0.0, 0.5, 49.5, 99.0, 99.5, 1e9):
        assert f(v) == int(np.searchsorted(arr, np.float32(v), "left"))
    dup = np.array([1, 1, 2, 2, 2, 3], dtype=np.float32)
    assert L.adbo_search_sorted_left_f32(dup.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(6), C.c_float(2.0)) == 2


def test_qtf_regions():
    """tests/unit_tests/search/scoring/test_quadrupole.py:58-79."""
    from alphadia_b200.scoring import SimpleQuadrupole, quadrupole_transfer_function_single

    fake_cycle = np.array([[780.0, 801], [801, 820]])
    fake_cycle = np.repeat(fake_cycle[:, np.newaxis, :], 10, axis=1)[np.newaxis, :, :, :]
    quad = SimpleQuadrupole(fake_cycle)
    isotope_mz = np.array([800.0, 800.1, 800.2, 802.42944, 802.9311, 803.1])
    qtf = quadrupole_transfer_function_single(quad.jit, np.array([0, 1]), np.arange(2, 9), isotope_mz)
    assert qtf.shape == (6, 2, 7)
    assert np.all(qtf[:3, 0, :] > 0.9) and np.all(qtf[:3, 1, :] < 0.1)
    assert np.all(qtf[3:, 0, :] < 0.1) and np.all(qtf[3:, 1, :] > 0.9)


def test_assemble_isotope_mz_typing(oracle_lib):
    """tests/unit_tests/search/selection/test_selection_utils.py:14-36: isotope m/z = float32(mz + i * 1.0033548 / charge)
    computed in float64 and stored as float32 (selection/utils.py:35-40); the oracle's selection uses exactly this value —
    checked through the golden candidate tables — here the arithmetic itself."""
    mz, charge = np.float32(500.123), 2
    off = np.arange(4) * 1.0033548350700006 / charge
    iso = (np.float64(mz) + off).astype(np.float32)
    assert iso.dtype == np.float32 and iso[0] == mz
    assert np.allclose(np.diff(iso.astype(np.float64)), 1.0033548350700006 / charge, atol=1e-4)


def test_collect_candidates_table():
    """tests/unit_tests/search/scoring/test_scoring.py:110-221 — the boundary schema of the feature table: the 46 feature
    columns in order, ids, candidate and precursor columns, delta_rt and residue counts.  Built like the reference's test:
    mocked DiaData and quadrupole, an OutputPsmDF-like object handing over (precursor_idx, rank, features)."""
    from unittest.mock import Mock

    import pandas as pd

    from alphadia_b200.config import CandidateScoringConfig
    from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS, CandidateScoring

    precursors = pd.DataFrame({
        "elution_group_idx": np.array([0, 1], dtype=np.uint32), "precursor_idx": np.array([0, 1], dtype=np.uint32),
        "channel": np.array([0, 0], dtype=np.uint32), "decoy": np.array([0, 1], dtype=np.uint8),
        "flat_frag_start_idx": np.array([0, 5], dtype=np.uint32), "flat_frag_stop_idx": np.array([5, 10], dtype=np.uint32),
        "charge": np.array([2, 3], dtype=np.uint8), "rt_library": np.array([100.0, 200.0], dtype=np.float32),
        "mobility_library": np.array([0.8, 0.9], dtype=np.float32), "mz_library": np.array([500.0, 600.0], dtype=np.float32),
        "proteins": ["P1", "P2"], "genes": ["G1", "G2"], "sequence": ["PEPTIDEK", "ANOTHERR"], "mods": ["", ""],
        "mod_sites": ["", ""], "i_0": np.array([1.0, 1.0], dtype=np.float32),
    })
    fragments = pd.DataFrame({
        "mz_library": np.array([200.0, 300.0, 400.0], dtype=np.float32), "intensity": np.array([1000.0, 2000.0, 1500.0], dtype=np.float32),
        "cardinality": np.array([1, 1, 1], dtype=np.uint8), "type": np.array([0, 1, 0], dtype=np.uint8),
        "loss_type": np.array([0, 0, 0], dtype=np.uint8), "charge": np.array([1, 1, 2], dtype=np.uint8),
        "number": np.array([1, 2, 3], dtype=np.uint8), "position": np.array([1, 2, 3], dtype=np.uint8),
    })
    dia_data = Mock()
    dia_data.cycle = Mock()
    quadrupole = Mock()
    quadrupole.jit = Mock()
    op = CandidateScoring(dia_data=dia_data, precursors_flat=precursors, fragments_flat=fragments, quadrupole_calibration=quadrupole,
                          config=CandidateScoringConfig(), rt_column="rt_library", mobility_column="mobility_library",
                          precursor_mz_column="mz_library", fragment_mz_column="mz_library")
    features = np.array([
        [1.0, 2.0, 100.5, 0.85, 1000.0, 800.0, 1500.0, 900.0, 0.1, 0.05, 500.1, 900.0, 700.0, 1400.0, 850.0,
         0.95, 0.90, 10.0, 0.85, 0.80, 0.75, 0.70, 0.65, 0.60, 0.55, 2000.0, 1800.0, 200.0, 0.15, 0.88, 0.82,
         0.78, 0.92, 0.86, 0.84, 5.0, 0.89, 6.0, 20.0, 15.0, 2.5, 0.02, 0.01, 3.0, 500.0, 0.03],
        [1.5, 2.5, 200.5, 0.90, 1100.0, 850.0, 1600.0, 950.0, 0.15, 0.08, 600.1, 950.0, 750.0, 1450.0, 900.0,
         0.96, 0.91, 12.0, 0.87, 0.82, 0.77, 0.72, 0.67, 0.62, 0.57, 2100.0, 1900.0, 200.0, 0.18, 0.90, 0.84,
         0.80, 0.94, 0.88, 0.86, 6.0, 0.91, 7.0, 22.0, 17.0, 3.0, 0.025, 0.015, 4.0, 550.0, 0.035],
    ], dtype=np.float32)
    psm = Mock()
    psm.to_precursor_df.return_value = (np.array([0, 1]), np.array([1, 1]), features)
    candidates = pd.DataFrame({
        "precursor_idx": [0, 1], "rank": [1, 1], "elution_group_idx": [0, 1], "scan_center": [100, 200],
        "scan_start": [90, 190], "scan_stop": [110, 210], "frame_center": [50, 100], "frame_start": [45, 95],
        "frame_stop": [55, 105],
    })
    result = op.collect_candidates(candidates, psm)

    assert list(result.columns[:46]) == DEFAULT_FEATURE_COLUMNS and list(result.columns[46:48]) == ["precursor_idx", "rank"]
    assert np.array_equal(result[DEFAULT_FEATURE_COLUMNS].values, features)
    expected = {
        "precursor_idx": [0, 1], "rank": [1, 1],
        "elution_group_idx": [0, 1], "frame_center": [50, 100], "frame_stop": [55, 105], "scan_stop": [110, 210],
        "frame_start": [45, 95], "scan_start": [90, 190], "scan_center": [100, 200],
        "flat_frag_stop_idx": [5, 10], "mod_sites": ["", ""], "mz_library": [500.0, 600.0], "decoy": [0, 1], "charge": [2, 3],
        "mods": ["", ""], "rt_library": [100.0, 200.0], "proteins": ["P1", "P2"], "channel": [0, 0], "genes": ["G1", "G2"],
        "flat_frag_start_idx": [0, 5], "mobility_library": [0.8, 0.9], "i_0": [1.0, 1.0], "sequence": ["PEPTIDEK", "ANOTHERR"],
        "delta_rt": [0.5, 0.5], "n_K": [1, 0], "n_R": [0, 2], "n_P": [2, 0],
    }
    assert set(result.columns) == set(DEFAULT_FEATURE_COLUMNS) | set(expected)
    assert list(result.columns[-4:]) == ["delta_rt", "n_K", "n_R", "n_P"]
    exp_df = pd.DataFrame(expected)[list(result.columns[46:])]
    pd.testing.assert_frame_equal(result[list(result.columns[46:])], exp_df, check_dtype=False)
    # the array form adb_score_candidates fills gives the same table, invalid rows dropped
    arrays = dict(features=np.vstack([features, np.zeros((1, 46), np.float32)]), valid=np.array([1, 1, 0], np.uint8),
                  precursor_idx=np.array([0, 1, 1], np.uint32), rank=np.array([1, 1, 2], np.uint8))
    pd.testing.assert_frame_equal(op.collect_candidates(candidates, arrays), result, check_dtype=False)


# ---- tests/unit_tests/fdr/test_fdr.py:12-123 keep_best, _fdr_to_q_values, get_q_values -----------------------
def test_fdr_known_answers(oracle_lib, monkeypatch):
    """The reference's own vectors through the host mirror alphadia_b200.fdr; the two device calls are replaced by the
    oracle here (no GPU), tests/test_gpu_parity.py::test_fdr_bookkeeping replays them on the device."""
    from tests.test_oracle_golden import host_fdr_with_oracle

    fdr = host_fdr_with_oracle(monkeypatch, oracle_lib)
    for check in FDR_KNOWN_ANSWERS:
        check(fdr)
    q = oracle_lib.fdr_to_q_values(np.array([0.2, 0.1, 0.05, 0.3, 0.26, 0.25, 0.5]))
    assert np.allclose(q, np.array([0.05, 0.05, 0.05, 0.25, 0.25, 0.25, 0.5]))


def _known_keep_best(fdr):
    test_df = pd.DataFrame({"precursor_idx": [0, 0, 0, 1, 1, 1, 2, 2, 2], "channel": [0, 0, 1, 0, 1, 1, 0, 0, 1],
                            "proba": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]})
    best = fdr.keep_best(test_df, score_column="proba", group_columns=["precursor_idx"])
    assert best.shape[0] == 3 and np.allclose(best["proba"].values, [0.1, 0.4, 0.7])
    best = fdr.keep_best(test_df, score_column="proba", group_columns=["channel", "precursor_idx"])
    assert best.shape[0] == 6 and np.allclose(best["proba"].values, [0.1, 0.3, 0.4, 0.5, 0.7, 0.9])


def _known_keep_best_2(fdr):
    test_df = pd.DataFrame({"channel": [0, 0, 0, 4, 4, 4, 8, 8, 8], "elution_group_idx": [0, 1, 2, 0, 1, 2, 0, 1, 2],
                            "proba": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.1, 0.2, 0.3]})
    pd.testing.assert_frame_equal(fdr.keep_best(test_df, group_columns=["channel", "elution_group_idx"]), test_df)
    for col in ("elution_group_idx", "precursor_idx"):
        test_df = pd.DataFrame({"channel": [0, 0, 0, 4, 4, 4, 8, 8, 8], col: [0, 0, 1, 0, 0, 1, 0, 0, 1],
                                "proba": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.1, 0.2, 0.3]})
        expected = pd.DataFrame({"channel": [0, 0, 4, 4, 8, 8], col: [0, 1, 0, 1, 0, 1], "proba": [0.1, 0.3, 0.4, 0.6, 0.1, 0.3]})
        pd.testing.assert_frame_equal(fdr.keep_best(test_df, group_columns=["channel", col]), expected)


def _known_q_values(fdr):
    test_df = pd.DataFrame({"precursor_idx": [0, 1, 2, 3, 4, 5, 6, 7, 8, 9],
                            "proba": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0], "_decoy": [0, 0, 0, 1, 0, 0, 1, 1, 1, 1]})
    out = fdr.get_q_values(test_df, "proba", "_decoy")
    assert np.allclose(out["qval"].values, [0.0, 0.0, 0.0, 0.2, 0.2, 0.2, 0.4, 0.6, 0.8, 1.0])
    assert list(out.columns) == ["precursor_idx", "proba", "_decoy", "qval"] and "qval" not in test_df.columns
    empty = fdr.get_q_values(test_df.iloc[:0], "proba", "_decoy")
    assert len(empty) == 0 and "qval" in empty.columns
    assert len(fdr.keep_best(test_df.iloc[:0], group_columns=["precursor_idx"])) == 0


FDR_KNOWN_ANSWERS = [_known_keep_best, _known_keep_best_2, _known_q_values]


def _plan_with_pandas(psm_df, frag_df, cycle):
    """The preparation of fragcomp.py:254-273 / fragcomp/utils.py:10-45 spelled out with group-by, merge and the n x positions
    window matrix - the specification the run-boundary implementation is held to."""
    psm_df["_candidate_idx"] = candidate_hash(psm_df["precursor_idx"].values, psm_df["rank"].values)
    frag_df["_candidate_idx"] = candidate_hash(frag_df["precursor_idx"].values, frag_df["rank"].values)
    frag_df["frag_idx"] = np.arange(len(frag_df))
    ext = frag_df.groupby("_candidate_idx", as_index=False).agg(_frag_start_idx=("frag_idx", "min"), _frag_stop_idx=("frag_idx", "max"))
    ext["_frag_stop_idx"] += 1
    psm_df = psm_df.merge(ext, "inner", on="_candidate_idx")
    lower = np.min(cycle[0, :, :, 0], axis=1, keepdims=True).T
    upper = np.max(cycle[0, :, :, 1], axis=1, keepdims=True).T
    mz = psm_df["mz_observed"].values[:, None]
    psm_df["window_idx"] = np.argmax((mz >= lower) & (mz < upper), axis=1)
    psm_df = psm_df.sort_values(by=["window_idx", "proba", "precursor_idx"])
    pos = pd.DataFrame({"window_idx": psm_df["window_idx"].values, "row": np.arange(len(psm_df))})
    win = pos.groupby("window_idx", as_index=False).agg(start=("row", "min"), stop=("row", "max"))
    return psm_df, win["start"].values, win["stop"].values + 1


def test_fragment_competition_preparation_vs_pandas_formulation():
    """Random adversarial tables: duplicate candidates, PSMs without fragments, fragments of one candidate in several runs,
    m/z outside every window or NaN, overlapping and empty (-1, -1) cycle positions, shuffled indices, empty inputs."""
    rng = np.random.default_rng(0)
    for it in range(120):
        n, n_pos, n_scan = int(rng.integers(0, 60)), int(rng.integers(1, 8)), int(rng.integers(1, 4))
        psm = pd.DataFrame({"precursor_idx": rng.integers(0, 25, n), "rank": rng.integers(0, 3, n).astype(np.uint8),
                            "rt_observed": rng.random(n) * 50, "proba": np.round(rng.random(n), 1),
                            "mz_observed": np.where(rng.random(n) < 0.1, np.nan, rng.uniform(50, 450, n))})
        if rng.random() < 0.5:
            psm = psm.drop_duplicates(["precursor_idx", "rank"])
        psm.index = rng.permutation(len(psm)) + 7
        m = int(rng.integers(0, 200))
        frag = pd.DataFrame({"precursor_idx": rng.integers(0, 28, m), "rank": rng.integers(0, 3, m).astype(np.uint8),
                             "mz_observed": rng.uniform(100, 1000, m)})
        mode = int(rng.integers(0, 3))
        if mode == 0:
            frag = frag.sort_values(["precursor_idx", "rank"]).reset_index(drop=True)
        elif mode == 1:
            frag = frag.sort_values(["rank", "precursor_idx"], kind="stable").reset_index(drop=True)
        cycle = np.zeros((1, n_pos, n_scan, 2))
        lo = np.sort(rng.uniform(0, 400, n_pos))
        for pos in range(n_pos):
            for s in range(n_scan):
                a = lo[pos] + rng.uniform(-20, 20)
                cycle[0, pos, s] = (a, a + rng.uniform(10, 150))
        if rng.random() < 0.5:
            cycle[0, 0] = -1
        psm_a, frag_a, psm_b, frag_b = psm.copy(), frag.copy(), psm.copy(), frag.copy()
        plan = FragmentCompetition().plan(psm_a, frag_a, cycle)
        want, start, stop = _plan_with_pandas(psm_b, frag_b, cycle)
        pd.testing.assert_frame_equal(plan.psm_df, want)
        pd.testing.assert_frame_equal(psm_a, psm_b)  # side effects on the caller's frames
        pd.testing.assert_frame_equal(frag_a, frag_b)
        assert np.array_equal(plan.window_start, start) and np.array_equal(plan.window_stop, stop), it
        assert np.array_equal(plan.frag_start, want["_frag_start_idx"].values) and np.array_equal(plan.frag_stop, want["_frag_stop_idx"].values)

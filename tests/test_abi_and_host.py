"""CPU-only checks: the C-ABI library loads and exports every symbol include/alphadia_b200.h declares; the
product fails loudly without a GPU; host-side mirrors (configs, schemas, kernel, sharding) behave like the
reference.  No device compute."""

import ctypes as C
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from alphadia_b200 import build

    build.build()
    from alphadia_b200 import _lib

    return _lib


def test_header_symbols_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "alphadia_b200.h")).read()
    declared = set(re.findall(r"\b(adb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 19
    lib = built_lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(built_lib.EXPORTED_SYMBOLS)
    assert lib.adb_version().decode().startswith("alphadia_b200")


def test_struct_sizes_match_header(built_lib):
    """ctypes mirrors must have the C layout (sizes computed by hand from the header's field lists)."""
    from alphadia_b200 import _abi

    assert C.sizeof(_abi.RawFile3DDesc) == 15 * 8
    assert C.sizeof(_abi.RawFile4DDesc) == 17 * 8
    assert C.sizeof(_abi.CandidateTable) == 11 * 8
    assert C.sizeof(_abi.LibraryDesc) == 8 + 8 * 8 + 8 + 8 + 9 * 8
    assert C.sizeof(_abi.CandidatesOut) == 10 * 8
    assert C.sizeof(_abi.CandidatesIn) == 9 * 8
    assert C.sizeof(_abi.ScoresOut) == 14 * 8
    assert C.sizeof(_abi.ScoringConfig) == 9 * 4 + 4 + 4 * 8
    assert C.sizeof(_abi.SelectionConfig) == 22 * 8  # the two adjacent int32 flags share one 8-byte slot


def test_no_cpu_fallback(built_lib):
    if built_lib.load().adb_device_count() > 0:
        pytest.skip("a GPU is visible")
    from alphadia_b200 import CandidateSelection, FragmentCompetition
    from tests import helpers as H

    raw, pdf, fdf, lib, p = H.workload("config1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        built_lib.DeviceRawFile(raw)
    sel = CandidateSelection(raw, pdf.copy(), fdf.copy(), H.selection_config(30.0), rt_column="rt_library",
                             mobility_column="mobility_library", precursor_mz_column="mz_library",
                             fragment_mz_column="mz_library")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sel()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        built_lib.fragment_competition([0], [1], np.zeros(1, np.float32), [0], [1], np.ones(1, np.float32), 3.0, 15.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        built_lib.transpose_csr(np.zeros(1, np.uint32), np.array([0, 1], np.int64), 4, np.ones(1, np.uint16))
    assert FragmentCompetition is not None


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "alphadia_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "adb_oracle" not in src, f
                assert "/root/reference" not in src, f


# ---- config semantics (jit_config.py:84-138) ------------------------------------------------------
def test_config_update_semantics():
    from alphadia_b200.config import CandidateScoringConfig, CandidateSelectionConfig

    c = CandidateScoringConfig()
    c.update({"precursor_mz_tolerance": 4.7, "quant_all": True})
    assert c.precursor_mz_tolerance == 4 and isinstance(c.precursor_mz_tolerance, int)  # int default truncates
    with pytest.raises(ValueError):
        c.update({"does_not_exist": 1})
    s = CandidateSelectionConfig()
    s.update({"rt_tolerance": 30, "candidate_count": 3.0})
    assert isinstance(s.rt_tolerance, float) and s.candidate_count == 3
    with pytest.raises(ValueError):
        s.update({"feature_std": np.ones(2)})
    st = s.to_struct()
    assert st.rt_tolerance == 30.0 and st.candidate_count == 3 and st.kernel_size == 30
    sc = c.to_struct()
    assert sc.top_k_fragments == 12 and sc.quant_all == 1 and abs(sc.quad_sigma[0] - 0.2) < 1e-12
    c.update({"top_k_fragments": 9999})  # transfer-library requantification: served by adb_score_candidates_ragged
    assert c.to_struct().top_k_fragments == 9999


def test_schema_validation_casts_and_raises():
    from alphadia_b200.validation import candidates_schema, fragments_flat_schema

    df = pd.DataFrame({"mz_library": [500.0], "intensity": [1.0], "cardinality": [1], "type": [98], "loss_type": [0],
                       "charge": [1], "number": [3], "position": [2]})
    fragments_flat_schema.validate(df)
    assert df["mz_library"].dtype == np.float32 and df["type"].dtype == np.uint8
    with pytest.raises(ValueError, match="rank"):
        candidates_schema.validate(pd.DataFrame({"elution_group_idx": [0], "precursor_idx": [0]}))


@pytest.mark.parametrize("kernel_size,fwhm_rt,sigma_scale_rt", [(30, 5.0, 0.5), (30, 10.0, 1.0), (20, 2.0, 0.5), (31, 5.0, 0.1)])
def test_gaussian_kernel_properties(kernel_size, fwhm_rt, sigma_scale_rt):
    """tests/unit_tests/search/selection/test_kernel.py:35-67: even dims, finite, float32."""
    from alphadia_b200.kernel import GaussianKernel
    from tests import helpers as H

    raw = H.workload("config1")[0]
    k = GaussianKernel(raw, fwhm_rt=fwhm_rt, sigma_scale_rt=sigma_scale_rt, kernel_width=kernel_size,
                       kernel_height=min(kernel_size, raw.scan_max_index + 1)).get_dense_matrix(verbose=False)
    assert k.dtype == np.float32 and k.shape[0] % 2 == 0 and k.shape[1] % 2 == 0
    assert np.all(np.isfinite(k)) and k.shape[0] == 2
    assert np.argmax(k[1]) == k.shape[1] // 2


# ---- sharding --------------------------------------------------------------------------------------
def test_shard_library_partitions_exactly():
    from alphadia_b200.sharding import shard_bounds, shard_library
    from tests import helpers as H

    raw, pdf, fdf, lib, p = H.workload("config1")
    assert list(shard_bounds(10, 3)) == [0, 4, 7, 10]
    assert list(shard_bounds(2, 4)) == [0, 1, 2, 2, 2]
    seen = []
    for r in range(3):
        sp, sf = shard_library(pdf, fdf, r, 3)
        seen.append(sp["precursor_idx"].values)
        s0, s1 = sp["flat_frag_start_idx"].values, sp["flat_frag_stop_idx"].values
        assert s0[0] == 0 and s1[-1] == len(sf)
        # fragments of every precursor are unchanged
        i = len(sp) // 2
        orig = pdf[pdf["precursor_idx"] == sp["precursor_idx"].values[i]].iloc[0]
        a = fdf["mz_library"].values[int(orig["flat_frag_start_idx"]): int(orig["flat_frag_stop_idx"])]
        assert np.array_equal(a, sf["mz_library"].values[s0[i]: s1[i]])
    assert np.array_equal(np.concatenate(seen), np.sort(pdf["precursor_idx"].values))


def test_pack_unpack_score_table():
    from alphadia_b200.sharding import pack_score_table, unpack_score_table

    rng = np.random.default_rng(0)
    f = rng.normal(size=(17, 46)).astype(np.float32)
    f[3, 5] = np.nan
    pidx = rng.integers(0, 2**32 - 1, 17, dtype=np.uint64).astype(np.uint32)
    rank = rng.integers(0, 5, 17).astype(np.uint8)
    valid = rng.integers(0, 2, 17).astype(np.uint8)
    u = unpack_score_table(pack_score_table(f, pidx, rank, valid))
    assert np.array_equal(u["features"].view(np.uint32), f.view(np.uint32))
    assert np.array_equal(u["precursor_idx"], pidx) and np.array_equal(u["rank"], rank) and np.array_equal(u["valid"], valid)


def test_sharded_selection_equals_unsharded(oracle_lib):
    """Any partition of the precursors gives the same candidates (disjoint output rows)."""
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.sharding import shard_library
    from tests import helpers as H

    raw, pdf, fdf, lib, p = H.workload("config1")
    cfg = H.selection_config(p["rt_tolerance"]).to_struct()
    kernel = H.default_kernel(raw)
    full = oracle_lib.select_candidates(raw, lib, cfg, kernel)
    parts = []
    for r in range(2):
        sp, sf = shard_library(pdf, fdf, r, 2)
        sl = assemble_library_arrays(sp, sf, "rt_library", "mobility_library", "mz_library", "mz_library")
        parts.append(oracle_lib.select_candidates(raw, sl, cfg, kernel))
    for c in full:
        assert np.array_equal(full[c], np.concatenate([q[c] for q in parts])), c


def test_fast_merge_equals_pandas_merge():
    """merge_missing_columns' gather path must return exactly what DataFrame.merge(how='left') returns."""
    import pandas as pd

    from alphadia_b200.scoring import _merge_by_sorted_key, merge_missing_columns

    rng = np.random.default_rng(0)
    right = pd.DataFrame({"precursor_idx": rng.permutation(5000).astype(np.uint32), "a": rng.normal(size=5000).astype(np.float32),
                          "b": rng.integers(0, 9, 5000).astype(np.uint8), "s": np.array([f"PEP{k}K" for k in range(5000)], dtype=object)})
    left = pd.DataFrame({"precursor_idx": rng.integers(0, 5000, 20000).astype(np.uint32), "x": rng.normal(size=20000)})
    left.index = left.index[::-1]  # a non-default index must not leak into the result
    fast = merge_missing_columns(left, right, ["a", "b", "s", "x"], on="precursor_idx")
    slow = left.merge(right[["precursor_idx", "a", "b", "s"]], on=["precursor_idx"], how="left")
    pd.testing.assert_frame_equal(fast, slow)
    # sorted right, two keys
    cand = pd.DataFrame({"precursor_idx": np.repeat(np.arange(3000, dtype=np.uint32), 3), "rank": np.tile(np.arange(3, dtype=np.uint8), 3000)})
    cand["score"] = rng.normal(size=len(cand)).astype(np.float32)
    psm = cand.sample(frac=0.7, random_state=1)[["precursor_idx", "rank"]].reset_index(drop=True)
    fast = merge_missing_columns(psm, cand, ["score"], on=["precursor_idx", "rank"])
    slow = psm.merge(cand[["precursor_idx", "rank", "score"]], on=["precursor_idx", "rank"], how="left")
    pd.testing.assert_frame_equal(fast, slow)
    # preconditions that must fall back to pandas: unmatched keys, duplicate right keys
    assert _merge_by_sorted_key(pd.DataFrame({"precursor_idx": np.array([7000], np.uint32)}), right, ["precursor_idx"], ["a"]) is None
    dup = pd.concat([right, right.iloc[:1]])
    assert _merge_by_sorted_key(left, dup, ["precursor_idx"], ["a"]) is None
    out = merge_missing_columns(pd.DataFrame({"precursor_idx": np.array([1, 7000], np.uint32)}), right, ["a"], on="precursor_idx")
    assert np.isnan(out["a"].values[1])


def test_count_residues_equals_str_count():
    import pandas as pd

    from alphadia_b200.scoring import count_residues

    rng = np.random.default_rng(1)
    alphabet = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    seqs = np.array(["".join(rng.choice(alphabet, size=rng.integers(0, 40))) for _ in range(3000)], dtype=object)
    col = seqs[rng.integers(0, len(seqs), 20000)]
    for r in "KRP":
        assert np.array_equal(count_residues(col, r), pd.Series(col).str.count(r).values)
    for got, r in zip(count_residues(col, ["K", "R", "P"]), "KRP"):
        assert np.array_equal(got, pd.Series(col).str.count(r).values)
    with_nan = col.copy()
    with_nan[5] = np.nan
    a, b = count_residues(with_nan, "K"), pd.Series(with_nan).str.count("K").values
    assert np.array_equal(np.isnan(a.astype(float)), np.isnan(b.astype(float))) and np.array_equal(a[:5], b[:5])


def test_calculate_score_groups_equals_pandas_sort():
    """The lexsort / already-sorted shortcut must reproduce the reference's two sort_values calls (scoring/utils.py:269-410)."""
    import pandas as pd

    from alphadia_b200.scoring import calculate_score_groups

    def reference(df, group_channels):
        df = df.sort_values(by=["elution_group_idx", "decoy", "rank", "precursor_idx"])
        if group_channels:
            eg, dc, rk = df["elution_group_idx"].values, df["decoy"].values, df["rank"].values
            change = np.zeros(len(df), dtype=bool)
            change[1:] = (eg[1:] != eg[:-1]) | (dc[1:] != dc[:-1]) | (rk[1:] != rk[:-1])
            df["score_group_idx"] = np.cumsum(change).astype(np.uint32)
        else:
            df["score_group_idx"] = np.arange(len(df), dtype=np.uint32)
        return df.sort_values(by=["score_group_idx", "precursor_idx"]).reset_index(drop=True)

    rng = np.random.default_rng(2)
    n = 5000
    base = pd.DataFrame({"precursor_idx": rng.permutation(n).astype(np.uint32), "elution_group_idx": rng.integers(0, 900, n).astype(np.uint32),
                         "decoy": rng.integers(0, 2, n).astype(np.uint8), "rank": rng.integers(0, 3, n).astype(np.uint8),
                         "x": rng.normal(size=n)})
    already = base.sort_values(by=["elution_group_idx", "decoy", "rank", "precursor_idx"]).reset_index(drop=True)
    for df in (base, already, base.iloc[:1], base.iloc[:0]):
        for gc in (False, True):
            pd.testing.assert_frame_equal(calculate_score_groups(df.copy(), group_channels=gc), reference(df.copy(), gc))


def test_collect_fragments_equals_mask_gather_and_merge():
    """collect_fragments (threaded column gathers, per-candidate lookup) == the reference's flat mask + merge
    (output.py:72-90, scoring.py:520-580) for ragged, full, empty and duplicated-library cases."""
    import pandas as pd

    from alphadia_b200 import _abi
    from alphadia_b200.scoring import FRAGMENT_COLUMNS, CandidateScoring
    from tests import helpers as H

    raw, pdf, fdf, lib, p = H.workload("config1")
    op = CandidateScoring(dia_data=raw, precursors_flat=pdf, fragments_flat=fdf, rt_column="rt_library",
                          mobility_column="mobility_library", precursor_mz_column="mz_library", fragment_mz_column="mz_library")

    def reference(psm, precursors):
        mask = psm["fragment_mz_library"].reshape(-1) > 0
        top_k = psm["fragment_mz_library"].shape[1]
        data = {"precursor_idx": np.repeat(psm["precursor_idx"], top_k)[mask], "rank": np.repeat(psm["rank"], top_k)[mask]}
        for col in FRAGMENT_COLUMNS[2:]:
            data[col] = psm["fragment_" + col].reshape(-1)[mask]
        return pd.DataFrame(data).merge(precursors[["precursor_idx", "elution_group_idx", "decoy"]], on=["precursor_idx"], how="left")

    def make_psm(n, top_k, fill, seed):
        rng = np.random.default_rng(seed)
        _, psm = _abi.alloc_scores_out(n, top_k)
        for k, v in psm.items():
            if k.startswith("fragment_"):
                v[:] = rng.integers(1, 200, v.shape).astype(v.dtype)
        psm["fragment_mz_library"][rng.random((n, top_k)) >= fill] = 0
        if n > 3:
            psm["fragment_mz_library"][n // 2] = 0  # a candidate without any row
        psm["precursor_idx"] = rng.integers(0, len(pdf), n).astype(np.uint32)
        psm["rank"] = rng.integers(0, 3, n).astype(np.uint8)
        return psm

    for n, top_k, fill in [(500, 12, 0.7), (500, 12, 1.1), (40, 6, 0.0), (0, 12, 0.5), (1, 12, 0.5), (100_000, 12, 0.9)]:
        psm = make_psm(n, top_k, fill, seed=n + top_k)
        if fill > 1:
            psm["fragment_mz_library"][:] = 1.0
        got, ref = op.collect_fragments(None, psm), reference(psm, op.precursors_flat_df)
        pd.testing.assert_frame_equal(got, ref)
        assert list(got.columns) == FRAGMENT_COLUMNS + ["elution_group_idx", "decoy"]
    # a library that repeats a precursor_idx multiplies rows in the reference's merge: same here
    dup = pd.concat([op.precursors_flat_df, op.precursors_flat_df.iloc[:3]], ignore_index=True)
    op._precursors_flat_df = dup
    psm = make_psm(300, 12, 0.5, seed=9)
    psm["precursor_idx"][:50] = dup["precursor_idx"].values[0]
    pd.testing.assert_frame_equal(op.collect_fragments(None, psm), reference(psm, dup))


def test_warn_on_critical_values_counts(caplog):
    """NaN / Inf warnings per float column (validation/base.py:120-152), also for strided views of one shared matrix."""
    import logging

    import pandas as pd

    from alphadia_b200.validation import Schema

    m = np.ones((1000, 5), dtype=np.float32)
    m[3, 1] = np.nan
    m[4:6, 1] = np.inf
    m[7, 4] = -np.inf
    df = pd.DataFrame(m, columns=list("abcde"), copy=False)
    df["f"] = np.full(1000, 3.0e38, dtype=np.float32)  # finite values whose f32 sum overflows: no warning
    df["g"] = np.arange(1000)
    df["h"] = pd.Series(["x"] * 1000, dtype=object)
    with caplog.at_level(logging.WARNING):
        Schema._warn_on_critical_values(df)
    msgs = [r.getMessage() for r in caplog.records]
    assert sum("b has 1 NaNs" in x for x in msgs) == 1 and sum("b has 2 Infs" in x for x in msgs) == 1
    assert sum("e has 1 Infs" in x for x in msgs) == 1 and len(msgs) == 3
    caplog.clear()
    clean = pd.DataFrame(np.ones((1000, 5), dtype=np.float32), columns=list("abcde"), copy=False)
    with caplog.at_level(logging.WARNING):
        Schema._warn_on_critical_values(clean)
        Schema._warn_on_critical_values(clean.iloc[:0])
    assert not caplog.records


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs without a GPU and prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "precursor candidates scored/sec" and d["unit"] == "candidates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_adapt_dia_data_variants():
    """The adapters accept the reference's objects: public or private attribute names, 3-D and timsTOF layouts."""
    from types import SimpleNamespace

    from alphadia_b200 import _abi
    from alphadia_b200.raw_data import adapt_dia_data
    from alphadia_b200.synthetic import make_config_3d, make_config_4d

    raw3 = make_config_3d("config1")[0]
    a = adapt_dia_data(raw3)
    assert not a.is_4d and a.rt_values.dtype == np.float32 and a.cycle.shape == raw3.cycle.shape
    # an AlphaRaw wrapper keeps its arrays under private names (alphadia/raw_data/alpharaw_wrapper.py:86-156)
    private = SimpleNamespace(cycle=raw3.cycle, rt_values=raw3.rt_values, mobility_values=raw3.mobility_values,
                              _peak_start_idx_list=raw3.peak_start_idx_list, _peak_stop_idx_list=raw3.peak_stop_idx_list,
                              _mz_values=raw3.mz_values, _intensity_values=raw3.intensity_values, _zeroth_frame=0,
                              _scan_max_index=1, has_mobility=False)
    b = adapt_dia_data(private)
    assert np.array_equal(b.mz_values, a.mz_values) and b.precursor_cycle_max_index == a.precursor_cycle_max_index
    assert b.frame_max_index == len(raw3.rt_values) - 1
    d3, _ = _abi.make_rawfile3d_desc(b)
    assert d3.cycle_len == raw3.cycle.shape[1] and d3.n_peaks == len(raw3.mz_values)
    raw4 = make_config_4d("parity_4d", n_precursors=8)[0]
    c = adapt_dia_data(raw4)
    assert c.is_4d and c.rt_values.dtype == np.float64 and c.push_indices.dtype == np.uint32 and c.intensity_values.dtype == np.uint16
    d4, _ = _abi.make_rawfile4d_desc(c)
    assert d4.frames_per_cycle == raw4.cycle.shape[1] and d4.scans == raw4.cycle.shape[2] and d4.n_events == raw4.n_events
    private4 = SimpleNamespace(_cycle=raw4.cycle, _rt_values=raw4.rt_values, _mobility_values=raw4.mobility_values, _mz_values=raw4.mz_values,
                               _tof_indptr=raw4.tof_indptr, _push_indices=raw4.push_indices, _intensity_values=raw4.intensity_values,
                               _dia_precursor_cycle=raw4.dia_precursor_cycle, _zeroth_frame=1, _scan_max_index=raw4.scan_max_index,
                               _frame_max_index=raw4.frame_max_index, has_mobility=True)
    e = adapt_dia_data(private4)
    assert e.is_4d and e.precursor_cycle_max_index == raw4.precursor_cycle_max_index
    with pytest.raises(ValueError, match="timsTOF CSR arrays"):
        adapt_dia_data(SimpleNamespace(has_mobility=True, cycle=raw4.cycle))
    bad = SimpleNamespace(**{**vars(c), "dia_precursor_cycle": c.dia_precursor_cycle[:-1]})
    with pytest.raises(ValueError):
        _abi.make_rawfile4d_desc(bad)


def test_count_residues_arrow_backed_strings_and_nan_columns_in_validation(caplog):
    """The Arrow-backed ``str`` columns of pandas >= 3 are counted without a trip through Python objects; the per-column
    NaN / Inf warnings of Schema.validate come out of one pass over a shared feature matrix."""
    import logging

    import pandas as pd

    from alphadia_b200.scoring import count_residues
    from alphadia_b200.validation import Schema

    rng = np.random.default_rng(2)
    alphabet = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    seqs = pd.Series(["".join(rng.choice(alphabet, size=rng.integers(0, 30))) for _ in range(5000)], dtype="str")
    for got, r in zip(count_residues(seqs.array, ["K", "R", "P"]), "KRP"):
        assert got.dtype == np.int64 and np.array_equal(got, seqs.str.count(r).values)
    block = rng.random((300_000, 16)).astype(np.float32)  # > 2^22 elements: the threaded column sums
    block[::7, 3] = np.nan
    block[5, 9] = np.inf
    df = pd.DataFrame(block, columns=[f"f{j}" for j in range(16)], copy=False)
    df["extra"] = np.float32(1.0)
    with caplog.at_level(logging.WARNING):
        Schema("t", [])._warn_on_critical_values(df)
    msgs = [r.getMessage() for r in caplog.records]
    assert any(m.startswith("f3 has 42858 NaNs") for m in msgs) and any(m.startswith("f9 has 1 Infs") for m in msgs)
    assert len(msgs) == 2


def test_bench_other_workload_lines_never_break_the_main_line(monkeypatch):
    """bench.py condenses the child runs of config 2 / config 4 into the main line; a failing or hanging child is recorded."""
    import json
    import subprocess
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    child = {"metric": "m", "value": 2.0, "unit": "candidates/s", "steps": 3, "warmup": 3, "ms_per_step": 1.5, "gpu_launches": 7,
             "config": {"workload": "w", "stage_ms": {"selection": 1.0, "scoring": 0.5}, "candidates_per_step": 3.0},
             "e2e": {"value": 1.0}, "roofline": {"frac": 0.1, "kernel": "k"},
             "parity": {"n": 5, "int_exact": True, "selection_score_bit_exact": True, "valid_exact": True, "max_rel": 0.0, "tolerance": 1e-4,
                        "checked": "long text that is not copied"}}
    calls = []

    def fake_run(cmd, **kw):
        calls.append(cmd)
        name = cmd[cmd.index("--workload") + 1]
        if name == "config4":
            raise subprocess.TimeoutExpired(cmd, kw.get("timeout"))
        return subprocess.CompletedProcess(cmd, 0, stdout="noise\n" + json.dumps(child) + "\n", stderr="")

    monkeypatch.setattr(subprocess, "run", fake_run)
    out = bench.other_workload_lines(("config2", "config4", "config9"))
    assert out["config2"]["value"] == 2.0 and out["config2"]["e2e"] == 1.0 and out["config2"]["parity"]["int_exact"] is True
    assert "checked" not in out["config2"]["parity"] and out["config2"]["stage_ms"] == {"selection": 1.0, "scoring": 0.5}
    assert "TimeoutExpired" in out["config4"]["error"]
    assert all("--no-other-workloads" in c and "--no-cpu-baseline" in c for c in calls)  # children never recurse
    monkeypatch.setattr(subprocess, "run", lambda cmd, **kw: subprocess.CompletedProcess(cmd, 1, stdout="", stderr="boom"))
    assert "error" in bench.other_workload_lines(("config2",))["config2"]

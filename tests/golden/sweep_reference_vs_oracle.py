#!/usr/bin/env python
"""Randomized pinning of the oracle against the LIVE reference (this container only: needs /root/reference and numba).

The committed golden vectors pin ``oracle/adb_oracle.c`` at fixed seeds and configurations.  This script draws random raw files,
libraries (some ragged), selection and scoring configurations and quadrupole parameters, runs the unmodified numba path through
``oracle/refshim.py`` and the oracle on the same inputs, and holds the oracle to the bar of tests/test_oracle_golden.py:
candidate table bit-exact (integer columns and f32 score), valid rows equal, features bit-exact except the BLAS-summed ones
(1e-4 relative + 1e-6 absolute), per-fragment columns bit-exact except the correlation (same tolerance).

    python tests/golden/sweep_reference_vs_oracle.py [n_cases] [seed] > tests/golden/sweep_reference_vs_oracle.json
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from alphadia_b200.library import assemble_library_arrays  # noqa: E402
from alphadia_b200.synthetic import make_config_3d, make_config_4d  # noqa: E402
from oracle import refshim  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_oracle_golden import BLAS_FEATURES, FRAG_MAP, INT_COLS  # noqa: E402


def _excess(a, b):
    """Largest |a - b| in units of the tolerance of BASELINE.json's north star for the BLAS-summed quantities (correlations, whose
    summation order numpy / BLAS leave open): 1e-4 relative plus an absolute 1e-6 (they live in [-1, 1]; a value of 1e-8 cannot be
    held to 1e-4 relative)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan, 0.0, np.abs(a - b))
    return float(np.max(d / (1e-4 * np.maximum(np.abs(a), np.abs(b)) + 1e-6))) if d.size else 0.0


def draw_case(rng, it):
    """All random choices of one case (cheap; lets a later case be reproduced without running the earlier ones)."""
    is4d = it % 3 == 2
    seed = int(rng.integers(1, 10**6))
    if is4d:
        name = str(rng.choice(["parity_4d", "parity_4d_overlap"]))
        n_prec, noise = int(rng.integers(40, 120)), 1.0
    else:
        name = str(rng.choice(["parity_small", "parity_f20", "config1"]))
        n_prec, noise = int(rng.integers(60, 250)), float(rng.choice([0.3, 1.0, 3.0]))
    ragged = bool(rng.random() < 0.4)
    rt_factor = float(rng.choice([0.5, 1.0, 2.0]))
    candidate_count = int(rng.integers(1, 6))
    fwhm_rt = float(rng.choice([2.0, 5.0, 10.0]))
    sc_kw = dict(top_k_fragments=int(rng.choice([6, 12, 20])), top_k_isotopes=int(rng.integers(2, 5)), quant_window=int(rng.integers(1, 5)),
                 quant_all=bool(rng.integers(0, 2)), experimental_xic=bool(rng.integers(0, 2)),
                 precursor_mz_tolerance=float(rng.choice([5, 15])), fragment_mz_tolerance=float(rng.choice([10, 30])))
    return dict(is4d=is4d, seed=seed, name=name, n_prec=n_prec, noise=noise, ragged=ragged, rt_factor=rt_factor,
                candidate_count=candidate_count, fwhm_rt=fwhm_rt, sc_kw=sc_kw)


def one_case(rng, it, threads, drawn=None, debug=None):
    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    sc_mod = refshim.ref("alphadia.search.scoring.scoring")
    sccfg_mod = refshim.ref("alphadia.search.scoring.config")
    d = drawn if drawn is not None else draw_case(rng, it)
    is4d, seed, name, fwhm_rt, sc_kw = d["is4d"], d["seed"], d["name"], d["fwhm_rt"], d["sc_kw"]
    if is4d:
        raw, pdf, fdf, p = make_config_4d(name, seed=seed, n_precursors=d["n_prec"])
        dia = refshim.RefDiaData4D(raw)
    else:
        raw, pdf, fdf, p = make_config_3d(name, seed=seed, n_precursors=d["n_prec"], scale_noise=d["noise"])
        dia = refshim.RefDiaData(raw)
    if d["ragged"]:
        pdf, fdf = H.ragged_library_frames(pdf, fdf, float(np.max(raw.rt_values)), seed=seed)
    sel_kw = {"rt_tolerance": float(p["rt_tolerance"]) * d["rt_factor"],
              "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)), "candidate_count": d["candidate_count"],
              "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0}
    info = {"case": it, "workload": name, "seed": seed, "n_precursors": len(pdf), "selection": sel_kw, "fwhm_rt": fwhm_rt, "scoring": sc_kw}
    # ---- live reference ----
    sel_cfg = cfg_mod.CandidateSelectionConfig()
    sel_cfg.update({**H.SELECTION_BASE, **sel_kw})
    sel = sel_mod.CandidateSelection(dia, pdf.copy(), fdf.copy(), sel_cfg, rt_column="rt_library", mobility_column="mobility_library",
                                     precursor_mz_column="mz_library", fragment_mz_column="mz_library", fwhm_rt=fwhm_rt, fwhm_mobility=0.01)
    cand = sel(thread_count=threads)
    sc_cfg = sccfg_mod.CandidateScoringConfig()
    sc_cfg.update({**H.SCORING_BASE, **sc_kw})
    feat = frag = None
    if len(cand):
        scorer = sc_mod.CandidateScoring(dia_data=dia, precursors_flat=pdf.copy(), fragments_flat=fdf.copy(), config=sc_cfg,
                                         rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                                         fragment_mz_column="mz_library")
        feat, frag = scorer(cand.copy(), thread_count=threads, include_decoy_fragment_features=True)
    # ---- oracle ----
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    osel = oracle.select_candidates_4d if is4d else oracle.select_candidates
    oscore = oracle.score_candidates_4d if is4d else oracle.score_candidates
    arrs = osel(raw, lib, H.selection_config(sel_kw["rt_tolerance"], **{k: v for k, v in sel_kw.items() if k != "rt_tolerance"}).to_struct(),
                np.asarray(sel.kernel))
    m = arrs["score"] > 0
    problems = []
    if m.sum() != len(cand):
        problems.append(f"candidate count {int(m.sum())} vs {len(cand)}")
    else:
        for c in INT_COLS:
            if not np.array_equal(arrs[c][m].astype(np.int64), cand[c].values.astype(np.int64)):
                problems.append(f"selection column {c}")
        if not np.array_equal(arrs["score"][m], cand["score"].values.astype(np.float32)):
            problems.append("selection score not bit-exact")
    info.update(candidates=int(len(cand)))
    if feat is not None and not problems:
        cin, keep = H.candidates_in_from_arrays(lib, {c: cand[c].values for c in INT_COLS})
        res = oscore(raw, lib, H.scoring_config(**sc_kw).to_struct(), cin)
        v = res["valid"].astype(bool)
        info.update(valid=int(v.sum()), fragment_rows=int(len(frag)))
        if not (np.array_equal(keep["precursor_idx"][v], feat["precursor_idx"].values) and np.array_equal(keep["rank"][v], feat["rank"].values)):
            problems.append("valid rows differ")
        else:
            F, G = res["features"][v], feat[sc_mod.DEFAULT_FEATURE_COLUMNS].values.astype(np.float32)
            worst = 0.0
            for j in range(46):
                same = (F[:, j] == G[:, j]) | (np.isnan(F[:, j]) & np.isnan(G[:, j]))
                if j in BLAS_FEATURES:
                    e = float(_excess(F[:, j], G[:, j])) if len(F) else 0.0
                    worst = max(worst, e)
                    if e > 1.0:
                        problems.append(f"feature {j}: {e:.2f} x the tolerance")
                elif not same.all():
                    problems.append(f"feature {j} not bit-exact ({int((~same).sum())} rows)")
            info["max_blas_feature_error_in_tolerances"] = worst
            mm = res["fragment_mz_library"] > 0
            if mm.sum() != len(frag):
                problems.append(f"fragment rows {int(mm.sum())} vs {len(frag)}")
            else:
                for k, v2 in FRAG_MAP.items():
                    a, b = res[v2][mm], frag[k].values
                    if k == "correlation":
                        if len(a) and float(_excess(a, b)) > 1.0:
                            problems.append("fragment correlation")
                    elif not np.array_equal(a, b):
                        problems.append(f"fragment column {k}")
    info["problems"] = problems
    if debug is not None:
        debug.update(raw=raw, pdf=pdf, fdf=fdf, lib=lib, cand=cand, feat=feat, frag=frag, res=locals().get("res"), keep=locals().get("keep"),
                     sc_mod=sc_mod, sc_kw=sc_kw, is4d=is4d)
    return info


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    threads = int(os.environ.get("ADB_THREADS", os.cpu_count() or 1))
    oracle.build()
    out, t0 = [], time.time()
    for it in range(n_cases):
        t = time.time()
        try:
            info = one_case(rng, it, threads)
        except Exception as e:  # noqa: BLE001 - a crash of either side is a finding, keep going
            info = {"case": it, "problems": [f"exception {type(e).__name__}: {e}"[:300]]}
        info["seconds"] = round(time.time() - t, 1)
        print(json.dumps(info), file=sys.stderr, flush=True)
        out.append(info)
    bad = [c for c in out if c["problems"]]
    print(json.dumps({"what": "oracle vs the live reference (numba through oracle/refshim.py) on random workloads and configurations",
                      "cases": len(out), "cases_with_problems": len(bad), "candidates_compared": sum(c.get("candidates", 0) for c in out),
                      "valid_rows_compared": sum(c.get("valid", 0) for c in out), "seconds": round(time.time() - t0, 1), "results": out}, indent=1))

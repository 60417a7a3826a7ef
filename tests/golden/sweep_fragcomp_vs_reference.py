#!/usr/bin/env python
"""Randomized pinning of the fragment competition against the LIVE reference (this container only).

Random PSM tables (sizes, window counts, collision rates, RT spans, rt / m/z dtypes, tolerances) through the unmodified
``_compete_for_fragments`` (fragcomp/fragcomp.py:51-143) and through the oracle; and random PSM / fragment DataFrames through the
reference's ``FragmentCompetition.__call__`` and through ``alphadia_b200.fragcomp`` with the oracle as its kernel.

    python tests/golden/sweep_fragcomp_vs_reference.py [n_cases] [seed] > tests/golden/sweep_fragcomp_vs_reference.json
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

import oracle  # noqa: E402
from oracle import refshim  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    fc_mod = refshim.ref("alphadia.fragcomp.fragcomp")
    refshim.ref("alphatims.utils").set_threads(8)
    oracle.build()
    from alphadia_b200 import _lib, fragcomp

    _lib.fragment_competition = lambda ws, we, rt, fs, fe, mz, rt_tol, ppm, device=None: oracle.fragment_competition(
        ws, we, rt, fs, fe, mz, rt_tol, ppm).astype(bool)
    results, t0 = [], time.time()
    for it in range(n_cases):
        dtype_rt = [np.float32, np.float64][int(rng.integers(0, 2))]
        dtype_mz = [np.float32, np.float64][int(rng.integers(0, 2))]
        n, nwin = int(rng.integers(1, 5000)), int(rng.integers(1, 9))
        rt_tol, ppm = float(rng.choice([0.5, 3.0, 10.0])), float(rng.choice([5.0, 15.0, 50.0]))
        ws, we, rt, fs, fe, mz = H.fragcomp_dense_inputs(dtype_rt, dtype_mz, seed=int(rng.integers(1, 10**6)), n=n, nwin=nwin,
                                                         rt_span=float(rng.choice([5.0, 40.0, 400.0])), p_replace=float(rng.choice([0.1, 0.45, 0.9])))
        valid = np.ones(len(rt)).astype(bool)
        fc_mod._compete_for_fragments(np.arange(len(ws)), ws, we, rt, fs, fe, mz, rt_tol, ppm, valid)
        got = oracle.fragment_competition(ws, we, rt, fs, fe, mz, rt_tol, ppm).astype(bool)
        problems = [] if np.array_equal(valid, got) else [f"kernel: {int((valid != got).sum())} of {len(valid)} PSMs differ"]
        # DataFrame level: PSMs spread over quadrupole windows by their precursor m/z, 4-12 fragments each
        n_psm = int(rng.integers(1, 1500))
        n_frag = rng.integers(4, 13, n_psm)
        base = rng.uniform(200, 1800, size=(60, 12))
        src = rng.integers(0, 60, n_psm)
        psm = pd.DataFrame({"precursor_idx": rng.permutation(n_psm * 2)[:n_psm], "rank": rng.integers(0, 3, n_psm).astype(np.uint8),
                            "rt_observed": rng.uniform(0, float(rng.choice([20.0, 300.0])), n_psm).astype(dtype_rt),
                            "proba": np.round(rng.random(n_psm), 2), "mz_observed": rng.uniform(390, 1010, n_psm).astype(np.float32)})
        rows = []
        for i in range(n_psm):
            f = base[src[i], : n_frag[i]] * (1 + rng.normal(0, 4e-6, n_frag[i]))
            rows.append(pd.DataFrame({"precursor_idx": psm["precursor_idx"].values[i], "rank": psm["rank"].values[i], "mz_observed": f.astype(dtype_mz)}))
        frag = pd.concat(rows, ignore_index=True)
        n_pos = int(rng.integers(2, 20))
        cycle = np.zeros((1, n_pos, 1, 2))
        edges = np.linspace(400, 1000, n_pos)
        cycle[0, 0, 0] = (-1, -1)
        cycle[0, 1:, 0, 0], cycle[0, 1:, 0, 1] = edges[:-1], edges[1:]
        kept_ref = fc_mod.FragmentCompetition(rt_tol_seconds=rt_tol, mass_tol_ppm=ppm, thread_count=8)(psm.copy(), frag.copy(), cycle)
        kept = fragcomp.FragmentCompetition(rt_tol_seconds=rt_tol, mass_tol_ppm=ppm)(psm.copy(), frag.copy(), cycle)
        try:
            pd.testing.assert_frame_equal(kept, kept_ref)
        except AssertionError as e:
            problems.append("dataframes: " + str(e)[:200])
        results.append({"case": it, "kernel_psms": len(rt), "kernel_kept": int(valid.sum()), "windows": nwin, "df_psms": n_psm, "df_kept": len(kept_ref),
                        "rt_dtype": np.dtype(dtype_rt).name, "mz_dtype": np.dtype(dtype_mz).name, "rt_tol": rt_tol, "ppm": ppm, "problems": problems})
        print(json.dumps(results[-1]), file=sys.stderr, flush=True)
    bad = [r for r in results if r["problems"]]
    print(json.dumps({"what": "fragment competition: oracle kernel and alphadia_b200.fragcomp (oracle as kernel) vs the live reference",
                      "cases": len(results), "cases_with_problems": len(bad), "psms_compared": sum(r["kernel_psms"] + r["df_psms"] for r in results),
                      "seconds": round(time.time() - t0, 1), "results": results}, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""How far do the candidate tables move when the reference's convolution runs on a single-precision pocketfft FFT
(fft.py:141-212, scipy.fft stands in for rocket-fft) instead of the direct fp64 circular convolution all golden vectors,
the oracle and the CUDA kernels use ("The one substituted dependency")?

Runs the UNMODIFIED reference selection through oracle/refshim.py in its ``ADB_REFSHIM_FFT=pocketfft`` mode (this container
only: it needs /root/reference) and compares with the committed golden candidate tables (direct convolution).

    python tests/golden/measure_fft_disagreement.py [threads] > tests/golden/fft_disagreement.json
"""
import json
import os
import sys
import time

os.environ.setdefault("ADB_REFSHIM_FFT", "pocketfft")  # or "separable": two 1-D fp64 passes instead of the 2-D sum
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import numpy as np  # noqa: E402

from alphadia_b200.synthetic import CONFIGS_4D, make_config_3d, make_config_4d  # noqa: E402
from oracle import refshim  # noqa: E402
from tests.helpers import SELECTION_BASE, input_checksum, load_golden  # noqa: E402

INT_COLS = ["scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]


def run(name: str, threads: int) -> dict:
    sel_mod = refshim.ref("alphadia.search.selection.selection")
    cfg_mod = refshim.ref("alphadia.search.selection.config_df")
    if name in CONFIGS_4D:
        raw, precursor_df, fragment_df, p = make_config_4d(name)
        dia = refshim.RefDiaData4D(raw)
    else:
        raw, precursor_df, fragment_df, p = make_config_3d(name)
        dia = refshim.RefDiaData(raw)
    g = load_golden(name)
    assert str(g["input_checksum"]) == input_checksum(raw, precursor_df, fragment_df), "golden was made from other inputs"
    cfg = cfg_mod.CandidateSelectionConfig()
    cfg.update({**SELECTION_BASE, "rt_tolerance": float(p["rt_tolerance"]), "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)),
                "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0})
    sel = sel_mod.CandidateSelection(dia, precursor_df.copy(), fragment_df.copy(), cfg, rt_column="rt_library",
                                     mobility_column="mobility_library", precursor_mz_column="mz_library",
                                     fragment_mz_column="mz_library", fwhm_rt=5.0, fwhm_mobility=0.01)
    t0 = time.perf_counter()
    cand = sel(thread_count=threads)
    secs = time.perf_counter() - t0
    # candidate = (precursor, peak position and limits); rank is compared separately
    def rows(pidx, cols):
        return {(int(pi), *[int(c[i]) for c in cols]) for i, pi in enumerate(pidx)}
    fft_cols = [cand[c].values for c in INT_COLS]
    dir_cols = [g["cand_" + c] for c in INT_COLS]
    a = rows(cand["precursor_idx"].values, fft_cols)
    b = rows(g["cand_precursor_idx"], dir_cols)
    a_rank = rows(cand["precursor_idx"].values, [cand["rank"].values, *fft_cols])
    b_rank = rows(g["cand_precursor_idx"], [g["cand_rank"], *dir_cols])
    top_a = rows(cand["precursor_idx"].values[cand["rank"].values == 0], [c[cand["rank"].values == 0] for c in fft_cols])
    top_b = rows(g["cand_precursor_idx"][g["cand_rank"] == 0], [c[g["cand_rank"] == 0] for c in dir_cols])
    # score differences of the candidates both runs hold
    key_a = {k: float(s) for k, s in zip(sorted(a), [0.0] * len(a))}
    sc_a = {(int(pi), *[int(c[i]) for c in fft_cols]): float(cand["score"].values[i]) for i, pi in enumerate(cand["precursor_idx"].values)}
    sc_b = {(int(pi), *[int(c[i]) for c in dir_cols]): float(g["cand_score"][i]) for i, pi in enumerate(g["cand_precursor_idx"])}
    common = sorted(a & b)
    rel = np.array([abs(sc_a[k] - sc_b[k]) / max(abs(sc_b[k]), 1e-12) for k in common]) if common else np.zeros(1)
    del key_a
    return {
        "workload": name, "reference_seconds": round(secs, 1),
        "candidates_direct": len(b), "candidates_pocketfft": len(a),
        "same_candidate_any_rank": len(a & b), "same_candidate_same_rank": len(a_rank & b_rank),
        "only_direct": len(b - a), "only_pocketfft": len(a - b),
        "rank0_direct": len(top_b), "rank0_same": len(top_a & top_b),
        "frac_same_any_rank": round(len(a & b) / max(len(b), 1), 5),
        "frac_same_same_rank": round(len(a_rank & b_rank) / max(len(b), 1), 5),
        "frac_rank0_same": round(len(top_a & top_b) / max(len(top_b), 1), 5),
        "score_rel_diff_common_median": float(np.median(rel)), "score_rel_diff_common_max": float(rel.max()),
    }


if __name__ == "__main__":
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    names = [a for a in sys.argv[2:]] or ["parity_small", "parity_4d"]
    out = [run(n, threads) for n in names]
    print(json.dumps({"mode": os.environ["ADB_REFSHIM_FFT"], "what": "reference selection with a single-precision pocketfft convolution (scipy.fft; mode separable: two 1-D fp64 passes) vs the committed "
                              "golden candidate tables (direct fp64 circular convolution); a candidate = (precursor_idx, "
                              "scan/frame centre, start, stop)", "results": out}, indent=1))

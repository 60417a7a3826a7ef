"""Time the HOST side of the operator classes with the device call replaced by a stand-in that only
allocates the output arrays, so that the pandas / numpy glue around ``adb_score_candidates`` can be profiled on a box
without a GPU.  The numbers say nothing about the kernels.

    python profiles/profile_host_glue.py [n_precursors] [--valid-frac=0.6] [--cprofile]
"""

from __future__ import annotations

import cProfile
import pstats
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

from alphadia_b200 import _abi, _lib, scoring  # noqa: E402
from alphadia_b200.synthetic import make_library  # noqa: E402


class _Raw:
    """Only what CandidateScoring touches before the device call."""

    def __init__(self):
        self.cycle = np.zeros((1, 76, 1, 2))
        self.has_mobility = False


def main():
    n_prec = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 500_000
    valid_frac = float(next((a.split("=")[1] for a in sys.argv if a.startswith("--valid-frac=")), 1.0))
    rng = np.random.default_rng(0)
    pdf, fdf = make_library(n_prec, rng, quad_lo=400, quad_hi=1000, rt_lo=60, rt_hi=1100)
    n_cand = 3 * n_prec
    pidx = np.repeat(np.arange(n_prec, dtype=np.uint32), 3)
    cand = pd.DataFrame({
        "precursor_idx": pidx, "rank": np.tile(np.arange(3, dtype=np.uint8), n_prec),
        "score": rng.random(n_cand).astype(np.float32),
        "scan_center": np.zeros(n_cand, np.int64), "scan_start": np.zeros(n_cand, np.int64), "scan_stop": np.full(n_cand, 2, np.int64),
        "frame_center": np.full(n_cand, 7600, np.int64), "frame_start": np.full(n_cand, 7000, np.int64),
        "frame_stop": np.full(n_cand, 8140, np.int64),
        "elution_group_idx": pidx.copy(), "decoy": (pidx % 2).astype(np.uint8),
    })

    def fake_score(dev_raw, dev_lib, cfg, cin):
        _, arrs = _abi.alloc_scores_out(int(cin.n), int(cfg.top_k_fragments))
        valid = np.random.default_rng(1).random(int(cin.n)) < valid_frac
        arrs["valid"][:] = valid
        arrs["features"][:] = 1.0
        arrs["fragment_mz_library"][valid, :12] = 500.0
        return arrs

    class FakeLib:
        def __init__(self, arrays, device=0):
            pass

        def close(self):
            pass

    class FakeDevRaw:
        device = 0

        def last_timing(self):
            return {}

    def fake_score_ragged(dev_raw, dev_lib, cfg, cin, bufs=None, max_fragments=None):
        """Stand-in for adb_score_candidates_ragged: valid rows with `valid_frac`, 4 kept fragment slots each."""
        n = int(cin.n)
        valid = np.flatnonzero(np.random.default_rng(1).random(n) < valid_frac).astype(np.int64)
        nr, per = len(valid), 4
        out = dict(n_rows=nr, n_fragments=nr * per, row_index=valid, features=np.ones((nr, _abi.NUM_FEATURES), np.float32),
                   frag_offset=np.arange(nr + 1, dtype=np.int64) * per)
        for k in _abi.FRAG_F32:
            out[k] = np.full(nr * per, 500.0, np.float32)
        for k in _abi.FRAG_U8:
            out[k] = np.ones(nr * per, np.uint8)
        return out

    _lib.score_candidates = fake_score
    _lib.score_candidates_ragged = fake_score_ragged
    _lib.DeviceLibrary = FakeLib
    _lib.device_rawfile_for = lambda dia, raw: FakeDevRaw()
    scoring.adapt_dia_data = lambda d: d

    t0 = time.perf_counter()
    op = scoring.CandidateScoring(dia_data=_Raw(), precursors_flat=pdf, fragments_flat=fdf, rt_column="rt_library",
                                  mobility_column="mobility_library", precursor_mz_column="mz_library",
                                  fragment_mz_column="mz_library")
    t1 = time.perf_counter()
    print(f"constructor {t1 - t0:.3f} s")
    for it in range(2):
        t0 = time.perf_counter()
        if "--cprofile" in sys.argv and it == 1:
            pr = cProfile.Profile()
            pr.enable()
        feat, frag = op(cand)
        if "--cprofile" in sys.argv and it == 1:
            pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
        t1 = time.perf_counter()
        print(f"call {it}: {t1 - t0:.3f} s for {n_cand} candidates -> {n_cand / (t1 - t0) / 1e6:.2f} M cand/s (host glue only), "
              f"features {feat.shape}, fragments {frag.shape}")


if __name__ == "__main__":
    main()

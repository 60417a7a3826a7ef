#!/usr/bin/env python
"""Timing of the FDR bookkeeping at config-3 scale (6 M candidate rows): adb_q_values / adb_keep_best with host
buffers in and out (copies inside the timed call) next to the reference's pandas formulation (sort_values + cumsum +
minimum.accumulate, sort_values + groupby.head; alphadia/fdr/fdr.py:195-297) and the C restatement on one host core.

    python profiles/bench_fdr.py [n_rows]
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from alphadia_b200 import _lib  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 6_000_000
rng = np.random.default_rng(0)
score = rng.random(n).astype(np.float32).astype(np.float64)  # classifier output: float32 probabilities
decoy = (rng.random(n) < 0.5).astype(np.uint8)
precursor_idx = rng.integers(0, n // 3, n).astype(np.uint64)
oracle.build()
_lib.require_device()
_lib.q_values(score[:1000], decoy[:1000], precursor_idx[:1000], device=0)  # context + module load


def best_of(f, reps=3):
    times, out = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        times.append(time.perf_counter() - t0)
    return min(times), out


t_q, (order, qval) = best_of(lambda: _lib.q_values(score, decoy, precursor_idx, device=0))
t_k, keep = best_of(lambda: _lib.keep_best(score, precursor_idx, device=0))
t0 = time.perf_counter()
o_ref, q_ref = oracle.q_values(score, decoy, precursor_idx)
k_ref = oracle.keep_best(score, precursor_idx)
t_port = time.perf_counter() - t0
df = pd.DataFrame({"proba": score, "_decoy": decoy, "precursor_idx": precursor_idx})
t0 = time.perf_counter()
s = df.sort_values(["proba", "_decoy", "precursor_idx"], ascending=True)
d = s["_decoy"].to_numpy()
with np.errstate(all="ignore"):
    f = np.cumsum(d) / np.cumsum(1 - d)
q_pd = np.flip(np.minimum.accumulate(np.flip(f)))
b = df.reset_index(drop=True).sort_values(["proba", "precursor_idx"], ascending=True).groupby(["precursor_idx"]).head(1).sort_index()
t_pandas = time.perf_counter() - t0
ok = (np.array_equal(order, o_ref) and np.array_equal(qval, q_ref) and np.array_equal(keep, k_ref)
      and np.array_equal(s.index.values, order) and np.array_equal(q_pd, qval) and np.array_equal(b.index.values, np.flatnonzero(keep)))
print(json.dumps({"op": "adb_q_values + adb_keep_best", "n_rows": n, "identical_to_oracle_and_pandas": bool(ok),
                  "gpu_q_values_s": t_q, "gpu_keep_best_s": t_k, "gpu_rows_per_s": n / (t_q + t_k),
                  "pandas_formulation_s": t_pandas, "cpu_port_s": t_port, "cpu_cores": 1,
                  "bytes_in": int(2 * score.nbytes + decoy.nbytes + 2 * precursor_idx.nbytes),
                  "bytes_out": int(order.nbytes + qval.nbytes + keep.nbytes),
                  "note": "GPU calls = pageable H2D + two stable 64-bit radix sorts + scans + D2H each"}))

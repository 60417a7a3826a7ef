#!/usr/bin/env python
"""Timing of the load-time CSR transpose at config-4 scale: adb_transpose_csr (host buffers in and out, so the
H2D/D2H copies are inside the timed call) next to the C restatement of the reference's _transpose on one host core.

    python profiles/bench_transpose.py [n_events] [n_push] [n_tof]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from alphadia_b200 import _lib  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000_000
n_push = int(float(sys.argv[2])) if len(sys.argv) > 2 else 7201 * 928
n_tof = int(float(sys.argv[3])) if len(sys.argv) > 3 else 634_744
rng = np.random.default_rng(0)
counts = np.bincount(rng.integers(0, n_push, n), minlength=n_push)
push_indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
tof = rng.integers(0, n_tof, n, dtype=np.uint32)
values = rng.integers(1, 60000, n).astype(np.uint16)
oracle.build()
_lib.require_device()
_lib.transpose_csr(tof[:1000], np.array([0, 1000], np.int64), n_tof, values[:1000], device=0)  # context + module load
times = []
for _ in range(3):
    t0 = time.perf_counter()
    got = _lib.transpose_csr(tof, push_indptr, n_tof, values, device=0)
    times.append(time.perf_counter() - t0)
t0 = time.perf_counter()
ref = oracle.transpose_csr(tof, push_indptr, n_tof, values)
t_cpu = time.perf_counter() - t0
ok = all(np.array_equal(a, b) for a, b in zip(got, ref))
print(json.dumps({"op": "adb_transpose_csr", "n_events": n, "n_push": n_push, "n_tof": n_tof, "identical_to_oracle": bool(ok),
                  "gpu_call_s": min(times), "gpu_events_per_s": n / min(times), "cpu_port_s": t_cpu, "cpu_cores": 1,
                  "bytes_in": int(tof.nbytes + push_indptr.nbytes + values.nbytes),
                  "bytes_out": int(sum(a.nbytes for a in got)),
                  "note": "GPU call = pageable H2D + stable radix sort + gathers + D2H; CPU = counting-sort restatement of bruker.py:202-274"}))

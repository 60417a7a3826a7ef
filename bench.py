#!/usr/bin/env python
"""bench.py — precursor candidates scored per second (BASELINE.json metric) on synthetic DIA data.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # B200 arm (default workload: config3)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # CPU arm: the oracle port on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, weak scaling

One "step" = candidate selection + candidate scoring of the whole library batch against one raw file
(SURVEY.md §8d).  `value` is measured with raw file and library resident in HBM (results stay in HBM);
`e2e` is measured through the C ABI with pinned HOST buffers: library batch H2D, then ONE call
adb_select_score_candidates_ragged (selection, candidate table D2H, scoring, ragged score + fragment tables D2H) every
step (`--e2e-two-calls` / `--e2e-dense`: the separate-call and dense-table variants); the raw file is uploaded once per
file, as `dia_data.to_jitclass()` is built once per file in the reference.  `e2e_operator` is the DataFrame-level operator
path, `fragcomp` the fragment competition host to host, `parity` an oracle spot check of the timed results, and
`other_workloads` (default N = 1 run only) condensed lines of BASELINE.json's config 2 and config 4, each from a short child run
of this script after the main measurement (`--no-other-workloads` skips them).

Baselines: `cpu_baseline` and `--impl reference` run the C/OpenMP PORT of the reference algorithm (oracle/adb_oracle.c,
`kind: "port"`), not the numba code (it cannot travel to the GPU box; BASELINE.md has the numba calibration).  Everywhere
in this repository "the reference" means alphaDIA with its float32 rocket-fft convolution replaced by the direct fp64
circular convolution (rocket-fft is not installed) - on the oracle, the golden vectors and the device alike.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "precursor candidates scored/sec"
UNIT = "candidates/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def build_workload(name: str, rank: int, n_precursors: int | None):
    from alphadia_b200.config import CandidateScoringConfig, CandidateSelectionConfig
    from alphadia_b200.kernel import GaussianKernel
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.synthetic import CONFIGS_3D, CONFIGS_4D, make_config_3d, make_config_4d

    is4d = name in CONFIGS_4D
    seed = (CONFIGS_4D if is4d else CONFIGS_3D)[name]["seed"] + rank
    t0 = time.time()
    make = make_config_4d if is4d else make_config_3d
    raw, pdf, fdf, p = make(name, seed=seed, n_precursors=n_precursors, with_strings=False)
    lib = assemble_library_arrays(pdf, fdf, "rt_library", "mobility_library", "mz_library", "mz_library")
    # ClassicExtractionHandler parameters (reference extraction_handler.py:349-409) at target tolerances
    sel = CandidateSelectionConfig()
    sel.update({
        "peak_len_rt": 10.0, "sigma_scale_rt": 0.5, "peak_len_mobility": 0.01, "sigma_scale_mobility": 1.0,
        "top_k_precursors": 3, "kernel_size": 30, "f_mobility": 1.0, "f_rt": 0.99, "center_fraction": 0.5,
        "min_size_mobility": 8, "min_size_rt": 3, "max_size_mobility": 20, "max_size_rt": 15,
        "group_channels": False, "use_weighted_score": True, "join_close_candidates": False,
        "join_close_candidates_scan_threshold": 0.6, "join_close_candidates_cycle_threshold": 0.6,
        "top_k_fragments": 12, "exclude_shared_ions": True, "rt_tolerance": float(p["rt_tolerance"]),
        "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)), "candidate_count": 3,
        "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0,
    })
    sc = CandidateScoringConfig()
    sc.update({
        "score_grouped": False, "top_k_isotopes": 3, "reference_channel": -1, "precursor_mz_tolerance": 5,
        "fragment_mz_tolerance": 10, "exclude_shared_ions": True, "quant_window": 3, "quant_all": True,
        "experimental_xic": True, "top_k_fragments": 12,
    })
    kernel = GaussianKernel(raw, fwhm_rt=5.0, sigma_scale_rt=0.5, fwhm_mobility=0.01, sigma_scale_mobility=1.0,
                            kernel_width=30, kernel_height=min(30, raw.scan_max_index + 1)).get_dense_matrix(verbose=False)
    if is4d:
        log(f"[rank {rank}] workload {name}: {len(pdf)} precursors, {len(fdf)} fragments, {len(raw.rt_values)} frames x "
            f"{raw.scan_max_index} scans, {raw.n_events} events ({(raw.n_events * 6) / 1e9:.2f} GB) generated in {time.time() - t0:.1f}s")
    else:
        log(f"[rank {rank}] workload {name}: {len(pdf)} precursors, {len(fdf)} fragments, {len(raw.rt_values)} spectra, "
            f"{raw.n_peaks} peaks ({(raw.n_peaks * 8) / 1e9:.2f} GB) generated in {time.time() - t0:.1f}s")
    return raw, pdf, fdf, lib, p, sel, sc, kernel


def algorithmic_bytes(raw, lib, c_sel_mean, c_sc_mean, n_obs=1.0, F=12, I=3, N=3, h=1.0):
    """SURVEY.md §8d formulas (bytes the algorithm must touch per precursor / per candidate)."""
    L = raw.cycle_len
    counts = (raw.peak_stop_idx_list - raw.peak_start_idx_list).astype(np.float64)
    is_ms1 = (np.arange(len(counts)) % L) == 0
    p_ms1 = float(counts[is_ms1].mean())
    p_ms2 = float(counts[~is_ms1].mean())
    probe2 = 4 * math.ceil(math.log2(max(p_ms2, 2)))
    probe1 = 4 * math.ceil(math.log2(max(p_ms1, 2)))
    n_iso = lib["isotopes"].shape[1]
    b_prec = (c_sel_mean * n_obs * (16 + F * (probe2 + 8 * h)) + c_sel_mean * (16 + I * (probe1 + 8 * h))
              + (18 * F + 40 + 4 * n_iso) + 33 * N)
    b_cand = (c_sc_mean * n_obs * (16 + F * (probe2 + 8 * h)) + c_sc_mean * (16 + I * (probe1 + 8 * h))
              + (18 * F + 40) + 72 + (184 + 38 * F + 8))
    return dict(b_prec=b_prec, b_cand=b_cand, p_ms1=p_ms1, p_ms2=p_ms2)


def algorithmic_bytes_4d(raw, lib, p, frames_sel, frames_sc, F=12, I=3, N=3):
    """SURVEY.md §8d, 4-D form: per query (fragment / isotope) T tof rows; per row the CSR pointer pair (16 B), two
    binary searches on the push axis (4 B probes) and the (push u32 + intensity u16) of the events inside the frame window."""
    R = raw.n_events / max(len(raw.mz_values), 1)
    probes = 2 * 4 * math.ceil(math.log2(max(R, 2)))
    t2 = 2 * 10.0 / p["tof_ppm"]
    t1 = 2 * 5.0 / p["tof_ppm"]
    n_frames = len(raw.rt_values)
    n_iso = lib["isotopes"].shape[1]

    def per_window(frames):
        h = R * frames / n_frames
        return F * t2 * (16 + probes + 6 * h) + I * t1 * (16 + probes + 6 * h)

    b_prec = per_window(frames_sel) + (18 * F + 40 + 4 * n_iso) + 33 * N
    b_cand = per_window(frames_sc) + (18 * F + 40) + 72 + (184 + 38 * F + 8)
    return dict(b_prec=b_prec, b_cand=b_cand, row_len=R, tof_rows_ms2=t2, tof_rows_ms1=t1)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def pinned_alloc_factory():
    import torch

    keep = []

    def alloc(shape, dtype):
        n = int(np.prod(shape))
        t = torch.empty(max(n, 1) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
        keep.append(t)
        return t.numpy()[: n * np.dtype(dtype).itemsize].view(dtype).reshape(shape)

    alloc.keep = keep
    return alloc


# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(raw, lib, sel, sc, kernel, rows, threads):
    """Oracle port (CPU) on library rows [rows]: selection + scoring; returns (#candidates, seconds)."""
    import oracle
    from alphadia_b200 import _abi

    sub = dict(lib)
    for k in ("precursor_idx", "frag_start_idx", "frag_stop_idx", "charge", "rt", "mobility", "mz", "isotopes"):
        sub[k] = np.ascontiguousarray(lib[k][rows])
    is4d = hasattr(raw, "tof_indptr")
    o_select = oracle.select_candidates_4d if is4d else oracle.select_candidates
    o_score = oracle.score_candidates_4d if is4d else oracle.score_candidates
    t0 = time.perf_counter()
    cont = o_select(raw, sub, sel.to_struct(), kernel, n_threads=threads)
    t_sel = time.perf_counter() - t0
    m = np.flatnonzero(cont["score"] > 0)
    cc = int(sel.candidate_count)
    cin, keep = _abi.make_candidates_in(m // cc, cont["rank"][m], cont["scan_start"][m], cont["scan_stop"][m],
                                        cont["scan_center"][m], cont["frame_start"][m], cont["frame_stop"][m],
                                        cont["frame_center"][m])
    t0 = time.perf_counter()
    out = o_score(raw, sub, sc.to_struct(), cin, n_threads=threads)
    t_sc = time.perf_counter() - t0
    return len(m), t_sel, t_sc, int(out["valid"].sum())


def parity_spot_check(hp, raw, lib, sel, sc, kernel, cont_gpu, n_prec=800, seed=7):
    """Outside the timed region: the oracle selects and scores a random subsample of the library (>= 2000 candidate rows) on
    the host and the result is compared with what the timed device path left in HBM for the same precursors — candidate
    container rows bit for bit, the 46 features within 1e-4 relative, the per-fragment tables."""
    import oracle
    from alphadia_b200 import _abi, _lib

    oracle.build()
    is4d = hasattr(raw, "tof_indptr")
    P, cc = int(hp.n_precursors), int(sel.candidate_count)
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.permutation(P)[: min(P, n_prec)])
    sub = dict(lib)
    for k in ("precursor_idx", "frag_start_idx", "frag_stop_idx", "charge", "rt", "mobility", "mz", "isotopes"):
        sub[k] = np.ascontiguousarray(lib[k][rows])
    threads = os.cpu_count() or 1
    ref = (oracle.select_candidates_4d if is4d else oracle.select_candidates)(raw, sub, sel.to_struct(), kernel, n_threads=threads)
    take = (rows[:, None] * cc + np.arange(cc)[None, :]).ravel()
    int_cols = ("precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop")
    int_exact = all(np.array_equal(cont_gpu[c][take], ref[c]) for c in int_cols)
    score_exact = bool(np.array_equal(cont_gpu["score"][take].view(np.uint32), ref["score"].view(np.uint32)))
    m = np.flatnonzero(ref["score"] > 0)
    cin, keep = _abi.make_candidates_in(m // cc, ref["rank"][m], ref["scan_start"][m], ref["scan_stop"][m], ref["scan_center"][m],
                                        ref["frame_start"][m], ref["frame_stop"][m], ref["frame_center"][m])
    ref_sc = (oracle.score_candidates_4d if is4d else oracle.score_candidates)(raw, sub, sc.to_struct(), cin, n_threads=threads)
    got = hp.fetch()  # every row the timed step scored, with its (library row, rank)
    key_gpu = got["lib_row"].astype(np.int64) * 256 + got["rank"].astype(np.int64)
    order = np.argsort(key_gpu, kind="stable")
    key_ref = rows[m // cc].astype(np.int64) * 256 + ref["rank"][m].astype(np.int64)
    at = np.searchsorted(key_gpu[order], key_ref)
    found = (at < len(order)) & (key_gpu[order][np.minimum(at, len(order) - 1)] == key_ref)
    sel_rows = order[np.minimum(at, len(order) - 1)]
    valid_exact = bool(found.all() and np.array_equal(got["valid"][sel_rows], ref_sc["valid"]))
    v = ref_sc["valid"].astype(bool) & found
    Fa, Fb = got["features"][sel_rows][v].astype(np.float64), ref_sc["features"][v].astype(np.float64)
    s_med = np.nanmedian(np.abs(Fb), axis=0) if len(Fb) else np.zeros(Fb.shape[1])
    floor = np.maximum(1e-6, 1e-6 * np.where(np.isfinite(s_med), s_med, 0.0))
    both_nan = np.isnan(Fa) & np.isnan(Fb)
    err = np.where(both_nan, 0.0, np.abs(Fa - Fb) / np.maximum(np.maximum(np.abs(Fa), np.abs(Fb)), floor[None, :]))
    err = np.where(np.isnan(err), np.inf, err)
    frag_u8 = all(np.array_equal(got[k][sel_rows][v], ref_sc[k][v]) for k in _abi.FRAG_U8)
    frag_rel = 0.0
    for k in _abi.FRAG_F32:
        a, b = got[k][sel_rows][v].astype(np.float64), ref_sc[k][v].astype(np.float64)
        d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-6)
        frag_rel = max(frag_rel, float(d.max()) if d.size else 0.0)
    return {"n": int(len(m)), "precursors": int(len(rows)), "valid": int(v.sum()), "int_exact": bool(int_exact and score_exact and frag_u8),
            "selection_score_bit_exact": score_exact, "valid_exact": valid_exact, "max_rel": float(err.max()) if err.size else 0.0,
            "fragment_table_max_rel": frag_rel, "features_bit_identical_frac": float((Fa == Fb).mean()) if Fa.size else 1.0,
            "tolerance": 1e-4,
            "checked": "candidate container rows (8 integer columns + f32 score, bit for bit), valid mask, 46 features, 12 per-fragment "
                       "columns of a random library subsample, device results of the timed step vs oracle/adb_oracle.c"}


def cpu_baseline(raw, lib, sel, sc, kernel, target_seconds=15.0):
    import oracle

    oracle.build()
    threads = os.cpu_count() or 1
    P = len(lib["precursor_idx"])
    rng = np.random.default_rng(99)
    perm = np.sort(rng.permutation(P)[: min(P, 2000)])
    n, t_sel, t_sc, _ = cpu_reference_pass(raw, lib, sel, sc, kernel, perm, threads)  # pilot (also warms the threads)
    rate = len(perm) / max(t_sel + t_sc, 1e-6)
    size = int(min(P, max(2000, rate * target_seconds)))
    rows = np.sort(rng.permutation(P)[:size])
    n, t_sel, t_sc, valid = cpu_reference_pass(raw, lib, sel, sc, kernel, rows, threads)
    return dict(value=n / (t_sel + t_sc), unit=UNIT, cores=threads, kind="port",
                sample=f"{size} of {P} precursors (random subset, same raw file), selection {t_sel:.2f}s + scoring {t_sc:.2f}s, "
                       f"{n} candidates, {valid} valid; oracle/adb_oracle.c with OpenMP",
                precursors_per_s=size / t_sel, scoring_candidates_per_s=n / t_sc)


def bench_operator(raw, pdf, fdf, sel_cfg, sc_cfg, steps=2):
    """The reference-facing operator path (what ClassicExtractionHandler runs, extraction_handler.py:411-486): DataFrames in,
    DataFrames out - CandidateSelection(...)() -> CandidateScoring(...)(candidates_df), constructors included, on the same
    workload.  Host pandas / numpy glue, schema validation, library marshalling and every copy are inside the timed region."""
    from alphadia_b200 import CandidateScoring, CandidateSelection

    pdf = pdf.copy()
    if "sequence" not in pdf.columns:  # string columns of a real library (a pool of distinct peptides; cheap to generate)
        rng = np.random.default_rng(5)
        aa = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
        pool = np.array(["".join(r) for r in aa[rng.integers(0, 20, size=(4096, 9))]], dtype=object)
        pdf["sequence"] = pool[rng.integers(0, 4096, size=len(pdf))]
        for col in ("mods", "mod_sites"):
            pdf[col] = np.full(len(pdf), "", dtype=object)
    cols = dict(rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
                fragment_mz_column="mz_library")
    phases = []
    n_cand = n_feat = n_frag = 0
    for step in range(steps + 1):  # step 0 is the warm-up (second device copy of the raw file, first allocations)
        t0 = time.perf_counter()
        selector = CandidateSelection(raw, pdf, fdf, sel_cfg, fwhm_rt=5.0, fwhm_mobility=0.01, **cols)
        t1 = time.perf_counter()
        cand_df = selector(thread_count=1)
        t2 = time.perf_counter()
        scorer = CandidateScoring(dia_data=raw, precursors_flat=pdf, fragments_flat=fdf, config=sc_cfg, **cols)
        t3 = time.perf_counter()
        feat_df, frag_df = scorer(cand_df, thread_count=1)
        t4 = time.perf_counter()
        n_cand, n_feat, n_frag = len(cand_df), len(feat_df), len(frag_df)
        if step:
            phases.append({"selection_ctor": 1e3 * (t1 - t0), "selection_call": 1e3 * (t2 - t1), "scoring_ctor": 1e3 * (t3 - t2),
                           "scoring_call": 1e3 * (t4 - t3), "total": 1e3 * (t4 - t0)})
        del selector, scorer, cand_df, feat_df, frag_df
    total_s = sum(p["total"] for p in phases) / 1e3 / len(phases)
    return {"value": n_cand / total_s, "unit": UNIT, "steps": steps, "candidates": n_cand, "feature_rows": n_feat,
            "fragment_rows": n_frag, "ms_per_step": 1e3 * total_s,
            "phases_ms": {k: float(np.mean([p[k] for p in phases])) for k in phases[0]},
            "path": "CandidateSelection(...)(): candidates DataFrame -> CandidateScoring(...)(candidates_df): feature + fragment DataFrames"}


def fragcomp_workload(n_psm=200_000, n_frag=12, n_windows=75, seed=11):
    """SURVEY.md's fragment-competition probe shape: 200 000 PSMs x 12 fragments in 75 DIA windows, sorted by
    (window, proba) as FragmentCompetition.plan leaves them (fragcomp.py:254-273).  A tenth of the PSMs are shadows of a better
    PSM of the same window: same retention time (+- 2 s) and 4-8 of its fragment m/z (+- 5 ppm), so the veto has work to do."""
    rng = np.random.default_rng(seed)
    window = np.sort(rng.integers(0, n_windows, n_psm))
    proba = rng.random(n_psm)
    order = np.lexsort((proba, window))
    window, proba = window[order], proba[order]
    rt = rng.uniform(0.0, 2400.0, n_psm).astype(np.float32)
    mz = np.sort(rng.uniform(200.0, 1800.0, (n_psm, n_frag)), axis=1).astype(np.float32)
    ws = np.searchsorted(window, np.arange(n_windows), side="left").astype(np.int64)
    we = np.searchsorted(window, np.arange(n_windows), side="right").astype(np.int64)
    shadows = rng.choice(n_psm, n_psm // 10, replace=False)
    for i in shadows:
        w = window[i]
        if i <= ws[w]:
            continue
        j = rng.integers(ws[w], i)  # a better PSM (lower proba) of the same window
        rt[i] = rt[j] + np.float32(rng.uniform(-2.0, 2.0))
        k = rng.integers(4, 9)
        cols = rng.choice(n_frag, k, replace=False)
        mz[i, cols] = mz[j, cols] * (1.0 + rng.uniform(-5e-6, 5e-6, k)).astype(np.float32)
    fs = (np.arange(n_psm, dtype=np.int64) * n_frag)
    return dict(window_start=ws, window_stop=we, rt=rt, frag_start=fs, frag_stop=fs + n_frag, fragment_mz=mz.reshape(-1))


def bench_fragcomp(repeats=5):
    """Fragment competition host-to-host through adb_fragment_competition (H2D + kernels + D2H) next to the C port on all
    host threads, same arrays; the surviving rows must be identical."""
    from alphadia_b200 import _lib

    w = fragcomp_workload()
    args = (w["window_start"], w["window_stop"], w["rt"], w["frag_start"], w["frag_stop"], w["fragment_mz"], 3.0, 15.0)
    valid = _lib.fragment_competition(*args)  # warm-up
    t0 = time.perf_counter()
    for _ in range(repeats):
        valid = _lib.fragment_competition(*args)
    t_gpu = (time.perf_counter() - t0) / repeats
    os.environ["ADB_FRAGCOMP_SERIAL"] = "1"
    try:
        t0 = time.perf_counter()
        valid_serial = _lib.fragment_competition(*args)
        t_serial = time.perf_counter() - t0
    finally:
        del os.environ["ADB_FRAGCOMP_SERIAL"]
    import oracle

    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    ref = oracle.fragment_competition(*args, n_threads=threads).astype(bool)
    t_cpu = time.perf_counter() - t0
    n = len(w["rt"])
    return {"n_psm": n, "fragments_per_psm": 12, "windows": 75, "removed": int(n - valid.sum()),
            "psm_per_s": n / t_gpu, "ms": 1e3 * t_gpu, "window_serial_kernel_ms": 1e3 * t_serial,
            "cpu_port_psm_per_s": n / t_cpu, "cpu_cores": threads, "identical_to_cpu_port": bool(np.array_equal(valid, ref)),
            "identical_to_window_serial_kernel": bool(np.array_equal(valid, valid_serial)),
            "path": "adb_fragment_competition, host arrays in and out (conflict graph over RT-sorted windows + greedy pass)"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    raw, pdf, fdf, lib, p, sel, sc, kernel = build_workload(args.workload, 0, args.precursors)
    import oracle

    oracle.build()
    threads = os.cpu_count() or 1
    P = len(lib["precursor_idx"])
    rng = np.random.default_rng(99)
    pilot = np.sort(rng.permutation(P)[: min(P, 2000)])
    n, t_sel, t_sc, _ = cpu_reference_pass(raw, lib, sel, sc, kernel, pilot, threads)
    rate = len(pilot) / max(t_sel + t_sc, 1e-6)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    size = int(min(P, max(2000, rate * min(budget, 20.0))))
    times, cands = [], []
    for s in range(args.warmup + args.steps):
        rows = np.sort(rng.permutation(P)[:size])
        n, t_sel, t_sc, valid = cpu_reference_pass(raw, lib, sel, sc, kernel, rows, threads)
        if s >= args.warmup:
            times.append(t_sel + t_sc); cands.append(n)
    value = sum(cands) / sum(times)
    sample = f"{size} of {P} precursors per step (random subset), all {threads} host threads (OpenMP), oracle port of the numba path"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": workload_description(args.workload, P), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def other_workload_lines(names, timeout_s=300):
    """BASELINE.json's other single-GPU configurations, each measured by a short run of this script in a child process after
    the main measurement (same code path: resident value, C-ABI e2e, oracle spot check of the timed results); the child's
    JSON line is condensed to the figures below.  A failure is recorded, it never breaks the main line."""
    import subprocess

    out = {}
    for name in names:
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", name, "--steps", "3", "--warmup", "3", "--e2e-steps", "3",
               "--no-cpu-baseline", "--no-operator", "--no-fragcomp", "--no-other-workloads"]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            par = d.get("parity") or {}
            out[name] = {
                "workload": d["config"]["workload"], "value": d["value"], "unit": d["unit"], "e2e": d["e2e"]["value"],
                "ms_per_step": d["ms_per_step"], "stage_ms": d["config"]["stage_ms"], "candidates_per_step": d["config"]["candidates_per_step"],
                "steps": d["steps"], "warmup": d["warmup"], "gpu_launches": d.get("gpu_launches"),
                "roofline_frac": d["roofline"]["frac"], "roofline_kernel": d["roofline"]["kernel"],
                "parity": {k: par.get(k) for k in ("n", "int_exact", "selection_score_bit_exact", "valid_exact", "max_rel", "tolerance")},
                "wall_s": round(time.time() - t0, 1),
            }
        except Exception as e:  # noqa: BLE001 - reported, not raised
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300], "wall_s": round(time.time() - t0, 1)}
        log(f"other workload {name}: {json.dumps(out[name])}")
    return out


def workload_description(name, P):
    from alphadia_b200.synthetic import CONFIGS_3D, CONFIGS_4D

    if name in CONFIGS_4D:
        c = CONFIGS_4D[name]
        return (f"{name}: {P} precursors x 12 fragments, 3 candidates/precursor, synthetic timsTOF-shape 4-D run "
                f"{c['n_cycles']} cycles x (1 MS1 + {c['n_ms2_frames']} diaPASEF frames) x {c['n_scans']} scans, "
                f"rt_tolerance {c['rt_tolerance']}s, mobility_tolerance {c['mobility_tolerance']}, ms1/ms2 5/10 ppm")
    c = CONFIGS_3D[name]
    return (f"{name}: {P} precursors x 12 fragments, 3 candidates/precursor, synthetic Thermo-shape 3-D run "
            f"{c['n_cycles']} cycles x (1 MS1 + {c['n_windows']} MS2), rt_tolerance {c['rt_tolerance']}s, ms1/ms2 5/10 ppm")


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ADB_BENCH_WORKLOAD", "config3"))
    ap.add_argument("--precursors", type=int, default=None, help="override the library size (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-two-calls", action="store_true", help="e2e with separate selection and scoring calls: the candidate table goes to the host and is uploaded again (A/B)")
    ap.add_argument("--e2e-dense", action="store_true", help="e2e with the dense [n, top_k] result tables of adb_score_candidates (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-operator", action="store_true", help="skip the DataFrame-level operator measurement")
    ap.add_argument("--operator-steps", type=int, default=2)
    ap.add_argument("--no-fragcomp", action="store_true", help="skip the fragment-competition side benchmark")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle spot check of the timed results")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="skip the short runs of BASELINE.json's other single-GPU configurations (config 2, config 4) that the default N = 1 run adds to its line")
    ap.add_argument("--parity-precursors", type=int, default=800)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from alphadia_b200 import _lib
    from alphadia_b200.engine import HotPath
    from alphadia_b200.sharding import ScoreTableGather

    _lib.require_device()  # no CPU fallback
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    raw, pdf, fdf, lib, p, sel, sc, kernel = build_workload(args.workload, rank, args.precursors)
    t0 = time.time()
    hp = HotPath(raw, lib, sel, sc, kernel, device=local_rank)
    log(f"[rank {rank}] raw file + library resident in HBM ({hp.dev_raw.device_bytes / 1e9:.2f} GB raw) in {time.time() - t0:.1f}s")
    lib_pidx_dev = torch.from_numpy(lib["precursor_idx"].astype(np.int64)).cuda(local_rank) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gather = None
    if world > 1:  # preallocated buffers of the single collective: one all-gather of the packed score table
        gather = ScoreTableGather(int(hp.n_precursors * sel.candidate_count), torch.device("cuda", local_rank))

    def one_step():
        r = hp.resident_step()
        if gather is not None:
            gather.pack_resident(hp, lib_pidx_dev)
            gather.allgather()
        return r

    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = hp.dev_raw.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stats = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        stats.append(one_step())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ev1.record(); torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = hp.dev_raw.kernel_launches - launches0
    # device time of the step = CUDA-event time on the engine's stream (selection + compaction + scoring)
    dev_ms = sum(s["select_ms"] + s["score_ms"] for s in stats)
    n_cand = stats[-1]["n_candidates"]
    # the calls block; wall time additionally contains the host-side launch/sync gaps (and the all-gather)
    t_rank = torch.tensor([t_wall, dev_ms / 1000.0, float(n_cand)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t_rank.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_rank.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_wall_max, dev_s_max, n_cand_total = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        t_wall_max, dev_s_max, n_cand_total = t_wall, dev_ms / 1000.0, float(n_cand)
    value = n_cand_total * args.steps / t_wall_max

    # ---- N > 1: the gathered table on every rank must hold every rank's local table (outside the timed region) ----
    gather_check = None
    if gather is not None:
        g_all, sizes = gather.allgather()
        n_loc = int(gather.n_local.item())
        local_sum = gather.local[:n_loc].to(torch.int64).sum().reshape(1)
        sums = torch.zeros(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(sums, local_sum)
        sizes_h = sizes.cpu().tolist()
        ok = all(int(g_all[r, : int(sizes_h[r])].to(torch.int64).sum().item()) == int(sums[r].item()) for r in range(world))
        ok = ok and bool(torch.equal(g_all[rank, :n_loc], gather.local[:n_loc]))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = {"ok": bool(flag.item() == 1), "rows_per_rank": [int(x) for x in sizes_h],
                        "what": "on every rank: word checksum of each gathered block == the owner's local checksum, own block bit-identical"}

    # ---- e2e through the C ABI with pinned host buffers ------------------------------------------
    alloc = pinned_alloc_factory()
    fused = not (args.e2e_dense or args.e2e_two_calls)
    hp.host_step(alloc, ragged=not args.e2e_dense, fused=fused)  # warm-up: allocates the pinned buffers
    barrier()
    t0 = time.perf_counter()
    e2e_stats = [hp.host_step(alloc, ragged=not args.e2e_dense, fused=fused) for _ in range(args.e2e_steps)]
    barrier()
    t_e2e = time.perf_counter() - t0
    t_e = torch.tensor([t_e2e, float(e2e_stats[-1]["n_candidates"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t_e.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t_e.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        t_e2e_max, n_e2e_total = float(tm[0]), float(ts[1])
    else:
        t_e2e_max, n_e2e_total = t_e2e, float(e2e_stats[-1]["n_candidates"])
    e2e_value = n_e2e_total * args.e2e_steps / t_e2e_max
    for k_step, st_ in enumerate(e2e_stats):
        ph = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st_.get("phases_ms", {}).items() if not isinstance(v, dict)}
        log(f"[rank {rank}] e2e step {k_step} phases (ms): {json.dumps(ph)}")
    log(f"[rank {rank}] e2e wall per step: {1000.0 * t_e2e / max(args.e2e_steps, 1):.1f} ms")

    if rank == 0:
        # ---- roofline of the dominant kernel ------------------------------------------------------
        is4d = hasattr(raw, "tof_indptr")
        L = raw.cycle.shape[1]
        cont = _lib.fetch_candidates(hp.dev_raw, int(hp.n_precursors * sel.candidate_count))
        m = cont["score"] > 0
        c_sc = (cont["frame_stop"][m].astype(np.int64) // L - cont["frame_start"][m].astype(np.int64) // L)
        c_sc_mean = float(c_sc.mean()) if m.any() else 0.0
        cyc_rt = raw.rt_values[int(raw.zeroth_frame)::L] if is4d else raw.rt_values[::L]
        cyc_s = float(np.mean(np.diff(cyc_rt))) if len(cyc_rt) > 1 else 1.0
        c_sel = 16 * math.ceil(max(2 * p["rt_tolerance"] / cyc_s, 30) / 16)
        c_sel = min(c_sel, raw.precursor_cycle_max_index)
        if is4d:
            ab = algorithmic_bytes_4d(raw, lib, p, c_sel * L, c_sc_mean * L)
        else:
            ab = algorithmic_bytes(raw, lib, c_sel, c_sc_mean)
        sel_k = float(np.mean([s["select_kernel_ms"] for s in stats]))
        sc_k = float(np.mean([s["score_kernel_ms"] for s in stats]))
        peak, peak_src = measured_peaks()
        if sel_k >= sc_k:
            dom, dur, bytes_launch, unit_desc = ("adb_select4d_kernel" if is4d else "adb_select_fused_kernel"), sel_k, ab["b_prec"] * hp.n_precursors, f"{ab['b_prec']:.0f} B/precursor x {hp.n_precursors} precursors (C_sel={c_sel})"
        else:
            dom, dur, bytes_launch, unit_desc = ("adb_score4d_kernel" if is4d else "dp_score_passes"), sc_k, ab["b_cand"] * n_cand, f"{ab['b_cand']:.0f} B/candidate x {n_cand} candidates (C_sc={c_sc_mean:.1f})"
        achieved = bytes_launch / (dur * 1e-3) / 1e9
        traffic = None  # GB per launch: ncu DRAM bytes per unit (profiles/dram_traffic.json) x the units of this launch
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            try:
                ent = json.load(open(tpath)).get(dom) or {}
                if ent.get("bytes_per_unit"):
                    units = hp.n_precursors if ent.get("unit") == "precursor" else n_cand
                    traffic = float(ent["bytes_per_unit"]) * units / 1e9
            except Exception:
                traffic = None
        # `achieved` follows the contract: the REFERENCE ALGORITHM's bytes over the measured duration.  The index
        # layouts make the kernels touch fewer bytes than that, so the DRAM-side reading is given next to it:
        # dram_gbps = ncu DRAM bytes of the same kernel(s) / the same duration, dram_frac = its share of the HBM peak.
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_unit": "GB per launch (ncu dram bytes per unit x units)",
                    "dram_gbps": (traffic / (dur * 1e-3)) if traffic else None,
                    "dram_frac": (traffic / (dur * 1e-3) / peak) if traffic else None,
                    "algorithmic_gb_per_launch": bytes_launch / 1e9, "kernel": dom, "kernel_ms": dur, "algorithmic_bytes": unit_desc, "peak_source": peak_src,
                    "other_kernel": {"selection_kernel_ms": sel_k, "scoring_passes_ms": sc_k,
                                     "select_frac": ab["b_prec"] * hp.n_precursors / (sel_k * 1e-3) / 1e9 / peak,
                                     "score_frac": ab["b_cand"] * n_cand / (sc_k * 1e-3) / 1e9 / peak,
                                     "note": "3-D scoring = the data-parallel passes dp_setup .. dp_write (8 kernels per batch of 2 M candidates); selection = one fused kernel"}}
        parity = None
        if not args.no_parity:
            t0 = time.time()
            parity = parity_spot_check(hp, raw, lib, sel, sc, kernel, cont, n_prec=args.parity_precursors)
            log(f"parity spot check done in {time.time() - t0:.1f}s: {json.dumps(parity)}")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.time()
            cpu = cpu_baseline(raw, lib, sel, sc, kernel)
            log(f"cpu baseline done in {time.time() - t0:.1f}s: {cpu['value']:.0f} candidates/s on {cpu['cores']} threads")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t_wall_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": workload_description(args.workload, hp.n_precursors) + (f", one file per GPU x {world}" if world > 1 else ""),
                       "candidates_per_step": n_cand_total, "l2": "inputs (raw file + library) exceed the 126 MB L2; no explicit flush",
                       "device_ms_per_step": 1000.0 * dev_s_max / args.steps,
                       "stage_ms": {"selection": float(np.mean([s["select_ms"] for s in stats])),
                                    "scoring": float(np.mean([s["score_ms"] for s in stats]))},
                       "precursors_per_s_selection": hp.n_precursors / (np.mean([s["select_ms"] for s in stats]) * 1e-3),
                       "candidates_per_s_scoring": n_cand / (np.mean([s["score_ms"] for s in stats]) * 1e-3)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_stats[-1]["h2d_bytes"],
                    "d2h_bytes_per_step": e2e_stats[-1]["d2h_bytes"], "steps": args.e2e_steps,
                    "valid_rows": e2e_stats[-1]["valid"], "fragment_rows": e2e_stats[-1]["n_fragments"],
                    "path": (("adb_library_create (H2D) + adb_select_score_candidates_ragged (selection, candidate table D2H while the first "
                              "scoring block runs, scoring from the resident table; feature rows of the valid candidates + their kept "
                              "fragment slots compacted on the device per row block and copied while the next block is scored)")
                             if fused else
                             ("adb_library_create (H2D) + adb_select_candidates_resident + adb_fetch_candidate_table (D2H) + "
                              + ("adb_score_candidates (candidate table H2D, dense score/fragment tables D2H in 4 row blocks overlapped with the kernel)"
                                 if args.e2e_dense else
                                 "adb_score_candidates_ragged (candidate table H2D; feature rows of the valid candidates + their kept fragment slots, "
                                 "compacted on the device per row block and copied while the next block is scored)")))
                            + ", pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if parity is not None:
            line["parity"] = parity
        if gather_check is not None:
            line["gather_check"] = gather_check
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_operator:
            op = bench_operator(raw, pdf, fdf, sel, sc, steps=args.operator_steps)
            log(f"operator path: {json.dumps(op)}")
            line["e2e_operator"] = op
        if world == 1 and not args.no_fragcomp:
            fc = bench_fragcomp()
            log(f"fragment competition: {json.dumps(fc)}")
            line["fragcomp"] = fc
        if world == 1 and args.workload == "config3" and args.precursors is None and not args.no_other_workloads:
            line["other_workloads"] = other_workload_lines(("config2", "config4"))
        print(json.dumps(line), flush=True)
    hp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

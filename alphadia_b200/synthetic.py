"""Deterministic synthetic DIA runs + spectral libraries (SURVEY.md §8d data model).

The arrays produced here have *exactly* the layouts of the reference's raw-file
views, so the same buffers feed the CUDA engine, the C oracle and (in this
container only) the reference's numba classes:

* 3-D (Thermo-shape) runs follow ``AlphaRawJIT``'s 15 fields
  (reference ``alphadia/search/jitclasses/alpharaw_jit.py:78-138``, built at
  ``alphadia/raw_data/alpharaw_wrapper.py:86-156``).
* Libraries follow ``precursors_flat_schema`` / ``fragments_flat_schema``
  (reference ``alphadia/validation/schemas.py:11-48``).

Nothing here touches the oracle or the reference.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import pandas as pd

ISOTOPE_MASS_DIFF = 1.0033548350700006  # reference selection/utils.py:35


@dataclass
class RawFile3D:
    """Host-side 3-D raw file; attribute names mirror the reference's ``AlphaRaw``
    wrapper (``raw_data/alpharaw_wrapper.py:22-71``) so either object can be
    handed to the engine."""

    cycle: np.ndarray  # f64 [1, L, 1, 2]
    rt_values: np.ndarray  # f32 [n_spec], seconds
    peak_start_idx_list: np.ndarray  # i64 [n_spec]
    peak_stop_idx_list: np.ndarray  # i64 [n_spec]
    mz_values: np.ndarray  # f32 [n_peaks] ascending within a spectrum
    intensity_values: np.ndarray  # f32 [n_peaks]
    mobility_values: np.ndarray = field(
        default_factory=lambda: np.array([1e-6, 0], dtype=np.float32)
    )
    zeroth_frame: int = 0
    scan_max_index: int = 1
    has_mobility: bool = False
    has_ms1: bool = True

    @property
    def cycle_len(self) -> int:
        return int(self.cycle.shape[1])

    @property
    def precursor_cycle_max_index(self) -> int:
        return len(self.rt_values) // self.cycle_len

    @property
    def frame_max_index(self) -> int:
        return len(self.rt_values) - 1

    @property
    def n_peaks(self) -> int:
        return int(self.mz_values.shape[0])

    # scalars only used for bookkeeping in the reference (alpharaw_wrapper.py:95-108)
    @property
    def max_mz_value(self) -> np.float32:
        return np.float32(self.cycle[self.cycle > 0].max()) if (self.cycle > 0).any() else np.float32(0)

    @property
    def min_mz_value(self) -> np.float32:
        return np.float32(self.cycle[self.cycle > 0].min()) if (self.cycle > 0).any() else np.float32(0)


def _float_sort_key(spec_idx: np.ndarray, mz: np.ndarray) -> np.ndarray:
    """(spectrum, m/z) -> uint64 key; positive f32 bit patterns are monotone."""
    return (spec_idx.astype(np.uint64) << np.uint64(32)) | mz.view(np.uint32).astype(
        np.uint64
    )


def make_library(
    n_precursors: int,
    rng: np.random.Generator,
    *,
    quad_lo: float,
    quad_hi: float,
    rt_lo: float,
    rt_hi: float,
    n_fragments: int = 12,
    n_isotopes: int = 4,
    with_strings: bool = True,
) -> tuple[pd.DataFrame, pd.DataFrame]:
    """Flat spectral library (precursor_df, fragment_df), SURVEY §8d."""
    P, F = n_precursors, n_fragments
    idx = np.arange(P, dtype=np.uint32)
    charge = rng.integers(2, 4, size=P).astype(np.uint8)
    mz = rng.uniform(quad_lo + 1.0, quad_hi - 3.0, size=P).astype(np.float32)
    rt = rng.uniform(rt_lo, rt_hi, size=P).astype(np.float32)
    iso = np.array([0.5, 0.3, 0.15, 0.05], dtype=np.float32)[:n_isotopes]

    prec = {
        "elution_group_idx": idx.copy(),
        "precursor_idx": idx.copy(),
        "channel": np.zeros(P, dtype=np.uint32),
        "decoy": (idx % 2).astype(np.uint8),
        "flat_frag_start_idx": (idx * F).astype(np.uint32),
        "flat_frag_stop_idx": ((idx + 1) * F).astype(np.uint32),
        "charge": charge,
        "rt_library": rt,
        "mobility_library": np.zeros(P, dtype=np.float32),
        "mz_library": mz,
    }
    for i in range(n_isotopes):
        prec[f"i_{i}"] = np.full(P, iso[i], dtype=np.float32)
    precursor_df = pd.DataFrame(prec)
    if not with_strings:  # schema-required object columns only (cheap constants)
        for col in ("proteins", "genes"):
            precursor_df[col] = pd.Series(np.full(P, "P0", dtype=object), dtype=object)
    if with_strings:
        aa = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
        letters = aa[rng.integers(0, 20, size=(P, 9))]
        seqs = np.array(["".join(r) for r in letters], dtype=object) if P <= 200_000 else None
        if seqs is None:
            # cheap path for multi-million libraries: a handful of distinct strings
            pool = np.array(["".join(r) for r in aa[rng.integers(0, 20, size=(4096, 9))]], dtype=object)
            seqs = pool[rng.integers(0, 4096, size=P)]
        precursor_df["sequence"] = pd.Series(seqs, dtype=object)
        for col in ("proteins", "genes"):
            precursor_df[col] = pd.Series(np.full(P, "P0", dtype=object), dtype=object)
        for col in ("mods", "mod_sites"):
            precursor_df[col] = pd.Series(np.full(P, "", dtype=object), dtype=object)

    fmz = rng.uniform(200.0, 1800.0, size=P * F).astype(np.float32)
    fint = rng.uniform(0.05, 1.0, size=P * F).astype(np.float32)
    ftype = np.where(rng.integers(0, 2, size=P * F) == 0, 98, 121).astype(np.uint8)
    fragment_df = pd.DataFrame(
        {
            "mz_library": fmz,
            "intensity": fint,
            "cardinality": np.ones(P * F, dtype=np.uint8),
            "type": ftype,
            "loss_type": np.zeros(P * F, dtype=np.uint8),
            "charge": np.ones(P * F, dtype=np.uint8),
            "number": rng.integers(1, 20, size=P * F).astype(np.uint8),
            "position": rng.integers(1, 20, size=P * F).astype(np.uint8),
        }
    )
    return precursor_df, fragment_df


def make_run_3d(
    precursor_df: pd.DataFrame,
    fragment_df: pd.DataFrame,
    rng: np.random.Generator,
    *,
    n_cycles: int,
    n_windows: int,
    quad_lo: float,
    quad_hi: float,
    cycle_seconds: float,
    n_noise_ms2: int,
    n_noise_ms1: int,
    planted_fraction: float = 0.5,
    max_planted: int | None = None,
    elution_sigma: float = 3.0,
    rt_jitter: float = 5.0,
) -> tuple[RawFile3D, np.ndarray]:
    """Thermo-shape run: every cycle = 1 MS1 + ``n_windows`` MS2 spectra.

    Returns the raw file and the apex RT planted for each precursor (NaN = not planted).
    """
    L = n_windows + 1
    n_spec = n_cycles * L
    width = (quad_hi - quad_lo) / n_windows
    cycle = np.zeros((1, L, 1, 2), dtype=np.float64)
    cycle[0, 0, 0, :] = -1.0
    cycle[0, 1:, 0, 0] = quad_lo + width * np.arange(n_windows)
    cycle[0, 1:, 0, 1] = quad_lo + width * (np.arange(n_windows) + 1)

    rt_values = (np.arange(n_spec, dtype=np.float64) * (cycle_seconds / L)).astype(np.float32)
    is_ms1 = (np.arange(n_spec) % L) == 0
    n_noise = np.where(is_ms1, n_noise_ms1, n_noise_ms2).astype(np.int64)

    # ---- noise: per-spectrum sorted m/z -------------------------------------------------
    total_noise = int(n_noise.sum())
    spec_of_noise = np.repeat(np.arange(n_spec, dtype=np.int64), n_noise)
    noise_mz = rng.uniform(150.0, 1900.0, size=total_noise).astype(np.float32)
    noise_int = rng.exponential(200.0, size=total_noise).astype(np.float32) + np.float32(1.0)

    # ---- planted signal -----------------------------------------------------------------
    P = len(precursor_df)
    decoy = precursor_df["decoy"].values
    targets = np.flatnonzero(decoy == 0)
    n_plant = int(len(targets) * planted_fraction)
    if max_planted is not None:
        n_plant = min(n_plant, max_planted)
    planted = np.sort(rng.choice(targets, size=n_plant, replace=False)) if n_plant else np.zeros(0, np.int64)
    apex = np.full(P, np.nan, dtype=np.float64)
    apex[planted] = precursor_df["rt_library"].values[planted] + rng.normal(0, rt_jitter, size=n_plant)

    sig_spec, sig_mz, sig_int = [], [], []
    if n_plant:
        pmz = precursor_df["mz_library"].values.astype(np.float64)[planted]
        pch = precursor_df["charge"].values.astype(np.float64)[planted]
        window = np.clip(((pmz - quad_lo) // width).astype(np.int64), 0, n_windows - 1)
        apex_cycle = apex[planted] / cycle_seconds
        half = int(np.ceil(3.0 * elution_sigma / cycle_seconds))
        offs = np.arange(-half, half + 1)
        cyc = np.rint(apex_cycle)[:, None].astype(np.int64) + offs[None, :]  # (n_plant, W)
        ok = (cyc >= 0) & (cyc < n_cycles)
        cyc_c = np.clip(cyc, 0, n_cycles - 1)

        # MS2: fragments in the precursor's window
        fs = precursor_df["flat_frag_start_idx"].values[planted].astype(np.int64)
        fe = precursor_df["flat_frag_stop_idx"].values[planted].astype(np.int64)
        nf = int((fe - fs).max())
        fmz_all = fragment_df["mz_library"].values
        fint_all = fragment_df["intensity"].values
        for k in range(nf):
            has = (fs + k) < fe
            fi = np.minimum(fs + k, len(fmz_all) - 1)
            spec = cyc_c * L + 1 + window[:, None]
            t = rt_values[spec].astype(np.float64)
            shape = np.exp(-0.5 * ((t - apex[planted][:, None]) / elution_sigma) ** 2)
            inten = shape * fint_all[fi][:, None] * 1e5
            jit = 1.0 + rng.normal(0, 2e-6, size=spec.shape)
            mzv = fmz_all[fi].astype(np.float64)[:, None] * jit
            m = ok & has[:, None] & (inten > 1.0)
            sig_spec.append(spec[m])
            sig_mz.append(mzv[m].astype(np.float32))
            sig_int.append(inten[m].astype(np.float32))
        # MS1: isotopes
        iso_cols = [c for c in precursor_df.columns if c.startswith("i_")]
        for i, col in enumerate(iso_cols):
            ab = precursor_df[col].values.astype(np.float64)[planted]
            spec = cyc_c * L
            t = rt_values[spec].astype(np.float64)
            shape = np.exp(-0.5 * ((t - apex[planted][:, None]) / elution_sigma) ** 2)
            inten = shape * ab[:, None] * 1e6
            jit = 1.0 + rng.normal(0, 2e-6, size=spec.shape)
            mzv = (pmz + i * ISOTOPE_MASS_DIFF / pch)[:, None] * jit
            m = ok & (inten > 1.0)
            sig_spec.append(spec[m])
            sig_mz.append(mzv[m].astype(np.float32))
            sig_int.append(inten[m].astype(np.float32))

    if sig_spec:
        all_spec = np.concatenate([spec_of_noise] + sig_spec)
        all_mz = np.concatenate([noise_mz] + sig_mz)
        all_int = np.concatenate([noise_int] + sig_int)
    else:
        all_spec, all_mz, all_int = spec_of_noise, noise_mz, noise_int
    del spec_of_noise, noise_mz, noise_int, sig_spec, sig_mz, sig_int

    order = np.argsort(_float_sort_key(all_spec, all_mz), kind="stable")
    mz_values = np.ascontiguousarray(all_mz[order])
    intensity_values = np.ascontiguousarray(all_int[order])
    counts = np.bincount(all_spec, minlength=n_spec).astype(np.int64)
    del order, all_spec, all_mz, all_int
    stop = np.cumsum(counts)
    start = stop - counts

    raw = RawFile3D(
        cycle=cycle,
        rt_values=rt_values,
        peak_start_idx_list=start.astype(np.int64),
        peak_stop_idx_list=stop.astype(np.int64),
        mz_values=mz_values,
        intensity_values=intensity_values,
    )
    return raw, apex


# --------------------------------------------------------------------------------------
# The named configurations of BASELINE.json / SURVEY.md §8d
# --------------------------------------------------------------------------------------
CONFIGS_3D = {
    # config 1: plumbing (reference-runnable on CPU)
    "config1": dict(
        seed=1, n_precursors=1000, n_cycles=50, n_windows=9, quad_lo=400.0, quad_hi=1200.0,
        cycle_seconds=1.5, n_noise_ms2=200, n_noise_ms1=200, rt_tolerance=30.0,
        planted_fraction=0.5, max_planted=None,
    ),
    # a mid-size case for parity tests (a few thousand candidates, C_sel = 64..)
    "parity_small": dict(
        seed=11, n_precursors=600, n_cycles=160, n_windows=12, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=400, n_noise_ms1=800, rt_tolerance=25.0,
        planted_fraction=0.6, max_planted=None,
    ),
    # 20 library fragments per precursor: selection smooths 20 fragment layers, scoring picks its top 12 (or 6) of 20
    "parity_f20": dict(
        seed=31, n_precursors=300, n_cycles=120, n_windows=12, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=400, n_noise_ms1=800, rt_tolerance=25.0,
        planted_fraction=0.6, max_planted=None, n_fragments=20,
    ),
    # 48 library fragments per precursor (transfer-library scale: top_k_fragments = 9999 keeps all of them)
    "parity_f48": dict(
        seed=37, n_precursors=200, n_cycles=120, n_windows=12, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=400, n_noise_ms1=800, rt_tolerance=25.0,
        planted_fraction=0.6, max_planted=None, n_fragments=48,
    ),
    # 96 library fragments per precursor: beyond every dense table, close to the device limit of 128 per precursor
    "parity_f96": dict(
        seed=41, n_precursors=120, n_cycles=120, n_windows=12, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=400, n_noise_ms1=800, rt_tolerance=25.0,
        planted_fraction=0.6, max_planted=None, n_fragments=96,
    ),
    # config 2: 50k precursors, Thermo shape (91 200 spectra, ~1.4e8 peaks)
    "config2": dict(
        seed=2, n_precursors=50_000, n_cycles=1200, n_windows=75, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=1500, n_noise_ms1=4000, rt_tolerance=100.0,
        planted_fraction=0.5, max_planted=None,
    ),
    # config 3: 2M precursors, same frames
    "config3": dict(
        seed=3, n_precursors=2_000_000, n_cycles=1200, n_windows=75, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=1500, n_noise_ms1=4000, rt_tolerance=100.0,
        planted_fraction=0.5, max_planted=60_000,
    ),
    # many short cycles, few peaks: with a huge rt_tolerance the selection window exceeds 1024 cycles (legacy kernel pair)
    "long_run": dict(
        seed=31, n_precursors=300, n_cycles=1400, n_windows=3, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=0.5, n_noise_ms2=60, n_noise_ms1=120, rt_tolerance=2000.0,
        planted_fraction=0.7, max_planted=None,
    ),
    # config 5 (per file): 500k precursors
    "config5": dict(
        seed=50, n_precursors=500_000, n_cycles=1200, n_windows=75, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=1.0, n_noise_ms2=1500, n_noise_ms1=4000, rt_tolerance=100.0,
        planted_fraction=0.5, max_planted=60_000,
    ),
}


def make_config_3d(name: str, *, seed: int | None = None, n_precursors: int | None = None,
                   with_strings: bool = True, scale_noise: float = 1.0):
    """Build (raw, precursor_df, fragment_df, params) for a named 3-D configuration."""
    p = dict(CONFIGS_3D[name])
    if seed is not None:
        p["seed"] = seed
    if n_precursors is not None:
        p["n_precursors"] = n_precursors
    rng = np.random.default_rng(p["seed"])
    run_s = p["n_cycles"] * p["cycle_seconds"]
    margin = min(60.0, run_s * 0.15)
    precursor_df, fragment_df = make_library(
        p["n_precursors"], rng, quad_lo=p["quad_lo"], quad_hi=p["quad_hi"],
        rt_lo=margin, rt_hi=run_s - margin, with_strings=with_strings, n_fragments=p.get("n_fragments", 12),
    )
    raw, apex = make_run_3d(
        precursor_df, fragment_df, rng,
        n_cycles=p["n_cycles"], n_windows=p["n_windows"], quad_lo=p["quad_lo"], quad_hi=p["quad_hi"],
        cycle_seconds=p["cycle_seconds"],
        n_noise_ms2=int(p["n_noise_ms2"] * scale_noise), n_noise_ms1=int(p["n_noise_ms1"] * scale_noise),
        planted_fraction=p["planted_fraction"], max_planted=p["max_planted"],
    )
    p["apex_rt"] = apex
    return raw, precursor_df, fragment_df, p


# ======================================================================================
# 4-D (timsTOF shape) runs: TimsTOFTransposeJIT layout
# (reference alphadia/search/jitclasses/bruker_jit.py:20-137, built at alphadia/raw_data/bruker.py:119-152)
# ======================================================================================
@dataclass
class RawFile4D:
    cycle: np.ndarray  # f64 [1, Fr, Sc, 2] quad window per (frame in cycle, scan); MS1 frame = (-1, -1)
    rt_values: np.ndarray  # f64 [n_frames] (frame 0 is the empty "zeroth" frame)
    mobility_values: np.ndarray  # f64 [Sc], descending
    mz_values: np.ndarray  # f64 [n_tof] ascending tof -> m/z grid
    tof_indptr: np.ndarray  # i64 [n_tof + 1]  CSR by tof index
    push_indices: np.ndarray  # u32 [n_events], ascending inside a tof row; push = frame * Sc + scan
    intensity_values: np.ndarray  # u16 [n_events]
    zeroth_frame: int = 1
    has_mobility: bool = True
    has_ms1: bool = True

    @property
    def scan_max_index(self) -> int:
        return int(self.cycle.shape[2])

    @property
    def frame_max_index(self) -> int:
        return int(len(self.rt_values))

    @property
    def precursor_cycle_max_index(self) -> int:
        return self.frame_max_index // int(self.cycle.shape[1])

    @property
    def dia_mz_cycle(self) -> np.ndarray:
        return np.ascontiguousarray(self.cycle.reshape(-1, 2))

    @property
    def dia_precursor_cycle(self) -> np.ndarray:
        fr, sc = self.cycle.shape[1], self.cycle.shape[2]
        return (np.arange(fr * sc, dtype=np.int64) // sc).astype(np.int64)

    @property
    def n_events(self) -> int:
        return int(self.push_indices.shape[0])


CONFIGS_4D = {
    # small case the reference can run on the CPU in seconds (golden vectors)
    "parity_4d": dict(
        seed=21, n_precursors=240, n_cycles=90, n_ms2_frames=3, n_scans=96, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=0.6, mob_hi=1.30, mob_lo=0.70, tof_ppm=4.0, mz_lo=150.0, mz_hi=1900.0,
        noise_per_push=6, rt_tolerance=12.0, mobility_tolerance=0.12, planted_fraction=0.7,
    ),
    # like parity_4d, but every MS2 frame isolates ONE window on all scans and neighbouring windows overlap by 60 Th: about a
    # third of the precursors are seen by two frames of the cycle (two observations per candidate in scoring)
    "parity_4d_overlap": dict(
        seed=23, n_precursors=240, n_cycles=90, n_ms2_frames=3, n_scans=96, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=0.6, mob_hi=1.30, mob_lo=0.70, tof_ppm=4.0, mz_lo=150.0, mz_hi=1900.0,
        noise_per_push=6, rt_tolerance=12.0, mobility_tolerance=0.12, planted_fraction=0.7,
        diagonal=False, window_overlap=30.0,
    ),
    # parity_4d with 20 library fragments per precursor
    "parity_4d_f20": dict(
        seed=27, n_precursors=200, n_cycles=90, n_ms2_frames=3, n_scans=96, quad_lo=400.0, quad_hi=1000.0,
        cycle_seconds=0.6, mob_hi=1.30, mob_lo=0.70, tof_ppm=4.0, mz_lo=150.0, mz_hi=1900.0,
        noise_per_push=6, rt_tolerance=12.0, mobility_tolerance=0.12, planted_fraction=0.7, n_fragments=20,
    ),
    # config 4 of BASELINE.json: 200k precursors, (1 + 8) x 928 cycle, 800 cycles
    "config4": dict(
        seed=4, n_precursors=200_000, n_cycles=800, n_ms2_frames=8, n_scans=928, quad_lo=400.0, quad_hi=1200.0,
        cycle_seconds=0.95, mob_hi=1.45, mob_lo=0.65, tof_ppm=4.0, mz_lo=150.0, mz_hi=1900.0,
        noise_per_push=40, rt_tolerance=50.0, mobility_tolerance=0.04, planted_fraction=0.5, sorted_noise=True,
    ),
}


def make_config_4d(name: str, *, seed: int | None = None, n_precursors: int | None = None, with_strings: bool = True):
    """(raw4d, precursor_df, fragment_df, params) for a named timsTOF-shape configuration.

    Geometry: every cycle = 1 MS1 frame + ``n_ms2_frames`` diaPASEF frames; each MS2 frame isolates two quadrupole
    windows, the upper half of the m/z band in the upper-mobility half of the scans and the lower half in the
    lower-mobility scans (a coarse diaPASEF diagonal).  m/z grid: geometric, ``tof_ppm`` spacing.
    """
    p = dict(CONFIGS_4D[name])
    if seed is not None:
        p["seed"] = seed
    if n_precursors is not None:
        p["n_precursors"] = n_precursors
    rng = np.random.default_rng(p["seed"])
    Fr, Sc, ncyc = p["n_ms2_frames"] + 1, p["n_scans"], p["n_cycles"]
    n_frames = ncyc * Fr + 1  # + zeroth frame
    frame_s = p["cycle_seconds"] / Fr
    rt_values = np.concatenate([[0.0], (np.arange(ncyc * Fr) + 1) * frame_s]).astype(np.float64)
    mobility_values = np.linspace(p["mob_hi"], p["mob_lo"], Sc).astype(np.float64)
    n_tof = int(np.log(p["mz_hi"] / p["mz_lo"]) / (p["tof_ppm"] * 1e-6)) + 1
    mz_grid = (p["mz_lo"] * np.exp(np.arange(n_tof) * p["tof_ppm"] * 1e-6)).astype(np.float64)

    # quad windows: MS2 frame f (1-based) covers band f; scans [0, Sc/2) -> upper half of the band, rest -> lower half
    band = (p["quad_hi"] - p["quad_lo"]) / p["n_ms2_frames"]
    cycle = np.full((1, Fr, Sc, 2), -1.0, dtype=np.float64)
    half = Sc // 2
    diagonal, overlap = p.get("diagonal", True), float(p.get("window_overlap", 0.0))
    for f in range(1, Fr):
        lo = p["quad_lo"] + band * (f - 1)
        if diagonal:
            cycle[0, f, :half, 0], cycle[0, f, :half, 1] = lo + band / 2, lo + band
            cycle[0, f, half:, 0], cycle[0, f, half:, 1] = lo, lo + band / 2
        else:  # one window per frame on all scans, widened so that neighbouring frames overlap
            cycle[0, f, :, 0], cycle[0, f, :, 1] = lo - overlap, lo + band + overlap

    # library: precursors live inside one (frame, half) window, mobility inside that half
    P, F = p["n_precursors"], p.get("n_fragments", 12)
    run_s = rt_values[-1]
    margin = min(60.0, run_s * 0.15)
    precursor_df, fragment_df = make_library(P, rng, quad_lo=p["quad_lo"], quad_hi=p["quad_hi"], rt_lo=margin,
                                             rt_hi=run_s - margin, with_strings=with_strings, n_fragments=F)
    wf = rng.integers(1, Fr, size=P)
    wh = rng.integers(0, 2, size=P)
    if diagonal:
        wlo = p["quad_lo"] + band * (wf - 1) + np.where(wh == 0, band / 2, 0.0)
        pmz = (wlo + rng.uniform(0.5, band / 2 - 3.5, size=P)).astype(np.float32)
    else:
        wlo = p["quad_lo"] + band * (wf - 1)
        pmz = (wlo + rng.uniform(0.5, band - 3.5, size=P)).astype(np.float32)
    precursor_df["mz_library"] = pmz
    mob_span = (p["mob_hi"] - p["mob_lo"]) / 2
    edge = p["mobility_tolerance"] * 1.2
    mob = np.where(wh == 0, p["mob_hi"] - rng.uniform(edge, mob_span - edge, size=P),
                   p["mob_hi"] - mob_span - rng.uniform(edge, mob_span - edge, size=P))
    precursor_df["mobility_library"] = mob.astype(np.float32)

    # ---- events: (tof, push, intensity) -------------------------------------------------------
    n_push = n_frames * Sc
    n_noise = int((n_push - Sc) * p["noise_per_push"])
    if p.get("sorted_noise", False):
        # large runs: draw the noise as a Poisson process over the (tof, push) key space, i.e. already in CSR order
        # (cumulative exponential gaps), so that no 3e8-element sort is needed; frame 0 stays empty
        span = n_push - Sc
        total = float(n_tof) * float(span)
        gaps = rng.exponential(total / n_noise, size=n_noise)
        keys = np.cumsum(gaps)
        del gaps
        keys = keys[keys < total].astype(np.int64)
        ev_tof = keys // span
        ev_push = keys - ev_tof * span + Sc
        del keys
        n_noise = len(ev_tof)
        ev_int = np.minimum(rng.exponential(60.0, size=n_noise).astype(np.float32) + 10.0, 60000.0)
    else:
        ev_push = rng.integers(Sc, n_push, size=n_noise).astype(np.int64)  # frame 0 stays empty
        ev_tof = rng.integers(0, n_tof, size=n_noise).astype(np.int64)
        ev_int = np.minimum(rng.exponential(60.0, size=n_noise) + 10.0, 60000.0)

    targets = np.flatnonzero(precursor_df["decoy"].values == 0)
    n_plant = int(len(targets) * p["planted_fraction"])
    planted = np.sort(rng.choice(targets, size=n_plant, replace=False)) if n_plant else np.zeros(0, np.int64)
    apex_rt = np.full(P, np.nan)
    apex_rt[planted] = precursor_df["rt_library"].values[planted] + rng.normal(0, 1.5, size=n_plant)
    sig_push, sig_tof, sig_int = [], [], []
    if n_plant:
        rt_sigma, scan_sigma = 1.2, 3.0
        hc = int(np.ceil(3 * rt_sigma / p["cycle_seconds"]))
        hs = int(np.ceil(2.5 * scan_sigma))
        c_off = np.arange(-hc, hc + 1)
        s_off = np.arange(-hs, hs + 1)
        apex_cycle = np.rint(apex_rt[planted] / p["cycle_seconds"]).astype(np.int64)
        apex_scan = np.rint((p["mob_hi"] - mob[planted]) / (p["mob_hi"] - p["mob_lo"]) * (Sc - 1)).astype(np.int64)
        cyc = apex_cycle[:, None, None] + c_off[None, :, None]      # (n, Wc, 1)
        scn = apex_scan[:, None, None] + s_off[None, None, :]       # (n, 1, Ws)
        ok = (cyc >= 0) & (cyc < ncyc) & (scn >= 0) & (scn < Sc)
        shape_s = np.exp(-0.5 * (s_off[None, None, :] / scan_sigma) ** 2)
        fs = precursor_df["flat_frag_start_idx"].values[planted].astype(np.int64)
        fmz_all, fint_all = fragment_df["mz_library"].values.astype(np.float64), fragment_df["intensity"].values
        pch = precursor_df["charge"].values[planted].astype(np.float64)

        def add(frame_in_cycle, mz_target, amp):
            frame = cyc * Fr + 1 + frame_in_cycle[:, None, None]
            t = rt_values[np.clip(frame, 0, n_frames - 1)]
            inten = amp[:, None, None] * np.exp(-0.5 * ((t - apex_rt[planted][:, None, None]) / rt_sigma) ** 2) * shape_s
            tof = np.rint(np.log(mz_target / p["mz_lo"]) / (p["tof_ppm"] * 1e-6)).astype(np.int64)
            tof = np.broadcast_to(tof[:, None, None], inten.shape)
            m = ok & (inten > 12.0) & (tof >= 0) & (tof < n_tof)
            push = frame * Sc + scn
            sig_push.append(np.broadcast_to(push, inten.shape)[m])
            sig_tof.append(tof[m])
            sig_int.append(np.minimum(inten[m], 60000.0))

        for k in range(F):
            jittered = fmz_all[fs + k] * (1 + rng.normal(0, 1.5e-6, size=n_plant))
            add(wf[planted], jittered, fint_all[fs + k] * 6000.0)
            if not diagonal:  # the neighbouring frames whose widened window also isolates the precursor see it too (weaker)
                for step in (-1, 1):
                    nf = wf[planted] + step
                    nlo = p["quad_lo"] + band * (nf - 1) - overlap
                    seen = (nf >= 1) & (nf < Fr) & (pmz[planted] >= nlo) & (pmz[planted] <= nlo + band + 2 * overlap)
                    add(np.where(seen, nf, wf[planted]), jittered, np.where(seen, fint_all[fs + k] * 3000.0, 0.0))
        for i, ab in enumerate([0.5, 0.3, 0.15]):
            add(np.zeros(n_plant, np.int64), (pmz[planted].astype(np.float64) + i * ISOTOPE_MASS_DIFF / pch)
                * (1 + rng.normal(0, 1.5e-6, size=n_plant)), np.full(n_plant, ab * 20000.0))
    if p.get("sorted_noise", False):
        # the noise is sorted; sort the (much smaller) signal part and merge (a stable sort of two sorted runs is a merge)
        s_key = np.concatenate(sig_tof) * np.int64(n_push) + np.concatenate(sig_push) if sig_push else np.zeros(0, np.int64)
        s_int = np.concatenate(sig_int).astype(np.float32) if sig_int else np.zeros(0, np.float32)
        so = np.argsort(s_key, kind="stable")
        s_key, s_int = s_key[so], s_int[so]
        n_key = ev_tof * np.int64(n_push) + ev_push
        del ev_tof, ev_push
        pos = np.searchsorted(n_key, s_key, side="left") + np.arange(len(s_key))
        key = np.empty(len(n_key) + len(s_key), np.int64)
        all_int = np.empty(len(key), np.float32)
        is_sig = np.zeros(len(key), bool)
        is_sig[pos] = True
        key[is_sig], key[~is_sig] = s_key, n_key
        all_int[is_sig], all_int[~is_sig] = s_int, ev_int
        del n_key, ev_int, is_sig
    else:
        all_push = np.concatenate([ev_push] + sig_push)
        all_tof = np.concatenate([ev_tof] + sig_tof)
        all_int = np.concatenate([ev_int] + sig_int)
        key = all_tof * np.int64(n_push) + all_push
        order = np.argsort(key, kind="stable")
        key, all_int = key[order], all_int[order]
    # merge duplicate (tof, push) events: detector events are unique per (push, tof)
    first = np.ones(len(key), dtype=bool)
    first[1:] = key[1:] != key[:-1]
    grp = np.cumsum(first) - 1
    summed = np.bincount(grp, weights=all_int)
    key = key[first]
    tof_u = key // n_push
    push_u = (key % n_push).astype(np.uint32)
    inten_u = np.minimum(np.rint(summed), 65535).astype(np.uint16)
    tof_indptr = np.concatenate([[0], np.cumsum(np.bincount(tof_u, minlength=n_tof))]).astype(np.int64)
    raw = RawFile4D(cycle=cycle, rt_values=rt_values, mobility_values=mobility_values, mz_values=mz_grid,
                    tof_indptr=tof_indptr, push_indices=np.ascontiguousarray(push_u), intensity_values=inten_u)
    p["apex_rt"] = apex_rt
    return raw, precursor_df, fragment_df, p

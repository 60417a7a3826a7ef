// alphadia_b200 — candidate scoring, 3-D raw files, as DATA-PARALLEL PASSES over a batch of candidates (sm_100a).
//
// Replaces Candidate.process (alphadia/search/scoring/containers/candidate.py:166-481) and everything it calls:
// AlphaRawJIT.get_dense(absolute_masses=True) (jitclasses/alpharaw_jit.py:208-337), the quadrupole transfer function /
// template / observation importance (scoring/quadrupole.py:80-115,261-335), profiles (scoring/utils.py:26-66) and the 46
// features (scoring/features/*.py).
//
// One pass = one kernel = one function below; a "thread" owns one candidate, one (candidate, fragment) row or one
// (candidate, cycle) column and runs the reference's loops for it sequentially, in the reference's order (sequential f32 /
// f64 accumulation; this file is compiled with --fmad=false, so nothing is contracted).  No pass BODY uses warp-level
// cooperation, which has two consequences: every lane of a warp does useful work (the r1 tile-per-candidate kernel kept 13.9
// of 32 lanes busy), and the very same source runs thread by thread on a CPU (tests/hostsim/) where it is compared with
// the oracle without a GPU.
//
//   dp_setup      per candidate           fragment selection / m/z sort, isotope windows, quad-window hits, cycle window,
//                                         quadrupole transfer function; size of the candidate's workspace block
//   (largest block of every tile of DP_W slots, exclusive scan -> tile offsets)
//   dp_extract    per (candidate, row)    one XIC row (fragment x observation, or isotope) through the time-blocked m/z index:
//                                         per 32-cycle block one bucket read + a scan of the ~1 peaks inside the ppm window;
//                                         fragment mask (rows with signal -> work list of dp_fragment / dp_corr)
//   dp_template   per candidate           template, observation importance, distance-weight tables, precursor features
//   dp_fragment   per row with signal     best profile + envelope, area, weighted centres, mass error, cosine,
//                                         normalised profile, template correlation, FWHM, frame peak
//   dp_median     per (candidate, cycle)  median profile over the fragments (experimental_xic)
//   dp_corr       per (candidate, frag)   correlation with the median profile / legacy F x F correlation rows
//   dp_aggregate  per candidate           the remaining aggregate features, feature row + valid flag
//   dp_write      per (candidate, frag)   the per-fragment output table
//
// Workspace block of one candidate (float offsets from DpLayout), cube layout [observation][cycle][fragment]; the blocks of
// DP_W neighbouring slots are interleaved element by element (see DpPtr below), and every pass maps its threads with the
// slot index fastest, so the lanes of a warp touch consecutive addresses.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "adb_common.cuh"

#if defined(__CUDACC__)
#define ADB_HD __host__ __device__ __forceinline__
#else
#define ADB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define ADB_LD(p) __ldg(p)
#else
#define ADB_LD(p) (*(p))
#endif

// Workspace tiles: the blocks of DP_W consecutive slots are interleaved - element x of slot j lives at
// tile_base + x * DP_W + (j % DP_W) - so that the lanes of a warp (consecutive slots, same loop index) read and write
// consecutive addresses.  The pass bodies index their block through the small pointer wrappers below and are written as if
// the block were contiguous (DP_W = 1 is exactly that layout).
#ifndef DP_W
#define DP_W 8
// dp_template keeps its loops rolled: unrolled by the compiler the kernel is 9 600 instructions (154 KB) and stalls on instruction
// fetch (7 of 15 stall cycles per issue); rolled it is 5 400 and config 3 scoring goes 45.5 -> 44.2 ms.  The same limit on
// dp_fragment / dp_aggregate costs memory-level parallelism and measured slower (45.0 / 45.1 ms).
#ifndef DP_TPL_UNROLL_N
#define DP_TPL_UNROLL_N 1
#endif
#define DP_PRAGMA_(x) _Pragma(#x)
#define DP_PRAGMA(x) DP_PRAGMA_(x)
#define DP_TPL_UNROLL DP_PRAGMA(unroll DP_TPL_UNROLL_N)   // measured on config 3: 1 (contiguous blocks) 54.1 ms, 4 / 8 / 16: 46.7 ms, 32: 48.2 ms (padding to the largest block of a tile)
#endif

template <typename T>
struct DpPtr {  // T = float or int
  T* p;
  ADB_HD T& operator[](int i) const { return p[(ptrdiff_t)i * DP_W]; }
  ADB_HD DpPtr operator+(int o) const { return DpPtr{p + (ptrdiff_t)o * DP_W}; }
};
typedef DpPtr<float> FPtr;
typedef DpPtr<int> IPtr;
ADB_HD IPtr dp_as_int(FPtr f) { return IPtr{(int*)f.p}; }

struct DpDRef {  // a double kept as its two 32-bit halves in neighbouring elements
  float* lo;
  ADB_HD operator double() const {
    union { double d; uint32_t u[2]; } v;
    v.u[0] = ((const uint32_t*)lo)[0];
    v.u[1] = ((const uint32_t*)lo)[DP_W];
    return v.d;
  }
  ADB_HD const DpDRef& operator=(double x) const {
    union { double d; uint32_t u[2]; } v;
    v.d = x;
    ((uint32_t*)lo)[0] = v.u[0];
    ((uint32_t*)lo)[DP_W] = v.u[1];
    return *this;
  }
  ADB_HD const DpDRef& operator=(const DpDRef& o) const { return *this = (double)o; }
};
struct DPtr {
  float* p;
  ADB_HD DpDRef operator[](int i) const { return DpDRef{p + (ptrdiff_t)(2 * i) * DP_W}; }
  ADB_HD DPtr operator+(int o) const { return DPtr{p + (ptrdiff_t)(2 * o) * DP_W}; }
};
ADB_HD DPtr dp_as_double(FPtr f) { return DPtr{f.p}; }

#define DP_MAXF ADB_MAX_LIB_FRAGMENTS  // fragments one candidate may keep (top_k_fragments is clamped to the library's widest precursor)
#define DP_MED_LANES 16                // threads per candidate in dp_median

struct DpParams {
  DevRaw raw;
  DevLib lib;
  adb_scoring_config cfg;
  DevCandidatesIn cand;
  DevScoresOut out;
  int32_t out_k;         // width of the per-fragment output tables
  const int32_t* order;  // slot -> candidate index is order[base + slot] (null: base + slot)
  int64_t base;
  int64_t n;             // slots in this batch
  int32_t KS;            // fragment stride: min(top_k_fragments, widest precursor of the library)
  int32_t nIcap;         // min(library isotope columns, top_k_isotopes)
  // plan (per slot)
  uint8_t* state;        // 0 dead, 1 live, 2 scored (valid)
  uint8_t* F;
  uint8_t* nobs;
  int32_t* C;
  int32_t* cs;
  uint16_t* pos;         // [n][ADB_MAX_OBS] cycle positions of the observations
  uint32_t* fsel;        // [n][KS] library fragment index of the selected fragments, m/z order
  double* qtf;           // [n][nIcap * ADB_MAX_OBS]
  float* qmask;          // [n][ADB_MAX_OBS]
  int64_t* need;         // [n] block size of a slot in floats
  int64_t* tile_need;    // [n_tiles + 1] DP_W * the largest block of the tile (input of the scan)
  const int64_t* off;    // [n_tiles + 1] tile offset in floats (output of the scan)
  float* cube;
  uint32_t* status;
  uint8_t* rowflag;      // [n][KS] 1: fragment row with signal (candidate.py:319-329), written by dp_extract
  uint32_t* work;        // slot * KS + k of the fragment rows with signal, ascending (compaction of rowflag)
  int32_t* n_work;       // their number
  const double* wtab_p;  // [2][DP_WTAB_P_STRIDE] dp_wtab_p_entry(s, c)
};

struct DpLayout {
  int dfi, dfm, bp, nrm, isl, dpi, dpm, tmpl, tfp, med, wtab, red, ms1, fd, ff, fi, sc, total;
};

// per-fragment scalar slots (index k < F): doubles in fd, floats in ff, ints in fi; per-candidate scalars in sc
enum { FD_AREA = 0, FD_OFH = 1, FD_MZOBS = 2, FD_MERR = 3, FD_N = 4 };
enum { FF_OFI = 0, FF_COS = 1, FF_TCORR = 2, FF_FWHM = 3, FF_CORR = 4, FF_N = 5 };  // + rfw[nobs] behind them (legacy)
enum { FI_VALID = 0, FI_FMAP = 1, FI_N = 2 };                                          // + frame_peak[nobs] behind them
enum { SC_STI = 0, SC_OI = 8, SC_YM = 16, SC_YSTD = 24, SC_FEAT = 32, SC_FV = 32 + ADB_NUM_FEATURES, SC_N = 80 };

ADB_HD DpLayout dp_layout(int F, int nobs, int C, int nI, bool experimental, int n_ms1_pos) {
  DpLayout l;
  const int nFC = F * nobs * C, FC = F * C;
  int o = 0;
  l.dfi = o; o += nFC;
  l.dfm = o; o += nFC;
  l.bp = o; o += FC;
  l.nrm = o; o += experimental ? FC : nFC;             // normalised profiles / legacy: centred profiles of every observation
  l.isl = o; o += (experimental && nobs > 1) ? FC : 0;  // fragments_frame_profile.sum(axis=1)
  l.dpi = o; o += nI * C;
  l.dpm = o; o += nI * C;
  l.tmpl = o; o += nobs * C;
  l.tfp = o; o += nobs * C;
  l.med = o; o += C;
  o += o & 1;
  l.wtab = o; o += 4 * nobs * C;                        // double [nobs][2][C]
  l.fd = o; o += 2 * FD_N * F;                          // double [FD_N][F]
  l.ms1 = o; o += (n_ms1_pos > 1) ? 5 * nI * C + ((nI * C) & 1) : 0;  // double smz[nI C]; float ta, tm [nI C]; int cnt[nI C]
  l.red = o; o += experimental ? 0 : F * F;
  l.ff = o; o += (FF_N + (experimental ? 0 : nobs)) * F;
  l.fi = o; o += (FI_N + nobs) * F;
  l.sc = o; o += SC_N;
  l.total = (o + 3) & ~3;
  return l;
}

ADB_HD void dp_status_or(uint32_t* status, uint32_t bit) {
#if defined(__CUDA_ARCH__)
  atomicOr(status, bit);
#else
  *status |= bit;
#endif
}

ADB_HD FPtr dp_block(const DpParams& P, int64_t j) { return FPtr{P.cube + P.off[j / DP_W] + (j % DP_W)}; }

// thread index <-> (slot, row) for passes with `per` rows per slot: inside a tile the slot index runs fastest, so the lanes
// of a warp hold the same row of consecutive slots
ADB_HD void dp_decode(uint32_t t, uint32_t per, uint32_t& j, uint32_t& r) {
  const uint32_t tile = t / (DP_W * per), rem = t - tile * (DP_W * per);
  r = rem / DP_W;
  j = tile * DP_W + (rem - r * DP_W);
}
ADB_HD uint32_t dp_encode(uint32_t j, uint32_t r, uint32_t per) { return (j / DP_W) * (DP_W * per) + r * DP_W + (j % DP_W); }
ADB_HD int64_t dp_padded_slots(int64_t n) { return (n + DP_W - 1) / DP_W * DP_W; }

ADB_HD int64_t dp_candidate_of(const DpParams& P, int64_t j) { return P.order ? (int64_t)P.order[P.base + j] : P.base + j; }

// jitclasses/utils.py:15-20 mass window with a float32 tolerance
ADB_HD void dp_window(float mz, float tol, float& lo, float& hi) {
  const double d = (double)(tol * mz) / 1000000.0;
  lo = (float)((double)mz - d);
  hi = (float)((double)mz + d);
}

// distance weight of cell (scan s, cycle c) around the fixed precursor centre (2, 1), features_utils.py:9-26
#define DP_WTAB_P_STRIDE 4096  // = the largest cycle window dp_setup accepts
ADB_HD double dp_wtab_p_entry(int s, int c) {
  const double ds = (double)s - 2.0, dc = (double)c - 1.0;
  return exp(-0.1 * sqrt(ds * ds + dc * dc));
}

ADB_HD float dp_twice(float x) { return x + x; }  // sum over the two identical scan rows of a 3-D file

// ------------------------------------------------------------------------------------------------------------------
// dp_setup: candidate.py:151-232, quadrupole.py:80-115,261-301 (n_scans == 1), candidate.py:287-289
// ------------------------------------------------------------------------------------------------------------------
ADB_HD void dp_setup(const DpParams& P, int64_t j) {
  P.state[j] = 0;
  P.need[j] = 0;
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int64_t ci = dp_candidate_of(P, j);
  const int64_t L = raw.cycle_len;
  const int64_t p = P.cand.lib_row[ci];
  const int64_t frame_start = P.cand.frame_start[ci], frame_stop = P.cand.frame_stop[ci], frame_center = P.cand.frame_center[ci];
  const int64_t scan_start = P.cand.scan_start[ci], scan_stop = P.cand.scan_stop[ci], scan_center = P.cand.scan_center[ci];

  // candidate.py:151-163 isotope m/z
  const int nI = min(min(lib.n_isotopes, (int)min(cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const double charge = (double)lib.charge[p];
  const float pmz = lib.mz[p];
  float iso_mz[ADB_MAX_ISOTOPES];
  for (int i = 0; i < nI; i++) iso_mz[i] = (float)((double)i * ADB_ISOTOPE_DIFF / charge) + pmz;

  // candidate.py:181-192 fragments: cardinality filter, top-k by intensity, sort by m/z
  const int64_t fs = lib.frag_start_idx[p], fe = lib.frag_stop_idx[p];
  int n_all = (int)(fe - fs);
  if (n_all < 0) n_all = 0;
  if (n_all > DP_MAXF) { dp_status_or(P.status, ADB_STATUS_TOO_MANY_LIB_FRAGMENTS); return; }
  float t_int[DP_MAXF], t_mz[DP_MAXF];
  uint8_t t_src[DP_MAXF], t_sel[DP_MAXF];
  int m = 0;
  for (int q = 0; q < n_all; q++) {
    if (!cfg.exclude_shared_ions || lib.frag_cardinality[fs + q] <= 1) {
      t_src[m] = (uint8_t)q;
      t_int[m] = lib.frag_intensity[fs + q];
      t_mz[m] = lib.frag_mz[fs + q];
      m++;
    }
  }
  const int F = min(m, (int)min(cfg.top_k_fragments, (uint32_t)P.KS));
  if (m > ADB_NUMBA_SMALL_SORT) {  // np.argsort(intensity)[::-1][:top_k], numba's quicksort order among ties
    uint8_t ord[DP_MAXF];
    adb_argsort_numba(t_int, m, ord);
    for (int r = 0; r < F; r++) t_sel[r] = ord[m - 1 - r];
  } else
  for (int u = 0; u < m; u++) {  // descending-intensity position (stable argsort, reversed)
    const float v = t_int[u];
    int rank_asc = 0;
    for (int q = 0; q < m; q++) rank_asc += (t_int[q] < v) || (t_int[q] == v && q < u);
    const int r = m - 1 - rank_asc;
    if (r < F) t_sel[r] = (uint8_t)u;
  }
  if (F <= 3) return;
  if (F > ADB_NUMBA_SMALL_SORT) {  // np.argsort(mz) of the selected fragments, numba's quicksort order among ties
    float sel_mz[DP_MAXF];
    uint8_t ord[DP_MAXF];
    for (int r = 0; r < F; r++) sel_mz[r] = t_mz[t_sel[r]];
    adb_argsort_numba(sel_mz, F, ord);
    for (int r = 0; r < F; r++) P.fsel[j * P.KS + r] = (uint32_t)(fs + t_src[t_sel[ord[r]]]);
  } else
  for (int r = 0; r < F; r++) {  // stable m/z order among the selected
    const int u = t_sel[r];
    const float v = t_mz[u];
    int rank2 = 0;
    for (int q = 0; q < F; q++) { const float vq = t_mz[t_sel[q]]; rank2 += (vq < v) || (vq == v && q < r); }
    P.fsel[j * P.KS + rank2] = (uint32_t)(fs + t_src[u]);
  }

  float mn = iso_mz[0], mx = iso_mz[0];
  for (int i = 1; i < nI; i++) { mn = fminf(mn, iso_mz[i]); mx = fmaxf(mx, iso_mz[i]); }
  const float q0 = (float)((double)mn - 0.5), q1 = (float)((double)mx + 0.5);  // candidate.py:203-205
  int nobs = 0;  // alpharaw_jit.py:19-50
  for (int64_t t = 0; t < L; t++) {
    if (((double)q0 <= ADB_LD(raw.cycle + 2 * t + 1)) && ((double)q1 >= ADB_LD(raw.cycle + 2 * t))) {
      if (nobs < ADB_MAX_OBS) P.pos[j * ADB_MAX_OBS + nobs] = (uint16_t)t;
      nobs++;
    }
  }
  if (nobs > ADB_MAX_OBS) { dp_status_or(P.status, ADB_STATUS_TOO_MANY_OBS); return; }
  const int64_t cs = frame_start / L;
  const int64_t C64 = frame_stop / L - cs;
  if (C64 <= 0 || nobs == 0) return;  // candidate.py:230-232, :323-325
  // 3-D files: np.arange(scan_start, scan_stop) indexes cycle[0, c, s]; only s == 0 exists
  if (scan_stop - scan_start != 1 || scan_start != 0) return;
  if (scan_center < 0 || scan_center >= raw.n_mobility || frame_stop < 1 || frame_stop > raw.n_spectra ||
      frame_center < 0 || frame_center >= raw.n_spectra || frame_start < 0)
    return;
  if ((cs + C64) * L > raw.n_spectra) return;
  if (C64 > DP_WTAB_P_STRIDE) { dp_status_or(P.status, ADB_STATUS_SCRATCH_OVERFLOW); return; }
  const int C = (int)C64;

  // quadrupole.py:80-115,261-301 transfer function of every (isotope, observation); candidate.py:287-289 its isotope mean
  double* qtf = P.qtf + j * (int64_t)(P.nIcap * ADB_MAX_OBS);
  for (int i = 0; i < nI; i++)
    for (int o = 0; o < nobs; o++) {
      const int ps = P.pos[j * ADB_MAX_OBS + o];
      const double mu1 = ADB_LD(raw.cycle + 2 * ps + 0) + cfg.quad_delta_mu[0];
      const double mu2 = ADB_LD(raw.cycle + 2 * ps + 1) + cfg.quad_delta_mu[1];
      const double x = (double)iso_mz[i];
      const double a1 = (x - mu1) / cfg.quad_sigma[0], a2 = (x - mu2) / cfg.quad_sigma[1];
      qtf[i * nobs + o] = 1.0 / (1.0 + exp(-a1)) - 1.0 / (1.0 + exp(-a2));
    }
  for (int o = 0; o < nobs; o++) {
    double s = 0;
    for (int i = 0; i < nI; i++) s = s + qtf[i * nobs + o];
    P.qmask[j * ADB_MAX_OBS + o] = (float)(s / (double)nI);
  }
  P.F[j] = (uint8_t)F;
  P.nobs[j] = (uint8_t)nobs;
  P.C[j] = C;
  P.cs[j] = (int32_t)cs;
  P.need[j] = dp_layout(F, nobs, C, nI, cfg.experimental_xic != 0, raw.n_ms1_pos).total;
  P.state[j] = 1;
}

// ------------------------------------------------------------------------------------------------------------------
// dp_extract: get_dense(absolute_masses=True), alpharaw_jit.py:208-337, one XIC row per thread through the time-blocked m/z index
// ------------------------------------------------------------------------------------------------------------------
// All peaks of cycle position `ps` and cycles [cs, cs + C) whose m/z lies in [lo, hi] (minus the ones the previous, overlapping
// window already consumed: the reference's search cursor only moves forward), through the time-blocked m/z index: per time
// block one bucket-table read, a branch-free positioning over at most 8 peaks and a scan of the (about one) peaks inside the
// ppm window, one 128-bit load per peak.  All lanes of a warp reach the scan loop at its first in-window peak, so the cell
// updates of a warp run together.  Inside a cell the peaks arrive in ascending m/z = the order the reference meets them in
// the spectrum.  Cell recurrence: alpharaw_jit.py:300-333.  [c_lo, c_hi] grows to the range of cells that were written.
ADB_HD float dp_cell_mz(float num32, float den32) {
  // (float)((double(num) + 1e-36) / (double(den) + 1e-36)): for operands in [1e-18, 1e18] the additions are absorbed and
  // rounding the correctly rounded binary64 quotient of two binary32 numbers to binary32 equals the correctly rounded
  // binary32 quotient (53 >= 2 * 24 + 2, no double-rounding error; quotient inside the normal range)
  if (num32 >= 1e-18f && num32 <= 1e18f && den32 >= 1e-18f && den32 <= 1e18f) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(num32, den32);
#else
    return num32 / den32;
#endif
  }
  return (float)(((double)num32 + 1e-36) / ((double)den32 + 1e-36));
}

ADB_HD float4 dp_ld_peak(const float4* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

ADB_HD uint32_t dp_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  union { float f; uint32_t u; } v; v.f = x; return v.u;
#endif
}

ADB_HD void dp_extract_row(const DevRaw& raw, int ps, float lo, float hi, float prev_hi, int cs, int C, FPtr di, FPtr dm, int stride,
                           int& c_lo, int& c_hi) {
  // first wanted peak: m/z >= lo, or m/z > prev_hi when the previous window reaches into this one (prev_hi >= lo)
  const bool overlap = prev_hi >= lo;
  const float v = overlap ? prev_hi : lo;
  const int bk = adb_tb_bucket_of(raw, v);
  const int tb0 = cs / ADB_TB_CYCLES, tb1 = (cs + C - 1) / ADB_TB_CYCLES;
  for (int tb = tb0; tb <= tb1; tb++) {
    const uint32_t* tab = raw.tb_bucket + (size_t)(ps * raw.tb_ntb + tb) * (size_t)(raw.tb_nb + 1);
    uint32_t a = ADB_LD(tab + bk), b = ADB_LD(tab + bk + 1);
    const uint32_t seg1 = ADB_LD(tab + raw.tb_nb);
    while (b - a > 8u) {
      const uint32_t mid = (a + b) >> 1;
      const float m = ADB_LD(&raw.tb_pk[mid].x);
      if (overlap ? (m <= v) : (m < v)) a = mid + 1; else b = mid;
    }
    uint32_t i = a;
    for (uint32_t q = 0; q < 8u; q++) {  // the records behind the last peak are padding: reading past b is safe
      const float m = ADB_LD(&raw.tb_pk[a + q].x);
      i += (a + q < b) && (overlap ? (m <= v) : (m < v));
    }
    // scan: four records per round trip (the records behind the last peak are padding), processed in order
    bool done = false;
    while (!done && i < seg1) {
      float4 pk4[4];
      for (int q = 0; q < 4; q++) pk4[q] = dp_ld_peak(raw.tb_pk + i + q);
      for (int q = 0; q < 4; q++) {
        const float nm = pk4[q].x;
        if (i + (uint32_t)q >= seg1 || !(nm <= hi)) { done = true; break; }
        const uint32_t c = dp_f2u(pk4[q].z) - (uint32_t)cs;
        if (c < (uint32_t)C) {
          float ni = pk4[q].y;
          ni = ni * (((double)ni > 1e-26) ? 1.0f : 0.0f);
          const float acc_i = di[c * stride], acc_m = dm[c * stride];
          const float num32 = acc_m * acc_i + ni * nm;
          const float den32 = acc_i + ni;
          di[c * stride] = den32;
          dm[c * stride] = dp_cell_mz(num32, den32);
          c_lo = min(c_lo, (int)c);
          c_hi = max(c_hi, (int)c);
        }
      }
      i += 4u;
    }
  }
}

// r < KS: fragment row r (all observations); r >= KS: isotope row r - KS.  Fragment rows and the isotope rows of a file
// with one MS1 spectrum per cycle run through the same extraction code (one call site: the lanes of a warp stay together).
ADB_HD void dp_extract(const DpParams& P, int64_t j, int r) {
  if (!P.state[j]) {
    if (r < P.KS) P.rowflag[dp_encode((uint32_t)j, (uint32_t)r, (uint32_t)P.KS)] = 0;
    return;
  }
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int F = P.F[j], nobs = P.nobs[j], C = P.C[j], cs = P.cs[j];
  const int nI = min(min(lib.n_isotopes, (int)min(cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const DpLayout l = dp_layout(F, nobs, C, nI, cfg.experimental_xic != 0, raw.n_ms1_pos);
  const bool is_frag = r < P.KS;
  const int k = is_frag ? r : r - P.KS;  // fragment or isotope index
  if (is_frag ? (k >= F) : (k >= nI)) {
    if (is_frag) P.rowflag[dp_encode((uint32_t)j, (uint32_t)k, (uint32_t)P.KS)] = 0;
    return;
  }
  // m/z window of the row and of its predecessor (the reference's search cursor only moves forward)
  float mz_row, mz_prev = 0.f, tol;
  if (is_frag) {
    mz_row = ADB_LD(lib.frag_mz + P.fsel[j * P.KS + k]);
    if (k > 0) mz_prev = ADB_LD(lib.frag_mz + P.fsel[j * P.KS + k - 1]);
    tol = cfg.fragment_mz_tolerance;
  } else {  // candidate.py:151-163 isotope m/z
    const int64_t p = P.cand.lib_row[dp_candidate_of(P, j)];
    const double charge = (double)lib.charge[p];
    const float pmz = lib.mz[p];
    mz_row = (float)((double)k * ADB_ISOTOPE_DIFF / charge) + pmz;
    if (k > 0) mz_prev = (float)((double)(k - 1) * ADB_ISOTOPE_DIFF / charge) + pmz;
    tol = cfg.precursor_mz_tolerance;
  }
  float lo, hi, plo, prev_hi = -1.0f;
  dp_window(mz_row, tol, lo, hi);
  // the previous window matters only if it reaches into this one; a float pre-check with a wide margin (tolerance + 1e-5
  // relative) skips its fp64 division for the usual well-separated neighbours (prev_hi < lo then, whatever its exact value)
  if (k > 0 && mz_prev * (1.0f + 1e-6f * tol + 1e-5f) >= lo) dp_window(mz_prev, tol, plo, prev_hi);

  if (is_frag || raw.n_ms1_pos == 1) {
    const int stride = is_frag ? F : nI;
    const int n_rows = is_frag ? nobs : 1;
    // candidate.py:319-329 fragment mask: any signal over observations, scans and cycles.  Cells outside the written range
    // are +0 and leave the f32 sums unchanged, so the sums run over the written range only.
    float t_o = 0.f;
    for (int o = 0; o < n_rows; o++) {
      const FPtr di = blk + ((is_frag ? l.dfi + (o * C) * F : l.dpi) + k);
      const FPtr dm = blk + ((is_frag ? l.dfm + (o * C) * F : l.dpm) + k);
      for (int c = 0; c < C; c++) { di[c * stride] = 0.f; dm[c * stride] = 0.f; }
      int c_lo = C, c_hi = -1;
      dp_extract_row(raw, is_frag ? (int)P.pos[j * ADB_MAX_OBS + o] : raw.ms1_pos[0], lo, hi, prev_hi, cs, C, di, dm, stride, c_lo, c_hi);
      if (is_frag) {
        // candidate.py:290; an untouched cell stays +0 (0 * qmask is a zero of either sign: nothing downstream tells them apart)
        const float qm = P.qmask[j * ADB_MAX_OBS + o];
        float t_c = 0.f;
        for (int c = c_lo; c <= c_hi; c++) { const float x = di[c * stride] * qm; di[c * stride] = x; t_c = t_c + x; }
        t_o = t_o + dp_twice(t_c);
      } else {
        // candidate.py:239-269 MS1 observation collapse with one observation; unwritten cells: 0 / (0 + 1e-6) == 0
        for (int c = c_lo; c <= c_hi; c++) {
          const float am = dm[c * stride];
          di[c * stride] = 0.f + di[c * stride];
          // a zero numerator sends the fp64 division down its slow path; the quotient is the (signed) zero itself
          dm[c * stride] = (am == 0.f) ? am : (float)((0.0 + (double)am) / ((double)(am > 0.f ? 1 : 0) + 1e-6));
        }
      }
    }
    if (is_frag) {
      const bool fvalid = t_o > 0.f;
      dp_as_int(blk + l.fi)[FI_VALID * F + k] = fvalid ? 1 : 0;
      P.rowflag[dp_encode((uint32_t)j, (uint32_t)k, (uint32_t)P.KS)] = fvalid ? 1 : 0;
    }
    return;
  }
  // candidate.py:239-269 MS1 cube with the observation collapse (sum of intensities, mean of the non-zero m/z)
  const int i = k;
  const FPtr di = blk + (l.dpi + i);
  const FPtr dm = blk + (l.dpm + i);
  const DPtr smz = dp_as_double(blk + l.ms1) + i * C;
  const FPtr ta = blk + (l.ms1 + 2 * nI * C + i * C);
  const FPtr tm = blk + (l.ms1 + 3 * nI * C + i * C);
  const IPtr cnt = dp_as_int(blk + (l.ms1 + 4 * nI * C)) + i * C;
  for (int c = 0; c < C; c++) { di[c * nI] = 0.f; smz[c] = 0.0; cnt[c] = 0; }
  for (int q = 0; q < raw.n_ms1_pos; q++) {
    for (int c = 0; c < C; c++) { ta[c] = 0.f; tm[c] = 0.f; }
    int c_lo = C, c_hi = -1;
    dp_extract_row(raw, raw.ms1_pos[q], lo, hi, prev_hi, cs, C, ta, tm, 1, c_lo, c_hi);
    for (int c = 0; c < C; c++) {
      di[c * nI] = di[c * nI] + ta[c];
      smz[c] = smz[c] + (double)tm[c];
      cnt[c] += tm[c] > 0.f;
    }
  }
  for (int c = 0; c < C; c++) dm[c * nI] = (float)(smz[c] / ((double)cnt[c] + 1e-6));
}

// features_utils.py:9-26 weighted_center_mean of the intensity row r and the m/z row rm (element stride `st`) of one
// (fragment, observation) cell over the two identical scan rows, with the tabulated distance weights wt[2][C].
// A cell that is not > 0 adds +0.0, which leaves the (non-negative) running sums bit-identical.
template <typename WT>  // weights: DPtr (table inside the block) or const double* (the shared precursor table)
ADB_HD void dp_weighted_center_mean_pair(FPtr r, FPtr rm, int st, WT wt, int wst, int C, double& h, double& mz) {
  double v1 = 0, w1 = 0, v2 = 0, w2 = 0;
  bool any1 = false, any2 = false;
  for (int s = 0; s < 2; s++)
    for (int c = 0; c < C; c++) {
      const double wgt = wt[s * wst + c];
      const float a = r[c * st], b = rm[c * st];
      const bool pa = a > 0.f, pb = b > 0.f;
      any1 |= pa; any2 |= pb;
      v1 = v1 + (pa ? (double)a * wgt : 0.0); w1 = w1 + (pa ? wgt : 0.0);
      v2 = v2 + (pb ? (double)b * wgt : 0.0); w2 = w2 + (pb ? wgt : 0.0);
    }
  h = (any1 && w1 > 0) ? v1 / w1 : 0.0;
  mz = (any2 && w2 > 0) ? v2 / w2 : 0.0;
}

// ------------------------------------------------------------------------------------------------------------------
// dp_template: quadrupole.py:304-335, scoring/utils.py:46-53, fragment_features.py:20-49, features_utils.py:9-26 tables,
//              location_features.py:9-33, precursor_features.py:14-102
// ------------------------------------------------------------------------------------------------------------------
ADB_HD void dp_template(const DpParams& P, int64_t j) {
  if (!P.state[j]) return;
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int F = P.F[j], nobs = P.nobs[j], C = P.C[j];
  const int nI = min(min(lib.n_isotopes, (int)min(cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const DpLayout l = dp_layout(F, nobs, C, nI, cfg.experimental_xic != 0, raw.n_ms1_pos);
  const int64_t ci = dp_candidate_of(P, j);
  const int64_t p = P.cand.lib_row[ci];
  const int64_t frame_start = P.cand.frame_start[ci], frame_stop = P.cand.frame_stop[ci], frame_center = P.cand.frame_center[ci];
  const int64_t scan_start = P.cand.scan_start[ci], scan_stop = P.cand.scan_stop[ci], scan_center = P.cand.scan_center[ci];
  const double* qtf = P.qtf + j * (int64_t)(P.nIcap * ADB_MAX_OBS);
  const FPtr dpi = blk + l.dpi;
  const FPtr dpm = blk + l.dpm;
  const FPtr tmpl = blk + l.tmpl;
  const FPtr tfp = blk + l.tfp;
  const DPtr wtab = dp_as_double(blk + l.wtab);
  const FPtr sc = blk + l.sc;
  float iso_int[ADB_MAX_ISOTOPES], iso_mz[ADB_MAX_ISOTOPES];
  const double charge = (double)lib.charge[p];
  const float pmz = lib.mz[p];
  for (int i = 0; i < nI; i++) {
    iso_int[i] = lib.isotopes[p * lib.n_isotopes + i];
    iso_mz[i] = (float)((double)i * ADB_ISOTOPE_DIFF / charge) + pmz;
  }
  // quadrupole.py:304-324 template
DP_TPL_UNROLL
  for (int o = 0; o < nobs; o++)
DP_TPL_UNROLL
    for (int c = 0; c < C; c++) {
      double acc = 0;
      for (int i = 0; i < nI; i++) acc = acc + (double)(dpi[c * nI + i] * iso_int[i]) * qtf[i * nobs + o];
      tmpl[o * C + c] = (float)acc;
    }
  // quadrupole.py:327-335 observation importance
  float tot = 0.f;
DP_TPL_UNROLL
  for (int o = 0; o < nobs; o++) {
    float s = 0.f;
DP_TPL_UNROLL
    for (int c = 0; c < C; c++) s = s + tmpl[o * C + c];
    sc[SC_STI + o] = dp_twice(s);  // sum_template_intensity, also used by the cosine score
    tot = tot + sc[SC_STI + o];
  }
DP_TPL_UNROLL
  for (int o = 0; o < nobs; o++) sc[SC_OI + o] = (tot == 0.f) ? 1.0f / (float)nobs : sc[SC_STI + o] / tot;
  // candidate.py:341 template frame profile with or_envelope (scoring/utils.py:46-53) and its statistics
DP_TPL_UNROLL
  for (int o = 0; o < nobs; o++) {
    float ys = 0.f;
DP_TPL_UNROLL
    for (int c = 0; c < C; c++) {
      const float x = dp_twice(tmpl[o * C + c]);
      float res = x;
      if (c >= 1 && c < C - 1) {
        const float xl = dp_twice(tmpl[o * C + c - 1]), xr = dp_twice(tmpl[o * C + c + 1]);
        if (x < xl || x < xr) res = (float)((double)(xl + xr) / 2);
      }
      tfp[o * C + c] = res;
      ys = ys + res;
    }
    const float ym = ys / (float)C;
    float yss = 0.f;
DP_TPL_UNROLL
    for (int c = 0; c < C; c++) { const float yc = tfp[o * C + c] - ym; yss = yss + yc * yc; }
    sc[SC_YM + o] = ym;
    sc[SC_YSTD + o] = sqrtf(yss / (float)C);
  }
  // distance-weight tables for weighted_center_mean (features_utils.py:9-26); the precursor "centres" are the constants
  // (n_scans, n_observations) = (2, 1): that table is the same for every candidate (P.wtab_p, dp_wtab_p_entry)
DP_TPL_UNROLL
  for (int o = 0; o < nobs; o++) {  // fragment_features.py:20-49 centre of mass of the template
    const FPtr r = tmpl + o * C;
    double isum = 0, ssum = 0, fsum = 0;
    bool any = false;
DP_TPL_UNROLL
    for (int s = 0; s < 2; s++)
DP_TPL_UNROLL
      for (int c = 0; c < C; c++) { const float v = r[c]; if (v > 0.f) { any = true; isum = isum + (double)v; } }
    if (any)
DP_TPL_UNROLL
      for (int s = 0; s < 2; s++)
DP_TPL_UNROLL
        for (int c = 0; c < C; c++) {
          const float v = r[c];
          if (v > 0.f) { ssum = ssum + (double)s * (double)v; fsum = fsum + (double)c * (double)v; }
        }
    const double esc = (any && isum > 0) ? ssum / isum : 0.0;
    const double efc = (any && isum > 0) ? fsum / isum : 0.0;
DP_TPL_UNROLL
    for (int t = 0; t < 2 * C; t++) {
      const int s = t / C, c = t % C;
      const double ds = (double)s - esc, dc = (double)c - efc;
      wtab[o * 2 * C + t] = exp(-0.1 * sqrt(ds * ds + dc * dc));
    }
  }
  const FPtr fa = sc + SC_FEAT;
DP_TPL_UNROLL
  for (int t = 0; t < ADB_NUM_FEATURES; t++) fa[t] = 0.f;
  // features/location_features.py:9-33
  fa[0] = raw.mobility_values[scan_start] - raw.mobility_values[scan_stop - 1];
  fa[1] = raw.rt_values[frame_stop - 1] - raw.rt_values[frame_start];
  fa[2] = raw.rt_values[frame_center];
  fa[3] = raw.mobility_values[scan_center];
  fa[17] = (float)nobs;
  // features/precursor_features.py:14-102
  float spi[ADB_MAX_ISOTOPES], wspi[ADB_MAX_ISOTOPES];
  double H[ADB_MAX_ISOTOPES], MZo[ADB_MAX_ISOTOPES];
  for (int i = 0; i < nI; i++) {
    float tc = 0.f;
DP_TPL_UNROLL
    for (int c = 0; c < C; c++) tc = tc + dpi[c * nI + i];
    spi[i] = dp_twice(tc);
    float wsp = 0.f;
DP_TPL_UNROLL
    for (int o = 0; o < nobs; o++) wsp = wsp + spi[i] * sc[SC_OI + o];
    wspi[i] = wsp;
    dp_weighted_center_mean_pair(dpi + i, dpm + i, nI, P.wtab_p, DP_WTAB_P_STRIDE, C, H[i], MZo[i]);
  }
  int amax = 0;
  for (int i = 1; i < nI; i++) if (iso_int[i] > iso_int[amax]) amax = i;
  fa[4] = wspi[0];
  fa[5] = wspi[amax];
  float t6 = 0.f, t7 = 0.f;
  for (int i = 0; i < nI; i++) { t6 = t6 + wspi[i]; t7 = t7 + wspi[i] * iso_int[i]; }
  fa[6] = t6; fa[7] = t7;
  double wme = 0;
  for (int i = 0; i < nI; i++) if (MZo[i] > 0) {
    const double me = (MZo[i] - (double)iso_mz[i]) / (double)iso_mz[i] * 1e6;
    wme = wme + me * (double)iso_int[i];
  }
  fa[8] = (float)wme;
  fa[9] = (float)fabs(wme);
  fa[10] = (float)((double)iso_mz[0] + (wme * 1e-6) * (double)iso_mz[0]);
  fa[11] = (float)H[0];
  fa[12] = (float)H[amax];
  double t13 = 0, t14 = 0;
  for (int i = 0; i < nI; i++) { t13 = t13 + H[i]; t14 = t14 + H[i] * (double)iso_int[i]; }
  fa[13] = (float)t13; fa[14] = (float)t14;
  const double hbar = t13 / (double)nI;
  float sx = 0.f, sy = 0.f;
  for (int i = 0; i < nI; i++) { sx = sx + iso_int[i]; sy = sy + spi[i]; }
  const double xbar = (double)sx / (double)nI, ybar = (double)sy / (double)nI;
  double num = 0, sxx = 0, syy = 0, num2 = 0, shh = 0;
  for (int i = 0; i < nI; i++) {
    const double a = (double)iso_int[i] - xbar, b = (double)spi[i] - ybar, h = H[i] - hbar;
    num = num + a * b; sxx = sxx + a * a; syy = syy + b * b;
    num2 = num2 + a * h; shh = shh + h * h;
  }
  fa[15] = (float)(num / (sqrt(sxx * syy) + 1e-12));
  fa[16] = (float)(num2 / (sqrt(sxx * shh) + 1e-12));
}

ADB_HD int dp_best_obs(FPtr sc, int nobs) {
  int best = 0;
  for (int o = 1; o < nobs; o++) if (sc[SC_OI + o] > sc[SC_OI + best]) best = o;
  return best;
}

// ------------------------------------------------------------------------------------------------------------------
// dp_fragment: candidate.py:319-329 mask, fragment_features.py:198-336, profile_features.py (per-fragment parts)
// ------------------------------------------------------------------------------------------------------------------
ADB_HD void dp_fragment(const DpParams& P, int64_t j, int k) {
  if (!P.state[j]) return;
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int F = P.F[j];
  if (k >= F) return;
  const int nobs = P.nobs[j], C = P.C[j];
  const int nI = min(min(lib.n_isotopes, (int)min(cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const bool experimental = cfg.experimental_xic != 0;
  const DpLayout l = dp_layout(F, nobs, C, nI, experimental, raw.n_ms1_pos);
  const FPtr d = blk + (l.dfi + k);   // d[(o * C + c) * F]
  const FPtr dmz = blk + (l.dfm + k);
  const FPtr b = blk + (l.bp + k);    // b[c * F]
  const FPtr sc = blk + l.sc;
  const IPtr fi = dp_as_int(blk + l.fi);
  const FPtr ff = blk + l.ff;
  const DPtr fd = dp_as_double(blk + l.fd);
  const int64_t L = raw.cycle_len;
  const int64_t ci = dp_candidate_of(P, j);
  const int64_t frame_start = P.cand.frame_start[ci], frame_stop = P.cand.frame_stop[ci];

  if (!fi[FI_VALID * F + k]) return;  // candidate.py:319-329 fragment mask, set by dp_extract
  float tcs[ADB_MAX_OBS];  // per-observation intensity sums over the cycles
  for (int o = 0; o < nobs; o++) {
    float t_c = 0.f;
    for (int c = 0; c < C; c++) t_c = t_c + d[(o * C + c) * F];
    tcs[o] = t_c;
  }

  const int best_obs = dp_best_obs(sc, nobs);
  const bool quant_all = cfg.quant_all != 0;
  int64_t qw = (int64_t)cfg.quant_window;
  if ((C / 2) - 1 < qw) qw = (C / 2) - 1;
  const int center = C / 2;
  int w0 = center - (int)qw, w1 = center + (int)qw + 1;
  if (qw < 0) { w0 = 0; w1 = 0; }
  if (w1 > C) w1 = C;
  if (w0 < 0) w0 = 0;
  const int wn = max(w1 - w0, 0);
  // fragment_features.py:225-250 best profile
  if (quant_all) {
    for (int c = 0; c < C; c++) { float t = 0.f; for (int o = 0; o < nobs; o++) t = t + dp_twice(d[(o * C + c) * F]); b[c * F] = t; }
  } else {
    for (int c = 0; c < C; c++) b[c * F] = dp_twice(d[(best_obs * C + c) * F]);
  }
  // center_envelope_1d, fragment_features.py:71-159
  if (C % 2 == 0) {
    const int cr = C / 2, cl = cr - 1;
    if (cl >= 0) {
      float left = b[cl * F], right = b[cr * F];
      for (int i = 1; i <= cl; i++) {
        b[(cl - i) * F] = fminf(left, b[(cl - i) * F]);
        left = (float)((double)(b[(cl - i) * F] + b[(cl - i + 1) * F]) * 0.5);
        b[(cr + i) * F] = fminf(right, b[(cr + i) * F]);
        right = (float)((double)(b[(cr + i) * F] + b[(cr + i - 1) * F]) * 0.5);
      }
    }
  } else if (C >= 3) {
    const int cc = C / 2;
    float left = (float)((double)(b[(cc - 1) * F] + b[cc * F]) * 0.5);
    float right = (float)((double)(b[(cc + 1) * F] + b[cc * F]) * 0.5);
    for (int i = 1; i <= cc; i++) {
      b[(cc - i) * F] = fminf(left, b[(cc - i) * F]);
      left = (float)((double)(b[(cc - i) * F] + b[(cc - i + 1) * F]) * 0.5);
      b[(cc + i) * F] = fminf(right, b[(cc + i) * F]);
      right = (float)((double)(b[(cc + i) * F] + b[(cc + i - 1) * F]) * 0.5);
    }
  }
  // trapezoid area over the quant window, fragment_features.py:253-273
  double area = 0;
  for (int t = 0; t + 1 < wn; t++) {
    const float drt = ADB_LD(raw.rt_values + frame_start + (int64_t)(w0 + t + 1) * L) - ADB_LD(raw.rt_values + frame_start + (int64_t)(w0 + t) * L);
    const float sum2 = b[(w0 + t + 1) * F] + b[(w0 + t) * F];
    area = area + (double)(sum2 * drt) * 0.5;
  }
  fd[FD_AREA * F + k] = area * (double)qw;
  float ofi = 0.f;
  for (int u = 0; u < wn; u++) ofi = ofi + b[(w0 + u) * F];
  ff[FF_OFI * F + k] = ofi;

  // per-observation: summed intensity (cosine score) and the observation mask.  A weighted-centre height is > 0 exactly
  // when the (fragment, observation) row has signal (all weights are positive), i.e. when its f32 sum is > 0.
  float fn2 = 0.f, dot = 0.f, wsum = 0.f, tn2 = 0.f;
  unsigned obs_mask = 0u;
  for (int o = 0; o < nobs; o++) {
    const float v = dp_twice(tcs[o]);
    fn2 = fn2 + v * v;
    dot = dot + v * sc[SC_STI + o];
    tn2 = tn2 + sc[SC_STI + o] * sc[SC_STI + o];
    const bool mm = tcs[o] > 0.f;
    if (mm) obs_mask |= 1u << o;
    wsum = wsum + (mm ? sc[SC_OI + o] : 0.0f);  // fragment_features.py:318-326
  }
  {  // cosine_similarity_a1, features_utils.py:40-47
    const double div = (double)(sqrtf(fn2) * sqrtf(tn2)) + 0.0001;
    ff[FF_COS * F + k] = (float)((double)dot / div);
  }
  // fragment_features.py:312-336 observation-weighted means of the weighted-centre height and m/z
  double wtot = 0;
  int cnt = 0;
  for (int o = 0; o < nobs; o++) {
    const double wv = (double)(((obs_mask >> o) & 1u) ? sc[SC_OI + o] : 0.0f) / ((double)wsum + 1e-20);
    if (wv > 0) { wtot = wtot + wv; cnt++; }
  }
  double a = 0, bsum = 0;
  if (cnt > 0) {
    const DPtr wtab = dp_as_double(blk + l.wtab);
    for (int o = 0; o < nobs; o++) {
      const double wv = (double)(((obs_mask >> o) & 1u) ? sc[SC_OI + o] : 0.0f) / ((double)wsum + 1e-20);
      if (wv > 0) {
        double h_o, mz_o;
        dp_weighted_center_mean_pair(d + (o * C) * F, dmz + (o * C) * F, F, wtab + o * 2 * C, C, C, h_o, mz_o);
        const double lw = wv / wtot;
        a = a + mz_o * lw;
        bsum = bsum + h_o * lw;
      }
    }
  }
  fd[FD_OFH * F + k] = bsum;
  fd[FD_MZOBS * F + k] = a;
  const double mzf = (double)ADB_LD(lib.frag_mz + P.fsel[j * P.KS + k]);
  fd[FD_MERR * F + k] = (a - mzf) / mzf * 1e6;

  // ---- profile_features.py:18-206, per-fragment parts.  fragments_frame_profile accessor: the best observation's rows were
  // enveloped in place when quant_all is off (fragment_features.py:248-250, view semantics)
#define DP_FFP(o, c) ((!quant_all && (o) == best_obs) ? b[(c) * F] : dp_twice(d[((o) * C + (c)) * F]))
  if (experimental) {
    // fragments_frame_profile.sum(axis=1)
    const FPtr islr = blk + (l.isl + k);
    if (nobs > 1)
      for (int c = 0; c < C; c++) { float t = 0.f; for (int o = 0; o < nobs; o++) t = t + DP_FFP(o, c); islr[c * F] = t; }
#define DP_ISL(c) (nobs == 1 ? (quant_all ? dp_twice(d[(c) * F]) : b[(c) * F]) : islr[(c) * F])
    int a0 = center - 1, a1 = center + 2;  // scoring_utils.py:100-110 python slice semantics
    if (a0 < 0) { a0 += C; if (a0 < 0) a0 = 0; }
    if (a1 > C) a1 = C;
    const int wnn = max(a1 - a0, 0);
    float t = 0.f;
    for (int c = a0; c < a1; c++) t = t + DP_ISL(c);
    const double cint = (double)t / (double)wnn;
    const FPtr nr = blk + (l.nrm + k);
    for (int c = 0; c < C; c++) {
      const float xv = DP_ISL(c);  // zero cells skip the fp64 division (0 / cint is the same signed zero)
      nr[c * F] = (cint > 0) ? ((xv == 0.f) ? xv : (float)((double)xv / cint)) : 0.f;
    }
#undef DP_ISL
  } else {
    // legacy (scoring/utils.py:513-571): centred profile and its std for every observation
    const FPtr cen = blk + (l.nrm + k);  // cen[(o * C + c) * F]
    for (int o = 0; o < nobs; o++) {
      float s = 0.f;
      for (int c = 0; c < C; c++) s = s + DP_FFP(o, c);
      const float mean = s / (float)C;
      float ss = 0.f;
      for (int c = 0; c < C; c++) { const float cv = DP_FFP(o, c) - mean; cen[(o * C + c) * F] = cv; ss = ss + cv * cv; }
      ff[(FF_N + o) * F + k] = sqrtf(ss / (float)C);
    }
  }
  // template correlation (profile_features.py:82-85), cycle fwhm (:142-144), frame peak (:193-204)
  const FPtr tfp = blk + l.tfp;
  const float rt_width = raw.rt_values[frame_stop - 1] - raw.rt_values[frame_start];
  float tcorr = 0.f, fwhm = 0.f;
  for (int o = 0; o < nobs; o++) {
    const FPtr y = tfp + o * C;
    const float ym = sc[SC_YM + o], ystd = sc[SC_YSTD + o];
    float xs = 0.f, mxv = 0.f;
    int am = 0;
    for (int c = 0; c < C; c++) {
      const float x = DP_FFP(o, c);
      xs = xs + x;
      if (c == 0 || x > mxv) { am = c; mxv = x; }
    }
    const float xm = xs / (float)C;
    float xss = 0.f, dotp = 0.f;
    int na = 0;
    const double half = (double)mxv / 2;
    for (int c = 0; c < C; c++) {
      const float x = DP_FFP(o, c);
      const float xc = x - xm;
      xss = xss + xc * xc;
      dotp = dotp + xc * (y[c] - ym);
      na += (double)x > half;
    }
    const float xstd = sqrtf(xss / (float)C);
    const float cov = dotp / (float)C;
    const float ct = (float)((double)cov / ((double)(xstd * ystd) + 1e-12));
    const float fw = (float)(((double)na / (double)C) * (double)rt_width);
    tcorr = tcorr + ct * sc[SC_OI + o];
    fwhm = fwhm + fw * sc[SC_OI + o];
    fi[(FI_N + o) * F + k] = am;
  }
#undef DP_FFP
  ff[FF_TCORR * F + k] = tcorr;
  ff[FF_FWHM * F + k] = fwhm;
}

// masked fragment list of a slot: fmap[w] = k of the w-th fragment with signal; returns Fv
ADB_HD int dp_fragment_mask(IPtr fi, int F, uint8_t* fmap) {
  int Fv = 0;
  for (int k = 0; k < F; k++) if (fi[FI_VALID * F + k]) fmap[Fv++] = (uint8_t)k;
  return Fv;
}

// ------------------------------------------------------------------------------------------------------------------
// dp_median: scoring_utils.py:127-152 median of the normalised profiles over the fragments, DP_MED_LANES threads per slot
// ------------------------------------------------------------------------------------------------------------------
ADB_HD void dp_median(const DpParams& P, int64_t j, int lane) {
  if (!P.state[j] || !P.cfg.experimental_xic) return;
  const int F = P.F[j], nobs = P.nobs[j], C = P.C[j];
  const int nI = min(min(P.lib.n_isotopes, (int)min(P.cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const DpLayout l = dp_layout(F, nobs, C, nI, true, P.raw.n_ms1_pos);
  uint8_t fmap[DP_MAXF];
  const int Fv = dp_fragment_mask(dp_as_int(blk + l.fi), F, fmap);
  if (Fv < 2) return;
  const FPtr nrm = blk + l.nrm;
  const FPtr med = blk + l.med;
  for (int c = lane; c < C; c += DP_MED_LANES) {
    float vlo = 0.f, vhi = 0.f;
    for (int w = 0; w < Fv; w++) {
      const float v = nrm[c * F + fmap[w]];
      int rk = 0;
      for (int u = 0; u < Fv; u++) { const float vu = nrm[c * F + fmap[u]]; rk += (vu < v) || (vu == v && u < w); }
      if (rk == (Fv - 1) / 2) vlo = v;
      if (rk == Fv / 2) vhi = v;
    }
    med[c] = (Fv & 1) ? vhi : (float)((double)(vlo + vhi) / 2);
  }
}

// fragment_container.py:119-120 renormalised library intensities of the masked fragments
ADB_HD void dp_masked_intensity(const DpParams& P, int64_t j, const uint8_t* fmap, int Fv, float* fint) {
  float isum = 0.f;
  for (int w = 0; w < Fv; w++) { fint[w] = ADB_LD(P.lib.frag_intensity + P.fsel[j * P.KS + fmap[w]]); isum = isum + fint[w]; }
  for (int w = 0; w < Fv; w++) fint[w] = fint[w] / isum;
}

// ------------------------------------------------------------------------------------------------------------------
// dp_corr: correlation_coefficient(median_profile, intensity_slice) (scoring_utils.py:20-76) or the legacy observation-
//          weighted F x F correlation matrix row (scoring/utils.py:513-571)
// ------------------------------------------------------------------------------------------------------------------
ADB_HD void dp_corr(const DpParams& P, int64_t j, int k) {
  if (!P.state[j]) return;
  const adb_scoring_config& cfg = P.cfg;
  const int F = P.F[j];
  if (k >= F) return;
  const int nobs = P.nobs[j], C = P.C[j];
  const int nI = min(min(P.lib.n_isotopes, (int)min(cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const bool experimental = cfg.experimental_xic != 0;
  const DpLayout l = dp_layout(F, nobs, C, nI, experimental, P.raw.n_ms1_pos);
  const IPtr fi = dp_as_int(blk + l.fi);
  if (!fi[FI_VALID * F + k]) return;
  const FPtr ff = blk + l.ff;
  const FPtr sc = blk + l.sc;
  if (experimental) {
    const FPtr med = blk + l.med;
    const bool quant_all = cfg.quant_all != 0;
    const FPtr d = blk + (l.dfi + k);
    const FPtr b = blk + (l.bp + k);
    const FPtr islr = blk + (l.isl + k);
#define DP_ISL(c) (nobs == 1 ? (quant_all ? dp_twice(d[(c) * F]) : b[(c) * F]) : islr[(c) * F])
    float sx = 0.f;
    for (int c = 0; c < C; c++) sx = sx + med[c];
    const double mxv = (double)sx / (double)C;
    double varx = 0;
    for (int c = 0; c < C; c++) { const double dd = (double)med[c] - mxv; varx = varx + dd * dd; }
    varx /= (double)C;
    float sy = 0.f;
    for (int c = 0; c < C; c++) sy = sy + DP_ISL(c);
    const float myv = (float)((double)sy / (double)C);
    double cov = 0;
    float vy32 = 0.f;
    for (int c = 0; c < C; c++) {
      const float ym = DP_ISL(c) - myv;
      cov = cov + ((double)med[c] - mxv) * (double)ym;
      vy32 = vy32 + ym * ym;
    }
#undef DP_ISL
    cov /= (double)C;
    const double vxy = varx * ((double)vy32 / (double)C);
    ff[FF_CORR * F + k] = (vxy == 0) ? 0.f : (float)(cov / sqrt(vxy));
    return;
  }
  uint8_t fmap[DP_MAXF];
  float fint[DP_MAXF];
  const int Fv = dp_fragment_mask(fi, F, fmap);
  if (Fv < 2) return;
  dp_masked_intensity(P, j, fmap, Fv, fint);
  const FPtr cen = blk + l.nrm;
  const FPtr red = blk + (l.red + k * F);  // row of fragment k, columns by fragment index
  for (int w = 0; w < Fv; w++) red[fmap[w]] = 0.f;
  for (int o = 0; o < nobs; o++) {
    const float rfa = ff[(FF_N + o) * F + k];
    for (int w = 0; w < Fv; w++) {
      const int kb = fmap[w];
      float dot = 0.f;
      for (int c = 0; c < C; c++) dot = dot + cen[(o * C + c) * F + k] * cen[(o * C + c) * F + kb];
      const float cov = dot / (float)C;
      const float smx = rfa * ff[(FF_N + o) * F + kb];
      const float corr = (float)((double)cov / ((double)smx + 1e-12));
      red[kb] = red[kb] + corr * sc[SC_OI + o];
    }
  }
  float t = 0.f;
  for (int g = 0; g < Fv; g++) t = t + red[fmap[g]] * fint[g];
  ff[FF_CORR * F + k] = t;
}

// np.corrcoef(x, y)[0, 1] as numba evaluates it (cov with 1/(n-1), divide by both std)
ADB_HD double dp_corrcoef01(const double* x, const float* yf, int n) {
  double mx = 0, my = 0;
  for (int i = 0; i < n; i++) { mx = mx + x[i]; my = my + (double)yf[i]; }
  mx /= n; my /= n;
  double cxx = 0, cyy = 0, cxy = 0;
  for (int i = 0; i < n; i++) {
    const double a = x[i] - mx, b = (double)yf[i] - my;
    cxx = cxx + a * a; cyy = cyy + b * b; cxy = cxy + a * b;
  }
  const double fact = 1.0 / (double)(n - 1);
  cxx *= fact; cyy *= fact; cxy *= fact;
  return (cxy / sqrt(cyy)) / sqrt(cxx);
}

// ------------------------------------------------------------------------------------------------------------------
// dp_aggregate: fragment_features.py:337-427, profile_features.py aggregates, candidate.py:362,475-481
// ------------------------------------------------------------------------------------------------------------------
// `stage`: when given, the finished feature row goes there instead of straight into the output table (the kernel then
// writes the rows of a warp cooperatively, full sectors instead of 46 scattered 4-byte stores per thread)
ADB_HD void dp_aggregate(const DpParams& P, int64_t j, float* stage = nullptr) {
  if (!P.state[j]) return;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int F = P.F[j], nobs = P.nobs[j], C = P.C[j];
  const int nI = min(min(lib.n_isotopes, (int)min(cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const bool experimental = cfg.experimental_xic != 0;
  const DpLayout l = dp_layout(F, nobs, C, nI, experimental, P.raw.n_ms1_pos);
  const IPtr fi = dp_as_int(blk + l.fi);
  const FPtr ff = blk + l.ff;
  const DPtr fd = dp_as_double(blk + l.fd);
  const FPtr sc = blk + l.sc;
  uint8_t fmap[DP_MAXF], sorted_idx[DP_MAXF];
  float fint[DP_MAXF], fin[DP_MAXF];
  const int Fv = dp_fragment_mask(fi, F, fmap);
  if (Fv < 2) return;
  dp_masked_intensity(P, j, fmap, Fv, fint);
  {  // fragment_features.py:218 fragment_intensity_norm
    float t = 0.f;
    for (int w = 0; w < Fv; w++) t = t + fint[w];
    for (int w = 0; w < Fv; w++) fin[w] = fint[w] / t;
  }
  if (Fv > ADB_NUMBA_SMALL_SORT) {  // np.argsort(fragments.intensity)[::-1], numba's quicksort order among ties
    uint8_t ord[DP_MAXF];
    adb_argsort_numba(fint, Fv, ord);
    for (int r = 0; r < Fv; r++) sorted_idx[r] = ord[Fv - 1 - r];
  } else
  for (int w = 0; w < Fv; w++) {  // np.argsort(fragments.intensity)[::-1]
    const float v = fint[w];
    int rank_asc = 0;
    for (int q = 0; q < Fv; q++) rank_asc += (fint[q] < v) || (fint[q] == v && q < w);
    sorted_idx[Fv - 1 - rank_asc] = (uint8_t)w;
  }
  // masked views of the per-fragment scalars
  double area_norm[DP_MAXF], ofh_mean[DP_MAXF], mass_error[DP_MAXF];
  float ofi[DP_MAXF], corr_list[DP_MAXF];
  uint8_t ftype[DP_MAXF], fpos[DP_MAXF];
  for (int w = 0; w < Fv; w++) {
    const int k = fmap[w];
    area_norm[w] = fd[FD_AREA * F + k];
    ofh_mean[w] = fd[FD_OFH * F + k];
    mass_error[w] = fd[FD_MERR * F + k];
    ofi[w] = ff[FF_OFI * F + k];
    corr_list[w] = ff[FF_CORR * F + k];
    const uint32_t g = P.fsel[j * P.KS + k];
    ftype[w] = ADB_LD(lib.frag_type + g);
    fpos[w] = ADB_LD(lib.frag_position + g);
  }
  const FPtr fa = sc + SC_FEAT;
  fa[28] = (float)((double)Fv / (double)F);  // candidate.py:362
  bool anyh = false;  // a weighted-centre height > 0 exists: some (fragment, observation) row has signal; true for every masked fragment
  anyh = Fv > 0;
  double sum_ofh = 0;
  for (int w = 0; w < Fv; w++) sum_ofh = sum_ofh + ofh_mean[w];
  if (anyh) fa[18] = (float)dp_corrcoef01(area_norm, fin, Fv);
  if (sum_ofh > 0.0) fa[19] = (float)dp_corrcoef01(ofh_mean, fin, Fv);
  int n20 = 0, n21 = 0;
  float s22 = 0.f, s23 = 0.f, cacc = 0.f;
  for (int w = 0; w < Fv; w++) if (ofi[w] > 0.f) { n20++; s22 = s22 + fin[w]; cacc = cacc + ff[FF_COS * F + fmap[w]]; }
  for (int w = 0; w < Fv; w++) if (ofh_mean[w] > 0.0) { n21++; s23 = s23 + fin[w]; }
  fa[20] = (float)((double)n20 / (double)Fv);
  fa[21] = (float)((double)n21 / (double)Fv);
  fa[22] = s22; fa[23] = s23;
  if (n20 > 0) fa[24] = (float)((double)cacc / (double)n20);
  float sb = 0.f, sy = 0.f;
  int nb = 0, ny = 0, min_y = 255, max_b = 0;
  for (int w = 0; w < Fv; w++) {
    const int ty = ftype[w], po = fpos[w];
    if (ty == 98) { sb = sb + ofi[w]; nb++; max_b = max(max_b, po); }
    if (ty == 121) { sy = sy + ofi[w]; ny++; min_y = min(min_y, po); }
  }
  fa[25] = nb > 0 ? (float)log((double)sb + 1.0) : 0.f;
  fa[26] = ny > 0 ? (float)log((double)sy + 1.0) : 0.f;
  fa[27] = fa[25] - fa[26];
  const int n3 = min(Fv, 3);
  double t41 = 0, t42 = 0;
  for (int r = 0; r < n3; r++) t41 = t41 + mass_error[sorted_idx[r]];
  for (int w = 0; w < Fv; w++) t42 = t42 + mass_error[w];
  fa[41] = (float)(t41 / (double)n3);
  fa[42] = (float)(t42 / (double)Fv);
  if (nb > 0 && ny > 0) {
    int n_ov = 0;
    double sa = 0, se = 0;
    for (int w = 0; w < Fv; w++) {
      const int ty = ftype[w], po = fpos[w];
      const bool ov = (ty == 121 && po < max_b) || (ty == 98 && po > min_y);
      if (ov) { n_ov++; sa = sa + area_norm[w]; se = se + mass_error[w]; }
    }
    fa[43] = (float)n_ov;
    if (n_ov > 0) { fa[44] = (float)(sa / (double)n_ov); fa[45] = (float)(se / (double)n_ov); }
    else { fa[44] = 0.f; fa[45] = 15.f; }
  }
  // profile_features.py:18-113
  if (experimental) {
    float t = 0.f;
    for (int r = 0; r < n3; r++) t = t + corr_list[sorted_idx[r]];
    fa[32] = (float)((double)t / (double)n3);
  } else {
    const FPtr red = blk + l.red;
    float t = 0.f;
    for (int a = 0; a < n3; a++)
      for (int b = 0; b < n3; b++) t = t + red[fmap[sorted_idx[a]] * F + fmap[sorted_idx[b]]];
    fa[32] = (float)((double)t / (double)(n3 * n3));
  }
  // median frame peak per observation (profile_features.py:193-204)
  double acc40 = 0.0;
  for (int o = 0; o < nobs; o++) {
    int vlo = 0, vhi = 0;
    for (int w = 0; w < Fv; w++) {
      const int v = fi[(FI_N + o) * F + fmap[w]];
      int rk = 0;
      for (int u = 0; u < Fv; u++) { const int vu = fi[(FI_N + o) * F + fmap[u]]; rk += (vu < v) || (vu == v && u < w); }
      if (rk == (Fv - 1) / 2) vlo = v;
      if (rk == Fv / 2) vhi = v;
    }
    const float medp = (float)((Fv & 1) ? (double)vhi : ((double)vlo + (double)vhi) / 2);
    const double delta = (double)medp - floor((double)C / 2);
    acc40 = (o == 0 ? 0.0 : acc40) + delta * (double)sc[SC_OI + o];
  }
  float t31 = 0.f, t33 = 0.f, t38 = 0.f;
  for (int w = 0; w < Fv; w++) {
    t31 = t31 + corr_list[w];
    t33 = t33 + ff[FF_TCORR * F + fmap[w]] * fint[w];
    t38 = t38 + ff[FF_FWHM * F + fmap[w]] * fint[w];
  }
  fa[31] = (float)((double)t31 / (double)Fv);
  fa[33] = t33;
  fa[38] = t38;
  fa[40] = (float)acc40;
  {  // profile_features.py:94-113 (the type mask indexes the sorted-index array by position)
    int nb2 = 0, ny2 = 0;
    float sb2 = 0.f, sy2 = 0.f;
    for (int r = 0; r < Fv; r++) {
      const int ty = ftype[r];
      if (ty == 98) { if (nb2 < 3) sb2 = sb2 + corr_list[sorted_idx[r]]; nb2++; }
      if (ty == 121) { if (ny2 < 3) sy2 = sy2 + corr_list[sorted_idx[r]]; ny2++; }
    }
    if (nb2 > 0) { fa[34] = (float)((double)sb2 / (double)min(nb2, 3)); fa[35] = (float)nb2; }
    if (ny2 > 0) { fa[36] = (float)((double)sy2 / (double)min(ny2, 3)); fa[37] = (float)ny2; }
  }
  // candidate.py:475-481
  const int64_t ci = dp_candidate_of(P, j);
  if (stage) {
    for (int t = 0; t < ADB_NUM_FEATURES; t++) stage[t] = fa[t];
  } else {
    for (int t = 0; t < ADB_NUM_FEATURES; t++) P.out.features[(size_t)ci * ADB_NUM_FEATURES + t] = fa[t];
  }
  P.out.valid[ci] = 1;
  for (int w = 0; w < Fv; w++) fi[FI_FMAP * F + w] = fmap[w];
  sc[SC_FV] = (float)Fv;
  P.state[j] = 2;
}

// ------------------------------------------------------------------------------------------------------------------
// dp_write: candidate.py:403-442,475-481 per-fragment output table, thread (slot, w) <-> masked fragment w
// ------------------------------------------------------------------------------------------------------------------
ADB_HD void dp_write(const DpParams& P, int64_t j, int w) {
  if (P.state[j] != 2 || !P.cfg.collect_fragments) return;
  const DevLib& lib = P.lib;
  const int F = P.F[j], nobs = P.nobs[j], C = P.C[j];
  const int nI = min(min(lib.n_isotopes, (int)min(P.cfg.top_k_isotopes, 1000u)), ADB_MAX_ISOTOPES);
  const FPtr blk = dp_block(P, j);
  const DpLayout l = dp_layout(F, nobs, C, nI, P.cfg.experimental_xic != 0, P.raw.n_ms1_pos);
  const int Fv = (int)blk[l.sc + SC_FV];
  if (w >= Fv || w >= P.out_k) return;
  const IPtr fi = dp_as_int(blk + l.fi);
  const FPtr ff = blk + l.ff;
  const DPtr fd = dp_as_double(blk + l.fd);
  const int k = fi[FI_FMAP * F + w];
  const uint32_t g = P.fsel[j * P.KS + k];
  const size_t o = (size_t)dp_candidate_of(P, j) * (size_t)P.out_k + (size_t)w;
  P.out.fragment_mz_library[o] = ADB_LD(lib.frag_mz_library + g);
  P.out.fragment_mz[o] = ADB_LD(lib.frag_mz + g);
  P.out.fragment_mz_observed[o] = (float)fd[FD_MZOBS * F + k];
  P.out.fragment_height[o] = (float)fd[FD_OFH * F + k];
  P.out.fragment_intensity[o] = (float)fd[FD_AREA * F + k];
  P.out.fragment_mass_error[o] = (float)fd[FD_MERR * F + k];
  P.out.fragment_correlation[o] = ff[FF_CORR * F + k];
  P.out.fragment_position[o] = ADB_LD(lib.frag_position + g);
  P.out.fragment_number[o] = ADB_LD(lib.frag_number + g);
  P.out.fragment_type[o] = ADB_LD(lib.frag_type + g);
  P.out.fragment_charge[o] = ADB_LD(lib.frag_charge + g);
  P.out.fragment_loss_type[o] = ADB_LD(lib.frag_loss_type + g);
}

/*
 * adb_oracle.c — CPU restatement of the reference's precursor-candidate hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product (alphadia_b200/)
 * never links, imports or calls it and fails loudly when its CUDA extension is missing.
 *
 * Parity status: PINNED against the live reference (numba path of /root/reference run in the
 * build container through oracle/refshim.py) by the golden vectors in tests/golden (npz files)
 * (tests/test_oracle_golden.py).  The one arithmetic dependency that is not available
 * (rocket-fft 0.2.5 / pocketfft behind alphadia/search/selection/fft.py) is replaced on BOTH
 * sides by its mathematical definition: direct circular same-size convolution with fp64 FMA
 * accumulation (conv_circular below; refshim._conv_layer) — see DESIGN.md §Oracle.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference
 * repository root).  Arithmetic dtypes follow numba's typing of the reference expressions
 * (float32 array (+,-,*,/) int scalar -> float32; float64 scalar -> float64; float32 scalar /
 * int64 scalar -> float64; np.sum / np.mean accumulate sequentially in the array dtype).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/__init__.py build()).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/alphadia_b200.h"

#ifdef _OPENMP
#include <omp.h>

/* fragments one candidate may keep in the oracle (the device's dense tables stop at ADB_MAX_FRAGMENTS = 32, its ragged
 * results at 64; the oracle covers both) */
#define ORACLE_MAX_FRAGMENTS 128

#endif

#define ISOTOPE_DIFF 1.0033548350700006

/* ------------------------------------------------------------------------------------------
 * np.log(smooth + 1) in selection.py:221-222.  numba fuses the array expression and types it per
 * element with SCALAR rules: float32 + int64 -> float64, log in float64, one rounding to float32
 * on store (verified against numba 0.65: bit-identical on 1536/1536 probes, whereas glibc logf on
 * the float32 sum differs in 0.7 % of values).
 * ---------------------------------------------------------------------------------------- */
static inline float log1p_feature(float x) { return (float)log((double)x + 1.0); }

/* ------------------------------------------------------------------------------------------
 * small helpers mirroring numba/numpy primitives
 * ---------------------------------------------------------------------------------------- */

/* numba quicksort falls back to insertion sort for <= 15 elements (numba/misc/quicksort.py,
 * SMALL_QUICKSORT = 15) => stable ascending argsort.  Used where ties cannot change the result (selection: layers
 * with equal m/z are identical rows; peak values); the fragment orders of the scoring path use argsort_numba_f32. */
static void argsort_f32(const float* v, int n, int* idx) {
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int i = 1; i < n; i++) {
    int k = idx[i];
    float x = v[k];
    int j = i;
    while (j > 0 && x < v[idx[j - 1]]) {
      idx[j] = idx[j - 1];
      j--;
    }
    idx[j] = k;
  }
}
/* np.argsort as numba compiles it (numba/misc/quicksort.py, is_argsort): ranges of more than 15 elements are partitioned
 * around a median-of-three pivot - NOT stable, so the order of exact ties depends on the algorithm - and ranges of up to 15
 * are insertion-sorted.  Restated step by step so that libraries with more than 15 fragments per precursor and tied m/z or
 * intensity values give the reference's fragment order (found by tests/golden/sweep_reference_vs_oracle.py). */
static void numba_insertion_f32(const float* v, int* R, int low, int high) {
  for (int i = low + 1; i <= high; i++) {
    int k = R[i];
    float x = v[k];
    int j = i;
    while (j > low && x < v[R[j - 1]]) { R[j] = R[j - 1]; j--; }
    R[j] = k;
  }
}
static int numba_partition_f32(const float* v, int* R, int low, int high) {
  int mid = (low + high) >> 1, t;
  if (v[R[mid]] < v[R[low]]) { t = R[low]; R[low] = R[mid]; R[mid] = t; }
  if (v[R[high]] < v[R[mid]]) { t = R[high]; R[high] = R[mid]; R[mid] = t; }
  if (v[R[mid]] < v[R[low]]) { t = R[low]; R[low] = R[mid]; R[mid] = t; }
  float pivot = v[R[mid]];
  t = R[high]; R[high] = R[mid]; R[mid] = t;
  int i = low, j = high - 1;
  for (;;) {
    while (i < high && v[R[i]] < pivot) i++;
    while (j >= low && pivot < v[R[j]]) j--;
    if (i >= j) break;
    t = R[i]; R[i] = R[j]; R[j] = t;
    i++; j--;
  }
  t = R[i]; R[i] = R[high]; R[high] = t;
  return i;
}
static void argsort_numba_f32(const float* v, int n, int* R) {
  for (int i = 0; i < n; i++) R[i] = i;
  if (n < 2) return;
  int lo_stack[64], hi_stack[64], sp = 0;
  lo_stack[0] = 0; hi_stack[0] = n - 1; sp = 1;
  while (sp > 0) {
    sp--;
    int low = lo_stack[sp], high = hi_stack[sp];
    while (high - low >= 15) {
      int i = numba_partition_f32(v, R, low, high);
      if (high - i > i - low) {
        if (high > i) { lo_stack[sp] = i + 1; hi_stack[sp] = high; sp++; }
        high = i - 1;
      } else {
        if (i > low) { lo_stack[sp] = low; hi_stack[sp] = i - 1; sp++; }
        low = i + 1;
      }
    }
    numba_insertion_f32(v, R, low, high);
  }
}
static void argsort_f64(const double* v, int n, int* idx) {
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int i = 1; i < n; i++) {
    int k = idx[i];
    double x = v[k];
    int j = i;
    while (j > 0 && x < v[idx[j - 1]]) {
      idx[j] = idx[j - 1];
      j--;
    }
    idx[j] = k;
  }
}

/* np.searchsorted(a, v, 'left') on float32 */
static int64_t searchsorted_left_f32(const float* a, int64_t n, float v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static double median_f64(double* tmp, int n) { /* np.median: sorts a copy */
  for (int i = 1; i < n; i++) {
    double x = tmp[i];
    int j = i;
    while (j > 0 && x < tmp[j - 1]) { tmp[j] = tmp[j - 1]; j--; }
    tmp[j] = x;
  }
  if (n & 1) return tmp[n >> 1];
  return (tmp[(n >> 1) - 1] + tmp[n >> 1]) / 2;
}
static float median_f32(float* tmp, int n) {
  for (int i = 1; i < n; i++) {
    float x = tmp[i];
    int j = i;
    while (j > 0 && x < tmp[j - 1]) { tmp[j] = tmp[j - 1]; j--; }
    tmp[j] = x;
  }
  if (n & 1) return tmp[n >> 1];
  return (float)((double)(float)(tmp[(n >> 1) - 1] + tmp[n >> 1]) / 2);
}

/* ------------------------------------------------------------------------------------------
 * raw-file access (3-D)
 * ---------------------------------------------------------------------------------------- */

/* alphadia/search/jitclasses/utils.py:24-88 get_frame_indices (+ alpharaw_jit.py:173-203) */
static void get_frame_indices_tolerance(const adb_rawfile3d_desc* raw, float rt, double tolerance,
                                        int64_t optimize_size, int64_t min_size, int64_t out[2]) {
  float lim[2] = {(float)((double)rt - tolerance), (float)((double)rt + tolerance)};
  int64_t fi0 = searchsorted_left_f32(raw->rt_values, raw->n_spectra, lim[0]);
  int64_t fi1 = searchsorted_left_f32(raw->rt_values, raw->n_spectra, lim[1]);
  int64_t L = raw->cycle_len;
  int64_t c0 = (fi0 + raw->zeroth_frame) / L;
  int64_t c1 = (fi1 + raw->zeroth_frame) / L;
  int64_t len = c1 - c0;
  int64_t opt = len > min_size ? len : min_size;
  opt = (int64_t)((double)optimize_size * ceil((double)opt / (double)optimize_size));
  int64_t l0 = c0, l1 = c0 + opt;
  int64_t pcmi = raw->precursor_cycle_max_index;
  if (l1 > pcmi) {
    l1 = pcmi;
    l0 = pcmi - opt;
    if (l0 < 0) l0 = (pcmi % 2 == 0) ? 0 : 1;
  }
  out[0] = l0 * L + raw->zeroth_frame;
  out[1] = l1 * L + raw->zeroth_frame;
}

/* alpharaw_jit.py:19-50 _calculate_valid_scans; quad limits are float32, cycle float64 */
static int calculate_valid_scans(const adb_rawfile3d_desc* raw, float q0, float q1, int64_t* out) {
  int n = 0;
  for (int64_t i = 0; i < raw->cycle_len; i++) {
    double mz_start = raw->cycle[2 * i], mz_stop = raw->cycle[2 * i + 1];
    if (((double)q0 <= mz_stop) && ((double)q1 >= mz_start)) out[n++] = i;
  }
  return n;
}

/* utils.py:15-20 mass_range with a float64 tolerance (selection config) */
static void mass_range_f64tol(const float* mz, int n, double ppm, float* lo, float* hi) {
  for (int i = 0; i < n; i++) {
    double d = ppm * (double)mz[i] / 1000000.0;
    lo[i] = (float)((double)mz[i] - d);
    hi[i] = (float)((double)mz[i] + d);
  }
}
/* utils.py:15-20 mass_range with a float32 tolerance (scoring config) */
static void mass_range_f32tol(const float* mz, int n, float ppm, float* lo, float* hi) {
  for (int i = 0; i < n; i++) { /* fused array expression: f32*f32 -> f32, / int64 -> f64, f32 -/+ f64 -> f64 */
    double d = (double)(float)(ppm * mz[i]) / 1000000.0;
    lo[i] = (float)((double)mz[i] - d);
    hi[i] = (float)((double)mz[i] + d);
  }
}

/* alpharaw_jit.py:339-425 get_dense_intensity.  out: [n_q][C] (the two scan rows are identical,
 * alpharaw_jit.py:420-421, so one is stored) */
static void get_dense_intensity(const adb_rawfile3d_desc* raw, int64_t frame_start, int64_t frame_stop,
                                const float* lo, const float* hi, int n_q, float q0, float q1,
                                float* out, int64_t C) {
  int64_t L = raw->cycle_len;
  int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
  int n_pos = calculate_valid_scans(raw, q0, q1, pos);
  int64_t cs = frame_start / L;
  memset(out, 0, sizeof(float) * (size_t)n_q * (size_t)C);
  for (int64_t i = 0; i < C; i++) {
    int64_t cyc = cs + i;
    for (int j = 0; j < n_pos; j++) {
      int64_t scan = pos[j] + cyc * L;
      int64_t idx = raw->peak_start_idx[scan], stop = raw->peak_stop_idx[scan];
      for (int k = 0; k < n_q; k++) {
        int64_t l = idx, r = stop; /* _search_sorted_reference_left, alpharaw_jit.py:67-75 */
        while (l < r) {
          int64_t mid = (l + r) >> 1;
          if (raw->mz_values[mid] < lo[k]) l = mid + 1; else r = mid;
        }
        idx = l;
        while (idx < stop && raw->mz_values[idx] <= hi[k]) {
          out[(int64_t)k * C + i] = out[(int64_t)k * C + i] + raw->intensity_values[idx];
          idx++;
        }
      }
    }
  }
  free(pos);
}

/* alpharaw_jit.py:208-337 get_dense(absolute_masses=True).
 * out_int / out_mz: [n_q][n_pos][C] (scan rows identical, :326-333).  Returns n_pos; pos_out = cycle positions. */
static int get_dense_abs(const adb_rawfile3d_desc* raw, int64_t frame_start, int64_t frame_stop,
                         const float* lo, const float* hi, int n_q, float q0, float q1,
                         float** out_int, float** out_mz, int64_t* C_out, int64_t* pos_out) {
  const double HIGH_EPSILON = 1e-26, LOW_EPSILON = 1e-36;
  int64_t L = raw->cycle_len;
  int n_pos = calculate_valid_scans(raw, q0, q1, pos_out);
  int64_t cs = frame_start / L, ce = frame_stop / L;
  int64_t C = ce - cs;
  if (C < 0) C = 0;
  *C_out = C;
  size_t tot = (size_t)n_q * (size_t)n_pos * (size_t)C;
  float* di = (float*)calloc(tot ? tot : 1, sizeof(float));
  float* dm = (float*)calloc(tot ? tot : 1, sizeof(float));
  for (int64_t i = 0; i < C; i++) {
    int64_t cyc = cs + i;
    for (int j = 0; j < n_pos; j++) {
      int64_t scan = pos_out[j] + cyc * L;
      int64_t idx = raw->peak_start_idx[scan], stop = raw->peak_stop_idx[scan];
      for (int k = 0; k < n_q; k++) {
        idx += searchsorted_left_f32(raw->mz_values + idx, stop - idx, lo[k]);
        size_t cell = ((size_t)k * (size_t)n_pos + (size_t)j) * (size_t)C + (size_t)i;
        while (idx < stop && raw->mz_values[idx] <= hi[k]) {
          float acc_i = di[cell], acc_m = dm[cell];
          float ni = raw->intensity_values[idx];
          ni = ni * (float)((double)ni > HIGH_EPSILON);
          float nm = raw->mz_values[idx];
          float num32 = (float)(acc_m * acc_i) + (float)(ni * nm);
          float den32 = acc_i + ni;
          double nd = ((double)num32 + LOW_EPSILON) / ((double)den32 + LOW_EPSILON);
          di[cell] = acc_i + ni;
          dm[cell] = (float)nd;
          idx++;
        }
      }
    }
  }
  *out_int = di;
  *out_mz = dm;
  return n_pos;
}

/* ------------------------------------------------------------------------------------------
 * SELECTION
 * ---------------------------------------------------------------------------------------- */

/* Definition standing in for alphadia/search/selection/fft.py:141-212 convolve_fourier:
 *   out[i,j] = sum_{a<k0} sum_{b<k1} k[a,b] * x[(i + k0/2 - a) mod n0, (j + k1/2 - b) mod n1]
 * accumulated in fp64 with fma(k, x, acc), a then b ascending, rounded once to f32. */
static void conv_circular(const float* x, int n0, int n1, const float* k, int k0, int k1, float* out) {
  int s0 = k0 / 2, s1 = k1 / 2;
  /* fma(k, 0, acc) == acc exactly, so input rows that are entirely zero are skipped (timsTOF XIC tiles are > 99 % zeros);
   * the result is bit-identical to the full double loop. */
  unsigned char* row_nz = (unsigned char*)calloc((size_t)(n0 > 0 ? n0 : 1), 1);
  for (int i = 0; i < n0; i++)
    for (int j = 0; j < n1; j++)
      if (x[(size_t)i * n1 + j] != 0.0f) { row_nz[i] = 1; break; }
  for (int i = 0; i < n0; i++)
    for (int j = 0; j < n1; j++) {
      double acc = 0.0;
      for (int a = 0; a < k0; a++) {
        int ii = ((i + s0 - a) % n0 + n0) % n0;
        if (!row_nz[ii]) continue;
        const float* xr = x + (size_t)ii * n1;
        const float* kr = k + (size_t)a * k1;
        int jj = (j + s1) % n1; /* (j + s1 - b) mod n1 for b = 0, then step down with wrap */
        for (int b = 0; b < k1; b++) {
          acc = fma((double)kr[b], (double)xr[jj], acc);
          jj = (jj == 0) ? n1 - 1 : jj - 1;
        }
      }
      out[i * n1 + j] = (float)acc;
    }
  free(row_nz);
}

/* alphadia/search/selection/utils.py:205-273 _symetric_limits_1d */
static void symetric_limits_1d(const double* a, int n, int center, double f, double center_fraction,
                               int min_size, int max_size, int out[2]) {
  if (n == 0 || center < 0 || center >= n) { out[0] = center; out[1] = center; return; }
  double center_intensity = a[center], trailing = center_intensity;
  int limit = min_size;
  for (int s = min_size + 1; s < max_size; s++) {
    int l = center - s; if (l < 0) l = 0;
    int r = center + s; if (r > n - 1) r = n - 1;
    double intensity = (a[l] + a[r]) / 2;
    if (intensity < f * trailing) {
      if (intensity > center_intensity * center_fraction) { limit = s; trailing = intensity; }
      else break;
    } else break;
  }
  out[0] = center - limit > 0 ? center - limit : 0;
  out[1] = center + limit + 1 < n ? center + limit + 1 : n;
}

/* alphadia/search/selection/utils.py:276-312 symetric_limits_2d on score[S][C] */
static void symetric_limits_2d(const double* a, int S, int C, int scan_center, int cycle_center,
                               const adb_selection_config* cfg, int scan_lim[2], int cyc_lim[2]) {
  int ml = scan_center - (int)cfg->min_size_mobility; if (ml < 0) ml = 0;
  int mu = scan_center + (int)cfg->min_size_mobility; if (mu > S) mu = S;
  int cl = cycle_center - (int)cfg->min_size_rt; if (cl < 0) cl = 0;
  int cu = cycle_center + (int)cfg->min_size_rt; if (cu > C) cu = C;
  double* ps = (double*)malloc(sizeof(double) * (size_t)(S > 0 ? S : 1));
  double* pc = (double*)malloc(sizeof(double) * (size_t)(C > 0 ? C : 1));
  for (int s = 0; s < S; s++) { double t = 0; for (int c = cl; c < cu; c++) t += a[s * C + c]; ps[s] = t; }
  for (int c = 0; c < C; c++) { double t = 0; for (int s = ml; s < mu; s++) t += a[s * C + c]; pc[c] = t; }
  symetric_limits_1d(ps, S, scan_center, cfg->f_mobility, cfg->center_fraction,
                     (int)cfg->min_size_mobility, (int)cfg->max_size_mobility, scan_lim);
  symetric_limits_1d(pc, C, cycle_center, cfg->f_rt, cfg->center_fraction,
                     (int)cfg->min_size_rt, (int)cfg->max_size_rt, cyc_lim);
  free(ps); free(pc);
}

static int64_t wrap0(int64_t v, int64_t limit) { if (v < 0) return 0; return v < limit ? v : limit; }

/* optional debug taps (tests only) */
typedef struct {
  int64_t precursor_row;  /* which library row to tap, -1 = none */
  float* dense_precursors; /* [I][C] */
  float* dense_fragments;  /* [F][C] */
  double* score;           /* [C]     */
  int64_t capacity;        /* elements available in each buffer */
  int64_t C, F, I;         /* filled */
} adbo_selection_tap;

static void select_one(const adb_rawfile3d_desc* raw, const adb_library_desc* lib,
                       const adb_selection_config* cfg, const float* kernel, int kh, int kw,
                       int64_t i, adb_candidates_out* out, adbo_selection_tap* tap) {
  /* selection.py:112-118 + utils.py:35-40 (float32 += float64) */
  int nI = lib->n_isotopes < cfg->top_k_precursors ? lib->n_isotopes : (int)cfg->top_k_precursors;
  float iso_mz[ADB_MAX_ISOTOPES];
  for (int j = 0; j < nI; j++) {
    double off = (double)j * ISOTOPE_DIFF / (double)lib->charge[i];
    iso_mz[j] = (float)((double)lib->mz[i] + off);
  }
  /* selection.py:120-137 fragments: slice, cardinality filter, sort by m/z */
  int64_t fs = lib->frag_start_idx[i], fe = lib->frag_stop_idx[i];
  int nf_all = (int)(fe - fs);
  if (nf_all < 0) nf_all = 0;
  float* fmz = (float*)malloc(sizeof(float) * (size_t)(nf_all + 1));
  int nF = 0;
  for (int64_t j = fs; j < fe; j++)
    if (!cfg->exclude_shared_ions || lib->frag_cardinality[j] <= 1) fmz[nF++] = lib->frag_mz[j];
  int* order = (int*)malloc(sizeof(int) * (size_t)(nF + 1));
  argsort_f32(fmz, nF, order);
  float* fmz_sorted = (float*)malloc(sizeof(float) * (size_t)(nF + 1));
  for (int j = 0; j < nF; j++) fmz_sorted[j] = fmz[order[j]];
  free(fmz); free(order);
  if (nF <= 3) { free(fmz_sorted); return; }

  /* selection.py:140-148 */
  int64_t fl[2];
  get_frame_indices_tolerance(raw, lib->rt[i], cfg->rt_tolerance, 16, cfg->kernel_size, fl);
  int64_t L = raw->cycle_len;
  int64_t C = fl[1] / L - fl[0] / L;
  const int S = 2; /* alpharaw_jit.py:205-206 */
  if (C <= 0) { free(fmz_sorted); return; }

  /* selection.py:152-170 */
  float* lo = (float*)malloc(sizeof(float) * (size_t)(nF + nI));
  float* hi = (float*)malloc(sizeof(float) * (size_t)(nF + nI));
  float* dp = (float*)malloc(sizeof(float) * (size_t)nI * (size_t)C);
  float* df = (float*)malloc(sizeof(float) * (size_t)nF * (size_t)C);
  mass_range_f64tol(iso_mz, nI, cfg->precursor_mz_tolerance, lo, hi);
  get_dense_intensity(raw, fl[0], fl[1], lo, hi, nI, -1.0f, -1.0f, dp, C);
  mass_range_f64tol(fmz_sorted, nF, cfg->fragment_mz_tolerance, lo, hi);
  get_dense_intensity(raw, fl[0], fl[1], lo, hi, nF, iso_mz[0], iso_mz[nI - 1], df, C);
  free(lo); free(hi);

  if (tap && tap->precursor_row == i) {
    tap->C = C; tap->F = nF; tap->I = nI;
    if ((int64_t)nI * C <= tap->capacity) memcpy(tap->dense_precursors, dp, sizeof(float) * (size_t)nI * (size_t)C);
    if ((int64_t)nF * C <= tap->capacity) memcpy(tap->dense_fragments, df, sizeof(float) * (size_t)nF * (size_t)C);
  }

  /* selection.py:40-75 _is_valid (shape[2] == 2 is even by construction) */
  if (S < kh || C < kw) { free(dp); free(df); free(fmz_sorted); return; }

  /* selection.py:389-428: smooth, log-sum, normalise.  Dense layers are [S=2][C] with identical rows. */
  float* layer = (float*)malloc(sizeof(float) * (size_t)S * (size_t)C);
  float* smooth = (float*)malloc(sizeof(float) * (size_t)S * (size_t)C);
  float* logf_acc = (float*)calloc((size_t)S * (size_t)C, sizeof(float));
  float* logp_acc = (float*)calloc((size_t)S * (size_t)C, sizeof(float));
  for (int l = 0; l < nF; l++) {
    for (int s = 0; s < S; s++) memcpy(layer + (size_t)s * C, df + (size_t)l * C, sizeof(float) * (size_t)C);
    conv_circular(layer, S, (int)C, kernel, kh, kw, smooth);
    for (int64_t t = 0; t < S * C; t++) logf_acc[t] = logf_acc[t] + log1p_feature(smooth[t]);
  }
  for (int l = 0; l < nI; l++) {
    for (int s = 0; s < S; s++) memcpy(layer + (size_t)s * C, dp + (size_t)l * C, sizeof(float) * (size_t)C);
    conv_circular(layer, S, (int)C, kernel, kh, kw, smooth);
    for (int64_t t = 0; t < S * C; t++) logp_acc[t] = logp_acc[t] + log1p_feature(smooth[t]);
  }
  double* score = (double*)malloc(sizeof(double) * (size_t)S * (size_t)C);
  double mean = cfg->use_weighted_score ? cfg->feature_mean : 0.0;
  double std = cfg->use_weighted_score ? cfg->feature_std : 0.0;
  double w = cfg->use_weighted_score ? cfg->feature_weight : 1.0;
  if (!cfg->use_weighted_score) { /* selection.py:405-417 amean1 / astd1 over the single feature */
    double m = 0; float acc = 0;
    for (int64_t t = 0; t < S * C; t++) acc = acc + (float)(logf_acc[t] + logp_acc[t]);
    m = (double)acc / (double)(S * C);
    double v = 0;
    for (int64_t t = 0; t < S * C; t++) { double d = (double)(float)(logf_acc[t] + logp_acc[t]) - m; v += d * d; }
    mean = m; std = sqrt(v / (double)(S * C));
  }
  for (int64_t t = 0; t < S * C; t++) {
    float feat = logf_acc[t] + logp_acc[t];
    score[t] = 0.0 + w * ((double)feat - mean) / (std + 1e-6);
  }
  free(layer); free(smooth); free(logf_acc); free(logp_acc); free(dp); free(df); free(fmz_sorted);
  if (tap && tap->precursor_row == i && C <= tap->capacity) memcpy(tap->score, score, sizeof(double) * (size_t)C);

  /* selection.py:529-544 _find_peaks -> utils.py:45-74 find_peaks_1d (S <= 2) */
  int cap = (int)C;
  int* pk_scan = (int*)malloc(sizeof(int) * (size_t)cap);
  int* pk_cyc = (int*)malloc(sizeof(int) * (size_t)cap);
  double* pk_val = (double*)malloc(sizeof(double) * (size_t)cap);
  int n_pk = 0;
  for (int p = 2; p < C - 2; p++) {
    const double* a = score;
    if (a[p - 2] < a[p - 1] && a[p - 1] < a[p] && a[p] > a[p + 1] && a[p + 1] > a[p + 2]) {
      pk_scan[n_pk] = 0; pk_cyc[n_pk] = p; pk_val[n_pk] = a[p]; n_pk++;
    }
  }
  int* ord = (int*)malloc(sizeof(int) * (size_t)(n_pk + 1));
  argsort_f64(pk_val, n_pk, ord);
  int top_n = (int)cfg->candidate_count < n_pk ? (int)cfg->candidate_count : n_pk;
  int* t_scan = (int*)malloc(sizeof(int) * (size_t)(top_n + 1));
  int* t_cyc = (int*)malloc(sizeof(int) * (size_t)(top_n + 1));
  double* t_val = (double*)malloc(sizeof(double) * (size_t)(top_n + 1));
  for (int r = 0; r < top_n; r++) { int k = ord[n_pk - 1 - r]; t_scan[r] = pk_scan[k]; t_cyc[r] = pk_cyc[k]; t_val[r] = pk_val[k]; }
  free(pk_scan); free(pk_cyc); free(pk_val); free(ord);

  /* selection.py:229-284 _join_close_peaks(…, 3, 3) */
  uint8_t* mask = (uint8_t*)malloc((size_t)(top_n + 1));
  for (int r = 0; r < top_n; r++) mask[r] = 1;
  for (int a = 0; a < top_n; a++) {
    if (!mask[a]) continue;
    for (int b = a + 1; b < top_n; b++) {
      if (!mask[b]) continue;
      if (abs(t_scan[a] - t_scan[b]) <= 3 && abs(t_cyc[a] - t_cyc[b]) <= 3) {
        if (t_val[a] > t_val[b]) mask[b] = 0; else mask[a] = 0;
      }
    }
  }
  int n_c = 0;
  for (int r = 0; r < top_n; r++) if (mask[r]) { t_scan[n_c] = t_scan[r]; t_cyc[n_c] = t_cyc[r]; t_val[n_c] = t_val[r]; n_c++; }
  free(mask);

  /* selection.py:442-462 limits */
  int (*slim)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(n_c + 1));
  int (*clim)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(n_c + 1));
  for (int r = 0; r < n_c; r++) symetric_limits_2d(score, S, (int)C, t_scan[r], t_cyc[r], cfg, slim[r], clim[r]);

  /* selection.py:287-364,465-477 optional joining of overlapping candidates */
  if (cfg->join_close_candidates) {
    uint8_t* jm = (uint8_t*)malloc((size_t)(n_c + 1));
    for (int r = 0; r < n_c; r++) jm[r] = 1;
    for (int a = 0; a < n_c; a++) {
      if (!jm[a]) continue;
      for (int b = a + 1; b < n_c; b++) {
        if (!jm[b]) continue;
        double cycle_len = (double)(clim[a][1] - clim[a][0]);
        int mn = clim[a][1] < clim[b][1] ? clim[a][1] : clim[b][1];
        int mx = clim[a][0] > clim[b][0] ? clim[a][0] : clim[b][0];
        double cycle_overlap = (double)(mn - mx) / cycle_len;
        double scan_len = (double)(slim[a][1] - slim[a][0]);
        mn = slim[a][1] < slim[b][1] ? slim[a][1] : slim[b][1];
        mx = slim[a][0] > slim[b][0] ? slim[a][0] : slim[b][0];
        double scan_overlap = (double)(mn - mx) / scan_len;
        if (scan_overlap < 0 || cycle_overlap < 0) continue;
        if (cycle_overlap > cfg->join_close_candidates_cycle_threshold &&
            scan_overlap > cfg->join_close_candidates_scan_threshold) {
          if (slim[b][0] < slim[a][0]) slim[a][0] = slim[b][0];
          if (slim[b][1] > slim[a][1]) slim[a][1] = slim[b][1];
          if (clim[b][0] < clim[a][0]) clim[a][0] = clim[b][0];
          if (clim[b][1] > clim[a][1]) clim[a][1] = clim[b][1];
          jm[b] = 0;
        }
      }
    }
    int m = 0;
    for (int r = 0; r < n_c; r++) if (jm[r]) {
      t_scan[m] = t_scan[r]; t_cyc[m] = t_cyc[r]; t_val[m] = t_val[r];
      slim[m][0] = slim[r][0]; slim[m][1] = slim[r][1]; clim[m][0] = clim[r][0]; clim[m][1] = clim[r][1]; m++;
    }
    n_c = m;
    free(jm);
  }

  /* selection.py:480-526 write-out; candidate_start_idx = i * candidate_count (selection.py:716-721) */
  int64_t scan_lo = 0; /* scan_limits[0,0] for 3-D */
  for (int r = 0; r < n_c; r++) {
    int64_t row = i * cfg->candidate_count + r;
    if (row >= out->n_rows) break;
    out->precursor_idx[row] = lib->precursor_idx[i];
    out->rank[row] = (uint8_t)r;
    out->score[row] = (float)t_val[r];
    out->scan_center[row] = (uint32_t)wrap0(t_scan[r] + scan_lo, raw->scan_max_index);
    out->scan_start[row] = (uint32_t)wrap0(slim[r][0] + scan_lo, raw->scan_max_index);
    out->scan_stop[row] = (uint32_t)wrap0(slim[r][1] + scan_lo, raw->scan_max_index);
    out->frame_center[row] = (uint32_t)wrap0((int64_t)t_cyc[r] * L + fl[0], raw->frame_max_index);
    out->frame_start[row] = (uint32_t)wrap0((int64_t)clim[r][0] * L + fl[0], raw->frame_max_index);
    out->frame_stop[row] = (uint32_t)wrap0((int64_t)clim[r][1] * L + fl[0], raw->frame_max_index);
  }
  free(t_scan); free(t_cyc); free(t_val); free(slim); free(clim); free(score);
}

static void zero_candidates(adb_candidates_out* out) {
  size_t n = (size_t)out->n_rows;
  memset(out->precursor_idx, 0, 4 * n); memset(out->rank, 0, n); memset(out->score, 0, 4 * n);
  memset(out->scan_center, 0, 4 * n); memset(out->scan_start, 0, 4 * n); memset(out->scan_stop, 0, 4 * n);
  memset(out->frame_center, 0, 4 * n); memset(out->frame_start, 0, 4 * n); memset(out->frame_stop, 0, 4 * n);
}

/* selection.py:78-203 over all precursors (row subset [row_begin, row_end)) */
int adbo_select_candidates(const adb_rawfile3d_desc* raw, const adb_library_desc* lib,
                           const adb_selection_config* cfg, const float* kernel, int32_t kh, int32_t kw,
                           adb_candidates_out* out, int64_t row_begin, int64_t row_end, int32_t n_threads,
                           adbo_selection_tap* tap) {
  if (row_begin == 0 && row_end >= lib->n_precursors) zero_candidates(out);
  if (row_end > lib->n_precursors) row_end = lib->n_precursors;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = row_begin; i < row_end; i++) select_one(raw, lib, cfg, kernel, kh, kw, i, out, tap);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * SCORING
 * ---------------------------------------------------------------------------------------- */

typedef struct { /* FragmentContainer subset, fragment_container.py:12-45 */
  int n;
  float mz_library[ORACLE_MAX_FRAGMENTS], mz[ORACLE_MAX_FRAGMENTS], intensity[ORACLE_MAX_FRAGMENTS];
  uint8_t type[ORACLE_MAX_FRAGMENTS], loss_type[ORACLE_MAX_FRAGMENTS], charge[ORACLE_MAX_FRAGMENTS],
      number[ORACLE_MAX_FRAGMENTS], position[ORACLE_MAX_FRAGMENTS];
} frag_set;

/* features/features_utils.py:9-26 weighted_center_mean on x[S][C] (row-major nonzero order) */
static double weighted_center_mean(const float* x, int S, int C, double scan_center, double frame_center) {
  double values = 0, weights = 0;
  int any = 0;
  for (int s = 0; s < S; s++)
    for (int c = 0; c < C; c++) {
      float v = x[s * C + c];
      if (v > 0) {
        any = 1;
        double ds = (double)s - scan_center, dc = (double)c - frame_center;
        double distance = sqrt(ds * ds + dc * dc);
        double weight = exp(-0.1 * distance);
        values += (double)v * weight;
        weights += weight;
      }
    }
  if (!any) return 0;
  return weights > 0 ? values / weights : 0;
}

/* features/fragment_features.py:20-49 weighted_center_of_mass (scan_mean, frame_mean only) */
static void weighted_center_of_mass(const float* x, int S, int C, double* scan_mean, double* frame_mean) {
  double isum = 0, ssum = 0, fsum = 0;
  int any = 0;
  for (int s = 0; s < S; s++)
    for (int c = 0; c < C; c++) {
      float v = x[s * C + c];
      if (v > 0) { any = 1; isum += (double)v; }
    }
  if (!any) { *scan_mean = 0; *frame_mean = 0; return; }
  for (int s = 0; s < S; s++)
    for (int c = 0; c < C; c++) {
      float v = x[s * C + c];
      if (v > 0) { ssum += (double)s * (double)v; fsum += (double)c * (double)v; }
    }
  *scan_mean = isum > 0 ? ssum / isum : 0;
  *frame_mean = isum > 0 ? fsum / isum : 0;
}

/* scoring/utils.py:46-66 or_envelope over the last axis, rows of length n */
static void or_envelope_rows(float* x, int rows, int n) {
  float* tmp = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  for (int r = 0; r < rows; r++) {
    float* row = x + (size_t)r * n;
    memcpy(tmp, row, sizeof(float) * (size_t)n);
    for (int i = 1; i < n - 1; i++)
      if (tmp[i] < tmp[i - 1] || tmp[i] < tmp[i + 1]) row[i] = (float)((double)(float)(tmp[i - 1] + tmp[i + 1]) / 2);
  }
  free(tmp);
}

/* features/fragment_features.py:71-159 center_envelope_1d, in place on rows of length n */
static void center_envelope_rows(float* x, int rows, int n) {
  if (n % 2 == 0) {
    int cr = n / 2, cl = cr - 1;
    if (cl < 0) return;
    for (int r = 0; r < rows; r++) {
      float* a = x + (size_t)r * n;
      float left = a[cl], right = a[cr];
      for (int i = 1; i <= cl; i++) {
        a[cl - i] = left < a[cl - i] ? left : a[cl - i];
        left = (float)((double)(float)(a[cl - i] + a[cl - i + 1]) * 0.5);
        a[cr + i] = right < a[cr + i] ? right : a[cr + i];
        right = (float)((double)(float)(a[cr + i] + a[cr + i - 1]) * 0.5);
      }
    }
  } else {
    int ci = n / 2;
    if (n < 3) return;
    for (int r = 0; r < rows; r++) {
      float* a = x + (size_t)r * n;
      float left = (float)((double)(float)(a[ci - 1] + a[ci]) * 0.5);
      float right = (float)((double)(float)(a[ci + 1] + a[ci]) * 0.5);
      for (int i = 1; i <= ci; i++) {
        a[ci - i] = left < a[ci - i] ? left : a[ci - i];
        left = (float)((double)(float)(a[ci - i] + a[ci - i + 1]) * 0.5);
        a[ci + i] = right < a[ci + i] ? right : a[ci + i];
        right = (float)((double)(float)(a[ci + i] + a[ci + i - 1]) * 0.5);
      }
    }
  }
}

/* numba np.corrcoef(x, y)[0, 1] (numba/np/arraymath.py np_cov_impl / np_corrcoef_impl), float64 */
static double corrcoef01(const double* x, const double* y, int n) {
  double mx = 0, my = 0;
  for (int i = 0; i < n; i++) { mx += x[i]; my += y[i]; }
  mx /= n; my /= n;
  double cxx = 0, cyy = 0, cxy = 0;
  for (int i = 0; i < n; i++) { double a = x[i] - mx, b = y[i] - my; cxx += a * a; cyy += b * b; cxy += a * b; }
  double fact = 1.0 / (double)(n - 1);
  cxx *= fact; cyy *= fact; cxy *= fact;
  double sx = sqrt(cxx), sy = sqrt(cyy);
  return (cxy / sy) / sx;
}

/* scoring/utils.py:478-510 save_corrcoeff in float64 */
static double save_corrcoeff(const double* x, const double* y, int n, double xbar, double ybar) {
  double num = 0, sxx = 0, syy = 0;
  for (int i = 0; i < n; i++) { double a = x[i] - xbar, b = y[i] - ybar; num += a * b; sxx += a * a; syy += b * b; }
  return num / (sqrt(sxx * syy) + 1e-12);
}

/* scoring/utils.py:574-647 fragment_correlation_different with y = one profile per observation.
 * x: [F][nobs][n]; y: [nobs][n]; out[o][f]  (float32 arithmetic as in the reference) */
static void corr_with_template(const float* x, const float* y, int F, int nobs, int n, float* out) {
  float* yc = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float* xc = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  for (int o = 0; o < nobs; o++) {
    float ys = 0;
    for (int t = 0; t < n; t++) ys = ys + y[o * n + t];
    float ym = ys / (float)n;
    float yss = 0;
    for (int t = 0; t < n; t++) { yc[t] = y[o * n + t] - ym; yss = yss + yc[t] * yc[t]; }
    float ystd = sqrtf(yss / (float)n);
    for (int f = 0; f < F; f++) {
      const float* xr = x + ((size_t)f * nobs + o) * n;
      float xs = 0;
      for (int t = 0; t < n; t++) xs = xs + xr[t];
      float xm = xs / (float)n;
      float xss = 0, dot = 0;
      for (int t = 0; t < n; t++) { xc[t] = xr[t] - xm; xss = xss + xc[t] * xc[t]; }
      for (int t = 0; t < n; t++) dot = dot + xc[t] * yc[t];
      float xstd = sqrtf(xss / (float)n);
      float cov = dot / (float)n;
      float sm = xstd * ystd;
      out[o * F + f] = (float)((double)cov / ((double)sm + 1e-12));
    }
  }
  free(yc); free(xc);
}

typedef struct {
  int64_t candidate;      /* index to tap, -1 none */
  float* dense_fragments; /* [2][F][nobs][C] before qtf mask & fragment mask */
  float* dense_precursors;/* [2][I][C] collapsed */
  float* template_;       /* [nobs][C] */
  int64_t capacity;
  int64_t F, nobs, C, I;
} adbo_scoring_tap;

/* raw-file view shared by the 3-D and 4-D scoring paths */
typedef struct {
  const adb_rawfile3d_desc* r3; /* exactly one of r3 / r4 is set */
  const adb_rawfile4d_desc* r4;
} rawview;

static int extract_cubes_4d(const adb_rawfile4d_desc* raw, int64_t frame_start, int64_t frame_stop, int64_t scan_start,
                            int64_t scan_stop, const float* mz, int n_q, float tol, float q0, float q1, float** out_i,
                            float** out_m, int64_t* C_out, int64_t** obs_out);

static void score_one(const rawview* rv, const adb_library_desc* lib, const adb_scoring_config* cfg,
                      const adb_candidates_in* cand, int64_t ci, adb_scores_out* out, adbo_scoring_tap* tap) {
  const adb_rawfile3d_desc* raw = rv->r3;
  const adb_rawfile4d_desc* raw4 = rv->r4;
  const int K = (int)cfg->top_k_fragments;
  int64_t p = cand->lib_row[ci];
  float* feat = out->features + (size_t)ci * ADB_NUM_FEATURES;

  /* candidate.py:151-163 assemble_isotope_mz (float32(offset) + float32 mz) */
  int nI = lib->n_isotopes < (int)cfg->top_k_isotopes ? lib->n_isotopes : (int)cfg->top_k_isotopes;
  float iso_mz[ADB_MAX_ISOTOPES], iso_int[ADB_MAX_ISOTOPES];
  for (int j = 0; j < nI; j++) {
    double off = (double)j * ISOTOPE_DIFF / (double)lib->charge[p];
    iso_mz[j] = (float)off + lib->mz[p];
    iso_int[j] = lib->isotopes[(size_t)p * lib->n_isotopes + j];
  }

  /* candidate.py:181-192 fragments: slice, cardinality filter, top-k by intensity, sort by m/z */
  frag_set fr;
  {
    int64_t fs = lib->frag_start_idx[p], fe = lib->frag_stop_idx[p];
    int n_all = (int)(fe - fs); if (n_all < 0) n_all = 0;
    int64_t* src = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_all + 1));
    float* inten = (float*)malloc(sizeof(float) * (size_t)(n_all + 1));
    int m = 0;
    for (int64_t j = fs; j < fe; j++)
      if (!cfg->exclude_shared_ions || lib->frag_cardinality[j] <= 1) { src[m] = j; inten[m] = lib->frag_intensity[j]; m++; }
    int* ord = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    argsort_numba_f32(inten, m, ord);
    int k = m < K ? m : K;
    if (k > ORACLE_MAX_FRAGMENTS) k = ORACLE_MAX_FRAGMENTS;
    int64_t sel[ORACLE_MAX_FRAGMENTS]; float selmz[ORACLE_MAX_FRAGMENTS]; int ord2[ORACLE_MAX_FRAGMENTS];
    for (int r = 0; r < k; r++) { sel[r] = src[ord[m - 1 - r]]; selmz[r] = lib->frag_mz[sel[r]]; }
    argsort_numba_f32(selmz, k, ord2);
    fr.n = k;
    for (int r = 0; r < k; r++) {
      int64_t j = sel[ord2[r]];
      fr.mz_library[r] = lib->frag_mz_library[j]; fr.mz[r] = lib->frag_mz[j]; fr.intensity[r] = lib->frag_intensity[j];
      fr.type[r] = lib->frag_type[j]; fr.loss_type[r] = lib->frag_loss_type[j]; fr.charge[r] = lib->frag_charge[j];
      fr.number[r] = lib->frag_number[j]; fr.position[r] = lib->frag_position[j];
    }
    free(src); free(inten); free(ord);
  }
  if (fr.n <= 3) return;

  int64_t frame_start = cand->frame_start[ci], frame_stop = cand->frame_stop[ci], frame_center = cand->frame_center[ci];
  int64_t scan_start = cand->scan_start[ci], scan_stop = cand->scan_stop[ci], scan_center = cand->scan_center[ci];
  const int64_t L = raw4 ? raw4->frames_per_cycle : raw->cycle_len;
  /* a 3-D file is carried as 2 identical scans (alpharaw_jit.py:205-206); a 4-D cube spans the candidate's scans */
  const int S = raw4 ? (int)(scan_stop - scan_start) : 2;
  if (S <= 0) return;

  /* candidate.py:203-205 quadrupole limit (float32 of float64 arithmetic) */
  float mn = iso_mz[0], mx = iso_mz[0];
  for (int j = 1; j < nI; j++) { if (iso_mz[j] < mn) mn = iso_mz[j]; if (iso_mz[j] > mx) mx = iso_mz[j]; }
  float q0 = (float)((double)mn - 0.5), q1 = (float)((double)mx + 0.5);

  /* candidate.py:216-246 dense cubes, always carried as [..][S][C] from here on */
  float lo[ORACLE_MAX_FRAGMENTS], hi[ORACLE_MAX_FRAGMENTS];
  int F = fr.n;
  int64_t C = 0;
  int nobs = 0;
  int64_t* pos_f = NULL; /* observation ids: cycle positions (3-D) / dia_precursor_cycle values (4-D) */
  float *dfi2 = NULL, *dfm2 = NULL, *dpi = NULL, *dpm = NULL;
  if (raw4) {
    int64_t* pos_p = NULL;
    float *dpi_raw = NULL, *dpm_raw = NULL;
    int64_t C2 = 0;
    nobs = extract_cubes_4d(raw4, frame_start, frame_stop, scan_start, scan_stop, fr.mz, F, cfg->fragment_mz_tolerance, q0, q1,
                            &dfi2, &dfm2, &C, &pos_f);
    if (nobs <= 0 || C <= 0 || F <= 1) { free(dfi2); free(dfm2); free(pos_f); return; } /* candidate.py:230-237 */
    int nobs_p = extract_cubes_4d(raw4, frame_start, frame_stop, scan_start, scan_stop, iso_mz, nI, cfg->precursor_mz_tolerance,
                                  -1.0f, -1.0f, &dpi_raw, &dpm_raw, &C2, &pos_p);
    if (nobs_p <= 0) { free(dfi2); free(dfm2); free(pos_f); free(dpi_raw); free(dpm_raw); free(pos_p); return; }
    /* candidate.py:248-269 collapse MS1 observations */
    dpi = (float*)calloc((size_t)nI * S * C, sizeof(float));
    dpm = (float*)calloc((size_t)nI * S * C, sizeof(float));
    for (int i = 0; i < nI; i++)
      for (int sc = 0; sc < S; sc++)
        for (int64_t c = 0; c < C; c++) {
          float s32 = 0; double sm = 0; int count = 0;
          for (int j = 0; j < nobs_p; j++) {
            size_t cell = ((((size_t)i * nobs_p + j) * S) + sc) * C + c;
            s32 = s32 + dpi_raw[cell];
            sm += (double)dpm_raw[cell];
            if (dpm_raw[cell] > 0) count++;
          }
          dpi[((size_t)i * S + sc) * C + c] = s32;
          dpm[((size_t)i * S + sc) * C + c] = (float)(sm / ((double)count + 1e-6));
        }
    free(dpi_raw); free(dpm_raw); free(pos_p);
  } else {
    pos_f = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
    int64_t* pos_p = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
    float *dfi, *dfm, *dpi_raw, *dpm_raw;
    int64_t C2;
    mass_range_f32tol(fr.mz, fr.n, cfg->fragment_mz_tolerance, lo, hi);
    nobs = get_dense_abs(raw, frame_start, frame_stop, lo, hi, fr.n, q0, q1, &dfi, &dfm, &C, pos_f);
    if (C == 0 || fr.n <= 1) { free(dfi); free(dfm); free(pos_f); free(pos_p); return; } /* candidate.py:230-237 */
    mass_range_f32tol(iso_mz, nI, cfg->precursor_mz_tolerance, lo, hi);
    int nobs_p = get_dense_abs(raw, frame_start, frame_stop, lo, hi, nI, -1.0f, -1.0f, &dpi_raw, &dpm_raw, &C2, pos_p);
    /* candidate.py:248-269 collapse MS1 observations; both scan rows are identical (alpharaw_jit.py:326-333) */
    dpi = (float*)calloc((size_t)nI * S * C, sizeof(float));
    dpm = (float*)calloc((size_t)nI * S * C, sizeof(float));
    for (int i = 0; i < nI; i++)
      for (int64_t c = 0; c < C; c++) {
        float s32 = 0; double sm = 0; int count = 0;
        for (int j = 0; j < nobs_p; j++) {
          size_t cell = ((size_t)i * nobs_p + j) * C + c;
          s32 = s32 + dpi_raw[cell];
          sm += (double)dpm_raw[cell];
          if (dpm_raw[cell] > 0) count++;
        }
        for (int sc = 0; sc < S; sc++) {
          dpi[((size_t)i * S + sc) * C + c] = s32;
          dpm[((size_t)i * S + sc) * C + c] = (float)(sm / ((double)count + 1e-6));
        }
      }
    free(dpi_raw); free(dpm_raw); free(pos_p);
    dfi2 = (float*)malloc(sizeof(float) * (size_t)F * (nobs > 0 ? nobs : 1) * S * C);
    dfm2 = (float*)malloc(sizeof(float) * (size_t)F * (nobs > 0 ? nobs : 1) * S * C);
    for (int f = 0; f < F; f++)
      for (int o = 0; o < nobs; o++)
        for (int sc = 0; sc < S; sc++)
          for (int64_t c = 0; c < C; c++) {
            size_t src = ((size_t)f * nobs + o) * C + c;
            size_t dst = (((size_t)f * nobs + o) * S + sc) * C + c;
            dfi2[dst] = dfi[src];
            dfm2[dst] = dfm[src];
          }
    free(dfi); free(dfm);
  }
  (void)tap;

  /* candidate.py:279-284 + quadrupole.py:80-115,261-301: qtf[i][o][s'] with s' over arange(scan_start, scan_stop) */
  int nsc = (int)(scan_stop - scan_start);
  int fail = 0;
  if (!(nsc == 1 || nsc == S) || nobs == 0) fail = 1; /* numba would raise a broadcast error; row stays invalid */
  const int64_t cyc_S = raw4 ? raw4->scans : 1; /* cycle.shape[2] */
  const double* cyc = raw4 ? raw4->cycle : raw->cycle;
  double* qtf = (double*)calloc((size_t)nI * (size_t)(nobs > 0 ? nobs : 1) * (size_t)(nsc > 0 ? nsc : 1), sizeof(double));
  if (!fail)
    for (int i = 0; i < nI; i++)
      for (int o = 0; o < nobs; o++)
        for (int s = 0; s < nsc; s++) {
          int64_t sc = scan_start + s;
          if (sc >= cyc_S || sc < 0) { fail = 1; continue; }
          double mu1 = cyc[(pos_f[o] * cyc_S + sc) * 2 + 0] + cfg->quad_delta_mu[0];
          double mu2 = cyc[(pos_f[o] * cyc_S + sc) * 2 + 1] + cfg->quad_delta_mu[1];
          double x = (double)iso_mz[i];
          double a1 = (x - mu1) / cfg->quad_sigma[0], a2 = (x - mu2) / cfg->quad_sigma[1];
          qtf[((size_t)i * nobs + o) * nsc + s] = 1 / (1 + exp(-a1)) - 1 / (1 + exp(-a2));
        }
  if (fail) { free(qtf); free(dfi2); free(dfm2); free(dpi); free(dpm); free(pos_f); return; }

  /* dense_fragments[0] *= qtf_mask (candidate.py:287-290); the mask broadcasts over scans when nsc == 1 */
  for (int f = 0; f < F; f++)
    for (int o = 0; o < nobs; o++)
      for (int s = 0; s < S; s++) {
        int sq = nsc == 1 ? 0 : s;
        double m = 0;
        for (int i = 0; i < nI; i++) m += qtf[((size_t)i * nobs + o) * nsc + sq];
        float mask = (float)(m / (double)nI);
        for (int64_t c = 0; c < C; c++) {
          size_t dst = (((size_t)f * nobs + o) * S + s) * C + c;
          dfi2[dst] = dfi2[dst] * mask;
        }
      }

  /* quadrupole.py:304-324 template[o][s][c] */
  float* tmpl = (float*)malloc(sizeof(float) * (size_t)nobs * S * C);
  for (int o = 0; o < nobs; o++)
    for (int s = 0; s < S; s++) {
      int sq = nsc == 1 ? 0 : s;
      for (int64_t c = 0; c < C; c++) {
        double t = 0;
        for (int i = 0; i < nI; i++) t += (double)(float)(dpi[((size_t)i * S + s) * C + c] * iso_int[i]) * qtf[((size_t)i * nobs + o) * nsc + sq];
        tmpl[((size_t)o * S + s) * C + c] = (float)t;
      }
    }
  free(qtf);

  /* quadrupole.py:327-335 observation importance (float32) */
  float* oi = (float*)malloc(sizeof(float) * (size_t)nobs);
  {
    float tot = 0;
    for (int o = 0; o < nobs; o++) {
      float so = 0;
      for (int s = 0; s < S; s++) { float sc = 0; for (int64_t c = 0; c < C; c++) sc = sc + tmpl[((size_t)o * S + s) * C + c]; so = so + sc; }
      oi[o] = so; tot = tot + so;
    }
    if (tot == 0) for (int o = 0; o < nobs; o++) oi[o] = 1.0f / (float)nobs;
    else for (int o = 0; o < nobs; o++) oi[o] = oi[o] / tot;
  }

  /* candidate.py:319-329 fragment mask */
  uint8_t fmask[ORACLE_MAX_FRAGMENTS];
  int Fv = 0;
  for (int f = 0; f < F; f++) {
    float t_o = 0;
    for (int o = 0; o < nobs; o++) {
      float t_s = 0;
      for (int s = 0; s < S; s++) { float t_c = 0; for (int64_t c = 0; c < C; c++) t_c = t_c + dfi2[(((size_t)f * nobs + o) * S + s) * C + c]; t_s = t_s + t_c; }
      t_o = t_o + t_s;
    }
    fmask[f] = t_o > 0; Fv += fmask[f];
  }
  if (Fv < 2) { free(dfi2); free(dfm2); free(tmpl); free(oi); free(dpi); free(dpm); free(pos_f); return; }

  /* compact cubes + fragment_container.py:104-120 apply_mask (renormalise intensities) */
  {
    int w = 0;
    for (int f = 0; f < F; f++) if (fmask[f]) {
      if (w != f) {
        memmove(dfi2 + (size_t)w * nobs * S * C, dfi2 + (size_t)f * nobs * S * C, 4 * (size_t)nobs * S * C);
        memmove(dfm2 + (size_t)w * nobs * S * C, dfm2 + (size_t)f * nobs * S * C, 4 * (size_t)nobs * S * C);
        fr.mz_library[w] = fr.mz_library[f]; fr.mz[w] = fr.mz[f]; fr.intensity[w] = fr.intensity[f];
        fr.type[w] = fr.type[f]; fr.loss_type[w] = fr.loss_type[f]; fr.charge[w] = fr.charge[f];
        fr.number[w] = fr.number[f]; fr.position[w] = fr.position[f];
      }
      w++;
    }
    float isum = 0;
    for (int f = 0; f < Fv; f++) isum = isum + fr.intensity[f];
    for (int f = 0; f < Fv; f++) fr.intensity[f] = fr.intensity[f] / isum;
  }
  int Fall = F;
  F = Fv; fr.n = Fv;

  /* candidate.py:333-347 profiles */
  float* ffp = (float*)malloc(sizeof(float) * (size_t)F * nobs * C); /* fragments_frame_profile [F][nobs][C] */
  float* fsp = (float*)malloc(sizeof(float) * (size_t)F * nobs * S); /* fragments_scan_profile  [F][nobs][S] */
  for (int f = 0; f < F; f++)
    for (int o = 0; o < nobs; o++) {
      for (int64_t c = 0; c < C; c++) { float t = 0; for (int s = 0; s < S; s++) t = t + dfi2[(((size_t)f * nobs + o) * S + s) * C + c]; ffp[((size_t)f * nobs + o) * C + c] = t; }
      for (int s = 0; s < S; s++) { float t = 0; for (int64_t c = 0; c < C; c++) t = t + dfi2[(((size_t)f * nobs + o) * S + s) * C + c]; fsp[((size_t)f * nobs + o) * S + s] = t; }
    }
  or_envelope_rows(fsp, F * nobs, S);
  float* tfp = (float*)malloc(sizeof(float) * (size_t)nobs * C); /* template_frame_profile [nobs][C] */
  float* tsp = (float*)malloc(sizeof(float) * (size_t)nobs * S);
  for (int o = 0; o < nobs; o++) {
    for (int64_t c = 0; c < C; c++) { float t = 0; for (int s = 0; s < S; s++) t = t + tmpl[((size_t)o * S + s) * C + c]; tfp[(size_t)o * C + c] = t; }
    for (int s = 0; s < S; s++) { float t = 0; for (int64_t c = 0; c < C; c++) t = t + tmpl[((size_t)o * S + s) * C + c]; tsp[(size_t)o * S + s] = t; }
  }
  or_envelope_rows(tfp, nobs, (int)C);
  or_envelope_rows(tsp, nobs, S);

  float fa[ADB_NUM_FEATURES];
  memset(fa, 0, sizeof(fa));
  fa[28] = (float)((double)Fv / (double)Fall); /* candidate.py:362 */

  /* features/location_features.py:9-33 */
  if (raw4) { /* float64 arrays in the timsTOF view */
    fa[0] = (float)(raw4->mobility_values[scan_start] - raw4->mobility_values[scan_stop - 1]);
    fa[1] = (float)(raw4->rt_values[frame_stop - 1] - raw4->rt_values[frame_start]);
    fa[2] = (float)raw4->rt_values[frame_center];
    fa[3] = (float)raw4->mobility_values[scan_center];
  } else {
    fa[0] = raw->mobility_values[scan_start] - raw->mobility_values[scan_stop - 1];
    fa[1] = raw->rt_values[frame_stop - 1] - raw->rt_values[frame_start];
    fa[2] = raw->rt_values[frame_center];
    fa[3] = raw->mobility_values[scan_center];
  }

  /* ---------------- features/precursor_features.py:14-102 ---------------- */
  {
    float spi[ADB_MAX_ISOTOPES]; /* sum_precursor_intensity [I][1] */
    for (int i = 0; i < nI; i++) {
      float t_s = 0;
      for (int s = 0; s < S; s++) { float t_c = 0; for (int64_t c = 0; c < C; c++) t_c = t_c + dpi[((size_t)i * S + s) * C + c]; t_s = t_s + t_c; }
      spi[i] = t_s;
    }
    float wspi[ADB_MAX_ISOTOPES];
    for (int i = 0; i < nI; i++) { float t = 0; for (int o = 0; o < nobs; o++) t = t + spi[i] * oi[o]; wspi[i] = t; }
    int amax = 0;
    for (int i = 1; i < nI; i++) if (iso_int[i] > iso_int[amax]) amax = i;
    fa[4] = wspi[0];
    fa[5] = wspi[amax];
    { float t = 0; for (int i = 0; i < nI; i++) t = t + wspi[i]; fa[6] = t; }
    { float t = 0; for (int i = 0; i < nI; i++) t = t + wspi[i] * iso_int[i]; fa[7] = t; }
    /* precursor_features.py:52-65: "centres" are the sizes (n_scans, n_observations = 1) */
    double H[ADB_MAX_ISOTOPES], MZo[ADB_MAX_ISOTOPES];
    for (int i = 0; i < nI; i++) {
      H[i] = weighted_center_mean(dpi + (size_t)i * S * C, S, (int)C, (double)S, 1.0);
      MZo[i] = weighted_center_mean(dpm + (size_t)i * S * C, S, (int)C, (double)S, 1.0);
    }
    double wme = 0;
    for (int i = 0; i < nI; i++) if (MZo[i] > 0) {
      double me = (MZo[i] - (double)iso_mz[i]) / (double)iso_mz[i] * 1e6;
      wme += me * (double)iso_int[i];
    }
    fa[8] = (float)wme;
    fa[9] = (float)fabs(wme);
    fa[10] = (float)((double)iso_mz[0] + wme * 1e-6 * (double)iso_mz[0]);
    fa[11] = (float)H[0];
    fa[12] = (float)H[amax];
    { double t = 0; for (int i = 0; i < nI; i++) t += H[i]; fa[13] = (float)t; }
    { double t = 0; for (int i = 0; i < nI; i++) t += H[i] * (double)iso_int[i]; fa[14] = (float)t; }
    double xi[ADB_MAX_ISOTOPES], yi[ADB_MAX_ISOTOPES];
    float sx = 0, sy = 0;
    for (int i = 0; i < nI; i++) { xi[i] = iso_int[i]; yi[i] = spi[i]; sx = sx + iso_int[i]; sy = sy + spi[i]; }
    double xbar = (double)sx / (double)nI;
    fa[15] = (float)save_corrcoeff(xi, yi, nI, xbar, (double)sy / (double)nI);
    double hbar = 0; for (int i = 0; i < nI; i++) hbar += H[i]; hbar /= (double)nI;
    fa[16] = (float)save_corrcoeff(xi, H, nI, xbar, hbar);
  }

  /* ---------------- features/fragment_features.py:198-427 ---------------- */
  double ofmm[ORACLE_MAX_FRAGMENTS], mass_error[ORACLE_MAX_FRAGMENTS], ofh_mean[ORACLE_MAX_FRAGMENTS], area_norm[ORACLE_MAX_FRAGMENTS];
  float fin[ORACLE_MAX_FRAGMENTS];
  {
    fa[17] = (float)nobs;
    { float t = 0; for (int f = 0; f < F; f++) t = t + fr.intensity[f]; for (int f = 0; f < F; f++) fin[f] = fr.intensity[f] / t; }
    double* esc = (double*)malloc(sizeof(double) * (size_t)nobs);
    double* efc = (double*)malloc(sizeof(double) * (size_t)nobs);
    for (int o = 0; o < nobs; o++) weighted_center_of_mass(tmpl + (size_t)o * S * C, S, (int)C, &esc[o], &efc[o]);

    /* best_profile [F][C] */
    float* bp = (float*)malloc(sizeof(float) * (size_t)F * C);
    int best_obs = 0;
    if (cfg->quant_all) {
      for (int f = 0; f < F; f++)
        for (int64_t c = 0; c < C; c++) { float t = 0; for (int o = 0; o < nobs; o++) t = t + ffp[((size_t)f * nobs + o) * C + c]; bp[(size_t)f * C + c] = t; }
      center_envelope_rows(bp, F, (int)C);
    } else {
      for (int o = 1; o < nobs; o++) if (oi[o] > oi[best_obs]) best_obs = o;
      for (int f = 0; f < F; f++) memcpy(bp + (size_t)f * C, ffp + ((size_t)f * nobs + best_obs) * C, 4 * (size_t)C);
      center_envelope_rows(bp, F, (int)C);
      /* fragment_features.py:248-250: best_profile is a VIEW, the envelope mutates fragments_frame_profile */
      for (int f = 0; f < F; f++) memcpy(ffp + ((size_t)f * nobs + best_obs) * C, bp + (size_t)f * C, 4 * (size_t)C);
    }
    int64_t qw = (int64_t)cfg->quant_window;
    if ((C / 2) - 1 < qw) qw = (C / 2) - 1;
    int64_t center = C / 2;
    int64_t w0 = center - qw, w1 = center + qw + 1;
    if (qw < 0) { w0 = 0; w1 = 0; } /* empty python slice */
    if (w1 > C) w1 = C;
    if (w0 < 0) w0 = 0;
    int64_t wn = w1 - w0; if (wn < 0) wn = 0;
    float ofi[ORACLE_MAX_FRAGMENTS];
    for (int f = 0; f < F; f++) {
      double area = 0;
      for (int64_t t = 0; t + 1 < wn; t++) {
        float sum2 = bp[(size_t)f * C + w0 + t + 1] + bp[(size_t)f * C + w0 + t];
        if (raw4) { /* float64 rt: (f32 + f32) * f64 * 0.5 */
          double drt = raw4->rt_values[frame_start + (w0 + t + 1) * L] - raw4->rt_values[frame_start + (w0 + t) * L];
          area += (double)sum2 * drt * 0.5;
        } else {
          float drt = raw->rt_values[frame_start + (w0 + t + 1) * L] - raw->rt_values[frame_start + (w0 + t) * L];
          area += (double)(float)(sum2 * drt) * 0.5;
        }
      }
      area_norm[f] = area * (double)qw;
      float t = 0;
      for (int64_t u = 0; u < wn; u++) t = t + bp[(size_t)f * C + w0 + u];
      ofi[f] = t;
    }
    free(bp);

    /* sum_fragment_intensity [F][nobs], sum_template_intensity [nobs] */
    float* sfi = (float*)malloc(sizeof(float) * (size_t)F * nobs);
    for (int f = 0; f < F; f++)
      for (int o = 0; o < nobs; o++) {
        float t_s = 0;
        for (int s = 0; s < S; s++) { float t_c = 0; for (int64_t c = 0; c < C; c++) t_c = t_c + dfi2[(((size_t)f * nobs + o) * S + s) * C + c]; t_s = t_s + t_c; }
        sfi[(size_t)f * nobs + o] = t_s;
      }
    float* sti = (float*)malloc(sizeof(float) * (size_t)nobs);
    for (int o = 0; o < nobs; o++) {
      float t_s = 0;
      for (int s = 0; s < S; s++) { float t_c = 0; for (int64_t c = 0; c < C; c++) t_c = t_c + tmpl[((size_t)o * S + s) * C + c]; t_s = t_s + t_c; }
      sti[o] = t_s;
    }

    double* ofmz = (double*)malloc(sizeof(double) * (size_t)F * nobs);
    double* ofh = (double*)malloc(sizeof(double) * (size_t)F * nobs);
    for (int f = 0; f < F; f++)
      for (int o = 0; o < nobs; o++) {
        ofmz[(size_t)f * nobs + o] = weighted_center_mean(dfm2 + ((size_t)f * nobs + o) * S * C, S, (int)C, esc[o], efc[o]);
        ofh[(size_t)f * nobs + o] = weighted_center_mean(dfi2 + ((size_t)f * nobs + o) * S * C, S, (int)C, esc[o], efc[o]);
      }
    free(esc); free(efc);

    int any_height = 0; double sum_ofh_mean = 0;
    for (int f = 0; f < F; f++) {
      /* fragment_features.py:312-336 */
      float wsum = 0; int anyh = 0;
      for (int o = 0; o < nobs; o++) { int m = ofh[(size_t)f * nobs + o] > 0; anyh |= m; wsum = wsum + (m ? oi[o] : 0.0f); }
      any_height += anyh;
      double wtot = 0; int cnt = 0;
      double* wrow = (double*)malloc(sizeof(double) * (size_t)nobs);
      int nn = nobs;
      for (int o = 0; o < nn; o++) {
        int m = ofh[(size_t)f * nobs + o] > 0;
        double wv = (double)(m ? oi[o] : 0.0f) / ((double)wsum + 1e-20);
        wrow[o] = wv;
        if (wv > 0) { wtot += wv; cnt++; }
      }
      double a = 0, b = 0;
      if (cnt > 0)
        for (int o = 0; o < nn; o++) if (wrow[o] > 0) { double lw = wrow[o] / wtot; a += ofmz[(size_t)f * nobs + o] * lw; b += ofh[(size_t)f * nobs + o] * lw; }
      ofmm[f] = a; ofh_mean[f] = b;
      sum_ofh_mean += b;
      free(wrow);
    }
    double find[ORACLE_MAX_FRAGMENTS];
    for (int f = 0; f < F; f++) find[f] = (double)fin[f];
    if (any_height > 0) fa[18] = (float)corrcoef01(area_norm, find, F);
    if (sum_ofh_mean > 0.0) fa[19] = (float)corrcoef01(ofh_mean, find, F);
    int n20 = 0, n21 = 0; float s22 = 0, s23 = 0;
    for (int f = 0; f < F; f++) { if (ofi[f] > 0.0f) { n20++; s22 = s22 + fin[f]; } }
    for (int f = 0; f < F; f++) { if (ofh_mean[f] > 0.0) { n21++; s23 = s23 + fin[f]; } }
    fa[20] = (float)((double)n20 / (double)F);
    fa[21] = (float)((double)n21 / (double)F);
    fa[22] = s22; fa[23] = s23;
    if (n20 > 0) { /* fragment_features.py:356-363 + features_utils.py:40-47 */
      float tn2 = 0;
      for (int o = 0; o < nobs; o++) tn2 = tn2 + sti[o] * sti[o];
      float tnorm = sqrtf(tn2);
      float acc = 0;
      for (int f = 0; f < F; f++) if (ofi[f] > 0) {
        float fn2 = 0, dot = 0;
        for (int o = 0; o < nobs; o++) { float v = sfi[(size_t)f * nobs + o]; fn2 = fn2 + v * v; dot = dot + v * sti[o]; }
        double div = (double)(float)(sqrtf(fn2) * tnorm) + 0.0001;
        acc = acc + (float)((double)dot / div);
      }
      fa[24] = (float)((double)acc / (double)n20);
    }
    free(sfi); free(sti); free(ofmz); free(ofh);
    /* fragment_features.py:367-383 */
    { float sb = 0, sy = 0; int nb = 0, ny = 0;
      for (int f = 0; f < F; f++) { if (fr.type[f] == 98) { sb = sb + ofi[f]; nb++; } if (fr.type[f] == 121) { sy = sy + ofi[f]; ny++; } }
      fa[25] = nb > 0 ? (float)log((double)sb + 1) : 0.0f;
      fa[26] = ny > 0 ? (float)log((double)sy + 1) : 0.0f;
      fa[27] = fa[25] - fa[26];
    }
    /* fragment_features.py:387-396 */
    for (int f = 0; f < F; f++) mass_error[f] = (ofmm[f] - (double)fr.mz[f]) / (double)fr.mz[f] * 1e6;
    int ord[ORACLE_MAX_FRAGMENTS];
    argsort_numba_f32(fr.intensity, F, ord);
    { int n3 = F < 3 ? F : 3; double t = 0; for (int r = 0; r < n3; r++) t += mass_error[ord[F - 1 - r]]; fa[41] = (float)(t / (double)n3); }
    { double t = 0; for (int f = 0; f < F; f++) t += mass_error[f]; fa[42] = (float)(t / (double)F); }
    /* fragment_features.py:400-420 */
    { int nb = 0, ny = 0, min_y = 255, max_b = 0;
      for (int f = 0; f < F; f++) { if (fr.type[f] == 98) { nb++; if (fr.position[f] > max_b) max_b = fr.position[f]; } if (fr.type[f] == 121) { ny++; if (fr.position[f] < min_y) min_y = fr.position[f]; } }
      if (nb > 0 && ny > 0) {
        int n_ov = 0; double sa = 0, se = 0;
        for (int f = 0; f < F; f++) {
          int ov = (fr.type[f] == 121 && fr.position[f] < max_b) || (fr.type[f] == 98 && fr.position[f] > min_y);
          if (ov) { n_ov++; sa += area_norm[f]; se += mass_error[f]; }
        }
        fa[43] = (float)n_ov;
        if (n_ov > 0) { fa[44] = (float)(sa / (double)n_ov); fa[45] = (float)(se / (double)n_ov); }
        else { fa[44] = 0; fa[45] = 15; }
      }
    }
  }

  /* ---------------- features/fragment_features.py:430-480 fragment_mobility_correlation (has_mobility) ------- */
  if (raw4) {
    int idx[ORACLE_MAX_FRAGMENTS], nz = 0;
    for (int f = 0; f < F; f++) {
      float t_o = 0;
      for (int o = 0; o < nobs; o++) { float t_s = 0; for (int sc = 0; sc < S; sc++) t_s = t_s + fsp[((size_t)f * nobs + o) * S + sc]; t_o = t_o + t_s; }
      if (t_o > 0) idx[nz++] = f;
    }
    if (nz >= 3) {
      float norm[ORACLE_MAX_FRAGMENTS];
      { float t = 0; for (int a = 0; a < nz; a++) t = t + fr.intensity[idx[a]]; for (int a = 0; a < nz; a++) norm[a] = fr.intensity[idx[a]] / t; }
      float* red = (float*)calloc((size_t)nz * nz, sizeof(float));
      float* cen = (float*)malloc(sizeof(float) * (size_t)nz * S);
      float stdv[ORACLE_MAX_FRAGMENTS];
      for (int o = 0; o < nobs; o++) { /* scoring/utils.py:513-571 fragment_correlation on the scan profiles */
        for (int a = 0; a < nz; a++) {
          const float* r = fsp + ((size_t)idx[a] * nobs + o) * S;
          float sm = 0; for (int sc = 0; sc < S; sc++) sm = sm + r[sc];
          float mean = sm / (float)S;
          float ss = 0;
          for (int sc = 0; sc < S; sc++) { cen[(size_t)a * S + sc] = r[sc] - mean; ss = ss + cen[(size_t)a * S + sc] * cen[(size_t)a * S + sc]; }
          stdv[a] = sqrtf(ss / (float)S);
        }
        for (int a = 0; a < nz; a++)
          for (int b = 0; b < nz; b++) {
            float dot = 0;
            for (int sc = 0; sc < S; sc++) dot = dot + cen[(size_t)a * S + sc] * cen[(size_t)b * S + sc];
            float cov = dot / (float)S;
            float smx = stdv[a] * stdv[b];
            float corr = (float)((double)cov / ((double)smx + 1e-12));
            red[(size_t)a * nz + b] = red[(size_t)a * nz + b] + corr * oi[o];
          }
      }
      float lsum = 0;
      for (int a = 0; a < nz; a++) { float t = 0; for (int b = 0; b < nz; b++) t = t + red[(size_t)a * nz + b] * norm[b]; lsum = lsum + t; }
      fa[29] = (float)((double)lsum / (double)nz);
      free(red); free(cen);
      /* template scan correlation: fragment_correlation_different against the template scan profile */
      float* sub = (float*)malloc(sizeof(float) * (size_t)nz * nobs * S);
      for (int a = 0; a < nz; a++) memcpy(sub + (size_t)a * nobs * S, fsp + (size_t)idx[a] * nobs * S, sizeof(float) * (size_t)nobs * S);
      float* ct = (float*)malloc(sizeof(float) * (size_t)nobs * nz);
      corr_with_template(sub, tsp, nz, nobs, S, ct);
      float t30 = 0;
      for (int a = 0; a < nz; a++) { float r = 0; for (int o = 0; o < nobs; o++) r = r + ct[(size_t)o * nz + a] * oi[o]; t30 = t30 + r * norm[a]; }
      fa[30] = t30;
      free(sub); free(ct);
    }
  }

  /* ---------------- features/profile_features.py:18-206 ---------------- */
  float corr_list[ORACLE_MAX_FRAGMENTS];
  {
    int ord[ORACLE_MAX_FRAGMENTS];
    argsort_numba_f32(fr.intensity, F, ord);
    int sorted_idx[ORACLE_MAX_FRAGMENTS];
    for (int r = 0; r < F; r++) sorted_idx[r] = ord[F - 1 - r];
    int n3 = F < 3 ? F : 3;
    float top3 = 0;
    if (cfg->experimental_xic) {
      /* scoring_utils.py:79-124,127-152,20-76 */
      float* isl = (float*)malloc(sizeof(float) * (size_t)F * C);
      float* nrm = (float*)calloc((size_t)F * C, sizeof(float));
      for (int f = 0; f < F; f++)
        for (int64_t c = 0; c < C; c++) { float t = 0; for (int o = 0; o < nobs; o++) t = t + ffp[((size_t)f * nobs + o) * C + c]; isl[(size_t)f * C + c] = t; }
      int64_t center = C / 2;
      int64_t a0 = center - 1, a1 = center + 2;
      if (a0 < 0) { a0 += C; if (a0 < 0) a0 = 0; }
      if (a1 > C) a1 = C;
      for (int f = 0; f < F; f++) {
        float t = 0; int64_t wn = a1 - a0 > 0 ? a1 - a0 : 0;
        for (int64_t c = a0; c < a1; c++) t = t + isl[(size_t)f * C + c];
        double ci_ = (double)t / (double)wn;
        if (ci_ > 0) for (int64_t c = 0; c < C; c++) nrm[(size_t)f * C + c] = (float)((double)isl[(size_t)f * C + c] / ci_);
      }
      float* med = (float*)malloc(sizeof(float) * (size_t)C);
      float tmpv[ORACLE_MAX_FRAGMENTS];
      for (int64_t c = 0; c < C; c++) { for (int f = 0; f < F; f++) tmpv[f] = nrm[(size_t)f * C + c]; med[c] = median_f32(tmpv, F); }
      /* correlation_coefficient(median_profile, intensity_slice) */
      float sx = 0; for (int64_t c = 0; c < C; c++) sx = sx + med[c];
      double mxv = (double)sx / (double)C;
      double varx = 0;
      for (int64_t c = 0; c < C; c++) { double d = (double)med[c] - mxv; varx += d * d; }
      varx /= (double)C;
      for (int f = 0; f < F; f++) {
        float sy = 0; for (int64_t c = 0; c < C; c++) sy = sy + isl[(size_t)f * C + c];
        float myv = (float)((double)sy / (double)C);
        double cov = 0; float vy32 = 0;
        for (int64_t c = 0; c < C; c++) {
          float ym = isl[(size_t)f * C + c] - myv;
          cov += ((double)med[c] - mxv) * (double)ym;
          vy32 = vy32 + ym * ym;
        }
        cov /= (double)C;
        double vary = (double)vy32 / (double)C;
        double vxy = varx * vary;
        corr_list[f] = vxy == 0 ? 0.0f : (float)(cov / sqrt(vxy));
      }
      free(isl); free(nrm); free(med);
      { float t = 0; for (int r = 0; r < n3; r++) t = t + corr_list[sorted_idx[r]]; top3 = (float)((double)t / (double)n3); }
    } else {
      /* scoring/utils.py:513-571 fragment_correlation, float32 throughout */
      float* red = (float*)calloc((size_t)F * F, sizeof(float));
      float* cen = (float*)malloc(sizeof(float) * (size_t)F * C);
      float stdv[ORACLE_MAX_FRAGMENTS];
      for (int o = 0; o < nobs; o++) {
        for (int f = 0; f < F; f++) {
          const float* r = ffp + ((size_t)f * nobs + o) * C;
          float s = 0; for (int64_t c = 0; c < C; c++) s = s + r[c];
          float m = s / (float)C;
          float ss = 0;
          for (int64_t c = 0; c < C; c++) { cen[(size_t)f * C + c] = r[c] - m; ss = ss + cen[(size_t)f * C + c] * cen[(size_t)f * C + c]; }
          stdv[f] = sqrtf(ss / (float)C);
        }
        for (int f = 0; f < F; f++)
          for (int g = 0; g < F; g++) {
            float dot = 0;
            for (int64_t c = 0; c < C; c++) dot = dot + cen[(size_t)f * C + c] * cen[(size_t)g * C + c];
            float cov = dot / (float)C;
            float sm = stdv[f] * stdv[g];
            float corr = (float)((double)cov / ((double)sm + 1e-12));
            red[(size_t)f * F + g] = red[(size_t)f * F + g] + corr * oi[o];
          }
      }
      for (int f = 0; f < F; f++) { float t = 0; for (int g = 0; g < F; g++) t = t + red[(size_t)f * F + g] * fr.intensity[g]; corr_list[f] = t; }
      { float t = 0; for (int a = 0; a < n3; a++) for (int b = 0; b < n3; b++) t = t + red[(size_t)sorted_idx[a] * F + sorted_idx[b]]; top3 = (float)((double)t / (double)(n3 * n3)); }
      free(red); free(cen);
    }
    { float t = 0; for (int f = 0; f < F; f++) t = t + corr_list[f]; fa[31] = (float)((double)t / (double)F); }
    fa[32] = top3;
    /* profile_features.py:75-90 */
    {
      float* ct = (float*)malloc(sizeof(float) * (size_t)nobs * F);
      corr_with_template(ffp, tfp, F, nobs, (int)C, ct);
      float t33 = 0;
      for (int f = 0; f < F; f++) { float r = 0; for (int o = 0; o < nobs; o++) r = r + ct[(size_t)o * F + f] * oi[o]; t33 = t33 + r * fr.intensity[f]; }
      fa[33] = t33;
      free(ct);
    }
    /* profile_features.py:94-113: the type mask indexes the SORTED-index array by position (quirk) */
    { int nb = 0, ny = 0; float sb = 0, sy = 0;
      for (int r = 0; r < F; r++) {
        if (fr.type[r] == 98) { if (nb < 3) sb = sb + corr_list[sorted_idx[r]]; nb++; }
        if (fr.type[r] == 121) { if (ny < 3) sy = sy + corr_list[sorted_idx[r]]; ny++; }
      }
      if (nb > 0) { fa[34] = (float)((double)sb / (double)(nb < 3 ? nb : 3)); fa[35] = (float)nb; }
      if (ny > 0) { fa[36] = (float)((double)sy / (double)(ny < 3 ? ny : 3)); fa[37] = (float)ny; }
    }
    /* profile_features.py:117-146 cycle_fwhm */
    {
      const double rt_width = raw4 ? raw4->rt_values[frame_stop - 1] - raw4->rt_values[frame_start]
                                   : (double)(float)(raw->rt_values[frame_stop - 1] - raw->rt_values[frame_start]);
      float agg = 0;
      for (int f = 0; f < F; f++) {
        float ml = 0;
        for (int o = 0; o < nobs; o++) {
          const float* r = ffp + ((size_t)f * nobs + o) * C;
          float mxv = r[0];
          for (int64_t c = 1; c < C; c++) if (r[c] > mxv) mxv = r[c];
          double half = (double)mxv / 2;
          int na = 0;
          for (int64_t c = 0; c < C; c++) if ((double)r[c] > half) na++;
          float fw = (float)(((double)na / (double)C) * (double)rt_width);
          ml = ml + fw * oi[o];
        }
        agg = agg + ml * fr.intensity[f];
      }
      fa[38] = agg;
    }
    /* profile_features.py:148-188 mobility_fwhm (has_mobility only) */
    if (raw4) {
      const double mobility_width = raw4->mobility_values[scan_start] - raw4->mobility_values[scan_stop - 1];
      float agg = 0;
      for (int f = 0; f < F; f++) {
        float ml = 0;
        for (int o = 0; o < nobs; o++) {
          const float* r = fsp + ((size_t)f * nobs + o) * S;
          float mxv = r[0];
          for (int sc = 1; sc < S; sc++) if (r[sc] > mxv) mxv = r[sc];
          double half = (double)mxv / 2;
          int na = 0;
          for (int sc = 0; sc < S; sc++) if ((double)r[sc] > half) na++;
          float fw = (float)(((double)na / (double)S) * mobility_width);
          ml = ml + fw * oi[o];
        }
        agg = agg + ml * fr.intensity[f];
      }
      fa[39] = agg;
    }
    /* profile_features.py:190-204 delta_frame_peak */
    {
      double acc = 0;
      double tmpd[ORACLE_MAX_FRAGMENTS];
      for (int o = 0; o < nobs; o++) {
        for (int f = 0; f < F; f++) {
          const float* r = ffp + ((size_t)f * nobs + o) * C;
          int am = 0;
          for (int64_t c = 1; c < C; c++) if (r[c] > r[am]) am = (int)c;
          tmpd[f] = (double)am;
        }
        float medp = (float)median_f64(tmpd, F);
        double delta = (double)medp - floor((double)C / 2);
        acc += delta * (double)oi[o];
      }
      fa[40] = (float)acc;
    }
  }

  /* candidate.py:403-442,475-481 outputs */
  if (cfg->collect_fragments) {
    size_t base = (size_t)ci * (size_t)K;
    for (int f = 0; f < F && f < K; f++) {
      out->fragment_mz_library[base + f] = fr.mz_library[f];
      out->fragment_mz[base + f] = fr.mz[f];
      out->fragment_mz_observed[base + f] = (float)ofmm[f];
      out->fragment_height[base + f] = (float)ofh_mean[f];
      out->fragment_intensity[base + f] = (float)area_norm[f];
      out->fragment_mass_error[base + f] = (float)mass_error[f];
      out->fragment_correlation[base + f] = corr_list[f];
      out->fragment_position[base + f] = fr.position[f];
      out->fragment_number[base + f] = fr.number[f];
      out->fragment_type[base + f] = fr.type[f];
      out->fragment_charge[base + f] = fr.charge[f];
      out->fragment_loss_type[base + f] = fr.loss_type[f];
    }
  }
  memcpy(feat, fa, sizeof(fa));
  out->valid[ci] = 1;

  free(dfi2); free(dfm2); free(tmpl); free(oi); free(dpi); free(dpm); free(pos_f);
  free(ffp); free(fsp); free(tfp); free(tsp);
}

static void zero_scores(const adb_scoring_config* cfg, int64_t n, adb_scores_out* out) {
  size_t K = cfg->top_k_fragments, N = (size_t)n;
  memset(out->features, 0, 4 * N * ADB_NUM_FEATURES);
  memset(out->valid, 0, N);
  float* f32s[] = {out->fragment_mz_library, out->fragment_mz, out->fragment_mz_observed, out->fragment_height,
                   out->fragment_intensity, out->fragment_mass_error, out->fragment_correlation};
  for (int i = 0; i < 7; i++) memset(f32s[i], 0, 4 * N * K);
  uint8_t* u8s[] = {out->fragment_position, out->fragment_number, out->fragment_type, out->fragment_charge, out->fragment_loss_type};
  for (int i = 0; i < 5; i++) memset(u8s[i], 0, N * K);
}

/* scoring.py:114-137 over all candidates */
int adbo_score_candidates(const adb_rawfile3d_desc* raw, const adb_library_desc* lib, const adb_scoring_config* cfg,
                          const adb_candidates_in* cand, adb_scores_out* out, int32_t n_threads, adbo_scoring_tap* tap) {
  if (cfg->top_k_fragments > ORACLE_MAX_FRAGMENTS) return 1;
  zero_scores(cfg, cand->n, out);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  rawview rv = {raw, NULL};
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t ci = 0; ci < cand->n; ci++) score_one(&rv, lib, cfg, cand, ci, out, tap);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * FRAGMENT COMPETITION — alphadia/fragcomp/fragcomp.py:19-143
 * ---------------------------------------------------------------------------------------- */
static int overlap_f32(const float* a, int na, const float* b, int nb, double tol) {
  int n = 0;
  for (int i = 0; i < na; i++)
    for (int j = 0; j < nb; j++) {
      float d = fabsf(a[i] - b[j]);
      double ppm = (double)(float)(d / a[i]) * 1e6;
      n += ppm < tol;
    }
  return n;
}
static int overlap_f64(const double* a, int na, const double* b, int nb, double tol) {
  int n = 0;
  for (int i = 0; i < na; i++)
    for (int j = 0; j < nb; j++) {
      double ppm = fabs(a[i] - b[j]) / a[i] * 1e6;
      n += ppm < tol;
    }
  return n;
}

int adbo_fragment_competition(int64_t n_windows, const int64_t* window_start, const int64_t* window_stop,
                              int64_t n_psm, const void* rt, const int64_t* frag_start_idx,
                              const int64_t* frag_stop_idx, int64_t n_frag, const void* fragment_mz,
                              int32_t is_f64, double rt_tol_seconds, double mass_tol_ppm, uint8_t* valid,
                              int32_t n_threads) {
  (void)n_psm; (void)n_frag;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t w = 0; w < n_windows; w++) {
    int64_t s = window_start[w], e = window_stop[w];
    for (int64_t i = s; i < e; i++) {
      if (!valid[i]) continue;
      for (int64_t j = s; j < e; j++) {
        if (i == j) continue;
        if (!valid[j]) continue;
        double drt;
        if (is_f64 & 1) drt = fabs(((const double*)rt)[i] - ((const double*)rt)[j]);
        else drt = (double)fabsf(((const float*)rt)[i] - ((const float*)rt)[j]);
        if (drt < rt_tol_seconds) {
          int ov;
          if (is_f64 & 2)
            ov = overlap_f64((const double*)fragment_mz + frag_start_idx[i], (int)(frag_stop_idx[i] - frag_start_idx[i]),
                             (const double*)fragment_mz + frag_start_idx[j], (int)(frag_stop_idx[j] - frag_start_idx[j]), mass_tol_ppm);
          else
            ov = overlap_f32((const float*)fragment_mz + frag_start_idx[i], (int)(frag_stop_idx[i] - frag_start_idx[i]),
                             (const float*)fragment_mz + frag_start_idx[j], (int)(frag_stop_idx[j] - frag_start_idx[j]), mass_tol_ppm);
          if (ov >= 3) valid[j] = 0;
        }
      }
    }
  }
  return 0;
}

/* exported helpers for unit tests */
void adbo_conv_circular(const float* x, int n0, int n1, const float* k, int k0, int k1, float* out) { conv_circular(x, n0, n1, k, k0, k1, out); }
void adbo_frame_indices(const adb_rawfile3d_desc* raw, float rt, double tol, int64_t optimize_size, int64_t min_size, int64_t* out) {
  get_frame_indices_tolerance(raw, rt, tol, optimize_size, min_size, out);
}
void adbo_symetric_limits_1d(const double* a, int n, int center, double f, double cf, int min_size, int max_size, int* out) {
  symetric_limits_1d(a, n, center, f, cf, min_size, max_size, out);
}
void adbo_center_envelope(float* x, int rows, int n) { center_envelope_rows(x, rows, n); }
int adbo_fragment_overlap_f64(const double* a, int na, const double* b, int nb, double tol) { return overlap_f64(a, na, b, nb, tol); }
int adbo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}


/* ==========================================================================================
 * 4-D (timsTOF) raw files — alphadia/search/jitclasses/bruker_jit.py
 * ======================================================================================== */

static int64_t searchsorted_left_f64(const double* a, int64_t n, double v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

/* utils.py:24-88 get_frame_indices on float64 rt_values (bruker_jit.py:172-202); limits are float32 */
static void get_frame_indices_tolerance_4d(const adb_rawfile4d_desc* raw, float rt, double tolerance, int64_t optimize_size,
                                           int64_t min_size, int64_t out[2]) {
  float lim[2] = {(float)((double)rt - tolerance), (float)((double)rt + tolerance)};
  int64_t fi0 = searchsorted_left_f64(raw->rt_values, raw->n_frames, (double)lim[0]);
  int64_t fi1 = searchsorted_left_f64(raw->rt_values, raw->n_frames, (double)lim[1]);
  int64_t L = raw->frames_per_cycle;
  int64_t c0 = (fi0 + raw->zeroth_frame) / L, c1 = (fi1 + raw->zeroth_frame) / L;
  int64_t len = c1 - c0;
  int64_t opt = len > min_size ? len : min_size;
  opt = (int64_t)((double)optimize_size * ceil((double)opt / (double)optimize_size));
  int64_t l0 = c0, l1 = c0 + opt;
  int64_t pcmi = raw->precursor_cycle_max_index;
  if (l1 > pcmi) { l1 = pcmi; l0 = pcmi - opt; if (l0 < 0) l0 = (pcmi % 2 == 0) ? 0 : 1; }
  out[0] = l0 * L + raw->zeroth_frame;
  out[1] = l1 * L + raw->zeroth_frame;
}

/* bruker_jit.py:204-271 get_scan_indices_tolerance: searchsorted(mobility_values[::-1], v, "right") */
static void get_scan_indices_tolerance_4d(const adb_rawfile4d_desc* raw, float mobility, double tolerance,
                                          int64_t optimize_size, int64_t out[2]) {
  float lim[2] = {(float)((double)mobility + tolerance), (float)((double)mobility - tolerance)};
  int64_t n = raw->scans, si[2];
  for (int k = 0; k < 2; k++) { /* "right" on the ascending reversed array */
    double v = (double)lim[k];
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (raw->mobility_values[n - 1 - mid] <= v) lo = mid + 1; else hi = mid; }
    si[k] = raw->scan_max_index - lo;
  }
  int64_t scan_len = si[0] - si[1];
  int64_t opt = (int64_t)((double)optimize_size * ceil((double)scan_len / (double)optimize_size));
  int64_t l0 = si[0], l1 = si[0] - opt;
  if (l1 < 0) { l1 = 0; l0 = opt; if (l0 > raw->scan_max_index) l0 = raw->scan_max_index; }
  out[0] = l0; out[1] = l1;
}

/* push query of bruker_jit.py:315-350 as a predicate + the observation id of a push */
typedef struct {
  int64_t f0, f1, s0, s1;  /* frame / scan limits (step 1) */
  const uint8_t* mask;     /* [Fr * Sc] cycle mask for the quadrupole window */
  int64_t ncyc_pos;        /* Fr * Sc */
} push_query;

static inline int push_in_query(const adb_rawfile4d_desc* raw, const push_query* q, int64_t push, int64_t* pos_out) {
  int64_t frame = push / raw->scan_max_index, scan = push % raw->scan_max_index;
  if (frame < q->f0 || frame >= q->f1 || scan < q->s0 || scan >= q->s1) return 0;
  int64_t cyclic = raw->zeroth_frame ? push - raw->scan_max_index : push;
  int64_t pos = ((cyclic % q->ncyc_pos) + q->ncyc_pos) % q->ncyc_pos;
  *pos_out = pos;
  return q->mask[pos];
}

static uint8_t* cycle_mask_4d(const adb_rawfile4d_desc* raw, float q0, float q1) { /* bruker_jit.py:280-313 */
  int64_t n = raw->frames_per_cycle * raw->scans;
  uint8_t* m = (uint8_t*)malloc((size_t)n);
  for (int64_t k = 0; k < n; k++) m[k] = ((double)q0 <= raw->cycle[2 * k + 1]) && ((double)q1 >= raw->cycle[2 * k]);
  return m;
}

/* sorted unique observation ids present in the push query (np.unique(precursor_index), bruker_jit.py:368) */
static int query_observations(const adb_rawfile4d_desc* raw, const push_query* q, int64_t** obs_out) {
  int64_t Fr = raw->frames_per_cycle;
  uint8_t* seen = (uint8_t*)calloc((size_t)(Fr > 0 ? Fr : 1) + 1, 1);
  int64_t max_id = Fr; /* ids are frame-in-cycle indices in practice; grow if needed */
  int64_t cap = max_id + 1;
  for (int64_t frame = q->f0; frame < q->f1; frame++)
    for (int64_t scan = q->s0; scan < q->s1; scan++) {
      int64_t pos;
      if (push_in_query(raw, q, frame * raw->scan_max_index + scan, &pos)) {
        int64_t id = raw->dia_precursor_cycle[pos];
        if (id < 0) continue;
        if (id >= cap) { int64_t ncap = id + 1; seen = (uint8_t*)realloc(seen, (size_t)ncap); memset(seen + cap, 0, (size_t)(ncap - cap)); cap = ncap; }
        seen[id] = 1;
      }
    }
  int n = 0;
  for (int64_t k = 0; k < cap; k++) n += seen[k];
  int64_t* obs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int w = 0;
  for (int64_t k = 0; k < cap; k++) if (seen[k]) obs[w++] = k;
  free(seen);
  *obs_out = obs;
  return n;
}

/* tof slices: searchsorted(mz_values f64, mass_range(mz f32, tol), "left") (bruker_jit.py:273-278,596-598) */
static void tof_limits_4d(const adb_rawfile4d_desc* raw, float lo, float hi, int64_t* t0, int64_t* t1) {
  *t0 = searchsorted_left_f64(raw->mz_values, raw->n_tof, (double)lo);
  *t1 = searchsorted_left_f64(raw->mz_values, raw->n_tof, (double)hi);
}

/* first event of tof row `t` with push >= p */
static int64_t row_lower_bound(const adb_rawfile4d_desc* raw, int64_t t, int64_t p) {
  int64_t lo = raw->tof_indptr[t], hi = raw->tof_indptr[t + 1];
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((int64_t)raw->push_indices[mid] < p) lo = mid + 1; else hi = mid; }
  return lo;
}

/* bruker_jit.py:586-615 get_dense(absolute_masses=True) -> _assemble_push (:352-504).
 * out_i / out_m: [n_q][nobs][S][C]; returns nobs (0 = empty query), obs ids in *obs_out. */
static int extract_cubes_4d(const adb_rawfile4d_desc* raw, int64_t frame_start, int64_t frame_stop, int64_t scan_start,
                            int64_t scan_stop, const float* mz, int n_q, float tol, float q0, float q1, float** out_i,
                            float** out_m, int64_t* C_out, int64_t** obs_out) {
  const double HIGH_EPSILON = 1e-26, LOW_EPSILON = 1e-36;
  *out_i = NULL; *out_m = NULL; *obs_out = NULL; *C_out = 0;
  int64_t Fr = raw->frames_per_cycle;
  push_query q = {frame_start, frame_stop, scan_start, scan_stop, cycle_mask_4d(raw, q0, q1), Fr * raw->scans};
  int64_t* obs = NULL;
  int nobs = query_observations(raw, &q, &obs);
  if (nobs == 0) { free((void*)q.mask); free(obs); return 0; }
  int64_t S = scan_stop - scan_start;
  int64_t cs = (frame_start - raw->zeroth_frame) / Fr, ce = (frame_stop - raw->zeroth_frame) / Fr;
  int64_t C = ce - cs;
  if (S <= 0 || C <= 0) { free((void*)q.mask); free(obs); return 0; }
  size_t tot = (size_t)n_q * nobs * S * C;
  float* di = (float*)calloc(tot, sizeof(float));
  float* dm = (float*)calloc(tot, sizeof(float));
  float lo[ORACLE_MAX_FRAGMENTS > ADB_MAX_ISOTOPES ? ORACLE_MAX_FRAGMENTS : ADB_MAX_ISOTOPES], hi[ORACLE_MAX_FRAGMENTS > ADB_MAX_ISOTOPES ? ORACLE_MAX_FRAGMENTS : ADB_MAX_ISOTOPES];
  mass_range_f32tol(mz, n_q, tol, lo, hi);
  for (int j = 0; j < n_q; j++) {
    int64_t t0, t1;
    tof_limits_4d(raw, lo[j], hi[j], &t0, &t1);
    for (int64_t t = t0; t < t1; t++) {
      double measured = raw->mz_values[t];
      int64_t e = row_lower_bound(raw, t, frame_start * raw->scan_max_index), end = raw->tof_indptr[t + 1];
      for (; e < end; e++) {
        int64_t push = raw->push_indices[e];
        if (push >= frame_stop * raw->scan_max_index) break;
        int64_t pos;
        if (!push_in_query(raw, &q, push, &pos)) continue;
        int64_t id = raw->dia_precursor_cycle[pos];
        int o = 0;
        while (o < nobs && obs[o] != id) o++;
        if (o == nobs) continue;
        int64_t frame = push / raw->scan_max_index, scan = push % raw->scan_max_index;
        int64_t rc = (frame - raw->zeroth_frame) / Fr - cs, rs = scan - scan_start;
        if (rc < 0 || rc >= C) continue;
        size_t cell = ((((size_t)j * nobs + o) * S) + rs) * C + rc;
        float acc_i = di[cell], acc_m = dm[cell];
        int64_t ni = (int64_t)raw->intensity_values[e] * (((double)raw->intensity_values[e]) > HIGH_EPSILON);
        double num = (double)(float)(acc_m * acc_i) + (double)ni * measured + LOW_EPSILON;
        double den = ((double)acc_i + (double)ni) + LOW_EPSILON;
        di[cell] = (float)((double)acc_i + (double)ni);
        dm[cell] = (float)(num / den);
      }
    }
  }
  free((void*)q.mask);
  *out_i = di; *out_m = dm; *obs_out = obs; *C_out = C;
  return nobs;
}

/* bruker_jit.py:617-645 get_dense_intensity -> _assemble_push_intensity (:506-584). out: [n_q][S][C].
 * Returns 0 when the push query is empty (the reference returns a 0-sized array -> _is_valid fails). */
static int dense_intensity_4d(const adb_rawfile4d_desc* raw, const int64_t fl[2], const int64_t sl[2], const float* lo,
                              const float* hi, int n_q, float q0, float q1, float* out, int64_t S, int64_t C) {
  int64_t Fr = raw->frames_per_cycle;
  push_query q = {fl[0], fl[1], sl[0], sl[1], cycle_mask_4d(raw, q0, q1), Fr * raw->scans};
  int64_t* obs = NULL;
  int nobs = query_observations(raw, &q, &obs);
  free(obs);
  memset(out, 0, sizeof(float) * (size_t)n_q * S * C);
  if (nobs == 0) { free((void*)q.mask); return 0; }
  int64_t cs = (fl[0] - raw->zeroth_frame) / Fr;
  for (int j = 0; j < n_q; j++) {
    int64_t t0, t1;
    tof_limits_4d(raw, lo[j], hi[j], &t0, &t1);
    for (int64_t t = t0; t < t1; t++) {
      int64_t e = row_lower_bound(raw, t, fl[0] * raw->scan_max_index), end = raw->tof_indptr[t + 1];
      for (; e < end; e++) {
        int64_t push = raw->push_indices[e];
        if (push >= fl[1] * raw->scan_max_index) break;
        int64_t pos;
        if (!push_in_query(raw, &q, push, &pos)) continue;
        int64_t frame = push / raw->scan_max_index, scan = push % raw->scan_max_index;
        int64_t rc = (frame - raw->zeroth_frame) / Fr - cs, rs = scan - sl[0];
        if (rc < 0 || rc >= C) continue;
        size_t cell = ((size_t)j * S + rs) * C + rc;
        out[cell] = out[cell] + (float)raw->intensity_values[e];
      }
    }
  }
  free((void*)q.mask);
  return 1;
}

/* selection.py:78-203 for one precursor of a 4-D file */
static void select_one_4d(const adb_rawfile4d_desc* raw, const adb_library_desc* lib, const adb_selection_config* cfg,
                          const float* kernel, int kh, int kw, int64_t i, adb_candidates_out* out) {
  int nI = lib->n_isotopes < cfg->top_k_precursors ? lib->n_isotopes : (int)cfg->top_k_precursors;
  float iso_mz[ADB_MAX_ISOTOPES];
  for (int j = 0; j < nI; j++) iso_mz[j] = (float)((double)lib->mz[i] + (double)j * ISOTOPE_DIFF / (double)lib->charge[i]);
  int64_t fs = lib->frag_start_idx[i], fe = lib->frag_stop_idx[i];
  int nf_all = (int)(fe - fs); if (nf_all < 0) nf_all = 0;
  float* fmz = (float*)malloc(sizeof(float) * (size_t)(nf_all + 1));
  int nF = 0;
  for (int64_t j = fs; j < fe; j++) if (!cfg->exclude_shared_ions || lib->frag_cardinality[j] <= 1) fmz[nF++] = lib->frag_mz[j];
  int* order = (int*)malloc(sizeof(int) * (size_t)(nF + 1));
  argsort_f32(fmz, nF, order);
  float* fsorted = (float*)malloc(sizeof(float) * (size_t)(nF + 1));
  for (int j = 0; j < nF; j++) fsorted[j] = fmz[order[j]];
  free(fmz); free(order);
  if (nF <= 3) { free(fsorted); return; }
  int64_t fl[2], sl[2];
  get_frame_indices_tolerance_4d(raw, lib->rt[i], cfg->rt_tolerance, 16, cfg->kernel_size, fl);
  get_scan_indices_tolerance_4d(raw, lib->mobility[i], cfg->mobility_tolerance, 16, sl);
  int64_t Fr = raw->frames_per_cycle;
  int64_t C = (fl[1] - raw->zeroth_frame) / Fr - (fl[0] - raw->zeroth_frame) / Fr;
  int64_t S = sl[1] - sl[0];
  if (C <= 0 || S <= 0 || sl[0] < 0 || sl[1] > raw->scan_max_index) { free(fsorted); return; }
  float* lo = (float*)malloc(sizeof(float) * (size_t)(nF + nI));
  float* hi = (float*)malloc(sizeof(float) * (size_t)(nF + nI));
  float* dp = (float*)malloc(sizeof(float) * (size_t)nI * S * C);
  float* df = (float*)malloc(sizeof(float) * (size_t)nF * S * C);
  mass_range_f64tol(iso_mz, nI, cfg->precursor_mz_tolerance, lo, hi);
  int okp = dense_intensity_4d(raw, fl, sl, lo, hi, nI, -1.0f, -1.0f, dp, S, C);
  mass_range_f64tol(fsorted, nF, cfg->fragment_mz_tolerance, lo, hi);
  int okf = dense_intensity_4d(raw, fl, sl, lo, hi, nF, iso_mz[0], iso_mz[nI - 1], df, S, C);
  free(lo); free(hi); free(fsorted);
  /* selection.py:40-75 _is_valid */
  if (!okp || !okf || (S % 2) != 0 || S < kh || C < kw) { free(dp); free(df); return; }
  float* smooth = (float*)malloc(sizeof(float) * (size_t)S * C);
  float* lf = (float*)calloc((size_t)S * C, sizeof(float));
  float* lp = (float*)calloc((size_t)S * C, sizeof(float));
  for (int l = 0; l < nF; l++) {
    conv_circular(df + (size_t)l * S * C, (int)S, (int)C, kernel, kh, kw, smooth);
    for (int64_t t = 0; t < S * C; t++) lf[t] = lf[t] + log1p_feature(smooth[t]);
  }
  for (int l = 0; l < nI; l++) {
    conv_circular(dp + (size_t)l * S * C, (int)S, (int)C, kernel, kh, kw, smooth);
    for (int64_t t = 0; t < S * C; t++) lp[t] = lp[t] + log1p_feature(smooth[t]);
  }
  double* score = (double*)malloc(sizeof(double) * (size_t)S * C);
  double mean = cfg->use_weighted_score ? cfg->feature_mean : 0.0;
  double std = cfg->use_weighted_score ? cfg->feature_std : 0.0;
  double w = cfg->use_weighted_score ? cfg->feature_weight : 1.0;
  if (!cfg->use_weighted_score) {
    float acc = 0;
    for (int64_t t = 0; t < S * C; t++) acc = acc + (float)(lf[t] + lp[t]);
    mean = (double)acc / (double)(S * C);
    double v = 0;
    for (int64_t t = 0; t < S * C; t++) { double d = (double)(float)(lf[t] + lp[t]) - mean; v += d * d; }
    std = sqrt(v / (double)(S * C));
  }
  for (int64_t t = 0; t < S * C; t++) score[t] = 0.0 + w * ((double)(float)(lf[t] + lp[t]) - mean) / (std + 1e-6);
  free(smooth); free(lf); free(lp); free(dp); free(df);

  /* selection/utils.py:77-110 find_peaks_2d (or find_peaks_1d when S <= 2) */
  size_t cap = (size_t)S * C;
  int* pk_scan = (int*)malloc(sizeof(int) * cap);
  int* pk_cyc = (int*)malloc(sizeof(int) * cap);
  double* pk_val = (double*)malloc(sizeof(double) * cap);
  int n_pk = 0;
  if (S <= 2) {
    for (int p = 2; p < C - 2; p++) {
      const double* a = score;
      if (a[p - 2] < a[p - 1] && a[p - 1] < a[p] && a[p] > a[p + 1] && a[p + 1] > a[p + 2]) { pk_scan[n_pk] = 0; pk_cyc[n_pk] = p; pk_val[n_pk] = a[p]; n_pk++; }
    }
  } else {
    for (int sidx = 2; sidx < S - 2; sidx++)
      for (int p = 2; p < C - 2; p++) {
        const double* a = score;
#define A2(ss, pp) a[(size_t)(ss) * C + (pp)]
        int pk = A2(sidx - 2, p) < A2(sidx - 1, p) && A2(sidx - 1, p) < A2(sidx, p) && A2(sidx, p) > A2(sidx + 1, p) && A2(sidx + 1, p) > A2(sidx + 2, p);
        pk = pk && A2(sidx, p - 2) < A2(sidx, p - 1) && A2(sidx, p - 1) < A2(sidx, p) && A2(sidx, p) > A2(sidx, p + 1) && A2(sidx, p + 1) > A2(sidx, p + 2);
        if (pk) { pk_scan[n_pk] = sidx; pk_cyc[n_pk] = p; pk_val[n_pk] = A2(sidx, p); n_pk++; }
#undef A2
      }
  }
  int* ord = (int*)malloc(sizeof(int) * (size_t)(n_pk + 1));
  argsort_f64(pk_val, n_pk, ord);
  int top_n = (int)cfg->candidate_count < n_pk ? (int)cfg->candidate_count : n_pk;
  int* t_scan = (int*)malloc(sizeof(int) * (size_t)(top_n + 1));
  int* t_cyc = (int*)malloc(sizeof(int) * (size_t)(top_n + 1));
  double* t_val = (double*)malloc(sizeof(double) * (size_t)(top_n + 1));
  for (int r = 0; r < top_n; r++) { int k = ord[n_pk - 1 - r]; t_scan[r] = pk_scan[k]; t_cyc[r] = pk_cyc[k]; t_val[r] = pk_val[k]; }
  free(pk_scan); free(pk_cyc); free(pk_val); free(ord);
  uint8_t* mask = (uint8_t*)malloc((size_t)(top_n + 1));
  for (int r = 0; r < top_n; r++) mask[r] = 1;
  for (int a = 0; a < top_n; a++) {
    if (!mask[a]) continue;
    for (int b = a + 1; b < top_n; b++) {
      if (!mask[b]) continue;
      if (abs(t_scan[a] - t_scan[b]) <= 3 && abs(t_cyc[a] - t_cyc[b]) <= 3) { if (t_val[a] > t_val[b]) mask[b] = 0; else mask[a] = 0; }
    }
  }
  int n_c = 0;
  for (int r = 0; r < top_n; r++) if (mask[r]) { t_scan[n_c] = t_scan[r]; t_cyc[n_c] = t_cyc[r]; t_val[n_c] = t_val[r]; n_c++; }
  free(mask);
  int (*slim)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(n_c + 1));
  int (*clim)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(n_c + 1));
  for (int r = 0; r < n_c; r++) symetric_limits_2d(score, (int)S, (int)C, t_scan[r], t_cyc[r], cfg, slim[r], clim[r]);
  if (cfg->join_close_candidates) {
    uint8_t* jm = (uint8_t*)malloc((size_t)(n_c + 1));
    for (int r = 0; r < n_c; r++) jm[r] = 1;
    for (int a = 0; a < n_c; a++) {
      if (!jm[a]) continue;
      for (int b = a + 1; b < n_c; b++) {
        if (!jm[b]) continue;
        double cycle_len = (double)(clim[a][1] - clim[a][0]);
        int mn = clim[a][1] < clim[b][1] ? clim[a][1] : clim[b][1];
        int mx = clim[a][0] > clim[b][0] ? clim[a][0] : clim[b][0];
        double cycle_overlap = (double)(mn - mx) / cycle_len;
        double scan_len = (double)(slim[a][1] - slim[a][0]);
        mn = slim[a][1] < slim[b][1] ? slim[a][1] : slim[b][1];
        mx = slim[a][0] > slim[b][0] ? slim[a][0] : slim[b][0];
        double scan_overlap = (double)(mn - mx) / scan_len;
        if (scan_overlap < 0 || cycle_overlap < 0) continue;
        if (cycle_overlap > cfg->join_close_candidates_cycle_threshold && scan_overlap > cfg->join_close_candidates_scan_threshold) {
          if (slim[b][0] < slim[a][0]) slim[a][0] = slim[b][0];
          if (slim[b][1] > slim[a][1]) slim[a][1] = slim[b][1];
          if (clim[b][0] < clim[a][0]) clim[a][0] = clim[b][0];
          if (clim[b][1] > clim[a][1]) clim[a][1] = clim[b][1];
          jm[b] = 0;
        }
      }
    }
    int m = 0;
    for (int r = 0; r < n_c; r++) if (jm[r]) {
      t_scan[m] = t_scan[r]; t_cyc[m] = t_cyc[r]; t_val[m] = t_val[r];
      slim[m][0] = slim[r][0]; slim[m][1] = slim[r][1]; clim[m][0] = clim[r][0]; clim[m][1] = clim[r][1]; m++;
    }
    n_c = m;
    free(jm);
  }
  for (int r = 0; r < n_c; r++) { /* selection.py:480-526 */
    int64_t row = i * cfg->candidate_count + r;
    if (row >= out->n_rows) break;
    out->precursor_idx[row] = lib->precursor_idx[i];
    out->rank[row] = (uint8_t)r;
    out->score[row] = (float)t_val[r];
    out->scan_center[row] = (uint32_t)wrap0(t_scan[r] + sl[0], raw->scan_max_index);
    out->scan_start[row] = (uint32_t)wrap0(slim[r][0] + sl[0], raw->scan_max_index);
    out->scan_stop[row] = (uint32_t)wrap0(slim[r][1] + sl[0], raw->scan_max_index);
    out->frame_center[row] = (uint32_t)wrap0((int64_t)t_cyc[r] * Fr + fl[0], raw->frame_max_index);
    out->frame_start[row] = (uint32_t)wrap0((int64_t)clim[r][0] * Fr + fl[0], raw->frame_max_index);
    out->frame_stop[row] = (uint32_t)wrap0((int64_t)clim[r][1] * Fr + fl[0], raw->frame_max_index);
  }
  free(t_scan); free(t_cyc); free(t_val); free(slim); free(clim); free(score);
}

int adbo_select_candidates_4d(const adb_rawfile4d_desc* raw, const adb_library_desc* lib, const adb_selection_config* cfg,
                              const float* kernel, int32_t kh, int32_t kw, adb_candidates_out* out, int32_t n_threads) {
  zero_candidates(out);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < lib->n_precursors; i++) select_one_4d(raw, lib, cfg, kernel, kh, kw, i, out);
  return 0;
}

int adbo_score_candidates_4d(const adb_rawfile4d_desc* raw, const adb_library_desc* lib, const adb_scoring_config* cfg,
                             const adb_candidates_in* cand, adb_scores_out* out, int32_t n_threads) {
  if (cfg->top_k_fragments > ORACLE_MAX_FRAGMENTS) return 1;
  zero_scores(cfg, cand->n, out);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  rawview rv = {NULL, raw};
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t ci = 0; ci < cand->n; ci++) score_one(&rv, lib, cfg, cand, ci, out, NULL);
  return 0;
}

void adbo_scan_indices_4d(const adb_rawfile4d_desc* raw, float mobility, double tol, int64_t* out) {
  get_scan_indices_tolerance_4d(raw, mobility, tol, 16, out);
}


/* ==========================================================================================
 * timsTOF load-time transpose — alphadia/raw_data/bruker.py:202-274 (_transpose, _transpose_chunk)
 * push-major CSR (rows = pushes, columns = tof indices) -> tof-major CSR; the chunked scatter of the
 * reference visits the events in push order for every tof chunk, i.e. a stable counting sort by tof index.
 * ======================================================================================== */
int adbo_transpose_csr(int64_t n_values, int64_t n_push, int64_t n_tof, const uint32_t* tof_indices, const int64_t* push_indptr,
                       const uint16_t* values, uint32_t* push_indices_out, int64_t* tof_indptr_out, uint16_t* values_out) {
  uint32_t* count = (uint32_t*)calloc((size_t)(n_tof > 0 ? n_tof : 1), sizeof(uint32_t));
  for (int64_t i = 0; i < n_values; i++) { if ((int64_t)tof_indices[i] >= n_tof) { free(count); return 1; } count[tof_indices[i]]++; } /* :238-240 */
  tof_indptr_out[0] = 0;
  for (int64_t t = 0; t < n_tof; t++) tof_indptr_out[t + 1] = tof_indptr_out[t] + count[t]; /* :243-246 */
  memset(count, 0, sizeof(uint32_t) * (size_t)(n_tof > 0 ? n_tof : 1));
  for (int64_t p = 0; p < n_push; p++) /* :171-182 */
    for (int64_t idx = push_indptr[p]; idx < push_indptr[p + 1]; idx++) {
      uint32_t t = tof_indices[idx];
      int64_t dst = tof_indptr_out[t] + count[t];
      push_indices_out[dst] = (uint32_t)p;
      values_out[dst] = values[idx];
      count[t]++;
    }
  free(count);
  return 0;
}


/* ==========================================================================================
 * FDR bookkeeping — alphadia/fdr/fdr.py:195-297 (keep_best, _fdr_to_q_values, get_q_values)
 * pandas' multi-column sort_values is a stable lexicographic sort; restated as a bottom-up merge sort of row indices.
 * ======================================================================================== */
typedef struct { const double* score; const uint8_t* decoy; const uint64_t* key; int mode; } FdrCmp;
/* mode 0: (score, decoy, key) — get_q_values; mode 1: (key, score) — per-group order of keep_best */
static int fdr_less_equal(const FdrCmp* c, int64_t a, int64_t b) {
  if (c->mode == 0) {
    if (c->score[a] != c->score[b]) return c->score[a] < c->score[b];
    if (c->decoy[a] != c->decoy[b]) return c->decoy[a] < c->decoy[b];
    return c->key[a] <= c->key[b];
  }
  if (c->key[a] != c->key[b]) return c->key[a] < c->key[b];
  return c->score[a] <= c->score[b];
}
static void fdr_stable_sort(const FdrCmp* c, int64_t n, int64_t* idx) {
  int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++) idx[i] = i;
  for (int64_t w = 1; w < n; w *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * w) {
      int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b = mid, k = lo;
      while (a < mid && b < hi) tmp[k++] = fdr_less_equal(c, idx[a], idx[b]) ? idx[a++] : idx[b++]; /* ties: left first */
      while (a < mid) tmp[k++] = idx[a++];
      while (b < hi) tmp[k++] = idx[b++];
    }
    memcpy(idx, tmp, sizeof(int64_t) * (size_t)n);
  }
  free(tmp);
}

/* fdr.py:211-214 */
void adbo_fdr_to_q_values(const double* fdr, int64_t n, double* q) {
  double m = 0;
  for (int64_t i = n - 1; i >= 0; i--) { m = (i == n - 1 || fdr[i] < m) ? fdr[i] : m; q[i] = m; }
}

/* fdr.py:280-296: order_out = row index of each sorted row, qval_out in sorted order */
void adbo_q_values(int64_t n, const double* score, const uint8_t* decoy, const uint64_t* extra_key, int64_t* order_out, double* qval_out) {
  FdrCmp c = {score, decoy, extra_key, 0};
  fdr_stable_sort(&c, n, order_out);
  double* fdr = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  int64_t d = 0, t = 0;
  for (int64_t i = 0; i < n; i++) { d += decoy[order_out[i]]; t += 1 - decoy[order_out[i]]; fdr[i] = (double)d / (double)t; }
  adbo_fdr_to_q_values(fdr, n, qval_out);
  free(fdr);
}

/* fdr.py:219-224: sort by (score, group), first row of every group, back in original order */
void adbo_keep_best(int64_t n, const double* score, const uint64_t* group_key, uint8_t* keep_out) {
  FdrCmp c = {score, NULL, group_key, 1};
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  fdr_stable_sort(&c, n, idx);
  for (int64_t i = 0; i < n; i++) keep_out[i] = 0;
  for (int64_t i = 0; i < n; i++)
    if (i == 0 || group_key[idx[i]] != group_key[idx[i - 1]]) keep_out[idx[i]] = 1;
  free(idx);
}


/* ---- test hooks for the reference's remaining known-answer unit tests ------------------------------------------- */
/* scoring/utils.py:478-510 save_corrcoeff on float32 inputs (np.mean / np.sum accumulate in float32 under numba) */
double adbo_save_corrcoeff_f32(const float* x, const float* y, int n) {
  float sx = 0, sy = 0;
  for (int i = 0; i < n; i++) { sx = sx + x[i]; sy = sy + y[i]; }
  float xb = sx / (float)n, yb = sy / (float)n;
  float num = 0, sxx = 0, syy = 0;
  for (int i = 0; i < n; i++) { float a = x[i] - xb, b = y[i] - yb; num = num + a * b; sxx = sxx + a * a; syy = syy + b * b; }
  return (double)num / ((double)sqrtf(sxx * syy) + 1e-12);
}

/* scoring/utils.py:513-571 fragment_correlation: x [F][nobs][n] -> out [nobs][F][F], float32 like the scoring path */
void adbo_fragment_correlation(const float* x, int F, int nobs, int n, float* out) {
  float* cen = (float*)malloc(sizeof(float) * (size_t)F * (size_t)(n > 0 ? n : 1));
  float* stdv = (float*)malloc(sizeof(float) * (size_t)(F > 0 ? F : 1));
  for (int o = 0; o < nobs; o++) {
    for (int f = 0; f < F; f++) {
      const float* r = x + ((size_t)f * nobs + o) * n;
      float s = 0; for (int c = 0; c < n; c++) s = s + r[c];
      float m = s / (float)n, ss = 0;
      for (int c = 0; c < n; c++) { cen[(size_t)f * n + c] = r[c] - m; ss = ss + cen[(size_t)f * n + c] * cen[(size_t)f * n + c]; }
      stdv[f] = sqrtf(ss / (float)n);
    }
    for (int f = 0; f < F; f++)
      for (int g = 0; g < F; g++) {
        float dot = 0;
        for (int c = 0; c < n; c++) dot = dot + cen[(size_t)f * n + c] * cen[(size_t)g * n + c];
        out[((size_t)o * F + f) * F + g] = (float)((double)(dot / (float)n) / ((double)(stdv[f] * stdv[g]) + 1e-12));
      }
  }
  free(cen); free(stdv);
}

/* scoring/utils.py:574-647 fragment_correlation_different with one y profile per observation: out [nobs][F] */
void adbo_corr_with_template(const float* x, const float* y, int F, int nobs, int n, float* out) { corr_with_template(x, y, F, nobs, n, out); }

/* alpharaw_jit.py:53-75 _search_sorted_left / _search_sorted_reference_left */
int64_t adbo_search_sorted_left_f32(const float* a, int64_t n, float v) { return searchsorted_left_f32(a, n, v); }

// alphadia_b200 — candidate selection (3-D raw files), sm_100a.
//
// Replaces _select_candidates_pjit (alphadia/search/selection/selection.py:78-203) and everything below it:
// AlphaRawJIT.get_dense_intensity (jitclasses/alpharaw_jit.py:339-425), get_frame_indices
// (jitclasses/utils.py:24-88), convolve_fourier (selection/fft.py:141-212, as its defining circular
// convolution with fp64 FMA accumulation — see DESIGN.md), _build_features/_build_candidates
// (selection.py:206-226,367-526), find_peaks_1d / symetric_limits_2d (selection/utils.py:45-74,205-312).
//
// Precursors are visited in (quad window, RT) order so concurrently running warps hit the same index segment in L2.
//   adb_select_plan_kernel     warp per precursor: isotope m/z, fragment filter + m/z sort, RT window -> cycle
//                              window, quad windows, ppm windows -> a 40-byte plan + lo/hi rows in HBM
//   adb_select_fused_kernel    WARP PER PRECURSOR (cycle windows <= 1024, kernel width <= 32): every XIC row (layer) is
//                              extracted through the m/z-major index (one warp-cooperative 32-ary search + a scan of
//                              the few dozen peaks inside the ppm window, any cycle), lands in shared memory with its
//                              circular halo and a bit mask of its non-zero cycles, and is smoothed over the non-zero
//                              taps only (bit-identical to the dense fp64 sum); log(x + 1) goes into per-cell f32 layer
//                              sums; the warp then finds strict 5-point peaks, top-N (warp arg-max), suppresses close
//                              peaks, computes the symmetric limits and writes the candidates (integer-exact tail).
//   legacy pair                longer windows / wider kernels: adb_select_extract_kernel (warp per XIC row, same index
//                              extraction, rows to a dense HBM buffer) + adb_select_smooth_kernel (CTA per precursor,
//                              dense 30-tap x 2-row fp64 smoothing, 4 cells per thread).
#include <algorithm>
#include <cstdlib>

#include "adb_common.cuh"

#define FULL 0xffffffffu
#define SEL_MAX_LAYERS (ADB_MAX_LIB_FRAGMENTS + ADB_MAX_ISOTOPES)
#define SEL_MAX_CAND 16
#define SEL_PLAN_THREADS 128
#define SEL_EXTRACT_THREADS 256

namespace {

struct __align__(8) PrecPlan {
  long long frame_lo;
  int row, cs, C;
  unsigned char ok, nF, nI, nobs;
  short pos[ADB_MAX_OBS];
};

struct SelectParams {
  DevRaw raw;
  DevLib lib;
  adb_selection_config cfg;
  double kern[2 * ADB_MAX_KERNEL_W];  // [2][kw] Gaussian kernel as doubles (constant bank)
  int kw;
  DevCandidatesOut out;
  long long chunk_begin, chunk_n;  // positions [chunk_begin, chunk_begin + chunk_n) of the processing order
  const int32_t* order;            // processing order of library rows (may be null = identity)
  uint32_t* status;
  int c_cap, layer_cap;            // strides of the dense buffer
  PrecPlan* plan;                  // [chunk_n]
  float* win_lo;                   // [chunk_n][layer_cap]
  float* win_hi;                   // [chunk_n][layer_cap]
  float* dense;                    // [chunk_n][layer_cap][c_cap]
};

// what the peak picking / write-out needs of a precursor (per warp in the fused kernel, per CTA in the smoothing kernel)
struct SlotFinish {
  int C;
  long long frame_lo, row;
  double* score;   // [C]
};
// per-warp staging of the plan kernel (the only user of the per-fragment arrays: the fused kernel keeps 8 of these small
// SlotFinish records in shared memory instead, which leaves room for 5 CTAs per SM whatever ADB_MAX_LIB_FRAGMENTS is)
struct SlotMeta : SlotFinish {
  float lo[SEL_MAX_LAYERS], hi[SEL_MAX_LAYERS];
  float tmp_mz[ADB_MAX_LIB_FRAGMENTS];
  float iso_mz[ADB_MAX_ISOTOPES];
  int pos[ADB_MAX_OBS];
  int nF, nI, nobs, ok;
  long long cs;
};

__device__ __forceinline__ double limits_value(const double* a, int idx, int nrows) {
  // a[mobility_lower:mobility_upper, :].sum(axis=0) over nrows identical rows (0 + x [+ x])
  return nrows == 2 ? __dadd_rn(a[idx], a[idx]) : (nrows == 1 ? a[idx] : 0.0);
}

// selection/utils.py:205-273 on an implicit array v(i) = limits_value(a, i, nrows)
__device__ void symetric_limits_1d(const double* a, int nrows, int n, int center, double f, double cf, int min_size,
                                   int max_size, int out[2]) {
  if (n == 0 || center < 0 || center >= n) { out[0] = center; out[1] = center; return; }
  double center_intensity = limits_value(a, center, nrows), trailing = center_intensity;
  int limit = min_size;
  for (int s = min_size + 1; s < max_size; s++) {
    int l = max(center - s, 0), r = min(center + s, n - 1);
    double intensity = __dadd_rn(limits_value(a, l, nrows), limits_value(a, r, nrows)) / 2;
    if (intensity < __dmul_rn(f, trailing)) {
      if (intensity > __dmul_rn(center_intensity, cf)) { limit = s; trailing = intensity; }
      else break;
    } else break;
  }
  out[0] = max(center - limit, 0);
  out[1] = min(center + limit + 1, n);
}

// phase 0 for one slot, executed by one warp
__device__ void slot_setup(const SelectParams& P, SlotMeta& sl, int64_t i, int lane) {
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_selection_config& cfg = P.cfg;
  const int64_t L = raw.cycle_len;
  int nI = (int)min((long long)lib.n_isotopes, (long long)cfg.top_k_precursors);
  nI = min(nI, ADB_MAX_ISOTOPES);
  if (lane < nI) {  // selection/utils.py:35-40: float32 += float64
    double off = (double)lane * ADB_ISOTOPE_DIFF / (double)lib.charge[i];
    sl.iso_mz[lane] = (float)((double)lib.mz[i] + off);
  }
  const int64_t fs = lib.frag_start_idx[i], fe = lib.frag_stop_idx[i];
  int n_all = (int)max((long long)(fe - fs), 0LL);
  int ok = 1;
  if (n_all > ADB_MAX_LIB_FRAGMENTS) { if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_LIB_FRAGMENTS); ok = 0; n_all = 0; }
  int m = 0;
  for (int base = 0; base < n_all; base += 32) {  // selection.py:131-134
    int j = base + lane;
    bool keep = j < n_all && (!cfg.exclude_shared_ions || lib.frag_cardinality[fs + j] <= 1);
    unsigned b = __ballot_sync(FULL, keep);
    if (keep) sl.tmp_mz[m + __popc(b & ((1u << lane) - 1u))] = lib.frag_mz[fs + j];
    m += __popc(b);
  }
  __syncwarp();
  if (m <= 3) ok = 0;  // selection.py:136-137
  // stable ascending m/z + windows (jitclasses/utils.py:15-20 with float64 tolerances)
  for (int u = lane; u < m; u += 32) {
    float v = sl.tmp_mz[u];
    int rk = 0;
    for (int q = 0; q < m; q++) rk += (sl.tmp_mz[q] < v) || (sl.tmp_mz[q] == v && q < u);
    double mz = (double)v, d = cfg.fragment_mz_tolerance * mz / 1000000.0;
    sl.lo[rk] = (float)(mz - d);
    sl.hi[rk] = (float)(mz + d);
  }
  if (lane < nI) {
    double mz = (double)sl.iso_mz[lane], d = cfg.precursor_mz_tolerance * mz / 1000000.0;
    sl.lo[m + lane] = (float)(mz - d);
    sl.hi[m + lane] = (float)(mz + d);
  }
  // jitclasses/utils.py:24-88 + alpharaw_jit.py:173-203
  long long f0 = 0, f1 = 0;
  if (lane < 2) {
    float rt = lib.rt[i];
    float lim = (lane == 0) ? (float)((double)rt - cfg.rt_tolerance) : (float)((double)rt + cfg.rt_tolerance);
    f0 = adb_lower_bound(raw.rt_values, 0, raw.n_spectra, lim);
  }
  f1 = __shfl_sync(FULL, f0, 1);
  f0 = __shfl_sync(FULL, f0, 0);
  long long c0 = (f0 + raw.zeroth_frame) / L, c1 = (f1 + raw.zeroth_frame) / L;
  long long len = c1 - c0;
  long long opt = max(len, (long long)cfg.kernel_size);
  opt = (long long)(16.0 * ceil((double)opt / 16.0));
  long long l0 = c0, l1 = c0 + opt;
  const long long pcmi = raw.precursor_cycle_max_index;
  if (l1 > pcmi) {
    l1 = pcmi;
    l0 = pcmi - opt;
    if (l0 < 0) l0 = (pcmi % 2 == 0) ? 0 : 1;
  }
  const long long frame_lo = l0 * L + raw.zeroth_frame, frame_hi = l1 * L + raw.zeroth_frame;
  const long long cs = frame_lo / L;
  const long long C = frame_hi / L - cs;
  if (C <= 0 || (cs + C) * L > raw.n_spectra) ok = 0;
  if (C < P.kw) ok = 0;  // selection.py:61-73 (scan extent 2 >= kernel height 2)
  __syncwarp();
  const float q0 = sl.iso_mz[0], q1 = sl.iso_mz[max(nI - 1, 0)];  // selection.py:152
  int nobs = 0;
  for (int64_t base = 0; base < L; base += 32) {  // alpharaw_jit.py:19-50
    int64_t j = base + lane;
    bool hit = j < L && ((double)q0 <= raw.cycle[2 * j + 1]) && ((double)q1 >= raw.cycle[2 * j]);
    unsigned b = __ballot_sync(FULL, hit);
    if (hit) { int u = nobs + __popc(b & ((1u << lane) - 1u)); if (u < ADB_MAX_OBS) sl.pos[u] = (int)j; }
    nobs += __popc(b);
  }
  if (nobs > ADB_MAX_OBS) { if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_OBS); ok = 0; }
  if (lane == 0) {
    sl.nF = m; sl.nI = nI; sl.nobs = nobs; sl.C = (int)min(C, 2000000000LL);
    sl.frame_lo = frame_lo; sl.cs = cs; sl.ok = ok; sl.row = i;
  }
}

// phase 3 for one slot, executed by one warp
__device__ void slot_finish(const SelectParams& P, SlotFinish& sl, int lane) {
  const DevRaw& raw = P.raw;
  const adb_selection_config& cfg = P.cfg;
  const int64_t L = raw.cycle_len;
  const int C = sl.C;
  const double* a = sl.score;
  int t_cyc[SEL_MAX_CAND];
  double t_val[SEL_MAX_CAND];
  int top_n = 0;
  const int want = (int)min((long long)cfg.candidate_count, (long long)SEL_MAX_CAND);
  // selection/utils.py:45-74: top-N of the strict 5-point maxima; argsort(...)[::-1] of a stable sort puts
  // the LATER index first among equal values.  Each round is a warp arg-max over the not-yet-taken peaks.
  double last_v = 0;
  int last_p = 0;
  for (int r = 0; r < want; r++) {
    int best = -1;
    double bv = 0;
    for (int p = 2 + lane; p < C - 2; p += 32) {
      if (!(a[p - 2] < a[p - 1] && a[p - 1] < a[p] && a[p] > a[p + 1] && a[p + 1] > a[p + 2])) continue;
      double v = a[p];
      if (r > 0 && !(v < last_v || (v == last_v && p < last_p))) continue;  // already taken
      if (best < 0 || v > bv || (v == bv && p > best)) { best = p; bv = v; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      int ob = __shfl_xor_sync(FULL, best, off);
      double ov = __shfl_xor_sync(FULL, bv, off);
      if (ob >= 0 && (best < 0 || ov > bv || (ov == bv && ob > best))) { best = ob; bv = ov; }
    }
    if (best < 0) break;
    t_cyc[top_n] = best; t_val[top_n] = bv; top_n++;
    last_v = bv; last_p = best;
  }
  if (lane != 0) return;
  // selection.py:229-284 _join_close_peaks(3, 3); the scan index is always 0
  bool mask[SEL_MAX_CAND];
  for (int r = 0; r < top_n; r++) mask[r] = true;
  for (int x = 0; x < top_n; x++) {
    if (!mask[x]) continue;
    for (int y = x + 1; y < top_n; y++) {
      if (!mask[y]) continue;
      if (abs(t_cyc[x] - t_cyc[y]) <= 3) { if (t_val[x] > t_val[y]) mask[y] = false; else mask[x] = false; }
    }
  }
  int n_c = 0;
  for (int r = 0; r < top_n; r++) if (mask[r]) { t_cyc[n_c] = t_cyc[r]; t_val[n_c] = t_val[r]; n_c++; }
  // selection/utils.py:276-312 symetric_limits_2d on the (2, C) map with identical rows
  int slim[SEL_MAX_CAND][2], clim[SEL_MAX_CAND][2];
  for (int r = 0; r < n_c; r++) {
    const int scan_center = 0, cc = t_cyc[r];
    int ml = max(0, scan_center - (int)cfg.min_size_mobility), mu = min(2, scan_center + (int)cfg.min_size_mobility);
    int cl = max(0, cc - (int)cfg.min_size_rt), cu = min(C, cc + (int)cfg.min_size_rt);
    double ps[2];
    double t = 0;
    for (int c = cl; c < cu; c++) t = __dadd_rn(t, a[c]);
    ps[0] = t; ps[1] = t;
    symetric_limits_1d(ps, 1, 2, scan_center, cfg.f_mobility, cfg.center_fraction, (int)cfg.min_size_mobility,
                       (int)cfg.max_size_mobility, slim[r]);
    symetric_limits_1d(a, max(mu - ml, 0), C, cc, cfg.f_rt, cfg.center_fraction, (int)cfg.min_size_rt,
                       (int)cfg.max_size_rt, clim[r]);
  }
  if (cfg.join_close_candidates) {  // selection.py:287-364
    bool jm[SEL_MAX_CAND];
    for (int r = 0; r < n_c; r++) jm[r] = true;
    for (int x = 0; x < n_c; x++) {
      if (!jm[x]) continue;
      for (int y = x + 1; y < n_c; y++) {
        if (!jm[y]) continue;
        double cycle_len = (double)(clim[x][1] - clim[x][0]);
        double cycle_overlap = (double)(min(clim[x][1], clim[y][1]) - max(clim[x][0], clim[y][0])) / cycle_len;
        double scan_len = (double)(slim[x][1] - slim[x][0]);
        double scan_overlap = (double)(min(slim[x][1], slim[y][1]) - max(slim[x][0], slim[y][0])) / scan_len;
        if (scan_overlap < 0 || cycle_overlap < 0) continue;
        if (cycle_overlap > cfg.join_close_candidates_cycle_threshold && scan_overlap > cfg.join_close_candidates_scan_threshold) {
          slim[x][0] = min(slim[x][0], slim[y][0]); slim[x][1] = max(slim[x][1], slim[y][1]);
          clim[x][0] = min(clim[x][0], clim[y][0]); clim[x][1] = max(clim[x][1], clim[y][1]);
          jm[y] = false;
        }
      }
    }
    int mm = 0;
    for (int r = 0; r < n_c; r++) if (jm[r]) {
      t_cyc[mm] = t_cyc[r]; t_val[mm] = t_val[r];
      slim[mm][0] = slim[r][0]; slim[mm][1] = slim[r][1]; clim[mm][0] = clim[r][0]; clim[mm][1] = clim[r][1]; mm++;
    }
    n_c = mm;
  }
  // selection.py:480-526 write-out
  const long long frame_lo = sl.frame_lo;
  const long long i = sl.row;
  for (int r = 0; r < n_c; r++) {
    long long row = i * cfg.candidate_count + r;
    if (row >= P.out.n_rows) break;
    P.out.precursor_idx[row] = P.lib.precursor_idx[i];
    P.out.rank[row] = (uint8_t)r;
    P.out.score[row] = (float)t_val[r];
    P.out.scan_center[row] = (uint32_t)adb_wrap0(0, raw.scan_max_index);
    P.out.scan_start[row] = (uint32_t)adb_wrap0(slim[r][0], raw.scan_max_index);
    P.out.scan_stop[row] = (uint32_t)adb_wrap0(slim[r][1], raw.scan_max_index);
    P.out.frame_center[row] = (uint32_t)adb_wrap0((long long)t_cyc[r] * L + frame_lo, raw.frame_max_index);
    P.out.frame_start[row] = (uint32_t)adb_wrap0((long long)clim[r][0] * L + frame_lo, raw.frame_max_index);
    P.out.frame_stop[row] = (uint32_t)adb_wrap0((long long)clim[r][1] * L + frame_lo, raw.frame_max_index);
  }
}


// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_PLAN_THREADS) adb_select_plan_kernel(const __grid_constant__ SelectParams P) {
  __shared__ SlotMeta slots[SEL_PLAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long it = (long long)blockIdx.x * (SEL_PLAN_THREADS / 32) + warp;
  if (it >= P.chunk_n) return;
  SlotMeta& sl = slots[warp];
  const long long pos_in_order = P.chunk_begin + it;
  const int64_t i = P.order ? (int64_t)P.order[pos_in_order] : (int64_t)pos_in_order;
  slot_setup(P, sl, i, lane);
  __syncwarp();
  int ok = sl.ok;
  if (ok && (sl.C > P.c_cap || sl.nF + sl.nI > P.layer_cap)) {  // cannot happen with the host-side bounds
    if (lane == 0) atomicOr(P.status, ADB_STATUS_SCRATCH_OVERFLOW);
    ok = 0;
  }
  if (lane == 0) {
    PrecPlan pl;
    pl.frame_lo = sl.frame_lo; pl.row = (int)sl.row; pl.cs = (int)sl.cs; pl.C = sl.C;
    pl.ok = (unsigned char)ok; pl.nF = (unsigned char)sl.nF; pl.nI = (unsigned char)sl.nI;
    pl.nobs = (unsigned char)min(sl.nobs, ADB_MAX_OBS);
    for (int o = 0; o < ADB_MAX_OBS; o++) pl.pos[o] = (short)((o < sl.nobs) ? sl.pos[o] : 0);
    P.plan[it] = pl;
  }
  if (ok) {
    const int nL = sl.nF + sl.nI;
    for (int k = lane; k < nL; k += 32) {
      P.win_lo[it * P.layer_cap + k] = sl.lo[k];
      P.win_hi[it * P.layer_cap + k] = sl.hi[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// XIC extraction (alpharaw_jit.py:398-423) through the m/z-major index: ONE WARP PER XIC ROW (precursor, layer).
// The reference walks every spectrum of the cycle window and binary-searches each m/z window in it; here the peaks of
// one cycle position are stored sorted by m/z across all cycles (DevRaw::s_pk records + the s_bucket table), so a row needs one
// warp-cooperative 32-ary search per observation and a scan over the few dozen peaks inside the m/z window; peaks whose
// cycle lies outside [cs, cs + C) are skipped.  Bit-exactness: a cell sums its peaks in ascending m/z order and its
// observations in ascending cycle-position order, exactly like the reference — the index is a STABLE sort, observations
// are processed one after the other, and peaks of one 32-peak chunk that fall into the same cell are added in lane
// order (match_any ranks).  The reference's forward-only cursor (a peak consumed by the previous, overlapping window is
// not seen again) becomes the extra lower bound mz > hi[k - 1].
#define SEL_ROWS_PER_CTA (SEL_EXTRACT_THREADS / 32)

__device__ __forceinline__ int64_t warp_lower_bound_window(const float4* __restrict__ pk, int64_t lo, int64_t hi, float v_lo, float prev_hi,
                                                           int lane) {
  // first index in [lo, hi] whose m/z is >= v_lo and > prev_hi ("before" = mz < v_lo || mz <= prev_hi, monotone)
  while (hi - lo > 32) {
    const int64_t n = hi - lo, step = (n + 31) >> 5;
    const int64_t last = lo + min((int64_t)(lane + 1) * step, n) - 1;  // last element of sub-block `lane`
    const float m = __ldg(&pk[last].x);
    const unsigned b = __ballot_sync(FULL, (m < v_lo) || (m <= prev_hi));
    const int c = __popc(b);  // leading sub-blocks that lie completely before the window
    const int64_t nlo = lo + min((int64_t)c * step, n);
    hi = (c < 32) ? min(lo + (int64_t)(c + 1) * step, hi) : hi;
    lo = nlo;
  }
  const int64_t i = lo + lane;
  float m = 3.0e38f;
  if (i < hi) m = __ldg(&pk[i].x);
  const unsigned b = __ballot_sync(FULL, i < hi && ((m < v_lo) || (m <= prev_hi)));
  return lo + __popc(b);
}

// accumulates XIC row k of the planned precursor `pl` into row[0 .. C) (zeroed by the caller), executed by one warp
__device__ __forceinline__ void extract_row(const SelectParams& P, const PrecPlan& pl, long long it, int k, float* row, int lane) {
  const DevRaw& raw = P.raw;
  const int nF = pl.nF, C = pl.C;
  const float lo = P.win_lo[it * P.layer_cap + k], hi = P.win_hi[it * P.layer_cap + k];
  float prev_hi = (k > 0 && k != nF) ? P.win_hi[it * P.layer_cap + k - 1] : -1.0f;
  if (!(prev_hi >= lo)) prev_hi = -1.0f;  // only an overlapping previous window moves the cursor past lo
  const bool ms1 = k >= nF;
  const int n_o = ms1 ? raw.n_ms1_pos : (int)pl.nobs;
  const uint32_t cs = (uint32_t)pl.cs;
  for (int o = 0; o < n_o; o++) {
    const int p = ms1 ? raw.ms1_pos[o] : (int)pl.pos[o];
    // bucket of the first wanted m/z (the previous window's upper edge when that window reaches into this one): the
    // first wanted peak lies inside the bucket or is the first peak of the next one
    const uint32_t* tab = raw.s_bucket + (size_t)p * (size_t)(raw.sb_nb + 1);
    const int bk = adb_bucket_of(raw.sb_lo, raw.sb_width, raw.sb_inv_width, raw.sb_nb, prev_hi >= lo ? prev_hi : lo);
    const int64_t seg1 = (int64_t)__ldg(tab + raw.sb_nb);
    int64_t idx = warp_lower_bound_window(raw.s_pk, (int64_t)__ldg(tab + bk), (int64_t)__ldg(tab + bk + 1), lo, prev_hi, lane);
    while (idx < seg1) {  // chunks of 32 peaks in ascending m/z
      const int64_t i = idx + lane;
      const float4 pk = __ldg(raw.s_pk + i);  // padded behind the last segment
      const float m = pk.x;
      const bool inw = i < seg1 && m <= hi;
      const uint32_t c = __float_as_uint(pk.z) - cs;
      const bool take = inw && c < (uint32_t)C;
      const unsigned tb = __ballot_sync(FULL, take);
      if (tb) {
        const float v = pk.y;
        const unsigned grp = __match_any_sync(FULL, take ? c : 0xFFFFFF00u + (uint32_t)lane);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        const unsigned multi = __ballot_sync(FULL, take && rank > 0);
        if (take && rank == 0) row[c] = __fadd_rn(row[c], v);
        if (multi) {  // several peaks of this chunk fall into one cell: add them in lane (= m/z) order
          int max_rank = rank;
          for (int off = 16; off > 0; off >>= 1) max_rank = max(max_rank, __shfl_xor_sync(FULL, max_rank, off));
          for (int r = 1; r <= max_rank; r++) {
            __syncwarp();
            if (take && rank == r) row[c] = __fadd_rn(row[c], v);
          }
        }
        __syncwarp();
      }
      if (__ballot_sync(FULL, inw) != FULL) break;  // the window ended inside this chunk
      idx += 32;
    }
    __syncwarp();
  }
}

// legacy pair (cycle windows too long for the fused kernel): rows to the dense HBM buffer, then adb_select_smooth_kernel
__global__ void __launch_bounds__(SEL_EXTRACT_THREADS) adb_select_extract_kernel(const __grid_constant__ SelectParams P) {
  __shared__ float s_row[SEL_ROWS_PER_CTA][1024];  // rows longer than 1024 cycles accumulate in the dense buffer itself
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_rows = P.chunk_n * (long long)P.layer_cap;
  for (long long task = (long long)blockIdx.x * SEL_ROWS_PER_CTA + warp; task < n_rows; task += (long long)gridDim.x * SEL_ROWS_PER_CTA) {
    const long long it = task / P.layer_cap;
    const int k = (int)(task - it * P.layer_cap);
    const PrecPlan pl = P.plan[it];
    const int nL = (int)pl.nF + (int)pl.nI, C = pl.C;
    if (!pl.ok || k >= nL) continue;  // warp-uniform
    float* out = P.dense + (it * (long long)P.layer_cap + k) * P.c_cap;
    const bool in_smem = C <= 1024;
    float* row = in_smem ? s_row[warp] : out;
    for (int c = lane; c < C; c += 32) row[c] = 0.f;
    __syncwarp();
    extract_row(P, pl, it, k, row, lane);
    if (in_smem) {
      for (int c = lane; c < C; c += 32) out[c] = row[c];
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// smoothing of one layer at cycles c0 .. c0+3: out[c] = sum_a sum_b k[a][b] * x[(c + kw/2 - b) mod C], a then b
// ascending (fp64 FMA).  ext[t] = x[(t - off) mod C] with off = kw - 1 - kw/2, so x[(c + kw/2 - b) mod C] =
// ext[c + kw - 1 - b].  Four adjacent cells share one register window of KW + 3 doubles (four independent FMA chains).
template <int KW>
__device__ __forceinline__ void smooth_cells4(const SelectParams& P, const double* ext, int c0, int kw, float out[4]) {
  if (KW > 0) {
    double v[(KW > 0 ? KW : 1) + 3];
#pragma unroll
    for (int u = 0; u < KW + 3; u++) v[u] = ext[c0 + u];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < KW; b++) acc = fma(P.kern[a * KW + b], v[q + KW - 1 - b], acc);
      out[q] = (float)acc;
    }
  } else {
    for (int q = 0; q < 4; q++) {
      double acc = 0.0;
      for (int a = 0; a < 2; a++)
        for (int b = 0; b < kw; b++) acc = fma(P.kern[a * kw + b], ext[c0 + q + kw - 1 - b], acc);
      out[q] = (float)acc;
    }
  }
}

#define SMOOTH_GROUPS 2  // groups of 4 adjacent cycles per thread (cycle windows up to 8 * blockDim)

// dynamic shared memory: score doubles [c_cap] | two fp64 layer rows [c_cap + kw - 1 + 4]
template <int KW>
__global__ void adb_select_smooth_kernel(const __grid_constant__ SelectParams P) {
  extern __shared__ __align__(16) double dyn_d[];
  __shared__ SlotMeta sl;
  const adb_selection_config& cfg = P.cfg;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int kw = (KW > 0) ? KW : P.kw;
  const int off = kw - 1 - kw / 2;
  const int row_alloc = P.c_cap + kw - 1 + 4;
  double* score = dyn_d;
  double* buf0 = dyn_d + ((P.c_cap + 3) & ~3);
  double* buf1 = buf0 + ((row_alloc + 3) & ~3);
  for (long long it = blockIdx.x; it < P.chunk_n; it += gridDim.x) {
    const PrecPlan pl = P.plan[it];
    if (!pl.ok) continue;  // uniform for the CTA
    const int C = pl.C, nF = pl.nF, nL = (int)pl.nF + (int)pl.nI;
    const int stride = C + kw - 1;
    const float* dense = P.dense + it * (long long)P.layer_cap * P.c_cap;
    __syncthreads();  // previous precursor's tail (warp 0) is done with the shared state
    // layer 0 -> buf0 (with circular halo: ext[t] = x[(t - off) mod C]); the 4-double pad stays finite
    for (int t = tid; t < stride + 4; t += nthr) {
      int j = t - off; if (j < 0) j += C; else if (j >= C) j -= C;
      buf0[t] = (t < stride) ? (double)__ldg(dense + j) : 0.0;
      if (t >= stride) buf1[t] = 0.0;
    }
    __syncthreads();
    float lf_r[SMOOTH_GROUPS][4], lp_r[SMOOTH_GROUPS][4];
#pragma unroll
    for (int g = 0; g < SMOOTH_GROUPS; g++)
#pragma unroll
      for (int q = 0; q < 4; q++) { lf_r[g][q] = 0.f; lp_r[g][q] = 0.f; }
    for (int l = 0; l < nL; l++) {
      const double* cur = (l & 1) ? buf1 : buf0;
      double* nxt = (l & 1) ? buf0 : buf1;
      // prefetch the next layer into registers (global latency overlaps the smoothing below)
      float pre[4] = {0.f, 0.f, 0.f, 0.f};
      if (l + 1 < nL) {
        const float* row = dense + (long long)(l + 1) * P.c_cap;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = tid + u * nthr;
          if (t < stride) { int j = t - off; if (j < 0) j += C; else if (j >= C) j -= C; pre[u] = __ldg(row + j); }
        }
      }
#pragma unroll
      for (int g = 0; g < SMOOTH_GROUPS; g++) {
        const int c0 = 4 * (tid + g * nthr);
        if (c0 < C) {
          float sm4[4];
          smooth_cells4<KW>(P, cur, c0, kw, sm4);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float lg = (float)log((double)sm4[q] + 1.0);
            if (l < nF) lf_r[g][q] = __fadd_rn(lf_r[g][q], lg); else lp_r[g][q] = __fadd_rn(lp_r[g][q], lg);
          }
        }
      }
      if (l + 1 < nL) {
        const float* row = dense + (long long)(l + 1) * P.c_cap;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = tid + u * nthr;
          if (t < stride) nxt[t] = (double)pre[u];
        }
        for (int t = tid + 4 * nthr; t < stride; t += nthr) {
          int j = t - off; if (j < 0) j += C; else if (j >= C) j -= C;
          nxt[t] = (double)__ldg(row + j);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < SMOOTH_GROUPS; g++)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int c = 4 * (tid + g * nthr) + q;
        if (c < C) score[c] = (double)__fadd_rn(lf_r[g][q], lp_r[g][q]);  // raw feature (selection.py:221-224)
      }
    __syncthreads();
    if (warp == 0) {
      // normalisation (selection.py:405-428)
      double mean = cfg.use_weighted_score ? cfg.feature_mean : 0.0;
      double stdv = cfg.use_weighted_score ? cfg.feature_std : 0.0;
      const double wgt = cfg.use_weighted_score ? cfg.feature_weight : 1.0;
      if (!cfg.use_weighted_score) {  // amean1 / astd1 over the (2, C) feature map; sequential, rare path
        float accf = 0.f;
        for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) accf = __fadd_rn(accf, (float)score[c]);
        mean = (double)accf / (double)(2 * C);
        double v = 0;
        for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) { double d = score[c] - mean; v = __dadd_rn(v, __dmul_rn(d, d)); }
        stdv = sqrt(v / (double)(2 * C));
      }
      __syncwarp();
      for (int c = lane; c < C; c += 32) score[c] = 0.0 + __dmul_rn(wgt, (score[c] - mean)) / (stdv + 1e-6);
      if (lane == 0) { sl.C = C; sl.frame_lo = pl.frame_lo; sl.row = pl.row; sl.score = score; }
      __syncwarp();
      slot_finish(P, sl, lane);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Fused selection, ONE WARP PER PRECURSOR (cycle windows up to SEL_FUSED_MAX_C, kernel width <= 32), no CTA barriers:
// the warp walks the XIC rows (layers) of its precursor in order; a row is extracted through the m/z-major index
// straight into shared memory (with its circular halo), a bit mask marks its non-zero cycles, and the 2 x kw Gaussian
// is evaluated over the NON-ZERO taps only — fma(k, 0, acc) == acc exactly, so skipping them reproduces the dense fp64
// sum bit for bit, in the reference's tap order (kernel rows, then columns ascending = cycles descending).  > 97 % of
// the XIC cells are zero; most smoothed cells have no tap at all and cost one funnel shift.  log(x + 1) is added to the
// per-cell f32 layer sums as the rows come (layer order, selection.py:206-226); then the warp normalises, picks peaks
// and writes the candidates (slot_finish).
#ifndef SEL_FUSED_THREADS
#define SEL_FUSED_THREADS 256
#endif
#define SEL_FUSED_WARPS (SEL_FUSED_THREADS / 32)
#define SEL_FUSED_MAX_C 1024

struct FusedLayout { int kern_bytes, warp_bytes, score_off, lf_off, lp_off, ext_off, mask_off; size_t bytes; };
__host__ __device__ inline FusedLayout fused_layout(int c_cap, int kw) {
  FusedLayout f;
  f.kern_bytes = (int)(sizeof(double) * 2 * (size_t)kw);
  size_t b = 0;
  f.score_off = (int)b; b += sizeof(double) * (size_t)((c_cap + 3) & ~3);
  f.lf_off = (int)b; b += sizeof(float) * (size_t)((c_cap + 3) & ~3);
  f.lp_off = (int)b; b += sizeof(float) * (size_t)((c_cap + 3) & ~3);
  f.ext_off = (int)b; b += sizeof(float) * (size_t)((c_cap + kw + 3) & ~3);
  f.mask_off = (int)b; b += sizeof(uint32_t) * (size_t)(((c_cap + kw + 31) / 32 + 2 + 1) & ~1);
  f.warp_bytes = (int)b;
  f.bytes = (size_t)f.kern_bytes + (size_t)SEL_FUSED_WARPS * b + 16;
  return f;
}

__global__ void __launch_bounds__(SEL_FUSED_THREADS) adb_select_fused_kernel(const __grid_constant__ SelectParams P) {
  extern __shared__ __align__(16) unsigned char dyn_f[];
  __shared__ SlotFinish slots[SEL_FUSED_WARPS];
  const adb_selection_config& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kw = P.kw, off = kw - 1 - kw / 2;
  const FusedLayout fl = fused_layout(P.c_cap, kw);
  double* skern = (double*)dyn_f;  // [2][kw]; divergent tap indices would serialise in the constant bank
  unsigned char* wb = dyn_f + fl.kern_bytes + (size_t)warp * fl.warp_bytes;
  double* score = (double*)(wb + fl.score_off);
  float* lf = (float*)(wb + fl.lf_off);
  float* lp = (float*)(wb + fl.lp_off);
  float* ext = (float*)(wb + fl.ext_off);
  uint32_t* mask = (uint32_t*)(wb + fl.mask_off);
  SlotFinish& sl = slots[warp];
  for (int t = tid; t < 2 * kw; t += SEL_FUSED_THREADS) skern[t] = P.kern[t];
  __syncthreads();
  const uint32_t kw_mask = (kw >= 32) ? 0xFFFFFFFFu : ((1u << kw) - 1u);
  const long long n_warps = (long long)gridDim.x * SEL_FUSED_WARPS;
  for (long long it = (long long)blockIdx.x * SEL_FUSED_WARPS + warp; it < P.chunk_n; it += n_warps) {
    const PrecPlan pl = P.plan[it];
    if (!pl.ok) continue;  // warp-uniform
    const int C = pl.C, nF = pl.nF, nL = (int)pl.nF + (int)pl.nI;
    const int stride = C + kw - 1;  // ext[t] = x[(t - off) mod C], so x[(c + kw/2 - b) mod C] = ext[c + kw - 1 - b]
    const int n_words = (stride + 31) >> 5;
    for (int c = lane; c < C; c += 32) { lf[c] = 0.f; lp[c] = 0.f; }
    for (int k = 0; k < nL; k++) {
      for (int t = lane; t < stride + 1; t += 32) ext[t] = 0.f;
      __syncwarp();
      extract_row(P, pl, it, k, ext + off, lane);
      for (int h = lane; h < kw - 1; h += 32) {  // circular halo: kw - 1 cells (C >= kw is guaranteed by the plan)
        const int t = h < off ? h : h + C;
        ext[t] = h < off ? ext[t + C] : ext[t - C];
      }
      __syncwarp();
      unsigned any_nz = 0;
      for (int w = 0; w <= n_words; w++) {
        const int t = w * 32 + lane;
        const unsigned b = __ballot_sync(FULL, t < stride && ext[t] != 0.f);
        if (lane == 0) mask[w] = b;
        any_nz |= b;
      }
      __syncwarp();
      if (!any_nz) continue;  // empty row: every smoothed cell is 0, log(0 + 1) adds nothing
      float* lacc = (k < nF) ? lf : lp;
      for (int c = lane; c < C; c += 32) {
        const uint32_t w = __funnelshift_r(mask[c >> 5], mask[(c >> 5) + 1], c & 31) & kw_mask;  // taps t = c .. c + kw - 1
        if (w) {
          double acc = 0.0;
#pragma unroll 1
          for (int a = 0; a < 2; a++) {
            const double* kr = skern + a * kw + (kw - 1);
            uint32_t ww = w;
            while (ww) {  // b ascending = t descending
              const int hb = 31 - __clz(ww);
              ww ^= 1u << hb;
              acc = fma(kr[-hb], (double)ext[c + hb], acc);
            }
          }
          const float sm = (float)acc;
          if (sm != 0.f) lacc[c] = __fadd_rn(lacc[c], (float)log((double)sm + 1.0));
        }
      }
      __syncwarp();
    }
    for (int c = lane; c < C; c += 32) score[c] = (double)__fadd_rn(lf[c], lp[c]);  // raw feature (selection.py:221-224)
    __syncwarp();
    // normalisation (selection.py:405-428)
    double mean = cfg.use_weighted_score ? cfg.feature_mean : 0.0;
    double stdv = cfg.use_weighted_score ? cfg.feature_std : 0.0;
    const double wgt = cfg.use_weighted_score ? cfg.feature_weight : 1.0;
    if (!cfg.use_weighted_score) {  // amean1 / astd1 over the (2, C) feature map; sequential, rare path
      float accf = 0.f;
      for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) accf = __fadd_rn(accf, (float)score[c]);
      mean = (double)accf / (double)(2 * C);
      double v = 0;
      for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) { double d = score[c] - mean; v = __dadd_rn(v, __dmul_rn(d, d)); }
      stdv = sqrt(v / (double)(2 * C));
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) score[c] = 0.0 + __dmul_rn(wgt, (score[c] - mean)) / (stdv + 1e-6);
    if (lane == 0) { sl.C = C; sl.frame_lo = pl.frame_lo; sl.row = pl.row; sl.score = score; }
    __syncwarp();
    slot_finish(P, sl, lane);
    __syncwarp();
  }
}

size_t smooth_smem_bytes(int c_cap, int kw) { return sizeof(double) * ((size_t)((c_cap + 3) & ~3) + 2 * (size_t)((c_cap + kw - 1 + 4 + 3) & ~3)) + 16; }

template <int KW>
void launch_smooth(const SelectParams& P, int threads, long long grid, size_t dyn, cudaStream_t stream) {
  cudaFuncSetAttribute(adb_select_smooth_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  adb_select_smooth_kernel<KW><<<(unsigned)grid, threads, dyn, stream>>>(P);
}

}  // namespace

// the fused kernel needs its rows in shared memory and a kernel width that fits one 32-bit tap mask
bool adb_select_fused(int c_cap, int max_layers, int kw) {
  if (getenv("ADB_SELECT_LEGACY")) return false;  // test hook: force the extract + dense-smoothing pair
  return kw <= 32 && c_cap <= SEL_FUSED_MAX_C && fused_layout(c_cap, kw).bytes + SEL_FUSED_WARPS * sizeof(SlotFinish) + 1024 <= 200 * 1024;
}

size_t adb_select_bytes_per_precursor(int c_cap, int max_layers, int kw) {
  const size_t base = sizeof(PrecPlan) + 2 * sizeof(float) * (size_t)max_layers;
  if (adb_select_fused(c_cap, max_layers, kw)) return base;
  return base + sizeof(float) * (size_t)max_layers * (size_t)c_cap;
}

// Runs the selection for positions [chunk_begin, chunk_begin + chunk_n) of the processing order.
// `workspace` must hold chunk_n * adb_select_bytes_per_precursor(c_cap, max_layers, kw) bytes (+ 1024 for alignment).
void adb_launch_select_chunk(const DevRaw& raw, const DevLib& lib, const adb_selection_config& cfg, const double* h_kernel,
                             int kw, DevCandidatesOut out, int64_t chunk_begin, int64_t chunk_n, const int32_t* d_order,
                             uint32_t* d_status, int c_cap, int max_layers, void* workspace, int sm_count,
                             cudaStream_t stream, int* n_launches) {
  if (chunk_n <= 0) return;
  const bool fused = adb_select_fused(c_cap, max_layers, kw);
  SelectParams P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.kw = kw; P.out = out;
  for (int t = 0; t < 2 * ADB_MAX_KERNEL_W; t++) P.kern[t] = (t < 2 * kw) ? h_kernel[t] : 0.0;
  P.chunk_begin = chunk_begin; P.chunk_n = chunk_n; P.order = d_order; P.status = d_status;
  P.c_cap = c_cap; P.layer_cap = max_layers;
  char* w = (char*)workspace;
  auto take = [&](size_t bytes) { char* r = w; w += (bytes + 255) & ~(size_t)255; return r; };
  P.plan = (PrecPlan*)take(sizeof(PrecPlan) * (size_t)chunk_n);
  P.win_lo = (float*)take(sizeof(float) * (size_t)chunk_n * max_layers);
  P.win_hi = (float*)take(sizeof(float) * (size_t)chunk_n * max_layers);
  P.dense = fused ? nullptr : (float*)take(sizeof(float) * (size_t)chunk_n * max_layers * c_cap);

  const int warps_per_block = SEL_PLAN_THREADS / 32;
  adb_select_plan_kernel<<<(unsigned)((chunk_n + warps_per_block - 1) / warps_per_block), SEL_PLAN_THREADS, 0, stream>>>(P);

  if (fused) {
    const size_t dyn = fused_layout(c_cap, kw).bytes;
    cudaFuncSetAttribute(adb_select_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adb_select_fused_kernel, SEL_FUSED_THREADS, dyn);
    if (per_sm < 1) per_sm = 1;
    const long long grid = std::min<long long>((chunk_n + SEL_FUSED_WARPS - 1) / SEL_FUSED_WARPS, (long long)sm_count * per_sm);  // persistent, warp per precursor
    adb_select_fused_kernel<<<(unsigned)grid, SEL_FUSED_THREADS, dyn, stream>>>(P);
    if (n_launches) (*n_launches) += 2;
    return;
  }
  const long long rows = chunk_n * (long long)max_layers;  // warp per XIC row, grid-stride beyond the resident wave
  long long blocks = std::min<long long>((rows + SEL_ROWS_PER_CTA - 1) / SEL_ROWS_PER_CTA, (long long)sm_count * 16);
  adb_select_extract_kernel<<<(unsigned)blocks, SEL_EXTRACT_THREADS, 0, stream>>>(P);

  int threads = (((c_cap + 3) / 4 + 31) / 32) * 32;  // 4 adjacent cycles per thread
  if (threads > 512) threads = 512;
  if (threads < 32) threads = 32;
  const size_t dyn = smooth_smem_bytes(c_cap, kw);
  long long grid = chunk_n;
  if (grid > (long long)sm_count * 64) grid = (long long)sm_count * 64;
  if (kw == 30) launch_smooth<30>(P, threads, grid, dyn, stream);
  else launch_smooth<0>(P, threads, grid, dyn, stream);
  if (n_launches) (*n_launches) += 3;
}

// alphadia_b200 — candidate selection on timsTOF (4-D) raw files, sm_100a.
//
// Replaces _select_candidates_pjit (alphadia/search/selection/selection.py:78-203) for data with ion mobility:
// TimsTOFTransposeJIT.get_dense_intensity -> _assemble_push_intensity (jitclasses/bruker_jit.py:617-645,506-584),
// get_frame_indices / get_scan_indices_tolerance (jitclasses/utils.py:24-88, bruker_jit.py:204-271),
// convolve_fourier (selection/fft.py:141-212, as its defining circular convolution with fp64 FMA accumulation —
// see DESIGN.md), _build_features / _build_candidates (selection.py:206-226,367-526), find_peaks_2d and
// symetric_limits_2d (selection/utils.py:77-110,205-312).
//
// One persistent CTA per precursor (grid = SMs x resident CTAs, precursors visited in (quad window, RT) order):
//   setup     warp 0: isotope m/z, fragment filter + m/z sort, ppm windows -> tof-row ranges (searchsorted on the
//             fp64 m/z grid), RT window -> cycle window, mobility window -> scan window; then all threads build the
//             quadrupole masks of the (frame-in-cycle, scan) positions of the window.
//   per layer (12 fragments + 3 isotopes), the dense (scan x cycle) XIC tile lives in SHARED MEMORY:
//     extract   warp per tof row: two binary searches bound the push range of the frame window, lanes stride over
//               the events, u32 atomicAdd into the tile.  Detector intensities are integers (u16), so the integer
//               sum equals the reference's sequential f32 sum whenever it is < 2^24 (every partial sum is exact);
//               the rare cell beyond that is recomputed sequentially in the reference's order.
//     smooth    the tile is > 99 % zeros.  fma(k, 0, acc) == acc exactly, so the 30 x 30 circular Gaussian is
//               evaluated over the NON-ZERO inputs only, in the reference's summation order (kernel rows, then
//               kernel columns, ascending): the tile rows carry their circular halo and a bit mask of their non-zero
//               columns; an output cell visits only the non-empty rows inside its window and, per row, the set bits
//               of its 30-column tap window (one funnel shift).  fp64 FMA, one rounding to f32, log(x + 1) in fp64
//               rounded to f32, f32 layer sums — only for outputs that are non-zero.
//   score     (x - mean) / (std + 1e-6) * w in fp64; strict 5-point maxima in both axes -> peak list; top-N by a
//             block arg-max with the reference's tie order; close-peak suppression, symmetric limits, optional
//             join of overlapping candidates, clamped write-out (integer-exact tail).
#include <algorithm>

#include "adb_common.cuh"

#define FULL 0xffffffffu
#ifndef S4_THREADS
#define S4_THREADS 256
#endif
#define S4_WARPS (S4_THREADS / 32)
#define S4_MAX_LAYERS (ADB_MAX_LIB_FRAGMENTS + ADB_MAX_ISOTOPES)
#define S4_MAX_CAND 16
#define S4_EVR_CAP 512            // cached (layer, tof row) event ranges per precursor; rows beyond are searched in place
#ifndef S4_RB
#define S4_RB 4                  // output rows per lane in the register-blocked smoothing
#endif
#ifndef S4_MIN_CTAS
#define S4_MIN_CTAS 3
#endif
#define S4_BLOCKED_MIN_ROWS 4   // layers with at least this many non-empty rows take the blocked path

namespace {

struct Select4Params {
  DevRaw4 raw;
  DevLib lib;
  adb_selection_config cfg;
  const double* kern;  // [kh][kw] in HBM, staged to shared memory per CTA
  int kh, kw;
  DevCandidatesOut out;
  long long n;
  const int32_t* order;
  uint32_t* status;
  int s_cap, c_cap, tile_in_smem;
  char* ws;
  unsigned long long ws_per_cta;
};

struct Sel4State {
  float lo[S4_MAX_LAYERS], hi[S4_MAX_LAYERS];
  int t0[S4_MAX_LAYERS], t1[S4_MAX_LAYERS];
  float tmp_mz[ADB_MAX_LIB_FRAGMENTS];
  float iso_mz[ADB_MAX_ISOTOPES];
  int nF, nI, C, S, ok;
  long long f0, f1, s0, s1, cs, row;
  int n_peaks, overflow, n_nzrows;
  int rowoff[S4_MAX_LAYERS + 1];  // prefix sums of the tof-row counts of the layers
  uint2 evr[S4_EVR_CAP];          // [first, last) event of every (layer, tof row) inside the frame window
  int red_idx[S4_WARPS];
  double red_val[S4_WARPS];
  int top_idx[S4_MAX_CAND];
  double top_val[S4_MAX_CAND];
  int top_n;
  double norm_mean, norm_std;
};

// setup of one precursor, executed by warp 0
__device__ void setup4(const Select4Params& P, Sel4State& st, int64_t i, int lane) {
  const DevRaw4& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_selection_config& cfg = P.cfg;
  int nI = (int)min((long long)lib.n_isotopes, (long long)cfg.top_k_precursors);
  nI = min(nI, ADB_MAX_ISOTOPES);
  if (lane < nI) {  // selection/utils.py:35-40: float32 += float64
    double off = (double)lane * ADB_ISOTOPE_DIFF / (double)lib.charge[i];
    st.iso_mz[lane] = (float)((double)lib.mz[i] + off);
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
    if (keep) st.tmp_mz[m + __popc(b & ((1u << lane) - 1u))] = lib.frag_mz[fs + j];
    m += __popc(b);
  }
  __syncwarp();
  if (m <= 3) ok = 0;  // selection.py:136-137
  for (int u = lane; u < m; u += 32) {  // stable ascending m/z + windows (jitclasses/utils.py:15-20, float64 tolerance)
    float v = st.tmp_mz[u];
    int rk = 0;
    for (int q = 0; q < m; q++) rk += (st.tmp_mz[q] < v) || (st.tmp_mz[q] == v && q < u);
    double mz = (double)v, d = cfg.fragment_mz_tolerance * mz / 1000000.0;
    st.lo[rk] = (float)(mz - d);
    st.hi[rk] = (float)(mz + d);
  }
  if (lane < nI) {
    double mz = (double)st.iso_mz[lane], d = cfg.precursor_mz_tolerance * mz / 1000000.0;
    st.lo[m + lane] = (float)(mz - d);
    st.hi[m + lane] = (float)(mz + d);
  }
  // frame window: jitclasses/utils.py:24-88 on the float64 rt_values (bruker_jit.py:172-202)
  long long fi = 0;
  if (lane < 2) {
    float rt = lib.rt[i];
    float lim = (lane == 0) ? (float)((double)rt - cfg.rt_tolerance) : (float)((double)rt + cfg.rt_tolerance);
    fi = adb_lower_bound_f64(raw.rt_values, raw.n_frames, (double)lim);
  }
  const long long fi1 = __shfl_sync(FULL, fi, 1), fi0 = __shfl_sync(FULL, fi, 0);
  const long long L = raw.Fr, z = raw.zeroth_frame;
  long long c0 = (fi0 + z) / L, c1 = (fi1 + z) / L;
  long long opt = max(c1 - c0, (long long)cfg.kernel_size);
  opt = (long long)(16.0 * ceil((double)opt / 16.0));
  long long l0 = c0, l1 = c0 + opt;
  const long long pcmi = raw.precursor_cycle_max_index;
  if (l1 > pcmi) {
    l1 = pcmi;
    l0 = pcmi - opt;
    if (l0 < 0) l0 = (pcmi % 2 == 0) ? 0 : 1;
  }
  const long long f0 = l0 * L + z, f1 = l1 * L + z;
  // scan window: bruker_jit.py:204-271, searchsorted(mobility_values[::-1], v, "right")
  long long si = 0;
  if (lane < 2) {
    float mob = lib.mobility[i];
    float lim = (lane == 0) ? (float)((double)mob + cfg.mobility_tolerance) : (float)((double)mob - cfg.mobility_tolerance);
    const double v = (double)lim;
    const long long n = raw.Sc;
    long long lo = 0, hi = n;
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (raw.mobility_values[n - 1 - mid] <= v) lo = mid + 1; else hi = mid;
    }
    si = raw.scan_max_index - lo;
  }
  const long long si1 = __shfl_sync(FULL, si, 1), si0 = __shfl_sync(FULL, si, 0);
  const long long scan_len = si0 - si1;
  const long long sopt = (long long)(16.0 * ceil((double)scan_len / 16.0));
  long long sl0 = si0, sl1 = si0 - sopt;
  if (sl1 < 0) { sl1 = 0; sl0 = sopt; if (sl0 > raw.scan_max_index) sl0 = raw.scan_max_index; }
  const long long cs = (f0 - z) / L;
  const long long C = (f1 - z) / L - cs;
  const long long S = sl1 - sl0;
  if (C <= 0 || S <= 0 || sl0 < 0 || sl1 > raw.scan_max_index) ok = 0;
  if (ok && ((S % 2) != 0 || S < P.kh || C < P.kw)) ok = 0;  // selection.py:40-75 _is_valid
  if (ok && (S > P.s_cap || C > P.c_cap)) { if (lane == 0) atomicOr(P.status, ADB_STATUS_SCRATCH_OVERFLOW); ok = 0; }
  __syncwarp();
  // tof slices: searchsorted(mz_values f64, window, "left") (bruker_jit.py:273-278,596-598)
  if (ok)
    for (int k = lane; k < m + nI; k += 32) {
      st.t0[k] = (int)adb_lower_bound_f64(raw.mz_values, raw.n_tof, (double)st.lo[k]);
      st.t1[k] = (int)adb_lower_bound_f64(raw.mz_values, raw.n_tof, (double)st.hi[k]);
    }
  if (lane == 0) {
    st.nF = m; st.nI = nI; st.C = (int)min(C, 2000000000LL); st.S = (int)min(max(S, 0LL), 2000000000LL);
    st.f0 = f0; st.f1 = f1; st.s0 = sl0; st.s1 = sl1; st.cs = cs; st.ok = ok; st.row = i;
  }
}

// reference-order f32 sum of one XIC cell (bruker_jit.py:555-582): tof rows ascending, events ascending
__device__ __noinline__ float seq_cell_sum(const DevRaw4& raw, const Sel4State& st, int l, int rs, int rc, unsigned bit,
                                           const unsigned char* smask) {
  const uint32_t smi = (uint32_t)raw.scan_max_index;
  const long long p_lo = st.f0 * (long long)smi, p_hi = st.f1 * (long long)smi;
  const int S = st.S;
  float acc = 0.f;
  for (int t = st.t0[l]; t < st.t1[l]; t++) {
    const int64_t r1 = raw.tof_indptr[t + 1];
    for (int64_t e = adb_row_lower_bound(raw.push, raw.tof_indptr[t], r1, p_lo); e < r1; e++) {
      const uint32_t push = raw.push[e];
      if ((long long)push >= p_hi) break;
      const uint32_t frame = push / smi, scan = push - frame * smi;
      if ((long long)scan - st.s0 != rs) continue;
      const uint32_t fz = frame - (uint32_t)raw.zeroth_frame;
      const uint32_t cyc = fz / (uint32_t)raw.Fr, fic = fz - cyc * (uint32_t)raw.Fr;
      if ((long long)cyc - st.cs != rc) continue;
      if (!(smask[fic * S + rs] & bit)) continue;
      acc = __fadd_rn(acc, (float)raw.intensity[e]);
    }
  }
  return acc;
}

// on-demand projections for symetric_limits_2d (selection/utils.py:276-312)
__device__ double proj_scan(const double* a, int C, int s, int cl, int cu) {
  double t = 0;
  for (int c = cl; c < cu; c++) t = __dadd_rn(t, a[s * C + c]);
  return t;
}
__device__ double proj_cycle(const double* a, int C, int c, int ml, int mu) {
  double t = 0;
  for (int s = ml; s < mu; s++) t = __dadd_rn(t, a[s * C + c]);
  return t;
}

// selection/utils.py:205-273 with the projection evaluated on demand (axis 0 = scans, 1 = cycles)
__device__ void sym_limits_1d(const double* a, int C, int axis, int lo_o, int hi_o, int n, int center, double f, double cf,
                              int min_size, int max_size, int out[2]) {
  if (n == 0 || center < 0 || center >= n) { out[0] = center; out[1] = center; return; }
  auto val = [&](int idx) { return axis == 0 ? proj_scan(a, C, idx, lo_o, hi_o) : proj_cycle(a, C, idx, lo_o, hi_o); };
  const double center_intensity = val(center);
  double trailing = center_intensity;
  int limit = min_size;
  for (int s = min_size + 1; s < max_size; s++) {
    int l = max(center - s, 0), r = min(center + s, n - 1);
    double intensity = __dadd_rn(val(l), val(r)) / 2;
    if (intensity < __dmul_rn(f, trailing)) {
      if (intensity > __dmul_rn(center_intensity, cf)) { limit = s; trailing = intensity; }
      else break;
    } else break;
  }
  out[0] = max(center - limit, 0);
  out[1] = min(center + limit + 1, n);
}

__global__ void __launch_bounds__(S4_THREADS, S4_MIN_CTAS) adb_select4d_kernel(const __grid_constant__ Select4Params P) {
  extern __shared__ __align__(16) unsigned char dyn4[];
  __shared__ Sel4State st;
  const DevRaw4& raw = P.raw;
  const adb_selection_config& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kh = P.kh, kw = P.kw, sh = kh / 2, sw = kw / 2;
  const size_t cells_cap = (size_t)P.s_cap * P.c_cap;
  // shared memory carve-up
  double* s_kern = (double*)dyn4;
  unsigned char* sp = dyn4 + sizeof(double) * (size_t)kh * kw;
  // tile rows carry the circular halo: trow[t] = x[(t - off) mod C], t in [0, C + kw - 1), row stride cx_cap
  const int off = kw - 1 - sw;
  const int cx_cap = P.c_cap + kw;
  const int mw_cap = (cx_cap + 31) / 32 + 1;  // mask words per row (+1 so the funnel shift may read one word beyond)
  uint32_t* tile = (uint32_t*)sp;
  if (P.tile_in_smem) sp += sizeof(uint32_t) * (size_t)P.s_cap * cx_cap;
  uint32_t* rmask = (uint32_t*)sp; sp += sizeof(uint32_t) * (size_t)P.s_cap * mw_cap;  // non-zero columns of every tile row
  unsigned short* rowcnt = (unsigned short*)sp; sp += sizeof(unsigned short) * (size_t)((P.s_cap + 7) & ~7);
  unsigned short* nzrows = (unsigned short*)sp; sp += sizeof(unsigned short) * (size_t)((P.s_cap + 7) & ~7);  // non-empty rows, ascending
  unsigned short* rowrank = (unsigned short*)sp; sp += sizeof(unsigned short) * (size_t)((P.s_cap + 7) & ~7);  // # non-empty rows <= r
  uint32_t* rowtouch = (uint32_t*)sp; sp += sizeof(uint32_t) * (size_t)(((P.s_cap + 31) / 32 + 1) & ~1);      // rows hit by the current layer
  unsigned char* smask = sp;  // [Fr][S] bit 0: fragment quad window, bit 1: MS1
  // HBM workspace of this CTA
  char* wp = P.ws + (size_t)blockIdx.x * P.ws_per_cta;
  float* lf = (float*)wp; wp += sizeof(float) * cells_cap;
  float* lp = (float*)wp; wp += sizeof(float) * cells_cap;
  double* score = (double*)wp; wp += sizeof(double) * cells_cap;
  double* pk_val = (double*)wp; wp += sizeof(double) * (cells_cap / 2 + 8);
  int* pk_idx = (int*)wp; wp += sizeof(int) * (cells_cap / 2 + 8);
  if (!P.tile_in_smem) tile = (uint32_t*)wp;

  for (int t = tid; t < kh * kw; t += S4_THREADS) s_kern[t] = P.kern[t];
  for (long long t = tid; t < (long long)P.s_cap * cx_cap; t += S4_THREADS) tile[t] = 0u;  // rows are re-zeroed after use
  const uint32_t smi = (uint32_t)raw.scan_max_index;
  const uint32_t Fr = (uint32_t)raw.Fr, z = (uint32_t)raw.zeroth_frame;

  for (long long it = blockIdx.x; it < P.n; it += gridDim.x) {
    __syncthreads();  // the previous precursor's tail is done with the shared state
    if (warp == 0) {
      const int64_t i = P.order ? (int64_t)P.order[it] : (int64_t)it;
      setup4(P, st, i, lane);
    }
    __syncthreads();
    if (!st.ok) continue;  // uniform
    const int S = st.S, C = st.C, nF = st.nF, nL = st.nF + st.nI;
    const int cells = S * C;
    // quadrupole masks of the window (bruker_jit.py:280-313) + non-empty query test (:368, _is_valid)
    int any_f = 0, any_p = 0;
    {
      const double q0 = (double)st.iso_mz[0], q1 = (double)st.iso_mz[max(st.nI - 1, 0)];  // selection.py:152
      for (int t = tid; t < (int)Fr * S; t += S4_THREADS) {
        const int fic = t / S, rs = t - fic * S;
        const long long pos = (long long)fic * raw.Sc + st.s0 + rs;
        const double wlo = raw.cycle[2 * pos], whi = raw.cycle[2 * pos + 1];
        const bool has_id = raw.dia_precursor_cycle[pos] >= 0;
        const bool mf = (q0 <= whi) && (q1 >= wlo), mp = (-1.0 <= whi) && (-1.0 >= wlo);
        smask[t] = (unsigned char)((mf ? 1 : 0) | (mp ? 2 : 0));
        any_f |= (mf && has_id);
        any_p |= (mp && has_id);
      }
      if (tid == 0) { st.n_peaks = 0; st.overflow = 0; }
    }
    any_f = __syncthreads_or(any_f);
    any_p = __syncthreads_or(any_p);
    if (!any_f || !any_p) continue;  // empty push query -> 0-sized dense array -> _is_valid fails
    for (int t = tid; t < cells; t += S4_THREADS) { lf[t] = 0.f; lp[t] = 0.f; }

    const long long p_lo = st.f0 * (long long)smi, p_hi = st.f1 * (long long)smi;
    // every (layer, tof row) needs two binary searches on the push axis (dependent, mostly L2-missing loads): all of them
    // are done here at once, one per thread, instead of a ~6 us latency chain in front of every layer
    if (tid == 0) {
      int acc = 0;
      for (int l = 0; l < nL; l++) { st.rowoff[l] = acc; acc += max(st.t1[l] - st.t0[l], 0); }
      st.rowoff[nL] = acc;
    }
    __syncthreads();
    {
      const int n_rows_all = min(st.rowoff[nL], S4_EVR_CAP);
      for (int it2 = tid; it2 < 2 * n_rows_all; it2 += S4_THREADS) {
        const int idx = it2 >> 1, which = it2 & 1;
        int l = 0;
        while (l + 1 < nL && st.rowoff[l + 1] <= idx) l++;
        const int t = st.t0[l] + (idx - st.rowoff[l]);
        const int64_t r0 = __ldg(raw.tof_indptr + t), r1 = __ldg(raw.tof_indptr + t + 1);
        const uint32_t e = (uint32_t)adb_row_lower_bound(raw.push, r0, r1, which ? p_hi : p_lo);
        if (which) st.evr[idx].y = e; else st.evr[idx].x = e;
      }
    }
    __syncthreads();
    const int cx = C + kw - 1;            // used length of a tile row
    const int n_words = (cx + 31) >> 5;
    const uint32_t kw_mask = (kw >= 32) ? 0xFFFFFFFFu : ((1u << kw) - 1u);
    for (int l = 0; l < nL; l++) {
      for (int t = tid; t < S; t += S4_THREADS) rowcnt[t] = 0;
      for (int t = tid; t < (S + 31) / 32; t += S4_THREADS) rowtouch[t] = 0u;
      __syncthreads();
      // ---- XIC extraction (bruker_jit.py:506-584) ---------------------------------------------------
      const unsigned bit = (l < nF) ? 1u : 2u;
      for (int t = st.t0[l] + warp; t < st.t1[l]; t += S4_WARPS) {
        const int ridx = st.rowoff[l] + (t - st.t0[l]);
        long long e0, e1;
        if (ridx < S4_EVR_CAP) { e0 = st.evr[ridx].x; e1 = st.evr[ridx].y; }
        else {
          const int64_t r0 = __ldg(raw.tof_indptr + t), r1 = __ldg(raw.tof_indptr + t + 1);
          long long e = 0;
          if (lane < 2) e = adb_row_lower_bound(raw.push, r0, r1, lane == 0 ? p_lo : p_hi);
          e1 = __shfl_sync(FULL, e, 1); e0 = __shfl_sync(FULL, e, 0);
        }
        for (long long k = e0 + lane; k < e1; k += 32) {
          const uint32_t push = __ldg(raw.push + k);
          const uint32_t frame = push / smi, scan = push - frame * smi;
          const long long rs = (long long)scan - st.s0;
          if (rs < 0 || rs >= S) continue;
          const uint32_t fz = frame - z;
          const uint32_t cyc = fz / Fr, fic = fz - cyc * Fr;
          const long long rc = (long long)cyc - st.cs;
          if (rc < 0 || rc >= C) continue;
          if (!(smask[fic * S + (int)rs] & bit)) continue;
          atomicAdd(&tile[(int)rs * cx_cap + off + (int)rc], (uint32_t)__ldg(raw.intensity + k));
          const uint32_t rbit = 1u << ((int)rs & 31);
          if (!(rowtouch[(int)rs >> 5] & rbit)) atomicOr(&rowtouch[(int)rs >> 5], rbit);
        }
      }
      __syncthreads();
      // ---- integer sums -> f32 bit patterns, circular halo, per-row masks of the non-zero columns ----------
      for (int r = warp; r < S; r += S4_WARPS) {
        if (!((rowtouch[r >> 5] >> (r & 31)) & 1u)) continue;  // untouched rows are all zero; warp-uniform
        uint32_t* trow = tile + r * cx_cap;
        for (int base = 0; base < C; base += 32) {
          const int c = base + lane;
          const uint32_t v = (c < C) ? trow[off + c] : 0u;
          if (v != 0u) {
            const float fv = (v < (1u << 24)) ? (float)v : seq_cell_sum(raw, st, l, r, c, bit, smask);
            trow[off + c] = __float_as_uint(fv);
          }
        }
        __syncwarp();
        for (int t = lane; t < cx; t += 32)  // C >= kw is guaranteed by setup4 (_is_valid)
          if (t < off) trow[t] = trow[t + C]; else if (t >= off + C) trow[t] = trow[t - C];
        __syncwarp();
        unsigned any = 0;
        for (int w = 0; w <= n_words; w++) {
          const int t = w * 32 + lane;
          const unsigned b = __ballot_sync(FULL, t < cx && trow[t] != 0u);
          if (lane == 0) rmask[r * mw_cap + w] = b;
          any |= b;
        }
        if (lane == 0) rowcnt[r] = any ? 1 : 0;
      }
      __syncthreads();
      int M = 0;
      {  // ascending list of the non-empty tile rows + rank of every row; every warp writes the same values, so no
         // CTA barrier is needed before the warp reads them back
        for (int base = 0; base < S; base += 32) {
          const int r = base + lane;
          const bool ne = r < S && rowcnt[r] != 0;
          const unsigned b = __ballot_sync(FULL, ne);
          if (ne) nzrows[M + __popc(b & ((1u << lane) - 1u))] = (unsigned short)r;
          if (r < S) rowrank[r] = (unsigned short)(M + __popc(b & ((2u << lane) - 1u)));
          M += __popc(b);
        }
        __syncwarp();
      }
      // ---- sparse circular smoothing + log-sum (fft.py:141-212, selection.py:206-226) ----------------
      // out[i][j] = sum_a sum_b k[a][b] x[(i + sh - a) mod S][(j + sw - b) mod C], a then b ascending: for one output
      // cell the input rows are visited downwards (circularly) from r0 = (i + sh) mod S, only the non-empty ones, and
      // inside a row x[(j + sw - b) mod C] = trow[j + kw - 1 - b]: the set bits of the row mask in [j, j + kw), from
      // the highest (b = 0) down.
      float* lacc = (l < nF) ? lf : lp;
      const int segs = (C + 31) >> 5;
      if (M >= S4_BLOCKED_MIN_ROWS && kw <= 32) {
        // Dense blobs (an eluting peptide covers ~17 scans x 9 cycles): one lane owns an output column of S4_RB
        // consecutive output rows, so every tap (row r, column t) is loaded once and feeds up to S4_RB accumulators with
        // k[a_q][b], a_q = (r0_q - r) mod S.  Each accumulator must still see its taps in (a, b) ascending order: input
        // rows are swept downwards twice — first the rows at or below r0_q (a = r0_q - r), then the circularly wrapped
        // rows above it (a = r0_q - r + S) — and inside a row the taps from the highest set bit (b = 0) down.
        const int n_blk = (S + S4_RB - 1) / S4_RB;
        for (int item = warp; item < n_blk * segs; item += S4_WARPS) {  // segment-major: the warps share a blob's rows evenly
          const int seg = item / n_blk, blk = item - seg * n_blk;
          const int i0 = blk * S4_RB;
          int r0q[S4_RB], r0_min = 1 << 30, r0_max = -1;
#pragma unroll
          for (int q = 0; q < S4_RB; q++) {
            int r0 = i0 + q + sh;
            if (r0 >= S) r0 -= S;
            const bool live = i0 + q < S;
            r0q[q] = live ? r0 : -(1 << 20);  // dead rows never match: a < 0 in sweep 0, a < 0 in sweep 1
            if (live) { r0_min = min(r0_min, r0); r0_max = max(r0_max, r0); }
          }
          // list ranges of the two sweeps: rows in [r0_min - kh + 1, r0_max] and rows > r0_min + S - kh
          const int k_top0 = (int)rowrank[r0_max] - 1;
          const int r_low0 = r0_min - kh + 1, r_low1 = r0_min + S - kh;
          const bool has0 = k_top0 >= 0 && (int)nzrows[k_top0] >= r_low0;
          const bool has1 = (int)nzrows[M - 1] > r_low1;
          if (!has0 && !has1) continue;  // warp-uniform
          {
            const int j = min(seg * 32 + lane, C - 1);
            double acc[S4_RB];
#pragma unroll
            for (int q = 0; q < S4_RB; q++) acc[q] = 0.0;
            uint32_t anyw = 0u;
#pragma unroll 1
            for (int sweep = 0; sweep < 2; sweep++) {
              const int add = sweep ? S : 0;
              const int k_top = sweep ? M - 1 : k_top0;
              const int r_low = sweep ? r_low1 + 1 : r_low0;
#pragma unroll 1
              for (int k = k_top; k >= 0; k--) {
                const int r = nzrows[k];
                if (r < r_low) break;  // warp-uniform
                const double* kq[S4_RB];
                bool anyq = false;
#pragma unroll
                for (int q = 0; q < S4_RB; q++) {
                  const int a = r0q[q] - r + add;  // sweep 0: r <= r0_q, sweep 1: r > r0_q (then a >= S - ... < kh only if wrapped)
                  const bool ok = (unsigned)a < (unsigned)kh;
                  kq[q] = ok ? s_kern + a * kw + (kw - 1) : nullptr;
                  anyq |= ok;
                }
                if (!anyq) continue;  // warp-uniform
                const uint32_t* mrow = rmask + r * mw_cap + (j >> 5);
                uint32_t w = __funnelshift_r(mrow[0], mrow[1], j & 31) & kw_mask;  // taps t = j .. j + kw - 1
                anyw |= w;
                const uint32_t* trow = tile + r * cx_cap + j;
                while (w) {  // b ascending = t descending
                  const int hb = 31 - __clz(w);
                  w ^= 1u << hb;
                  const double v = (double)__uint_as_float(trow[hb]);
#pragma unroll
                  for (int q = 0; q < S4_RB; q++)
                    if (kq[q]) acc[q] = fma(kq[q][-hb], v, acc[q]);
                }
              }
            }
            if (anyw && seg * 32 + lane < C) {
#pragma unroll
              for (int q = 0; q < S4_RB; q++) {
                const float sm = (float)acc[q];
                if (i0 + q < S && sm != 0.f) {
                  const float lg = (float)log((double)sm + 1.0);
                  lacc[(i0 + q) * C + j] = __fadd_rn(lacc[(i0 + q) * C + j], lg);
                }
              }
            }
          }
        }
      } else if (M > 0)
        for (int i = warp; i < S; i += S4_WARPS) {
          int r0 = i + sh;
          if (r0 >= S) r0 -= S;
          const int k0 = rowrank[r0];
          {  // no non-empty row inside the window of this output row?
            int k = k0 - 1;
            if (k < 0) k += M;
            int a = r0 - (int)nzrows[k];
            if (a < 0) a += S;
            if (a >= kh) continue;
          }
          for (int seg = 0; seg < segs; seg++) {
            const int j = min(seg * 32 + lane, C - 1);  // lanes beyond the row compute a discarded duplicate of the last cell
            double acc = 0.0;
            bool any = false;
            for (int step = 0; step < M; step++) {
              int k = k0 - 1 - step;
              if (k < 0) k += M;
              const int r = nzrows[k];
              int a = r0 - r;
              if (a < 0) a += S;
              if (a >= kh) break;  // a grows along the sequence; warp-uniform
              const double* krow = s_kern + a * kw + (kw - 1);
              const uint32_t* trow = tile + r * cx_cap + j;
              const uint32_t* mrow = rmask + r * mw_cap + (j >> 5);
              if (kw <= 32) {
                uint32_t w = __funnelshift_r(mrow[0], mrow[1], j & 31) & kw_mask;  // taps t = j .. j + kw - 1
                any |= (w != 0u);
                while (w) {  // b ascending = t descending
                  const int hb = 31 - __clz(w);
                  w ^= 1u << hb;
                  acc = fma(krow[-hb], (double)__uint_as_float(trow[hb]), acc);
                }
              } else {
                for (int b = 0; b < kw; b++) {
                  const uint32_t v = trow[kw - 1 - b];
                  if (v != 0u) { any = true; acc = fma(krow[-(kw - 1 - b)], (double)__uint_as_float(v), acc); }
                }
              }
            }
            if (any && seg * 32 + lane < C) {
              const float sm = (float)acc;
              if (sm != 0.f) {
                const float lg = (float)log((double)sm + 1.0);
                lacc[i * C + j] = __fadd_rn(lacc[i * C + j], lg);
              }
            }
          }
        }
      __syncthreads();
      for (int r = warp; r < S; r += S4_WARPS) {  // leave the tile all-zero for the next layer / precursor
        if (!((rowtouch[r >> 5] >> (r & 31)) & 1u)) continue;
        uint32_t* trow = tile + r * cx_cap;
        for (int t = lane; t < cx; t += 32) trow[t] = 0u;
      }
      __syncthreads();
    }
    // ---- score normalisation (selection.py:401-428) ---------------------------------------------------
    if (!cfg.use_weighted_score) {  // amean1 / astd1 over the feature map, sequential (rare path)
      if (tid == 0) {
        float accf = 0.f;
        for (int t = 0; t < cells; t++) accf = __fadd_rn(accf, __fadd_rn(lf[t], lp[t]));
        const double mean = (double)accf / (double)cells;
        double v = 0;
        for (int t = 0; t < cells; t++) { double d = (double)__fadd_rn(lf[t], lp[t]) - mean; v = __dadd_rn(v, __dmul_rn(d, d)); }
        st.norm_mean = mean;
        st.norm_std = sqrt(v / (double)cells);
      }
      __syncthreads();
    }
    {
      const double mean = cfg.use_weighted_score ? cfg.feature_mean : st.norm_mean;
      const double stdv = cfg.use_weighted_score ? cfg.feature_std : st.norm_std;
      const double wgt = cfg.use_weighted_score ? cfg.feature_weight : 1.0;
      for (int t = tid; t < cells; t += S4_THREADS)
        score[t] = 0.0 + __dmul_rn(wgt, ((double)__fadd_rn(lf[t], lp[t]) - mean)) / (stdv + 1e-6);
    }
    __syncthreads();
    // ---- find_peaks_2d (selection/utils.py:77-110) -> peak list -------------------------------------------
    {
      const int inner_c = C - 4, inner_s = S - 4;
      const int pk_cap = (int)(cells_cap / 2);
      for (int t = tid; t < inner_c * inner_s; t += S4_THREADS) {
        const int s = 2 + t / inner_c, p = 2 + t % inner_c;
        const double* a = score + s * C + p;
        const double v = a[0];
        bool pk = a[-2 * C] < a[-C] && a[-C] < v && v > a[C] && a[C] > a[2 * C];
        pk = pk && a[-2] < a[-1] && a[-1] < v && v > a[1] && a[1] > a[2];
        if (pk) {
          const int slot = atomicAdd(&st.n_peaks, 1);
          if (slot < pk_cap) { pk_idx[slot] = s * C + p; pk_val[slot] = v; }
        }
      }
    }
    __syncthreads();
    // top-N = argsort(values)[::-1][:N] of a stable sort: among equal values the LATER peak comes first
    {
      const int n_pk = min(st.n_peaks, (int)(cells_cap / 2));
      const int want = (int)min((long long)cfg.candidate_count, (long long)S4_MAX_CAND);
      int top_n = 0;
      double last_v = 0;
      int last_i = 0;
      for (int r = 0; r < want; r++) {
        int best = -1;
        double bv = 0;
        for (int t = tid; t < n_pk; t += S4_THREADS) {
          const double v = pk_val[t];
          const int ix = pk_idx[t];
          if (r > 0 && !(v < last_v || (v == last_v && ix < last_i))) continue;  // already taken
          if (best < 0 || v > bv || (v == bv && ix > best)) { best = ix; bv = v; }
        }
        for (int off = 16; off > 0; off >>= 1) {
          const int ob = __shfl_xor_sync(FULL, best, off);
          const double ov = __shfl_xor_sync(FULL, bv, off);
          if (ob >= 0 && (best < 0 || ov > bv || (ov == bv && ob > best))) { best = ob; bv = ov; }
        }
        if (lane == 0) { st.red_idx[warp] = best; st.red_val[warp] = bv; }
        __syncthreads();
        best = st.red_idx[0]; bv = st.red_val[0];
        for (int w = 1; w < S4_WARPS; w++) {
          const int ob = st.red_idx[w];
          const double ov = st.red_val[w];
          if (ob >= 0 && (best < 0 || ov > bv || (ov == bv && ob > best))) { best = ob; bv = ov; }
        }
        __syncthreads();
        if (best < 0) break;  // uniform
        if (tid == 0) { st.top_idx[top_n] = best; st.top_val[top_n] = bv; }
        top_n++;
        last_v = bv; last_i = best;
      }
      if (tid == 0) st.top_n = top_n;
    }
    __syncthreads();
    if (tid == 0) {
      // selection.py:229-284 _join_close_peaks(3, 3)
      const int top_n = st.top_n;
      int t_scan[S4_MAX_CAND], t_cyc[S4_MAX_CAND];
      double t_val[S4_MAX_CAND];
      bool mask[S4_MAX_CAND];
      for (int r = 0; r < top_n; r++) { t_scan[r] = st.top_idx[r] / C; t_cyc[r] = st.top_idx[r] % C; t_val[r] = st.top_val[r]; mask[r] = true; }
      for (int x = 0; x < top_n; x++) {
        if (!mask[x]) continue;
        for (int y = x + 1; y < top_n; y++) {
          if (!mask[y]) continue;
          if (abs(t_scan[x] - t_scan[y]) <= 3 && abs(t_cyc[x] - t_cyc[y]) <= 3) { if (t_val[x] > t_val[y]) mask[y] = false; else mask[x] = false; }
        }
      }
      int n_c = 0;
      for (int r = 0; r < top_n; r++) if (mask[r]) { t_scan[n_c] = t_scan[r]; t_cyc[n_c] = t_cyc[r]; t_val[n_c] = t_val[r]; n_c++; }
      int slim[S4_MAX_CAND][2], clim[S4_MAX_CAND][2];
      for (int r = 0; r < n_c; r++) {  // selection/utils.py:276-312
        const int ml = max(0, t_scan[r] - (int)cfg.min_size_mobility), mu = min(S, t_scan[r] + (int)cfg.min_size_mobility);
        const int cl = max(0, t_cyc[r] - (int)cfg.min_size_rt), cu = min(C, t_cyc[r] + (int)cfg.min_size_rt);
        sym_limits_1d(score, C, 0, cl, cu, S, t_scan[r], cfg.f_mobility, cfg.center_fraction, (int)cfg.min_size_mobility,
                      (int)cfg.max_size_mobility, slim[r]);
        sym_limits_1d(score, C, 1, ml, mu, C, t_cyc[r], cfg.f_rt, cfg.center_fraction, (int)cfg.min_size_rt,
                      (int)cfg.max_size_rt, clim[r]);
      }
      if (cfg.join_close_candidates) {  // selection.py:287-364
        bool jm[S4_MAX_CAND];
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
          t_scan[mm] = t_scan[r]; t_cyc[mm] = t_cyc[r]; t_val[mm] = t_val[r];
          slim[mm][0] = slim[r][0]; slim[mm][1] = slim[r][1]; clim[mm][0] = clim[r][0]; clim[mm][1] = clim[r][1]; mm++;
        }
        n_c = mm;
      }
      // selection.py:480-526 write-out
      const long long i = st.row, L = raw.Fr;
      for (int r = 0; r < n_c; r++) {
        const long long row = i * cfg.candidate_count + r;
        if (row >= P.out.n_rows) break;
        P.out.precursor_idx[row] = P.lib.precursor_idx[i];
        P.out.rank[row] = (uint8_t)r;
        P.out.score[row] = (float)t_val[r];
        P.out.scan_center[row] = (uint32_t)adb_wrap0(t_scan[r] + st.s0, raw.scan_max_index);
        P.out.scan_start[row] = (uint32_t)adb_wrap0(slim[r][0] + st.s0, raw.scan_max_index);
        P.out.scan_stop[row] = (uint32_t)adb_wrap0(slim[r][1] + st.s0, raw.scan_max_index);
        P.out.frame_center[row] = (uint32_t)adb_wrap0((long long)t_cyc[r] * L + st.f0, raw.frame_max_index);
        P.out.frame_start[row] = (uint32_t)adb_wrap0((long long)clim[r][0] * L + st.f0, raw.frame_max_index);
        P.out.frame_stop[row] = (uint32_t)adb_wrap0((long long)clim[r][1] * L + st.f0, raw.frame_max_index);
      }
    }
  }
}

size_t select4d_smem_bytes(const DevRaw4& raw, const Select4Geometry& g, int kh, int kw, bool tile_in_smem) {
  const size_t cx_cap = (size_t)g.c_cap + kw, mw_cap = (cx_cap + 31) / 32 + 1;
  size_t b = sizeof(double) * (size_t)kh * kw;
  if (tile_in_smem) b += sizeof(uint32_t) * (size_t)g.s_cap * cx_cap;
  b += sizeof(uint32_t) * (size_t)g.s_cap * mw_cap;
  b += 3 * sizeof(unsigned short) * (size_t)((g.s_cap + 7) & ~7);
  b += sizeof(uint32_t) * (size_t)(((g.s_cap + 31) / 32 + 1) & ~1);
  b += (size_t)raw.Fr * g.s_cap;
  return b + 16;
}

}  // namespace

size_t adb_select4d_ws_bytes_per_cta(const Select4Geometry& g) {
  const size_t cells = (size_t)g.s_cap * g.c_cap;
  size_t b = 2 * sizeof(float) * cells + sizeof(double) * cells + (sizeof(double) + sizeof(int)) * (cells / 2 + 8) +
             sizeof(uint32_t) * (size_t)g.s_cap * ((size_t)g.c_cap + ADB_MAX_KERNEL_W);  // tile with halo when it does not fit smem
  return (b + 255) & ~(size_t)255;
}

// grid size (resident CTAs) and shared-memory plan; returns 0 when even the smallest plan does not fit
int adb_select4d_grid(int device, const DevRaw4& raw, const Select4Geometry& g, int kh, int kw, size_t* dyn_smem, int* tile_in_smem) {
  int sms = 148, max_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  size_t want = select4d_smem_bytes(raw, g, kh, kw, true);
  int in_smem = 1;
  if (want + sizeof(Sel4State) + 1024 > (size_t)max_optin) {
    in_smem = 0;
    want = select4d_smem_bytes(raw, g, kh, kw, false);
    if (want + sizeof(Sel4State) + 1024 > (size_t)max_optin) return 0;
  }
  cudaFuncSetAttribute(adb_select4d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adb_select4d_kernel, S4_THREADS, want);
  if (per_sm < 1) per_sm = 1;
  *dyn_smem = want;
  *tile_in_smem = in_smem;
  return sms * per_sm;
}

void adb_launch_select4d(const DevRaw4& raw, const DevLib& lib, const adb_selection_config& cfg, const double* d_kernel, int kh,
                         int kw, DevCandidatesOut out, int64_t n, const int32_t* d_order, uint32_t* d_status,
                         const Select4Geometry& g, void* workspace, size_t ws_per_cta, int grid, size_t dyn_smem,
                         int tile_in_smem, cudaStream_t stream, int* n_launches) {
  if (n <= 0) return;
  Select4Params P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.kern = d_kernel; P.kh = kh; P.kw = kw; P.out = out; P.n = n;
  P.order = d_order; P.status = d_status; P.s_cap = g.s_cap; P.c_cap = g.c_cap; P.tile_in_smem = tile_in_smem;
  P.ws = (char*)workspace; P.ws_per_cta = ws_per_cta;
  long long blocks = std::min<long long>(n, grid);
  adb_select4d_kernel<<<(unsigned)blocks, S4_THREADS, dyn_smem, stream>>>(P);
  if (n_launches) (*n_launches)++;
}

// alphadia_b200 — candidate selection kernel (3-D raw files), sm_100a.
//
// Replaces _select_candidates_pjit (alphadia/search/selection/selection.py:78-203) and everything below it:
// AlphaRawJIT.get_dense_intensity (jitclasses/alpharaw_jit.py:339-425), get_frame_indices
// (jitclasses/utils.py:24-88), convolve_fourier (selection/fft.py:141-212, as its defining circular
// convolution with fp64 FMA accumulation — see DESIGN.md), _build_features/_build_candidates
// (selection.py:206-226,367-526), find_peaks_1d / symetric_limits_2d (selection/utils.py:45-74,205-312).
//
// Mapping: a CTA of SEL_SLOTS warps works on SEL_SLOTS precursors at a time ("slots"); persistent CTAs stride
// over groups of precursors in (quad window, RT) order so co-resident CTAs hit the same spectra in L2.
//   phase 0  warp w <-> slot w: isotope m/z, fragment filter + m/z sort, RT window -> cycle window, quad windows
//   phase 1  all threads over (slot, cycle, layer) items, layer fastest (neighbouring threads search the same
//            spectrum); SEL_ILP independent searches in flight per thread; per-spectrum lower bounds go
//            through the L2-resident m/z bucket index; XIC layers land in shared memory with a circular halo
//   phase 2  all threads over (slot, cycle) cells: circular Gaussian smoothing of every layer (fp64 FMA from
//            the kernel in the constant bank, kernel rows then columns ascending), log(x + 1) in fp64 rounded
//            to f32, f32 layer sums -> score[slot][cycle]
//   phase 3  warp w <-> slot w: strict 5-point peaks, top-N (warp arg-max), close-peak suppression,
//            symmetric limits, write-out (integer-exact tail).
#include "adb_common.cuh"

#define FULL 0xffffffffu
#define SEL_SLOTS 4
#define SEL_THREADS (SEL_SLOTS * 32)
#define SEL_ILP 4
#define SEL_MAX_LAYERS (ADB_MAX_LIB_FRAGMENTS + ADB_MAX_ISOTOPES)
#define SEL_MAX_CAND 16

namespace {

struct SelectParams {
  DevRaw raw;
  DevLib lib;
  adb_selection_config cfg;
  double kern[2 * ADB_MAX_KERNEL_W];  // [2][kw] Gaussian kernel as doubles (constant bank)
  int kw;
  DevCandidatesOut out;
  long long row_begin, row_end;
  const int32_t* order;  // optional processing order of library rows
  uint32_t* status;
  int c_cap;       // cycles per slot the shared-memory layout can hold
  int layer_cap;   // layers per slot the shared-memory layout can hold
  float* workspace;      // per-slot HBM fallback for oversized windows
  long long ws_floats_per_slot;
};

struct SlotMeta {
  float lo[SEL_MAX_LAYERS], hi[SEL_MAX_LAYERS];
  float tmp_mz[ADB_MAX_LIB_FRAGMENTS];
  float iso_mz[ADB_MAX_ISOTOPES];
  int pos[ADB_MAX_OBS];
  int nF, nI, nobs, C, ok;
  long long frame_lo, cs, row;
  float* dense;    // [nL][C + kw - 1] with halo
  double* score;   // [C]
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
__device__ void slot_finish(const SelectParams& P, SlotMeta& sl, int lane) {
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

// smoothing of one layer at cycle c: out = sum_a sum_b k[a][b] * x[(c + kw/2 - b) mod C], a then b ascending.
// ext[t] = x[(t - off) mod C] with off = kw - 1 - kw/2, so x[(c + kw/2 - b) mod C] = ext[c + kw - 1 - b].
template <int KW>
__device__ __forceinline__ float smooth_cell(const SelectParams& P, const float* ext, int c, int kw) {
  double acc = 0.0;
  if (KW > 0) {
    double v[KW > 0 ? KW : 1];
#pragma unroll
    for (int u = 0; u < KW; u++) v[u] = (double)ext[c + u];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < KW; b++) acc = fma(P.kern[a * KW + b], v[KW - 1 - b], acc);
  } else {
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < kw; b++) acc = fma(P.kern[a * kw + b], (double)ext[c + kw - 1 - b], acc);
  }
  return (float)acc;
}

template <int KW>
__global__ void __launch_bounds__(SEL_THREADS) adb_select_kernel(const __grid_constant__ SelectParams P) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ SlotMeta slots[SEL_SLOTS];
  __shared__ long long item_prefix[SEL_SLOTS + 1];
  __shared__ int cell_prefix[SEL_SLOTS + 1];
  const DevRaw& raw = P.raw;
  const adb_selection_config& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t L = raw.cycle_len;
  const int kw = (KW > 0) ? KW : P.kw;
  const int ext_stride_cap = P.c_cap + kw - 1;
  // dynamic shared memory per slot: score doubles [c_cap] | dense floats [layer_cap][c_cap + kw - 1]
  const size_t slot_bytes = sizeof(double) * (size_t)P.c_cap + sizeof(float) * (size_t)P.layer_cap * ext_stride_cap;
  const size_t slot_bytes_al = (slot_bytes + 15) & ~(size_t)15;

  const long long n_groups = (P.row_end - P.row_begin + SEL_SLOTS - 1) / SEL_SLOTS;
  for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
    __syncthreads();
    // ---------------- phase 0 ----------------
    {
      SlotMeta& sl = slots[warp];
      long long it = P.row_begin + g * SEL_SLOTS + warp;
      if (it < P.row_end) {
        const int64_t i = P.order ? (int64_t)P.order[it] : (int64_t)it;
        slot_setup(P, sl, i, lane);
      } else if (lane == 0) {
        sl.ok = 0; sl.nF = 0; sl.nI = 0; sl.C = 0; sl.nobs = 0;
      }
      __syncwarp();
      if (lane == 0) {
        unsigned char* base = dyn + (size_t)warp * slot_bytes_al;
        sl.score = (double*)base;
        sl.dense = (float*)(base + sizeof(double) * (size_t)P.c_cap);
        if (sl.ok && (sl.C > P.c_cap || sl.nF + sl.nI > P.layer_cap)) {  // HBM fallback for oversized windows
          long long need = (long long)(sl.nF + sl.nI) * (sl.C + kw - 1) + 2LL * sl.C + 4;
          if (P.workspace == nullptr || need > P.ws_floats_per_slot) {
            atomicOr(P.status, ADB_STATUS_SCRATCH_OVERFLOW);
            sl.ok = 0;
          } else {
            float* ws = P.workspace + ((size_t)blockIdx.x * SEL_SLOTS + warp) * (size_t)P.ws_floats_per_slot;
            sl.score = (double*)ws;  // workspace slots are 16-byte aligned
            sl.dense = ws + 2 * (size_t)sl.C;
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      long long acc = 0;
      int cacc = 0;
      for (int s = 0; s < SEL_SLOTS; s++) {
        item_prefix[s] = acc;
        cell_prefix[s] = cacc;
        if (slots[s].ok) { acc += (long long)(slots[s].nF + slots[s].nI) * slots[s].C; cacc += slots[s].C; }
      }
      item_prefix[SEL_SLOTS] = acc;
      cell_prefix[SEL_SLOTS] = cacc;
    }
    __syncthreads();
    // ---------------- phase 1: XICs (alpharaw_jit.py:398-423), SEL_ILP searches in flight per thread ----
    const long long n_items = item_prefix[SEL_SLOTS];
    for (long long t0 = tid; t0 < n_items; t0 += (long long)SEL_THREADS * SEL_ILP) {
      float lo[SEL_ILP], hi[SEL_ILP], prev_hi[SEL_ILP], acc[SEL_ILP];
      float* dst[SEL_ILP];
      const int* posv[SEL_ILP];
      long long cyc_base[SEL_ILP];
      int n_o[SEL_ILP];
      int max_o = 0;
#pragma unroll
      for (int q = 0; q < SEL_ILP; q++) {
        long long t = t0 + (long long)q * SEL_THREADS;
        n_o[q] = 0; acc[q] = 0.f; dst[q] = nullptr; posv[q] = nullptr; lo[q] = 0.f; hi[q] = 0.f; prev_hi[q] = -1.f; cyc_base[q] = 0;
        if (t < n_items) {
          int s = 0;
#pragma unroll
          for (int z = 1; z < SEL_SLOTS; z++) s += (t >= item_prefix[z]);
          const SlotMeta& sl = slots[s];
          long long local = t - item_prefix[s];
          const int nL = sl.nF + sl.nI;
          const int li = (int)local;
          int k = li % nL, c = li / nL;
          lo[q] = sl.lo[k]; hi[q] = sl.hi[k];
          prev_hi[q] = (k > 0 && k != sl.nF) ? sl.hi[k - 1] : -1.0f;
          const bool ms1 = k >= sl.nF;
          n_o[q] = ms1 ? raw.n_ms1_pos : sl.nobs;
          posv[q] = ms1 ? raw.ms1_pos : sl.pos;
          cyc_base[q] = (sl.cs + c) * L;
          dst[q] = sl.dense + (size_t)k * (sl.C + kw - 1) + (kw - 1 - kw / 2) + c;
          max_o = max(max_o, n_o[q]);
        }
      }
      for (int o = 0; o < max_o; o++) {
        uint32_t l[SEL_ILP], h[SEL_ILP], stop[SEL_ILP];
#pragma unroll
        for (int q = 0; q < SEL_ILP; q++) {
          l[q] = 0; h[q] = 0; stop[q] = 0;
          if (o < n_o[q]) adb_bucket_range(raw, (int64_t)posv[q][o] + cyc_base[q], lo[q], l[q], h[q], stop[q]);
        }
        bool any = true;
        while (any) {  // interleaved binary searches: the SEL_ILP loads of a round are independent
          any = false;
          float v[SEL_ILP];
          uint32_t mid[SEL_ILP];
#pragma unroll
          for (int q = 0; q < SEL_ILP; q++) {
            mid[q] = (l[q] + h[q]) >> 1;
            v[q] = (h[q] - l[q] > 8u) ? __ldg(raw.mz + mid[q]) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < SEL_ILP; q++)
            if (h[q] - l[q] > 8u) {
              if (v[q] < lo[q]) l[q] = mid[q] + 1; else h[q] = mid[q];
              any |= (h[q] - l[q] > 8u);
            }
        }
        uint32_t idx[SEL_ILP];
#pragma unroll
        for (int q = 0; q < SEL_ILP; q++) idx[q] = (o < n_o[q]) ? adb_finish_lower_bound(raw.mz, l[q], h[q], lo[q]) : 0u;
#pragma unroll
        for (int q = 0; q < SEL_ILP; q++)
          if (o < n_o[q]) {
            uint32_t i2 = idx[q];
            if (prev_hi[q] >= lo[q])
              while (i2 < stop[q] && __ldg(raw.mz + i2) <= prev_hi[q]) i2++;
            while (i2 < stop[q] && __ldg(raw.mz + i2) <= hi[q]) {
              acc[q] = __fadd_rn(acc[q], __ldg(raw.intensity + i2));
              i2++;
            }
          }
      }
#pragma unroll
      for (int q = 0; q < SEL_ILP; q++)
        if (dst[q]) *dst[q] = acc[q];
    }
    __syncthreads();
    // circular halo: ext[t] = x[(t - off) mod C], t in [0, C + kw - 1)
    for (int s = 0; s < SEL_SLOTS; s++) {
      const SlotMeta& sl = slots[s];
      if (!sl.ok) continue;
      const int C = sl.C, nL = sl.nF + sl.nI, stride = C + kw - 1, off = kw - 1 - kw / 2;
      const int halo = kw - 1;
      for (int t = tid; t < nL * halo; t += SEL_THREADS) {
        int k = t / halo, u = t % halo;
        float* row = sl.dense + (size_t)k * stride;
        if (u < off) row[u] = row[u + C];
        else row[C + u] = row[u];  // u in [off, kw-1): positions C+off .. C+kw-2 mirror off .. kw-2
      }
    }
    __syncthreads();
    // ---------------- phase 2: smooth + log-sum (selection.py:389-428) ----------------
    {
      const int n_cells = cell_prefix[SEL_SLOTS];
      for (int t = tid; t < n_cells; t += SEL_THREADS) {
        int s = 0;
#pragma unroll
        for (int z = 1; z < SEL_SLOTS; z++) s += (t >= cell_prefix[z]);
        const SlotMeta& sl = slots[s];
        const int c = t - cell_prefix[s];
        const int stride = sl.C + kw - 1, nL = sl.nF + sl.nI;
        float lf = 0.f, lp = 0.f;
        for (int l = 0; l < nL; l++) {
          float smooth = smooth_cell<KW>(P, sl.dense + (size_t)l * stride, c, kw);
          float lg = (float)log((double)smooth + 1.0);
          if (l < sl.nF) lf = __fadd_rn(lf, lg); else lp = __fadd_rn(lp, lg);
        }
        sl.score[c] = (double)__fadd_rn(lf, lp);  // raw feature, normalised below
      }
    }
    __syncthreads();
    // normalisation (selection.py:405-428), warp w <-> slot w
    {
      SlotMeta& sl = slots[warp];
      if (sl.ok) {
        const int C = sl.C;
        double mean = cfg.use_weighted_score ? cfg.feature_mean : 0.0;
        double stdv = cfg.use_weighted_score ? cfg.feature_std : 0.0;
        const double wgt = cfg.use_weighted_score ? cfg.feature_weight : 1.0;
        if (!cfg.use_weighted_score) {  // amean1 / astd1 over the (2, C) feature map; sequential, rare path
          float accf = 0.f;
          for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) accf = __fadd_rn(accf, (float)sl.score[c]);
          mean = (double)accf / (double)(2 * C);
          double v = 0;
          for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) { double d = sl.score[c] - mean; v = __dadd_rn(v, __dmul_rn(d, d)); }
          stdv = sqrt(v / (double)(2 * C));
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32) sl.score[c] = 0.0 + __dmul_rn(wgt, (sl.score[c] - mean)) / (stdv + 1e-6);
        __syncwarp();
        // ---------------- phase 3 ----------------
        slot_finish(P, sl, lane);
      }
    }
  }
}

size_t select_smem_bytes(int c_cap, int layer_cap, int kw) {
  size_t slot = sizeof(double) * (size_t)c_cap + sizeof(float) * (size_t)layer_cap * (size_t)(c_cap + kw - 1);
  slot = (slot + 15) & ~(size_t)15;
  return slot * SEL_SLOTS;
}

}  // namespace

size_t adb_select_smem_bytes(int c_cap, int max_layers, int kw) { return select_smem_bytes(c_cap, max_layers, kw); }

int adb_select_resident_ctas(int device, int c_cap, int max_layers, int kw) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  size_t dyn = select_smem_bytes(c_cap, max_layers, kw);
  int per_sm = 0;
  if (kw == 30) {
    cudaFuncSetAttribute(adb_select_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adb_select_kernel<30>, SEL_THREADS, dyn);
  } else {
    cudaFuncSetAttribute(adb_select_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adb_select_kernel<0>, SEL_THREADS, dyn);
  }
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

int adb_select_slots(void) { return SEL_SLOTS; }

void adb_launch_select_ex(const DevRaw& raw, const DevLib& lib, const adb_selection_config& cfg, const double* h_kernel,
                          int kh, int kw, DevCandidatesOut out, int64_t row_begin, int64_t row_end, const int32_t* d_order,
                          uint32_t* d_status, int c_cap, int max_layers, float* d_workspace, int64_t ws_floats_per_slot,
                          int grid, cudaStream_t stream, int* n_launches) {
  if (row_end <= row_begin) return;
  (void)kh;
  SelectParams P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.kw = kw; P.out = out;
  for (int t = 0; t < 2 * ADB_MAX_KERNEL_W; t++) P.kern[t] = (t < 2 * kw) ? h_kernel[t] : 0.0;
  P.row_begin = row_begin; P.row_end = row_end; P.order = d_order; P.status = d_status;
  P.c_cap = c_cap; P.layer_cap = max_layers; P.workspace = d_workspace; P.ws_floats_per_slot = ws_floats_per_slot;
  size_t dyn = select_smem_bytes(c_cap, max_layers, kw);
  long long n_groups = (row_end - row_begin + SEL_SLOTS - 1) / SEL_SLOTS;
  if (grid > n_groups) grid = (int)n_groups;
  if (grid < 1) grid = 1;
  if (kw == 30) {
    cudaFuncSetAttribute(adb_select_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    adb_select_kernel<30><<<grid, SEL_THREADS, dyn, stream>>>(P);
  } else {
    cudaFuncSetAttribute(adb_select_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    adb_select_kernel<0><<<grid, SEL_THREADS, dyn, stream>>>(P);
  }
  if (n_launches) (*n_launches)++;
}

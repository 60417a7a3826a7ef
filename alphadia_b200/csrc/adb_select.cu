// alphadia_b200 — candidate selection kernel (3-D raw files), sm_100a.
//
// Replaces _select_candidates_pjit (alphadia/search/selection/selection.py:78-203) and everything below it:
// AlphaRawJIT.get_dense_intensity (jitclasses/alpharaw_jit.py:339-425), get_frame_indices
// (jitclasses/utils.py:24-88), convolve_fourier (selection/fft.py:141-212, as its defining circular
// convolution with fp64 FMA accumulation — see DESIGN.md), _build_features/_build_candidates
// (selection.py:206-226,367-526), find_peaks_1d / symetric_limits_2d (selection/utils.py:45-74,205-312).
//
// Mapping: ONE CTA PER PRECURSOR (persistent CTAs stride over precursors).
//   phase 0  warp 0: isotope m/z, fragment filter + m/z sort, RT window -> cycle window, quad windows
//   phase 1  all threads: XIC extraction, items = (cycle, layer) with the layer fastest so neighbouring
//            threads binary-search the same spectrum; dense layers [layer][cycle] in shared memory
//   phase 2  thread <-> cycle: circular Gaussian smoothing of every layer (fp64 FMA, kernel rows then
//            columns ascending), log(x + 1) in fp64 rounded to f32, f32 layer sums -> score[cycle]
//   phase 3  thread 0: strict 5-point peaks, top-N, close-peak suppression, symmetric limits, write-out
//            (integer-exact tail).
#include "adb_common.cuh"

#define FULL 0xffffffffu
#define SEL_THREADS 128
#define SEL_MAX_LAYERS (ADB_MAX_LIB_FRAGMENTS + ADB_MAX_ISOTOPES)
#define SEL_MAX_CAND 16
#define SEL_MAX_PEAKS 1024

namespace {

struct SelectParams {
  DevRaw raw;
  DevLib lib;
  adb_selection_config cfg;
  const double* kernel;  // [kh][kw] as doubles
  int kh, kw;
  DevCandidatesOut out;
  long long row_begin, row_end;
  const int32_t* order;  // optional processing order of library rows
  uint32_t* status;
  int c_cap;             // cycles the shared-memory layout can hold
  float* workspace;      // per-CTA HBM fallback for the dense layers
  long long ws_floats_per_cta;
};

struct SelShared {
  float frag_mz[ADB_MAX_LIB_FRAGMENTS];
  float tmp_mz[ADB_MAX_LIB_FRAGMENTS];
  float lo[SEL_MAX_LAYERS], hi[SEL_MAX_LAYERS];
  float iso_mz[ADB_MAX_ISOTOPES];
  int pos[ADB_MAX_OBS];
  int nF, nI, nobs, C;
  long long frame_lo, cs;
  int ok;
};

__device__ __forceinline__ float extract_intensity(const DevRaw& raw, int64_t scan, float lo, float hi, float prev_hi, float acc) {
  int64_t start = __ldg(raw.peak_start + scan), stop = __ldg(raw.peak_stop + scan);
  int64_t idx = adb_lower_bound(raw.mz, start, stop, lo);
  if (prev_hi >= lo)
    while (idx < stop && __ldg(raw.mz + idx) <= prev_hi) idx++;
  while (idx < stop && __ldg(raw.mz + idx) <= hi) {
    acc = __fadd_rn(acc, __ldg(raw.intensity + idx));
    idx++;
  }
  return acc;
}

__device__ void symetric_limits_1d(const double* a, int n, int center, double f, double cf, int min_size, int max_size, int out[2]) {
  if (n == 0 || center < 0 || center >= n) { out[0] = center; out[1] = center; return; }
  double center_intensity = a[center], trailing = center_intensity;
  int limit = min_size;
  for (int s = min_size + 1; s < max_size; s++) {
    int l = max(center - s, 0), r = min(center + s, n - 1);
    double intensity = __dadd_rn(a[l], a[r]) / 2;
    if (intensity < __dmul_rn(f, trailing)) {
      if (intensity > __dmul_rn(center_intensity, cf)) { limit = s; trailing = intensity; }
      else break;
    } else break;
  }
  out[0] = max(center - limit, 0);
  out[1] = min(center + limit + 1, n);
}

__global__ void __launch_bounds__(SEL_THREADS) adb_select_kernel(const __grid_constant__ SelectParams P) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ SelShared sh;
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_selection_config& cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t L = raw.cycle_len;
  const int kw = P.kw;
  // dynamic shared memory: kernel doubles | score doubles [c_cap] | proj doubles [c_cap] | dense floats
  double* kern = (double*)dyn;
  double* score = kern + 2 * ADB_MAX_KERNEL_W;
  double* proj = score + P.c_cap;
  float* dense_smem = (float*)(proj + P.c_cap);
  for (int t = tid; t < P.kh * kw; t += SEL_THREADS) kern[t] = P.kernel[t];

  for (long long it = P.row_begin + blockIdx.x; it < P.row_end; it += gridDim.x) {
    const int64_t i = P.order ? (int64_t)P.order[it] : (int64_t)it;
    __syncthreads();
    // ---------------- phase 0 ----------------
    if (warp == 0) {
      int nI = (int)min((long long)lib.n_isotopes, (long long)cfg.top_k_precursors);
      nI = min(nI, ADB_MAX_ISOTOPES);
      if (lane < nI) {  // selection/utils.py:35-40: float32 += float64
        double off = (double)lane * ADB_ISOTOPE_DIFF / (double)lib.charge[i];
        sh.iso_mz[lane] = (float)((double)lib.mz[i] + off);
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
        if (keep) sh.tmp_mz[m + __popc(b & ((1u << lane) - 1u))] = lib.frag_mz[fs + j];
        m += __popc(b);
      }
      __syncwarp();
      for (int u = lane; u < m; u += 32) {  // stable ascending m/z
        float v = sh.tmp_mz[u];
        int rk = 0;
        for (int q = 0; q < m; q++) rk += (sh.tmp_mz[q] < v) || (sh.tmp_mz[q] == v && q < u);
        sh.frag_mz[rk] = v;
      }
      __syncwarp();
      if (m <= 3) ok = 0;  // selection.py:136-137
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
      if (C < kw) ok = 0;  // selection.py:61-73 (scan extent 2 >= kernel height 2)
      // windows: jitclasses/utils.py:15-20 with float64 tolerances
      if (lane < nI) {
        double mz = (double)sh.iso_mz[lane], d = cfg.precursor_mz_tolerance * mz / 1000000.0;
        sh.lo[m + lane] = (float)(mz - d);
        sh.hi[m + lane] = (float)(mz + d);
      }
      for (int u = lane; u < m; u += 32) {
        double mz = (double)sh.frag_mz[u], d = cfg.fragment_mz_tolerance * mz / 1000000.0;
        sh.lo[u] = (float)(mz - d);
        sh.hi[u] = (float)(mz + d);
      }
      __syncwarp();
      const float q0 = sh.iso_mz[0], q1 = sh.iso_mz[max(nI - 1, 0)];  // selection.py:152
      int nobs = 0;
      for (int64_t base = 0; base < L; base += 32) {
        int64_t j = base + lane;
        bool hit = j < L && ((double)q0 <= raw.cycle[2 * j + 1]) && ((double)q1 >= raw.cycle[2 * j]);
        unsigned b = __ballot_sync(FULL, hit);
        if (hit) { int u = nobs + __popc(b & ((1u << lane) - 1u)); if (u < ADB_MAX_OBS) sh.pos[u] = (int)j; }
        nobs += __popc(b);
      }
      if (nobs > ADB_MAX_OBS) { if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_OBS); ok = 0; }
      if (lane == 0) {
        sh.nF = m; sh.nI = nI; sh.nobs = nobs; sh.C = (int)min(C, 2000000000LL);
        sh.frame_lo = frame_lo; sh.cs = cs; sh.ok = ok;
      }
    }
    __syncthreads();
    if (!sh.ok) continue;
    const int nF = sh.nF, nI = sh.nI, nobs = sh.nobs, C = sh.C;
    const int nL = nF + nI;
    const long long cs = sh.cs;
    float* dense = dense_smem;
    if (C > P.c_cap) {
      if (P.workspace == nullptr || (long long)nL * C + 3LL * 2 * C > P.ws_floats_per_cta) {
        if (tid == 0) atomicOr(P.status, ADB_STATUS_SCRATCH_OVERFLOW);
        continue;
      }
      dense = P.workspace + (size_t)blockIdx.x * (size_t)P.ws_floats_per_cta;
    }
    double* score_p = score;
    double* proj_p = proj;
    if (C > P.c_cap) {  // fallback: doubles live behind the dense layers in the workspace
      float* after = dense + (size_t)nL * C;
      after += ((uintptr_t)after & 7u) ? 1 : 0;
      score_p = (double*)after;
      proj_p = score_p + C;
    }
    // ---------------- phase 1: XICs (alpharaw_jit.py:398-423) ----------------
    for (long long t = tid; t < (long long)nF * C; t += SEL_THREADS) {
      int k = (int)(t % nF), c = (int)(t / nF);
      float acc = 0.f;
      float prev_hi = k > 0 ? sh.hi[k - 1] : -1.0f;
      for (int o = 0; o < nobs; o++)
        acc = extract_intensity(raw, (int64_t)sh.pos[o] + (cs + c) * L, sh.lo[k], sh.hi[k], prev_hi, acc);
      dense[(size_t)k * C + c] = acc;
    }
    for (long long t = tid; t < (long long)nI * C; t += SEL_THREADS) {
      int k = (int)(t % nI), c = (int)(t / nI);
      float acc = 0.f;
      float prev_hi = k > 0 ? sh.hi[nF + k - 1] : -1.0f;
      for (int o = 0; o < raw.n_ms1_pos; o++)
        acc = extract_intensity(raw, (int64_t)raw.ms1_pos[o] + (cs + c) * L, sh.lo[nF + k], sh.hi[nF + k], prev_hi, acc);
      dense[(size_t)(nF + k) * C + c] = acc;
    }
    __syncthreads();
    // ---------------- phase 2: smooth + log-sum (selection.py:389-428) ----------------
    {
      const int s1 = kw / 2;
      double mean = cfg.use_weighted_score ? cfg.feature_mean : 0.0;
      double stdv = cfg.use_weighted_score ? cfg.feature_std : 0.0;
      const double wgt = cfg.use_weighted_score ? cfg.feature_weight : 1.0;
      for (int c = tid; c < C; c += SEL_THREADS) {
        float lf = 0.f, lp = 0.f;
        for (int l = 0; l < nL; l++) {
          const float* x = dense + (size_t)l * C;
          double acc = 0.0;
          for (int a = 0; a < 2; a++) {
            const double* kr = kern + a * kw;
            int jj = c + s1;
            if (jj >= C) jj -= C;
            for (int b = 0; b < kw; b++) {
              acc = fma(kr[b], (double)x[jj], acc);
              jj = (jj == 0) ? C - 1 : jj - 1;
            }
          }
          float smooth = (float)acc;
          float lg = (float)log((double)smooth + 1.0);
          if (l < nF) lf = __fadd_rn(lf, lg); else lp = __fadd_rn(lp, lg);
        }
        float feat = __fadd_rn(lf, lp);
        proj_p[c] = (double)feat;  // raw feature, normalised below
      }
      __syncthreads();
      if (!cfg.use_weighted_score) {  // selection.py:405-417 amean1/astd1 over the (2, C) feature map
        // every thread computes the same sequential statistics (rare path)
        float accf = 0.f;
        for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) accf = __fadd_rn(accf, (float)proj_p[c]);
        mean = (double)accf / (double)(2 * C);
        double v = 0;
        for (int s = 0; s < 2; s++) for (int c = 0; c < C; c++) { double d = proj_p[c] - mean; v = __dadd_rn(v, __dmul_rn(d, d)); }
        stdv = sqrt(v / (double)(2 * C));
      }
      for (int c = tid; c < C; c += SEL_THREADS)
        score_p[c] = 0.0 + __dmul_rn(wgt, (proj_p[c] - mean)) / (stdv + 1e-6);
    }
    __syncthreads();
    // ---------------- phase 3: peaks -> candidates (thread 0) ----------------
    if (tid == 0) {
      const double* a = score_p;
      int t_cyc[SEL_MAX_CAND];
      double t_val[SEL_MAX_CAND];
      int top_n = 0;
      const int want = (int)min((long long)cfg.candidate_count, (long long)SEL_MAX_CAND);
      // top-N of the strict 5-point maxima; argsort(...)[::-1] of a stable sort: ties -> later index first
      double last_v = 0;
      int last_p = 0;
      for (int r = 0; r < want; r++) {
        int best = -1;
        double bv = 0;
        for (int p = 2; p < C - 2; p++) {
          if (!(a[p - 2] < a[p - 1] && a[p - 1] < a[p] && a[p] > a[p + 1] && a[p + 1] > a[p + 2])) continue;
          double v = a[p];
          if (r > 0 && !(v < last_v || (v == last_v && p < last_p))) continue;  // already taken
          if (best < 0 || v > bv || (v == bv && p > best)) { best = p; bv = v; }
        }
        if (best < 0) break;
        t_cyc[top_n] = best; t_val[top_n] = bv; top_n++;
        last_v = bv; last_p = best;
      }
      // selection.py:229-284 _join_close_peaks(3, 3); scan index is always 0
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
        symetric_limits_1d(ps, 2, scan_center, cfg.f_mobility, cfg.center_fraction, (int)cfg.min_size_mobility, (int)cfg.max_size_mobility, slim[r]);
        const int nrows = max(mu - ml, 0);
        // cycle projection: sum over nrows identical rows, evaluated lazily inside the limits walk
        // (materialise once into proj)
        for (int c = 0; c < C; c++) proj_p[c] = (nrows == 2) ? __dadd_rn(a[c], a[c]) : (nrows == 1 ? a[c] : 0.0);
        symetric_limits_1d(proj_p, C, cc, cfg.f_rt, cfg.center_fraction, (int)cfg.min_size_rt, (int)cfg.max_size_rt, clim[r]);
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
      const long long frame_lo = sh.frame_lo;
      for (int r = 0; r < n_c; r++) {
        long long row = (long long)i * cfg.candidate_count + r;
        if (row >= P.out.n_rows) break;
        P.out.precursor_idx[row] = lib.precursor_idx[i];
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
  }
}

}  // namespace

// host-side launch wrapper with explicit geometry (called from adb_api.cu)
void adb_launch_select_ex(const DevRaw& raw, const DevLib& lib, const adb_selection_config& cfg, const double* d_kernel,
                          int kh, int kw, DevCandidatesOut out, int64_t row_begin, int64_t row_end, const int32_t* d_order,
                          uint32_t* d_status, int c_cap, int max_layers, float* d_workspace, int64_t ws_floats_per_cta,
                          int grid, cudaStream_t stream, int* n_launches) {
  if (row_end <= row_begin) return;
  SelectParams P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.kernel = d_kernel; P.kh = kh; P.kw = kw; P.out = out;
  P.row_begin = row_begin; P.row_end = row_end; P.order = d_order; P.status = d_status;
  P.c_cap = c_cap; P.workspace = d_workspace; P.ws_floats_per_cta = ws_floats_per_cta;
  size_t dyn = sizeof(double) * (2 * ADB_MAX_KERNEL_W + 2 * (size_t)c_cap) + sizeof(float) * (size_t)max_layers * (size_t)c_cap;
  cudaFuncSetAttribute(adb_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  long long n = row_end - row_begin;
  if (grid > n) grid = (int)n;
  if (grid < 1) grid = 1;
  adb_select_kernel<<<grid, SEL_THREADS, dyn, stream>>>(P);
  if (n_launches) (*n_launches)++;
}

size_t adb_select_smem_bytes(int c_cap, int max_layers) {
  return sizeof(double) * (2 * ADB_MAX_KERNEL_W + 2 * (size_t)c_cap) + sizeof(float) * (size_t)max_layers * (size_t)c_cap;
}

int adb_select_resident_ctas(int device, int c_cap, int max_layers) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  size_t dyn = adb_select_smem_bytes(c_cap, max_layers);
  cudaFuncSetAttribute(adb_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adb_select_kernel, SEL_THREADS, dyn);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

// alphadia_b200 — device-side common definitions (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/alphadia_b200.h"

#define ADB_MAX_OBS 8            // quad windows one candidate may hit
#define ADB_MAX_LIB_FRAGMENTS 128 // fragments per precursor in the flat library (before top-k)
#define ADB_MAX_MS1_POS 8        // MS1 spectra per DIA cycle
#define ADB_MAX_KERNEL_W 64
#define ADB_ISOTOPE_DIFF 1.0033548350700006
#define ADB_N_BUCKETS 256        // m/z buckets per spectrum in the derived search index
#define ADB_MZ_PAD 16            // floats of padding behind mz/intensity so vector tail reads stay in bounds

// status bits reported by kernels through a device word (-> adb_last_error on the host)
#define ADB_STATUS_TOO_MANY_OBS 1u
#define ADB_STATUS_TOO_MANY_LIB_FRAGMENTS 2u
#define ADB_STATUS_SCRATCH_OVERFLOW 4u
#define ADB_STATUS_TOO_MANY_PEAKS 8u

// Raw file resident in HBM: the reference's AlphaRawJIT arrays, field for field (SoA, m/z and
// intensity kept separate so binary-search probes touch m/z sectors only).
struct DevRaw {
  const double* cycle;  // [L][2]
  int64_t cycle_len;
  const float* rt_values;
  int64_t n_spectra;
  const float* mobility_values;
  int64_t n_mobility;
  const int64_t* peak_start;
  const int64_t* peak_stop;
  const float* mz;
  const float* intensity;
  int64_t n_peaks;
  int64_t zeroth_frame;
  int64_t precursor_cycle_max_index;
  int64_t scan_max_index;
  int64_t frame_max_index;
  int32_t n_ms1_pos;                // cycle positions whose window overlaps [-1,-1]
  int32_t ms1_pos[ADB_MAX_MS1_POS];
  // derived m/z bucket index (built once per file on the device, sized to stay L2-resident):
  // bucket_pair[scan][b] = {first, last+1} ABSOLUTE peak indices of the peaks of spectrum `scan` that fall into
  // m/z bucket b (edges bucket_lo + b * bucket_width; bucket 0 starts at the spectrum start, the last bucket ends
  // at the spectrum stop).  One 8-byte read replaces the peak_start/peak_stop reads of the reference and all but
  // the last levels of the binary search.
  const uint2* bucket_pair;         // [n_spectra][ADB_N_BUCKETS]
  float bucket_lo, bucket_width, bucket_inv_width;
  // derived m/z-major index (built once per file on the device): for every cycle position p the peaks of ALL its
  // spectra, stably sorted by m/z: segment [pos_start[p], pos_start[p + 1]) of s_pk records.
  // One search answers "which peaks of this quad window fall into this m/z window, in any cycle" — candidate selection
  // extracts a whole XIC row (hundreds of cycles) with it instead of one binary search per spectrum.
  const float4* s_pk;               // {m/z, intensity, cycle index (bit pattern), 0}: one 128-bit load per peak
  const int64_t* pos_start;         // [cycle_len + 1]
  // s_bucket[p][b]: first peak of position p at or above the lower edge of m/z bucket b (entry sb_nb = segment end),
  // about 16 peaks per bucket: the row search of candidate selection is one table read + one 32-wide probe
  const uint32_t* s_bucket;         // [cycle_len][sb_nb + 1]
  int32_t sb_nb;
  float sb_lo, sb_width, sb_inv_width;
  // derived time-blocked m/z index (built once per file on the device), used by candidate scoring: segment
  // (cycle position p, time block t = cycle / ADB_TB_CYCLES) holds the peaks of the <= ADB_TB_CYCLES spectra of position p in
  // that block, stably sorted by m/z.  A scoring window of ~10 cycles touches 1-2 segments and finds ~1 peak of its ppm
  // window in each; tb_bucket[seg][b] is the absolute index of the first peak of the segment at or above the lower edge
  // of m/z bucket b (entry tb_nb = segment end), so a query is one table read + a search over a handful of peaks.
  const float4* tb_pk;              // {m/z, intensity, cycle index (bit pattern), 0}: one 128-bit load per peak
  const uint32_t* tb_bucket;        // [cycle_len * tb_ntb][tb_nb + 1]
  int32_t tb_ntb, tb_nb;
  float tb_lo, tb_width, tb_inv_width;
};

#ifndef ADB_TB_MAX_BUCKETS
#define ADB_TB_MAX_BUCKETS 16384  // m/z buckets per segment of the time-blocked index (about 4 peaks per bucket)
#endif
#ifndef ADB_TB_CYCLES
#define ADB_TB_CYCLES 32
#endif

// m/z edge of bucket b of a bucket table (the SAME float expression builds the table and routes the queries)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline float adb_bucket_edge(float lo, float width, int b) { return fmaf((float)b, width, lo); }

#if defined(__CUDACC__)
__host__ __device__
#endif
inline int adb_bucket_of(float lo, float width, float inv_width, int nb, float v) {
  int b = (int)((v - lo) * inv_width);
  b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
  while (b > 0 && adb_bucket_edge(lo, width, b) > v) b--;
  while (b < nb - 1 && adb_bucket_edge(lo, width, b + 1) <= v) b++;
  return b;
}

#if defined(__CUDACC__)
__host__ __device__
#endif
inline float adb_tb_edge(const DevRaw& raw, int b) { return adb_bucket_edge(raw.tb_lo, raw.tb_width, b); }

#if defined(__CUDACC__)
__host__ __device__
#endif
inline int adb_tb_bucket_of(const DevRaw& raw, float v) { return adb_bucket_of(raw.tb_lo, raw.tb_width, raw.tb_inv_width, raw.tb_nb, v); }

// timsTOF (4-D) raw file resident in HBM: the TimsTOFTransposeJIT arrays the hot path reads
// (alphadia/search/jitclasses/bruker_jit.py:20-137), CSR by tof index.
struct DevRaw4 {
  const double* cycle;                 // [Fr][Sc][2]
  int64_t Fr, Sc;                      // cycle.shape[1], cycle.shape[2]
  const int64_t* dia_precursor_cycle;  // [Fr * Sc]
  const double* rt_values;             // [n_frames]
  int64_t n_frames;
  const double* mobility_values;       // [Sc], descending
  const double* mz_values;             // [n_tof]
  int64_t n_tof;
  const int64_t* tof_indptr;           // [n_tof + 1]
  const uint32_t* push;                // [n_events] ascending inside a tof row
  const uint16_t* intensity;           // [n_events]
  int64_t n_events;
  int64_t zeroth_frame;
  int64_t precursor_cycle_max_index;
  int64_t scan_max_index;
  int64_t frame_max_index;
  // 1 when no observation id occurs in two different frames of the cycle at the same scan: then the events of one
  // tof row always fall into distinct cube cells and may be processed concurrently (adb_score4d.cu)
  int32_t obs_unique_per_scan;
};

#define ADB_MAX_OBS4 16  // observation ids (frames of the cycle) one 4-D candidate may hit

struct DevLib {
  int64_t n_precursors;
  const uint32_t* precursor_idx;
  const uint32_t* frag_start_idx;
  const uint32_t* frag_stop_idx;
  const uint8_t* charge;
  const float* rt;
  const float* mobility;
  const float* mz;
  const float* isotopes;
  int32_t n_isotopes;
  int64_t n_fragments;
  const float* frag_mz_library;
  const float* frag_mz;
  const float* frag_intensity;
  const uint8_t* frag_type;
  const uint8_t* frag_loss_type;
  const uint8_t* frag_charge;
  const uint8_t* frag_number;
  const uint8_t* frag_position;
  const uint8_t* frag_cardinality;
};

struct DevCandidatesOut {  // CandidateContainer on the device
  int64_t n_rows;
  uint32_t* precursor_idx;
  uint8_t* rank;
  float* score;
  uint32_t* scan_center;
  uint32_t* scan_start;
  uint32_t* scan_stop;
  uint32_t* frame_center;
  uint32_t* frame_start;
  uint32_t* frame_stop;
};

struct DevCandidatesIn {
  int64_t n;
  const int64_t* lib_row;
  const uint8_t* rank;
  const int64_t* scan_start;
  const int64_t* scan_stop;
  const int64_t* scan_center;
  const int64_t* frame_start;
  const int64_t* frame_stop;
  const int64_t* frame_center;
};

struct DevScoresOut {
  float* features;
  uint8_t* valid;
  float* fragment_mz_library;
  float* fragment_mz;
  float* fragment_mz_observed;
  float* fragment_height;
  float* fragment_intensity;
  float* fragment_mass_error;
  float* fragment_correlation;
  uint8_t* fragment_position;
  uint8_t* fragment_number;
  uint8_t* fragment_type;
  uint8_t* fragment_charge;
  uint8_t* fragment_loss_type;
};

// ---- launchers (implemented in the .cu files) --------------------------------------------
size_t adb_select_bytes_per_precursor(int c_cap, int max_layers, int kw);
void adb_launch_select_chunk(const DevRaw& raw, const DevLib& lib, const adb_selection_config& cfg, const double* h_kernel,
                             int kw, DevCandidatesOut out, int64_t chunk_begin, int64_t chunk_n, const int32_t* d_order,
                             uint32_t* d_status, int c_cap, int max_layers, void* workspace, int sm_count,
                             cudaStream_t stream, int* n_launches);

void adb_launch_score(const DevRaw& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand,
                      DevScoresOut out, float* d_workspace, int64_t workspace_floats_per_tile, int n_resident_tiles,
                      const int32_t* d_order, uint32_t* d_status, cudaStream_t stream, int* n_launches);
int adb_score_resident_tiles(int device, int top_k);
// data-parallel scoring passes (adb_score_dp.cu)
size_t adb_score_dp_plan_bytes(int64_t nb, int KS, int nIcap, size_t* scan_tmp_bytes);
int adb_launch_score_dp(const DevRaw& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand, DevScoresOut out,
                        int out_k, int KS, const int32_t* d_order, int64_t batch, void* plan, float** cube, size_t* cube_floats,
                        int (*grow)(void* owner, size_t floats), void* owner, uint32_t* d_status, cudaStream_t stream,
                        int* n_launches);
int64_t adb_score_workspace_floats(int top_k, int64_t c_max);

int adb_run_fragcomp_graph(int64_t n_windows, const int64_t* d_ws, const int64_t* d_we, int64_t n_psm, const void* d_rt,
                           const int64_t* d_fs, const int64_t* d_fe, const void* d_mz, int is_f64, double rt_tol, double ppm_tol,
                           uint8_t* d_valid, size_t pair_cap, cudaStream_t stream, int* n_launches);
void adb_launch_fragcomp(int64_t n_windows, const int64_t* d_ws, const int64_t* d_we, const void* d_rt,
                         const int64_t* d_fs, const int64_t* d_fe, const void* d_mz, int is_f64, double rt_tol,
                         double ppm_tol, uint8_t* d_valid, cudaStream_t stream, int* n_launches);

// 4-D (timsTOF) launchers
struct Select4Geometry { int s_cap, c_cap, max_layers; };
size_t adb_select4d_ws_bytes_per_cta(const Select4Geometry& g);
int adb_select4d_grid(int device, const DevRaw4& raw, const Select4Geometry& g, int kh, int kw, size_t* dyn_smem, int* tile_in_smem);
void adb_launch_select4d(const DevRaw4& raw, const DevLib& lib, const adb_selection_config& cfg, const double* d_kernel, int kh,
                         int kw, DevCandidatesOut out, int64_t n, const int32_t* d_order, uint32_t* d_status,
                         const Select4Geometry& g, void* workspace, size_t ws_per_cta, int grid, size_t dyn_smem,
                         int tile_in_smem, cudaStream_t stream, int* n_launches);
int adb_score4d_resident_warps(int device);
int64_t adb_score4d_workspace_floats(int top_k, int n_iso, int64_t s_max, int64_t c_max, int nobs_cap);
void adb_launch_score4d(const DevRaw4& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand,
                        DevScoresOut out, float* d_workspace, int64_t workspace_floats_per_warp, int n_resident_warps,
                        int s_cap, int c_cap, const int32_t* d_order, uint32_t* d_status, cudaStream_t stream, int* n_launches);

size_t adb_compact_temp_bytes(int64_t n_rows);
void adb_launch_compact_ex(DevCandidatesOut cont, int64_t candidate_count, int* d_flags, int* d_offs, void* d_tmp,
                           size_t tmp_bytes, int64_t* d_lib_row, uint8_t* d_rank, int64_t* d_scan_start,
                           int64_t* d_scan_stop, int64_t* d_scan_center, int64_t* d_frame_start, int64_t* d_frame_stop,
                           int64_t* d_frame_center, uint32_t* d_precursor_idx, float* d_score, int64_t* d_count,
                           cudaStream_t stream, int* n_launches, float score_cutoff = -INFINITY);

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ int64_t adb_lower_bound(const float* __restrict__ a, int64_t lo, int64_t hi, float v) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// m/z edge of bucket b (the SAME float expression builds the index and routes the queries)
__device__ __forceinline__ float adb_bucket_edge(const DevRaw& raw, int b) {
  return __fmaf_rn((float)b, raw.bucket_width, raw.bucket_lo);
}

__device__ __forceinline__ int adb_bucket_of(const DevRaw& raw, float v) {
  int b = (int)((v - raw.bucket_lo) * raw.bucket_inv_width);
  b = max(0, min(b, ADB_N_BUCKETS - 1));
  while (b > 0 && adb_bucket_edge(raw, b) > v) b--;
  while (b < ADB_N_BUCKETS - 1 && adb_bucket_edge(raw, b + 1) <= v) b++;
  return b;
}

// Result of a lower-bound search inside one spectrum.
struct AdbFound {
  uint32_t idx;      // absolute index of the first peak with mz >= v (may equal the spectrum stop)
  float mz_at_idx;   // mz[idx] when `inside`
  bool inside;       // idx is known to lie inside the spectrum (idx < bucket end <= stop)
};

// Finish a lower-bound search once the range [lo, hi) is <= 8 peaks: three independent aligned 16-byte reads
// cover it; the answer is lo + #(elements of [lo, hi) below v) because the spectrum is sorted.  The value at the
// answer is taken from the registers already loaded when possible.
__device__ __forceinline__ AdbFound adb_finish_lower_bound(const float* __restrict__ mz, uint32_t lo, uint32_t hi, float v) {
  const uint32_t A = lo & ~3u;
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(mz + A));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(mz + A + 4));
  const float4 x2 = __ldg(reinterpret_cast<const float4*>(mz + A + 8));
  const float e[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    uint32_t idx = A + j;
    cnt += (idx >= lo && idx < hi && e[j] < v) ? 1u : 0u;
  }
  AdbFound f;
  f.idx = lo + cnt;
  f.inside = f.idx < hi;
  const uint32_t j = f.idx - A;
  float val = 0.f;
#pragma unroll
  for (int q = 0; q < 12; q++) val = (j == (uint32_t)q) ? e[q] : val;
  f.mz_at_idx = val;
  if (f.inside && j >= 12u) f.mz_at_idx = __ldg(mz + f.idx);  // cannot happen for ranges <= 8, kept for safety
  return f;
}

__device__ __forceinline__ uint2 adb_bucket_pair(const DevRaw& raw, int64_t scan, float v) {
  return __ldg(raw.bucket_pair + scan * ADB_N_BUCKETS + adb_bucket_of(raw, v));
}

__device__ __forceinline__ uint32_t adb_spectrum_stop(const DevRaw& raw, int64_t scan) {
  return __ldg(&raw.bucket_pair[scan * ADB_N_BUCKETS + (ADB_N_BUCKETS - 1)].y);
}

// lower bound of v inside spectrum `scan`; same result as np.searchsorted(mz[start:stop], v, "left") + start /
// _search_sorted_reference_left (alpharaw_jit.py:53-75).
__device__ __forceinline__ AdbFound adb_spectrum_lower_bound(const DevRaw& raw, int64_t scan, float v) {
  const uint2 r = adb_bucket_pair(raw, scan, v);
  uint32_t lo = r.x, hi = r.y;
  while (hi - lo > 8u) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(raw.mz + mid) < v) lo = mid + 1; else hi = mid;
  }
  AdbFound f = adb_finish_lower_bound(raw.mz, lo, hi, v);
  f.inside = f.idx < r.y;
  return f;
}

// ---- 4-D helpers -----------------------------------------------------------------------------
// np.searchsorted(a, v, "left") on float64
__device__ __forceinline__ int64_t adb_lower_bound_f64(const double* __restrict__ a, int64_t n, double v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// first event in [lo, hi) of a tof row whose push index is >= p (rows are sorted by push)
__device__ __forceinline__ int64_t adb_row_lower_bound(const uint32_t* __restrict__ push, int64_t lo, int64_t hi, int64_t p) {
  if (p <= 0) return lo;
  if (p > 0xFFFFFFFFLL) return hi;
  const uint32_t pv = (uint32_t)p;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(push + mid) < pv) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int64_t adb_wrap0(int64_t v, int64_t limit) {
  if (v < 0) return 0;
  return v < limit ? v : limit;
}

// np.argsort as numba compiles it (numba/misc/quicksort.py, argsort variant): ranges of more than 15 elements are partitioned
// around a median-of-three pivot (NOT stable: the order of exact ties follows the algorithm), ranges of up to 15 elements are
// insertion-sorted (stable).  The fragment orders of the reference (top-k by intensity, m/z order, intensity order of the masked
// fragments) come from it, so for more than 15 elements it is followed step by step; up to 15 the callers keep their stable
// rank counting, which gives the same order.
#if defined(__CUDACC__)
#define ADB_HOSTDEV __host__ __device__ inline
#else
#define ADB_HOSTDEV inline
#endif
#define ADB_NUMBA_SMALL_SORT 15
ADB_HOSTDEV void adb_argsort_numba(const float* v, int n, uint8_t* R) {
  for (int i = 0; i < n; i++) R[i] = (uint8_t)i;
  if (n < 2) return;
  uint8_t lo_stack[16], hi_stack[16];  // the larger side of every split is pushed: depth <= log2(n) + 1
  int sp = 1;
  lo_stack[0] = 0; hi_stack[0] = (uint8_t)(n - 1);
  while (sp > 0) {
    sp--;
    int low = lo_stack[sp], high = hi_stack[sp];
    while (high - low >= ADB_NUMBA_SMALL_SORT) {
      const int mid = (low + high) >> 1;
      uint8_t t;
      if (v[R[mid]] < v[R[low]]) { t = R[low]; R[low] = R[mid]; R[mid] = t; }
      if (v[R[high]] < v[R[mid]]) { t = R[high]; R[high] = R[mid]; R[mid] = t; }
      if (v[R[mid]] < v[R[low]]) { t = R[low]; R[low] = R[mid]; R[mid] = t; }
      const float pivot = v[R[mid]];
      t = R[high]; R[high] = R[mid]; R[mid] = t;
      int i = low, k = high - 1;
      for (;;) {
        while (i < high && v[R[i]] < pivot) i++;
        while (k >= low && pivot < v[R[k]]) k--;
        if (i >= k) break;
        t = R[i]; R[i] = R[k]; R[k] = t;
        i++; k--;
      }
      t = R[i]; R[i] = R[high]; R[high] = t;
      if (high - i > i - low) {
        if (high > i && sp < 16) { lo_stack[sp] = (uint8_t)(i + 1); hi_stack[sp] = (uint8_t)high; sp++; }
        high = i - 1;
      } else {
        if (i > low && sp < 16) { lo_stack[sp] = (uint8_t)low; hi_stack[sp] = (uint8_t)(i - 1); sp++; }
        low = i + 1;
      }
    }
    for (int i = low + 1; i <= high; i++) {
      const uint8_t key = R[i];
      const float x = v[key];
      int k = i;
      while (k > low && x < v[R[k - 1]]) { R[k] = R[k - 1]; k--; }
      R[k] = key;
    }
  }
}

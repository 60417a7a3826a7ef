// alphadia_b200 — device-side common definitions (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/alphadia_b200.h"

#define ADB_MAX_OBS 8            // quad windows one candidate may hit
#define ADB_MAX_LIB_FRAGMENTS 64 // fragments per precursor in the flat library (before top-k)
#define ADB_MAX_MS1_POS 8        // MS1 spectra per DIA cycle
#define ADB_MAX_KERNEL_W 64
#define ADB_ISOTOPE_DIFF 1.0033548350700006
#define ADB_N_BUCKETS 64         // m/z buckets per spectrum in the derived search index

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
  // bucket_idx[scan][b] = first peak (relative to peak_start[scan]) with mz >= bucket_lo + b * bucket_width
  const int32_t* bucket_idx;        // [n_spectra][ADB_N_BUCKETS]
  float bucket_lo, bucket_width, bucket_inv_width;
};

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
void adb_launch_select_ex(const DevRaw& raw, const DevLib& lib, const adb_selection_config& cfg, const double* h_kernel,
                          int kh, int kw, DevCandidatesOut out, int64_t row_begin, int64_t row_end, const int32_t* d_order,
                          uint32_t* d_status, int c_cap, int max_layers, float* d_workspace, int64_t ws_floats_per_slot,
                          int grid, cudaStream_t stream, int* n_launches);
size_t adb_select_smem_bytes(int c_cap, int max_layers, int kw);
int adb_select_resident_ctas(int device, int c_cap, int max_layers, int kw);
int adb_select_slots(void);

void adb_launch_score(const DevRaw& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand,
                      DevScoresOut out, float* d_workspace, int64_t workspace_floats_per_tile, int n_resident_tiles,
                      const int32_t* d_order, uint32_t* d_status, cudaStream_t stream, int* n_launches);
int adb_score_resident_tiles(int device, int top_k);
int64_t adb_score_workspace_floats(int top_k, int64_t c_max);

void adb_launch_fragcomp(int64_t n_windows, const int64_t* d_ws, const int64_t* d_we, const void* d_rt,
                         const int64_t* d_fs, const int64_t* d_fe, const void* d_mz, int is_f64, double rt_tol,
                         double ppm_tol, uint8_t* d_valid, cudaStream_t stream, int* n_launches);

size_t adb_compact_temp_bytes(int64_t n_rows);
void adb_launch_compact_ex(DevCandidatesOut cont, int64_t candidate_count, int* d_flags, int* d_offs, void* d_tmp,
                           size_t tmp_bytes, int64_t* d_lib_row, uint8_t* d_rank, int64_t* d_scan_start,
                           int64_t* d_scan_stop, int64_t* d_scan_center, int64_t* d_frame_start, int64_t* d_frame_stop,
                           int64_t* d_frame_center, int64_t* d_count, cudaStream_t stream, int* n_launches);

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

// One spectrum as a (pointer, length) view with 32-bit offsets.
struct AdbSpectrum {
  const float* mz;
  const float* intensity;
  int n;
};

__device__ __forceinline__ AdbSpectrum adb_spectrum(const DevRaw& raw, int64_t scan) {
  int64_t start = __ldg(raw.peak_start + scan), stop = __ldg(raw.peak_stop + scan);
  AdbSpectrum s;
  s.mz = raw.mz + start;
  s.intensity = raw.intensity + start;
  s.n = (int)(stop - start);
  return s;
}

// Search range [lo, hi] that contains the lower bound of v in one spectrum: the bucket index narrows it to
// ~n/64 peaks with one 8-byte read of an L2-resident table.
__device__ __forceinline__ void adb_bucket_range(const DevRaw& raw, int64_t scan, const AdbSpectrum& s, float v, int& lo, int& hi) {
  int b = (int)((v - raw.bucket_lo) * raw.bucket_inv_width);
  b = max(0, min(b, ADB_N_BUCKETS - 1));
  while (b > 0 && adb_bucket_edge(raw, b) > v) b--;
  while (b < ADB_N_BUCKETS - 1 && adb_bucket_edge(raw, b + 1) <= v) b++;
  const int32_t* row = raw.bucket_idx + scan * ADB_N_BUCKETS;
  lo = (b == 0) ? 0 : __ldg(row + b);
  hi = (b == ADB_N_BUCKETS - 1) ? s.n : __ldg(row + b + 1);
}

// lower bound of v inside one spectrum (first index with mz >= v).  Same result as
// np.searchsorted(mz[start:stop], v, "left") / _search_sorted_reference_left (alpharaw_jit.py:53-75).
__device__ __forceinline__ int adb_spectrum_lower_bound(const DevRaw& raw, int64_t scan, const AdbSpectrum& s, float v) {
  int lo, hi;
  adb_bucket_range(raw, scan, s, v, lo, hi);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(s.mz + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int64_t adb_wrap0(int64_t v, int64_t limit) {
  if (v < 0) return 0;
  return v < limit ? v : limit;
}

// alphadia_b200 — fragment competition kernel + candidate-container compaction, sm_100a.
#include <cub/device/device_scan.cuh>

#include "adb_common.cuh"

#define FC_THREADS 256
#define FC_MAX_FRAG 64

namespace {

// alphadia/fragcomp/fragcomp.py:19-48 _get_fragment_overlap in the array dtype
template <typename T>
__device__ __forceinline__ int fragment_overlap(const T* a, int na, const T* b, int nb, double tol);

template <>
__device__ __forceinline__ int fragment_overlap<float>(const float* a, int na, const float* b, int nb, double tol) {
  int n = 0;
  for (int i = 0; i < na; i++) {
    float ai = a[i];
    for (int j = 0; j < nb; j++) {
      float d = fabsf(__fsub_rn(ai, b[j]));
      double ppm = __dmul_rn((double)__fdiv_rn(d, ai), 1e6);
      n += ppm < tol;
    }
  }
  return n;
}
template <>
__device__ __forceinline__ int fragment_overlap<double>(const double* a, int na, const double* b, int nb, double tol) {
  int n = 0;
  for (int i = 0; i < na; i++) {
    double ai = a[i];
    for (int j = 0; j < nb; j++) {
      double ppm = __dmul_rn(__ddiv_rn(fabs(__dsub_rn(ai, b[j])), ai), 1e6);
      n += ppm < tol;
    }
  }
  return n;
}

// alphadia/fragcomp/fragcomp.py:51-143 _compete_for_fragments: ONE CTA PER DIA WINDOW.  The outer loop over i is
// the greedy order (PSMs sorted by proba) and stays sequential; the inner loop over j is spread over the CTA.
template <typename TR, typename T>
__global__ void __launch_bounds__(FC_THREADS) adb_fragcomp_kernel(int64_t n_windows, const int64_t* __restrict__ ws,
                                                                  const int64_t* __restrict__ we, const TR* __restrict__ rt,
                                                                  const int64_t* __restrict__ fs, const int64_t* __restrict__ fe,
                                                                  const T* __restrict__ mz, double rt_tol, double ppm_tol,
                                                                  uint8_t* valid) {
  __shared__ T frag_i[FC_MAX_FRAG];
  volatile uint8_t* vvalid = valid;
  for (int64_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
    const int64_t s = ws[w], e = we[w];
    for (int64_t i = s; i < e; i++) {
      __syncthreads();  // writes of the previous round are visible; frag_i is free
      if (!vvalid[i]) continue;
      const int64_t fsi = fs[i];
      const int ni = (int)(fe[i] - fsi);
      const bool staged = ni <= FC_MAX_FRAG;
      if (staged)
        for (int t = threadIdx.x; t < ni; t += FC_THREADS) frag_i[t] = mz[fsi + t];
      __syncthreads();
      const TR rti = rt[i];
      const T* a = staged ? frag_i : (mz + fsi);
      for (int64_t j = s + threadIdx.x; j < e; j += FC_THREADS) {
        if (j == i || !vvalid[j]) continue;
        double drt;
        if (sizeof(TR) == 4) drt = (double)fabsf(__fsub_rn((float)rti, (float)rt[j]));
        else drt = fabs(__dsub_rn((double)rti, (double)rt[j]));
        if (drt < rt_tol) {
          const int64_t fsj = fs[j];
          int ov = fragment_overlap<T>(a, ni, mz + fsj, (int)(fe[j] - fsj), ppm_tol);
          if (ov >= 3) vvalid[j] = 0;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void adb_flag_kernel(const float* __restrict__ score, int64_t n, int* __restrict__ flags) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flags[t] = score[t] > 0.f;  // CandidateContainer.get_candidate_df_data, config_df.py:270-284
}

__global__ void adb_scatter_kernel(DevCandidatesOut c, int64_t candidate_count, const int* __restrict__ flags,
                                   const int* __restrict__ offs, int64_t* lib_row, uint8_t* rank, int64_t* scan_start,
                                   int64_t* scan_stop, int64_t* scan_center, int64_t* frame_start, int64_t* frame_stop,
                                   int64_t* frame_center, uint32_t* precursor_idx, float* score, int64_t* count) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c.n_rows) return;
  if (flags[t]) {
    int o = offs[t];
    lib_row[o] = t / candidate_count;
    precursor_idx[o] = c.precursor_idx[t];
    score[o] = c.score[t];
    rank[o] = c.rank[t];
    scan_start[o] = c.scan_start[t]; scan_stop[o] = c.scan_stop[t]; scan_center[o] = c.scan_center[t];
    frame_start[o] = c.frame_start[t]; frame_stop[o] = c.frame_stop[t]; frame_center[o] = c.frame_center[t];
  }
  if (t == c.n_rows - 1) *count = (int64_t)offs[t] + flags[t];
}

}  // namespace

void adb_launch_fragcomp(int64_t n_windows, const int64_t* d_ws, const int64_t* d_we, const void* d_rt,
                         const int64_t* d_fs, const int64_t* d_fe, const void* d_mz, int is_f64, double rt_tol,
                         double ppm_tol, uint8_t* d_valid, cudaStream_t stream, int* n_launches) {
  if (n_windows <= 0) return;
  unsigned grid = (unsigned)(n_windows < 65535 ? n_windows : 65535);
  // is_f64: bit 0 = rt is float64, bit 1 = fragment_mz is float64
  switch (is_f64 & 3) {
    case 0: adb_fragcomp_kernel<float, float><<<grid, FC_THREADS, 0, stream>>>(n_windows, d_ws, d_we, (const float*)d_rt, d_fs, d_fe, (const float*)d_mz, rt_tol, ppm_tol, d_valid); break;
    case 1: adb_fragcomp_kernel<double, float><<<grid, FC_THREADS, 0, stream>>>(n_windows, d_ws, d_we, (const double*)d_rt, d_fs, d_fe, (const float*)d_mz, rt_tol, ppm_tol, d_valid); break;
    case 2: adb_fragcomp_kernel<float, double><<<grid, FC_THREADS, 0, stream>>>(n_windows, d_ws, d_we, (const float*)d_rt, d_fs, d_fe, (const double*)d_mz, rt_tol, ppm_tol, d_valid); break;
    default: adb_fragcomp_kernel<double, double><<<grid, FC_THREADS, 0, stream>>>(n_windows, d_ws, d_we, (const double*)d_rt, d_fs, d_fe, (const double*)d_mz, rt_tol, ppm_tol, d_valid); break;
  }
  if (n_launches) (*n_launches)++;
}

// scratch: flags[n_rows] + offs[n_rows] ints + cub temp storage, all caller-provided
size_t adb_compact_temp_bytes(int64_t n_rows) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int*)nullptr, (int*)nullptr, (int)n_rows);
  return tmp;
}

void adb_launch_compact_ex(DevCandidatesOut cont, int64_t candidate_count, int* d_flags, int* d_offs, void* d_tmp,
                           size_t tmp_bytes, int64_t* d_lib_row, uint8_t* d_rank, int64_t* d_scan_start,
                           int64_t* d_scan_stop, int64_t* d_scan_center, int64_t* d_frame_start, int64_t* d_frame_stop,
                           int64_t* d_frame_center, uint32_t* d_precursor_idx, float* d_score, int64_t* d_count,
                           cudaStream_t stream, int* n_launches) {
  if (cont.n_rows <= 0) return;
  unsigned blocks = (unsigned)((cont.n_rows + 255) / 256);
  adb_flag_kernel<<<blocks, 256, 0, stream>>>(cont.score, cont.n_rows, d_flags);
  cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flags, d_offs, (int)cont.n_rows, stream);
  adb_scatter_kernel<<<blocks, 256, 0, stream>>>(cont, candidate_count, d_flags, d_offs, d_lib_row, d_rank, d_scan_start,
                                                 d_scan_stop, d_scan_center, d_frame_start, d_frame_stop, d_frame_center,
                                                 d_precursor_idx, d_score, d_count);
  if (n_launches) (*n_launches) += 3;
}

// alphadia_b200 — fragment competition kernel + candidate-container compaction, sm_100a.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

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

// ---- fragment competition as a conflict graph ----------------------------------------------------------------------
// The veto "i removes j" (fragcomp.py:122-143) depends only on the pair: |rt_i - rt_j| < rt_tol and >= 3 overlapping fragments
// (asymmetric: the ppm distance is relative to i's m/z).  All pairs inside the RT tolerance are evaluated in parallel from
// an RT-sorted view of every DIA window; what stays sequential is the greedy pass "in proba order, every PSM that is still
// valid removes its out-neighbours", which touches only the (few) PSMs that have an edge at all.
__device__ __forceinline__ uint32_t fc_ordered_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void fc_segment_kernel(int64_t n_windows, const int64_t* __restrict__ ws, const int64_t* __restrict__ we, uint32_t* seg_of) {
  const int64_t w = blockIdx.x;
  if (w >= n_windows) return;
  for (int64_t i = ws[w] + threadIdx.x; i < we[w]; i += blockDim.x) seg_of[i] = (uint32_t)w;
}

template <typename TR>
__global__ void fc_keys_kernel(int64_t n, const TR* __restrict__ rt, const uint32_t* __restrict__ seg_of, uint64_t* keys, int32_t* vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = ((uint64_t)seg_of[i] << 32) | fc_ordered_bits((float)rt[i]);
  vals[i] = (int32_t)i;
}

// sb[w] = first sorted position of segment w (w = n_windows: end of the covered PSMs)
__global__ void fc_bounds_kernel(int64_t n_windows, int64_t n, const uint64_t* __restrict__ keys, int64_t* sb) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w > n_windows) return;
  const uint64_t v = (uint64_t)w << 32;
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < v) lo = mid + 1; else hi = mid; }
  sb[w] = lo;
}

// sorted range [lo_i, lo_i + cnt_i) of PSM i: every PSM of its window whose RT can be inside the tolerance (a superset: the
// float32 view of the RT and a widened interval; the exact test in the array dtype follows per pair)
template <typename TR>
__global__ void fc_range_kernel(int64_t n, int64_t n_windows, const TR* __restrict__ rt, const uint32_t* __restrict__ seg_of,
                                const uint64_t* __restrict__ keys, const int64_t* __restrict__ sb, double rt_tol, int64_t* lo_out,
                                int64_t* cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t w = seg_of[i];
  const double r = (double)rt[i];
  if ((int64_t)w >= n_windows || !(r == r) || !(rt_tol > 0)) { lo_out[i] = 0; cnt[i] = 0; return; }
  const double pad = rt_tol * 1e-6 + fabs(r) * 1e-6;
  const float lf = nextafterf((float)(r - rt_tol - pad), -INFINITY), uf = nextafterf((float)(r + rt_tol + pad), INFINITY);
  const uint64_t kl = ((uint64_t)w << 32) | fc_ordered_bits(lf), ku = ((uint64_t)w << 32) | fc_ordered_bits(uf);
  int64_t a = sb[w], b = sb[w + 1];
  int64_t lo = a, hi = b;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < kl) lo = mid + 1; else hi = mid; }
  const int64_t first = lo;
  hi = b;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] <= ku) lo = mid + 1; else hi = mid; }
  lo_out[i] = first;
  cnt[i] = lo - first;
}

// edges of PSM i: the j it would remove; ec[i] = their number, stored from off[i]
template <typename TR, typename T>
__global__ void fc_edges_kernel(int64_t n, const TR* __restrict__ rt, const int64_t* __restrict__ fs, const int64_t* __restrict__ fe,
                                const T* __restrict__ mz, const int32_t* __restrict__ order, const int64_t* __restrict__ lo,
                                const int64_t* __restrict__ cnt, const int64_t* __restrict__ off, double rt_tol, double ppm_tol,
                                int32_t* edges, int32_t* ec) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = cnt[i];
  int32_t m = 0;
  if (c > 1) {
    const TR rti = rt[i];
    const int64_t fsi = fs[i];
    const int ni = (int)(fe[i] - fsi);
    int32_t* out = edges + off[i];
    const int32_t* cand = order + lo[i];
    for (int64_t t = 0; t < c; t++) {
      const int64_t j = cand[t];
      if (j == i) continue;
      double drt;
      if (sizeof(TR) == 4) drt = (double)fabsf(__fsub_rn((float)rti, (float)rt[j]));
      else drt = fabs(__dsub_rn((double)rti, (double)rt[j]));
      if (drt < rt_tol) {
        const int64_t fsj = fs[j];
        if (fragment_overlap<T>(mz + fsi, ni, mz + fsj, (int)(fe[j] - fsj), ppm_tol) >= 3) out[m++] = (int32_t)j;
      }
    }
  }
  ec[i] = m;
}

// one warp per DIA window: the greedy pass in proba order over the PSMs that have an edge
__global__ void fc_resolve_kernel(int64_t n_windows, const int64_t* __restrict__ ws, const int64_t* __restrict__ we,
                                  const int32_t* __restrict__ ec, const int64_t* __restrict__ off, const int32_t* __restrict__ edges,
                                  uint8_t* valid) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_windows) return;
  volatile uint8_t* vvalid = valid;
  const int64_t s = ws[w], e = we[w];
  for (int64_t base = s; base < e; base += 32) {
    const int64_t i = base + lane;
    unsigned has = __ballot_sync(0xFFFFFFFFu, i < e && ec[i] > 0);
    while (has) {
      const int b = __ffs(has) - 1;
      has &= has - 1;
      const int64_t i0 = base + b;
      if (vvalid[i0]) {
        const int32_t m = ec[i0];
        const int32_t* ed = edges + off[i0];
        for (int32_t t = lane; t < m; t += 32) vvalid[ed[t]] = 0;
      }
      __syncwarp();
    }
  }
}

__global__ void adb_flag_kernel(const float* __restrict__ score, int64_t n, float cutoff, int* __restrict__ flags) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // CandidateContainer.get_candidate_df_data, config_df.py:270-284 (score > 0), optionally followed by the handler's score
  // cutoff (extraction_handler.py:177-202, score > cutoff on the float32 column)
  if (t < n) flags[t] = score[t] > 0.f && score[t] > cutoff;
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

namespace {
// grow-only scratch of the conflict-graph passes, one per device (cudaMalloc / cudaFree per call cost more than the kernels)
struct FcScratch { char* ptr = nullptr; size_t bytes = 0; };
FcScratch g_fc_scratch[2][64];
char* fc_scratch(int which, size_t need) {
  int dev = 0;
  cudaGetDevice(&dev);
  FcScratch& s = g_fc_scratch[which][dev & 63];
  if (need <= s.bytes) return s.ptr;
  if (s.ptr) cudaFree(s.ptr);
  s.ptr = nullptr; s.bytes = 0;
  const size_t cap = need + need / 4 + 4096;
  if (cudaMalloc((void**)&s.ptr, cap) != cudaSuccess) { cudaGetLastError(); s.ptr = nullptr; return nullptr; }
  s.bytes = cap;
  return s.ptr;
}

template <typename TR, typename T>
int fragcomp_graph_typed(int64_t n_windows, const int64_t* d_ws, const int64_t* d_we, int64_t n, const TR* d_rt, const int64_t* d_fs,
                         const int64_t* d_fe, const T* d_mz, double rt_tol, double ppm_tol, uint8_t* d_valid, size_t pair_cap,
                         cudaStream_t st, int* n_launches) {
  const size_t N = (size_t)n;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int64_t)n, 0, 64, st);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n + 1), st);
  const size_t bytes = al(4 * N) + 2 * al(8 * N) + 2 * al(4 * N) + al(8 * ((size_t)n_windows + 2)) + 2 * al(8 * (N + 1)) +
                       al(8 * (N + 1)) + al(4 * N) + al(std::max(sort_tmp, scan_tmp)) + 256;
  char* base = fc_scratch(0, bytes);
  if (!base) return 2;
  char* p = base;
  auto take = [&](size_t b) { char* r = p; p += al(b); return r; };
  uint32_t* seg_of = (uint32_t*)take(4 * N);
  uint64_t* k_in = (uint64_t*)take(8 * N);
  uint64_t* k_out = (uint64_t*)take(8 * N);
  int32_t* v_in = (int32_t*)take(4 * N);
  int32_t* v_out = (int32_t*)take(4 * N);
  int64_t* sb = (int64_t*)take(8 * ((size_t)n_windows + 2));
  int64_t* lo = (int64_t*)take(8 * (N + 1));
  int64_t* cnt = (int64_t*)take(8 * (N + 1));
  int64_t* off = (int64_t*)take(8 * (N + 1));
  int32_t* ec = (int32_t*)take(4 * N);
  void* tmp = take(std::max(sort_tmp, scan_tmp));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  int rc = 0;
  int32_t* edges = nullptr;
  // uncovered PSMs get segment n_windows: they sort behind every window and take no part
  cudaMemsetAsync(seg_of, 0xFF, 4 * N, st);
  fc_segment_kernel<<<(unsigned)std::min<int64_t>(n_windows, 65535), 256, 0, st>>>(n_windows, d_ws, d_we, seg_of);
  fc_keys_kernel<TR><<<blocks, 256, 0, st>>>(n, d_rt, seg_of, k_in, v_in);
  size_t tb = sort_tmp;
  cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, v_out, (int64_t)n, 0, 64, st);
  fc_bounds_kernel<<<(unsigned)((n_windows + 1 + 255) / 256), 256, 0, st>>>(n_windows, n, k_out, sb);
  fc_range_kernel<TR><<<blocks, 256, 0, st>>>(n, n_windows, d_rt, seg_of, k_out, sb, rt_tol, lo, cnt);
  cudaMemsetAsync(cnt + n, 0, 8, st);
  tb = scan_tmp;
  cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, (int)(n + 1), st);
  int64_t total = 0;
  if (cudaMemcpyAsync(&total, off + n, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
  if (!rc && (size_t)total > pair_cap) rc = 2;  // pathological co-elution: the caller falls back to the window-serial kernel
  if (!rc && !(edges = (int32_t*)fc_scratch(1, 4 * (size_t)std::max<int64_t>(total, 1)))) rc = 2;
  if (!rc) {
    fc_edges_kernel<TR, T><<<blocks, 256, 0, st>>>(n, d_rt, d_fs, d_fe, d_mz, v_out, lo, cnt, off, rt_tol, ppm_tol, edges, ec);
    fc_resolve_kernel<<<(unsigned)((n_windows * 32 + 127) / 128), 128, 0, st>>>(n_windows, d_ws, d_we, ec, off, edges, d_valid);
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
    if (n_launches) *n_launches += 12;
  }
  return rc;
}
}  // namespace

// conflict-graph formulation; returns 0 on success, 1 on a CUDA error, 2 when the pair list would not fit (the caller then
// uses adb_launch_fragcomp, the window-serial kernel)
int adb_run_fragcomp_graph(int64_t n_windows, const int64_t* d_ws, const int64_t* d_we, int64_t n_psm, const void* d_rt,
                           const int64_t* d_fs, const int64_t* d_fe, const void* d_mz, int is_f64, double rt_tol, double ppm_tol,
                           uint8_t* d_valid, size_t pair_cap, cudaStream_t stream, int* n_launches) {
  if (n_windows <= 0 || n_psm <= 0) return 0;
  if (n_psm >= 2000000000LL) return 2;
  switch (is_f64 & 3) {
    case 0: return fragcomp_graph_typed<float, float>(n_windows, d_ws, d_we, n_psm, (const float*)d_rt, d_fs, d_fe, (const float*)d_mz, rt_tol, ppm_tol, d_valid, pair_cap, stream, n_launches);
    case 1: return fragcomp_graph_typed<double, float>(n_windows, d_ws, d_we, n_psm, (const double*)d_rt, d_fs, d_fe, (const float*)d_mz, rt_tol, ppm_tol, d_valid, pair_cap, stream, n_launches);
    case 2: return fragcomp_graph_typed<float, double>(n_windows, d_ws, d_we, n_psm, (const float*)d_rt, d_fs, d_fe, (const double*)d_mz, rt_tol, ppm_tol, d_valid, pair_cap, stream, n_launches);
    default: return fragcomp_graph_typed<double, double>(n_windows, d_ws, d_we, n_psm, (const double*)d_rt, d_fs, d_fe, (const double*)d_mz, rt_tol, ppm_tol, d_valid, pair_cap, stream, n_launches);
  }
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
                           cudaStream_t stream, int* n_launches, float score_cutoff) {
  if (cont.n_rows <= 0) return;
  unsigned blocks = (unsigned)((cont.n_rows + 255) / 256);
  adb_flag_kernel<<<blocks, 256, 0, stream>>>(cont.score, cont.n_rows, score_cutoff, d_flags);
  cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flags, d_offs, (int)cont.n_rows, stream);
  adb_scatter_kernel<<<blocks, 256, 0, stream>>>(cont, candidate_count, d_flags, d_offs, d_lib_row, d_rank, d_scan_start,
                                                 d_scan_stop, d_scan_center, d_frame_start, d_frame_stop, d_frame_center,
                                                 d_precursor_idx, d_score, d_count);
  if (n_launches) (*n_launches) += 3;
}

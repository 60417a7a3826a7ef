// alphadia_b200 — C ABI implementation: handles, H2D/D2H staging, launches (see include/alphadia_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "adb_common.cuh"

#ifndef ADB_SCORE_DP_BATCH
#define ADB_SCORE_DP_BATCH (3 << 20)
#define ADB_SCORE_DP_BATCH_MAX (1 << 22)  // candidates per batch of the data-parallel scoring passes (x 72 rows < 2^32)
#endif
#ifndef ADB_SCORE_BLOCKS
#define ADB_SCORE_BLOCKS 4  // row blocks of a scoring call whose D2H overlaps the next block's kernel (<= ADB_MAX_BLOCKS)
#endif
#ifndef ADB_RAGGED_BLOCKS
#define ADB_RAGGED_BLOCKS 5  // same for ragged results: a block's copy starts one block late (<= ADB_MAX_BLOCKS)
#endif
#ifndef ADB_RAGGED_BLOCK_RATIO
#define ADB_RAGGED_BLOCK_RATIO 0.7  // size of block k + 1 relative to block k
#endif
#ifndef ADB_MAX_BLOCKS
#define ADB_MAX_BLOCKS 15  // 4 key bits
#endif

namespace {

thread_local std::string g_error;

int fail(const std::string& msg) {
  g_error = msg;
  return 1;
}

}  // namespace
int adb_set_error(const std::string& msg) { return fail(msg); }  // for the other translation units of the library
namespace {

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                  std::to_string(__LINE__) + ")");                                            \
  } while (0)

struct DeviceBuffer {  // grow-only cached device allocation
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    size_t cap = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, cap);
    if (e != cudaSuccess) return fail(std::string("cudaMalloc(") + std::to_string(cap) + ") failed: " + cudaGetErrorString(e));
    bytes = cap;
    return 0;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return (T*)ptr; }
};

template <typename T>
int upload(const T* host, int64_t n, T** dev, std::vector<void*>& allocs, int64_t& total, cudaStream_t stream, int64_t pad = 0) {
  size_t bytes = sizeof(T) * (size_t)((n > 0 ? n : 1) + pad);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
  allocs.push_back(p);
  total += (int64_t)bytes;
  if (pad > 0) cudaMemsetAsync((char*)p + sizeof(T) * (size_t)(n > 0 ? n : 0), 0, sizeof(T) * (size_t)pad, stream);
  if (n > 0) {
    e = cudaMemcpyAsync(p, host, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return fail(std::string("cudaMemcpy H2D failed: ") + cudaGetErrorString(e));
  }
  *dev = (T*)p;
  return 0;
}

}  // namespace

struct adb_library {
  int device = 0;
  DevLib dev{};
  std::vector<void*> allocs;
  int64_t bytes = 0;
  int max_lib_fragments = 0;
  void* pool = nullptr;   // the single device allocation behind `dev`
  size_t pool_bytes = 0;
};

namespace {
// The operators upload a library batch per call; cudaMalloc/cudaFree of a 0.5 GB pool cost 10-100 ms and synchronise the
// device, so one released pool per device is kept for the next adb_library_create.
struct LibraryPoolCache { void* ptr = nullptr; size_t bytes = 0; };
LibraryPoolCache g_pool_cache[64];
std::mutex g_pool_mutex;
}  // namespace

struct adb_rawfile {
  int device = 0;
  cudaStream_t stream = nullptr;
  DevRaw dev{};
  int is4d = 0;    // timsTOF handle: dev4 is set instead of dev
  DevRaw4 dev4{};
  std::vector<void*> allocs;
  int64_t bytes = 0;
  std::vector<float> rt_host;  // host copy of rt_values (to bound the cycle window on the host)
  std::vector<double> rt4_host, mob4_host;  // 4-D: rt_values / mobility_values
  uint32_t* d_status = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 4,5 bracket the main kernel
  float h2d_ms = 0, kernel_ms = 0, d2h_ms = 0, main_kernel_ms = 0;
  int launches = 0;
  int sm_count = 148;
  size_t ws_budget = 0;
  // cached workspaces
  DeviceBuffer kern, order_keys, order_vals, order_tmp, sel_ws;
  DeviceBuffer cont;       // candidate container (9 columns)
  DeviceBuffer cand_in;    // compacted candidates (8 columns)
  DeviceBuffer flags, offs, scan_tmp, count;
  DeviceBuffer scores;     // score outputs
  DeviceBuffer score_ws;
  DeviceBuffer dp_plan;    // per-batch plan arrays of the data-parallel scoring passes
  float* dp_cube = nullptr;   // = score_ws.ptr, the batch workspace of the passes
  size_t dp_cube_floats = 0;
  DeviceBuffer staging;    // generic H2D staging
  DeviceBuffer extent;     // 4-D: max scan / cycle extent of the resident candidates
  // resident state
  int64_t cont_rows = 0, cont_count = 0;
  DevCandidatesOut d_cont{};
  int64_t n_cand = 0;
  DevCandidatesIn d_cand{};
  const uint32_t* d_cand_pidx = nullptr;  // compacted precursor_idx / score of the resident candidates
  const float* d_cand_score = nullptr;
  cudaStream_t copy_stream = nullptr;     // D2H of finished scoring chunks overlaps the next chunk's kernel
  cudaEvent_t chunk_ev[ADB_MAX_BLOCKS + 1] = {};  // [ADB_MAX_BLOCKS]: end-of-copies marker
  DevScoresOut d_scores{};
  int64_t scores_n = 0;
  int scores_k = 0;
  // ragged results (adb_score_candidates_ragged): device-compacted tables + per-block running totals
  DeviceBuffer rag_rows, rag_frags, rag_scan, rag_scan_tmp, rag_tot;
  int64_t* rag_host_tot = nullptr;  // pinned [2 * (ADB_RAGGED_MAX_BLOCKS + 1)]
  cudaEvent_t rag_ev[ADB_MAX_BLOCKS + 1] = {};
};

namespace {

int set_device(int device) {
  CUDA_TRY(cudaSetDevice(device));
  return 0;
}

int check_status(adb_rawfile* raw, const char* where) {
  uint32_t st = 0;
  CUDA_TRY(cudaMemcpyAsync(&st, raw->d_status, sizeof(st), cudaMemcpyDeviceToHost, raw->stream));
  CUDA_TRY(cudaStreamSynchronize(raw->stream));
  CUDA_TRY(cudaGetLastError());
  if (st == 0) return 0;
  std::string msg = std::string(where) + ": unsupported input on device:";
  if (st & ADB_STATUS_TOO_MANY_OBS) msg += " a precursor overlaps more than " + std::to_string(ADB_MAX_OBS) + " quadrupole windows;";
  if (st & ADB_STATUS_TOO_MANY_LIB_FRAGMENTS) msg += " a precursor has more than " + std::to_string(ADB_MAX_LIB_FRAGMENTS) + " library fragments;";
  if (st & ADB_STATUS_SCRATCH_OVERFLOW) msg += " a candidate/precursor window exceeds the device scratch;";
  if (st & ADB_STATUS_TOO_MANY_PEAKS) msg += " a spectrum segment exceeds the index limits;";
  CUDA_TRY(cudaMemsetAsync(raw->d_status, 0, sizeof(uint32_t), raw->stream));
  return fail(msg);
}

// candidate container columns carved from one buffer
DevCandidatesOut carve_container(void* base, int64_t n) {
  DevCandidatesOut c{};
  c.n_rows = n;
  size_t N = (size_t)n;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  c.precursor_idx = (uint32_t*)take(4 * N);
  c.score = (float*)take(4 * N);
  c.scan_center = (uint32_t*)take(4 * N);
  c.scan_start = (uint32_t*)take(4 * N);
  c.scan_stop = (uint32_t*)take(4 * N);
  c.frame_center = (uint32_t*)take(4 * N);
  c.frame_start = (uint32_t*)take(4 * N);
  c.frame_stop = (uint32_t*)take(4 * N);
  c.rank = (uint8_t*)take(N);
  return c;
}
size_t container_bytes(int64_t n) { return 9 * (((size_t)n * 4 + 255) & ~(size_t)255) + 512; }

struct CandInPtrs {
  int64_t *lib_row, *scan_start, *scan_stop, *scan_center, *frame_start, *frame_stop, *frame_center;
  uint8_t* rank;
  uint32_t* precursor_idx;  // compacted container columns kept for adb_fetch_candidate_table
  float* score;
};
CandInPtrs carve_cand_in(void* base, int64_t n) {
  CandInPtrs c{};
  size_t N = (size_t)n;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  c.lib_row = (int64_t*)take(8 * N);
  c.scan_start = (int64_t*)take(8 * N);
  c.scan_stop = (int64_t*)take(8 * N);
  c.scan_center = (int64_t*)take(8 * N);
  c.frame_start = (int64_t*)take(8 * N);
  c.frame_stop = (int64_t*)take(8 * N);
  c.frame_center = (int64_t*)take(8 * N);
  c.rank = (uint8_t*)take(N);
  c.precursor_idx = (uint32_t*)take(4 * N);
  c.score = (float*)take(4 * N);
  return c;
}
size_t cand_in_bytes(int64_t n) { return 10 * (((size_t)n * 8 + 255) & ~(size_t)255) + 512; }

DevScoresOut carve_scores(void* base, int64_t n, int k) {
  DevScoresOut s{};
  size_t N = (size_t)n, K = (size_t)k;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  s.features = (float*)take(4 * N * ADB_NUM_FEATURES);
  s.fragment_mz_library = (float*)take(4 * N * K);
  s.fragment_mz = (float*)take(4 * N * K);
  s.fragment_mz_observed = (float*)take(4 * N * K);
  s.fragment_height = (float*)take(4 * N * K);
  s.fragment_intensity = (float*)take(4 * N * K);
  s.fragment_mass_error = (float*)take(4 * N * K);
  s.fragment_correlation = (float*)take(4 * N * K);
  s.fragment_position = (uint8_t*)take(N * K);
  s.fragment_number = (uint8_t*)take(N * K);
  s.fragment_type = (uint8_t*)take(N * K);
  s.fragment_charge = (uint8_t*)take(N * K);
  s.fragment_loss_type = (uint8_t*)take(N * K);
  s.valid = (uint8_t*)take(N);
  return s;
}
size_t scores_bytes(int64_t n, int k) {
  size_t N = (size_t)n, K = (size_t)k;
  return 4 * N * ADB_NUM_FEATURES + 7 * 4 * N * K + 5 * N * K + N + 15 * 256;
}

// m/z range of the file: spectra are sorted, so first/last peaks bound it
__global__ void mz_range_kernel(DevRaw raw, float* out /* [2] = {min, max}, pre-set to {+inf, -inf} */) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= raw.n_spectra) return;
  int64_t s = raw.peak_start[i], e = raw.peak_stop[i];
  if (e <= s) return;
  float lo = raw.mz[s], hi = raw.mz[e - 1];
  // positive floats order like their bit patterns
  atomicMin((unsigned int*)out, __float_as_uint(fmaxf(lo, 0.f)));
  atomicMax((unsigned int*)(out + 1), __float_as_uint(fmaxf(hi, 0.f)));
}

__global__ void bucket_index_kernel(DevRaw raw, uint2* table) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= raw.n_spectra * ADB_N_BUCKETS) return;
  int64_t scan = t / ADB_N_BUCKETS;
  int b = (int)(t % ADB_N_BUCKETS);
  int64_t s = raw.peak_start[scan], e = raw.peak_stop[scan];
  uint32_t first = (uint32_t)s, last = (uint32_t)e;
  for (int side = 0; side < 2; side++) {
    int bb = b + side;  // lower edge of bucket b, lower edge of bucket b + 1
    if (bb == 0 || bb == ADB_N_BUCKETS) continue;
    float edge = adb_bucket_edge(raw, bb);
    int64_t lo = s, hi = e;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (raw.mz[mid] < edge) lo = mid + 1; else hi = mid;
    }
    if (side == 0) first = (uint32_t)lo; else last = (uint32_t)lo;
  }
  table[t] = make_uint2(first, last);
}

__global__ void order_key_kernel(DevRaw raw, DevLib lib, uint64_t* keys, int32_t* vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lib.n_precursors) return;
  float mz = lib.mz[i];
  uint32_t win = 0xFFFFu;
  for (int64_t j = 0; j < raw.cycle_len; j++)
    if ((double)mz <= raw.cycle[2 * j + 1] && (double)mz >= raw.cycle[2 * j]) { win = (uint32_t)j; break; }
  float rt = lib.rt[i];
  uint32_t rb = __float_as_uint(rt);
  rb = (rb & 0x80000000u) ? ~rb : (rb | 0x80000000u);  // order-preserving float -> uint
  keys[i] = ((uint64_t)win << 32) | rb;
  vals[i] = (int32_t)i;
}

// processing order of the candidates: (output row block, quad window, time bucket of 32 cycles, cost class).  Candidates
// that are resident together read the same spectra (L2 reuse) and the candidates of one CTA round cost about the same
// (the scoring kernel runs its tiles in lock step).  Results do not depend on the order (disjoint output rows).
struct RowBlocks { int n; int64_t start[ADB_MAX_BLOCKS + 2]; };  // row blocks [start[k], start[k + 1]) of a scoring call

__global__ void score_order_key_kernel(DevRaw raw, DevLib lib, DevCandidatesIn cand, RowBlocks blocks, int n_iso,
                                       uint64_t* keys, int32_t* vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cand.n) return;
  const int64_t row = cand.lib_row[i];
  const float mz = lib.mz[row];
  const double q0 = (double)mz - 0.5, q1 = (double)mz + (double)(n_iso - 1) * ADB_ISOTOPE_DIFF / (double)lib.charge[row] + 0.5;
  uint32_t win = 0xFFFFu;
  int nobs = 0;
  for (int64_t j = 0; j < raw.cycle_len; j++) {
    const double lo = raw.cycle[2 * j], hi = raw.cycle[2 * j + 1];
    if (win == 0xFFFFu && (double)mz <= hi && (double)mz >= lo) win = (uint32_t)j;
    nobs += (q0 <= hi) && (q1 >= lo);
  }
  const int64_t fs = cand.frame_start[i], fe = cand.frame_stop[i];
  const int64_t L = raw.cycle_len;
  const uint64_t bucket = fs < 0 ? 0ull : (uint64_t)min((long long)(fs / (L * 32)), 0xFFFFFFLL);
  const long long cyc = max((long long)(fe / L - fs / L), 0LL);
  // cost class = (observations, cycles): the two candidates that share a warp and the tiles of a lock-step round then
  // run loops of identical trip counts (less divergence than sorting by the product)
  const uint64_t cost = ((uint64_t)min(max(nobs, 1), 3) << 6) | (uint64_t)min(cyc, 63LL);
  uint64_t blk = 0;
  for (int k = 1; k < blocks.n; k++) blk += i >= blocks.start[k];
  keys[i] = (blk << 48) | ((uint64_t)win << 32) | (bucket << 8) | cost;
  vals[i] = (int32_t)i;
}



// ---- CSR transpose of a timsTOF raw file (push-major -> tof-major), alphadia/raw_data/bruker.py:155-274 ----------
__global__ void transpose_push_of_event_kernel(const int64_t* __restrict__ push_indptr, int64_t n_push, uint32_t* push_of_event,
                                               uint32_t* idx) {
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per push
  const int lane = threadIdx.x & 31;
  if (p >= n_push) return;
  for (int64_t i = push_indptr[p] + lane; i < push_indptr[p + 1]; i += 32) { push_of_event[i] = (uint32_t)p; idx[i] = (uint32_t)i; }
}

__global__ void transpose_gather_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ push_of_event,
                                        const uint16_t* __restrict__ values, int64_t n, uint32_t* push_out, uint16_t* values_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t i = idx[j];
  push_out[j] = push_of_event[i];
  values_out[j] = values[i];
}

__global__ void transpose_indptr_kernel(const uint32_t* __restrict__ sorted_tof, int64_t n, int64_t n_tof, int64_t* tof_indptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tof) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted_tof[mid] < t) lo = mid + 1; else hi = mid;
  }
  tof_indptr[t] = lo;
}

// ---- m/z-major index of a 3-D raw file ---------------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_bits(float v) {
  uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one warp per spectrum: key = (cycle position, m/z), value = peak index
__global__ void mzindex_keys_kernel(DevRaw raw, uint64_t* keys, uint32_t* vals) {
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= raw.n_spectra) return;
  const uint64_t pos = (uint64_t)(s % raw.cycle_len) << 32;
  for (int64_t i = raw.peak_start[s] + lane; i < raw.peak_stop[s]; i += 32) {
    keys[i] = pos | ordered_bits(raw.mz[i]);
    vals[i] = (uint32_t)i;
  }
}

__global__ void mzindex_segments_kernel(const uint64_t* keys, int64_t n, int64_t L, int64_t* pos_start) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > L) return;
  const uint64_t v = (uint64_t)p << 32;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  pos_start[p] = lo;
}

// ---- time-blocked m/z index of a 3-D raw file (candidate scoring) -------------------------------------------
// one warp per spectrum: key = (cycle position * n_time_blocks + time block, m/z), value = peak index
__global__ void tbindex_keys_kernel(DevRaw raw, int ntb, uint64_t* keys, uint32_t* vals) {
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= raw.n_spectra) return;
  const uint64_t seg = (uint64_t)((s % raw.cycle_len) * ntb + (s / raw.cycle_len) / ADB_TB_CYCLES) << 32;
  for (int64_t i = raw.peak_start[s] + lane; i < raw.peak_stop[s]; i += 32) {
    keys[i] = seg | ordered_bits(raw.mz[i]);
    vals[i] = (uint32_t)i;
  }
}

// sorted peak j of an m/z-sorted index as one 16-byte record; the `pad` records behind the last peak are padding with a huge
// m/z that is never inside a window
__global__ void tbindex_gather_kernel(DevRaw raw, const uint64_t* keys, const uint32_t* vals, int64_t n, float4* pk, uint64_t n_segments,
                                      int pad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n + pad) return;
  if (j >= n || (keys[j] >> 32) >= n_segments) { pk[j] = make_float4(3.0e38f, 0.f, __uint_as_float(0xFFFFFFFFu), 0.f); return; }
  const uint32_t i = vals[j];
  int64_t lo = 0, hi = raw.n_spectra;  // spectrum of peak i: last s with peak_start[s] <= i and i < peak_stop[s]
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (raw.peak_stop[mid] <= (int64_t)i) lo = mid + 1; else hi = mid;
  }
  pk[j] = make_float4(raw.mz[i], raw.intensity[i], __uint_as_float((uint32_t)(lo / raw.cycle_len)), 0.f);
}

// tb_bucket[seg][b] = first sorted peak whose key is >= (seg, lower edge of bucket b); b == nb: end of the segment
__global__ void tbindex_bucket_kernel(const uint64_t* keys, int64_t n, int64_t n_seg, int nb, float edge_lo, float edge_width,
                                      uint32_t* table) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = nb + 1;
  if (t >= n_seg * per) return;
  const int64_t seg = t / per;
  const int b = (int)(t - seg * per);
  uint64_t v;
  if (b == 0) v = (uint64_t)seg << 32;
  else if (b == nb) v = (uint64_t)(seg + 1) << 32;
  else v = ((uint64_t)seg << 32) | ordered_bits(adb_bucket_edge(edge_lo, edge_width, b));
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  table[t] = (uint32_t)lo;
}

// upper bound of the selection cycle window: jitclasses/utils.py:62-70 over every possible precursor RT
int64_t cycle_window_upper_bound(const adb_rawfile* raw, double rt_tol, int64_t kernel_size) {
  const std::vector<float>& rt = raw->rt_host;
  const int64_t n = (int64_t)rt.size(), L = raw->dev.cycle_len;
  int64_t max_span = 0, j = 0;
  for (int64_t i = 0; i < n; i++) {
    if (j < i) j = i;
    while (j < n && (double)rt[j] <= (double)rt[i] + 2.0 * rt_tol + 1e-3) j++;
    max_span = std::max(max_span, j - i);
  }
  int64_t len = max_span / L + 2;
  int64_t opt = std::max(len, kernel_size);
  opt = 16 * ((opt + 15) / 16);
  return std::min<int64_t>(opt, std::max<int64_t>(raw->dev.precursor_cycle_max_index, 1));
}


// ---- 4-D (timsTOF) ---------------------------------------------------------------------------------
// processing order by time: a tof-major raw file has no per-candidate locality except along the push axis, so
// co-resident CTAs should work on the same time slice of every tof row (the slice then stays in L2)
__global__ void order_key4_kernel(DevLib lib, uint64_t* keys, int32_t* vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lib.n_precursors) return;
  uint32_t rb = __float_as_uint(lib.rt[i]);
  rb = (rb & 0x80000000u) ? ~rb : (rb | 0x80000000u);
  keys[i] = (uint64_t)rb;
  vals[i] = (int32_t)i;
}

__global__ void score_order_key4_kernel(DevCandidatesIn cand, uint64_t* keys, int32_t* vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cand.n) return;
  int64_t fs = cand.frame_start[i];
  keys[i] = fs < 0 ? 0ull : (uint64_t)fs;
  vals[i] = (int32_t)i;
}

// max scan extent and frame extent of the candidates -> out[0], out[1]; out[2] != 0: a library row is out of range
__global__ void cand_extent_kernel(DevCandidatesIn cand, int64_t n_precursors, unsigned long long* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long ms = 0ull, mf = 0ull, bad = 0ull;
  if (i < cand.n) {
    int64_t s = cand.scan_stop[i] - cand.scan_start[i], f = cand.frame_stop[i] - cand.frame_start[i];
    if (cand.frame_start[i] < 0 || cand.frame_stop[i] < 0) f = 0;
    ms = s > 0 ? (unsigned long long)s : 0ull;
    mf = f > 0 ? (unsigned long long)f : 0ull;
    bad = (cand.lib_row[i] < 0 || cand.lib_row[i] >= n_precursors) ? 1ull : 0ull;
  }
  for (int off = 16; off > 0; off >>= 1) {
    ms = max(ms, __shfl_xor_sync(0xffffffffu, ms, off));
    mf = max(mf, __shfl_xor_sync(0xffffffffu, mf, off));
    bad |= __shfl_xor_sync(0xffffffffu, bad, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (ms) atomicMax(out, ms);
    if (mf) atomicMax(out + 1, mf);
    if (bad) atomicOr(out + 2, 1ull);
  }
}

// upper bounds of the selection window of any precursor: cycles (jitclasses/utils.py:62-70) and scans (bruker_jit.py:227-233)
void window_upper_bounds_4d(const adb_rawfile* raw, double rt_tol, double mob_tol, int64_t kernel_size, int64_t* c_cap, int64_t* s_cap) {
  const std::vector<double>& rt = raw->rt4_host;
  const int64_t n = (int64_t)rt.size(), L = raw->dev4.Fr;
  int64_t max_span = 0, j = 0;
  for (int64_t i = 0; i < n; i++) {
    if (j < i) j = i;
    while (j < n && rt[j] <= rt[i] + 2.0 * rt_tol + 1e-3) j++;
    max_span = std::max(max_span, j - i);
  }
  int64_t opt = std::max(max_span / L + 2, kernel_size);
  opt = 16 * ((opt + 15) / 16);
  *c_cap = std::min<int64_t>(opt, std::max<int64_t>(raw->dev4.precursor_cycle_max_index, 1));
  const std::vector<double>& mob = raw->mob4_host;  // descending
  const int64_t m = (int64_t)mob.size();
  int64_t max_sc = 0;
  j = 0;
  for (int64_t i = 0; i < m; i++) {  // scans whose mobility lies within 2 * tol below mob[i]
    if (j < i) j = i;
    while (j < m && mob[j] >= mob[i] - 2.0 * mob_tol - 1e-6) j++;
    max_sc = std::max(max_sc, j - i);
  }
  int64_t so = 16 * ((max_sc + 2 + 15) / 16);
  *s_cap = std::min<int64_t>(so, std::max<int64_t>(raw->dev4.scan_max_index, 1));
}

int run_selection4d(adb_rawfile* raw, adb_library* lib, const adb_selection_config* cfg, const float* kernel, int32_t kh, int32_t kw) {
  if (kh < 1 || kh > ADB_MAX_KERNEL_W || kw < 1 || kw > ADB_MAX_KERNEL_W)
    return fail("kernel height and width must be in [1, " + std::to_string(ADB_MAX_KERNEL_W) + "]");
  cudaStream_t st = raw->stream;
  const int64_t P = lib->dev.n_precursors;
  const int64_t rows = P * cfg->candidate_count;
  CUDA_TRY(cudaEventRecord(raw->ev[0], st));
  std::vector<double> kd((size_t)kh * kw);
  for (size_t t = 0; t < kd.size(); t++) kd[t] = (double)kernel[t];
  if (raw->kern.reserve(sizeof(double) * kd.size())) return 1;
  CUDA_TRY(cudaMemcpyAsync(raw->kern.ptr, kd.data(), sizeof(double) * kd.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // kd is a local
  if (raw->cont.reserve(container_bytes(rows))) return 1;
  raw->d_cont = carve_container(raw->cont.ptr, rows);
  raw->cont_rows = rows;
  raw->cont_count = cfg->candidate_count;
  CUDA_TRY(cudaEventRecord(raw->ev[1], st));
  CUDA_TRY(cudaMemsetAsync(raw->cont.ptr, 0, container_bytes(rows), st));
  int32_t* d_order = nullptr;
  if (P > 1) {
    if (raw->order_keys.reserve(sizeof(uint64_t) * 2 * (size_t)P)) return 1;
    if (raw->order_vals.reserve(sizeof(int32_t) * 2 * (size_t)P)) return 1;
    uint64_t* k_in = raw->order_keys.as<uint64_t>();
    uint64_t* k_out = k_in + P;
    int32_t* v_in = raw->order_vals.as<int32_t>();
    int32_t* v_out = v_in + P;
    order_key4_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(lib->dev, k_in, v_in);
    raw->launches++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)P, 0, 32, st);
    if (raw->order_tmp.reserve(tmp)) return 1;
    cub::DeviceRadixSort::SortPairs(raw->order_tmp.ptr, tmp, k_in, k_out, v_in, v_out, (int)P, 0, 32, st);
    raw->launches += 4;
    d_order = v_out;
  }
  int64_t c_cap = 0, s_cap = 0;
  window_upper_bounds_4d(raw, cfg->rt_tolerance, cfg->mobility_tolerance, cfg->kernel_size, &c_cap, &s_cap);
  if (c_cap > 4096 || s_cap > 4096 || c_cap * s_cap > (1 << 22)) return fail("selection window too large for the device workspace");
  const int nI = (int)std::min<int64_t>(std::min<int64_t>(lib->dev.n_isotopes, cfg->top_k_precursors), ADB_MAX_ISOTOPES);
  Select4Geometry g{(int)s_cap, (int)c_cap, std::min(lib->max_lib_fragments, (int)ADB_MAX_LIB_FRAGMENTS) + nI};
  size_t dyn = 0;
  int tile_in_smem = 0;
  const int grid = adb_select4d_grid(raw->device, raw->dev4, g, kh, kw, &dyn, &tile_in_smem);
  if (grid <= 0) return fail("selection kernel does not fit the shared memory of this device");
  const size_t per_cta = adb_select4d_ws_bytes_per_cta(g);
  if (raw->sel_ws.reserve(per_cta * (size_t)grid + 4096)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[4], st));
  adb_launch_select4d(raw->dev4, lib->dev, *cfg, raw->kern.as<double>(), kh, kw, raw->d_cont, P, d_order, raw->d_status, g,
                      raw->sel_ws.ptr, per_cta, grid, dyn, tile_in_smem, st, &raw->launches);
  CUDA_TRY(cudaEventRecord(raw->ev[5], st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int run_scoring4d(adb_rawfile* raw, adb_library* lib, const adb_scoring_config* cfg, int64_t s_max, int64_t f_max) {
  cudaStream_t st = raw->stream;
  const int64_t n = raw->d_cand.n;
  const int K = (int)cfg->top_k_fragments;
  if (raw->scores.reserve(scores_bytes(std::max<int64_t>(n, 1), K))) return 1;
  raw->d_scores = carve_scores(raw->scores.ptr, std::max<int64_t>(n, 1), K);
  raw->scores_n = n;
  raw->scores_k = K;
  CUDA_TRY(cudaMemsetAsync(raw->scores.ptr, 0, scores_bytes(std::max<int64_t>(n, 1), K), st));
  const int64_t L = raw->dev4.Fr;
  const int64_t c_max = std::max<int64_t>(f_max / L + 2, 4);
  s_max = std::max<int64_t>(std::min<int64_t>(s_max, raw->dev4.Sc), 1);
  if (c_max > 4096) return fail("a candidate spans more than 4096 cycles");
  const int nI = (int)std::min<int64_t>(std::min<int64_t>(lib->dev.n_isotopes, cfg->top_k_isotopes), ADB_MAX_ISOTOPES);
  const int nobs_cap = (int)std::min<int64_t>(L, ADB_MAX_OBS4);
  const int64_t ws_floats = adb_score4d_workspace_floats(K, nI, s_max, c_max, nobs_cap);
  int warps = adb_score4d_resident_warps(raw->device);
  if (raw->ws_budget == 0) {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    raw->ws_budget = std::min<size_t>((size_t)8 << 30, free_b / 3);
  }
  const size_t budget = std::max(raw->ws_budget, raw->score_ws.bytes);
  const size_t per_warp = sizeof(float) * (size_t)ws_floats;
  if (per_warp * 4 > budget) return fail("a candidate window exceeds the device scratch");
  warps = (int)std::min<size_t>((size_t)warps, budget / per_warp);
  warps = std::max(4, warps - warps % 4);
  if (raw->score_ws.reserve(per_warp * (size_t)warps)) return 1;
  int32_t* d_order = nullptr;
  if (n > 1 && n < 2000000000LL) {
    if (raw->order_keys.reserve(sizeof(uint64_t) * 2 * (size_t)n)) return 1;
    if (raw->order_vals.reserve(sizeof(int32_t) * 2 * (size_t)n)) return 1;
    uint64_t* k_in = raw->order_keys.as<uint64_t>();
    uint64_t* k_out = k_in + n;
    int32_t* v_in = raw->order_vals.as<int32_t>();
    int32_t* v_out = v_in + n;
    score_order_key4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw->d_cand, k_in, v_in);
    raw->launches++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)n, 0, 32, st);
    if (raw->order_tmp.reserve(tmp)) return 1;
    cub::DeviceRadixSort::SortPairs(raw->order_tmp.ptr, tmp, k_in, k_out, v_in, v_out, (int)n, 0, 32, st);
    raw->launches += 4;
    d_order = v_out;
  }
  CUDA_TRY(cudaEventRecord(raw->ev[4], st));
  adb_launch_score4d(raw->dev4, lib->dev, *cfg, raw->d_cand, raw->d_scores, raw->score_ws.as<float>(), ws_floats, warps,
                     (int)s_max, (int)c_max, d_order, raw->d_status, st, &raw->launches);
  CUDA_TRY(cudaEventRecord(raw->ev[5], st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// extents of the device-resident candidate table (scratch sizing) + library-row validation
int resident_extents(adb_rawfile* raw, int64_t* s_max, int64_t* f_max, int64_t n_precursors = (int64_t)1 << 62, int64_t* bad_row = nullptr) {
  cudaStream_t st = raw->stream;
  if (raw->extent.reserve(4 * sizeof(unsigned long long))) return 1;
  CUDA_TRY(cudaMemsetAsync(raw->extent.ptr, 0, 4 * sizeof(unsigned long long), st));
  const int64_t n = raw->d_cand.n;
  if (n > 0) {
    cand_extent_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw->d_cand, n_precursors, raw->extent.as<unsigned long long>());
    raw->launches++;
  }
  unsigned long long h[3] = {0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(h, raw->extent.ptr, sizeof(h), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *s_max = (int64_t)h[0];
  *f_max = (int64_t)h[1];
  if (bad_row) *bad_row = (int64_t)h[2];
  return 0;
}

int run_selection(adb_rawfile* raw, adb_library* lib, const adb_selection_config* cfg, const float* kernel,
                  int32_t kh, int32_t kw) {
  if (raw->device != lib->device) return fail("raw file and library live on different devices");
  if (lib->dev.n_precursors * std::max<int64_t>(cfg->candidate_count, 1) >= 2000000000LL)
    return fail("library batch too large: n_precursors * candidate_count must stay below 2e9 (split the library)");
  if (raw->is4d) {
    if (cfg->candidate_count < 1 || cfg->candidate_count > 16) return fail("candidate_count must be in [1, 16]");
    if (cfg->top_k_precursors < 1) return fail("top_k_precursors must be >= 1");
    if (set_device(raw->device)) return 1;
    return run_selection4d(raw, lib, cfg, kernel, kh, kw);
  }
  if (kh != 2) return fail("3-D selection expects a kernel of height 2 (GaussianKernel with scan_max_index + 1 == 2)");
  if (kw < 1 || kw > ADB_MAX_KERNEL_W) return fail("kernel width must be in [1, " + std::to_string(ADB_MAX_KERNEL_W) + "]");
  if (cfg->candidate_count < 1 || cfg->candidate_count > 16) return fail("candidate_count must be in [1, 16]");
  if (cfg->top_k_precursors < 1) return fail("top_k_precursors must be >= 1");
  if (set_device(raw->device)) return 1;
  cudaStream_t st = raw->stream;
  const int64_t P = lib->dev.n_precursors;
  const int64_t rows = P * cfg->candidate_count;

  CUDA_TRY(cudaEventRecord(raw->ev[0], st));
  // kernel as doubles; it travels in the kernel parameters (constant bank)
  std::vector<double> kd((size_t)kh * kw);
  for (size_t t = 0; t < kd.size(); t++) kd[t] = (double)kernel[t];
  if (raw->cont.reserve(container_bytes(rows))) return 1;
  raw->d_cont = carve_container(raw->cont.ptr, rows);
  raw->cont_rows = rows;
  raw->cont_count = cfg->candidate_count;
  CUDA_TRY(cudaEventRecord(raw->ev[1], st));
  // CandidateContainer.__init__ zero-fills (config_df.py:241-254)
  CUDA_TRY(cudaMemsetAsync(raw->cont.ptr, 0, container_bytes(rows), st));

  // processing order: (quad window of the precursor, library RT) so that concurrently resident CTAs read
  // the same spectra (L2 reuse); results do not depend on it (disjoint output rows)
  int32_t* d_order = nullptr;
  if (P > 1) {
    if (raw->order_keys.reserve(sizeof(uint64_t) * 2 * (size_t)P)) return 1;
    if (raw->order_vals.reserve(sizeof(int32_t) * 2 * (size_t)P)) return 1;
    uint64_t* k_in = raw->order_keys.as<uint64_t>();
    uint64_t* k_out = k_in + P;
    int32_t* v_in = raw->order_vals.as<int32_t>();
    int32_t* v_out = v_in + P;
    order_key_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(raw->dev, lib->dev, k_in, v_in);
    raw->launches++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)P, 0, 64, st);
    if (raw->order_tmp.reserve(tmp)) return 1;
    cub::DeviceRadixSort::SortPairs(raw->order_tmp.ptr, tmp, k_in, k_out, v_in, v_out, (int)P, 0, 64, st);
    raw->launches += 4;
    d_order = v_out;
  }

  // geometry: dense XIC buffer strides + chunking of the precursor list so the HBM workspace stays bounded
  const int64_t c_upper = cycle_window_upper_bound(raw, cfg->rt_tolerance, cfg->kernel_size);
  const int nI = (int)std::min<int64_t>(std::min<int64_t>(lib->dev.n_isotopes, cfg->top_k_precursors), ADB_MAX_ISOTOPES);
  const int max_layers = std::min(lib->max_lib_fragments, (int)ADB_MAX_LIB_FRAGMENTS) + nI;
  const int c_cap = (int)c_upper;
  if (c_cap > 4096) return fail("selection cycle window larger than 4096 cycles is not supported");
  const size_t per_prec = adb_select_bytes_per_precursor(c_cap, max_layers, kw);
  if (raw->ws_budget == 0) {  // once per handle: cudaMemGetInfo is slow
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    raw->ws_budget = std::min<size_t>((size_t)8 << 30, free_b / 3);
  }
  size_t budget = std::max(raw->ws_budget, raw->sel_ws.bytes);
  int64_t chunk = (int64_t)std::max<size_t>(budget / per_prec, 1);
  chunk = std::min<int64_t>(chunk, P);
  if (per_prec > budget) return fail("selection window too large for the device workspace");
  if (raw->sel_ws.reserve((size_t)chunk * per_prec + 4096)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[4], st));
  for (int64_t begin = 0; begin < P; begin += chunk)
    adb_launch_select_chunk(raw->dev, lib->dev, *cfg, kd.data(), kw, raw->d_cont, begin, std::min<int64_t>(chunk, P - begin),
                            d_order, raw->d_status, c_cap, max_layers, raw->sel_ws.ptr, raw->sm_count, st, &raw->launches);
  CUDA_TRY(cudaEventRecord(raw->ev[5], st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int run_compaction(adb_rawfile* raw, float score_cutoff = -INFINITY) {
  cudaStream_t st = raw->stream;
  const int64_t rows = raw->cont_rows;
  if (raw->flags.reserve(sizeof(int) * (size_t)std::max<int64_t>(rows, 1))) return 1;
  if (raw->offs.reserve(sizeof(int) * (size_t)std::max<int64_t>(rows, 1))) return 1;
  size_t tmp = adb_compact_temp_bytes(rows);
  if (raw->scan_tmp.reserve(tmp + 16)) return 1;
  if (raw->count.reserve(sizeof(int64_t))) return 1;
  if (raw->cand_in.reserve(cand_in_bytes(rows))) return 1;
  CandInPtrs c = carve_cand_in(raw->cand_in.ptr, rows);
  CUDA_TRY(cudaMemsetAsync(raw->count.ptr, 0, sizeof(int64_t), st));
  adb_launch_compact_ex(raw->d_cont, raw->cont_count, raw->flags.as<int>(), raw->offs.as<int>(), raw->scan_tmp.ptr, tmp,
                        c.lib_row, c.rank, c.scan_start, c.scan_stop, c.scan_center, c.frame_start, c.frame_stop,
                        c.frame_center, c.precursor_idx, c.score, raw->count.as<int64_t>(), st, &raw->launches, score_cutoff);
  raw->d_cand_pidx = c.precursor_idx;
  raw->d_cand_score = c.score;
  CUDA_TRY(cudaGetLastError());
  int64_t n = 0;
  CUDA_TRY(cudaMemcpyAsync(&n, raw->count.ptr, sizeof(n), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  raw->n_cand = n;
  raw->d_cand = DevCandidatesIn{n, c.lib_row, c.rank, c.scan_start, c.scan_stop, c.scan_center, c.frame_start, c.frame_stop, c.frame_center};
  return 0;
}

// grows the batch workspace of the data-parallel scoring passes (called between two batches, stream idle)
int grow_dp_cube(void* owner, size_t floats) {
  adb_rawfile* raw = (adb_rawfile*)owner;
  if (raw->score_ws.reserve(sizeof(float) * floats)) return 1;
  raw->dp_cube = raw->score_ws.as<float>();
  raw->dp_cube_floats = raw->score_ws.bytes / sizeof(float);
  return 0;
}

bool use_tile_scoring() {  // ADB_SCORE_TILE=1: the r1 tile-per-candidate kernel (A/B measurements only)
  const char* e = getenv("ADB_SCORE_TILE");
  return e && e[0] == '1';
}

// ---- ragged results: device-side compaction to what collect_candidates / collect_fragments keep ------------------------
// (scoring/output.py:72-97): rows with valid != 0 and, of those, the fragment slots with mz_library > 0
struct RaggedDev {
  int64_t* row_index;   // [n]
  int64_t* frag_offset; // [n + 1]
  float* features;      // [n, 46]
  float* f32[7];        // [n * K] each, order of adb_scores_ragged
  uint8_t* u8[5];
};

size_t ragged_rows_bytes(int64_t n) { return (size_t)n * (8 + 4 * ADB_NUM_FEATURES) + ((size_t)n + 1) * 8 + 3 * 256; }
size_t ragged_frags_bytes(int64_t n, int k) { return ((size_t)n * (size_t)k * 4 + 256) * 7 + ((size_t)n * (size_t)k + 256) * 5; }

RaggedDev carve_ragged(void* rows, void* frags, int64_t n, int k) {
  RaggedDev r{};
  const size_t N = (size_t)n, NK = (size_t)n * (size_t)k;
  char* p = (char*)rows;
  auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) & ~(size_t)255; return q; };
  r.row_index = (int64_t*)take(8 * N);
  r.frag_offset = (int64_t*)take(8 * (N + 1));
  r.features = (float*)take(4 * N * ADB_NUM_FEATURES);
  p = (char*)frags;
  for (int i = 0; i < 7; i++) r.f32[i] = (float*)take(4 * NK);
  for (int i = 0; i < 5; i++) r.u8[i] = (uint8_t*)take(NK);
  return r;
}

// per row of the block: (valid, kept fragment slots) packed as (1 << 32) | n_slots
__global__ void ragged_count_kernel(DevScoresOut s, int k, int64_t r0, int64_t r1, unsigned long long* packed) {
  const int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  unsigned long long v = 0;
  if (s.valid[i]) {
    unsigned c = 0;
    const float* m = s.fragment_mz_library + (size_t)i * (size_t)k;
    for (int q = 0; q < k; q++) c += m[q] > 0.f;
    v = (1ull << 32) | c;
  }
  packed[i - r0] = v;
}

// one warp per row: position = running totals of the earlier blocks + exclusive prefix inside the block
__global__ void ragged_scatter_kernel(DevScoresOut s, int k, int64_t r0, int64_t r1, const unsigned long long* packed,
                                      const unsigned long long* prefix, const int64_t* tot, RaggedDev out) {
  const int64_t i = r0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= r1) return;
  if (!(packed[i - r0] >> 32)) return;
  const unsigned long long pre = prefix[i - r0];
  const int64_t p = tot[0] + (int64_t)(pre >> 32);
  int64_t q = tot[1] + (int64_t)(pre & 0xFFFFFFFFull);
  if (lane == 0) { out.row_index[p] = i; out.frag_offset[p] = q; }
  for (int t = lane; t < ADB_NUM_FEATURES; t += 32) out.features[(size_t)p * ADB_NUM_FEATURES + t] = s.features[(size_t)i * ADB_NUM_FEATURES + t];
  const float* const src_f[7] = {s.fragment_mz_library, s.fragment_mz, s.fragment_mz_observed, s.fragment_height,
                                 s.fragment_intensity, s.fragment_mass_error, s.fragment_correlation};
  const uint8_t* const src_u[5] = {s.fragment_position, s.fragment_number, s.fragment_type, s.fragment_charge, s.fragment_loss_type};
  for (int base = 0; base < k; base += 32) {
    const int slot = base + lane;
    const size_t a = (size_t)i * (size_t)k + (size_t)slot;
    const bool keep = slot < k && s.fragment_mz_library[a] > 0.f;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    if (keep) {
      const int64_t d = q + __popc(m & ((1u << lane) - 1u));
#pragma unroll
      for (int c = 0; c < 7; c++) out.f32[c][d] = src_f[c][a];
#pragma unroll
      for (int c = 0; c < 5; c++) out.u8[c][d] = src_u[c][a];
    }
    q += __popc(m);
  }
}

// tot += totals of the block; snap[0..1] = tot after the block
__global__ void ragged_advance_kernel(const unsigned long long* packed, const unsigned long long* prefix, int64_t nb, int64_t* tot, int64_t* snap) {
  if (nb > 0) {
    const unsigned long long t = prefix[nb - 1] + packed[nb - 1];
    tot[0] += (int64_t)(t >> 32);
    tot[1] += (int64_t)(t & 0xFFFFFFFFull);
  }
  snap[0] = tot[0];
  snap[1] = tot[1];
}

struct RaggedRun {  // host state of one adb_score_candidates_ragged call
  adb_scores_ragged* out = nullptr;
  RaggedDev dev{};
  int k = 0;
  int n_blocks = 0, copied = 0;
  bool overflow = false;
};

int ragged_begin(adb_rawfile* raw, RaggedRun& R, int64_t n, int k) {
  R.k = k;
  R.n_blocks = 0;
  R.copied = 0;
  const int64_t N = std::max<int64_t>(n, 1);
  if (raw->rag_rows.reserve(ragged_rows_bytes(N)) || raw->rag_frags.reserve(ragged_frags_bytes(N, k)) ||
      raw->rag_scan.reserve(16 * (size_t)N + 512) || raw->rag_tot.reserve(256))
    return 1;
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)N);
  if (raw->rag_scan_tmp.reserve(tmp + 256)) return 1;
  if (!raw->rag_host_tot) {
    CUDA_TRY(cudaHostAlloc((void**)&raw->rag_host_tot, sizeof(int64_t) * 2 * (ADB_MAX_BLOCKS + 2), cudaHostAllocDefault));
    for (int i = 0; i <= ADB_MAX_BLOCKS; i++) CUDA_TRY(cudaEventCreateWithFlags(&raw->rag_ev[i], cudaEventDisableTiming));
  }
  R.dev = carve_ragged(raw->rag_rows.ptr, raw->rag_frags.ptr, N, k);
  CUDA_TRY(cudaMemsetAsync(raw->rag_tot.ptr, 0, 256, raw->stream));
  raw->rag_host_tot[0] = raw->rag_host_tot[1] = 0;
  return 0;
}

// on `st` (after the rows [r0, r1) are scored): compaction of the block + its running totals to the host
int ragged_compact_block(adb_rawfile* raw, RaggedRun& R, int64_t r0, int64_t r1, cudaStream_t st) {
  const int b = R.n_blocks++;
  const int64_t nb = r1 - r0;
  unsigned long long* packed = raw->rag_scan.as<unsigned long long>() + r0;
  unsigned long long* prefix = packed + std::max<int64_t>(raw->scores_n, 1);
  int64_t* tot = raw->rag_tot.as<int64_t>();
  if (nb > 0) {
    ragged_count_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(raw->d_scores, R.k, r0, r1, packed);
    size_t tmp = raw->rag_scan_tmp.bytes;
    cub::DeviceScan::ExclusiveSum(raw->rag_scan_tmp.ptr, tmp, packed, prefix, (int)nb, st);
    ragged_scatter_kernel<<<(unsigned)((nb * 32 + 255) / 256), 256, 0, st>>>(raw->d_scores, R.k, r0, r1, packed, prefix, tot, R.dev);
    raw->launches += 4;
  }
  ragged_advance_kernel<<<1, 1, 0, st>>>(packed, prefix, nb, tot, tot + 2 + 2 * b);
  raw->launches++;
  CUDA_TRY(cudaMemcpyAsync(raw->rag_host_tot + 2 * (b + 1), tot + 2 + 2 * b, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(raw->rag_ev[b], st));
  return 0;
}

// D2H of the compacted blocks [R.copied, upto) whose totals have arrived (blocks wait for their event)
int ragged_copy_blocks(adb_rawfile* raw, RaggedRun& R, int upto, cudaStream_t st) {
  adb_scores_ragged* o = R.out;
  for (; R.copied < upto; R.copied++) {
    const int b = R.copied;
    CUDA_TRY(cudaEventSynchronize(raw->rag_ev[b]));
    const int64_t p0 = raw->rag_host_tot[2 * b], p1 = raw->rag_host_tot[2 * b + 2];
    const int64_t q0 = raw->rag_host_tot[2 * b + 1], q1 = raw->rag_host_tot[2 * b + 3];
    if (p1 > o->row_capacity || q1 > o->frag_capacity) { R.overflow = true; continue; }
    if (p1 > p0) {
      const size_t a = (size_t)p0, N = (size_t)(p1 - p0);
      CUDA_TRY(cudaMemcpyAsync(o->row_index + a, R.dev.row_index + a, 8 * N, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(o->frag_offset + a, R.dev.frag_offset + a, 8 * N, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(o->features + a * ADB_NUM_FEATURES, R.dev.features + a * ADB_NUM_FEATURES, 4 * N * ADB_NUM_FEATURES, cudaMemcpyDeviceToHost, st));
    }
    if (q1 > q0) {
      const size_t a = (size_t)q0, N = (size_t)(q1 - q0);
      float* hf[7] = {o->fragment_mz_library, o->fragment_mz, o->fragment_mz_observed, o->fragment_height,
                      o->fragment_intensity, o->fragment_mass_error, o->fragment_correlation};
      uint8_t* hu[5] = {o->fragment_position, o->fragment_number, o->fragment_type, o->fragment_charge, o->fragment_loss_type};
      for (int c = 0; c < 7; c++) CUDA_TRY(cudaMemcpyAsync(hf[c] + a, R.dev.f32[c] + a, 4 * N, cudaMemcpyDeviceToHost, st));
      for (int c = 0; c < 5; c++) CUDA_TRY(cudaMemcpyAsync(hu[c] + a, R.dev.u8[c] + a, N, cudaMemcpyDeviceToHost, st));
    }
  }
  return 0;
}

// D2H of rows [r0, r1) of the resident score tables into the caller's (row-major) host tables
int copy_score_rows(adb_rawfile* raw, adb_scores_out* out, int64_t r0, int64_t r1, cudaStream_t st) {
  if (r1 <= r0) return 0;
  const size_t K = (size_t)raw->scores_k, a = (size_t)r0, N = (size_t)(r1 - r0);
  const DevScoresOut& s = raw->d_scores;
  CUDA_TRY(cudaMemcpyAsync(out->features + a * ADB_NUM_FEATURES, s.features + a * ADB_NUM_FEATURES, 4 * N * ADB_NUM_FEATURES, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->valid + a, s.valid + a, N, cudaMemcpyDeviceToHost, st));
  float* hf[] = {out->fragment_mz_library, out->fragment_mz, out->fragment_mz_observed, out->fragment_height,
                 out->fragment_intensity, out->fragment_mass_error, out->fragment_correlation};
  float* df[] = {s.fragment_mz_library, s.fragment_mz, s.fragment_mz_observed, s.fragment_height,
                 s.fragment_intensity, s.fragment_mass_error, s.fragment_correlation};
  for (int i = 0; i < 7; i++) CUDA_TRY(cudaMemcpyAsync(hf[i] + a * K, df[i] + a * K, 4 * N * K, cudaMemcpyDeviceToHost, st));
  uint8_t* hu[] = {out->fragment_position, out->fragment_number, out->fragment_type, out->fragment_charge, out->fragment_loss_type};
  uint8_t* du[] = {s.fragment_position, s.fragment_number, s.fragment_type, s.fragment_charge, s.fragment_loss_type};
  for (int i = 0; i < 5; i++) CUDA_TRY(cudaMemcpyAsync(hu[i] + a * K, du[i] + a * K, N * K, cudaMemcpyDeviceToHost, st));
  return 0;
}

// host_out != nullptr: the candidates are scored in up to 4 row blocks and every finished block is copied to the host on a
// second stream while the next block is being scored
int run_scoring(adb_rawfile* raw, adb_library* lib, const adb_scoring_config* cfg_in, int64_t c_max_hint, int64_t s_max_hint = 0,
                adb_scores_out* host_out = nullptr, RaggedRun* rag = nullptr) {
  if (raw->device != lib->device) return fail("raw file and library live on different devices");
  if (cfg_in->top_k_fragments < 1) return fail("top_k_fragments must be >= 1");
  if (cfg_in->top_k_isotopes < 1) return fail("top_k_isotopes must be >= 1");
  adb_scoring_config cfg_eff = *cfg_in;
  const adb_scoring_config* cfg = &cfg_eff;
  const bool tile_path = use_tile_scoring();
  if (rag) {
    // a candidate keeps at most the fragments its precursor has: the device tables are as wide as the widest precursor
    cfg_eff.top_k_fragments = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(cfg_in->top_k_fragments, lib->max_lib_fragments));
    const int cap = (raw->is4d || tile_path) ? ADB_MAX_FRAGMENTS : ADB_MAX_LIB_FRAGMENTS;
    if ((int)cfg_eff.top_k_fragments > cap)
      return fail("the library's widest precursor has more than " + std::to_string(cap) + " fragments: not supported on this path");
  } else if (cfg->top_k_fragments > ADB_MAX_FRAGMENTS) {
    return fail("top_k_fragments must be in [1, " + std::to_string(ADB_MAX_FRAGMENTS) + "] for dense score tables (use adb_score_candidates_ragged)");
  }
  if (raw->is4d) {
    if (run_scoring4d(raw, lib, cfg, s_max_hint, c_max_hint /* frames */)) return 1;
    if (host_out) {
      CUDA_TRY(cudaEventRecord(raw->ev[2], raw->stream));
      if (copy_score_rows(raw, host_out, 0, raw->scores_n, raw->stream)) return 1;
    }
    if (rag) {
      CUDA_TRY(cudaEventRecord(raw->ev[2], raw->stream));
      if (ragged_begin(raw, *rag, raw->scores_n, raw->scores_k)) return 1;
      if (ragged_compact_block(raw, *rag, 0, raw->scores_n, raw->stream)) return 1;
      if (ragged_copy_blocks(raw, *rag, rag->n_blocks, raw->stream)) return 1;
    }
    return 0;
  }
  cudaStream_t st = raw->stream;
  const int64_t n = raw->d_cand.n;
  const int K = (int)cfg->top_k_fragments;
  if (raw->scores.reserve(scores_bytes(std::max<int64_t>(n, 1), K))) return 1;
  raw->d_scores = carve_scores(raw->scores.ptr, std::max<int64_t>(n, 1), K);
  raw->scores_n = n;
  raw->scores_k = K;
  // OutputPsmDF.__init__ zero-fills (scoring/output.py:42-70)
  CUDA_TRY(cudaMemsetAsync(raw->scores.ptr, 0, scores_bytes(std::max<int64_t>(n, 1), K), st));
  if (rag && ragged_begin(raw, *rag, n, K)) return 1;
  int tiles = adb_score_resident_tiles(raw->device, K);
  int64_t c_max = std::max<int64_t>(c_max_hint, 32);
  int64_t ws_floats = (adb_score_workspace_floats(K, c_max) + 3) & ~(int64_t)3;
  const int KS = std::max(1, std::min(K, std::min(lib->max_lib_fragments, (int)ADB_MAX_LIB_FRAGMENTS)));
  const int nIcap = (int)std::min<int64_t>(std::min<int64_t>(lib->dev.n_isotopes, cfg->top_k_isotopes), ADB_MAX_ISOTOPES);
  int64_t dp_batch_cap = ADB_SCORE_DP_BATCH;
  if (const char* e = getenv("ADB_DP_BATCH")) dp_batch_cap = std::max<int64_t>(std::min<int64_t>(atoll(e), ADB_SCORE_DP_BATCH_MAX), 256);  // tuning
  const int64_t dp_batch = std::max<int64_t>(std::min<int64_t>(n, dp_batch_cap), 1);
  if (tile_path) {
    // HBM fallback scratch for candidates whose cube exceeds the shared-memory budget
    if (raw->score_ws.reserve(sizeof(float) * (size_t)ws_floats * (size_t)tiles)) return 1;
  } else {
    if (raw->dp_plan.reserve(adb_score_dp_plan_bytes(dp_batch, KS, nIcap, nullptr))) return 1;
    raw->dp_cube = raw->score_ws.as<float>();
    raw->dp_cube_floats = raw->score_ws.bytes / sizeof(float);
  }
  auto launch = [&](DevCandidatesIn part, const int32_t* order) -> int {
    if (tile_path) {
      adb_launch_score(raw->dev, lib->dev, *cfg, part, raw->d_scores, raw->score_ws.as<float>(), ws_floats, tiles, order,
                       raw->d_status, st, &raw->launches);
      return 0;
    }
    if (adb_launch_score_dp(raw->dev, lib->dev, *cfg, part, raw->d_scores, K, KS, order, dp_batch, raw->dp_plan.ptr, &raw->dp_cube,
                            &raw->dp_cube_floats, grow_dp_cube, raw, raw->d_status, st, &raw->launches))
      return g_error.empty() ? fail("data-parallel scoring failed") : 1;
    return 0;
  };
  // processing order: (quad window of the precursor, frame_start) so that co-resident tiles read the same
  // spectra; results do not depend on it (disjoint output rows)
  int32_t* d_order = nullptr;
  int n_chunks = ((host_out || rag) && n >= 200000) ? (rag ? ADB_RAGGED_BLOCKS : ADB_SCORE_BLOCKS) : 1;
  double block_ratio = ADB_RAGGED_BLOCK_RATIO;
  if (rag && n_chunks > 1) {  // tuning knobs
    if (const char* e = getenv("ADB_RAGGED_BLOCKS_N")) n_chunks = std::max(1, std::min(atoi(e), (int)ADB_MAX_BLOCKS));
    if (const char* e = getenv("ADB_RAGGED_RATIO")) block_ratio = std::max(0.1, std::min(atof(e), 2.0));
  }
  // row blocks: equal for the dense tables; geometrically shrinking for the ragged result, whose copy starts one block
  // late - every copy hides behind the next block and the un-overlapped tail is the copy of the last, smallest blocks.
  // Measured on config 3 (fused call, ms per e2e step): 3 blocks / ratio 0.6: 160.8, 4 / 0.7: 158.2, 5 / 0.7: 157.4, 6 / 0.7: 157.2,
  // 6 / 0.8: 159.7, 7 / 0.6: 162.6, 8 / 0.9: 162.1, 12 / 0.95: 165.5, 15 / 1.0: 168.5 - every block costs about a millisecond
  RowBlocks blocks{};
  blocks.n = n_chunks;
  {
    double w[ADB_MAX_BLOCKS + 1], tot = 0;
    for (int k = 0; k < n_chunks; k++) { w[k] = rag ? pow(block_ratio, k) : 1.0; tot += w[k]; }
    double acc = 0;
    for (int k = 0; k < n_chunks; k++) { blocks.start[k] = std::min<int64_t>((int64_t)(acc / tot * (double)n), n); acc += w[k]; }
    blocks.start[0] = 0;
    blocks.start[n_chunks] = n;
    for (int k = n_chunks + 1; k < ADB_MAX_BLOCKS + 2; k++) blocks.start[k] = n;
  }
  if (n > 1 && n < 2000000000LL) {
    if (raw->order_keys.reserve(sizeof(uint64_t) * 2 * (size_t)n)) return 1;
    if (raw->order_vals.reserve(sizeof(int32_t) * 2 * (size_t)n)) return 1;
    uint64_t* k_in = raw->order_keys.as<uint64_t>();
    uint64_t* k_out = k_in + n;
    int32_t* v_in = raw->order_vals.as<int32_t>();
    int32_t* v_out = v_in + n;
    score_order_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw->dev, lib->dev, raw->d_cand, blocks,
                                                                      (int)std::min<int64_t>(std::min<int64_t>(lib->dev.n_isotopes, cfg->top_k_isotopes), ADB_MAX_ISOTOPES), k_in, v_in);
    raw->launches++;
    const int end_bit = n_chunks > 1 ? 52 : 48;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)n, 0, end_bit, st);
    if (raw->order_tmp.reserve(tmp)) return 1;
    cub::DeviceRadixSort::SortPairs(raw->order_tmp.ptr, tmp, k_in, k_out, v_in, v_out, (int)n, 0, end_bit, st);
    raw->launches += 4;
    d_order = v_out;
  }
  CUDA_TRY(cudaEventRecord(raw->ev[4], st));
  if (n_chunks == 1 || d_order == nullptr) {
    if (launch(raw->d_cand, d_order)) return 1;
    CUDA_TRY(cudaEventRecord(raw->ev[5], st));
    if (host_out) {
      CUDA_TRY(cudaEventRecord(raw->ev[2], st));
      if (copy_score_rows(raw, host_out, 0, n, st)) return 1;
    }
    if (rag) {
      CUDA_TRY(cudaEventRecord(raw->ev[2], st));
      if (ragged_compact_block(raw, *rag, 0, n, st)) return 1;
      if (ragged_copy_blocks(raw, *rag, rag->n_blocks, st)) return 1;
    }
  } else {
    for (int k = 0; k < n_chunks; k++) {
      const int64_t c0 = blocks.start[k], c1 = blocks.start[k + 1];
      if (c1 <= c0) continue;
      DevCandidatesIn part = raw->d_cand;
      part.n = c1 - c0;  // the kernel visits order[c0 .. c1): the (window, time)-sorted rows of block k
      if (launch(part, d_order + c0)) return 1;
      CUDA_TRY(cudaEventRecord(raw->chunk_ev[k], st));
      // ragged: a block is compacted on the copy stream; its D2H is issued once its totals are on the host, i.e. after
      // the next block has been launched (launching blocks the host on the batch sizes anyway)
      if (rag && ragged_copy_blocks(raw, *rag, rag->n_blocks, raw->copy_stream)) return 1;
      CUDA_TRY(cudaStreamWaitEvent(raw->copy_stream, raw->chunk_ev[k], 0));
      if (host_out && copy_score_rows(raw, host_out, c0, c1, raw->copy_stream)) return 1;
      if (rag && ragged_compact_block(raw, *rag, c0, c1, raw->copy_stream)) return 1;
    }
    if (rag && ragged_copy_blocks(raw, *rag, rag->n_blocks, raw->copy_stream)) return 1;
    CUDA_TRY(cudaEventRecord(raw->ev[5], st));
    CUDA_TRY(cudaEventRecord(raw->ev[2], st));
    CUDA_TRY(cudaEventRecord(raw->chunk_ev[ADB_MAX_BLOCKS], raw->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, raw->chunk_ev[ADB_MAX_BLOCKS], 0));  // the handle's stream is done when the copies are
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// H2D of a caller's candidate table into the handle (becomes the resident candidate set)
int upload_candidates(adb_rawfile* raw, const adb_candidates_in* cand) {
  cudaStream_t st = raw->stream;
  const int64_t n = cand->n;
  if (raw->cand_in.reserve(cand_in_bytes(std::max<int64_t>(n, 1)))) return 1;
  CandInPtrs c = carve_cand_in(raw->cand_in.ptr, std::max<int64_t>(n, 1));
  if (n > 0) {
    const size_t N = (size_t)n;
    CUDA_TRY(cudaMemcpyAsync(c.lib_row, cand->lib_row, 8 * N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.rank, cand->rank, N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.scan_start, cand->scan_start, 8 * N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.scan_stop, cand->scan_stop, 8 * N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.scan_center, cand->scan_center, 8 * N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.frame_start, cand->frame_start, 8 * N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.frame_stop, cand->frame_stop, 8 * N, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.frame_center, cand->frame_center, 8 * N, cudaMemcpyHostToDevice, st));
  }
  raw->d_cand = DevCandidatesIn{n, c.lib_row, c.rank, c.scan_start, c.scan_stop, c.scan_center, c.frame_start, c.frame_stop, c.frame_center};
  raw->n_cand = n;
  return 0;
}

void finish_timing(adb_rawfile* raw) {
  cudaEventSynchronize(raw->ev[3]);
  cudaEventElapsedTime(&raw->h2d_ms, raw->ev[0], raw->ev[1]);
  cudaEventElapsedTime(&raw->kernel_ms, raw->ev[1], raw->ev[2]);
  cudaEventElapsedTime(&raw->d2h_ms, raw->ev[2], raw->ev[3]);
  cudaEventElapsedTime(&raw->main_kernel_ms, raw->ev[4], raw->ev[5]);
}

}  // namespace

extern "C" {

const char* adb_last_error(void) { return g_error.c_str(); }
const char* adb_version(void) { return "alphadia_b200 0.1 (sm_100a)"; }

int adb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int adb_rawfile3d_create(const adb_rawfile3d_desc* d, int device, adb_rawfile_t** out) {
  if (!d || !out) return fail("null argument");
  if (d->cycle_len < 1 || d->n_spectra < 1) return fail("empty raw file");
  if (d->n_mobility < 1) return fail("mobility_values must not be empty");
  if (d->n_peaks >= 4294967000LL) return fail("raw files with more than 2^32 peaks are not supported");
  if (set_device(device)) return 1;
  adb_rawfile* r = new adb_rawfile();
  r->device = device;
  CUDA_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 6; i++) CUDA_TRY(cudaEventCreate(&r->ev[i]));
  CUDA_TRY(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i <= ADB_MAX_BLOCKS; i++) CUDA_TRY(cudaEventCreateWithFlags(&r->chunk_ev[i], cudaEventDisableTiming));
  cudaDeviceGetAttribute(&r->sm_count, cudaDevAttrMultiProcessorCount, device);
  DevRaw& v = r->dev;
  double* cyc; float *rt, *mob, *mz, *it; int64_t *ps, *pe;
  if (upload(d->cycle, d->cycle_len * 2, &cyc, r->allocs, r->bytes, r->stream) ||
      upload(d->rt_values, d->n_spectra, &rt, r->allocs, r->bytes, r->stream) ||
      upload(d->mobility_values, d->n_mobility, &mob, r->allocs, r->bytes, r->stream) ||
      upload(d->peak_start_idx, d->n_spectra, &ps, r->allocs, r->bytes, r->stream) ||
      upload(d->peak_stop_idx, d->n_spectra, &pe, r->allocs, r->bytes, r->stream) ||
      upload(d->mz_values, d->n_peaks, &mz, r->allocs, r->bytes, r->stream, ADB_MZ_PAD) ||
      upload(d->intensity_values, d->n_peaks, &it, r->allocs, r->bytes, r->stream, ADB_MZ_PAD)) {
    adb_rawfile_destroy(r);
    return 1;
  }
  v.cycle = cyc; v.cycle_len = d->cycle_len; v.rt_values = rt; v.n_spectra = d->n_spectra;
  v.mobility_values = mob; v.n_mobility = d->n_mobility; v.peak_start = ps; v.peak_stop = pe;
  v.mz = mz; v.intensity = it; v.n_peaks = d->n_peaks; v.zeroth_frame = d->zeroth_frame;
  v.precursor_cycle_max_index = d->precursor_cycle_max_index; v.scan_max_index = d->scan_max_index;
  v.frame_max_index = d->frame_max_index;
  // MS1 positions: windows overlapping the query [-1, -1] (alpharaw_jit.py:46-48)
  v.n_ms1_pos = 0;
  for (int64_t j = 0; j < d->cycle_len; j++)
    if ((-1.0 <= d->cycle[2 * j + 1]) && (-1.0 >= d->cycle[2 * j])) {
      if (v.n_ms1_pos >= ADB_MAX_MS1_POS) { adb_rawfile_destroy(r); return fail("more than 8 MS1 spectra per cycle"); }
      v.ms1_pos[v.n_ms1_pos++] = (int32_t)j;
    }
  r->rt_host.assign(d->rt_values, d->rt_values + d->n_spectra);
  for (int64_t i = 0; i < d->n_spectra; i++)
    if (d->peak_stop_idx[i] - d->peak_start_idx[i] > 2000000000LL || d->peak_stop_idx[i] < d->peak_start_idx[i] ||
        d->peak_stop_idx[i] > d->n_peaks || d->peak_start_idx[i] < 0) {
      adb_rawfile_destroy(r);
      return fail("spectrum " + std::to_string(i) + " has an invalid peak index range");
    }
  {  // derived bucket index, built on the device
    float init[2];
    unsigned int inf_bits = 0x7f800000u, zero_bits = 0u;
    memcpy(&init[0], &inf_bits, 4);
    memcpy(&init[1], &zero_bits, 4);
    float* d_rng = nullptr;
    uint2* d_tab = nullptr;
    if (upload(init, 2, &d_rng, r->allocs, r->bytes, r->stream)) { adb_rawfile_destroy(r); return 1; }
    size_t tab_n = (size_t)d->n_spectra * ADB_N_BUCKETS;
    void* tp = nullptr;
    if (cudaMalloc(&tp, tab_n * sizeof(uint2)) != cudaSuccess) { adb_rawfile_destroy(r); return fail("cudaMalloc bucket index failed"); }
    r->allocs.push_back(tp);
    r->bytes += (int64_t)(tab_n * sizeof(uint2));
    d_tab = (uint2*)tp;
    mz_range_kernel<<<(unsigned)((d->n_spectra + 255) / 256), 256, 0, r->stream>>>(v, d_rng);
    float rng[2] = {0.f, 0.f};
    cudaMemcpyAsync(rng, d_rng, sizeof(rng), cudaMemcpyDeviceToHost, r->stream);
    if (cudaStreamSynchronize(r->stream) != cudaSuccess) { adb_rawfile_destroy(r); return fail("m/z range kernel failed"); }
    if (!(rng[1] > rng[0])) { rng[0] = 0.f; rng[1] = 1.f; }
    v.bucket_lo = rng[0];
    v.bucket_width = (rng[1] - rng[0]) / (float)ADB_N_BUCKETS * 1.0001f;
    if (!(v.bucket_width > 0.f)) v.bucket_width = 1.f;
    v.bucket_inv_width = 1.0f / v.bucket_width;
    v.bucket_pair = d_tab;
    bucket_index_kernel<<<(unsigned)((tab_n + 255) / 256), 256, 0, r->stream>>>(v, d_tab);
    r->launches += 2;
  }
  {  // derived m/z-major index (see DevRaw): stable radix sort of (cycle position, m/z) over all peaks
    const int64_t n = d->n_peaks;
    const size_t N = (size_t)std::max<int64_t>(n, 1);
    float4* s_pk = nullptr; uint32_t* s_tab = nullptr; int64_t* pstart = nullptr;
    int nb = 64;  // about 16 peaks per bucket of an average position
    while (nb < (1 << 18) && (int64_t)nb * 16 * d->cycle_len < n) nb *= 2;
    v.sb_nb = nb;
    v.sb_lo = v.bucket_lo;
    v.sb_width = v.bucket_width * (float)ADB_N_BUCKETS / (float)nb;
    if (!(v.sb_width > 0.f)) v.sb_width = 1.f;
    v.sb_inv_width = 1.0f / v.sb_width;
    const size_t tab_n = (size_t)d->cycle_len * (size_t)(nb + 1);
    auto dalloc = [&](void** p, size_t bytes) { if (cudaMalloc(p, bytes) != cudaSuccess) return 1; r->allocs.push_back(*p); r->bytes += (int64_t)bytes; return 0; };
    if (dalloc((void**)&s_pk, 16 * (N + 64)) || dalloc((void**)&s_tab, 4 * tab_n) ||
        dalloc((void**)&pstart, 8 * (size_t)(d->cycle_len + 1))) { adb_rawfile_destroy(r); return fail("cudaMalloc m/z index failed"); }
    uint64_t *k_in = nullptr, *k_out = nullptr; uint32_t *v_in = nullptr, *v_out = nullptr; void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const int end_bit = 64;  // all-ones filler keys (peaks outside every spectrum) must sort last
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)n, 0, end_bit, r->stream);
    bool ok = cudaMalloc((void**)&k_in, 8 * N) == cudaSuccess && cudaMalloc((void**)&k_out, 8 * N) == cudaSuccess &&
              cudaMalloc((void**)&v_in, 4 * N) == cudaSuccess && cudaMalloc((void**)&v_out, 4 * N) == cudaSuccess &&
              cudaMalloc(&tmp, tmp_bytes + 16) == cudaSuccess;
    if (ok) {
      cudaMemsetAsync(k_in, 0xFF, 8 * N, r->stream);  // peaks outside every spectrum sort to the end
      cudaMemsetAsync(v_in, 0, 4 * N, r->stream);
      mzindex_keys_kernel<<<(unsigned)((d->n_spectra * 32 + 255) / 256), 256, 0, r->stream>>>(v, k_in, v_in);
      cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)n, 0, end_bit, r->stream);
      tbindex_gather_kernel<<<(unsigned)((n + 64 + 255) / 256), 256, 0, r->stream>>>(v, k_out, v_out, n, s_pk, (uint64_t)d->cycle_len, 64);
      mzindex_segments_kernel<<<(unsigned)((d->cycle_len + 1 + 255) / 256), 256, 0, r->stream>>>(k_out, n, d->cycle_len, pstart);
      tbindex_bucket_kernel<<<(unsigned)((tab_n + 255) / 256), 256, 0, r->stream>>>(k_out, n, d->cycle_len, nb, v.sb_lo, v.sb_width, s_tab);
      r->launches += 5;
      ok = cudaStreamSynchronize(r->stream) == cudaSuccess;
    }
    cudaFree(k_in); cudaFree(k_out); cudaFree(v_in); cudaFree(v_out); cudaFree(tmp);
    if (!ok) { cudaGetLastError(); adb_rawfile_destroy(r); return fail("building the m/z-major index failed (out of device memory?)"); }
    v.s_pk = s_pk; v.s_bucket = s_tab; v.pos_start = pstart;
  }
  {  // derived time-blocked m/z index (see DevRaw): stable radix sort of (cycle position, time block, m/z) over all peaks
    const int64_t n = d->n_peaks;
    const size_t N = (size_t)std::max<int64_t>(n, 1);
    const int64_t n_cycles = (d->n_spectra + d->cycle_len - 1) / d->cycle_len;
    const int ntb = (int)std::max<int64_t>((n_cycles + ADB_TB_CYCLES - 1) / ADB_TB_CYCLES, 1);
    const int64_t n_seg = d->cycle_len * ntb;
    int nb = 64;  // about 4 peaks per bucket of an average segment
    while (nb < ADB_TB_MAX_BUCKETS && (int64_t)nb * 4 * n_seg < n) nb *= 2;
    v.tb_ntb = ntb; v.tb_nb = nb;
    v.tb_lo = v.bucket_lo;
    v.tb_width = v.bucket_width * (float)ADB_N_BUCKETS / (float)nb;
    if (!(v.tb_width > 0.f)) v.tb_width = 1.f;
    v.tb_inv_width = 1.0f / v.tb_width;
    float4* t_pk = nullptr; uint32_t* t_tab = nullptr;
    auto dalloc = [&](void** p, size_t bytes) { if (cudaMalloc(p, bytes) != cudaSuccess) return 1; r->allocs.push_back(*p); r->bytes += (int64_t)bytes; return 0; };
    const size_t tab_n = (size_t)n_seg * (size_t)(nb + 1);
    if (dalloc((void**)&t_pk, 16 * (N + 16)) || dalloc((void**)&t_tab, 4 * tab_n)) { adb_rawfile_destroy(r); return fail("cudaMalloc time-blocked m/z index failed"); }
    uint64_t *k_in = nullptr, *k_out = nullptr; uint32_t *v_in = nullptr, *v_out = nullptr; void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)n, 0, 64, r->stream);
    bool ok = cudaMalloc((void**)&k_in, 8 * N) == cudaSuccess && cudaMalloc((void**)&k_out, 8 * N) == cudaSuccess &&
              cudaMalloc((void**)&v_in, 4 * N) == cudaSuccess && cudaMalloc((void**)&v_out, 4 * N) == cudaSuccess &&
              cudaMalloc(&tmp, tmp_bytes + 16) == cudaSuccess;
    if (ok) {
      cudaMemsetAsync(k_in, 0xFF, 8 * N, r->stream);  // peaks outside every spectrum sort to the end
      cudaMemsetAsync(v_in, 0, 4 * N, r->stream);
      tbindex_keys_kernel<<<(unsigned)((d->n_spectra * 32 + 255) / 256), 256, 0, r->stream>>>(v, ntb, k_in, v_in);
      cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)n, 0, 64, r->stream);
      tbindex_gather_kernel<<<(unsigned)((n + 16 + 255) / 256), 256, 0, r->stream>>>(v, k_out, v_out, n, t_pk, (uint64_t)n_seg, 16);
      tbindex_bucket_kernel<<<(unsigned)((tab_n + 255) / 256), 256, 0, r->stream>>>(k_out, n, n_seg, nb, v.tb_lo, v.tb_width, t_tab);
      r->launches += 4;
      ok = cudaStreamSynchronize(r->stream) == cudaSuccess;
    }
    cudaFree(k_in); cudaFree(k_out); cudaFree(v_in); cudaFree(v_out); cudaFree(tmp);
    if (!ok) { cudaGetLastError(); adb_rawfile_destroy(r); return fail("building the time-blocked m/z index failed (out of device memory?)"); }
    v.tb_pk = t_pk; v.tb_bucket = t_tab;
  }
  void* st = nullptr;
  if (cudaMalloc(&st, sizeof(uint32_t)) != cudaSuccess) { adb_rawfile_destroy(r); return fail("cudaMalloc status failed"); }
  r->d_status = (uint32_t*)st;
  cudaMemsetAsync(r->d_status, 0, sizeof(uint32_t), r->stream);
  cudaError_t e = cudaStreamSynchronize(r->stream);
  if (e != cudaSuccess) { adb_rawfile_destroy(r); return fail(std::string("raw file upload failed: ") + cudaGetErrorString(e)); }
  *out = r;
  return 0;
}

int adb_rawfile4d_create(const adb_rawfile4d_desc* d, int device, adb_rawfile_t** out) {
  if (!d || !out) return fail("null argument");
  if (d->frames_per_cycle < 1 || d->scans < 1 || d->n_frames < 1 || d->n_tof < 1) return fail("empty raw file");
  if (d->scan_max_index != d->scans) return fail("scan_max_index must equal cycle.shape[2] (push = frame * scans + scan)");
  if (d->n_events >= 4294967000LL) return fail("raw files with more than 2^32 events are not supported");
  if (d->n_tof >= 2000000000LL) return fail("more than 2^31 tof bins are not supported");
  if ((d->n_frames + 1) * d->scans > 0xFFFFFFFFLL) return fail("push indices exceed 32 bits");
  const int64_t npos = d->frames_per_cycle * d->scans;
  int unique = 1;
  {  // observation ids: small non-negative integers; is every id bound to one frame of the cycle per scan?
    std::vector<int> seen(64);
    for (int64_t s = 0; s < d->scans; s++) {
      std::fill(seen.begin(), seen.end(), 0);
      for (int64_t f = 0; f < d->frames_per_cycle; f++) {
        const int64_t id = d->dia_precursor_cycle[f * d->scans + s];
        if (id >= 0 && id < 64 && seen[id]++) unique = 0;
      }
    }
  }
  for (int64_t t = 0; t < d->n_tof; t++)
    if (d->tof_indptr[t + 1] < d->tof_indptr[t] || d->tof_indptr[t] < 0 || d->tof_indptr[t + 1] > d->n_events)
      return fail("tof_indptr is not a valid CSR index (row " + std::to_string(t) + ")");
  if (set_device(device)) return 1;
  adb_rawfile* r = new adb_rawfile();
  r->device = device;
  r->is4d = 1;
  CUDA_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 6; i++) CUDA_TRY(cudaEventCreate(&r->ev[i]));
  CUDA_TRY(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i <= ADB_MAX_BLOCKS; i++) CUDA_TRY(cudaEventCreateWithFlags(&r->chunk_ev[i], cudaEventDisableTiming));
  cudaDeviceGetAttribute(&r->sm_count, cudaDevAttrMultiProcessorCount, device);
  DevRaw4& v = r->dev4;
  double *cyc, *rt, *mob, *mz; int64_t *dpc, *ip; uint32_t* push; uint16_t* it;
  if (upload(d->cycle, npos * 2, &cyc, r->allocs, r->bytes, r->stream) ||
      upload(d->dia_precursor_cycle, npos, &dpc, r->allocs, r->bytes, r->stream) ||
      upload(d->rt_values, d->n_frames, &rt, r->allocs, r->bytes, r->stream) ||
      upload(d->mobility_values, d->scans, &mob, r->allocs, r->bytes, r->stream) ||
      upload(d->mz_values, d->n_tof, &mz, r->allocs, r->bytes, r->stream) ||
      upload(d->tof_indptr, d->n_tof + 1, &ip, r->allocs, r->bytes, r->stream) ||
      upload(d->push_indices, d->n_events, &push, r->allocs, r->bytes, r->stream, 16) ||
      upload(d->intensity_values, d->n_events, &it, r->allocs, r->bytes, r->stream, 16)) {
    adb_rawfile_destroy(r);
    return 1;
  }
  v.cycle = cyc; v.Fr = d->frames_per_cycle; v.Sc = d->scans; v.dia_precursor_cycle = dpc; v.rt_values = rt;
  v.n_frames = d->n_frames; v.mobility_values = mob; v.mz_values = mz; v.n_tof = d->n_tof; v.tof_indptr = ip;
  v.push = push; v.intensity = it; v.n_events = d->n_events; v.zeroth_frame = d->zeroth_frame;
  v.precursor_cycle_max_index = d->precursor_cycle_max_index; v.scan_max_index = d->scan_max_index;
  v.frame_max_index = d->frame_max_index; v.obs_unique_per_scan = unique;
  r->rt4_host.assign(d->rt_values, d->rt_values + d->n_frames);
  r->mob4_host.assign(d->mobility_values, d->mobility_values + d->scans);
  void* st = nullptr;
  if (cudaMalloc(&st, sizeof(uint32_t)) != cudaSuccess) { adb_rawfile_destroy(r); return fail("cudaMalloc status failed"); }
  r->d_status = (uint32_t*)st;
  cudaMemsetAsync(r->d_status, 0, sizeof(uint32_t), r->stream);
  cudaError_t e = cudaStreamSynchronize(r->stream);
  if (e != cudaSuccess) { adb_rawfile_destroy(r); return fail(std::string("raw file upload failed: ") + cudaGetErrorString(e)); }
  *out = r;
  return 0;
}

void adb_rawfile_destroy(adb_rawfile_t* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  if (r->stream) cudaStreamSynchronize(r->stream);
  for (void* p : r->allocs) cudaFree(p);
  if (r->d_status) cudaFree(r->d_status);
  DeviceBuffer* bufs[] = {&r->kern, &r->order_keys, &r->order_vals, &r->order_tmp, &r->sel_ws, &r->cont, &r->cand_in,
                          &r->flags, &r->offs, &r->scan_tmp, &r->count, &r->scores, &r->score_ws, &r->dp_plan, &r->staging, &r->extent};
  for (DeviceBuffer* b : bufs) b->release();
  for (int i = 0; i < 6; i++) if (r->ev[i]) cudaEventDestroy(r->ev[i]);
  for (int i = 0; i <= ADB_MAX_BLOCKS; i++) if (r->chunk_ev[i]) cudaEventDestroy(r->chunk_ev[i]);
  if (r->copy_stream) { cudaStreamSynchronize(r->copy_stream); cudaStreamDestroy(r->copy_stream); }
  if (r->stream) cudaStreamDestroy(r->stream);
  delete r;
}

int64_t adb_rawfile_device_bytes(const adb_rawfile_t* r) { return r ? r->bytes : 0; }
void* adb_rawfile_stream(const adb_rawfile_t* r) { return r ? (void*)r->stream : nullptr; }

int adb_library_create(const adb_library_desc* d, int device, adb_library_t** out) {
  if (!d || !out) return fail("null argument");
  if (d->n_isotopes < 1) return fail("library needs at least one isotope column");
  if (set_device(device)) return 1;
  const int64_t P = d->n_precursors, NF = d->n_fragments;
  int mx = 0;
  for (int64_t i = 0; i < P; i++) {
    const int64_t s = d->frag_start_idx[i], e = d->frag_stop_idx[i];
    if (e > NF || s > e) return fail("fragment index range of precursor row " + std::to_string(i) + " is outside the fragment table");
    mx = std::max<int>(mx, (int)(e - s));
  }
  adb_library* l = new adb_library();
  l->device = device;
  l->max_lib_fragments = mx;
  // one pooled allocation, asynchronous copies (they overlap when the caller's arrays are pinned), one synchronise
  struct Item { const void* host; size_t bytes; size_t off; };
  Item items[17];
  size_t total = 0;
  auto add = [&](int k, const void* h, size_t bytes) { items[k] = Item{h, bytes, total}; total += (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255; };
  const size_t Pn = (size_t)std::max<int64_t>(P, 0), Fn = (size_t)std::max<int64_t>(NF, 0);
  add(0, d->precursor_idx, 4 * Pn); add(1, d->frag_start_idx, 4 * Pn); add(2, d->frag_stop_idx, 4 * Pn); add(3, d->charge, Pn);
  add(4, d->rt, 4 * Pn); add(5, d->mobility, 4 * Pn); add(6, d->mz, 4 * Pn); add(7, d->isotopes, 4 * Pn * (size_t)d->n_isotopes);
  // uncalibrated searches pass the SAME host column as library m/z and search m/z: it is uploaded once and shared
  const bool mz_alias = d->frag_mz == d->frag_mz_library;
  add(8, d->frag_mz_library, 4 * Fn); add(9, d->frag_mz, mz_alias ? 0 : 4 * Fn); add(10, d->frag_intensity, 4 * Fn); add(11, d->frag_type, Fn);
  add(12, d->frag_loss_type, Fn); add(13, d->frag_charge, Fn); add(14, d->frag_number, Fn); add(15, d->frag_position, Fn);
  add(16, d->frag_cardinality, Fn);
  void* pool = nullptr;
  size_t pool_bytes = 0;
  cudaError_t e = cudaSuccess;
  if (device >= 0 && device < 64) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    LibraryPoolCache& c = g_pool_cache[device];
    if (c.ptr && c.bytes >= total + 256) { pool = c.ptr; pool_bytes = c.bytes; c.ptr = nullptr; c.bytes = 0; }
  }
  if (!pool) {
    pool_bytes = total + total / 16 + 256;
    e = cudaMalloc(&pool, pool_bytes);
    if (e != cudaSuccess) { delete l; return fail(std::string("cudaMalloc(library) failed: ") + cudaGetErrorString(e)); }
  }
  l->pool = pool;
  l->pool_bytes = pool_bytes;
  l->bytes = (int64_t)total;
  for (int k = 0; k < 17 && e == cudaSuccess; k++)
    if (items[k].bytes) e = cudaMemcpyAsync((char*)pool + items[k].off, items[k].host, items[k].bytes, cudaMemcpyHostToDevice, cudaStreamPerThread);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
  if (e != cudaSuccess) { adb_library_destroy(l); return fail(std::string("library upload failed: ") + cudaGetErrorString(e)); }
  char* b = (char*)pool;
  DevLib& v = l->dev;
  v.n_precursors = P; v.n_isotopes = d->n_isotopes; v.n_fragments = NF;
  v.precursor_idx = (const uint32_t*)(b + items[0].off); v.frag_start_idx = (const uint32_t*)(b + items[1].off);
  v.frag_stop_idx = (const uint32_t*)(b + items[2].off); v.charge = (const uint8_t*)(b + items[3].off);
  v.rt = (const float*)(b + items[4].off); v.mobility = (const float*)(b + items[5].off); v.mz = (const float*)(b + items[6].off);
  v.isotopes = (const float*)(b + items[7].off); v.frag_mz_library = (const float*)(b + items[8].off);
  v.frag_mz = (const float*)(b + items[mz_alias ? 8 : 9].off); v.frag_intensity = (const float*)(b + items[10].off);
  v.frag_type = (const uint8_t*)(b + items[11].off); v.frag_loss_type = (const uint8_t*)(b + items[12].off);
  v.frag_charge = (const uint8_t*)(b + items[13].off); v.frag_number = (const uint8_t*)(b + items[14].off);
  v.frag_position = (const uint8_t*)(b + items[15].off); v.frag_cardinality = (const uint8_t*)(b + items[16].off);
  *out = l;
  return 0;
}

void adb_library_destroy(adb_library_t* l) {
  if (!l) return;
  cudaSetDevice(l->device);
  for (void* p : l->allocs) cudaFree(p);
  if (l->pool) {
    void* to_free = l->pool;
    if (l->device >= 0 && l->device < 64) {
      std::lock_guard<std::mutex> lock(g_pool_mutex);
      LibraryPoolCache& c = g_pool_cache[l->device];
      if (!c.ptr || c.bytes < l->pool_bytes) { to_free = c.ptr; c.ptr = l->pool; c.bytes = l->pool_bytes; }  // keep the larger one
    }
    if (to_free) cudaFree(to_free);
  }
  delete l;
}

int adb_fetch_candidates(adb_rawfile_t* raw, adb_candidates_out* out) {
  if (!raw || !out) return fail("null argument");
  if (out->n_rows != raw->cont_rows) return fail("candidate container size mismatch");
  if (set_device(raw->device)) return 1;
  cudaStream_t st = raw->stream;
  const size_t N = (size_t)raw->cont_rows;
  const DevCandidatesOut& c = raw->d_cont;
  CUDA_TRY(cudaMemcpyAsync(out->precursor_idx, c.precursor_idx, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->rank, c.rank, N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->score, c.score, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->scan_center, c.scan_center, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->scan_start, c.scan_start, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->scan_stop, c.scan_stop, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->frame_center, c.frame_center, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->frame_start, c.frame_start, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out->frame_stop, c.frame_stop, 4 * N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
namespace {
// asynchronous D2H of the resident candidate table on `st`
int copy_candidate_table(adb_rawfile* raw, adb_candidate_table* out, cudaStream_t st) {
  const size_t N = (size_t)raw->n_cand;
  const DevCandidatesIn& c = raw->d_cand;
  if (N > 0) {
    CUDA_TRY(cudaMemcpyAsync(out->lib_row, c.lib_row, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->rank, c.rank, N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->scan_start, c.scan_start, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->scan_stop, c.scan_stop, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->scan_center, c.scan_center, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->frame_start, c.frame_start, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->frame_stop, c.frame_stop, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->frame_center, c.frame_center, 8 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->precursor_idx, raw->d_cand_pidx, 4 * N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->score, raw->d_cand_score, 4 * N, cudaMemcpyDeviceToHost, st));
  }
  return 0;
}
}  // namespace
extern "C" {

int adb_fetch_candidate_table(adb_rawfile_t* raw, adb_candidate_table* out) {
  if (!raw || !out) return fail("null argument");
  if (out->n != raw->n_cand) return fail("candidate table size mismatch (expected the count returned by adb_select_candidates_resident)");
  if (!raw->d_cand_pidx) return fail("no resident candidate table (call adb_select_candidates_resident first)");
  if (set_device(raw->device)) return 1;
  if (copy_candidate_table(raw, out, raw->stream)) return 1;
  CUDA_TRY(cudaStreamSynchronize(raw->stream));
  return 0;
}

int adb_fetch_scores(adb_rawfile_t* raw, adb_scores_out* out, int64_t* lib_row, uint8_t* rank) {
  if (!raw || !out) return fail("null argument");
  if (set_device(raw->device)) return 1;
  cudaStream_t st = raw->stream;
  const size_t N = (size_t)raw->scores_n, K = (size_t)raw->scores_k;
  const DevScoresOut& s = raw->d_scores;
  if (N > 0) {
    CUDA_TRY(cudaMemcpyAsync(out->features, s.features, 4 * N * ADB_NUM_FEATURES, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out->valid, s.valid, N, cudaMemcpyDeviceToHost, st));
    float* hf[] = {out->fragment_mz_library, out->fragment_mz, out->fragment_mz_observed, out->fragment_height,
                   out->fragment_intensity, out->fragment_mass_error, out->fragment_correlation};
    float* df[] = {s.fragment_mz_library, s.fragment_mz, s.fragment_mz_observed, s.fragment_height,
                   s.fragment_intensity, s.fragment_mass_error, s.fragment_correlation};
    for (int i = 0; i < 7; i++) CUDA_TRY(cudaMemcpyAsync(hf[i], df[i], 4 * N * K, cudaMemcpyDeviceToHost, st));
    uint8_t* hu[] = {out->fragment_position, out->fragment_number, out->fragment_type, out->fragment_charge, out->fragment_loss_type};
    uint8_t* du[] = {s.fragment_position, s.fragment_number, s.fragment_type, s.fragment_charge, s.fragment_loss_type};
    for (int i = 0; i < 5; i++) CUDA_TRY(cudaMemcpyAsync(hu[i], du[i], N * K, cudaMemcpyDeviceToHost, st));
    if (lib_row) CUDA_TRY(cudaMemcpyAsync(lib_row, raw->d_cand.lib_row, 8 * N, cudaMemcpyDeviceToHost, st));
    if (rank) CUDA_TRY(cudaMemcpyAsync(rank, raw->d_cand.rank, N, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int adb_select_candidates(adb_rawfile_t* raw, adb_library_t* lib, const adb_selection_config* cfg, const float* kernel,
                          int32_t kh, int32_t kw, adb_candidates_out* out) {
  if (!raw || !lib || !cfg || !kernel || !out) return fail("null argument");
  if (out->n_rows != lib->dev.n_precursors * cfg->candidate_count) return fail("candidate container must have n_precursors * candidate_count rows");
  if (run_selection(raw, lib, cfg, kernel, kh, kw)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[2], raw->stream));
  if (check_status(raw, "adb_select_candidates")) return 1;
  if (adb_fetch_candidates(raw, out)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[3], raw->stream));
  finish_timing(raw);
  return 0;
}

int adb_select_candidates_resident(adb_rawfile_t* raw, adb_library_t* lib, const adb_selection_config* cfg,
                                   const float* kernel, int32_t kh, int32_t kw, int64_t* n_candidates) {
  if (!raw || !lib || !cfg || !kernel) return fail("null argument");
  if (run_selection(raw, lib, cfg, kernel, kh, kw)) return 1;
  if (run_compaction(raw)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[2], raw->stream));
  if (check_status(raw, "adb_select_candidates_resident")) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[3], raw->stream));
  finish_timing(raw);
  if (n_candidates) *n_candidates = raw->n_cand;
  return 0;
}

int adb_score_candidates(adb_rawfile_t* raw, adb_library_t* lib, const adb_scoring_config* cfg,
                         const adb_candidates_in* cand, adb_scores_out* out) {
  if (!raw || !lib || !cfg || !cand || !out) return fail("null argument");
  if (set_device(raw->device)) return 1;
  cudaStream_t st = raw->stream;
  CUDA_TRY(cudaEventRecord(raw->ev[0], st));
  if (upload_candidates(raw, cand)) return 1;
  // validation and the scratch extents (longest candidate in frames / scans) are computed on the device
  int64_t s_max = 0, f_max = 0, bad_row = 0;
  if (resident_extents(raw, &s_max, &f_max, lib->dev.n_precursors, &bad_row)) return 1;
  if (bad_row) return fail("a candidate refers to a precursor outside the library");
  const int64_t c_max = raw->is4d ? f_max : f_max / raw->dev.cycle_len + 1;
  CUDA_TRY(cudaEventRecord(raw->ev[1], st));
  if (!raw->is4d && c_max > 4096) return fail("a candidate spans more than 4096 cycles");
  if (run_scoring(raw, lib, cfg, c_max, s_max, out)) return 1;  // records ev[2] between the kernels and the (remaining) D2H
  CUDA_TRY(cudaEventRecord(raw->ev[3], st));
  if (check_status(raw, "adb_score_candidates")) return 1;
  finish_timing(raw);
  return 0;
}

int adb_score_candidates_ragged(adb_rawfile_t* raw, adb_library_t* lib, const adb_scoring_config* cfg,
                                const adb_candidates_in* cand, adb_scores_ragged* out) {
  if (!raw || !lib || !cfg || !cand || !out) return fail("null argument");
  if (set_device(raw->device)) return 1;
  cudaStream_t st = raw->stream;
  CUDA_TRY(cudaEventRecord(raw->ev[0], st));
  if (upload_candidates(raw, cand)) return 1;
  int64_t s_max = 0, f_max = 0, bad_row = 0;
  if (resident_extents(raw, &s_max, &f_max, lib->dev.n_precursors, &bad_row)) return 1;
  if (bad_row) return fail("a candidate refers to a precursor outside the library");
  const int64_t c_max = raw->is4d ? f_max : f_max / raw->dev.cycle_len + 1;
  CUDA_TRY(cudaEventRecord(raw->ev[1], st));
  if (!raw->is4d && c_max > 4096) return fail("a candidate spans more than 4096 cycles");
  RaggedRun R;
  R.out = out;
  out->n_rows = out->n_fragments = 0;
  if (run_scoring(raw, lib, cfg, c_max, s_max, nullptr, &R)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[3], st));
  if (check_status(raw, "adb_score_candidates_ragged")) return 1;
  CUDA_TRY(cudaStreamSynchronize(raw->copy_stream));
  finish_timing(raw);
  out->n_rows = raw->rag_host_tot[2 * R.n_blocks];
  out->n_fragments = raw->rag_host_tot[2 * R.n_blocks + 1];
  if (R.overflow || out->n_rows > out->row_capacity || out->n_fragments > out->frag_capacity)
    return fail("adb_score_candidates_ragged: output capacity too small (needs " + std::to_string(out->n_rows) + " rows, " +
                std::to_string(out->n_fragments) + " fragment entries)");
  out->frag_offset[out->n_rows] = out->n_fragments;
  return 0;
}

int adb_select_score_candidates_ragged(adb_rawfile_t* raw, adb_library_t* lib, const adb_selection_config* sel_cfg, const float* kernel,
                                       int32_t kh, int32_t kw, float score_cutoff, const adb_scoring_config* cfg,
                                       adb_candidate_table* table, adb_scores_ragged* out) {
  if (!raw || !lib || !sel_cfg || !kernel || !cfg || !out) return fail("null argument");
  if (run_selection(raw, lib, sel_cfg, kernel, kh, kw)) return 1;
  if (run_compaction(raw, score_cutoff)) return 1;
  if (check_status(raw, "adb_select_score_candidates_ragged (selection)")) return 1;
  cudaStream_t st = raw->stream;
  if (table) {
    if (table->n < raw->n_cand) {
      const int64_t need = raw->n_cand;
      table->n = need;
      return fail("adb_select_score_candidates_ragged: candidate table capacity too small (needs " + std::to_string(need) + " rows)");
    }
    table->n = raw->n_cand;
    // the table travels on the copy stream while the first scoring block runs
    CUDA_TRY(cudaEventRecord(raw->chunk_ev[ADB_MAX_BLOCKS], st));
    CUDA_TRY(cudaStreamWaitEvent(raw->copy_stream, raw->chunk_ev[ADB_MAX_BLOCKS], 0));
    if (copy_candidate_table(raw, table, raw->copy_stream)) return 1;
  }
  int64_t s_max = 0, f_max = 0;
  if (resident_extents(raw, &s_max, &f_max)) return 1;
  const int64_t c_max = raw->is4d ? f_max : f_max / raw->dev.cycle_len + 1;
  if (!raw->is4d && c_max > 4096) return fail("a candidate spans more than 4096 cycles");
  RaggedRun R;
  R.out = out;
  out->n_rows = out->n_fragments = 0;
  if (run_scoring(raw, lib, cfg, c_max, s_max, nullptr, &R)) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[3], st));
  if (check_status(raw, "adb_select_score_candidates_ragged")) return 1;
  CUDA_TRY(cudaStreamSynchronize(raw->copy_stream));
  out->n_rows = raw->rag_host_tot[2 * R.n_blocks];
  out->n_fragments = raw->rag_host_tot[2 * R.n_blocks + 1];
  if (R.overflow || out->n_rows > out->row_capacity || out->n_fragments > out->frag_capacity)
    return fail("adb_select_score_candidates_ragged: output capacity too small (needs " + std::to_string(out->n_rows) + " rows, " +
                std::to_string(out->n_fragments) + " fragment entries)");
  out->frag_offset[out->n_rows] = out->n_fragments;
  return 0;
}

int adb_score_candidates_resident(adb_rawfile_t* raw, adb_library_t* lib, const adb_scoring_config* cfg) {
  if (!raw || !lib || !cfg) return fail("null argument");
  if (set_device(raw->device)) return 1;
  cudaStream_t st = raw->stream;
  CUDA_TRY(cudaEventRecord(raw->ev[0], st));
  CUDA_TRY(cudaEventRecord(raw->ev[1], st));
  // extents of the resident candidate table (join_close_candidates and a user-set max_size_rt make windows of any length)
  // and the library-row bounds check, as adb_score_candidates does for a caller's table
  int64_t s_max = 0, f_max = 0, bad_row = 0;
  if (resident_extents(raw, &s_max, &f_max, lib->dev.n_precursors, &bad_row)) return 1;
  if (bad_row) return fail("a resident candidate refers to a precursor outside the library");
  if (raw->is4d) {
    if (run_scoring(raw, lib, cfg, f_max, s_max)) return 1;
  } else {
    const int64_t c_max = f_max / raw->dev.cycle_len + 1;
    if (c_max > 4096) return fail("a candidate spans more than 4096 cycles");
    if (run_scoring(raw, lib, cfg, c_max)) return 1;
  }
  CUDA_TRY(cudaEventRecord(raw->ev[2], st));
  if (check_status(raw, "adb_score_candidates_resident")) return 1;
  CUDA_TRY(cudaEventRecord(raw->ev[3], st));
  finish_timing(raw);
  return 0;
}

int adb_resident_score_table(adb_rawfile_t* raw, void** features, void** valid, void** lib_row, void** rank, int64_t* n_rows) {
  if (!raw) return fail("null argument");
  if (features) *features = raw->d_scores.features;
  if (valid) *valid = raw->d_scores.valid;
  if (lib_row) *lib_row = (void*)raw->d_cand.lib_row;
  if (rank) *rank = (void*)raw->d_cand.rank;
  if (n_rows) *n_rows = raw->scores_n;
  return 0;
}

int adb_last_timing(const adb_rawfile_t* raw, float* h2d_ms, float* kernel_ms, float* d2h_ms) {
  if (!raw) return fail("null argument");
  if (h2d_ms) *h2d_ms = raw->h2d_ms;
  if (kernel_ms) *kernel_ms = raw->kernel_ms;
  if (d2h_ms) *d2h_ms = raw->d2h_ms;
  return 0;
}

int64_t adb_kernel_launches(const adb_rawfile_t* raw) { return raw ? raw->launches : 0; }
float adb_last_main_kernel_ms(const adb_rawfile_t* raw) { return raw ? raw->main_kernel_ms : 0.f; }

int adb_fragment_competition(int device, int64_t n_windows, const int64_t* window_start, const int64_t* window_stop,
                             int64_t n_psm, const void* rt, const int64_t* frag_start_idx, const int64_t* frag_stop_idx,
                             int64_t n_frag, const void* fragment_mz, int32_t is_f64, double rt_tol_seconds,
                             double mass_tol_ppm, uint8_t* valid) {
  if (n_windows < 0 || n_psm < 0 || n_frag < 0) return fail("negative size");
  if (n_windows == 0 || n_psm == 0) return 0;
  if (!window_start || !window_stop || !rt || !frag_start_idx || !frag_stop_idx || !valid) return fail("null argument");
  if (n_frag > 0 && !fragment_mz) return fail("null argument");
  for (int64_t w = 0; w < n_windows; w++)
    if (window_start[w] < 0 || window_stop[w] > n_psm || window_start[w] > window_stop[w]) return fail("window range outside the PSM table");
  for (int64_t i = 0; i < n_psm; i++)
    if (frag_start_idx[i] < 0 || frag_stop_idx[i] > n_frag || frag_start_idx[i] > frag_stop_idx[i]) return fail("fragment range outside the fragment table");
  if (set_device(device)) return 1;
  const size_t es_rt = (is_f64 & 1) ? 8 : 4, es = (is_f64 & 2) ? 8 : 4;
  // grow-only device copies of the arguments, kept per device between calls (the call is a few ms in total)
  static std::mutex fc_mutex;
  static DeviceBuffer fc_bufs[64][7];
  std::lock_guard<std::mutex> fc_lock(fc_mutex);
  DeviceBuffer* fb = fc_bufs[device & 63];
  DeviceBuffer &b_ws = fb[0], &b_we = fb[1], &b_rt = fb[2], &b_fs = fb[3], &b_fe = fb[4], &b_mz = fb[5], &b_valid = fb[6];
  auto cleanup = [&]() {};
  if (b_ws.reserve(8 * (size_t)n_windows) || b_we.reserve(8 * (size_t)n_windows) || b_rt.reserve(es_rt * (size_t)n_psm) ||
      b_fs.reserve(8 * (size_t)n_psm) || b_fe.reserve(8 * (size_t)n_psm) || b_mz.reserve(es * (size_t)std::max<int64_t>(n_frag, 1)) ||
      b_valid.reserve((size_t)n_psm)) { cleanup(); return 1; }
  cudaError_t e = cudaSuccess;
  auto cp = [&](void* d, const void* h, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); };
  cp(b_ws.ptr, window_start, 8 * (size_t)n_windows); cp(b_we.ptr, window_stop, 8 * (size_t)n_windows);
  cp(b_rt.ptr, rt, es_rt * (size_t)n_psm); cp(b_fs.ptr, frag_start_idx, 8 * (size_t)n_psm); cp(b_fe.ptr, frag_stop_idx, 8 * (size_t)n_psm);
  cp(b_mz.ptr, fragment_mz, es * (size_t)n_frag); cp(b_valid.ptr, valid, (size_t)n_psm);
  if (e != cudaSuccess) { cleanup(); return fail(std::string("fragcomp H2D failed: ") + cudaGetErrorString(e)); }
  int launches = 0;
  // conflict graph over the RT-sorted windows; the window-serial kernel only when the pair list would not fit (or for A/B
  // measurements with ADB_FRAGCOMP_SERIAL=1)
  const char* serial = getenv("ADB_FRAGCOMP_SERIAL");
  int rc = 2;
  if (!(serial && serial[0] == '1')) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    rc = adb_run_fragcomp_graph(n_windows, b_ws.as<int64_t>(), b_we.as<int64_t>(), n_psm, b_rt.ptr, b_fs.as<int64_t>(), b_fe.as<int64_t>(),
                                b_mz.ptr, is_f64, rt_tol_seconds, mass_tol_ppm, b_valid.as<uint8_t>(), free_b / 16, nullptr, &launches);
    if (rc == 1) { cleanup(); return fail(std::string("fragcomp kernels failed: ") + cudaGetErrorString(cudaGetLastError())); }
  }
  if (rc == 2)
    adb_launch_fragcomp(n_windows, b_ws.as<int64_t>(), b_we.as<int64_t>(), b_rt.ptr, b_fs.as<int64_t>(), b_fe.as<int64_t>(),
                        b_mz.ptr, is_f64, rt_tol_seconds, mass_tol_ppm, b_valid.as<uint8_t>(), nullptr, &launches);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(valid, b_valid.ptr, (size_t)n_psm, cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(std::string("fragcomp kernel failed: ") + cudaGetErrorString(e));
  return 0;
}

int adb_transpose_csr(int device, int64_t n_values, int64_t n_push, int64_t n_tof, const uint32_t* tof_indices,
                      const int64_t* push_indptr, const uint16_t* values, uint32_t* push_indices_out, int64_t* tof_indptr_out,
                      uint16_t* values_out) {
  if (n_values < 0 || n_push < 0 || n_tof < 0) return fail("negative size");
  if (!push_indptr || !tof_indptr_out) return fail("null argument");
  if (n_values > 0 && (!tof_indices || !values || !push_indices_out || !values_out)) return fail("null argument");
  if (n_values >= 4294967000LL || n_push >= 4294967000LL) return fail("more than 2^32 events or pushes are not supported");
  if (push_indptr[0] != 0 || push_indptr[n_push] != n_values) return fail("push_indptr does not span the value array");
  for (int64_t p = 0; p < n_push; p++)
    if (push_indptr[p + 1] < push_indptr[p]) return fail("push_indptr is not monotone (push " + std::to_string(p) + ")");
  if (set_device(device)) return 1;
  const size_t N = (size_t)std::max<int64_t>(n_values, 1);
  DeviceBuffer b_tof, b_tof_sorted, b_idx, b_idx_sorted, b_push, b_ptr, b_val, b_push_out, b_val_out, b_indptr, b_tmp;
  DeviceBuffer* all[] = {&b_tof, &b_tof_sorted, &b_idx, &b_idx_sorted, &b_push, &b_ptr, &b_val, &b_push_out, &b_val_out, &b_indptr, &b_tmp};
  auto cleanup = [&]() { for (DeviceBuffer* b : all) b->release(); };
  const int end_bit = 32;  // all key bits: out-of-range tof indices must sort last so they can be detected
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int64_t)n_values, 0, end_bit, (cudaStream_t)0);
  if (b_tof.reserve(4 * N) || b_tof_sorted.reserve(4 * N) || b_idx.reserve(4 * N) || b_idx_sorted.reserve(4 * N) || b_push.reserve(4 * N) ||
      b_ptr.reserve(8 * (size_t)(n_push + 1)) || b_val.reserve(2 * N) || b_push_out.reserve(4 * N) || b_val_out.reserve(2 * N) ||
      b_indptr.reserve(8 * (size_t)(n_tof + 1)) || b_tmp.reserve(tmp_bytes + 16)) { cleanup(); return 1; }
  cudaError_t e = cudaSuccess;
  auto h2d = [&](void* d, const void* h, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); };
  h2d(b_tof.ptr, tof_indices, 4 * (size_t)n_values); h2d(b_ptr.ptr, push_indptr, 8 * (size_t)(n_push + 1)); h2d(b_val.ptr, values, 2 * (size_t)n_values);
  if (e != cudaSuccess) { cleanup(); return fail(std::string("transpose H2D failed: ") + cudaGetErrorString(e)); }
  if (n_values > 0) {
    transpose_push_of_event_kernel<<<(unsigned)((n_push * 32 + 255) / 256), 256>>>(b_ptr.as<int64_t>(), n_push, b_push.as<uint32_t>(), b_idx.as<uint32_t>());
    // stable radix sort by tof index: the input is push-major, so pushes stay ascending inside every tof row (bruker.py:165-182)
    cub::DeviceRadixSort::SortPairs(b_tmp.ptr, tmp_bytes, b_tof.as<uint32_t>(), b_tof_sorted.as<uint32_t>(), b_idx.as<uint32_t>(),
                                    b_idx_sorted.as<uint32_t>(), (int64_t)n_values, 0, end_bit, (cudaStream_t)0);
    transpose_gather_kernel<<<(unsigned)((n_values + 255) / 256), 256>>>(b_idx_sorted.as<uint32_t>(), b_push.as<uint32_t>(), b_val.as<uint16_t>(),
                                                                       n_values, b_push_out.as<uint32_t>(), b_val_out.as<uint16_t>());
  }
  transpose_indptr_kernel<<<(unsigned)((n_tof + 1 + 255) / 256), 256>>>(b_tof_sorted.as<uint32_t>(), n_values, n_tof, b_indptr.as<int64_t>());
  uint32_t last_tof = 0;
  if (n_values > 0) e = cudaMemcpy(&last_tof, b_tof_sorted.as<uint32_t>() + (n_values - 1), 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess && n_values > 0 && (int64_t)last_tof >= n_tof) { cleanup(); return fail("a tof index is >= n_tof_indices"); }
  auto d2h = [&](void* h, const void* d, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost); };
  d2h(push_indices_out, b_push_out.ptr, 4 * (size_t)n_values); d2h(values_out, b_val_out.ptr, 2 * (size_t)n_values);
  d2h(tof_indptr_out, b_indptr.ptr, 8 * (size_t)(n_tof + 1));
  cleanup();
  if (e != cudaSuccess) return fail(std::string("transpose failed: ") + cudaGetErrorString(e));
  return 0;
}

}  // extern "C"

// =====================================================================================================================
// FDR bookkeeping next to fragment competition: q-values and best-row-per-group, alphadia/fdr/fdr.py:195-297.
// Both are "stable multi-column sort, then a scan"; the sorts are least-significant-key-first passes of the stable
// cub::DeviceRadixSort over (key, row index) pairs.
namespace {

__device__ __forceinline__ uint64_t ordered_bits64(double v) {
  uint64_t b = (uint64_t)__double_as_longlong(v);
  if ((b << 1) == 0) b = 0;  // -0.0 and +0.0 compare equal in pandas' sort: give them one key
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void fdr_minor_key_kernel(const uint8_t* __restrict__ decoy, const uint64_t* __restrict__ extra, int64_t n, uint64_t* keys,
                                     uint32_t* idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = ((uint64_t)(decoy ? decoy[i] : 0) << 63) | extra[i];  // (decoy, extra) in one key; extra < 2^63 is checked on the host
  idx[i] = (uint32_t)i;
}

__global__ void fdr_score_key_kernel(const double* __restrict__ score, const uint32_t* __restrict__ perm, int64_t n, uint64_t* keys,
                                     uint32_t* idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = perm ? perm[i] : (uint32_t)i;
  keys[i] = ordered_bits64(score[r]);
  if (idx) idx[i] = (uint32_t)i;
}

__global__ void fdr_gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n, uint64_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[perm[i]];
}

__global__ void fdr_gather_decoy_kernel(const uint8_t* __restrict__ decoy, const uint32_t* __restrict__ perm, int64_t n, int64_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)decoy[perm[i]];
}

// fdr.py:288-293: decoy_cumsum / target_cumsum in float64, written back to front for the running minimum
__global__ void fdr_ratio_reversed_kernel(const int64_t* __restrict__ decoy_cumsum, int64_t n, double* fdr_reversed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t d = decoy_cumsum[i];
  fdr_reversed[n - 1 - i] = (double)d / (double)((i + 1) - d);
}

__global__ void fdr_unreverse_kernel(const double* __restrict__ q_reversed, const uint32_t* __restrict__ perm, int64_t n, double* qval,
                                     int64_t* order) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  qval[i] = q_reversed[n - 1 - i];
  order[i] = (int64_t)perm[i];
}

__global__ void fdr_group_head_kernel(const uint64_t* __restrict__ sorted_group, const uint32_t* __restrict__ perm, int64_t n, uint8_t* keep) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || sorted_group[i] != sorted_group[i - 1]) keep[perm[i]] = 1;
}

struct RunningMin {  // np.minimum.accumulate (fdr.py:213); no NaN can occur: i + 1 >= 1 makes 0/0 impossible
  __host__ __device__ __forceinline__ double operator()(double a, double b) const { return b < a ? b : a; }
};

int fdr_check_scores(const double* score, int64_t n, const char* what) {
  for (int64_t i = 0; i < n; i++)
    if (score[i] != score[i]) return fail(std::string(what) + ": NaN score in row " + std::to_string(i));
  return 0;
}

}  // namespace

extern "C" {

int adb_q_values(int device, int64_t n, const double* score, const uint8_t* decoy, const uint64_t* extra_key, int64_t* order_out,
                 double* qval_out) {
  if (n < 0) return fail("negative size");
  if (n == 0) return 0;
  if (!score || !decoy || !extra_key || !order_out || !qval_out) return fail("null argument");
  if (n >= 2147483000LL) return fail("more than 2^31 rows are not supported");
  if (fdr_check_scores(score, n, "adb_q_values")) return 1;
  for (int64_t i = 0; i < n; i++) {
    if (decoy[i] > 1) return fail("decoy column must hold 0 (target) or 1 (decoy)");
    if (extra_key[i] >> 63) return fail("extra sort key must be below 2^63");
  }
  if (set_device(device)) return 1;
  const size_t N = (size_t)n;
  DeviceBuffer b_score, b_decoy, b_extra, b_key, b_key2, b_idx, b_idx2, b_cum, b_f, b_tmp;
  DeviceBuffer* all[] = {&b_score, &b_decoy, &b_extra, &b_key, &b_key2, &b_idx, &b_idx2, &b_cum, &b_f, &b_tmp};
  auto cleanup = [&]() { for (DeviceBuffer* b : all) b->release(); };
  size_t sort_bytes = 0, sum_bytes = 0, min_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, n, 0, 64, (cudaStream_t)0);
  cub::DeviceScan::InclusiveSum(nullptr, sum_bytes, (const int64_t*)nullptr, (int64_t*)nullptr, n, (cudaStream_t)0);
  cub::DeviceScan::InclusiveScan(nullptr, min_bytes, (const double*)nullptr, (double*)nullptr, RunningMin(), n, (cudaStream_t)0);
  const size_t tmp_bytes = std::max(sort_bytes, std::max(sum_bytes, min_bytes));
  if (b_score.reserve(8 * N) || b_decoy.reserve(N) || b_extra.reserve(8 * N) || b_key.reserve(8 * N) || b_key2.reserve(8 * N) ||
      b_idx.reserve(4 * N) || b_idx2.reserve(4 * N) || b_cum.reserve(8 * N) || b_f.reserve(8 * N) || b_tmp.reserve(tmp_bytes + 16)) {
    cleanup();
    return 1;
  }
  cudaError_t e = cudaSuccess;
  auto h2d = [&](void* d, const void* h, size_t bytes) { if (e == cudaSuccess) e = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); };
  h2d(b_score.ptr, score, 8 * N); h2d(b_decoy.ptr, decoy, N); h2d(b_extra.ptr, extra_key, 8 * N);
  if (e != cudaSuccess) { cleanup(); return fail(std::string("adb_q_values H2D failed: ") + cudaGetErrorString(e)); }
  const unsigned grid = (unsigned)((n + 255) / 256);
  size_t tb = tmp_bytes;
  // df.sort_values([score, decoy, *extra]) (fdr.py:283-285): minor keys first, then the score; both passes are stable
  fdr_minor_key_kernel<<<grid, 256>>>(b_decoy.as<uint8_t>(), b_extra.as<uint64_t>(), n, b_key.as<uint64_t>(), b_idx.as<uint32_t>());
  cub::DeviceRadixSort::SortPairs(b_tmp.ptr, tb, b_key.as<uint64_t>(), b_key2.as<uint64_t>(), b_idx.as<uint32_t>(), b_idx2.as<uint32_t>(),
                                  n, 0, 64, (cudaStream_t)0);
  fdr_score_key_kernel<<<grid, 256>>>(b_score.as<double>(), b_idx2.as<uint32_t>(), n, b_key.as<uint64_t>(), nullptr);
  tb = tmp_bytes;
  cub::DeviceRadixSort::SortPairs(b_tmp.ptr, tb, b_key.as<uint64_t>(), b_key2.as<uint64_t>(), b_idx2.as<uint32_t>(), b_idx.as<uint32_t>(),
                                  n, 0, 64, (cudaStream_t)0);
  // cumulative decoys / cumulative targets, then the running minimum from the back (fdr.py:286-293, 211-214)
  fdr_gather_decoy_kernel<<<grid, 256>>>(b_decoy.as<uint8_t>(), b_idx.as<uint32_t>(), n, b_key.as<int64_t>());
  tb = tmp_bytes;
  cub::DeviceScan::InclusiveSum(b_tmp.ptr, tb, b_key.as<int64_t>(), b_cum.as<int64_t>(), n, (cudaStream_t)0);
  fdr_ratio_reversed_kernel<<<grid, 256>>>(b_cum.as<int64_t>(), n, b_f.as<double>());
  tb = tmp_bytes;
  cub::DeviceScan::InclusiveScan(b_tmp.ptr, tb, b_f.as<double>(), b_key2.as<double>(), RunningMin(), n, (cudaStream_t)0);
  fdr_unreverse_kernel<<<grid, 256>>>(b_key2.as<double>(), b_idx.as<uint32_t>(), n, b_f.as<double>(), b_cum.as<int64_t>());
  e = cudaGetLastError();
  auto d2h = [&](void* h, const void* d, size_t bytes) { if (e == cudaSuccess) e = cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost); };
  d2h(qval_out, b_f.ptr, 8 * N); d2h(order_out, b_cum.ptr, 8 * N);
  cleanup();
  if (e != cudaSuccess) return fail(std::string("adb_q_values failed: ") + cudaGetErrorString(e));
  return 0;
}

int adb_keep_best(int device, int64_t n, const double* score, const uint64_t* group_key, uint8_t* keep_out) {
  if (n < 0) return fail("negative size");
  if (n == 0) return 0;
  if (!score || !group_key || !keep_out) return fail("null argument");
  if (n >= 2147483000LL) return fail("more than 2^31 rows are not supported");
  if (fdr_check_scores(score, n, "adb_keep_best")) return 1;
  if (set_device(device)) return 1;
  const size_t N = (size_t)n;
  DeviceBuffer b_score, b_group, b_key, b_key2, b_idx, b_idx2, b_keep, b_tmp;
  DeviceBuffer* all[] = {&b_score, &b_group, &b_key, &b_key2, &b_idx, &b_idx2, &b_keep, &b_tmp};
  auto cleanup = [&]() { for (DeviceBuffer* b : all) b->release(); };
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, n, 0, 64, (cudaStream_t)0);
  if (b_score.reserve(8 * N) || b_group.reserve(8 * N) || b_key.reserve(8 * N) || b_key2.reserve(8 * N) || b_idx.reserve(4 * N) ||
      b_idx2.reserve(4 * N) || b_keep.reserve(N) || b_tmp.reserve(tmp_bytes + 16)) {
    cleanup();
    return 1;
  }
  cudaError_t e = cudaSuccess;
  auto h2d = [&](void* d, const void* h, size_t bytes) { if (e == cudaSuccess) e = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); };
  h2d(b_score.ptr, score, 8 * N); h2d(b_group.ptr, group_key, 8 * N);
  if (e == cudaSuccess) e = cudaMemset(b_keep.ptr, 0, N);
  if (e != cudaSuccess) { cleanup(); return fail(std::string("adb_keep_best H2D failed: ") + cudaGetErrorString(e)); }
  const unsigned grid = (unsigned)((n + 255) / 256);
  size_t tb = tmp_bytes;
  // sort_values([score, *group]) then groupby(group).head(1) (fdr.py:219-224) keeps, per group, the lowest score and among
  // equal scores the earliest row: order the rows by (group, score, row) and take every group's first
  fdr_score_key_kernel<<<grid, 256>>>(b_score.as<double>(), nullptr, n, b_key.as<uint64_t>(), b_idx.as<uint32_t>());
  cub::DeviceRadixSort::SortPairs(b_tmp.ptr, tb, b_key.as<uint64_t>(), b_key2.as<uint64_t>(), b_idx.as<uint32_t>(), b_idx2.as<uint32_t>(),
                                  n, 0, 64, (cudaStream_t)0);
  fdr_gather_u64_kernel<<<grid, 256>>>(b_group.as<uint64_t>(), b_idx2.as<uint32_t>(), n, b_key.as<uint64_t>());
  tb = tmp_bytes;
  cub::DeviceRadixSort::SortPairs(b_tmp.ptr, tb, b_key.as<uint64_t>(), b_key2.as<uint64_t>(), b_idx2.as<uint32_t>(), b_idx.as<uint32_t>(),
                                  n, 0, 64, (cudaStream_t)0);
  fdr_group_head_kernel<<<grid, 256>>>(b_key2.as<uint64_t>(), b_idx.as<uint32_t>(), n, b_keep.as<uint8_t>());
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(keep_out, b_keep.ptr, N, cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(std::string("adb_keep_best failed: ") + cudaGetErrorString(e));
  return 0;
}

}  // extern "C"

// alphadia_b200 — candidate scoring, 3-D raw files: kernels and launcher of the data-parallel passes (adb_score_dp.cuh).
//
// Compiled with --fmad=false: the pass bodies are written in plain arithmetic and must not be contracted (the reference
// accumulates sequentially in f32 / f64 without FMA).
#include <algorithm>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "adb_score_dp.cuh"

namespace {

constexpr int DP_THREADS = 256;
#ifndef DP_LB_EXTRACT
#define DP_LB_EXTRACT 5
#endif
#ifndef DP_LB_TEMPLATE
#define DP_LB_TEMPLATE 4
#endif
#ifndef DP_LB_FRAGMENT
#define DP_LB_FRAGMENT 4
#endif

__global__ void dp_wtab_p_kernel(double* tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 2 * DP_WTAB_P_STRIDE) tab[t] = dp_wtab_p_entry(t / DP_WTAB_P_STRIDE, t % DP_WTAB_P_STRIDE);
}

__global__ void __launch_bounds__(DP_THREADS) dp_setup_kernel(const __grid_constant__ DpParams P) {
  const int64_t j = (int64_t)blockIdx.x * DP_THREADS + threadIdx.x;
  if (j < P.n) dp_setup(P, j);
}

// tile_need[t] = DP_W * the largest block of tile t; entry n_tiles = 0 (the scan runs over n_tiles + 1 entries: off[n_tiles] = total)
__global__ void __launch_bounds__(DP_THREADS) dp_tile_need_kernel(const __grid_constant__ DpParams P) {
  const int64_t t = (int64_t)blockIdx.x * DP_THREADS + threadIdx.x;
  const int64_t n_tiles = (P.n + DP_W - 1) / DP_W;
  if (t > n_tiles) return;
  int64_t m = 0;
  if (t < n_tiles)
    for (int64_t j = t * DP_W; j < min((t + 1) * (int64_t)DP_W, P.n); j++) m = max(m, P.need[j]);
  P.tile_need[t] = m * DP_W;
}

// thread t <-> (slot, row): rows 0 .. KS-1 are the fragment rows, KS .. KS+nIcap-1 the isotope rows
__global__ void __launch_bounds__(DP_THREADS, DP_LB_EXTRACT) dp_extract_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;  // a batch has < 2^32 rows: 32-bit index arithmetic
  uint32_t j, r;
  dp_decode(t, (uint32_t)(P.KS + P.nIcap), j, r);
  if (j < P.n) dp_extract(P, j, (int)r);
  else if (r < (uint32_t)P.KS && j < (uint32_t)dp_padded_slots(P.n)) P.rowflag[dp_encode(j, r, (uint32_t)P.KS)] = 0;  // padding slots of the last tile
}

__global__ void __launch_bounds__(DP_THREADS, DP_LB_TEMPLATE) dp_template_kernel(const __grid_constant__ DpParams P) {
  const int64_t j = (int64_t)blockIdx.x * DP_THREADS + threadIdx.x;
  if (j < P.n) dp_template(P, j);
}

// thread t <-> t-th fragment row with signal of the batch (work list; inside a tile ordered by fragment, then slot, so the
// lanes of a warp mostly hold the same fragment index of neighbouring slots).  A variant that staged the fragment cubes in
// shared memory with cp.async.bulk + mbarrier was measured slower and removed (profiles/r2_dp_fragment_staged_variant.cu.txt).
__global__ void __launch_bounds__(DP_THREADS, DP_LB_FRAGMENT) dp_fragment_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  if (t >= (uint32_t)*P.n_work) return;
  uint32_t j, k;
  dp_decode(P.work[t], (uint32_t)P.KS, j, k);
  dp_fragment(P, j, (int)k);
}

__global__ void __launch_bounds__(DP_THREADS) dp_median_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  uint32_t j, lane;
  dp_decode(t, DP_MED_LANES, j, lane);
  if (j < P.n) dp_median(P, j, (int)lane);
}

__global__ void __launch_bounds__(DP_THREADS) dp_corr_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  if (t >= (uint32_t)*P.n_work) return;
  uint32_t j, k;
  dp_decode(P.work[t], (uint32_t)P.KS, j, k);
  dp_corr(P, j, (int)k);
}

// thread <-> slot; the finished feature rows of a warp are staged in shared memory and written row by row with the 32 lanes
// on consecutive floats (46 scattered 4-byte stores per thread were a quarter of this kernel's stall samples)
constexpr int DPA_THREADS = 128;
__global__ void __launch_bounds__(DPA_THREADS) dp_aggregate_kernel(const __grid_constant__ DpParams P) {
  __shared__ float rows[DPA_THREADS / 32][32][ADB_NUM_FEATURES + 1];
  const int64_t j = (int64_t)blockIdx.x * DPA_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool scored = false;
  if (j < P.n) {
    dp_aggregate(P, j, rows[warp][lane]);
    scored = P.state[j] == 2;
  }
  const long long ci = scored ? (long long)dp_candidate_of(P, j) : -1;
  __syncwarp();  // the staged rows are visible to the whole warp
  unsigned m = __ballot_sync(0xFFFFFFFFu, scored);
  while (m) {
    const int r = __ffs(m) - 1;
    m &= m - 1;
    const long long row = __shfl_sync(0xFFFFFFFFu, ci, r);
    float* dst = P.out.features + (size_t)row * ADB_NUM_FEATURES;
    dst[lane] = rows[warp][r][lane];
    if (lane + 32 < ADB_NUM_FEATURES) dst[lane + 32] = rows[warp][r][lane + 32];
  }
}

__global__ void __launch_bounds__(DP_THREADS) dp_write_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  uint32_t j, w;
  dp_decode(t, (uint32_t)P.KS, j, w);
  if (j < P.n) dp_write(P, j, (int)w);
}

inline unsigned blocks_for(int64_t threads) { return (unsigned)((threads + DP_THREADS - 1) / DP_THREADS); }

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

// bytes of the per-batch plan arrays (everything except the cube)
size_t adb_score_dp_plan_bytes(int64_t nb, int KS, int nIcap, size_t* scan_tmp_bytes) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(nb + 1));
  size_t tmp2 = 0;
  cub::DeviceSelect::Flagged(nullptr, tmp2, cub::CountingInputIterator<uint32_t>(0), (const uint8_t*)nullptr, (uint32_t*)nullptr,
                             (int32_t*)nullptr, (int)(dp_padded_slots(nb) * KS));
  tmp = std::max(tmp, tmp2);
  if (scan_tmp_bytes) *scan_tmp_bytes = tmp;
  const size_t N = (size_t)dp_padded_slots(nb);
  return align256(8 * (N + 1)) + align256(N * (size_t)KS) + align256(4 * N * (size_t)KS) + 256 + align256(16 * DP_WTAB_P_STRIDE) + align256(N) * 3 + align256(4 * N) * 2 + align256(2 * N * ADB_MAX_OBS) + align256(4 * N * (size_t)KS) +
         align256(8 * N * (size_t)nIcap * ADB_MAX_OBS) + align256(4 * N * ADB_MAX_OBS) + align256(8 * (N + 1)) * 2 + align256(tmp) + 256;
}

// Scores candidates order[0 .. cand_n) in batches of `batch` slots.  plan: adb_score_dp_plan_bytes(batch, ...) bytes;
// cube: grow-only float workspace (*cube / *cube_floats, reallocated through `grow` when a batch needs more).
// Returns 0, or 1 when the workspace could not be grown (message through `grow`'s owner).
int adb_launch_score_dp(const DevRaw& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand, DevScoresOut out,
                        int out_k, int KS, const int32_t* d_order, int64_t batch, void* plan, float** cube, size_t* cube_floats,
                        int (*grow)(void* owner, size_t floats), void* owner, uint32_t* d_status, cudaStream_t stream,
                        int* n_launches) {
  if (cand.n <= 0) return 0;
  DpParams P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.cand = cand; P.out = out;
  P.out_k = out_k;
  P.order = d_order;
  P.KS = KS;
  P.nIcap = (int)std::min<int64_t>(std::min<int64_t>(lib.n_isotopes, cfg.top_k_isotopes), ADB_MAX_ISOTOPES);
  P.status = d_status;
  size_t tmp = 0;
  adb_score_dp_plan_bytes(batch, KS, P.nIcap, &tmp);
  const size_t N = (size_t)dp_padded_slots(batch);
  char* p = (char*)plan;
  auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
  P.state = (uint8_t*)take(N);
  P.F = (uint8_t*)take(N);
  P.nobs = (uint8_t*)take(N);
  P.C = (int32_t*)take(4 * N);
  P.cs = (int32_t*)take(4 * N);
  P.pos = (uint16_t*)take(2 * N * ADB_MAX_OBS);
  P.fsel = (uint32_t*)take(4 * N * (size_t)KS);
  P.qtf = (double*)take(8 * N * (size_t)P.nIcap * ADB_MAX_OBS);
  P.qmask = (float*)take(4 * N * ADB_MAX_OBS);
  P.need = (int64_t*)take(8 * (N + 1));
  P.tile_need = (int64_t*)take(8 * (N + 1));
  int64_t* off = (int64_t*)take(8 * (N + 1));
  P.off = off;
  void* scan_tmp = take(tmp);
  P.rowflag = (uint8_t*)take(N * (size_t)KS);
  P.work = (uint32_t*)take(4 * N * (size_t)KS);
  P.n_work = (int32_t*)take(4);
  double* wtab_p = (double*)take(16 * DP_WTAB_P_STRIDE);
  P.wtab_p = wtab_p;
  dp_wtab_p_kernel<<<blocks_for(2 * DP_WTAB_P_STRIDE), DP_THREADS, 0, stream>>>(wtab_p);
  if (n_launches) *n_launches += 1;
  for (int64_t base = 0; base < cand.n; base += batch) {
    P.base = base;
    P.n = std::min<int64_t>(batch, cand.n - base);
    P.cube = *cube;
    const int64_t n_tiles = (P.n + DP_W - 1) / DP_W, n_pad = n_tiles * DP_W;
    dp_setup_kernel<<<blocks_for(P.n), DP_THREADS, 0, stream>>>(P);
    dp_tile_need_kernel<<<blocks_for(n_tiles + 1), DP_THREADS, 0, stream>>>(P);
    cub::DeviceScan::ExclusiveSum(scan_tmp, tmp, (const int64_t*)P.tile_need, off, (int)(n_tiles + 1), stream);
    int64_t total = 0;
    if (cudaMemcpyAsync(&total, off + n_tiles, sizeof(total), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;
    if ((size_t)total > *cube_floats) {
      if (grow(owner, (size_t)total)) return 1;
    }
    P.cube = *cube;
    dp_extract_kernel<<<blocks_for(n_pad * (P.KS + P.nIcap)), DP_THREADS, 0, stream>>>(P);
    {
      size_t tb = tmp;
      cub::DeviceSelect::Flagged(scan_tmp, tb, cub::CountingInputIterator<uint32_t>(0), (const uint8_t*)P.rowflag, P.work, P.n_work,
                                 (int)(n_pad * P.KS), stream);
    }
    dp_template_kernel<<<blocks_for(P.n), DP_THREADS, 0, stream>>>(P);
    dp_fragment_kernel<<<blocks_for(n_pad * P.KS), DP_THREADS, 0, stream>>>(P);
    if (cfg.experimental_xic) dp_median_kernel<<<blocks_for(n_pad * DP_MED_LANES), DP_THREADS, 0, stream>>>(P);
    dp_corr_kernel<<<blocks_for(n_pad * P.KS), DP_THREADS, 0, stream>>>(P);
    dp_aggregate_kernel<<<(unsigned)((P.n + DPA_THREADS - 1) / DPA_THREADS), DPA_THREADS, 0, stream>>>(P);
    if (cfg.collect_fragments) dp_write_kernel<<<blocks_for(n_pad * P.KS), DP_THREADS, 0, stream>>>(P);
    if (n_launches) *n_launches += 10 + (cfg.experimental_xic ? 1 : 0) + (cfg.collect_fragments ? 1 : 0);
  }
  return 0;
}

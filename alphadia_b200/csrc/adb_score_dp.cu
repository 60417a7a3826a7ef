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
  if (j == P.n) P.need[j] = 0;  // the scan runs over n + 1 entries: off[n] = total
}

// thread t <-> (slot, row): rows 0 .. KS-1 are the fragment rows, KS .. KS+nIcap-1 the isotope rows
__global__ void __launch_bounds__(DP_THREADS, DP_LB_EXTRACT) dp_extract_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;  // a batch has < 2^32 rows: 32-bit index arithmetic
  const uint32_t rows = (uint32_t)(P.KS + P.nIcap);
  const uint32_t j = t / rows;
  if (j < P.n) dp_extract(P, j, (int)(t - j * rows));
}

__global__ void __launch_bounds__(DP_THREADS, DP_LB_TEMPLATE) dp_template_kernel(const __grid_constant__ DpParams P) {
  const int64_t j = (int64_t)blockIdx.x * DP_THREADS + threadIdx.x;
  if (j < P.n) dp_template(P, j);
}

// ---- dp_fragment: work list + bulk-asynchronous staging of the fragment cubes ------------------------------------------
// thread t <-> t-th fragment row with signal of the batch (work list, ascending slot).  The rows of one CTA belong to a run
// of consecutive slots; the first row of every slot ("head") fetches that candidate's fragment cube [dfi | dfm] - one
// contiguous, 16-byte aligned piece of the batch workspace whose address is known before any arithmetic - into the CTA's
// shared-memory arena with ONE cp.async.bulk (TMA engine, completion counted on an mbarrier); the pass then reads its
// intensity / m/z rows at shared-memory latency.  Cubes that do not fit the arena are read in place.
// Measured on config 3 (r2, gpurun_out/w_*.json): staging ON with 128 threads x 36 KB x 6 CTAs/SM: scoring 58.6 ms, 128 x 26 KB
// x 8: 57.9 ms; staging OFF with 256 threads x 4 CTAs/SM: 53.2 ms.  The pass still chases its other arrays (best profile,
// template, weight tables) through L1/L2, the arena costs a quarter of the resident threads, and the cube rows are re-read
// from L1 anyway - so the default build reads in place; -DDPF_STAGE=1 -DDPF_THREADS_N=128 -DDPF_ARENA_KB=36 -DDPF_MIN_BLOCKS=6
// compiles the staged variant (SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64).
#ifndef DPF_STAGE
#define DPF_STAGE 0
#endif
#ifndef DPF_THREADS_N
#define DPF_THREADS_N 256
#endif
#ifndef DPF_ARENA_KB
#define DPF_ARENA_KB 0
#endif
#ifndef DPF_MIN_BLOCKS
#define DPF_MIN_BLOCKS 4
#endif
constexpr int DPF_THREADS = DPF_THREADS_N;
constexpr int DPF_ARENA_BYTES = DPF_ARENA_KB * 1024;

__device__ __forceinline__ uint32_t dp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(DPF_THREADS, DPF_MIN_BLOCKS) dp_fragment_kernel(const __grid_constant__ DpParams P) {
  extern __shared__ __align__(128) unsigned char dpf_arena[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int s_off[DPF_THREADS];  // arena byte offset of the i-th slot run of this CTA, -1: read in place
  __shared__ int s_warp_bytes[DPF_THREADS / 32], s_warp_heads[DPF_THREADS / 32];
  const uint32_t t = blockIdx.x * DPF_THREADS + threadIdx.x;
  const uint32_t n_work = (uint32_t)*P.n_work;
  if (blockIdx.x * DPF_THREADS >= n_work) return;  // whole CTA idle
  const bool active = t < n_work;
  const uint32_t w = active ? P.work[t] : 0u;
  const uint32_t j = w / (uint32_t)P.KS;
  const int k = (int)(w - j * (uint32_t)P.KS);
  const bool head = active && (threadIdx.x == 0 || P.work[t - 1] / (uint32_t)P.KS != j);
  int bytes = 0;
  if (head) bytes = (2 * (int)P.F[j] * (int)P.nobs[j] * P.C[j] * 4 + 15) & ~15;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dp_smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // CTA-wide exclusive scans of (bytes, head flags): warp shuffles + one pass over the warp totals
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc_b = bytes, inc_h = head ? 1 : 0;
  for (int d = 1; d < 32; d <<= 1) {
    const int ub = __shfl_up_sync(0xFFFFFFFFu, inc_b, d), uh = __shfl_up_sync(0xFFFFFFFFu, inc_h, d);
    if (lane >= d) { inc_b += ub; inc_h += uh; }
  }
  if (lane == 31) { s_warp_bytes[wid] = inc_b; s_warp_heads[wid] = inc_h; }
  __syncthreads();
  int base_b = 0, base_h = 0;
  for (int q = 0; q < wid; q++) { base_b += s_warp_bytes[q]; base_h += s_warp_heads[q]; }
  const int off = base_b + inc_b - bytes;        // exclusive
  const int run = base_h + inc_h - 1;            // ordinal of this thread's slot run inside the CTA
  const bool fits = DPF_STAGE && head && off + bytes <= DPF_ARENA_BYTES && bytes > 0;
  if (head) s_off[run] = fits ? off : -1;
  // total bytes in flight = those of the fitting heads (a prefix of the runs: offsets only grow)
  int staged_bytes = fits ? bytes : 0;
  for (int d = 16; d > 0; d >>= 1) staged_bytes += __shfl_xor_sync(0xFFFFFFFFu, staged_bytes, d);
  if (lane == 0) s_warp_bytes[wid] = staged_bytes;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int q = 0; q < DPF_THREADS / 32; q++) total += s_warp_bytes[q];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dp_smem_u32(&mbar)), "r"(total) : "memory");
  }
  if (fits) {
    const float* src = P.cube + P.off[j];  // the block starts with [dfi | dfm]
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dp_smem_u32(dpf_arena + off)),
                 "l"(src), "r"(bytes), "r"(dp_smem_u32(&mbar))
                 : "memory");
  }
  {  // every thread waits for phase 0 of the barrier: all staged cubes have landed
    const uint32_t bar = dp_smem_u32(&mbar);
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(0) : "memory");
    }
  }
  if (!active) return;
  const int so = s_off[run];
  dp_fragment(P, j, k, so >= 0 ? (const float*)(dpf_arena + so) : nullptr);
}

__global__ void __launch_bounds__(DP_THREADS) dp_median_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  const uint32_t j = t / DP_MED_LANES;
  if (j < P.n) dp_median(P, j, (int)(t % DP_MED_LANES));
}

__global__ void __launch_bounds__(DP_THREADS) dp_corr_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  if (t >= (uint32_t)*P.n_work) return;
  const uint32_t w = P.work[t];
  const uint32_t j = w / (uint32_t)P.KS;
  dp_corr(P, j, (int)(w - j * (uint32_t)P.KS));
}

__global__ void __launch_bounds__(DP_THREADS) dp_aggregate_kernel(const __grid_constant__ DpParams P) {
  const int64_t j = (int64_t)blockIdx.x * DP_THREADS + threadIdx.x;
  if (j < P.n) dp_aggregate(P, j);
}

__global__ void __launch_bounds__(DP_THREADS) dp_write_kernel(const __grid_constant__ DpParams P) {
  const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
  const uint32_t j = t / (uint32_t)P.KS;
  if (j < P.n) dp_write(P, j, (int)(t - j * (uint32_t)P.KS));
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
                             (int32_t*)nullptr, (int)(nb * KS));
  tmp = std::max(tmp, tmp2);
  if (scan_tmp_bytes) *scan_tmp_bytes = tmp;
  const size_t N = (size_t)nb;
  return align256(N * (size_t)KS) + align256(4 * N * (size_t)KS) + 256 + align256(16 * DP_WTAB_P_STRIDE) + align256(N) * 3 + align256(4 * N) * 2 + align256(2 * N * ADB_MAX_OBS) + align256(4 * N * (size_t)KS) +
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
  const size_t N = (size_t)batch;
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
    dp_setup_kernel<<<blocks_for(P.n + 1), DP_THREADS, 0, stream>>>(P);
    cub::DeviceScan::ExclusiveSum(scan_tmp, tmp, (const int64_t*)P.need, off, (int)(P.n + 1), stream);
    int64_t total = 0;
    if (cudaMemcpyAsync(&total, off + P.n, sizeof(total), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;
    if ((size_t)total > *cube_floats) {
      if (grow(owner, (size_t)total)) return 1;
    }
    P.cube = *cube;
    dp_extract_kernel<<<blocks_for(P.n * (P.KS + P.nIcap)), DP_THREADS, 0, stream>>>(P);
    {
      size_t tb = tmp;
      cub::DeviceSelect::Flagged(scan_tmp, tb, cub::CountingInputIterator<uint32_t>(0), (const uint8_t*)P.rowflag, P.work, P.n_work,
                                 (int)(P.n * P.KS), stream);
    }
    dp_template_kernel<<<blocks_for(P.n), DP_THREADS, 0, stream>>>(P);
    dp_fragment_kernel<<<(unsigned)((P.n * P.KS + DPF_THREADS - 1) / DPF_THREADS), DPF_THREADS, DPF_ARENA_BYTES, stream>>>(P);
    if (cfg.experimental_xic) dp_median_kernel<<<blocks_for(P.n * DP_MED_LANES), DP_THREADS, 0, stream>>>(P);
    dp_corr_kernel<<<blocks_for(P.n * P.KS), DP_THREADS, 0, stream>>>(P);
    dp_aggregate_kernel<<<blocks_for(P.n), DP_THREADS, 0, stream>>>(P);
    if (cfg.collect_fragments) dp_write_kernel<<<blocks_for(P.n * P.KS), DP_THREADS, 0, stream>>>(P);
    if (n_launches) *n_launches += 9 + (cfg.experimental_xic ? 1 : 0) + (cfg.collect_fragments ? 1 : 0);
  }
  return 0;
}

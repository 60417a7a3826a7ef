// alphadia_b200 — candidate scoring kernel (3-D raw files), sm_100a.
//
// Replaces Candidate.process (alphadia/search/scoring/containers/candidate.py:166-481) and everything it
// calls: AlphaRawJIT.get_dense(absolute_masses=True) (jitclasses/alpharaw_jit.py:208-337), the quadrupole
// transfer function / template / observation importance (scoring/quadrupole.py:80-115,261-335), profiles
// (scoring/utils.py:26-66) and the 46 features (scoring/features/*.py).
//
// Mapping: ONE TILE PER CANDIDATE — a 16-lane half-warp when top_k_fragments <= 16 (two candidates per warp, so
// the fragment-per-lane phases keep 12 of 16 lanes busy), a full warp otherwise.  Candidates are visited in
// (quad window, frame_start) order so neighbouring tiles read the same spectra (L2/L1 reuse).
//   * extraction: lanes stride over (spectrum, fragment) items with the fragment index fastest, so the 12
//     lanes probing one spectrum share the first levels of their binary searches (same sectors);
//     the per-cell m/z recurrence runs in ascending peak order inside one lane (order-dependent f32).
//   * the dense cube (intensity + m/z channel), template and profiles live in the warp's shared-memory
//     scratch; a 3-D file's two scan rows are identical (alpharaw_jit.py:326-333), one row is stored and
//     the scan sums are formed as x + x (exact).
//   * per-fragment features: lane w <-> fragment w, loops in the reference's order (sequential f32/f64
//     accumulation, no FMA contraction) so results track the CPU path to the last bits;
//     cross-fragment statistics are short sequential loops over shared memory, executed by all lanes.
//   * candidates whose cube does not fit the shared-memory scratch fall back to a per-warp HBM workspace
//     through the same generic pointer.
#include <cooperative_groups.h>

#include "adb_common.cuh"

namespace cg = cooperative_groups;

#ifndef SCORE_THREADS
#define SCORE_THREADS 128  // threads per CTA; the tiles of a CTA move through the code in lock step (see adb_score_kernel)
#endif
#ifndef SCORE_CTAS_PER_SM
#define SCORE_CTAS_PER_SM (512 / SCORE_THREADS)
#endif
#define SCORE_PHASES 3     // CTA barriers inside one candidate round
#ifndef SMEM_FLOATS_PER_TILE
#define SMEM_FLOATS_PER_TILE 1024  // 4 KB dynamic scratch per candidate tile; larger cubes use the HBM workspace
#endif

namespace {

template <int TILE>
struct TileSmall {
  // selected fragments, m/z sorted (FragmentContainer, fragment_container.py:12-45)
  float mz_library[TILE], mz[TILE], intensity[TILE];
  float lo[TILE], hi[TILE];
  uint8_t type[TILE], loss_type[TILE], charge[TILE], number[TILE], position[TILE];
  int fmap[TILE];        // masked fragment w -> selected fragment f
  int sorted_idx[TILE];  // np.argsort(intensity)[::-1] over masked fragments
  int frame_peak[TILE];
  float fint[TILE];      // fragments.intensity after apply_mask (sum 1)
  float fin[TILE];       // fragment_intensity_norm
  float ofi[TILE];       // observed_fragment_intensity
  float cosv[TILE];
  float corr_list[TILE];
  float rfw[TILE];
  double area_norm[TILE], ofh_mean[TILE], mass_error[TILE], ci[TILE];
  double qtf[ADB_MAX_ISOTOPES * ADB_MAX_OBS];
  double esc[ADB_MAX_OBS], efc[ADB_MAX_OBS];
  double H[ADB_MAX_ISOTOPES], MZo[ADB_MAX_ISOTOPES];
  float oi[ADB_MAX_OBS], sti[ADB_MAX_OBS], qmask[ADB_MAX_OBS];
  float iso_mz[ADB_MAX_ISOTOPES], iso_int[ADB_MAX_ISOTOPES], lo_p[ADB_MAX_ISOTOPES], hi_p[ADB_MAX_ISOTOPES];
  float spi[ADB_MAX_ISOTOPES], wspi[ADB_MAX_ISOTOPES];
  int pos[ADB_MAX_OBS];
  float feat[ADB_NUM_FEATURES];
  int t_sel[TILE];
};

struct ScoreParams {
  DevRaw raw;
  DevLib lib;
  adb_scoring_config cfg;
  DevCandidatesIn cand;
  DevScoresOut out;
  float* workspace;
  long long ws_floats_per_tile;
  uint32_t* status;
  const int32_t* order;  // processing order of the candidates (locality), may be null
};

// one (spectrum, query window) cell of get_dense(absolute_masses=True): alpharaw_jit.py:290-335.
// prev_hi: upper bound of the previous (lower m/z) window when it overlaps this one, else -1.
__device__ __forceinline__ void extract_cell(const DevRaw& raw, int64_t scan, float lo, float hi, float prev_hi,
                                             float& acc_i, float& acc_m) {
  const AdbFound f = adb_spectrum_lower_bound(raw, scan, lo);
  // common case: the first candidate peak is known (still in registers) and lies above the window -> no hit
  if (f.inside && !(f.mz_at_idx <= hi) && !(prev_hi >= lo)) return;
  const uint32_t stop = adb_spectrum_stop(raw, scan);
  uint32_t idx = f.idx;
  if (prev_hi >= lo)  // the search cursor only moves forward: peaks taken by the previous window are gone
    while (idx < stop && __ldg(raw.mz + idx) <= prev_hi) idx++;
  while (idx < stop) {
    float nm = __ldg(raw.mz + idx);
    if (!(nm <= hi)) break;
    float ni = __ldg(raw.intensity + idx);
    ni = __fmul_rn(ni, ((double)ni > 1e-26) ? 1.0f : 0.0f);
    float num32 = __fadd_rn(__fmul_rn(acc_m, acc_i), __fmul_rn(ni, nm));
    float den32 = __fadd_rn(acc_i, ni);
    double nd = __ddiv_rn(__dadd_rn((double)num32, 1e-36), __dadd_rn((double)den32, 1e-36));
    acc_i = den32;
    acc_m = (float)nd;
    idx++;
  }
}

__device__ __forceinline__ float twice(float x) { return __fadd_rn(x, x); }  // sum over the 2 identical scan rows

// np.corrcoef(x, y)[0, 1] as numba evaluates it (cov with 1/(n-1), divide by both std)
__device__ __noinline__ double corrcoef01(const double* x, const float* yf, int n) {
  double mx = 0, my = 0;
  for (int i = 0; i < n; i++) { mx = __dadd_rn(mx, x[i]); my = __dadd_rn(my, (double)yf[i]); }
  mx /= n; my /= n;
  double cxx = 0, cyy = 0, cxy = 0;
  for (int i = 0; i < n; i++) {
    double a = x[i] - mx, b = (double)yf[i] - my;
    cxx = __dadd_rn(cxx, __dmul_rn(a, a)); cyy = __dadd_rn(cyy, __dmul_rn(b, b)); cxy = __dadd_rn(cxy, __dmul_rn(a, b));
  }
  double fact = 1.0 / (double)(n - 1);
  cxx *= fact; cyy *= fact; cxy *= fact;
  return (cxy / sqrt(cyy)) / sqrt(cxx);
}

// features_utils.py:9-26 weighted_center_mean of the intensity row r and the m/z row rm of one (fragment,
// observation) cell over the two identical scan rows, with the tabulated distance weights wt[2][C]
__device__ __noinline__ void weighted_center_mean_pair(const float* r, const float* rm, const double* wt, int C,
                                                       double& h, double& mz) {
  double v1 = 0, w1 = 0, v2 = 0, w2 = 0;
  bool any1 = false, any2 = false;
  // branch-free: a cell that is not > 0 adds +0.0, which leaves the (non-negative) running sums bit-identical
#pragma unroll 1
  for (int s = 0; s < 2; s++)
#pragma unroll 4
    for (int c = 0; c < C; c++) {
      const double wgt = wt[s * C + c];
      const float a = r[c], b = rm[c];
      const bool pa = a > 0.f, pb = b > 0.f;
      any1 |= pa; any2 |= pb;
      v1 = __dadd_rn(v1, pa ? __dmul_rn((double)a, wgt) : 0.0); w1 = __dadd_rn(w1, pa ? wgt : 0.0);
      v2 = __dadd_rn(v2, pb ? __dmul_rn((double)b, wgt) : 0.0); w2 = __dadd_rn(w2, pb ? wgt : 0.0);
    }
  h = (any1 && w1 > 0) ? v1 / w1 : 0.0;
  mz = (any2 && w2 > 0) ? v2 / w2 : 0.0;
}

// sum over observations of fragments_frame_profile[f, :, c]; the best observation's row lives in bp when the
// centre envelope mutated it (fragment_features.py:248-250)
__device__ __noinline__ float frame_profile_obs_sum(const float* dfi_f, const float* bp_w, int nobs, int C, int c, int mutated_obs) {
  float t = 0.f;
#pragma unroll 1
  for (int o = 0; o < nobs; o++) t = __fadd_rn(t, (o == mutated_obs) ? bp_w[c] : __fadd_rn(dfi_f[o * C + c], dfi_f[o * C + c]));
  return t;
}

// CTA-wide barrier that tiles of one warp may reach at different times (non-.aligned form)
__device__ __forceinline__ void cta_phase_barrier() { asm volatile("barrier.sync 1;" ::: "memory"); }
#define PHASE_BARRIER() do { cta_phase_barrier(); phase++; } while (0)

template <int TILE>
__device__ void score_one_body(const ScoreParams& P, int64_t ci, const cg::thread_block_tile<TILE>& tile, TileSmall<TILE>& sm,
                               float* smem_scratch, float* ws_scratch, int& phase) {
  const int lane = (int)tile.thread_rank();
  const DevRaw& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int K = (int)cfg.top_k_fragments;
  const int64_t L = raw.cycle_len;

  const int64_t p = P.cand.lib_row[ci];
  const int64_t frame_start = P.cand.frame_start[ci], frame_stop = P.cand.frame_stop[ci], frame_center = P.cand.frame_center[ci];
  const int64_t scan_start = P.cand.scan_start[ci], scan_stop = P.cand.scan_stop[ci], scan_center = P.cand.scan_center[ci];

  // ---- candidate.py:151-163 isotope m/z ------------------------------------------------------
  const int nI = min(min(lib.n_isotopes, (int)cfg.top_k_isotopes), ADB_MAX_ISOTOPES);
  const double charge = (double)lib.charge[p];
  const float pmz = lib.mz[p];
  if (lane < nI) {
    sm.iso_mz[lane] = __fadd_rn((float)((double)lane * ADB_ISOTOPE_DIFF / charge), pmz);
    sm.iso_int[lane] = lib.isotopes[p * lib.n_isotopes + lane];
  }

  // ---- candidate.py:181-192 fragments: cardinality filter, top-k by intensity, sort by m/z ----
  const int64_t fs = lib.frag_start_idx[p], fe = lib.frag_stop_idx[p];
  int n_all = (int)(fe - fs);
  if (n_all < 0) n_all = 0;
  if (n_all > ADB_MAX_LIB_FRAGMENTS) {
    if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_LIB_FRAGMENTS);
    return;
  }
  // staging of the library fragments (before top-k) in the tile's scratch
  float* t_int = smem_scratch;
  float* t_mz = smem_scratch + ADB_MAX_LIB_FRAGMENTS;
  int* t_src = (int*)(smem_scratch + 2 * ADB_MAX_LIB_FRAGMENTS);
  int m = 0;
  _Pragma("unroll 1") for (int base = 0; base < n_all; base += TILE) {
    int j = base + lane;
    bool keep = j < n_all && (!cfg.exclude_shared_ions || lib.frag_cardinality[fs + j] <= 1);
    unsigned b = tile.ballot(keep);
    if (keep) {
      int u = m + __popc(b & ((1u << lane) - 1u));
      t_src[u] = j;
      t_int[u] = lib.frag_intensity[fs + j];
      t_mz[u] = lib.frag_mz[fs + j];
    }
    m += __popc(b);
  }
  tile.sync();
  const int F0 = min(min(m, K), TILE);
  _Pragma("unroll 1") for (int u = lane; u < m; u += TILE) {  // descending-intensity position (stable argsort reversed)
    float v = t_int[u];
    int rank_asc = 0;
    _Pragma("unroll 1") for (int q = 0; q < m; q++) rank_asc += (t_int[q] < v) || (t_int[q] == v && q < u);
    int r = m - 1 - rank_asc;
    if (r < F0) sm.t_sel[r] = u;
  }
  tile.sync();
  _Pragma("unroll 1") for (int r = lane; r < F0; r += TILE) {  // stable m/z order among the selected
    int u = sm.t_sel[r];
    float v = t_mz[u];
    int rank2 = 0;
    _Pragma("unroll 1") for (int q = 0; q < F0; q++) { float vq = t_mz[sm.t_sel[q]]; rank2 += (vq < v) || (vq == v && q < r); }
    int64_t g = fs + t_src[u];
    sm.mz_library[rank2] = lib.frag_mz_library[g];
    sm.mz[rank2] = v;
    sm.intensity[rank2] = t_int[u];
    sm.type[rank2] = lib.frag_type[g];
    sm.loss_type[rank2] = lib.frag_loss_type[g];
    sm.charge[rank2] = lib.frag_charge[g];
    sm.number[rank2] = lib.frag_number[g];
    sm.position[rank2] = lib.frag_position[g];
  }
  tile.sync();
  const int F = F0;
  if (F <= 3) return;

  // ---- windows (jitclasses/utils.py:15-20 with a float32 tolerance) + quadrupole limits --------
  if (lane < F) {
    float mz = sm.mz[lane];
    double d = (double)__fmul_rn(cfg.fragment_mz_tolerance, mz) / 1000000.0;
    sm.lo[lane] = (float)((double)mz - d);
    sm.hi[lane] = (float)((double)mz + d);
  }
  if (lane < nI) {
    float mz = sm.iso_mz[lane];
    double d = (double)__fmul_rn(cfg.precursor_mz_tolerance, mz) / 1000000.0;
    sm.lo_p[lane] = (float)((double)mz - d);
    sm.hi_p[lane] = (float)((double)mz + d);
  }
  tile.sync();
  float mn = sm.iso_mz[0], mx = sm.iso_mz[0];
  _Pragma("unroll 1") for (int i = 1; i < nI; i++) { mn = fminf(mn, sm.iso_mz[i]); mx = fmaxf(mx, sm.iso_mz[i]); }
  const float q0 = (float)((double)mn - 0.5), q1 = (float)((double)mx + 0.5);  // candidate.py:203-205

  int nobs = 0;  // alpharaw_jit.py:19-50
  _Pragma("unroll 1") for (int64_t base = 0; base < L; base += TILE) {
    int64_t j = base + lane;
    bool hit = j < L && ((double)q0 <= raw.cycle[2 * j + 1]) && ((double)q1 >= raw.cycle[2 * j]);
    unsigned b = tile.ballot(hit);
    if (hit) {
      int u = nobs + __popc(b & ((1u << lane) - 1u));
      if (u < ADB_MAX_OBS) sm.pos[u] = (int)j;
    }
    nobs += __popc(b);
  }
  if (nobs > ADB_MAX_OBS) {
    if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_OBS);
    return;
  }
  const int64_t cs = frame_start / L;
  const int64_t C64 = frame_stop / L - cs;
  if (C64 <= 0 || nobs == 0) return;  // candidate.py:230-232, :323-325
  // 3-D files: np.arange(scan_start, scan_stop) indexes cycle[0, c, s]; only s == 0 exists
  if (scan_stop - scan_start != 1 || scan_start != 0) return;
  if (scan_center < 0 || scan_center >= raw.n_mobility || frame_stop < 1 || frame_stop > raw.n_spectra ||
      frame_center < 0 || frame_center >= raw.n_spectra || frame_start < 0)
    return;
  if ((cs + C64) * L > raw.n_spectra) return;
  const int C = (int)C64;

  // ---- scratch carve-up ----------------------------------------------------------------------
  const int nFC = F * nobs * C;  // <= 32 * 8 * 4096: 32-bit index arithmetic in the hot loops
  long long need = 2LL * nFC + 1LL * F * C + 2LL * nI * C + 2LL * nobs * C + C + 2 /*align*/ + 4LL * nobs * C + 4LL * C;
  if (!cfg.experimental_xic) need += (long long)F * F;
  float* scratch = smem_scratch;
  if (need > SMEM_FLOATS_PER_TILE) {
    if (ws_scratch == nullptr || need > P.ws_floats_per_tile) {
      if (lane == 0) atomicOr(P.status, ADB_STATUS_SCRATCH_OVERFLOW);
      return;
    }
    scratch = ws_scratch;
  }
  float* dfi = scratch;               // [F][nobs][C] intensity (one scan row)
  float* dfm = dfi + nFC;             // [F][nobs][C] m/z channel; dead after the fragment features ...
  float* nrm = dfm;                   // ... then reused: [F][C] normalised / centred profiles
  float* bp = dfm + nFC;              // [F][C] best profile (enveloped)
  float* dpi = bp + (long long)F * C; // [I][C]
  float* dpm = dpi + (long long)nI * C;
  float* tmpl = dpm + (long long)nI * C;  // [nobs][C]
  float* tfp = tmpl + (long long)nobs * C;// [nobs][C] template frame profile
  float* med = tfp + (long long)nobs * C; // [C]
  float* after = med + C;
  after += ((uintptr_t)after & 7u) ? 1 : 0;  // 8-byte align
  double* wtab = (double*)after;             // [nobs][2][C] exp(-0.1 dist) tables for fragments
  double* wtab_p = wtab + 2LL * nobs * C;    // [2][C] for precursors
  float* red = (float*)(wtab_p + 2LL * C);   // [F][F] legacy correlation accumulator
  tile.sync();  // the staging area in the scratch is dead from here on

  // ---- candidate.py:216-223 fragment cube -------------------------------------------------------
  _Pragma("unroll 1") for (int t = lane; t < nFC; t += TILE) {
    const int k = t % F, oc = t / F;
    const int c = oc % C, o = oc / C;
    int64_t scan = (int64_t)sm.pos[o] + (cs + c) * L;
    float ai = 0.f, am = 0.f;
    float prev_hi = (k > 0) ? sm.hi[k - 1] : -1.0f;
    extract_cell(raw, scan, sm.lo[k], sm.hi[k], prev_hi, ai, am);
    const int cell = (k * nobs + o) * C + c;
    dfi[cell] = ai;
    dfm[cell] = am;
  }
  // ---- candidate.py:239-269 MS1 cube with the observation collapse --------------------------------
  _Pragma("unroll 1") for (int t = lane; t < nI * C; t += TILE) {
    int i = t % nI, c = t / nI;
    float s32 = 0.f;
    double smz = 0.0;
    int count = 0;
    float prev_hi = (i > 0) ? sm.hi_p[i - 1] : -1.0f;
    _Pragma("unroll 1") for (int j = 0; j < raw.n_ms1_pos; j++) {
      int64_t scan = (int64_t)raw.ms1_pos[j] + (cs + c) * L;
      float ai = 0.f, am = 0.f;
      extract_cell(raw, scan, sm.lo_p[i], sm.hi_p[i], prev_hi, ai, am);
      s32 = __fadd_rn(s32, ai);
      smz = __dadd_rn(smz, (double)am);
      count += am > 0.f;
    }
    dpi[i * C + c] = s32;
    dpm[i * C + c] = (float)(smz / ((double)count + 1e-6));
  }

  PHASE_BARRIER();  // ======== end of code region A (setup + extraction) ========
  // ---- quadrupole.py:80-115,261-301 transfer function (n_scans == 1) --------------------------------
  _Pragma("unroll 1") for (int t = lane; t < nI * nobs; t += TILE) {
    int i = t / nobs, o = t % nobs;
    double mu1 = raw.cycle[2 * sm.pos[o] + 0] + cfg.quad_delta_mu[0];
    double mu2 = raw.cycle[2 * sm.pos[o] + 1] + cfg.quad_delta_mu[1];
    double x = (double)sm.iso_mz[i];
    double a1 = (x - mu1) / cfg.quad_sigma[0], a2 = (x - mu2) / cfg.quad_sigma[1];
    sm.qtf[i * nobs + o] = 1.0 / (1.0 + exp(-a1)) - 1.0 / (1.0 + exp(-a2));
  }
  tile.sync();
  if (lane < nobs) {  // candidate.py:287-289
    double s = 0;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++) s = __dadd_rn(s, sm.qtf[i * nobs + lane]);
    sm.qmask[lane] = (float)(s / (double)nI);
  }
  tile.sync();
  _Pragma("unroll 1") for (int t = lane; t < nFC; t += TILE) {  // candidate.py:290
    const int o = (t / C) % nobs;
    dfi[t] = __fmul_rn(dfi[t], sm.qmask[o]);
  }
  _Pragma("unroll 1") for (int t = lane; t < nobs * C; t += TILE) {  // quadrupole.py:304-324 template
    int o = t / C, c = t % C;
    double acc = 0;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++)
      acc = __dadd_rn(acc, __dmul_rn((double)__fmul_rn(dpi[i * C + c], sm.iso_int[i]), sm.qtf[i * nobs + o]));
    tmpl[t] = (float)acc;
  }
  tile.sync();
  // ---- quadrupole.py:327-335 observation importance ------------------------------------------------
  if (lane < nobs) {
    float sc = 0.f;
    _Pragma("unroll 1") for (int c = 0; c < C; c++) sc = __fadd_rn(sc, tmpl[lane * C + c]);
    sm.sti[lane] = twice(sc);  // sum_template_intensity, also used by the cosine score
  }
  tile.sync();
  {
    float tot = 0.f;
    _Pragma("unroll 1") for (int o = 0; o < nobs; o++) tot = __fadd_rn(tot, sm.sti[o]);
    if (lane < nobs) sm.oi[lane] = (tot == 0.f) ? __fdiv_rn(1.0f, (float)nobs) : __fdiv_rn(sm.sti[lane], tot);
  }
  // ---- candidate.py:319-329 fragment mask --------------------------------------------------------
  bool fvalid = false;
  if (lane < F) {
    float t_o = 0.f;
    _Pragma("unroll 1") for (int o = 0; o < nobs; o++) {
      float t_c = 0.f;
      const float* r = dfi + (lane * nobs + o) * C;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) t_c = __fadd_rn(t_c, r[c]);
      t_o = __fadd_rn(t_o, twice(t_c));
    }
    fvalid = t_o > 0.f;
  }
  const unsigned vb = tile.ballot(fvalid);
  const int Fv = __popc(vb);
  if (Fv < 2) return;
  if (fvalid) sm.fmap[__popc(vb & ((1u << lane) - 1u))] = lane;
  tile.sync();
  const bool act = lane < Fv;          // lane w <-> masked fragment w
  const int f = act ? sm.fmap[lane] : 0;
  {  // fragment_container.py:119-120 renormalise, fragment_features.py:218
    float isum = 0.f;
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) isum = __fadd_rn(isum, sm.intensity[sm.fmap[w]]);
    if (act) sm.fint[lane] = __fdiv_rn(sm.intensity[f], isum);
    tile.sync();
    float t = 0.f;
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) t = __fadd_rn(t, sm.fint[w]);
    if (act) sm.fin[lane] = __fdiv_rn(sm.fint[lane], t);
  }
  // ---- candidate.py:341 template frame profile with or_envelope (scoring/utils.py:46-53) -----------
  _Pragma("unroll 1") for (int t = lane; t < nobs * C; t += TILE) {
    int c = t % C;
    float x = twice(tmpl[t]);
    float res = x;
    if (c >= 1 && c < C - 1) {
      float xl = twice(tmpl[t - 1]), xr = twice(tmpl[t + 1]);
      if (x < xl || x < xr) res = (float)((double)__fadd_rn(xl, xr) / 2);
    }
    tfp[t] = res;
  }
  // distance-weight tables for weighted_center_mean (features_utils.py:9-26)
  _Pragma("unroll 1") for (int t = lane; t < 2 * C; t += TILE) {  // precursor "centres" = (n_scans, n_observations) = (2, 1)
    int s = t / C, c = t % C;
    double ds = (double)s - 2.0, dc = (double)c - 1.0;
    wtab_p[t] = exp(-0.1 * sqrt(ds * ds + dc * dc));
  }
  if (lane < nobs) {  // fragment_features.py:20-49 centre of mass of the template
    const float* r = tmpl + lane * C;
    double isum = 0, ssum = 0, fsum = 0;
    bool any = false;
    _Pragma("unroll 1") for (int s = 0; s < 2; s++)
      _Pragma("unroll 1") for (int c = 0; c < C; c++) { float v = r[c]; if (v > 0.f) { any = true; isum = __dadd_rn(isum, (double)v); } }
    if (any)
      _Pragma("unroll 1") for (int s = 0; s < 2; s++)
        _Pragma("unroll 1") for (int c = 0; c < C; c++) {
          float v = r[c];
          if (v > 0.f) { ssum = __dadd_rn(ssum, __dmul_rn((double)s, (double)v)); fsum = __dadd_rn(fsum, __dmul_rn((double)c, (double)v)); }
        }
    sm.esc[lane] = (any && isum > 0) ? ssum / isum : 0.0;
    sm.efc[lane] = (any && isum > 0) ? fsum / isum : 0.0;
  }
  tile.sync();
  _Pragma("unroll 1") for (int t = lane; t < nobs * 2 * C; t += TILE) {
    int o = t / (2 * C), s = (t / C) % 2, c = t % C;
    double ds = (double)s - sm.esc[o], dc = (double)c - sm.efc[o];
    wtab[t] = exp(-0.1 * sqrt(ds * ds + dc * dc));
  }
  tile.sync();

  float* fa = sm.feat;
  _Pragma("unroll 1") for (int t = lane; t < ADB_NUM_FEATURES; t += TILE) fa[t] = 0.f;
  tile.sync();
  if (lane == 0) {
    fa[28] = (float)((double)Fv / (double)F);  // candidate.py:362
    // features/location_features.py:9-33
    fa[0] = __fsub_rn(raw.mobility_values[scan_start], raw.mobility_values[scan_stop - 1]);
    fa[1] = __fsub_rn(raw.rt_values[frame_stop - 1], raw.rt_values[frame_start]);
    fa[2] = raw.rt_values[frame_center];
    fa[3] = raw.mobility_values[scan_center];
    fa[17] = (float)nobs;
  }

  // ================= features/precursor_features.py:14-102 =================
  if (lane < nI) {
    const float* r = dpi + lane * C;
    float tc = 0.f;
    _Pragma("unroll 1") for (int c = 0; c < C; c++) tc = __fadd_rn(tc, r[c]);
    float spi = twice(tc);
    float wsp = 0.f;
    _Pragma("unroll 1") for (int o = 0; o < nobs; o++) wsp = __fadd_rn(wsp, __fmul_rn(spi, sm.oi[o]));
    sm.spi[lane] = spi;
    sm.wspi[lane] = wsp;
    const float* rm = dpm + lane * C;
    double hh, mm;
    weighted_center_mean_pair(r, rm, wtab_p, C, hh, mm);
    sm.H[lane] = hh;
    sm.MZo[lane] = mm;
  }
  tile.sync();
  if (lane == 0) {
    int amax = 0;
    _Pragma("unroll 1") for (int i = 1; i < nI; i++) if (sm.iso_int[i] > sm.iso_int[amax]) amax = i;
    fa[4] = sm.wspi[0];
    fa[5] = sm.wspi[amax];
    float t6 = 0.f, t7 = 0.f;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++) { t6 = __fadd_rn(t6, sm.wspi[i]); t7 = __fadd_rn(t7, __fmul_rn(sm.wspi[i], sm.iso_int[i])); }
    fa[6] = t6; fa[7] = t7;
    double wme = 0;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++) if (sm.MZo[i] > 0) {
      double me = (sm.MZo[i] - (double)sm.iso_mz[i]) / (double)sm.iso_mz[i] * 1e6;
      wme = __dadd_rn(wme, __dmul_rn(me, (double)sm.iso_int[i]));
    }
    fa[8] = (float)wme;
    fa[9] = (float)fabs(wme);
    fa[10] = (float)__dadd_rn((double)sm.iso_mz[0], __dmul_rn(__dmul_rn(wme, 1e-6), (double)sm.iso_mz[0]));
    fa[11] = (float)sm.H[0];
    fa[12] = (float)sm.H[amax];
    double t13 = 0, t14 = 0, hbar = 0;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++) { t13 = __dadd_rn(t13, sm.H[i]); t14 = __dadd_rn(t14, __dmul_rn(sm.H[i], (double)sm.iso_int[i])); }
    fa[13] = (float)t13; fa[14] = (float)t14;
    hbar = t13 / (double)nI;
    float sx = 0.f, sy = 0.f;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++) { sx = __fadd_rn(sx, sm.iso_int[i]); sy = __fadd_rn(sy, sm.spi[i]); }
    double xbar = (double)sx / (double)nI, ybar = (double)sy / (double)nI;
    double num = 0, sxx = 0, syy = 0, num2 = 0, shh = 0;
    _Pragma("unroll 1") for (int i = 0; i < nI; i++) {
      double a = (double)sm.iso_int[i] - xbar, b = (double)sm.spi[i] - ybar, h = sm.H[i] - hbar;
      num = __dadd_rn(num, __dmul_rn(a, b)); sxx = __dadd_rn(sxx, __dmul_rn(a, a)); syy = __dadd_rn(syy, __dmul_rn(b, b));
      num2 = __dadd_rn(num2, __dmul_rn(a, h)); shh = __dadd_rn(shh, __dmul_rn(h, h));
    }
    fa[15] = (float)(num / (sqrt(sxx * syy) + 1e-12));
    fa[16] = (float)(num2 / (sqrt(sxx * shh) + 1e-12));
  }

  PHASE_BARRIER();  // ======== end of code region B (template, profiles, precursor features) ========
  // ================= features/fragment_features.py:198-427 =================
  int best_obs = 0;
  _Pragma("unroll 1") for (int o = 1; o < nobs; o++) if (sm.oi[o] > sm.oi[best_obs]) best_obs = o;
  const bool quant_all = cfg.quant_all != 0;
  int64_t qw = (int64_t)cfg.quant_window;
  if ((C / 2) - 1 < qw) qw = (C / 2) - 1;
  const int center = C / 2;
  int w0 = center - (int)qw, w1 = center + (int)qw + 1;
  if (qw < 0) { w0 = 0; w1 = 0; }
  if (w1 > C) w1 = C;
  if (w0 < 0) w0 = 0;
  const int wn = max(w1 - w0, 0);
  bool anyh = false;
  if (act) {
    float* b = bp + lane * C;
    const float* d = dfi + f * nobs * C;
    if (quant_all) {
      _Pragma("unroll 1") for (int c = 0; c < C; c++) { float t = 0.f; _Pragma("unroll 1") for (int o = 0; o < nobs; o++) t = __fadd_rn(t, twice(d[o * C + c])); b[c] = t; }
    } else {
      _Pragma("unroll 1") for (int c = 0; c < C; c++) b[c] = twice(d[best_obs * C + c]);
    }
    // center_envelope_1d, fragment_features.py:71-159
    if (C % 2 == 0) {
      int cr = C / 2, cl = cr - 1;
      if (cl >= 0) {
        float left = b[cl], right = b[cr];
        _Pragma("unroll 1") for (int i = 1; i <= cl; i++) {
          b[cl - i] = fminf(left, b[cl - i]);
          left = (float)((double)__fadd_rn(b[cl - i], b[cl - i + 1]) * 0.5);
          b[cr + i] = fminf(right, b[cr + i]);
          right = (float)((double)__fadd_rn(b[cr + i], b[cr + i - 1]) * 0.5);
        }
      }
    } else if (C >= 3) {
      int cc = C / 2;
      float left = (float)((double)__fadd_rn(b[cc - 1], b[cc]) * 0.5);
      float right = (float)((double)__fadd_rn(b[cc + 1], b[cc]) * 0.5);
      _Pragma("unroll 1") for (int i = 1; i <= cc; i++) {
        b[cc - i] = fminf(left, b[cc - i]);
        left = (float)((double)__fadd_rn(b[cc - i], b[cc - i + 1]) * 0.5);
        b[cc + i] = fminf(right, b[cc + i]);
        right = (float)((double)__fadd_rn(b[cc + i], b[cc + i - 1]) * 0.5);
      }
    }
    // trapezoid area over the quant window, fragment_features.py:253-273
    double area = 0;
    _Pragma("unroll 1") for (int t = 0; t + 1 < wn; t++) {
      float drt = __fsub_rn(__ldg(raw.rt_values + frame_start + (int64_t)(w0 + t + 1) * L), __ldg(raw.rt_values + frame_start + (int64_t)(w0 + t) * L));
      float sum2 = __fadd_rn(b[w0 + t + 1], b[w0 + t]);
      area = __dadd_rn(area, __dmul_rn((double)__fmul_rn(sum2, drt), 0.5));
    }
    sm.area_norm[lane] = __dmul_rn(area, (double)qw);
    float ofi = 0.f;
    _Pragma("unroll 1") for (int u = 0; u < wn; u++) ofi = __fadd_rn(ofi, b[w0 + u]);
    sm.ofi[lane] = ofi;

    // per-observation: summed intensity (cosine score) and the observation mask.  A weighted-centre height is
    // > 0 exactly when the (f, o) cell row has signal (all weights are positive), i.e. when its f32 sum is > 0.
    float fn2 = 0.f, dot = 0.f, wsum = 0.f;
    unsigned obs_mask = 0u;
    const float* dm = dfm + f * nobs * C;
    _Pragma("unroll 1") for (int o = 0; o < nobs; o++) {
      const float* r = d + o * C;
      float tc = 0.f;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) tc = __fadd_rn(tc, r[c]);
      float v = twice(tc);
      fn2 = __fadd_rn(fn2, __fmul_rn(v, v));
      dot = __fadd_rn(dot, __fmul_rn(v, sm.sti[o]));
      const bool mm = tc > 0.f;
      if (mm && o < 32) obs_mask |= 1u << o;
      wsum = __fadd_rn(wsum, mm ? sm.oi[o] : 0.0f);  // fragment_features.py:318-326
    }
    anyh = obs_mask != 0u;
    {  // cosine_similarity_a1, features_utils.py:40-47
      float tn2 = 0.f;
      _Pragma("unroll 1") for (int o = 0; o < nobs; o++) tn2 = __fadd_rn(tn2, __fmul_rn(sm.sti[o], sm.sti[o]));
      double div = (double)__fmul_rn(sqrtf(fn2), sqrtf(tn2)) + 0.0001;
      sm.cosv[lane] = (float)((double)dot / div);
    }
    // fragment_features.py:312-336 observation-weighted means of the weighted-centre height and m/z
    double wtot = 0;
    int cnt = 0;
    _Pragma("unroll 1") for (int o = 0; o < nobs; o++) {
      double wv = (double)(((obs_mask >> o) & 1u) ? sm.oi[o] : 0.0f) / ((double)wsum + 1e-20);
      if (wv > 0) { wtot = __dadd_rn(wtot, wv); cnt++; }
    }
    double a = 0, bsum = 0;
    if (cnt > 0) {
      _Pragma("unroll 1") for (int o = 0; o < nobs; o++) {
        double wv = (double)(((obs_mask >> o) & 1u) ? sm.oi[o] : 0.0f) / ((double)wsum + 1e-20);
        if (wv > 0) {
          double h_o, mz_o;
          weighted_center_mean_pair(d + o * C, dm + o * C, wtab + o * 2 * C, C, h_o, mz_o);
          double lw = wv / wtot;
          a = __dadd_rn(a, __dmul_rn(mz_o, lw));
          bsum = __dadd_rn(bsum, __dmul_rn(h_o, lw));
        }
      }
    }
    sm.ofh_mean[lane] = bsum;
    sm.ci[lane] = a;  // observed_fragment_mz_mean (slot reused below for centre intensities)
    double mzf = (double)sm.mz[f];
    sm.mass_error[lane] = (a - mzf) / mzf * 1e6;
    // np.argsort(fragments.intensity)[::-1]
    float v = sm.fint[lane];
    int rank_asc = 0;
    _Pragma("unroll 1") for (int q = 0; q < Fv; q++) rank_asc += (sm.fint[q] < v) || (sm.fint[q] == v && q < lane);
    sm.sorted_idx[Fv - 1 - rank_asc] = lane;
  }
  const unsigned anyh_b = tile.ballot(anyh);
  tile.sync();
  // fragment-level outputs, candidate.py:403-442
  const size_t obase = (size_t)ci * (size_t)K;
  if (act && cfg.collect_fragments && lane < K) {
    P.out.fragment_mz_library[obase + lane] = sm.mz_library[f];
    P.out.fragment_mz[obase + lane] = sm.mz[f];
    P.out.fragment_mz_observed[obase + lane] = (float)sm.ci[lane];
    P.out.fragment_height[obase + lane] = (float)sm.ofh_mean[lane];
    P.out.fragment_intensity[obase + lane] = (float)sm.area_norm[lane];
    P.out.fragment_mass_error[obase + lane] = (float)sm.mass_error[lane];
    P.out.fragment_position[obase + lane] = sm.position[f];
    P.out.fragment_number[obase + lane] = sm.number[f];
    P.out.fragment_type[obase + lane] = sm.type[f];
    P.out.fragment_charge[obase + lane] = sm.charge[f];
    P.out.fragment_loss_type[obase + lane] = sm.loss_type[f];
  }
  tile.sync();
  if (lane == 0) {
    double sum_ofh = 0;
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) sum_ofh = __dadd_rn(sum_ofh, sm.ofh_mean[w]);
    if (anyh_b != 0u) fa[18] = (float)corrcoef01(sm.area_norm, sm.fin, Fv);
    if (sum_ofh > 0.0) fa[19] = (float)corrcoef01(sm.ofh_mean, sm.fin, Fv);
    int n20 = 0, n21 = 0;
    float s22 = 0.f, s23 = 0.f, cacc = 0.f;
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) if (sm.ofi[w] > 0.f) { n20++; s22 = __fadd_rn(s22, sm.fin[w]); cacc = __fadd_rn(cacc, sm.cosv[w]); }
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) if (sm.ofh_mean[w] > 0.0) { n21++; s23 = __fadd_rn(s23, sm.fin[w]); }
    fa[20] = (float)((double)n20 / (double)Fv);
    fa[21] = (float)((double)n21 / (double)Fv);
    fa[22] = s22; fa[23] = s23;
    if (n20 > 0) fa[24] = (float)((double)cacc / (double)n20);
    float sb = 0.f, sy = 0.f;
    int nb = 0, ny = 0, min_y = 255, max_b = 0;
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) {
      int ty = sm.type[sm.fmap[w]], po = sm.position[sm.fmap[w]];
      if (ty == 98) { sb = __fadd_rn(sb, sm.ofi[w]); nb++; max_b = max(max_b, po); }
      if (ty == 121) { sy = __fadd_rn(sy, sm.ofi[w]); ny++; min_y = min(min_y, po); }
    }
    fa[25] = nb > 0 ? (float)log((double)sb + 1.0) : 0.f;
    fa[26] = ny > 0 ? (float)log((double)sy + 1.0) : 0.f;
    fa[27] = __fsub_rn(fa[25], fa[26]);
    int n3 = min(Fv, 3);
    double t41 = 0, t42 = 0;
    _Pragma("unroll 1") for (int r = 0; r < n3; r++) t41 = __dadd_rn(t41, sm.mass_error[sm.sorted_idx[r]]);
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) t42 = __dadd_rn(t42, sm.mass_error[w]);
    fa[41] = (float)(t41 / (double)n3);
    fa[42] = (float)(t42 / (double)Fv);
    if (nb > 0 && ny > 0) {
      int n_ov = 0;
      double sa = 0, se = 0;
      _Pragma("unroll 1") for (int w = 0; w < Fv; w++) {
        int ty = sm.type[sm.fmap[w]], po = sm.position[sm.fmap[w]];
        bool ov = (ty == 121 && po < max_b) || (ty == 98 && po > min_y);
        if (ov) { n_ov++; sa = __dadd_rn(sa, sm.area_norm[w]); se = __dadd_rn(se, sm.mass_error[w]); }
      }
      fa[43] = (float)n_ov;
      if (n_ov > 0) { fa[44] = (float)(sa / (double)n_ov); fa[45] = (float)(se / (double)n_ov); }
      else { fa[44] = 0.f; fa[45] = 15.f; }
    }
  }
  tile.sync();

  PHASE_BARRIER();  // ======== end of code region C (fragment features) ========
  // ================= features/profile_features.py:18-206 =================
  // fragments_frame_profile accessor: the best observation's rows were enveloped in place when
  // quant_all is off (fragment_features.py:248-250, view semantics)
  auto ffp = [&](int w, int fidx, int o, int c) -> float {
    if (!quant_all && o == best_obs) return bp[w * C + c];
    return twice(dfi[(fidx * nobs + o) * C + c]);
  };
  // fragments_frame_profile.sum(axis=1)
  // with several observations every lane sums its row once into the second F x C block of the (dead) m/z channel
  float* isl_rows = dfm + F * C;  // valid when nobs >= 2: nFC >= 2 F C, and nrm only uses the first block
  if (cfg.experimental_xic && nobs > 1 && act)
    _Pragma("unroll 1") for (int c = 0; c < C; c++)
      isl_rows[lane * C + c] = frame_profile_obs_sum(dfi + f * nobs * C, bp + lane * C, nobs, C, c, quant_all ? -1 : best_obs);
  auto isl = [&](int w, int fidx, int c) -> float {
    if (nobs == 1)  // single observation: 0 + x == x exactly, no call
      return quant_all ? twice(dfi[fidx * C + c]) : bp[w * C + c];
    return isl_rows[w * C + c];
  };
  if (cfg.experimental_xic) {
    int a0 = center - 1, a1 = center + 2;  // scoring_utils.py:100-110 python slice semantics
    if (a0 < 0) { a0 += C; if (a0 < 0) a0 = 0; }
    if (a1 > C) a1 = C;
    const int wnn = max(a1 - a0, 0);
    if (act) {
      float t = 0.f;
      _Pragma("unroll 1") for (int c = a0; c < a1; c++) t = __fadd_rn(t, isl(lane, f, c));
      double cint = (double)t / (double)wnn;
      float* nr = nrm + lane * C;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) nr[c] = (cint > 0) ? (float)((double)isl(lane, f, c) / cint) : 0.f;
    }
    tile.sync();
    _Pragma("unroll 1") for (int c = lane; c < C; c += TILE) {  // median over fragments (scoring_utils.py:127-152)
      float vlo = 0.f, vhi = 0.f;
      _Pragma("unroll 1") for (int w = 0; w < Fv; w++) {
        float v = nrm[w * C + c];
        int rk = 0;
        _Pragma("unroll 1") for (int u = 0; u < Fv; u++) { float vu = nrm[u * C + c]; rk += (vu < v) || (vu == v && u < w); }
        if (rk == (Fv - 1) / 2) vlo = v;
        if (rk == Fv / 2) vhi = v;
      }
      med[c] = (Fv & 1) ? vhi : (float)((double)__fadd_rn(vlo, vhi) / 2);
    }
    tile.sync();
    // correlation_coefficient(median_profile, intensity_slice), scoring_utils.py:20-76
    float sx = 0.f;
    _Pragma("unroll 1") for (int c = 0; c < C; c++) sx = __fadd_rn(sx, med[c]);
    const double mxv = (double)sx / (double)C;
    double varx = 0;
    _Pragma("unroll 1") for (int c = 0; c < C; c++) { double dd = (double)med[c] - mxv; varx = __dadd_rn(varx, __dmul_rn(dd, dd)); }
    varx /= (double)C;
    if (act) {
      float sy = 0.f;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) sy = __fadd_rn(sy, isl(lane, f, c));
      float myv = (float)((double)sy / (double)C);
      double cov = 0;
      float vy32 = 0.f;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) {
        float ym = __fsub_rn(isl(lane, f, c), myv);
        cov = __dadd_rn(cov, __dmul_rn((double)med[c] - mxv, (double)ym));
        vy32 = __fadd_rn(vy32, __fmul_rn(ym, ym));
      }
      cov /= (double)C;
      double vxy = varx * ((double)vy32 / (double)C);
      sm.corr_list[lane] = (vxy == 0) ? 0.f : (float)(cov / sqrt(vxy));
    }
    tile.sync();
    if (lane == 0) {
      int n3 = min(Fv, 3);
      float t = 0.f;
      _Pragma("unroll 1") for (int r = 0; r < n3; r++) t = __fadd_rn(t, sm.corr_list[sm.sorted_idx[r]]);
      fa[32] = (float)((double)t / (double)n3);
    }
  } else {
    // legacy: observation-weighted F x F correlation matrix (scoring/utils.py:513-571), float32
    _Pragma("unroll 1") for (int t = lane; t < Fv * Fv; t += TILE) red[t] = 0.f;
    _Pragma("unroll 1") for (int o = 0; o < nobs; o++) {
      tile.sync();
      if (act) {
        float s = 0.f;
        _Pragma("unroll 1") for (int c = 0; c < C; c++) s = __fadd_rn(s, ffp(lane, f, o, c));
        float mean = __fdiv_rn(s, (float)C);
        float ss = 0.f;
        float* cen = nrm + lane * C;
        _Pragma("unroll 1") for (int c = 0; c < C; c++) { float cv = __fsub_rn(ffp(lane, f, o, c), mean); cen[c] = cv; ss = __fadd_rn(ss, __fmul_rn(cv, cv)); }
        sm.rfw[lane] = sqrtf(__fdiv_rn(ss, (float)C));
      }
      tile.sync();
      _Pragma("unroll 1") for (int t = lane; t < Fv * Fv; t += TILE) {
        int a = t / Fv, b = t % Fv;
        const float* ca = nrm + a * C;
        const float* cb = nrm + b * C;
        float dot = 0.f;
        _Pragma("unroll 1") for (int c = 0; c < C; c++) dot = __fadd_rn(dot, __fmul_rn(ca[c], cb[c]));
        float cov = __fdiv_rn(dot, (float)C);
        float smx = __fmul_rn(sm.rfw[a], sm.rfw[b]);
        float corr = (float)((double)cov / ((double)smx + 1e-12));
        red[t] = __fadd_rn(red[t], __fmul_rn(corr, sm.oi[o]));
      }
    }
    tile.sync();
    if (act) {
      float t = 0.f;
      _Pragma("unroll 1") for (int g = 0; g < Fv; g++) t = __fadd_rn(t, __fmul_rn(red[lane * Fv + g], sm.fint[g]));
      sm.corr_list[lane] = t;
    }
    tile.sync();
    if (lane == 0) {
      int n3 = min(Fv, 3);
      float t = 0.f;
      _Pragma("unroll 1") for (int a = 0; a < n3; a++) _Pragma("unroll 1") for (int b = 0; b < n3; b++) t = __fadd_rn(t, red[sm.sorted_idx[a] * Fv + sm.sorted_idx[b]]);
      fa[32] = (float)((double)t / (double)(n3 * n3));
    }
  }
  // template correlation, cycle fwhm, frame peak — lane w <-> fragment w
  _Pragma("unroll 1") for (int o = 0; o < nobs; o++) {
    // y statistics of the template frame profile (all lanes, sequential)
    const float* y = tfp + o * C;
    float ys = 0.f;
    _Pragma("unroll 1") for (int c = 0; c < C; c++) ys = __fadd_rn(ys, y[c]);
    const float ym = __fdiv_rn(ys, (float)C);
    float yss = 0.f;
    _Pragma("unroll 1") for (int c = 0; c < C; c++) { float yc = __fsub_rn(y[c], ym); yss = __fadd_rn(yss, __fmul_rn(yc, yc)); }
    const float ystd = sqrtf(__fdiv_rn(yss, (float)C));
    if (act) {
      float xs = 0.f, mxv = 0.f;
      int am = 0;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) {
        float x = ffp(lane, f, o, c);
        xs = __fadd_rn(xs, x);
        if (c == 0 || x > mxv) { am = c; mxv = x; }
      }
      float xm = __fdiv_rn(xs, (float)C);
      float xss = 0.f;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) { float xc = __fsub_rn(ffp(lane, f, o, c), xm); xss = __fadd_rn(xss, __fmul_rn(xc, xc)); }
      float dot = 0.f;
      int na = 0;
      double half = (double)mxv / 2;
      _Pragma("unroll 1") for (int c = 0; c < C; c++) {
        float x = ffp(lane, f, o, c);
        dot = __fadd_rn(dot, __fmul_rn(__fsub_rn(x, xm), __fsub_rn(y[c], ym)));
        na += (double)x > half;
      }
      float xstd = sqrtf(__fdiv_rn(xss, (float)C));
      float cov = __fdiv_rn(dot, (float)C);
      float ct = (float)((double)cov / ((double)__fmul_rn(xstd, ystd) + 1e-12));
      float rt_width = __fsub_rn(raw.rt_values[frame_stop - 1], raw.rt_values[frame_start]);
      float fw = (float)(((double)na / (double)C) * (double)rt_width);
      // accumulate over observations: template corr (profile_features.py:82-85), fwhm (:142-144)
      if (o == 0) { sm.rfw[lane] = 0.f; sm.cosv[lane] = 0.f; }
      sm.cosv[lane] = __fadd_rn(sm.cosv[lane], __fmul_rn(ct, sm.oi[o]));
      sm.rfw[lane] = __fadd_rn(sm.rfw[lane], __fmul_rn(fw, sm.oi[o]));
      sm.frame_peak[lane] = am;
    }
    tile.sync();
    if (act) {  // median frame peak of this observation (profile_features.py:193-204): every lane ranks its own value
      const int v = sm.frame_peak[lane];
      int rk = 0;
      _Pragma("unroll 1") for (int u = 0; u < Fv; u++) rk += (sm.frame_peak[u] < v) || (sm.frame_peak[u] == v && u < lane);
      if (rk == (Fv - 1) / 2) sm.t_sel[0] = v;  // ranks are a permutation: exactly one lane holds each order statistic
      if (rk == Fv / 2) sm.t_sel[1] = v;
    }
    tile.sync();
    if (lane == 0) {
      const double vlo = (double)sm.t_sel[0], vhi = (double)sm.t_sel[1];
      float medp = (float)((Fv & 1) ? vhi : (vlo + vhi) / 2);
      double delta = (double)medp - floor((double)C / 2);
      double prev = (o == 0) ? 0.0 : sm.esc[0];
      sm.esc[0] = __dadd_rn(prev, __dmul_rn(delta, (double)sm.oi[o]));  // esc no longer needed
    }
    tile.sync();
  }
  if (lane == 0) {
    float t31 = 0.f, t33 = 0.f, t38 = 0.f;
    _Pragma("unroll 1") for (int w = 0; w < Fv; w++) {
      t31 = __fadd_rn(t31, sm.corr_list[w]);
      t33 = __fadd_rn(t33, __fmul_rn(sm.cosv[w], sm.fint[w]));
      t38 = __fadd_rn(t38, __fmul_rn(sm.rfw[w], sm.fint[w]));
    }
    fa[31] = (float)((double)t31 / (double)Fv);
    fa[33] = t33;
    fa[38] = t38;
    fa[40] = (float)sm.esc[0];
    // profile_features.py:94-113 (the type mask indexes the sorted-index array by position)
    int nb = 0, ny = 0;
    float sb = 0.f, sy = 0.f;
    _Pragma("unroll 1") for (int r = 0; r < Fv; r++) {
      int ty = sm.type[sm.fmap[r]];
      if (ty == 98) { if (nb < 3) sb = __fadd_rn(sb, sm.corr_list[sm.sorted_idx[r]]); nb++; }
      if (ty == 121) { if (ny < 3) sy = __fadd_rn(sy, sm.corr_list[sm.sorted_idx[r]]); ny++; }
    }
    if (nb > 0) { fa[34] = (float)((double)sb / (double)min(nb, 3)); fa[35] = (float)nb; }
    if (ny > 0) { fa[36] = (float)((double)sy / (double)min(ny, 3)); fa[37] = (float)ny; }
  }
  tile.sync();
  // ---- candidate.py:475-481 ---------------------------------------------------------------------
  if (act && cfg.collect_fragments && lane < K) P.out.fragment_correlation[obase + lane] = sm.corr_list[lane];
  _Pragma("unroll 1") for (int t = lane; t < ADB_NUM_FEATURES; t += TILE) P.out.features[(size_t)ci * ADB_NUM_FEATURES + t] = fa[t];
  if (lane == 0) P.out.valid[ci] = 1;
}

// The kernel body is ~110 KB of almost straight-line SASS, far beyond the 32 KB per-SM instruction cache: with independent
// warps every warp streams the whole program from L2 for every candidate and the per-GPC instruction cache saturates
// (ncu r1: gcc__cache_requests_type_instruction at 86 % of peak, SM i-cache hit rate 59 %).  So ONE CTA per SM runs all
// its tiles through the program in lock step: the code is cut into four regions of <= 32 KB and a CTA barrier separates
// them, so a region is fetched once per SM and round instead of once per warp.  The candidates of one round have the
// same cost class (the processing order sorts by it), which keeps the barriers cheap.
template <int TILE>
__global__ void __launch_bounds__(SCORE_THREADS, SCORE_CTAS_PER_SM) adb_score_kernel(const __grid_constant__ ScoreParams P) {
  extern __shared__ __align__(16) float dyn_smem[];
  constexpr int TILES = SCORE_THREADS / TILE;
  TileSmall<TILE>* small = reinterpret_cast<TileSmall<TILE>*>(dyn_smem + (size_t)TILES * SMEM_FLOATS_PER_TILE);
  cg::thread_block block = cg::this_thread_block();
  cg::thread_block_tile<TILE> tile = cg::tiled_partition<TILE>(block);
  const int tib = (int)(threadIdx.x / TILE);
  const long long gt = (long long)blockIdx.x * TILES + tib;
  float* scratch = dyn_smem + (size_t)tib * SMEM_FLOATS_PER_TILE;
  float* ws = P.workspace ? P.workspace + (size_t)gt * (size_t)P.ws_floats_per_tile : nullptr;
  for (long long base = (long long)blockIdx.x * TILES; base < P.cand.n; base += (long long)gridDim.x * TILES) {
    const long long it = base + tib;
    int phase = 0;
    if (it < P.cand.n) {
      const long long ci = P.order ? (long long)P.order[it] : it;
      score_one_body<TILE>(P, ci, tile, small[tib], scratch, ws, phase);
    }
    for (; phase < SCORE_PHASES + 1; phase++) cta_phase_barrier();  // regions skipped by this tile + end of round
  }
}

template <int TILE>
size_t score_smem_bytes() {
  constexpr int TILES = SCORE_THREADS / TILE;
  return (size_t)TILES * SMEM_FLOATS_PER_TILE * sizeof(float) + (size_t)TILES * sizeof(TileSmall<TILE>);
}

template <int TILE>
int resident_tiles(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * SCORE_CTAS_PER_SM * (SCORE_THREADS / TILE);
}

}  // namespace

int adb_score_tile(int top_k) { return top_k <= 16 ? 16 : 32; }

int adb_score_resident_tiles(int device, int top_k) {
  return top_k <= 16 ? resident_tiles<16>(device) : resident_tiles<32>(device);
}

int64_t adb_score_workspace_floats(int top_k, int64_t c_max) {
  int64_t F = top_k, nobs = ADB_MAX_OBS, I = ADB_MAX_ISOTOPES, C = c_max;
  return 2 * F * nobs * C + 2 * F * C + 2 * I * C + 2 * nobs * C + C + 2 + 4 * nobs * C + 4 * C + F * F + 16;
}

void adb_launch_score(const DevRaw& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand,
                      DevScoresOut out, float* d_workspace, int64_t workspace_floats_per_tile, int n_resident_tiles,
                      const int32_t* d_order, uint32_t* d_status, cudaStream_t stream, int* n_launches) {
  if (cand.n <= 0) return;
  ScoreParams P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.cand = cand; P.out = out;
  P.workspace = d_workspace; P.ws_floats_per_tile = workspace_floats_per_tile; P.status = d_status; P.order = d_order;
  const int tile = adb_score_tile((int)cfg.top_k_fragments);
  const int tiles_per_block = SCORE_THREADS / tile;
  long long blocks = n_resident_tiles / tiles_per_block;  // persistent: one CTA per SM, rounds of tiles_per_block candidates
  long long needed = (cand.n + tiles_per_block - 1) / tiles_per_block;
  if (blocks > needed) blocks = needed;
  if (blocks < 1) blocks = 1;
  if (tile == 16) {
    const size_t dyn = score_smem_bytes<16>();
    cudaFuncSetAttribute(adb_score_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    adb_score_kernel<16><<<(unsigned)blocks, SCORE_THREADS, dyn, stream>>>(P);
  } else {
    const size_t dyn = score_smem_bytes<32>();
    cudaFuncSetAttribute(adb_score_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    adb_score_kernel<32><<<(unsigned)blocks, SCORE_THREADS, dyn, stream>>>(P);
  }
  if (n_launches) (*n_launches)++;
}

// alphadia_b200 — candidate scoring on timsTOF (4-D) raw files, sm_100a.
//
// Replaces Candidate.process (alphadia/search/scoring/containers/candidate.py:166-481) for data with ion mobility:
// TimsTOFTransposeJIT.get_dense(absolute_masses=True) -> _assemble_push (jitclasses/bruker_jit.py:586-615,352-504),
// the quadrupole transfer function per (isotope, observation, scan) / template / observation importance
// (scoring/quadrupole.py:80-115,261-335), frame and scan profiles (scoring/utils.py:26-66) and the 46 features
// including the mobility ones (features/fragment_features.py:430-480 -> 29/30, features/profile_features.py:148-188 -> 39).
//
// One warp per candidate, persistent grid, candidates visited in (tof row of the precursor, frame_start) order.
//   extract   the cube cell recurrence (intensity sum + intensity-weighted mean m/z) is order dependent: tof rows
//             ascending, events ascending.  Inside ONE tof row every event has a distinct push index, hence a distinct
//             cell, so the events of a row are independent: each query (fragment / isotope) owns a group of lanes that
//             strides over the events of the current row, rows advance in lock step with a warp barrier between them.
//   cube      [query][observation][scan][cycle] in a per-warp HBM workspace (it is 20-500 KB and > 95 % zeros: it lives
//             in L2); per-fragment / per-observation statistics in shared memory.
//   features  lanes <-> independent rows (profiles, weighted centres) or <-> fragments; every f32 accumulation keeps the
//             reference's order (explicit __fadd_rn/__fmul_rn, no contraction).
#include "adb_common.cuh"

#define FULL 0xffffffffu
#define SC4_THREADS 128
#define SC4_WARPS (SC4_THREADS / 32)
#define SC4_F 32  // == ADB_MAX_FRAGMENTS

namespace {

struct Small4 {
  float mz_library[SC4_F], mz[SC4_F], intensity[SC4_F];
  float lo[SC4_F], hi[SC4_F];
  uint8_t type[SC4_F], loss_type[SC4_F], charge[SC4_F], number[SC4_F], position[SC4_F];
  int t0[SC4_F], t1[SC4_F];
  int fmap[SC4_F], sorted_idx[SC4_F], frame_peak[SC4_F], zidx[SC4_F];
  float fint[SC4_F], fin[SC4_F], ofi[SC4_F], cosv[SC4_F], corr_list[SC4_F], rfw[SC4_F], acc_a[SC4_F], acc_b[SC4_F], znorm[SC4_F];
  double area_norm[SC4_F], ofh_mean[SC4_F], mass_error[SC4_F], ofmm[SC4_F];
  double esc[ADB_MAX_OBS4], efc[ADB_MAX_OBS4];
  double H[ADB_MAX_ISOTOPES], MZo[ADB_MAX_ISOTOPES];
  float oi[ADB_MAX_OBS4], sti[ADB_MAX_OBS4];
  int obs_id[ADB_MAX_OBS4];
  float iso_mz[ADB_MAX_ISOTOPES], iso_int[ADB_MAX_ISOTOPES], lo_p[ADB_MAX_ISOTOPES], hi_p[ADB_MAX_ISOTOPES];
  float spi[ADB_MAX_ISOTOPES], wspi[ADB_MAX_ISOTOPES];
  int t0p[ADB_MAX_ISOTOPES], t1p[ADB_MAX_ISOTOPES];
  float feat[ADB_NUM_FEATURES];
  // staging of the library fragments before top-k
  float t_int[ADB_MAX_LIB_FRAGMENTS], t_mz[ADB_MAX_LIB_FRAGMENTS];
  int t_src[ADB_MAX_LIB_FRAGMENTS], t_sel[SC4_F];
  double delta_acc;
};

struct Score4Params {
  DevRaw4 raw;
  DevLib lib;
  adb_scoring_config cfg;
  DevCandidatesIn cand;
  DevScoresOut out;
  float* workspace;
  long long ws_floats_per_warp;
  uint32_t* status;
  const int32_t* order;
};

// workspace floats of one candidate (shared by the kernel carve-up and the host sizing)
__host__ __device__ inline long long score4d_need(long long F, long long nI, long long nobs, long long nobsP, long long S, long long C) {
  const long long nSC = S * C, mx = S > C ? S : C;
  long long n = 2 * F * nobs * nSC + 2 * nI * nobsP * nSC + 2 * nI * nSC + nobs * nSC + nobs * S;
  n += 2 * F * nobs * S + F * nobs * C + 2 * nobs * C + 2 * nobs * S;
  n += 3 * F * C + C + F * mx + F * F + 2 * nobs * F;
  n += 2 + 2 * (nI * nobs * S) + 4 * F * nobs;  // doubles: qtf, ofmz, ofh
  return n + 8;
}

// frame-in-cycle index of a frame ((push - zeroth * scans) mod (Fr * Sc) of bruker_jit.py:302-313, python modulo)
__device__ __forceinline__ int fic_of(long long frame, long long z, long long Fr) {
  long long r = (frame - z) % Fr;
  return (int)(r < 0 ? r + Fr : r);
}

// One query group extracts its tof rows into the cube (bruker_jit.py:352-504).  `grp` lanes share query j.
__device__ __forceinline__ void extract4d(const DevRaw4& raw, const int* t0, const int* t1, int nq, int nobs, unsigned long long obs_mask,
                                          int bit, long long frame_start, long long frame_stop, long long scan_start, int S, int C,
                                          long long cs, float q0, float q1, float* di, float* dm, int lane) {
  const int G = raw.obs_unique_per_scan ? max(32 / nq, 1) : 1;
  const int j = lane / G, sub = lane - j * G;
  const bool has = j < nq;
  int rows = has ? (t1[j] - t0[j]) : 0;
  int max_rows = rows;
  for (int off = 16; off > 0; off >>= 1) max_rows = max(max_rows, __shfl_xor_sync(FULL, max_rows, off));
  const long long smi = raw.scan_max_index, Fr = raw.Fr, z = raw.zeroth_frame;
  const long long p_lo = frame_start * smi, p_hi = frame_stop * smi;
  const long long nSC = (long long)S * C;
  (void)bit;
  for (int r = 0; r < max_rows; r++) {
    if (r < rows) {
      const int t = t0[j] + r;
      const double measured = __ldg(raw.mz_values + t);
      const int64_t r0 = __ldg(raw.tof_indptr + t), r1 = __ldg(raw.tof_indptr + t + 1);
      int64_t e = adb_row_lower_bound(raw.push, r0, r1, p_lo) + sub;
      for (; e < r1; e += G) {
        const uint32_t push = __ldg(raw.push + e);
        if ((long long)push >= p_hi) break;
        if (e > r0 && __ldg(raw.push + e - 1) == push) continue;  // not the first of a duplicate run
        const long long frame = push / smi, scan = push - frame * smi;
        if (scan < scan_start || scan >= scan_start + S) continue;
        const int fic = fic_of(frame, z, Fr);
        const long long pos = (long long)fic * raw.Sc + scan;
        if (!(((double)q0 <= __ldg(raw.cycle + 2 * pos + 1)) && ((double)q1 >= __ldg(raw.cycle + 2 * pos)))) continue;
        const long long id = __ldg(raw.dia_precursor_cycle + pos);
        if (id < 0 || id >= 64 || !((obs_mask >> id) & 1ull)) continue;
        const int o = __popcll(obs_mask & ((1ull << id) - 1ull));
        const long long rc = (frame - z) / Fr - cs;
        if (rc < 0 || rc >= C) continue;
        const long long cell = (((long long)j * nobs + o) * S + (scan - scan_start)) * C + rc;
        float acc_i = di[cell], acc_m = dm[cell];
        for (int64_t k = e; k < r1 && __ldg(raw.push + k) == push; k++) {  // a run of equal pushes is one lane's job
          const unsigned iv = __ldg(raw.intensity + k);
          const double ni = (double)iv;  // intensity * (intensity > 1e-26)
          const double num = __dadd_rn(__dadd_rn((double)__fmul_rn(acc_m, acc_i), __dmul_rn(ni, measured)), 1e-36);
          const double den = __dadd_rn(__dadd_rn((double)acc_i, ni), 1e-36);
          acc_i = (float)__dadd_rn((double)acc_i, ni);
          acc_m = (float)__ddiv_rn(num, den);
        }
        di[cell] = acc_i;
        dm[cell] = acc_m;
      }
    }
    __syncwarp();
  }
  (void)nSC;
}

// features_utils.py:9-26 weighted_center_mean over x[S][C] in row-major non-zero order
__device__ __noinline__ double wcm4(const float* x, int S, int C, double scan_center, double frame_center) {
  double values = 0, weights = 0;
  bool any = false;
#pragma unroll 1
  for (int s = 0; s < S; s++)
#pragma unroll 1
    for (int c = 0; c < C; c++) {
      const float v = x[s * C + c];
      if (v > 0.f) {
        any = true;
        const double ds = (double)s - scan_center, dc = (double)c - frame_center;
        const double w = exp(-0.1 * sqrt(__dadd_rn(__dmul_rn(ds, ds), __dmul_rn(dc, dc))));
        values = __dadd_rn(values, __dmul_rn((double)v, w));
        weights = __dadd_rn(weights, w);
      }
    }
  if (!any) return 0.0;
  return weights > 0 ? values / weights : 0.0;
}

__device__ __noinline__ double corrcoef01_4(const double* x, const float* yf, int n) {
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

// scoring/utils.py:46-66 or_envelope of one element (rows of length n, t = index inside the row)
__device__ __forceinline__ float or_env(const float* row, int n, int t) {
  const float x = row[t];
  if (t >= 1 && t < n - 1) {
    const float xl = row[t - 1], xr = row[t + 1];
    if (x < xl || x < xr) return (float)((double)__fadd_rn(xl, xr) / 2);
  }
  return x;
}

// fragment_features.py:71-159 center_envelope_1d in place
__device__ __noinline__ void center_envelope(float* b, int C) {
  if (C % 2 == 0) {
    const int cr = C / 2, cl = cr - 1;
    if (cl < 0) return;
    float left = b[cl], right = b[cr];
#pragma unroll 1
    for (int i = 1; i <= cl; i++) {
      b[cl - i] = fminf(left, b[cl - i]);
      left = (float)((double)__fadd_rn(b[cl - i], b[cl - i + 1]) * 0.5);
      b[cr + i] = fminf(right, b[cr + i]);
      right = (float)((double)__fadd_rn(b[cr + i], b[cr + i - 1]) * 0.5);
    }
  } else if (C >= 3) {
    const int cc = C / 2;
    float left = (float)((double)__fadd_rn(b[cc - 1], b[cc]) * 0.5);
    float right = (float)((double)__fadd_rn(b[cc + 1], b[cc]) * 0.5);
#pragma unroll 1
    for (int i = 1; i <= cc; i++) {
      b[cc - i] = fminf(left, b[cc - i]);
      left = (float)((double)__fadd_rn(b[cc - i], b[cc - i + 1]) * 0.5);
      b[cc + i] = fminf(right, b[cc + i]);
      right = (float)((double)__fadd_rn(b[cc + i], b[cc + i - 1]) * 0.5);
    }
  }
}

// sequential f32 sum of n elements with a stride
__device__ __forceinline__ float seq_sum(const float* x, int n, int stride) {
  float t = 0.f;
#pragma unroll 1
  for (int i = 0; i < n; i++) t = __fadd_rn(t, x[(long long)i * stride]);
  return t;
}

__device__ void score_one_4d(const Score4Params& P, int64_t ci, Small4& sm, float* ws, int lane) {
  const DevRaw4& raw = P.raw;
  const DevLib& lib = P.lib;
  const adb_scoring_config& cfg = P.cfg;
  const int K = (int)cfg.top_k_fragments;
  const long long L = raw.Fr, z = raw.zeroth_frame;

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
  int m = 0;
#pragma unroll 1
  for (int base = 0; base < n_all; base += 32) {
    const int j = base + lane;
    const bool keep = j < n_all && (!cfg.exclude_shared_ions || lib.frag_cardinality[fs + j] <= 1);
    const unsigned b = __ballot_sync(FULL, keep);
    if (keep) {
      const int u = m + __popc(b & ((1u << lane) - 1u));
      sm.t_src[u] = j;
      sm.t_int[u] = lib.frag_intensity[fs + j];
      sm.t_mz[u] = lib.frag_mz[fs + j];
    }
    m += __popc(b);
  }
  __syncwarp();
  const int F = min(min(m, K), SC4_F);
  if (m > ADB_NUMBA_SMALL_SORT) {  // more than 15 elements: numba's quicksort order among ties (adb_argsort_numba), one lane
    if (lane == 0) {
      uint8_t ord[ADB_MAX_LIB_FRAGMENTS];
      adb_argsort_numba(sm.t_int, m, ord);
      for (int r = 0; r < F; r++) sm.t_sel[r] = ord[m - 1 - r];
    }
  } else {
#pragma unroll 1
    for (int u = lane; u < m; u += 32) {  // descending-intensity position (stable argsort reversed)
      const float v = sm.t_int[u];
      int rank_asc = 0;
#pragma unroll 1
      for (int q = 0; q < m; q++) rank_asc += (sm.t_int[q] < v) || (sm.t_int[q] == v && q < u);
      const int r = m - 1 - rank_asc;
      if (r < F) sm.t_sel[r] = u;
    }
  }
  __syncwarp();
  if (F > ADB_NUMBA_SMALL_SORT) {  // m/z order of the selected fragments, same rule; sorted_idx is free until the fragment features
    if (lane == 0) {
      float sel_mz[SC4_F];
      uint8_t ord[SC4_F];
      for (int r = 0; r < F; r++) sel_mz[r] = sm.t_mz[sm.t_sel[r]];
      adb_argsort_numba(sel_mz, F, ord);
      for (int k = 0; k < F; k++) sm.sorted_idx[ord[k]] = k;
    }
    __syncwarp();
  }
#pragma unroll 1
  for (int r = lane; r < F; r += 32) {  // stable m/z order among the selected
    const int u = sm.t_sel[r];
    const float v = sm.t_mz[u];
    int rank2 = 0;
    if (F > ADB_NUMBA_SMALL_SORT) {
      rank2 = sm.sorted_idx[r];
    } else {
#pragma unroll 1
      for (int q = 0; q < F; q++) { const float vq = sm.t_mz[sm.t_sel[q]]; rank2 += (vq < v) || (vq == v && q < r); }
    }
    const int64_t g = fs + sm.t_src[u];
    sm.mz_library[rank2] = lib.frag_mz_library[g];
    sm.mz[rank2] = v;
    sm.intensity[rank2] = sm.t_int[u];
    sm.type[rank2] = lib.frag_type[g];
    sm.loss_type[rank2] = lib.frag_loss_type[g];
    sm.charge[rank2] = lib.frag_charge[g];
    sm.number[rank2] = lib.frag_number[g];
    sm.position[rank2] = lib.frag_position[g];
  }
  __syncwarp();
  if (F <= 3) return;

  // ---- candidate validity (the reference trusts its own selection output; rows outside the file stay invalid) ----
  const long long S64 = scan_stop - scan_start;
  if (S64 <= 0 || scan_start < 0 || scan_stop > raw.Sc || scan_center < 0 || scan_center >= raw.Sc) return;
  if (frame_start < 0 || frame_stop < 1 || frame_stop > raw.n_frames || frame_center < 0 || frame_center >= raw.n_frames ||
      frame_stop <= frame_start)
    return;
  const long long cs = (frame_start - z) / L;
  const long long C64 = (frame_stop - z) / L - cs;
  if (C64 <= 0) return;
  if (frame_start + (C64 - 1) * L >= raw.n_frames) return;
  const int S = (int)S64, C = (int)C64;
  const int nSC = S * C;

  // ---- windows (jitclasses/utils.py:15-20 with a float32 tolerance) -> tof slices (bruker_jit.py:273-278) ---------
  if (lane < F) {
    const float mz = sm.mz[lane];
    const double d = (double)__fmul_rn(cfg.fragment_mz_tolerance, mz) / 1000000.0;
    const float lo = (float)((double)mz - d), hi = (float)((double)mz + d);
    sm.t0[lane] = (int)adb_lower_bound_f64(raw.mz_values, raw.n_tof, (double)lo);
    sm.t1[lane] = (int)adb_lower_bound_f64(raw.mz_values, raw.n_tof, (double)hi);
  }
  if (lane < nI) {
    const float mz = sm.iso_mz[lane];
    const double d = (double)__fmul_rn(cfg.precursor_mz_tolerance, mz) / 1000000.0;
    const float lo = (float)((double)mz - d), hi = (float)((double)mz + d);
    sm.t0p[lane] = (int)adb_lower_bound_f64(raw.mz_values, raw.n_tof, (double)lo);
    sm.t1p[lane] = (int)adb_lower_bound_f64(raw.mz_values, raw.n_tof, (double)hi);
  }
  __syncwarp();
  float mn = sm.iso_mz[0], mx = sm.iso_mz[0];
#pragma unroll 1
  for (int i = 1; i < nI; i++) { mn = fminf(mn, sm.iso_mz[i]); mx = fmaxf(mx, sm.iso_mz[i]); }
  const float q0 = (float)((double)mn - 0.5), q1 = (float)((double)mx + 0.5);  // candidate.py:203-205

  // ---- observation ids present in the push query: np.unique(precursor_index) (bruker_jit.py:315-368) ---------------
  unsigned long long mask_f = 0ull, mask_p = 0ull;
  {
    const int fic0 = fic_of(frame_start, z, L);
    const long long span = frame_stop - frame_start;
    int bad = 0;
#pragma unroll 1
    for (long long t = lane; t < L * S; t += 32) {
      const int fic = (int)(t / S);
      const long long scan = scan_start + (t - (long long)fic * S);
      int d = fic - fic0;
      if (d < 0) d += (int)L;
      if ((long long)d >= span) continue;  // no frame of the window has this frame-in-cycle index
      const long long pos = (long long)fic * raw.Sc + scan;
      const double wlo = __ldg(raw.cycle + 2 * pos), whi = __ldg(raw.cycle + 2 * pos + 1);
      const long long id = __ldg(raw.dia_precursor_cycle + pos);
      if (id < 0) continue;
      const bool mf = ((double)q0 <= whi) && ((double)q1 >= wlo), mp = (-1.0 <= whi) && (-1.0 >= wlo);
      if ((mf || mp) && id >= 64) { bad = 1; continue; }
      if (mf) mask_f |= 1ull << id;
      if (mp) mask_p |= 1ull << id;
    }
    for (int off = 16; off > 0; off >>= 1) {
      mask_f |= __shfl_xor_sync(FULL, mask_f, off);
      mask_p |= __shfl_xor_sync(FULL, mask_p, off);
      bad |= __shfl_xor_sync(FULL, bad, off);
    }
    if (bad) { if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_OBS); return; }
  }
  const int nobs = __popcll(mask_f), nobsP = __popcll(mask_p);
  if (nobs == 0 || nobsP == 0) return;  // candidate.py:230-237
  if (nobs > ADB_MAX_OBS4 || nobsP > ADB_MAX_OBS4) { if (lane == 0) atomicOr(P.status, ADB_STATUS_TOO_MANY_OBS); return; }
  if (lane == 0) {
    int o = 0;
    for (int id = 0; id < 64; id++) if ((mask_f >> id) & 1ull) sm.obs_id[o++] = id;
  }
  __syncwarp();
  {  // quadrupole.py:105-110 indexes cycle[0, observation id, scan]: ids must be frame-in-cycle indices
    bool oob = false;
    for (int o = 0; o < nobs; o++) oob |= sm.obs_id[o] >= (int)L;
    if (oob) return;
  }

  // ---- scratch carve-up ----------------------------------------------------------------------
  const long long need = score4d_need(F, nI, nobs, nobsP, S, C);
  if (ws == nullptr || need > P.ws_floats_per_warp) {
    if (lane == 0) atomicOr(P.status, ADB_STATUS_SCRATCH_OVERFLOW);
    return;
  }
  const long long nFC = (long long)F * nobs * nSC;
  const int mxSC = max(S, C);
  float* dfi = ws;
  float* dfm = dfi + nFC;
  float* pri = dfm + nFC;
  float* prm = pri + (long long)nI * nobsP * nSC;
  float* dpi = prm + (long long)nI * nobsP * nSC;
  float* dpm = dpi + (long long)nI * nSC;
  float* tmpl = dpm + (long long)nI * nSC;
  float* qmask = tmpl + (long long)nobs * nSC;
  float* fsp_pre = qmask + nobs * S;
  float* fsp = fsp_pre + F * nobs * S;
  float* ffp = fsp + F * nobs * S;
  float* tfp_pre = ffp + F * nobs * C;
  float* tfp = tfp_pre + nobs * C;
  float* tsp_pre = tfp + nobs * C;
  float* tsp = tsp_pre + nobs * S;
  float* bp = tsp + nobs * S;
  float* isl = bp + F * C;
  float* nrm = isl + F * C;
  float* med = nrm + F * C;
  float* cen = med + C;
  float* red = cen + F * mxSC;
  float* ct = red + F * F;
  float* sfi = ct + nobs * F;
  float* after = sfi + nobs * F;
  after += ((uintptr_t)after & 7u) ? 1 : 0;
  double* qtf = (double*)after;
  double* ofmz = qtf + (long long)nI * nobs * S;
  double* ofh = ofmz + F * nobs;

  // ---- candidate.py:216-246 cubes -------------------------------------------------------------
  {
    const long long nz = 2 * nFC + 2LL * nI * nobsP * nSC;  // dfi, dfm, pri, prm are contiguous
#pragma unroll 1
    for (long long t = lane; t < nz; t += 32) dfi[t] = 0.f;
  }
  __syncwarp();
  extract4d(raw, sm.t0, sm.t1, F, nobs, mask_f, 1, frame_start, frame_stop, scan_start, S, C, cs, q0, q1, dfi, dfm, lane);
  extract4d(raw, sm.t0p, sm.t1p, nI, nobsP, mask_p, 2, frame_start, frame_stop, scan_start, S, C, cs, -1.0f, -1.0f, pri, prm, lane);
  __syncwarp();
  // ---- candidate.py:248-269 collapse the MS1 observations ----------------------------------------
#pragma unroll 1
  for (int t = lane; t < nI * nSC; t += 32) {
    const int i = t / nSC, cell = t - i * nSC;
    float s32 = 0.f;
    double smz = 0.0;
    int count = 0;
#pragma unroll 1
    for (int j = 0; j < nobsP; j++) {
      const long long idx = ((long long)i * nobsP + j) * nSC + cell;
      s32 = __fadd_rn(s32, pri[idx]);
      const float mzv = prm[idx];
      smz = __dadd_rn(smz, (double)mzv);
      count += mzv > 0.f;
    }
    dpi[t] = s32;
    dpm[t] = (float)(smz / ((double)count + 1e-6));
  }
  // ---- quadrupole.py:80-115,261-301 transfer function per (isotope, observation, scan) ---------------
#pragma unroll 1
  for (int t = lane; t < nI * nobs * S; t += 32) {
    const int i = t / (nobs * S), o = (t / S) % nobs, s = t % S;
    const long long pos = (long long)sm.obs_id[o] * raw.Sc + (scan_start + s);
    const double mu1 = __ldg(raw.cycle + 2 * pos) + cfg.quad_delta_mu[0];
    const double mu2 = __ldg(raw.cycle + 2 * pos + 1) + cfg.quad_delta_mu[1];
    const double x = (double)sm.iso_mz[i];
    const double a1 = (x - mu1) / cfg.quad_sigma[0], a2 = (x - mu2) / cfg.quad_sigma[1];
    qtf[t] = 1.0 / (1.0 + exp(-a1)) - 1.0 / (1.0 + exp(-a2));
  }
  __syncwarp();
#pragma unroll 1
  for (int t = lane; t < nobs * S; t += 32) {  // candidate.py:287-289
    double s = 0;
#pragma unroll 1
    for (int i = 0; i < nI; i++) s = __dadd_rn(s, qtf[(long long)i * nobs * S + t]);
    qmask[t] = (float)(s / (double)nI);
  }
  __syncwarp();
#pragma unroll 1
  for (long long t = lane; t < nFC; t += 32) {  // candidate.py:290
    const float v = dfi[t];
    if (v != 0.f) dfi[t] = __fmul_rn(v, qmask[(t / C) % (nobs * S)]);
  }
#pragma unroll 1
  for (int t = lane; t < nobs * nSC; t += 32) {  // quadrupole.py:304-324 template
    const int o = t / nSC, sc = t - o * nSC, s = sc / C;
    double acc = 0;
#pragma unroll 1
    for (int i = 0; i < nI; i++)
      acc = __dadd_rn(acc, __dmul_rn((double)__fmul_rn(dpi[i * nSC + sc], sm.iso_int[i]), qtf[((long long)i * nobs + o) * S + s]));
    tmpl[t] = (float)acc;
  }
  __syncwarp();
  // ---- profiles before the envelopes (scoring/utils.py:26-44): sums over cycles / over scans -----------
#pragma unroll 1
  for (int t = lane; t < nobs * S; t += 32) tsp_pre[t] = seq_sum(tmpl + (long long)t * C, C, 1);
#pragma unroll 1
  for (int t = lane; t < nobs * C; t += 32) { const int o = t / C, c = t - o * C; tfp_pre[t] = seq_sum(tmpl + (long long)o * nSC + c, S, C); }
#pragma unroll 1
  for (int t = lane; t < F * nobs * S; t += 32) fsp_pre[t] = seq_sum(dfi + (long long)t * C, C, 1);
#pragma unroll 1
  for (int t = lane; t < F * nobs * C; t += 32) { const int fo = t / C, c = t - fo * C; ffp[t] = seq_sum(dfi + (long long)fo * nSC + c, S, C); }
  __syncwarp();
  // ---- quadrupole.py:327-335 observation importance ------------------------------------------------
  if (lane < nobs) sm.sti[lane] = seq_sum(tsp_pre + lane * S, S, 1);  // sum_template_intensity
#pragma unroll 1
  for (int t = lane; t < F * nobs; t += 32) sfi[t] = seq_sum(fsp_pre + t * S, S, 1);  // sum_fragment_intensity
  __syncwarp();
  {
    float tot = 0.f;
#pragma unroll 1
    for (int o = 0; o < nobs; o++) tot = __fadd_rn(tot, sm.sti[o]);
    if (lane < nobs) sm.oi[lane] = (tot == 0.f) ? __fdiv_rn(1.0f, (float)nobs) : __fdiv_rn(sm.sti[lane], tot);
  }
  // ---- candidate.py:319-329 fragment mask --------------------------------------------------------
  bool fvalid = false;
  if (lane < F) fvalid = seq_sum(sfi + lane * nobs, nobs, 1) > 0.f;
  const unsigned vb = __ballot_sync(FULL, fvalid);
  const int Fv = __popc(vb);
  if (Fv < 2) return;
  if (fvalid) sm.fmap[__popc(vb & ((1u << lane) - 1u))] = lane;
  __syncwarp();
  const bool act = lane < Fv;  // lane w <-> masked fragment w
  const int f = act ? sm.fmap[lane] : 0;
  {  // fragment_container.py:119-120 renormalise, fragment_features.py:218
    float isum = 0.f;
#pragma unroll 1
    for (int w = 0; w < Fv; w++) isum = __fadd_rn(isum, sm.intensity[sm.fmap[w]]);
    if (act) sm.fint[lane] = __fdiv_rn(sm.intensity[f], isum);
    __syncwarp();
    float t = 0.f;
#pragma unroll 1
    for (int w = 0; w < Fv; w++) t = __fadd_rn(t, sm.fint[w]);
    if (act) sm.fin[lane] = __fdiv_rn(sm.fint[lane], t);
  }
  // ---- candidate.py:333-347 envelopes: fragment scan profile, template frame + scan profiles --------
#pragma unroll 1
  for (int t = lane; t < F * nobs * S; t += 32) { const int r = t / S; fsp[t] = or_env(fsp_pre + r * S, S, t - r * S); }
#pragma unroll 1
  for (int t = lane; t < nobs * C; t += 32) { const int r = t / C; tfp[t] = or_env(tfp_pre + r * C, C, t - r * C); }
#pragma unroll 1
  for (int t = lane; t < nobs * S; t += 32) { const int r = t / S; tsp[t] = or_env(tsp_pre + r * S, S, t - r * S); }

  float* fa = sm.feat;
#pragma unroll 1
  for (int t = lane; t < ADB_NUM_FEATURES; t += 32) fa[t] = 0.f;
  __syncwarp();
  const double rt_width = __ldg(raw.rt_values + frame_stop - 1) - __ldg(raw.rt_values + frame_start);
  const double mobility_width = __ldg(raw.mobility_values + scan_start) - __ldg(raw.mobility_values + scan_stop - 1);
  if (lane == 0) {
    fa[28] = (float)((double)Fv / (double)F);  // candidate.py:362
    // features/location_features.py:9-33 (float64 arrays in the timsTOF view)
    fa[0] = (float)mobility_width;
    fa[1] = (float)rt_width;
    fa[2] = (float)__ldg(raw.rt_values + frame_center);
    fa[3] = (float)__ldg(raw.mobility_values + scan_center);
    fa[17] = (float)nobs;
  }

  // ================= features/precursor_features.py:14-102 =================
  if (lane < nI) {
    float t_s = 0.f;
#pragma unroll 1
    for (int s = 0; s < S; s++) t_s = __fadd_rn(t_s, seq_sum(dpi + (long long)lane * nSC + s * C, C, 1));
    float wsp = 0.f;
#pragma unroll 1
    for (int o = 0; o < nobs; o++) wsp = __fadd_rn(wsp, __fmul_rn(t_s, sm.oi[o]));
    sm.spi[lane] = t_s;
    sm.wspi[lane] = wsp;
  }
  if (lane < 2 * nI) {  // precursor_features.py:52-65: the "centres" are the sizes (n_scans, n_observations = 1)
    const int i = lane >> 1;
    const double v = wcm4(((lane & 1) ? dpm : dpi) + (long long)i * nSC, S, C, (double)S, 1.0);
    if (lane & 1) sm.MZo[i] = v; else sm.H[i] = v;
  }
  __syncwarp();
  if (lane == 0) {
    int amax = 0;
#pragma unroll 1
    for (int i = 1; i < nI; i++) if (sm.iso_int[i] > sm.iso_int[amax]) amax = i;
    fa[4] = sm.wspi[0];
    fa[5] = sm.wspi[amax];
    float t6 = 0.f, t7 = 0.f;
#pragma unroll 1
    for (int i = 0; i < nI; i++) { t6 = __fadd_rn(t6, sm.wspi[i]); t7 = __fadd_rn(t7, __fmul_rn(sm.wspi[i], sm.iso_int[i])); }
    fa[6] = t6; fa[7] = t7;
    double wme = 0;
#pragma unroll 1
    for (int i = 0; i < nI; i++) if (sm.MZo[i] > 0) {
      const double me = (sm.MZo[i] - (double)sm.iso_mz[i]) / (double)sm.iso_mz[i] * 1e6;
      wme = __dadd_rn(wme, __dmul_rn(me, (double)sm.iso_int[i]));
    }
    fa[8] = (float)wme;
    fa[9] = (float)fabs(wme);
    fa[10] = (float)__dadd_rn((double)sm.iso_mz[0], __dmul_rn(__dmul_rn(wme, 1e-6), (double)sm.iso_mz[0]));
    fa[11] = (float)sm.H[0];
    fa[12] = (float)sm.H[amax];
    double t13 = 0, t14 = 0;
#pragma unroll 1
    for (int i = 0; i < nI; i++) { t13 = __dadd_rn(t13, sm.H[i]); t14 = __dadd_rn(t14, __dmul_rn(sm.H[i], (double)sm.iso_int[i])); }
    fa[13] = (float)t13; fa[14] = (float)t14;
    const double hbar = t13 / (double)nI;
    float sx = 0.f, sy = 0.f;
#pragma unroll 1
    for (int i = 0; i < nI; i++) { sx = __fadd_rn(sx, sm.iso_int[i]); sy = __fadd_rn(sy, sm.spi[i]); }
    const double xbar = (double)sx / (double)nI, ybar = (double)sy / (double)nI;
    double num = 0, sxx = 0, syy = 0, num2 = 0, shh = 0;
#pragma unroll 1
    for (int i = 0; i < nI; i++) {
      const double a = (double)sm.iso_int[i] - xbar, b = (double)sm.spi[i] - ybar, h = sm.H[i] - hbar;
      num = __dadd_rn(num, __dmul_rn(a, b)); sxx = __dadd_rn(sxx, __dmul_rn(a, a)); syy = __dadd_rn(syy, __dmul_rn(b, b));
      num2 = __dadd_rn(num2, __dmul_rn(a, h)); shh = __dadd_rn(shh, __dmul_rn(h, h));
    }
    fa[15] = (float)(num / (sqrt(sxx * syy) + 1e-12));
    fa[16] = (float)(num2 / (sqrt(sxx * shh) + 1e-12));
  }

  // ================= features/fragment_features.py:198-427 =================
  if (lane < nobs) {  // fragment_features.py:20-49 centre of mass of the template
    const float* r = tmpl + (long long)lane * nSC;
    double isum = 0, ssum = 0, fsum = 0;
    bool any = false;
#pragma unroll 1
    for (int t = 0; t < nSC; t++) { const float v = r[t]; if (v > 0.f) { any = true; isum = __dadd_rn(isum, (double)v); } }
    if (any)
#pragma unroll 1
      for (int s = 0; s < S; s++)
#pragma unroll 1
        for (int c = 0; c < C; c++) {
          const float v = r[s * C + c];
          if (v > 0.f) { ssum = __dadd_rn(ssum, __dmul_rn((double)s, (double)v)); fsum = __dadd_rn(fsum, __dmul_rn((double)c, (double)v)); }
        }
    sm.esc[lane] = (any && isum > 0) ? ssum / isum : 0.0;
    sm.efc[lane] = (any && isum > 0) ? fsum / isum : 0.0;
  }
  __syncwarp();
  int best_obs = 0;
#pragma unroll 1
  for (int o = 1; o < nobs; o++) if (sm.oi[o] > sm.oi[best_obs]) best_obs = o;
  const bool quant_all = cfg.quant_all != 0;
  // weighted-centre height and m/z of every (masked fragment, observation) cell (fragment_features.py:287-310);
  // a height is > 0 exactly when the cell has signal, i.e. when its f32 sum is > 0
#pragma unroll 1
  for (int t = lane; t < 2 * Fv * nobs; t += 32) {
    const int ch = t & 1, wo = t >> 1, w = wo / nobs, o = wo - w * nobs;
    const int ff = sm.fmap[w];
    double v = 0.0;
    if (sfi[ff * nobs + o] > 0.f) v = wcm4((ch ? dfm : dfi) + ((long long)ff * nobs + o) * nSC, S, C, sm.esc[o], sm.efc[o]);
    (ch ? ofmz : ofh)[w * nobs + o] = v;
  }
  long long qw = (long long)cfg.quant_window;
  if ((C / 2) - 1 < qw) qw = (C / 2) - 1;
  const int center = C / 2;
  int w0 = center - (int)qw, w1 = center + (int)qw + 1;
  if (qw < 0) { w0 = 0; w1 = 0; }
  if (w1 > C) w1 = C;
  if (w0 < 0) w0 = 0;
  const int wn = max(w1 - w0, 0);
  __syncwarp();
  bool anyh = false;
  if (act) {
    float* b = bp + lane * C;
    float* d = ffp + (long long)f * nobs * C;
    if (quant_all) {
#pragma unroll 1
      for (int c = 0; c < C; c++) b[c] = seq_sum(d + c, nobs, C);
    } else {
#pragma unroll 1
      for (int c = 0; c < C; c++) b[c] = d[best_obs * C + c];
    }
    center_envelope(b, C);
    if (!quant_all)  // fragment_features.py:248-250: best_profile is a VIEW, the envelope mutates the frame profile
#pragma unroll 1
      for (int c = 0; c < C; c++) d[best_obs * C + c] = b[c];
    double area = 0;  // trapezoid over the quant window on the float64 rt axis (fragment_features.py:253-273)
#pragma unroll 1
    for (int t = 0; t + 1 < wn; t++) {
      const double drt = __ldg(raw.rt_values + frame_start + (long long)(w0 + t + 1) * L) - __ldg(raw.rt_values + frame_start + (long long)(w0 + t) * L);
      const float sum2 = __fadd_rn(b[w0 + t + 1], b[w0 + t]);
      area = __dadd_rn(area, __dmul_rn(__dmul_rn((double)sum2, drt), 0.5));
    }
    sm.area_norm[lane] = __dmul_rn(area, (double)qw);
    sm.ofi[lane] = seq_sum(b + w0, wn, 1);
    // fragment_features.py:312-336
    float wsum = 0.f, fn2 = 0.f, dot = 0.f, tn2 = 0.f;
#pragma unroll 1
    for (int o = 0; o < nobs; o++) {
      const bool mm = ofh[lane * nobs + o] > 0;
      anyh |= mm;
      wsum = __fadd_rn(wsum, mm ? sm.oi[o] : 0.0f);
      const float v = sfi[f * nobs + o];
      fn2 = __fadd_rn(fn2, __fmul_rn(v, v));
      dot = __fadd_rn(dot, __fmul_rn(v, sm.sti[o]));
      tn2 = __fadd_rn(tn2, __fmul_rn(sm.sti[o], sm.sti[o]));
    }
    {  // cosine_similarity_a1, features_utils.py:40-47
      const double div = (double)__fmul_rn(sqrtf(fn2), sqrtf(tn2)) + 0.0001;
      sm.cosv[lane] = (float)((double)dot / div);
    }
    double wtot = 0;
    int cnt = 0;
#pragma unroll 1
    for (int o = 0; o < nobs; o++) {
      const double wv = (double)((ofh[lane * nobs + o] > 0) ? sm.oi[o] : 0.0f) / ((double)wsum + 1e-20);
      if (wv > 0) { wtot = __dadd_rn(wtot, wv); cnt++; }
    }
    double a = 0, bsum = 0;
    if (cnt > 0)
#pragma unroll 1
      for (int o = 0; o < nobs; o++) {
        const double wv = (double)((ofh[lane * nobs + o] > 0) ? sm.oi[o] : 0.0f) / ((double)wsum + 1e-20);
        if (wv > 0) {
          const double lw = wv / wtot;
          a = __dadd_rn(a, __dmul_rn(ofmz[lane * nobs + o], lw));
          bsum = __dadd_rn(bsum, __dmul_rn(ofh[lane * nobs + o], lw));
        }
      }
    sm.ofh_mean[lane] = bsum;
    sm.ofmm[lane] = a;
    const double mzf = (double)sm.mz[f];
    sm.mass_error[lane] = (a - mzf) / mzf * 1e6;
    const float v = sm.fint[lane];  // np.argsort(fragments.intensity)[::-1]
    int rank_asc = 0;
#pragma unroll 1
    for (int q = 0; q < Fv; q++) rank_asc += (sm.fint[q] < v) || (sm.fint[q] == v && q < lane);
    sm.sorted_idx[Fv - 1 - rank_asc] = lane;
  }
  const unsigned anyh_b = __ballot_sync(FULL, anyh);
  __syncwarp();
  if (Fv > ADB_NUMBA_SMALL_SORT) {  // more than 15 masked fragments: numba's quicksort order among ties
    if (lane == 0) {
      uint8_t ord[SC4_F];
      adb_argsort_numba(sm.fint, Fv, ord);
      for (int r = 0; r < Fv; r++) sm.sorted_idx[r] = ord[Fv - 1 - r];
    }
    __syncwarp();
  }
  // fragment-level outputs, candidate.py:403-442
  const size_t obase = (size_t)ci * (size_t)K;
  if (act && cfg.collect_fragments && lane < K) {
    P.out.fragment_mz_library[obase + lane] = sm.mz_library[f];
    P.out.fragment_mz[obase + lane] = sm.mz[f];
    P.out.fragment_mz_observed[obase + lane] = (float)sm.ofmm[lane];
    P.out.fragment_height[obase + lane] = (float)sm.ofh_mean[lane];
    P.out.fragment_intensity[obase + lane] = (float)sm.area_norm[lane];
    P.out.fragment_mass_error[obase + lane] = (float)sm.mass_error[lane];
    P.out.fragment_position[obase + lane] = sm.position[f];
    P.out.fragment_number[obase + lane] = sm.number[f];
    P.out.fragment_type[obase + lane] = sm.type[f];
    P.out.fragment_charge[obase + lane] = sm.charge[f];
    P.out.fragment_loss_type[obase + lane] = sm.loss_type[f];
  }
  if (lane == 0) {
    double sum_ofh = 0;
#pragma unroll 1
    for (int w = 0; w < Fv; w++) sum_ofh = __dadd_rn(sum_ofh, sm.ofh_mean[w]);
    if (anyh_b != 0u) fa[18] = (float)corrcoef01_4(sm.area_norm, sm.fin, Fv);
    if (sum_ofh > 0.0) fa[19] = (float)corrcoef01_4(sm.ofh_mean, sm.fin, Fv);
    int n20 = 0, n21 = 0;
    float s22 = 0.f, s23 = 0.f, cacc = 0.f;
#pragma unroll 1
    for (int w = 0; w < Fv; w++) if (sm.ofi[w] > 0.f) { n20++; s22 = __fadd_rn(s22, sm.fin[w]); cacc = __fadd_rn(cacc, sm.cosv[w]); }
#pragma unroll 1
    for (int w = 0; w < Fv; w++) if (sm.ofh_mean[w] > 0.0) { n21++; s23 = __fadd_rn(s23, sm.fin[w]); }
    fa[20] = (float)((double)n20 / (double)Fv);
    fa[21] = (float)((double)n21 / (double)Fv);
    fa[22] = s22; fa[23] = s23;
    if (n20 > 0) fa[24] = (float)((double)cacc / (double)n20);
    float sb = 0.f, sy = 0.f;
    int nb = 0, ny = 0, min_y = 255, max_b = 0;
#pragma unroll 1
    for (int w = 0; w < Fv; w++) {
      const int ty = sm.type[sm.fmap[w]], po = sm.position[sm.fmap[w]];
      if (ty == 98) { sb = __fadd_rn(sb, sm.ofi[w]); nb++; max_b = max(max_b, po); }
      if (ty == 121) { sy = __fadd_rn(sy, sm.ofi[w]); ny++; min_y = min(min_y, po); }
    }
    fa[25] = nb > 0 ? (float)log((double)sb + 1.0) : 0.f;
    fa[26] = ny > 0 ? (float)log((double)sy + 1.0) : 0.f;
    fa[27] = __fsub_rn(fa[25], fa[26]);
    const int n3 = min(Fv, 3);
    double t41 = 0, t42 = 0;
#pragma unroll 1
    for (int r = 0; r < n3; r++) t41 = __dadd_rn(t41, sm.mass_error[sm.sorted_idx[r]]);
#pragma unroll 1
    for (int w = 0; w < Fv; w++) t42 = __dadd_rn(t42, sm.mass_error[w]);
    fa[41] = (float)(t41 / (double)n3);
    fa[42] = (float)(t42 / (double)Fv);
    if (nb > 0 && ny > 0) {
      int n_ov = 0;
      double sa = 0, se = 0;
#pragma unroll 1
      for (int w = 0; w < Fv; w++) {
        const int ty = sm.type[sm.fmap[w]], po = sm.position[sm.fmap[w]];
        const bool ov = (ty == 121 && po < max_b) || (ty == 98 && po > min_y);
        if (ov) { n_ov++; sa = __dadd_rn(sa, sm.area_norm[w]); se = __dadd_rn(se, sm.mass_error[w]); }
      }
      fa[43] = (float)n_ov;
      if (n_ov > 0) { fa[44] = (float)(sa / (double)n_ov); fa[45] = (float)(se / (double)n_ov); }
      else { fa[44] = 0.f; fa[45] = 15.f; }
    }
  }
  __syncwarp();

  // ================= features/fragment_features.py:430-480 fragment_mobility_correlation =================
  {
    bool zin = false;
    if (act) {  // fragments whose (enveloped) scan profile has signal
      float t_o = 0.f;
#pragma unroll 1
      for (int o = 0; o < nobs; o++) t_o = __fadd_rn(t_o, seq_sum(fsp + ((long long)f * nobs + o) * S, S, 1));
      zin = t_o > 0.f;
    }
    const unsigned zb = __ballot_sync(FULL, zin);
    const int nz = __popc(zb);
    if (zin) sm.zidx[__popc(zb & ((1u << lane) - 1u))] = lane;  // index among the masked fragments
    __syncwarp();
    if (nz >= 3) {
      const bool za = lane < nz;
      const int zw = za ? sm.zidx[lane] : 0, zf = sm.fmap[zw];
      {
        float t = 0.f;
#pragma unroll 1
        for (int a = 0; a < nz; a++) t = __fadd_rn(t, sm.fint[sm.zidx[a]]);
        if (za) sm.znorm[lane] = __fdiv_rn(sm.fint[zw], t);
      }
#pragma unroll 1
      for (int t = lane; t < nz * nz; t += 32) red[t] = 0.f;
#pragma unroll 1
      for (int o = 0; o < nobs; o++) {  // scoring/utils.py:513-571 fragment_correlation on the scan profiles
        __syncwarp();
        if (za) {
          const float* r = fsp + ((long long)zf * nobs + o) * S;
          const float mean = __fdiv_rn(seq_sum(r, S, 1), (float)S);
          float ss = 0.f;
          float* cr = cen + lane * mxSC;
#pragma unroll 1
          for (int s = 0; s < S; s++) { const float cv = __fsub_rn(r[s], mean); cr[s] = cv; ss = __fadd_rn(ss, __fmul_rn(cv, cv)); }
          sm.rfw[lane] = sqrtf(__fdiv_rn(ss, (float)S));
        }
        __syncwarp();
#pragma unroll 1
        for (int t = lane; t < nz * nz; t += 32) {
          const int a = t / nz, b = t - a * nz;
          const float* ca = cen + a * mxSC;
          const float* cb = cen + b * mxSC;
          float dot = 0.f;
#pragma unroll 1
          for (int s = 0; s < S; s++) dot = __fadd_rn(dot, __fmul_rn(ca[s], cb[s]));
          const float cov = __fdiv_rn(dot, (float)S);
          const float smx = __fmul_rn(sm.rfw[a], sm.rfw[b]);
          const float corr = (float)((double)cov / ((double)smx + 1e-12));
          red[t] = __fadd_rn(red[t], __fmul_rn(corr, sm.oi[o]));
        }
      }
      __syncwarp();
      if (za) {
        float t = 0.f;
#pragma unroll 1
        for (int b = 0; b < nz; b++) t = __fadd_rn(t, __fmul_rn(red[lane * nz + b], sm.znorm[b]));
        sm.acc_a[lane] = t;
      }
      // template scan correlation (scoring/utils.py:574-647 against the template scan profile)
      if (za) sm.acc_b[lane] = 0.f;
#pragma unroll 1
      for (int o = 0; o < nobs; o++) {
        const float* y = tsp + o * S;
        const float ym = __fdiv_rn(seq_sum(y, S, 1), (float)S);
        float yss = 0.f;
#pragma unroll 1
        for (int s = 0; s < S; s++) { const float yc = __fsub_rn(y[s], ym); yss = __fadd_rn(yss, __fmul_rn(yc, yc)); }
        const float ystd = sqrtf(__fdiv_rn(yss, (float)S));
        if (za) {
          const float* xr = fsp + ((long long)zf * nobs + o) * S;
          const float xm = __fdiv_rn(seq_sum(xr, S, 1), (float)S);
          float xss = 0.f, dot = 0.f;
#pragma unroll 1
          for (int s = 0; s < S; s++) { const float xc = __fsub_rn(xr[s], xm); xss = __fadd_rn(xss, __fmul_rn(xc, xc)); }
#pragma unroll 1
          for (int s = 0; s < S; s++) dot = __fadd_rn(dot, __fmul_rn(__fsub_rn(xr[s], xm), __fsub_rn(y[s], ym)));
          const float xstd = sqrtf(__fdiv_rn(xss, (float)S));
          const float cov = __fdiv_rn(dot, (float)S);
          const float c1 = (float)((double)cov / ((double)__fmul_rn(xstd, ystd) + 1e-12));
          sm.acc_b[lane] = __fadd_rn(sm.acc_b[lane], __fmul_rn(c1, sm.oi[o]));
        }
      }
      __syncwarp();
      if (lane == 0) {
        float lsum = 0.f, t30 = 0.f;
#pragma unroll 1
        for (int a = 0; a < nz; a++) { lsum = __fadd_rn(lsum, sm.acc_a[a]); t30 = __fadd_rn(t30, __fmul_rn(sm.acc_b[a], sm.znorm[a])); }
        fa[29] = (float)((double)lsum / (double)nz);
        fa[30] = t30;
      }
    }
    __syncwarp();
  }

  // ================= features/profile_features.py:18-206 =================
  const float* ffr = ffp + (long long)f * nobs * C;  // this lane's frame profiles [nobs][C]
  if (cfg.experimental_xic) {
    int a0 = center - 1, a1 = center + 2;  // scoring_utils.py:100-110 python slice semantics
    if (a0 < 0) { a0 += C; if (a0 < 0) a0 = 0; }
    if (a1 > C) a1 = C;
    const int wnn = max(a1 - a0, 0);
    if (act) {
      float* is = isl + lane * C;
#pragma unroll 1
      for (int c = 0; c < C; c++) is[c] = seq_sum(ffr + c, nobs, C);
      const double cint = (double)seq_sum(is + a0, wnn, 1) / (double)wnn;
      float* nr = nrm + lane * C;
#pragma unroll 1
      for (int c = 0; c < C; c++) nr[c] = (cint > 0) ? (float)((double)is[c] / cint) : 0.f;
    }
    __syncwarp();
#pragma unroll 1
    for (int c = lane; c < C; c += 32) {  // median over fragments (scoring_utils.py:127-152)
      float vlo = 0.f, vhi = 0.f;
#pragma unroll 1
      for (int w = 0; w < Fv; w++) {
        const float v = nrm[w * C + c];
        int rk = 0;
#pragma unroll 1
        for (int u = 0; u < Fv; u++) { const float vu = nrm[u * C + c]; rk += (vu < v) || (vu == v && u < w); }
        if (rk == (Fv - 1) / 2) vlo = v;
        if (rk == Fv / 2) vhi = v;
      }
      med[c] = (Fv & 1) ? vhi : (float)((double)__fadd_rn(vlo, vhi) / 2);
    }
    __syncwarp();
    const double mxv = (double)seq_sum(med, C, 1) / (double)C;  // correlation_coefficient, scoring_utils.py:20-76
    double varx = 0;
#pragma unroll 1
    for (int c = 0; c < C; c++) { const double dd = (double)med[c] - mxv; varx = __dadd_rn(varx, __dmul_rn(dd, dd)); }
    varx /= (double)C;
    if (act) {
      const float* is = isl + lane * C;
      const float myv = (float)((double)seq_sum(is, C, 1) / (double)C);
      double cov = 0;
      float vy32 = 0.f;
#pragma unroll 1
      for (int c = 0; c < C; c++) {
        const float ym = __fsub_rn(is[c], myv);
        cov = __dadd_rn(cov, __dmul_rn((double)med[c] - mxv, (double)ym));
        vy32 = __fadd_rn(vy32, __fmul_rn(ym, ym));
      }
      cov /= (double)C;
      const double vxy = varx * ((double)vy32 / (double)C);
      sm.corr_list[lane] = (vxy == 0) ? 0.f : (float)(cov / sqrt(vxy));
    }
    __syncwarp();
    if (lane == 0) {
      const int n3 = min(Fv, 3);
      float t = 0.f;
#pragma unroll 1
      for (int r = 0; r < n3; r++) t = __fadd_rn(t, sm.corr_list[sm.sorted_idx[r]]);
      fa[32] = (float)((double)t / (double)n3);
    }
  } else {
    // legacy: observation-weighted F x F correlation matrix (scoring/utils.py:513-571), float32
#pragma unroll 1
    for (int t = lane; t < Fv * Fv; t += 32) red[t] = 0.f;
#pragma unroll 1
    for (int o = 0; o < nobs; o++) {
      __syncwarp();
      if (act) {
        const float* r = ffr + o * C;
        const float mean = __fdiv_rn(seq_sum(r, C, 1), (float)C);
        float ss = 0.f;
        float* cr = cen + lane * mxSC;
#pragma unroll 1
        for (int c = 0; c < C; c++) { const float cv = __fsub_rn(r[c], mean); cr[c] = cv; ss = __fadd_rn(ss, __fmul_rn(cv, cv)); }
        sm.rfw[lane] = sqrtf(__fdiv_rn(ss, (float)C));
      }
      __syncwarp();
#pragma unroll 1
      for (int t = lane; t < Fv * Fv; t += 32) {
        const int a = t / Fv, b = t - a * Fv;
        const float* ca = cen + a * mxSC;
        const float* cb = cen + b * mxSC;
        float dot = 0.f;
#pragma unroll 1
        for (int c = 0; c < C; c++) dot = __fadd_rn(dot, __fmul_rn(ca[c], cb[c]));
        const float cov = __fdiv_rn(dot, (float)C);
        const float smx = __fmul_rn(sm.rfw[a], sm.rfw[b]);
        const float corr = (float)((double)cov / ((double)smx + 1e-12));
        red[t] = __fadd_rn(red[t], __fmul_rn(corr, sm.oi[o]));
      }
    }
    __syncwarp();
    if (act) {
      float t = 0.f;
#pragma unroll 1
      for (int g = 0; g < Fv; g++) t = __fadd_rn(t, __fmul_rn(red[lane * Fv + g], sm.fint[g]));
      sm.corr_list[lane] = t;
    }
    __syncwarp();
    if (lane == 0) {
      const int n3 = min(Fv, 3);
      float t = 0.f;
#pragma unroll 1
      for (int a = 0; a < n3; a++)
#pragma unroll 1
        for (int b = 0; b < n3; b++) t = __fadd_rn(t, red[sm.sorted_idx[a] * Fv + sm.sorted_idx[b]]);
      fa[32] = (float)((double)t / (double)(n3 * n3));
    }
  }
  // template correlation, cycle / mobility fwhm, frame peak — lane w <-> masked fragment w
  if (act) { sm.acc_a[lane] = 0.f; sm.acc_b[lane] = 0.f; sm.rfw[lane] = 0.f; }
  if (lane == 0) sm.delta_acc = 0.0;
#pragma unroll 1
  for (int o = 0; o < nobs; o++) {
    const float* y = tfp + o * C;
    const float ym = __fdiv_rn(seq_sum(y, C, 1), (float)C);
    float yss = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; c++) { const float yc = __fsub_rn(y[c], ym); yss = __fadd_rn(yss, __fmul_rn(yc, yc)); }
    const float ystd = sqrtf(__fdiv_rn(yss, (float)C));
    if (act) {
      const float* x = ffr + o * C;
      float mxv = x[0];
      int am = 0;
#pragma unroll 1
      for (int c = 1; c < C; c++) if (x[c] > mxv) { am = c; mxv = x[c]; }
      const float xm = __fdiv_rn(seq_sum(x, C, 1), (float)C);
      float xss = 0.f, dot = 0.f;
      int na = 0;
      const double half = (double)mxv / 2;
#pragma unroll 1
      for (int c = 0; c < C; c++) { const float xc = __fsub_rn(x[c], xm); xss = __fadd_rn(xss, __fmul_rn(xc, xc)); }
#pragma unroll 1
      for (int c = 0; c < C; c++) {
        dot = __fadd_rn(dot, __fmul_rn(__fsub_rn(x[c], xm), __fsub_rn(y[c], ym)));
        na += (double)x[c] > half;
      }
      const float xstd = sqrtf(__fdiv_rn(xss, (float)C));
      const float cov = __fdiv_rn(dot, (float)C);
      const float c1 = (float)((double)cov / ((double)__fmul_rn(xstd, ystd) + 1e-12));
      const float fw = (float)(((double)na / (double)C) * rt_width);
      sm.acc_a[lane] = __fadd_rn(sm.acc_a[lane], __fmul_rn(c1, sm.oi[o]));  // profile_features.py:82-85
      sm.rfw[lane] = __fadd_rn(sm.rfw[lane], __fmul_rn(fw, sm.oi[o]));      // :142-144
      sm.frame_peak[lane] = am;
      // profile_features.py:148-188 mobility fwhm on the scan profile
      const float* xs = fsp + ((long long)f * nobs + o) * S;
      float mxs = xs[0];
#pragma unroll 1
      for (int s = 1; s < S; s++) if (xs[s] > mxs) mxs = xs[s];
      const double half_s = (double)mxs / 2;
      int nas = 0;
#pragma unroll 1
      for (int s = 0; s < S; s++) nas += (double)xs[s] > half_s;
      const float fws = (float)(((double)nas / (double)S) * mobility_width);
      sm.acc_b[lane] = __fadd_rn(sm.acc_b[lane], __fmul_rn(fws, sm.oi[o]));
    }
    __syncwarp();
    if (lane == 0) {  // median frame peak of this observation (profile_features.py:193-204)
      double vlo = 0, vhi = 0;
#pragma unroll 1
      for (int w = 0; w < Fv; w++) {
        const int v = sm.frame_peak[w];
        int rk = 0;
#pragma unroll 1
        for (int u = 0; u < Fv; u++) rk += (sm.frame_peak[u] < v) || (sm.frame_peak[u] == v && u < w);
        if (rk == (Fv - 1) / 2) vlo = (double)v;
        if (rk == Fv / 2) vhi = (double)v;
      }
      const float medp = (float)((Fv & 1) ? vhi : (vlo + vhi) / 2);
      const double delta = (double)medp - floor((double)C / 2);
      sm.delta_acc = __dadd_rn(sm.delta_acc, __dmul_rn(delta, (double)sm.oi[o]));
    }
    __syncwarp();
  }
  if (lane == 0) {
    float t31 = 0.f, t33 = 0.f, t38 = 0.f, t39 = 0.f;
#pragma unroll 1
    for (int w = 0; w < Fv; w++) {
      t31 = __fadd_rn(t31, sm.corr_list[w]);
      t33 = __fadd_rn(t33, __fmul_rn(sm.acc_a[w], sm.fint[w]));
      t38 = __fadd_rn(t38, __fmul_rn(sm.rfw[w], sm.fint[w]));
      t39 = __fadd_rn(t39, __fmul_rn(sm.acc_b[w], sm.fint[w]));
    }
    fa[31] = (float)((double)t31 / (double)Fv);
    fa[33] = t33;
    fa[38] = t38;
    fa[39] = t39;
    fa[40] = (float)sm.delta_acc;
    // profile_features.py:94-113 (the type mask indexes the sorted-index array by position)
    int nb = 0, ny = 0;
    float sb = 0.f, sy = 0.f;
#pragma unroll 1
    for (int r = 0; r < Fv; r++) {
      const int ty = sm.type[sm.fmap[r]];
      if (ty == 98) { if (nb < 3) sb = __fadd_rn(sb, sm.corr_list[sm.sorted_idx[r]]); nb++; }
      if (ty == 121) { if (ny < 3) sy = __fadd_rn(sy, sm.corr_list[sm.sorted_idx[r]]); ny++; }
    }
    if (nb > 0) { fa[34] = (float)((double)sb / (double)min(nb, 3)); fa[35] = (float)nb; }
    if (ny > 0) { fa[36] = (float)((double)sy / (double)min(ny, 3)); fa[37] = (float)ny; }
  }
  __syncwarp();
  // ---- candidate.py:475-481 ---------------------------------------------------------------------
  if (act && cfg.collect_fragments && lane < K) P.out.fragment_correlation[obase + lane] = sm.corr_list[lane];
#pragma unroll 1
  for (int t = lane; t < ADB_NUM_FEATURES; t += 32) P.out.features[(size_t)ci * ADB_NUM_FEATURES + t] = fa[t];
  if (lane == 0) P.out.valid[ci] = 1;
}

// 4 resident CTAs (16 warps) per SM: 126 registers per thread without spills; left to itself ptxas takes 178 registers and
// only 2 CTAs fit, which costs this latency-bound kernel a third of its speed (config 4: 117 -> 86 ms; 5 and 6 CTAs: same)
#ifndef SC4_MIN_CTAS
#define SC4_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(SC4_THREADS, SC4_MIN_CTAS) adb_score4d_kernel(const __grid_constant__ Score4Params P) {
  __shared__ Small4 small[SC4_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * SC4_WARPS + warp;
  const long long n_warps = (long long)gridDim.x * SC4_WARPS;
  float* ws = P.workspace ? P.workspace + (size_t)gw * (size_t)P.ws_floats_per_warp : nullptr;
  for (long long it = gw; it < P.cand.n; it += n_warps) {
    const long long ci = P.order ? (long long)P.order[it] : it;
    score_one_4d(P, ci, small[warp], ws, lane);
    __syncwarp();
  }
}

}  // namespace

int adb_score4d_resident_warps(int device) {
  int sms = 148, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adb_score4d_kernel, SC4_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm * SC4_WARPS;
}

int64_t adb_score4d_workspace_floats(int top_k, int n_iso, int64_t s_max, int64_t c_max, int nobs_cap) {
  return (score4d_need(top_k, n_iso, nobs_cap, nobs_cap, s_max, c_max) + 3) & ~(int64_t)3;
}

void adb_launch_score4d(const DevRaw4& raw, const DevLib& lib, const adb_scoring_config& cfg, DevCandidatesIn cand,
                        DevScoresOut out, float* d_workspace, int64_t workspace_floats_per_warp, int n_resident_warps,
                        int s_cap, int c_cap, const int32_t* d_order, uint32_t* d_status, cudaStream_t stream, int* n_launches) {
  (void)s_cap; (void)c_cap;
  if (cand.n <= 0) return;
  Score4Params P;
  P.raw = raw; P.lib = lib; P.cfg = cfg; P.cand = cand; P.out = out;
  P.workspace = d_workspace; P.ws_floats_per_warp = workspace_floats_per_warp; P.status = d_status; P.order = d_order;
  long long blocks = n_resident_warps / SC4_WARPS;
  const long long needed = (cand.n + SC4_WARPS - 1) / SC4_WARPS;
  if (blocks > needed) blocks = needed;
  if (blocks < 1) blocks = 1;
  adb_score4d_kernel<<<(unsigned)blocks, SC4_THREADS, 0, stream>>>(P);
  if (n_launches) (*n_launches)++;
}

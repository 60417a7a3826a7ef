// TEST INFRASTRUCTURE — runs the pass bodies of alphadia_b200/csrc/adb_score_dp.cuh thread by thread on the CPU.
//
// The scoring kernels are plain data-parallel passes (no warp-level cooperation), so the same source compiles as host
// code; tests/test_hostsim.py compares this emulation with the oracle on machines without a GPU.  The product never loads
// this library (it is not part of libalphadia_b200.so).  Built by tests/hostsim/__init__.py with the nvcc host compiler.
#include <algorithm>
#include <numeric>
#include <vector>

#include "../../alphadia_b200/csrc/adb_score_dp.cuh"

extern "C" int adb_hostsim_score(const adb_rawfile3d_desc* d, const adb_library_desc* ld, const adb_scoring_config* cfg,
                                 const adb_candidates_in* cand, adb_scores_out* out, int32_t out_k, int32_t KS, int64_t batch,
                                 const int32_t* order, int32_t nb_override, uint32_t* status_out) {
  DevRaw raw{};
  raw.cycle = d->cycle; raw.cycle_len = d->cycle_len; raw.rt_values = d->rt_values; raw.n_spectra = d->n_spectra;
  raw.mobility_values = d->mobility_values; raw.n_mobility = d->n_mobility; raw.peak_start = d->peak_start_idx;
  raw.peak_stop = d->peak_stop_idx; raw.mz = d->mz_values; raw.intensity = d->intensity_values; raw.n_peaks = d->n_peaks;
  raw.zeroth_frame = d->zeroth_frame; raw.precursor_cycle_max_index = d->precursor_cycle_max_index;
  raw.scan_max_index = d->scan_max_index; raw.frame_max_index = d->frame_max_index;
  raw.n_ms1_pos = 0;
  for (int64_t j = 0; j < d->cycle_len; j++)
    if ((-1.0 <= d->cycle[2 * j + 1]) && (-1.0 >= d->cycle[2 * j]) && raw.n_ms1_pos < ADB_MAX_MS1_POS) raw.ms1_pos[raw.n_ms1_pos++] = (int32_t)j;
  // time-blocked m/z index: peaks of every (cycle position, block of ADB_TB_CYCLES cycles), stably sorted by m/z, plus the
  // bucket table (adb_api.cu builds the same with a radix sort and binary searches); nb_override exercises several bucket widths
  const int64_t L = d->cycle_len;
  const int64_t n_cycles = (d->n_spectra + L - 1) / L;
  const int ntb = (int)std::max<int64_t>((n_cycles + ADB_TB_CYCLES - 1) / ADB_TB_CYCLES, 1);
  const int64_t n_seg = L * ntb;
  float mz_min = 3.0e38f, mz_max = 0.f;
  for (int64_t s = 0; s < d->n_spectra; s++)
    if (d->peak_stop_idx[s] > d->peak_start_idx[s]) {
      mz_min = std::min(mz_min, std::max(d->mz_values[d->peak_start_idx[s]], 0.f));
      mz_max = std::max(mz_max, std::max(d->mz_values[d->peak_stop_idx[s] - 1], 0.f));
    }
  if (!(mz_max > mz_min)) { mz_min = 0.f; mz_max = 1.f; }
  int nb = 64;
  while (nb < ADB_TB_MAX_BUCKETS && (int64_t)nb * 4 * n_seg < d->n_peaks) nb *= 2;
  if (nb_override > 0) nb = nb_override;
  raw.tb_ntb = ntb; raw.tb_nb = nb; raw.tb_lo = mz_min;
  raw.tb_width = (mz_max - mz_min) / (float)nb * 1.0001f;
  if (!(raw.tb_width > 0.f)) raw.tb_width = 1.f;
  raw.tb_inv_width = 1.0f / raw.tb_width;
  std::vector<int64_t> idx, spec_of;
  idx.reserve((size_t)d->n_peaks); spec_of.reserve((size_t)d->n_peaks);
  std::vector<uint32_t> table((size_t)n_seg * (size_t)(nb + 1));
  for (int64_t seg = 0; seg < n_seg; seg++) {
    const int64_t pos = seg / ntb, tb = seg % ntb;
    const size_t seg0 = idx.size();
    for (int64_t cyc = tb * ADB_TB_CYCLES; cyc < std::min<int64_t>((tb + 1) * ADB_TB_CYCLES, n_cycles); cyc++) {
      const int64_t s = cyc * L + pos;
      if (s >= d->n_spectra) continue;
      for (int64_t i = d->peak_start_idx[s]; i < d->peak_stop_idx[s]; i++) { idx.push_back(i); spec_of.push_back(s); }
    }
    std::vector<size_t> perm(idx.size() - seg0);
    std::iota(perm.begin(), perm.end(), (size_t)0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return d->mz_values[idx[seg0 + a]] < d->mz_values[idx[seg0 + b]]; });
    std::vector<int64_t> i2(perm.size()), s2(perm.size());
    for (size_t t = 0; t < perm.size(); t++) { i2[t] = idx[seg0 + perm[t]]; s2[t] = spec_of[seg0 + perm[t]]; }
    std::copy(i2.begin(), i2.end(), idx.begin() + seg0);
    std::copy(s2.begin(), s2.end(), spec_of.begin() + seg0);
    uint32_t* tab = table.data() + (size_t)seg * (size_t)(nb + 1);
    size_t cur = seg0;
    tab[0] = (uint32_t)seg0;
    for (int b = 1; b < nb; b++) {
      const float edge = adb_tb_edge(raw, b);
      while (cur < idx.size() && d->mz_values[idx[cur]] < edge) cur++;
      tab[b] = (uint32_t)cur;
    }
    tab[nb] = (uint32_t)idx.size();
  }
  std::vector<float4> s_pk(idx.size() + 16);
  for (size_t t = 0; t < s_pk.size(); t++) {
    union { uint32_t u; float f; } cyc;
    cyc.u = t < idx.size() ? (uint32_t)(spec_of[t] / L) : 0xFFFFFFFFu;
    s_pk[t] = t < idx.size() ? make_float4(d->mz_values[idx[t]], d->intensity_values[idx[t]], cyc.f, 0.f) : make_float4(3.0e38f, 0.f, cyc.f, 0.f);
  }
  raw.tb_pk = s_pk.data(); raw.tb_bucket = table.data();

  DevLib lib{};
  lib.n_precursors = ld->n_precursors; lib.precursor_idx = ld->precursor_idx; lib.frag_start_idx = ld->frag_start_idx;
  lib.frag_stop_idx = ld->frag_stop_idx; lib.charge = ld->charge; lib.rt = ld->rt; lib.mobility = ld->mobility; lib.mz = ld->mz;
  lib.isotopes = ld->isotopes; lib.n_isotopes = ld->n_isotopes; lib.n_fragments = ld->n_fragments;
  lib.frag_mz_library = ld->frag_mz_library; lib.frag_mz = ld->frag_mz; lib.frag_intensity = ld->frag_intensity;
  lib.frag_type = ld->frag_type; lib.frag_loss_type = ld->frag_loss_type; lib.frag_charge = ld->frag_charge;
  lib.frag_number = ld->frag_number; lib.frag_position = ld->frag_position; lib.frag_cardinality = ld->frag_cardinality;

  DpParams P{};
  P.raw = raw; P.lib = lib; P.cfg = *cfg;
  P.cand = DevCandidatesIn{cand->n, cand->lib_row, cand->rank, cand->scan_start, cand->scan_stop, cand->scan_center,
                           cand->frame_start, cand->frame_stop, cand->frame_center};
  P.out = DevScoresOut{out->features, out->valid, out->fragment_mz_library, out->fragment_mz, out->fragment_mz_observed,
                       out->fragment_height, out->fragment_intensity, out->fragment_mass_error, out->fragment_correlation,
                       out->fragment_position, out->fragment_number, out->fragment_type, out->fragment_charge, out->fragment_loss_type};
  P.out_k = out_k; P.order = order; P.KS = KS;
  P.nIcap = (int)std::min<int64_t>(std::min<int64_t>(lib.n_isotopes, cfg->top_k_isotopes), ADB_MAX_ISOTOPES);
  uint32_t status = 0;
  P.status = &status;
  const size_t N = (size_t)dp_padded_slots(std::max<int64_t>(batch, 1));
  std::vector<uint8_t> state(N), F(N), nobs(N);
  std::vector<int32_t> C(N), cs(N);
  std::vector<uint16_t> pos(N * ADB_MAX_OBS);
  std::vector<uint32_t> fsel(N * (size_t)KS);
  std::vector<double> qtf(N * (size_t)P.nIcap * ADB_MAX_OBS);
  std::vector<float> qmask(N * ADB_MAX_OBS);
  std::vector<int64_t> need(N + 1), tile_need(N + 1), off(N + 1);
  std::vector<uint8_t> rowflag(N * (size_t)KS);
  std::vector<uint32_t> work(N * (size_t)KS);
  int32_t n_work = 0;
  P.rowflag = rowflag.data(); P.work = work.data(); P.n_work = &n_work;
  P.state = state.data(); P.F = F.data(); P.nobs = nobs.data(); P.C = C.data(); P.cs = cs.data(); P.pos = pos.data();
  P.fsel = fsel.data(); P.qtf = qtf.data(); P.qmask = qmask.data(); P.need = need.data(); P.tile_need = tile_need.data(); P.off = off.data();
  std::vector<double> wtab_p(2 * DP_WTAB_P_STRIDE);
  for (int t = 0; t < 2 * DP_WTAB_P_STRIDE; t++) wtab_p[(size_t)t] = dp_wtab_p_entry(t / DP_WTAB_P_STRIDE, t % DP_WTAB_P_STRIDE);
  P.wtab_p = wtab_p.data();
  std::vector<float> cube;
  for (int64_t base = 0; base < cand->n; base += batch) {
    P.base = base;
    P.n = std::min<int64_t>(batch, cand->n - base);
    for (int64_t j = 0; j < P.n; j++) dp_setup(P, j);
    const int64_t n_tiles = (P.n + DP_W - 1) / DP_W, n_pad = n_tiles * DP_W;
    int64_t run = 0;  // dp_tile_need_kernel + the exclusive scan
    for (int64_t t = 0; t <= n_tiles; t++) {
      int64_t m = 0;
      for (int64_t j = t * DP_W; t < n_tiles && j < std::min<int64_t>((t + 1) * (int64_t)DP_W, P.n); j++) m = std::max(m, need[(size_t)j]);
      off[(size_t)t] = run;
      run += m * DP_W;
    }
    cube.assign((size_t)run + 4, -12345.0f);  // poison: a pass that reads what no pass wrote shows up as a mismatch
    P.cube = cube.data();
    const uint32_t rows = (uint32_t)(P.KS + P.nIcap), KS32 = (uint32_t)P.KS;
    uint32_t j, r;
    for (uint32_t t = 0; t < (uint32_t)(n_pad * rows); t++) {
      dp_decode(t, rows, j, r);
      if (j < P.n) dp_extract(P, j, (int)r);
      else if (r < KS32) rowflag[dp_encode(j, r, KS32)] = 0;
    }
    for (int64_t jj = 0; jj < P.n; jj++) dp_template(P, jj);
    n_work = 0;  // cub::DeviceSelect::Flagged on the device
    for (int64_t t = 0; t < n_pad * P.KS; t++) if (rowflag[(size_t)t]) work[(size_t)n_work++] = (uint32_t)t;
    for (int32_t t = 0; t < n_work; t++) { dp_decode(work[(size_t)t], KS32, j, r); dp_fragment(P, j, (int)r); }
    if (cfg->experimental_xic)
      for (uint32_t t = 0; t < (uint32_t)(n_pad * DP_MED_LANES); t++) { dp_decode(t, DP_MED_LANES, j, r); if (j < P.n) dp_median(P, j, (int)r); }
    for (int32_t t = 0; t < n_work; t++) { dp_decode(work[(size_t)t], KS32, j, r); dp_corr(P, j, (int)r); }
    for (int64_t jj = 0; jj < P.n; jj++) dp_aggregate(P, jj);
    if (cfg->collect_fragments)
      for (uint32_t t = 0; t < (uint32_t)(n_pad * P.KS); t++) { dp_decode(t, KS32, j, r); if (j < P.n) dp_write(P, j, (int)r); }
  }
  if (status_out) *status_out = status;
  return 0;
}

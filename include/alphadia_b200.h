/*
 * alphadia_b200 — C ABI of the B200-native precursor-candidate hot path.
 *
 * This is the drop-in boundary: the three numba entry points of the reference
 * become three blocking C calls on device-resident objects.  Plain pointers and sizes only;
 * the caller owns every host buffer (in and out); the library copies in, never keeps a host
 * pointer after return, and never calls back into the host language.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   adb_rawfile3d_create   <- DiaData.to_jitclass() -> AlphaRawJIT(...)        alphadia/raw_data/alpharaw_wrapper.py:138-156,
 *                                                                               alphadia/search/jitclasses/alpharaw_jit.py:98-138
 *   adb_rawfile4d_create   <- DiaData.to_jitclass() -> TimsTOFTransposeJIT(...) alphadia/raw_data/bruker.py:119-152,
 *                                                                               alphadia/search/jitclasses/bruker_jit.py:20-137
 *   adb_library_create     <- PrecursorFlatContainer / FragmentContainer       alphadia/search/selection/config_df.py:184-223,
 *                                                                               alphadia/search/jitclasses/fragment_container.py:12-45
 *   adb_select_candidates  <- _select_candidates_pjit(range(n), ...)           alphadia/search/selection/selection.py:78-203,656-666
 *   adb_score_candidates   <- _process_score_groups(range(n), ...)             alphadia/search/scoring/scoring.py:114-137,633-643
 *                             (Candidate.process                               alphadia/search/scoring/containers/candidate.py:166-481)
 *   adb_fragment_competition <- _compete_for_fragments(np.arange(n_win), ...)  alphadia/fragcomp/fragcomp.py:51-143,275-289
 *
 * Error convention: every call returns 0 on success, non-zero otherwise; adb_last_error() then
 * holds a message (thread-local).  Per-item failures follow the reference: the output row is left
 * untouched (score == 0 / valid == 0) and the call still returns 0.
 *
 * Threading: one host thread per handle.  Kernels run on a per-handle CUDA stream; calls block
 * until results are in the caller's host buffers.
 */
#ifndef ALPHADIA_B200_H
#define ALPHADIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADB_NUM_FEATURES 46 /* alphadia/constants/settings.py:5 */
#define ADB_MAX_FRAGMENTS 32 /* hard cap on top_k_fragments handled on the device */
#define ADB_MAX_ISOTOPES 8

typedef struct adb_rawfile adb_rawfile_t; /* opaque, device resident */
typedef struct adb_library adb_library_t; /* opaque, device resident */

/* ---- raw file, 3-D (RT x m/z), AlphaRawJIT field for field ------------------------------- */
typedef struct {
  const double* cycle;               /* f64 [1, cycle_len, 1, 2]  (quad lower, upper); MS1 row = (-1,-1) */
  int64_t cycle_len;                 /* L = spectra per DIA cycle */
  const float* rt_values;            /* f32 [n_spectra], seconds */
  int64_t n_spectra;
  const float* mobility_values;      /* f32 [n_mobility]; 3-D files: {1e-6, 0} */
  int64_t n_mobility;
  const int64_t* peak_start_idx;     /* i64 [n_spectra] */
  const int64_t* peak_stop_idx;      /* i64 [n_spectra] */
  const float* mz_values;            /* f32 [n_peaks], ascending inside a spectrum */
  const float* intensity_values;     /* f32 [n_peaks] */
  int64_t n_peaks;
  int64_t zeroth_frame;              /* 0/1 */
  int64_t precursor_cycle_max_index; /* n_spectra / cycle_len */
  int64_t scan_max_index;            /* 1 for 3-D files */
  int64_t frame_max_index;           /* n_spectra - 1 */
} adb_rawfile3d_desc;

/* ---- raw file, 4-D (RT x ion mobility x m/z), the TimsTOFTransposeJIT fields the hot path reads
 *      (alphadia/search/jitclasses/bruker_jit.py:20-137; CSR by tof index, built by raw_data/bruker.py:202-274) */
typedef struct {
  const double* cycle;                /* f64 [1, frames_per_cycle, scans, 2] quad window per (frame in cycle, scan) */
  int64_t frames_per_cycle;           /* cycle.shape[1] */
  int64_t scans;                      /* cycle.shape[2] */
  const int64_t* dia_precursor_cycle; /* i64 [frames_per_cycle * scans] observation id of every cycle position */
  const double* rt_values;            /* f64 [n_frames] */
  int64_t n_frames;
  const double* mobility_values;      /* f64 [scans], descending */
  const double* mz_values;            /* f64 [n_tof] */
  int64_t n_tof;
  const int64_t* tof_indptr;          /* i64 [n_tof + 1] */
  const uint32_t* push_indices;       /* u32 [n_events], ascending inside a tof row; push = frame * scan_max_index + scan */
  const uint16_t* intensity_values;   /* u16 [n_events] */
  int64_t n_events;
  int64_t zeroth_frame;               /* 0/1 */
  int64_t precursor_cycle_max_index;  /* frame_max_index / frames_per_cycle */
  int64_t scan_max_index;
  int64_t frame_max_index;
} adb_rawfile4d_desc;

/* ---- spectral library (flat), precursors sorted by precursor_idx -------------------------- */
typedef struct {
  int64_t n_precursors;
  const uint32_t* precursor_idx;   /* [P] */
  const uint32_t* frag_start_idx;  /* [P] */
  const uint32_t* frag_stop_idx;   /* [P] */
  const uint8_t* charge;           /* [P] */
  const float* rt;                 /* [P] rt_column      */
  const float* mobility;           /* [P] mobility_column */
  const float* mz;                 /* [P] precursor_mz_column */
  const float* isotopes;           /* [P, n_isotopes] row-major (i_0 .. i_k) */
  int32_t n_isotopes;
  int64_t n_fragments;
  const float* frag_mz_library;    /* [NF] */
  const float* frag_mz;            /* [NF] fragment_mz_column */
  const float* frag_intensity;     /* [NF] */
  const uint8_t* frag_type;        /* [NF] 98 = b, 121 = y */
  const uint8_t* frag_loss_type;   /* [NF] */
  const uint8_t* frag_charge;      /* [NF] */
  const uint8_t* frag_number;      /* [NF] */
  const uint8_t* frag_position;    /* [NF] */
  const uint8_t* frag_cardinality; /* [NF] */
} adb_library_desc;

/* ---- selection: CandidateSelectionConfigJIT (alphadia/search/selection/config_df.py:14-124) */
typedef struct {
  double rt_tolerance;
  double precursor_mz_tolerance;
  double fragment_mz_tolerance;
  double mobility_tolerance;
  int64_t candidate_count;
  int64_t top_k_precursors;
  int64_t top_k_fragments; /* carried, unused by the reference kernel too */
  int32_t exclude_shared_ions;
  int64_t kernel_size;
  double f_mobility;
  double f_rt;
  double center_fraction;
  int64_t min_size_mobility;
  int64_t min_size_rt;
  int64_t max_size_mobility;
  int64_t max_size_rt;
  int32_t use_weighted_score;
  int32_t join_close_candidates;
  double join_close_candidates_scan_threshold;
  double join_close_candidates_cycle_threshold;
  double feature_std;    /* feature_std[0]    */
  double feature_mean;   /* feature_mean[0]   */
  double feature_weight; /* feature_weight[0] */
} adb_selection_config;

/* CandidateContainer (config_df.py:226-254): n_precursors * candidate_count rows, caller-zeroed or not:
 * the library zero-fills them first, exactly like CandidateContainer.__init__. */
typedef struct {
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
} adb_candidates_out;

/* ---- scoring: CandidateScoringConfigJIT (alphadia/search/scoring/config.py:13-65) -------- */
typedef struct {
  int32_t collect_fragments;
  int32_t exclude_shared_ions;
  uint32_t top_k_fragments;
  uint32_t top_k_isotopes;
  uint32_t quant_window;
  int32_t quant_all;
  float precursor_mz_tolerance;
  float fragment_mz_tolerance;
  int32_t experimental_xic;
  /* SimpleQuadrupoleJit state (alphadia/search/scoring/quadrupole.py:58-78) */
  double quad_sigma[2];
  double quad_delta_mu[2];
} adb_scoring_config;

/* One row per candidate, in output order (= sorted by score_group_idx, precursor_idx). */
typedef struct {
  int64_t n;
  const int64_t* lib_row; /* row of the candidate's precursor in adb_library_desc arrays */
  const uint8_t* rank;
  const int64_t* scan_start;
  const int64_t* scan_stop;
  const int64_t* scan_center;
  const int64_t* frame_start;
  const int64_t* frame_stop;
  const int64_t* frame_center;
} adb_candidates_in;

/* OutputPsmDF (alphadia/search/scoring/output.py:18-70).  Fragment tables are [n, top_k_fragments]. */
typedef struct {
  float* features;   /* [n, 46] */
  uint8_t* valid;    /* [n] */
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
} adb_scores_out;

/* Ragged result of candidate scoring: exactly the rows OutputPsmDF.to_precursor_df / to_fragment_df keep
 * (alphadia/search/scoring/output.py:72-97, consumed by collect_candidates / collect_fragments,
 * alphadia/search/scoring/scoring.py:394-580): the feature rows of the valid candidates, in candidate order, and - flattened
 * in the same order - their fragment slots with mz_library > 0.  Only these bytes cross the bus (the dense [n, top_k]
 * tables are mostly empty slots), and top_k_fragments is not capped by ADB_MAX_FRAGMENTS here: the transfer-library
 * requantification calls with top_k_fragments = 9999
 * (alphadia/workflow/peptidecentric/transfer_library_requantification_handler.py:102-124); a candidate keeps at most
 * the fragments its precursor has.  All buffers are caller-owned host memory (pinned memory makes the copies overlap the
 * scoring kernels). */
typedef struct {
  int64_t row_capacity;  /* in: rows the row-wise buffers hold; adb_candidates_in.n always suffices */
  int64_t frag_capacity; /* in: entries the fragment buffers hold; n * min(top_k_fragments, widest precursor) always suffices */
  int64_t n_rows;        /* out: valid candidates */
  int64_t n_fragments;   /* out: fragment entries */
  int64_t* row_index;    /* [row_capacity] row of the candidate in adb_candidates_in */
  float* features;       /* [row_capacity, 46] */
  int64_t* frag_offset;  /* [row_capacity + 1] fragments of valid row r: [frag_offset[r], frag_offset[r + 1]) */
  float* fragment_mz_library; /* [frag_capacity] each */
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
} adb_scores_ragged;

/* ---- entry points ------------------------------------------------------------------------- */
const char* adb_last_error(void);
const char* adb_version(void);
int adb_device_count(void);

int adb_rawfile3d_create(const adb_rawfile3d_desc* desc, int device, adb_rawfile_t** out);
/* timsTOF raw file: DiaData.to_jitclass() -> TimsTOFTransposeJIT(...) (alphadia/raw_data/bruker.py:119-152).  The
 * same adb_select_candidates / adb_score_candidates entry points then run the 4-D kernels on the handle. */
int adb_rawfile4d_create(const adb_rawfile4d_desc* desc, int device, adb_rawfile_t** out);
void adb_rawfile_destroy(adb_rawfile_t* raw);
int64_t adb_rawfile_device_bytes(const adb_rawfile_t* raw);

int adb_library_create(const adb_library_desc* desc, int device, adb_library_t** out);
void adb_library_destroy(adb_library_t* lib);

/* kernel: GaussianKernel.get_dense_matrix() output, f32 [kernel_h, kernel_w] (selection/kernel.py:141-218) */
int adb_select_candidates(adb_rawfile_t* raw, adb_library_t* lib, const adb_selection_config* cfg,
                          const float* kernel, int32_t kernel_h, int32_t kernel_w,
                          adb_candidates_out* out);

int adb_score_candidates(adb_rawfile_t* raw, adb_library_t* lib, const adb_scoring_config* cfg,
                         const adb_candidates_in* cand, adb_scores_out* out);

/* same scoring, ragged result (see adb_scores_ragged); fails with the needed sizes in n_rows / n_fragments when a
 * capacity is too small */
int adb_score_candidates_ragged(adb_rawfile_t* raw, adb_library_t* lib, const adb_scoring_config* cfg,
                                const adb_candidates_in* cand, adb_scores_ragged* out);

/* dtype flags (the reference computes in the array dtypes): is_f64 bit 0 = rt is f64 (else f32),
 * bit 1 = fragment_mz is f64 (else f32).
 * valid: u8 [n_psm], in/out (the reference starts from all-true). */
int adb_fragment_competition(int device, int64_t n_windows, const int64_t* window_start,
                             const int64_t* window_stop, int64_t n_psm, const void* rt,
                             const int64_t* frag_start_idx, const int64_t* frag_stop_idx,
                             int64_t n_frag, const void* fragment_mz, int32_t is_f64,
                             double rt_tol_seconds, double mass_tol_ppm, uint8_t* valid);

/* Load-time CSR transpose of a timsTOF raw file, push-major -> tof-major: replaces _transpose / _transpose_chunk
 * (alphadia/raw_data/bruker.py:155-274).  In: tof_indices u32 [n_values] (column index of every event, events
 * ordered by push), push_indptr i64 [n_push + 1], values u16 [n_values].  Out (caller-allocated): push_indices u32 [n_values]
 * ascending inside every tof row, tof_indptr i64 [n_tof + 1], values u16 [n_values] — one stable device radix sort. */
int adb_transpose_csr(int device, int64_t n_values, int64_t n_push, int64_t n_tof, const uint32_t* tof_indices,
                      const int64_t* push_indptr, const uint16_t* values, uint32_t* push_indices_out,
                      int64_t* tof_indptr_out, uint16_t* values_out);

/* FDR bookkeeping around fragment competition (both called by perform_fdr, alphadia/fdr/fdr.py:157-186).
 *
 * adb_q_values replaces get_q_values + _fdr_to_q_values (fdr.py:195-297): rows are ordered like
 * df.sort_values([score, decoy, *extra_sort_columns]) (stable), fdr = cumsum(decoy) / cumsum(1 - decoy) in float64 and
 * qval = running minimum of fdr from the back.  In: score f64 [n] (no NaN), decoy u8 [n] in {0, 1}, extra_key u64 [n] < 2^63
 * (the tie-break columns packed order-preservingly, usually precursor_idx).  Out (caller-allocated): order i64 [n] = row
 * index of the i-th sorted row (df.iloc[order] is the sorted frame), qval f64 [n] in sorted order. */
int adb_q_values(int device, int64_t n, const double* score, const uint8_t* decoy, const uint64_t* extra_key,
                 int64_t* order_out, double* qval_out);
/* adb_keep_best replaces keep_best (fdr.py:195-224): per group the row with the lowest score, ties to the earliest row.
 * In: score f64 [n] (no NaN), group_key u64 [n] (equal key <=> same group).  Out: keep u8 [n], 1 for the surviving rows
 * (df[keep].reset_index(drop=True) is the reference's result). */
int adb_keep_best(int device, int64_t n, const double* score, const uint64_t* group_key, uint8_t* keep_out);

/* FDR classifier inference: BinaryClassifierLegacyNewBatching.predict_proba (alphadia/fdr/classifiers.py:441-470)
 * = FeedForwardNN.forward in eval mode (classifiers.py:473-532): BatchNorm1d(running statistics) -> [Linear -> ReLU] per hidden
 * layer -> Linear -> softmax, float32.  Weights in torch's layout (Linear.weight is [out, in]); x is [n, input_dim] and
 * proba_out [n, layer_dims[n_layers - 1]], both host memory. */
typedef struct {
  int32_t input_dim;
  int32_t n_layers;             /* linear layers, the output layer included (<= 8, widths <= 128) */
  const int32_t* layer_dims;    /* [n_layers] output width of each linear layer */
  const float* bn_weight;       /* [input_dim], NULL = 1 */
  const float* bn_bias;         /* [input_dim], NULL = 0 */
  const float* bn_mean;         /* running_mean */
  const float* bn_var;          /* running_var */
  float bn_eps;
  const float* const* weights;  /* [n_layers] pointers, each [out, in] row-major */
  const float* const* biases;   /* [n_layers] pointers, each [out] */
} adb_classifier_desc;
int adb_classifier_predict_proba(int device, const adb_classifier_desc* net, int64_t n, const float* x, float* proba_out);

/* ---- resident variants (bench.py `value`, the sharded driver) -------------------------------
 * Same kernels; the raw file and library are already in HBM, results stay in HBM inside the raw
 * handle's workspace until fetched.  (Not needed by a reference-side binding.) */
/* selection + device-side compaction of rows with score > 0 (CandidateContainer.get_candidate_df_data,
 * config_df.py:270-284) into scoring input order (library row, rank); returns the candidate count */
int adb_select_candidates_resident(adb_rawfile_t* raw, adb_library_t* lib, const adb_selection_config* cfg,
                                   const float* kernel, int32_t kernel_h, int32_t kernel_w, int64_t* n_candidates);
/* the compacted candidate table of the last adb_select_candidates_resident call: rows of the candidate container with
 * score > 0 in container order (candidate_container_to_df, config_df.py:270-298), index columns already widened to
 * int64 as Schema.validate does (validation/schemas.py:51-73), plus the library row of each precursor.  The int64
 * columns have the layout of adb_candidates_in, so the table can be handed to adb_score_candidates as it is. */
typedef struct {
  int64_t n; /* must equal the count returned by adb_select_candidates_resident */
  int64_t* lib_row;
  uint8_t* rank;
  int64_t* scan_start;
  int64_t* scan_stop;
  int64_t* scan_center;
  int64_t* frame_start;
  int64_t* frame_stop;
  int64_t* frame_center;
  uint32_t* precursor_idx;
  float* score;
} adb_candidate_table;
int adb_fetch_candidate_table(adb_rawfile_t* raw, adb_candidate_table* out);
/* Selection and scoring of one library batch in ONE call - what the workflow does back to back
 * (alphadia/workflow/peptidecentric/peptidecentric.py:196-207: extraction_handler.select_candidates(apply_cutoff=True) ->
 * extraction_handler.score_and_quantify_candidates(candidates_df, ...)): candidate selection, the `score > 0` filter of
 * candidate_container_to_df (config_df.py:270-298) and the handler's score cutoff (`score > score_cutoff` on the float32
 * column, extraction_handler.py:177-202; pass -INFINITY for none), scoring of the surviving candidates from the resident
 * table.  `table` (host, capacity in table->n, row count out) receives the candidate table while the first scoring block
 * runs and may be NULL; `out` is the ragged result of adb_score_candidates_ragged, its row_index refers to `table`. */
int adb_select_score_candidates_ragged(adb_rawfile_t* raw, adb_library_t* lib, const adb_selection_config* sel_cfg,
                                       const float* kernel, int32_t kernel_h, int32_t kernel_w, float score_cutoff,
                                       const adb_scoring_config* score_cfg, adb_candidate_table* table,
                                       adb_scores_ragged* out);
/* scores the resident candidates; outputs stay on the device */
int adb_score_candidates_resident(adb_rawfile_t* raw, adb_library_t* lib, const adb_scoring_config* cfg);
/* D2H of the resident results: the full candidate container (n_precursors * candidate_count rows) ... */
int adb_fetch_candidates(adb_rawfile_t* raw, adb_candidates_out* container);
/* ... and the score tables of the n resident candidates with their identity (lib_row / rank may be NULL) */
int adb_fetch_scores(adb_rawfile_t* raw, adb_scores_out* out, int64_t* lib_row, uint8_t* rank);
/* device pointers of the resident score table: features f32 [n, 46], valid u8 [n], lib_row i64 [n], rank u8 [n] */
int adb_resident_score_table(adb_rawfile_t* raw, void** features, void** valid, void** lib_row, void** rank,
                             int64_t* n_rows);

/* timing of the last call on this handle, CUDA events on the handle's stream (ms) + kernel launches so far */
int adb_last_timing(const adb_rawfile_t* raw, float* h2d_ms, float* kernel_ms, float* d2h_ms);
int64_t adb_kernel_launches(const adb_rawfile_t* raw);
/* duration (ms, CUDA events on the handle's stream) of the selection / scoring kernel of the last call */
float adb_last_main_kernel_ms(const adb_rawfile_t* raw);
/* CUDA stream the handle launches on (cudaStream_t as void*) */
void* adb_rawfile_stream(const adb_rawfile_t* raw);

#ifdef __cplusplus
}
#endif
#endif /* ALPHADIA_B200_H */

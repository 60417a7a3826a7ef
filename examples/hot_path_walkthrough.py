#!/usr/bin/env python
"""The hot path of one raw file through the reference-shaped operator classes (needs a B200 and the built library):

    python examples/hot_path_walkthrough.py [config1|parity_small|parity_4d|config2]

1. ``CandidateSelection(dia_data, precursors_flat, fragments_flat, config, ...)()``  -> candidates DataFrame
2. ``CandidateScoring(dia_data=..., ...)(candidates_df)``                            -> (features_df, fragments_df)
3. ``FragmentCompetition()(psm_df, fragments_df, dia_data.cycle)``                    -> surviving PSMs
4. ``perform_fdr(BinaryClassifierLegacyNewBatching(...), ...)`` (alphadia_b200.fdr)     -> best row per group with q-values
   (classifier inference, q-values, fragment competition and best-row-per-group on the device)

`dia_data` is whatever the reference passes around (an AlphaRaw / TimsTOFTranspose wrapper or its jitclass); here it is a
synthetic run from ``alphadia_b200.synthetic`` so that the script has no external inputs.  The classes, their arguments and
the returned columns are the reference's (alphadia/search/selection/selection.py:547-737, search/scoring/scoring.py:139-660,
fragcomp/fragcomp.py:146-299); see INTEGRATION.md for the ~25-line ExtractionHandler that plugs them into the workflow.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from alphadia_b200 import (  # noqa: E402
    CandidateScoring, CandidateScoringConfig, CandidateSelection, CandidateSelectionConfig, FragmentCompetition,
)
from alphadia_b200.synthetic import CONFIGS_4D, make_config_3d, make_config_4d  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config1"
raw, precursors_flat, fragments_flat, p = (make_config_4d if name in CONFIGS_4D else make_config_3d)(name)
columns = dict(rt_column="rt_library", mobility_column="mobility_library", precursor_mz_column="mz_library",
               fragment_mz_column="mz_library")

selection_config = CandidateSelectionConfig()
selection_config.update({"rt_tolerance": float(p["rt_tolerance"]), "mobility_tolerance": float(p.get("mobility_tolerance", 0.1)),
                         "candidate_count": 3, "precursor_mz_tolerance": 5.0, "fragment_mz_tolerance": 10.0,
                         "sigma_scale_rt": 0.5, "max_size_mobility": 20})
t0 = time.perf_counter()
candidates_df = CandidateSelection(raw, precursors_flat, fragments_flat, selection_config, fwhm_rt=5.0, fwhm_mobility=0.01,
                                   **columns)(thread_count=8)
t1 = time.perf_counter()
print(f"selection: {len(precursors_flat)} precursors -> {len(candidates_df)} candidates in {t1 - t0:.2f} s "
      f"(first call includes the raw-file upload and index build)")
print(candidates_df.head(3).to_string())

scoring_config = CandidateScoringConfig()
scoring_config.update({"top_k_isotopes": 3, "precursor_mz_tolerance": 5, "fragment_mz_tolerance": 10, "quant_all": True,
                       "experimental_xic": True, "top_k_fragments": 12})
features_df, fragments_df = CandidateScoring(dia_data=raw, precursors_flat=precursors_flat, fragments_flat=fragments_flat,
                                             config=scoring_config, **columns)(candidates_df, thread_count=8)
t2 = time.perf_counter()
print(f"scoring: {len(features_df)} scored candidates x {features_df.shape[1]} columns, {len(fragments_df)} fragment rows in {t2 - t1:.2f} s")
print(features_df[["precursor_idx", "rank", "rt_observed", "intensity_correlation", "mean_observation_score", "delta_rt"]].head(3).to_string())

# the FDR step between scoring and the final table (alphadia/fdr/fdr.py:25-192): the reference's feed-forward classifier is
# trained on the 46 features + the candidate columns (torch on the device), its inference runs in adb_classifier_predict_proba,
# and q-values, fragment competition (data without ion mobility only, fdr.py:157) and best-row-per-group on the device
from alphadia_b200.classifier import BinaryClassifierLegacyNewBatching  # noqa: E402
from alphadia_b200.fdr import perform_fdr  # noqa: E402
from alphadia_b200.scoring import DEFAULT_FEATURE_COLUMNS  # noqa: E402

available = [c for c in DEFAULT_FEATURE_COLUMNS if features_df[c].notna().all() and features_df[c].std() > 0]
psm_df = features_df.assign(channel=0)
classifier = BinaryClassifierLegacyNewBatching(input_dim=len(available), epochs=40, batch_size=64, learning_rate=0.005, random_state=0)
result = perform_fdr(classifier, available, psm_df[psm_df["decoy"] == 0].copy(), psm_df[psm_df["decoy"] == 1].copy(),
                     competitive=True, df_fragments=fragments_df.copy(), dia_cycle=raw.cycle, random_state=0)
t3 = time.perf_counter()
targets = result[result["_decoy"] == 0]
print(f"perform_fdr: {len(available)} features, {len(psm_df)} PSMs -> {len(result)} best-per-elution-group rows, "
      f"{int((targets['qval'] <= 0.01).sum())} targets at q <= 0.01 in {t3 - t2:.2f} s")
print(result[["precursor_idx", "rank", "decoy", "proba", "qval"]].head(5).to_string())

if not getattr(raw, "has_mobility", False):  # the pieces perform_fdr calls, on their own
    kept = FragmentCompetition(rt_tol_seconds=3, mass_tol_ppm=15)(result.copy(), fragments_df.copy(), raw.cycle)
    print(f"fragment competition on its own: {len(kept)} of {len(result)} PSMs keep their fragments")

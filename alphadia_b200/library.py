"""Flat spectral library -> SoA arrays for the C ABI.

Restates the marshalling the reference does before entering numba:
``CandidateSelection._assemble_precursor_container/_assemble_fragment_container``
(alphadia/search/selection/selection.py:678-737) and ``CandidateScoring.assemble_fragments``
(alphadia/search/scoring/scoring.py:355-392): schema validation (dtype casts in place),
precursors sorted by ``precursor_idx``, cardinality defaulted to 1, isotope columns ``i_*``
stacked into one [P, n_isotopes] matrix.
"""

from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from alphadia_b200.validation import fragments_flat_schema, get_isotope_columns, precursors_flat_schema

logger = logging.getLogger()


def assemble_library_arrays(
    precursors_flat: pd.DataFrame,
    fragments_flat: pd.DataFrame,
    rt_column: str,
    mobility_column: str,
    precursor_mz_column: str,
    fragment_mz_column: str,
) -> dict[str, np.ndarray]:
    """Validated, precursor_idx-sorted SoA arrays (keys = fields of ``adb_library_desc``)."""
    precursors_flat_schema.validate(precursors_flat, warn_on_critical_values=True)
    if not precursors_flat["precursor_idx"].is_monotonic_increasing:
        precursors_flat = precursors_flat.sort_values("precursor_idx").reset_index(drop=True)

    if "cardinality" not in fragments_flat.columns:
        logger.warning("Fragment cardinality column not found in fragment dataframe. Setting cardinality to 1.")
        fragments_flat["cardinality"] = np.ones(len(fragments_flat), dtype=np.uint8)
    fragments_flat_schema.validate(fragments_flat, warn_on_critical_values=True)

    iso_cols = [f"i_{i}" for i in get_isotope_columns(precursors_flat.columns)]
    if iso_cols:
        isotopes = np.ascontiguousarray(precursors_flat[iso_cols].values, dtype=np.float32)
    else:  # scoring.py:321-323: monoisotopic abundance defaults to 1
        isotopes = np.ones((len(precursors_flat), 1), dtype=np.float32)

    def c(df, col, dt):
        return np.ascontiguousarray(df[col].values, dtype=dt)

    return dict(
        precursor_idx=c(precursors_flat, "precursor_idx", np.uint32),
        frag_start_idx=c(precursors_flat, "flat_frag_start_idx", np.uint32),
        frag_stop_idx=c(precursors_flat, "flat_frag_stop_idx", np.uint32),
        charge=c(precursors_flat, "charge", np.uint8),
        rt=c(precursors_flat, rt_column, np.float32),
        mobility=c(precursors_flat, mobility_column, np.float32),
        mz=c(precursors_flat, precursor_mz_column, np.float32),
        isotopes=isotopes,
        frag_mz_library=c(fragments_flat, "mz_library", np.float32),
        frag_mz=c(fragments_flat, fragment_mz_column, np.float32),
        frag_intensity=c(fragments_flat, "intensity", np.float32),
        frag_type=c(fragments_flat, "type", np.uint8),
        frag_loss_type=c(fragments_flat, "loss_type", np.uint8),
        frag_charge=c(fragments_flat, "charge", np.uint8),
        frag_number=c(fragments_flat, "number", np.uint8),
        frag_position=c(fragments_flat, "position", np.uint8),
        frag_cardinality=c(fragments_flat, "cardinality", np.uint8),
    )

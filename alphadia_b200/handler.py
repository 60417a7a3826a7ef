"""``B200ExtractionHandler`` — the reference's ``ClassicExtractionHandler`` with the numba operators replaced by the sm_100a
engine (alphadia/workflow/peptidecentric/extraction_handler.py:344-508).

The reference instantiates ``CandidateSelection`` / ``CandidateScoring`` by name inside ``_select_candidates`` (:411-449) and
``score_and_quantify_candidates`` (:451-487), so the handler overrides exactly these two methods; everything else — the public
``select_candidates`` with its score cutoff (:119-154, :177-202), ``quantify_candidates`` (:488-508, which calls
``score_and_quantify_candidates`` and therefore runs on the device too), the parameter logging — is inherited.

The reference package is imported lazily: ``B200ExtractionHandler`` resolves its base class on first use, so this module can
be imported (and unit-tested with a stand-in base class through ``make_handler_class``) where alphaDIA is not installed.

    from alphadia_b200.handler import install
    install()        # registers extraction_backend: "b200" with ExtractionHandler.create_handler (:70-117)
"""

from __future__ import annotations

import pandas as pd

from alphadia_b200.config import CandidateScoringConfig, CandidateSelectionConfig
from alphadia_b200.scoring import CandidateScoring
from alphadia_b200.selection import CandidateSelection

BACKEND_NAME = "b200"


def make_handler_class(base):
    """``class B200ExtractionHandler(base)`` for a given ``ClassicExtractionHandler``-like base class."""

    class B200ExtractionHandler(base):
        """Extraction handler using the B200 engine (same flow and tables as the classic backend)."""

        _selection_cls = CandidateSelection
        _scoring_cls = CandidateScoring

        def __init__(self, config, optimization_manager, fdr_manager, reporter, column_name_handler):
            super().__init__(config, optimization_manager, fdr_manager, reporter, column_name_handler)
            # extraction_handler.py:387-409 with C-struct marshalling configs instead of numba jitclass ones
            self._selection_config = CandidateSelectionConfig()
            self._selection_config.update(
                {
                    **self._base_selection_config,
                    "top_k_fragments": config["search"]["top_k_fragments_selection"],
                    "exclude_shared_ions": config["search"]["exclude_shared_ions"],
                    "min_size_rt": config["search"]["quant_window"],
                }
            )
            self._scoring_config = CandidateScoringConfig()
            self._scoring_config.update(
                {
                    **self._base_scoring_config,
                    "exclude_shared_ions": config["search"]["exclude_shared_ions"],
                    "quant_window": config["search"]["quant_window"],
                    "quant_all": config["search"]["quant_all"],
                    "experimental_xic": config["search"]["experimental_xic"],
                }
            )

        def _select_candidates(self, dia_data, spectral_library) -> pd.DataFrame:
            """extraction_handler.py:411-449 on ``alphadia_b200.CandidateSelection``."""
            self._log_parameters()
            om = self._optimization_manager
            self._selection_config.update(
                {
                    "rt_tolerance": om.rt_error,
                    "mobility_tolerance": om.mobility_error,
                    "candidate_count": om.num_candidates,
                    "precursor_mz_tolerance": om.ms1_error,
                    "fragment_mz_tolerance": om.ms2_error,
                }
            )
            cols = self._column_name_handler
            extraction = self._selection_cls(
                dia_data,
                spectral_library.precursor_df,
                spectral_library.fragment_df,
                self._selection_config,
                rt_column=cols.get_rt_column(),
                mobility_column=cols.get_mobility_column(),
                precursor_mz_column=cols.get_precursor_mz_column(),
                fragment_mz_column=cols.get_fragment_mz_column(),
                fwhm_rt=om.fwhm_rt,
                fwhm_mobility=om.fwhm_mobility,
            )
            return extraction(thread_count=self._config["general"]["thread_count"])

        def score_and_quantify_candidates(self, candidates_df, dia_data, spectral_library, top_k_fragments=None):
            """extraction_handler.py:451-487 on ``alphadia_b200.CandidateScoring``."""
            om = self._optimization_manager
            self._scoring_config.update(
                {
                    "precursor_mz_tolerance": om.ms1_error,
                    "fragment_mz_tolerance": om.ms2_error,
                    "top_k_fragments": top_k_fragments
                    if top_k_fragments is not None
                    else self._config["search"]["top_k_fragments_scoring"],
                }
            )
            cols = self._column_name_handler
            candidate_scoring = self._scoring_cls(
                dia_data=dia_data,
                precursors_flat=spectral_library.precursor_df,
                fragments_flat=spectral_library.fragment_df,
                config=self._scoring_config,
                rt_column=cols.get_rt_column(),
                mobility_column=cols.get_mobility_column(),
                precursor_mz_column=cols.get_precursor_mz_column(),
                fragment_mz_column=cols.get_fragment_mz_column(),
            )
            return candidate_scoring(
                candidates_df,
                thread_count=self._config["general"]["thread_count"],
                include_decoy_fragment_features=True,
            )

    B200ExtractionHandler.__qualname__ = "B200ExtractionHandler"
    return B200ExtractionHandler


_HANDLER_CLASS = None


def handler_class():
    """``B200ExtractionHandler`` derived from the installed reference's ``ClassicExtractionHandler``."""
    global _HANDLER_CLASS
    if _HANDLER_CLASS is None:
        try:
            from alphadia.workflow.peptidecentric.extraction_handler import ClassicExtractionHandler
        except ImportError as e:  # pragma: no cover - depends on the environment
            raise ImportError(
                "alphadia_b200.handler needs alphaDIA (alphadia.workflow.peptidecentric.extraction_handler); "
                "use make_handler_class(base) with your own base class otherwise"
            ) from e
        _HANDLER_CLASS = make_handler_class(ClassicExtractionHandler)
    return _HANDLER_CLASS


def __getattr__(name):  # `from alphadia_b200.handler import B200ExtractionHandler` resolves the reference lazily
    if name == "B200ExtractionHandler":
        return handler_class()
    raise AttributeError(name)


def install(extraction_handler_module=None):
    """Makes ``search.extraction_backend: b200`` selectable: wraps ``ExtractionHandler.create_handler``
    (extraction_handler.py:70-117) so that the new backend name returns a ``B200ExtractionHandler`` and every other name is
    passed on unchanged.  The four call sites that compare the backend string with ``"python"`` are listed in
    INTEGRATION.md.  Idempotent; returns the handler class."""
    if extraction_handler_module is None:
        from alphadia.workflow.peptidecentric import extraction_handler as extraction_handler_module
    mod = extraction_handler_module
    base = mod.ExtractionHandler
    if getattr(base.create_handler, "_adb_installed", False):
        return base.create_handler._adb_handler_cls
    cls = make_handler_class(mod.ClassicExtractionHandler)
    original = base.create_handler

    def create_handler(config, optimization_manager, fdr_manager, reporter, column_name_handler):
        backend = str(config["search"]["extraction_backend"]).lower()
        if backend == BACKEND_NAME:
            reporter.log_string(f"Using {backend} extraction backend", verbosity="info")
            return cls(config, optimization_manager, fdr_manager, reporter, column_name_handler)
        return original(config, optimization_manager, fdr_manager, reporter, column_name_handler)

    create_handler._adb_installed = True
    create_handler._adb_handler_cls = cls
    base.create_handler = staticmethod(create_handler)
    return cls

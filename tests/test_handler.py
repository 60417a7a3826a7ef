"""alphadia_b200.handler.B200ExtractionHandler against a stand-in for the reference's ClassicExtractionHandler
(alphadia/workflow/peptidecentric/extraction_handler.py:344-508): the reference module cannot be imported on the test boxes
(it needs alphadia_search_rs), so the base class below restates its constructor and the three inherited entry points; the
device calls are replaced by the oracle (CPU test of the host logic)."""

from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest

from tests import helpers as H

INT_COLS = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]


class Reporter:
    def __init__(self):
        self.lines = []

    def log_string(self, s, verbosity="info"):
        self.lines.append((verbosity, s))


class StandInClassicHandler:
    """What B200ExtractionHandler inherits from the reference (extraction_handler.py:41-68,119-202,349-409,488-508)."""

    _base_selection_config = {k: v for k, v in H.SELECTION_BASE.items() if k not in ("top_k_fragments", "exclude_shared_ions")}
    _base_scoring_config = {"score_grouped": False, "top_k_isotopes": 3, "reference_channel": -1, "precursor_mz_tolerance": 10,
                            "fragment_mz_tolerance": 15}

    def __init__(self, config, optimization_manager, fdr_manager, reporter, column_name_handler):
        self._config, self._optimization_manager, self._fdr_manager = config, optimization_manager, fdr_manager
        self._reporter, self._column_name_handler = reporter, column_name_handler
        self._selection_config = self._scoring_config = "numba configs of the reference"

    def _log_parameters(self):
        self._reporter.log_string("=== Search parameters used ===", verbosity="info")

    def select_candidates(self, dia_data, spectral_library, apply_cutoff=False):
        df = self._select_candidates(dia_data, spectral_library)
        return df[df["score"] > self._optimization_manager.score_cutoff] if apply_cutoff else df

    def quantify_candidates(self, candidates_df, precursor_fdr_df, dia_data, spectral_library, top_k_fragments=None):
        _, fragments_df = self.score_and_quantify_candidates(candidates_df, dia_data, spectral_library, top_k_fragments)
        return None, fragments_df


def _make(name="parity_small"):
    from alphadia_b200.handler import make_handler_class

    raw, pdf, fdf, lib, p = H.workload(name)
    config = {"search": {"top_k_fragments_selection": 12, "top_k_fragments_scoring": 12, "exclude_shared_ions": True, "quant_window": 3,
                         "quant_all": True, "experimental_xic": True, "extraction_backend": "b200"},
              "general": {"thread_count": 4}}
    om = SimpleNamespace(rt_error=float(p["rt_tolerance"]), mobility_error=0.1, num_candidates=3, ms1_error=5.0, ms2_error=10.0,
                         fwhm_rt=5.0, fwhm_mobility=0.01, score_cutoff=50.0)
    cols = SimpleNamespace(get_rt_column=lambda: "rt_library", get_mobility_column=lambda: "mobility_library",
                           get_precursor_mz_column=lambda: "mz_library", get_fragment_mz_column=lambda: "mz_library")
    reporter = Reporter()
    handler = make_handler_class(StandInClassicHandler)(config, om, None, reporter, cols)
    speclib = SimpleNamespace(precursor_df=pdf.copy(), fragment_df=fdf.copy())
    return handler, raw, speclib, reporter


def test_handler_flow_equals_the_operator_classes(oracle_lib, monkeypatch):
    from alphadia_b200.config import CandidateScoringConfig, CandidateSelectionConfig

    H.patch_device_with_oracle(monkeypatch, oracle_lib)
    handler, raw, speclib, reporter = _make()
    assert isinstance(handler._selection_config, CandidateSelectionConfig) and isinstance(handler._scoring_config, CandidateScoringConfig)
    assert handler._selection_config.min_size_rt == 3 and handler._selection_config.top_k_fragments == 12
    g = H.load_golden("parity_small")
    cand = handler.select_candidates(raw, speclib)
    assert ("info", "=== Search parameters used ===") in reporter.lines
    # the table the reference's CandidateSelection returned for this configuration (tests/golden/parity_small.npz)
    for c in INT_COLS + ["score", "elution_group_idx", "decoy"]:
        assert cand[c].dtype == g["cand_" + c].dtype and np.array_equal(cand[c].values, g["cand_" + c]), c
    cut = handler.select_candidates(raw, speclib, apply_cutoff=True)
    assert 0 < len(cut) < len(cand) and (cut["score"] > 50.0).all()
    feat, frag = handler.score_and_quantify_candidates(cand, raw, speclib)
    assert np.array_equal(feat["precursor_idx"].values, g["feat_precursor_idx"]) and np.array_equal(feat["rank"].values, g["feat_rank"])
    assert np.array_equal(frag["precursor_idx"].values, g["frag_precursor_idx"])
    assert np.allclose(frag["mz_observed"].values, g["frag_mz_observed"], rtol=1e-6)
    # quantify_candidates (transfer-library / multiplexing requantification) goes through the same device scoring
    none, frag6 = handler.quantify_candidates(cand, None, raw, speclib, top_k_fragments=6)
    assert none is None and handler._scoring_config.top_k_fragments == 6
    assert frag6.groupby(["precursor_idx", "rank"]).size().max() <= 6 < frag.groupby(["precursor_idx", "rank"]).size().max()


def test_install_registers_the_backend():
    from alphadia_b200 import handler as adb_handler

    calls = []

    class ExtractionHandler:
        @staticmethod
        def create_handler(config, optimization_manager, fdr_manager, reporter, column_name_handler):
            calls.append(config["search"]["extraction_backend"])
            return "classic"

    mod = SimpleNamespace(ExtractionHandler=ExtractionHandler, ClassicExtractionHandler=StandInClassicHandler)
    cls = adb_handler.install(mod)
    assert adb_handler.install(mod) is cls  # idempotent
    rep = Reporter()
    cfg = {"search": {"top_k_fragments_selection": 12, "top_k_fragments_scoring": 12, "exclude_shared_ions": True, "quant_window": 3,
                      "quant_all": True, "experimental_xic": True, "extraction_backend": "B200"}, "general": {"thread_count": 1}}
    h = mod.ExtractionHandler.create_handler(cfg, None, None, rep, None)
    assert isinstance(h, cls) and isinstance(h, StandInClassicHandler) and rep.lines == [("info", "Using b200 extraction backend")]
    cfg["search"]["extraction_backend"] = "python"
    assert mod.ExtractionHandler.create_handler(cfg, None, None, rep, None) == "classic" and calls == ["python"]
    with pytest.raises(ImportError):
        adb_handler.handler_class()  # alphaDIA itself is not installed here

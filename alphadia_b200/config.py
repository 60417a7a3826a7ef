"""Host-side mirrors of the reference's hyper-parameter objects.

Same attribute names, defaults and ``update`` semantics as the reference so existing call
sites (``ClassicExtractionHandler``, extraction_handler.py:349-409) work unchanged:

* ``CandidateSelectionConfig`` <- alphadia/search/selection/config_df.py:127-181
* ``CandidateScoringConfig``   <- alphadia/search/scoring/config.py:68-222
* ``JITConfig.update``         <- alphadia/search/jitclasses/jit_config.py:84-138
  (values are coerced to the *type of the current value* — an int default therefore truncates a
  float tolerance exactly as in the reference).

``to_jitclass()`` in the reference builds a numba jitclass; here ``to_struct()`` builds the
C-ABI struct of ``include/alphadia_b200.h``.  ``to_jitclass`` is kept as an alias.
"""

from __future__ import annotations

import numpy as np

from alphadia_b200 import _abi

MAX_FRAGMENT_MZ_TOLERANCE = 200  # alphadia/constants/settings.py


class _Config:
    def update(self, input_dict: dict):
        for key, value in input_dict.items():
            if not hasattr(self, key):
                raise ValueError(f"Parameter {key} does not exist in {self.__class__.__name__}")
            current = getattr(self, key)
            if not isinstance(value, type(current)):
                try:
                    value = type(current)(value)
                except Exception as e:
                    raise ValueError(f"Parameter {key} has wrong type {type(value)}") from e
            if isinstance(value, np.ndarray) and value.dtype != current.dtype:
                try:
                    value = value.astype(current.dtype)
                except Exception as e:
                    raise ValueError(f"Parameter {key} has wrong dtype {value.dtype}") from e
            if isinstance(value, np.ndarray) and value.shape != current.shape:
                raise ValueError(f"Parameter {key} has wrong shape {value.shape}")
            setattr(self, key, value)

    def validate(self):
        pass

    def to_jitclass(self):
        return self.to_struct()

    def __repr__(self) -> str:
        body = " \n".join(f"{k}={v}" for k, v in self.__dict__.items())
        return f"<{self.__class__.__name__}, \n{body} \n>"


class CandidateSelectionConfig(_Config):
    def __init__(self):
        self.rt_tolerance = 60.0
        self.precursor_mz_tolerance = 10.0
        self.fragment_mz_tolerance = 15.0
        self.mobility_tolerance = 0.1
        self.isotope_tolerance = 0.01

        self.peak_len_rt = 10.0
        self.sigma_scale_rt = 0.1
        self.peak_len_mobility = 0.013
        self.sigma_scale_mobility = 1.0

        self.candidate_count = 5

        self.top_k_precursors = 3
        self.top_k_fragments = 12
        self.exclude_shared_ions = True
        self.kernel_size = 30

        self.f_mobility = 1.0
        self.f_rt = 0.99
        self.center_fraction = 0.5
        self.min_size_mobility = 8
        self.min_size_rt = 3
        self.max_size_mobility = 30
        self.max_size_rt = 15

        self.group_channels = False
        self.use_weighted_score = True

        self.join_close_candidates = True
        self.join_close_candidates_scan_threshold = 0.01
        self.join_close_candidates_cycle_threshold = 0.6

        self.feature_std = np.ones(1, np.float64)
        self.feature_mean = np.zeros(1, np.float64)
        self.feature_weight = np.ones(1, np.float64)

    def to_struct(self) -> _abi.SelectionConfig:
        self.validate()
        s = _abi.SelectionConfig()
        for name in ("rt_tolerance", "precursor_mz_tolerance", "fragment_mz_tolerance", "mobility_tolerance",
                     "f_mobility", "f_rt", "center_fraction", "join_close_candidates_scan_threshold",
                     "join_close_candidates_cycle_threshold"):
            setattr(s, name, float(getattr(self, name)))
        for name in ("candidate_count", "top_k_precursors", "top_k_fragments", "kernel_size", "min_size_mobility",
                     "min_size_rt", "max_size_mobility", "max_size_rt"):
            setattr(s, name, int(getattr(self, name)))
        s.exclude_shared_ions = int(bool(self.exclude_shared_ions))
        s.use_weighted_score = int(bool(self.use_weighted_score))
        s.join_close_candidates = int(bool(self.join_close_candidates))
        s.feature_std = float(self.feature_std[0])
        s.feature_mean = float(self.feature_mean[0])
        s.feature_weight = float(self.feature_weight[0])
        return s


class CandidateScoringConfig(_Config):
    def __init__(self):
        self.collect_fragments = True
        self.score_grouped = False
        self.exclude_shared_ions = True
        self.top_k_fragments = 12
        self.top_k_isotopes = 4
        self.reference_channel = -1
        self.quant_window = 3
        self.quant_all = False
        self.precursor_mz_tolerance = 15
        self.fragment_mz_tolerance = 15
        self.experimental_xic = False

    def validate(self):
        assert isinstance(self.score_grouped, bool), "score_grouped must be a boolean"
        assert self.top_k_fragments > 0, "top_k_fragments must be greater than 0"
        assert self.top_k_isotopes > 0, "top_k_isotopes must be greater than 0"
        assert self.reference_channel >= -1, "reference_channel must be greater than or equal to -1"
        assert self.precursor_mz_tolerance >= 0, "precursor_mz_tolerance must be greater than or equal to 0"
        assert self.precursor_mz_tolerance < 200, "precursor_mz_tolerance must be less than 200"
        assert self.fragment_mz_tolerance >= 0, "fragment_mz_tolerance must be greater than or equal to 0"
        assert (
            self.fragment_mz_tolerance <= MAX_FRAGMENT_MZ_TOLERANCE
        ), f"fragment_mz_tolerance must be less than or equal {MAX_FRAGMENT_MZ_TOLERANCE}"

    def to_struct(self, quad_sigma=(0.2, 0.2), quad_delta_mu=(0.0, 0.0)) -> _abi.ScoringConfig:
        self.validate()
        s = _abi.ScoringConfig()
        s.collect_fragments = int(bool(self.collect_fragments))
        s.exclude_shared_ions = int(bool(self.exclude_shared_ions))
        s.top_k_fragments = int(self.top_k_fragments)
        s.top_k_isotopes = int(self.top_k_isotopes)
        s.quant_window = int(self.quant_window)
        s.quant_all = int(bool(self.quant_all))
        s.precursor_mz_tolerance = float(np.float32(self.precursor_mz_tolerance))
        s.fragment_mz_tolerance = float(np.float32(self.fragment_mz_tolerance))
        s.experimental_xic = int(bool(self.experimental_xic))
        s.quad_sigma[0], s.quad_sigma[1] = float(quad_sigma[0]), float(quad_sigma[1])
        s.quad_delta_mu[0], s.quad_delta_mu[1] = float(quad_delta_mu[0]), float(quad_delta_mu[1])
        return s

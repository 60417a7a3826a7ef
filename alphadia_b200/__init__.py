"""alphadia_b200 — B200-native (sm_100a) engine for alphaDIA's precursor-candidate hot path.

Drop-in operators (same names and call signatures as the reference):

* ``CandidateSelection``   (alphadia/search/selection/selection.py:547)
* ``CandidateScoring``     (alphadia/search/scoring/scoring.py:140)
* ``FragmentCompetition``  (alphadia/fragcomp/fragcomp.py:146)

all backed by hand-written CUDA kernels behind the C ABI of ``include/alphadia_b200.h``.
There is no CPU fallback: the operators raise if the extension or a CUDA device is missing.
"""

from alphadia_b200.config import CandidateScoringConfig, CandidateSelectionConfig  # noqa: F401
from alphadia_b200.fragcomp import FragmentCompetition  # noqa: F401
from alphadia_b200.scoring import CandidateScoring  # noqa: F401
from alphadia_b200.selection import CandidateSelection  # noqa: F401

__version__ = "0.1.0"

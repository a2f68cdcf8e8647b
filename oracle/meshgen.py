"""Moved: the synthetic input generators live in ``densematcher_b200/synth.py`` (they are input generation, not part of
the checker).  This alias keeps ``from oracle import meshgen`` working for the test-suite."""
from densematcher_b200.synth import *  # noqa: F401,F403
from densematcher_b200.synth import (bandlimited_features, cotan_stiffness, deform, icosphere, lbo_basis,  # noqa: F401
                                     lumped_area, random_unit_features, synthetic_basis)

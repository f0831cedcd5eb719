"""Seeded synthetic configs / weights / inputs now live in tests/synth.py (they are shared by the tests, the golden
generator, bench.py and the scripts, and are NOT part of the oracle's arithmetic).  This module re-exports them so
that `from oracle import synth` keeps working inside the test infrastructure."""
from tests.synth import *  # noqa: F401,F403
from tests.synth import make_config, gpt2_weights, lorenz_embedding_weights, cylinder_embedding_weights  # noqa: F401

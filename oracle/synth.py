"""Kept for the tests' import path: the seeded input / weight generators live in tools/synth.py (they generate
inputs only and are also used by bench.py's GPU arm, which must not import oracle/)."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import _key_seed  # noqa: F401

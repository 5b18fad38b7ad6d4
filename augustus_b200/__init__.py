"""augustus_b200 — B200-native GHMM Viterbi decoder behind AUGUSTUS' NAMGene / StateModel seam.

The compute path is ``libaugb200.so`` (hand-written sm_100a CUDA, C ABI in ``include/augb200.h``);
this package is the thin Python binding used by the tests and ``bench.py``.  There is no CPU
fallback: constructing a :class:`Decoder` without the CUDA library or without a GPU raises.
"""
from .decoder import Decoder, StatePath, State, AugB200Error, library_path, STATE_TYPE_NAMES  # noqa: F401

__all__ = ["Decoder", "StatePath", "State", "AugB200Error", "library_path", "STATE_TYPE_NAMES"]

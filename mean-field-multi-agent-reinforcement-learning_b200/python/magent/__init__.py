"""Host-side mirror of the reference's `magent` Python package, for the battle hot path only.

Same import surface the MFRL scripts use (reference: examples/battle_model/python/magent/__init__.py:1-8):
`magent.GridWorld`, `magent.gridworld` (Config / CircleRange / AgentSymbol / Event) and the builtin
`battle` config.  The engine behind it is the B200 CUDA library `build/libmagent.so` of this package
(C ABI declared in include/mfmarl_magent.h); there is no CPU fallback.
"""
from . import gridworld
from .environment import Environment

GridWorld = gridworld.GridWorld

__all__ = ["gridworld", "GridWorld", "Environment"]

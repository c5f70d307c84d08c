"""PyTorch re-expression of the reference's learning algorithms (examples/battle_model/algo/), consuming the
engine's observations and the group mean action.  Same duck type as the TF1 models -- `act(state=, prob=, eps=)`,
`flush_buffer(**)`, `train()`, `save/load` -- so the reference's play loop can drive them unchanged.

    MFQ   mean-field Q-learning      algo/q_learning.py:74-143 + base.py (use_mf=True)
    IL    independent Q-learning     algo/q_learning.py:9-71   (DQN)
    AC    actor-critic               algo/ac.py:8-176
    MFAC  mean-field actor-critic    algo/ac.py:178-361
"""
from . import ac, q_learning, tools
from .ac import MFAC, ActorCritic
from .q_learning import DQN, MFQ

AC = ActorCritic
IL = DQN


def spawn_ai(algo_name, env, handle, human_name, max_steps, device=None):
    """algo/__init__.py:10-19 without the TF session argument."""
    if algo_name == "mfq":
        return MFQ(human_name, handle, env, max_steps, memory_size=80000, device=device)
    if algo_name == "mfac":
        return MFAC(human_name, handle, env, device=device)
    if algo_name == "ac":
        return AC(human_name, handle, env, device=device)
    if algo_name == "il":
        return IL(human_name, handle, env, max_steps, memory_size=80000, device=device)
    raise ValueError("unknown algorithm %r (choose from mfq, mfac, ac, il)" % (algo_name,))


__all__ = ["spawn_ai", "MFQ", "DQN", "IL", "ActorCritic", "AC", "MFAC", "tools", "ac", "q_learning"]

"""Where one step of the single-env play loop goes (bench.py --workload c2): wall time per call group, for the CUDA
library and for the reference CPU engine.   python profiles/c2_breakdown.py [cuda|reference] [steps]"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mean-field-multi-agent-reinforcement-learning_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "tests"))
os.environ.setdefault("OMP_NUM_THREADS", "1")
kind = sys.argv[1] if len(sys.argv) > 1 else "cuda"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
from mfmarl_b200.scenarios import generate_map_positions  # noqa: E402
import conftest  # noqa: E402,F401
from engines import CudaEngine, RefEngine  # noqa: E402

eng = CudaEngine(40) if kind == "cuda" else RefEngine(40)
left, right = generate_map_positions(40)
rng = np.random.RandomState(7)
eye = np.eye(21)
acc = {}


def timed(name, fn, *a):
    t = time.perf_counter()
    r = fn(*a)
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
    return r


def reset():
    eng.reset(); eng.add_agents(0, left); eng.add_agents(1, right)


reset()
ep = 0
t_all = None
for s in range(200 + steps):
    if s == 200:
        acc.clear(); t_all = time.perf_counter()
    n = [timed("get_num", eng.get_num, g) for g in range(2)]
    timed("get_observation(0)", eng.env.get_observation, eng.h[0])
    timed("get_observation(1)", eng.env.get_observation, eng.h[1])
    acts = [timed("randint", lambda g=g: rng.randint(0, 21, size=n[g]).astype(np.int32)) for g in range(2)]
    for g in range(2):
        timed("set_action", eng.set_action, g, acts[g])
    done = timed("step", eng.step)
    for g in range(2):
        timed("get_reward", eng.get_reward, g); timed("get_alive", eng.get_alive, g)
        timed("mean_action (numpy)", lambda g=g: np.mean(eye[acts[g]], axis=0, keepdims=True))
    timed("clear_dead", eng.clear_dead)
    ep += 1
    if done or ep >= 400:
        reset(); ep = 0
total = time.perf_counter() - t_all
print("%s: %.1f us per step" % (kind, total / steps * 1e6))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  %-22s %7.1f us" % (k, v / steps * 1e6))

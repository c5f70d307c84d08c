"""Small end-to-end case for compute-sanitizer: single-env ABI + batched engine + Ising, a few steps each."""
import os, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(R, "mean-field-multi-agent-reinforcement-learning_b200", "python")); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from engines import CudaEngine
from scenarios import generate_map_positions, fight_actions, c4_positions
from mfmarl_b200 import BatchedGridWorld, IsingMFQ
left, right = generate_map_positions(40)
cu = CudaEngine(40); cu.reset(); cu.add_agents(0, left[:13]); cu.add_agents(1, right[:7])
rng = np.random.RandomState(0)
for s in range(15):
    for g in range(2): cu.get_observation(g)
    for g in range(2): cu.set_action(g, fight_actions(rng, cu.get_pos(g), 40))
    cu.step(); [cu.get_reward(g) for g in range(2)]; cu.clear_dead()
for (ms, cap, pos) in [(40, 64, (left, right)), (80, 512, c4_positions())]:
    env = BatchedGridWorld(3, map_size=ms, capacity=cap, rng="philox", auto_reset=True, max_steps=6)
    env.reset(); env.add_agents(0, pos[0]); env.add_agents(1, pos[1])
    for s in range(8):
        env.observe()
        p, n = env.get("pos"), env.get_num()
        a = np.zeros((3, 2, cap), np.int32)
        for e in range(3):
            for g in range(2): a[e, g, :n[e, g]] = fight_actions(rng, p[e, g, :n[e, g]], ms)
        env.step(torch.from_numpy(a).cuda())
for L in (20, 64, 128, 256):
    for rpt in ("0", "4", "8"):          # generic cluster-barrier kernel, specialised st.async/mbarrier kernel (4 / 8 rows per thread)
        os.environ["MFMARL_ISING_RPT"] = rpt
        m = IsingMFQ(2, L); m.run([0.8] * 3, resident=False); m.run([0.8] * 5, resident=True)
env = BatchedGridWorld(3, map_size=40, capacity=64, rng="minstd")
env.reset(); env.add_agents(0, left); env.add_agents(1, right)
env.observe_groups(); env.observe_groups(groups=(1,)); env.device_state("id")
tmp = "/tmp/_trace"; os.makedirs(tmp, exist_ok=True)
cu.env.set_render_dir(tmp)
for s in range(4):
    for g in range(2): cu.set_action(g, fight_actions(rng, cu.get_pos(g), 40))
    cu.step(); cu.env.render(); cu.clear_dead()
torch.cuda.synchronize(); print("sanitize case done")

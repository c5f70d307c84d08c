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
# round 2: bf16 rows, observation record forced on at cap 64, per-env placement + random sides, large-map band order,
# the Ising environment step, the persistent Ising kernel (default for 256 / 128) with and without act groups
from scenarios import block_positions
for cached in (0, 1):
    env = BatchedGridWorld(3, map_size=40, capacity=64, rng="philox", auto_reset=True, max_steps=3, obs_record=cached,
                           random_sides=True, concurrent_step_envs=3)
    env.reset()
    env.add_agents_per_env(0, np.stack([left[:, :2] + [0, e % 2] for e in range(3)]))
    env.add_agents(1, right)
    for s in range(5):
        env.observe_groups(dtype=torch.bfloat16); env.observe()
        env.step(torch.from_numpy(rng.randint(0, 21, size=(3, 2, 64)).astype(np.int32)).cuda())
big = CudaEngine(104); big.reset()
big.add_agents(0, block_positions(20, 30, 20, 10, stride=1)); big.add_agents(1, block_positions(54, 30, 20, 10, stride=1))
for s in range(4):
    for g in range(2): big.get_observation(g)
    for g in range(2): big.set_action(g, fight_actions(rng, big.get_pos(g), 104))
    big.step(); big.get_observation(0); big.clear_dead()          # (an observation before clear_dead: the rollback path)
os.environ.pop("MFMARL_ISING_RPT", None)
for persist in (None, "1"):                        # K6s (default; 64 / 128 / 256), then its predecessor K6p (128 / 256)
    if persist: os.environ["MFMARL_ISING_PERSIST"] = persist
    for L in (64, 128, 256):
        m = IsingMFQ(11, L)                        # 11 lattices: a slot sweeps more than one (the state-number hand-over)
        m.run([0.8] * 4, resident=True)
        masks = (torch.rand((4, 11, L * L), device="cuda") < 0.5).to(torch.uint8).contiguous()
        m.run([0.8] * 4, resident=True, update_mask=masks)
        m.run([0.8] * 35, resident=True)           # more than 32 sweeps: the per-warp statistics flush inside the loop
os.environ.pop("MFMARL_ISING_PERSIST", None)
import ctypes
from mfmarl_b200.lib import load_library
lib = load_library(); lib.mfi_env_step.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6
sp = torch.randint(0, 2, (2, 33 * 33), dtype=torch.int8, device="cuda"); ac = torch.randint(-1, 2, (2, 33 * 33), dtype=torch.int32, device="cuda")
ob = torch.zeros((2, 33 * 33, 4), dtype=torch.uint8, device="cuda"); rw = torch.zeros((2, 33 * 33), device="cuda"); nu = torch.zeros((2,), dtype=torch.int32, device="cuda")
lib.mfi_env_step(2, 33, sp.data_ptr(), ac.data_ptr(), ob.data_ptr(), rw.data_ptr(), nu.data_ptr(), None)
torch.cuda.synchronize(); print("sanitize case done")

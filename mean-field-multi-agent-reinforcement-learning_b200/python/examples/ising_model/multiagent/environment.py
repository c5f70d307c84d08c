"""IsingMultiAgentEnv over the B200 kernels (reference: examples/ising_model/multiagent/environment.py).

Same constructor, attributes and return values as the reference class, so main_MFQ_Ising.py runs unchanged:
  reset()            -> obs_n                                   (environment.py:80-88)
  step(action_n)     -> obs_n, reward_n, done_n, order_param, n_up, n_down     (:49-78)
with obs_n[i] the four neighbour spins of agent i (float64, ascending flat index of the neighbour), reward_n[i] a
1-element float64 array (what the reference's np.where-indexed arithmetic yields), done_n[i] = (order_param == 1.0).
One `mfi_env_step` call (include/mfmarl_batched.h) applies the whole action vector and evaluates every agent's
reward and observation on the new lattice; the lattice never leaves HBM except as the values returned here.
gym is not needed: `step` / `reset` are the methods themselves, and the two spaces are plain holders of `.n`."""
import ctypes

import numpy as np
import torch

from mfmarl_b200.lib import check, load_library


class _Space(object):
    def __init__(self, n):
        self.n = n


class IsingMultiAgentEnv(object):
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, world, reset_callback=None, reward_callback=None, observation_callback=None,
                 info_callback=None, done_callback=None):
        self.world = world
        self.agents = self.world.policy_agents
        self.n = len(world.policy_agents)
        assert self.n == len(world.agents)
        self.reset_callback, self.reward_callback = reset_callback, reward_callback
        self.observation_callback, self.info_callback, self.done_callback = observation_callback, info_callback, done_callback
        for name, cb in (("reward", reward_callback), ("observation", observation_callback), ("done", done_callback)):
            owner = getattr(cb, "__self__", None)
            if cb is not None and not (type(owner).__name__ == "Scenario" and type(owner).__module__.endswith("ising_model.Ising")):
                raise NotImplementedError("%s_callback: only the built-in Ising scenario's callbacks run on the device" % name)
        self.discrete_action_space = True
        self.shared_reward = False
        self.time = 0
        self.action_space = [_Space(self.world.dim_spin)]                    # Discrete(2)
        self.observation_space = [_Space(4 * self.world.agent_view_sight)]   # MultiBinary(4)
        self._lib = load_library()
        self._lib.mfi_env_step.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6
        self._lib.mfi_env_step.restype = ctypes.c_int
        dev, N = world.device, self.n
        self._actions = torch.zeros((1, N), dtype=torch.int32, device=dev)
        self._obs = torch.zeros((1, N, 4), dtype=torch.uint8, device=dev)
        self._reward = torch.zeros((1, N), dtype=torch.float32, device=dev)
        self._n_up = torch.zeros((1,), dtype=torch.int32, device=dev)

    # ---- device call ----
    def _device_step(self, actions):
        w = self.world
        with torch.cuda.device(w.device):
            check(self._lib.mfi_env_step(1, w.shape_size, w.spins.data_ptr(),
                                         self._actions.data_ptr() if actions is not None else None,
                                         self._obs.data_ptr(), self._reward.data_ptr(), self._n_up.data_ptr(),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if actions is not None:
            w.invalidate_host()
        obs = self._obs[0].to(torch.float64).cpu().numpy()                   # [N, 4] of 0. / 1.
        reward = self._reward[0].to(torch.float64).cpu().numpy()[:, None]    # [N, 1]: reward_n[i] is a 1-element array
        return obs, reward, int(self._n_up[0])

    # ---- the reference interface ----
    def step(self, action_n):
        action_n = np.asarray(action_n)
        assert action_n.shape[0] == self.n and action_n.reshape(self.n, -1).shape[1] == 1, "action dimenion error!"
        self._actions.copy_(torch.from_numpy(np.ascontiguousarray(action_n.reshape(1, self.n).astype(np.int32))))
        obs, reward, n_up = self._device_step(action_n)
        self.world.update_order_param(n_up)
        done = self.world.order_param == 1.0 if self.done_callback is not None else False
        obs_n = list(obs) if self.observation_callback is not None else [np.zeros(0)] * self.n
        reward_n = list(reward) if self.reward_callback is not None else [0.0] * self.n
        if self.shared_reward:
            reward_n = [np.sum(reward_n)] * self.n
        return obs_n, reward_n, [done] * self.n, self.world.order_param, self.world.n_up, self.world.n_down

    def reset(self):
        self.reset_callback(self.world)
        self.agents = self.world.policy_agents
        if self.observation_callback is None:
            return [np.zeros(0)] * self.n
        return list(self._device_step(None)[0])

    _step, _reset = step, reset        # gym 0.9 names (environment.py:49,80)

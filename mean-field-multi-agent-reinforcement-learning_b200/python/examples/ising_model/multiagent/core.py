"""World state of the Ising environment (reference: examples/ising_model/multiagent/core.py).

The reference keeps one Python object per spin and a dense float64 `global_state`; here the lattice lives in HBM as
int8 [1, L, L] (`IsingWorld.spins`) and the per-agent objects are thin views of it, kept because the reference's
callbacks take `(agent, world)`.  `global_state` is a host mirror, refreshed on demand."""
import numpy as np
import torch


class IsingAction(object):
    def __init__(self):
        self.a = None
        self.a_range = [0, 1]


class IsingAgentState(object):
    """id / p_pos / spin of one site; `spin` reads (and writes) the world lattice."""
    __slots__ = ("id", "p_pos", "_world", "spin_range")

    def __init__(self, world=None, index=None):
        self._world, self.id, self.p_pos, self.spin_range = world, index, None, [0, 1]

    @property
    def spin(self):
        return int(self._world.global_state.flat[self.id])

    @spin.setter
    def spin(self, value):
        self._world.set_spin(self.id, value)


class IsingAgent(object):
    def __init__(self, view_sight=1, world=None, index=None):
        self.name = "" if index is None else "agent %d" % index
        self.size, self.movable, self.color = 0.050, False, None
        self.view_sight = view_sight          # 1: the four torus neighbours (the only sight the kernels implement)
        self.spin_mask = None                 # the reference's dense N-vector mask is never materialised
        self.state = IsingAgentState(world, index)
        self.action = IsingAction()
        self.action_callback = None


class IsingWorld(object):
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("the Ising environment needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.agents = []
        self.n_agents, self.agent_view_sight = 1, 1
        self.dim_pos, self.dim_spin, self.dim_color = 2, 2, 3
        self.shape_size = 1
        self.moment, self.field, self.temperature, self.interaction = 1, None, .1, 1
        self.order_param, self.order_param_delta = 1.0, 0.01
        self.n_up = self.n_down = 0
        self.spins = None                     # int8 [1, L, L] on the device: THE lattice
        self._host = None                     # float64 [L, L] mirror; None = stale

    # ---- the lattice ----
    def allocate(self, shape_size):
        self.shape_size = shape_size
        self.spins = torch.zeros((1, shape_size, shape_size), dtype=torch.int8, device=self.device)
        self._host = np.zeros((shape_size, shape_size))

    def upload(self, host_spins):
        """host lattice (any integer/float array [L, L] of 0/1) -> device"""
        self._host = np.asarray(host_spins, dtype=np.float64).reshape(self.shape_size, self.shape_size).copy()
        self.spins.copy_(torch.from_numpy(self._host.astype(np.int8)).view_as(self.spins))

    def invalidate_host(self):
        self._host = None

    @property
    def global_state(self):
        """float64 [L, L] of 0. / 1. as in the reference (core.py:72), downloaded when stale"""
        if self._host is None:
            self._host = self.spins[0].to(torch.float64).cpu().numpy()
        return self._host

    def set_spin(self, index, value):
        host = self.global_state
        host.flat[index] = 1.0 if value else 0.0
        self.spins.view(-1)[index] = 1 if value else 0

    # ---- reference properties ----
    @property
    def entities(self):
        return self.agents

    @property
    def policy_agents(self):
        return [agent for agent in self.agents if agent.action_callback is None]

    @property
    def scripted_agents(self):
        return [agent for agent in self.agents if agent.action_callback is not None]

    def update_order_param(self, n_up):
        """core.py:106-110"""
        self.n_up = int(n_up)
        self.n_down = self.n_agents - self.n_up
        self.order_param = abs(self.n_up - self.n_down) / (self.n_agents + 0.0)

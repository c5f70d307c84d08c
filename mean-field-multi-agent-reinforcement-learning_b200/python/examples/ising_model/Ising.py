"""The Ising scenario (reference: examples/ising_model/Ising.py): L x L torus, each spin sees its four neighbours,
reward 0.5 * sigma_i * sum_nbr sigma_j.  `make_world` / `reset_world` draw the initial spins from numpy's global
generator with the reference's call sequence (one np.random.choice(2) per agent, Ising.py:79-99), so a seeded script
starts from the same lattice.  The per-agent `reward` / `observation` / `done` callbacks exist for interface
compatibility; IsingMultiAgentEnv recognises them and answers for all agents with one kernel instead."""
import numpy as np

from examples.ising_model.multiagent.core import IsingAgent, IsingWorld


class Scenario():
    def make_world(self, num_agents=100, agent_view=1, device=None):
        if agent_view != 1:
            raise NotImplementedError("the B200 Ising environment implements agent_view = 1 (four torus neighbours), "
                                      "the only value main_MFQ_Ising.py uses")
        world = IsingWorld(device=device)
        world.agent_view_sight = agent_view
        world.n_agents = num_agents
        side = int(np.ceil(np.power(num_agents, 1.0 / world.dim_pos)))
        if side * side != num_agents or side < 3:
            raise ValueError("num_agents must be a perfect square >= 9 (got %d)" % num_agents)
        world.allocate(side)
        world.field = np.zeros((side, side))
        world.agents = [IsingAgent(view_sight=agent_view, world=world, index=i) for i in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.color = np.array([0.35, 0.35, 0.85])
            agent.state.p_pos = (np.array([i // side]), np.array([i % side]))   # what np.where(world_mat == i) returns
        self.reset_world(world)
        return world

    def reset_world(self, world):
        spins = np.empty(world.n_agents, dtype=np.int64)
        for i in range(world.n_agents):          # one draw per agent, in agent order (Ising.py:88)
            spins[i] = np.random.choice(world.dim_spin)
        world.upload(spins)
        world.update_order_param(int(spins.sum()))

    # ---- per-agent callbacks (host mirror; the environment batches them on the device) ----
    @staticmethod
    def _neighbours(agent, world):
        side = world.shape_size
        r, c = divmod(agent.state.id, side)
        ids = sorted({((r - 1) % side) * side + c, ((r + 1) % side) * side + c,
                      r * side + (c - 1) % side, r * side + (c + 1) % side})
        return np.array(ids)

    def reward(self, agent, world):
        gs = world.global_state
        sigma = 2.0 * gs.flat[agent.state.id] - 1.0
        nb = 2.0 * gs.flatten()[self._neighbours(agent, world)] - 1.0
        return np.array([0.5 * sigma * nb.sum()])

    def observation(self, agent, world):
        return world.global_state.flatten()[self._neighbours(agent, world)]

    def done(self, agent, world):
        return world.order_param == 1.0

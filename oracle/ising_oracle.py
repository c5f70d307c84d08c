"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy, float64, O(N) per step) of the reference's Ising tabular MFQ step:
  main_MFQ_Ising.py:105-134   loop body (Boltzmann action, env.step, Q update, mse)
  main_MFQ_Ising.py:55-67     boltzman_explore + np.random.choice(2, 1, p)  ==  a = [u >= cdf_0]
  examples/ising_model/Ising.py:7-58,101-118   4-neighbour torus mask, reward, observation
  examples/ising_model/multiagent/core.py:99-125   spin <- action, order parameter
The reference itself is O(N^2) per step (a dense N-vector mask per agent).

Parity pinning: the reference ships no golden vectors.  tests/test_ising_oracle.py runs the UNMODIFIED
reference classes (under oracle/ising_ref_shim: gym 0.9.2 / imp stand-ins) beside this restatement with
the same injected uniforms and requires exact float64 equality of actions, rewards, Q and the order
parameter; tests/golden/ising_*.npz hold trajectories generated that way (tests/golden/make_golden.py).
"""
import numpy as np

REWARD_TARGET = np.array([[2, -2], [1, -1], [0, 0], [-1, 1], [-2, 2]], dtype=np.float64)  # main_MFQ_Ising.py:77-81


def up_neighbours(spins):
    """s_i: number of up spins among the 4 torus neighbours (Ising.py:7-58 mask, :113-118 observation,
    main_MFQ_Ising.py:115 count_nonzero(obs == 1)).  spins: int array [..., L, L] in {0, 1}."""
    s = spins.astype(np.int64)
    return (np.roll(s, 1, -2) + np.roll(s, -1, -2) + np.roll(s, 1, -1) + np.roll(s, -1, -1))


def action_threshold(q0, q1, temperature):
    """cdf_0 of np.random.choice(2, 1, p): p_a = exp(Q_a/T)/denom (main_MFQ_Ising.py:55-65), then numpy's
    legacy choice renormalises: cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, u, 'right')."""
    e0, e1 = np.exp(q0 / temperature), np.exp(q1 / temperature)
    denom = 0 + e0 + e1
    p0, p1 = e0 / denom, e1 / denom
    return p0 / (p0 + p1)


def step(spins, Q, temperature, lr, u, update_mask=None):
    """One fused step for a batch of lattices.
    spins int [B, L, L] {0,1};  Q float64 [B, 5, L*L, 2] (plane per s, action pair per site);  u float64 [B, L*L].
    Returns (new_spins, new_Q, info) with info = dict(action, s, reward, n_up, order, reward_sum, mse)."""
    B, L, _ = spins.shape
    N = L * L
    s = up_neighbours(spins).reshape(B, N)                                   # on the OLD lattice
    idx = np.arange(N)
    q0 = np.take_along_axis(Q[:, :, :, 0], s[:, None, :], axis=1)[:, 0, :]
    q1 = np.take_along_axis(Q[:, :, :, 1], s[:, None, :], axis=1)[:, 0, :]
    a = (u >= action_threshold(q0, q1, temperature)).astype(np.int64)        # [B, N]
    new_spins = a.reshape(B, L, L)                                           # core.py:118-125
    sigma = 2 * new_spins - 1
    nb = np.roll(sigma, 1, -2) + np.roll(sigma, -1, -2) + np.roll(sigma, 1, -1) + np.roll(sigma, -1, -1)
    reward = (0.5 * sigma * nb).astype(np.float64).reshape(B, N)             # Ising.py:101-111 (sign folded)
    qsel = np.where(a == 1, q1, q0)
    qn = qsel + lr * (reward - qsel)                                         # main_MFQ_Ising.py:129-131
    upd = np.ones((B, N), bool) if update_mask is None else update_mask.reshape(B, N).astype(bool)
    new_Q = Q.copy()
    bb, ii = np.nonzero(upd)
    new_Q[bb, s[bb, ii], ii, a[bb, ii]] = qn[bb, ii]
    target = REWARD_TARGET[s, a]
    mse = np.where(upd, (qn - target) ** 2, 0.0).sum(axis=1) / N             # main_MFQ_Ising.py:127-134
    n_up = new_spins.reshape(B, N).sum(axis=1)
    order = np.abs(n_up - (N - n_up)) / (N + 0.0)                            # core.py:106-110
    info = dict(action=a, s=s, reward=reward, n_up=n_up, order=order, reward_sum=reward.sum(axis=1), mse=mse,
                threshold=action_threshold(q0, q1, temperature))
    return new_spins.astype(spins.dtype), new_Q, info


def temperature_schedule(t, current_t, floor, decay_rate=0.99, decay_gap=2000):
    """main_MFQ_Ising.py:108-112."""
    if t % decay_gap == 0:
        current_t *= decay_rate
    if current_t < floor:
        current_t = floor
    return current_t

"""GPU parity tests for the Ising MFQ kernel (K6) against the numpy oracle (oracle/ising_oracle.py),
which is itself pinned against the unmodified reference classes (tests/test_ising_oracle.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "oracle"))
import ising_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def run_pair(B, L, steps, T, dtype, seed, act_rate=1.0, lr=0.1):
    from mfmarl_b200 import IsingMFQ
    rng = np.random.RandomState(seed)
    spins = rng.randint(0, 2, size=(B, L, L)).astype(np.int8)
    m = IsingMFQ(B, L, dtype=dtype, spins=torch.from_numpy(spins), lr=lr)
    Q = np.zeros((B, 5, L * L, 2))
    np_dtype = np.float64 if dtype == torch.float64 else np.float32
    for t in range(steps):
        u = rng.random_sample((B, L * L)).astype(np_dtype)
        mask = None
        if act_rate < 1.0:
            mask = np.zeros((B, L * L), np.uint8)
            for b in range(B):
                mask[b, rng.choice(L * L, int(act_rate * L * L), replace=False)] = 1
        yield t, m, spins, Q, u, mask, rng


@pytest.mark.parametrize("L,T", [(20, 0.8), (20, 2.0), (7, 0.8), (33, 1.2), (64, 0.297), (3, 0.8), (256, 0.8), (512, 1.0)])
def test_fp64_trajectory_matches_oracle_exactly(L, T):
    """fp64 mode reproduces the reference precision: identical actions/spins for the whole trajectory with
    injected uniforms, Q equal to ~1 ulp (CUDA exp vs libm exp may differ in the last bit)."""
    B, steps = (3, 40) if L <= 64 else (2, 6)      # smallest (3) and largest (512 in fp64) supported sides included
    for t, m, spins, Q, u, mask, rng in run_pair(B, L, steps, T, torch.float64, seed=L):
        n_up, rsum, mse = m.step(T, uniforms=torch.from_numpy(u).cuda())
        new_spins, new_Q, info = ising_oracle.step(spins, Q, T, 0.1, u.astype(np.float64))
        # a draw closer to the threshold than 1e-12 may legitimately flip (last-bit exp difference)
        assert np.abs(u - info["threshold"]).min() > 1e-12
        assert np.array_equal(m.spins.cpu().numpy(), new_spins), "spins differ at step %d" % t
        np.testing.assert_allclose(m.Q.cpu().numpy(), new_Q, rtol=1e-13, atol=1e-15)
        assert np.array_equal(n_up.cpu().numpy(), info["n_up"])
        np.testing.assert_allclose(rsum.cpu().numpy(), info["reward_sum"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(mse.cpu().numpy(), info["mse"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(m.order_param().cpu().numpy(), info["order"], rtol=0, atol=0)
        spins[...] = new_spins
        Q[...] = new_Q


def _select_kernel(monkeypatch, variant):
    """variant: "0" generic resident kernel, "4"/"8" the shape-specialised fp32 kernels with 4 / 8 rows per thread
    (at 256 x 256 the default "8" is K6s, the persistent kernel with SWAR neighbour counts); a "p" prefix keeps the
    previous persistent kernel K6p (MFMARL_ISING_PERSIST=1), a "c" prefix the cluster kernel K6r (=0)."""
    if not variant:
        return
    if variant[0] in "pc":
        monkeypatch.setenv("MFMARL_ISING_PERSIST", "1" if variant[0] == "p" else "0")
        variant = variant[1:]
    monkeypatch.setenv("MFMARL_ISING_RPT", variant)


def _neighbourhood(mask):
    """sites whose reward (hence Q update) can see a site of `mask`: the site itself and its 4 torus neighbours"""
    return mask | np.roll(mask, 1, -2) | np.roll(mask, -1, -2) | np.roll(mask, 1, -1) | np.roll(mask, -1, -1)


@pytest.mark.parametrize("L,B,mode", [(20, 2, "stream"), (256, 2, "stream"), (256, 2, "0"), (256, 2, "4"), (256, 2, "8"),
                                        (256, 2, "p8"), (256, 2, "c8"), (64, 3, "8"), (128, 2, "8"), (512, 1, "")])
def test_fp32_every_sweep_from_the_oracle_state(L, B, mode, monkeypatch, capsys):
    """Production precision, production kernels (BASELINE config 5 shape, main_MFQ_Ising.py:55-67,105-134): 20 sweeps,
    each started from the oracle's state and driven by the same injected uniforms.  The fp32 decision
    u (1 + e) >= 1 uses ex2.approx, so a draw closer than 1e-5 to the fp64 threshold may legitimately fall on the other
    side: such flips are COUNTED and reported, every flip must be such a draw, and Q is checked to 1e-6 relative
    on every site whose update cannot see a flipped spin -- unconditionally.
    mode: "stream" = mfi_step (K6); "0"/"4"/"8" = mfi_run (K6r) generic / 4 / 8 rows per thread, one sweep per launch."""
    from mfmarl_b200 import IsingMFQ
    if mode != "stream":
        _select_kernel(monkeypatch, mode)
    T, sweeps = 0.8, 20
    rng = np.random.RandomState(L + 3)
    spins = rng.randint(0, 2, size=(B, L, L)).astype(np.int8)
    Q = np.zeros((B, 5, L * L, 2))
    m = IsingMFQ(B, L, dtype=torch.float32, spins=torch.from_numpy(spins))
    flips = near = checked = 0
    for t in range(sweeps):
        u = rng.random_sample((B, L * L)).astype(np.float32)
        m.spins.copy_(torch.from_numpy(spins)); m.Q.copy_(torch.from_numpy(Q.astype(np.float32)))
        Q32 = Q.astype(np.float32).astype(np.float64)     # the state the kernel actually starts from
        ud = torch.from_numpy(u).cuda()
        if mode == "stream":
            m.step(T, uniforms=ud)
        else:
            m.run([T], uniforms=ud.unsqueeze(0).contiguous(), resident=True)
        new_spins, new_Q, info = ising_oracle.step(spins, Q32, T, 0.1, u.astype(np.float64))
        close = (np.abs(u - info["threshold"]) < 1e-5).reshape(B, L, L)
        got = m.spins.cpu().numpy()
        flipped = got != new_spins
        assert not (flipped & ~close).any(), "sweep %d: an action differs although the draw is not near the threshold" % t
        flips += int(flipped.sum()); near += int(close.sum())
        ok = ~_neighbourhood(flipped).reshape(B, 1, L * L, 1)
        ok = np.broadcast_to(ok, new_Q.shape)
        np.testing.assert_allclose(m.Q.cpu().numpy()[ok], new_Q[ok], rtol=1e-6, atol=1e-7)
        checked += int(ok.sum())
        spins, Q = new_spins.astype(np.int8), new_Q
    with capsys.disabled():
        print("\n[ising fp32 %s L=%d] %d sweeps x %d sites: %d draws within 1e-5 of the threshold, %d flipped; "
              "Q checked at 1e-6 on %d entries" % (mode, L, sweeps, B * L * L, near, flips, checked))
    assert flips <= near
    assert flips <= max(4, near // 2)       # the fp32 decision is far more accurate than the 1e-5 window


@pytest.mark.parametrize("L,B,mode", [(256, 2, "8"), (256, 11, "8"), (256, 2, "p8"), (256, 2, "4"), (256, 2, "0"), (64, 2, "8"),
                                        (20, 3, "")])
def test_fp32_resident_run_of_20_sweeps_follows_the_oracle_trajectory(L, B, mode, monkeypatch, capsys):
    """K = 20 sweeps in ONE launch of the resident kernel against the fp64 oracle trajectory.  The uniforms are fixed
    up front: the oracle walks its own trajectory and every draw that comes closer than 1e-5 to its threshold is
    moved 1e-3 away (counted), so the two trajectories cannot part on a rounding tie.  Then spins and per-sweep up
    counts must be EQUAL and Q within 1e-6 relative everywhere, with no exception."""
    from mfmarl_b200 import IsingMFQ
    _select_kernel(monkeypatch, mode)
    T, K = 0.8, 20
    rng = np.random.RandomState(L + 11)
    spins0 = rng.randint(0, 2, size=(B, L, L)).astype(np.int8)
    spins, Q = spins0.copy(), np.zeros((B, 5, L * L, 2))
    us, ups, moved = [], [], 0
    for k in range(K):
        u = rng.random_sample((B, L * L)).astype(np.float32)
        thr = ising_oracle.step(spins, Q, T, 0.1, u.astype(np.float64))[2]["threshold"]
        close = np.abs(u - thr) < 1e-5
        u = np.where(close, np.where(thr < 0.5, thr + 1e-3, thr - 1e-3), u).astype(np.float32)
        moved += int(close.sum())
        spins, Q, info = ising_oracle.step(spins, Q, T, 0.1, u.astype(np.float64))
        assert np.abs(u - info["threshold"]).min() > 1e-5
        us.append(u); ups.append(info["n_up"])
    m = IsingMFQ(B, L, dtype=torch.float32, spins=torch.from_numpy(spins0))
    n_up, _ = m.run([T] * K, uniforms=torch.from_numpy(np.stack(us)).cuda().contiguous(), resident=True)
    assert np.array_equal(n_up.cpu().numpy(), np.stack(ups))
    assert np.array_equal(m.spins.cpu().numpy(), spins)
    np.testing.assert_allclose(m.Q.cpu().numpy(), Q, rtol=1e-6, atol=1e-7)
    with capsys.disabled():
        print("\n[ising fp32 resident %r L=%d] %d sweeps in one launch: spins equal, Q within 1e-6 on all %d entries "
              "(%d near-threshold draws moved)" % (mode, L, K, Q.size, moved))


def test_act_rate_mask_and_lr():
    for t, m, spins, Q, u, mask, rng in run_pair(2, 20, 25, 0.8, torch.float64, seed=9, act_rate=0.6, lr=0.25):
        m.step(0.8, uniforms=torch.from_numpy(u).cuda(), update_mask=torch.from_numpy(mask).cuda())
        new_spins, new_Q, info = ising_oracle.step(spins, Q, 0.8, 0.25, u, update_mask=mask)
        assert np.array_equal(m.spins.cpu().numpy(), new_spins)
        np.testing.assert_allclose(m.Q.cpu().numpy(), new_Q, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(m.mse.cpu().numpy(), info["mse"], rtol=1e-10, atol=1e-14)
        spins[...] = new_spins
        Q[...] = new_Q


def test_philox_is_deterministic_and_independent_of_sharding():
    from mfmarl_b200 import IsingMFQ
    rng = np.random.RandomState(1)
    spins = rng.randint(0, 2, size=(6, 32, 32)).astype(np.int8)
    full = IsingMFQ(6, 32, seed=5, spins=torch.from_numpy(spins))
    again = IsingMFQ(6, 32, seed=5, spins=torch.from_numpy(spins))
    shard = IsingMFQ(2, 32, seed=5, lattice_base=4, spins=torch.from_numpy(spins[4:]))
    for t in range(30):
        for m in (full, again, shard):
            m.step(0.9)
    assert torch.equal(full.spins, again.spins) and torch.equal(full.Q, again.Q)
    assert torch.equal(full.spins[4:], shard.spins) and torch.equal(full.Q[4:], shard.Q)
    assert not torch.equal(full.spins[0], full.spins[1])


def test_low_temperature_orders_and_q_converges():
    """Physics sanity at the reference's headline setting (20x20, tau = 0.8, paper Fig. 5) with the
    production Philox draws: the lattices order and Q approaches the reward table, at the same pace as the
    CPU oracle driven by numpy uniforms (mean order ~0.49, mse ~0.14 after 1500 sweeps of 16 lattices)."""
    from mfmarl_b200 import IsingMFQ
    m = IsingMFQ(64, 20, seed=13)
    first = None
    for t in range(1500):
        m.step(0.8)
        if t == 0:
            first = (float(m.order_param().mean()), float(m.mse.mean()))
    order, mse = float(m.order_param().mean()), float(m.mse.mean())
    assert first[0] < 0.1 and first[1] > 0.8
    assert 0.35 < order < 0.7, order
    assert mse < 0.25, mse


def test_full_size_256x256_properties():
    """BASELINE config 5 geometry (256 x 256) on a bounded batch: size-independent properties.
    spins stay in {0,1}; n_up equals the spin sum; each sweep touches exactly one Q entry per site;
    reward_sum equals the bond sum identity sum_i r_i = sum over the 2N bonds of sigma_i sigma_j."""
    from mfmarl_b200 import IsingMFQ
    B, L = 8, 256
    m = IsingMFQ(B, L, seed=13)
    for t in range(5):
        q_before = m.Q.clone()
        m.step(0.8)
        s = m.spins.to(torch.int64)
        assert int(s.min()) >= 0 and int(s.max()) <= 1
        assert torch.equal(m.n_up.to(torch.int64), s.sum(dim=(1, 2)))
        changed = (m.Q != q_before).sum(dim=(1, 3))          # per site: how many of its 10 entries moved
        assert int(changed.max()) <= 1
        sig = (2 * s - 1).to(torch.float32)
        bonds = (sig * torch.roll(sig, 1, 1)).sum(dim=(1, 2)) + (sig * torch.roll(sig, 1, 2)).sum(dim=(1, 2))
        assert torch.allclose(m.reward_sum, bonds, rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("L,B,K,variant", [(20, 5, 30, ""), (48, 2, 9, ""), (64, 3, 12, "0"), (64, 3, 12, "4"),
                                            (64, 2, 5, "8"), (128, 2, 8, "0"), (128, 2, 8, "4"), (128, 2, 7, "8"),
                                            (256, 3, 6, "0"), (256, 3, 6, "4"), (256, 2, 9, "8"), (256, 2, 1, "4"),
                                            (256, 2, 1, "8"), (256, 2, 2, "8"), (256, 12, 5, "8"), (256, 3, 7, "p8"),
                                            (256, 3, 7, "c8"), (64, 2, 5, "c8"), (64, 5, 33, "8"), (128, 2, 7, "p8"),
                                            (128, 2, 7, "c8"), (128, 40, 3, "8"), (512, 3, 5, ""), (512, 1, 34, ""),
                                            (64, 2, 4100, "8")])
def test_resident_kernel_equals_streaming_bit_for_bit(L, B, K, variant, monkeypatch):
    """K sweeps in one launch with Q resident in shared memory (cluster of 1 / 1 / 4 / 16 CTAs per lattice, halo
    rows pushed through DSMEM) give exactly the spins, Q and per-sweep statistics of K streaming launches (same
    Philox keys).  `variant` picks the kernel: "0" the generic one, "4"/"8" the shape-specialised fp32 kernel with
    4 / 8 rows per thread (MFMARL_ISING_RPT); "" leaves the default.  K = 4100 exceeds the persistent kernels' sweeps per
    launch (the temperature table in shared memory): mfi_run splits it into two launches."""
    from mfmarl_b200 import IsingMFQ
    _select_kernel(monkeypatch, variant)
    rng = np.random.RandomState(L)
    spins = torch.from_numpy(rng.randint(0, 2, size=(B, L, L)).astype(np.int8))
    a = IsingMFQ(B, L, seed=21, spins=spins)
    b = IsingMFQ(B, L, seed=21, spins=spins)
    assert a.resident_cluster == {20: 1, 48: 1, 64: 1, 128: 4, 256: 16, 512: 64}[L]
    temps = [max(0.8, 0.3 * 0.99)] * K
    for rounds in range(2):            # two launches: state carries over (step counter, spins, Q)
        n_res, r_res = a.run(temps, resident=True)
        n_str, r_str = b.run(temps, resident=False)
        assert torch.equal(a.spins, b.spins)
        assert torch.equal(a.Q, b.Q)
        assert torch.equal(n_res, n_str)
        assert torch.allclose(r_res, r_str, rtol=1e-6, atol=1e-3)   # float sums, different reduction order
    assert not torch.equal(a.spins[0], spins[0].cuda())


@pytest.mark.parametrize("L,variant", [(20, ""), (64, "8"), (256, "8"), (256, "p8"), (256, "0"), (512, "")])
def test_resident_kernel_with_act_groups_equals_streaming(L, variant, monkeypatch):
    """act_rate < 1 (main_MFQ_Ising.py:126): only a random subset of the sites updates Q each sweep -- the per-sweep
    masks go through the resident kernel exactly as through K streaming launches."""
    from mfmarl_b200 import IsingMFQ
    _select_kernel(monkeypatch, variant)
    B, K = 2, 7
    gen = torch.Generator(device="cuda"); gen.manual_seed(L)
    masks = (torch.rand((K, B, L * L), generator=gen, device="cuda") < 0.6).to(torch.uint8).contiguous()
    a, b = IsingMFQ(B, L, seed=4), IsingMFQ(B, L, seed=4)
    n_res, r_res = a.run([0.7] * K, update_mask=masks, resident=True)
    n_str, r_str = b.run([0.7] * K, update_mask=masks, resident=False)
    assert torch.equal(a.spins, b.spins) and torch.equal(a.Q, b.Q) and torch.equal(n_res, n_str)
    c = IsingMFQ(B, L, seed=4)
    c.run([0.7] * K, resident=True)
    assert not torch.equal(c.Q, a.Q)                      # the mask really held back updates
    # a masked-out site keeps its Q row for that sweep: with a single sweep, untouched sites stay exactly 0
    d = IsingMFQ(B, L, seed=4)
    d.run([0.7], update_mask=masks[:1].contiguous(), resident=True)
    touched = (d.Q != 0).sum(dim=(1, 3))                  # [B, N]
    assert int((touched * (1 - masks[0].to(torch.int64))).sum()) == 0


def test_resident_kernel_matches_oracle_fp64_with_injected_uniforms():
    from mfmarl_b200 import IsingMFQ
    B, L, K, T = 3, 20, 25, 0.8
    rng = np.random.RandomState(77)
    spins = rng.randint(0, 2, size=(B, L, L)).astype(np.int8)
    m = IsingMFQ(B, L, dtype=torch.float64, spins=torch.from_numpy(spins))
    u = rng.random_sample((K, B, L * L))
    n_up, rsum = m.run([T] * K, uniforms=torch.from_numpy(u).cuda(), resident=True)
    Q = np.zeros((B, 5, L * L, 2))
    for k in range(K):
        spins, Q, info = ising_oracle.step(spins, Q, T, 0.1, u[k])
        assert np.array_equal(n_up[k].cpu().numpy(), info["n_up"])
        np.testing.assert_allclose(rsum[k].cpu().numpy(), info["reward_sum"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(m.spins.cpu().numpy(), spins)
    np.testing.assert_allclose(m.Q.cpu().numpy(), Q, rtol=1e-13, atol=1e-15)


def test_cli_chunked_quiet_mode_equals_stepwise(capsys):
    """`python -m mfmarl_b200.ising --quiet` runs --chunk sweeps per launch (resident kernel) and applies the reference's
    stagnation stop to the per-sweep up counts: same MaxO / step as the one-launch-per-step loop, with the schedule of
    main_MFQ_Ising.py:103-112."""
    from mfmarl_b200 import ising
    args = ["-n", "400", "-t", "0.25", "-ts", "700", "-dg", "50", "-dr", "0.97"]
    stepwise = ising.run(args)
    out = capsys.readouterr().out
    assert out.count("E: 0/") >= 100 and "Order" in out
    chunked = ising.run(args + ["--quiet", "--chunk", "64"])
    assert chunked == stepwise
    # act_rate < 1: the act groups come from the same generator in both modes when the chunk is one sweep
    assert ising.run(args + ["-ac", "0.5", "-ts", "60"]) == ising.run(args + ["-ac", "0.5", "-ts", "60", "--quiet", "--chunk", "1"])
    assert len(ising.run(args + ["-ac", "0.5", "-ts", "130", "--quiet", "--chunk", "50"])) == 1
    sched = ising.temperature_schedule(0, 120, 0.97, 50, 0.25)
    cur, want = 0.3, []          # independent restatement of :103-112
    for t in range(120):
        if t % 50 == 0:
            cur *= 0.97
        if cur < 0.25:
            cur = 0.25
        want.append(cur)
    assert sched == want and ising.temperature_schedule(60, 60, 0.97, 50, 0.25) == want[60:]


@pytest.mark.parametrize("L,B,resident", [(256, 3, False), (256, 11, True), (512, 3, True), (128, 2, True), (64, 2, True), (20, 2, False),
                                          (20, 2, True), (7, 2, False)])
def test_production_draws_are_philox4x32_10_with_the_documented_keys(L, B, resident):
    """include/mfmarl_batched.h: key (seed, lattice_base + lattice), counter (column, row / 4, step, 0), output word
    row % 4, u = word / 2^32.  With Q = 0 both actions are equally likely and a = [u >= 1/2] is the word's top bit, so
    the lattice after one sweep IS the Philox4x32-10 stream -- computed here in numpy from the published round function
    (tests/philox_ref.py, pinned to the Random123 known-answer vectors)."""
    from mfmarl_b200 import IsingMFQ
    from philox_ref import philox4x32_10
    seed, base, step0 = 77, 5, 3
    m = IsingMFQ(B, L, seed=seed, lattice_base=base)
    m.t = step0
    if resident:
        m.run([0.8], resident=True)
    else:
        m.step(0.8)
    got = m.spins.cpu().numpy()
    x = np.arange(L, dtype=np.uint32)[None, None, :]
    band = (np.arange(L, dtype=np.uint32) // 4)[None, :, None]
    lat = (base + np.arange(B, dtype=np.uint32))[:, None, None]
    ctr = [np.broadcast_to(a, (B, L, L)).astype(np.uint32) for a in (x, band, np.uint32(step0), np.uint32(0))]
    key = [np.broadcast_to(np.uint32(seed), (B, L, L)).astype(np.uint32), np.broadcast_to(lat, (B, L, L)).astype(np.uint32)]
    out = philox4x32_10(ctr, key)                                # four words per (lattice, row, column)
    comp = (np.arange(L) % 4)[None, :, None]
    word = np.choose(np.broadcast_to(comp, (B, L, L)), out)
    assert np.array_equal(got, (word >> 31).astype(np.int8))


@pytest.mark.parametrize("L,T", [(256, 0.05), (64, 0.02), (512, 0.05)])
def test_persistent_kernel_at_very_low_temperature_equals_streaming(L, T):
    """Deep in the ordered phase exp((q1 - q0) / T) overflows to +inf (or underflows to 0) in fp32 for most sites: the
    decision u (1 + e) >= 1 must still come out the same in every kernel (they share the function), the lattice orders
    locally (aligned neighbours: the reward sum, 0 for a random lattice and 2 N for a uniform one, is well above 0; the
    global order parameter stays small while the domains coarsen), and nothing turns into NaN."""
    from mfmarl_b200 import IsingMFQ
    B, K = 3, 40
    a, b = IsingMFQ(B, L, seed=31), IsingMFQ(B, L, seed=31)
    n_res, r_res = a.run([T] * K, resident=True)
    n_str, r_str = b.run([T] * K, resident=False)
    assert torch.equal(a.spins, b.spins) and torch.equal(a.Q, b.Q) and torch.equal(n_res, n_str)
    assert torch.isfinite(a.Q).all()
    assert float(r_res[-1].min()) > 0.5 * L * L, r_res[-1]

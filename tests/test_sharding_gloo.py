"""CPU, world_size 2 over gloo: the host-side multi-GPU logic of bench.py / the sharding contract
(SURVEY.md section 8e): contiguous env ranges per rank, max-over-ranks timing, sum-over-ranks work,
and only rank 0 reporting.  No collective touches the env data path."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def shard_range(total, world, rank):
    """Rank r owns the contiguous env-id range [r * per, (r + 1) * per) (weak scaling: per is fixed)."""
    per = total // world
    return rank * per, (rank + 1) * per


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(8192, world, rank)
    # each rank "measures" its own time and work; the job reports max time and summed work
    t = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([(hi - lo) * 128, lo], dtype=torch.int64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        out.put((t.tolist(), c.tolist(), (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_reduction_contract():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    [p.start() for p in procs]
    res = out.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    t, c, (lo, hi) = res
    assert t == [11.0, 20.0]                    # max over ranks
    assert c[0] == 8192 * 128                   # whole-job agent-steps per lockstep step
    assert (lo, hi) == (0, 4096)


def test_shards_are_contiguous_disjoint_and_cover():
    for world in (1, 2, 4, 8):
        ranges = [shard_range(4096 * world, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 4096 * world
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(hi - lo == 4096 for lo, hi in ranges)


def test_reference_arm_only_rank0_reports(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 silently."""
    import subprocess
    import sys
    from conftest import REPO
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "3",
                        "--warmup", "3"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _grad_worker(rank, world, port, out):
    """two ranks train the same MF-Q net on different halves of a batch with grad_sync: identical parameters after
    the step, and equal to one process stepping on the whole batch (masked-mean loss, equal mask counts)."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mfmarl_b200.algo import MFQ
    from mfmarl_b200.algo.base import ValueNet

    class Spaces:
        def get_view_space(self, h): return (13, 13, 7)
        def get_feature_space(self, h): return (34,)
        def get_action_space(self, h): return (21,)

    torch.manual_seed(0)
    m = MFQ("m", 0, Spaces(), 10, device="cpu")
    m.grad_sync = True
    rng = np.random.RandomState(0)
    n = 16
    view, feat = rng.rand(n, 13, 13, 7).astype(np.float32), rng.rand(n, 34).astype(np.float32)
    prob = rng.dirichlet(np.ones(21), size=n).astype(np.float32)
    acts, tq = rng.randint(0, 21, n).astype(np.int32), rng.randn(n).astype(np.float32)
    masks = np.ones(n, bool)
    sl = slice(rank * n // world, (rank + 1) * n // world)
    ValueNet.train(m, state=[view[sl], feat[sl]], target_q=tq[sl], prob=prob[sl], acts=acts[sl], masks=masks[sl])
    flat = torch.cat([p.detach().reshape(-1) for p in m.eval_net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.manual_seed(0)
        solo = MFQ("s", 0, Spaces(), 10, device="cpu")
        ValueNet.train(solo, state=[view, feat], target_q=tq, prob=prob, acts=acts, masks=masks)
        ref = torch.cat([p.detach().reshape(-1) for p in solo.eval_net.parameters()])
        out.put((bool(torch.equal(gathered[0], gathered[1])), float((gathered[0] - ref).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_optional_gradient_allreduce_matches_single_process_training():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, out)) for r in range(world)]
    [p.start() for p in procs]
    same, err = out.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert same                      # both ranks hold the same parameters after the synchronised step
    assert err < 2e-6                # ... and they are the single-process result on the whole batch (Adam step 1e-4)

"""Render trace: env.set_render_dir + env.render write config.json and video_N.txt in the
format of RenderGenerator.cc:57-185.  The CUDA engine's files must equal, byte for byte, the ones the unmodified
reference engine writes for the same seeded fight stream (and the committed golden copy of those)."""
import os

import numpy as np
import pytest

from engines import CudaEngine, RefEngine, have_ref
from scenarios import fight_actions, generate_map_positions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_trace")


def write_trace(engine_cls, out_dir, steps=12, episodes=2):
    os.makedirs(out_dir, exist_ok=True)
    eng = engine_cls(40)
    eng.env.set_render_dir(out_dir)
    rng = np.random.RandomState(5)
    left, right = generate_map_positions(40)
    for ep in range(episodes):
        eng.reset(); eng.add_agents(0, left); eng.add_agents(1, right)
        for s in range(steps):
            for g in range(2):
                eng.set_action(g, fight_actions(rng, eng.get_pos(g), 40))
            eng.step()
            if s % 2 == 0 or ep == 1:          # frames with and without a preceding un-rendered step
                eng.env.render()
            eng.clear_dead()
    return sorted(os.listdir(out_dir))


def read_all(d):
    return {name: open(os.path.join(d, name), "rb").read() for name in sorted(os.listdir(d))}


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_reference_engine_reproduces_the_golden_trace(tmp_path):
    """pins the fixture: regenerating it from the reference engine gives the committed bytes"""
    names = write_trace(RefEngine, str(tmp_path))
    assert names == sorted(os.listdir(GOLD)) == ["config.json", "video_1.txt", "video_2.txt"]
    assert read_all(str(tmp_path)) == read_all(GOLD)


@pytest.mark.gpu
def test_cuda_engine_writes_the_reference_trace(tmp_path):
    write_trace(CudaEngine, str(tmp_path))
    got, want = read_all(str(tmp_path)), read_all(GOLD)
    assert sorted(got) == sorted(want)
    for name in want:
        assert got[name] == want[name], name
    # frames carry attack events and dead agents (hp 0) at some point
    video = want["video_2.txt"].decode().splitlines()
    frames = [l for l in video if l.startswith("F ")]
    assert len(frames) == 12 and any(int(l.split()[2]) > 0 for l in frames)

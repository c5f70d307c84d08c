"""CPU: the product path has no CPU fallback -- without a CUDA device (or without the library) it fails
loudly instead of computing something else."""
import os
import subprocess
import sys

import pytest

from conftest import PKG

SO = os.path.join(PKG, "build", "libmagent.so")
PY = os.path.join(PKG, "python")


def run(code, **env):
    e = dict(os.environ, PYTHONPATH=PY, CUDA_VISIBLE_DEVICES="", **env)
    return subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=300)


@pytest.mark.skipif(not os.path.exists(SO), reason="library not built")
def test_single_env_abi_aborts_without_a_gpu():
    r = run("import magent; magent.GridWorld('battle', map_size=40)")
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.skipif(not os.path.exists(SO), reason="library not built")
def test_batched_abi_reports_the_error_without_a_gpu():
    code = ("import ctypes, mfmarl_b200.lib as L\n"
            "lib = L.load_library(); cfg = L.MfbConfig(); lib.mfb_default_config(ctypes.byref(cfg))\n"
            "h = ctypes.c_void_p(); rc = lib.mfb_create(ctypes.byref(cfg), ctypes.byref(h))\n"
            "print(rc, lib.mfb_last_error().decode())")
    r = run(code)
    assert r.returncode == 0 and r.stdout.startswith("-1") and "no CUDA device" in r.stdout


def test_missing_library_is_a_hard_error(tmp_path):
    r = run("import magent; magent.GridWorld('battle', map_size=40)", MAGENT_LIB=str(tmp_path / "nope.so"))
    assert r.returncode != 0 and "not found" in r.stderr and "no CPU fallback" in r.stderr
    r = run("import torch, mfmarl_b200; mfmarl_b200.BatchedGridWorld(2)")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.skipif(not os.path.exists(SO), reason="library not built")
def test_return_codes_are_not_dropped_by_the_binding():
    """MAGENT_ERRORS=return turns the abort into a return code; the binding must then raise with the engine's message
    (the reference's binding ignores every code)."""
    r = run("import magent\n"
            "try:\n"
            "    magent.GridWorld('battle', map_size=40)\n"
            "except RuntimeError as ex:\n"
            "    print('raised:', ex)\n", MAGENT_ERRORS="return")
    assert r.returncode == 0, r.stderr[-500:]
    assert "raised:" in r.stdout and "env_new_game failed" in r.stdout and "no CUDA device" in r.stdout

"""pytest wiring: the `gpu` marker and import paths for the package's python/ tree."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "mean-field-multi-agent-reinforcement-learning_b200")
for p in (os.path.join(PKG, "python"), REPO, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

# the reference engine is only reproducible single-threaded
os.environ.setdefault("OMP_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")

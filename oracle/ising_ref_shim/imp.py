"""TEST INFRASTRUCTURE: `imp.load_source` (removed in Python 3.12) for examples/ising_model/__init__.py."""
import importlib.util


def load_source(name, pathname):
    spec = importlib.util.spec_from_file_location(name or "_ising_scenario", pathname)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module

"""examples.ising_model -- the Ising environment of the reference (examples/ising_model/__init__.py:1-7) over the
B200 kernels.  `load("Ising.py").Scenario()` is the entry point main_MFQ_Ising.py:29 uses.

`examples/` itself has no __init__.py, here as in the reference: it is a namespace package, so whichever tree comes first
on sys.path provides `examples.ising_model` (this one for the CUDA environment)."""
import importlib
import importlib.util
import os.path as osp


def load(name):
    """The reference loads `name` as a source file next to this package (imp.load_source).  The built-in scenario is
    returned as the module it already is; any other file is executed from its path."""
    if osp.basename(name) == "Ising.py" and not osp.isabs(name):
        return importlib.import_module(__name__ + ".Ising")
    path = name if osp.isabs(name) else osp.join(osp.dirname(__file__), name)
    spec = importlib.util.spec_from_file_location("ising_scenario", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module

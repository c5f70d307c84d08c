"""ctypes loader for the engine library.

Reference: examples/battle_model/python/magent/c_lib.py:13-55 loads `<pkg>/../../build/libmagent.so`
with RTLD_GLOBAL and exposes `_LIB` plus three ndarray->pointer helpers.  Same here, with two
additions: the path can be overridden with $MAGENT_LIB, and `load(path)` returns an independent handle
(tests load the reference engine beside the CUDA one).  A missing library is a hard error: the CUDA
engine is the product, there is nothing to fall back to.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "build", "libmagent.so"))


def declare_abi(lib):
    """Attach argtypes/restype for the runtime_api.h symbols (include/mfmarl_magent.h).

    The reference declares none (ints default to C int); declaring them lets plain Python ints,
    c_int32 handles and byref() results all pass unchanged, and catches arity mistakes early.
    """
    vp, ci, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
    sig = {
        "env_new_game": [ctypes.POINTER(vp), cp],
        "env_delete_game": [vp],
        "env_config_game": [vp, cp, vp],
        "env_reset": [vp],
        "env_get_observation": [vp, ci, vp],
        "env_set_action": [vp, ci, vp],
        "env_step": [vp, vp],
        "env_get_reward": [vp, ci, vp],
        "env_get_info": [vp, ci, cp, vp],
        "env_render": [vp],
        "env_render_next_file": [vp],
        "gridworld_register_agent_type": [vp, cp, ci, vp, vp],
        "gridworld_new_group": [vp, cp, vp],
        "gridworld_add_agents": [vp, ci, ci, cp, vp, vp, vp],
        "gridworld_clear_dead": [vp],
        "gridworld_set_goal": [vp, ci, cp, vp],
        "gridworld_define_agent_symbol": [vp, ci, ci, ci],
        "gridworld_define_event_node": [vp, ci, ci, vp, ci],
        "gridworld_add_reward_rule": [vp, ci, vp, vp, ci, ctypes.c_bool, ctypes.c_bool],
    }
    last_error = getattr(lib, "mfmarl_last_error", None)      # (the reference engine has no such symbol and never fails softly)
    if last_error is not None:
        last_error.restype = cp

    def raise_on_error(result, func, args):
        # The reference aborts the process on a fatal error; this engine does the same unless MAGENT_ERRORS=return
        # asks for return codes -- then a failure must not pass silently (the reference's binding ignores the codes).
        if result != 0:
            text = last_error().decode("utf-8", "replace") if last_error is not None else "error %d" % result
            raise RuntimeError("%s failed: %s" % (func.__name__, text))
        return result

    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ci
        fn.errcheck = raise_on_error
    return lib


def load(path=None):
    path = path or os.environ.get("MAGENT_LIB") or DEFAULT_LIB_PATH
    if not os.path.exists(path):
        raise OSError(
            "magent engine library not found at %s -- build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a); there is no CPU fallback" % path)
    return declare_abi(ctypes.CDLL(path, ctypes.RTLD_GLOBAL))


def as_float_c_array(buf):
    return buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def as_int32_c_array(buf):
    return buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def as_bool_c_array(buf):
    return buf.ctypes.data_as(ctypes.POINTER(ctypes.c_bool))


class _LazyLib:
    """`_LIB` resolves on first attribute access so that importing `magent` for its Config classes
    (e.g. to drive another engine build in a test) does not require the CUDA library."""

    _lib = None

    def __getattr__(self, name):
        if _LazyLib._lib is None:
            _LazyLib._lib = load()
        return getattr(_LazyLib._lib, name)


_LIB = _LazyLib()

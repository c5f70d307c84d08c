"""CPU: the built library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import PKG, REPO

SO = os.path.join(PKG, "build", "libmagent.so")


def declared_symbols():
    names = set()
    for hdr in sorted(os.listdir(os.path.join(REPO, "include"))):
        text = open(os.path.join(REPO, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        for m in re.finditer(r"\b(?:int|const char \*)\s*\*?\s*((?:env|gridworld|mfb|mfi|mfmarl)_\w+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_headers_declare_the_reference_abi():
    names = declared_symbols()
    for ref in ["env_new_game", "env_delete_game", "env_config_game", "env_reset", "env_get_observation",
                "env_set_action", "env_step", "env_get_reward", "env_get_info", "env_render",
                "env_render_next_file", "gridworld_register_agent_type", "gridworld_new_group",
                "gridworld_add_agents", "gridworld_clear_dead", "gridworld_set_goal",
                "gridworld_define_agent_symbol", "gridworld_define_event_node", "gridworld_add_reward_rule"]:
        assert ref in names


@pytest.mark.skipif(not os.path.exists(SO), reason="library not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(SO)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing

"""The battle scenario: two armies of identical `small` agents.

Reference: examples/battle_model/python/magent/builtin/config/battle.py:6-44 -- the only config
train_battle.py:85 / battle.py use.  Every constant below is part of the parity contract
(SURVEY.md section 8): 13x13x7 view, 34 features, 21 actions.
"""
import magent

SMALL = {
    "width": 1, "length": 1, "hp": 10, "speed": 2,
    "damage": 2, "step_recover": 0.1,
    "step_reward": -0.005, "kill_reward": 5, "dead_penalty": -0.1, "attack_penalty": -0.1,
}
ATTACK_BONUS = 0.2


def get_config(map_size):
    gw = magent.gridworld
    cfg = gw.Config()
    cfg.set({"map_width": map_size, "map_height": map_size})
    cfg.set({"minimap_mode": True})
    cfg.set({"embedding_size": 10})

    attrs = dict(SMALL)
    attrs["view_range"] = gw.CircleRange(6)
    attrs["attack_range"] = gw.CircleRange(1.5)
    small = cfg.register_agent_type("small", attrs)

    g0 = cfg.add_group(small)
    g1 = cfg.add_group(small)
    a = gw.AgentSymbol(g0, index="any")
    b = gw.AgentSymbol(g1, index="any")
    # shaping: landing a (non-lethal) hit on the other army pays the attacker
    cfg.add_reward_rule(gw.Event(a, "attack", b), receiver=a, value=ATTACK_BONUS)
    cfg.add_reward_rule(gw.Event(b, "attack", a), receiver=b, value=ATTACK_BONUS)
    return cfg

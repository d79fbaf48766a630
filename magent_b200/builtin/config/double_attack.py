"""Cooperative hunt: two tigers must hit the same deer in one step
(parameters of reference python/magent/builtin/config/double_attack.py:8-42)."""
from ... import gridworld as gw


def get_config(map_size):
    cfg = gw.Config()
    cfg.set({"map_width": map_size, "map_height": map_size, "embedding_size": 10})

    deer = cfg.register_agent_type("deer", dict(
        width=1, length=1, hp=5, speed=1, step_recover=0.2, kill_supply=8,
        view_range=gw.CircleRange(1), attack_range=gw.CircleRange(0)))
    tiger = cfg.register_agent_type("tiger", dict(
        width=1, length=1, hp=10, speed=1, damage=1, step_recover=-0.2,
        view_range=gw.CircleRange(4), attack_range=gw.CircleRange(1)))
    deer_group = cfg.add_group(deer)
    tiger_group = cfg.add_group(tiger)

    first = gw.AgentSymbol(tiger_group, index='any')
    second = gw.AgentSymbol(tiger_group, index='any')
    target = gw.AgentSymbol(deer_group, index='any')
    both = gw.Event(first, 'attack', target) & gw.Event(second, 'attack', target)
    cfg.add_reward_rule(both, receiver=[first, second], value=[1, 1])
    return cfg

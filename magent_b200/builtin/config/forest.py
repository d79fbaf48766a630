"""Tigers eat deer (parameters of reference python/magent/builtin/config/forest.py:6-34)."""
from ... import gridworld as gw


def get_config(map_size):
    cfg = gw.Config()
    cfg.set({"map_width": map_size, "map_height": map_size, "embedding_size": 10})

    deer = cfg.register_agent_type("deer", dict(
        width=1, length=1, hp=5, speed=1, damage=0, step_recover=0.2,
        food_supply=0, kill_supply=8,
        view_range=gw.CircleRange(1), attack_range=gw.CircleRange(0)))
    tiger = cfg.register_agent_type("tiger", dict(
        width=1, length=1, hp=10, speed=1, damage=3, step_recover=-0.5,
        food_supply=0, kill_supply=0, step_reward=1, attack_penalty=-0.1,
        view_range=gw.CircleRange(4), attack_range=gw.CircleRange(1)))
    cfg.add_group(deer)
    cfg.add_group(tiger)
    return cfg

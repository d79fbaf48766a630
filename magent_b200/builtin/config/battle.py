"""Two armies fight (parameters of reference python/magent/builtin/config/battle.py:6-35)."""
from ... import gridworld as gw

SMALL = dict(width=1, length=1, hp=10, speed=2, damage=2, step_recover=0.1,
             step_reward=-0.005, kill_reward=5, dead_penalty=-0.1, attack_penalty=-0.1)


def get_config(map_size):
    cfg = gw.Config()
    cfg.set({"map_width": map_size, "map_height": map_size,
             "minimap_mode": True, "embedding_size": 10})

    small = cfg.register_agent_type(
        "small", dict(SMALL, view_range=gw.CircleRange(6), attack_range=gw.CircleRange(1.5)))
    armies = [cfg.add_group(small), cfg.add_group(small)]

    # reward shaping: +0.2 for hitting somebody of the other army
    who = [gw.AgentSymbol(g, index='any') for g in armies]
    for me, foe in ((who[0], who[1]), (who[1], who[0])):
        cfg.add_reward_rule(gw.Event(me, 'attack', foe), receiver=me, value=0.2)
    return cfg

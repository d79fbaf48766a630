"""Predators chase prey (parameters of reference python/magent/builtin/config/pursuit.py:4-34)."""
from ... import gridworld as gw


def get_config(map_size):
    cfg = gw.Config()
    cfg.set({"map_width": map_size, "map_height": map_size})

    types = {
        "predator": dict(width=2, length=2, hp=1, speed=1, attack_penalty=-0.2,
                         view_range=gw.CircleRange(5), attack_range=gw.CircleRange(2)),
        "prey": dict(width=1, length=1, hp=1, speed=1.5,
                     view_range=gw.CircleRange(4), attack_range=gw.CircleRange(0)),
    }
    handle = {name: cfg.add_group(cfg.register_agent_type(name, attr))
              for name, attr in types.items()}

    hunter = gw.AgentSymbol(handle["predator"], index='any')
    victim = gw.AgentSymbol(handle["prey"], index='any')
    cfg.add_reward_rule(gw.Event(hunter, 'attack', victim), receiver=[hunter, victim], value=[1, -1])
    return cfg

"""Abstract environment surface (reference: python/magent/environment.py:4-43)."""


class Environment:
    """Method set every environment of the package answers to; see GridWorld for semantics."""

    def __init__(self):
        pass

    def _abstract(self, *_a, **_k):
        return None

    reset = get_observation = set_action = step = render = render_next_file = _abstract
    get_reward = get_num = get_action_space = get_view_space = get_feature_space = _abstract

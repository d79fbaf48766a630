"""Built-in game configurations (reference: python/magent/builtin/config/)."""
from . import config

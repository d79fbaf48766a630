"""Alias package: ``import magent`` -> :mod:`magent_b200` (the B200-native engine).

Lets scripts written against the reference package (``import magent``; ``magent.GridWorld``;
``magent.gridworld.Config``; ``magent.builtin.config.*``) run unchanged from the repository root.
Attribute access is forwarded lazily, sub-modules are aliased in ``sys.modules``.
"""
import sys as _sys

import magent_b200 as _impl

_prefix = _impl.__name__ + "."
for _name, _mod in list(_sys.modules.items()):
    if _name.startswith(_prefix):
        _sys.modules["magent." + _name[len(_prefix):]] = _mod


def __getattr__(name):
    return getattr(_impl, name)


def __dir__():
    return dir(_impl)

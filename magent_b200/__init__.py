"""magent_b200 -- B200-native drop-in for the MAgent grid-world hot path.

The package is the host-side mirror of the reference ``magent`` Python surface for the
GridWorld step path (observe -> act -> step -> reward -> cull).  All simulation runs in
hand-written sm_100a CUDA kernels behind the reference's own C ABI (``env_*`` / ``gridworld_*``,
reference: src/runtime_api.h:20-61) exported by ``magent_b200/lib/libmagent.so``.

``import magent`` resolves to this package through the thin alias package at the repository root
(or :func:`install_as_magent`), so ``examples/train_{battle,pursuit,gather}.py`` run unchanged.
"""
import sys as _sys

from . import utility
from . import gridworld
from . import builtin
from .c_lib import load_library

GridWorld = gridworld.GridWorld
round = utility.rec_round


def install_as_magent():
    """Register this package under the name ``magent`` (and its sub-modules) in sys.modules."""
    me = _sys.modules[__name__]
    _sys.modules.setdefault("magent", me)
    prefix = __name__ + "."
    for name, mod in list(_sys.modules.items()):
        if name.startswith(prefix):
            _sys.modules.setdefault("magent." + name[len(prefix):], mod)
    return me

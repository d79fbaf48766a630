"""Micro scenarios for compute-sanitizer: a few steps of the whole loop on tiny maps, so that memcheck /
racecheck / synccheck see every kernel and every phase of the step kernel (shared-memory scratch, atomicExch lists, aliased
shuffle / mover storage) at least once.  Usage:
    compute-sanitizer --tool racecheck python profiles/scripts/sanitize_micro.py [scenario ...]
"""
import os
import sys

os.environ.setdefault("OMP_NUM_THREADS", "1")      # the compiled reference is deterministic only single-threaded

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import parity_common as pc  # noqa: E402

SCENARIOS = {
    "dense_blocks": (lambda lib: pc.make_battle_blocks(lib, 30), 8),      # heavy attack / move contention
    "arrange": (lambda lib: pc.make_arrange(lib), 8),                     # absorbers, goals that move
    "food_mode": (lambda lib: pc.make_food(lib), 8),                      # food timelines
    "pursuit": (lambda lib: pc.make_pursuit(lib, 30, 0), 6),              # 2x2 bodies, walls
}


def main():
    names = sys.argv[1:] or ["dense_blocks", "arrange", "food_mode"]
    for name in names:
        make, steps = SCENARIOS[name]
        got = pc.run_trace(make(pc.CUDA_LIB), steps, 1, keep_obs=True)
        checker = next((p for p in (pc.REF_LIB, pc.PORT_LIB) if os.path.exists(p)), None)
        if checker:
            pc.compare_traces(pc.run_trace(make(checker), steps, 1, keep_obs=True), got, name)
        print("scenario %s: %d steps, parity ok" % (name, len(got)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-phase timeline of step_kernel_cta (profiling variant built with -DMG_PHASE_TIMING):
    MAGENT_B200_LIB=magent_b200/lib/variants/libmagent_timing.so python profiles/scripts/phase_timeline.py battle1
Thread 0 of each of the first 8 arenas stamps clock64() after every phase; this prints the median per-phase
duration over a few steps (SM cycles / sm clock)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from magent_b200.c_lib import load_library  # noqa: E402

NAMES = ["enum", "init", "attack_scan", "rng", "rank_target", "attack_relax", "apply_starve", "move_register",
         "move_relax", "move_collide", "move_clear", "move_fill", "rules", "done+flush"]


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "battle1"
    wl = dict(bench.WORKLOADS[wl_name])
    lib = load_library()
    env, act = bench.build_env(wl, lib.path, wl["arenas"])
    raw = ctypes.CDLL(lib.path)
    mhz = 1965.0          # SM max clock (MEASURED_PEAKS.json); a lightly loaded SM runs at it
    rows, sweeps, first = [], [], []
    for s in range(12):
        for h in act:
            env.set_random_actions(h, s)
        env.step()
        buf = (ctypes.c_longlong * 256)()
        raw.magent_b200_debug_phase_clocks(buf)
        t = np.array(buf[:], dtype=np.int64).reshape(8, 32)
        env.clear_dead()
        if s < 2:
            continue
        n_ar = min(8, wl["arenas"])
        for a in range(n_ar):
            d = np.diff(t[a, :14]).astype(np.float64)
            d[d < 0] = np.nan                      # phases skipped this step keep an old stamp
            rows.append(d)
            sweeps.append(t[a, 14:16])
            first.append(float(t[a, 16] - t[a, 7]))
    rows = np.array(rows)
    med = np.nanmedian(rows, axis=0)
    print("workload %s: per-arena phase durations (median over %d samples), us at %.0f MHz" % (wl_name, len(rows), mhz))
    for nm, cyc in zip(NAMES[1:], med):
        print("  %-14s %8.0f cycles  %6.2f us" % (nm, cyc, cyc / mhz))
    print("  %-14s %8.0f cycles  %6.2f us" % ("TOTAL", np.nansum(med), np.nansum(med) / mhz))
    print("  %-14s %8.0f cycles  %6.2f us   (first sweep of move_relax incl. its barrier)" % ("relax sweep 1", np.median(first), np.median(first) / mhz))
    sw = np.array(sweeps)
    print("  relaxation sweeps: attack median %d max %d, move median %d max %d" % (np.median(sw[:, 0]), sw[:, 0].max(), np.median(sw[:, 1]), sw[:, 1].max()))


if __name__ == "__main__":
    main()

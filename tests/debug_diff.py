"""debug helper: first differing observation cell between checker and CUDA engine for a scenario"""
import os, sys
os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_common as pc

def main():
    mk = lambda lib: pc.make_arrange(lib, 24, 14, n_goal=60, n_agent=150)
    other = os.environ.get("DEBUG_LIB", pc.CUDA_LIB)
    a = pc.run_trace(mk(pc.REF_LIB), 40, 14, keep_obs=True, stop_on_done=False)
    b = pc.run_trace(mk(other), 40, 14, keep_obs=True, stop_on_done=False)
    for t, (x, y) in enumerate(zip(a, b)):
        for g in x["obs"]:
            va, vb = x["obs"][g][0], y["obs"][g][0]
            if va.shape != vb.shape or not np.array_equal(va.view(np.uint32), vb.view(np.uint32)):
                bad = np.argwhere(va.view(np.uint32) != vb.view(np.uint32))
                print("step", t, "group", g, "n diffs", len(bad))
                for i, r, c, ch in bad[:12]:
                    print("  agent", i, "pos", x["pos"][g][i], "cell", (r, c), "ch", ch, "ref", va[i, r, c, ch], "got", vb[i, r, c, ch])
                # compare state
                for gg in range(len(x["num"])):
                    print("  group", gg, "pos equal", np.array_equal(x["pos"][gg], y["pos"][gg]), "num", x["num"][gg], y["num"][gg])
                return
        for gg in range(len(x["num"])):
            if not np.array_equal(x["pos_after"][gg], y["pos_after"][gg]) or not np.array_equal(x["alive"][gg], y["alive"][gg]):
                print("step", t, "state diff group", gg)
                return
    print("no diff")

if __name__ == "__main__":
    main()

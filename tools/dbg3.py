import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, edyn_b200 as E
scene = E.scenes.mixed_pile(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
w = E.scenes.build_world(scene); w.step(150)
c = w.contacts(); n = scene["dynamic"]
act = c["num"] > 0
deg = np.bincount(c["pairs"][act].reshape(-1), minlength=n + 5)[:n]
print("active manifolds", act.sum(), "points", c["num"].sum(), "pts/manifold hist", np.bincount(c["num"]))
print("degree hist", np.bincount(deg), "max", deg.max(), "colors", w.stats()["contact_colors"])
alld = np.bincount(c["pairs"].reshape(-1), minlength=n + 5)[:n]
print("all-manifold degree max", alld.max(), "mean", alld.mean())

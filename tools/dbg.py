import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, edyn_b200 as E
from oracle import oracle as O
scene = E.scenes.spheres_in_box(16, 8, 16)
w = E.scenes.build_world(scene)
for it in range(6):
    w.step(25)
    st = w.download_state(aabb=True); pairs = w.pairs()
    a, b = st["aabb"][pairs[:, 0]], st["aabb"][pairs[:, 1]]
    gap = np.maximum(a[:, 0:3] - b[:, 3:6], b[:, 0:3] - a[:, 3:6]).max(axis=1)
    i = int(gap.argmax())
    print(it, "npairs", len(pairs), "max gap", gap[i], "pair", pairs[i], "kinds", scene["bodies"]["kind"][pairs[i]], "aabbs", a[i], b[i], "n bad", (gap > 0.3).sum(), w.stats())
# box stacks
scene = E.scenes.boxes_on_plane(3)
w = E.scenes.build_world(scene)
o = O.OracleWorld(vel_iters=10, pos_iters=3); o.add_bodies(scene["bodies"])
for k in range(10):
    w.step(100); o.step(100)
    g, c = w.download_state(), o.state()
    print(k, "max pos diff", np.abs(g["pos"][:27] - c["pos"][:27]).max(), "gpu y", g["pos"][:3, 1], "cpu y", c["pos"][:3, 1])

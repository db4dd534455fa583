"""Development aid: solver time of one workload under different block settings (environment knobs of b2d_create).
usage: python tools/solver_variants.py WORKLOAD 'B2D_SLOT_CAP=0' 'B2D_BLOCKS=148' ...   (each argument = one run; '' = defaults)"""
import os, subprocess, sys, json
wl = sys.argv[1]
for var in sys.argv[2:]:
    env = dict(os.environ)
    for kv in var.split():
        k, v = kv.split("="); env[k] = v
    code = f"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench, edyn_b200 as E
class A: scale=1.0
scene = bench.make_scene('{wl}')
w = E.scenes.build_world(scene, max_manifolds=bench.capacity('{wl}', len(scene['bodies']['kind'])))
w.step(150); w.sync(); w.reset_timers()
w.step(20); w.sync()
st = w.stats()
tl = w.debug_tiles()
print('{var!s:44s}', 'step %.3f ms  solve %.3f ms' % (st['last_step_ms'] / 20, st['solve_ms']), tl, ' err', st['error_flags'])
"""
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-800:])

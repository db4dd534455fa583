"""The C++17 host adapter (edyn_b200/csrc/host/edyn_adapter.hpp: edyn::attach / make_rigidbody / update / detach over the
C ABI) executed on the device: the reference's examples/hello_world/hello_world.cpp restated over it must bring the box
to rest on the plane exactly like the Python adapter and the CPU oracle do."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, src, name):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path / name
    libdir = os.path.join(ROOT, "edyn_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "edyn_b200", "csrc", "host"),
           str(src), "-o", str(exe), "-L" + libdir, "-lb2d", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def test_cpp_hello_world_runs_on_the_device(gpu, E, O, tmp_path):
    exe = _build(tmp_path, os.path.join(ROOT, "edyn_b200", "csrc", "host", "hello_world.cpp"), "hello_b2d")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    rows = [ln for ln in r.stdout.splitlines() if ln.startswith("pos")]
    assert len(rows) == 10
    traj = np.array([[float(x) for x in ln[ln.index("(") + 1:ln.index(")")].split(",")] for ln in rows])
    # the same scene through the oracle: the printed trajectory (3 decimals) must agree sample by sample
    o = O.OracleWorld(vel_iters=8, pos_iters=3)
    scene = E.scenes.hello_world()
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in scene["bodies"].items()}
    b["orn"][:] = [0, 0, 0, 1]                       # hello_world.cpp leaves the box unrotated
    o.add_bodies(b)
    dyn = int(np.where(b["kind"] == 0)[0][0])
    ys = [o.state()["pos"][dyn].copy()]
    for i in range(202):
        o.step(1)
        ys.append(o.state()["pos"][dyn].copy())
    ys = np.array(ys)                                # ys[k] = position after k fixed steps
    for k, row in enumerate(traj):                   # printed after update number 20 k + 1 (3 decimals); edyn::update runs
        cand = ys[20 * k:20 * k + 3]                 # floor(elapsed / fixed_dt) steps, so allow the one-step phase of its clock
        assert np.abs(cand - row).max(axis=1).min() <= 1e-3, (k, row, cand)
    assert abs(traj[-1, 1] - 0.5) < 0.02

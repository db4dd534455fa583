"""Parity AT THE BENCHMARKED SIZES (BASELINE.json configs 2-5): the settled device state -- bodies, persistent contact
manifolds with their warm-start impulses, joint impulses -- is loaded into the CPU oracle (exactly what bench.py's
cpu_baseline leg does), then both sides step in lock step and every phase is compared: ordered broadphase pair list,
contact sets (point counts, attachments, lifetimes bit-equal; pivots / normals / distances <= 1e-5), island partition,
and the post-solve state <= 1e-5 with the oracle replaying the device's colour order (within a colour the constraints
touch disjoint dynamic bodies, so the coloured parallel solve IS a serial sweep in that order)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
STEP_TOL = 1e-5
LOCKSTEP = 3

FULL = {
    # name: (scene factory, settle steps on the device, manifold capacity per body)
    "boxes_4096": (lambda E: E.scenes.boxes_on_plane(16), 150, 6.0),
    "spheres_65536": (lambda E: E.scenes.spheres_in_box(), 150, 7.0),
    "mixed_262144": (lambda E: E.scenes.mixed_pile(64), 150, 9.0),
    "chains_1048576": (lambda E: E.scenes.hinge_chains(512, 512), 60, 2.0),
}


def _keys(pairs):
    p = np.asarray(pairs, np.uint64).reshape(-1, 2)
    return (p[:, 0] << np.uint64(32)) | p[:, 1]


def _oracle_from_device(O, scene, w):
    o = O.OracleWorld(vel_iters=scene["settings"]["velocity_iterations"], pos_iters=scene["settings"]["position_iterations"],
                      threads=os.cpu_count() or 1)
    o.add_bodies(scene["bodies"])
    if scene["hinges"]:
        h = scene["hinges"]
        o.add_hinges(h["a"], h["b"], h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
        o.set_hinge_impulses(w.hinge_impulses())
    if scene["exclusions"] is not None:
        o.add_exclusions(*scene["exclusions"])
    st = w.download_state(aabb=False)
    o.set_state(st["pos"], st["orn"], st["linvel"], st["angvel"])
    c = w.contacts()
    o.set_contacts(c["pairs"], c["num"], c["pts"], c["att"], c["lifetime"])
    return o


def _same_partition(a, b):
    """Island labels are names, not values: two labelings describe the same partition iff the label pairs are in bijection."""
    pa = np.unique(np.stack([a.astype(np.int64), b.astype(np.int64)], 1), axis=0)
    return len(np.unique(pa[:, 0])) == len(pa) and len(np.unique(pa[:, 1])) == len(pa)


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_lockstep_parity(gpu, E, O, name):
    make, settle, mpb = FULL[name]
    scene = make(E)
    n_all = len(scene["bodies"]["kind"])
    w = E.scenes.build_world(scene, max_manifolds=max(4096, int(mpb * n_all)))
    w.step(settle)
    o = _oracle_from_device(O, scene, w)
    exact = 0
    for s in range(LOCKSTEP):
        w.run_phases(E.world.PH_BROAD); o.run_phases(O.PH_BROAD)
        gk, ok = _keys(w.pairs()), _keys(o.pairs())
        assert len(gk) == len(ok) and np.array_equal(np.sort(gk), np.sort(ok)), f"{name} step {s}: ordered broadphase pair lists differ"
        w.run_phases(E.world.PH_NARROW); o.run_phases(O.PH_NARROW)
        gc, oc = w.contacts(), o.contacts()
        gi, oi = np.argsort(_keys(gc["pairs"])), np.argsort(_keys(oc["pairs"]))
        assert np.array_equal(_keys(gc["pairs"])[gi], _keys(oc["pairs"])[oi])
        assert np.array_equal(gc["num"][gi], oc["num"][oi]), f"{name} step {s}: contact point counts differ"
        valid = np.arange(4)[None, :] < gc["num"][gi][:, None]
        assert np.array_equal(gc["att"][gi][valid], oc["att"][oi][valid]) and np.array_equal(gc["lifetime"][gi][valid], oc["lifetime"][oi][valid])
        assert np.abs(gc["pts"][gi][valid] - oc["pts"][oi][valid]).max(initial=0) <= STEP_TOL
        w.run_phases(E.world.PH_ISLANDS); o.run_phases(O.PH_ISLANDS)
        assert _same_partition(w.islands(), o.islands()), f"{name} step {s}: island partition differs"
        w.run_phases(E.world.PH_SOLVE)
        hi, pr = w.solver_order()
        o.set_order(hi, pr)
        o.run_phases(O.PH_SOLVE)
        g, c = w.download_state(aabb=True), o.state()
        errs = {k: float(np.abs(g[k] - c[k]).max()) for k in ("pos", "orn", "linvel", "angvel", "aabb")}
        assert max(errs.values()) <= STEP_TOL, f"{name} step {s}: {errs}"
        exact += all(v == 0.0 for v in errs.values())
        o.set_state(g["pos"], g["orn"], g["linvel"], g["angvel"])
        gc = w.contacts()
        o.set_contacts(gc["pairs"], gc["num"], gc["pts"], gc["att"], gc["lifetime"])
        if scene["hinges"]:
            o.set_hinge_impulses(w.hinge_impulses())
    st = w.stats()
    assert st["error_flags"] == 0, st
    assert st["contact_points"] > scene["dynamic"] // 4


# ----------------------------------------------------------------------------- 1000 free-running steps

def _free_pair(E, O, scene, mpb):
    n_all = len(scene["bodies"]["kind"])
    w = E.scenes.build_world(scene, max_manifolds=max(4096, int(mpb * n_all)))
    o = O.OracleWorld(vel_iters=scene["settings"]["velocity_iterations"], pos_iters=scene["settings"]["position_iterations"],
                      threads=os.cpu_count() or 1)
    o.add_bodies(scene["bodies"])
    if scene["hinges"]:
        h = scene["hinges"]
        o.add_hinges(h["a"], h["b"], h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
    if scene["exclusions"] is not None:
        o.add_exclusions(*scene["exclusions"])
    return w, o


def _free_step(E, O, w, o):
    """One step on both sides with NO state copied across: the only thing the oracle takes from the device is the ORDER
    in which rows are swept (any order is a valid Gauss-Seidel sweep; the reference's own order is EnTT's pool order,
    which SURVEY A.11 could not pin), so what accumulates is the floating-point difference of the whole pipeline."""
    w.run_phases(E.world.PH_ALL)
    hi, pr = w.solver_order()
    o.run_phases(O.PH_BROAD | O.PH_NARROW | O.PH_ISLANDS)
    o.set_order(hi, pr)
    o.run_phases(O.PH_SOLVE)


def _rel_err(w, o, n):
    g, c = w.download_state(aabb=False), o.state()
    return float(np.abs(g["pos"][:n] - c["pos"][:n]).max() / np.abs(c["pos"][:n]).max()), g, c


def test_chain_slice_1000_free_running_steps(gpu, E, O):
    """north_star: positions within 1e-4 relative of the CPU stepper after 1000 steps -- a 4 096-chain slice of config 5
    (16 384 bodies, 12 288 hinges), both sides free-running from the same initial scene."""
    scene = E.scenes.hinge_chains(64, 64)
    w, o = _free_pair(E, O, scene, 2.0)
    n = scene["dynamic"]
    for _ in range(1000):
        _free_step(E, O, w, o)
    rel, g, c = _rel_err(w, o, n)
    assert rel <= 1e-4, f"relative position error after 1000 steps {rel:.3e}"
    assert float(np.abs(g["linvel"][:n] - c["linvel"][:n]).max()) <= 1e-3
    assert np.array_equal(np.sort(_keys(w.pairs())), np.sort(_keys(o.pairs())))
    assert w.stats()["error_flags"] == 0


def test_boxes_4096_1000_free_running_steps(gpu, E, O):
    """Config 2 at full size, free-running.  4 096 boxes dropped as 256 sixteen-high stacks are a CHAOTIC system: the
    stacks lean, and around step 600-800 many of them topple (peak speeds 13 m/s).  Measured (tools/free_run_drift.py):
    the two sides agree to 7e-5 relative through the drop-and-impact phase (step 50), 4e-4 at step 100 (20 bodies beyond
    1e-4), and once the first stack falls a different way the trajectories are unrelated (1e-1) -- as they are for any
    two runs of the reference that differ in one rounding.  So the 1e-4 bound is asserted where it is attainable (the
    first 50 steps); after 1000 steps the piles must agree as PILES: same resting statistics, nothing through the floor."""
    scene = E.scenes.boxes_on_plane(16)
    w, o = _free_pair(E, O, scene, 6.0)
    n = scene["dynamic"]
    for s in range(1000):
        _free_step(E, O, w, o)
        if s == 49:
            rel, _, _ = _rel_err(w, o, n)
            assert rel <= 1e-4, f"relative position error after 50 steps {rel:.3e}"
    _, g, c = _rel_err(w, o, n)
    for st in (g, c):
        assert st["pos"][:n, 1].min() > 0.45                                   # resting on the plane, not in it
        assert np.linalg.norm(st["linvel"][:n], axis=1).mean() < 0.25           # settled
    hg, hc = np.sort(g["pos"][:n, 1]), np.sort(c["pos"][:n, 1])
    assert abs(hg.mean() - hc.mean()) <= 0.03 * hc.mean()                       # same pile height profile
    assert np.abs(hg - hc).mean() <= 0.25
    cg, co = int(w.contacts()["num"].sum()), int(o.contacts()["num"].sum())
    assert abs(cg - co) <= 0.05 * co, (cg, co)                                  # same amount of contact
    assert w.stats()["error_flags"] == 0

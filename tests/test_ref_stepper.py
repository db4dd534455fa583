"""The oracle against the reference's REAL whole-step code.

oracle/_ref/libedyn_stepper.so is every translation unit of /root/reference/src/edyn (minus networking) compiled unmodified
against oracle/entt_lite, the functional stand-in for the absent EnTT dependency; oracle/ref_stepper.cpp drives
edyn::attach / make_rigidbody / make_constraint<hinge_constraint> / exclude_collision / edyn::step_simulation
(stepper_sequential.cpp:71-102).  The only thing the oracle takes from the reference is the ORDER in which the island
solver swept the rows that step (island_solver.cpp:181-222 walks island.edges, a sparse set filled by entity_graph
traversals); broadphase, narrowphase incl. contact persistence, islands, row preparation, warm starting, velocity and
position iterations, integration and the AABB / inertia refresh are all the oracle's own -- and must agree BIT FOR BIT.

  * live lock-step (needs the library: built here by `make -C oracle stepper`, shipped prebuilt to the GPU box),
  * the same comparison against trajectories recorded in tests/golden/whole_step.npz (always runs)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.golden import make_whole_step as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "whole_step.npz")


@pytest.fixture(scope="module")
def refstep(O):
    if O.ref_stepper() is None:
        pytest.skip("oracle/_ref/libedyn_stepper.so not available (needs /root/reference at build time)")
    return O


def _keys(p):
    p = np.asarray(p, np.uint64).reshape(-1, 2)
    return (p[:, 0] << np.uint64(32)) | p[:, 1]


def _oracle_step(O, o, hinge_order, contact_order):
    o.run_phases(O.PH_BROAD | O.PH_NARROW | O.PH_ISLANDS)
    o.set_point_order(hinge_order, contact_order)
    o.run_phases(O.PH_SOLVE)


def _same_partition(a, b):
    pa = np.unique(np.stack([a.astype(np.int64), b.astype(np.int64)], 1), axis=0)
    return len(np.unique(pa[:, 0])) == len(pa) and len(np.unique(pa[:, 1])) == len(pa)


def test_entt_lite_semantics(tmp_path):
    """Pool order, view order, swap-and-pop, identifier recycling, signals, paged storage: oracle/entt_lite/selftest.cpp."""
    cxx = shutil.which("g++")
    assert cxx, "g++ is part of the image"
    exe = str(tmp_path / "entt_selftest")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "oracle", "entt_lite"), "-o", exe,
                    os.path.join(ROOT, "oracle", "entt_lite", "selftest.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "entt_lite ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("name", list(G.SCENES))
def test_oracle_lockstep_with_real_stepper(refstep, E, name):
    O = refstep
    scene = G.build_scene(E, name)
    st = scene["settings"]
    r = O.RefWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"])
    o = O.OracleWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"])
    G.populate(r, scene); G.populate(o, scene)
    touched = 0
    for s in range(90):
        r.step(1)
        hi, ct = r.solver_order()
        _oracle_step(O, o, hi, ct)
        a, b = r.state(), o.state()
        for k in ("pos", "orn", "linvel", "angvel", "aabb"):
            assert np.array_equal(a[k], b[k]), f"{name} step {s}: {k} differs by {np.abs(a[k] - b[k]).max():.3e}"
        rc, oc = r.contacts(), o.contacts()
        ri, oi = np.argsort(_keys(rc["pairs"])), np.argsort(_keys(oc["pairs"]))
        assert np.array_equal(_keys(rc["pairs"])[ri], _keys(oc["pairs"])[oi]), f"{name} step {s}: manifold sets (ordered pairs) differ"
        assert np.array_equal(rc["num"][ri], oc["num"][oi]), f"{name} step {s}: point counts differ"
        assert _same_partition(r.islands(), o.islands()), f"{name} step {s}: island partition differs"
        touched = max(touched, int(rc["num"].sum()))
    assert touched > 0                                   # the scene did come to rest on its contacts
    # make_rigidbody derives the inverse inertia from mass and shape (dynamics/moment_of_inertia.cpp); the scene generators
    # (edyn_b200/rigidbody.py) must hand the device the same numbers
    dyn = np.asarray(scene["bodies"]["kind"]) == 0
    assert np.array_equal(r.inertia_inv()[dyn], np.asarray(scene["bodies"]["inv_inertia"], np.float32).reshape(-1, 9)[dyn])


def test_real_stepper_multithreaded_matches_sequential(refstep, E):
    """execution_mode::sequential_multithreaded (what bench.py's reference arm times) gives the sequential mode's results."""
    O = refstep
    scene = G.build_scene(E, "mixed_125")
    st = scene["settings"]
    worlds = [O.RefWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"], threads=t) for t in (0, 4)]
    for w in worlds:
        G.populate(w, scene)
        w.step(60)
    a, b = worlds[0].state(), worlds[1].state()
    for k in ("pos", "orn", "linvel", "angvel"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name", list(G.SCENES))
def test_oracle_replays_reference_trajectories(O, E, name):
    """Same comparison against the committed recording (tests/golden/make_whole_step.py): runs wherever the repo does."""
    g = np.load(GOLDEN)
    steps, keep = (int(v) for v in g["steps"])
    scene = G.build_scene(E, name)
    st = scene["settings"]
    o = O.OracleWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"])
    G.populate(o, scene)
    oh, oc, hoff, coff = g[f"{name}.order_h"], g[f"{name}.order_c"], g[f"{name}.order_h_off"], g[f"{name}.order_c_off"]
    for s in range(steps):
        _oracle_step(O, o, oh[hoff[s]:hoff[s + 1]], oc[coff[s]:coff[s + 1]])
        if (s + 1) % keep == 0:
            x = o.state()
            got = np.concatenate([x["pos"], x["orn"], x["linvel"], x["angvel"]], axis=1)
            want = g[f"{name}.states"][(s + 1) // keep - 1]
            assert np.array_equal(got, want), f"{name} step {s}: |oracle - reference| = {np.abs(got - want).max():.3e}"
    c = o.contacts()
    assert np.array_equal(np.sort(_keys(c["pairs"])), np.sort(_keys(g[f"{name}.final_pairs"])))
    assert int(c["num"].sum()) == int(g[f"{name}.final_num"].sum())
    assert _same_partition(o.islands(), g[f"{name}.islands"])

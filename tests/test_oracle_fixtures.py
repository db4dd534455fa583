"""CPU: the oracle restatement (and the product's host-side inertia code) against the committed golden vectors
generated from the real reference functions (tests/golden/make_golden.py), and -- where oracle/_ref is present --
against the reference library itself on fresh random inputs.  Integer/bit-level work must match exactly."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_collide_matches_reference_vectors(O):
    g = load("collide.npz")
    o = O.ora_fns()
    for i in range(len(g["num"])):
        pts, att = o.collide(int(g["kinds"][i, 0]), g["params"][i, 0], int(g["kinds"][i, 1]), g["params"][i, 1],
                             g["pos"][i, 0], g["orn"][i, 0], g["pos"][i, 1], g["orn"][i, 1])
        n = int(g["num"][i])
        assert len(pts) == n, f"case {i}: {len(pts)} points, reference {n}"
        assert np.array_equal(pts, g["pts"][i, :n]), f"case {i} differs"
        assert np.array_equal(att, g["att"][i, :n])


def test_shape_aabb_matches_reference_vectors(O):
    g = load("aabb.npz")
    o = O.ora_fns()
    for i in range(len(g["kind"])):
        assert np.array_equal(o.shape_aabb(int(g["kind"][i]), g["params"][i], g["pos"][i], g["orn"][i]), g["aabb"][i])
    for p, bb in zip(g["planes"], g["plane_aabb"]):
        assert np.array_equal(o.shape_aabb(6, p, [0, 0, 0], [0, 0, 0, 1]), bb)


def test_body_math_matches_reference_vectors(O, E):
    g = load("body_math.npz")
    o = O.ora_fns()
    from edyn_b200 import rigidbody as rb
    for i in range(len(g["q"])):
        assert np.array_equal(o.integrate(g["q"][i], g["w"][i], float(g["dt"][i])), g["q_out"][i])
        I = o.moment_of_inertia(int(g["kind"][i]), g["params"][i], float(g["mass"][i]))
        assert np.array_equal(I.reshape(9), g["moi"][i])
        assert np.array_equal(o.inverse_symmetric(I).reshape(9), g["inv_inertia"][i])
        assert np.array_equal(o.world_inertia(g["q"][i], g["inv_inertia"][i]).reshape(9), g["inv_inertia_world"][i])
        # the product's host adapter (make_rigidbody mirror) must stage the same inertia_inv
        Ih = rb.moment_of_inertia(rb.Shape(int(g["kind"][i]), tuple(g["params"][i])), g["mass"][i])
        assert np.array_equal(Ih.reshape(9), g["moi"][i]), f"host moment_of_inertia differs at {i}"
        assert np.array_equal(rb.inverse_matrix_symmetric(Ih).reshape(9), g["inv_inertia"][i])


def test_rows_match_reference_vectors(O):
    g = load("rows.npz")
    o = O.ora_fns()
    for i in range(len(g["J"])):
        out = o.prepare_row(g["J"][i], g["inv_mA"][i], g["inv_IA"][i], g["inv_mB"][i], g["inv_IB"][i], g["error"][i], 0.2,
                            g["restitution"][i], g["vel"][i])
        assert np.array_equal(out, g["prepared"][i])
        d, r = o.solve_row(g["J"][i], g["row5"][i], g["dv"][i])
        assert d == g["delta"][i] and r[4] == g["impulse"][i]
        p, q = o.plane_space(g["normal"][i])
        assert np.array_equal(np.concatenate([p, q]), g["plane_space"][i])
    for i in range(len(g["hinge_J"])):
        h = g["hinge_params"][i]
        n, J = O.hinge_rows("ora", h[0:3], h[3:6], h[6:9], h[9:12], h[12:15], h[15:19], h[19:22], h[22:26])
        assert n == 5 and np.array_equal(J, g["hinge_J"][i])


def test_friction_rows_match_reference_vectors(O):
    """solve_friction and warm_start of constraint_row_friction (constraint_row_friction.cpp:11-66): 600 committed cases
    produced by the reference's own object code, two thirds of them clamped to the friction circle, some at zero load."""
    g = load("friction.npz")
    o = O.ora_fns()
    clamped = 0
    for i in range(len(g["J"])):
        imp, dv = o.solve_friction(g["J"][i], g["fr"][i], g["mu"][i], g["normal_impulse"][i], g["masses"][i], g["dv"][i])
        assert np.array_equal(imp, g["impulse"][i]) and np.array_equal(dv, g["dv_out"][i]), i
        _, wdv = o.solve_friction(g["J"][i], g["fr"][i], g["mu"][i], g["normal_impulse"][i], g["masses"][i], g["dv"][i], warm=True)
        assert np.array_equal(wdv, g["dv_warm"][i]), i
        lim = g["mu"][i] * g["normal_impulse"][i]
        clamped += bool(lim > 0 and abs(np.hypot(*imp) - lim) <= 1e-5 * max(1.0, lim))
    assert clamped > 300


def test_contact_rows_and_position_solve_match_reference_vectors(O):
    """contact_constraint::prepare and ::solve_position with position_solver::solve (contact_constraint.cpp:15-90,
    position_solver.hpp:16-51): 500 committed cases produced by the reference's own object code."""
    g = load("contacts.npz")
    o = O.ora_fns()
    assert 150 < int(g["solved"].sum()) < 400, "both the penetrating and the separated branch are exercised"
    for i in range(len(g["cp15"])):
        p = o.contact_prepare(g["cp15"][i], 1.0 / 60, g["bodyA23"][i], g["bodyB23"][i])
        for k in ("nJ", "n5", "fJ", "fr6"):
            assert np.array_equal(p[k], g[k][i]), (i, k)
        assert p["mu"] == g["mu"][i]
        solved, a, b, out = o.contact_solve_position(g["cp13"][i], g["bodyA26"][i], g["bodyB26"][i])
        assert solved == g["solved"][i] and np.array_equal(a, g["outA"][i]) and np.array_equal(out, g["out5"][i]), i
        if g["bodyB26"][i][7] != 0:        # a static second body: the reference re-normalises its orientation, the oracle leaves it alone
            assert np.array_equal(b, g["outB"][i]), i


def test_manifold_decisions_match_reference_vectors(O):
    """find_nearest_contact, find_nearest_contact_rolling, should_remove_point (util/collision_util.cpp:233-280, :397-413):
    the per-point decisions of process_collision, 800 committed cases around the caching / breaking thresholds."""
    g = load("manifold.npz")
    o = O.ora_fns()
    for i in range(len(g["n"])):
        n = int(g["n"][i])
        assert o.find_nearest_contact(g["cpA"][i], g["cpB"][i], g["resA"][i, :n], g["resB"][i, :n]) == g["nearest"][i], i
        assert o.find_nearest_contact_rolling(g["resA"][i, :n], g["cpA"][i], g["origin"][i], g["orn"][i], g["angvel"][i], 1.0 / 60) == g["nearest_rolling"][i], i
        assert o.should_remove_point(g["cpA"][i], g["pivB"][i], g["normal"][i], g["origin"][i], g["orn"][i], g["posB"][i], g["ornB"][i]) == bool(g["remove"][i]), i
    found, kept = (g["nearest"] < g["n"]).mean(), 1 - g["remove"].mean()
    assert 0.5 < found < 0.98 and 0.3 < kept < 0.8, "both outcomes of every decision are exercised"


def test_hinge_position_material_mix_aabb_test_match_reference_vectors(O):
    """hinge_constraint::solve_position (hinge_constraint.cpp:180-213, three position_solver::solve calls per hinge),
    material mixing (material_mixing.hpp:12-18) and intersect(AABB, AABB) (geom.cpp:762-770, closed intervals: boxes
    that touch on a face intersect, one ulp apart they do not)."""
    g = load("misc.npz")
    o = O.ora_fns()
    for i in range(len(g["hinge"])):
        e, a, b = o.hinge_solve_position(g["hinge"][i], g["bodyA26"][i], g["bodyB26"][i])
        assert e == g["hinge_err"][i] and np.array_equal(a, g["outA"][i]) and np.array_equal(b, g["outB"][i]), i
        assert np.array_equal(o.material_mix(*g["materials"][i]), g["mixed"][i])
        assert o.intersect_aabb(g["aabb_a"][i], g["aabb_b"][i]) == bool(g["hit"][i]), i
    touching = np.arange(len(g["hit"])) % 3 == 0
    ulp_off = np.arange(len(g["hit"])) % 6 == 0
    assert g["hit"][touching & ~ulp_off].any() and not g["hit"][ulp_off].all()


def _oracle_island_labels(O, E, static, edges):
    from edyn_b200.rigidbody import RigidBodyDef, bodies_soa, sphere_shape
    n = len(static)
    defs = [RigidBodyDef(kind=E.STATIC if static[i] else E.DYNAMIC, position=(3.0 * i, 0, 0), mass=1.0, shape=sphere_shape(0.1)) for i in range(n)]
    o = O.OracleWorld()
    o.add_bodies(bodies_soa(defs, (0.0, 0.0, 0.0)))
    m = len(edges)
    o.set_contacts(edges, np.zeros(m, np.uint32), np.zeros((m, 4, 18), f32), np.zeros((m, 4), np.uint32))
    o.run_phases(O.PH_ISLANDS)
    return o.islands()


def test_island_partition_matches_reference_entity_graph(O, E):
    """The oracle's island labels == entity_graph::connected_components of the reference (core/entity_graph.cpp) on 60
    committed random graphs (908 components): static nodes belong to no island and do not connect their neighbours."""
    g = load("graphs.npz")
    off = eo = 0
    for n, ne in zip(g["sizes"].tolist(), g["edge_counts"].tolist()):
        st, ed, lab = g["static"][off:off + n], g["edges"][eo:eo + ne], g["labels"][off:off + n]
        assert np.array_equal(_oracle_island_labels(O, E, st, ed), lab)
        assert (lab[st != 0] == 0xFFFFFFFF).all()
        off += n; eo += ne


def _broadphase_cases():
    g = load("broadphase.npz")
    keys = [k[4:] for k in g.files if k.startswith("soa_")]
    off = po = 0
    for n, c in zip(g["sizes"].tolist(), g["counts"].tolist()):
        yield {k: g["soa_" + k][off:off + n] for k in keys}, g["aabb1"][off:off + n], {tuple(p) for p in g["pairs"][po:po + c].tolist()}
        off += n; po += c


def test_broadphase_pairs_match_reference_trees(O):
    """The oracle's ORDERED broadphase pair lists == the pair search of broadphase::update run around the reference's real
    dynamic AABB trees (collision/dynamic_tree.cpp: fat leaves created elsewhere and moved, SAH insertion, rotations):
    12 committed crowded scenes, 5 578 pairs.  What stays an assumption is the order in which EnTT visits the bodies."""
    total = 0
    for soa, aabb, want in _broadphase_cases():
        o = O.OracleWorld()
        o.add_bodies(soa)
        assert np.array_equal(o.state()["aabb"], aabb)
        o.run_phases(O.PH_BROAD)
        assert {tuple(p) for p in o.pairs().tolist()} == want
        total += len(want)
    assert total > 5000


def _rq(rng):
    q = rng.normal(size=4)
    return (q / np.linalg.norm(q)).astype(f32)


def test_random_against_reference_library(O, ref):
    """Fresh random inputs each run of the suite (seeded): restatement == reference, bit for bit."""
    o = O.ora_fns()
    rng = np.random.default_rng(7)
    kinds = [0, 2, 3]
    from tests.golden.make_golden import shape_params
    for it in range(1500):
        ka, kb = int(rng.choice(kinds)), int(rng.choice(kinds + [6]))
        pA, pB = shape_params(ka, rng), shape_params(kb, rng)
        posA, posB = (rng.random(3) * 0.7).astype(f32), (rng.random(3) * 0.7).astype(f32)
        qa, qb = _rq(rng), _rq(rng)
        if kb == 6:
            posB, qb = np.zeros(3, f32), np.array([0, 0, 0, 1], f32)
        a, b = o.collide(ka, pA, kb, pB, posA, qa, posB, qb), ref.collide(ka, pA, kb, pB, posA, qa, posB, qb)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (ka, kb, it)
        a, b = o.collide(kb, pB, ka, pA, posB, qb, posA, qa), ref.collide(kb, pB, ka, pA, posB, qb, posA, qa)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (kb, ka, it)
    for it in range(500):
        p1, q1, p2, q2 = [rng.normal(size=3).astype(f32) for _ in range(4)]
        if it % 4 == 0:
            q2 = (p2 + (q1 - p1) * f32(rng.random() + 0.2)).astype(f32)          # parallel segments
        a, b = o.closest_segment_segment(p1, q1, p2, q2), ref.closest_segment_segment(p1, q1, p2, q2)
        assert a[0] == b[0] and a[2] == b[2]
        k = 8 if b[0] < 2 else 16
        assert np.array_equal(a[1][:k], b[1][:k])
        pa = (rng.normal(size=(6, 3)) * 0.3).astype(f32)
        pb = (rng.normal(size=(6, 3)) * 0.3).astype(f32)
        if it % 3 == 0:
            pa[:, 1] = 0                                                             # coplanar: area/collinearity rules
        a, b = o.maybe_add_points(pa, pb), ref.maybe_add_points(pa, pb)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    from tests.golden.make_golden import broadphase_scene
    for _ in range(4):
        soa = broadphase_scene(rng, int(rng.integers(40, 200)), int(rng.integers(0, 4)))
        ow = O.OracleWorld()
        ow.add_bodies(soa)
        bb = ow.state()["aabb"]
        ow.run_phases(O.PH_BROAD)
        want = O.ref_broadphase_pairs(bb, bb, (soa["kind"] == 0).astype(np.uint8))
        assert {tuple(p) for p in ow.pairs().tolist()} == {tuple(p) for p in want.tolist()}
    from tests.golden.make_golden import graph_inputs
    import edyn_b200 as E
    for st, ed in graph_inputs(rng, 25):
        assert np.array_equal(_oracle_island_labels(O, E, st, ed), O.ref_connected_components(st, ed))
    from tests.golden.make_golden import misc_inputs
    hinge, bA, bB, mats, a6, b6 = misc_inputs(rng, 300)
    for i in range(len(hinge)):
        x, y = o.hinge_solve_position(hinge[i], bA[i], bB[i]), ref.hinge_solve_position(hinge[i], bA[i], bB[i])
        assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]), i
        assert np.array_equal(o.material_mix(*mats[i]), ref.material_mix(*mats[i]))
        assert o.intersect_aabb(a6[i], b6[i]) == ref.intersect_aabb(a6[i], b6[i])
    from tests.golden.make_golden import manifold_decision_inputs
    n, cpA, cpB, resA, resB, origin, orn, angvel, posB, ornB, normal = manifold_decision_inputs(rng, 600)
    for i in range(len(n)):
        k = int(n[i])
        assert o.find_nearest_contact(cpA[i], cpB[i], resA[i, :k], resB[i, :k]) == ref.find_nearest_contact(cpA[i], cpB[i], resA[i, :k], resB[i, :k])
        assert (o.find_nearest_contact_rolling(resA[i, :k], cpA[i], origin[i], orn[i], angvel[i], 1.0 / 60)
                == ref.find_nearest_contact_rolling(resA[i, :k], cpA[i], origin[i], orn[i], angvel[i], 1.0 / 60))
        piv = (cpB[i] * np.float32(0.1)).astype(f32)
        assert (o.should_remove_point(cpA[i] * np.float32(0.05), piv, normal[i], origin[i], orn[i], origin[i], ornB[i])
                == ref.should_remove_point(cpA[i] * np.float32(0.05), piv, normal[i], origin[i], orn[i], origin[i], ornB[i]))
    from tests.golden.make_golden import contact_inputs
    cp15, bA23, bB23, cp13, bA26, bB26 = contact_inputs(rng, 400)
    for i in range(len(cp15)):
        a, b = o.contact_prepare(cp15[i], 1.0 / 60, bA23[i], bB23[i]), ref.contact_prepare(cp15[i], 1.0 / 60, bA23[i], bB23[i])
        assert all(np.array_equal(a[k], b[k]) for k in ("nJ", "n5", "fJ", "fr6")) and a["mu"] == b["mu"], i
        a, b = o.contact_solve_position(cp13[i], bA26[i], bB26[i]), ref.contact_solve_position(cp13[i], bA26[i], bB26[i])
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3]), i
        if bB26[i][7] != 0:
            assert np.array_equal(a[2], b[2]), i
    from tests.golden.make_golden import friction_inputs
    J, fr, mu, nimp, masses, dv = friction_inputs(rng, 800)
    for i in range(len(J)):
        for warm in (False, True):
            a, b = o.solve_friction(J[i], fr[i], mu[i], nimp[i], masses[i], dv[i], warm), ref.solve_friction(J[i], fr[i], mu[i], nimp[i], masses[i], dv[i], warm)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (i, warm)


def test_golden_generator_is_committed():
    assert os.path.exists(os.path.join(GOLD, "make_golden.py"))
    for f in ("collide.npz", "aabb.npz", "body_math.npz", "rows.npz", "friction.npz", "contacts.npz", "manifold.npz", "misc.npz", "graphs.npz", "broadphase.npz"):
        assert os.path.exists(os.path.join(GOLD, f))

"""The oracle against the reference's own known-answer tests.

Each case restates a gtest of /root/reference/test without gtest; the cited file:line is the source of the
expected values.  ASSERT_SCALAR_EQ is ASSERT_FLOAT_EQ, i.e. 4 ULP (test/edyn/common/common.hpp:7-17).
"""
import numpy as np

f32 = np.float32
EPS = np.finfo(np.float32).eps
BOX, PLANE = 3, 6


def ulp_close(a, b, ulps=4):
    a, b = f32(a), f32(b)
    return abs(a - b) <= ulps * np.spacing(max(abs(a), abs(b), f32(1e-30)))


def _match_points(got, expected):
    expected = [np.array(e, f32) for e in expected]
    for g in got:
        hit = [i for i, e in enumerate(expected) if np.linalg.norm(g - e) < EPS]
        assert hit, f"unexpected point {g}"
        expected.pop(hit[0])
    assert not expected


def test_collide_box_box_face_face(O):
    """test/edyn/collision/test_collision.cpp:13-43: 4 points at the +-0.5 corners, y = 0.5."""
    o = O.ora_fns()
    pts, att = o.collide(BOX, [0.5, 0.5, 0.5, 0], BOX, [0.5, 0.5, 0.5, 0], [0, 0, 0], [0, 0, 0, 1], [0, 1.0, 0], [0, 0, 0, 1])
    assert len(pts) == 4
    _match_points(pts[:, 0:3], [(0.5, 0.5, 0.5), (-0.5, 0.5, 0.5), (-0.5, 0.5, -0.5), (0.5, 0.5, -0.5)])


def test_collide_box_box_face_edge(O):
    """test/edyn/collision/test_collision.cpp:45-90: pivotA {(+-0.5, 0.5, 0)}, pivotB {(+-0.5, -0.5, 0.5)}."""
    o = O.ora_fns()
    a = np.pi / 4
    ornB = [np.sin(a / 2), 0, 0, np.cos(a / 2)]                      # quaternion_axis_angle({1,0,0}, pi/4)
    pts, att = o.collide(BOX, [0.5, 0.5, 0.5, 0], BOX, [0.5, 0.5, 0.5, 0], [0, 0, 0], [0, 0, 0, 1], [0, 1.2, 0], ornB)
    assert len(pts) == 2
    _match_points(pts[:, 0:3], [(0.5, 0.5, 0), (-0.5, 0.5, 0)])
    _match_points(pts[:, 3:6], [(0.5, -0.5, 0.5), (-0.5, -0.5, 0.5)])


def test_intersect_line_aabb(O):
    """test/edyn/math/test_geom.cpp:3-58, five cases."""
    o = O.ora_fns()
    n, s = o.intersect_line_aabb([0, 0.5], [1, 1.5], [-1, -0.5], [2, 1])
    assert n == 2 and ulp_close(s[0], -1) and ulp_close(s[1], 0.5)
    n, s = o.intersect_line_aabb([2, 1], [1, 1.5], [-1, -0.5], [2, 1])
    assert n == 1 and ulp_close(s[0], 0)
    n, s = o.intersect_line_aabb([0, 0], [1, 1.5], [-1, 0.5], [0, 1])
    assert n == 0
    n, s = o.intersect_line_aabb([1, 1], [1, -1], [-2, -0.5], [1, 0.5])
    assert n == 2 and ulp_close(s[0], 0.75) and ulp_close(s[1], 0.25)
    n, s = o.intersect_line_aabb([0, -0.25], [1, -0.25], [-2, -0.5], [1, 0.5])
    assert n == 2 and ulp_close(s[0], -2) and ulp_close(s[1], 1)


def _body(E, **kw):
    d = E.RigidBodyDef(**kw)
    return d


def test_apply_gravity(O, E):
    """test/edyn/sys/test_apply_gravity.cpp:4-23: 10 x apply_gravity(dt = 0.1666) == g * dt * n (4 ULP)."""
    dt, n = 0.1666, 10
    w = O.OracleWorld(dt=dt, vel_iters=1, pos_iters=0)
    soa = E.bodies_soa([_body(E, shape=None, inertia=np.eye(3, dtype=f32), gravity=(0, -9.8, 0))])
    w.add_bodies(soa)
    for _ in range(n):
        w.run_phases(O.PH_SOLVE)
    v = w.state()["linvel"][0]
    g = np.array([0, -9.8, 0], f32)
    expect = g * f32(dt) * f32(n)
    for k in range(3):
        assert ulp_close(v[k], expect[k]), (v, expect)


def test_collision_filter_truth_table(O, E):
    """test/edyn/collision/test_broadphase.cpp:16-31 (should_collide_default with filters and an exclusion)."""
    box = E.box_shape((0.5, 0.5, 0.5))
    defs = [_body(E, shape=box, collision_group=0x1, collision_mask=(~0x2) & 0xFFFFFFFFFFFFFFFF),
            _body(E, shape=box, collision_group=0x2, collision_mask=(~0x1) & 0xFFFFFFFFFFFFFFFF),
            _body(E, shape=box)]
    w = O.OracleWorld()
    w.add_bodies(E.bodies_soa(defs))
    first, second, third = 0, 1, 2
    assert not w.should_collide(first, second)
    assert w.should_collide(first, third)
    assert w.should_collide(second, third)
    w.add_exclusions([first], [third])
    assert not w.should_collide(first, second)
    assert not w.should_collide(first, third)
    assert w.should_collide(second, third)
    # without filters everything collides (the first three assertions of the reference test)
    w2 = O.OracleWorld()
    w2.add_bodies(E.bodies_soa([_body(E, shape=box) for _ in range(3)]))
    assert w2.should_collide(0, 1) and w2.should_collide(0, 2) and w2.should_collide(1, 2)


def test_connected_components(O, E):
    """test/edyn/core/test_entity_graph.cpp:52-125: nodes {0, 1} joined by two edges + a lone node -> 2 components
    of sizes {2 nodes, 2 edges} and {1, 0}."""
    w = O.OracleWorld()
    w.add_bodies(E.bodies_soa([_body(E, shape=None, inertia=np.eye(3, dtype=f32)) for _ in range(3)]))
    z = [[0, 0, 0]] * 2
    ax = [[0, 0, 1]] * 2
    w.add_hinges([0, 0], [1, 1], z, z, ax, ax)          # two parallel edges between node 0 and node 1
    w.run_phases(O.PH_ISLANDS)
    lab = w.islands()
    assert lab[0] == lab[1] and lab[2] != lab[0]
    comps = {}
    for i, l in enumerate(lab):
        comps.setdefault(int(l), []).append(i)
    assert sorted(len(v) for v in comps.values()) == [1, 2]


def test_hello_world_rests_on_plane(O, E):
    """the box comes to rest at y = 0.5 on four contact points."""
    s = E.scenes.hello_world()
    w = O.OracleWorld(vel_iters=8, pos_iters=3)
    w.add_bodies(s["bodies"])
    w.step(300)
    st, c = w.state(), w.contacts()
    assert abs(st["pos"][0, 1] - 0.5) < 1e-3
    assert np.abs(st["linvel"][0]).max() < 1e-3
    assert c["num"].tolist() == [4]


def test_island_sleeping_and_wake_up(E, O):
    """put_islands_to_sleep / wake_up_island restated (island_manager.cpp:524-623): a resting box falls asleep after
    island_time_to_sleep = 2 s below the velocity thresholds, is frozen while asleep, wakes the step a falling box makes
    a manifold with it (new edge into a sleeping island), and both go back to sleep together.  Off by default."""
    scene = E.scenes.sleep_and_wake()
    o = O.OracleWorld(vel_iters=10, pos_iters=3)
    o.add_bodies(scene["bodies"])
    o.step(200)
    assert not o.sleeping().any(), "sleeping is opt-in (benchmark configurations carry sleeping_disabled)"
    o = O.OracleWorld(vel_iters=10, pos_iters=3)
    o.add_bodies(scene["bodies"])
    o.set_sleeping(True)
    events, prev, frozen = [], np.zeros(2, bool), None
    for k in range(420):
        o.step(1)
        s = o.sleeping()[:2]
        if (s != prev).any():
            events.append((k, s.tolist()))
            prev = s
        st = o.state()
        if s[0]:
            if frozen is None:
                frozen = st["pos"][0].copy()
            assert np.array_equal(st["pos"][0], frozen) and not st["linvel"][0].any(), "a sleeping body is not simulated"
        else:
            frozen = None
    # dt = 1/60: the box is below the thresholds from the first update; timestamp taken at update 1 (time of update 0),
    # sleep_dt first exceeds 2 s at update 122
    assert events[0] == (121, [True, False])
    assert events[1][1] == [False, False] and 190 < events[1][0] < 205, "the landing box wakes the sleeper"
    assert events[2][1] == [True, True] and events[2][0] - events[1][0] > 120, "both rest for 2 s before sleeping again"
    assert len(events) == 3
    lab = o.islands()
    assert lab[0] == lab[1] == 0
    # wake_up_entity
    o.wake_bodies([1])
    o.step(1)
    assert not o.sleeping()[:2].any(), "waking one member wakes the island at the next update"


def _three_overlapping_boxes(E):
    from edyn_b200.rigidbody import RigidBodyDef, bodies_soa, box_shape
    defs = [RigidBodyDef(position=(0.1 * i, 0, 0), mass=1.0, shape=box_shape((0.2, 0.2, 0.2))) for i in range(3)]
    return bodies_soa(defs, (0.0, 0.0, 0.0))


def test_collision_exclusion_add_remove_clear(O, E):
    """test/edyn/collision/test_exclusion.cpp restated on behaviour: exclude e0-e1 and e0-e2, remove e1-e0, clear e0;
    after each change the broadphase makes exactly the pairs that are not excluded."""
    def pairs_after(excl_ops):
        o = O.OracleWorld()
        o.add_bodies(_three_overlapping_boxes(E))
        for op, a, b in excl_ops:
            (o.add_exclusions if op == "+" else o.remove_exclusions)([a], [b])
        o.run_phases(O.PH_BROAD)
        return {tuple(sorted(p)) for p in o.pairs().tolist()}
    assert pairs_after([]) == {(0, 1), (0, 2), (1, 2)}
    assert pairs_after([("+", 0, 1), ("+", 0, 2)]) == {(1, 2)}
    assert pairs_after([("+", 0, 1), ("+", 0, 2), ("-", 1, 0)]) == {(0, 1), (1, 2)}, "removal is symmetric in the pair"
    assert pairs_after([("+", 0, 1), ("+", 0, 2), ("-", 1, 0), ("-", 0, 2)]) == {(0, 1), (0, 2), (1, 2)}


def test_issue_76_destroy_then_recreate(O, E):
    """test/edyn/issues/issue76.cpp: make a static floor, destroy it, make it again, update -- must simply work, and the
    destroyed floor must no longer hold anything up."""
    from edyn_b200.rigidbody import RigidBodyDef, bodies_soa, box_shape, plane_shape
    floor = bodies_soa([RigidBodyDef(kind=E.STATIC, shape=plane_shape((0, 1, 0), 0.0))], (0.0, -9.8, 0.0))
    box = bodies_soa([RigidBodyDef(position=(0, 0.5, 0), mass=1.0, shape=box_shape((0.5, 0.5, 0.5)))], (0.0, -9.8, 0.0))
    o = O.OracleWorld()
    f0 = o.add_bodies(floor)
    o.remove_bodies([f0])
    o.add_bodies(box)
    o.step(30)
    assert o.state()["pos"][1, 1] < 0.0 and len(o.pairs()) == 0, "nothing left to rest on"
    f1 = o.add_bodies(floor)
    st = o.state()
    o.set_state(np.array([[0, 0, 0], [0, 0.5, 0], [0, 0, 0]], np.float32), st["orn"], np.zeros((3, 3), np.float32), np.zeros((3, 3), np.float32))
    o.step(30)
    assert abs(o.state()["pos"][1, 1] - 0.5) < 1e-3 and {tuple(p) for p in o.pairs().tolist()} == {(1, f1)}

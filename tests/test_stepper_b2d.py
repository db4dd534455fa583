"""The reference-side binding, for real: edyn_b200/csrc/host/stepper_b2d.hpp compiled against the reference's own
headers (and oracle/entt_lite for EnTT) steps a REAL Edyn registry -- built with edyn::attach, make_rigidbody,
make_constraint<hinge_constraint>, exclude_collision -- through the C ABI of include/b2d.h.

CPU suite: the C ABI is answered by tests/integration/b2d_mock.cpp (the oracle behind b2d.h), so what is under test is the
host logic of the binding: entity <-> body id maps, the SoA staging of make_rigidbody's components (inverse inertia,
shape parameters, filters, materials), hinge frames -> axes, exclusions added and removed, patched entities travelling
as dirty subsets, destroyed entities leaving, results scattered back into the registry's components.  The same harness
linked against the real edyn_b200/libb2d.so runs on the device in tests/test_zz_gpu_stepper_b2d.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.golden import make_whole_step as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "integration", "_build")
_f = np.float32
_u = C.c_uint32


def build_integration():
    """(Re)build where the reference is present; elsewhere the shipped libraries are used as they are."""
    if os.path.isdir("/root/reference/include/edyn") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "edyn_obj")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "integration")], check=True)


def load(O, variant):
    path = os.path.join(BUILD, f"libedyn_b2d_{variant}.so")
    if not os.path.exists(path):
        return None
    l = O.load_refs_library(path)
    l.eb2d_world.restype = C.c_void_p
    l.eb2d_body_id.restype = C.c_uint32
    l.eb2d_update.argtypes = [C.c_void_p, C.c_double]
    if variant == "mock":
        l.b2d_mock_patched.restype = C.c_uint32
        l.b2d_mock_patched.argtypes = [C.c_void_p]
    return l


class EdynB2dWorld:
    """A real Edyn registry (O.RefWorld's construction path) stepped by edyn::stepper_b2d."""

    def __init__(self, O, lib, scene, attach_first=False, max_manifolds=1 << 16, sleeping=False, **kw):
        st = scene["settings"]
        self.lib, self.sleeping = lib, sleeping
        self.r = O.RefWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"], library=lib, **kw)
        n = len(scene["bodies"]["kind"])
        nh = len(scene["hinges"]["a"]) if scene["hinges"] else 0
        if attach_first:
            self._attach(n, max_manifolds, nh)
        if sleeping:                                    # bodies without sleeping_disabled_tag on the registry side as well
            assert not scene["hinges"] and scene["exclusions"] is None
            self.r.add_bodies(scene["bodies"], sleeping_disabled=False)
        else:
            G.populate(self.r, scene)
        if not attach_first:
            self._attach(n, max_manifolds, nh)

    def _attach(self, n, max_manifolds, nh):
        rc = self.lib.eb2d_attach(self.r.h, 0, _u(n + 8), _u(max_manifolds), _u(max(nh, 1)), C.c_int(1 if self.sleeping else 0))
        assert rc == 0, "stepper_b2d could not be created"

    def step(self, n=1):
        assert self.lib.eb2d_step(self.r.h, _u(n)) == 0

    def state(self):
        return self.r.state()

    def state_of_survivors(self):
        n = self.r.num_bodies
        out = dict(pos=np.zeros((n, 3), _f), orn=np.zeros((n, 4), _f), linvel=np.zeros((n, 3), _f), angvel=np.zeros((n, 3), _f))
        self.lib.eb2d_get_state(self.r.h, *[out[k].ctypes.data_as(C.c_void_p) for k in ("pos", "orn", "linvel", "angvel")])
        return out

    def patch_velocity(self, body, lv, av):
        lv, av = np.asarray(lv, _f), np.asarray(av, _f)
        self.lib.eb2d_patch_velocity(self.r.h, _u(body), lv.ctypes.data_as(C.c_void_p), av.ctypes.data_as(C.c_void_p))

    def destroy_body(self, body):
        self.lib.eb2d_destroy_body(self.r.h, _u(body))

    def remove_exclusion(self, a, b):
        self.lib.eb2d_remove_exclusion(self.r.h, _u(a), _u(b))

    def wake_up(self, body):
        assert self.lib.eb2d_wake_up(self.r.h, _u(body)) == 0

    def mirror_contacts(self):
        """stepper_b2d::mirror_contacts; returns (manifolds, points, created, destroyed) of this call and the running totals
        a user's on_construct<contact_started_tag> / on_destroy<contact_point> listeners have seen."""
        out, seen = (C.c_uint32 * 4)(), (C.c_uint32 * 2)()
        assert self.lib.eb2d_mirror_contacts(self.r.h, out, seen) == 0
        return tuple(out), tuple(seen)

    def contact_entities(self, capacity=1 << 16):
        ent, life = np.zeros(capacity, np.uint32), np.zeros(capacity, np.uint32)
        self.lib.eb2d_contact_entities.restype = C.c_uint32
        k = self.lib.eb2d_contact_entities(self.r.h, _u(capacity), ent.ctypes.data_as(C.c_void_p), life.ctypes.data_as(C.c_void_p))
        return ent[:k].copy(), life[:k].copy()

    def close(self):
        if self.r is not None:
            self.lib.eb2d_detach(self.r.h)          # the stepper disconnects from the registry before the registry goes
            self.r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@pytest.fixture(scope="module")
def mock(O):
    if O.ref_stepper() is None:
        pytest.skip("needs the reference's objects (oracle/_ref)")
    build_integration()
    lib = load(O, "mock")
    if lib is None:
        pytest.skip("tests/integration/_build/libedyn_b2d_mock.so not available")
    return lib


def _plain_oracle(O, scene):
    st = scene["settings"]
    o = O.OracleWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"])
    G.populate(o, scene)
    return o


def _assert_same(a, b, what, keys=("pos", "orn", "linvel", "angvel", "aabb")):
    for k in keys:
        assert np.array_equal(a[k], b[k]), f"{what}: {k} differs by {np.abs(a[k] - b[k]).max():.3e}"


@pytest.mark.parametrize("name", list(G.SCENES))
@pytest.mark.parametrize("attach_first", [False, True])
def test_registry_through_the_binding_equals_arrays_through_the_abi(mock, O, E, name, attach_first):
    """A registry filled by make_rigidbody & co. and staged by stepper_b2d must reach the C ABI as exactly the arrays the
    scene generators hand over directly: 80 steps, component values in the registry == oracle state, bit for bit.
    attach_first: bodies arrive through the construction signals instead of the initial sweep of the registry."""
    scene = G.build_scene(E, name)
    w = EdynB2dWorld(O, mock, scene, attach_first=attach_first)
    o = _plain_oracle(O, scene)
    for s in range(80):
        w.step(1); o.step(1)
        if s % 10 == 9:
            _assert_same(w.state(), o.state(), f"{name} step {s}")
    w.close()


def test_random_registries_through_the_binding(mock, O, E):
    """The staging on broad content: random scenes (all shape kinds on dynamic, kinematic and static bodies, filters, bodies
    with their own gravity, materials, tilted planes, hinges with random axes, exclusions) built in a real registry and
    staged by stepper_b2d == the same arrays handed to the ABI directly, bit for bit."""
    from tests.test_ref_stepper import random_scene_wide
    for seed in range(12):
        scene = random_scene_wide(E, O, seed)
        w = EdynB2dWorld(O, mock, scene, attach_first=bool(seed & 1))
        o = _plain_oracle(O, scene)
        for s in range(6):
            w.step(10); o.step(10)
            _assert_same(w.state(), o.state(), f"seed {seed} step {10 * s + 9}")
        w.close()


def test_binding_follows_the_real_stepper(mock, O, E):
    """Same user code, two steppers: the reference's stepper_sequential and stepper_b2d (over the mock).  Identical while
    nothing touches (no row order involved), and the same resting pile afterwards."""
    scene = G.build_scene(E, "boxes_27")
    w = EdynB2dWorld(O, mock, scene)
    st = scene["settings"]
    r = O.RefWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"])
    G.populate(r, scene)
    w.step(5); r.step(5)
    _assert_same(w.state(), r.state(), "free fall")
    w.step(175); r.step(175)
    a, b = w.state(), r.state()
    assert np.abs(a["pos"] - b["pos"]).max() < 2e-3 and np.abs(a["linvel"]).max() < 0.05
    w.close()


def test_patched_entities_travel_as_dirty_subset(mock, O, E):
    scene = G.build_scene(E, "spheres_96")
    w = EdynB2dWorld(O, mock, scene)
    o = _plain_oracle(O, scene)
    w.step(20); o.step(20)
    handle = C.c_void_p(mock.eb2d_world(w.r.h))
    assert mock.b2d_mock_patched(handle) == 0               # our own write-back is not "dirty"
    lv, av = np.array([1.5, 2.0, -0.5], _f), np.array([0.0, 3.0, 0.0], _f)
    for body in (3, 40):
        w.patch_velocity(body, lv, av)                       # registry.patch<linvel> / <angvel>
    x = o.state()
    for body in (3, 40):
        x["linvel"][body], x["angvel"][body] = lv, av
    o.set_state(x["pos"], x["orn"], x["linvel"], x["angvel"])
    w.step(30); o.step(30)
    assert mock.b2d_mock_patched(handle) == 2                # two entities travelled, not the scene
    _assert_same(w.state(), o.state(), "after patch")
    w.close()


def test_exclusions_and_destroyed_bodies(mock, O, E):
    scene = G.build_scene(E, "chains_16")
    w = EdynB2dWorld(O, mock, scene)
    o = _plain_oracle(O, scene)
    w.step(10); o.step(10)
    a, b = int(scene["exclusions"][0][0]), int(scene["exclusions"][1][0])
    w.remove_exclusion(a, b)                                 # edyn::remove_collision_exclusion
    o.remove_exclusions([a], [b])
    w.step(10); o.step(10)
    _assert_same(w.state(), o.state(), "after removing an exclusion")
    victim = 7
    keep = np.arange(len(scene["bodies"]["kind"])) != victim
    w.destroy_body(victim)                                   # registry.destroy(entity): joints and manifolds go with it
    o.remove_bodies([victim])
    w.step(40); o.step(40)
    got, want = w.state_of_survivors(), o.state()
    for k in ("pos", "orn", "linvel", "angvel"):
        assert np.array_equal(got[k][keep], want[k][keep]), k
        assert not got[k][victim].any()                      # the entity is gone from the registry
    w.close()


def _contacts_by_pair(c, cols):
    """{(body0, body1): (points in list order restricted to `cols`)} of a contacts() dictionary."""
    out = {}
    for i, (a, b) in enumerate(np.asarray(c["pairs"]).reshape(-1, 2).tolist()):
        out[(a, b)] = np.asarray(c["pts"][i, :c["num"][i]][:, cols], np.float32)
    return out


@pytest.mark.parametrize("name", ["boxes_27", "mixed_125", "chains_16"])
def test_contacts_mirrored_into_the_registry_on_demand(mock, O, E, name):
    """mirror_contacts(): afterwards the registry reads -- through the reference's own component layout, walked the way user
    code walks it (contact_manifold_state.contact_entity -> contact_point_list.next) -- exactly like the device's manifolds:
    ordered pairs incl. those without points, points in list order, pivots / normal / distance / applied impulses /
    lifetime bit for bit.  Points keep their entity while they persist; new ones raise contact_started_tag, vanished
    ones on_destroy<contact_point>."""
    scene = G.build_scene(E, name)
    w = EdynB2dWorld(O, mock, scene)
    o = _plain_oracle(O, scene)
    PT18 = [0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 15, 16, 17]      # pivotA pivotB normal distance impulses of the 18-float device point
    PT14 = list(range(13))                                   # the same fields as refs_get_contacts lays them out
    seen_started = seen_ended = 0
    before = {}
    for stage, steps in enumerate((45, 1, 7, 40)):
        w.step(steps); o.step(steps)
        (manifolds, points, created, destroyed), (started, ended) = w.mirror_contacts()
        got, want = w.r.contacts(), o.contacts()
        assert manifolds == len(want["pairs"]) and points == int(want["num"].sum())
        g, t = _contacts_by_pair(got, PT14), _contacts_by_pair(want, PT18)
        assert g.keys() == t.keys(), f"{name} stage {stage}: manifold sets differ"
        for k in t:
            assert np.array_equal(g[k], t[k]), f"{name} stage {stage}: points of manifold {k} differ"
        life = {k: want["lifetime"][i, :want["num"][i]] for i, k in enumerate(map(tuple, want["pairs"].tolist()))}
        got_life = _contacts_by_pair(got, [13])
        for k in life:
            assert np.array_equal(got_life[k][:, 0].astype(np.uint32), life[k])
        # events as a user's listeners saw them
        assert started - seen_started == created and ended - seen_ended == destroyed
        seen_started, seen_ended = started, ended
        # identity: a point that is `steps` older than at the last look lives in the same entity
        ent, lt = w.contact_entities()
        assert len(ent) == points and len(set(ent.tolist())) == points
        now = dict(zip(ent.tolist(), lt.tolist()))
        if stage:
            kept = [e for e, l in now.items() if l >= steps]
            assert all(e in before and before[e] + steps == now[e] for e in kept), f"{name} stage {stage}: a persisting point changed entity"
            assert created == len(now) - len(kept)
        else:
            assert created == points and destroyed == 0
        before = now
    assert seen_started > 0
    w.close()


def test_sleeping_tags_follow_the_device(mock, O, E):
    """B2D_FLAG_SLEEPING: sleeping_tag on the registry's bodies after every step == the device's flags (the box falls
    asleep after 2 s, is woken by the second box landing on it, both sleep again), and stepper_b2d::wake_up."""
    scene = E.scenes.sleep_and_wake()
    w = EdynB2dWorld(O, mock, scene, sleeping=True)
    o = _plain_oracle(O, scene)
    o.set_sleeping(True)
    changes, prev = 0, np.zeros(3, bool)
    for s in range(460):
        w.step(1); o.step(1)
        tags = w.r.sleeping()
        assert np.array_equal(tags, o.sleeping().astype(bool)), f"step {s}"
        changes += not np.array_equal(tags, prev)
        prev = tags
    assert changes >= 3 and prev[:2].all()
    _assert_same(w.state(), o.state(), "after sleeping", keys=("pos", "orn", "linvel", "angvel"))
    w.wake_up(0)
    o.wake_bodies([0])
    assert not w.r.sleeping()[0]
    w.step(3); o.step(3)
    assert np.array_equal(w.r.sleeping(), o.sleeping().astype(bool))
    w.close()


def test_default_restitution_settings_select_the_restitution_solver(mock, O, E):
    """settings.num_restitution_iterations = 8 (the reference's default) -> B2D_FLAG_RESTITUTION_SOLVER; 0 -> restitution
    through the row rhs.  Seen from the registry the two give different trajectories, each equal to the oracle's."""
    scene = G.build_scene(E, "mixed_125")           # e = 0.2
    ends = []
    for iters in (0, 8):
        w = EdynB2dWorld(O, mock, scene, restitution_iters=iters)
        o = _plain_oracle(O, scene)
        o.set_restitution_iterations(iters)
        w.step(60); o.step(60)
        _assert_same(w.state(), o.state(), f"restitution iterations {iters}")
        ends.append(w.state()["pos"])
        w.close()
    assert np.abs(ends[0] - ends[1]).max() > 1e-3

"""Generates tests/golden/*.npz from the REAL reference functions (oracle/_ref/libedyn_ref.so, compiled in place
from /root/reference by oracle/Makefile).  /root/reference does not exist on the GPU box, so the vectors are
committed; re-run this script here to regenerate them:

    python tests/golden/make_golden.py

Every file holds seeded random inputs plus the reference's outputs for one function of the hot path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
SPHERE, CAPSULE, BOX, PLANE = 0, 2, 3, 6


def rq(rng):
    q = rng.normal(size=4)
    return (q / np.linalg.norm(q)).astype(f32)


def shape_params(kind, rng):
    if kind == SPHERE:
        return np.array([0.2 + 0.25 * rng.random(), 0, 0, 0], f32)
    if kind == CAPSULE:
        return np.array([0.1 + 0.1 * rng.random(), 0.15 + 0.25 * rng.random(), float(rng.integers(3)), 0], f32)
    if kind == BOX:
        return np.concatenate([0.15 + 0.35 * rng.random(3), [0]]).astype(f32)
    n = rng.normal(size=3) if rng.random() < 0.3 else np.array([0, 1, 0])
    n = n / np.linalg.norm(n)
    return np.array([n[0], n[1], n[2], 0.1 * rng.normal()], f32)


def collide_cases(rng, count):
    pairs = [(SPHERE, SPHERE), (SPHERE, PLANE), (BOX, PLANE), (CAPSULE, PLANE), (SPHERE, BOX), (CAPSULE, CAPSULE),
             (CAPSULE, SPHERE), (CAPSULE, BOX), (BOX, BOX), (PLANE, SPHERE), (BOX, SPHERE), (BOX, CAPSULE),
             (SPHERE, CAPSULE), (PLANE, BOX), (PLANE, CAPSULE)]
    out = []
    ident = np.array([0, 0, 0, 1], f32)
    for ka, kb in pairs:
        for it in range(count):
            pA, pB = shape_params(ka, rng), shape_params(kb, rng)
            posA, posB = (rng.random(3) * 0.8).astype(f32), (rng.random(3) * 0.8).astype(f32)
            qa, qb = rq(rng), rq(rng)
            if kb == PLANE:
                posB, qb = np.zeros(3, f32), ident
                posA = (pB[:3] * (pB[3] + 0.5 * rng.random()) + 0.3 * rng.normal(size=3) * (1 - np.abs(pB[:3]))).astype(f32)
            if ka == PLANE:
                posA, qa = np.zeros(3, f32), ident
                posB = (pA[:3] * (pA[3] + 0.5 * rng.random()) + 0.3 * rng.normal(size=3) * (1 - np.abs(pA[:3]))).astype(f32)
            if it % 3 == 0 and kb == BOX and ka in (BOX, CAPSULE):        # stacked / parallel: face-face, 4 points
                qa = ident
                qb = ident if it % 2 else rq(rng)
                hA = pA[1] if ka == BOX else pA[0]
                if ka == CAPSULE:
                    pA[2] = 0.0
                posB = (posA + np.array([0.1 * rng.random(), -(hA + pB[1]) + 0.005 * rng.normal(), 0.1 * rng.random()])).astype(f32)
            if it % 3 == 1 and ka == CAPSULE and kb == CAPSULE:           # parallel capsules: 2 points
                pA[2] = pB[2] = 0.0
                qa = qb = ident
                posB = (posA + np.array([0.2 * rng.normal(), pA[0] + pB[0] + 0.005 * rng.normal(), 0])).astype(f32)
            out.append((ka, pA, kb, pB, posA, qa, posB, qb))
    return out


def main():
    r = O.ref_fns()
    if r is None:
        raise SystemExit("oracle/_ref/libedyn_ref.so missing: run `make -C oracle` where /root/reference exists")
    rng = np.random.default_rng(20260922)

    # ---- collide(): 15 ordered shape pairs
    cases = collide_cases(rng, 160)
    n = len(cases)
    kinds = np.zeros((n, 2), np.uint32)
    params = np.zeros((n, 2, 4), f32)
    poses = np.zeros((n, 2, 3), f32)
    orns = np.zeros((n, 2, 4), f32)
    num = np.zeros(n, np.uint32)
    pts = np.zeros((n, 4, 10), f32)
    att = np.zeros((n, 4), np.uint32)
    for i, (ka, pA, kb, pB, posA, qa, posB, qb) in enumerate(cases):
        kinds[i] = (ka, kb); params[i, 0], params[i, 1] = pA, pB
        poses[i, 0], poses[i, 1] = posA, posB; orns[i, 0], orns[i, 1] = qa, qb
        p, a = r.collide(ka, pA, kb, pB, posA, qa, posB, qb)
        num[i] = len(p); pts[i, :len(p)] = p; att[i, :len(p)] = a
    np.savez_compressed(os.path.join(HERE, "collide.npz"), kinds=kinds, params=params, pos=poses, orn=orns, num=num, pts=pts, att=att)
    print("collide.npz:", n, "cases, point histogram", np.bincount(num, minlength=5).tolist())

    # ---- shape AABBs
    m = 600
    k = rng.choice([SPHERE, CAPSULE, BOX], size=m).astype(np.uint32)
    sp = np.stack([shape_params(int(x), rng) for x in k])
    pos = rng.normal(size=(m, 3)).astype(f32) * 3
    orn = np.stack([rq(rng) for _ in range(m)])
    bb = np.stack([r.shape_aabb(int(k[i]), sp[i], pos[i], orn[i]) for i in range(m)])
    planes = np.array([[0, 1, 0, 0], [1, 0, 0, -0.5], [-1, 0, 0, -35.5], [0, 0, 1, 0.25], [0, 0, -1, 2], [0, -1, 0, 1],
                       [0.6, 0.8, 0, 0.3]], f32)
    pbb = np.stack([r.shape_aabb(PLANE, p, [0, 0, 0], [0, 0, 0, 1]) for p in planes])
    np.savez_compressed(os.path.join(HERE, "aabb.npz"), kind=k, params=sp, pos=pos, orn=orn, aabb=bb, planes=planes, plane_aabb=pbb)

    # ---- quaternion integrate (both branches), world inertia, moment of inertia
    q = np.stack([rq(rng) for _ in range(m)])
    w = rng.normal(size=(m, 3)).astype(f32) * 4
    w[: m // 6] *= f32(1e-4)                                           # Taylor branch (|w| < 1e-3)
    dts = np.where(rng.random(m) < 0.5, 1.0 / 60, -1.0 / 60).astype(f32)
    qo = np.stack([r.integrate(q[i], w[i], float(dts[i])) for i in range(m)])
    invI = np.zeros((m, 9), f32)
    moi = np.zeros((m, 9), f32)
    mass = (0.5 + 10 * rng.random(m)).astype(f32)
    for i in range(m):
        I = r.moment_of_inertia(int(k[i]), sp[i], float(mass[i]))
        moi[i] = I.reshape(9)
        invI[i] = r.inverse_symmetric(I).reshape(9)
    iw = np.stack([r.world_inertia(q[i], invI[i]).reshape(9) for i in range(m)])
    np.savez_compressed(os.path.join(HERE, "body_math.npz"), q=q, w=w, dt=dts, q_out=qo, kind=k, params=sp, mass=mass, moi=moi,
                        inv_inertia=invI, inv_inertia_world=iw)

    # ---- constraint rows: prepare_row / solve (constraint_row.cpp), plane_space, hinge Jacobians
    J = rng.normal(size=(m, 12)).astype(f32)
    IA = np.stack([np.diag(1 + rng.random(3)).reshape(9) for _ in range(m)]).astype(f32)
    IB = np.stack([np.diag(1 + rng.random(3)).reshape(9) for _ in range(m)]).astype(f32)
    mA, mB = rng.random(m).astype(f32), rng.random(m).astype(f32)
    mB[::5] = 0; IB[::5] = 0                                          # static second body
    err, rest = rng.normal(size=m).astype(f32) * 0.1, (rng.random(m) * 0.5).astype(f32)
    vel = rng.normal(size=(m, 12)).astype(f32)
    prep = np.stack([r.prepare_row(J[i], mA[i], IA[i], mB[i], IB[i], err[i], 0.2, rest[i], vel[i]) for i in range(m)])
    row5 = np.stack([prep[:, 0], prep[:, 1], np.where(rng.random(m) < 0.5, 0, -1e30).astype(f32),
                     np.where(rng.random(m) < 0.5, 1e18, 0.05).astype(f32), (rng.normal(size=m) * 0.1).astype(f32)], axis=1).astype(f32)
    dv = rng.normal(size=(m, 12)).astype(f32) * 0.3
    sol = [r.solve_row(J[i], row5[i], dv[i]) for i in range(m)]
    delta = np.array([s[0] for s in sol], f32)
    imp = np.array([s[1][4] for s in sol], f32)
    nrm = np.stack([x / np.linalg.norm(x) for x in rng.normal(size=(m, 3))]).astype(f32)
    ps = np.stack([np.concatenate(r.plane_space(nrm[i])) for i in range(m)])
    hin = np.zeros((100, 5, 4, 3), f32)
    hpar = np.zeros((100, 26), f32)
    for i in range(100):
        piv = rng.normal(size=6).astype(f32) * 0.4
        axA = nrm[i]; axB = nrm[i + 100]
        pa, pb = rng.normal(size=3).astype(f32), rng.normal(size=3).astype(f32)
        qa, qb = rq(rng), rq(rng)
        hpar[i] = np.concatenate([piv, axA, axB, pa, qa, pb, qb])
        _, hin[i] = O.hinge_rows("ref", piv[:3], piv[3:], axA, axB, pa, qa, pb, qb)
    np.savez_compressed(os.path.join(HERE, "rows.npz"), J=J, inv_mA=mA, inv_IA=IA, inv_mB=mB, inv_IB=IB, error=err, restitution=rest,
                        vel=vel, prepared=prep, row5=row5, dv=dv, delta=delta, impulse=imp, normal=nrm, plane_space=ps,
                        hinge_params=hpar, hinge_J=hin)
    make_friction(r)
    make_manifold_decisions(r)
    make_contacts(r)
    make_misc(r)
    make_graphs()
    make_broadphase()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


def friction_inputs(rng, m):
    """Random constraint_row_friction cases: tangent Jacobians, prepared rows, warm impulses, normal loads that put about
    half of the cases outside the friction circle (clamped) and a few at zero load (the |impulse| <= epsilon branch)."""
    J = rng.normal(size=(m, 24)).astype(f32)
    fr = np.concatenate([(0.2 + rng.random((m, 2))), rng.normal(size=(m, 2)), rng.normal(size=(m, 2)) * 0.2], axis=1).astype(f32)
    mu = (0.1 + rng.random(m)).astype(f32)
    nimp = np.abs(rng.normal(size=m)).astype(f32) * np.where(rng.random(m) < 0.5, 0.05, 3.0).astype(f32)
    nimp[::17] = 0
    fr[::34, 2:6] = 0                                                   # zero rhs and impulse with zero load
    masses = np.zeros((m, 20), f32)
    masses[:, 0], masses[:, 10] = rng.random(m), rng.random(m)
    for i in range(m):
        masses[i, 1:10] = np.diag(1 + rng.random(3)).reshape(9)
        masses[i, 11:20] = np.diag(1 + rng.random(3)).reshape(9)
    masses[::5, 10:20] = 0                                              # static second body
    dv = (rng.normal(size=(m, 12)) * 0.3).astype(f32)
    dv[::5, 6:12] = 0
    return J, fr, mu, nimp, masses, dv


def make_friction(r):
    """solve_friction + warm_start (constraints/constraint_row_friction.cpp:11-66)."""
    rng = np.random.default_rng(4321)
    m = 600
    J, fr, mu, nimp, masses, dv = friction_inputs(rng, m)
    out_imp, out_dv, warm_dv = np.zeros((m, 2), f32), np.zeros((m, 12), f32), np.zeros((m, 12), f32)
    for i in range(m):
        out_imp[i], out_dv[i] = r.solve_friction(J[i], fr[i], mu[i], nimp[i], masses[i], dv[i])
        _, warm_dv[i] = r.solve_friction(J[i], fr[i], mu[i], nimp[i], masses[i], dv[i], warm=True)
    np.savez_compressed(os.path.join(HERE, "friction.npz"), J=J, fr=fr, mu=mu, normal_impulse=nimp, masses=masses, dv=dv,
                        impulse=out_imp, dv_out=out_dv, dv_warm=warm_dv)


def contact_inputs(rng, m):
    """Random contact points between two bodies a few centimetres apart: prepare inputs (cp15, two body23) and
    solve_position inputs (cp13, two body26).  Every fifth second body is static; a third of the points is separated
    (distance > 0: the error = distance / dt branch, and the early return of solve_position)."""
    def bodies(k):
        pos = (rng.normal(size=(m, 3)) * 0.5).astype(f32)
        orn = np.stack([rq(rng) for _ in range(m)])
        inv_m = (0.2 + rng.random(m)).astype(f32)
        inv_I = np.stack([np.diag(0.5 + rng.random(3)).reshape(9) for _ in range(m)]).astype(f32)
        return pos, orn, inv_m, inv_I
    pa, qa, ma, Ia = bodies(0)
    pb, qb, mb, Ib = bodies(1)
    mb[::5] = 0; Ib[::5] = 0
    lv = lambda: (rng.normal(size=(m, 3)) * 0.5).astype(f32)
    va, wa, vb, wb = lv(), lv(), lv(), lv()
    vb[::5] = 0; wb[::5] = 0
    nrm = np.stack([x / np.linalg.norm(x) for x in rng.normal(size=(m, 3))]).astype(f32)
    pivA, pivB = (rng.normal(size=(m, 3)) * 0.4).astype(f32), (rng.normal(size=(m, 3)) * 0.4).astype(f32)
    dist = (rng.normal(size=m) * 0.02).astype(f32)
    cp15 = np.concatenate([pivA, pivB, nrm, dist[:, None], (0.1 + rng.random((m, 1))), rng.random((m, 1)) * 0.5,
                           np.abs(rng.normal(size=(m, 1))), rng.normal(size=(m, 2)) * 0.2], axis=1).astype(f32)
    bA23 = np.concatenate([pa, qa, va, wa, ma[:, None], Ia], axis=1).astype(f32)
    bB23 = np.concatenate([pb, qb, vb, wb, mb[:, None], Ib], axis=1).astype(f32)
    # position solve: world inertia consistent with the orientation, like update_inertias leaves it
    o = O.ora_fns()
    IWa = np.stack([o.world_inertia(qa[i], Ia[i]).reshape(9) for i in range(m)])
    IWb = np.stack([o.world_inertia(qb[i], Ib[i]).reshape(9) for i in range(m)])
    att = rng.integers(0, 3, size=m).astype(f32)
    local_n = np.stack([x / np.linalg.norm(x) for x in rng.normal(size=(m, 3))]).astype(f32)
    cp13 = np.concatenate([pivA, pivB, nrm, local_n, att[:, None]], axis=1).astype(f32)
    # put pivotB where pivotA is (in world space) shifted along the normal, so that distances are a few centimetres
    bA26 = np.concatenate([pa, qa, ma[:, None], IWa, Ia], axis=1).astype(f32)
    bB26 = np.concatenate([pb, qb, mb[:, None], IWb, Ib], axis=1).astype(f32)
    return cp15, bA23, bB23, cp13, bA26, bB26


def make_contacts(r):
    """contact_constraint::prepare and ::solve_position (constraints/contact_constraint.cpp:15-90) with
    position_solver::solve (dynamics/position_solver.hpp:16-51)."""
    rng = np.random.default_rng(97531)
    m = 500
    cp15, bA23, bB23, cp13, bA26, bB26 = contact_inputs(rng, m)
    prep = [r.contact_prepare(cp15[i], 1.0 / 60, bA23[i], bB23[i]) for i in range(m)]
    pos = [r.contact_solve_position(cp13[i], bA26[i], bB26[i]) for i in range(m)]
    np.savez_compressed(os.path.join(HERE, "contacts.npz"), cp15=cp15, bodyA23=bA23, bodyB23=bB23, cp13=cp13, bodyA26=bA26, bodyB26=bB26,
                        nJ=np.stack([p["nJ"] for p in prep]), n5=np.stack([p["n5"] for p in prep]), fJ=np.stack([p["fJ"] for p in prep]),
                        fr6=np.stack([p["fr6"] for p in prep]), mu=np.array([p["mu"] for p in prep], f32),
                        solved=np.array([p[0] for p in pos], np.int32), outA=np.stack([p[1] for p in pos]), outB=np.stack([p[2] for p in pos]),
                        out5=np.stack([p[3] for p in pos]))


def manifold_decision_inputs(rng, m):
    """Inputs for the three per-point decisions of process_collision, scaled around their thresholds (caching 0.04,
    breaking 0.02) so that both outcomes and ties between result points occur."""
    n = rng.integers(1, 5, size=m)
    cpA, cpB = (rng.normal(size=(m, 3)) * 0.3).astype(f32), (rng.normal(size=(m, 3)) * 0.3).astype(f32)
    resA = (cpA[:, None, :] + rng.normal(size=(m, 4, 3)) * rng.choice([0.01, 0.03, 0.08], size=(m, 4, 1))).astype(f32)
    resB = (cpB[:, None, :] + rng.normal(size=(m, 4, 3)) * rng.choice([0.01, 0.03, 0.08], size=(m, 4, 1))).astype(f32)
    resA[::7, 1] = resA[::7, 0]                                         # exact ties: the first one must win
    origin = (rng.normal(size=(m, 3))).astype(f32)
    orn = np.stack([rq(rng) for _ in range(m)])
    angvel = (rng.normal(size=(m, 3)) * rng.choice([0.01, 1.0, 6.0], size=(m, 1))).astype(f32)
    posB = (origin + rng.normal(size=(m, 3)) * 0.5).astype(f32)
    ornB = np.stack([rq(rng) for _ in range(m)])
    normal = np.stack([x / np.linalg.norm(x) for x in rng.normal(size=(m, 3))]).astype(f32)
    return n, cpA, cpB, resA, resB, origin, orn, angvel, posB, ornB, normal


def make_manifold_decisions(r):
    """find_nearest_contact, find_nearest_contact_rolling, should_remove_point (util/collision_util.cpp:233-280, :397-413)."""
    rng = np.random.default_rng(24680)
    m = 800
    n, cpA, cpB, resA, resB, origin, orn, angvel, posB, ornB, normal = manifold_decision_inputs(rng, m)
    o = O.ora_fns()
    near = np.array([r.find_nearest_contact(cpA[i], cpB[i], resA[i, :n[i]], resB[i, :n[i]]) for i in range(m)], np.uint32)
    roll = np.array([r.find_nearest_contact_rolling(resA[i, :n[i]], cpA[i], origin[i], orn[i], angvel[i], 1.0 / 60) for i in range(m)], np.uint32)
    # should_remove_point: pivotB placed so that the world-space separation is a few centimetres in a random direction
    pAw = np.stack([o.integrate(orn[i], [0, 0, 0], 0.0) for i in range(m)])        # (normalised orientation, unused)
    sep = (rng.normal(size=(m, 3)) * 0.015).astype(f32)
    pivB = np.zeros((m, 3), f32)
    for i in range(m):
        R = lambda q, v: np.asarray(v, np.float64) + 2 * np.cross(q[:3], np.cross(q[:3], v) + q[3] * np.asarray(v, np.float64))
        world_A = origin[i].astype(np.float64) + R(orn[i].astype(np.float64), cpA[i])
        target = world_A - sep[i]
        qc = ornB[i].astype(np.float64) * np.array([-1, -1, -1, 1])
        pivB[i] = R(qc, target - posB[i]).astype(f32)
    rem = np.array([r.should_remove_point(cpA[i], pivB[i], normal[i], origin[i], orn[i], posB[i], ornB[i]) for i in range(m)], np.uint8)
    np.savez_compressed(os.path.join(HERE, "manifold.npz"), n=n.astype(np.uint32), cpA=cpA, cpB=cpB, resA=resA, resB=resB, origin=origin, orn=orn,
                        angvel=angvel, posB=posB, ornB=ornB, normal=normal, pivB=pivB, nearest=near, nearest_rolling=roll, remove=rem)
    del pAw


def misc_inputs(rng, m):
    """Hinge position solve inputs (hinge12 + two body26: bodies a little off their hinge), material pairs, AABB pairs
    that touch exactly on a face (closed-interval test) or miss by an ulp."""
    o = O.ora_fns()
    def body(pos, static=False):
        # a static body keeps the identity orientation: position_solver::solve re-normalises the orientation of BOTH
        # bodies in place (position_solver.hpp:26-32), also of a static one, which the restatement deliberately does not
        # (DESIGN.md section 6); with an exactly unit quaternion the two agree bit for bit
        q = np.array([0, 0, 0, 1], f32) if static else rq(rng)
        inv_m = f32(0) if static else f32(0.2 + rng.random())
        I = np.zeros(9, f32) if static else np.diag(0.5 + rng.random(3)).reshape(9).astype(f32)
        IW = o.world_inertia(q, I).reshape(9)
        return np.concatenate([pos, q, [inv_m], IW, I]).astype(f32)
    hinge, bA, bB = np.zeros((m, 12), f32), np.zeros((m, 26), f32), np.zeros((m, 26), f32)
    for i in range(m):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        hinge[i] = np.concatenate([rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.3, ax, ax + rng.normal(size=3) * 0.05])
        hinge[i, 9:12] /= np.linalg.norm(hinge[i, 9:12])
        pa = rng.normal(size=3)
        bA[i] = body(pa.astype(f32))
        bB[i] = body((pa + rng.normal(size=3) * 0.4).astype(f32), static=(i % 5 == 0))
    mats = rng.random((m, 4)).astype(f32)
    a = np.concatenate([rng.normal(size=(m, 3)), np.zeros((m, 3))], axis=1).astype(f32)
    a[:, 3:] = a[:, :3] + (0.1 + rng.random((m, 3))).astype(f32)
    b = a.copy()
    shift = (rng.normal(size=(m, 3)) * 0.6).astype(f32)
    b[:, :3] += shift; b[:, 3:] += shift
    k = np.arange(m) % 3 == 0                                            # touching exactly: b.min.x == a.max.x
    b[k, 3] = a[k, 3] + (b[k, 3] - b[k, 0]); b[k, 0] = a[k, 3]
    k2 = np.arange(m) % 6 == 0                                           # ... or one ulp beyond
    b[k2, 0] = np.nextafter(b[k2, 0], f32(np.inf))
    return hinge, bA, bB, mats, a, b


def make_misc(r):
    """hinge_constraint::solve_position (hinge_constraint.cpp:180-213), material mixing (material_mixing.hpp:12-18),
    intersect(AABB, AABB) (geom.cpp:762-770)."""
    rng = np.random.default_rng(1357)
    m = 400
    hinge, bA, bB, mats, a, b = misc_inputs(rng, m)
    hs = [r.hinge_solve_position(hinge[i], bA[i], bB[i]) for i in range(m)]
    mix = np.stack([r.material_mix(*mats[i]) for i in range(m)])
    hit = np.array([r.intersect_aabb(a[i], b[i]) for i in range(m)], np.uint8)
    np.savez_compressed(os.path.join(HERE, "misc.npz"), hinge=hinge, bodyA26=bA, bodyB26=bB, hinge_err=np.array([h[0] for h in hs], f32),
                        outA=np.stack([h[1] for h in hs]), outB=np.stack([h[2] for h in hs]), materials=mats, mixed=mix, aabb_a=a, aabb_b=b, hit=hit)


def graph_inputs(rng, count):
    """Random body graphs: 20-120 nodes, a fifth of them non-connecting (static), sparse to dense edge sets, with
    duplicate edges and edges between two static nodes left out (a manifold always has a dynamic body)."""
    out = []
    for _ in range(count):
        n = int(rng.integers(20, 121))
        static = rng.random(n) < 0.2
        ne = int(rng.integers(n // 3, 2 * n))
        e = rng.integers(0, n, size=(ne, 2))
        e = e[(e[:, 0] != e[:, 1]) & ~(static[e[:, 0]] & static[e[:, 1]])]
        e = np.unique(np.sort(e, axis=1), axis=0)
        out.append((static.astype(np.uint8), e.astype(np.uint32)))
    return out


def make_graphs():
    """entity_graph::connected_components (core/entity_graph.cpp) on 60 random graphs, flattened into one file."""
    rng = np.random.default_rng(8642)
    graphs = graph_inputs(rng, 60)
    labels = [O.ref_connected_components(s, e) for s, e in graphs]
    np.savez_compressed(os.path.join(HERE, "graphs.npz"), sizes=np.array([len(s) for s, _ in graphs], np.uint32),
                        edge_counts=np.array([len(e) for _, e in graphs], np.uint32), static=np.concatenate([s for s, _ in graphs]),
                        edges=np.concatenate([e for _, e in graphs]), labels=np.concatenate(labels))


def broadphase_scene(rng, n_dyn, n_static):
    """A crowded box of unit-ish boxes, spheres and capsules (many AABBs within the 0.02 margin of each other), a floor
    plane and a few big static slabs: bodies SoA for OracleWorld.add_bodies."""
    from edyn_b200.rigidbody import DYNAMIC, STATIC, RigidBodyDef, bodies_soa, box_shape, capsule_shape, plane_shape, sphere_shape
    side = max(2.0, (n_dyn ** (1 / 3)) * 0.9)
    defs = []
    for _ in range(n_dyn):
        kind = rng.integers(0, 3)
        shape = (box_shape(tuple((0.2 + 0.3 * rng.random(3)).tolist())) if kind == 0 else
                 sphere_shape(float(0.2 + 0.3 * rng.random())) if kind == 1 else capsule_shape(float(0.15 + 0.1 * rng.random()), float(0.2 + 0.2 * rng.random())))
        q = rq(rng) if kind != 1 else np.array([0, 0, 0, 1], f32)
        defs.append(RigidBodyDef(kind=DYNAMIC, position=tuple((rng.random(3) * side).tolist()), orientation=tuple(q.tolist()), mass=1.0, shape=shape))
    defs.append(RigidBodyDef(kind=STATIC, shape=plane_shape((0, 1, 0), 0.0)))
    for _ in range(n_static):
        defs.append(RigidBodyDef(kind=STATIC, position=tuple((rng.random(3) * side).tolist()), shape=box_shape((float(side), 0.3, 0.3))))
    return bodies_soa(defs, (0.0, 0.0, 0.0))


def make_broadphase():
    """Ordered broadphase pair lists produced with the reference's real dynamic trees (ref_broadphase_pairs) for 12
    crowded scenes; the AABBs the trees were built from are stored alongside (they come from the pinned shape_aabb)."""
    rng = np.random.default_rng(112233)
    sizes, procs, boxes0, boxes1, pairs, counts, scenes = [], [], [], [], [], [], []
    for k in range(12):
        soa = broadphase_scene(rng, int(rng.integers(40, 260)), int(rng.integers(0, 4)))
        o = O.OracleWorld(); o.add_bodies(soa)
        bb1 = o.state()["aabb"]
        bb0 = (bb1 + np.tile((rng.normal(size=(len(bb1), 3)) * 0.3).astype(f32), 2)).astype(f32)       # where the leaves were created
        proc = (soa["kind"] == 0).astype(np.uint8)
        p = O.ref_broadphase_pairs(bb0, bb1, proc)
        sizes.append(len(proc)); procs.append(proc); boxes0.append(bb0); boxes1.append(bb1); pairs.append(p); counts.append(len(p))
        scenes.append({k2: v for k2, v in soa.items() if v is not None})
    np.savez_compressed(os.path.join(HERE, "broadphase.npz"), sizes=np.array(sizes, np.uint32), counts=np.array(counts, np.uint32),
                        procedural=np.concatenate(procs), aabb0=np.concatenate(boxes0), aabb1=np.concatenate(boxes1), pairs=np.concatenate(pairs),
                        **{"soa_" + k2: np.concatenate([s[k2] for s in scenes]) for k2 in scenes[0]})


if __name__ == "__main__":
    main()

// TEST INFRASTRUCTURE -- oracle/_ref: thin extern "C" wrappers over the REAL reference functions,
// compiled in place from /root/reference by oracle/Makefile (no reference source is copied into
// this repo).  Signatures match the ora_* pure functions of ora_api.cpp one to one, so tests can
// diff the restatement against the reference bit for bit.
#include <edyn/collision/collide.hpp>
#include <edyn/collision/collision_result.hpp>
#include <edyn/constraints/constraint_row.hpp>
#include <edyn/constraints/constraint_row_options.hpp>
#include <edyn/constraints/hinge_constraint.hpp>
#include <edyn/constraints/constraint_body.hpp>
#include <edyn/dynamics/moment_of_inertia.hpp>
#include <edyn/dynamics/row_cache.hpp>
#include <edyn/dynamics/position_solver.hpp>
#include <edyn/constraints/constraint_row_friction.hpp>
#include <edyn/constraints/contact_constraint.hpp>
#include <edyn/comp/position.hpp>
#include <edyn/comp/orientation.hpp>
#include <edyn/comp/inertia.hpp>
#include <edyn/collision/contact_point.hpp>
#include <edyn/util/collision_util.hpp>
#include <edyn/dynamics/material_mixing.hpp>
#include <edyn/comp/aabb.hpp>
#include <edyn/core/entity_graph.hpp>
#include <edyn/collision/dynamic_tree.hpp>
#include <edyn/config/constants.hpp>
#include <vector>
#include <edyn/math/geom.hpp>
#include <edyn/math/quaternion.hpp>
#include <edyn/math/matrix3x3.hpp>
#include <edyn/util/aabb_util.hpp>
#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))

using namespace edyn;

namespace {
vector3 v3(const float *p) { return {p[0], p[1], p[2]}; }
quaternion q4(const float *p) { return {p[0], p[1], p[2], p[3]}; }
void put3(float *p, const vector3 &v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
void put4(float *p, const quaternion &q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
matrix3x3 m9(const float *p) { return {vector3{p[0], p[1], p[2]}, vector3{p[3], p[4], p[5]}, vector3{p[6], p[7], p[8]}}; }
void put9(float *p, const matrix3x3 &m) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) p[i * 3 + j] = m[i][j]; }

// kinds follow shapes.hpp:23-37: sphere 0, capsule 2, box 3, plane 6
template<typename F> void with_shape(uint32_t kind, const float *p, F f) {
    switch (kind) {
    case 0: f(sphere_shape{p[0]}); break;
    case 2: { capsule_shape c; c.radius = p[0]; c.half_length = p[1]; c.axis = static_cast<coordinate_axis>(int(p[2])); f(c); break; }
    case 3: f(box_shape{vector3{p[0], p[1], p[2]}}); break;
    case 6: f(plane_shape{vector3{p[0], p[1], p[2]}, p[3]}); break;
    default: break;
    }
}
template<typename S> AABB aabb_of(const S &s, const vector3 &pos, const quaternion &orn) { return shape_aabb(s, pos, orn); }
}

REF_API int ref_collide(uint32_t kindA, const float *pA, uint32_t kindB, const float *pB, const float *posA, const float *ornA,
                        const float *posB, const float *ornB, float *out10, uint32_t *att) {
    collision_result result;
    with_shape(kindA, pA, [&](auto &&shA) {
        with_shape(kindB, pB, [&](auto &&shB) {
            auto aabbA = aabb_of(shA, v3(posA), q4(ornA));
            auto aabbB = aabb_of(shB, v3(posB), q4(ornB));
            auto ctx = collision_context{v3(posA), q4(ornA), aabbA, v3(posB), q4(ornB), aabbB, collision_threshold};
            collide(shA, shB, ctx, result);
        });
    });
    for (size_t i = 0; i < result.num_points; ++i) {
        auto &pt = result.point[i];
        put3(out10 + 10 * i, pt.pivotA); put3(out10 + 10 * i + 3, pt.pivotB); put3(out10 + 10 * i + 6, pt.normal);
        out10[10 * i + 9] = pt.distance;
        att[i] = static_cast<uint32_t>(pt.normal_attachment);
    }
    return int(result.num_points);
}

REF_API void ref_shape_aabb(uint32_t kind, const float *p, const float *pos, const float *orn, float *out6) {
    with_shape(kind, p, [&](auto &&sh) {
        auto bb = aabb_of(sh, v3(pos), q4(orn));
        put3(out6, bb.min); put3(out6 + 3, bb.max);
    });
}

REF_API void ref_integrate(const float *q, const float *w, float dt, float *out4) { put4(out4, integrate(q4(q), v3(w), dt)); }

REF_API void ref_plane_space(const float *n, float *p, float *q) { vector3 a, b; plane_space(v3(n), a, b); put3(p, a); put3(q, b); }

REF_API int ref_intersect_line_aabb(const float *p0, const float *p1, const float *mn, const float *mx, float *s) {
    return int(intersect_line_aabb(vector2{p0[0], p0[1]}, vector2{p1[0], p1[1]}, vector2{mn[0], mn[1]}, vector2{mx[0], mx[1]}, s[0], s[1]));
}

REF_API int ref_closest_segment_segment(const float *p1, const float *q1, const float *p2, const float *q2, float *out16, float *dist) {
    scalar s, t, sp = 0, tp = 0; vector3 c1, c2, c1p{0, 0, 0}, c2p{0, 0, 0}; size_t np = 0;
    *dist = closest_point_segment_segment(v3(p1), v3(q1), v3(p2), v3(q2), s, t, c1, c2, &np, &sp, &tp, &c1p, &c2p);
    out16[0] = s; out16[1] = t; put3(out16 + 2, c1); put3(out16 + 5, c2);
    out16[8] = sp; out16[9] = tp; put3(out16 + 10, c1p); put3(out16 + 13, c2p);
    return int(np);
}

REF_API int ref_maybe_add_points(uint32_t n, const float *pivotA, const float *pivotB, float *outA, float *outB) {
    collision_result r;
    for (uint32_t i = 0; i < n; ++i) {
        collision_result::collision_point p{};
        p.pivotA = v3(pivotA + 3 * i); p.pivotB = v3(pivotB + 3 * i); p.normal = vector3_y;
        p.distance = 0; p.normal_attachment = contact_normal_attachment::none;
        r.maybe_add_point(p);
    }
    for (size_t i = 0; i < r.num_points; ++i) { put3(outA + 3 * i, r.point[i].pivotA); put3(outB + 3 * i, r.point[i].pivotB); }
    return int(r.num_points);
}

REF_API void ref_moment_of_inertia(uint32_t kind, const float *p, float mass, float *out9) {
    with_shape(kind, p, [&](auto &&sh) { put9(out9, moment_of_inertia(sh, mass)); });
}
REF_API void ref_inverse_symmetric(const float *m, float *out9) { put9(out9, inverse_matrix_symmetric(m9(m))); }
REF_API void ref_world_inertia(const float *orn, const float *inv_I, float *out9) {
    auto basis = to_matrix3x3(q4(orn));
    put9(out9, basis * m9(inv_I) * transpose(basis));
}

REF_API void ref_prepare_row(const float *J, float inv_mA, const float *inv_IA, float inv_mB, const float *inv_IB, float error,
                             float erp, float restitution, const float *vels12, float *out2) {
    constraint_row row{};
    for (int i = 0; i < 4; ++i) row.J[i] = v3(J + 3 * i);
    row.inv_mA = inv_mA; row.inv_IA = m9(inv_IA); row.inv_mB = inv_mB; row.inv_IB = m9(inv_IB);
    constraint_row_options opt{};
    opt.error = error; opt.erp = erp; opt.restitution = restitution;
    prepare_row(row, opt, v3(vels12), v3(vels12 + 3), v3(vels12 + 6), v3(vels12 + 9));
    out2[0] = row.eff_mass; out2[1] = row.rhs;
}

REF_API float ref_solve_row(const float *J, float *row5, const float *dv12) {
    constraint_row row{};
    for (int i = 0; i < 4; ++i) row.J[i] = v3(J + 3 * i);
    row.eff_mass = row5[0]; row.rhs = row5[1]; row.lower_limit = row5[2]; row.upper_limit = row5[3]; row.impulse = row5[4];
    delta_linvel dvA{v3(dv12)}, dvB{v3(dv12 + 6)};
    delta_angvel dwA{v3(dv12 + 3)}, dwB{v3(dv12 + 9)};
    row.dvA = &dvA; row.dwA = &dwA; row.dvB = &dvB; row.dwB = &dwB;
    float d = solve(row);
    row5[4] = row.impulse;
    return d;
}

// solve_friction / warm_start(constraint_row_friction&), constraint_row_friction.cpp:11-66.  Same argument layout as
// ora_solve_friction: the friction pair points at a one-element row cache holding the normal row (impulse, masses, deltas).
REF_API void ref_solve_friction(const float *J24, float *fr6, float mu, float normal_impulse, const float *masses20, float *dv12, int warm) {
    std::vector<constraint_row> cache(1);
    constraint_row &nr = cache[0];
    nr = constraint_row{};
    nr.impulse = normal_impulse;
    nr.inv_mA = masses20[0]; nr.inv_IA = m9(masses20 + 1); nr.inv_mB = masses20[10]; nr.inv_IB = m9(masses20 + 11);
    delta_linvel dvA{v3(dv12)}, dvB{v3(dv12 + 6)};
    delta_angvel dwA{v3(dv12 + 3)}, dwB{v3(dv12 + 9)};
    nr.dvA = &dvA; nr.dwA = &dwA; nr.dvB = &dvB; nr.dwB = &dwB;
    constraint_row_friction fr{};
    for (int i = 0; i < 2; ++i) {
        for (int k = 0; k < 4; ++k) fr.row[i].J[k] = v3(J24 + 12 * i + 3 * k);
        fr.row[i].eff_mass = fr6[i]; fr.row[i].rhs = fr6[2 + i]; fr.row[i].impulse = fr6[4 + i];
    }
    fr.friction_coefficient = mu; fr.normal_row_index = 0;
    if (warm) warm_start(fr, cache); else solve_friction(fr, cache);
    fr6[4] = fr.row[0].impulse; fr6[5] = fr.row[1].impulse;
    put3(dv12, dvA); put3(dv12 + 3, dwA); put3(dv12 + 6, dvB); put3(dv12 + 9, dwB);
}

// contact_constraint::prepare (contact_constraint.cpp:15-56).  cp14 = pivotA pivotB normal distance friction restitution
// applied_normal_impulse applied_friction_impulse[2]; body23 = pos(3) orn(4) linvel(3) angvel(3) inv_m inv_I(9), origin == pos.
// out: nJ12, n5 = lower upper impulse options.error options.restitution, fJ24, fr6 = eff_mass[2] rhs[2] impulse[2], mu.
static constraint_body glue_body(const float *b) {
    constraint_body cb{};
    cb.origin = v3(b); cb.pos = v3(b); cb.orn = q4(b + 3); cb.linvel = v3(b + 7); cb.angvel = v3(b + 10); cb.inv_m = b[13]; cb.inv_I = m9(b + 14);
    return cb;
}
REF_API void ref_contact_prepare(const float *cp14, float dt, const float *bodyA23, const float *bodyB23,
                                 float *nJ12, float *n5, float *fJ24, float *fr6, float *mu) {
    contact_constraint con{};
    con.pivotA = v3(cp14); con.pivotB = v3(cp14 + 3); con.normal = v3(cp14 + 6); con.distance = cp14[9];
    con.friction = cp14[10]; con.restitution = cp14[11]; con.applied_normal_impulse = cp14[12];
    con.applied_friction_impulse = {cp14[13], cp14[14]};
    constraint_row_prep_cache cache;
    cache.add_constraint();
    con.prepare(cache, dt, glue_body(bodyA23), glue_body(bodyB23));
    const auto &e = cache.rows[0];
    for (int k = 0; k < 4; ++k) put3(nJ12 + 3 * k, e.row.J[k]);
    n5[0] = e.row.lower_limit; n5[1] = e.row.upper_limit; n5[2] = e.row.impulse; n5[3] = e.options.error; n5[4] = e.options.restitution;
    for (int i = 0; i < 2; ++i) {
        for (int k = 0; k < 4; ++k) put3(fJ24 + 12 * i + 3 * k, e.friction.row[i].J[k]);
        fr6[i] = e.friction.row[i].eff_mass; fr6[2 + i] = e.friction.row[i].rhs; fr6[4 + i] = e.friction.row[i].impulse;
    }
    *mu = e.friction.friction_coefficient;
}

// contact_constraint::solve_position (contact_constraint.cpp:58-90) through position_solver::solve
// (dynamics/position_solver.hpp:16-51).  cp13 = pivotA pivotB normal local_normal attachment(as float 0/1/2);
// body26 = pos(3) orn(4) inv_m inv_IW(9) inv_I_local(9), updated in place; out5 = normal(3) distance max_error.
REF_API int ref_contact_solve_position(const float *cp13, float *bodyA26, float *bodyB26, float *out5) {
    contact_constraint con{};
    con.pivotA = v3(cp13); con.pivotB = v3(cp13 + 3); con.normal = v3(cp13 + 6); con.local_normal = v3(cp13 + 9);
    con.normal_attachment = cp13[12] == 1.0f ? contact_normal_attachment::normal_on_A
                          : cp13[12] == 2.0f ? contact_normal_attachment::normal_on_B : contact_normal_attachment::none;
    position posA{v3(bodyA26)}, posB{v3(bodyB26)};
    orientation ornA{q4(bodyA26 + 3)}, ornB{q4(bodyB26 + 3)};
    inertia_world_inv iwA{m9(bodyA26 + 8)}, iwB{m9(bodyB26 + 8)};
    inertia_inv ilA{m9(bodyA26 + 17)}, ilB{m9(bodyB26 + 17)};
    position_solver solver{};
    solver.originA = nullptr; solver.originB = nullptr; solver.comA = vector3_zero; solver.comB = vector3_zero;
    solver.posA = &posA; solver.posB = &posB; solver.ornA = &ornA; solver.ornB = &ornB;
    solver.inv_mA = bodyA26[7]; solver.inv_mB = bodyB26[7];
    solver.inv_IA = &iwA; solver.inv_IB = &iwB; solver.inv_IA_local = &ilA; solver.inv_IB_local = &ilB;
    const scalar before = solver.max_error;
    con.solve_position(solver);
    put3(bodyA26, posA); put4(bodyA26 + 3, ornA); put9(bodyA26 + 8, iwA);
    put3(bodyB26, posB); put4(bodyB26 + 3, ornB); put9(bodyB26 + 8, iwB);
    put3(out5, con.normal); out5[3] = con.distance; out5[4] = solver.max_error;
    (void)before;
    return con.distance > -EDYN_EPSILON ? 0 : 1;
}

// find_nearest_contact / find_nearest_contact_rolling / should_remove_point (util/collision_util.cpp:233-280, :397-413;
// cut out of that unit at build time, see oracle/Makefile).
REF_API uint32_t ref_find_nearest_contact(const float *cpA, const float *cpB, uint32_t n, const float *resA, const float *resB) {
    contact_point cp{}; cp.pivotA = v3(cpA); cp.pivotB = v3(cpB);
    collision_result res{}; res.num_points = n;
    for (uint32_t i = 0; i < n; ++i) { res.point[i].pivotA = v3(resA + 3 * i); res.point[i].pivotB = v3(resB + 3 * i); }
    return uint32_t(find_nearest_contact(cp, res));
}
REF_API uint32_t ref_find_nearest_contact_rolling(uint32_t n, const float *resA, const float *cp_pivot, const float *origin, const float *orn,
                                                  const float *angvel, float dt) {
    collision_result res{}; res.num_points = n;
    for (uint32_t i = 0; i < n; ++i) res.point[i].pivotA = v3(resA + 3 * i);
    return uint32_t(find_nearest_contact_rolling(res, v3(cp_pivot), v3(origin), q4(orn), v3(angvel), dt));
}
REF_API int ref_should_remove_point(const float *pivotA, const float *pivotB, const float *normal, const float *posA, const float *ornA,
                                    const float *posB, const float *ornB) {
    contact_point cp{}; cp.pivotA = v3(pivotA); cp.pivotB = v3(pivotB); cp.normal = v3(normal);
    return should_remove_point(cp, v3(posA), q4(ornA), v3(posB), q4(ornB)) ? 1 : 0;
}

// hinge_constraint::solve_position (hinge_constraint.cpp:180-213) through position_solver::solve; layouts of
// ora_hinge_solve_position.
REF_API float ref_hinge_solve_position(const float *hinge12, float *bodyA26, float *bodyB26) {
    hinge_constraint con{};
    con.pivot[0] = v3(hinge12); con.pivot[1] = v3(hinge12 + 3);
    con.set_axes(v3(hinge12 + 6), v3(hinge12 + 9));
    position posA{v3(bodyA26)}, posB{v3(bodyB26)};
    orientation ornA{q4(bodyA26 + 3)}, ornB{q4(bodyB26 + 3)};
    inertia_world_inv iwA{m9(bodyA26 + 8)}, iwB{m9(bodyB26 + 8)};
    inertia_inv ilA{m9(bodyA26 + 17)}, ilB{m9(bodyB26 + 17)};
    position_solver solver{};
    solver.originA = nullptr; solver.originB = nullptr; solver.comA = vector3_zero; solver.comB = vector3_zero;
    solver.posA = &posA; solver.posB = &posB; solver.ornA = &ornA; solver.ornB = &ornB;
    solver.inv_mA = bodyA26[7]; solver.inv_mB = bodyB26[7];
    solver.inv_IA = &iwA; solver.inv_IB = &iwB; solver.inv_IA_local = &ilA; solver.inv_IB_local = &ilB;
    con.solve_position(solver);
    put3(bodyA26, posA); put4(bodyA26 + 3, ornA); put9(bodyA26 + 8, iwA);
    put3(bodyB26, posB); put4(bodyB26 + 3, ornB); put9(bodyB26 + 8, iwB);
    return solver.max_error;
}

// material mixing (dynamics/material_mixing.hpp:12-18) and the closed-interval AABB test (math/geom.cpp:762-770)
REF_API void ref_material_mix(float frictionA, float frictionB, float restitutionA, float restitutionB, float *out2) {
    out2[0] = material_mix_friction(frictionA, frictionB); out2[1] = material_mix_restitution(restitutionA, restitutionB);
}
REF_API int ref_intersect_aabb(const float *a6, const float *b6) {
    AABB a{v3(a6), v3(a6 + 3)}, b{v3(b6), v3(b6 + 3)};
    return intersect(a, b) ? 1 : 0;
}

// entity_graph::connected_components (core/entity_graph.cpp): the island partition the reference's island manager
// maintains incrementally.  Nodes 0..n-1 (non_connecting[i] != 0 = static / kinematic body), edges as node index pairs.
// out_label[i] = smallest connecting node of i's component, 0xFFFFFFFF for non-connecting nodes.
REF_API uint32_t ref_connected_components(uint32_t n, const uint8_t *non_connecting, uint32_t ne, const uint32_t *edges, uint32_t *out_label) {
    entity_graph graph;
    std::vector<entity_graph::index_type> node(n);
    for (uint32_t i = 0; i < n; ++i) node[i] = graph.insert_node(entt::entity{i}, non_connecting[i] != 0);
    for (uint32_t e = 0; e < ne; ++e) graph.insert_edge(entt::entity{n + e}, node[edges[2 * e]], node[edges[2 * e + 1]]);
    for (uint32_t i = 0; i < n; ++i) out_label[i] = 0xFFFFFFFFu;
    auto comps = graph.connected_components();
    uint32_t count = 0;
    for (auto &c : comps) {
        uint32_t lab = 0xFFFFFFFFu;
        for (auto ent : c.nodes) { uint32_t i = static_cast<uint32_t>(ent); if (!non_connecting[i] && i < lab) lab = i; }
        if (lab == 0xFFFFFFFFu) continue;
        ++count;
        for (auto ent : c.nodes) { uint32_t i = static_cast<uint32_t>(ent); if (!non_connecting[i]) out_label[i] = lab; }
    }
    return count;
}

// The pair search of broadphase::update (broadphase.cpp:177-195) around the reference's REAL dynamic AABB trees
// (collision/dynamic_tree.cpp: fat leaves, SAH insertion, rotations).  broadphase.cpp itself needs a registry, so the
// ten lines of collide_tree (:136-155) are restated here: query the procedural tree, then the non-procedural one, with
// the body's AABB inset by -0.02; for every leaf hit that is not the body itself, not paired yet and whose EXACT AABB
// intersects the query, make the pair (body, other).  Bodies are created at aabb0 and moved to aabb1 (dynamic_tree::move,
// what move_aabbs does) before the queries; procedural bodies are visited in descending index order.
REF_API uint32_t ref_broadphase_pairs(uint32_t n, const float *aabb0, const float *aabb1, const uint8_t *procedural, uint32_t cap, uint32_t *out_pairs) {
    dynamic_tree tree, np_tree;
    std::vector<tree_node_id_t> id(n);
    auto box = [](const float *p) { return AABB{v3(p), v3(p + 3)}; };
    for (uint32_t i = 0; i < n; ++i) id[i] = (procedural[i] ? tree : np_tree).create(box(aabb0 + 6 * i), entt::entity{i});
    for (uint32_t i = 0; i < n; ++i) (procedural[i] ? tree : np_tree).move(id[i], box(aabb1 + 6 * i));      // move_aabbs, broadphase.cpp:99-117
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    auto paired = [&](uint32_t a, uint32_t b) {
        for (auto &p : pairs) if ((p.first == a && p.second == b) || (p.first == b && p.second == a)) return true;
        return false;
    };
    const vector3 offset = vector3_one * -contact_breaking_threshold;           // m_aabb_offset, broadphase.hpp:15
    for (uint32_t i = n; i-- > 0;) {
        if (!procedural[i]) continue;
        const AABB query = box(aabb1 + 6 * i).inset(offset);
        auto visit = [&](const dynamic_tree &t) {
            t.query(query, [&](tree_node_id_t nid) {
                const uint32_t j = static_cast<uint32_t>(t.get_node(nid).entity);
                if (j == i || paired(i, j)) return;
                if (intersect(query, box(aabb1 + 6 * j))) pairs.emplace_back(i, j);
            });
        };
        visit(tree); visit(np_tree);
    }
    uint32_t k = 0;
    for (auto &p : pairs) { if (k < cap) { out_pairs[2 * k] = p.first; out_pairs[2 * k + 1] = p.second; } ++k; }
    return k;
}

// hinge_constraint::prepare (hinge_constraint.cpp:26-69): returns the 5 Jacobians (60 floats) built for
// pivot/axes given in body space.  bodies: pos(3) orn(4) per body.
REF_API int ref_hinge_rows(const float *pivotA, const float *pivotB, const float *axisA, const float *axisB,
                           const float *posA, const float *ornA, const float *posB, const float *ornB, float *J60) {
    hinge_constraint con{};
    con.pivot[0] = v3(pivotA); con.pivot[1] = v3(pivotB);
    con.set_axes(v3(axisA), v3(axisB));
    constraint_row_prep_cache cache{};
    cache.add_constraint();
    constraint_body bA{v3(posA), v3(posA), q4(ornA), vector3_zero, vector3_zero, 1, matrix3x3_identity};
    constraint_body bB{v3(posB), v3(posB), q4(ornB), vector3_zero, vector3_zero, 1, matrix3x3_identity};
    con.prepare(cache, scalar(1.0 / 60), bA, bB);
    int n = int(cache.num_rows);
    for (int i = 0; i < n && i < 5; ++i) for (int k = 0; k < 4; ++k) put3(J60 + i * 12 + k * 3, cache.rows[i].row.J[k]);
    return n;
}

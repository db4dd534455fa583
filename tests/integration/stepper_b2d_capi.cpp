// TEST INFRASTRUCTURE.  C interface over a REAL Edyn registry stepped by edyn::stepper_b2d
// (edyn_b200/csrc/host/stepper_b2d.hpp) instead of the reference's stepper_sequential.  Scene construction is the
// reference's own public API, shared with oracle/ref_stepper.cpp (refs_create / refs_add_bodies / refs_add_hinges /
// refs_add_exclusions: edyn::attach, make_rigidbody, make_constraint<hinge_constraint>, exclude_collision; refs_get_state
// reads the registry's components back), so what is tested is exactly what a user of the reference would run after
// swapping the stepper.  Linked once against tests/integration/b2d_mock.cpp (CPU suite) and once against
// edyn_b200/libb2d.so (-m gpu).
#include "../../oracle/ref_stepper.cpp"
#include "stepper_b2d.hpp"
#include <edyn/util/exclude_collision.hpp>
#include <map>
#include <memory>

namespace {
struct contact_observer {                      // what a user of the reference connects to see contacts come and go
    uint32_t started = 0, ended = 0;
    void on_started(entt::registry &, entt::entity) { ++started; }
    void on_ended(entt::registry &, entt::entity) { ++ended; }
};
std::map<void *, std::unique_ptr<contact_observer>> g_observers;
std::map<void *, std::unique_ptr<edyn::stepper_b2d>> g_steppers;
edyn::stepper_b2d &stepper(void *h) { return *g_steppers.at(h); }
}

// attaches the device stepper to a world made by refs_create(); bodies added before or after are both picked up
REFS_API int eb2d_attach(void *h, int device, uint32_t max_bodies, uint32_t max_manifolds, uint32_t max_hinges, int sleeping) {
    auto *w = static_cast<World *>(h);
    try {
        edyn::b2d_capacities cap;
        cap.device = device; cap.max_bodies = max_bodies; cap.max_manifolds = max_manifolds; cap.max_hinges = max_hinges;
        cap.sleeping = sleeping != 0;
        g_steppers[h] = std::make_unique<edyn::stepper_b2d>(w->registry, w->time, cap);
    } catch (const std::exception &e) { std::fprintf(stderr, "eb2d_attach: %s\n", e.what()); return -1; }
    return 0;
}
REFS_API void eb2d_detach(void *h) {
    if (auto it = g_observers.find(h); it != g_observers.end()) {
        auto &r = static_cast<World *>(h)->registry;
        r.on_construct<edyn::contact_started_tag>().disconnect<&contact_observer::on_started>(*it->second);
        r.on_destroy<edyn::contact_point>().disconnect<&contact_observer::on_ended>(*it->second);
        g_observers.erase(it);
    }
    g_steppers.erase(h);
}
// out4 = manifolds, points, points created, points destroyed by this call; observed2 = running totals of
// on_construct<contact_started_tag> and on_destroy<contact_point> as a user's listeners see them
REFS_API int eb2d_mirror_contacts(void *h, uint32_t *out4, uint32_t *observed2) {
    auto *w = static_cast<World *>(h);
    if (!g_observers.count(h)) {
        auto obs = std::make_unique<contact_observer>();
        w->registry.on_construct<edyn::contact_started_tag>().connect<&contact_observer::on_started>(*obs);
        w->registry.on_destroy<edyn::contact_point>().connect<&contact_observer::on_ended>(*obs);
        g_observers[h] = std::move(obs);
    }
    try {
        const auto st = stepper(h).mirror_contacts();
        out4[0] = st.manifolds; out4[1] = st.points; out4[2] = st.points_created; out4[3] = st.points_destroyed;
    } catch (const std::exception &e) { std::fprintf(stderr, "eb2d_mirror_contacts: %s\n", e.what()); return -1; }
    observed2[0] = g_observers[h]->started; observed2[1] = g_observers[h]->ended;
    return 0;
}
// the contact point entities in the order refs_get_contacts reports the points (manifold by manifold, list order)
REFS_API uint32_t eb2d_contact_entities(void *h, uint32_t capacity, uint32_t *entities, uint32_t *lifetime) {
    auto &r = static_cast<World *>(h)->registry;
    uint32_t k = 0;
    for (auto [me, manifold, state] : r.view<edyn::contact_manifold, edyn::contact_manifold_state>().each()) {
        for (auto e = state.contact_entity; e != entt::null && k < capacity; e = r.get<edyn::contact_point_list>(e).next) {
            entities[k] = entt::to_integral(e); lifetime[k] = r.get<edyn::contact_point>(e).lifetime; ++k;
        }
    }
    return k;
}
REFS_API int eb2d_step(void *h, uint32_t n) {
    auto *w = static_cast<World *>(h);
    try {
        for (uint32_t i = 0; i < n; ++i) { w->time = double(w->steps++) * w->dt; stepper(h).step_simulation(w->time); }
    } catch (const std::exception &e) { std::fprintf(stderr, "eb2d_step: %s\n", e.what()); return -1; }
    return 0;
}
// edyn::update semantics: as many fixed steps as fit in `time - last time` (stepper_sequential.cpp:28-69)
REFS_API int eb2d_update(void *h, double time) {
    try { stepper(h).update(time); } catch (const std::exception &e) { std::fprintf(stderr, "eb2d_update: %s\n", e.what()); return -1; }
    return 0;
}
REFS_API int eb2d_wake_up(void *h, uint32_t body) {
    try { stepper(h).wake_up(static_cast<World *>(h)->bodies[body]); } catch (const std::exception &e) { std::fprintf(stderr, "eb2d_wake_up: %s\n", e.what()); return -1; }
    return 0;
}
REFS_API uint32_t eb2d_body_id(void *h, uint32_t body) { return stepper(h).body_id(static_cast<World *>(h)->bodies[body]); }
REFS_API void *eb2d_world(void *h) { return stepper(h).world(); }
// what user code does between steps
REFS_API void eb2d_patch_velocity(void *h, uint32_t body, const float *lv, const float *av) {
    auto *w = static_cast<World *>(h);
    w->registry.patch<edyn::linvel>(w->bodies[body], [&](edyn::linvel &v) { v.x = lv[0]; v.y = lv[1]; v.z = lv[2]; });
    w->registry.patch<edyn::angvel>(w->bodies[body], [&](edyn::angvel &v) { v.x = av[0]; v.y = av[1]; v.z = av[2]; });
}
REFS_API void eb2d_patch_position(void *h, uint32_t body, const float *p) {
    auto *w = static_cast<World *>(h);
    w->registry.patch<edyn::position>(w->bodies[body], [&](edyn::position &x) { x.x = p[0]; x.y = p[1]; x.z = p[2]; });
}
REFS_API void eb2d_destroy_body(void *h, uint32_t body) {
    auto *w = static_cast<World *>(h);
    w->registry.destroy(w->bodies[body]);
}
REFS_API void eb2d_remove_exclusion(void *h, uint32_t a, uint32_t b) {
    auto *w = static_cast<World *>(h);
    edyn::remove_collision_exclusion(w->registry, w->bodies[a], w->bodies[b]);
}
// refs_get_state for registries in which some of the bodies were destroyed: their slots read as zeros
REFS_API void eb2d_get_state(void *h, float *pos, float *orn, float *lv, float *av) {
    auto *w = static_cast<World *>(h);
    auto &r = w->registry;
    for (size_t i = 0; i < w->bodies.size(); ++i) {
        const auto e = w->bodies[i];
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = lv[3 * i + k] = av[3 * i + k] = 0;
        for (int k = 0; k < 4; ++k) orn[4 * i + k] = 0;
        if (!r.valid(e)) continue;
        const auto &p = r.get<edyn::position>(e); const auto &q = r.get<edyn::orientation>(e);
        pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
        orn[4 * i] = q.x; orn[4 * i + 1] = q.y; orn[4 * i + 2] = q.z; orn[4 * i + 3] = q.w;
        if (auto *v = r.try_get<edyn::linvel>(e)) { lv[3 * i] = v->x; lv[3 * i + 1] = v->y; lv[3 * i + 2] = v->z; }
        if (auto *v = r.try_get<edyn::angvel>(e)) { av[3 * i] = v->x; av[3 * i + 1] = v->y; av[3 * i + 2] = v->z; }
    }
}

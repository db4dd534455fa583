// TEST INFRASTRUCTURE.  Stand-in for libb2d.so on machines without a GPU: the part of include/b2d.h that
// edyn_b200/csrc/host/stepper_b2d.hpp calls, answered by the CPU oracle (oracle/liboracle.so).  It exists so that the
// host-side staging logic of the EnTT binding -- entity <-> body id maps, SoA layout, dirty subsets, exclusions, removal --
// can be exercised by the CPU test-suite (tests/test_stepper_b2d.py).  Never shipped, never loaded by edyn_b200.
#include "b2d.h"
#include <cstdint>
#include <string>
#include <vector>

extern "C" {
void *ora_create(float dt, int vel_iters, int pos_iters, int threads);
void ora_destroy(void *h);
int ora_add_bodies(void *h, uint32_t n, const float *pos, const float *orn, const float *linvel, const float *angvel, const float *inv_mass,
                   const float *inv_inertia, const float *gravity, const uint32_t *kind, const uint32_t *shape_kind, const float *shape_params,
                   const float *friction, const float *restitution, const uint64_t *group, const uint64_t *mask);
int ora_add_hinges(void *h, uint32_t n, const uint32_t *a, const uint32_t *b, const float *pivotA, const float *pivotB, const float *axisA, const float *axisB);
void ora_remove_bodies(void *h, uint32_t n, const uint32_t *ids);
void ora_add_exclusions(void *h, uint32_t n, const uint32_t *a, const uint32_t *b);
void ora_remove_exclusions(void *h, uint32_t n, const uint32_t *a, const uint32_t *b);
void ora_step(void *h, int n);
uint32_t ora_num_bodies(void *h);
void ora_get_state(void *h, float *pos, float *orn, float *linvel, float *angvel, float *aabb6, float *inv_IW9);
void ora_set_state(void *h, const float *pos, const float *orn, const float *linvel, const float *angvel);
uint32_t ora_num_manifolds(void *h);
void ora_set_sleeping(void *h, int enabled);
void ora_set_restitution_iterations(void *h, int iters, int individual);
void ora_wake_bodies(void *h, uint32_t n, const uint32_t *ids);
void ora_get_sleeping(void *h, uint32_t *asleep);
void ora_get_pairs(void *h, uint32_t *pairs);
void ora_get_contacts(void *h, uint32_t *num, float *pt18, uint32_t *pt_u2);
}

struct b2d_world { void *ora; std::string err; uint32_t patched = 0; };
static std::string g_create_error;

extern "C" {
b2d_world *b2d_create(const b2d_config *c) {
    if (!c || c->max_bodies == 0) { g_create_error = "mock: bad configuration"; return nullptr; }
    auto *w = new b2d_world();
    w->ora = ora_create(c->fixed_dt, int(c->velocity_iterations), int(c->position_iterations), 1);
    if (c->flags & B2D_FLAG_SLEEPING) ora_set_sleeping(w->ora, 1);
    if (c->flags & B2D_FLAG_RESTITUTION_SOLVER) ora_set_restitution_iterations(w->ora, 8, 3);
    return w;
}
void b2d_destroy(b2d_world *w) { if (w) { ora_destroy(w->ora); delete w; } }
const char *b2d_last_error(const b2d_world *w) { return w ? w->err.c_str() : g_create_error.c_str(); }
int b2d_add_bodies(b2d_world *w, const b2d_bodies *b, uint32_t *first_id) {
    for (uint32_t i = 0; i < b->count; ++i) {
        const uint32_t k = b->shape_kind[i];
        if (k != B2D_SHAPE_SPHERE && k != B2D_SHAPE_CAPSULE && k != B2D_SHAPE_BOX && k != B2D_SHAPE_PLANE && k != B2D_SHAPE_NONE) { w->err = "unsupported shape"; return B2D_ERR_UNSUPPORTED; }
    }
    const int first = ora_add_bodies(w->ora, b->count, b->pos, b->orn, b->linvel, b->angvel, b->inv_mass, b->inv_inertia, b->gravity, b->kind,
                                     b->shape_kind, b->shape_params, b->friction, b->restitution, b->group, b->mask);
    if (first_id) *first_id = uint32_t(first);
    return B2D_OK;
}
int b2d_add_hinges(b2d_world *w, uint32_t n, const uint32_t *a, const uint32_t *b, const float *pa, const float *pb, const float *xa, const float *xb) {
    ora_add_hinges(w->ora, n, a, b, pa, pb, xa, xb);
    return B2D_OK;
}
int b2d_remove_bodies(b2d_world *w, const uint32_t *ids, uint32_t n) { ora_remove_bodies(w->ora, n, ids); return B2D_OK; }
int b2d_add_exclusions(b2d_world *w, uint32_t n, const uint32_t *a, const uint32_t *b) { ora_add_exclusions(w->ora, n, a, b); return B2D_OK; }
int b2d_remove_exclusions(b2d_world *w, uint32_t n, const uint32_t *a, const uint32_t *b) { ora_remove_exclusions(w->ora, n, a, b); return B2D_OK; }
// mock limitation: transform and velocities only (the oracle has no per-body mass / material setter)
int b2d_upload_bodies(b2d_world *w, uint32_t n, const uint32_t *ids, const b2d_body_patch *p) {
    const uint32_t nb = ora_num_bodies(w->ora);
    std::vector<float> pos(3 * nb), orn(4 * nb), lv(3 * nb), av(3 * nb);
    ora_get_state(w->ora, pos.data(), orn.data(), lv.data(), av.data(), nullptr, nullptr);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t id = ids[i];
        if (id >= nb) { w->err = "body id out of range"; return B2D_ERR_ARGUMENT; }
        for (int k = 0; k < 3; ++k) {
            if (p->pos) pos[3 * id + k] = p->pos[3 * i + k];
            if (p->linvel) lv[3 * id + k] = p->linvel[3 * i + k];
            if (p->angvel) av[3 * id + k] = p->angvel[3 * i + k];
        }
        if (p->orn) for (int k = 0; k < 4; ++k) orn[4 * id + k] = p->orn[4 * i + k];
    }
    ora_set_state(w->ora, pos.data(), orn.data(), lv.data(), av.data());
    w->patched += n;
    return B2D_OK;
}
int b2d_upload_state(b2d_world *w, const float *pos, const float *orn, const float *lv, const float *av) { ora_set_state(w->ora, pos, orn, lv, av); return B2D_OK; }
int b2d_step(b2d_world *w, uint32_t n) { ora_step(w->ora, int(n)); return B2D_OK; }
int b2d_download_state(b2d_world *w, float *pos, float *orn, float *lv, float *av, float *aabb, float *iw) { ora_get_state(w->ora, pos, orn, lv, av, aabb, iw); return B2D_OK; }
int b2d_num_manifolds(b2d_world *w, uint32_t *n) { *n = ora_num_manifolds(w->ora); return B2D_OK; }
int b2d_download_contacts(b2d_world *w, uint32_t capacity, uint32_t *pairs, uint32_t *num, float *pt18, uint32_t *pt_u2, uint32_t *n) {
    const uint32_t m = ora_num_manifolds(w->ora);
    if (m > capacity) { w->err = "capacity"; return B2D_ERR_CAPACITY; }
    ora_get_pairs(w->ora, pairs);
    ora_get_contacts(w->ora, num, pt18, pt_u2);
    *n = m;
    return B2D_OK;
}
int b2d_download_sleeping(b2d_world *w, uint32_t *asleep) { ora_get_sleeping(w->ora, asleep); return B2D_OK; }
int b2d_wake_bodies(b2d_world *w, const uint32_t *ids, uint32_t n) {
    if (!ids) { w->err = "mock: wake-all is not supported"; return B2D_ERR_UNSUPPORTED; }
    ora_wake_bodies(w->ora, n, ids);
    return B2D_OK;
}
int b2d_sync(b2d_world *) { return B2D_OK; }
// test hook (not in b2d.h): how many bodies travelled through b2d_upload_bodies so far
__attribute__((visibility("default"))) uint32_t b2d_mock_patched(b2d_world *w) { return w->patched; }
}

// The reference's examples/hello_world/hello_world.cpp:17-39 restated over the adapter (a box dropped on a plane,
// sequential stepping).  Build: g++ -std=c++17 hello_world.cpp -L../.. -lb2d -Wl,-rpath,'$ORIGIN/../..' -o hello_world
#include "edyn_adapter.hpp"
#include <cstdio>

int main() {
    edyn::registry registry;
    edyn::init_config cfg;
    cfg.max_bodies = 16; cfg.max_manifolds = 64;
    edyn::attach(registry, cfg);

    auto plane = edyn::rigidbody_def{};
    plane.kind = edyn::rigidbody_kind::rb_static;
    plane.shape = edyn::plane_shape{{0, 1, 0}, 0};
    edyn::make_rigidbody(registry, plane);

    auto def = edyn::rigidbody_def{};
    def.position = {0, 3, 0};
    def.mass = 10;
    def.shape = edyn::box_shape{{0.5f, 0.5f, 0.5f}};
    def.material->friction = 0.8f;
    auto box = edyn::make_rigidbody(registry, def);

    for (int i = 0; i < 200; ++i) {
        edyn::update(registry, 1.0 / 60);
        if (i % 20 == 0) { auto p = registry.get_position(box); std::printf("pos (%.3f, %.3f, %.3f)\n", p.x, p.y, p.z); }
    }
    auto p = registry.get_position(box);
    edyn::detach(registry);
    return (p.y > 0.49f && p.y < 0.51f) ? 0 : 1;
}

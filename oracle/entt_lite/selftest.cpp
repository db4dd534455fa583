// TEST INFRASTRUCTURE.  Behaviour checks of oracle/entt_lite against EnTT's documented semantics -- the ones the
// reference's results depend on (iteration order decides the Gauss-Seidel row order).  tests/test_ref_stepper.py
// compiles and runs this file; exit code 0 and "entt_lite ok" = pass.
#include <entt/entity/registry.hpp>
#include <entt/signal/sigh.hpp>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); std::exit(1); } } while (0)

struct pos { float x; };
struct vel { float v; };
struct tag {};
struct big { double d[100]; };
struct listener {
    int n = 0, short_calls = 0;
    void on(entt::registry &, entt::entity) { ++n; }
    void on_short(entt::registry &) { ++short_calls; }           // fewer parameters than the signal: trailing ones are dropped
};
static int free_calls = 0;
static void free_listener(entt::registry &, entt::entity) { ++free_calls; }

int main() {
    entt::registry r;
    listener L;
    r.on_construct<pos>().connect<&listener::on>(L);
    r.on_construct<pos>().connect<&listener::on_short>(L);
    r.on_construct<vel>().connect<&free_listener>();
    entt::scoped_connection c = r.on_destroy<pos>().connect<&listener::on>(L);
    std::vector<entt::entity> es;
    for (int i = 0; i < 5; ++i) {
        auto e = r.create(); es.push_back(e);
        r.emplace<pos>(e, float(i));
        if (i % 2 == 0) r.emplace<vel>(e, float(10 * i));
        if (i != 3) r.emplace<tag>(e);
    }
    CHECK(L.n == 5 && L.short_calls == 5 && free_calls == 3);
    // a single-component view iterates its pool newest first
    int k = 4;
    for (auto e : r.view<pos>()) { CHECK(r.get<pos>(e).x == float(k)); --k; }
    // a multi-component view is led by its smallest pool (vel: 0, 2, 4), newest first; each() skips the empty type
    std::vector<float> got;
    for (auto [e, p, v] : r.view<pos, vel, tag>().each()) { got.push_back(p.x); CHECK(v.v == 10 * p.x); }
    CHECK((got == std::vector<float>{4, 2, 0}));
    r.view<pos, tag>(entt::exclude<vel>).each([&](pos &p) { CHECK(p.x == 1); });
    r.view<pos, tag>().each([&](entt::entity e, pos &p) { CHECK(r.valid(e)); (void)p; });
    // swap and pop: destroying entity 1 moves the last element (4) into its slot: pool = [0, 4, 2, 3] -> iteration 3, 2, 4, 0
    r.destroy(es[1]);
    CHECK(L.n == 6 && !r.valid(es[1]));
    got.clear();
    for (auto e : r.view<pos>()) got.push_back(r.get<pos>(e).x);
    CHECK((got == std::vector<float>{3, 2, 4, 0}));
    // identifiers are recycled with a new version
    auto e2 = r.create();
    CHECK(entt::to_entity(e2) == entt::to_entity(es[1]) && e2 != es[1]);
    CHECK(r.try_get<vel>(es[3]) == nullptr && r.try_get<vel>(es[2]) != nullptr);
    r.patch<pos>(es[0], [](pos &p) { p.x = 42; });
    CHECK(r.get<pos>(es[0]).x == 42);
    CHECK((r.all_of<pos, tag>(es[0]) && !r.any_of<vel>(es[3])));
    r.ctx().emplace<int>(7);
    CHECK(r.ctx().get<int>() == 7 && r.ctx().find<float>() == nullptr);
    // a free-standing sparse set: newest first, paged sparse array (a high index touches one page, not idx entries)
    entt::sparse_set s;
    s.push(es[0]); s.push(es[2]);
    CHECK(s.contains(es[2]) && *s.begin() == es[2]);
    s.erase(es[0]);
    CHECK(s.size() == 1 && !s.contains(es[0]));
    const entt::entity far = entt::entity{900000u};
    s.push(far);
    CHECK(s.contains(far) && !s.contains(entt::entity{899999u}) && s.index(far) == 1);
    s.clear();
    CHECK(s.empty() && !s.contains(far));
    CHECK(entt::entity{entt::null} == entt::null);
    // references to components survive appends (paged storage)
    std::vector<entt::entity> many;
    auto first = r.create();
    big *addr = &r.emplace<big>(first);
    addr->d[0] = 3.5;
    for (int i = 0; i < 5000; ++i) { auto e = r.create(); many.push_back(e); r.emplace<big>(e).d[0] = i; }
    CHECK(addr == &r.get<big>(first) && addr->d[0] == 3.5 && r.get<big>(many[4999]).d[0] == 4999);
    r.destroy(many.begin(), many.end());
    CHECK(r.view<big>().size() == 1);
    // disconnect by (function, instance)
    r.on_construct<pos>().disconnect<&listener::on>(L);
    r.emplace<pos>(r.create(), 1.f);
    CHECK(L.n == 6 && L.short_calls == 6);
    r.clear<vel>();
    CHECK(r.view<vel>().size() == 0);
    std::puts("entt_lite ok");
    return 0;
}

// Declaration-level stand-in for EnTT 3.15 (see entt/entity/fwd.hpp in this shim).
#pragma once
#include "fwd.hpp"
#include "../signal/fwd.hpp"
#include <cstddef>
#include <tuple>
#include <utility>
#include <algorithm>
#include <vector>
#include <memory>
namespace entt {
template<typename T> class storage_stub;
template<typename Get, typename Exclude> class basic_view;
template<typename... G, typename... E>
class basic_view<get_t<G...>, exclude_t<E...>> {
public:
    using iterator = const entity *;
    iterator begin() const; iterator end() const;
    template<typename... T> decltype(auto) get(entity) const;
    bool contains(entity) const;
    std::size_t size_hint() const; std::size_t size() const;
    template<typename F> void each(F) const;
    struct iterable { struct it { }; };
    explicit operator bool() const;
    entity front() const;
};
struct ctx_stub {
    template<typename T, typename... A> T & emplace(A&&...);
    template<typename T> T & get(); template<typename T> const T & get() const;
    template<typename T> T * find(); template<typename T> const T * find() const;
    template<typename T> bool contains() const; template<typename T> bool erase();
};
class registry {
public:
    template<typename T> using storage_for_type = storage_stub<T>;
    template<typename... T, typename... E> basic_view<get_t<storage_stub<T>...>, exclude_t<storage_stub<E>...>> view(exclude_t<E...> = exclude_t<>{});
    template<typename... T, typename... E> basic_view<get_t<storage_stub<T>...>, exclude_t<storage_stub<E>...>> view(exclude_t<E...> = exclude_t<>{}) const;
    template<typename... T> decltype(auto) get(entity);
    template<typename... T> decltype(auto) get(entity) const;
    template<typename... T> auto try_get(entity);
    template<typename... T> auto try_get(entity) const;
    template<typename... T> bool all_of(entity) const;
    template<typename... T> bool any_of(entity) const;
    template<typename T, typename... A> decltype(auto) emplace(entity, A&&...);
    template<typename T, typename... A> decltype(auto) emplace_or_replace(entity, A&&...);
    template<typename T, typename... A> decltype(auto) replace(entity, A&&...);
    template<typename T, typename... F> decltype(auto) patch(entity, F&&...);
    template<typename... T> std::size_t remove(entity);
    template<typename... T> void erase(entity);
    template<typename... T> void clear();
    entity create(); void destroy(entity); bool valid(entity) const;
    template<typename It> void destroy(It, It);
    ctx_stub & ctx(); const ctx_stub & ctx() const;
    template<typename T> storage_stub<T> & storage();
    template<typename T> sink<sigh<void(registry&, entity)>> on_construct();
    template<typename T> sink<sigh<void(registry&, entity)>> on_destroy();
    template<typename T> sink<sigh<void(registry&, entity)>> on_update();
};
}

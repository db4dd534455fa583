// Declaration-level stand-in for EnTT 3.15 (absent from this image, reference conanfile.py:77).
// Written from scratch for oracle/_ref: just enough names for the reference's EnTT-free
// translation units (collide_*.cpp, geom.cpp, constraint_row*.cpp, ...) to parse.  Nothing here is
// ever instantiated or called by those units; it is not an ECS.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <vector>
namespace entt {
enum class entity : std::uint32_t {};
struct null_t {
    constexpr operator entity() const noexcept { return entity{0xFFFFFFFFu}; }
    constexpr bool operator==(null_t) const noexcept { return true; }
    constexpr bool operator!=(null_t) const noexcept { return false; }
    constexpr bool operator==(entity e) const noexcept { return static_cast<std::uint32_t>(e) == 0xFFFFFFFFu; }
    constexpr bool operator!=(entity e) const noexcept { return !(*this == e); }
};
constexpr bool operator==(entity e, null_t n) noexcept { return n == e; }
constexpr bool operator!=(entity e, null_t n) noexcept { return n != e; }
inline constexpr null_t null{};
class registry;
template<typename...> struct type_list {};
template<typename... T> struct exclude_t : type_list<T...> {};
template<typename... T> struct get_t : type_list<T...> {};
template<typename... T> inline constexpr exclude_t<T...> exclude{};
class sparse_set {
public:
    using iterator = std::vector<entity>::const_reverse_iterator;
    iterator begin() const; iterator end() const;
    bool contains(entity) const; std::size_t size() const; bool empty() const;
    iterator push(entity); template<typename It> iterator push(It, It);
    void erase(entity); bool remove(entity); void clear();
    entity operator[](std::size_t) const; iterator find(entity) const;
    void swap(sparse_set &); void reserve(std::size_t);
};
}

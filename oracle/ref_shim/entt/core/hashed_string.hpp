// Declaration-level stand-in for EnTT 3.15 (see entt/entity/fwd.hpp in this shim).
#pragma once
#include "fwd.hpp"
#include <cstddef>
namespace entt {
// FNV-1a 32-bit, the published algorithm entt::hashed_string uses.
class hashed_string {
    id_type h; const char *s;
public:
    static constexpr id_type value(const char *str) noexcept {
        id_type v = 2166136261u;
        while(*str) { v = (v ^ static_cast<id_type>(*str++)) * 16777619u; }
        return v;
    }
    constexpr hashed_string(const char *str) noexcept : h{value(str)}, s{str} {}
    constexpr id_type value() const noexcept { return h; }
    constexpr operator id_type() const noexcept { return h; }
    constexpr const char *data() const noexcept { return s; }
};
namespace literals { constexpr hashed_string operator"" _hs(const char *str, std::size_t) noexcept { return hashed_string{str}; } }
}

// Declaration-level stand-in for EnTT 3.15 (see entt/entity/fwd.hpp in this shim).
#pragma once
#include "fwd.hpp"
namespace entt {
template<typename R, typename... A> class sigh<R(A...)> { public: void publish(A...) const; };
template<typename R, typename... A> class sink<sigh<R(A...)>> { public:
    sink(sigh<R(A...)> &);
    template<auto C, typename... T> void connect(T&&...);
    template<auto C, typename... T> void disconnect(T&&...);
    template<typename T> void disconnect(T&&);
};
template<typename R, typename... A> sink(sigh<R(A...)> &) -> sink<sigh<R(A...)>>;
struct scoped_connection { void release(); };
}

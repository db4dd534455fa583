// Declaration-level stand-in for EnTT 3.15 (see entt/entity/fwd.hpp in this shim).
#pragma once
namespace entt {
template<typename> class delegate;
template<typename> class sigh;
template<typename> class sink;
class dispatcher;
struct scoped_connection;
}

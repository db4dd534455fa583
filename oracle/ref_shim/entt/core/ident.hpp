// Declaration-level stand-in for EnTT 3.15 (see entt/entity/fwd.hpp in this shim).
#pragma once
#include "fwd.hpp"

#pragma once
#include "../entt_lite.hpp"

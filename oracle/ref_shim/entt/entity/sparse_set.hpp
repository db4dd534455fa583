#pragma once
#include "fwd.hpp"

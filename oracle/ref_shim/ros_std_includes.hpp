// Test-infrastructure shim (NOT product code): the standard headers that the real
// rosidl-generated message headers pull in transitively and that the reference
// sources silently rely on (e.g. std::abs in State.h, std::unordered_map in HybridAstar.cpp).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

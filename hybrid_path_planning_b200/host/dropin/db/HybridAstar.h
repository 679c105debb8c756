// Forwarding header: copy over dubins_hybrid_star/include/HybridAstar.h; the hpp-b200 host
// directory must be on the include path.
#pragma once
#define HPP_SHELL_DEFAULT_DUBINS 1
#include "HybridAstarShell.h"

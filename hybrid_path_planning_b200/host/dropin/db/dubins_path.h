// Forwarding header: copy over dubins_hybrid_star/include/dubins_path.h; the hpp-b200 host
// directory must be on the include path.
#pragma once
#include "PathShells.h"

// Forwarding header: copy over reeds_shepp_hybrid_astar/include/reeds_shepp_path.h; the
// hpp-b200 host directory must be on the include path.
#pragma once
#include "PathShells.h"

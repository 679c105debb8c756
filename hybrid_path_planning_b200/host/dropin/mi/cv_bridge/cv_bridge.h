// Stand-in for cv_bridge/cv_bridge.h in the map_inflate drop-in: map_inflate_node.cpp includes it and uses nothing of it.
#pragma once

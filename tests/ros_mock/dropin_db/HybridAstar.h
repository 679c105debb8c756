#pragma once
#define HPP_SHELL_DEFAULT_DUBINS 1
#include "HybridAstarShell.h"

#pragma once
#include "HybridAstarShell.h"

#pragma once
#include "GridMapShell.h"

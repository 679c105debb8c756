#pragma once
#include "CarDataShell.h"

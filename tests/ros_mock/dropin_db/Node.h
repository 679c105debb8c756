#pragma once
#include "hpp_types.h"

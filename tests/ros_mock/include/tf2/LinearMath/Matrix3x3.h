#pragma once
#include "tf2/LinearMath/Quaternion.h"

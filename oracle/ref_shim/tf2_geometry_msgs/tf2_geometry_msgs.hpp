// Shim: included but unused on the reference hot path.
#pragma once

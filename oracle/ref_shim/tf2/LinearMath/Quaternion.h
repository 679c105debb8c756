// Shim: tf2 Quaternion is included but unused on the reference hot path.
#pragma once
namespace tf2 { class Quaternion {}; }

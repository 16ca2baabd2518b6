// The one header a maintainer swaps: the node's `#include "monocular_pose_estimator_lib/pose_estimator.h"` now
// reaches the class shim over the C ABI instead of the CPU library.
#pragma once
#include "monocular_pose_estimator_b200/shim.h"

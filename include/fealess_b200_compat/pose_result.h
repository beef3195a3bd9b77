// Forwarding header: lets `#include "pose_result.h"` (reference: ICP/pose_result.h) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_POSE_RESULT_H
#define FEALESS_B200_COMPAT_POSE_RESULT_H
#include "../fealess_b200/icp.hpp"
#endif

// Forwarding header: lets `#include "detection.h"` (reference: ICP/detection.h) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_DETECTION_H
#define FEALESS_B200_COMPAT_DETECTION_H
#include "../fealess_b200/icp.hpp"
#endif

// Forwarding header: lets `#include "NMS.h"` (reference: ICP/NMS.h) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_NMS_H
#define FEALESS_B200_COMPAT_NMS_H
#include "../fealess_b200/icp.hpp"
#endif

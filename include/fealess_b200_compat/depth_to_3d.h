// Forwarding header: lets `#include "depth_to_3d.h"` (reference: ICP/depth_to_3d.h) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_DEPTH_TO_3D_H
#define FEALESS_B200_COMPAT_DEPTH_TO_3D_H
#include "../fealess_b200/icp.hpp"
#endif

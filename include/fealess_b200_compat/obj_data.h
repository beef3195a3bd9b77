// Forwarding header: lets `#include "obj_data.h"` (reference: ICP/obj_data.h) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_OBJ_DATA_H
#define FEALESS_B200_COMPAT_OBJ_DATA_H
#include "../fealess_b200/icp.hpp"
#endif

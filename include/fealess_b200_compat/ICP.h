// Forwarding header: lets `#include "ICP.h"` (reference: ICP/ICP.h) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_ICP_H
#define FEALESS_B200_COMPAT_ICP_H
#include "../fealess_b200/icp.hpp"
#endif

// Forwarding header: lets `#include "linemod.hpp"` (reference: linemod/linemod.hpp) resolve to the fealess_b200 mirror.
#ifndef FEALESS_B200_COMPAT_LINEMOD_HPP
#define FEALESS_B200_COMPAT_LINEMOD_HPP
#include "../fealess_b200/linemod.hpp"
#endif

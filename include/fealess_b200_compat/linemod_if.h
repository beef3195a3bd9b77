// Forwarding header for the reference's linemod/linemod_if.h:15-23.  Detector::match and the accessors come from the
// fealess_b200 mirror; readLinemod / writeLinemod are the mirror's own (fealess_b200/linemod_io.hpp: same signatures, same
// file layout, no cv::FileStorage needed).  drawResponse (viewer) is out of scope.
#ifndef FEALESS_B200_COMPAT_LINEMOD_IF_H
#define FEALESS_B200_COMPAT_LINEMOD_IF_H
#include <string>
#include "../fealess_b200/linemod.hpp"
#include "../fealess_b200/linemod_io.hpp"
#endif

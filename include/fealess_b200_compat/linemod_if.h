// Forwarding header for the reference's linemod/linemod_if.h:15-23.  Detector::match and the accessors come from the
// fealess_b200 mirror; readLinemod / writeLinemod (cv::FileStorage YAML, out of scope here) stay the reference's own
// host functions and are only declared, exactly as in the reference header.
#ifndef FEALESS_B200_COMPAT_LINEMOD_IF_H
#define FEALESS_B200_COMPAT_LINEMOD_IF_H
#include <string>
#include "../fealess_b200/linemod.hpp"
cv::Ptr<cup_linemod::Detector> readLinemod(const std::string& filename);
void writeLinemod(const cv::Ptr<cup_linemod::Detector>& detector, const std::string& filename);
#endif

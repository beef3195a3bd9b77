// TEST INFRASTRUCTURE: forwards <opencv2/calib3d/calib3d.hpp> to the stand-in (see ../cvshim.hpp; the image has no OpenCV C++ SDK).
#include "../../cvshim.hpp"

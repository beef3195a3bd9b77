// TEST INFRASTRUCTURE: forwards <opencv2/core.hpp> to the stand-in (see ../cvshim.hpp; the image has no OpenCV C++ SDK).
#include "../cvshim.hpp"

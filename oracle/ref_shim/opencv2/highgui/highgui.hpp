// TEST INFRASTRUCTURE: forwards <opencv2/highgui/highgui.hpp> to the stand-in (see ../cvshim.hpp; the image has no OpenCV C++ SDK).
#include "../../cvshim.hpp"

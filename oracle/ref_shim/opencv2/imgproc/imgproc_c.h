// TEST INFRASTRUCTURE: forwards <opencv2/imgproc/imgproc_c.h> to the stand-in (see ../cvshim.hpp; the image has no OpenCV C++ SDK).
#include "../../cvshim.hpp"

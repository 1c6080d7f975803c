// oracle/cvshim/opencv2/core.hpp -- test infrastructure: the stand-in for <opencv2/core.hpp> is the same header as
// <opencv2/opencv.hpp> (this image has no OpenCV C++).
#pragma once
#include "opencv.hpp"

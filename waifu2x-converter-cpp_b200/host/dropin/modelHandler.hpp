// dropin/modelHandler.hpp -- INTEGRATION.md option A: put this directory in front of the reference's src/ on the include
// path and its UNMODIFIED sources (src/main.cpp, src/test.cpp) pick up the GPU implementation of w2xc::Model,
// w2xc::modelUtility (reference src/modelHandler.hpp:24-113) instead of the CPU one; src/modelHandler.cpp and
// src/convertRoutine.cpp are then simply not compiled.  Needs OpenCV (cv::Mat is the reference's plane type).
#ifndef W2X_DROPIN_MODEL_HANDLER_HPP_
#define W2X_DROPIN_MODEL_HANDLER_HPP_
#ifndef W2X_WITH_OPENCV
#define W2X_WITH_OPENCV
#endif
#include "../w2xc.hpp"
#endif

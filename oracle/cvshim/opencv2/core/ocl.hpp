// oracle/cvshim/opencv2/core/ocl.hpp -- TEST INFRASTRUCTURE ONLY (see ../opencv.hpp): cv::ocl::setUseOpenCL is a no-op.
#ifndef W2X_ORACLE_CVSHIM_OCL_HPP_
#define W2X_ORACLE_CVSHIM_OCL_HPP_
namespace cv { namespace ocl { inline void setUseOpenCL(bool) {} } }
#endif

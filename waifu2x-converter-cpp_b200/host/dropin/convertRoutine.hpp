// dropin/convertRoutine.hpp -- INTEGRATION.md option A: stands in for the reference's src/convertRoutine.hpp
// (w2xc::convertWithModels, :25-28); see dropin/modelHandler.hpp.
#ifndef W2X_DROPIN_CONVERT_ROUTINE_HPP_
#define W2X_DROPIN_CONVERT_ROUTINE_HPP_
#include "modelHandler.hpp"
#endif

// host simulation shim (tests/ only): the constants of CUDA's math_constants.h the kernels use
#pragma once
#include <limits>
#define CUDART_INF_F std::numeric_limits<float>::infinity()
#define CUDART_NAN_F std::numeric_limits<float>::quiet_NaN()

// densecrf_cpu.h -- drop-in mirror of Thirdparty/DenseCRF/include/densecrf_cpu.h: DenseCRFCPU<M>,
// the image-oriented twin used by examples/example_cpu.cpp:80.  (The name says CPU because the
// reference's does; the arithmetic runs on the B200.)  See densecrf_base.h.
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

#include "densecrf_base.h"

namespace DenseCRF {
#define LCCRF_VARIANT_NAME DenseCRFCPU
#include "densecrf_variant.inl"
#undef LCCRF_VARIANT_NAME
}  // namespace DenseCRF

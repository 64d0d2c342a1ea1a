// densecrf3d.h -- drop-in mirror of Thirdparty/DenseCRF/include/densecrf3d.h: DenseCRF3D<M>,
// the variant src/Tracking.cc:1920 instantiates.  See densecrf_base.h.
#ifndef DENSECRF3D_H
#define DENSECRF3D_H

#include <cmath>
#include <cstring>
#include <vector>

#include "densecrf_base.h"

namespace DenseCRF {
#define LCCRF_VARIANT_NAME DenseCRF3D
#define LCCRF_VARIANT_HAS_POKE
#include "densecrf_variant.inl"
#undef LCCRF_VARIANT_HAS_POKE
#undef LCCRF_VARIANT_NAME
}  // namespace DenseCRF

#endif

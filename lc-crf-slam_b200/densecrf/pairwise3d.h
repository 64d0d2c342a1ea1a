// pairwise3d.h -- drop-in mirror of Thirdparty/DenseCRF/include/pairwise3d.h: PottsPotential3D<M,F>
// with the appearanceKernel / smoothKernel factories called at src/Tracking.cc:1923,1926.
#pragma once

#include <opencv2/core/core.hpp>
#include <vector>

#include "densecrf_base.h"

using namespace std;
using namespace cv;

namespace DenseCRF {

template <int M, int F>
class PottsPotential3D : public PairwisePotential {
#define LCCRF_POTTS_NAME PottsPotential3D
#include "potts_variant.inl"
#undef LCCRF_POTTS_NAME

    // appearance kernel: f = (observation count / posdev1, reprojection error / posdev2)   pairwise3d.h:38-48
    template <class T = float>
    static PottsPotential3D<M, F> *appearanceKernel(int N, float weight, vector<float> &vobserv, vector<float> &verror,
                                                    float posdev1, float posdev2) {
        vector<float> feat((size_t)N * F);
        for (int idx = 0; idx < N; ++idx) {
            feat[(size_t)idx * F + 0] = vobserv[idx] / posdev1;
            feat[(size_t)idx * F + 1] = verror[idx] / posdev2;
        }
        return new PottsPotential3D<M, F>(feat.data(), N, weight);
    }

    // smoothness kernel: f = keypoint (x, y) / posdev2; the 3-D branch is disabled in the reference
    // (pairwise3d.h:56-61), so points3d and posdev1 are accepted and ignored here too   pairwise3d.h:52-71
    template <class T = float>
    static PottsPotential3D<M, F> *smoothKernel(int N, float weight, vector<Point3f> &points3d, vector<Point2f> &points2d,
                                                float posdev1, float posdev2) {
        (void)points3d;
        (void)posdev1;
        vector<float> feat((size_t)N * F);
        for (int idx = 0; idx < N; ++idx) {
            Point2f point2d = points2d[idx];
            feat[(size_t)idx * F + 0] = point2d.x / posdev2;
            feat[(size_t)idx * F + 1] = point2d.y / posdev2;
        }
        return new PottsPotential3D<M, F>(feat.data(), N, weight);
    }
};

}  // namespace DenseCRF

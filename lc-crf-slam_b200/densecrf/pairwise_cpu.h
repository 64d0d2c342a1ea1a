// pairwise_cpu.h -- drop-in mirror of Thirdparty/DenseCRF/include/pairwise_cpu.h: PottsPotentialCPU<M,F>
// with the FromImage factory (Gaussian F=2 / bilateral F=5) used by examples/example_cpu.cpp:86-94.
#pragma once

#include <vector>

#include "densecrf_base.h"

namespace DenseCRF {

template <int M, int F>
class PottsPotentialCPU : public PairwisePotential {
#define LCCRF_POTTS_NAME PottsPotentialCPU
#include "potts_variant.inl"
#undef LCCRF_POTTS_NAME

    // Image potential: f = (x/posdev, y/posdev, features/featuredev...); features == nullptr -> Gaussian only
    //   pairwise_cpu.h:34-50
    template <class T = float>
    static PottsPotentialCPU<M, F> *FromImage(int w, int h, float weight, float posdev, const T *features = nullptr,
                                              float featuredev = 0.0) {
        std::vector<float> feat((size_t)w * h * F);
        for (int hi = 0; hi < h; ++hi)
            for (int wi = 0; wi < w; ++wi) {
                const size_t idx = (size_t)hi * w + wi;
                feat[idx * F + 0] = (float)wi / posdev;
                feat[idx * F + 1] = (float)hi / posdev;
                for (int i = 2; i < F; ++i) feat[idx * F + i] = (float)features[idx * (F - 2) + (i - 2)] / featuredev;
            }
        return new PottsPotentialCPU<M, F>(feat.data(), w * h, weight);
    }
};

}  // namespace DenseCRF

// densecrf_base.h -- drop-in mirror of Thirdparty/DenseCRF/include/densecrf_base.h (LC-CRF-SLAM).
// Same namespace, class names, method signatures and ownership rules; the arithmetic runs on a
// B200 through the C ABI of include/lccrf.h (liblccrf.so).  Host code stays C++.
//
//   PairwisePotential            densecrf_base.h:12-19   plugin interface, unchanged
//   DenseCRF                     densecrf_base.h:22-92   mean-field driver
//
// Built-in potentials (PottsPotential3D / PottsPotentialCPU) attach themselves to the device CRF
// in addPairwiseEnergy(); when every potential is built-in the whole inference() runs on the GPU
// and only the marginals / MAP come back.  A user-defined PairwisePotential keeps working: the
// driver then walks the potentials on host arrays exactly like the reference (stepInit, apply...,
// expAndNormalize), with the built-in pieces still computed by the GPU.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lccrf.h"

namespace DenseCRF {

enum Device { CPU, GPU };

namespace lccrf_detail {
inline void check(int rc, const char *what) {
    if (rc != LCCRF_OK) {  // the reference API is all-void: there is nobody to return an error to
        std::fprintf(stderr, "lccrf: %s failed (%d): %s\n", what, rc, lccrf_last_error());
        std::abort();
    }
}
// one context per host thread, created on first use on device $LCCRF_DEVICE (default 0)
inline lccrf_ctx *context() {
    struct Holder {
        lccrf_ctx *ctx = nullptr;
        ~Holder() { if (ctx) lccrf_ctx_destroy(ctx); }
    };
    static thread_local Holder h;
    if (!h.ctx) {
        const char *dev = std::getenv("LCCRF_DEVICE");
        check(lccrf_ctx_create(dev ? std::atoi(dev) : 0, &h.ctx), "lccrf_ctx_create");
    }
    return h.ctx;
}
}  // namespace lccrf_detail

// Plugin interface of the mean-field driver (densecrf_base.h:12-19): a potential adds its message onto `out_values`.
class PairwisePotential {
protected:
    int N_;
public:
    PairwisePotential(int N) : N_(N) {}
    virtual ~PairwisePotential() = default;
    virtual void apply(float *out_values, const float *in_values, float *tmp) const = 0;
    // lccrf extension point: a potential that can live on the device registers itself with the
    // CRF here and returns true.  Foreign potentials inherit this default and stay on the host.
    virtual bool lccrfAttach(lccrf_crf * /*crf*/, int /*M*/) { return false; }
};

class DenseCRF {
protected:
    // point count; the label count M_ is fixed by the subclass template (lccrfCreate)
    int N_;
    // Host views, materialised lazily: only the plugin path and the getters touch them
    float *unary_, *current_, *next_, *tmp_;
    short *map_;
    // potentials handed over by addPairwiseEnergy; deleted by the destructor (densecrf_base.h:41-45)
    std::vector<PairwisePotential *> pairwise_;

    lccrf_crf *crf_ = nullptr;  // device-side state
    int M_ = 0;
    bool all_on_device_ = true;

    virtual void expAndNormalize(float *out, const float *in, float scale = 1.0, float relax = 1.0) = 0;
    virtual void buildMap() = 0;
    virtual void stepInit() = 0;

    void lccrfCreate(int M) {
        M_ = M;
        lccrf_detail::check(lccrf_crf_create(lccrf_detail::context(), N_, M, &crf_), "lccrf_crf_create");
    }
    void lccrfDestroy() {
        if (crf_) lccrf_crf_destroy(crf_);
        crf_ = nullptr;
    }

public:
    DenseCRF(int N) : N_(N), unary_(nullptr), current_(nullptr), next_(nullptr), tmp_(nullptr), map_(nullptr) {}
    virtual ~DenseCRF() {
        for (auto *p : pairwise_) delete p;  // ownership was transferred by addPairwiseEnergy
    }

    template <int M>
    static DenseCRF *Create(int N);  // declared for source compatibility; the reference's has no body either

    // takes ownership (densecrf_base.h:53-54); a built-in potential moves onto the device here
    void addPairwiseEnergy(PairwisePotential *potential) {
        pairwise_.push_back(potential);
        if (!potential->lccrfAttach(crf_, M_)) all_on_device_ = false;
    }

    virtual void setUnaryEnergy(const float *unary) = 0;
    virtual void setUnaryEnergyFromLabel(const short *label, float *confidences) = 0;
    virtual void setUnaryEnergyFromLabel(const short *label, float confidence = 0.5) = 0;

    // densecrf_base.h:65-73; getMap() / getProbability() stay valid until destruction (class-owned host arrays)
    virtual void inference(int n_iterations, bool with_map = false, float relax = 1.0) {
        if (all_on_device_) {
            lccrf_detail::check(lccrf_crf_inference(crf_, n_iterations, with_map ? 1 : 0, relax), "lccrf_crf_inference");
            return;
        }
        startInference();
        for (int it = 0; it < n_iterations; ++it) stepInference(relax);
        if (with_map) buildMap();
    }
    short *getMap() const { return const_cast<short *>(lccrf_crf_map(crf_)); }
    float *getProbability() const { return const_cast<float *>(lccrf_crf_prob(crf_)); }

    // stepwise API (densecrf_base.h:78-91)
    virtual void startInference() { lccrf_detail::check(lccrf_crf_start(crf_), "lccrf_crf_start"); }

    virtual void stepInference(float relax = 1.0) {
        if (all_on_device_) {
            lccrf_detail::check(lccrf_crf_step(crf_, relax), "lccrf_crf_step");
            return;
        }
        // plugin path: the reference's loop on host arrays (densecrf_base.h:82-91)
        const size_t n = (size_t)N_ * M_;
        if (!next_) {
            next_ = new float[n ? n : 1];
            tmp_ = new float[n ? n : 1];
            current_ = new float[n ? n : 1];
        }
        const float *q = lccrf_crf_prob(crf_);
        for (size_t i = 0; i < n; i++) current_[i] = q[i];
        stepInit();
        for (unsigned int i = 0; i < pairwise_.size(); i++) pairwise_[i]->apply(next_, current_, tmp_);
        expAndNormalize(current_, next_, 1.0, relax);
    }
};

}  // namespace DenseCRF

// potts_variant.inl -- shared body of PottsPotential3D<M,F> (pairwise3d.h:14-79) and
// PottsPotentialCPU<M,F> (pairwise_cpu.h:10-58).  Included with LCCRF_POTTS_NAME defined.
//
// The reference builds the lattice and norm_ in the constructor (pairwise3d.h:20-28).  Here the
// constructor keeps the features; the device lattice + norm are built when the potential is handed
// to a CRF (lccrfAttach, called by addPairwiseEnergy) or on the first stand-alone apply().

protected:
    float w_;
    float *features_;              // [N*F], released once the device owns the lattice
    mutable lccrf_crf *crf_;       // CRF this potential is attached to (or a private one for stand-alone use)
    mutable int index_;            // potential index inside crf_
    mutable bool private_crf_;

    void ensureDevice() const {
        if (crf_) return;
        lccrf_detail::check(lccrf_crf_create(lccrf_detail::context(), N_, M, &crf_), "lccrf_crf_create");
        private_crf_ = true;
        lccrf_detail::check(lccrf_crf_add_potts(crf_, features_, F, w_), "lccrf_crf_add_potts");
        index_ = 0;
    }

public:
    LCCRF_POTTS_NAME(const float *features, int N, float w)
        : PairwisePotential(N), w_(w), features_(nullptr), crf_(nullptr), index_(-1), private_crf_(false) {
        const size_t n = (size_t)N * F;
        features_ = new float[n ? n : 1];
        for (size_t i = 0; i < n; i++) features_[i] = features[i];
    }

    ~LCCRF_POTTS_NAME() {
        delete[] features_;
        if (private_crf_ && crf_) lccrf_crf_destroy(crf_);
    }

    LCCRF_POTTS_NAME(const LCCRF_POTTS_NAME &o) = delete;

    bool lccrfAttach(lccrf_crf *crf, int labels) override {
        if (!crf || labels != M || private_crf_) return false;
        lccrf_detail::check(lccrf_crf_add_potts(crf, features_, F, w_), "lccrf_crf_add_potts");
        crf_ = crf;
        index_ = lccrf_crf_num_potts(crf) - 1;
        delete[] features_;
        features_ = nullptr;
        return true;
    }

    // out_values[k] += w_*norm_[i]*filter(in_values)[k] on host arrays (pairwise3d.h:73-78)
    void apply(float *out_values, const float *in_values, float *tmp) const override {
        ensureDevice();
        lccrf_detail::check(lccrf_crf_potts_apply(crf_, index_, out_values, in_values, tmp), "lccrf_crf_potts_apply");
    }

// densecrf_variant.inl -- shared body of DenseCRF3D<M> (densecrf3d.h:13-158) and DenseCRFCPU<M>
// (densecrf_cpu.h:12-152): the two reference classes are the same driver under two names.
// Included with LCCRF_VARIANT_NAME defined; not a public header.

template <int M>
class LCCRF_VARIANT_NAME : public DenseCRF {
protected:
    // expAndNormalize / stepInit on host arrays: only reached on the plugin path; computed on the GPU
    void expAndNormalize(float *out, const float *in, float scale = 1.0, float relax = 1.0) override {
        lccrf_detail::check(lccrf_exp_and_normalize(lccrf_detail::context(), out, in, N_, M, scale, relax),
                            "lccrf_exp_and_normalize");
        lccrf_detail::check(lccrf_crf_set_prob(crf_, out), "lccrf_crf_set_prob");
    }
    void buildMap() override { lccrf_detail::check(lccrf_crf_build_map(crf_), "lccrf_crf_build_map"); }
    void stepInit() override { lccrf_detail::check(lccrf_crf_step_init(crf_, next_), "lccrf_crf_step_init"); }

public:
    // Create a dense CRF model of size N with M labels
    explicit LCCRF_VARIANT_NAME(int N) : DenseCRF(N) { lccrfCreate(M); }

    ~LCCRF_VARIANT_NAME() override {
        delete[] next_;
        delete[] tmp_;
        delete[] current_;
        next_ = tmp_ = current_ = nullptr;
        // potentials reference the device CRF: release them before it (the base dtor then sees none)
        for (auto *p : pairwise_) delete p;
        pairwise_.clear();
        lccrfDestroy();
    }

    LCCRF_VARIANT_NAME(LCCRF_VARIANT_NAME &o) = delete;

    // memory order is [x0l0 x0l1 x0l2 .. x1l0 x1l1 ...]
    void setUnaryEnergy(const float *unary) override {
        lccrf_detail::check(lccrf_crf_set_unary(crf_, unary), "lccrf_crf_set_unary");
    }

    void setUnaryEnergyFromLabel(const short *label, float confidence = 0.5) override {
        float confidences[M];
        for (int i = 0; i < M; ++i) confidences[i] = confidence;
        setUnaryEnergyFromLabel(label, confidences);
    }

    // label -1 = unknown.  The three energy tables use the reference's expressions (densecrf3d.h:109-114)
    // evaluated HERE, in the caller's translation unit, so log() resolves exactly as it would there.
    void setUnaryEnergyFromLabel(const short *label, float *confidences) override {
        float u_energy = -log(1.0f / M);
        float n_energies[M], p_energies[M];
        for (int i = 0; i < M; ++i) {
            n_energies[i] = -log((1.0f - confidences[i]) / (M - 1));
            p_energies[i] = -log(confidences[i]);
        }
        lccrf_detail::check(lccrf_crf_set_unary_from_label(crf_, label, u_energy, n_energies, p_energies),
                            "lccrf_crf_set_unary_from_label");
    }

#ifdef LCCRF_VARIANT_HAS_POKE
    void SetUnaryEnergtForPositiveNode(int idx, int m, float value) {  // (sic) densecrf3d.h:132-134
        lccrf_detail::check(lccrf_crf_set_unary_entry(crf_, idx, m, value), "lccrf_crf_set_unary_entry");
    }
#endif
};

// ref_shim.cpp -- C wrapper around the UNMODIFIED reference headers, compiled in place from
// /root/reference/Thirdparty/DenseCRF/include (never copied).  Output: oracle/_ref/libref.so.
// Test infrastructure (the "real reference" arm of the oracle), not product code.
//
// Include order mirrors src/Tracking.cc:21-41: `using namespace std;` is already active
// (include/Tracking.h:55) when densecrf3d.h / pairwise3d.h are parsed, so the unqualified
// log() in densecrf3d.h:109-113 binds to std::log(float).
#include <cmath>
#include <cstring>
#include <vector>
using namespace std;
// densecrf3d.h and densecrf_cpu.h both define DenseCRF::fast_exp, so the two variants live in
// two translation units of the same source file (-DREF_IMAGE_VARIANT selects the image one).
#ifndef REF_IMAGE_VARIANT
#include "densecrf3d.h"
#include "pairwise3d.h"
#else
#include "densecrf_cpu.h"
#include "pairwise_cpu.h"
#endif

using namespace DenseCRF;

namespace {
#ifndef REF_IMAGE_VARIANT
// protected members are reachable only from a subclass (permutohedral_cpu.h:174-187)
struct LatticeView : PermutohedralLatticeCPU {
    int V() const { return M_; }
    int Npts() const { return N_; }
    int dim() const { return d_; }
    const int *off() const { return offset_; }
    const float *bar() const { return barycentric_; }
    void nbr(int *out) const {
        for (int i = 0; i < (d_ + 1) * M_; i++) {
            out[2 * i] = blur_neighbors_[i].n1;
            out[2 * i + 1] = blur_neighbors_[i].n2;
        }
    }
};

template <int M>
int run3d(int N, const float *unary, const short *label, float conf, int K, const float *const *feat,
          const int *dims, const float *w, int iters, float relax, float *Q, short *map) {
    DenseCRF3D<M> crf(N);
    if (unary) crf.setUnaryEnergy(unary);
    else crf.setUnaryEnergyFromLabel(label, conf);
    for (int k = 0; k < K; k++) {
        if (dims[k] == 2) crf.addPairwiseEnergy(new PottsPotential3D<M, 2>(feat[k], N, w[k]));
        else if (dims[k] == 3) crf.addPairwiseEnergy(new PottsPotential3D<M, 3>(feat[k], N, w[k]));
        else if (dims[k] == 5) crf.addPairwiseEnergy(new PottsPotential3D<M, 5>(feat[k], N, w[k]));
        else return -1;
    }
    crf.inference(iters, map != nullptr, relax);
    memcpy(Q, crf.getProbability(), sizeof(float) * (size_t)N * M);
    if (map) memcpy(map, crf.getMap(), sizeof(short) * (size_t)N);
    return 0;
}

#else
template <int M>
int runimg(int W, int H, const short *label, float conf, float w_g, float sd_g, float w_b, float sd_b,
           const unsigned char *img, float sd_rgb, int iters, float *Q, short *map) {
    // example_cpu.cpp:80-98
    DenseCRFCPU<M> crf(W * H);
    crf.setUnaryEnergyFromLabel(label, conf);
    if (w_g > 0) crf.addPairwiseEnergy(PottsPotentialCPU<M, 2>::template FromImage<>(W, H, w_g, sd_g));
    if (w_b > 0)
        crf.addPairwiseEnergy(PottsPotentialCPU<M, 5>::template FromImage<unsigned char>(W, H, w_b, sd_b, img, sd_rgb));
    crf.inference(iters, map != nullptr);
    if (Q) memcpy(Q, crf.getProbability(), sizeof(float) * (size_t)W * H * M);
    if (map) memcpy(map, crf.getMap(), sizeof(short) * (size_t)W * H);
    return 0;
}
#endif
}  // namespace

extern "C" {
#ifndef REF_IMAGE_VARIANT

void *ref_lattice_init(const float *feature, int d, int N) {
    auto *l = new LatticeView();
    l->init(feature, d, N);
    return l;
}
void ref_lattice_free(void *h) { delete static_cast<LatticeView *>(h); }
int ref_lattice_V(void *h) { return static_cast<LatticeView *>(h)->V(); }
void ref_lattice_export(void *h, int *offset, float *bary, int *nbr) {
    auto *l = static_cast<LatticeView *>(h);
    size_t n = (size_t)l->Npts() * (l->dim() + 1);
    memcpy(offset, l->off(), n * sizeof(int));
    memcpy(bary, l->bar(), n * sizeof(float));
    l->nbr(nbr);
}
void ref_lattice_filter(void *h, float *out, const float *in, int L) {
    static_cast<LatticeView *>(h)->compute(out, in, L);
}
// the windowing arguments of PermutohedralLatticeCPU::compute (permutohedral_cpu.h:634-637)
void ref_lattice_filter_window(void *h, float *out, const float *in, int L, int in_offset, int out_offset, int in_size,
                               int out_size) {
    static_cast<LatticeView *>(h)->compute(out, in, L, in_offset, out_offset, in_size, out_size);
}

// DenseCRF3D<M> + PottsPotential3D<M,d>: generic unary / features (M in {2,3,4,21})
int ref_crf3d(int N, int M, const float *unary, const short *label, float conf, int K,
              const float *const *feat, const int *dims, const float *w, int iters, float relax,
              float *Q, short *map) {
    switch (M) {
        case 2: return run3d<2>(N, unary, label, conf, K, feat, dims, w, iters, relax, Q, map);
        case 3: return run3d<3>(N, unary, label, conf, K, feat, dims, w, iters, relax, Q, map);
        case 4: return run3d<4>(N, unary, label, conf, K, feat, dims, w, iters, relax, Q, map);
        case 21: return run3d<21>(N, unary, label, conf, K, feat, dims, w, iters, relax, Q, map);
    }
    return -1;
}

// The exact call sequence of src/Tracking.cc:1919-1930 (factories included).
int ref_slam_crf(int N, const float *observs, const float *error, const float *kp2d,
                 const short *init_label, float conf, float w1, float w2, float observ_stdev,
                 float rpj_stdev, float p3d_stdev, float p2d_stdev, int iters, float *Q, short *map) {
    vector<float> vobservs(observs, observs + N), verrors(error, error + N);
    vector<Point3f> vpoints(N);
    vector<Point2f> vcorrd2d(N);
    for (int i = 0; i < N; i++) vcorrd2d[i] = Point2f(kp2d[2 * i], kp2d[2 * i + 1]);
    const int M = 2;
    DenseCRF3D<M> crf(N);
    crf.setUnaryEnergyFromLabel(init_label, conf);
    auto *appearancePairwise = PottsPotential3D<M, 2>::appearanceKernel(N, w1, vobservs, verrors, observ_stdev, rpj_stdev);
    crf.addPairwiseEnergy(appearancePairwise);
    auto *smoothnessPairwise = PottsPotential3D<M, 2>::smoothKernel(N, w2, vpoints, vcorrd2d, p3d_stdev, p2d_stdev);
    crf.addPairwiseEnergy(smoothnessPairwise);
    crf.inference(iters, true);
    memcpy(Q, crf.getProbability(), sizeof(float) * (size_t)N * M);
    memcpy(map, crf.getMap(), sizeof(short) * (size_t)N);
    return 0;
}

// energies as the reference computes them in this translation unit (densecrf3d.h:109-113)
void ref_label_energies(int M, float conf, float *out3) {
    out3[0] = -log(1.0f / M);
    out3[1] = -log((1.0f - conf) / (M - 1));
    out3[2] = -log(conf);
}
#else
// DenseCRFCPU<M> + FromImage (example_cpu.cpp:80-98); M in {2,21}
int ref_image_crf(int W, int H, int M, const short *label, float conf, float w_g, float sd_g, float w_b,
                  float sd_b, const unsigned char *img, float sd_rgb, int iters, float *Q, short *map) {
    if (M == 2) return runimg<2>(W, H, label, conf, w_g, sd_g, w_b, sd_b, img, sd_rgb, iters, Q, map);
    if (M == 21) return runimg<21>(W, H, label, conf, w_g, sd_g, w_b, sd_b, img, sd_rgb, iters, Q, map);
    return -1;
}

#endif
}

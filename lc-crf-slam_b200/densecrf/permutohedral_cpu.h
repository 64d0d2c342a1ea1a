// permutohedral_cpu.h -- drop-in mirror of Thirdparty/DenseCRF/include/permutohedral_cpu.h:
// PermutohedralLatticeCPU::init / compute (float overload) forwarding to lccrf_lattice_*.
// HashTableCPU (:66-167) is an implementation detail of the reference's init() and has no
// counterpart here: vertex ids come from a device-side hash + prefix scan.
#pragma once

#include <cstdio>
#include <cstdlib>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

#include "densecrf_base.h"

namespace DenseCRF {

class PermutohedralLatticeCPU {
protected:
    lccrf_lattice *lat_;
    float *feature_copy_;  // kept so that copies can rebuild (the reference copies its arrays, :192-232)
    int N_, M_, d_;

    void rebuild(const float *feature, int d, int N) {
        release();
        N_ = N;
        d_ = d;
        const size_t n = (size_t)N * d;
        feature_copy_ = new float[n ? n : 1];
        for (size_t i = 0; i < n; i++) feature_copy_[i] = feature[i];
        lccrf_detail::check(lccrf_lattice_create(lccrf_detail::context(), feature_copy_, d, N, &lat_), "lccrf_lattice_create");
        lccrf_detail::check(lccrf_lattice_sizes(lat_, nullptr, nullptr, &M_), "lccrf_lattice_sizes");
    }
    void release() {
        if (lat_) lccrf_lattice_destroy(lat_);
        delete[] feature_copy_;
        lat_ = nullptr;
        feature_copy_ = nullptr;
    }

public:
    PermutohedralLatticeCPU() : lat_(nullptr), feature_copy_(nullptr), N_(0), M_(0), d_(0) {}
    PermutohedralLatticeCPU(const PermutohedralLatticeCPU &o) : lat_(nullptr), feature_copy_(nullptr), N_(0), M_(0), d_(0) {
        if (o.lat_) rebuild(o.feature_copy_, o.d_, o.N_);
    }
    PermutohedralLatticeCPU &operator=(const PermutohedralLatticeCPU &o) {
        if (&o == this) return *this;
        release();
        N_ = M_ = d_ = 0;
        if (o.lat_) rebuild(o.feature_copy_, o.d_, o.N_);
        return *this;
    }
    ~PermutohedralLatticeCPU() { release(); }

    void init(const float *feature, int feature_size, int N) { rebuild(feature, feature_size, N); }

    // compute(out, in, value_size, in_offset, out_offset, in_size, out_size)   permutohedral_cpu.h:634-699
    void compute(float *out, const float *in, int value_size, int in_offset = 0, int out_offset = 0, int in_size = -1,
                 int out_size = -1) const {
        lccrf_detail::check(lccrf_lattice_filter_window(lat_, out, in, value_size, in_offset, out_offset, in_size, out_size),
                            "lccrf_lattice_filter_window");
    }
#if defined(__SSE__)
    // the __m128 overload (permutohedral_cpu.h:572-633): value_size counts vectors of four floats
    void compute(__m128 *out, const __m128 *in, int value_size, int in_offset = 0, int out_offset = 0, int in_size = -1,
                 int out_size = -1) const {
        compute(reinterpret_cast<float *>(out), reinterpret_cast<const float *>(in), 4 * value_size, in_offset, out_offset,
                in_size, out_size);
    }
#endif

    // lccrf extensions (parity tests): reference-identical arrays
    int lccrfVertices() const { return M_; }
    void lccrfExport(int *offset, float *barycentric, int *neighbours) const {
        lccrf_detail::check(lccrf_lattice_export(lat_, offset, barycentric, neighbours), "lccrf_lattice_export");
    }
};

}  // namespace DenseCRF

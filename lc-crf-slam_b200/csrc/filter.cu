// filter.cu -- splat / blur / slice over a batch of permutohedral lattices.
// Replaces PermutohedralLatticeCPU::compute (Thirdparty/DenseCRF/include/permutohedral_cpu.h:634-699).
//
//   k_splat   (:653-661)  values[v] += bary * in[i]  as an order-independent integer reduction:
//             fp32 product (rounded like the reference's mulps), converted to 2^-40 fixed point,
//             segment-reduced inside the warp (match.any + redux.sync) and added with 64-bit
//             integer atomics -- deterministic for any thread order, no float atomics.
//   k_blur    (:663-679)  new[v] = old[v] + 0.5*(old[n1] + old[n2]), one pass per lattice axis,
//             coalesced over [vertex][label]; pass 0 converts the fixed-point sums, pass 1 re-zeroes them.
//   k_slice   (:684-694)  out[i] = sum_r (bary*alpha) * values[v_r]  (association of the SSE overload).
// The mean-field update fuses PottsPotential3D::apply (pairwise3d.h:73-78) into the slice.
#include <cfloat>

#include "engine.cuh"

namespace lccrf {

namespace {

// ---------------------------------------------------------------- splat
// one thread per point; all 32 lanes stay in the loop so the warp-level segmented reduction is legal
template <int LT>
__global__ void __launch_bounds__(kThreads)
k_splat(const int *__restrict__ offset, const float *__restrict__ bary, const float *__restrict__ in,
        long long *__restrict__ acc, int NT, int D, int L_rt, const float *__restrict__ in_scale) {
    const int L = LT > 0 ? LT : L_rt;
    const float sc = in_scale ? __ldg(in_scale) : 1.0f;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const bool valid = i < NT;
    const unsigned lane = threadIdx.x & 31;
    for (int r = 0; r < D; r++) {
        int id = -1;
        float w = 0.f;
        if (valid) {
            id = __ldg(offset + (size_t)i * D + r);
            w = __ldg(bary + (size_t)i * D + r);
        }
        const unsigned grp = __match_any_sync(0xffffffffu, id);
        const bool leader = (__ffs(grp) - 1) == (int)lane;
        for (int l = 0; l < L; l++) {
            long long fx = 0;
            if (valid) fx = to_fix(__fmul_rn(w, __fmul_rn(__ldg(in + (size_t)i * L + l), sc)));
            // |fx| < 2^42: 27-bit low part (sum of 32 fits u32), signed high part
            unsigned lo = (unsigned)(fx & 0x7ffffffll);
            int hi = (int)(fx >> 27);
            unsigned slo = __reduce_add_sync(grp, lo);
            int shi = __reduce_add_sync(grp, hi);
            if (leader && id >= 0) {
                long long tot = ((long long)shi << 27) + (long long)slo;
                if (tot != 0) atomicAdd((unsigned long long *)(acc + (size_t)id * L + l), (unsigned long long)tot);
            }
        }
    }
}

// ---------------------------------------------------------------- blur
// element-parallel over [vertex][label]; persistent grid-stride because V lives on the device
template <bool FROM_ACC, bool ZERO_ACC>
__global__ void __launch_bounds__(kThreads)
k_blur(const int2 *__restrict__ nbr_j, const long long *__restrict__ acc_in, long long *__restrict__ acc_zero,
       const float *__restrict__ src, float *__restrict__ dst, const int *__restrict__ vtotal, int L) {
    const long long total = (long long)__ldg(vtotal) * L;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (long long)gridDim.x * kThreads) {
        const int v = (int)(t / L), l = (int)(t - (long long)v * L);
        const int2 nb = __ldg(nbr_j + v);
        float o, a = 0.f, b = 0.f;
        if (FROM_ACC) {
            o = from_fix(__ldg(acc_in + t));
            if (nb.x >= 0) a = from_fix(__ldg(acc_in + (size_t)nb.x * L + l));
            if (nb.y >= 0) b = from_fix(__ldg(acc_in + (size_t)nb.y * L + l));
        } else {
            o = __ldg(src + t);
            if (nb.x >= 0) a = __ldg(src + (size_t)nb.x * L + l);
            if (nb.y >= 0) b = __ldg(src + (size_t)nb.y * L + l);
        }
        dst[t] = __fadd_rn(o, __fmul_rn(0.5f, __fadd_rn(a, b)));
        if (ZERO_ACC) acc_zero[t] = 0;
    }
}

// ---------------------------------------------------------------- slice (+ fused Potts apply)
enum SliceMode { kPlain = 0, kApplyFirst = 1, kApplyAdd = 2 };

template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_slice(const int *__restrict__ offset, const float *__restrict__ bary, const float *__restrict__ val,
        float *__restrict__ out, int NT, int D, int L, float alpha, const float *__restrict__ out_scale,
        float w, const float *__restrict__ norm, const float *__restrict__ unary) {
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= (long long)NT * L) return;
    const int i = (int)(t / L), l = (int)(t - (long long)i * L);
    float s = 0.0f;
    for (int r = 0; r < D; r++) {
        const int id = __ldg(offset + (size_t)i * D + r);
        const float wa = __fmul_rn(__ldg(bary + (size_t)i * D + r), alpha);
        s = __fadd_rn(s, __fmul_rn(wa, __ldg(val + (size_t)id * L + l)));
    }
    if (MODE == kPlain) {
        out[t] = out_scale ? __fmul_rn(s, __ldg(out_scale + 1)) : s;
    } else {
        const float m = __fmul_rn(__fmul_rn(w, __ldg(norm + i)), s);  // (w_*norm_[i])*tmp[k]
        const float base = MODE == kApplyFirst ? -__ldg(unary + t) : out[t];
        out[t] = __fadd_rn(base, m);
    }
}

__global__ void k_norm_finish(float *__restrict__ norm, int NT) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < NT) norm[i] = __fdiv_rn(1.0f, __fadd_rn(norm[i], 1e-20f));  // pairwise3d.h:26
}

__global__ void k_fill(float *__restrict__ x, float v, int n) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < n) x[i] = v;
}

// max |x| -> power-of-two scale pair {2^-e, 2^e} so that |x * 2^-e| <= 1 (generic-range filter inputs)
__global__ void __launch_bounds__(1024) k_absmax_scale(const float *__restrict__ x, long long n, float *scale2) {
    __shared__ float red[32];
    float m = 0.f;
    for (long long i = threadIdx.x; i < n; i += 1024) {
        float a = fabsf(x[i]);
        if (a > m && a <= FLT_MAX) m = a;
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) {
            int e = 0;
            if (m > 0.f) {
                frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
            }
            if (e < -100) e = -100;
            if (e > 100) e = 100;
            scale2[0] = ldexpf(1.0f, -e);
            scale2[1] = ldexpf(1.0f, e);
        }
    }
}

}  // namespace

// splat + all blur passes; *values_out is the buffer holding the blurred vertex values
int filter_splat_blur(Ctx *ctx, const Batch &b, LatticeSet *ls, const float *in_dev, int L,
                      const float *scale2_dev, const float **values_out) {
    cudaStream_t st = ctx->stream;
    if (L > ls->Lmax) return fail(LCCRF_ERR_ARG, "filter: L exceeds the lattice workspace");
    const int NT = b.NT, D = ls->D;
    if (NT > 0) {
        const int grid = cdiv(NT, kThreads);
        if (L == 1) { LCCRF_KERNEL(ctx, "k_splat"); k_splat<1><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, in_dev, ls->acc, NT, D, L, scale2_dev); }
        else if (L == 2) { LCCRF_KERNEL(ctx, "k_splat"); k_splat<2><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, in_dev, ls->acc, NT, D, L, scale2_dev); }
        else { LCCRF_KERNEL(ctx, "k_splat"); k_splat<0><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, in_dev, ls->acc, NT, D, L, scale2_dev); }
    }
    const int bgrid = persistent_grid((long long)ls->Vcap * L, kThreads, 8);
    const int *vt = ls->vbase + ls->B;
    float *src = ls->valA, *dst = ls->valB;
    for (int j = 0; j < D; j++) {
        const int2 *nb = ls->nbr + (size_t)j * ls->Vcap;
        if (j == 0) { LCCRF_KERNEL(ctx, "k_blur"); k_blur<true, false><<<bgrid, kThreads, 0, st>>>(nb, ls->acc, nullptr, nullptr, dst, vt, L); }
        else if (j == 1) { LCCRF_KERNEL(ctx, "k_blur"); k_blur<false, true><<<bgrid, kThreads, 0, st>>>(nb, nullptr, ls->acc, src, dst, vt, L); }
        else { LCCRF_KERNEL(ctx, "k_blur"); k_blur<false, false><<<bgrid, kThreads, 0, st>>>(nb, nullptr, nullptr, src, dst, vt, L); }
        float *t = src;
        src = dst;
        dst = t;
    }
    *values_out = src;
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

static int launch_slice(Ctx *ctx, int mode, const Batch &b, LatticeSet *ls, const float *values, float *out,
                        int L, const float *scale2, const float *unary) {
    if (b.NT == 0) return LCCRF_OK;
    const int grid = cdiv((long long)b.NT * L, kThreads);
    cudaStream_t st = ctx->stream;
    if (mode == kPlain)
        { LCCRF_KERNEL(ctx, "k_slice"); k_slice<kPlain><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, values, out, b.NT, ls->D, L, ls->alpha,
                                                    scale2, 0.f, nullptr, nullptr); }
    else if (mode == kApplyFirst)
        { LCCRF_KERNEL(ctx, "k_slice"); k_slice<kApplyFirst><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, values, out, b.NT, ls->D, L,
                                                         ls->alpha, nullptr, ls->w, ls->norm, unary); }
    else
        { LCCRF_KERNEL(ctx, "k_slice"); k_slice<kApplyAdd><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, values, out, b.NT, ls->D, L,
                                                       ls->alpha, nullptr, ls->w, ls->norm, nullptr); }
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

// PermutohedralLatticeCPU::compute(out, in, L); generic_range: inputs are arbitrary floats
int filter_full(Ctx *ctx, const Batch &b, LatticeSet *ls, float *out_dev, const float *in_dev, int L,
                bool generic_range) {
    const float *scale2 = nullptr;
    if (generic_range && b.NT > 0) {
        LCCRF_TRY(ctx_scratch(ctx, ctx->misc, 64));
        float *s2 = (float *)ctx->misc.p + 8;
        { LCCRF_KERNEL(ctx, "k_absmax_scale"); k_absmax_scale<<<1, 1024, 0, ctx->stream>>>(in_dev, (long long)b.NT * L, s2); }
        scale2 = s2;
    }
    const float *vals = nullptr;
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, in_dev, L, scale2, &vals));
    return launch_slice(ctx, kPlain, b, ls, vals, out_dev, L, scale2, nullptr);
}

// PottsPotential3D ctor: norm_ = 1/(filter(1)+1e-20)   pairwise3d.h:22-27
int potts_norm(Ctx *ctx, const Batch &b, LatticeSet *ls) {
    if (b.NT == 0) return LCCRF_OK;
    { LCCRF_KERNEL(ctx, "k_fill"); k_fill<<<cdiv(b.NT, kThreads), kThreads, 0, ctx->stream>>>(ls->norm, 1.0f, b.NT); }
    const float *vals = nullptr;
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, ls->norm, 1, nullptr, &vals));
    LCCRF_TRY(launch_slice(ctx, kPlain, b, ls, vals, ls->norm, 1, nullptr, nullptr));
    { LCCRF_KERNEL(ctx, "k_norm_finish"); k_norm_finish<<<cdiv(b.NT, kThreads), kThreads, 0, ctx->stream>>>(ls->norm, b.NT); }
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

// PottsPotential3D::apply on arbitrary device arrays: out += (w*norm)*filter(in)
int mf_potts_apply(Ctx *ctx, const Batch &b, LatticeSet *ls, float *out, const float *in, float *tmp, int L) {
    LCCRF_TRY(ctx_scratch(ctx, ctx->misc, 64));
    const float *scale2 = nullptr;
    const float *vals = nullptr;
    // generic inputs: the plugin path may hand us anything; scale to |x| <= 1, undo inside the apply
    // (power-of-two scaling commutes with every rounding step)
    if (b.NT > 0) {
        float *s2 = (float *)ctx->misc.p + 8;
        { LCCRF_KERNEL(ctx, "k_absmax_scale"); k_absmax_scale<<<1, 1024, 0, ctx->stream>>>(in, (long long)b.NT * L, s2); }
        scale2 = s2;
    }
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, in, L, scale2, &vals));
    // tmp = filter(in) (unscaled), then out += (w*norm)*tmp
    LCCRF_TRY(launch_slice(ctx, kPlain, b, ls, vals, tmp, L, scale2, nullptr));
    if (b.NT > 0) LCCRF_TRY(launch_axpy_norm(ctx, out, tmp, ls->norm, ls->w, b.NT, L));
    return LCCRF_OK;
}

// mean-field step pieces used by meanfield.cu
int mf_apply_fused(Ctx *ctx, const Batch &b, LatticeSet *ls, const float *cur, float *next, const float *unary,
                   bool first) {
    const float *vals = nullptr;
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, cur, b.L, nullptr, &vals));
    return launch_slice(ctx, first ? kApplyFirst : kApplyAdd, b, ls, vals, next, b.L, nullptr, unary);
}

}  // namespace lccrf

// filter.cu -- splat / blur / slice over a batch of permutohedral lattices.
// Replaces PermutohedralLatticeCPU::compute (Thirdparty/DenseCRF/include/permutohedral_cpu.h:634-699).
//
//   k_splat   (:653-661)  values[v] += bary * in[i]  as a segmented reduction over the vertex-sorted
//             entries of csr.cu, each row summed in point order: no atomics, bit-exact with the
//             reference's sequential float accumulation.
//   k_blur    (:663-679)  new[v] = old[v] + 0.5*(old[n1] + old[n2]), one pass per lattice axis,
//             coalesced over [vertex][label].
//   k_slice   (:684-694)  out[i] = sum_r (bary*alpha) * values[v_r]  (association of the SSE overload).
// The mean-field update fuses PottsPotential3D::apply (pairwise3d.h:73-78) into the slice.
#include <cfloat>

#include "engine.cuh"

namespace lccrf {

namespace {

// ---------------------------------------------------------------- splat
// Segmented reduction over the vertex-sorted entries (csr.cu).  A lane owns one (vertex row, label) pair
// and walks the row front to back: values[v][l] = (((0 + w0*x0) + w1*x1) + ...) in point order, every
// product and sum individually rounded -- the reference's splat loop (:653-661) bit for bit.  32/LP rows
// share a warp (LP = label count rounded up to a power of two), the labels of a row sit in adjacent lanes
// so the gathers of in[point][0..L) coalesce.
template <int LP>
__global__ void __launch_bounds__(kThreads)
k_splat(const int *__restrict__ row_ptr, const int2 *__restrict__ ent, const float *__restrict__ in,
        float *__restrict__ val, const int *__restrict__ vtotal, int L, int l0) {
    constexpr int G = 32 / LP;
    const int V = __ldg(vtotal);
    const int lane = threadIdx.x & 31;
    const int g = lane / LP, l = l0 + lane % LP;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    for (long long rb = (long long)warp * G; rb < V; rb += (long long)nwarps * G) {
        const int v = (int)rb + g;
        const bool act = v < V && l < L;
        int e = act ? __ldg(row_ptr + v) : 0;
        int e1 = act ? __ldg(row_ptr + v + 1) : 0;
        const bool mine = e1 - e < kMedRow;  // longer rows belong to k_splat_scan
        if (!mine) e1 = e;
        float acc = 0.0f;
        for (; e + 4 <= e1; e += 4) {  // independent loads first, then the ordered chain
            const int2 t0 = __ldg(ent + e), t1 = __ldg(ent + e + 1), t2 = __ldg(ent + e + 2), t3 = __ldg(ent + e + 3);
            const float x0 = __ldg(in + (size_t)t0.x * L + l), x1 = __ldg(in + (size_t)t1.x * L + l);
            const float x2 = __ldg(in + (size_t)t2.x * L + l), x3 = __ldg(in + (size_t)t3.x * L + l);
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(t0.y), x0));
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(t1.y), x1));
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(t2.y), x2));
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(t3.y), x3));
        }
        for (; e < e1; e++) {
            const int2 t = __ldg(ent + e);
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(t.y), __ldg(in + (size_t)t.x * L + l)));
        }
        if (act && mine) val[(size_t)v * L + l] = acc;
    }
}

// ---------------------------------------------------------------- exact ordered scan for long rows
// A lane-sequential walk costs one dependent FADD (4 cycles) plus load latency per entry, which is too slow for
// rows with thousands of entries.  The sequential fp32 sum s <- RN(s + c_k) can nevertheless be evaluated in
// parallel EXACTLY: while s stays inside one binade [2^E, 2^(E+1)) its ulp u = 2^(E-23) is constant, s = m*u with
// an integer m, and RN(s + c) = (m + n + r)*u where c/u = n + f (n = floor) and r = [f > 1/2], or, on a tie
// f == 1/2, whatever makes the result even (round-half-even).  "m -> m + a_parity(m)" maps compose
// associatively as pairs (a_even, a_odd), so a block of entries is an inclusive scan of such pairs.  The first
// entry whose prefix leaves the binade is added with one real FADD and the scan restarts behind it; a row of
// 100k entries needs ~17 restarts.  Verified against sequential summation incl. ties, signs, zeros and jumps.
struct ScanPair {
    int a0, a1;
};
__device__ __forceinline__ ScanPair scan_combine(ScanPair l, ScanPair r) {  // apply l first, then r
    ScanPair o;
    o.a0 = l.a0 + ((l.a0 & 1) ? r.a1 : r.a0);
    o.a1 = l.a1 + ((l.a1 & 1) ? r.a0 : r.a1);
    return o;
}

template <int NW>
struct ScanShared {
    ScanPair warp_tot[NW > 1 ? NW : 1];
    int first_bad[NW > 1 ? NW : 1];
    float bcast[2];
};

// Exact sequential sum of c over entries [e0, e1) for label l, computed by a group of NW warps; every thread owns
// IT consecutive entries of each chunk of NW*32*IT.  All threads of the group call this with identical arguments;
// tid = thread index inside the group.  Returns the same value in every thread.
template <int NW, int IT>
__device__ float row_sum_exact(const int2 *__restrict__ ent, int e0, int e1, const float *__restrict__ in, int L, int l,
                               int tid, ScanShared<NW> &sh) {
    constexpr int T = NW * 32, CH = T * IT;
    const int lane = tid & 31, wid = tid >> 5;
    float s = 0.0f;
    for (int cb = e0; cb < e1; cb += CH) {
        float c[IT];
        const int base = cb + tid * IT;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            c[k] = 0.0f;
            if (base + k < e1) {
                const int2 t = __ldg(ent + base + k);
                c[k] = __fmul_rn(__int_as_float(t.y), __ldg(in + (size_t)t.x * L + l));
            }
        }
        const int n_chunk = min(CH, e1 - cb);
        const int li0 = tid * IT;  // chunk-local index of this thread's first entry
        int st = 0;                // entries [0, st) of the chunk are already folded into s
        while (st < n_chunk) {
            const int E = ((__float_as_int(s) >> 23) & 0xff) - 127;
            const bool regular = (s > 0.0f) && E >= -100 && E <= 100;
            if (!regular) {
                // rare path (row start, zero / negative / huge running sums): one real addition at a time,
                // except that a run of exact zeros is skipped while s == +0
                if (NW > 1) __syncthreads();
                if (s == 0.0f) {
                    int first = CH;
#pragma unroll
                    for (int k = IT - 1; k >= 0; k--)
                        if (li0 + k >= st && li0 + k < n_chunk && c[k] != 0.0f) first = li0 + k;
#pragma unroll
                    for (int o = 16; o; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
                    if (NW > 1) {
                        if (lane == 0) sh.first_bad[wid] = first;
                        __syncthreads();
                        first = CH;
#pragma unroll
                        for (int w = 0; w < NW; w++) first = min(first, sh.first_bad[w]);
                        __syncthreads();
                    }
                    if (first >= n_chunk) {
                        st = n_chunk;
                        continue;
                    }
                    st = first;
                }
                float cj = 0.0f;
#pragma unroll
                for (int k = 0; k < IT; k++)
                    if (li0 + k == st) cj = c[k];
                if (NW == 1) cj = __shfl_sync(0xffffffffu, cj, st / IT);
                else {
                    if (tid == st / IT) sh.bcast[1] = cj;
                    __syncthreads();
                    cj = sh.bcast[1];
                }
                s = __fadd_rn(s, cj);
                st++;
                continue;
            }
            const float u = __int_as_float((E - 23 + 127) << 23);
            const float inv_u = __int_as_float((23 - E + 127) << 23);
            const int m0 = (int)__fmul_rn(s, inv_u);  // in [2^23, 2^24)
            // thread-local composition of this thread's entries + excursion bounds for either start parity
            ScanPair tot;
            tot.a0 = tot.a1 = 0;
            int hi0 = 0, lo0 = 0, hi1 = 0, lo1 = 0;
#pragma unroll
            for (int k = 0; k < IT; k++) {
                if (li0 + k >= st && li0 + k < n_chunk) {
                    ScanPair a;
                    const float q = __fmul_rn(c[k], inv_u);
                    if (!(fabsf(q) < 16777216.0f)) {  // also NaN / inf: forces "leaves the binade"
                        a.a0 = a.a1 = q > 0.0f ? (1 << 24) : -(1 << 24);
                    } else {
                        const float nf = floorf(q);
                        const float fr = __fsub_rn(q, nf);  // exact, in [0, 1)
                        const int ni = (int)nf;
                        if (fr == 0.5f) {
                            a.a0 = ni + (ni & 1);
                            a.a1 = ni + ((ni + 1) & 1);
                        } else {
                            a.a0 = a.a1 = ni + (fr > 0.5f ? 1 : 0);
                        }
                    }
                    tot = scan_combine(tot, a);
                    hi0 = max(hi0, tot.a0);
                    lo0 = min(lo0, tot.a0);
                    hi1 = max(hi1, tot.a1);
                    lo1 = min(lo1, tot.a1);
                }
            }
            // exclusive scan of the thread totals in thread order
            ScanPair inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                ScanPair lft;
                lft.a0 = __shfl_up_sync(0xffffffffu, inc.a0, o);
                lft.a1 = __shfl_up_sync(0xffffffffu, inc.a1, o);
                if (lane >= o) inc = scan_combine(lft, inc);
            }
            ScanPair exc;
            exc.a0 = __shfl_up_sync(0xffffffffu, inc.a0, 1);
            exc.a1 = __shfl_up_sync(0xffffffffu, inc.a1, 1);
            if (lane == 0) exc.a0 = exc.a1 = 0;
            if (NW > 1) {
                if (lane == 31) sh.warp_tot[wid] = inc;
                __syncthreads();
                // every warp scans the NW warp totals itself (no second barrier)
                ScanPair w = sh.warp_tot[lane < NW ? lane : 0];
                if (lane >= NW) w.a0 = w.a1 = 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    ScanPair lft;
                    lft.a0 = __shfl_up_sync(0xffffffffu, w.a0, o);
                    lft.a1 = __shfl_up_sync(0xffffffffu, w.a1, o);
                    if (lane >= o) w = scan_combine(lft, w);
                }
                ScanPair wp;  // composite of all warps before this one
                wp.a0 = __shfl_sync(0xffffffffu, w.a0, wid > 0 ? wid - 1 : 0);
                wp.a1 = __shfl_sync(0xffffffffu, w.a1, wid > 0 ? wid - 1 : 0);
                if (wid > 0) exc = scan_combine(wp, exc);
            }
            const int m_start = m0 + ((m0 & 1) ? exc.a1 : exc.a0);
            const int hi = (m_start & 1) ? hi1 : hi0, lo = (m_start & 1) ? lo1 : lo0;
            const bool ok = (m_start + lo >= (1 << 23)) && (m_start + hi < (1 << 24));
            const int m_end = m_start + ((m_start & 1) ? tot.a1 : tot.a0);
            const unsigned badmask = __ballot_sync(0xffffffffu, !ok);
            int tb;  // first thread whose entries leave the binade
            if (NW == 1) {
                tb = badmask ? __ffs(badmask) - 1 : T;
            } else {
                if (lane == 0) sh.first_bad[wid] = badmask ? wid * 32 + __ffs(badmask) - 1 : T;
                if (tid == T - 1) sh.bcast[0] = __fmul_rn((float)m_end, u);
                __syncthreads();
                tb = T;
#pragma unroll
                for (int w = NW - 1; w >= 0; w--) {
                    const int fb = sh.first_bad[w];
                    if (fb < T) tb = fb;
                }
            }
            if (tb == T) {  // the rest of the chunk stayed inside the binade
                if (NW == 1) s = __shfl_sync(0xffffffffu, __fmul_rn((float)m_end, u), 31);
                else s = sh.bcast[0];
                st = n_chunk;
            } else {  // threads before tb are consumed by the scan; thread tb adds its entries for real
                float sq = __fmul_rn((float)m_start, u);
                if (tid == tb) {
#pragma unroll
                    for (int k = 0; k < IT; k++)
                        if (li0 + k >= st && li0 + k < n_chunk) sq = __fadd_rn(sq, c[k]);
                }
                if (NW == 1) s = __shfl_sync(0xffffffffu, sq, tb);
                else {
                    if (tid == tb) sh.bcast[1] = sq;
                    __syncthreads();
                    s = sh.bcast[1];
                }
                st = min((tb + 1) * IT, n_chunk);
            }
        }
    }
    return s;
}

// one group (warp or CTA) per (row, label) task; rows come from a device-side list
template <int NW, int IT>
__global__ void __launch_bounds__(NW * 32 > kThreads ? NW * 32 : kThreads)
k_splat_scan(const int *__restrict__ row_ptr, const int2 *__restrict__ ent, const float *__restrict__ in,
             float *__restrict__ val, const int *__restrict__ list, const int *__restrict__ count, int L) {
    constexpr int GROUPS = NW == 1 ? kThreads / 32 : 1;  // groups per CTA
    __shared__ ScanShared<NW> sh[GROUPS];
    const long long n = (long long)__ldg(count) * L;
    const int gid = NW == 1 ? threadIdx.x >> 5 : 0;
    const int tid = NW == 1 ? (threadIdx.x & 31) : threadIdx.x;
    for (long long k = (long long)blockIdx.x * GROUPS + gid; k < n; k += (long long)gridDim.x * GROUPS) {
        const int v = __ldg(list + (int)(k / L)), l = (int)(k % L);
        const int e0 = __ldg(row_ptr + v), e1 = __ldg(row_ptr + v + 1);
        const float s = row_sum_exact<NW, IT>(ent, e0, e1, in, L, l, tid, sh[gid]);
        if (tid == 0) val[(size_t)v * L + l] = s;
        if (NW > 1) __syncthreads();
    }
}

// ---------------------------------------------------------------- blur
// element-parallel over [vertex][label]; persistent grid-stride because V lives on the device
__global__ void __launch_bounds__(kThreads)
k_blur(const int2 *__restrict__ nbr_j, const float *__restrict__ src, float *__restrict__ dst,
       const int *__restrict__ vtotal, int L) {
    const long long total = (long long)__ldg(vtotal) * L;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (long long)gridDim.x * kThreads) {
        const int v = (int)(t / L), l = (int)(t - (long long)v * L);
        const int2 nb = __ldg(nbr_j + v);
        const float o = __ldg(src + t);
        const float a = nb.x >= 0 ? __ldg(src + (size_t)nb.x * L + l) : 0.f;
        const float b = nb.y >= 0 ? __ldg(src + (size_t)nb.y * L + l) : 0.f;
        dst[t] = __fadd_rn(o, __fmul_rn(0.5f, __fadd_rn(a, b)));
    }
}

// ---------------------------------------------------------------- slice (+ fused Potts apply)
enum SliceMode { kPlain = 0, kApplyFirst = 1, kApplyAdd = 2 };

template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_slice(const int *__restrict__ offset, const float *__restrict__ bary, const float *__restrict__ val,
        float *__restrict__ out, int NT, int D, int L, float alpha, float w, const float *__restrict__ norm,
        const float *__restrict__ unary) {
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
        out[t] = s;
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

}  // namespace

// splat + all blur passes; *values_out is the buffer holding the blurred vertex values
int filter_splat_blur(Ctx *ctx, const Batch &b, LatticeSet *ls, const float *in_dev, int L, const float **values_out) {
    cudaStream_t st = ctx->stream;
    if (L > ls->Lmax) return fail(LCCRF_ERR_ARG, "filter: L exceeds the lattice workspace");
    (void)b;
    const int D = ls->D;
    const int *vt = ls->vbase + ls->B;
    float *src = ls->valA, *dst = ls->valB;
    {
        int LP = 1;
        while (LP < L && LP < 32) LP <<= 1;
        const int grid = persistent_grid((long long)ls->Vcap * LP, kThreads, 8);
        for (int l0 = 0; l0 < L; l0 += 32) {
            LCCRF_KERNEL(ctx, "k_splat");
            switch (LP) {
                case 1: k_splat<1><<<grid, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, vt, L, l0); break;
                case 2: k_splat<2><<<grid, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, vt, L, l0); break;
                case 4: k_splat<4><<<grid, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, vt, L, l0); break;
                case 8: k_splat<8><<<grid, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, vt, L, l0); break;
                case 16: k_splat<16><<<grid, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, vt, L, l0); break;
                default: k_splat<32><<<grid, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, vt, L, l0); break;
            }
        }
    }
    {   // medium rows: one warp each; long rows: one CTA each (exact ordered scan)
        LCCRF_KERNEL(ctx, "k_splat_scan_warp");
        k_splat_scan<1, 4><<<kNumSMs * 8, kThreads, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, ls->row_list_med, ls->row_counts, L);
    }
    {
        LCCRF_KERNEL(ctx, "k_splat_scan_cta");
        k_splat_scan<32, 8><<<kNumSMs * 2, 1024, 0, st>>>(ls->row_ptr, ls->csr_ent, in_dev, src, ls->row_list_long, ls->row_counts + 1, L);
    }
    const int bgrid = persistent_grid((long long)ls->Vcap * L, kThreads, 8);
    for (int j = 0; j < D; j++) {
        const int2 *nb = ls->nbr + (size_t)j * ls->Vcap;
        { LCCRF_KERNEL(ctx, "k_blur"); k_blur<<<bgrid, kThreads, 0, st>>>(nb, src, dst, vt, L); }
        float *t = src;
        src = dst;
        dst = t;
    }
    *values_out = src;
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

static int launch_slice(Ctx *ctx, int mode, const Batch &b, LatticeSet *ls, const float *values, float *out,
                        int L, const float *unary) {
    if (b.NT == 0) return LCCRF_OK;
    const int grid = cdiv((long long)b.NT * L, kThreads);
    cudaStream_t st = ctx->stream;
    LCCRF_KERNEL(ctx, "k_slice");
    if (mode == kPlain)
        k_slice<kPlain><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, values, out, b.NT, ls->D, L, ls->alpha, 0.f, nullptr, nullptr);
    else if (mode == kApplyFirst)
        k_slice<kApplyFirst><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, values, out, b.NT, ls->D, L, ls->alpha, ls->w, ls->norm, unary);
    else
        k_slice<kApplyAdd><<<grid, kThreads, 0, st>>>(ls->offset, ls->bary, values, out, b.NT, ls->D, L, ls->alpha, ls->w, ls->norm, nullptr);
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

// PermutohedralLatticeCPU::compute(out, in, L)
int filter_full(Ctx *ctx, const Batch &b, LatticeSet *ls, float *out_dev, const float *in_dev, int L) {
    const float *vals = nullptr;
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, in_dev, L, &vals));
    return launch_slice(ctx, kPlain, b, ls, vals, out_dev, L, nullptr);
}

// PottsPotential3D ctor: norm_ = 1/(filter(1)+1e-20)   pairwise3d.h:22-27
int potts_norm(Ctx *ctx, const Batch &b, LatticeSet *ls) {
    if (b.NT == 0) return LCCRF_OK;
    { LCCRF_KERNEL(ctx, "k_fill"); k_fill<<<cdiv(b.NT, kThreads), kThreads, 0, ctx->stream>>>(ls->norm, 1.0f, b.NT); }
    const float *vals = nullptr;
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, ls->norm, 1, &vals));
    LCCRF_TRY(launch_slice(ctx, kPlain, b, ls, vals, ls->norm, 1, nullptr));
    { LCCRF_KERNEL(ctx, "k_norm_finish"); k_norm_finish<<<cdiv(b.NT, kThreads), kThreads, 0, ctx->stream>>>(ls->norm, b.NT); }
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

// PottsPotential3D::apply on arbitrary device arrays: tmp = filter(in); out += (w*norm)*tmp
int mf_potts_apply(Ctx *ctx, const Batch &b, LatticeSet *ls, float *out, const float *in, float *tmp, int L) {
    LCCRF_TRY(filter_full(ctx, b, ls, tmp, in, L));
    if (b.NT > 0) LCCRF_TRY(launch_axpy_norm(ctx, out, tmp, ls->norm, ls->w, b.NT, L));
    return LCCRF_OK;
}

// one potential of a mean-field step: next = (first ? -unary : next) + (w*norm)*filter(cur)
int mf_apply_fused(Ctx *ctx, const Batch &b, LatticeSet *ls, const float *cur, float *next, const float *unary,
                   bool first) {
    const float *vals = nullptr;
    LCCRF_TRY(filter_splat_blur(ctx, b, ls, cur, b.L, &vals));
    return launch_slice(ctx, first ? kApplyFirst : kApplyAdd, b, ls, vals, next, b.L, unary);
}

}  // namespace lccrf

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
// Segmented reduction over the vertex-sorted entries (csr.cu).  values[v][l] = (((0 + w0*x0) + w1*x1) + ...)
// over the row of v in point order, every product and sum individually rounded -- the reference's splat loop
// (:653-661) bit for bit.  Three kernels:
//   k_products      prod[e][l] = bary[e] * in[point[e]][l] for every entry, in vertex-sorted order: one fully
//                   parallel, coalesced pass (the only gather of the splat)
//   k_splat_staged  rows shorter than kLongRow: a warp owns G = 32/LP consecutive rows (LP = label count rounded up
//                   to a power of two) = one contiguous slab of prod; the slab is copied to shared memory with wide
//                   coalesced loads, then lane (row, label) adds its row front to back -- ordered, no DRAM latency
//                   in the dependent chain
//   k_splat_scan    rows of kLongRow entries and more: exact parallel scan (below)
template <int LT>
__global__ void __launch_bounds__(kThreads)
k_products(const int2 *__restrict__ ent, const float *__restrict__ in, float *__restrict__ prod, long long E, int L_rt) {
    const int L = LT > 0 ? LT : L_rt;
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= E) return;
    const int2 t = __ldg(ent + e);
    const float w = __int_as_float(t.y);
    if (LT == 2) {
        const float2 x = __ldg((const float2 *)in + t.x);
        ((float2 *)prod)[e] = make_float2(__fmul_rn(w, x.x), __fmul_rn(w, x.y));
    } else if (LT == 1) {
        prod[e] = __fmul_rn(w, __ldg(in + t.x));
    } else {
        for (int l = 0; l < L; l++) prod[e * L + l] = __fmul_rn(w, __ldg(in + (size_t)t.x * L + l));
    }
}

constexpr int kStageFloats = 2048;  // per warp
__device__ __forceinline__ int spad(int idx) { return idx + (idx >> 5); }

template <int LP>
__global__ void __launch_bounds__(kThreads)
k_splat_staged(const int *__restrict__ row_ptr, const float *__restrict__ prod, float *__restrict__ val,
               const int *__restrict__ vtotal, int L) {
    constexpr int G = 32 / LP;
    extern __shared__ float s_all[];
    const int V = __ldg(vtotal);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *s_prod = s_all + (size_t)wid * (kStageFloats + kStageFloats / 32);
    const int g = lane / LP;
    const int cap = (kStageFloats - 4) / L;  // entries per staged chunk (4 floats of alignment slack)
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    for (int l0 = 0; l0 < L; l0 += LP) {  // L > 32: several label passes over the same rows
        const int l = l0 + lane % LP;
        for (long long rb = (long long)warp * G; rb < V; rb += (long long)nwarps * G) {
            const int v = (int)rb + g;
            const int rs = v < V ? __ldg(row_ptr + v) : 0;
            const int re = v < V ? __ldg(row_ptr + v + 1) : rs;
            const bool mine = v < V && l < L && (re - rs) < kLongRow;  // longer rows belong to k_splat_scan
            const int E0 = __shfl_sync(0xffffffffu, rs, 0);
            const int E1 = __ldg(row_ptr + min((int)rb + G, V));
            float acc = 0.0f;
            int gcur = 0;  // row of the group that contains cb (moves forward only)
            for (int cb = E0; cb < E1;) {
                int rs_c, re_c;
                for (;;) {
                    rs_c = __shfl_sync(0xffffffffu, rs, gcur * LP);
                    re_c = __shfl_sync(0xffffffffu, re, gcur * LP);
                    if (re_c > cb || gcur == G - 1) break;
                    gcur++;
                }
                if (re_c - rs_c >= kLongRow && cb >= rs_c && cb < re_c) {  // inside a long row: jump over it
                    cb = re_c;
                    continue;
                }
                const int ce = min(cb + cap, E1);
                // phase 1: the slab prod[cb*L, ce*L) -> shared memory with 128-bit loads from the enclosing
                // 16-byte aligned window, all lanes, every load independent (one memory round trip per chunk)
                const long long f0 = (long long)cb * L;
                const long long a0 = f0 & ~3ll;
                const int shift = (int)(f0 - a0);
                const int nq = ((ce - cb) * L + shift + 3) >> 2;  // float4 count (prod is padded by 4 floats)
                {   // nq <= 512: at most 16 float4 per lane, all loaded before the first shared store
                    float4 x[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int q = lane + 32 * j;
                        if (q < nq) x[j] = __ldg((const float4 *)(prod + a0) + q);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int q = lane + 32 * j;
                        if (q < nq) {
                            const int o = spad(q * 4);
                            s_prod[o] = x[j].x;
                            s_prod[o + 1] = x[j].y;
                            s_prod[o + 2] = x[j].z;
                            s_prod[o + 3] = x[j].w;
                        }
                    }
                }
                __syncwarp();
                // phase 2: ordered accumulation
                if (mine) {
                    const int a = max(rs, cb), z = min(re, ce);
                    int e = a;
                    for (; e + 8 <= z; e += 8) {  // independent shared loads first, then the ordered chain
                        float x[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) x[q] = s_prod[spad((e + q - cb) * L + l + shift)];
#pragma unroll
                        for (int q = 0; q < 8; q++) acc = __fadd_rn(acc, x[q]);
                    }
                    for (; e < z; e++) acc = __fadd_rn(acc, s_prod[spad((e - cb) * L + l + shift)]);
                }
                __syncwarp();
                cb = ce;
            }
            if (mine) val[(size_t)v * L + l] = acc;
        }
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
__device__ float row_sum_exact(const float *__restrict__ prod, int e0, int e1, int L, int l, int tid, ScanShared<NW> &sh) {
    constexpr int T = NW * 32, CH = T * IT;
    const int lane = tid & 31, wid = tid >> 5;
    float s = 0.0f;
    float cn[IT];  // next chunk's contributions, loaded while the current chunk is being scanned
#pragma unroll
    for (int k = 0; k < IT; k++) {
        cn[k] = 0.0f;
        if (e0 + tid * IT + k < e1) cn[k] = __ldg(prod + (size_t)(e0 + tid * IT + k) * L + l);
    }
    for (int cb = e0; cb < e1; cb += CH) {
        float c[IT];
#pragma unroll
        for (int k = 0; k < IT; k++) c[k] = cn[k];
        {
            const int nbase = cb + CH + tid * IT;
#pragma unroll
            for (int k = 0; k < IT; k++) {
                cn[k] = 0.0f;
                if (nbase + k < e1) cn[k] = __ldg(prod + (size_t)(nbase + k) * L + l);
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
                    const float q = __fmul_rn(c[k], inv_u);
                    int ni;
                    float fr = 0.0f;
                    if (!(fabsf(q) < 16777216.0f)) {  // also NaN / inf: forces "leaves the binade"
                        ni = q > 0.0f ? (1 << 24) : -(1 << 24);
                    } else {
                        ni = __float2int_rd(q);
                        fr = __fsub_rn(q, (float)ni);  // exact, in [0, 1)
                    }
                    if (fr != 0.5f) {  // common case: the same increment for either parity
                        const int inc1 = ni + (fr > 0.5f ? 1 : 0);
                        tot.a0 += inc1;
                        tot.a1 += inc1;
                    } else {           // tie: round half to even
                        ScanPair a;
                        a.a0 = ni + (ni & 1);
                        a.a1 = ni + ((ni + 1) & 1);
                        tot = scan_combine(tot, a);
                    }
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
k_splat_scan(const int *__restrict__ row_ptr, const float *__restrict__ prod, float *__restrict__ val,
             const int *__restrict__ list, const int *__restrict__ count, int L) {
    constexpr int GROUPS = NW == 1 ? kThreads / 32 : 1;  // groups per CTA
    __shared__ ScanShared<NW> sh[GROUPS];
    const long long n = (long long)__ldg(count) * L;
    const int gid = NW == 1 ? threadIdx.x >> 5 : 0;
    const int tid = NW == 1 ? (threadIdx.x & 31) : threadIdx.x;
    for (long long k = (long long)blockIdx.x * GROUPS + gid; k < n; k += (long long)gridDim.x * GROUPS) {
        const int v = __ldg(list + (int)(k / L)), l = (int)(k % L);
        const int e0 = __ldg(row_ptr + v), e1 = __ldg(row_ptr + v + 1);
        const float s = row_sum_exact<NW, IT>(prod, e0, e1, L, l, tid, sh[gid]);
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

// all D blur passes of one problem inside one CTA (shared-memory ping-pong when the lattice fits, global otherwise):
// replaces D launch-bound passes when a batch holds several problems or the lattice is small
constexpr int kBlurFusedFloats = 24576;  // 96 KB
__global__ void __launch_bounds__(1024)
k_blur_fused(const int2 *__restrict__ nbr, int Vcap, const int *__restrict__ vbase, float *__restrict__ A,
             float *__restrict__ B, int L, int D) {
    extern __shared__ float s_blur[];
    const int b = blockIdx.x;
    const int vb = __ldg(vbase + b);
    const int n = (__ldg(vbase + b + 1) - vb) * L;
    float *gA = A + (size_t)vb * L, *gB = B + (size_t)vb * L;
    if (2 * n <= kBlurFusedFloats) {
        float *src = s_blur, *dst = s_blur + n;
        for (int t = threadIdx.x; t < n; t += 1024) src[t] = gA[t];
        __syncthreads();
        for (int j = 0; j < D; j++) {
            const int2 *nb_j = nbr + (size_t)j * Vcap + vb;
            for (int t = threadIdx.x; t < n; t += 1024) {
                const int v = t / L, l = t - v * L;
                const int2 nb = __ldg(nb_j + v);
                const float a = nb.x >= 0 ? src[(nb.x - vb) * L + l] : 0.f;
                const float c = nb.y >= 0 ? src[(nb.y - vb) * L + l] : 0.f;
                dst[t] = __fadd_rn(src[t], __fmul_rn(0.5f, __fadd_rn(a, c)));
            }
            __syncthreads();
            float *tmp = src;
            src = dst;
            dst = tmp;
        }
        for (int t = threadIdx.x; t < n; t += 1024) gB[t] = src[t];
    } else {
        float *src = gA, *dst = gB;
        for (int j = 0; j < D; j++) {
            const int2 *nb_j = nbr + (size_t)j * Vcap + vb;
            for (int t = threadIdx.x; t < n; t += 1024) {
                const int v = t / L, l = t - v * L;
                const int2 nb = __ldg(nb_j + v);
                const float a = nb.x >= 0 ? src[(size_t)(nb.x - vb) * L + l] : 0.f;
                const float c = nb.y >= 0 ? src[(size_t)(nb.y - vb) * L + l] : 0.f;
                dst[t] = __fadd_rn(src[t], __fmul_rn(0.5f, __fadd_rn(a, c)));
            }
            __syncthreads();
            float *tmp = src;
            src = dst;
            dst = tmp;
        }
        if (src != gB)
            for (int t = threadIdx.x; t < n; t += 1024) gB[t] = src[t];
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
    const int D = ls->D;
    const int *vt = ls->vbase + ls->B;
    float *src = ls->valA, *dst = ls->valB;
    const long long E = (long long)b.NT * D;
    if (E > 0) {
        LCCRF_KERNEL(ctx, "k_products");
        const int grid = cdiv(E, kThreads);
        if (L == 1) k_products<1><<<grid, kThreads, 0, st>>>(ls->csr_ent, in_dev, ls->prod, E, L);
        else if (L == 2) k_products<2><<<grid, kThreads, 0, st>>>(ls->csr_ent, in_dev, ls->prod, E, L);
        else k_products<0><<<grid, kThreads, 0, st>>>(ls->csr_ent, in_dev, ls->prod, E, L);
    }
    {
        int LP = 1;
        while (LP < L && LP < 32) LP <<= 1;
        const int grid = persistent_grid((long long)ls->Vcap * LP, kThreads, 3);
        const size_t smem = (size_t)(kThreads / 32) * (kStageFloats + kStageFloats / 32) * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            LCCRF_CUDA(cudaFuncSetAttribute(k_splat_staged<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LCCRF_CUDA(cudaFuncSetAttribute(k_splat_staged<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LCCRF_CUDA(cudaFuncSetAttribute(k_splat_staged<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LCCRF_CUDA(cudaFuncSetAttribute(k_splat_staged<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LCCRF_CUDA(cudaFuncSetAttribute(k_splat_staged<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LCCRF_CUDA(cudaFuncSetAttribute(k_splat_staged<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        LCCRF_KERNEL(ctx, "k_splat_staged");
        switch (LP) {
            case 1: k_splat_staged<1><<<grid, kThreads, smem, st>>>(ls->row_ptr, ls->prod, src, vt, L); break;
            case 2: k_splat_staged<2><<<grid, kThreads, smem, st>>>(ls->row_ptr, ls->prod, src, vt, L); break;
            case 4: k_splat_staged<4><<<grid, kThreads, smem, st>>>(ls->row_ptr, ls->prod, src, vt, L); break;
            case 8: k_splat_staged<8><<<grid, kThreads, smem, st>>>(ls->row_ptr, ls->prod, src, vt, L); break;
            case 16: k_splat_staged<16><<<grid, kThreads, smem, st>>>(ls->row_ptr, ls->prod, src, vt, L); break;
            default: k_splat_staged<32><<<grid, kThreads, smem, st>>>(ls->row_ptr, ls->prod, src, vt, L); break;
        }
    }
    {   // long rows: one CTA each, exact ordered scan
        LCCRF_KERNEL(ctx, "k_splat_scan_cta");
        k_splat_scan<8, 8><<<kNumSMs * 8, kThreads, 0, st>>>(ls->row_ptr, ls->prod, src, ls->row_list_long, ls->row_counts + 1, L);
    }
    if (b.B >= 2 || b.maxN <= 32768) {  // one CTA per problem runs all D passes
        static bool attr_set = false;
        if (!attr_set) {
            LCCRF_CUDA(cudaFuncSetAttribute(k_blur_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kBlurFusedFloats * (int)sizeof(float)));
            attr_set = true;
        }
        LCCRF_KERNEL(ctx, "k_blur_fused");
        k_blur_fused<<<b.B, 1024, kBlurFusedFloats * sizeof(float), st>>>(ls->nbr, ls->Vcap, ls->vbase, src, dst, L, D);
        *values_out = dst;
        LCCRF_CUDA(cudaGetLastError());
        return LCCRF_OK;
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

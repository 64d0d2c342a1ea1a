// frontend.cu -- the per-frame feeders of the CRF that LC-CRF-SLAM adds in front of it:
//   k_epipolar_prior   Tracking::GetFeature2EpipolarDis   src/Tracking.cc:2030-2047
//                      + FundamentalMatrixEstimator::symmetricEpipolarDistance
//                        Thirdparty/graph-cut-ransac-master/include/fundamental_estimator.h:90-127
//   k_bf_knn2 / k_bf_finish   Tracking::BfMatch   src/Tracking.cc:1747-1766
//                      (cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2) + ratio test 0.6)
// Both are embarrassingly parallel per match / per query descriptor; all arithmetic that reaches a result is
// IEEE double with every operation individually rounded (the reference is built without FMA contraction).
#include <climits>

#include "engine.cuh"

namespace lccrf {

namespace {

struct F33 {
    double f[9];  // row-major fundamental matrix, descriptor(r, c) = f[3 * r + c]
};

// one thread per match; dis/prob are scattered by feature id into the flat form of the reference's
// std::map<int,double> mvFeatureMatchDis / mvFeatureMatchProb (absent ids stay 0.0 = what operator[] inserts, :2003)
__global__ void __launch_bounds__(kThreads)
k_epipolar_prior(int M, const int *__restrict__ fid1, const float2 *__restrict__ pt1, const float2 *__restrict__ pt2,
                 F33 F, float u_gamma, float stdev_gamma, int nFeat, double *__restrict__ dis_by_fid,
                 double *__restrict__ prob_by_fid, double *__restrict__ dis_m, double *__restrict__ prob_m) {
    const int m = blockIdx.x * kThreads + threadIdx.x;
    if (m >= M) return;
    const float2 a = __ldg(pt1 + m), b = __ldg(pt2 + m);
    // Point2d(point_1) / Point2d(point_2): float -> double (:2037-2039)
    const double x1 = (double)a.x, y1 = (double)a.y, x2 = (double)b.x, y2 = (double)b.y;
    const double f11 = F.f[0], f12 = F.f[1], f13 = F.f[2], f21 = F.f[3], f22 = F.f[4], f23 = F.f[5], f31 = F.f[6],
                 f32 = F.f[7], f33 = F.f[8];
    // fundamental_estimator.h:110-116 (left-to-right, each op rounded)
    const double l1 = __dadd_rn(__dadd_rn(__dmul_rn(f11, x2), __dmul_rn(f21, y2)), f31);
    const double l2 = __dadd_rn(__dadd_rn(__dmul_rn(f12, x2), __dmul_rn(f22, y2)), f32);
    const double l3 = __dadd_rn(__dadd_rn(__dmul_rn(f13, x2), __dmul_rn(f23, y2)), f33);
    const double t1 = __dadd_rn(__dadd_rn(__dmul_rn(f11, x1), __dmul_rn(f12, y1)), f13);
    const double t2 = __dadd_rn(__dadd_rn(__dmul_rn(f21, x1), __dmul_rn(f22, y1)), f23);
    const double t3 = __dadd_rn(__dadd_rn(__dmul_rn(f31, x1), __dmul_rn(f32, y1)), f33);
    // :118-122.  NOTE b1 uses y1 (not y2) exactly as the reference does (:121) -- kept for parity.
    const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(l1, x1), __dmul_rn(l2, y1)), l3);
    const double a2 = __dsqrt_rn(__dadd_rn(__dmul_rn(l1, l1), __dmul_rn(l2, l2)));
    const double b1 = __dadd_rn(__dadd_rn(__dmul_rn(t1, x2), __dmul_rn(t2, y1)), t3);
    const double b2 = __dsqrt_rn(__dadd_rn(__dmul_rn(t1, t1), __dmul_rn(t2, t2)));
    const double d1 = __ddiv_rn(a1, a2), d2 = __ddiv_rn(b1, b2);
    const double dis = fabs(__dmul_rn(0.5, __dadd_rn(d1, d2)));  // :127
    // exp(-(dis-mGcMean)*(dis-mGcMean)/(2*mGcStdev*mGcStdev))  Tracking.cc:2043; the denominator is float arithmetic
    const double c = __dsub_rn(dis, (double)u_gamma);
    const float den = __fmul_rn(__fmul_rn(2.0f, stdev_gamma), stdev_gamma);
    const double prob = exp(__ddiv_rn(__dmul_rn(-c, c), (double)den));
    if (dis_m) dis_m[m] = dis;
    if (prob_m) prob_m[m] = prob;
    if (fid1) {
        const int f = __ldg(fid1 + m);
        if (f >= 0 && f < nFeat) {  // duplicates cannot occur: asso is a std::map keyed by fid1
            if (dis_by_fid) dis_by_fid[f] = dis;
            if (prob_by_fid) prob_by_fid[f] = prob;
        }
    }
}

// ---------------------------------------------------------------- brute-force Hamming kNN (k = 2)
constexpr int kBfQ = 128;     // queries per CTA (one per thread, descriptor in 8 registers)
constexpr int kBfTile = 256;  // train descriptors staged in shared memory per round (8 KB)

struct Top2 {
    int d0, i0, d1, i1;
};

// candidate (d, j) arriving in ascending j: strict comparisons keep the first-seen minimum first, which is the
// order cv::batchDistance's K-best insertion produces (pinned against cv2 4.13, tests/golden/make_golden_frontend.py)
__device__ __forceinline__ void top2_push(Top2 &t, int d, int j) {
    if (d < t.d0) {
        t.d1 = t.d0;
        t.i1 = t.i0;
        t.d0 = d;
        t.i0 = j;
    } else if (d < t.d1) {
        t.d1 = d;
        t.i1 = j;
    }
}

// grid (query tiles, train splits, frame pairs).  Every thread owns one query descriptor; the CTA streams its split
// of the train descriptors through shared memory (all lanes read the same 32 bytes: broadcast, conflict-free) and
// keeps the two best candidates in registers.  Work per pair: 8 XOR + 8 POPC + 7 IADD -- bound by the POPC pipe.
__global__ void __launch_bounds__(kBfQ)
k_bf_knn2(const int *__restrict__ q_ptr, const uint4 *__restrict__ desc_q, const int *__restrict__ t_ptr,
          const uint4 *__restrict__ desc_t, int4 *__restrict__ part, int S) {
    __shared__ uint4 s_t[kBfTile * 2];
    const int b = blockIdx.z, s = blockIdx.y;
    const int q0 = __ldg(q_ptr + b), nq = __ldg(q_ptr + b + 1) - q0;
    const int t0 = __ldg(t_ptr + b), nt = __ldg(t_ptr + b + 1) - t0;
    const int qb = blockIdx.x * kBfQ;
    if (qb >= nq) return;
    const int q = qb + threadIdx.x;
    const bool qv = q < nq;
    uint4 qa = make_uint4(0, 0, 0, 0), qc = qa;
    if (qv) {
        qa = __ldg(desc_q + 2 * (size_t)(q0 + q));
        qc = __ldg(desc_q + 2 * (size_t)(q0 + q) + 1);
    }
    // contiguous split of the train range, tile-aligned so that splits stay in index order
    const int per = ((nt + S - 1) / S + kBfTile - 1) / kBfTile * kBfTile;
    const int ts = min(s * per, nt), te = min(ts + per, nt);
    Top2 best = {INT_MAX, -1, INT_MAX, -1};
    for (int tb = ts; tb < te; tb += kBfTile) {
        const int n = min(kBfTile, te - tb);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += kBfQ) s_t[i] = __ldg(desc_t + 2 * (size_t)(t0 + tb) + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; j++) {
            const uint4 ta = s_t[2 * j], tc = s_t[2 * j + 1];
            const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w) +
                          __popc(qc.x ^ tc.x) + __popc(qc.y ^ tc.y) + __popc(qc.z ^ tc.z) + __popc(qc.w ^ tc.w);
            top2_push(best, d, tb + j);
        }
    }
    if (qv) part[(size_t)(q0 + q) * S + s] = make_int4(best.d0, best.i0, best.d1, best.i1);
}

// merge the S partial lists of every query in split order (stable), then the ratio test of Tracking.cc:1755:
// match[0].distance < match[1].distance * 0.6  -- float distances promoted to double, the product rounded once
__global__ void __launch_bounds__(kThreads)
k_bf_finish(int NQ, const int4 *__restrict__ part, int S, double ratio, int *__restrict__ match, int4 *__restrict__ knn,
            int *__restrict__ n_match) {
    const int q = blockIdx.x * kThreads + threadIdx.x;
    int accepted = 0;
    if (q < NQ) {
        Top2 best = {INT_MAX, -1, INT_MAX, -1};
        for (int s = 0; s < S; s++) {
            const int4 p = __ldg(part + (size_t)q * S + s);
            if (p.y >= 0) top2_push(best, p.x, p.y);
            if (p.w >= 0) top2_push(best, p.z, p.w);
        }
        int m = -1;
        // knnMatch returns fewer than 2 neighbours when the train set has fewer than 2 rows: "match.size() == 2" fails
        if (best.i1 >= 0 && (double)best.d0 < __dmul_rn((double)best.d1, ratio)) m = best.i0;
        match[q] = m;
        if (knn) knn[q] = make_int4(best.i0 >= 0 ? best.d0 : -1, best.i0, best.i1 >= 0 ? best.d1 : -1, best.i1);
        accepted = m >= 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, accepted);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_match, __popc(bal));
}

}  // namespace

int epipolar_prior(Ctx *ctx, int M, const int *fid1, const float *pt1, const float *pt2, const double *F9, float u_gamma,
                   float stdev_gamma, int nFeat, double *dis_by_fid, double *prob_by_fid, double *dis_m, double *prob_m) {
    if (M <= 0) return LCCRF_OK;
    F33 F;
    for (int i = 0; i < 9; i++) F.f[i] = F9[i];
    LCCRF_KERNEL(ctx, "k_epipolar_prior");
    k_epipolar_prior<<<cdiv(M, kThreads), kThreads, 0, ctx->stream>>>(M, fid1, (const float2 *)pt1, (const float2 *)pt2, F,
                                                                      u_gamma, stdev_gamma, nFeat, dis_by_fid, prob_by_fid,
                                                                      dis_m, prob_m);
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

// number of train splits: enough CTAs to cover the GPU when there are few frame pairs, 1 for large batches
int bf_match_splits(int B, int max_nq, int max_nt) {
    const long long tiles = (long long)B * ((max_nq + kBfQ - 1) / kBfQ);
    if (tiles <= 0) return 1;
    long long S = (2LL * kNumSMs + tiles - 1) / tiles;
    const long long cap = (max_nt + kBfTile - 1) / kBfTile;
    if (S > cap) S = cap;
    if (S > 32) S = 32;
    return (int)(S < 1 ? 1 : S);
}

int bf_match(Ctx *ctx, int B, int NQ, int max_nq, int max_nt, const int *q_ptr, const void *desc_q, const int *t_ptr,
             const void *desc_t, double ratio, int S, void *part, int *match, int *knn, int *n_match) {
    cudaStream_t st = ctx->stream;
    LCCRF_CUDA(cudaMemsetAsync(n_match, 0, sizeof(int), st));
    if (NQ <= 0) return LCCRF_OK;
    (void)max_nt;
    {
        LCCRF_KERNEL(ctx, "k_bf_knn2");
        dim3 grid((max_nq + kBfQ - 1) / kBfQ, S, B);
        k_bf_knn2<<<grid, kBfQ, 0, st>>>(q_ptr, (const uint4 *)desc_q, t_ptr, (const uint4 *)desc_t, (int4 *)part, S);
    }
    {
        LCCRF_KERNEL(ctx, "k_bf_finish");
        k_bf_finish<<<cdiv(NQ, kThreads), kThreads, 0, st>>>(NQ, (const int4 *)part, S, ratio, match, (int4 *)knn, n_match);
    }
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

}  // namespace lccrf

/*
 * lccrf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See lccrf_oracle.h.
 *
 * Scalar fp32 restatement of the reference hot path.  Build:
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared lccrf_oracle.c -lm
 * (x86-64 baseline ISA: SSE2 scalar math, no FMA, MXCSR round-to-nearest-even, no FTZ/DAZ --
 * the same arithmetic the reference's SSE branch performs lane by lane.)
 */
#include "lccrf_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Vertex dictionary: key (d shorts) -> dense id in first-insertion order.
 * Replaces HashTableCPU (permutohedral_cpu.h:66-167).  Only the id assignment rule
 * ("id = number of distinct keys inserted before", :146) is observable; the reference's hash
 * function, capacity and growth policy (:79-111) do not influence ids and are not restated.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int d, cap, size, key_cap;
    int *slot;   /* cap entries, -1 = empty, else vertex id */
    short *keys; /* key_cap*d */
} vdict;

static uint64_t vd_hash(const short *k, int d) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < d; i++) {
        h ^= (uint16_t)k[i];
        h *= 0x100000001B3ull;
        h ^= h >> 29;
    }
    return h;
}

static void vd_init(vdict *t, int d, int expect) {
    t->d = d;
    t->cap = 1024;
    while (t->cap < 2 * expect) t->cap *= 2;
    t->size = 0;
    t->key_cap = expect > 16 ? expect : 16;
    t->slot = (int *)malloc(sizeof(int) * (size_t)t->cap);
    memset(t->slot, -1, sizeof(int) * (size_t)t->cap);
    t->keys = (short *)malloc(sizeof(short) * (size_t)t->key_cap * (size_t)d);
}

static void vd_free(vdict *t) {
    free(t->slot);
    free(t->keys);
}

static void vd_grow(vdict *t) {
    int ncap = t->cap * 2;
    int *ns = (int *)malloc(sizeof(int) * (size_t)ncap);
    memset(ns, -1, sizeof(int) * (size_t)ncap);
    for (int id = 0; id < t->size; id++) {
        uint64_t h = vd_hash(t->keys + (size_t)id * t->d, t->d) & (uint64_t)(ncap - 1);
        while (ns[h] >= 0) h = (h + 1) & (uint64_t)(ncap - 1);
        ns[h] = id;
    }
    free(t->slot);
    t->slot = ns;
    t->cap = ncap;
}

/* find(k, create): id of k, or -1 when absent and !create (permutohedral_cpu.h:134-161) */
static int vd_find(vdict *t, const short *k, int create) {
    if (create && 2 * (t->size + 1) > t->cap) vd_grow(t);
    uint64_t h = vd_hash(k, t->d) & (uint64_t)(t->cap - 1);
    for (;;) {
        int e = t->slot[h];
        if (e < 0) {
            if (!create) return -1;
            if (t->size == t->key_cap) {
                t->key_cap *= 2;
                t->keys = (short *)realloc(t->keys, sizeof(short) * (size_t)t->key_cap * (size_t)t->d);
            }
            memcpy(t->keys + (size_t)t->size * t->d, k, sizeof(short) * (size_t)t->d);
            t->slot[h] = t->size;
            return t->size++;
        }
        if (memcmp(t->keys + (size_t)e * t->d, k, sizeof(short) * (size_t)t->d) == 0) return e;
        h = (h + 1) & (uint64_t)(t->cap - 1);
    }
}

/* ------------------------------------------------------------------------------------------
 * PermutohedralLatticeCPU::init, SSE branch (permutohedral_cpu.h:241-424), one lane at a time.
 * The SSE build walks points four at a time; lanes >= N carry feature 0.0 (:299) and ARE
 * inserted into the hash table (:364-377): "phantom points" k in [N, ceil4(N)).
 * ------------------------------------------------------------------------------------------ */
#define ORC_MAXD 16

orc_lattice *orc_lattice_init(const float *feature, int d, int N) {
    if (d < 1 || d > ORC_MAXD || N < 0) return NULL;
    const int D = d + 1;
    orc_lattice *l = (orc_lattice *)calloc(1, sizeof(*l));
    l->N = N;
    l->d = d;
    l->offset = (int *)calloc((size_t)(N + 4) * D, sizeof(int));
    l->bary = (float *)calloc((size_t)(N + 4) * D, sizeof(float));

    vdict tab;
    vd_init(&tab, d, N > 0 ? N : 1);

    /* constants, permutohedral_cpu.h:249-250,274-285 */
    const float invD = 1.0f / (float)D;
    const float fD = (float)D;
    short canonical[(ORC_MAXD + 1) * (ORC_MAXD + 1)];
    for (int i = 0; i <= d; i++) {
        for (int j = 0; j <= d - i; j++) canonical[i * D + j] = (short)i;
        for (int j = d - i + 1; j <= d; j++) canonical[i * D + j] = (short)(i - D);
    }
    float scale[ORC_MAXD];
    {
        float inv_std_dev = (float)(sqrt(2.0 / 3.0) * D); /* :282, double math, one rounding */
        for (int i = 0; i < d; i++)                       /* :285, double math then float */
            scale[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);
    }

    const int Npad = (N + 3) / 4 * 4; /* :294 blocksize 4 */
    float el[ORC_MAXD + 1], rem[ORC_MAXD + 1], rank[ORC_MAXD + 1], b[ORC_MAXD + 2];
    short key[ORC_MAXD + 1];
    for (int k = 0; k < Npad; k++) {
        /* elevate, :304-310 */
        float sm = 0.0f;
        for (int j = d; j > 0; j--) {
            float f = k < N ? feature[(size_t)k * d + (j - 1)] : 0.0f; /* :299 */
            float cf = f * scale[j - 1];
            el[j] = sm - (float)j * cf;
            sm = sm + cf;
        }
        el[0] = sm;
        /* nearest 0-coloured simplex, :313-323; cvtps_epi32 under MXCSR nearest == rintf (half-even) */
        float sum = 0.0f;
        for (int i = 0; i <= d; i++) {
            float v = rintf(invD * el[i]);
            rem[i] = v * fD;
            sum = sum + v;
        }
        /* rank, :326-336; ties go to the higher index */
        for (int i = 0; i <= d; i++) rank[i] = 0.0f;
        for (int i = 0; i < d; i++) {
            float di = el[i] - rem[i];
            for (int j = i + 1; j <= d; j++) {
                float dj = el[j] - rem[j];
                float c = (di < dj) ? 1.0f : 0.0f;
                rank[i] = rank[i] + c;
                rank[j] = rank[j] + (1.0f - c);
            }
        }
        /* back onto the plane, :339-345; both masks from the same (pre-update) value */
        for (int i = 0; i <= d; i++) {
            rank[i] = rank[i] + sum;
            float add = (rank[i] < 0.0f) ? fD : 0.0f;
            float sub = (rank[i] >= fD) ? fD : 0.0f;
            rank[i] = rank[i] + (add - sub);
            rem[i] = rem[i] + (add - sub);
        }
        /* barycentric, :348-366 */
        for (int i = 0; i <= d + 1; i++) b[i] = 0.0f;
        for (int i = 0; i <= d; i++) {
            float v = (el[i] - rem[i]) * invD;
            int p = (int)((float)d - rank[i]);
            b[p] = b[p] + v;
            b[p + 1] = b[p + 1] - v;
        }
        b[0] = b[0] + (1.0f + b[d + 1]);
        /* vertices + ids, :371-377 */
        for (int r = 0; r <= d; r++) {
            for (int i = 0; i < d; i++) key[i] = (short)(rem[i] + (float)canonical[r * D + (int)rank[i]]);
            int id = vd_find(&tab, key, 1);
            l->offset[(size_t)k * D + r] = id;
            l->bary[(size_t)k * D + r] = b[r];
        }
    }

    /* neighbours, :398-421 */
    l->V = tab.size;
    const int V = l->V;
    l->nbr = (int *)malloc(sizeof(int) * 2 * (size_t)D * (size_t)(V > 0 ? V : 1));
    l->keys = (short *)malloc(sizeof(short) * (size_t)d * (size_t)(V > 0 ? V : 1));
    memcpy(l->keys, tab.keys, sizeof(short) * (size_t)d * (size_t)V);
    short n1[ORC_MAXD + 1], n2[ORC_MAXD + 1];
    for (int j = 0; j <= d; j++) {
        for (int i = 0; i < V; i++) {
            const short *kk = l->keys + (size_t)i * d;
            for (int c = 0; c < d; c++) {
                n1[c] = (short)(kk[c] - 1);
                n2[c] = (short)(kk[c] + 1);
            }
            if (j < d) { /* for j == d the write lands outside the d hashed coords (:415-416) */
                n1[j] = (short)(kk[j] + d);
                n2[j] = (short)(kk[j] - d);
            }
            l->nbr[2 * ((size_t)j * V + i) + 0] = vd_find(&tab, n1, 0);
            l->nbr[2 * ((size_t)j * V + i) + 1] = vd_find(&tab, n2, 0);
        }
    }
    vd_free(&tab);
    return l;
}

void orc_lattice_free(orc_lattice *l) {
    if (!l) return;
    free(l->offset);
    free(l->bary);
    free(l->nbr);
    free(l->keys);
    free(l);
}

/* ------------------------------------------------------------------------------------------
 * PermutohedralLatticeCPU::compute(float*,...) SSE overload (permutohedral_cpu.h:634-699),
 * default windowing (in_offset = out_offset = 0, sizes = N).
 * ------------------------------------------------------------------------------------------ */
void orc_lattice_filter(const orc_lattice *l, float *out, const float *in, int L) {
    const int N = l->N, d = l->d, D = d + 1, V = l->V;
    /* slot 0 == "neighbour -1" (:640-650) */
    float *val = (float *)calloc((size_t)(V + 2) * L, sizeof(float));
    float *nval = (float *)calloc((size_t)(V + 2) * L, sizeof(float));
    /* splat in point order, :653-661 */
    for (int i = 0; i < N; i++)
        for (int j = 0; j <= d; j++) {
            int o = l->offset[(size_t)i * D + j] + 1;
            float w = l->bary[(size_t)i * D + j];
            for (int k = 0; k < L; k++) {
                float t = w * in[(size_t)i * L + k];
                val[(size_t)o * L + k] = val[(size_t)o * L + k] + t;
            }
        }
    /* blur, :663-679 : new = old + 0.5*(n1 + n2) */
    for (int j = 0; j <= d; j++) {
        for (int i = 0; i < V; i++) {
            int n1 = l->nbr[2 * ((size_t)j * V + i)] + 1;
            int n2 = l->nbr[2 * ((size_t)j * V + i) + 1] + 1;
            for (int k = 0; k < L; k++) {
                float s = val[(size_t)n1 * L + k] + val[(size_t)n2 * L + k];
                float h = 0.5f * s;
                nval[(size_t)(i + 1) * L + k] = val[(size_t)(i + 1) * L + k] + h;
            }
        }
        float *t = val;
        val = nval;
        nval = t;
    }
    /* alpha, :681 ; slice, :684-694 : weight = bary*alpha first, then times value */
    float alpha = 1.0f / (1 + powf(2, (float)-d));
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < L; k++) {
            float acc = 0.0f;
            for (int j = 0; j <= d; j++) {
                int o = l->offset[(size_t)i * D + j] + 1;
                float w = l->bary[(size_t)i * D + j] * alpha;
                float t = w * val[(size_t)o * L + k];
                acc = acc + t;
            }
            out[(size_t)i * L + k] = acc;
        }
    }
    free(val);
    free(nval);
}

/* PottsPotential3D ctor, pairwise3d.h:20-28 */
void orc_potts_norm(const orc_lattice *l, float *norm) {
    for (int i = 0; i < l->N; i++) norm[i] = 1.0f;
    orc_lattice_filter(l, norm, norm, 1);
    for (int i = 0; i < l->N; i++) norm[i] = 1.0f / (norm[i] + 1e-20f);
}

/* PottsPotential3D::apply, pairwise3d.h:73-78 */
void orc_potts_apply(const orc_lattice *l, const float *norm, float w, float *out, const float *in,
                     float *tmp, int L) {
    orc_lattice_filter(l, tmp, in, L);
    for (int i = 0, k = 0; i < l->N; i++)
        for (int j = 0; j < L; j++, k++) {
            float wn = w * norm[i];
            float t = wn * tmp[k];
            out[k] = out[k] + t;
        }
}

/* very_fast_exp / fast_exp, densecrf3d.h:51-67 */
static float orc_very_fast_exp(float x) {
    float p = 0.0001413161f;
    p = 0.0013298820f - x * p;
    p = 0.0083013598f - x * p;
    p = 0.0416573475f - x * p;
    p = 0.1666653019f - x * p;
    p = 0.4999999206f - x * p;
    p = 0.9999999995f - x * p;
    return 1.0f - x * p;
}

float orc_fast_exp(float x) {
    int lessZero = 1;
    if (x < 0) {
        lessZero = 0;
        x = -x;
    }
    if (x > 20) return 0;
    int mult = 0;
    while ((double)x > 0.69 * 2 * 2 * 2) { /* thresholds are double constants, :60-62 */
        mult += 3;
        x = x / 8.0f;
    }
    while ((double)x > 0.69 * 2 * 2) {
        mult += 2;
        x = x / 4.0f;
    }
    while ((double)x > 0.69) {
        mult++;
        x = x / 2.0f;
    }
    x = orc_very_fast_exp(x);
    while (mult) {
        mult--;
        x = x * x;
    }
    return lessZero ? 1.0f / x : x;
}

/* DenseCRF3D<M>::expAndNormalize, densecrf3d.h:71-98 */
void orc_exp_and_normalize(float *out, const float *in, int N, int L, float scale, float relax) {
    float *Vv = (float *)malloc(sizeof(float) * (size_t)(L > 0 ? L : 1));
    for (int i = 0; i < N; i++) {
        const float *b = in + (size_t)i * L;
        float mx = scale * b[0];
        for (int j = 1; j < L; j++)
            if (mx < scale * b[j]) mx = scale * b[j];
        float tt = 0;
        for (int j = 0; j < L; j++) {
            Vv[j] = orc_fast_exp(scale * b[j] - mx);
            tt = tt + Vv[j];
        }
        for (int j = 0; j < L; j++) Vv[j] = Vv[j] / tt;
        float *a = out + (size_t)i * L;
        for (int j = 0; j < L; j++)
            if (relax == 1)
                a[j] = Vv[j];
            else
                a[j] = (1 - relax) * a[j] + relax * Vv[j];
    }
    free(Vv);
}

/* DenseCRF3D<M>::setUnaryEnergyFromLabel, densecrf3d.h:108-130 (energies computed by caller) */
void orc_unary_from_label(float *unary, const short *label, int N, int L, float u_energy,
                          const float *n_energies, const float *p_energies) {
    for (int i = 0; i < N; i++) {
        short t = label[i];
        if (t == -1) {
            for (int m = 0; m < L; m++) unary[(size_t)i * L + m] = u_energy;
        } else {
            for (int m = 0; m < L; m++) unary[(size_t)i * L + m] = n_energies[t];
            unary[(size_t)i * L + t] = p_energies[t];
        }
    }
}

/* DenseCRF3D<M>::buildMap, densecrf3d.h:137-151 : first maximum wins */
void orc_build_map(short *map, const float *Q, int N, int L) {
    for (int i = 0; i < N; i++) {
        const float *p = Q + (size_t)i * L;
        float mx = p[0];
        short imx = 0;
        for (short m = 1; m < L; m++)
            if (mx < p[m]) {
                mx = p[m];
                imx = m;
            }
        map[i] = imx;
    }
}

/* DenseCRF::inference, densecrf_base.h:65-91 */
void orc_meanfield(int N, int L, const float *unary, int K, const orc_lattice *const *lat,
                   const float *const *norm, const float *w, int iters, float relax, float *Q,
                   short *map) {
    size_t n = (size_t)N * L;
    float *next = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *tmp = (float *)malloc(sizeof(float) * (n ? n : 1));
    orc_exp_and_normalize(Q, unary, N, L, -1.0f, 1.0f); /* startInference, :78-80 */
    for (int it = 0; it < iters; it++) {
        for (size_t i = 0; i < n; i++) next[i] = -unary[i]; /* stepInit, densecrf3d.h:155-158 */
        for (int k = 0; k < K; k++) orc_potts_apply(lat[k], norm[k], w[k], next, Q, tmp, L);
        orc_exp_and_normalize(Q, next, N, L, 1.0f, relax);
    }
    if (map) orc_build_map(map, Q, N, L);
    free(next);
    free(tmp);
}

/* appearanceKernel / smoothKernel feature assembly, pairwise3d.h:38-48,52-71 */
void orc_features_div2(float *feat, const float *a, const float *b, int N, float sa, float sb,
                       int stride_a, int stride_b) {
    for (int i = 0; i < N; i++) {
        feat[2 * (size_t)i + 0] = a[(size_t)i * stride_a] / sa;
        feat[2 * (size_t)i + 1] = b[(size_t)i * stride_b] / sb;
    }
}

/* PottsPotentialCPU::FromImage, pairwise_cpu.h:34-50 */
void orc_features_image(float *feat, int W, int H, int F, float posdev, const unsigned char *img_u8,
                        const float *img_f32, float featuredev) {
    for (int hi = 0; hi < H; hi++)
        for (int wi = 0; wi < W; wi++) {
            size_t idx = (size_t)hi * W + wi;
            feat[idx * F + 0] = (float)wi / posdev;
            feat[idx * F + 1] = (float)hi / posdev;
            for (int i = 2; i < F; i++) {
                float v = img_u8 ? (float)img_u8[idx * (F - 2) + (i - 2)] : img_f32[idx * (F - 2) + (i - 2)];
                feat[idx * F + i] = v / featuredev;
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * Tracking::ComputeMapPointErrAndObserv, src/Tracking.cc:1803-1839, on flat arrays.
 * Rcw*x3Dw+tcw (:1818) is OpenCV gemm, restated as the sequential fp32 expression
 * ((r0*x0 + r1*x1) + r2*x2) + t; pinned bit for bit by the cv::gemm-based fixture
 * tests/golden/golden_unary.npz (see header); observation order = CSR order.
 * ------------------------------------------------------------------------------------------ */
void orc_map_point_unary(int N, const float *xyz, const int *obs_ptr, const int *obs_kf,
                         const float *obs_uv, const float *kf_pose, const float *kf_intr,
                         const float *kf_bounds, float *observs, float *error, float *depth) {
    for (int i = 0; i < N; i++) {
        int n = obs_ptr[i + 1] - obs_ptr[i]; /* :1808 */
        float err = 0.0f, dep = 0.0f;
        float x0 = xyz[3 * (size_t)i], x1 = xyz[3 * (size_t)i + 1], x2 = xyz[3 * (size_t)i + 2];
        for (int e = obs_ptr[i]; e < obs_ptr[i + 1]; e++) {
            const float *P = kf_pose + 12 * (size_t)obs_kf[e];
            const float *K = kf_intr + 4 * (size_t)obs_kf[e];
            const float *B = kf_bounds + 4 * (size_t)obs_kf[e];
            float xc = ((P[0] * x0 + P[1] * x1) + P[2] * x2) + P[3]; /* :1818 */
            float yc = ((P[4] * x0 + P[5] * x1) + P[6] * x2) + P[7];
            float zc = ((P[8] * x0 + P[9] * x1) + P[10] * x2) + P[11];
            float invz = (float)(1.0 / (double)zc); /* :1821 */
            if (invz < 0) continue;                 /* :1823 */
            float u = K[0] * xc * invz + K[2];      /* :1825 */
            float v = K[1] * yc * invz + K[3];      /* :1826 */
            if (u < B[0] || u > B[1] || v < B[2] || v > B[3]) continue; /* :1828 */
            double kx = (double)obs_uv[2 * (size_t)e], ky = (double)obs_uv[2 * (size_t)e + 1]; /* :1832 */
            double du = (double)u - kx, dv = (double)v - ky;
            float e_ = (float)sqrt(du * du + dv * dv); /* :1833 */
            err = err + e_;                            /* :1834 */
            dep = dep + zc;                            /* :1835 */
        }
        if (n > 0) {
            err = err / (float)n; /* :1837 divides by ALL observations */
            dep = dep / (float)n; /* :1838 */
        }
        observs[i] = (float)n; /* vobservs is vector<float>, :1867 */
        error[i] = err;
        depth[i] = dep;
    }
}

/* Tracking::RroughClassify, src/Tracking.cc:1961-2013 */
void orc_rough_classify(int N, const float *observs, const float *error, const float *depth,
                        const double *p4, const orc_slam_params *prm, short *label) {
    float observ_sigma2 = prm->stdev_beta * prm->stdev_beta;        /* :1964 */
    float rpjerror_sigma2 = prm->stdev_alpha * prm->stdev_alpha;    /* :1965 */
    float depth_sigma2 = prm->point3d_stdev * prm->point3d_stdev;   /* :1966 */
    for (int i = 0; i < N; i++) {
        float a = observs[i] - prm->u_beta;
        float k1 = a * a / (2 * observ_sigma2); /* :1972 */
        float b = error[i] - prm->u_alpha;
        float k2 = b * b / (2 * rpjerror_sigma2); /* :1973 */
        float c = depth[i] - prm->u_depth;
        float k3 = c * c / (2 * depth_sigma2); /* :1974 */
        float p1 = expf(-k1), p2 = expf(-k2), p3 = expf(-k3); /* :1975 (std::exp(float)) */
        if (!p4) {
            label[i] = (p1 + p2 + p3 <= prm->pth) ? 0 : 1; /* :1996-1999 */
        } else {
            label[i] = ((double)(p1 + p2 + p3) + p4[i] <= (double)prm->pth + 0.2) ? 0 : 1; /* :2003-2009 */
        }
    }
}

/* src/Tracking.cc:1919-1930 */
void orc_slam_crf(int N, const float *observs, const float *error, const float *kp2d,
                  const short *init_label, const float *energies, const orc_slam_params *prm,
                  float *Q, short *map, int *V_out) {
    const int L = 2;
    float *unary = (float *)malloc(sizeof(float) * (size_t)(N ? N : 1) * L);
    float n_en[2] = {energies[1], energies[1]}, p_en[2] = {energies[2], energies[2]};
    orc_unary_from_label(unary, init_label, N, L, energies[0], n_en, p_en);
    float *feat = (float *)malloc(sizeof(float) * 2 * (size_t)(N ? N : 1));
    orc_lattice *lat[2];
    float *norm[2];
    /* appearanceKernel(N, w1, vobservs, verrors, mObservStdev, mRpjErrorStdev), :1923 */
    orc_features_div2(feat, observs, error, N, prm->stdev_beta, prm->stdev_alpha, 1, 1);
    lat[0] = orc_lattice_init(feat, 2, N);
    /* smoothKernel(N, w2, vpoints, vcorrd2d, mPoint3dStdev, mPoint2dStdev): 2-D branch only, :1926 */
    orc_features_div2(feat, kp2d, kp2d + 1, N, prm->point2d_stdev, prm->point2d_stdev, 2, 2);
    lat[1] = orc_lattice_init(feat, 2, N);
    for (int k = 0; k < 2; k++) {
        norm[k] = (float *)malloc(sizeof(float) * (size_t)(N ? N : 1));
        orc_potts_norm(lat[k], norm[k]);
    }
    float w[2] = {prm->w1, prm->w2};
    orc_meanfield(N, L, unary, 2, (const orc_lattice *const *)lat, (const float *const *)norm, w,
                  prm->iters, 1.0f, Q, map);
    if (V_out) {
        V_out[0] = lat[0]->V;
        V_out[1] = lat[1]->V;
    }
    for (int k = 0; k < 2; k++) {
        orc_lattice_free(lat[k]);
        free(norm[k]);
    }
    free(feat);
    free(unary);
}

/* ------------------------------------------------------------------------------------------
 * Frontend feeders: epipolar prior and brute-force descriptor matching
 * ------------------------------------------------------------------------------------------ */

/* fundamental_estimator.h:90-127 + Tracking.cc:2037-2045 */
void orc_epipolar_prior(int M, const float *pt1, const float *pt2, const double *F, float u_gamma,
                        float stdev_gamma, double *dis_out, double *prob_out) {
    const double f11 = F[0], f12 = F[1], f13 = F[2], f21 = F[3], f22 = F[4], f23 = F[5], f31 = F[6], f32 = F[7],
                 f33 = F[8]; /* :99-107 */
    for (int m = 0; m < M; m++) {
        const double x1 = (double)pt1[2 * m], y1 = (double)pt1[2 * m + 1]; /* Tracking.cc:2037-2039 */
        const double x2 = (double)pt2[2 * m], y2 = (double)pt2[2 * m + 1];
        const double l1 = f11 * x2 + f21 * y2 + f31; /* :109-111 */
        const double l2 = f12 * x2 + f22 * y2 + f32;
        const double l3 = f13 * x2 + f23 * y2 + f33;
        const double t1 = f11 * x1 + f12 * y1 + f13; /* :113-115 */
        const double t2 = f21 * x1 + f22 * y1 + f23;
        const double t3 = f31 * x1 + f32 * y1 + f33;
        const double a1 = l1 * x1 + l2 * y1 + l3; /* :117 */
        const double a2 = sqrt(l1 * l1 + l2 * l2); /* :118 */
        const double b1 = t1 * x2 + t2 * y1 + t3; /* :120 -- y1 as in the reference */
        const double b2 = sqrt(t1 * t1 + t2 * t2); /* :121 */
        const double d1 = a1 / a2, d2 = b1 / b2; /* :123-124 */
        const double dis = fabs(0.5 * (d1 + d2)); /* :126 */
        /* Tracking.cc:2043: exp(- (dis-mGcMean)*(dis-mGcMean)/(2*mGcStdev*mGcStdev)), float denominator */
        const float den = 2 * stdev_gamma * stdev_gamma;
        const double prob = exp(-(dis - u_gamma) * (dis - u_gamma) / den);
        if (dis_out) dis_out[m] = dis;
        if (prob_out) prob_out[m] = prob;
    }
}

static int popcount8(unsigned char x) {
    int c = 0;
    while (x) {
        c += x & 1;
        x >>= 1;
    }
    return c;
}

/* Tracking.cc:1747-1766 */
int orc_bf_match(int nq, const unsigned char *dq, int nt, const unsigned char *dt, double ratio, int *match,
                 int *knn) {
    int accepted = 0;
    for (int q = 0; q < nq; q++) {
        int d0 = -1, i0 = -1, d1 = -1, i1 = -1; /* K-best list, K = 2 */
        for (int t = 0; t < nt; t++) {
            int d = 0;
            for (int k = 0; k < 32; k++) d += popcount8((unsigned char)(dq[32 * q + k] ^ dt[32 * t + k]));
            /* insertion with strict comparisons: an equal distance never displaces an earlier row */
            if (i0 < 0 || d < d0) {
                d1 = d0;
                i1 = i0;
                d0 = d;
                i0 = t;
            } else if (i1 < 0 || d < d1) {
                d1 = d;
                i1 = t;
            }
        }
        int m = -1;
        /* :1755  match.size() == 2 && match[0].distance < match[1].distance * 0.6  (float * double) */
        if (i1 >= 0 && (double)(float)d0 < (double)(float)d1 * ratio) m = i0;
        match[q] = m;
        accepted += m >= 0;
        if (knn) {
            knn[4 * q] = d0;
            knn[4 * q + 1] = i0;
            knn[4 * q + 2] = d1;
            knn[4 * q + 3] = i1;
        }
    }
    return accepted;
}

/* Tracking.cc:1945-1955: the loop that acts on every point labelled moving.  One problem; `dyn` receives, in loop
 * order, the element the loop body touches for each res_label[i] == 0 (fid = featureMapAssos[i].fid, or i itself
 * when fid is NULL), `stat` the points the loop leaves alone.  Returns the number of moving points. */
int orc_label_partition(int N, const short *res_label, const int *fid, int *dyn, int *stat) {
    int nd = 0, ns = 0;
    for (int i = 0; i < N; i++) {          /* :1946 */
        const int v = fid ? fid[i] : i;    /* :1948 */
        if (res_label[i] == 0) dyn[nd++] = v; /* :1949-1954 */
        else stat[ns++] = v;
    }
    return nd;
}

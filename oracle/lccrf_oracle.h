/*
 * lccrf_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's algorithm for the LC-CRF hot path
 * (SURVEY.md section 8a).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product path
 * (lc-crf-slam_b200/csrc) never links, loads or calls anything in oracle/.
 *
 * Parity status:
 *   - lattice / filter / mean-field / MAP: PINNED.  Checked bit-for-bit against the
 *     reference headers compiled in place (oracle/_ref/libref.so, built by
 *     oracle/Makefile from /root/reference/Thirdparty/DenseCRF/include) and against
 *     the reference's one golden vector (examples/res1_cpu.ppm, committed as
 *     tests/golden/golden_im1.npz by tests/golden/make_golden.py).
 *   - long-term unary (Tracking.cc:1803-1839, 1961-2013): Tracking.cc needs OpenCV
 *     C++/Eigen/Pangolin/g2o and cannot be compiled here, and the reference has no test or
 *     fixture for it.  PINNED TO THE REFERENCE'S THIRD-PARTY ARITHMETIC instead: the
 *     committed fixture tests/golden/golden_unary.npz (tests/golden/make_golden_unary.py)
 *     evaluates :1803-1839 statement by statement with Rcw*x3Dw+tcw computed by the real
 *     cv::gemm (cv2 4.13) and every scalar in its declared C++ type; this restatement
 *     reproduces it bit for bit (tests/test_oracle.py::test_unary_vs_opencv_golden).
 *     UNPINNED remains the summation order: the reference walks a
 *     std::map<KeyFrame*,size_t> (pointer order, not reproducible); the restatement fixes
 *     observation order = CSR order.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Arithmetic is IEEE fp32 with each operation individually rounded
 * (compile with -ffp-contract=off, no -march, no -ffast-math), mirroring the
 * reference's effective build (CMakeLists.txt:11-12,20: -O3, -march=native commented out).
 */
#ifndef LCCRF_ORACLE_H
#define LCCRF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- permutohedral lattice (permutohedral_cpu.h:173-761) ---- */
typedef struct orc_lattice {
    int N, d, V;
    int *offset;  /* [N*(d+1)]   vertex id per (point, remainder)      permutohedral_cpu.h:375 */
    float *bary;  /* [N*(d+1)]   barycentric weight                    permutohedral_cpu.h:376 */
    int *nbr;     /* [(d+1)*V*2] {n1,n2} per [axis j][vertex i], -1 = absent   :418-419 */
    short *keys;  /* [V*d]       lattice key of each vertex (first d coords)  :144-145 */
} orc_lattice;

orc_lattice *orc_lattice_init(const float *feature, int d, int N);
void orc_lattice_free(orc_lattice *l);
/* splat / blur / slice, SSE-overload semantics (permutohedral_cpu.h:634-699). out may alias in. */
void orc_lattice_filter(const orc_lattice *l, float *out, const float *in, int L);

/* ---- Potts potential (pairwise3d.h:20-28,73-78 == pairwise_cpu.h:15-23,52-57) ---- */
void orc_potts_norm(const orc_lattice *l, float *norm /*[N]*/);
void orc_potts_apply(const orc_lattice *l, const float *norm, float w, float *out, const float *in,
                     float *tmp, int L);

/* ---- DenseCRF driver (densecrf_base.h:65-91, densecrf3d.h:51-158) ---- */
float orc_fast_exp(float x);
void orc_exp_and_normalize(float *out, const float *in, int N, int L, float scale, float relax);
void orc_unary_from_label(float *unary, const short *label, int N, int L, float u_energy,
                          const float *n_energies, const float *p_energies);
void orc_build_map(short *map, const float *Q, int N, int L);
/* inference(iters, with_map, relax): Q [N*L] out, map [N] out (may be NULL) */
void orc_meanfield(int N, int L, const float *unary, int K, const orc_lattice *const *lat,
                   const float *const *norm, const float *w, int iters, float relax, float *Q,
                   short *map);

/* ---- feature assembly (pairwise3d.h:38-71, pairwise_cpu.h:34-50) ---- */
void orc_features_div2(float *feat /*[N*2]*/, const float *a, const float *b, int N, float sa,
                       float sb, int stride_a, int stride_b);
void orc_features_image(float *feat /*[W*H*F]*/, int W, int H, int F, float posdev,
                        const unsigned char *img_u8, const float *img_f32, float featuredev);

/* ---- long-term unary (src/Tracking.cc:1803-1839) : pinned by tests/golden/golden_unary.npz (cv::gemm), see header ---- */
void orc_map_point_unary(int N, const float *xyz, const int *obs_ptr, const int *obs_kf,
                         const float *obs_uv, const float *kf_pose /*[nKF][12] R|t rows*/,
                         const float *kf_intr /*[nKF][4] fx fy cx cy*/,
                         const float *kf_bounds /*[nKF][4] minx maxx miny maxy*/, float *observs,
                         float *error, float *depth);

typedef struct orc_slam_params {
    float w1, w2;                    /* TUM3.yaml:81-82 */
    float u_alpha, stdev_alpha;      /* reprojection error mean / stdev   Tracking.cc:155-156 */
    float u_beta, stdev_beta;        /* observation count mean / stdev    :158-159 */
    float u_gamma, stdev_gamma;      /* epipolar prior (unused by U2 itself) */
    float point3d_stdev, point2d_stdev;
    float u_depth, pth, confidence;
    int iters;                       /* Tracking.cc:1929 -> 5 */
} orc_slam_params;

/* src/Tracking.cc:1961-2013.  p4 may be NULL (mvFeatureMatchProb.empty()). */
void orc_rough_classify(int N, const float *observs, const float *error, const float *depth,
                        const double *p4, const orc_slam_params *prm, short *label);

/* src/Tracking.cc:1919-1930 on flat arrays: label->unary, two PottsPotential3D<2,2>, inference(iters,true).
 * energies[3] = {u_energy, n_energy, p_energy} for L=2 and one shared confidence. */
void orc_slam_crf(int N, const float *observs, const float *error, const float *kp2d,
                  const short *init_label, const float *energies, const orc_slam_params *prm,
                  float *Q /*[N*2]*/, short *map /*[N]*/, int *V_out /*[2] or NULL*/);

/* ---- frontend feeders of the CRF ---- */
/* Tracking::GetFeature2EpipolarDis, src/Tracking.cc:2030-2047, with
 * FundamentalMatrixEstimator::symmetricEpipolarDistance,
 * Thirdparty/graph-cut-ransac-master/include/fundamental_estimator.h:90-127 (needs Eigen + OpenCV: not
 * compilable here -> PARITY UNPINNED, plain double arithmetic restated line by line, including the reference's
 * use of y1 where y2 is meant at :121).  pt1/pt2 [M*2] float keypoints, F9 row-major.  dis/prob [M]. */
void orc_epipolar_prior(int M, const float *pt1, const float *pt2, const double *F9, float u_gamma,
                        float stdev_gamma, double *dis, double *prob);
/* Tracking::BfMatch, src/Tracking.cc:1747-1766: cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio test.
 * OpenCV (opencv_features2d, version un-pinned by CMakeLists.txt:33-39) is not in the tree; the published
 * algorithm is restated (exhaustive Hamming distances, K-best insertion with strict comparisons => ties keep
 * the lower train index first) and PINNED against cv2 4.13 in this container (tests/golden/golden_frontend.npz,
 * tests/golden/make_golden_frontend.py).  match [nq] (-1 = rejected), knn [nq*4] {d0,i0,d1,i1} (may be NULL).
 * Returns the number of accepted matches. */
int orc_bf_match(int nq, const unsigned char *desc_q, int nt, const unsigned char *desc_t, double ratio,
                 int *match, int *knn);

/* Tracking.cc:1945-1955 label application: ordered lists of moving (label 0) and static points of one problem */
int orc_label_partition(int N, const short *res_label, const int *fid, int *dyn, int *stat);

#ifdef __cplusplus
}
#endif
#endif

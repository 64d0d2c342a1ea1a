// slam_callsite.cpp -- the DenseCRF call sequence of src/Tracking.cc:1919-1930 as a stand-alone
// program.  The SAME source is compiled twice by tests/test_gpu_dropin.py:
//   (a) against the reference headers (/root/reference/Thirdparty/DenseCRF/include, build container only)
//   (b) against the drop-in mirror (lc-crf-slam_b200/densecrf) + liblccrf.so
// and the two outputs are compared.  Input/output are raw binary files.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace std;  // include/Tracking.h:55 does this before the DenseCRF headers are parsed
#include "densecrf3d.h"
#include "pairwise3d.h"

using namespace DenseCRF;

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int N = 0;
    float prm[7];  // conf, w1, w2, observ_stdev, rpj_stdev, p3d_stdev, p2d_stdev
    if (fread(&N, 4, 1, f) != 1 || fread(prm, 4, 7, f) != 7) return 4;
    vector<float> vobservs(N), verrors(N), kp(2 * (size_t)N);
    vector<short> label(N);
    if (N && (fread(vobservs.data(), 4, N, f) != (size_t)N || fread(verrors.data(), 4, N, f) != (size_t)N ||
              fread(kp.data(), 4, 2 * (size_t)N, f) != 2 * (size_t)N || fread(label.data(), 2, N, f) != (size_t)N))
        return 5;
    fclose(f);
    vector<Point3f> vpoints(N);
    vector<Point2f> vcorrd2d(N);
    for (int i = 0; i < N; i++) vcorrd2d[i] = Point2f(kp[2 * i], kp[2 * i + 1]);
    short *init_label = label.data();
    float mConf = prm[0], mW1 = prm[1], mW2 = prm[2], mObservStdev = prm[3], mRpjErrorStdev = prm[4],
          mPoint3dStdev = prm[5], mPoint2dStdev = prm[6];

    // ---- verbatim call sequence, src/Tracking.cc:1919-1930 ----
    const int M = 2;
    DenseCRF3D<M> crf(N);
    crf.setUnaryEnergyFromLabel(init_label, mConf);

    auto *appearancePairwise = PottsPotential3D<M, 2>::appearanceKernel(N, mW1, vobservs, verrors, mObservStdev, mRpjErrorStdev);
    crf.addPairwiseEnergy(appearancePairwise);

    auto *smoothnessPairwise = PottsPotential3D<M, 2>::smoothKernel(N, mW2, vpoints, vcorrd2d, mPoint3dStdev, mPoint2dStdev);
    crf.addPairwiseEnergy(smoothnessPairwise);

    crf.inference(5, true);
    short *res_label = crf.getMap();
    // ------------------------------------------------------------

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 6;
    fwrite(res_label, 2, N, o);
    fwrite(crf.getProbability(), 4, 2 * (size_t)N, o);
    fclose(o);
    return 0;
}

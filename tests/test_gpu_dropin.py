"""GPU tests of the drop-in boundary: the reference's own call sites, compiled against the C++
header mirror (lc-crf-slam_b200/densecrf) + liblccrf.so, versus the same source compiled against
the reference headers (oracle/_ref/slam_callsite_ref, prebuilt in the build container)."""
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

from util import assert_bit_exact, assert_map, assert_marginals

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
synth = importlib.import_module("lc-crf-slam_b200.synth")


def build_dropin(tmp_path, src, name, extra=()):
    exe = tmp_path / name
    cmd = ["g++", "-O2", "-std=c++14", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "lc-crf-slam_b200", "densecrf"), *extra, "-o", str(exe), src,
           "-L" + os.path.join(ROOT, "lc-crf-slam_b200"), "-llccrf", "-Wl,-rpath," + os.path.join(ROOT, "lc-crf-slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(exe)


def write_case(path, fr, lab, prm):
    with open(path, "wb") as f:
        f.write(struct.pack("i", fr.n))
        f.write(np.array([prm.confidence, prm.w1, prm.w2, prm.stdev_beta, prm.stdev_alpha, prm.point3d_stdev,
                          prm.point2d_stdev], dtype=np.float32).tobytes())
        for a in (fr.observs, fr.error, fr.kp2d):
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
        f.write(np.ascontiguousarray(lab, np.int16).tobytes())


def read_result(path, n):
    raw = open(path, "rb").read()
    m = np.frombuffer(raw[:2 * n], dtype=np.int16)
    q = np.frombuffer(raw[2 * n:], dtype=np.float32).reshape(n, 2)
    return m, q


@pytest.mark.parametrize("N", [0, 3001, 20000])
def test_tracking_callsite_dropin(pkg, oracle, tmp_path, N):
    from oracle.pyoracle import slam_params
    prm = pkg.SlamParams.make()
    fr = synth.slam_frame(N, seed=31 + N)
    lab = oracle.rough_classify(fr.observs, fr.error, fr.depth, slam_params(**synth.SLAM_PARAMS))
    case = str(tmp_path / "case.bin")
    write_case(case, fr, lab, prm)
    exe = build_dropin(tmp_path, os.path.join(ROOT, "tests", "cpp", "slam_callsite.cpp"), "slam_callsite_lccrf")
    out = str(tmp_path / "out_lccrf.bin")
    r = subprocess.run([exe, case, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m, q = read_result(out, N)
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "slam_callsite_ref")
    if os.path.exists(ref_exe):  # the very same source against the reference headers
        out_r = str(tmp_path / "out_ref.bin")
        assert subprocess.run([ref_exe, case, out_r]).returncode == 0
        mr, qr = read_result(out_r, N)
    else:
        qr, mr, _ = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab, pkg.label_energies(2, prm.confidence),
                                    slam_params(**synth.SLAM_PARAMS))
    if N:
        assert_bit_exact(q, qr)
        assert np.array_equal(m, mr)


PLUGIN_SRC = r"""
// A user-defined PairwisePotential mixed with a built-in one: the plugin interface of
// densecrf_base.h:12-19 must keep working on host pointers.
#include <cmath>
#include <cstdio>
#include <vector>
using namespace std;
#include "densecrf3d.h"
#include "pairwise3d.h"
using namespace DenseCRF;
struct Bias : PairwisePotential {          // out[i,l] += 0.25 * in[i,l]
    Bias(int N) : PairwisePotential(N) {}
    void apply(float *out, const float *in, float *) const override { for (int k = 0; k < 2 * N_; k++) out[k] += 0.25f * in[k]; }
};
int main(int argc, char **argv) {
    const int N = 2000;
    vector<float> feat(2 * N), unary(2 * N);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (auto &v : feat) v = rnd() * 20;
    for (auto &v : unary) v = rnd() * 2;
    DenseCRF3D<2> crf(N);
    crf.setUnaryEnergy(unary.data());
    crf.addPairwiseEnergy(new PottsPotential3D<2, 2>(feat.data(), N, 4.0f));
#ifdef WITH_PLUGIN
    crf.addPairwiseEnergy(new Bias(N));
#endif
    crf.inference(3, true);
    FILE *o = fopen(argv[1], "wb");
    fwrite(crf.getMap(), 2, N, o);
    fwrite(crf.getProbability(), 4, 2 * N, o);
    fclose(o);
    // stand-alone potential, never attached to a CRF (pairwise3d.h:73-78 on host arrays)
    PottsPotential3D<2, 2> p(feat.data(), N, 2.0f);
    vector<float> out(2 * N, 1.0f), tmp(2 * N);
    p.apply(out.data(), unary.data(), tmp.data());
    double acc = 0; for (float v : out) acc += v;
    printf("%.6f\n", acc);
    return 0;
}
"""


def test_plugin_potential_on_host_pointers(pkg, ctx, oracle, tmp_path):
    src = tmp_path / "plugin.cpp"
    src.write_text(PLUGIN_SRC)
    N = 2000
    res = {}
    for tag, extra in (("plain", ()), ("plugin", ("-DWITH_PLUGIN",))):
        exe = build_dropin(tmp_path, str(src), "plugin_" + tag, extra)
        out = str(tmp_path / ("out_%s.bin" % tag))
        r = subprocess.run([exe, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        res[tag] = read_result(out, N) + (float(r.stdout.strip()),)
    # regenerate the inputs (same LCG) and run the oracle with/without the extra term
    s = 12345
    vals = []
    for _ in range(4 * N):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        vals.append(np.float32((s >> 8) * np.float32(1.0 / 16777216.0)))
    feat = (np.array(vals[:2 * N], np.float32) * np.float32(20)).reshape(N, 2)
    unary = (np.array(vals[2 * N:], np.float32) * np.float32(2)).reshape(N, 2)
    Qo, mo, _ = oracle.meanfield(unary, [feat], [4.0], 3)
    assert_bit_exact(res["plain"][1], Qo)
    assert np.array_equal(res["plain"][0], mo)
    # with the plugin: emulate the host loop with oracle pieces
    lat = oracle.lattice(feat)
    norm = oracle.potts_norm(lat)
    Q = oracle.exp_and_normalize(unary, -1.0)
    for _ in range(3):
        nxt = -unary
        nxt = nxt + (np.float32(4.0) * norm)[:, None] * oracle.filter(lat, Q)
        nxt = nxt + np.float32(0.25) * Q
        Q = oracle.exp_and_normalize(nxt, 1.0)
    assert_marginals(res["plugin"][1], Q)
    assert np.abs(res["plugin"][1] - res["plain"][1]).max() > 1e-3  # the plugin term really took part
    # stand-alone apply
    lat2 = oracle.lattice(feat)
    exp = (1.0 + (np.float32(2.0) * oracle.potts_norm(lat2))[:, None] * oracle.filter(lat2, unary)).astype(np.float64).sum()
    assert abs(res["plain"][2] - exp) <= 1e-4 * abs(exp)
    oracle.lattice_free(lat)
    oracle.lattice_free(lat2)


def test_example_cpu_main_dropin(pkg, tmp_path):
    """examples/example_cpu.cpp's CRF section (lines 80-98), verbatim, against the mirror: golden image KAT."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_im1.npz"))
    W, H = int(g["W"]), int(g["H"])
    src = tmp_path / "example.cpp"
    src.write_text(r"""
#include "densecrf_cpu.h"
#include "pairwise_cpu.h"
#include <cstdio>
#include <vector>
using namespace DenseCRF;
int main(int argc, char **argv) {
    const int M = 21; int W = %d, H = %d; const float GT_PROB = 0.5;
    std::vector<unsigned char> imv(W * H * 3); std::vector<short> labelv(W * H);
    FILE *f = fopen(argv[1], "rb"); fread(imv.data(), 1, imv.size(), f); fread(labelv.data(), 2, labelv.size(), f); fclose(f);
    unsigned char *im = imv.data(); short *label = labelv.data();
    // ---- example_cpu.cpp:80-98 ----
    DenseCRFCPU<M> crf(W * H);
    crf.setUnaryEnergyFromLabel( label, GT_PROB );
    auto* smoothnessPairwise = PottsPotentialCPU<M, 2>::FromImage<>(W, H, 3.0, 3.0);
    crf.addPairwiseEnergy( smoothnessPairwise );
    auto* appearancePairwise = PottsPotentialCPU<M, 5>::FromImage<unsigned char>(W, H, 10.0, 60.0, im, 20.0);
    crf.addPairwiseEnergy( appearancePairwise );
    crf.inference(10, true);
    short * map = crf.getMap();
    // --------------------------------
    FILE *o = fopen(argv[2], "wb"); fwrite(map, 2, W * H, o); fclose(o);
    return 0;
}
""" % (W, H))
    exe = build_dropin(tmp_path, str(src), "example_dropin")
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.ascontiguousarray(g["im"], np.uint8).tobytes())
        f.write(np.ascontiguousarray(g["label"], np.int16).tobytes())
    r = subprocess.run([exe, inp, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    m = np.frombuffer(open(out, "rb").read(), dtype=np.int16)
    assert np.array_equal(m, g["map"]), int((m != g["map"]).sum())


def test_map_mirror_binding_runs(tmp_path):
    """the C++ binding of the resident map (tests/cpp/map_mirror.cpp) end to end on the device: two
    keyframes, eight points, one frame through lccrf_frames_submit_visible; the two drifting points come out moving"""
    exe = tmp_path / "map_mirror"
    cmd = ["g++", "-O1", "-std=c++14", "-I" + os.path.join(ROOT, "include"), "-o", str(exe),
           os.path.join(ROOT, "tests", "cpp", "map_mirror.cpp"), "-L" + os.path.join(ROOT, "lc-crf-slam_b200"), "-llccrf",
           "-Wl,-rpath," + os.path.join(ROOT, "lc-crf-slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    assert "map: 2 keyframes, 8 points, 12 observations" in run.stdout
    labels = [int(x) for x in run.stdout.strip().split("labels:")[1].split()]
    assert len(labels) == 8 and set(labels) <= {0, 1}

"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle
(oracle/liboracle.so) and the committed golden fixtures.  Bit-exact for integer/index data (vertex ids,
neighbour tables) and for barycentric bit patterns.  north_star asks for marginals within 1e-4 relative
and MAP labels identical except at near-ties; the CUDA path keeps the reference's operation order
(ordered segmented splat), so the tests assert the stronger property: marginals BIT-identical, MAP
identical.  Only the init-label classifier (glibc expf vs device exp) is tolerance-based."""
import importlib
import os

import numpy as np
import pytest

from util import assert_bit_exact, assert_map, assert_marginals, bits, rel_err, tie_features

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
synth = importlib.import_module("lc-crf-slam_b200.synth")


def oracle_params():
    from oracle.pyoracle import slam_params
    return slam_params(**synth.SLAM_PARAMS)


# ------------------------------------------------------------------ lattice: bit-exact
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7])
def test_lattice_bit_exact(pkg, ctx, oracle, d):
    rng = np.random.default_rng(100 + d)
    for N in (0, 1, 3, 4, 5, 6, 7, 8, 257, 1000, 5003):
        f = tie_features(rng, N, d)
        lo = oracle.lattice(f)
        lg = pkg.Lattice(ctx, f)
        off, bary, nbr = lg.export()
        assert lg.V == lo["V"], (d, N)
        assert np.array_equal(off, lo["offset"]), (d, N)
        assert np.array_equal(bits(bary), bits(lo["bary"])), (d, N)
        assert np.array_equal(nbr, lo["nbr"]), (d, N)
        oracle.lattice_free(lo)
        lg.close()


def test_lattice_phantom_lanes(pkg, ctx):
    got = []
    for N in (4, 5, 6, 7, 8):
        f = np.array([[100 + 0.01 * i, 100] for i in range(N)], dtype=np.float32)
        lg = pkg.Lattice(ctx, f)
        got.append(lg.V)
        lg.close()
    assert got == [3, 6, 6, 6, 3]


def test_lattice_golden_reference_vectors(pkg, ctx):
    """fixtures produced by the unmodified reference headers (tests/golden/make_golden.py)"""
    g = np.load(os.path.join(GOLD, "golden_ref.npz"))
    for d, N in ((2, 6), (2, 257), (3, 130), (5, 203)):
        k = "lat_d%d_n%d_" % (d, N)
        lg = pkg.Lattice(ctx, g[k + "feat"])
        off, bary, nbr = lg.export()
        assert np.array_equal(off, g[k + "offset"])
        assert np.array_equal(bits(bary), bits(g[k + "bary"]))
        assert np.array_equal(nbr, g[k + "nbr"])
        assert_bit_exact(lg.filter(g[k + "x"]), g[k + "y"], what=k)
        lg.close()


def test_lattice_large_image_shapes(pkg, ctx, oracle):
    """C2-shaped lattices: 640x480 Gaussian (d=2) and bilateral (d=5)."""
    W, H = 640, 480
    img, _ = synth.image_problem(W, H, 11)
    for F, posdev, fd in ((2, 3.0, 0.0), (5, 60.0, 20.0)):
        f = oracle.features_image(W, H, F, posdev, img if F == 5 else None, fd)
        lo = oracle.lattice(f)
        lg = pkg.Lattice(ctx, f)
        off, bary, nbr = lg.export()
        assert lg.V == lo["V"]
        assert np.array_equal(off, lo["offset"]) and np.array_equal(bits(bary), bits(lo["bary"])) and np.array_equal(nbr, lo["nbr"])
        oracle.lattice_free(lo)
        lg.close()


def test_lattice_key_range_error(pkg, ctx):
    f = np.full((16, 2), 3.0e4, dtype=np.float32)  # elevates beyond the reference's short keys
    with pytest.raises(pkg.LccrfError, match="short range"):
        pkg.Lattice(ctx, f)
    lg = pkg.Lattice(ctx, np.zeros((4, 2), np.float32))  # the context stays usable
    assert lg.V == 3
    lg.close()


# ------------------------------------------------------------------ filter
@pytest.mark.parametrize("d,N", [(2, 1000), (3, 777), (5, 2049)])
def test_filter_parity_and_properties(pkg, ctx, oracle, d, N):
    rng = np.random.default_rng(d * N)
    f = tie_features(rng, N, d, 2.0)
    lo = oracle.lattice(f)
    lg = pkg.Lattice(ctx, f)
    for L in (1, 2, 3, 21):
        x = (rng.random((N, L)) * 7 - 2).astype(np.float32)
        yo, yg = oracle.filter(lo, x), lg.filter(x)
        assert_bit_exact(yg, yo, what="filter d=%d N=%d L=%d" % (d, N, L))
        # linearity (size-independent property): filter(2x) == 2*filter(x) exactly (power-of-two scaling)
        assert np.array_equal(bits(lg.filter(2 * x)), bits(2 * yg))
        # determinism: integer accumulation -> run-to-run bit identical
        assert np.array_equal(bits(lg.filter(x)), bits(yg))
    oracle.lattice_free(lo)
    lg.close()


def test_filter_window_arguments(pkg, ctx, oracle):
    """compute(out, in, value_size, in_offset, out_offset, in_size, out_size) (permutohedral_cpu.h:634-637) against the
    compiled reference where it travelled with the snapshot, and against the oracle's zero-pad + crop restatement
    (pinned to the reference by tests/test_oracle.py::test_compute_window_is_zero_padding_plus_crop)"""
    from oracle.pyoracle import Ref
    rng = np.random.default_rng(13)
    N, d = 2203, 3
    f = tie_features(rng, N, d, 2.0)
    lo, lg = oracle.lattice(f), pkg.Lattice(ctx, f)
    r = Ref() if Ref.available() and hasattr(Ref().lib, "ref_lattice_filter_window") else None
    lr = r.lattice(f) if r else None
    for L, (io, oo, isz, osz) in ((1, (0, 0, -1, -1)), (2, (100, 0, 700, -1)), (3, (0, 37, -1, 900)), (21, (500, 1400, 1001, 101)),
                                  (2, (N - 1, 0, 1, 1)), (4, (0, 0, 0, -1))):
        n_in = N - io if isz == -1 else isz
        x = rng.normal(0, 1, (n_in, L)).astype(np.float32)
        got = lg.filter_window(x, L, io, oo, isz, osz)
        full = np.zeros((N, L), np.float32)
        full[io:io + n_in] = x
        n_out = N - oo if osz == -1 else osz
        assert_bit_exact(got, oracle.filter(lo, full)[oo:oo + n_out], what="window %s" % ((L, io, oo, isz, osz),))
        if r:
            assert_bit_exact(got, r.filter_window(lr, x, L, io, oo, isz, osz), what="window vs reference")
    with pytest.raises(pkg.LccrfError, match="window"):
        lg.filter_window(np.zeros((10, 1), np.float32), 1, N - 5, 0, 10, -1)
    if r:
        r.lattice_free(lr)
    oracle.lattice_free(lo)
    lg.close()


def test_filter_large_lattice_vector_blur(pkg, ctx, oracle):
    """Lattices beyond one CTA (N > 32768, one problem) take the element-parallel blur: float4 / float2 / float
    vectors per vertex depending on L.  Every width against the oracle, bit for bit."""
    W, H = 250, 180
    f = oracle.features_image(W, H, 2, 1.5)
    lo = oracle.lattice(f)
    lg = pkg.Lattice(ctx, f)
    assert lg.V == lo["V"]
    rng = np.random.default_rng(99)
    for L in (1, 2, 3, 4, 6, 8, 12):
        x = (rng.random((W * H, L)) * 3 - 1).astype(np.float32)
        for bulk in (1, 0):   # L = 2, 4: operands streamed by cp.async.bulk (k_blur_bulk) / by plain vector loads
            ctx.set_option("bulk_blur", bulk)
            assert_bit_exact(lg.filter(x), oracle.filter(lo, x), what="vector blur L=%d bulk=%d" % (L, bulk))
    ctx.set_option("bulk_blur", 1)
    oracle.lattice_free(lo)
    lg.close()


def test_image_crf_c2_full_size(pkg, ctx, oracle):
    """BASELINE configs[1] (C2) at its full size: 640x480, 2 labels, Gaussian + 5-D bilateral, 10 iterations."""
    W, H = 640, 480
    img, lab = synth.image_problem(W, H, 21)
    en = pkg.label_energies(2, 0.7)
    unary = oracle.unary_from_label(lab, 2, en[0], np.full(2, en[1], np.float32), np.full(2, en[2], np.float32))
    Qo, mo, _ = oracle.meanfield(unary, [oracle.features_image(W, H, 2, 3.0), oracle.features_image(W, H, 5, 60.0, img, 20.0)],
                                 [3.0, 10.0], 10)
    crf = pkg.DenseCRF(ctx, W * H, 2)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.addPairwiseFromImage(W, H, 3.0, 3.0)
    crf.addPairwiseFromImage(W, H, 10.0, 60.0, img, 20.0)
    crf.inference(10, True)
    assert_bit_exact(crf.getProbability(), Qo, what="C2 full-size marginals")
    assert np.array_equal(crf.getMap(), mo)
    crf.close()


# ------------------------------------------------------------------ driver pieces
def test_exp_and_normalize_bit_exact(ctx, oracle):
    rng = np.random.default_rng(3)
    for L in (2, 3, 21):
        x = (rng.normal(0, 8, (4001, L))).astype(np.float32)
        x[::13] *= 4  # exercise the < -20 cut-off and the range-reduction loops
        for scale, relax in ((-1.0, 1.0), (1.0, 1.0), (1.0, 0.5)):
            prev = rng.random((4001, L)).astype(np.float32)
            a = ctx.exp_and_normalize(x, scale, relax, prev)
            b = oracle.exp_and_normalize(x, scale, relax, prev)
            assert np.array_equal(bits(a), bits(b)), (L, scale, relax)


# ------------------------------------------------------------------ SLAM CRF (Tracking.cc:1919-1930)
def run_gpu_slam_crf(pkg, ctx, fr, lab, en, prm, iters=5):
    crf = pkg.DenseCRF(ctx, fr.n, 2)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.addPairwiseEnergy(np.stack([fr.observs / np.float32(prm.stdev_beta), fr.error / np.float32(prm.stdev_alpha)], 1), prm.w1)
    crf.addPairwiseEnergy(fr.kp2d / np.float32(prm.point2d_stdev), prm.w2)
    crf.inference(iters, True)
    out = crf.getProbability(), crf.getMap(), (crf.potts_vertices(0), crf.potts_vertices(1))
    crf.close()
    return out


@pytest.mark.parametrize("N", [0, 1, 2, 5, 2999, 3000, 3001, 3002, 100000])
def test_slam_crf_parity(pkg, ctx, oracle, N):
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    fr = synth.slam_frame(N, seed=N)
    lab = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o)
    Qo, mo, Vo = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab, en, prm_o)
    Q, m, V = run_gpu_slam_crf(pkg, ctx, fr, lab, en, prm)
    assert tuple(Vo) == V
    if N:
        assert_marginals(Q, Qo, what="N=%d" % N)   # the north_star gate
        assert_bit_exact(Q, Qo, what="N=%d" % N)   # what the ordered splat actually delivers
        assert np.array_equal(m, mo)


def test_slam_crf_golden_reference(pkg, ctx):
    """marginals + MAP produced by the unmodified reference on a seeded 1003-point frame"""
    g = np.load(os.path.join(GOLD, "golden_ref.npz"))
    prm = pkg.SlamParams.make()
    fr = synth.SlamFrame(g["slam_observs"], g["slam_error"], np.zeros_like(g["slam_error"]), g["slam_kp2d"], None)
    Q, m, _ = run_gpu_slam_crf(pkg, ctx, fr, g["slam_label"], g["slam_energies"], prm)
    assert_bit_exact(Q, g["slam_Q"], what="golden_ref")
    assert np.array_equal(m, g["slam_map"])


def test_generic_labels_dims_relax_and_stepwise(pkg, ctx, oracle):
    rng = np.random.default_rng(5)
    N = 1500
    # L = 2 with lattices goes through the fused point pass: generic kernel (K = 1, K = 3) and the two-lattice
    # specialisations (D = 3,3 and D = 3,6), each also with relax != 1
    for L, dims in ((3, (2, 3)), (4, (5,)), (21, (2,)), (2, ()), (2, (2,)), (2, (2, 2)), (2, (2, 5)), (2, (3, 2, 4))):
        feats = [tie_features(rng, N, d, 2.0) for d in dims]
        unary = rng.random((N, L)).astype(np.float32) * 3
        if L == 2:
            unary[::40, 1] = unary[::40, 0]  # label ties at the start (strict-< MAP rule, zero softmax arguments)
        w = [3.0 + k for k in range(len(dims))]
        for relax in (1.0, 0.5):
            Qo, mo, _ = oracle.meanfield(unary, feats, w, 4, relax)
            crf = pkg.DenseCRF(ctx, N, L)
            crf.setUnaryEnergy(unary)
            for f, wk in zip(feats, w):
                crf.addPairwiseEnergy(f, wk)
            crf.startInference()            # step-by-step API (densecrf_base.h:78-91)
            for _ in range(4):
                crf.stepInference(relax)
            assert crf.getMap() is None     # map_ does not exist before buildMap (densecrf3d.h:139)
            crf.buildMap()
            assert_bit_exact(crf.getProbability(), Qo, what="L=%d dims=%s relax=%g" % (L, dims, relax))
            assert np.array_equal(crf.getMap(), mo)
            crf.close()


def test_two_label_symmetric_problem_ties_everywhere(pkg, ctx, oracle):
    """Equal energies for both labels at every point: every message is label-symmetric, so both softmax arguments are
    exactly zero at every point and every iteration (the tie path of the fused point pass); MAP must be label 0
    everywhere (strict <, densecrf3d.h:143-148)."""
    rng = np.random.default_rng(12)
    N = 3001
    e = rng.random(N).astype(np.float32) * 2
    unary = np.stack([e, e], axis=1)
    feats = [tie_features(rng, N, 2, 2.0), tie_features(rng, N, 2, 6.0)]
    Qo, mo, _ = oracle.meanfield(unary, feats, [10.0, 30.0], 5)
    crf = pkg.DenseCRF(ctx, N, 2)
    crf.setUnaryEnergy(unary)
    for f, wk in zip(feats, (10.0, 30.0)):
        crf.addPairwiseEnergy(f, wk)
    crf.inference(5, True)
    assert_bit_exact(crf.getProbability(), Qo)
    assert np.array_equal(crf.getProbability(), np.full((N, 2), 0.5, np.float32))
    assert np.array_equal(crf.getMap(), mo) and not crf.getMap().any()
    crf.close()


def test_unary_entry_poke_and_unknown_labels(pkg, ctx, oracle):
    N, L = 600, 3
    rng = np.random.default_rng(8)
    lab = rng.integers(-1, L, N).astype(np.int16)
    en = pkg.label_energies(L, 0.6)
    f = tie_features(rng, N, 2, 2.0)
    unary = oracle.unary_from_label(lab, L, en[0], np.full(L, en[1], np.float32), np.full(L, en[2], np.float32))
    unary[17, 2] = 0.25
    Qo, mo, _ = oracle.meanfield(unary, [f], [4.0], 3)
    crf = pkg.DenseCRF(ctx, N, L)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.SetUnaryEnergtForPositiveNode(17, 2, 0.25)
    crf.addPairwiseEnergy(f, 4.0)
    crf.inference(3, True)
    assert_bit_exact(crf.getProbability(), Qo)
    assert np.array_equal(crf.getMap(), mo)
    crf.close()


def test_plugin_pieces_potts_apply(pkg, ctx, oracle):
    """PottsPotential3D::apply on host arrays (the plugin path of the C++ mirror)."""
    N, L = 900, 2
    rng = np.random.default_rng(12)
    f = tie_features(rng, N, 2, 2.0)
    lo = oracle.lattice(f)
    norm = oracle.potts_norm(lo)
    x = rng.random((N, L)).astype(np.float32)
    out0 = rng.normal(0, 1, (N, L)).astype(np.float32)
    tmp_o = oracle.filter(lo, x)
    exp = out0 + (np.float32(5.0) * norm)[:, None] * tmp_o
    crf = pkg.DenseCRF(ctx, N, L)
    crf.addPairwiseEnergy(f, 5.0)
    out, tmp = crf.potts_apply(0, out0, x)
    assert_bit_exact(tmp, tmp_o)
    assert_bit_exact(out, exp)
    crf.close()
    oracle.lattice_free(lo)


# ------------------------------------------------------------------ golden image (the reference's own KAT)
def test_golden_image_kat(pkg, ctx, oracle):
    g = np.load(os.path.join(GOLD, "golden_im1.npz"))
    W, H, M = int(g["W"]), int(g["H"]), 21
    c = np.float32(0.5)
    en = np.array([-np.log(np.float64(np.float32(1.0) / np.float32(M))), -np.log(np.float64((np.float32(1.0) - c) / np.float32(M - 1))),
                   -np.log(np.float64(c))]).astype(np.float32)
    crf = pkg.DenseCRF(ctx, W * H, M)
    crf.setUnaryEnergyFromLabel(g["label"], energies=en)
    crf.addPairwiseFromImage(W, H, 3.0, 3.0)
    crf.addPairwiseFromImage(W, H, 10.0, 60.0, g["im"], 20.0)
    crf.inference(10, True)
    m, Q = crf.getMap(), crf.getProbability()
    crf.close()
    # oracle marginals (bit-equal to the reference's) to qualify any mismatch as a near-tie
    unary = oracle.unary_from_label(g["label"], M, en[0], np.full(M, en[1], np.float32), np.full(M, en[2], np.float32))
    Qo, mo, _ = oracle.meanfield(unary, [oracle.features_image(W, H, 2, 3.0), oracle.features_image(W, H, 5, 60.0, g["im"], 20.0)],
                                 [3.0, 10.0], 10)
    assert np.array_equal(mo, g["map"])
    assert np.array_equal(m, g["map"]), "%d px differ from res1_cpu.ppm" % int((m != g["map"]).sum())
    assert_bit_exact(Q, Qo, what="golden image marginals")


def test_image_crf_c2_small(pkg, ctx, oracle):
    """C2 recipe (DenseCRFCPU<2>, Gaussian + bilateral, 10 iterations) at 160x120."""
    W, H = 160, 120
    img, lab = synth.image_problem(W, H, 4)
    en = pkg.label_energies(2, 0.7)
    unary = oracle.unary_from_label(lab, 2, en[0], np.full(2, en[1], np.float32), np.full(2, en[2], np.float32))
    Qo, mo, _ = oracle.meanfield(unary, [oracle.features_image(W, H, 2, 3.0), oracle.features_image(W, H, 5, 60.0, img, 20.0)],
                                 [3.0, 10.0], 10)
    crf = pkg.DenseCRF(ctx, W * H, 2)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.addPairwiseFromImage(W, H, 3.0, 3.0)
    crf.addPairwiseFromImage(W, H, 10.0, 60.0, img, 20.0)
    crf.inference(10, True)
    assert_bit_exact(crf.getProbability(), Qo)
    assert np.array_equal(crf.getMap(), mo)
    assert 0 < mo.sum() < W * H
    crf.close()


# ------------------------------------------------------------------ long-term unary
# uniform observation counts >= 16 take the round layout of the unary kernel (70, 100: partial last round; 16: one round;
# 20: not well filled -> chunk layout); ragged or short lists take the chunk layout
@pytest.mark.parametrize("N,obs,ragged", [(1, 1, False), (31, 3, True), (5000, 64, True), (20000, 64, False), (300, 2500, True),
                                          (3000, 70, False), (640, 16, False), (1000, 100, False), (2050, 20, False), (64, 4096, False)])
def test_map_point_unary_bit_exact(ctx, oracle, N, obs, ragged):
    snap = synth.map_snapshot(N, obs, seed=N + obs, ragged=ragged)
    ob, er, de = oracle.map_point_unary(snap)
    gob, ger, gde = ctx.map_point_unary(snap)
    assert np.array_equal(ob, gob)
    assert np.array_equal(bits(er), bits(ger))
    assert np.array_equal(bits(de), bits(gde))


@pytest.mark.parametrize("name", ["ragged", "uniform64", "uniform70", "mixed_cameras"])
def test_map_point_unary_vs_opencv_golden(ctx, name):
    """The CUDA unary against the committed cv2.gemm-based evaluation of Tracking.cc:1803-1839
    (tests/golden/golden_unary.npz, made by tests/golden/make_golden_unary.py): bit-identical."""
    from util import golden_unary_case
    s, ob, er, de = golden_unary_case(name)
    gob, ger, gde = ctx.map_point_unary(s)
    assert np.array_equal(ob, gob)
    assert np.array_equal(bits(er), bits(ger)) and np.array_equal(bits(de), bits(gde))


def test_map_point_unary_mixed_cameras(pkg, ctx, oracle):
    """Keyframes with different intrinsics / image bounds take the per-keyframe path (one shared camera is served
    from kernel parameters); both must match the restatement bit for bit, stand-alone and inside a frames batch."""
    snap = synth.map_snapshot(6000, 24, seed=21, ragged=True)
    snap.kf_intr[::3] = np.array(synth.BONN_INTR, np.float32)
    snap.kf_bounds[1::4] = np.array([8, 600, 4, 470], np.float32)
    ob, er, de = oracle.map_point_unary(snap)
    gob, ger, gde = ctx.map_point_unary(snap)
    assert np.array_equal(ob, gob) and np.array_equal(bits(er), bits(ger)) and np.array_equal(bits(de), bits(gde))
    F = pkg.Frames(ctx, [snap.n])
    F.set_map_inputs(snap.xyz, snap.obs_ptr, snap.obs_kf, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F.run(); F.run()
    d = F.get_debug()
    assert np.array_equal(bits(er), bits(d["error"])) and np.array_equal(bits(de), bits(d["depth"]))
    # back to one camera on the same object: the captured graph must be rebuilt with the new kernel parameters
    snap2 = synth.map_snapshot(6000, 24, seed=22, ragged=True)
    ob2, er2, de2 = oracle.map_point_unary(snap2)
    F.set_map_inputs(snap2.xyz, snap2.obs_ptr, snap2.obs_kf, snap2.obs_uv, snap2.kf_pose, snap2.kf_intr, snap2.kf_bounds, snap2.kp2d)
    F.run(); F.run()
    d = F.get_debug()
    assert np.array_equal(bits(er2), bits(d["error"])) and np.array_equal(bits(de2), bits(d["depth"]))
    F.close()


def classify_sum(fr_observs, fr_error, fr_depth, prm):
    """p1 + p2 + p3 of Tracking.cc:1972-1975 in double (the likelihood sum the threshold is compared with)"""
    f = np.float32
    k1 = (fr_observs - f(prm.u_beta)) ** 2 / (f(2) * f(prm.stdev_beta) * f(prm.stdev_beta))
    k2 = (fr_error - f(prm.u_alpha)) ** 2 / (f(2) * f(prm.stdev_alpha) * f(prm.stdev_alpha))
    k3 = (fr_depth - f(prm.u_depth)) ** 2 / (f(2) * f(prm.point3d_stdev) * f(prm.point3d_stdev))
    return np.exp(-k1.astype(np.float64)) + np.exp(-k2.astype(np.float64)) + np.exp(-k3.astype(np.float64))


def test_rough_classify(pkg, ctx, oracle):
    """Init labels against the oracle (glibc expf on the host, exp in double rounded once on the device): every label
    that differs is REPORTED with the distance of its likelihood sum from the threshold and must be a threshold tie
    (|p1+p2+p3(+p4) - threshold| below two float ulps of the sum), i.e. a different last bit of one exp()."""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    worst, flips = 0.0, 0
    for seed in (21, 22, 23, 24):
        fr = synth.slam_frame(50000, seed)
        for p4 in (None, np.random.default_rng(seed).random(50000)):
            lo = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o, p4)
            lg = ctx.rough_classify(fr.observs, fr.error, fr.depth, prm, p4)
            assert 0 < lo.sum() < lo.size
            bad = np.nonzero(lo != lg)[0]
            if bad.size:
                ssum = classify_sum(fr.observs[bad], fr.error[bad], fr.depth[bad], prm)
                thr = float(np.float32(prm.pth)) if p4 is None else float(np.float32(prm.pth)) + 0.2
                dist = np.abs(ssum + (0 if p4 is None else p4[bad]) - thr)
                print("RroughClassify: %d of 50000 labels differ (seed %d, p4 %s): |sum - threshold| = %s" % (
                    bad.size, seed, "no" if p4 is None else "yes", ", ".join("%.2e" % d for d in dist)))
                assert (dist <= 2 * np.spacing(np.float32(thr))).all(), dist
                worst, flips = max(worst, float(dist.max())), flips + int(bad.size)
    print("RroughClassify: %d label differences in 400000 points, largest distance from the threshold %.2e" % (flips, worst))
    assert flips <= 8


@pytest.mark.parametrize("case", ["positive", "ties", "mixed_sign", "sparse_zero", "wild"])
def test_filter_long_rows_bit_exact(pkg, ctx, oracle, case):
    """Rows of 10^5 entries (all points share a few lattice vertices): the speculative parallel scan
    (chunk composites, crossing windows, fallbacks) must reproduce the sequential fp32 sums bit for bit --
    including round-half-even ties, zeros, sign changes and magnitude jumps."""
    rng = np.random.default_rng({"positive": 1, "ties": 2, "mixed_sign": 3, "sparse_zero": 4, "wild": 5}[case])
    N = 150001
    f = np.zeros((N, 2), np.float32)
    f[:] = rng.normal(0, 0.02, (N, 2))            # one simplex: 3 rows of ~N entries
    f[::5000] += rng.normal(0, 5, (N // 5000 + 1, 2)).astype(np.float32)   # plus a few short rows
    far = rng.random(N) < 0.3
    f[far] += np.float32(7.0)                      # a second cluster: rows of ~45k entries
    lo, lg = oracle.lattice(f), pkg.Lattice(ctx, f)
    assert lo["V"] == lg.V
    for L in (1, 2, 3):
        if case == "positive":
            x = rng.random((N, L)).astype(np.float32)
        elif case == "ties":      # multiples of 2^-10: exact ties whenever the running sum's ulp exceeds 2^-10
            x = (rng.integers(0, 2048, (N, L)) / 1024.0).astype(np.float32)
        elif case == "mixed_sign":
            x = rng.normal(0, 1, (N, L)).astype(np.float32)
        elif case == "sparse_zero":
            x = (rng.random((N, L)) * (rng.random((N, L)) < 0.01)).astype(np.float32)
            x[: N // 3] = 0
        else:                     # magnitudes from 1e-30 to 1e+20 with zeros and negatives in between
            x = (10.0 ** rng.uniform(-30, 20, (N, L)) * rng.choice([-1.0, 0.0, 1.0, 1.0], (N, L))).astype(np.float32)
        yo, yg = oracle.filter(lo, x), lg.filter(x)
        assert_bit_exact(yg, yo, what="%s L=%d" % (case, L))
    oracle.lattice_free(lo)
    lg.close()


# ------------------------------------------------------------------ batched frames (C4) and the C3 pipeline
def test_frames_batch_parity(pkg, ctx, oracle):
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    rng = np.random.default_rng(42)
    sizes = [0, 1, 4001, 4002, 4003, 4004] + rng.integers(4000, 6001, 10).tolist()
    frames = [synth.slam_frame(n, seed=1000 + i) for i, n in enumerate(sizes)]
    F = pkg.Frames(ctx, sizes, prm, en)
    cat = lambda k: np.concatenate([getattr(f, k) for f in frames])
    F.set_inputs(cat("observs"), cat("error"), cat("depth"), cat("kp2d"))
    for graphs in (0, 1, 1):  # plain launches, capture, replay
        ctx.set_option("graphs", graphs)
        F.run()
        mp, pr = F.get_outputs()
        dbg = F.get_debug()
        o = 0
        for b, fr in enumerate(frames):
            lab = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o)
            assert np.array_equal(lab, dbg["init_label"][o:o + fr.n])
            Qo, mo, Vo = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab, en, prm_o)
            assert tuple(Vo) == tuple(dbg["V"][b]), b
            if fr.n:
                assert_bit_exact(pr[o:o + fr.n], Qo, what="problem %d" % b)
                assert np.array_equal(mp[o:o + fr.n], mo)
            o += fr.n
    ab = F.algorithmic_bytes()
    assert ab["total"] > 0 and ab["per_iteration"] > 0
    F.close()


def test_frames_label_partition(pkg, ctx, oracle):
    """Label application (Tracking.cc:1945-1955): the device's stable partition of the MAP labels lists, per problem
    and in point order, exactly the elements the reference loop acts on -- with feature ids and with local indices;
    empty problems (leading, inner, trailing) and tile-crossing problems included."""
    prm = pkg.SlamParams.make()
    sizes = [0, 0, 1, 3000, 0, 2047, 2048, 2049, 5000, 1, 0, 0]
    frames = [synth.slam_frame(n, seed=300 + i) for i, n in enumerate(sizes)]
    F = pkg.Frames(ctx, sizes, prm)
    with pytest.raises(pkg.LccrfError):
        F.partition()  # before a run
    cat = lambda k: np.concatenate([getattr(f, k) for f in frames])
    F.set_inputs(cat("observs"), cat("error"), cat("depth"), cat("kp2d"))
    F.run()
    mp, _ = F.get_outputs()
    assert 0 < int((mp == 0).sum()) < mp.size  # both classes present
    rng = np.random.default_rng(3)
    fid = np.concatenate([rng.permutation(4 * n)[:n] for n in sizes]).astype(np.int32)
    for f in (None, fid):
        dp, dl, sp, sl = F.partition(f)
        assert dp[0] == 0 and sp[0] == 0 and dp[-1] + sp[-1] == mp.size
        o = 0
        for b, n in enumerate(sizes):
            d_ref, s_ref = oracle.label_partition(mp[o:o + n], None if f is None else f[o:o + n])
            assert np.array_equal(dl[dp[b]:dp[b + 1]], d_ref), b
            assert np.array_equal(sl[sp[b]:sp[b + 1]], s_ref), b
            o += n
    F.close()
    # a batch of nothing
    F0 = pkg.Frames(ctx, [0, 0], prm)
    F0.set_inputs(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 2), np.float32))
    F0.run()
    dp, dl, sp, sl = F0.partition()
    assert dp.tolist() == [0, 0, 0] and sp.tolist() == [0, 0, 0] and dl.size == 0 and sl.size == 0
    F0.close()


def test_frames_from_map_snapshot_c3_small(pkg, ctx, oracle):
    """C3 pipeline at reduced size: unary from the map snapshot -> classify -> CRF, all on the device."""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    snap = synth.map_snapshot(20000, 64, seed=5)
    F = pkg.Frames(ctx, [snap.n], prm, en)
    F.set_map_inputs(snap.xyz, snap.obs_ptr, snap.obs_kf, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F.run()
    F.run()
    mp, pr = F.get_outputs()
    dbg = F.get_debug()
    ob, er, de = oracle.map_point_unary(snap)
    assert np.array_equal(bits(er), bits(dbg["error"])) and np.array_equal(bits(de), bits(dbg["depth"]))
    lab = oracle.rough_classify(ob, er, de, prm_o)
    assert (lab != dbg["init_label"]).sum() <= 1
    Qo, mo, Vo = oracle.slam_crf(ob, er, snap.kp2d, dbg["init_label"], en, prm_o)
    assert tuple(Vo) == tuple(dbg["V"][0])
    assert_bit_exact(pr, Qo)
    assert np.array_equal(mp, mo)
    assert 0 < mo.sum() < snap.n
    F.close()


def test_launch_structure_options_do_not_change_a_bit(pkg, ctx):
    """graphs / concurrent / split_splat / fused only change HOW the launch sequence runs (graph replay, the two pairwise
    kernels on two branches, the short-row splat beside the long-row scan kernels, the fused point pass): marginals and MAP
    of a C3-shaped batch (long rows in the appearance lattice) must be bit-identical in every combination."""
    snaps = [synth.map_snapshot(30000, 64, seed=40 + i) for i in range(2)]
    cat = pkg.concat_frames(snaps)
    ref = None
    try:
        for graphs, concurrent, split, fused in [(1, 1, 1, 1), (0, 1, 1, 1), (1, 0, 1, 1), (1, 1, 0, 1), (0, 0, 0, 0), (1, 1, 1, 0)]:
            for k, v in (("graphs", graphs), ("concurrent", concurrent), ("split_splat", split), ("fused", fused)):
                ctx.set_option(k, v)
            F = pkg.Frames(ctx, [s.n for s in snaps])
            F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"],
                             cat["kf_bounds"], cat["kp2d"], cat["kf_ptr"])
            for _ in range(3):  # plain launches, capture, replay
                F.run()
            mp, pr = F.get_outputs()
            F.close()
            if ref is None:
                ref = (mp.copy(), pr.copy())
                assert 0 < mp.sum() < mp.size
            else:
                assert np.array_equal(mp, ref[0]) and np.array_equal(bits(pr), bits(ref[1])), (graphs, concurrent, split, fused)
    finally:
        for k in ("graphs", "concurrent", "split_splat", "fused"):
            ctx.set_option(k, 1)


def test_frames_rejects_points_without_observations(pkg, ctx):
    snap = synth.map_snapshot(64, 4, seed=1)
    ptr = snap.obs_ptr.copy()
    ptr[10] = ptr[9]  # point 9 has no observation: Tracking.cc:1858 drops it before the CRF
    F = pkg.Frames(ctx, [snap.n])
    with pytest.raises(pkg.LccrfError, match="1858"):
        F.set_map_inputs(snap.xyz, ptr, snap.obs_kf, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F.close()


def test_full_size_c3_against_oracle(pkg, ctx, oracle):
    """BASELINE configs[2] at its full size (N = 100k map points x 64 observations, 6.4 M observations): the whole
    pipeline -- unary from the map snapshot, classification, both lattices, 5 iterations, MAP -- against the oracle,
    bit for bit; then the same problem inside a batch of 3 (problems share no state)."""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    snap = synth.map_snapshot(100000, 64, seed=77)
    F = pkg.Frames(ctx, [snap.n], prm, en)
    F.set_map_inputs(snap.xyz, snap.obs_ptr, snap.obs_kf, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F.run()
    m1, p1 = F.get_outputs()
    F.run()
    m2, p2 = F.get_outputs()
    assert np.array_equal(bits(p1), bits(p2)) and np.array_equal(m1, m2)   # deterministic
    d1 = F.get_debug()
    F.close()
    ob, er, de = oracle.map_point_unary(snap)
    assert np.array_equal(ob, d1["observs"])
    assert np.array_equal(bits(er), bits(d1["error"])) and np.array_equal(bits(de), bits(d1["depth"]))
    lab = oracle.rough_classify(ob, er, de, prm_o)
    assert (lab != d1["init_label"]).sum() <= 2   # threshold ties of exp(), see test_rough_classify
    Qo, mo, Vo = oracle.slam_crf(ob, er, snap.kp2d, d1["init_label"], en, prm_o)
    assert tuple(Vo) == tuple(d1["V"][0])
    assert_bit_exact(p1, Qo, what="C3 full size")
    assert np.array_equal(m1, mo) and 0 < mo.sum() < snap.n
    # the same problem inside a batch of 3 gives bit-identical results
    fr = synth.slam_frame(3000, 5)
    F3 = pkg.Frames(ctx, [fr.n, snap.n, fr.n], prm)
    cat = lambda a, b: np.concatenate([a, b, a])
    F3.set_inputs(cat(fr.observs, d1["observs"]), cat(fr.error, d1["error"]), cat(fr.depth, d1["depth"]), cat(fr.kp2d, snap.kp2d))
    F3.run()
    m3, p3 = F3.get_outputs()
    assert np.array_equal(bits(p3[fr.n:fr.n + snap.n]), bits(p1))
    assert np.array_equal(bits(p3[:fr.n]), bits(p3[fr.n + snap.n:]))
    F3.close()


def test_frames_c4_full_size_1024_problems(pkg, ctx, oracle):
    """BASELINE configs[3] at its stated size: 1024 independent frame CRFs of N ~ U[4000, 6000] in one launch sequence;
    first, last and 32 sampled problems against the oracle bit for bit, every problem normalised and deterministic."""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    rng = np.random.default_rng(1000)
    sizes = rng.integers(4000, 6001, 1024)
    frames = [synth.slam_frame(int(n), seed=5000 + i, dyn_frac=float(rng.uniform(0.15, 0.3))) for i, n in enumerate(sizes)]
    F = pkg.Frames(ctx, sizes.tolist(), prm, en)
    cat = lambda k: np.concatenate([getattr(f, k) for f in frames])
    F.set_inputs(cat("observs"), cat("error"), cat("depth"), cat("kp2d"))
    F.run()
    F.run()
    mp, pr = F.get_outputs()
    dbg = F.get_debug()
    F.run()
    mp2, pr2 = F.get_outputs()
    assert np.array_equal(bits(pr), bits(pr2)) and np.array_equal(mp, mp2)
    assert np.abs(pr.sum(1) - 1).max() < 1e-6
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    check = sorted(set([0, 1023] + rng.choice(1024, 32, replace=False).tolist()))
    for b in check:
        fr, a, z = frames[b], int(ptr[b]), int(ptr[b + 1])
        lab = dbg["init_label"][a:z]
        assert (oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o) != lab).sum() <= 1
        Qo, mo, Vo = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab, en, prm_o)
        assert tuple(Vo) == tuple(dbg["V"][b]), b
        assert_bit_exact(pr[a:z], Qo, what="C4 problem %d" % b)
        assert np.array_equal(mp[a:z], mo)
    F.close()


def test_frames_map_batch_keyframe_slices(pkg, ctx, oracle):
    """Batched C3-style problems with per-problem keyframe slices (kf_ptr): shared-memory slice path,
    global-table path and the oracle must agree bit for bit."""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    snaps = [synth.map_snapshot(n, o, seed=90 + i, n_kf=300, ragged=r) for i, (n, o, r) in
             enumerate(((2500, 64, False), (130, 7, True), (4000, 20, True), (31, 64, False)))]
    sys_path_bench = importlib.import_module("bench")
    cat = sys_path_bench.concat_snapshots(snaps)
    assert cat["kf_pose"].shape[0] == 1200  # > 384: the whole table does not fit shared memory
    F = pkg.Frames(ctx, [s.n for s in snaps], prm, en)
    outs = []
    for kf_ptr in (cat["kf_ptr"], None):
        F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"],
                         cat["kf_bounds"], cat["kp2d"], kf_ptr)
        F.run()
        F.run()
        outs.append((F.get_outputs(), F.get_debug()))
    (m0, p0), d0 = outs[0]
    (m1, p1), d1 = outs[1]
    assert np.array_equal(bits(p0), bits(p1)) and np.array_equal(m0, m1)
    o = 0
    for s in snaps:
        ob, er, de = oracle.map_point_unary(s)
        assert np.array_equal(bits(er), bits(d0["error"][o:o + s.n])) and np.array_equal(bits(de), bits(d0["depth"][o:o + s.n]))
        Qo, mo, _ = oracle.slam_crf(ob, er, s.kp2d, d0["init_label"][o:o + s.n], en, prm_o)
        assert_bit_exact(p0[o:o + s.n], Qo)
        assert np.array_equal(m0[o:o + s.n], mo)
        o += s.n
    F.close()


def test_frames_pipelined_submit_matches_run(pkg, ctx, oracle):
    """lccrf_frames_submit_map / submit / wait (two input slots, copy stream, uint16 keyframe indices) deliver the
    bits of set_inputs + run + get_outputs, for alternating inputs in both slots."""
    prm = pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    snapA = synth.map_snapshot(6001, 24, seed=11, ragged=True)
    snapB = synth.map_snapshot(6001, 24, seed=12, ragged=True)
    F = pkg.Frames(ctx, [snapA.n], prm, en)
    want = []
    for s in (snapA, snapB):
        F.set_map_inputs(s.xyz, s.obs_ptr, s.obs_kf, s.obs_uv, s.kf_pose, s.kf_intr, s.kf_bounds, s.kp2d)
        F.run()
        want.append(F.get_outputs())
    assert not np.array_equal(want[0][0], want[1][0])
    outs = [(np.empty(snapA.n, np.int16), np.empty((snapA.n, 2), np.float32)) for _ in range(2)]
    seq = [(snapA, np.int32), (snapB, np.uint16), (snapB, np.int32), (snapA, np.uint16), (snapA, np.int32)]
    for i, (s, dt) in enumerate(seq):
        slot = i & 1
        if i >= 2:
            F.wait(slot)
            ps, _ = seq[i - 2]
            w = want[0] if ps is snapA else want[1]
            assert np.array_equal(outs[slot][0], w[0]) and np.array_equal(bits(outs[slot][1]), bits(w[1])), i
        kf = np.ascontiguousarray(s.obs_kf.astype(dt))
        F.submit_map(slot, s.xyz, s.obs_ptr, kf, s.obs_uv, s.kf_pose, s.kf_intr, s.kf_bounds, s.kp2d, None, *outs[slot])
        F._keep_kf = getattr(F, "_keep_kf", []) + [kf]
    F.wait(0)
    F.wait(1)
    with pytest.raises(pkg.LccrfError):
        F.wait(2)
    # direct-vector variant
    fr = synth.slam_frame(snapA.n, seed=3)
    F.set_inputs(fr.observs, fr.error, fr.depth, fr.kp2d)
    F.run()
    wm, wp = F.get_outputs()
    F.submit(1, fr.observs, fr.error, fr.depth, fr.kp2d, *outs[1])
    F.wait(1)
    assert np.array_equal(outs[1][0], wm) and np.array_equal(bits(outs[1][1]), bits(wp))
    F.close()

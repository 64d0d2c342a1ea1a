"""GPU tests (-m gpu) of the tolerance mode of the splat: lccrf_ctx_set_option("ordered_splat", 0) replaces the
point-ordered row sums (bit-identical to permutohedral_cpu.h:653-661) by a fixed-shape tree reduction per row.  The
lattice structure stays bit-exact; marginals are gated at north_star's 1e-4 relative and MAP labels at its near-tie rule
(|dQ| < 1e-5) against the oracle; the measured maxima are printed (run with -s) and asserted well inside the gates."""
import importlib

import numpy as np
import pytest

from util import NEAR_TIE, REL_TOL, assert_map, assert_marginals, bits, rel_err, tie_features

pytestmark = pytest.mark.gpu
synth = importlib.import_module("lc-crf-slam_b200.synth")


@pytest.fixture()
def tree_ctx(ctx):
    ctx.set_option("ordered_splat", 0)
    yield ctx
    ctx.set_option("ordered_splat", 1)


def oracle_params():
    from oracle.pyoracle import slam_params
    return slam_params(**synth.SLAM_PARAMS)


@pytest.mark.parametrize("d,N", [(2, 1000), (3, 777), (5, 2049), (2, 40000), (2, 2048 * 3 + 1)])
def test_tree_filter_against_oracle(pkg, tree_ctx, oracle, d, N):
    ctx = tree_ctx
    rng = np.random.default_rng(d * N + 1)
    f = tie_features(rng, N, d, 2.0 if N < 5000 else 6.0)
    lo = oracle.lattice(f)
    lg = pkg.Lattice(ctx, f)
    assert lg.V == lo["V"]
    for L in (1, 2, 3, 4, 21):
        x = rng.random((N, L)).astype(np.float32)
        yo, yg = oracle.filter(lo, x), lg.filter(x)
        r = rel_err(yg, yo).max()
        assert r < 2e-5, "d=%d N=%d L=%d: rel err %.3e" % (d, N, L, r)   # positive inputs: no cancellation
        assert np.array_equal(bits(lg.filter(x)), bits(yg))               # deterministic
        # signed inputs: the error is relative to the magnitude of the terms, not of the (cancelling) sum
        xs = (rng.normal(0, 1, (N, L))).astype(np.float32)
        yo, yg = oracle.filter(lo, xs), lg.filter(xs)
        ya = oracle.filter(lo, np.abs(xs))
        assert (np.abs(yg.astype(np.float64) - yo) <= 2e-5 * np.maximum(ya, 1e-30)).all()
    ctx.set_option("ordered_splat", 1)
    x = rng.random((N, 2)).astype(np.float32)
    assert np.array_equal(bits(lg.filter(x)), bits(oracle.filter(lo, x)))  # the option switches per call
    oracle.lattice_free(lo)
    lg.close()


@pytest.mark.parametrize("case", ["positive", "sparse_zero", "tiny_rows"])
def test_tree_filter_long_and_tiny_rows(pkg, tree_ctx, oracle, case):
    """rows of 10^5 entries spanning ~50 tiles next to short rows / lattices where almost every vertex has one entry"""
    ctx = tree_ctx
    rng = np.random.default_rng(11)
    if case == "tiny_rows":
        N = 30011
        f = rng.uniform(0, 4000, (N, 2)).astype(np.float32)   # ~N*D distinct vertices
    else:
        N = 150001
        f = rng.normal(0, 0.02, (N, 2)).astype(np.float32)
        f[::5000] += rng.normal(0, 5, (N // 5000 + 1, 2)).astype(np.float32)
        f[rng.random(N) < 0.3] += np.float32(7.0)
    lo, lg = oracle.lattice(f), pkg.Lattice(ctx, f)
    assert lo["V"] == lg.V
    worst = 0.0
    for L in (1, 2, 3):
        x = rng.random((N, L)).astype(np.float32)
        if case == "sparse_zero":
            x *= rng.random((N, L)) < 0.01
            x[: N // 3] = 0
        yo, yg = oracle.filter(lo, x), lg.filter(x)
        r = rel_err(yg, yo)
        worst = max(worst, float(r.max()))
        assert np.isfinite(yg).all() and r.max() < REL_TOL, "%s L=%d: %.3e" % (case, L, r.max())
    print("tree filter %s: max rel err vs sequential fp32 %.3e" % (case, worst))
    oracle.lattice_free(lo)
    lg.close()


def run_gpu_slam_crf(pkg, ctx, fr, lab, en, prm, iters=5):
    crf = pkg.DenseCRF(ctx, fr.n, 2)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.addPairwiseEnergy(np.stack([fr.observs / np.float32(prm.stdev_beta), fr.error / np.float32(prm.stdev_alpha)], 1), prm.w1)
    crf.addPairwiseEnergy(fr.kp2d / np.float32(prm.point2d_stdev), prm.w2)
    crf.inference(iters, True)
    out = crf.getProbability(), crf.getMap(), (crf.potts_vertices(0), crf.potts_vertices(1))
    crf.close()
    return out


@pytest.mark.parametrize("N", [1, 2, 5, 2999, 3000, 3001, 3002, 5000, 100000])
def test_tree_slam_crf_gates(pkg, tree_ctx, oracle, N):
    """every BASELINE point-set shape: marginals <= 1e-4 relative, MAP identical except near-ties"""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    worst, flips = 0.0, 0
    for seed in range(3):
        fr = synth.slam_frame(N, seed=N + 17 * seed)
        lab = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o)
        Qo, mo, Vo = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab, en, prm_o)
        Q, m, V = run_gpu_slam_crf(pkg, tree_ctx, fr, lab, en, prm)
        assert tuple(Vo) == V                                   # lattice structure stays bit-exact
        worst = max(worst, assert_marginals(Q, Qo, what="tree N=%d" % N))
        flips += assert_map(m, mo, Qo, what="tree N=%d" % N)
    print("tree splat N=%d: max rel marginal err %.3e (gate %.0e), MAP near-tie flips %d" % (N, worst, REL_TOL, flips))


def test_tree_frames_batch(pkg, tree_ctx, oracle):
    """the batched engine (C4 shape, graphs on) in tolerance mode: the 1e-4 gate and the near-tie rule hold"""
    ctx = tree_ctx
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    rng = np.random.default_rng(4)
    sizes = [0, 1, 4001, 4002] + rng.integers(4000, 6001, 8).tolist()
    frames = [synth.slam_frame(n, seed=2000 + i) for i, n in enumerate(sizes)]
    F = pkg.Frames(ctx, sizes, prm, en)
    cat = lambda k: np.concatenate([getattr(f, k) for f in frames])
    F.set_inputs(cat("observs"), cat("error"), cat("depth"), cat("kp2d"))
    for _ in range(3):
        F.run()
    mp, pr = F.get_outputs()
    o, worst = 0, 0.0
    for fr in frames:
        if fr.n:
            lab = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o)
            Qo, mo, _ = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab, en, prm_o)
            worst = max(worst, assert_marginals(pr[o:o + fr.n], Qo))
            assert_map(mp[o:o + fr.n], mo, Qo)
        o += fr.n
    print("tree splat, batched frames: max rel marginal err %.3e" % worst)
    F.close()


def filter_f64(lat, x):
    """splat / blur / slice of permutohedral_cpu.h:634-699 in double precision on the oracle's lattice arrays"""
    off, bary, nbr, V, d = lat["offset"], lat["bary"].astype(np.float64), lat["nbr"], lat["V"], lat["d"]
    val = np.zeros((V + 1, x.shape[1]))           # row V stands for an absent neighbour (-1)
    for r in range(d + 1):
        np.add.at(val, off[:, r], bary[:, r, None] * x.astype(np.float64))
    for j in range(d + 1):
        n1, n2 = nbr[j, :, 0], nbr[j, :, 1]
        new = val.copy()
        new[:V] = val[:V] + 0.5 * (val[np.where(n1 < 0, V, n1)] + val[np.where(n2 < 0, V, n2)])
        val = new
    alpha = 1.0 / (1.0 + 2.0 ** -d)
    return sum((bary[:, r, None] * alpha) * val[off[:, r]] for r in range(d + 1))


def test_tree_c3_shape_is_outside_the_gate_because_the_reference_rounds(pkg, tree_ctx, oracle):
    """The C3 shape (every point has 64 observations, so the appearance lattice has ~30 vertices with rows of 10^4..10^5
    entries) is where the tolerance mode does NOT meet north_star's 1e-4 against the reference -- and why: the
    reference's own sequential fp32 sum over such a row is 1e-4..1e-3 away from the exact sum, the tree sum 1e-7.
    Matching the reference there means reproducing its rounding, which is what the ordered kernels do (default mode;
    bench.py's headline).  This test pins both facts: filter error against a float64 evaluation, and the resulting
    marginal deviation of the whole pipeline, reported and bounded."""
    ctx = tree_ctx
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    snap = synth.map_snapshot(20000, 64, seed=70)
    ob, er, de = oracle.map_point_unary(snap)
    feat = np.stack([ob / np.float32(prm.stdev_beta), er / np.float32(prm.stdev_alpha)], 1).astype(np.float32)
    lo, lg = oracle.lattice(feat), pkg.Lattice(ctx, feat)
    x = np.random.default_rng(2).random((snap.n, 2)).astype(np.float32)
    truth = filter_f64(lo, x)
    e_ref = float((np.abs(oracle.filter(lo, x) - truth) / np.abs(truth)).max())
    e_tree = float((np.abs(lg.filter(x) - truth) / np.abs(truth)).max())
    print("C3-shaped appearance lattice (V=%d): filter error vs float64: reference order %.2e, tree %.2e" % (lo["V"], e_ref, e_tree))
    assert e_tree < 2e-6 and e_tree < e_ref
    oracle.lattice_free(lo)
    lg.close()
    F = pkg.Frames(ctx, [snap.n], prm, en)
    F.set_map_inputs(snap.xyz, snap.obs_ptr, snap.obs_kf, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F.run()
    F.run()
    mp, pr = F.get_outputs()
    lab = F.get_debug()["init_label"]
    Qo, mo, _ = oracle.slam_crf(ob, er, snap.kp2d, lab, en, prm_o)
    dev = float(rel_err(pr, Qo).max())
    diff = np.nonzero(mp != mo)[0]
    gap = np.abs(Qo[diff, 0] - Qo[diff, 1]) if diff.size else np.zeros(0)
    print("C3 pipeline (20000 x 64), tree splat vs reference order: max rel marginal deviation %.2e (gate %.0e); "
          "%d MAP differences, largest |dQ| among them %.2e" % (dev, REL_TOL, diff.size, gap.max() if diff.size else 0.0))
    assert np.isfinite(pr).all() and dev < 5e-2 and (gap < 2e-2).all()
    F.close()


def test_tree_image_crf_c2(pkg, tree_ctx, oracle):
    """C2 at full size (640x480, Gaussian + 5-D bilateral, 10 iterations) in tolerance mode"""
    W, H = 640, 480
    img, lab = synth.image_problem(W, H, 21)
    en = pkg.label_energies(2, 0.7)
    unary = oracle.unary_from_label(lab, 2, en[0], np.full(2, en[1], np.float32), np.full(2, en[2], np.float32))
    Qo, mo, _ = oracle.meanfield(unary, [oracle.features_image(W, H, 2, 3.0), oracle.features_image(W, H, 5, 60.0, img, 20.0)],
                                 [3.0, 10.0], 10)
    crf = pkg.DenseCRF(tree_ctx, W * H, 2)
    crf.setUnaryEnergyFromLabel(lab, energies=en)
    crf.addPairwiseFromImage(W, H, 3.0, 3.0)
    crf.addPairwiseFromImage(W, H, 10.0, 60.0, img, 20.0)
    crf.inference(10, True)
    worst = assert_marginals(crf.getProbability(), Qo, what="C2 tree")
    flips = assert_map(crf.getMap(), mo, Qo, what="C2 tree")
    print("tree splat C2: max rel marginal err %.3e, near-tie flips %d (|dQ| < %.0e)" % (worst, flips, NEAR_TIE))
    crf.close()

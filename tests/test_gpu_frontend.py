"""GPU parity tests (run with -m gpu) for the CRF's per-frame feeders, through the C ABI:
  lccrf_bf_match / _batch   Tracking::BfMatch               src/Tracking.cc:1747-1766   (integer work: bit-exact)
  lccrf_epipolar_prior      Tracking::GetFeature2EpipolarDis src/Tracking.cc:2030-2047   (double: distance bit-exact,
                            likelihood within 2 ulp of glibc exp -- the two libm exp() implementations are each < 1 ulp)
"""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
synth = importlib.import_module("lc-crf-slam_b200.synth")


def test_bf_match_opencv_golden(ctx):
    """knn lists (incl. tie order) and accepted correspondences equal what cv::BFMatcher produced (golden fixture)."""
    g = np.load(os.path.join(GOLD, "golden_frontend.npz"))
    for name in ("orb", "ties", "one_train_row", "two_train_rows", "ragged"):
        match, knn, n = ctx.bf_match(g[name + "_dq"], g[name + "_dt"], 0.6)
        assert np.array_equal(knn, g[name + "_knn"]), name
        assert np.array_equal(match, g[name + "_match"]), name
        assert n == int((g[name + "_match"] >= 0).sum())


@pytest.mark.parametrize("nq,nt,eb", [(1, 1, 32), (1, 5000, 32), (2000, 2000, 32), (129, 255, 2), (3000, 257, 4), (50, 0, 32), (0, 50, 32)])
def test_bf_match_vs_oracle(ctx, oracle, nq, nt, eb):
    """Frame-sized and ragged shapes, low-entropy descriptors (ties everywhere), empty sets; train ranges that split
    across several CTAs must merge back in index order."""
    dq, dt = synth.orb_frame_pair(nq, nt, seed=nq * 7 + nt, entropy_bytes=eb)
    mo, ko, no = oracle.bf_match(dq, dt, 0.6)
    mg, kg, ng = ctx.bf_match(dq, dt, 0.6)
    assert np.array_equal(kg, ko) and np.array_equal(mg, mo) and ng == no
    if nq and nt >= 2:
        # size-independent properties: d0 <= d1; equal distances keep index order; match implies the ratio test
        assert (kg[:, 0] <= kg[:, 2]).all()
        tie = kg[:, 0] == kg[:, 2]
        assert (kg[tie, 1] < kg[tie, 3]).all()
        acc = mg >= 0
        assert (kg[acc, 0].astype(np.float64) < kg[acc, 2].astype(np.float64) * 0.6).all()
        assert np.array_equal(mg[acc], kg[acc, 1])


def test_bf_match_ratio_boundaries(ctx, oracle):
    """d0 = 3, d1 = 5: 3 < 5 * 0.6 is FALSE in double (5 * 0.6 rounds to 3.0); d0 = 2, d1 = 5 passes."""
    dt = np.zeros((2, 32), np.uint8)
    dq = np.zeros((2, 32), np.uint8)
    dt[0, 0] = 0b00000111  # distance 3 from query 0
    dt[1, 0] = 0b11111000  # distance 5
    dq[1, 0] = 0b00000001  # distances 2 and 6 -> 2 < 3.6 accepted
    mg, kg, ng = ctx.bf_match(dq, dt, 0.6)
    mo, ko, no = oracle.bf_match(dq, dt, 0.6)
    assert kg.tolist() == [[3, 0, 5, 1], [2, 0, 6, 1]] and np.array_equal(kg, ko)
    assert mg.tolist() == [-1, 0] and np.array_equal(mg, mo) and ng == 1


def test_bf_match_batch_matches_single_pairs(ctx, oracle):
    """Sequence replay shape: many frame pairs in one launch == the pairs one by one (indices local to the pair)."""
    rng = np.random.default_rng(5)
    pairs = [synth.orb_frame_pair(int(rng.integers(1, 700)), int(rng.integers(0, 700)), seed=100 + i,
                                  entropy_bytes=int(rng.choice([3, 32]))) for i in range(37)]
    q_ptr = np.zeros(len(pairs) + 1, np.int32)
    t_ptr = np.zeros(len(pairs) + 1, np.int32)
    np.cumsum([p[0].shape[0] for p in pairs], out=q_ptr[1:])
    np.cumsum([p[1].shape[0] for p in pairs], out=t_ptr[1:])
    mg, kg, ng = ctx.bf_match_batch(q_ptr, np.concatenate([p[0] for p in pairs]), t_ptr, np.concatenate([p[1] for p in pairs]),
                                    0.6, want_knn=True)
    tot = 0
    for i, (dq, dt) in enumerate(pairs):
        mo, ko, no = oracle.bf_match(dq, dt, 0.6)
        assert np.array_equal(mg[q_ptr[i]:q_ptr[i + 1]], mo), i
        assert np.array_equal(kg[q_ptr[i]:q_ptr[i + 1]], ko), i
        tot += no
    assert ng == tot


def test_epipolar_prior_parity(ctx, oracle):
    x1, x2, F, _ = synth.epipolar_matches(3001, 8)
    rng = np.random.default_rng(1)
    n_feat = 5000
    fid = rng.permutation(n_feat)[:3001].astype(np.int32)
    do, po = oracle.epipolar_prior(x1, x2, F, 0.5, 1.1)
    dg, pg, dbf, pbf = ctx.epipolar_prior(fid, x1, x2, F, 0.5, 1.1, n_feat)
    assert np.array_equal(dg.view(np.int64), do.view(np.int64)), "symmetric epipolar distance must be bit-identical"
    assert np.allclose(pg, po, rtol=4.5e-16, atol=0), "likelihood beyond 2 ulp"
    # flat form of mvFeatureMatchDis / mvFeatureMatchProb: matched features carry their value, the others read 0.0
    assert np.array_equal(dbf[fid], dg) and np.array_equal(pbf[fid], pg)
    rest = np.ones(n_feat, bool)
    rest[fid] = False
    assert (dbf[rest] == 0).all() and (pbf[rest] == 0).all()
    # degenerate inputs: empty match list is a no-op
    d0, p0, _, _ = ctx.epipolar_prior(None, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), F, 0.5, 1.1)
    assert d0.size == 0 and p0.size == 0


def test_rough_classify_with_epipolar_prior(pkg, ctx, oracle):
    """The four-likelihood branch of RroughClassify (Tracking.cc:2001-2010) fed by the device-computed prior."""
    from oracle.pyoracle import slam_params
    fr = synth.slam_frame(3000, seed=12)
    x1, x2, F, _ = synth.epipolar_matches(3000, 9)
    prm, prm_o = pkg.SlamParams.make(), slam_params(**synth.SLAM_PARAMS)
    _, pg, _, _ = ctx.epipolar_prior(None, x1, x2, F, prm.u_gamma, prm.stdev_gamma)
    _, po = oracle.epipolar_prior(x1, x2, F, prm_o.u_gamma, prm_o.stdev_gamma)
    lg = ctx.rough_classify(fr.observs, fr.error, fr.depth, prm, pg)
    lo = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o, po)
    # labels may differ only where the decision sum sits within the likelihood tolerance of the threshold
    assert int((lg != lo).sum()) == 0
    l3 = ctx.rough_classify(fr.observs, fr.error, fr.depth, prm, None)
    assert (lg != l3).any(), "the prior must change some labels (otherwise this test exercises nothing)"

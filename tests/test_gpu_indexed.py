"""GPU tests (-m gpu) of the resident-keyframe path: observations as (keyframe, feature index) pairs -- the reference's
own MapPoint::mObservations entries (include/MapPoint.h:115) -- with the keypoints gathered on the device from the
resident per-keyframe arrays (KeyFrame::mvKeysUn, src/Tracking.cc:1831).  The indexed path must deliver the bits of
the flat-snapshot path (which the other tests pin against the oracle) and of the oracle itself."""
import importlib

import numpy as np
import pytest

from util import assert_bit_exact, bits

pytestmark = pytest.mark.gpu
synth = importlib.import_module("lc-crf-slam_b200.synth")


def oracle_params():
    from oracle.pyoracle import slam_params
    return slam_params(**synth.SLAM_PARAMS)


def indexed(cat, n_kf, seed, stride=None, dtype=np.uint16):
    fid, table, uvc = synth.index_observations(cat["obs_kf"], cat["obs_uv"], n_kf, seed=seed, stride=stride)
    ref = np.ascontiguousarray(np.stack([cat["obs_kf"], fid], axis=1).astype(dtype))
    return ref, table, uvc


@pytest.mark.parametrize("dtype", [np.uint16, np.int32])
@pytest.mark.parametrize("stride", [None, 64])
def test_indexed_matches_flat_and_oracle(pkg, ctx, oracle, dtype, stride):
    """Batch of ragged problems with per-problem keyframe slices; stride 64 forces keypoint re-use inside a keyframe."""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    snaps = [synth.map_snapshot(n, o, seed=190 + i, n_kf=300, ragged=r) for i, (n, o, r) in
             enumerate(((2500, 64, False), (130, 7, True), (4000, 20, True), (31, 64, False)))]
    cat = pkg.concat_frames(snaps)
    n_kf = cat["kf_pose"].shape[0]
    ref, table, uvc = indexed(cat, n_kf, seed=5, stride=stride, dtype=dtype)
    assert (stride is None) == np.array_equal(uvc, cat["obs_uv"])
    F = pkg.Frames(ctx, [s.n for s in snaps], prm, en)
    F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], uvc, cat["kf_pose"], cat["kf_intr"], cat["kf_bounds"],
                     cat["kp2d"], cat["kf_ptr"])
    F.run()
    (m0, p0), d0 = F.get_outputs(), F.get_debug()
    F.set_keyframe_keypoints(table)
    for kf_ptr in (cat["kf_ptr"], None):  # shared-memory keyframe slices / global keyframe table
        F.set_map_inputs_indexed(cat["xyz"], cat["obs_ptr"], ref, cat["kf_pose"], cat["kf_intr"], cat["kf_bounds"],
                                 cat["kp2d"], kf_ptr)
        F.run()
        F.run()  # graph replay
        (m1, p1), d1 = F.get_outputs(), F.get_debug()
        for k in ("observs", "error", "depth"):
            assert np.array_equal(bits(d0[k]), bits(d1[k])), k
        assert np.array_equal(d0["init_label"], d1["init_label"])
        assert np.array_equal(bits(p0), bits(p1)) and np.array_equal(m0, m1)
    # and against the oracle, problem by problem, on the flat restatement of the same problem
    o = e = 0
    for s in snaps:
        s.obs_uv = np.ascontiguousarray(uvc[e:e + s.nnz])
        ob, er, de = oracle.map_point_unary(s)
        assert np.array_equal(bits(er), bits(d1["error"][o:o + s.n])) and np.array_equal(bits(de), bits(d1["depth"][o:o + s.n]))
        Qo, mo, _ = oracle.slam_crf(ob, er, s.kp2d, d1["init_label"][o:o + s.n], en, prm_o)
        assert_bit_exact(p1[o:o + s.n], Qo)
        assert np.array_equal(m1[o:o + s.n], mo)
        o += s.n
        e += s.nnz
    F.close()


def test_indexed_mixed_cameras_small_table(pkg, ctx, oracle):
    """Keyframes with different intrinsics (no uniform-camera shortcut) and a table that fits the whole-table
    shared-memory mode (nKF <= 384)."""
    prm = pkg.SlamParams.make()
    s = synth.map_snapshot(3001, 16, seed=21, n_kf=40, ragged=True)
    s.kf_intr = s.kf_intr.copy()
    s.kf_intr[::3, 0] *= np.float32(1.01)
    s.kf_bounds = s.kf_bounds.copy()
    s.kf_bounds[1::4, 1] -= np.float32(7)
    fid, table, uvc = synth.index_observations(s.obs_kf, s.obs_uv, 40, seed=9)
    s.obs_uv = uvc
    ref = np.ascontiguousarray(np.stack([s.obs_kf, fid], axis=1).astype(np.uint16))
    F = pkg.Frames(ctx, [s.n], prm)
    F.set_keyframe_keypoints(table)
    F.set_map_inputs_indexed(s.xyz, s.obs_ptr, ref, s.kf_pose, s.kf_intr, s.kf_bounds, s.kp2d)
    F.run()
    d = F.get_debug()
    ob, er, de = oracle.map_point_unary(s)
    assert np.array_equal(bits(er), bits(d["error"])) and np.array_equal(bits(de), bits(d["depth"]))
    assert np.array_equal(ob, d["observs"])
    F.close()


def test_indexed_pipelined_submit_and_keyframe_insertion(pkg, ctx):
    """submit_map_indexed in both slots delivers the bits of set_map_inputs + run; the table grows keyframe by
    keyframe (earlier rows stay) and a moved table invalidates the captured graphs."""
    prm = pkg.SlamParams.make()
    n_kf = 96
    A = synth.map_snapshot(5001, 24, seed=31, n_kf=n_kf, ragged=True)
    B = synth.map_snapshot(5001, 24, seed=32, n_kf=n_kf, ragged=True)
    F = pkg.Frames(ctx, [A.n], prm)
    with pytest.raises(pkg.LccrfError):  # no resident table yet
        F.set_map_inputs_indexed(A.xyz, A.obs_ptr, np.zeros((A.nnz, 2), np.uint16), A.kf_pose, A.kf_intr, A.kf_bounds, A.kp2d)
    stride = 2048
    want, refs = [], []
    for i, s in enumerate((A, B)):
        fid, table, uvc = synth.index_observations(s.obs_kf, s.obs_uv, n_kf, seed=40 + i, stride=stride)
        assert np.array_equal(uvc, s.obs_uv)
        F.set_map_inputs(s.xyz, s.obs_ptr, s.obs_kf, s.obs_uv, s.kf_pose, s.kf_intr, s.kf_bounds, s.kp2d)
        F.run()
        want.append(F.get_outputs())
        # problem A uses keyframes [0, 96), problem B the keyframes [96, 192) of the same resident table
        refs.append((np.ascontiguousarray(np.stack([s.obs_kf + i * n_kf, fid], axis=1).astype(np.uint16)), table))
    assert not np.array_equal(want[0][0], want[1][0])
    # keyframe insertion in pieces: first A's keyframes in two calls, run, then B's (the table grows and moves)
    F.set_keyframe_keypoints(refs[0][1][:50], kf_first=0)
    F.set_keyframe_keypoints(refs[0][1][50:], kf_first=50)
    outs = [(np.empty(A.n, np.int16), np.empty((A.n, 2), np.float32)) for _ in range(2)]
    pose2 = lambda s, i: (np.concatenate([np.zeros_like(s.kf_pose)] * i + [s.kf_pose]),
                          np.concatenate([s.kf_intr] * (i + 1)), np.concatenate([s.kf_bounds] * (i + 1)))
    pa = pose2(A, 0)
    F.submit_map_indexed(0, A.xyz, A.obs_ptr, refs[0][0], *pa, A.kp2d, None, *outs[0])
    F.wait(0)
    assert np.array_equal(outs[0][0], want[0][0]) and np.array_equal(bits(outs[0][1]), bits(want[0][1]))
    with pytest.raises(pkg.LccrfError):  # stride is fixed once the table exists
        F.set_keyframe_keypoints(refs[1][1][:, :1024], kf_first=n_kf)
    F.set_keyframe_keypoints(refs[1][1], kf_first=n_kf)
    pb = pose2(B, 1)
    seq = [(A, 0, pa), (B, 1, pb), (B, 1, pb), (A, 0, pa), (B, 1, pb)]
    for i, (s, w, pp) in enumerate(seq):
        slot = i & 1
        if i >= 2:
            F.wait(slot)
            pw = seq[i - 2][1]
            assert np.array_equal(outs[slot][0], want[pw][0]) and np.array_equal(bits(outs[slot][1]), bits(want[pw][1])), i
        F.submit_map_indexed(slot, s.xyz, s.obs_ptr, refs[w][0], *pp, s.kp2d, None, *outs[slot])
    with pytest.raises(pkg.LccrfError):  # not while a submission is in flight
        F.set_keyframe_keypoints(refs[1][1], kf_first=n_kf)
    F.wait(0)
    F.wait(1)
    assert np.array_equal(outs[0][0], want[1][0]) and np.array_equal(outs[1][0], want[0][0])
    F.close()

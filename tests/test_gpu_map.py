"""GPU tests (-m gpu) of the device-resident map (lccrf_map_*) and the frame batches that name their points by map
point ids (lccrf_frames_set_visible / submit_visible).  The checker is a host model of the reference's own mutators
(MapPoint::AddObservation / EraseObservation / SetBadFlag / SetWorldPos, KeyFrame::SetPose -- src/MapPoint.cc:73-168,
src/KeyFrame.cc:70) kept in plain Python lists; after every delta the map's content must equal the model's, the unary
must equal the oracle's on the model's CSR snapshot bit for bit, and the whole frame CRF must be bit-identical to
lccrf_frames_set_map_inputs on the same snapshot."""
import importlib

import numpy as np
import pytest

from util import assert_bit_exact, bits

pytestmark = pytest.mark.gpu
synth = importlib.import_module("lc-crf-slam_b200.synth")
STRIDE = 2048


def oracle_params():
    from oracle.pyoracle import slam_params
    return slam_params(**synth.SLAM_PARAMS)


class HostMap:
    """the reference's map state, restated: per point a list of (keyframe, feature index) in insertion order"""

    def __init__(self, stride):
        self.stride = stride
        self.pose = np.zeros((0, 12), np.float32)
        self.intr = np.zeros((0, 4), np.float32)
        self.bounds = np.zeros((0, 4), np.float32)
        self.kp = np.zeros((0, stride, 2), np.float32)
        self.xyz = np.zeros((0, 3), np.float32)
        self.obs = []

    def add_keyframes(self, pose, intr, bounds, kp):
        self.pose = np.concatenate([self.pose, pose])
        self.intr = np.concatenate([self.intr, intr])
        self.bounds = np.concatenate([self.bounds, bounds])
        self.kp = np.concatenate([self.kp, kp])

    def set_xyz(self, ids, xyz):
        n = int(max(ids)) + 1
        if n > self.xyz.shape[0]:
            self.xyz = np.concatenate([self.xyz, np.zeros((n - self.xyz.shape[0], 3), np.float32)])
            self.obs += [[] for _ in range(n - len(self.obs))]
        self.xyz[ids] = xyz

    def add(self, p, kf, fid):           # MapPoint::AddObservation, src/MapPoint.cc:98-109
        if all(k != kf for k, _ in self.obs[p]):
            self.obs[p].append((kf, fid))

    def erase(self, p, kf):              # MapPoint::EraseObservation, :111-141
        self.obs[p] = [(k, f) for k, f in self.obs[p] if k != kf]

    def bad(self, p):                    # MapPoint::SetBadFlag, :151-168
        self.obs[p] = []

    def snapshot(self, ids, kp2d):
        """the lccrf_frames_set_map_inputs layout of the given points"""
        ptr = np.zeros(len(ids) + 1, np.int32)
        kf, uv = [], []
        for i, p in enumerate(ids):
            for k, f in self.obs[p]:
                kf.append(k)
                uv.append(self.kp[k, f])
            ptr[i + 1] = len(kf)
        return synth.MapSnapshot(self.xyz[ids].copy(), ptr, np.asarray(kf, np.int32), np.asarray(uv, np.float32).reshape(-1, 2),
                                 self.pose.copy(), self.intr.copy(), self.bounds.copy(), kp2d, np.zeros(len(ids), bool))


def make_keyframes(rng, n, intr=synth.TUM_INTR, first=0):
    """n keyframes on a smooth trajectory + a random keypoint table"""
    pose = np.zeros((n, 12), np.float64)
    for k in range(n):
        t = (first + k) / 64.0
        R = synth._rot(0.05 * np.sin(6.28 * t), 0.08 * np.sin(6.28 * t + 1.0), 0.03 * np.cos(6.28 * t))
        c = np.array([0.3 * np.sin(6.28 * t), 0.1 * np.cos(6.28 * t), 0.2 * t])
        pose[k] = np.concatenate([R, (-R @ c)[:, None]], axis=1).reshape(-1)
    kp = np.stack([rng.uniform(0, 640, (n, STRIDE)), rng.uniform(0, 480, (n, STRIDE))], axis=2).astype(np.float32)
    return (pose.astype(np.float32), np.tile(np.array(intr, np.float32), (n, 1)),
            np.tile(np.array([0, 640, 0, 480], np.float32), (n, 1)), kp)


def make_points(rng, n, intr=synth.TUM_INTR):
    fx, fy, cx, cy = intr
    xy = np.stack([rng.uniform(20, 620, n), rng.uniform(20, 460, n)], axis=1)
    z = np.clip(rng.normal(2.75, 0.45, n), 1.0, 5.0)
    xyz = np.stack([(xy[:, 0] - cx) / fx * z, (xy[:, 1] - cy) / fy * z, z], axis=1).astype(np.float32)
    return xyz, xy.astype(np.float32)


def project_keypoints(hm, kf, pts, fids, rng, noise=1.2):
    """make the keypoints that observations (kf, fid) refer to plausible: the projection of the point + noise"""
    P = hm.pose[kf].reshape(3, 4).astype(np.float64)
    X = hm.xyz[pts].astype(np.float64)
    Xc = X @ P[:, :3].T + P[:, 3]
    fx, fy, cx, cy = hm.intr[kf]
    uv = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], axis=1) + rng.normal(0, noise, (len(pts), 2))
    hm.kp[kf, fids] = uv.astype(np.float32)


def check_against_model(pkg, ctx, oracle, mp, hm, F1, F2, ids, kp2d, delta=None, kf_ptr=None):
    """map content == model; unary == oracle; visible CRF == set_map_inputs CRF, all bit for bit"""
    snap = hm.snapshot(ids, kp2d)
    F1.set_visible(mp, ids, kp2d, delta=delta, kf_ptr=kf_ptr)
    F1.run()
    m1, p1 = F1.get_outputs()
    d1 = F1.get_debug()
    ptr, kf, uv, xyz = mp.export(ids)
    assert np.array_equal(ptr, snap.obs_ptr) and np.array_equal(kf, snap.obs_kf)
    assert np.array_equal(bits(uv), bits(snap.obs_uv)) and np.array_equal(bits(xyz), bits(snap.xyz))
    ob, er, de = oracle.map_point_unary(snap)
    assert np.array_equal(ob, d1["observs"])
    assert np.array_equal(bits(er), bits(d1["error"])) and np.array_equal(bits(de), bits(d1["depth"]))
    F2.set_map_inputs(snap.xyz, snap.obs_ptr, snap.obs_kf, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d, kf_ptr)
    F2.run()
    m2, p2 = F2.get_outputs()
    assert_bit_exact(p1, p2, what="visible vs set_map_inputs")
    assert np.array_equal(m1, m2)
    return m1, p1, snap


def test_map_bulk_load_matches_snapshot(pkg, ctx, oracle):
    """a map loaded in one go (keyframes + CSR of observations) gives the results of the flat snapshot"""
    snap = synth.map_snapshot(5003, 16, seed=3, n_kf=64, ragged=True)
    fid, table, uvc = synth.index_observations(snap.obs_kf, snap.obs_uv, 64, seed=9, stride=STRIDE)
    mp = pkg.Map(ctx, STRIDE)
    mp.apply(kf_pose=snap.kf_pose, kf_intr=snap.kf_intr, kf_bounds=snap.kf_bounds, kf_keypoints=table, xyz=snap.xyz)
    mp.set_observations(snap.obs_ptr, np.stack([snap.obs_kf, fid], axis=1))
    sz = mp.sizes()
    assert sz["n_kf"] == 64 and sz["n_points"] == snap.n and sz["n_obs"] == snap.nnz and sz["pool_used"] >= snap.nnz
    ids = np.arange(snap.n, dtype=np.int32)
    F1, F2 = pkg.Frames(ctx, [snap.n]), pkg.Frames(ctx, [snap.n])
    F1.set_visible(mp, ids, snap.kp2d)
    for graphs in (0, 1, 1):
        ctx.set_option("graphs", graphs)
        F1.run()
    m1, p1 = F1.get_outputs()
    F2.set_map_inputs(snap.xyz, snap.obs_ptr, snap.obs_kf, uvc, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F2.run()
    m2, p2 = F2.get_outputs()
    assert_bit_exact(p1, p2)
    assert np.array_equal(m1, m2)
    # a permuted subset of the points as one frame
    sub = np.random.default_rng(1).permutation(snap.n)[:3001].astype(np.int32)
    F3, F4 = pkg.Frames(ctx, [sub.size]), pkg.Frames(ctx, [sub.size])
    hm_ptr = np.zeros(sub.size + 1, np.int32)
    cnt = np.diff(snap.obs_ptr)[sub]
    np.cumsum(cnt, out=hm_ptr[1:])
    sel = np.concatenate([np.arange(snap.obs_ptr[p], snap.obs_ptr[p + 1]) for p in sub])
    F3.set_visible(mp, sub, snap.kp2d[sub])
    F3.run()
    F4.set_map_inputs(snap.xyz[sub], hm_ptr, snap.obs_kf[sel], uvc[sel], snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d[sub])
    F4.run()
    assert_bit_exact(F3.get_outputs()[1], F4.get_outputs()[1])
    ob, er, de = oracle.map_point_unary(synth.MapSnapshot(snap.xyz[sub], hm_ptr, snap.obs_kf[sel], uvc[sel], snap.kf_pose,
                                                          snap.kf_intr, snap.kf_bounds, snap.kp2d[sub], None))
    assert np.array_equal(bits(er), bits(F3.get_debug()["error"]))
    for f in (F1, F2, F3, F4):
        f.close()
    mp.close()


@pytest.mark.parametrize("n_pts,mixed_cameras", [(3000, False), (1200, True)])
def test_map_incremental_deltas_against_host_model(pkg, ctx, oracle, n_pts, mixed_cameras):
    """keyframe insertions, culling, bad points, bundle-adjustment updates and new points over six deltas"""
    rng = np.random.default_rng(n_pts)
    hm = HostMap(STRIDE)
    mp = pkg.Map(ctx, STRIDE)
    # initial map: 8 keyframes, n_pts points, 1..6 observations each, loaded in bulk
    pose, intr, bounds, kp = make_keyframes(rng, 8)
    if mixed_cameras:
        intr[1::2] = np.array(synth.BONN_INTR, np.float32)
    hm.add_keyframes(pose, intr, bounds, kp)
    xyz, _ = make_points(rng, n_pts)
    hm.set_xyz(np.arange(n_pts), xyz)
    for k in range(8):
        pts = np.nonzero(rng.random(n_pts) < 0.45)[0]
        fids = rng.permutation(STRIDE)[:pts.size]
        project_keypoints(hm, k, pts, fids, rng)
        for p, f in zip(pts, fids):
            hm.add(int(p), k, int(f))
    mp.apply(kf_pose=hm.pose, kf_intr=hm.intr, kf_bounds=hm.bounds, kf_keypoints=hm.kp, xyz=hm.xyz)
    ptr = np.zeros(n_pts + 1, np.int32)
    np.cumsum([len(o) for o in hm.obs], out=ptr[1:])
    ref = np.array([kf_f for o in hm.obs for kf_f in o], np.int32).reshape(-1, 2)
    mp.set_observations(ptr, ref)
    frames = {}

    def frames_for(n):
        if n not in frames:
            frames[n] = (pkg.Frames(ctx, [n]), pkg.Frames(ctx, [n]))
        return frames[n]

    def visible():
        ids = np.array([p for p in range(len(hm.obs)) if hm.obs[p]], np.int32)
        ids = ids[rng.permutation(ids.size)]
        kp2d = np.stack([rng.uniform(0, 640, ids.size), rng.uniform(0, 480, ids.size)], axis=1).astype(np.float32)
        return ids, kp2d

    ids, kp2d = visible()
    check_against_model(pkg, ctx, oracle, mp, hm, *frames_for(ids.size), ids, kp2d)
    for step in range(6):
        nk = hm.pose.shape[0]
        # one new keyframe observing ~40% of the live points (+ a repeated (point, keyframe) pair: ignored, :101-102)
        kpose, kintr, kbounds, kkp = make_keyframes(rng, 1, first=nk)
        if mixed_cameras and step % 2:
            kintr[:] = np.array(synth.BONN_INTR, np.float32)
        hm.add_keyframes(kpose, kintr, kbounds, kkp)
        live = np.array([p for p in range(len(hm.obs)) if hm.obs[p]])
        add_pt = live[rng.random(live.size) < 0.4]
        add_kf = np.full(add_pt.size, nk)
        add_fid = rng.permutation(STRIDE)[:add_pt.size]
        project_keypoints(hm, nk, add_pt, add_fid, rng)
        dup = rng.random(add_pt.size) < 0.05   # these name a keyframe the point already has
        for i in np.nonzero(dup)[0]:
            add_kf[i], add_fid[i] = hm.obs[int(add_pt[i])][0]
        # new points: created with a position and observed by the new keyframe
        n_new = 50
        new_ids = np.arange(len(hm.obs), len(hm.obs) + n_new)
        new_xyz, _ = make_points(rng, n_new)
        # bundle adjustment: all poses jitter, a third of the points move
        pose_kf = rng.permutation(nk)[: max(1, nk // 2)]
        new_pose = hm.pose[pose_kf] + rng.normal(0, 1e-3, (pose_kf.size, 12)).astype(np.float32)
        mv = live[rng.random(live.size) < 0.33]
        mv_xyz = hm.xyz[mv] + rng.normal(0, 2e-3, (mv.size, 3)).astype(np.float32)
        # culling: keyframe `step` loses its observations in 60% of its points; 3% of the points go bad
        obs_in = np.array([p for p in live if any(k == step for k, _ in hm.obs[p])])
        er_pt = obs_in[rng.random(obs_in.size) < 0.6] if obs_in.size else obs_in
        er_pt = np.concatenate([er_pt, live[:3][~np.isin(live[:3], er_pt)]]).astype(np.int64)  # incl. absent pairs: no-op
        bad = live[rng.random(live.size) < 0.03]
        # --- the model, in the delta's order
        hm.pose[pose_kf] = new_pose
        hm.set_xyz(np.concatenate([mv, new_ids]), np.concatenate([mv_xyz, new_xyz]))
        for p in er_pt:
            hm.erase(int(p), step)
        for p in bad:
            hm.bad(int(p))
        all_pt = np.concatenate([add_pt, new_ids])
        all_kf = np.concatenate([add_kf, np.full(n_new, nk)])
        all_fid = np.concatenate([add_fid, STRIDE - 1 - np.arange(n_new)])
        project_keypoints(hm, nk, new_ids, all_fid[add_pt.size:], rng)
        for p, k, f in zip(all_pt, all_kf, all_fid):
            hm.add(int(p), int(k), int(f))
        delta = pkg.MapDelta.make(kf_first=nk, kf_pose=kpose, kf_intr=kintr, kf_bounds=kbounds, kf_keypoints=hm.kp[nk:nk + 1],
                                  pose_kf=pose_kf, pose=new_pose, xyz_id=np.concatenate([mv, new_ids]),
                                  xyz=np.concatenate([mv_xyz, new_xyz]), erase_pt=er_pt, erase_kf=np.full(er_pt.size, step),
                                  bad_pt=bad, add_pt=all_pt, add_kf=all_kf, add_fid=all_fid)
        ids, kp2d = visible()
        if step % 2:   # through lccrf_map_apply, then a frame without delta
            mp.apply(delta)
            delta = None
        check_against_model(pkg, ctx, oracle, mp, hm, *frames_for(ids.size), ids, kp2d, delta=delta)
    sz = mp.sizes()
    assert sz["n_kf"] == hm.pose.shape[0] and sz["n_points"] == len(hm.obs) and sz["n_obs"] == sum(len(o) for o in hm.obs)
    for a, b in frames.values():
        a.close()
        b.close()
    mp.close()


def test_map_batch_with_keyframe_slices(pkg, ctx, oracle):
    """four independent problems in one batch, 200 keyframes each (800 > the 384 that fit shared memory): the unary
    kernel keeps the current problem's keyframe slice in shared memory (kf_ptr), C3's shape at reduced size"""
    snaps = [synth.map_snapshot(4000 + 7 * i, 24, seed=60 + i, n_kf=200) for i in range(4)]
    cat = pkg.concat_frames(snaps)
    fids, tabs, uvs, ko = [], [], [], 0
    for i, s in enumerate(snaps):
        fid, tab, uvc = synth.index_observations(s.obs_kf, s.obs_uv, 200, seed=3 + i, stride=STRIDE)
        fids.append(np.stack([s.obs_kf + ko, fid], axis=1))
        tabs.append(tab)
        uvs.append(uvc)
        ko += 200
    mp = pkg.Map(ctx, STRIDE)
    mp.apply(kf_pose=cat["kf_pose"], kf_intr=cat["kf_intr"], kf_bounds=cat["kf_bounds"], kf_keypoints=np.concatenate(tabs), xyz=cat["xyz"])
    mp.set_observations(cat["obs_ptr"], np.concatenate(fids))
    sizes = [s.n for s in snaps]
    NT = sum(sizes)
    F1, F2 = pkg.Frames(ctx, sizes), pkg.Frames(ctx, sizes)
    F1.set_visible(mp, np.arange(NT, dtype=np.int32), cat["kp2d"], kf_ptr=cat["kf_ptr"])
    F1.run()
    F1.run()
    m1, p1 = F1.get_outputs()
    F2.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], np.concatenate(uvs), cat["kf_pose"], cat["kf_intr"], cat["kf_bounds"],
                      cat["kp2d"], cat["kf_ptr"])
    F2.run()
    m2, p2 = F2.get_outputs()
    assert_bit_exact(p1, p2)
    assert np.array_equal(m1, m2)
    d1 = F1.get_debug()
    o = 0
    for s, uvc in zip(snaps, uvs):
        ob, er, de = oracle.map_point_unary(synth.MapSnapshot(s.xyz, s.obs_ptr, s.obs_kf, uvc, s.kf_pose, s.kf_intr, s.kf_bounds, s.kp2d, None))
        assert np.array_equal(bits(er), bits(d1["error"][o:o + s.n])) and np.array_equal(bits(de), bits(d1["depth"][o:o + s.n]))
        o += s.n
    ab = F1.algorithmic_bytes()
    assert ab["unary"] == cat["obs_kf"].size * 12 + NT * 24 + 800 * 80
    # an observation that names a keyframe of another problem's slice is reported, not read out of bounds
    mp.apply(add_pt=[5], add_kf=[799], add_fid=[0])
    F1.set_visible(mp, np.arange(NT, dtype=np.int32), cat["kp2d"], kf_ptr=cat["kf_ptr"])
    F1.run()
    with pytest.raises(pkg.LccrfError, match="outside its table"):
        F1.get_outputs()
    F1.close()
    F2.close()
    mp.close()


def test_map_list_growth_many_appends(pkg, ctx, oracle):
    """a few points observed by 300 keyframes one after the other: the lists move (4 -> 8 -> ... -> 512 entries)"""
    rng = np.random.default_rng(8)
    hm, mp = HostMap(STRIDE), pkg.Map(ctx, STRIDE)
    pose, intr, bounds, kp = make_keyframes(rng, 300)
    hm.add_keyframes(pose, intr, bounds, kp)
    xyz, kp2d = make_points(rng, 40)
    hm.set_xyz(np.arange(40), xyz)
    mp.apply(kf_pose=pose, kf_intr=intr, kf_bounds=bounds, kf_keypoints=kp, xyz=xyz)
    for k in range(300):
        pts = np.arange(40)[rng.random(40) < (0.9 if k % 3 else 0.3)]
        fids = rng.permutation(STRIDE)[:pts.size]
        project_keypoints(hm, k, pts, fids, rng)
        for p, f in zip(pts, fids):
            hm.add(int(p), k, int(f))
        # the keypoints of keyframe k changed in the model: refresh the row together with the observations
        mp.apply(kf_first=k, kf_pose=pose[k:k + 1], kf_intr=intr[k:k + 1], kf_bounds=bounds[k:k + 1], kf_keypoints=hm.kp[k:k + 1],
                 add_pt=pts, add_kf=np.full(pts.size, k), add_fid=fids)
    ids = np.arange(40, dtype=np.int32)
    F1, F2 = pkg.Frames(ctx, [40]), pkg.Frames(ctx, [40])
    check_against_model(pkg, ctx, oracle, mp, hm, F1, F2, ids, kp2d)
    sz = mp.sizes()
    assert sz["n_obs"] == sum(len(o) for o in hm.obs) and sz["pool_used"] > sz["n_obs"]
    F1.close()
    F2.close()
    mp.close()


def test_map_errors_are_reported(pkg, ctx):
    rng = np.random.default_rng(2)
    mp = pkg.Map(ctx, STRIDE)
    pose, intr, bounds, kp = make_keyframes(rng, 4)
    xyz, kp2d = make_points(rng, 100)
    mp.apply(kf_pose=pose, kf_intr=intr, kf_bounds=bounds, kf_keypoints=kp, xyz=xyz)
    mp.apply(add_pt=np.arange(50), add_kf=np.zeros(50), add_fid=np.arange(50))
    with pytest.raises(pkg.LccrfError, match="twice"):
        mp.apply(add_pt=[3, 4, 3], add_kf=[1, 1, 2], add_fid=[0, 1, 2])
    with pytest.raises(pkg.LccrfError, match="outside its table"):
        mp.apply(add_pt=[5], add_kf=[4], add_fid=[0])          # keyframe 4 does not exist
    with pytest.raises(pkg.LccrfError, match="outside its table"):
        mp.apply(add_pt=[5], add_kf=[1], add_fid=[STRIDE])     # feature index beyond the keyframe's row
    with pytest.raises(pkg.LccrfError, match="outside its table"):
        mp.apply(add_pt=[100000], add_kf=[1], add_fid=[0])     # a point that was never created
    F = pkg.Frames(ctx, [60])
    F.set_visible(mp, np.arange(60, dtype=np.int32), kp2d[:60])  # points 50..59 have no observation
    F.run()
    with pytest.raises(pkg.LccrfError, match="1858"):
        F.get_outputs()
    F.set_visible(mp, list(range(50)) + list(range(10)), kp2d[:60])  # fine again (a point may be visible twice)
    F.run()
    F.get_outputs()
    F.close()
    # the snapshot path validates keyframe indices on the device as well
    snap = synth.map_snapshot(500, 4, seed=1, n_kf=16)
    bad = snap.obs_kf.copy()
    bad[77] = 16
    F = pkg.Frames(ctx, [snap.n])
    F.set_map_inputs(snap.xyz, snap.obs_ptr, bad, snap.obs_uv, snap.kf_pose, snap.kf_intr, snap.kf_bounds, snap.kp2d)
    F.run()
    with pytest.raises(pkg.LccrfError, match="outside its table"):
        F.get_outputs()
    F.close()
    mp.close()


def test_map_pipelined_submissions_with_deltas(pkg, ctx, oracle):
    """lccrf_frames_submit_visible, two slots, a delta per step, label-application lists per slot: every step's
    results equal the synchronous path on a second map that receives the same deltas"""
    rng = np.random.default_rng(5)
    B, n = 3, 1500
    maps = [pkg.Map(ctx, STRIDE), pkg.Map(ctx, STRIDE)]
    pose, intr, bounds, kp = make_keyframes(rng, 12)
    xyz, _ = make_points(rng, B * n)
    ptr = np.zeros(B * n + 1, np.int32)
    cnt = rng.integers(2, 7, B * n)
    np.cumsum(cnt, out=ptr[1:])
    ref = np.stack([np.concatenate([rng.permutation(12)[:c] for c in cnt]), rng.integers(0, STRIDE, int(ptr[-1]))], axis=1)
    for mp in maps:
        mp.apply(kf_pose=pose, kf_intr=intr, kf_bounds=bounds, kf_keypoints=kp, xyz=xyz)
        mp.set_observations(ptr, ref)
    Fp, Fs = pkg.Frames(ctx, [n] * B), pkg.Frames(ctx, [n] * B)
    parts = [Fp.set_partition_outputs(s, fid=np.arange(B * n) % 1000) for s in (0, 1)]
    outs = [(np.zeros(B * n, np.int16), np.zeros((B * n, 2), np.float32)) for _ in (0, 1)]
    expected = []
    steps = 7
    ids_all = [rng.permutation(B * n).astype(np.int32) for _ in range(steps)]
    kp_all = [np.stack([rng.uniform(0, 640, B * n), rng.uniform(0, 480, B * n)], axis=1).astype(np.float32) for _ in range(steps)]
    deltas = []
    for s in range(steps):
        k = 12 + s
        kpose, kintr, kbounds, kkp = make_keyframes(rng, 1, first=k)
        add_pt = rng.permutation(B * n)[: n // 2]
        er_pt = rng.permutation(B * n)[: n // 4]
        mv = rng.permutation(B * n)[: n]
        deltas.append(pkg.MapDelta.make(kf_first=k, kf_pose=kpose, kf_intr=kintr, kf_bounds=kbounds, kf_keypoints=kkp,
                                        pose=pose + rng.normal(0, 1e-3, pose.shape).astype(np.float32),
                                        xyz_id=mv, xyz=xyz[mv] + rng.normal(0, 1e-3, (n, 3)).astype(np.float32),
                                        erase_pt=er_pt, erase_kf=np.full(er_pt.size, k - 1),
                                        add_pt=add_pt, add_kf=np.full(add_pt.size, k), add_fid=rng.integers(0, STRIDE, add_pt.size)))
    for s in range(steps):   # synchronous path
        Fs.set_visible(maps[1], ids_all[s], kp_all[s], delta=deltas[s])
        Fs.run()
        m, p = Fs.get_outputs()
        expected.append((m.copy(), p.copy()) + Fs.partition(fid=np.arange(B * n) % 1000))
    got = {}

    def collect(s):
        Fp.wait(s & 1)
        m, p = outs[s & 1]
        pt = parts[s & 1]
        got[s] = (m.copy(), p.copy(), pt["dyn_ptr"].copy(), pt["dyn"][: pt["dyn_ptr"][-1]].copy(), pt["stat_ptr"].copy(),
                  pt["stat"][: pt["stat_ptr"][-1]].copy())

    for s in range(steps):   # pipelined path
        if s >= 2:
            collect(s - 2)
        Fp.submit_visible(s & 1, maps[0], ids_all[s], kp_all[s], outs[s & 1][0], outs[s & 1][1], delta=deltas[s])
    collect(steps - 2)
    collect(steps - 1)
    for s in range(steps):
        for a, b in zip(got[s], expected[s]):
            assert np.array_equal(a, b), "step %d" % s
    with pytest.raises(pkg.LccrfError, match="in flight"):
        Fp.submit_visible(0, maps[0], ids_all[0], kp_all[0], outs[0][0], outs[0][1])
        Fp.get_outputs()   # shared result buffers while a submission is in flight
    Fp.wait(0)
    assert maps[0].sizes() == maps[1].sizes() or maps[0].sizes()["n_obs"] == maps[1].sizes()["n_obs"]
    Fp.close()
    Fs.close()
    for mp in maps:
        mp.close()


def test_frames_prior_branch(pkg, ctx, oracle):
    """RroughClassify's second branch in the batched path (Tracking.cc:2001-2010): p4 per point, per-problem flags"""
    prm_o, prm = oracle_params(), pkg.SlamParams.make()
    en = pkg.label_energies(2, prm.confidence)
    sizes = [3000, 2500, 3100]
    frames = [synth.slam_frame(n, seed=40 + i) for i, n in enumerate(sizes)]
    rng = np.random.default_rng(6)
    p4 = rng.random(sum(sizes)) * 0.9
    has = np.array([1, 0, 1], np.uint8)
    F = pkg.Frames(ctx, sizes, prm, en)
    cat = lambda k: np.concatenate([getattr(f, k) for f in frames])
    F.set_prior(0, p4, has)
    F.set_inputs(cat("observs"), cat("error"), cat("depth"), cat("kp2d"))
    F.run()
    mp_, pr = F.get_outputs()
    lab = F.get_debug()["init_label"]
    o = 0
    n_diff = 0
    for b, fr in enumerate(frames):
        lo = oracle.rough_classify(fr.observs, fr.error, fr.depth, prm_o, p4[o:o + fr.n] if has[b] else None)
        n_diff += int((lo != lab[o:o + fr.n]).sum())
        Qo, mo, _ = oracle.slam_crf(fr.observs, fr.error, fr.kp2d, lab[o:o + fr.n], en, prm_o)
        assert_bit_exact(pr[o:o + fr.n], Qo)
        assert np.array_equal(mp_[o:o + fr.n], mo)
        o += fr.n
    assert n_diff <= 2   # device exp vs glibc expf at a threshold tie (test_rough_classify)
    # the prior changes labels, and clearing it restores the first branch
    l_no = np.concatenate([oracle.rough_classify(f.observs, f.error, f.depth, prm_o) for f in frames])
    assert (l_no != lab).sum() > 0
    F.set_prior(0, None)
    F.set_inputs(cat("observs"), cat("error"), cat("depth"), cat("kp2d"))
    F.run()
    assert np.array_equal(F.get_debug()["init_label"], l_no)
    F.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_map_random_delta_sequences(pkg, ctx, oracle, seed):
    """forty random deltas on a small map (every mutator, random sizes incl. empty ones, several segments per delta,
    erasing first / middle / last / absent entries, re-observing bad points): after each one the exported lists equal the
    host model's; at the end a frame over all live points matches the flat-snapshot path and the oracle"""
    rng = np.random.default_rng(100 + seed)
    stride, P = 64, 300
    hm, mp = HostMap(stride), pkg.Map(ctx, stride)

    def new_kf(n, first):
        pose, intr, bounds, _ = make_keyframes(rng, n, first=first)
        kp = np.stack([rng.uniform(0, 640, (n, stride)), rng.uniform(0, 480, (n, stride))], axis=2).astype(np.float32)
        return pose, intr, bounds, kp

    pose, intr, bounds, kp = new_kf(3, 0)
    hm.add_keyframes(pose, intr, bounds, kp)
    xyz, _ = make_points(rng, P)
    hm.set_xyz(np.arange(P), xyz)
    mp.apply(kf_pose=pose, kf_intr=intr, kf_bounds=bounds, kf_keypoints=kp, xyz=xyz)
    for step in range(40):
        nk = hm.pose.shape[0]
        kw = {}
        n_new = int(rng.integers(0, 3))
        if n_new:
            kpose, kintr, kbounds, kkp = new_kf(n_new, nk)
            hm.add_keyframes(kpose, kintr, kbounds, kkp)
            kw.update(kf_first=nk, kf_pose=kpose, kf_intr=kintr, kf_bounds=kbounds, kf_keypoints=kkp)
        n_kf = hm.pose.shape[0]
        if rng.random() < 0.5:
            ids = rng.permutation(n_kf)[: int(rng.integers(1, n_kf + 1))]
            newp = hm.pose[ids] + rng.normal(0, 1e-3, (ids.size, 12)).astype(np.float32)
            hm.pose[ids] = newp
            kw.update(pose_kf=ids, pose=newp)
        if rng.random() < 0.5:
            grow = int(rng.integers(0, 4))
            ids = np.concatenate([rng.permutation(len(hm.obs))[: int(rng.integers(0, 40))], np.arange(len(hm.obs), len(hm.obs) + grow)]).astype(np.int64)
            if ids.size:
                nx, _ = make_points(rng, ids.size)
                hm.set_xyz(ids, nx)
                kw.update(xyz_id=ids, xyz=nx)
        # erase: up to 3 segments, each a random keyframe and a random set of points (with or without that keyframe)
        er_pt, er_kf, er_seg = [], [], [0]
        for _ in range(int(rng.integers(0, 4))):
            k = int(rng.integers(0, n_kf))
            pts = rng.permutation(len(hm.obs))[: int(rng.integers(0, 60))]
            for p in pts:
                hm.erase(int(p), k)
            er_pt.append(pts)
            er_kf.append(np.full(pts.size, k))
            er_seg.append(er_seg[-1] + pts.size)
        if er_pt:
            kw.update(erase_pt=np.concatenate(er_pt), erase_kf=np.concatenate(er_kf), erase_seg_ptr=er_seg)
        bad = rng.permutation(len(hm.obs))[: int(rng.integers(0, 5))] if rng.random() < 0.4 else np.zeros(0, np.int64)
        for p in bad:
            hm.bad(int(p))
        if bad.size:
            kw.update(bad_pt=bad)
        ad_pt, ad_kf, ad_fid, ad_seg = [], [], [], [0]
        for _ in range(int(rng.integers(0, 4))):
            k = int(rng.integers(0, n_kf))
            pts = rng.permutation(len(hm.obs))[: int(rng.integers(0, 120))]
            fids = rng.integers(0, stride, pts.size)
            for p, f in zip(pts, fids):
                hm.add(int(p), k, int(f))
            ad_pt.append(pts)
            ad_kf.append(np.full(pts.size, k))
            ad_fid.append(fids)
            ad_seg.append(ad_seg[-1] + pts.size)
        if ad_pt:
            kw.update(add_pt=np.concatenate(ad_pt), add_kf=np.concatenate(ad_kf), add_fid=np.concatenate(ad_fid), add_seg_ptr=ad_seg)
        mp.apply(**kw)
        ids_all = np.arange(len(hm.obs), dtype=np.int32)
        ptr, kf, uv, xyz_d = mp.export(ids_all)
        want = hm.snapshot(ids_all, np.zeros((ids_all.size, 2), np.float32))
        assert np.array_equal(ptr, want.obs_ptr) and np.array_equal(kf, want.obs_kf), step
        assert np.array_equal(bits(uv), bits(want.obs_uv)) and np.array_equal(bits(xyz_d), bits(want.xyz)), step
    sz = mp.sizes()
    assert sz["n_obs"] == sum(len(o) for o in hm.obs) and sz["n_kf"] == hm.pose.shape[0] and sz["n_points"] == len(hm.obs)
    live = np.array([p for p in range(len(hm.obs)) if hm.obs[p]], np.int32)
    kp2d = np.stack([rng.uniform(0, 640, live.size), rng.uniform(0, 480, live.size)], axis=1).astype(np.float32)
    F1, F2 = pkg.Frames(ctx, [live.size]), pkg.Frames(ctx, [live.size])
    check_against_model(pkg, ctx, oracle, mp, hm, F1, F2, live, kp2d)
    F1.close()
    F2.close()
    mp.close()

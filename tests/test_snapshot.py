"""CRF-input snapshot / replay files: host-only format tests (no GPU needed) and the replay of a
file through the device path."""
import importlib
import os

import numpy as np
import pytest

pkg_mod = importlib.import_module("lc-crf-slam_b200")
synth = pkg_mod.synth

FIELDS = ("xyz", "obs_ptr", "obs_kf", "obs_uv", "kf_pose", "kf_intr", "kf_bounds", "kp2d")


def _frames():
    return [synth.map_snapshot(n, o, seed=40 + i, n_kf=k, ragged=r) for i, (n, o, k, r) in
            enumerate(((300, 12, 40, True), (1, 1, 4, False), (0, 1, 4, False), (257, 64, 300, False)))]


def _same(a, b):
    for k in FIELDS:
        x, y = np.asarray(getattr(a, k)), np.asarray(getattr(b, k))
        assert x.size == y.size and np.array_equal(x.reshape(-1).view(np.int32) if x.dtype == np.float32 else x.reshape(-1),
                                                   y.reshape(-1).view(np.int32) if y.dtype == np.float32 else y.reshape(-1)), k


def test_round_trip_append_and_index(tmp_path):
    path = str(tmp_path / "seq.snp")
    frames = _frames()
    rng = np.random.default_rng(0)
    fids = [rng.permutation(2 * f.n + 1)[:f.n].astype(np.int32) for f in frames]
    with pkg_mod.SnapshotWriter(path) as w:
        for i, f in enumerate(frames[:2]):
            w.write(f, frame_id=100 + i, timestamp=0.5 * i, fid=fids[i])
    with pkg_mod.SnapshotWriter(path, append=True) as w:  # a second session continues the file
        for i, f in enumerate(frames[2:], start=2):
            w.write(f, frame_id=100 + i, timestamp=0.5 * i, fid=fids[i])
    with pkg_mod.SnapshotReader(path) as r:
        assert len(r) == len(frames) and not r.truncated
        for i in (3, 0, 2, 1):  # random access
            inf = r.info(i)
            assert (inf.N, inf.nKF, inf.nnz, inf.frame_id, inf.timestamp) == (
                frames[i].n, frames[i].kf_pose.shape[0], frames[i].nnz, 100 + i, 0.5 * i)
            assert inf.stored_kf_bytes == 2
            g = r.read(i)
            _same(g, frames[i])
            assert np.array_equal(g.fid, fids[i])
            g16 = r.read(i, kf_dtype=np.uint16)  # the compact form lccrf_frames_submit_map takes
            assert g16.obs_kf.dtype == np.uint16 and np.array_equal(g16.obs_kf.astype(np.int32), frames[i].obs_kf)
        with pytest.raises(pkg_mod.LccrfError):
            r.read(len(frames))


def test_truncated_file_keeps_complete_frames_and_corruption_is_detected(tmp_path):
    path = str(tmp_path / "cut.snp")
    frames = _frames()
    with pkg_mod.SnapshotWriter(path) as w:
        for f in frames:
            w.write(f)
    size = os.path.getsize(path)
    blob = open(path, "rb").read()
    # the writer died inside the last frame: every complete frame is still readable
    cut = str(tmp_path / "cut2.snp")
    open(cut, "wb").write(blob[:size - 1000])
    with pkg_mod.SnapshotReader(cut) as r:
        assert len(r) == len(frames) - 1 and r.truncated
        _same(r.read(0), frames[0])
    # appending to the cut file continues behind the last complete frame (the partial record is cut off), so the file
    # stays readable and the new frame follows the old ones
    with pkg_mod.SnapshotWriter(cut, append=True) as w:
        w.write(frames[-1], frame_id=99)
    with pkg_mod.SnapshotReader(cut) as r:
        assert len(r) == len(frames) and not r.truncated
        _same(r.read(len(frames) - 2), frames[-2])
        _same(r.read(len(frames) - 1), frames[-1])
        assert r.info(len(frames) - 1).frame_id == 99
    # a damaged record header in the middle ends the index there: the frames in front of it stay readable
    import struct
    hdr_hit = bytearray(blob)
    payload0 = struct.unpack_from("<Q", blob, 32 + 40)[0]
    off1 = 32 + 64 + payload0                      # header of the second record
    hdr_hit[off1] ^= 0xFF
    hp = str(tmp_path / "hdr.snp")
    open(hp, "wb").write(bytes(hdr_hit))
    with pkg_mod.SnapshotReader(hp) as r:
        assert len(r) == 1 and r.truncated
        _same(r.read(0), frames[0])
    # a flipped payload byte fails the checksum of that frame only
    bad = bytearray(blob)
    bad[32 + 64 + 200] ^= 0x40
    badp = str(tmp_path / "bad.snp")
    open(badp, "wb").write(bytes(bad))
    with pkg_mod.SnapshotReader(badp) as r:
        with pytest.raises(pkg_mod.LccrfError, match="checksum"):
            r.read(0)
        _same(r.read(3), frames[3])
    # not a snapshot file at all
    junk = str(tmp_path / "junk.snp")
    open(junk, "wb").write(b"P6\n640 480\n255\n" + bytes(100))
    with pytest.raises(pkg_mod.LccrfError):
        pkg_mod.SnapshotReader(junk)
    with pytest.raises(pkg_mod.LccrfError):
        pkg_mod.SnapshotReader(str(tmp_path / "missing.snp"))


def test_writer_validates_the_csr(tmp_path):
    f = _frames()[0]
    with pkg_mod.SnapshotWriter(str(tmp_path / "v.snp")) as w:
        class Bad:
            pass
        b = Bad()
        for k in FIELDS:
            setattr(b, k, getattr(f, k).copy())
        b.obs_kf[5] = f.kf_pose.shape[0]  # out of range
        with pytest.raises(pkg_mod.LccrfError, match="obs_kf"):
            w.write(b)
        b.obs_kf[5] = 0
        b.obs_ptr[3] = b.obs_ptr[2] - 1  # decreasing
        with pytest.raises(pkg_mod.LccrfError, match="obs_ptr"):
            w.write(b)


def test_concat_frames_rebases(tmp_path):
    frames = _frames()
    cat = pkg_mod.concat_frames(frames)
    assert cat["kf_ptr"].tolist() == [0, 40, 44, 48, 348]
    assert cat["obs_ptr"][-1] == sum(f.nnz for f in frames) and cat["xyz"].shape[0] == sum(f.n for f in frames)
    o, e = 0, 0
    for b, f in enumerate(frames):
        assert np.array_equal(cat["obs_ptr"][o:o + f.n + 1] - e, f.obs_ptr)
        assert np.array_equal(cat["obs_kf"][e:e + f.nnz] - cat["kf_ptr"][b], f.obs_kf)
        o += f.n
        e += f.nnz


@pytest.mark.gpu
def test_replay_file_through_device_path(tmp_path, ctx, oracle):
    """A sequence written to disk and replayed (read -> batch -> submit_map with uint16 keyframe indices -> label
    application) gives, frame by frame, the oracle's marginals / MAP / moving-point lists."""
    from oracle.pyoracle import slam_params
    from util import assert_bit_exact
    prm_o, prm = slam_params(**synth.SLAM_PARAMS), pkg_mod.SlamParams.make()
    en = pkg_mod.label_energies(2, prm.confidence)
    seq = [synth.map_snapshot(n, o, seed=70 + i, n_kf=64, ragged=True) for i, (n, o) in enumerate(((1500, 20), (1800, 32), (900, 8), (2100, 16)))]
    rng = np.random.default_rng(1)
    fids = [rng.permutation(3000)[:f.n].astype(np.int32) for f in seq]
    path = str(tmp_path / "replay.snp")
    with pkg_mod.SnapshotWriter(path) as w:
        for i, f in enumerate(seq):
            w.write(f, frame_id=i, timestamp=i / 30.0, fid=fids[i])
    with pkg_mod.SnapshotReader(path) as r:
        got = [r.read(i, kf_dtype=np.uint16) for i in range(len(r))]
    cat = pkg_mod.concat_frames(got)
    F = pkg_mod.Frames(ctx, [g.n for g in got], prm, en)
    NT = sum(g.n for g in got)
    mp, pr = np.empty(NT, np.int16), np.empty((NT, 2), np.float32)
    kf16 = cat["obs_kf"].astype(np.uint16)  # host buffers stay alive until wait()
    F.submit_map(0, cat["xyz"], cat["obs_ptr"], kf16, cat["obs_uv"], cat["kf_pose"], cat["kf_intr"],
                 cat["kf_bounds"], cat["kp2d"], cat["kf_ptr"], mp, pr)
    F.wait(0)
    dp, dl, sp, sl = F.partition(cat["fid"])
    o = 0
    for b, f in enumerate(seq):
        ob, er, de = oracle.map_point_unary(f)
        lab = oracle.rough_classify(ob, er, de, prm_o)
        Qo, mo, _ = oracle.slam_crf(ob, er, f.kp2d, lab, en, prm_o)
        assert_bit_exact(pr[o:o + f.n], Qo, what="frame %d" % b)
        assert np.array_equal(mp[o:o + f.n], mo)
        d_ref, s_ref = oracle.label_partition(mo, fids[b])
        assert np.array_equal(dl[dp[b]:dp[b + 1]], d_ref) and np.array_equal(sl[sp[b]:sp[b + 1]], s_ref)
        o += f.n
    F.close()

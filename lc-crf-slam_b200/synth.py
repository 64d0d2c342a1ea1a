"""Seeded synthetic inputs of the BASELINE.json shapes.

Host-side plumbing shared by tests/ and bench.py.  Parameters follow the reference's
Examples/RGB-D/TUM3.yaml:8-11,81-101 (BONN.yaml:8-11 for the Bonn-shaped intrinsics).
Pure numpy; no oracle and no CUDA here.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# TUM3.yaml:81-101
SLAM_PARAMS = dict(
    w1=10.0, w2=30.0,
    u_alpha=1.7, stdev_alpha=0.6,
    u_beta=5.4, stdev_beta=1.5,
    u_gamma=0.3, stdev_gamma=0.2,
    point3d_stdev=0.5, point2d_stdev=18.0,
    u_depth=2.75, pth=0.8, confidence=0.7,
    iters=5,
)
TUM_INTR = (535.4, 539.2, 320.1, 247.6)            # TUM3.yaml:8-11
BONN_INTR = (542.822841, 542.576870, 315.593520, 237.756098)  # BONN.yaml:8-11
IMG_W, IMG_H = 640, 480


@dataclasses.dataclass
class SlamFrame:
    """Flat per-frame CRF inputs, the vectors gathered at src/Tracking.cc:1849-1870."""
    observs: np.ndarray   # [N] f32  (vobservs)
    error: np.ndarray     # [N] f32  (verrors)
    depth: np.ndarray     # [N] f32  (vdepths)
    kp2d: np.ndarray      # [N,2] f32 (vcorrd2d)
    dynamic: np.ndarray   # [N] bool ground truth of the generator (not an input)

    @property
    def n(self) -> int:
        return int(self.observs.shape[0])


def _dynamic_mask(rng: np.random.Generator, xy: np.ndarray, frac: float, blobs: int = 1) -> np.ndarray:
    """Elliptic 'moving object' regions covering roughly `frac` of the image."""
    mask = np.zeros(xy.shape[0], dtype=bool)
    area = frac * IMG_W * IMG_H / blobs
    for _ in range(blobs):
        cx, cy = rng.uniform(120, IMG_W - 120), rng.uniform(100, IMG_H - 100)
        aspect = rng.uniform(0.6, 1.6)
        ry = np.sqrt(area / (np.pi * aspect))
        rx = aspect * ry
        mask |= ((xy[:, 0] - cx) / rx) ** 2 + ((xy[:, 1] - cy) / ry) ** 2 <= 1.0
    return mask


def slam_frame(n: int, seed: int, dyn_frac: float = 0.2, blobs: int = 1) -> SlamFrame:
    """C1 / C4 unit problem: the (observs, error, depth, keypoint) vectors of one frame."""
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(0, IMG_W, n), rng.uniform(0, IMG_H, n)], axis=1)
    dyn = _dynamic_mask(rng, xy, dyn_frac, blobs)
    # static points: long tracks (observation count around u_beta), reprojection residual around
    # u_alpha, depth around u_depth; dynamic points: short tracks, 4x residual, any depth.
    observs = np.clip(np.rint(rng.normal(5.4, 2.5, n)), 1, 14)
    observs[dyn] = rng.integers(1, 5, int(dyn.sum()))
    error = np.abs(rng.normal(1.3, 0.7, n))
    error[dyn] = np.abs(rng.normal(0, 1, int(dyn.sum()))) * 3.2 + 2.0
    depth = np.clip(rng.normal(2.75, 0.8, n), 0.5, 5.0)
    depth[dyn] = rng.uniform(0.5, 5.0, int(dyn.sum()))
    return SlamFrame(observs.astype(np.float32), error.astype(np.float32), depth.astype(np.float32),
                     xy.astype(np.float32), dyn)


@dataclasses.dataclass
class MapSnapshot:
    """Flat SoA restatement of the pointer graph read by Tracking::ComputeMapPointErrAndObserv
    (src/Tracking.cc:1803-1839): map points, their keyframe observations (CSR), keyframes."""
    xyz: np.ndarray        # [N,3] f32   MapPoint::mWorldPos
    obs_ptr: np.ndarray    # [N+1] i32   CSR over observations
    obs_kf: np.ndarray     # [nnz] i32   keyframe index of each observation
    obs_uv: np.ndarray     # [nnz,2] f32 KeyFrame::mvKeysUn[fid].pt
    kf_pose: np.ndarray    # [nKF,12] f32 rows of [Rcw|tcw]
    kf_intr: np.ndarray    # [nKF,4] f32 fx fy cx cy
    kf_bounds: np.ndarray  # [nKF,4] f32 mnMinX mnMaxX mnMinY mnMaxY
    kp2d: np.ndarray       # [N,2] f32   current-frame keypoints (mCurrentFrame.mvKeysUn[i].pt)
    dynamic: np.ndarray    # [N] bool

    @property
    def n(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def nnz(self) -> int:
        return int(self.obs_kf.shape[0])


def _rot(rx: float, ry: float, rz: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def map_snapshot(n: int, obs_per_point: int, seed: int, n_kf: int = 256, intr=TUM_INTR,
                 dyn_frac: float = 0.2, ragged: bool = False, bad_frac: float = 0.05, unique_kf: bool = False) -> MapSnapshot:
    """C3: N map points x `obs_per_point` keyframe observations on a smooth trajectory.
    ~bad_frac of the observations are deliberately behind the camera / out of bounds to
    exercise the skip rules (Tracking.cc:1823,1828).
    unique_kf: every point is observed at most once per keyframe, as in the reference's std::map<KeyFrame*, size_t>
    (MapPoint.h:115) -- what a device-resident map filled through AddObservation / EraseObservation needs."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = intr
    # keyframes: small smooth motion around the origin, looking down +z
    t = np.linspace(0, 1, n_kf)
    pose = np.zeros((n_kf, 12), dtype=np.float64)
    for k in range(n_kf):
        R = _rot(0.05 * np.sin(2 * np.pi * t[k]), 0.08 * np.sin(2 * np.pi * t[k] + 1.0), 0.03 * np.cos(2 * np.pi * t[k]))
        c = np.array([0.3 * np.sin(2 * np.pi * t[k]), 0.1 * np.cos(2 * np.pi * t[k]), 0.2 * t[k]])
        tcw = -R @ c
        pose[k] = np.concatenate([R, tcw[:, None]], axis=1).reshape(-1)
    # current-frame keypoints and depths -> world points (current frame = identity pose)
    xy = np.stack([rng.uniform(20, IMG_W - 20, n), rng.uniform(20, IMG_H - 20, n)], axis=1)
    z = np.clip(rng.normal(2.75, 0.45, n), 1.0, 5.0)  # around u_depth (TUM3.yaml:98)
    xyz = np.stack([(xy[:, 0] - cx) / fx * z, (xy[:, 1] - cy) / fy * z, z], axis=1)
    dyn = _dynamic_mask(rng, xy, dyn_frac)
    if ragged:
        cnt = rng.integers(1, obs_per_point + 1, n)
        cnt[dyn] = np.minimum(cnt[dyn], rng.integers(1, 5, int(dyn.sum())))
    else:
        cnt = np.full(n, obs_per_point)
    obs_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=obs_ptr[1:])
    nnz = int(obs_ptr[-1])
    pt = np.repeat(np.arange(n), cnt)
    if unique_kf:
        # observation j of point p sits in keyframe perm[row_p][(j + off_p) % M]: distinct keyframes within a point
        M = n_kf - 2 if (bad_frac > 0 and n_kf >= 4) else n_kf
        assert obs_per_point <= M, "more observations per point than keyframes"
        perms = np.stack([rng.permutation(M) for _ in range(257)])
        j = np.arange(nnz) - obs_ptr[:-1][pt]
        obs_kf = perms[rng.integers(0, 257, n)[pt], (j + rng.integers(0, M, n)[pt]) % M]
    else:
        obs_kf = rng.integers(0, n_kf, nnz)
    P = pose[obs_kf].reshape(nnz, 3, 4)
    X = xyz[pt]
    Xc = np.einsum('nij,nj->ni', P[:, :, :3], X) + P[:, :, 3]
    u = fx * Xc[:, 0] / Xc[:, 2] + cx
    v = fy * Xc[:, 1] / Xc[:, 2] + cy
    uv = np.stack([u, v], axis=1) + rng.normal(0, 1.25, (nnz, 2))  # residual ~ u_alpha for static points
    d = dyn[pt]
    drift = rng.uniform(5, 20, nnz) * d
    ang = rng.uniform(0, 2 * np.pi, nnz)
    uv += np.stack([drift * np.cos(ang), drift * np.sin(ang)], axis=1)
    # skip-rule exercisers: move the world point of a few observations' keyframe far away is not
    # possible per observation, so instead give those observations a keyframe looking backwards.
    xyz32 = xyz.astype(np.float32)
    kf_pose = pose.astype(np.float32)
    if bad_frac > 0 and n_kf >= 4:
        # the last two keyframes are "bad": one looks backwards (z<0), one is far off-axis (out of bounds)
        Rb = _rot(0, np.pi, 0)
        kf_pose[n_kf - 1] = np.concatenate([Rb, np.zeros((3, 1))], axis=1).reshape(-1).astype(np.float32)
        Ro = _rot(0, 1.2, 0)
        kf_pose[n_kf - 2] = np.concatenate([Ro, np.zeros((3, 1))], axis=1).reshape(-1).astype(np.float32)
        bad = rng.random(nnz) < bad_frac
        if unique_kf:
            # at most one observation per point in each of the two bad keyframes: the first two marked ones
            cs = np.cumsum(bad)
            rank = cs - (cs[obs_ptr[:-1]] - bad[obs_ptr[:-1]])[pt]   # 1-based rank of a marked observation inside its point
            obs_kf = np.where(bad & (rank == 1), n_kf - 1, np.where(bad & (rank == 2), n_kf - 2, obs_kf))
        else:
            obs_kf = np.where(bad, rng.integers(n_kf - 2, n_kf, nnz), np.minimum(obs_kf, n_kf - 3))
    kf_intr = np.tile(np.array(intr, dtype=np.float32), (n_kf, 1))
    kf_bounds = np.tile(np.array([0, IMG_W, 0, IMG_H], dtype=np.float32), (n_kf, 1))
    return MapSnapshot(xyz32, obs_ptr.astype(np.int32), obs_kf.astype(np.int32), uv.astype(np.float32),
                       kf_pose, kf_intr, kf_bounds, xy.astype(np.float32), dyn)


def index_observations(obs_kf, obs_uv, n_kf: int, seed: int, stride: int | None = None, stride_max: int = 32768):
    """Restate a snapshot's observations the way the reference stores them: (keyframe, feature index) pairs
    (MapPoint::mObservations, MapPoint.h:115) plus one keypoint array per keyframe (KeyFrame::mvKeysUn).
    Returns (obs_fid int32 [nnz], kp_table float32 [n_kf][stride][2], obs_uv_consistent float32 [nnz][2]).
    Feature indices are scattered over the keyframe's row (a multiplicative permutation with a random per-keyframe
    offset), i.e. uncorrelated with map-point order, as in a real keyframe.  A keyframe referenced by more than
    `stride` observations re-uses keypoints (first occurrence defines the keypoint); obs_uv_consistent = kp_table[kf, fid]
    is the flat restatement of exactly the same problem (equal to obs_uv wherever no keypoint is re-used)."""
    obs_kf = np.asarray(obs_kf, dtype=np.int64)
    obs_uv = np.asarray(obs_uv, dtype=np.float32)
    nnz = obs_kf.size
    cnt = np.bincount(obs_kf, minlength=n_kf)
    if stride is None:
        need = int(min(cnt.max() if nnz else 1, stride_max))
        stride = 1 << max(int(np.ceil(np.log2(max(need, 1)))), 0)
    assert stride & (stride - 1) == 0, "stride must be a power of two (the permutation is multiplicative mod stride)"
    order = np.argsort(obs_kf, kind="stable")
    start = np.zeros(n_kf + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    rank = np.empty(nnz, dtype=np.int64)
    rank[order] = np.arange(nnz) - start[obs_kf[order]]
    rng = np.random.default_rng(seed)
    off = rng.integers(0, stride, n_kf)
    fid = ((rank % stride) * 40503 + off[obs_kf]) % stride  # 40503 is odd: a bijection modulo a power of two
    table = np.zeros((n_kf, stride, 2), dtype=np.float32)
    table[obs_kf[::-1], fid[::-1]] = obs_uv[::-1]            # last write wins -> reversed: the first occurrence wins
    return fid.astype(np.int32), table, table[obs_kf, fid]


def image_problem(w: int, h: int, seed: int, unknown_frac: float = 0.7, n_labels: int = 2):
    """C2: smooth blobs + N(0,10) noise uint8 RGB image, noisy fg/bg labels, `unknown_frac` unknown (-1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    fg = np.zeros((h, w), dtype=bool)
    for _ in range(3):
        cx, cy = rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h)
        rx, ry = rng.uniform(0.1 * w, 0.25 * w), rng.uniform(0.1 * h, 0.25 * h)
        fg |= ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1
    base = np.where(fg[..., None], np.array([200, 80, 60]), np.array([60, 120, 180])).astype(np.float64)
    base += 25 * np.sin(xx / 37.0)[..., None] + 15 * np.cos(yy / 23.0)[..., None]
    img = np.clip(base + rng.normal(0, 10, (h, w, 3)), 0, 255).astype(np.uint8)
    lab = fg.astype(np.int16) % n_labels
    flip = rng.random((h, w)) < 0.1
    lab = np.where(flip, (lab + 1) % n_labels, lab).astype(np.int16)
    lab[rng.random((h, w)) < unknown_frac] = -1
    return img.reshape(-1, 3).copy(), lab.reshape(-1).copy()


def orb_frame_pair(nq: int, nt: int, seed: int, match_frac: float = 0.5, flip_bits: int = 20, entropy_bytes: int = 32):
    """Two frames' ORB descriptor sets (256 bit, rows of a CV_8U x 32 matrix) for BfMatch (Tracking.cc:1747-1766):
    `match_frac` of the queries are noisy copies (up to `flip_bits` flipped bits) of random train rows, the rest are
    unrelated.  `entropy_bytes` < 32 zeroes the tail of every descriptor, which makes distance ties common."""
    rng = np.random.default_rng(seed)
    dt = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    dq = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    if nt > 0 and nq > 0:
        sel = np.nonzero(rng.random(nq) < match_frac)[0]
        src = rng.integers(0, nt, sel.size)
        dq[sel] = dt[src]
        for i in sel:
            nb = int(rng.integers(0, flip_bits + 1))
            pos = rng.integers(0, 256, nb)
            np.bitwise_xor.at(dq[i], pos >> 3, (1 << (pos & 7)).astype(np.uint8))
    if entropy_bytes < 32:
        dt[:, entropy_bytes:] = 0
        dq[:, entropy_bytes:] = 0
    return dq, dt


def epipolar_matches(m: int, seed: int, outlier_frac: float = 0.3):
    """Matched keypoints of two views of static 3-D points plus the fundamental matrix of the pair (double, row-major)
    for GetFeature2EpipolarDis (Tracking.cc:2030-2047); `outlier_frac` of the matches are displaced (moving points)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = TUM_INTR
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    # consecutive frames: small inter-frame motion (the reference evaluates the second distance with y1 in place of
    # y2, fundamental_estimator.h:121, which is only harmless when the vertical parallax is small)
    R = _rot(*(rng.normal(0, 0.004, 3)))
    t = rng.normal(0, 0.01, 3)
    X = np.stack([rng.uniform(-1.5, 1.5, m), rng.uniform(-1.0, 1.0, m), rng.uniform(1.0, 5.0, m)], axis=1)
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    X2 = (R @ X.T).T + t
    x2 = (K @ X2.T).T
    x2 = x2[:, :2] / x2[:, 2:]
    x1 += rng.normal(0, 0.5, x1.shape)
    x2 += rng.normal(0, 0.5, x2.shape)
    out = rng.random(m) < outlier_frac
    x2[out] += rng.normal(0, 15.0, (int(out.sum()), 2))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Kinv = np.linalg.inv(K)
    F = Kinv.T @ tx @ R @ Kinv
    F = F / np.linalg.norm(F)
    return x1.astype(np.float32), x2.astype(np.float32), np.ascontiguousarray(F, dtype=np.float64), out


class SequenceReplay:
    """C5: one synthetic RGB-D sequence replayed as batches of frame CRFs against a map that changes as it would in
    LC-CRF-SLAM: a keyframe is inserted every `kf_interval` frames (KeyFrame ctor + MapPoint::AddObservation for the
    points it sees), the oldest keyframe is culled once more than `kf_window` are alive (MapPoint::EraseObservation in
    every point that holds it), a few points go bad per batch (SetBadFlag), and bundle adjustment moves all poses and
    positions a little.  The host model is a boolean point x keyframe matrix; everything is seeded and vectorised.

    kind 'tum': TUM3 intrinsics (TUM3.yaml:8-11), one moving blob (~25% of the points);
    kind 'bonn': Bonn intrinsics (BONN.yaml:8-11), two larger blobs (~35%)."""

    def __init__(self, seed: int, kind: str = "tum", n_points: int = 40000, frames_per_batch: int = 64,
                 kf_interval: int = 8, kf_window: int = 24, stride: int = 8192, n_frame_points=(4000, 6000), kf0: int = 16,
                 kf_obs: int = 6000):
        self.rng = np.random.default_rng(seed)
        rng = self.rng
        self.kind, self.P, self.FB, self.kfi, self.win, self.stride = kind, n_points, frames_per_batch, kf_interval, kf_window, stride
        self.kf_obs = min(kf_obs, stride)
        self.intr = np.array(TUM_INTR if kind == "tum" else BONN_INTR, np.float32)
        self.bounds = np.array([0, IMG_W, 0, IMG_H], np.float32)
        self.xyz = np.stack([rng.uniform(-2.6, 2.6, n_points), rng.uniform(-1.9, 1.9, n_points), rng.uniform(2.0, 5.5, n_points)], 1).astype(np.float32)
        blobs = 1 if kind == "tum" else 2
        self.dynamic = np.zeros(n_points, bool)
        for _ in range(blobs):
            c = np.array([rng.uniform(-1.0, 1.0), rng.uniform(-0.6, 0.6), rng.uniform(3.0, 4.5)])
            r = np.array([1.75, 1.4, 1.5]) if kind == "tum" else np.array([1.7, 1.35, 1.5])
            self.dynamic |= (((self.xyz - c) / r) ** 2).sum(1) <= 1.0
        self.sizes = rng.integers(n_frame_points[0], n_frame_points[1] + 1, frames_per_batch).tolist()
        self.frame = 0
        self.n_kf = 0
        self.has = np.zeros((n_points, 0), bool)        # has[p, k]: point p holds an observation in keyframe k
        self.fid_of = np.zeros((n_points, 0), np.int32)  # its feature index there
        self.kp_rows = []                                # keypoint row of every keyframe
        self.kf_pose = np.zeros((0, 12), np.float32)     # pose of every keyframe as last sent to the map
        self.alive = []                                  # keyframes not yet culled, oldest first
        self.bad = np.zeros(n_points, bool)
        # initial map: kf0 keyframes before frame 0, loaded in bulk by the caller (initial_map)
        self._init = self._make_keyframes(range(-kf0 * kf_interval, 0, kf_interval))

    # ---- camera: a smooth hand-held path looking down +z
    def pose_of(self, f: float) -> np.ndarray:
        t = f / 400.0
        R = _rot(0.06 * np.sin(2 * np.pi * t), 0.09 * np.sin(2 * np.pi * t + 1.0), 0.03 * np.cos(2 * np.pi * t))
        c = np.array([0.5 * np.sin(2 * np.pi * t), 0.15 * np.cos(2 * np.pi * t), 0.3 * np.sin(np.pi * t)])
        return np.concatenate([R, (-R @ c)[:, None]], axis=1).reshape(-1).astype(np.float32)

    def project(self, pose12: np.ndarray, ids: np.ndarray):
        P = pose12.reshape(3, 4).astype(np.float64)
        Xc = self.xyz[ids].astype(np.float64) @ P[:, :3].T + P[:, 3]
        fx, fy, cx, cy = self.intr.astype(np.float64)
        z = np.maximum(Xc[:, 2], 1e-6)
        return np.stack([fx * Xc[:, 0] / z + cx, fy * Xc[:, 1] / z + cy], 1), Xc[:, 2]

    def in_view(self, pose12: np.ndarray) -> np.ndarray:
        uv, z = self.project(pose12, np.arange(self.P))
        return np.nonzero((z > 0.3) & (uv[:, 0] > 4) & (uv[:, 0] < IMG_W - 4) & (uv[:, 1] > 4) & (uv[:, 1] < IMG_H - 4) & ~self.bad)[0]

    def _observe(self, pose12, ids, noise=1.0):
        """keypoints of the given points in a camera: projection + noise (moving points drift 5-20 px)"""
        uv, _ = self.project(pose12, ids)
        uv = uv + self.rng.normal(0, noise, uv.shape)
        d = self.dynamic[ids]
        drift = self.rng.uniform(5, 20, ids.size) * d
        ang = self.rng.uniform(0, 2 * np.pi, ids.size)
        return (uv + np.stack([drift * np.cos(ang), drift * np.sin(ang)], 1)).astype(np.float32)

    def _make_keyframes(self, frames):
        """new keyframes at the given frame numbers: each observes up to kf_obs of the points in its view"""
        out = dict(pose=[], kp=[], pt=[], kf=[], fid=[], seg=[0])
        for f in frames:
            pose = self.pose_of(f)
            vis = self.in_view(pose)
            vis = np.sort(vis[self.rng.permutation(vis.size)[: self.kf_obs]])
            fid = self.rng.permutation(self.stride)[: vis.size]
            row = np.zeros((self.stride, 2), np.float32)
            row[fid] = self._observe(pose, vis)
            k = self.n_kf
            self.n_kf += 1
            self.has = np.concatenate([self.has, np.zeros((self.P, 1), bool)], axis=1)
            self.fid_of = np.concatenate([self.fid_of, np.zeros((self.P, 1), np.int32)], axis=1)
            self.has[vis, k] = True
            self.fid_of[vis, k] = fid
            self.kp_rows.append(row)
            self.kf_pose = np.concatenate([self.kf_pose, pose[None]])
            self.alive.append(k)
            out["pose"].append(pose)
            out["kp"].append(row)
            out["pt"].append(vis.astype(np.int32))
            out["kf"].append(np.full(vis.size, k, np.int32))
            out["fid"].append(fid.astype(np.int32))
            out["seg"].append(out["seg"][-1] + vis.size)
        n = len(out["pose"])
        cat = lambda k, dt: np.concatenate(out[k]).astype(dt) if n else np.zeros(0, dt)
        return dict(first=self.n_kf - n, pose=np.stack(out["pose"]) if n else np.zeros((0, 12), np.float32),
                    intr=np.tile(self.intr, (n, 1)), bounds=np.tile(self.bounds, (n, 1)),
                    kp=np.stack(out["kp"]) if n else np.zeros((0, self.stride, 2), np.float32),
                    pt=cat("pt", np.int32), kf=cat("kf", np.int32), fid=cat("fid", np.int32), seg=np.asarray(out["seg"], np.int32))

    def initial_map(self):
        """(keyframes dict, xyz, obs_ptr, obs_ref): the bulk load of the map as it is before frame 0"""
        kfs = self._init
        order = np.argsort(kfs["pt"], kind="stable")   # per point, keyframes in insertion order
        ptr = np.zeros(self.P + 1, np.int32)
        np.cumsum(np.bincount(kfs["pt"], minlength=self.P), out=ptr[1:])
        ref = np.stack([kfs["kf"][order], kfs["fid"][order]], 1).astype(np.int32)
        return kfs, self.xyz.copy(), ptr, ref

    def snapshot(self, ids, kp2d) -> MapSnapshot:
        """the model's current state for the given frame points in the lccrf_frames_set_map_inputs layout
        (observations of a point in keyframe insertion order, which is the order the map appends them in)"""
        ids = np.asarray(ids)
        rows, kfi = np.nonzero(self.has[ids])
        ptr = np.zeros(ids.size + 1, np.int32)
        np.cumsum(np.bincount(rows, minlength=ids.size), out=ptr[1:])
        kp = np.stack(self.kp_rows)
        uv = kp[kfi, self.fid_of[ids[rows], kfi]]
        n = self.n_kf
        return MapSnapshot(self.xyz[ids].copy(), ptr, kfi.astype(np.int32), uv.astype(np.float32), self.kf_pose.copy(),
                           np.tile(self.intr, (n, 1)), np.tile(self.bounds, (n, 1)), np.asarray(kp2d, np.float32),
                           self.dynamic[ids].copy())

    def next_batch(self):
        """frames [frame, frame + FB): returns dict(ids [NT] int32, kp2d [NT,2], delta kwargs for MapDelta.make)"""
        rng = self.rng
        f0 = self.frame
        self.frame += self.FB
        # map changes of this span, in the delta's order: new keyframes, poses, positions, erase, bad, add
        kfs = self._make_keyframes([f for f in range(f0, f0 + self.FB) if f % self.kfi == 0])
        er_pt, er_kf, er_seg = [], [], [0]
        while len(self.alive) > self.win:                  # culling of the oldest keyframes
            k = self.alive.pop(0)
            pts = np.nonzero(self.has[:, k])[0]
            self.has[pts, k] = False
            er_pt.append(pts.astype(np.int32))
            er_kf.append(np.full(pts.size, k, np.int32))
            er_seg.append(er_seg[-1] + pts.size)
        cand = np.nonzero(self.has.any(1) & ~self.bad)[0]
        bad = cand[rng.random(cand.size) < 0.004]          # SetBadFlag: the point loses all observations
        self.bad[bad] = True
        self.has[bad] = False
        # bundle adjustment: every pose and every position moves a little
        self.xyz = (self.xyz + rng.normal(0, 5e-4, self.xyz.shape)).astype(np.float32)
        poses = np.stack([self.pose_of(-16 * self.kfi + k * self.kfi) for k in range(self.n_kf)]) if self.n_kf else np.zeros((0, 12), np.float32)
        poses = (poses + rng.normal(0, 2e-4, poses.shape)).astype(np.float32)
        self.kf_pose[: kfs["first"]] = poses[: kfs["first"]]   # the new keyframes keep the pose they were created with
        # the frames of the batch: visible points = points in view that still hold an observation
        live = self.has.any(1) & ~self.bad
        ids, kp2d = [], []
        for j in range(self.FB):
            pose = self.pose_of(f0 + j + 0.37)             # tracked frames lie between keyframes
            vis = self.in_view(pose)
            vis = vis[live[vis]]
            n = self.sizes[j]
            vis = vis[rng.permutation(vis.size)[:n]] if vis.size >= n else rng.choice(np.nonzero(live)[0], n, replace=False)
            ids.append(vis.astype(np.int32))
            kp2d.append(self._observe(pose, vis, noise=0.8))
        delta = dict(kf_first=kfs["first"], kf_pose=kfs["pose"] if kfs["pose"].size else None,
                     kf_intr=kfs["intr"] if kfs["pose"].size else None, kf_bounds=kfs["bounds"] if kfs["pose"].size else None,
                     kf_keypoints=kfs["kp"] if kfs["pose"].size else None,
                     pose=poses[: kfs["first"]] if kfs["first"] > 0 else None, xyz=self.xyz.copy(),
                     erase_pt=np.concatenate(er_pt) if er_pt else None, erase_kf=np.concatenate(er_kf) if er_kf else None,
                     erase_seg_ptr=np.asarray(er_seg, np.int32) if er_pt else None, bad_pt=bad.astype(np.int32) if bad.size else None,
                     add_pt=kfs["pt"] if kfs["pt"].size else None, add_kf=kfs["kf"] if kfs["pt"].size else None,
                     add_fid=kfs["fid"] if kfs["pt"].size else None, add_seg_ptr=kfs["seg"] if kfs["pt"].size else None)
        return dict(ids=np.concatenate(ids), kp2d=np.concatenate(kp2d), delta=delta)

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
                 dyn_frac: float = 0.2, ragged: bool = False, bad_frac: float = 0.05) -> MapSnapshot:
    """C3: N map points x `obs_per_point` keyframe observations on a smooth trajectory.
    ~bad_frac of the observations are deliberately behind the camera / out of bounds to
    exercise the skip rules (Tracking.cc:1823,1828)."""
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

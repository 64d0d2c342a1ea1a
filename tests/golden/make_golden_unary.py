"""Regenerates tests/golden/golden_unary.npz (run in the build container; needs the cv2 wheel, which is absent on the
GPU box).  Pins the restatement of Tracking::ComputeMapPointErrAndObserv (src/Tracking.cc:1803-1839) against a
statement-by-statement evaluation with the reference's own third-party arithmetic: `Rcw * x3Dw + tcw` (:1818) is an
OpenCV MatExpr that evaluates as cv::gemm(Rcw, x3Dw, 1, tcw, 1) -- computed here by the real cv2.gemm on CV_32F
matrices (cv2 4.13; the reference links OpenCV 3.x, source not under /root/reference) -- and every scalar statement
is evaluated in the C++ type the reference declares (float products and sums rounded one by one, `1.0 / z` and the
square root in double, `error /= observs` as a float division by the converted int).  Observation order = CSR order
(the reference walks a std::map<KeyFrame*, size_t>, i.e. pointer order: not reproducible).
Stored per case: the snapshot arrays and observs / error / depth."""
import importlib
import math
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
synth = importlib.import_module("lc-crf-slam_b200.synth")
f32 = np.float32

CASES = {  # name: (N, observations per point, seed, n_kf, ragged, mixed cameras)
    "ragged": (400, 12, 1, 40, True, False),
    "uniform64": (64, 64, 2, 256, False, False),     # the C3 shape: all points of a warp have 64 observations
    "uniform70": (33, 70, 3, 64, False, False),      # uniform but not a multiple of 16
    "mixed_cameras": (150, 9, 4, 24, True, True),    # keyframes with different intrinsics and image bounds
}


def unary_cv2(s):
    n = s.n
    observs, error_out, depth_out = np.zeros(n, f32), np.zeros(n, f32), np.zeros(n, f32)
    with np.errstate(all="ignore"):
        for i in range(n):
            a, b = int(s.obs_ptr[i]), int(s.obs_ptr[i + 1])
            nobs = b - a                                                    # observs = observations.size()   :1808
            observs[i] = nobs
            if nobs == 0:                                                   # :1809
                continue
            x3Dw = np.ascontiguousarray(s.xyz[i].reshape(3, 1), dtype=f32)  # :1812
            error, depth = f32(0), f32(0)
            for o in range(a, b):
                k = int(s.obs_kf[o])
                T = s.kf_pose[k].reshape(3, 4)
                Rcw = np.ascontiguousarray(T[:, :3], dtype=f32)             # :1816
                tcw = np.ascontiguousarray(T[:, 3:4], dtype=f32)            # :1817
                x3Dc = cv2.gemm(Rcw, x3Dw, 1.0, tcw, 1.0)                   # Rcw * x3Dw + tcw   :1818
                assert x3Dc.dtype == f32
                xc, yc, zc = f32(x3Dc[0, 0]), f32(x3Dc[1, 0]), f32(x3Dc[2, 0])
                invzc = f32(np.float64(1.0) / np.float64(zc))               # float invzc = 1.0 / z   :1821
                if invzc < 0:                                               # :1823
                    continue
                fx, fy, cx, cy = (f32(v) for v in s.kf_intr[k])
                u = f32(f32(f32(fx * xc) * invzc) + cx)                     # :1825
                v = f32(f32(f32(fy * yc) * invzc) + cy)                     # :1826
                mnx, mxx, mny, mxy = (f32(int(t)) for t in s.kf_bounds[k])  # int members compared with a float
                if u < mnx or u > mxx or v < mny or v > mxy:                # :1828
                    continue
                kx, ky = float(s.obs_uv[o, 0]), float(s.obs_uv[o, 1])       # Point2d kp = mvKeysUn[fid].pt   :1832
                du, dv = float(u) - kx, float(v) - ky
                error_ = f32(math.sqrt(du * du + dv * dv))                  # :1833
                error = f32(error + error_)                                 # :1834
                depth = f32(depth + zc)                                     # :1835
            error_out[i] = f32(error / f32(nobs))                           # :1837
            depth_out[i] = f32(depth / f32(nobs))                           # :1838
    return observs, error_out, depth_out


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (n, obs, seed, n_kf, ragged, mixed) in CASES.items():
        s = synth.map_snapshot(n, obs, seed=seed, n_kf=n_kf, ragged=ragged)
        if mixed:
            s.kf_intr = s.kf_intr.copy()
            s.kf_intr[::3, 0] *= f32(1.013)
            s.kf_intr[1::3, 3] += f32(2.5)
            s.kf_bounds = s.kf_bounds.copy()
            s.kf_bounds[1::4, 1] -= f32(37)
            s.kf_bounds[2::5, 2] += f32(21)
        ob, er, de = unary_cv2(s)
        for k in ("xyz", "obs_ptr", "obs_kf", "obs_uv", "kf_pose", "kf_intr", "kf_bounds", "kp2d"):
            out[name + "_" + k] = getattr(s, k)
        out[name + "_observs"], out[name + "_error"], out[name + "_depth"] = ob, er, de
        skipped = int((er == 0).sum())
        print(name, "points", n, "observations", s.nnz, "mean error %.3f" % float(er.mean()), "zero-error points", skipped)
    np.savez_compressed(os.path.join(HERE, "golden_unary.npz"), **out)


if __name__ == "__main__":
    main()

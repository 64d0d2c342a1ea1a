"""ctypes loader for the CPU ORACLE (test infrastructure, NOT product code).

  liboracle.so      plain-C restatement (oracle/lccrf_oracle.c)          -> class Oracle
  _ref/libref.so    the reference headers compiled in place (ref_shim.cpp) -> class Ref (may be absent)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_short_p = C.POINTER(C.c_short)


def build(quiet: bool = True) -> None:
    """make liboracle.so (+ _ref/ when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _sp(a):
    return a.ctypes.data_as(c_short_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _SlamParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "w1", "w2", "u_alpha", "stdev_alpha", "u_beta", "stdev_beta", "u_gamma", "stdev_gamma",
        "point3d_stdev", "point2d_stdev", "u_depth", "pth", "confidence")] + [("iters", C.c_int)]


class _Lattice(C.Structure):
    _fields_ = [("N", C.c_int), ("d", C.c_int), ("V", C.c_int), ("offset", c_int_p),
                ("bary", c_float_p), ("nbr", c_int_p), ("keys", c_short_p)]


def slam_params(**kw) -> _SlamParams:
    p = _SlamParams()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.orc_lattice_init.restype = C.POINTER(_Lattice)
        L.orc_lattice_init.argtypes = [c_float_p, C.c_int, C.c_int]
        L.orc_lattice_free.argtypes = [C.POINTER(_Lattice)]
        L.orc_lattice_filter.argtypes = [C.POINTER(_Lattice), c_float_p, c_float_p, C.c_int]
        L.orc_potts_norm.argtypes = [C.POINTER(_Lattice), c_float_p]
        L.orc_fast_exp.restype = C.c_float
        L.orc_fast_exp.argtypes = [C.c_float]
        L.orc_exp_and_normalize.argtypes = [c_float_p, c_float_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_unary_from_label.argtypes = [c_float_p, c_short_p, C.c_int, C.c_int, C.c_float, c_float_p, c_float_p]
        L.orc_build_map.argtypes = [c_short_p, c_float_p, C.c_int, C.c_int]
        L.orc_meanfield.argtypes = [C.c_int, C.c_int, c_float_p, C.c_int, C.POINTER(C.POINTER(_Lattice)),
                                    C.POINTER(c_float_p), c_float_p, C.c_int, C.c_float, c_float_p, c_short_p]
        L.orc_features_image.argtypes = [c_float_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                         C.c_void_p, C.c_float]
        L.orc_map_point_unary.argtypes = [C.c_int, c_float_p, c_int_p, c_int_p, c_float_p, c_float_p,
                                          c_float_p, c_float_p, c_float_p, c_float_p, c_float_p]
        L.orc_rough_classify.argtypes = [C.c_int, c_float_p, c_float_p, c_float_p, C.c_void_p,
                                         C.POINTER(_SlamParams), c_short_p]
        L.orc_slam_crf.argtypes = [C.c_int, c_float_p, c_float_p, c_float_p, c_short_p, c_float_p,
                                   C.POINTER(_SlamParams), c_float_p, c_short_p, c_int_p]

        L.orc_epipolar_prior.argtypes = [C.c_int, c_float_p, c_float_p, C.c_void_p, C.c_float, C.c_float,
                                         C.c_void_p, C.c_void_p]
        L.orc_bf_match.restype = C.c_int
        L.orc_bf_match.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, c_int_p, c_int_p]

    # -- label application (Tracking.cc:1945-1955) --
    def label_partition(self, res_label, fid=None):
        lab = np.ascontiguousarray(res_label, dtype=np.int16)
        f = None if fid is None else np.ascontiguousarray(fid, dtype=np.int32)
        dyn, stat = np.empty(lab.size, dtype=np.int32), np.empty(lab.size, dtype=np.int32)
        self.lib.orc_label_partition.restype = C.c_int
        self.lib.orc_label_partition.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        nd = self.lib.orc_label_partition(lab.size, lab.ctypes.data, None if f is None else f.ctypes.data, dyn.ctypes.data, stat.ctypes.data)
        return dyn[:nd].copy(), stat[:lab.size - nd].copy()

    # -- frontend feeders --
    def epipolar_prior(self, pt1, pt2, F, u_gamma, stdev_gamma):
        pt1, pt2 = _f32(pt1), _f32(pt2)
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
        M = pt1.shape[0]
        dis, prob = np.empty(M, dtype=np.float64), np.empty(M, dtype=np.float64)
        self.lib.orc_epipolar_prior(M, _fp(pt1), _fp(pt2), F.ctypes.data, u_gamma, stdev_gamma, dis.ctypes.data, prob.ctypes.data)
        return dis, prob

    def bf_match(self, desc_q, desc_t, ratio=0.6):
        dq = np.ascontiguousarray(desc_q, dtype=np.uint8).reshape(-1, 32)
        dt = np.ascontiguousarray(desc_t, dtype=np.uint8).reshape(-1, 32)
        match = np.empty(dq.shape[0], dtype=np.int32)
        knn = np.empty((dq.shape[0], 4), dtype=np.int32)
        n = self.lib.orc_bf_match(dq.shape[0], dq.ctypes.data, dt.shape[0], dt.ctypes.data, ratio, _ip(match), _ip(knn))
        return match, knn, n

    # -- lattice --
    def lattice(self, feat: np.ndarray):
        feat = _f32(feat)
        N, d = feat.shape
        h = self.lib.orc_lattice_init(_fp(feat), d, N)
        l = h.contents
        D = d + 1
        out = dict(
            N=N, d=d, V=l.V,
            offset=np.ctypeslib.as_array(l.offset, shape=(max(N * D, 1),))[:N * D].copy().reshape(N, D),
            bary=np.ctypeslib.as_array(l.bary, shape=(max(N * D, 1),))[:N * D].copy().reshape(N, D),
            nbr=np.ctypeslib.as_array(l.nbr, shape=(max(2 * D * l.V, 1),))[:2 * D * l.V].copy().reshape(D, l.V, 2),
            keys=np.ctypeslib.as_array(l.keys, shape=(max(l.V * d, 1),))[:l.V * d].copy().reshape(l.V, d),
            handle=h,
        )
        return out

    def lattice_free(self, lat):
        self.lib.orc_lattice_free(lat["handle"])

    def filter(self, lat, x: np.ndarray) -> np.ndarray:
        x = _f32(x)
        N = lat["N"]
        L = x.size // max(N, 1) if N else 1
        out = np.empty_like(x)
        self.lib.orc_lattice_filter(lat["handle"], _fp(out), _fp(x), L)
        return out

    def potts_norm(self, lat) -> np.ndarray:
        out = np.empty(lat["N"], dtype=np.float32)
        self.lib.orc_potts_norm(lat["handle"], _fp(out))
        return out

    # -- driver --
    def fast_exp(self, x: float) -> float:
        return float(self.lib.orc_fast_exp(C.c_float(x)))

    def exp_and_normalize(self, x: np.ndarray, scale: float, relax: float = 1.0, prev=None) -> np.ndarray:
        x = _f32(x)
        N, L = x.shape
        out = _f32(prev).copy() if prev is not None else np.zeros_like(x)
        self.lib.orc_exp_and_normalize(_fp(out), _fp(x), N, L, scale, relax)
        return out

    def unary_from_label(self, label: np.ndarray, L: int, u_energy: float, n_en, p_en) -> np.ndarray:
        label = np.ascontiguousarray(label, dtype=np.int16)
        n_en, p_en = _f32(n_en), _f32(p_en)
        out = np.empty((label.size, L), dtype=np.float32)
        self.lib.orc_unary_from_label(_fp(out), _sp(label), label.size, L, u_energy, _fp(n_en), _fp(p_en))
        return out

    def meanfield(self, unary: np.ndarray, feats, weights, iters: int, relax: float = 1.0, with_map=True):
        """DenseCRF::inference over Potts potentials built from `feats` (list of [N,d] arrays)."""
        unary = _f32(unary)
        N, L = unary.shape
        lats = [self.lattice(f) for f in feats]
        norms = [self.potts_norm(l) for l in lats]
        K = len(lats)
        lat_arr = (C.POINTER(_Lattice) * K)(*[l["handle"] for l in lats])
        norm_arr = (c_float_p * K)(*[_fp(n) for n in norms])
        w = _f32(weights)
        Q = np.empty((N, L), dtype=np.float32)
        mp = np.empty(N, dtype=np.int16)
        self.lib.orc_meanfield(N, L, _fp(unary), K, lat_arr, norm_arr, _fp(w), iters, relax, _fp(Q),
                               _sp(mp) if with_map else None)
        V = [l["V"] for l in lats]
        for l in lats:
            self.lattice_free(l)
        return Q, (mp if with_map else None), V

    def features_image(self, W, H, F, posdev, img=None, featuredev=0.0) -> np.ndarray:
        feat = np.empty((W * H, F), dtype=np.float32)
        u8 = f32 = None
        if img is not None:
            if img.dtype == np.uint8:
                img = np.ascontiguousarray(img)
                u8 = img.ctypes.data
            else:
                img = _f32(img)
                f32 = img.ctypes.data
        self.lib.orc_features_image(_fp(feat), W, H, F, posdev, u8, f32, featuredev)
        return feat

    # -- SLAM path --
    def map_point_unary(self, snap):
        N = snap.n
        ob, er, de = (np.empty(N, dtype=np.float32) for _ in range(3))
        self.lib.orc_map_point_unary(N, _fp(snap.xyz), _ip(snap.obs_ptr), _ip(snap.obs_kf), _fp(snap.obs_uv),
                                     _fp(snap.kf_pose), _fp(snap.kf_intr), _fp(snap.kf_bounds),
                                     _fp(ob), _fp(er), _fp(de))
        return ob, er, de

    def rough_classify(self, observs, error, depth, prm: _SlamParams, p4=None) -> np.ndarray:
        observs, error, depth = _f32(observs), _f32(error), _f32(depth)
        lab = np.empty(observs.size, dtype=np.int16)
        p4p = None
        if p4 is not None:
            p4 = np.ascontiguousarray(p4, dtype=np.float64)
            p4p = p4.ctypes.data
        self.lib.orc_rough_classify(observs.size, _fp(observs), _fp(error), _fp(depth), p4p, C.byref(prm), _sp(lab))
        return lab

    def slam_crf(self, observs, error, kp2d, init_label, energies, prm: _SlamParams):
        observs, error, kp2d = _f32(observs), _f32(error), _f32(kp2d)
        init_label = np.ascontiguousarray(init_label, dtype=np.int16)
        energies = _f32(energies)
        N = observs.size
        Q = np.empty((N, 2), dtype=np.float32)
        mp = np.empty(N, dtype=np.int16)
        V = np.zeros(2, dtype=np.int32)
        self.lib.orc_slam_crf(N, _fp(observs), _fp(error), _fp(kp2d), _sp(init_label), _fp(energies),
                              C.byref(prm), _fp(Q), _sp(mp), _ip(V))
        return Q, mp, V


class Ref:
    """The reference's own headers compiled in place (oracle/_ref/libref.so)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libref.so"))

    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libref.so"))
        L.ref_lattice_init.restype = C.c_void_p
        L.ref_lattice_init.argtypes = [c_float_p, C.c_int, C.c_int]
        L.ref_lattice_free.argtypes = [C.c_void_p]
        L.ref_lattice_V.argtypes = [C.c_void_p]
        L.ref_lattice_export.argtypes = [C.c_void_p, c_int_p, c_float_p, c_int_p]
        L.ref_lattice_filter.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int]
        if hasattr(L, "ref_lattice_filter_window"):
            L.ref_lattice_filter_window.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int] + [C.c_int] * 4
        L.ref_crf3d.argtypes = [C.c_int, C.c_int, c_float_p, c_short_p, C.c_float, C.c_int, C.POINTER(c_float_p),
                                c_int_p, c_float_p, C.c_int, C.c_float, c_float_p, c_short_p]
        L.ref_slam_crf.argtypes = [C.c_int, c_float_p, c_float_p, c_float_p, c_short_p] + [C.c_float] * 7 + [
            C.c_int, c_float_p, c_short_p]
        L.ref_image_crf.argtypes = [C.c_int, C.c_int, C.c_int, c_short_p, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_int, c_float_p, c_short_p]
        L.ref_label_energies.argtypes = [C.c_int, C.c_float, c_float_p]

    def lattice(self, feat: np.ndarray):
        feat = _f32(feat)
        N, d = feat.shape
        D = d + 1
        h = self.lib.ref_lattice_init(_fp(feat), d, N)
        V = self.lib.ref_lattice_V(h)
        off = np.empty((N, D), dtype=np.int32)
        bary = np.empty((N, D), dtype=np.float32)
        nbr = np.empty((D, V, 2), dtype=np.int32)
        self.lib.ref_lattice_export(h, _ip(off), _fp(bary), _ip(nbr))
        return dict(N=N, d=d, V=V, offset=off, bary=bary, nbr=nbr, handle=h)

    def lattice_free(self, lat):
        self.lib.ref_lattice_free(lat["handle"])

    def filter(self, lat, x: np.ndarray) -> np.ndarray:
        x = _f32(x)
        L = x.size // max(lat["N"], 1) if lat["N"] else 1
        out = np.empty_like(x)
        self.lib.ref_lattice_filter(lat["handle"], _fp(out), _fp(x), L)
        return out

    def filter_window(self, lat, x: np.ndarray, L: int, in_offset=0, out_offset=0, in_size=-1, out_size=-1) -> np.ndarray:
        """PermutohedralLatticeCPU::compute with its windowing arguments (permutohedral_cpu.h:634-637): x holds the
        in_size input points, the result the out_size output points"""
        x = _f32(x)
        n_out = lat["N"] - out_offset if out_size == -1 else out_size
        out = np.empty((n_out, L), dtype=np.float32)
        self.lib.ref_lattice_filter_window(lat["handle"], _fp(out), _fp(x), L, in_offset, out_offset, in_size, out_size)
        return out

    def crf3d(self, L, feats, weights, iters, unary=None, label=None, conf=0.5, relax=1.0, with_map=True):
        feats = [_f32(f) for f in feats]
        N = feats[0].shape[0]
        K = len(feats)
        farr = (c_float_p * K)(*[_fp(f) for f in feats])
        dims = np.array([f.shape[1] for f in feats], dtype=np.int32)
        w = _f32(weights)
        Q = np.empty((N, L), dtype=np.float32)
        mp = np.empty(N, dtype=np.int16)
        up = lp = None
        if unary is not None:
            unary = _f32(unary)
            up = _fp(unary)
        else:
            label = np.ascontiguousarray(label, dtype=np.int16)
            lp = _sp(label)
        rc = self.lib.ref_crf3d(N, L, up, lp, conf, K, farr, _ip(dims), _fp(w), iters, relax, _fp(Q),
                                _sp(mp) if with_map else None)
        assert rc == 0, "unsupported (M, d) instantiation in ref_shim.cpp"
        return Q, (mp if with_map else None)

    def slam_crf(self, observs, error, kp2d, init_label, prm: _SlamParams):
        observs, error, kp2d = _f32(observs), _f32(error), _f32(kp2d)
        init_label = np.ascontiguousarray(init_label, dtype=np.int16)
        N = observs.size
        Q = np.empty((N, 2), dtype=np.float32)
        mp = np.empty(N, dtype=np.int16)
        self.lib.ref_slam_crf(N, _fp(observs), _fp(error), _fp(kp2d), _sp(init_label), prm.confidence, prm.w1,
                              prm.w2, prm.stdev_beta, prm.stdev_alpha, prm.point3d_stdev, prm.point2d_stdev,
                              prm.iters, _fp(Q), _sp(mp))
        return Q, mp

    def image_crf(self, W, H, L, label, conf, w_g, sd_g, w_b, sd_b, img_u8, sd_rgb, iters, want_q=True):
        label = np.ascontiguousarray(label, dtype=np.int16)
        img_u8 = np.ascontiguousarray(img_u8, dtype=np.uint8)
        Q = np.empty((W * H, L), dtype=np.float32) if want_q else None
        mp = np.empty(W * H, dtype=np.int16)
        rc = self.lib.ref_image_crf(W, H, L, _sp(label), conf, w_g, sd_g, w_b, sd_b, img_u8.ctypes.data, sd_rgb,
                                    iters, _fp(Q) if want_q else None, _sp(mp))
        assert rc == 0
        return Q, mp

    def label_energies(self, L, conf) -> np.ndarray:
        out = np.empty(3, dtype=np.float32)
        self.lib.ref_label_energies(L, conf, _fp(out))
        return out

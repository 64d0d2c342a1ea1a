"""Regenerates the golden fixtures in tests/golden/ (run in the build container, where
/root/reference and oracle/_ref/libref.so exist).  Nothing here runs on the GPU box.

  golden_im1.npz   the reference's ONE known-answer test: Thirdparty/DenseCRF/examples/{im1,anno1}.ppm ->
                   res1_cpu.ppm (DenseCRFCPU<21>, Gaussian F=2 + bilateral F=5, 10 iterations,
                   example_cpu.cpp:32,60,80-98).  Stored as arrays: image, label (classify(), :34-51),
                   golden MAP labels (res1_cpu.ppm colours mapped back through the same colour table).
  golden_ref.npz   outputs of the UNMODIFIED reference headers (oracle/_ref/libref.so) on small seeded
                   inputs: lattices (offset_/barycentric_/blur_neighbors_) for d in {2,3,5} incl. N%4 != 0,
                   filter outputs, and a full DenseCRF3D<2> + PottsPotential3D<2,2> run of the
                   Tracking.cc:1919-1930 call sequence (marginals + MAP).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
EX = "/root/reference/Thirdparty/DenseCRF/examples"


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        line = f.readline()
        while line.startswith(b"#"):
            line = f.readline()
        w, h = map(int, line.split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h * 3), dtype=np.uint8).reshape(h * w, 3).copy(), w, h


def classify(anno, M):
    """example_cpu.cpp:34-51: colour -> label in order of first appearance, black -> -1."""
    colors = []
    res = np.empty(anno.shape[0], dtype=np.int16)
    a64 = anno.astype(np.int64)
    code = a64[:, 0] + 256 * a64[:, 1] + 65536 * a64[:, 2]
    for k, c in enumerate(code):
        c = int(c)
        if c in colors:
            i = colors.index(c)
        else:
            i = len(colors)
            if c:
                if i < M:
                    colors.append(c)
                else:
                    c = 0
        res[k] = i if c else -1
    return res, colors


def main():
    from oracle.pyoracle import Ref, slam_params
    synth = importlib.import_module("lc-crf-slam_b200.synth")
    im, W, H = read_ppm(os.path.join(EX, "im1.ppm"))
    anno, _, _ = read_ppm(os.path.join(EX, "anno1.ppm"))
    res, _, _ = read_ppm(os.path.join(EX, "res1_cpu.ppm"))
    label, colors = classify(anno, 21)
    r64 = res.astype(np.int64)
    code = r64[:, 0] + 256 * r64[:, 1] + 65536 * r64[:, 2]
    lut = {c: i for i, c in enumerate(colors)}
    # colorize() (example_cpu.cpp:22-29) writes colors[map[k]]; labels >= nColors read colors[] zeros
    gmap = np.array([lut.get(int(c), -2) for c in code], dtype=np.int16)
    assert (gmap >= 0).all(), "golden image contains a colour outside the annotation colour table"
    np.savez_compressed(os.path.join(HERE, "golden_im1.npz"), im=im, label=label, map=gmap, W=W, H=H,
                        n_colors=len(colors))
    print("golden_im1.npz: %dx%d, %d colours, %d labelled px" % (W, H, len(colors), int((label >= 0).sum())))

    r = Ref()
    out = {}
    rng = np.random.default_rng(1234)
    for d, N in ((2, 6), (2, 257), (3, 130), (5, 203)):
        f = rng.normal(0, 2.5, (N, d)).astype(np.float32)
        f[::5] = np.round(f[::5])
        lat = r.lattice(f)
        x = rng.random((N, 3)).astype(np.float32)
        y = r.filter(lat, x)
        k = "lat_d%d_n%d_" % (d, N)
        out[k + "feat"], out[k + "offset"], out[k + "bary"], out[k + "nbr"] = f, lat["offset"], lat["bary"], lat["nbr"]
        out[k + "x"], out[k + "y"] = x, y
        r.lattice_free(lat)
    prm = slam_params(**synth.SLAM_PARAMS)
    fr = synth.slam_frame(1003, seed=77)
    # init labels by a fixed rule so the fixture does not depend on any exp() implementation
    lab = ((fr.observs > 3) & (fr.error < 2.6)).astype(np.int16)
    Q, mp = r.slam_crf(fr.observs, fr.error, fr.kp2d, lab, prm)
    out.update(slam_observs=fr.observs, slam_error=fr.error, slam_kp2d=fr.kp2d, slam_label=lab, slam_Q=Q, slam_map=mp,
               slam_energies=r.label_energies(2, prm.confidence))
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **out)
    print("golden_ref.npz written,", int(mp.sum()), "static of", mp.size)


if __name__ == "__main__":
    main()

"""Regenerates tests/golden/golden_frontend.npz (run in the build container; needs the cv2 wheel, which is absent
on the GPU box).  Pins the restatement of Tracking::BfMatch (src/Tracking.cc:1747-1766) against the real
cv::BFMatcher(cv::NORM_HAMMING).knnMatch(query, train, 2) of OpenCV (cv2 4.13 here; the reference links OpenCV 3.x,
whose brute-force matcher has the same K-best insertion).  Stored per case: descriptors, the two nearest train
rows {d0, i0, d1, i1} as OpenCV returns them, and the accepted correspondences of the 0.6 ratio test evaluated
exactly as :1755 writes it (float distances, double product)."""
import importlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
synth = importlib.import_module("lc-crf-slam_b200.synth")

CASES = {  # name: (nq, nt, seed, match_frac, flip_bits, entropy_bytes)
    "orb": (700, 900, 1, 0.6, 24, 32),
    "ties": (300, 500, 2, 0.5, 3, 3),       # 24-bit descriptors: many equal distances
    "one_train_row": (40, 1, 3, 0.5, 4, 32),  # knnMatch returns a single neighbour -> no correspondence
    "two_train_rows": (40, 2, 4, 0.5, 4, 32),
    "ragged": (257, 513, 5, 0.8, 40, 32),
}


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    for name, (nq, nt, seed, mf, fb, eb) in CASES.items():
        dq, dt = synth.orb_frame_pair(nq, nt, seed, mf, fb, eb)
        knn = np.full((nq, 4), -1, dtype=np.int32)
        match = np.full(nq, -1, dtype=np.int32)
        for q, mm in enumerate(bf.knnMatch(dq, dt, 2)):
            for j, m in enumerate(mm):
                assert m.queryIdx == q and float(m.distance) == int(m.distance)
                knn[q, 2 * j], knn[q, 2 * j + 1] = int(m.distance), m.trainIdx
            if len(mm) == 2 and float(np.float32(mm[0].distance)) < float(np.float32(mm[1].distance)) * 0.6:
                match[q] = mm[0].trainIdx
        out[name + "_dq"], out[name + "_dt"], out[name + "_knn"], out[name + "_match"] = dq, dt, knn, match
        print(name, "accepted", int((match >= 0).sum()), "of", nq)
    np.savez_compressed(os.path.join(HERE, "golden_frontend.npz"), **out)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Diagnose the global-keyframe-table path (kf_ptr=None, nKF > 640) of the unary kernel against the oracle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("lc-crf-slam_b200")
synth = pkg.synth
from oracle.pyoracle import Oracle
o = Oracle()
prm = pkg.SlamParams.make()
bits = lambda a: np.ascontiguousarray(a).view(np.int32)
snaps = [synth.map_snapshot(n, ob, seed=90 + i, n_kf=300, ragged=r) for i, (n, ob, r) in
         enumerate(((2500, 64, False), (130, 7, True), (4000, 20, True), (31, 64, False)))]
cat = bench.concat_snapshots(snaps)
er_or = np.concatenate([o.map_point_unary(s)[1] for s in snaps])
ctx = pkg.Context(0)
ctx.set_option("graphs", 0)
F = pkg.Frames(ctx, [s.n for s in snaps], prm)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for rep in range(reps):
    F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"],
                     cat["kf_bounds"], cat["kp2d"], None)
    F.run()
    d = F.get_debug()
    bad = np.nonzero(bits(d["error"]) != bits(er_or))[0]
    print("rep", rep, "bad points", bad.tolist())
    common = None
    for p in bad:
        ks = set(cat["obs_kf"][cat["obs_ptr"][p]:cat["obs_ptr"][p + 1]].tolist())
        common = ks if common is None else (common & ks)
    print("   keyframes common to all bad points:", sorted(common) if common else None, flush=True)
F.close(); ctx.close()

# ---- stand-alone unary API on the same concatenated snapshot and on single snapshots with many keyframes
from types import SimpleNamespace
ctx = pkg.Context(0)
big = SimpleNamespace(**cat, n=cat["xyz"].shape[0])
for name, s in (("concat(1200 kf)", big), ("single 700 kf N=37", synth.map_snapshot(37, 64, seed=5, n_kf=700)),
                ("single 700 kf N=256+5 ragged", synth.map_snapshot(261, 40, seed=6, n_kf=700, ragged=True))):
    s2 = SimpleNamespace(xyz=s.xyz, obs_ptr=s.obs_ptr, obs_kf=s.obs_kf, obs_uv=s.obs_uv, kf_pose=s.kf_pose, kf_intr=s.kf_intr,
                         kf_bounds=s.kf_bounds, n=s.xyz.shape[0])
    ob, er, de = o.map_point_unary(s2)
    for rep in range(3):
        gob, ger, gde = ctx.map_point_unary(s2)
        bad = np.nonzero((bits(ger) != bits(er)) | (bits(gde) != bits(de)))[0]
        print(name, "rep", rep, "bad", bad.tolist()[:12])
        for p in bad[:3]:
            a, z = int(s2.obs_ptr[p]), int(s2.obs_ptr[p + 1])
            print("   p", p, "obs", z - a, "err exp/got", er[p], ger[p], "sum diff", (ger[p] - er[p]) * (z - a),
                  "depth exp/got", de[p], gde[p], "sum diff", (gde[p] - de[p]) * (z - a))
ctx.close()

#!/usr/bin/env python
"""Diagnose test_frames_map_batch_keyframe_slices: which outputs differ between the kf_ptr / no-kf_ptr runs and the oracle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("lc-crf-slam_b200")
synth = pkg.synth
from oracle.pyoracle import Oracle, slam_params
o = Oracle()
prm_o = slam_params(**synth.SLAM_PARAMS)
prm = pkg.SlamParams.make()
en = pkg.label_energies(2, prm.confidence)
bits = lambda a: np.ascontiguousarray(a).view(np.int32)
snaps = [synth.map_snapshot(n, ob, seed=90 + i, n_kf=300, ragged=r) for i, (n, ob, r) in
         enumerate(((2500, 64, False), (130, 7, True), (4000, 20, True), (31, 64, False)))]
cat = bench.concat_snapshots(snaps)
ref = {}
oo = 0
Q_or, m_or, er_or, de_or, ob_or = [], [], [], [], []
for s in snaps:
    ob, er, de = o.map_point_unary(s)
    lab = o.rough_classify(ob, er, de, prm_o)
    Qo, mo, _ = o.slam_crf(ob, er, s.kp2d, lab, en, prm_o)
    Q_or.append(Qo); m_or.append(mo); er_or.append(er); de_or.append(de); ob_or.append(ob)
Q_or = np.concatenate(Q_or); er_or = np.concatenate(er_or); de_or = np.concatenate(de_or)
for opts in ((), (("graphs", 0),), (("concurrent", 0),), (("graphs", 0), ("concurrent", 0))):
    ctx = pkg.Context(0)
    for k, v in opts:
        try:
            ctx.set_option(k, v)
        except Exception as e:
            print("option", k, "rejected:", e)
    F = pkg.Frames(ctx, [s.n for s in snaps], prm, en)
    for rep in range(3):
        for name, kf_ptr in (("kf_ptr", cat["kf_ptr"]), ("none", None)):
            F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"],
                             cat["kf_bounds"], cat["kp2d"], kf_ptr)
            for run in range(2):
                F.run()
                m, p = F.get_outputs()
                d = F.get_debug()
                ne = int((bits(d["error"]) != bits(er_or)).sum()); nd = int((bits(d["depth"]) != bits(de_or)).sum())
                nq = int((bits(p) != bits(Q_or)).any(axis=1).sum())
                where = np.nonzero((bits(p) != bits(Q_or)).any(axis=1))[0]
                print(opts, "rep", rep, name, "run", run, "err_diff", ne, "depth_diff", nd, "Q_diff_points", nq,
                      "first", where[:5].tolist(), "V", d["V"].tolist(), flush=True)
    F.close(); ctx.close()

#!/usr/bin/env python
"""Experiment: do two independent frame batches on two contexts/streams overlap on one GPU?
(device-resident C3 batches; wall clock over many steps with a full sync on both sides)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("lc-crf-slam_b200")

def make(batch, seed0, nctx):
    out = []
    for c in range(nctx):
        st = torch.cuda.Stream()
        ctx = pkg.Context(0, stream=st.cuda_stream)
        probs = bench.make_problems("c3", batch, seed0 + 1000 * c)
        F = pkg.Frames(ctx, [p.n for p in probs], pkg.SlamParams.make())
        cat = bench.concat_snapshots(probs)
        F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"],
                         cat["kf_bounds"], cat["kp2d"], cat["kf_ptr"])
        out.append((st, ctx, F))
    return out

def run(sets, steps):
    for _ in range(3):
        for st, ctx, F in sets:
            F.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for st, ctx, F in sets:
            F.run()
    torch.cuda.synchronize()
    return time.perf_counter() - t0

if __name__ == "__main__":
    steps = 40
    for batch, nctx in ((16, 1), (8, 2), (16, 2), (8, 3), (32, 1)):
        sets = make(batch, 1000, nctx)
        dt = run(sets, steps)
        print("batch %d x %d contexts: %.3f ms per round, %.0f problems/s" % (batch, nctx, 1e3 * dt / steps, steps * batch * nctx / dt), flush=True)
        for st, ctx, F in sets:
            F.close(); ctx.close()

"""GPU experiment: throughput of C3 batches when two contexts (two streams, two workspaces) run their steps concurrently on
one GPU, against one context alone.  Flat snapshot inputs, no result checks.
Usage: python scripts/dual_ctx_probe.py [problems] [steps] [contexts]"""
import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("lc-crf-slam_b200")
synth = pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
NC = int(sys.argv[3]) if len(sys.argv) > 3 else 2
snaps = [synth.map_snapshot(100000, 64, seed=1000 + i, unique_kf=True) for i in range(B)]
cat = pkg.concat_frames(snaps)
ctxs, Fs = [], []
for c in range(NC):
    ctx = pkg.Context(0)
    F = pkg.Frames(ctx, [s.n for s in snaps])
    F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"], cat["kf_bounds"], cat["kp2d"], cat["kf_ptr"])
    for _ in range(4):
        F.run()
    ctx.sync()
    ctxs.append(ctx); Fs.append(F)
for n_active in range(1, NC + 1):
    for c in ctxs:
        c.sync()
    t0 = time.perf_counter()
    for _ in range(K):
        for F in Fs[:n_active]:
            F.run()
    for c in ctxs:
        c.sync()
    dt = time.perf_counter() - t0
    print("%d context(s): %.3f ms per step, %.0f problems/s" % (n_active, dt * 1e3 / (K * n_active), B * K * n_active / dt), flush=True)

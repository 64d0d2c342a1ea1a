"""GPU experiment: device-resident time of one C3 batch (flat snapshot inputs, no result checks) -- graph replay with CUDA
events plus the per-kernel profile.  For timing experimental builds of liblccrf.so; not a bench.
Usage: python scripts/step_probe.py [problems] [steps]"""
import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("lc-crf-slam_b200")
synth = pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = pkg.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
snaps = [synth.map_snapshot(100000, 64, seed=1000 + i, unique_kf=True) for i in range(B)]
cat = pkg.concat_frames(snaps)
F = pkg.Frames(ctx, [s.n for s in snaps])
F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"], cat["kf_bounds"], cat["kp2d"], cat["kf_ptr"])
for _ in range(4):
    F.run()
ctx.sync()
t0 = time.perf_counter()
for _ in range(K):
    F.run()
ctx.sync()
ms = (time.perf_counter() - t0) * 1e3 / K
ctx.set_option("profile", 1); ctx.profile_report()
for _ in range(3):
    F.run()
rep = ctx.profile_report(); ctx.set_option("profile", 0)
tot = sum(v[1] for v in rep.values()) / 3
keys = ("k_scan_compose", "k_splat_rows", "k_map_point_unary", "k_csr_fill", "k_csr_count", "k_csr_prefix", "k_csr_zero", "k_embed")
print("step %.3f ms (%d problems) | kernel sum %.3f ms |" % (ms, B, tot), " ".join("%s %.4f" % (k, rep[k][1] / rep[k][0]) for k in keys if k in rep), flush=True)

"""GPU probe: achieved GB/s of the element-parallel blur (bulk-copy and plain-load kernels) on the 2048 x 2048 stress lattice."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("lc-crf-slam_b200")
peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]) if os.path.exists("MEASURED_PEAKS.json") else 6650.0
ctx = pkg.Context(0)
W = H = 2048
yy, xx = np.mgrid[0:H, 0:W]
feat = np.stack([xx.ravel() / np.float32(0.5), yy.ravel() / np.float32(0.5)], axis=1).astype(np.float32)
lat = pkg.Lattice(ctx, feat)
rng = np.random.default_rng(7)
for L in (2, 4):
    x = rng.random((W * H, L), dtype=np.float32)
    for bulk in (1, 0):
        ctx.set_option("bulk_blur", bulk)
        lat.filter(x)
        ctx.set_option("profile", 1); ctx.profile_report()
        for _ in range(5):
            lat.filter(x)
        rep = ctx.profile_report(); ctx.set_option("profile", 0)
        cnt, tms = rep["k_blur"]
        gbs = lat.V * (8 * L + 8) / (tms / cnt * 1e-3) / 1e9
        print("L=%d %s: %.4f ms  %.0f GB/s  %.3f of peak" % (L, "bulk" if bulk else "vec ", tms / cnt, gbs, gbs / peak), flush=True)

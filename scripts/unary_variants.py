"""GPU experiment: time of k_map_point_unary for the input variants of one C3-shaped batch (profile mode, CUDA events):
flat snapshot (random keyframes / unique keyframes per point) vs the resident map (bulk-loaded lists with 25% room)."""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("lc-crf-slam_b200")
synth = pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = pkg.Context(0)
ctx.set_option("ordered_splat", 0)


def t_unary(F, n=3):
    F.run(); F.run()
    ctx.set_option("profile", 1); ctx.profile_report()
    for _ in range(n):
        F.run()
    rep = ctx.profile_report(); ctx.set_option("profile", 0)
    c, ms = rep["k_map_point_unary"]
    return ms / c


for uniq in (False, True):
    snaps = [synth.map_snapshot(100000, 64, seed=1000 + i, unique_kf=uniq) for i in range(B)]
    cat = pkg.concat_frames(snaps)
    F = pkg.Frames(ctx, [s.n for s in snaps])
    F.set_map_inputs(cat["xyz"], cat["obs_ptr"], cat["obs_kf"], cat["obs_uv"], cat["kf_pose"], cat["kf_intr"], cat["kf_bounds"], cat["kp2d"], cat["kf_ptr"])
    print("flat snapshot, unique_kf=%s: %.4f ms per %d observations" % (uniq, t_unary(F), cat["obs_kf"].size), flush=True)
    F.close()
STRIDE = 32768
fids, tabs, ko = [], [], 0
for i, s in enumerate(snaps):
    fid, tab, uvc = synth.index_observations(s.obs_kf, s.obs_uv, 256, seed=3 + i, stride=STRIDE)
    fids.append(np.stack([s.obs_kf + ko, fid], axis=1)); tabs.append(tab); ko += 256
tab_all, ref_all = np.concatenate(tabs), np.concatenate(fids)
for slack in (25, 0):
    ctx.set_option("map_slack", slack)
    mp = pkg.Map(ctx, STRIDE)
    mp.apply(kf_pose=cat["kf_pose"], kf_intr=cat["kf_intr"], kf_bounds=cat["kf_bounds"], kf_keypoints=tab_all, xyz=cat["xyz"])
    mp.set_observations(cat["obs_ptr"], ref_all)
    F = pkg.Frames(ctx, [s.n for s in snaps])
    F.set_visible(mp, np.arange(cat["xyz"].shape[0], dtype=np.int32), cat["kp2d"], kf_ptr=cat["kf_ptr"])
    print("resident map, list room +%d%% (visible ids 0..N-1): %.4f ms" % (slack, t_unary(F)), flush=True)
    F.close(); mp.close()

"""GPU probe: where the time of one C5 replay step goes (host clock inside wait / submit, with and without the map delta)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("lc-crf-slam_b200")
synth = pkg.synth
ctx = pkg.Context(0)
keep = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); keep.append(t); return t.numpy()
gen = synth.SequenceReplay(seed=5000, kind="tum", frames_per_batch=64)
kfs, xyz, ptr, ref = gen.initial_map()
mp = pkg.Map(ctx, gen.stride)
mp.apply(kf_pose=kfs["pose"], kf_intr=kfs["intr"], kf_bounds=kfs["bounds"], kf_keypoints=kfs["kp"], xyz=xyz)
mp.set_observations(ptr, ref)
F = pkg.Frames(ctx, gen.sizes)
NT = int(sum(gen.sizes))
steps = []
for _ in range(14):
    b = gen.next_batch()
    steps.append((pin(b["ids"]), pin(b["kp2d"]), pkg.MapDelta.make(pin=pin, **b["delta"])))
outs = [(pin(np.zeros(NT, np.int16)), pin(np.zeros((NT, 2), np.float32))) for _ in (0, 1)]
raw = [gen.next_batch() for _ in range(10)]   # further steps for the timed part (the map keeps evolving)
def variants(b):
    d = b["delta"]
    keys = {"full": None,
            "xyz+pose": ("pose", "xyz"),
            "keyframes": ("kf_first", "kf_pose", "kf_intr", "kf_bounds", "kf_keypoints"),
            "adds": ("kf_first", "kf_pose", "kf_intr", "kf_bounds", "kf_keypoints", "add_pt", "add_kf", "add_fid", "add_seg_ptr"),
            "erases": ("erase_pt", "erase_kf", "erase_seg_ptr")}
    return keys
# warm-up: the first 14 steps (captures, allocations, pool growth)
for i in range(14):
    ids, kp, d = steps[i]
    F.wait(i & 1); F.submit_visible(i & 1, mp, ids, kp, outs[i & 1][0], outs[i & 1][1], delta=d)
F.wait(0); F.wait(1)
full = [(pin(b["ids"]), pin(b["kp2d"]), pkg.MapDelta.make(pin=pin, **b["delta"])) for b in raw]
def timed(name, seq):
    tw = ts = 0.0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i, (ids, kp, d) in enumerate(seq):
        a = time.perf_counter(); F.wait(i & 1); b_ = time.perf_counter()
        F.submit_visible(i & 1, mp, ids, kp, outs[i & 1][0], outs[i & 1][1], delta=d)
        c = time.perf_counter(); tw += b_ - a; ts += c - b_
    F.wait(0); F.wait(1); torch.cuda.synchronize()
    n = len(seq); tot = time.perf_counter() - t0
    print("%-10s %.3f ms per step; host inside wait %.3f, inside submit %.3f" % (name, 1e3 * tot / n, 1e3 * tw / n, 1e3 * ts / n), flush=True)
ctx.set_option("trace", 1)
timed("full", full)
ctx.set_option("trace", 0)
last = full[-1]
timed("no delta", [(last[0], last[1], None)] * 10)
sub = lambda b, ks: pkg.MapDelta.make(pin=pin, **{k: v for k, v in b["delta"].items() if k in ks})
b = raw[-1]
timed("xyz+pose", [(last[0], last[1], sub(b, ("pose", "xyz")))] * 10)
timed("erases(noop)", [(last[0], last[1], sub(b, ("erase_pt", "erase_kf", "erase_seg_ptr")))] * 10)
timed("adds(dup)", [(last[0], last[1], sub(b, ("add_pt", "add_kf", "add_fid", "add_seg_ptr")))] * 10)
timed("kf rows", [(last[0], last[1], sub(b, ("kf_first", "kf_pose", "kf_intr", "kf_bounds", "kf_keypoints")))] * 10)

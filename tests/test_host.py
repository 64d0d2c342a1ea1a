"""CPU tests of the host-side logic: synthetic generators, problem sharding, and the N>1 launch
path on gloo (world_size 2) -- the data path has no collective, only barrier + max-reduce of timings."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
synth = importlib.import_module("lc-crf-slam_b200.synth")
shard = importlib.import_module("lc-crf-slam_b200.shard")


def test_shard_contiguous_covers_everything_once():
    rng = np.random.default_rng(0)
    for n, world in ((1024, 8), (1024, 3), (5, 8), (8, 8), (1, 2), (0, 4), (17, 4)):
        w = rng.integers(4000, 6001, n).tolist()
        parts = shard.shard_contiguous(w, world)
        assert len(parts) == world
        assert parts[0][0] == 0 and parts[-1][1] == n
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        if n >= world:
            loads = [sum(w[a:b]) for a, b in parts]
            assert min(loads) > 0
            assert max(loads) <= 1.25 * (sum(w) / world) + max(w)


def test_shard_round_robin():
    for world in (1, 2, 4, 8):
        seen = sorted(sum((shard.shard_round_robin(8, world, r) for r in range(world)), []))
        assert seen == list(range(8))


def test_synth_shapes_and_determinism():
    a, b = synth.slam_frame(3001, 7), synth.slam_frame(3001, 7)
    assert a.n == 3001 and np.array_equal(a.error, b.error) and a.kp2d.shape == (3001, 2)
    assert a.observs.min() >= 1 and 0.05 < a.dynamic.mean() < 0.5
    s = synth.map_snapshot(1000, 64, 3)
    assert s.nnz == 64000 and s.obs_ptr[-1] == s.nnz and s.obs_kf.max() < s.kf_pose.shape[0]
    r = synth.map_snapshot(500, 32, 3, ragged=True)
    assert np.diff(r.obs_ptr).min() >= 1
    img, lab = synth.image_problem(64, 48, 1)
    assert img.shape == (64 * 48, 3) and img.dtype == np.uint8 and set(np.unique(lab)) <= {-1, 0, 1}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = r"""
import importlib, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, %r)
shard = importlib.import_module("lc-crf-slam_b200.shard")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sizes = [4000 + 37 * i %% 2000 for i in range(64)]
a, b = shard.shard_contiguous(sizes, world)[rank]
mine = torch.zeros(64, dtype=torch.int64); mine[a:b] = 1
dist.barrier()
t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)      # per-rank elapsed time
dist.all_reduce(t, op=dist.ReduceOp.MAX)                          # max over ranks, as bench.py does
cov = mine.clone(); dist.all_reduce(cov)                          # test-only: every problem owned exactly once
n = torch.tensor([b - a]); dist.all_reduce(n)
if rank == 0:
    assert bool((cov == 1).all()) and int(n) == 64 and abs(float(t) - 0.010 * world) < 1e-12
    print("OK", float(t), int(n))
dist.destroy_process_group()
"""


def test_gloo_world2_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_index_observations_restates_the_same_problem():
    """(keyframe, feature index) pairs + per-keyframe keypoint rows reproduce the flat observations; with a stride
    smaller than a keyframe's observation count keypoints are re-used and the consistent flat form says so."""
    synth = importlib.import_module("lc-crf-slam_b200.synth")
    s = synth.map_snapshot(2000, 12, seed=4, n_kf=32, ragged=True)
    fid, table, uvc = synth.index_observations(s.obs_kf, s.obs_uv, 32, seed=1)
    assert table.shape[0] == 32 and table.shape[1] & (table.shape[1] - 1) == 0 and table.shape[2] == 2
    assert fid.min() >= 0 and fid.max() < table.shape[1]
    assert np.array_equal(table[s.obs_kf, fid], s.obs_uv) and np.array_equal(uvc, s.obs_uv)
    pairs = np.stack([s.obs_kf, fid], axis=1)
    assert np.unique(pairs, axis=0).shape[0] == pairs.shape[0]  # one keypoint per observation
    fid2, table2, uvc2 = synth.index_observations(s.obs_kf, s.obs_uv, 32, seed=1, stride=16)
    assert table2.shape == (32, 16, 2) and fid2.max() < 16
    assert np.array_equal(table2[s.obs_kf, fid2], uvc2) and not np.array_equal(uvc2, s.obs_uv)
    first = np.unique(np.stack([s.obs_kf, fid2], axis=1), axis=0, return_index=True)[1]
    assert np.array_equal(uvc2[first], s.obs_uv[first])  # the first occurrence defines the keypoint


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the
    contract keys; here on the small C1 workload so that it finishes in seconds."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "crf_problems_per_s" and d["unit"] == "problems/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C1")


def test_bench_reference_arm_other_ranks_stay_silent():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_unary_fast_sqrt_window_argument():
    """The unary kernel's straight-line path (csrc/unary.cu observe_fast) takes (float)sqrt(s) from one double Newton
    step on an fp32 rsqrt seed and trusts it unless the low 29 mantissa bits of the result lie within 2^13 ulps of a
    float rounding boundary.  Model of that arithmetic with a seed that is WORSE than the hardware's (relative error
    up to 2^-21 instead of <= 2^-22; the window holds up to 2^-20.5): wherever the window test passes, the result
    equals the reference's (float)sqrt(double) bit for bit -- on random inputs and on inputs constructed next to
    rounding boundaries (squares of float midpoints, nudged by a few double ulps)."""
    rng = np.random.default_rng(42)
    n = 2_000_000
    mant = rng.random(n) + 1.0
    s = mant * np.exp2(rng.integers(-58, 58, n).astype(np.float64))
    # adversarial: squares of float midpoints, nudged by a few double ulps either way
    f = (rng.random(n // 4).astype(np.float32) + np.float32(1)) * np.exp2(rng.integers(-20, 20, n // 4)).astype(np.float32)
    mid = (f.astype(np.float64) + np.nextafter(f, np.float32(np.inf)).astype(np.float64)) * 0.5
    adv = mid * mid
    for k in (-3, -1, 0, 1, 3):
        t = adv.copy()
        for _ in range(abs(k)):
            t = np.nextafter(t, np.inf if k > 0 else 0.0)
        s = np.concatenate([s, t])
    want = np.sqrt(s).astype(np.float32)                        # reference: double sqrt, then rounded to float
    sf = s.astype(np.float32)
    worst_flagged = 0.0
    for delta in (-2.0 ** -21, -2.0 ** -22, 0.0, 2.0 ** -22, 2.0 ** -21, None):
        d = rng.uniform(-2.0 ** -21, 2.0 ** -21, s.size) if delta is None else delta
        rs = ((1.0 / np.sqrt(sf.astype(np.float64))) * (1.0 + d)).astype(np.float32)   # rsqrt.approx stand-in
        g = (sf * rs).astype(np.float64)                         # __fmul_rn(sf, rs): fp32 product
        h = (np.float32(0.5) * rs).astype(np.float64)
        e2 = s - g * g                                           # fma(-g, g, s): g*g is exact (24-bit g), one rounding
        y = (e2.astype(np.longdouble) * h.astype(np.longdouble) + g.astype(np.longdouble)).astype(np.float64)  # fma(e2, h, g)
        low = (y.view(np.uint64) & np.uint64(0x1fffffff)).astype(np.int64)
        amb = ((low - 0x0fffe000) & 0xffffffff) < 0x4000          # the kernel's unsigned 32-bit window test
        ok = y.astype(np.float32) == want
        assert ok[~amb].all(), "fast sqrt disagrees outside the ambiguity window (delta=%r)" % (delta,)
        worst_flagged = max(worst_flagged, float(amb[:n].mean()))
    assert worst_flagged < 2.0 ** -13  # the library fallback stays rare (expected 2^-15 on random inputs)


def test_bench_c2_gpu_arm_against_a_mock_device(monkeypatch, capsys):
    """bench.py --workload c2 (image CRF through the per-object API) cannot be timed without a GPU; its host logic --
    step loop, event/host timing, profile folding, algorithmic bytes, JSON contract -- runs here against a mock device."""
    import contextlib
    import json
    import time
    import torch
    bench = importlib.import_module("bench")
    pkg = importlib.import_module("lc-crf-slam_b200")

    class FakeStream:
        cuda_stream = 0

    class FakeEvent:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self, s=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t)

    class FakeCtx:
        kernel_launches = 0

        def __init__(self, dev, stream=None):
            pass

        def set_option(self, k, v):
            pass

        def profile_report(self):
            return {"k_blur": (99, 1.5), "k_embed": (2, 0.4), "k_mf_point_l2": (10, 0.5)}

    class FakeCRF:
        def __init__(self, ctx, N, L):
            self.N = N
            ctx.kernel_launches += 100

        def setUnaryEnergyFromLabel(self, lab, energies=None):
            assert lab.shape == (self.N,) and lab.dtype == np.int16

        def addPairwiseFromImage(self, W, H, w, sd, img=None, fd=0.0):
            assert img is None or (img.shape == (W * H, 3) and img.dtype == np.uint8)

        def inference(self, T, with_map):
            assert T == 10 and with_map

        def getMap(self):
            return np.zeros(self.N, np.int16)

        def potts_vertices(self, k):
            return (39890, 11600)[k]

        def close(self):
            pass

    class FakeSampler:
        def __init__(self, i):
            pass

        def stop(self):
            return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}

    real_tensor = torch.tensor
    monkeypatch.setattr(torch.cuda, "Stream", lambda: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch, "tensor", lambda data, dtype=None, device=None: real_tensor(data, dtype=dtype))
    monkeypatch.setattr(pkg, "Context", FakeCtx)
    monkeypatch.setattr(pkg, "DenseCRF", FakeCRF)
    monkeypatch.setattr(bench, "ClockSampler", FakeSampler)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "c2", "--steps", "2", "--warmup", "1", "--batch", "2", "--no-cpu-baseline"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "crf_problems_per_s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["config"]["workload"].startswith("C2") and d["config"]["problems_per_step_per_gpu"] == 2
    assert d["gpu_launches"] == 2 * 2 * 100
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 640 * 480 * 5 and d["e2e"]["value"] <= d["value"] * 1.0000001
    assert d["roofline"]["kernel"] == "k_blur" and abs(d["roofline"]["share_of_step"] - 0.625) < 1e-9
    # k_blur: D passes per filter call, T calls with 2 labels + 1 norm call with 1 label, both lattices (SURVEY 8d B_blur)
    want = (3 * 39890 + 6 * 11600) * ((8 * 2 + 8) * 10 + 16) / (9 * 11)
    assert abs(d["roofline"]["algorithmic_bytes_per_launch"] - want) < 1e-6


def _mock_device(monkeypatch):
    """Replace CUDA (torch.cuda streams/events, pinned memory) and the device-side classes of the package by host mocks
    that deliver fixed results; returns (bench module, call counters)."""
    import contextlib
    import json
    import time
    import torch
    bench = importlib.import_module("bench")
    pkg = importlib.import_module("lc-crf-slam_b200")
    calls = {"run": 0, "indexed": 0, "flat": 0, "table": 0, "visible": 0, "map_apply": 0, "map_bulk": 0}

    class FakeStream:
        cuda_stream = 0

    class FakeEvent:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self, s=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t) + 1e-3

    class FakeCtx:
        kernel_launches = 0

        def __init__(self, dev, stream=None):
            pass

        def sync(self):
            pass

        def set_option(self, k, v):
            pass

        def profile_report(self):
            return {"k_splat_rows": (36, 1.2), "k_scan_compose": (36, 1.0), "k_scan_walk": (40, 0.4),
                    "k_map_point_unary": (3, 2.0), "k_mf_point_l2": (15, 0.6), "k_blur": (9, 0.5)}

    class FakeFrames:
        def __init__(self, ctx, sizes, prm, energies=None):
            self.ctx, self.NT, self.B = ctx, int(sum(sizes)), len(sizes)
            self.map = (np.arange(self.NT) % 2).astype(np.int16)
            self.prob = np.stack([1.0 - self.map, self.map], axis=1).astype(np.float32)

        def set_map_inputs(self, xyz, obs_ptr, obs_kf, obs_uv, kf_pose, kf_intr, kf_bounds, kp2d, kf_ptr=None):
            assert obs_kf.dtype == np.int32 and obs_uv.shape == (obs_kf.size, 2) and obs_ptr[-1] == obs_kf.size

        def set_keyframe_keypoints(self, table, kf_first=0):
            assert table.ndim == 3 and table.shape[2] == 2 and table.dtype == np.float32
            calls["table"] += 1

        def run(self):
            calls["run"] += 1
            self.ctx.kernel_launches += 111

        def _deliver(self, m, p):
            m[:] = self.map
            p[:] = self.prob

        def get_outputs(self, m=None, p=None, want_prob=True):
            self._deliver(m, p)
            return m, p

        def submit_map(self, slot, xyz, obs_ptr, obs_kf, obs_uv, kf_pose, kf_intr, kf_bounds, kp2d, kf_ptr, m, p):
            assert obs_kf.dtype == np.uint16
            calls["flat"] += 1
            self._deliver(m, p)

        def submit_map_indexed(self, slot, xyz, obs_ptr, obs_ref, kf_pose, kf_intr, kf_bounds, kp2d, kf_ptr, m, p):
            assert obs_ref.dtype == np.uint16 and obs_ref.shape == (obs_ptr[-1], 2)
            calls["indexed"] += 1
            self._deliver(m, p)

        def set_inputs(self, observs, error, depth, kp2d):
            assert observs.shape == (self.NT,) and kp2d.shape == (self.NT, 2)

        def set_visible(self, mp, ids, kp2d, delta=None, kf_ptr=None):
            assert ids.dtype == np.int32 and ids.shape == (self.NT,) and kp2d.shape == (self.NT, 2)

        def submit_visible(self, slot, mp, ids, kp2d, m, p, delta=None, kf_ptr=None):
            assert ids.dtype == np.int32 and delta is not None and delta.n_add == delta.n_erase > 0 and delta.n_xyz == self.NT
            calls["visible"] += 1
            self._deliver(m, p)

        def submit(self, slot, observs, error, depth, kp2d, m, p):
            calls["direct"] = calls.get("direct", 0) + 1
            self._deliver(m, p)

        def wait(self, slot):
            pass

        def get_debug(self):
            return {"V": np.array([[29, 1200]] * self.B, dtype=np.int32)}

        def debug_counters(self, k):
            return {"long_rows": 1}

        def algorithmic_bytes(self):
            return {"total": 3.0e8, "per_iteration": 2.0e7, "unary": 1.0e8}

    class FakeLattice:
        V = 1000

        def __init__(self, ctx, feat):
            pass

        def filter(self, x):
            return x

        def close(self):
            pass

    class FakeMap:
        def __init__(self, ctx, stride):
            self.stride = stride

        def apply(self, delta=None, **kw):
            assert kw["kf_keypoints"].shape[1:] == (self.stride, 2) and kw["kf_pose"].shape[0] == kw["kf_keypoints"].shape[0]
            calls["map_apply"] += 1

        def set_observations(self, obs_ptr, obs_ref, pt_first=0):
            assert obs_ref.dtype == np.int32 and obs_ref.shape == (obs_ptr[-1], 2)
            calls["map_bulk"] += 1

    class FakeSampler:
        def __init__(self, i):
            pass

        def stop(self):
            return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}

    real_tensor = torch.tensor
    monkeypatch.setattr(torch.cuda, "Stream", lambda: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch, "tensor", lambda data, dtype=None, device=None: real_tensor(data, dtype=dtype))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(pkg, "Context", FakeCtx)
    monkeypatch.setattr(pkg, "Frames", FakeFrames)
    monkeypatch.setattr(pkg, "Map", FakeMap)
    monkeypatch.setattr(pkg, "Lattice", FakeLattice)
    monkeypatch.setattr(bench, "ClockSampler", FakeSampler)
    return bench, calls


def test_bench_c3_gpu_arm_against_a_mock_device(monkeypatch, capsys):
    """The default bench arm (C3: device-resident loop, indexed and flat end-to-end loops, per-kernel profile, roofline,
    cpu_baseline) on a reduced problem size against a mock device: guards the host logic and the JSON contract of the
    line the driver parses."""
    import json
    bench, calls = _mock_device(monkeypatch)
    monkeypatch.setitem(bench.WORKLOADS, "c3", ("C3 (reduced for the mock test)", 3, 1500, 8))
    monkeypatch.setattr(bench, "KP_STRIDE", 1024)
    # the blur stress lattice is a 2048 x 2048 grid: shrink the host-side feature assembly, the lattice is a mock anyway
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "1"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["steps"] == 4 and d["warmup"] == 3 and d["gpu_launches"] == 4 * 111 and d["vs_baseline"] is None
    assert d["config"]["problems_per_step_per_gpu"] == 3 and d["config"]["points_per_step_per_gpu"] == 4500
    nnz = 4500 * 8
    # resident map: ids + keypoints + kf_ptr + delta (all poses, all positions, a quarter of the points erased and re-added)
    churn = 3 * 375
    assert d["e2e"]["h2d_bytes_per_step"] == 4500 * 4 + 4500 * 8 + 4 * 4 + 3 * 256 * 48 + 4500 * 12 + churn * (8 + 12)
    indexed = 4500 * 12 + 4501 * 4 + nnz * 4 + 3 * 256 * (48 + 16 + 16) + 4500 * 8 + 4 * 4
    assert d["e2e_indexed_snapshot"]["h2d_bytes_per_step"] == indexed
    assert d["e2e_full_snapshot"]["h2d_bytes_per_step"] == indexed + nnz * (2 + 8 - 4)
    assert d["e2e"]["d2h_bytes_per_step"] == 4500 * (2 + 8)
    assert calls["map_apply"] == 1 and calls["map_bulk"] == 1 and calls["visible"] == 4 + 4
    assert calls["table"] == 1 and calls["indexed"] == 4 + 4 and calls["flat"] == 4 + 4
    assert d["roofline"]["kernel"] == "k_splat_rows+k_scan_compose+k_scan_walk"
    assert abs(d["roofline"]["share_of_step"] - 2.6 / 5.7) < 1e-3 and d["roofline"]["bound"] == "hbm"
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["cpu_baseline"]["value"] > 0


def test_bench_c4_gpu_arm_against_a_mock_device(monkeypatch, capsys):
    """The batched per-frame workload (C4: direct per-frame vectors, lccrf_frames_submit) through the same mock."""
    import json
    bench, calls = _mock_device(monkeypatch)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "c4", "--batch", "6", "--steps", "3", "--warmup", "1"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    NT = d["run_info"]["points_per_step_rank0"]
    assert d["config"]["workload"].startswith("C4") and d["config"]["problems_per_step"] == 6 and d["scaling"] == "strong"
    assert 6 * 4000 <= NT <= 6 * 6000
    assert d["e2e"]["h2d_bytes_per_step"] == NT * (4 + 4 + 4 + 8) and d["e2e"]["d2h_bytes_per_step"] == NT * 10
    assert d["e2e_full_snapshot"] is None and calls["direct"] == 4 + 3 and calls["indexed"] == 0
    assert d["cpu_baseline"]["value"] > 0 and "fits the 126 MB L2" not in d["config"]["l2_policy"]

"""CPU tests: liblccrf.so loads, exports every symbol include/lccrf.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lccrf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lccrf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(pkg):
    lib = pkg.load_library()
    decl = declared_symbols()
    assert len(decl) >= 40
    nm = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (lccrf_[a-z0-9_]+)", nm))
    missing = [s for s in decl if s not in exported]
    assert not missing, "declared in lccrf.h but not exported: %s" % missing
    unbound = [s for s in decl if s not in pkg.SYMBOLS]
    assert not unbound, "declared in lccrf.h but not bound by the ctypes layer: %s" % unbound
    for s in decl:
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.lccrf_version()


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(pkg):
    """Without a usable device the library must fail loudly instead of computing on the host."""
    lib = pkg.load_library()
    if lib.lccrf_device_count() > 0:
        pytest.skip("a GPU is visible; the no-device path is exercised on the CPU box")
    h = ctypes.c_void_p()
    rc = lib.lccrf_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in lib.lccrf_last_error()
    with pytest.raises(pkg.LccrfError):
        pkg.Context(0)


def test_product_does_not_touch_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pk = os.path.join(ROOT, "lc-crf-slam_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dp, f)
    ldd = subprocess.run(["ldd", os.path.join(pk, "liblccrf.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libref" not in ldd


def test_dropin_headers_compile_with_reference_call_sites(tmp_path):
    """tests/cpp/slam_callsite.cpp holds src/Tracking.cc:1919-1930 verbatim; it must compile against the mirror."""
    exe = tmp_path / "slam_callsite"
    cmd = ["g++", "-O1", "-std=c++14", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "lc-crf-slam_b200", "densecrf"), "-o", str(exe),
           os.path.join(ROOT, "tests", "cpp", "slam_callsite.cpp"), "-L" + os.path.join(ROOT, "lc-crf-slam_b200"), "-llccrf",
           "-Wl,-rpath," + os.path.join(ROOT, "lc-crf-slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_map_mirror_binding_compiles_and_links(tmp_path):
    """tests/cpp/map_mirror.cpp is the reference-side binding of the resident map written out in full;
    it must compile against include/lccrf.h and link against liblccrf.so.  Without a B200 the program stops at
    lccrf_ctx_create with exit code 2 (no CPU fallback); tests/test_gpu_dropin.py runs it for real."""
    exe = tmp_path / "map_mirror"
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", str(exe),
           os.path.join(ROOT, "tests", "cpp", "map_mirror.cpp"), "-L" + os.path.join(ROOT, "lc-crf-slam_b200"), "-llccrf",
           "-Wl,-rpath," + os.path.join(ROOT, "lc-crf-slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([str(exe)], capture_output=True, text=True)
        assert run.returncode == 2 and "no CPU fallback" in run.stderr

"""The drop-in C++ shim (include/compat/opencl.h, cl_kernels.h): compiled as C++20 like the reference (CMakeLists.txt:27),
replaying the reference's call sites (tests/cpp/replay_callsites.cpp).  CPU: it must compile and link against
libvp_b200.so.  GPU: its results must equal the CPU oracle's."""
import os
import subprocess

import numpy as np
import pytest

import common
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "vision-processor_b200", "lib")


def build(tmp_path) -> str:
    exe = str(tmp_path / "replay_callsites")
    cmd = ["g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "replay_callsites.cpp"), "-o", exe, "-L", LIBDIR, "-lvp_b200", f"-Wl,-rpath,{LIBDIR}", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_compiles_and_links_as_cxx20(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)  # no arguments: the reference's FATAL = message + exit(1)
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(wq=160, hq=120, fmt=0, k2=0.08, tilt=0.15, n_robots=2, n_balls=2, seed=21), dict(wq=96, hq=64, fmt=1, seed=22),
                                dict(wq=128, hq=96, fmt=2, seed=23)])
def test_call_site_replay_matches_oracle(tmp_path, port, kw):
    exe = build(tmp_path)
    p, raw, _ = common.make_case(**kw)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(bytes(common.to_vp(p)))
        f.write(raw.tobytes())
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0 and "replay ok" in r.stdout, r.stdout + r.stderr
    assert "ms" in r.stdout  # OpenCL::printRuntimes
    want = port.detect(raw, p)
    ch = port.raw2quad(raw, p.fmt, p.wq, p.hq)
    nf, nq = p.wf * p.hf, p.wq * p.hq
    buf = np.fromfile(fout, np.uint8)
    o = 0

    def take(n):
        nonlocal o
        a = buf[o:o + n]
        o += n
        return a

    np.testing.assert_array_equal(take(nf * 4).reshape(p.hf, p.wf, 4), want["flat"])
    np.testing.assert_array_equal(take(nf * 4).view(np.float32).reshape(p.hf, p.wf), want["grad"])
    circ = take(nf * 4).view(np.float32).reshape(p.hf, p.wf)
    common.assert_float_images_equal(circ, want["circ"])
    np.testing.assert_array_equal(take(12).view(np.int32), want["counter"])
    n = int(take(4).view(np.int32)[0])
    common.assert_matches_equal(take(22 * n).view(O.MATCH_DTYPE), want["matches"])
    np.testing.assert_array_equal(take(nf * 3 // 2), port.rgba2nv12(want["flat"])[: nf * 3 // 2])
    np.testing.assert_array_equal(take(nf * 3 // 2), port.f2nv12(want["grad"])[: nf * 3 // 2])
    np.testing.assert_array_equal(take(nq * 3 // 2), port.quad2nv12(ch, p.fmt, 0)[: nq * 3 // 2])
    np.testing.assert_array_equal(take(nq * 4).reshape(p.hq, p.wq, 4), port.quad2rgba(ch, p.fmt, 0))
    pct = take(4).view(np.float32)[0]
    flat_sorted = np.sort(want["circ"].reshape(-1))
    assert pct == flat_sorted[int(nf * np.float32(0.99))]  # blob_benchmark.cpp:190-192
    assert o == buf.size

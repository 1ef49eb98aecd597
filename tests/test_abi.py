"""CPU: the C-ABI library loads, exports every symbol include/vp_b200.h declares, agrees with the Python binding on
struct layouts, and refuses to run without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import subprocess
import sys

import pytest

from vpb200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    declared = lib.declared_symbols()
    assert len(declared) >= 55
    assert lib.bound_symbols() == declared
    dll = lib.load()
    for name in declared:
        assert hasattr(dll, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported
    assert not [s for s in exported if not s.startswith("vp_") and not s.startswith("_")], "only the vp_ C ABI is exported"


def test_pixel_sizes_match_the_reference_table():
    dll = lib.load()  # src/opencl.cpp:24-31: stride * rowStride
    want = {lib.FMT_RGBA8: 4, lib.FMT_U8: 1, lib.FMT_F32: 4, lib.FMT_NV12: 2, lib.FMT_RGGB8: 4, lib.FMT_GRBG8: 4, lib.FMT_BGR8: 3}
    for fmt, size in want.items():
        assert dll.vp_format_pixel_size(fmt) == size
    assert dll.vp_format_pixel_size(99) == 0


def test_struct_layouts_match_the_header(tmp_path):
    exe = tmp_path / "abi_sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "helpers", "abi_sizes.c"), "-o", str(exe)], check=True)
    cam, match, params, o_model, o_h, o_mode, o_circ = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert cam == 72 == C.sizeof(lib.CameraModel)          # src/Perspective.h:22-29
    assert match == 22 == lib.MATCH_DTYPE.itemsize and o_circ == 14  # src/main.cpp:33-41
    assert params == C.sizeof(lib.Params)
    assert o_model == lib.Params.model.offset and o_h == lib.Params.max_robot_height.offset and o_mode == lib.Params.sample_mode.offset
    import oracle as O
    assert C.sizeof(O.Params) == params  # tests convert between the two by memmove


def test_no_gpu_means_no_context():
    """Without a CUDA device vp_ctx_create must fail loudly (VP_ERR_NO_DEVICE), never fall back to the CPU."""
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lib.VpError) as e:
        lib.Context(0)
    assert e.value.code == 5 and "no CPU fallback" in str(e.value)


def test_product_package_does_not_import_the_oracle():
    code = ("import sys; sys.path.insert(0, %r); import vpb200.lib, vpb200.geometry, vpb200.synth, vpb200.shard; "
            "assert 'oracle' not in sys.modules; print('ok')" % os.path.join(ROOT, "vision-processor_b200", "python"))
    assert subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.strip() == "ok"
    for root, _, files in os.walk(os.path.join(ROOT, "vision-processor_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, f)).read()
                assert "vp_oracle" not in text and "import oracle" not in text and "libvp_clref" not in text, f


def test_nv12_surface_contract():
    """the layout an encoder session takes from the batched NV12 conversions -- src/rtpstreamer.cpp:120-121
    (linesize = W for both planes, chroma at +W*H) -- stated by vp_nv12_surface_of; pure arithmetic, no GPU."""
    import ctypes as C
    from vpb200 import lib
    L = lib.load()
    s = lib.Nv12Surface()
    base, w, h, stride = 0x7f0000000000, 1224, 1024, 2 * 1224 * 1024
    assert L.vp_nv12_surface_of(C.c_void_p(base), w, h, stride, 3, C.byref(s)) == 0
    assert s.y == base + 3 * stride and s.uv == s.y + w * h and (s.pitch_y, s.pitch_uv) == (w, w)
    assert (s.width, s.height, s.bytes_used) == (w, h, w * h * 3 // 2) and s.aligned16 == 0   # 1224 is not a multiple of 16
    assert L.vp_nv12_surface_of(C.c_void_p(base), 2048, 1500, 2 * 2048 * 1500, 1, C.byref(s)) == 0 and s.aligned16 == 1
    assert L.vp_nv12_surface_of(C.c_void_p(base), 1223, 1024, stride, 0, C.byref(s)) == 1        # odd width: VP_ERR_INVALID
    assert L.vp_nv12_surface_of(C.c_void_p(base), w, h, w * h, 1, C.byref(s)) == 1               # stride below a frame

"""The C++ host side (include/sdfgpu_viewer.hpp: SDFSurface / SDFViewer / LoadingManager / SDFDemo
mirroring /root/reference/src/sdf/mod.rs:33-126 and src/app/scene/sdf/{mod,loading}.rs) compiled with
g++ against libsdfgpu.so and driven by tests/cpp/host_viewer.cpp the way the reference's scene drives
its SDFViewer (src/app/scene/mod.rs:154-215)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


@pytest.fixture(scope="module")
def host_viewer(S, oracle, tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "host_viewer"
    lib_dir, orc_dir = os.path.join(ROOT, "sdf-viewer_b200"), os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_viewer.cpp"), "-o", str(out),
           "-L" + lib_dir, "-lsdfgpu", "-L" + orc_dir, "-loracle",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + orc_dir, "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return str(out)


def test_reference_loading_tests_in_cpp(host_viewer):
    """loading.rs:117-171, the reference's own unit tests, against sdfgpu::LoadingManager."""
    r = subprocess.run([host_viewer, "loading"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "loading ok" in r.stdout, r.stdout + r.stderr


def test_cpp_demo_tape_equals_python_tape(S, host_viewer):
    for args, want in ((["tape"], S.tape.demo_tape()), (["tape", "nosphere"], S.tape.demo_tape(disable_sphere=True))):
        r = subprocess.run([host_viewer] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        assert bytes.fromhex(r.stdout.strip()) == want


def test_cpp_scalar_program_equals_python(S, host_viewer):
    import test_scalar_programs as P
    r = subprocess.run([host_viewer, "scalar"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert bytes.fromhex(r.stdout.strip()) == P.build_tape(S.tape, P.sphere_with_bands(S.tape))[1]


def test_cpp_wasm_sdf_lowers_like_python(S, host_viewer, tmp_path):
    """sdfgpu::WasmSDF (C++) and sdf_viewer_b200.WasmSDF (Python) are the same call into the library."""
    import test_wasm_lower as W
    wasm = W.guest_csg_calls().build()
    path = tmp_path / "guest.wasm"
    path.write_bytes(wasm)
    r = subprocess.run([host_viewer, "wasm", str(path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    summary, bb, tape_hex = r.stdout.strip().split("\n")
    tape, want_bb, want_summary = S.wasm.lower(wasm)
    assert summary == want_summary and bytes.fromhex(tape_hex) == tape
    assert [float(v) for v in bb.split()] == list(want_bb[0]) + list(want_bb[1])
    path.write_bytes(b"\0asm\x01\0\0\0")            # an empty module: no exports
    r = subprocess.run([host_viewer, "wasm", str(path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and "cannot lower (-1)" in r.stdout


def test_headers_are_self_contained_cxx(tmp_path):
    src = tmp_path / "tu.cpp"
    src.write_text('#include "sdfgpu_viewer.hpp"\nint main() { return sizeof(sdfgpu::SDFSample) == 28 ? 0 : 1; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I" + os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_host_on_gpu(S, oracle, host_viewer, tmp_path):
    r = subprocess.run([host_viewer, "gpu", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stdout + r.stderr

    def vol(name, dims):
        return np.fromfile(tmp_path / name, np.float32).reshape(dims[2], dims[1], dims[0], 4)

    def same(a, b):
        return np.array_equal(a.view(np.uint32), np.ascontiguousarray(b).view(np.uint32))

    # (1) the tape surface: 32^3, 2 passes == the oracle's update of the same tape
    dims = (32, 32, 32)
    o = oracle.Viewer(BB, dims, 2)
    o.update(oracle.Sampler(tape=S.tape.demo_tape()))
    assert same(vol("tape_tex0.bin", dims), o.tex0) and same(vol("tape_tex1.bin", dims), o.tex1)
    w, h = 160, 120
    cam = S.default_camera(w, h)
    P = oracle.trace_params(S.camera_rays(cam, w, h), BB, dims, lod=1.0, filter_linear=1)
    ro, do, _ = oracle.trace(P, o.tex0, o.tex1, w, h)
    rgba8 = np.fromfile(tmp_path / "tape_rgba8.bin", np.uint8).reshape(h, w, 4)
    depth = np.fromfile(tmp_path / "tape_depth.bin", np.float32).reshape(h, w)
    want8 = np.floor(np.clip(ro, 0.0, 1.0) * 255.0 + 0.5)
    assert np.abs(rgba8.astype(np.int32) - want8.astype(np.int32)).max() <= 1   # 1e-5 on the float frame ~ <= 1 code
    assert (rgba8 == want8).mean() > 0.999
    np.testing.assert_allclose(depth, do, rtol=1e-5)
    # after the parameter edit (whole box reported): the new tape everywhere
    o2 = oracle.Viewer(BB, dims, 1)
    o2.fill_all(oracle.Sampler(tape=S.tape.demo_tape(sphere_radius=0.9)))
    assert same(vol("tape_changed_tex0.bin", dims), o2.tex0) and same(vol("tape_changed_tex1.bin", dims), o2.tex1)
    # (2) the surface without a tape, sampled on 4 host threads a millisecond at a time
    dims = (24, 20, 16)
    o3 = oracle.Viewer(BB, dims, 3)
    o3.update(oracle.Sampler())
    assert same(vol("host_tex0.bin", dims), o3.tex0) and same(vol("host_tex1.bin", dims), o3.tex1)

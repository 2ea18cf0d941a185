"""The drop-in boundary: libsdfgpu.so loads, exports every function include/sdfgpu.h declares,
and the device-free entry points behave.  No compute is attempted without a GPU -- and the
library must fail loudly (never fall back to a CPU path) when there is none."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sdfgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdfgpu_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(S):
    from sdf_viewer_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdfgpu.h but not exported by libsdfgpu.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding lists functions the header does not declare"


def test_tape_header_layout_matches_python(S):
    """sdft_instr is 16 bytes, sdft_prim 48, header 32 (include/sdfgpu_tape.h)."""
    T = S.tape
    t = T.demo_tape()
    magic, version, n_instr, n_prims, n_consts = np.frombuffer(t[:20], "<u4")
    assert magic == 0x54464453 and version == 1
    assert (n_instr, n_prims, n_consts) == (5, 2, 7)
    assert len(t) == 32 + 16 * n_instr + 48 * n_prims + 4 * n_consts
    hdr = open(os.path.join(ROOT, "include", "sdfgpu_tape.h")).read()
    for name in ("OP_END", "OP_PRIM", "OP_UNION_PRIM", "OP_INTER_PRIM", "OP_UNION_RANGE", "OP_PUSH", "OP_POP_UNION",
                 "OP_POP_INTER", "OP_POP_DEMO_DIFF", "OP_D_NEG", "OP_D_ABS", "OP_D_ADD", "OP_D_MUL", "OP_D_MAX",
                 "OP_D_MIN", "OP_M_SET", "OP_P_RESET", "OP_P_SUB", "OP_P_MUL", "OP_P_ABS"):
        m = re.search(r"SDFT_%s\s*=\s*(\d+)" % name, hdr)
        assert m and int(m.group(1)) == getattr(T, name), name


def test_device_free_entry_points(S, oracle):
    from sdf_viewer_b200 import _lib
    lib = _lib.load()
    assert lib.sdfgpu_air_dist() == oracle.lib().orc_air_dist() == np.float32(0.1) + np.float32(0.001234)
    # from_bb's voxel rule (scene/sdf/mod.rs:47-68) against the oracle, incl. ties (last max wins) and truncation
    for bb in [((-1, -1, -1), (1, 1, 1)), ((0, 0, 0), (1, 2, 3)), ((0, 0, 0), (3, 2, 1)), ((-1, 0, 0), (1, 2, 0.7)),
               ((0, 0, 0), (2, 2, 1)), ((0, 0, 0), (1, 0.333, 0.9999))]:
        for side in (1, 7, 64, 100, 512):
            od = (C.c_uint32 * 3)()
            oracle.lib().orc_dims_from_bb(oracle._bb6(bb), side, od)
            assert S.dims_from_bb(bb, side) == tuple(od)
    # camera: cgmath look_at_rh / perspective as the oracle restates them
    cam = S.default_camera(640, 480)
    v, p = (C.c_float * 16)(), (C.c_float * 16)()
    oracle.lib().orc_look_at_rh(oracle._f([2.5, 3, 5]), oracle._f([0, 0, 0]), oracle._f([0, 1, 0]), v)
    oracle.lib().orc_perspective(np.float32(45.0 * np.pi / 180.0), np.float32(640 / 480), 0.1, 1000.0, p)
    np.testing.assert_allclose(list(cam.view), list(v), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(list(cam.projection), list(p), rtol=1e-6, atol=1e-7)
    assert (cam.tone_mapping, cam.color_mapping, cam.gamma) == (2, 1, 0.0)
    rays = S.camera_rays(cam, 640, 480)
    # the centre ray looks at the origin: camera at (2.5,3,5) -> direction ~ -position
    d = np.array(rays.base) + np.array(rays.dx) * 320 + np.array(rays.dy) * 240
    pos = np.array(cam.position)
    assert np.allclose(d / np.linalg.norm(d), -pos / np.linalg.norm(pos), atol=1e-5)
    # unprojecting then projecting with BVP lands on the pixel (bias maps NDC to [0,1])
    for (i, j) in ((0.5, 0.5), (100.5, 300.5), (639.5, 479.5)):
        dd = np.array(rays.base) + np.array(rays.dx) * i + np.array(rays.dy) * j
        q = np.array(rays.bvp).reshape(4, 4).T @ np.append(pos + 3.0 * dd, 1.0)
        assert np.allclose(q[:2] / q[3], [i / 640, j / 480], atol=1e-4)


def test_no_cpu_fallback_without_device(S):
    """On a box without a CUDA device every constructor fails with SDFGPU_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SdfGpuError) as e:
        S.SDFViewer.from_bb(((-1, -1, -1), (1, 1, 1)), 16, 2)
    assert e.value.code == -2 and "no CPU fallback" in e.value.message
    # ... the slab and group constructors included (every compute entry point needs a handle they would have made)
    with pytest.raises(S.SdfGpuError) as e:
        S.SDFViewer.new_voxels((16, 16, 16), ((-1, -1, -1), (1, 1, 1)), 1, z_range=(0, 8))
    assert e.value.code == -2
    for make in (lambda: S.SDFViewerGroup.from_bb(((-1, -1, -1), (1, 1, 1)), 16, 1, device_mask=3),
                 lambda: S.SDFViewerGroup.new_voxels((16, 16, 16), ((-1, -1, -1), (1, 1, 1)), 1, [0, 0])):
        with pytest.raises(S.SdfGpuError) as e:
            make()
        assert e.value.code == -2 and "no CPU fallback" in e.value.message


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under sdf-viewer_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sdf-viewer_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(d, f)).read()
                assert "liboracle" not in text and "import orc" not in text and "sdf_oracle" not in text, f


def test_specialised_kernel_compiles_for_sm100a(S):
    """sdfgpu_jit_check: the tape structure -> straight-line kernel source -> NVRTC cubin for sm_100a,
    from the same device source as the ahead-of-time kernels.  Needs no GPU."""
    src = S.jit_check(S.tape.demo_tape(), 2)
    assert "SDFGPU_STEP(5, 0) SDFGPU_STEP(20, 1) SDFGPU_STEP(3, 2) SDFGPU_STEP(24, 3)" in src
    assert "PROG_JIT" in src
    src = S.jit_check(S.tape.csg_tape(S.tape.csg_primitive_table(20)), 8)
    assert "SDFGPU_STEP(19, 0) SDFGPU_STEP(13, 1)" in src
    with pytest.raises(S.SdfGpuError) as e:
        S.jit_check(b"\0" * 40, 2)
    assert e.value.code == -3


C_CLIENT = r"""
#include "sdfgpu.h"
#include "sdfgpu_tape.h"
#include <stdio.h>
int main(void) {
    float bb[6] = {-1, -1, -1, 1, 2, 0.5f};
    uint32_t d[3];
    sdfgpu_camera cam;
    sdfgpu_rays rays;
    sdfgpu_ctx* ctx = 0;
    int rc = sdfgpu_dims_from_bb(bb, 64, d);
    sdfgpu_camera_default(&cam, 640, 480);
    rc |= sdfgpu_camera_rays(&cam, 640, 480, &rays);
    printf("%d %u %u %u %zu %zu %zu %zu %.9g\n", rc, d[0], d[1], d[2], sizeof(sdft_header), sizeof(sdft_instr),
           sizeof(sdft_prim), sizeof(sdfgpu_camera), (double)sdfgpu_air_dist());
    rc = sdfgpu_create(bb, 64, 2, 0, &ctx);   /* fails without a device: no CPU fallback */
    printf("%d %s\n", rc, sdfgpu_last_error(ctx));
    sdfgpu_destroy(ctx);
    return 0;
}
"""


def test_plain_c_client_links_against_the_abi(S, tmp_path):
    """include/*.h are C99 headers and libsdfgpu.so is a plain C ABI: a C program compiled with gcc
    -std=c99 -pedantic links and calls it (what a cgo / Rust `extern "C"` binding does)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "client.c"
    src.write_text(C_CLIENT)
    exe = tmp_path / "client"
    libdir = os.path.join(ROOT, "sdf-viewer_b200")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", libdir, "-lsdfgpu", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines()
    f = out[0].split()
    assert f[:4] == ["0", "42", "64", "32"]                      # from_bb: sizes (2, 3, 1.5) -> (64*2/3 truncated, 64, 64*1.5/3)
    assert f[4:8] == ["32", "16", "48", "180"]                   # struct sizes of the tape format and the camera
    assert abs(float(f[8]) - 0.101234004) < 1e-9
    import torch
    if not torch.cuda.is_available():
        assert out[1].startswith("-2 ") and "no CPU fallback" in out[1]
    else:
        assert out[1].startswith("0 ")


def test_mutated_tapes_are_rejected_or_accepted_never_crash(S):
    """Tapes are host input: random corruptions of valid tapes (primitive, CSG and scalar-program ones) go through
    the validator without crashing; truncations are always rejected."""
    T = S.tape
    p = T.ScalarProgram()
    x = p.px()
    p.out(0, p.op("FSUB", p.op("FABS", x), p.const(0.5)))
    t = T.TapeBuilder()
    t.scalar(p).emit(T.OP_END)
    seeds = [T.demo_tape(), T.csg_tape(T.csg_primitive_table(20)), t.build()]
    for s in seeds:
        S.tape_validate(s)
    rng = np.random.default_rng(99)
    rejected = 0
    for it in range(6000):
        w = bytearray(seeds[it % 3])
        if it % 3 == 0:
            w = w[:int(rng.integers(0, len(w)))]
        else:
            for _ in range(int(rng.integers(1, 5))):
                w[int(rng.integers(0, len(w)))] = int(rng.integers(0, 256))
        try:
            S.tape_validate(bytes(w))
        except S.SdfGpuError as e:
            rejected += 1
            assert e.code in (-1, -3)
            assert it % 3 != 0 or e.code in (-1, -3)
        else:
            assert it % 3 != 0, "a truncated tape was accepted"
    assert rejected > 2500


def test_surface_callback_table_from_a_python_surface(S, oracle):
    """`sdfgpu_surface` built from a Python SDFSurface (viewer._surface_struct): the trampolines marshal points,
    samples, the changed box and the tape exactly, and park exceptions instead of letting them cross the C ABI
    (the failing samples take the reference's benign value, src/sdf/wasm/native.rs:202)."""
    from sdf_viewer_b200 import viewer

    class Host(S.SDFSurface):
        def bounding_box(self):
            return ((-1, -2, -3), (1, 2, 3))

        def sample(self, pts, distance_only=False):
            return oracle.demo_sample(pts)

        def changed(self):
            return (0, 0, 0, 0.5, 0.25, 1)

    s = viewer._surface_struct(Host())
    pts = np.random.default_rng(1).uniform(-1, 1, (17, 3)).astype(np.float32)
    out = np.zeros((17, 7), np.float32)
    fp = C.POINTER(C.c_float)
    s.sample_batch(None, pts.ctypes.data_as(fp), 17, 0, out.ctypes.data_as(fp))
    assert np.array_equal(out.view(np.uint32), oracle.demo_sample(pts).view(np.uint32))
    box = (C.c_float * 6)()
    assert s.changed(None, box) == 1 and list(box) == [0, 0, 0, 0.5, 0.25, 1]
    bb = (C.c_float * 6)()
    s.bounding_box(None, bb)
    assert list(bb) == [-1, -2, -3, 1, 2, 3]
    ptr, n = C.c_void_p(), C.c_size_t()
    assert s.tape(None, C.byref(ptr), C.byref(n)) == 0 and not bool(s.sample)      # no tape, no single-point form
    assert s.sample_threads == 1 and s._py_error[0] is None
    # a surface with a tape hands its bytes over unchanged
    d = viewer._surface_struct(S.SDFDemo())
    assert d.tape(None, C.byref(ptr), C.byref(n)) == 1 and C.string_at(ptr.value, n.value) == S.SDFDemo().tape()

    class Broken(Host):
        def sample(self, pts, distance_only=False):
            raise RuntimeError("guest trapped")

        def changed(self):
            raise ValueError("also broken")

    b = viewer._surface_struct(Broken())
    out[:] = 7
    b.sample_batch(None, pts.ctypes.data_as(fp), 17, 0, out.ctypes.data_as(fp))
    assert np.all(out[:, 0] == 1.0) and np.all(out[:, 1:] == 0.0)
    assert b.changed(None, box) == 0
    assert isinstance(b._py_error[0], RuntimeError)          # the first exception is the one re-raised by update_surface


def test_camera_rays_match_the_oracles_own_derivation(S, oracle):
    """sdfgpu_camera_rays against the ray basis the oracle derives from its own cgmath restatement (orc.camera_rays):
    the frame tests hand the product's rays to the oracle, this pins the rays themselves."""
    for (w, h, eye, target) in ((640, 480, (2.5, 3.0, 5.0), (0, 0, 0)), (1920, 1080, (2.5, 3.0, 5.0), (0, 0, 0)),
                                (333, 211, (0.2, 0.1, 0.3), (1, 0.2, -0.4))):
        a = oracle.camera_rays(eye, target, (0, 1, 0), 45.0, w, h)
        b = S.camera_rays(S.look_at_camera(eye, target, w, h), w, h)
        for k in ("origin", "base", "dx", "dy", "bvp"):
            np.testing.assert_allclose(np.array(list(getattr(a, k)), np.float32), np.array(list(getattr(b, k)), np.float32),
                                       rtol=2e-6, atol=1e-7, err_msg=k)

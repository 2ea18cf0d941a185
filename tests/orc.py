"""ctypes wrapper of the CPU oracle (oracle/sdf_oracle.cpp).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")
GBUF_FLOATS = 16


def build(force=False):
    src = os.path.join(ORACLE_DIR, "sdf_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "sdfgpu_tape.h")
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in (src, hdr))
    if force or stale:
        r = subprocess.run(["make", "-C", ORACLE_DIR, "-B", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


class DemoParams(C.Structure):
    _fields_ = [("cube_half_side", C.c_float), ("cube_material", C.c_uint32), ("sphere_radius", C.c_float),
                ("sphere_material", C.c_uint32), ("max_distance_custom_material", C.c_float),
                ("disable_sphere", C.c_uint32)]


class TraceParams(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("base", C.c_float * 3), ("dx", C.c_float * 3), ("dy", C.c_float * 3),
                ("bvp", C.c_float * 16), ("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("dims", C.c_uint32 * 3),
                ("lod", C.c_float), ("filter_linear", C.c_uint32), ("tint", C.c_float * 4),
                ("tone_mapping", C.c_uint32), ("color_mapping", C.c_uint32), ("gamma", C.c_float),
                ("ambient", C.c_float * 3), ("z_lo", C.c_uint32), ("z_hi", C.c_uint32),
                ("clip_min", C.c_float * 3), ("clip_max", C.c_float * 3)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, fp = C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_float)
    sig = {
        "orc_air_dist": (C.c_float, []),
        "orc_srgb_lut": (None, [fp]),
        "orc_f32_to_u8": (u32, [C.c_float]),
        "orc_demo_params_default": (None, [C.POINTER(DemoParams)]),
        "orc_demo_sample": (None, [C.POINTER(DemoParams), vp, u64, C.c_int, vp]),
        "orc_tape_sample": (C.c_int, [vp, u64, vp, u64, vp]),
        "orc_lm_new": (vp, [C.POINTER(u64), u64]),
        "orc_lm_free": (None, [vp]),
        "orc_lm_next": (C.c_int, [vp, C.POINTER(u64)]),
        "orc_lm_len": (u64, [vp]),
        "orc_lm_total_iterations": (u64, [vp]),
        "orc_lm_passes_left": (u64, [vp]),
        "orc_prev_power_of_2": (u32, [u32]),
        "orc_dims_from_bb": (None, [fp, u32, C.POINTER(u32)]),
        "orc_viewer_new": (vp, [fp, C.POINTER(u32), u64]),
        "orc_viewer_new_noalloc": (vp, [fp, C.POINTER(u32), u64]),
        "orc_viewer_free": (None, [vp]),
        "orc_viewer_tex0": (fp, [vp]),
        "orc_viewer_tex1": (fp, [vp]),
        "orc_viewer_len": (u64, [vp]),
        "orc_viewer_total_iterations": (u64, [vp]),
        "orc_viewer_passes_left": (u64, [vp]),
        "orc_voxel_pos": (None, [vp, u64, u64, u64, fp]),
        "orc_sampler_demo": (vp, [C.POINTER(DemoParams)]),
        "orc_sampler_tape": (vp, [vp, u64]),
        "orc_sampler_free": (None, [vp]),
        "orc_viewer_update": (u64, [vp, vp, fp, u64]),
        "orc_viewer_fill_all": (u64, [vp, vp, u32, u32, C.c_int]),
        "orc_viewer_sample_voxels": (None, [vp, vp, vp, u64, vp, vp, C.c_int]),
        "orc_max_threads": (C.c_int, []),
        "orc_look_at_rh": (None, [fp, fp, fp, fp]),
        "orc_perspective": (None, [C.c_float, C.c_float, C.c_float, C.c_float, fp]),
        "orc_trace_params_size": (u64, []),
        "orc_trace": (None, [C.POINTER(TraceParams), vp, vp, u32, u32, u32, u32, vp, vp, vp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    assert L.orc_trace_params_size() == C.sizeof(TraceParams), "TraceParams layout drifted"
    _lib = L
    return L


def _f(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def _bb6(bb):
    return _f(list(bb[0]) + list(bb[1]) if len(bb) == 2 else list(bb))


def demo_params(**kw):
    P = DemoParams()
    lib().orc_demo_params_default(C.byref(P))
    for k, v in kw.items():
        setattr(P, k, v)
    return P


def demo_sample(points, params=None, distance_only=False):
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.empty((len(pts), 7), np.float32)
    P = params or demo_params()
    lib().orc_demo_sample(C.byref(P), pts.ctypes.data, len(pts), int(distance_only), out.ctypes.data)
    return out


def tape_sample(tape, points):
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.empty((len(pts), 7), np.float32)
    buf = (C.c_char * len(tape)).from_buffer_copy(tape)
    rc = lib().orc_tape_sample(buf, len(tape), pts.ctypes.data, len(pts), out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle rejected the tape")
    return out


class Sampler:
    def __init__(self, tape=None, params=None):
        L = lib()
        if tape is not None:
            self._buf = (C.c_char * len(tape)).from_buffer_copy(tape)
            self.h = L.orc_sampler_tape(self._buf, len(tape))
            if not self.h:
                raise ValueError("oracle rejected the tape")
        else:
            P = params or demo_params()
            self.h = L.orc_sampler_demo(C.byref(P))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sampler_free(self.h)
            self.h = None


class Viewer:
    """The oracle's SDFViewer (scene/sdf/mod.rs)."""

    def __init__(self, bb, dims, passes, alloc=True):
        self.bb, self.dims = bb, tuple(int(d) for d in dims)
        new = lib().orc_viewer_new if alloc else lib().orc_viewer_new_noalloc
        self.h = new(_bb6(bb), (C.c_uint32 * 3)(*self.dims), int(passes))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_viewer_free(self.h)
            self.h = None

    def _tex(self, fn):
        n = self.dims[0] * self.dims[1] * self.dims[2] * 4
        if n == 0:
            return np.zeros((self.dims[2], self.dims[1], self.dims[0], 4), np.float32)
        a = np.ctypeslib.as_array(fn(self.h), shape=(n,))
        return a.reshape(self.dims[2], self.dims[1], self.dims[0], 4)

    @property
    def tex0(self):
        return self._tex(lib().orc_viewer_tex0)

    @property
    def tex1(self):
        return self._tex(lib().orc_viewer_tex1)

    def update(self, sampler, changed=None, max_iterations=0):
        box = _bb6(changed) if changed is not None else None
        return lib().orc_viewer_update(self.h, sampler.h, box, int(max_iterations))

    def fill_all(self, sampler, z0=0, z1=None, threads=0):
        return lib().orc_viewer_fill_all(self.h, sampler.h, int(z0), int(self.dims[2] if z1 is None else z1), int(threads))

    def sample_voxels(self, sampler, idx, threads=0):
        """Stored texels (tex0, tex1) of the voxels idx[n,3] -- computed point by point, no volume needed."""
        idx = np.ascontiguousarray(idx, np.uint32).reshape(-1, 3)
        o0 = np.empty((len(idx), 4), np.float32)
        o1 = np.empty((len(idx), 4), np.float32)
        lib().orc_viewer_sample_voxels(self.h, sampler.h, idx.ctypes.data, len(idx), o0.ctypes.data, o1.ctypes.data,
                                       int(threads))
        return o0, o1

    def len(self):
        return lib().orc_viewer_len(self.h)

    def total_iterations(self):
        return lib().orc_viewer_total_iterations(self.h)

    def passes_left(self):
        return lib().orc_viewer_passes_left(self.h)

    def voxel_pos(self, x, y, z):
        out = (C.c_float * 3)()
        lib().orc_voxel_pos(self.h, x, y, z, out)
        return tuple(out)


def trace_params(rays, bb, dims, lod=1.0, filter_linear=1, tint=(1, 1, 1, 1), tone_mapping=2, color_mapping=1,
                 gamma=0.0, ambient=(1, 1, 1), z_lo=0, z_hi=None, clip_min=None, clip_max=None):
    """Build the oracle's TraceParams from an sdfgpu_rays-like object (origin/base/dx/dy/bvp)."""
    P = TraceParams()
    bmin, bmax = (list(bb[0]), list(bb[1])) if len(bb) == 2 else (list(bb[:3]), list(bb[3:]))
    P.origin[:] = list(rays.origin); P.base[:] = list(rays.base)
    P.dx[:] = list(rays.dx); P.dy[:] = list(rays.dy); P.bvp[:] = list(rays.bvp)
    P.bmin[:] = bmin; P.bmax[:] = bmax
    P.dims[:] = [int(d) for d in dims]
    P.lod = lod; P.filter_linear = int(filter_linear)
    P.tint[:] = list(tint)
    P.tone_mapping, P.color_mapping, P.gamma = tone_mapping, color_mapping, gamma
    P.ambient[:] = list(ambient)
    P.z_lo, P.z_hi = int(z_lo), int(dims[2] if z_hi is None else z_hi)
    P.clip_min[:] = list(clip_min if clip_min is not None else bmin)
    P.clip_max[:] = list(clip_max if clip_max is not None else bmax)
    return P


def trace(P, tex0, tex1, width, height, rows=None, rgba=True, depth=True, gbuf=True, threads=0):
    t0 = np.ascontiguousarray(tex0, np.float32)
    t1 = np.ascontiguousarray(tex1, np.float32)
    r = np.zeros((height, width, 4), np.float32) if rgba else None
    d = np.ones((height, width), np.float32) if depth else None
    g = np.zeros((height, width, GBUF_FLOATS), np.float32) if gbuf else None
    r0, r1 = rows if rows is not None else (0, height)
    lib().orc_trace(C.byref(P), t0.ctypes.data, t1.ctypes.data, width, height, r0, r1,
                    r.ctypes.data if r is not None else None, d.ctypes.data if d is not None else None,
                    g.ctypes.data if g is not None else None, int(threads))
    return r, d, g


class LM:
    def __init__(self, limits, passes):
        self.h = lib().orc_lm_new((C.c_uint64 * 3)(*limits), passes)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lm_free(self.h)
            self.h = None

    def next(self):
        out = (C.c_uint64 * 3)()
        return tuple(out) if lib().orc_lm_next(self.h, out) else None

    def len(self):
        return lib().orc_lm_len(self.h)

    def total_iterations(self):
        return lib().orc_lm_total_iterations(self.h)

    def passes_left(self):
        return lib().orc_lm_passes_left(self.h)


class RaysOracle:
    """origin / base / dx / dy / bvp of a camera, as trace_params() takes them (the fields of sdfgpu_rays)."""


def camera_rays(eye, center, up, fovy_deg, width, height, z_near=0.1, z_far=1000.0):
    """Per-pixel ray basis and the biased view-projection matrix from the oracle's own cgmath restatement
    (orc_look_at_rh / orc_perspective; scene/mod.rs:82-95, material.rs:88-95), in f32: pixel (i, j) looks along
    base + dx * (i + 0.5) + dy * (j + 0.5).  Independent of the product's sdfgpu_camera_rays (bench.py's reference
    arm must not load the product; tests compare the two)."""
    f32 = np.float32
    V, Pm = (C.c_float * 16)(), (C.c_float * 16)()
    lib().orc_look_at_rh(_f(eye), _f(center), _f(up), V)
    lib().orc_perspective(f32(fovy_deg) * f32(np.pi) / f32(180.0), f32(width) / f32(height), z_near, z_far, Pm)
    V = np.array(list(V), f32); Pm = np.array(list(Pm), f32)
    s, u, f = V[[0, 4, 8]], V[[1, 5, 9]], -V[[2, 6, 10]]
    sx, sy = f32(1.0) / Pm[0], f32(1.0) / Pm[5]
    r = RaysOracle()
    r.origin = [f32(x) for x in eye]
    r.base = list((f - s * sx) - u * sy)
    r.dx = list(s * (f32(2.0) * sx / f32(width)))
    r.dy = list(u * (f32(2.0) * sy / f32(height)))
    bias = np.array([0.5, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 0.5, 0, 0.5, 0.5, 0.5, 1.0], f32).reshape(4, 4).T  # column-major
    pv = Pm.reshape(4, 4).T @ V.reshape(4, 4).T
    r.bvp = list((bias @ pv).T.reshape(-1).astype(f32))
    return r


def default_rays(width, height):
    """The scene's default camera (scene/mod.rs:89-94): eye (2.5, 3, 5), target origin, 45 degrees."""
    return camera_rays((2.5, 3.0, 5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, width, height)

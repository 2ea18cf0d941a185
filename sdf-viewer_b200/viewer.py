"""Host mirror of `SDFViewer` (/root/reference/src/app/scene/sdf/mod.rs:21-251) over the C ABI
of include/sdfgpu.h: same constructor names, `update` / `commit`, `loading_mgr` fields read by the
scene (`src/app/scene/mod.rs:149-153,229-239`), plus `trace`, which stands where the scene calls
`volume.render(&camera, lights)` (`scene/mod.rs:213-215`) with `SDFViewerMaterial`.

Everything here is a thin call into libsdfgpu.so; no arithmetic of the hot path happens in Python.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Camera, Rays, GBUF_FLOATS, LINK_BLOB_BYTES, LINK_GBUF, LINK_HALO_PUSH, LINK_ROUNDS, LINK_STREAM, SdfGpuError, check, check_group  # noqa: F401


def _f6(bb):
    flat = [float(v) for v in (list(bb[0]) + list(bb[1]) if len(bb) == 2 else list(bb))]
    if len(flat) != 6:
        raise ValueError("bounding box must be ((minx,miny,minz),(maxx,maxy,maxz)) or 6 floats")
    return (C.c_float * 6)(*flat)


def _host_ptr(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


def dims_from_bb(bb, max_voxels_side):
    """SDFViewer::from_bb's voxel-count rule (scene/sdf/mod.rs:47-68)."""
    out = (C.c_uint32 * 3)()
    check(_lib.load().sdfgpu_dims_from_bb(_f6(bb), int(max_voxels_side), out))
    return tuple(out)


def default_camera(width, height):
    cam = Camera()
    _lib.load().sdfgpu_camera_default(C.byref(cam), int(width), int(height))
    return cam


def look_at_camera(eye, target, width, height, up=(0.0, 1.0, 0.0), fovy_deg=45.0, z_near=0.1, z_far=1000.0):
    """A camera like the scene's (scene/mod.rs:82-95) at another pose (CameraController output)."""
    lib = _lib.load()
    cam = default_camera(width, height)
    e, t, u = ((C.c_float * 3)(*[float(x) for x in v]) for v in (eye, target, up))
    cam.position[:] = list(e)
    lib.sdfgpu_look_at_rh(e, t, u, cam.view)
    lib.sdfgpu_perspective(C.c_float(np.float32(fovy_deg) * np.float32(np.pi) / np.float32(180.0)),
                           C.c_float(float(width) / float(height)), z_near, z_far, cam.projection)
    return cam


def camera_rays(cam, width, height):
    rays = Rays()
    check(_lib.load().sdfgpu_camera_rays(C.byref(cam), int(width), int(height), C.byref(rays)))
    return rays


def _surface_struct(sdf):
    """An `sdf.SDFSurface` as the callback table of include/sdfgpu.h (`sdfgpu_surface`).  Exceptions raised
    by the Python callbacks cannot cross the C ABI: they are kept and re-raised by the caller, and the
    callback returns the reference's benign value (distance 1.0, src/sdf/wasm/native.rs:202)."""
    surf = _lib.Surface()
    error = [None]
    keep = []

    def bounding_box(_self, out):
        try:
            bb = sdf.bounding_box()
            flat = list(bb[0]) + list(bb[1]) if len(bb) == 2 else list(bb)
            for i in range(6):
                out[i] = float(flat[i])
        except Exception as e:  # noqa: BLE001
            error[0] = error[0] or e

    def sample_batch(_self, xyz, n, distance_only, out):
        o = np.ctypeslib.as_array(out, shape=(int(n), 7))
        try:
            pts = np.ctypeslib.as_array(xyz, shape=(int(n), 3))
            o[:] = np.asarray(sdf.sample(pts, bool(distance_only)), np.float32).reshape(int(n), 7)
        except Exception as e:  # noqa: BLE001
            error[0] = error[0] or e
            o[:] = 0.0
            o[:, 0] = 1.0

    def changed(_self, out):
        try:
            bb = sdf.changed()
        except Exception as e:  # noqa: BLE001
            error[0] = error[0] or e
            return 0
        if bb is None:
            return 0
        flat = list(bb[0]) + list(bb[1]) if len(bb) == 2 else list(bb)
        for i in range(6):
            out[i] = float(flat[i])
        return 1

    def tape(_self, bytes_out, len_out):
        try:
            t = sdf.tape()
        except NotImplementedError:
            return 0
        except Exception as e:  # noqa: BLE001
            error[0] = error[0] or e
            return 0
        if not t:
            return 0
        buf = C.create_string_buffer(bytes(t), len(t))
        keep.append(buf)
        bytes_out[0] = C.cast(buf, C.c_void_p).value
        len_out[0] = len(t)
        return 1

    surf.self = None
    surf.bounding_box = _lib.BOUNDING_BOX_FN(bounding_box)
    surf.sample_batch = _lib.SAMPLE_BATCH_FN(sample_batch)
    surf.changed = _lib.CHANGED_FN(changed)
    surf.tape = _lib.TAPE_FN(tape)
    surf.sample_threads = 1  # Python callbacks hold the GIL
    surf._py_error = error
    surf._keep = keep
    return surf


class _LoadingView:
    """`viewer.loading_mgr` as the scene reads it (scene/mod.rs:153,229-239)."""

    def __init__(self, viewer):
        self._v = viewer

    def _state(self):
        ln, tot = C.c_uint64(), C.c_uint64()
        left, passes = C.c_uint32(), C.c_uint32()
        check(self._v._lib.sdfgpu_loading_state(self._v._h, C.byref(ln), C.byref(tot), C.byref(left), C.byref(passes)),
              self._v._h)
        return ln.value, tot.value, left.value, passes.value

    def __len__(self):
        return self._state()[0]

    def total_iterations(self):
        return self._state()[1]

    def passes_left(self):
        return self._state()[2]

    @property
    def passes(self):
        return self._state()[3]

    @property
    def limits(self):
        return self._v.dims


class SDFViewer:
    def __init__(self, handle, bounding_box):
        self._lib = _lib.load()
        self._h = handle
        self.bounding_box = bounding_box
        d = (C.c_uint32 * 3)()
        check(self._lib.sdfgpu_dims(self._h, d), self._h)
        self.dims = tuple(d)
        zb, ze, zl, zh = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(self._lib.sdfgpu_slab(self._h, C.byref(zb), C.byref(ze), C.byref(zl), C.byref(zh)), self._h)
        self.z_begin, self.z_end, self.z_lo, self.z_hi = zb.value, ze.value, zl.value, zh.value
        self.loading_mgr = _LoadingView(self)
        self._tape = None

    # ---- constructors (scene/sdf/mod.rs:46-101)
    @classmethod
    def from_bb(cls, bb, max_voxels_side, loading_passes, device=0):
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.sdfgpu_create(_f6(bb), int(max_voxels_side), int(loading_passes), int(device), C.byref(h)))
        return cls(h, bb)

    @classmethod
    def new_voxels(cls, voxels, bb, loading_passes, device=0, z_range=None):
        lib = _lib.load()
        h = C.c_void_p()
        v = (C.c_uint32 * 3)(*[int(x) for x in voxels])
        if z_range is None:
            check(lib.sdfgpu_create_voxels(_f6(bb), v, int(loading_passes), int(device), C.byref(h)))
        else:
            check(lib.sdfgpu_create_slab(_f6(bb), v, int(loading_passes), int(device), int(z_range[0]),
                                         int(z_range[1]), C.byref(h)))
        return cls(h, bb)

    def close(self):
        if self._h:
            self._lib.sdfgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- tex0.width/height/depth (scene/mod.rs:149-150)
    @property
    def width(self):
        return self.dims[0]

    @property
    def height(self):
        return self.dims[1]

    @property
    def depth(self):
        return self.dims[2]

    # ---- fill
    def set_tape(self, tape_bytes):
        buf = (C.c_char * len(tape_bytes)).from_buffer_copy(tape_bytes)
        check(self._lib.sdfgpu_set_tape(self._h, buf, len(tape_bytes)), self._h)
        self._tape = bytes(tape_bytes)

    def update(self, sdf, max_passes=0):
        """SDFViewer::update (scene/sdf/mod.rs:128-217).  `sdf` is an `sdf.SDFSurface` (its tape is
        re-sent when it reports a change) or None to keep the current tape.  `max_passes` replaces
        `max_delta_time`: a LoadingManager pass is one kernel launch.  Returns the iterations done."""
        changed = None
        if sdf is not None:
            changed = sdf.changed()
            # the reference samples whatever surface it is handed (scene/sdf/mod.rs:128): another surface, or the
            # same one with new parameters, must not be evaluated through the previous tape
            tape = bytes(sdf.tape())
            if self._tape is None or tape != self._tape:
                self.set_tape(tape)
        it = C.c_uint64()
        box = _f6(changed) if changed is not None else None
        check(self._lib.sdfgpu_update(self._h, box, int(max_passes), C.byref(it)), self._h)
        return it.value

    def update_surface(self, sdf, max_delta_time=0.030):
        """SDFViewer::update(sdf, max_delta_time) with the trait object itself (scene/sdf/mod.rs:128-217;
        the scene passes 30 ms, scene/mod.rs:168).  A surface whose `tape()` returns bytes is evaluated on
        the GPU; any other surface is sampled on the host through `sample(points)` in the reference's visit
        order for at most `max_delta_time` seconds and the results are scattered into the volumes.
        Returns the LoadingManager iterations done."""
        surf = _surface_struct(sdf)
        it = C.c_uint64()
        check(self._lib.sdfgpu_update_surface(self._h, C.byref(surf), float(max_delta_time), C.byref(it)), self._h)
        err = getattr(surf, "_py_error", None)
        if err and err[0] is not None:
            raise err[0]
        return it.value

    def fill_all(self):
        check(self._lib.sdfgpu_fill_all(self._h), self._h)

    def resample_box(self, box, count=False):
        n = C.c_uint64()
        check(self._lib.sdfgpu_resample_box(self._h, _f6(box), C.byref(n) if count else None), self._h)
        return n.value if count else None

    def voxel_positions(self, first_flat, count):
        """Positions (count, 3) of the voxels [first_flat, first_flat + count) in flat order."""
        xyz = np.empty((int(count), 3), np.float32)
        check(self._lib.sdfgpu_voxel_positions(self._h, int(first_flat), int(count), _host_ptr(xyz)), self._h)
        return xyz

    def ingest_samples(self, first_flat, samples):
        """Batched ingest of host-computed `SDFSample` records ((n, 7) float32) for the voxels that start
        at flat index `first_flat` (SDFs without a tape, e.g. any existing .wasm SDFSurface)."""
        s = np.ascontiguousarray(samples, np.float32).reshape(-1, 7)
        check(self._lib.sdfgpu_ingest_samples(self._h, int(first_flat), len(s), _host_ptr(s)), self._h)

    def commit(self):  # scene/sdf/mod.rs:220-239
        check(self._lib.sdfgpu_commit(self._h), self._h)

    def reset(self, loading_passes):
        check(self._lib.sdfgpu_reset(self._h, int(loading_passes)), self._h)
        self._tape = None  # set_sdf rebuilds the viewer (scene/mod.rs:154-155): the next update sends its surface's tape

    def download(self, tex0=True, tex1=True, out0=None, out1=None):
        """The CPU-side `Vec<[f32;4]>` volumes (scene/sdf/mod.rs:23-25) of this handle's own slices,
        shaped (depth, height, width, 4)."""
        shape = (self.z_end - self.z_begin, self.dims[1], self.dims[0], 4)
        a0 = (out0 if out0 is not None else np.empty(shape, np.float32)) if tex0 else None
        a1 = (out1 if out1 is not None else np.empty(shape, np.float32)) if tex1 else None
        check(self._lib.sdfgpu_download(self._h, _host_ptr(a0), _host_ptr(a1)), self._h)
        return a0, a1

    def device_ptrs(self):
        p0, p1 = C.c_void_p(), C.c_void_p()
        check(self._lib.sdfgpu_device_ptrs(self._h, C.byref(p0), C.byref(p1)), self._h)
        return p0.value, p1.value

    # ---- trace
    def trace(self, cam, width, height, rgba=True, depth=True, gbuf=False, out=None):
        """material.frag main() per pixel; returns (rgba[h,w,4], depth[h,w], gbuf[h,w,16]); row 0 is
        the bottom row (GL window coordinates).  `out` may hold preallocated (pinned) arrays."""
        out = out or {}
        r = out.get("rgba", np.empty((height, width, 4), np.float32)) if rgba else None
        d = out.get("depth", np.empty((height, width), np.float32)) if depth else None
        g = out.get("gbuf", np.empty((height, width, GBUF_FLOATS), np.float32)) if gbuf else None
        check(self._lib.sdfgpu_trace(self._h, C.byref(cam), int(width), int(height), _host_ptr(r), _host_ptr(d),
                                     _host_ptr(g)), self._h)
        return r, d, g

    def trace_rgba8(self, cam, width, height, out_rgba8=None, out_depth=None):
        """The frame as an RGBA8 framebuffer holds it (+ depth): (rgba8[h,w,4] uint8, depth[h,w])."""
        r = out_rgba8 if out_rgba8 is not None else np.empty((height, width, 4), np.uint8)
        d = out_depth if out_depth is not None else np.empty((height, width), np.float32)
        check(self._lib.sdfgpu_trace_rgba8(self._h, C.byref(cam), int(width), int(height), _host_ptr(r), _host_ptr(d)),
              self._h)
        return r, d

    def trace_device(self, cam, width, height, want_gbuf=False):
        """Enqueue the trace; the frame stays in HBM.  Returns device pointers (rgba, depth, gbuf)."""
        r, d, g = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._lib.sdfgpu_trace_device(self._h, C.byref(cam), int(width), int(height), int(want_gbuf),
                                            C.byref(r), C.byref(d), C.byref(g)), self._h)
        return r.value, d.value, g.value

    def trace_params(self, cam, width, height, slab_clip=False):
        cmin, cmax = (C.c_float * 3)(), (C.c_float * 3)()
        lod, lin = C.c_float(), C.c_uint32()
        check(self._lib.sdfgpu_trace_params(self._h, C.byref(cam), int(width), int(height), int(slab_clip), cmin, cmax,
                                            C.byref(lod), C.byref(lin)), self._h)
        return list(cmin), list(cmax), lod.value, lin.value

    def trace_slab_keys(self, cam, width, height):
        k = C.c_void_p()
        check(self._lib.sdfgpu_trace_slab_keys(self._h, C.byref(cam), int(width), int(height), C.byref(k)), self._h)
        return k.value

    # ---- exact multi-GPU trace (replicated distance volume, hits shaded by their owner)
    def exact_trace_prepare(self):
        """(device pointer of the full-grid distance volume, first float of this handle's own range, count)."""
        p, first, count = C.c_void_p(), C.c_uint64(), C.c_uint64()
        check(self._lib.sdfgpu_exact_trace_prepare(self._h, C.byref(p), C.byref(first), C.byref(count)), self._h)
        return p.value, first.value, count.value

    def dist_volume_read(self, first, count):
        out = np.empty(int(count), np.float32)
        check(self._lib.sdfgpu_dist_volume_read(self._h, int(first), int(count), _host_ptr(out)), self._h)
        return out

    def dist_volume_write(self, first, values):
        a = np.ascontiguousarray(values, np.float32).reshape(-1)
        check(self._lib.sdfgpu_dist_volume_write(self._h, int(first), len(a), _host_ptr(a)), self._h)

    def trace_exact_keys(self, cam, width, height):
        k = C.c_void_p()
        check(self._lib.sdfgpu_trace_exact_keys(self._h, C.byref(cam), int(width), int(height), C.byref(k)), self._h)
        return k.value

    def keys_download(self, keys_dev, width, height, out_rgba8=None, out_depth=None):
        """Unpack a (depth, RGBA8) key frame into caller-provided (pinned) buffers, or fresh ones."""
        rgba8 = out_rgba8 if out_rgba8 is not None and out_rgba8.dtype == np.uint8 else np.empty((height, width, 4), np.uint8)
        depth = out_depth if out_depth is not None else np.empty((height, width), np.float32)
        check(self._lib.sdfgpu_keys_download(self._h, C.c_void_p(keys_dev), int(width), int(height), _host_ptr(rgba8),
                                             _host_ptr(depth)), self._h)
        return rgba8, depth

    # ---- fused halo exchange (multi-GPU, one process per GPU)
    def ipc_export(self):
        buf = C.create_string_buffer(128)
        check(self._lib.sdfgpu_ipc_export(self._h, buf, len(buf)), self._h)
        return buf.raw

    def ipc_attach(self, side, handles, peer_z_lo, peer_z_hi):
        buf = C.create_string_buffer(bytes(handles), len(handles))
        check(self._lib.sdfgpu_ipc_attach(self._h, int(side), buf, len(handles), int(peer_z_lo), int(peer_z_hi)), self._h)

    def ipc_detach(self):
        check(self._lib.sdfgpu_ipc_detach(self._h), self._h)

    # ---- SDFSurface::sample at arbitrary positions, mesh export (src/sdf/meshers)
    def sample_points(self, points):
        """`sample(p, false)` of the current tape for points (n, 3) -> (n, 7) raw SDFSamples."""
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.empty((len(pts), 7), np.float32)
        check(self._lib.sdfgpu_sample_points(self._h, _host_ptr(pts), len(pts), _host_ptr(out)), self._h)
        return out

    def mesh(self, download=True):
        """Marching cubes over the resident volume + Mesh::postproc through the tape (meshers/mod.rs:66-87).
        Returns (vertices (n, 12) float32, triangles (m, 3) uint32), or the two counts with download=False."""
        nv, nt = C.c_uint64(), C.c_uint64()
        check(self._lib.sdfgpu_mesh(self._h, C.byref(nv), C.byref(nt)), self._h)
        if not download:
            return nv.value, nt.value
        v = np.empty((nv.value, 12), np.float32)
        t = np.empty((nt.value, 3), np.uint32)
        check(self._lib.sdfgpu_mesh_download(self._h, _host_ptr(v), _host_ptr(t)), self._h)
        return v, t

    def mesh_write_ply(self, path, comment="Created with sdf-viewer_b200"):
        """Mesh::serialize_ply (meshers/mesh.rs:38-129) of the last mesh(); returns the bytes written."""
        n = C.c_uint64()
        check(self._lib.sdfgpu_mesh_write_ply(self._h, str(path).encode(), comment.encode() if comment else None, C.byref(n)), self._h)
        return n.value

    def cull_stats(self):
        """Per-tile culling of the current tape's UNION_RANGE over one fill of every voxel (sdfgpu_cull_stats)."""
        t, s, m, n = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        check(self._lib.sdfgpu_cull_stats(self._h, C.byref(t), C.byref(s), C.byref(m), C.byref(n)), self._h)
        return {"primitives": n.value, "tiles": t.value, "survivors_mean": s.value / t.value if t.value else 0.0,
                "survivors_max": m.value}

    # ---- linked slabs (multi-GPU behind the C ABI: include/sdfgpu.h "linked slabs")
    def link_export(self, rank, world, max_width, max_height, gbuf=False, halo_push=False, trace_mode=0):
        """This rank's link blob (CUDA IPC handles of its volumes and arena); gather the blobs of all ranks.
        `halo_push`: the neighbours push the halo slices (default: every rank fills its own); `trace_mode`: 0 auto,
        1 rounds, 2 stream (include/sdfgpu.h)."""
        buf = C.create_string_buffer(LINK_BLOB_BYTES)
        check(self._lib.sdfgpu_link_export(self._h, int(rank), int(world), int(max_width), int(max_height),
                                           link_flags(gbuf, halo_push, trace_mode), buf, len(buf)), self._h)
        return buf.raw

    def link_attach(self, blobs):
        """`blobs`: the link blobs of ranks 0 .. world-1 in rank order.  From here on fills and traces are collective."""
        raw = b"".join(bytes(b) for b in blobs)
        buf = C.create_string_buffer(raw, len(raw))
        check(self._lib.sdfgpu_link_attach(self._h, buf, len(raw) // LINK_BLOB_BYTES), self._h)

    def link_detach(self):
        check(self._lib.sdfgpu_link_detach(self._h), self._h)

    def trace_linked(self, cam, width, height, out_rgba8=None, out_depth=None, out_gbuf=None, gbuf=False, presenter=True):
        """The collective exact trace of a linked handle.  Rank 0 (`presenter`) receives the frame; the others pass
        nothing and get (None, None, None)."""
        r = d = g = None
        if presenter:
            r = out_rgba8 if out_rgba8 is not None else np.empty((height, width, 4), np.uint8)
            d = out_depth if out_depth is not None else np.empty((height, width), np.float32)
            if gbuf:
                g = out_gbuf if out_gbuf is not None else np.empty((height, width, GBUF_FLOATS), np.float32)
        check(self._lib.sdfgpu_trace_linked(self._h, C.byref(cam), int(width), int(height), int(bool(gbuf)), _host_ptr(r),
                                            _host_ptr(d), _host_ptr(g)), self._h)
        return r, d, g

    def trace_linked_device(self, cam, width, height, gbuf=False):
        """Enqueue the collective trace; device pointers (rgba8, depth, gbuf) on rank 0, None elsewhere."""
        r, d, g = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._lib.sdfgpu_trace_linked_device(self._h, C.byref(cam), int(width), int(height), int(bool(gbuf)),
                                                   C.byref(r), C.byref(d), C.byref(g)), self._h)
        return r.value, d.value, g.value

    # ---- stream
    def sync(self):
        check(self._lib.sdfgpu_sync(self._h), self._h)

    @property
    def stream(self):
        return self._lib.sdfgpu_stream(self._h)

    @property
    def launch_count(self):
        return self._lib.sdfgpu_launch_count(self._h)

    def get_info(self, key):
        v = C.c_int64()
        check(self._lib.sdfgpu_get_info(self._h, key.encode(), C.byref(v)), self._h)
        return v.value

    def set_option(self, key, value):
        check(self._lib.sdfgpu_set_option(self._h, key.encode(), int(value)), self._h)


class _RankView(SDFViewer):
    """A group's handle of one rank: owned by the group (never destroyed from here)."""

    def close(self):
        self._h = None


def link_flags(gbuf=False, halo_push=False, trace_mode=0):
    return ((LINK_GBUF if gbuf else 0) | (LINK_HALO_PUSH if halo_push else 0)
            | {0: 0, 1: LINK_ROUNDS, 2: LINK_STREAM}[int(trace_mode)])


class SDFViewerGroup:
    """`SDFViewer` over several GPUs driven by ONE process and one thread, as the reference's scene is
    (scene/mod.rs:22-31,158-225): the grid is sharded along z over `devices`, every call below runs on all of
    them, and `trace*` returns the frame of one SDFViewer holding the whole grid (include/sdfgpu.h, sdfgpu_group_*)."""

    def __init__(self, handle, bounding_box):
        self._lib = _lib.load()
        self._g = handle
        self.bounding_box = bounding_box
        self.size = self._lib.sdfgpu_group_size(self._g)
        self.ranks = [_RankView(C.c_void_p(self._lib.sdfgpu_group_rank(self._g, r)), bounding_box) for r in range(self.size)]
        self.dims = self.ranks[0].dims
        self._tape = None

    @classmethod
    def new_voxels(cls, voxels, bb, loading_passes, devices, max_width=1920, max_height=1080, gbuf=False, halo_push=False,
                   trace_mode=0):
        lib = _lib.load()
        g = C.c_void_p()
        v = (C.c_uint32 * 3)(*[int(x) for x in voxels])
        dv = (C.c_int * len(devices))(*[int(d) for d in devices])
        check_group(lib.sdfgpu_group_create(_f6(bb), v, int(loading_passes), dv, len(devices), int(max_width), int(max_height),
                                            link_flags(gbuf, halo_push, trace_mode), C.byref(g)))
        return cls(g, bb)

    @classmethod
    def from_bb(cls, bb, max_voxels_side, loading_passes, device_mask=1, max_width=1920, max_height=1080):
        lib = _lib.load()
        g = C.c_void_p()
        check_group(lib.sdfgpu_group_create_mask(_f6(bb), int(max_voxels_side), int(loading_passes), int(device_mask),
                                                 int(max_width), int(max_height), C.byref(g)))
        return cls(g, bb)

    def close(self):
        if self._g:
            for r in self.ranks:
                r._h = None
            self._lib.sdfgpu_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        check_group(rc, self._g)

    def set_tape(self, tape_bytes):
        buf = (C.c_char * len(tape_bytes)).from_buffer_copy(tape_bytes)
        self._ck(self._lib.sdfgpu_group_set_tape(self._g, buf, len(tape_bytes)))
        self._tape = bytes(tape_bytes)

    def update(self, sdf, max_passes=0):
        changed = None
        if sdf is not None:
            changed = sdf.changed()
            tape = bytes(sdf.tape())
            if self._tape is None or tape != self._tape:
                self.set_tape(tape)
        it = C.c_uint64()
        box = _f6(changed) if changed is not None else None
        self._ck(self._lib.sdfgpu_group_update(self._g, box, int(max_passes), C.byref(it)))
        return it.value

    def update_surface(self, sdf, max_delta_time=0.030):
        surf = _surface_struct(sdf)
        it = C.c_uint64()
        self._ck(self._lib.sdfgpu_group_update_surface(self._g, C.byref(surf), float(max_delta_time), C.byref(it)))
        err = getattr(surf, "_py_error", None)
        if err and err[0] is not None:
            raise err[0]
        return it.value

    def fill_all(self):
        self._ck(self._lib.sdfgpu_group_fill_all(self._g))

    def resample_box(self, box, count=False):
        n = C.c_uint64()
        self._ck(self._lib.sdfgpu_group_resample_box(self._g, _f6(box), C.byref(n) if count else None))
        return n.value if count else None

    def commit(self):
        self._ck(self._lib.sdfgpu_group_commit(self._g))

    def reset(self, loading_passes):
        self._ck(self._lib.sdfgpu_group_reset(self._g, int(loading_passes)))
        self._tape = None

    def set_option(self, key, value):
        self._ck(self._lib.sdfgpu_group_set_option(self._g, key.encode(), int(value)))

    def sync(self):
        self._ck(self._lib.sdfgpu_group_sync(self._g))

    def loading_state(self):
        ln, tot = C.c_uint64(), C.c_uint64()
        left, passes = C.c_uint32(), C.c_uint32()
        self._ck(self._lib.sdfgpu_group_loading_state(self._g, C.byref(ln), C.byref(tot), C.byref(left), C.byref(passes)))
        return ln.value, tot.value, left.value, passes.value

    def download(self, out0=None, out1=None):
        """The whole grid's volumes, shaped (depth, height, width, 4)."""
        shape = (self.dims[2], self.dims[1], self.dims[0], 4)
        a0 = out0 if out0 is not None else np.empty(shape, np.float32)
        a1 = out1 if out1 is not None else np.empty(shape, np.float32)
        self._ck(self._lib.sdfgpu_group_download(self._g, _host_ptr(a0), _host_ptr(a1)))
        return a0, a1

    def trace_rgba8(self, cam, width, height, out_rgba8=None, out_depth=None):
        r = out_rgba8 if out_rgba8 is not None else np.empty((height, width, 4), np.uint8)
        d = out_depth if out_depth is not None else np.empty((height, width), np.float32)
        self._ck(self._lib.sdfgpu_group_trace_rgba8(self._g, C.byref(cam), int(width), int(height), _host_ptr(r), _host_ptr(d)))
        return r, d

    def trace(self, cam, width, height, gbuf=True):
        """(rgba8, depth, gbuf); the G-buffer needs a group created with gbuf=True."""
        r = np.empty((height, width, 4), np.uint8)
        d = np.empty((height, width), np.float32)
        g = np.empty((height, width, GBUF_FLOATS), np.float32) if gbuf else None
        self._ck(self._lib.sdfgpu_group_trace(self._g, C.byref(cam), int(width), int(height), _host_ptr(r), _host_ptr(d), _host_ptr(g)))
        return r, d, g

    @property
    def launch_count(self):
        return sum(r.launch_count for r in self.ranks)


def ply_serialize(vertices, triangles, path, comment="Created with sdf-viewer_b200"):
    """Mesh::serialize_ply for host arrays: vertices (n, 12) float32, triangles (m, 3) uint32."""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 12)
    t = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    n = C.c_uint64()
    check(_lib.load().sdfgpu_ply_serialize(_host_ptr(v), len(v), _host_ptr(t), len(t), comment.encode() if comment else None,
                                           str(path).encode(), C.byref(n)))
    return n.value


def tape_validate(tape_bytes):
    """Device-free validation of a tape; raises SdfGpuError with the reason."""
    buf = (C.c_char * len(tape_bytes)).from_buffer_copy(tape_bytes) if len(tape_bytes) else None
    rc = _lib.load().sdfgpu_tape_validate(buf, len(tape_bytes))
    if rc != 0:
        raise SdfGpuError(rc, _lib.load().sdfgpu_last_error(None).decode("utf-8", "replace"))


def jit_check(tape_bytes, voxels_per_thread=2):
    """Compile the specialised fill kernel for this tape's structure with NVRTC (no GPU needed).
    Returns the generated translation unit; raises SdfGpuError with the compiler log on failure."""
    buf = (C.c_char * len(tape_bytes)).from_buffer_copy(tape_bytes)
    log = C.create_string_buffer(1 << 22)  # the generated translation unit: long for scalar programs
    rc = _lib.load().sdfgpu_jit_check(buf, len(tape_bytes), int(voxels_per_thread), log, len(log))
    if rc != 0:
        raise SdfGpuError(rc, log.value.decode("utf-8", "replace") or _lib.load().sdfgpu_last_error(None).decode())
    return log.value.decode()

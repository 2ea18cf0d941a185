"""ctypes binding of libsdfgpu.so (include/sdfgpu.h).

This is the binding a host in another language would write (cgo / JNI / Rust
`extern "C"`): plain pointers and sizes, no torch types.  The library is
loaded from this directory; if it is missing the import of the compute API
fails loudly -- there is no CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdfgpu.so")

SDFGPU_OK = 0
SDFGPU_ERR_INVALID = -1
SDFGPU_ERR_CUDA = -2
SDFGPU_ERR_TAPE = -3
SDFGPU_ERR_STATE = -4
GBUF_FLOATS = 16
LINK_BLOB_BYTES = 320
LINK_GBUF = 1
LINK_HALO_PUSH = 2
LINK_ROUNDS = 4
LINK_STREAM = 8


class SdfGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"sdfgpu error {code}: {message}")
        self.code = code
        self.message = message


class Camera(C.Structure):  # sdfgpu_camera
    _fields_ = [
        ("position", C.c_float * 3),
        ("view", C.c_float * 16),
        ("projection", C.c_float * 16),
        ("tint", C.c_float * 4),
        ("tone_mapping", C.c_uint32),
        ("color_mapping", C.c_uint32),
        ("gamma", C.c_float),
        ("ambient", C.c_float * 3),
    ]


class Rays(C.Structure):  # sdfgpu_rays
    _fields_ = [
        ("origin", C.c_float * 3),
        ("base", C.c_float * 3),
        ("dx", C.c_float * 3),
        ("dy", C.c_float * 3),
        ("bvp", C.c_float * 16),
    ]


BOUNDING_BOX_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float))
SAMPLE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float))
SAMPLE_BATCH_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_uint64, C.c_int, C.POINTER(C.c_float))
CHANGED_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float))
TAPE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))


class Surface(C.Structure):  # sdfgpu_surface
    _fields_ = [
        ("self", C.c_void_p),
        ("bounding_box", BOUNDING_BOX_FN),
        ("sample", SAMPLE_FN),
        ("sample_batch", SAMPLE_BATCH_FN),
        ("changed", CHANGED_FN),
        ("tape", TAPE_FN),
        ("sample_threads", C.c_uint32),
    ]


_vp = C.c_void_p
_u32 = C.c_uint32
_u64 = C.c_uint64
_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every function include/sdfgpu.h declares
SIGNATURES = {
    "sdfgpu_air_dist": (C.c_float, []),
    "sdfgpu_dims_from_bb": (C.c_int, [_fp, _u32, _u32p]),
    "sdfgpu_create": (C.c_int, [_fp, _u32, _u32, C.c_int, _vpp]),
    "sdfgpu_create_voxels": (C.c_int, [_fp, _u32p, _u32, C.c_int, _vpp]),
    "sdfgpu_create_slab": (C.c_int, [_fp, _u32p, _u32, C.c_int, _u32, _u32, _vpp]),
    "sdfgpu_ipc_export": (C.c_int, [_vp, _vp, C.c_size_t]),
    "sdfgpu_ipc_attach": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, _u32, _u32]),
    "sdfgpu_ipc_detach": (C.c_int, [_vp]),
    "sdfgpu_link_export": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, _vp, C.c_size_t]),
    "sdfgpu_link_attach": (C.c_int, [_vp, _vp, _u32]),
    "sdfgpu_link_detach": (C.c_int, [_vp]),
    "sdfgpu_trace_linked": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, C.c_int, _vp, _vp, _vp]),
    "sdfgpu_trace_linked_device": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, C.c_int, _vpp, _vpp, _vpp]),
    "sdfgpu_group_create": (C.c_int, [_fp, _u32p, _u32, C.POINTER(C.c_int), _u32, _u32, _u32, _u32, _vpp]),
    "sdfgpu_group_create_mask": (C.c_int, [_fp, _u32, _u32, _u32, _u32, _u32, _vpp]),
    "sdfgpu_group_destroy": (None, [_vp]),
    "sdfgpu_group_size": (_u32, [_vp]),
    "sdfgpu_group_rank": (_vp, [_vp, _u32]),
    "sdfgpu_group_last_error": (C.c_char_p, [_vp]),
    "sdfgpu_group_set_tape": (C.c_int, [_vp, _vp, C.c_size_t]),
    "sdfgpu_group_update": (C.c_int, [_vp, _fp, _u32, _u64p]),
    "sdfgpu_group_update_surface": (C.c_int, [_vp, C.POINTER(Surface), C.c_double, _u64p]),
    "sdfgpu_group_fill_all": (C.c_int, [_vp]),
    "sdfgpu_group_resample_box": (C.c_int, [_vp, _fp, _u64p]),
    "sdfgpu_group_commit": (C.c_int, [_vp]),
    "sdfgpu_group_reset": (C.c_int, [_vp, _u32]),
    "sdfgpu_group_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "sdfgpu_group_sync": (C.c_int, [_vp]),
    "sdfgpu_group_loading_state": (C.c_int, [_vp, _u64p, _u64p, _u32p, _u32p]),
    "sdfgpu_group_download": (C.c_int, [_vp, _vp, _vp]),
    "sdfgpu_group_trace_rgba8": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, _vp, _vp]),
    "sdfgpu_group_trace": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, _vp, _vp, _vp]),
    "sdfgpu_sample_points": (C.c_int, [_vp, _vp, _u64, _vp]),
    "sdfgpu_mesh": (C.c_int, [_vp, _u64p, _u64p]),
    "sdfgpu_mesh_download": (C.c_int, [_vp, _vp, _vp]),
    "sdfgpu_mesh_device_ptrs": (C.c_int, [_vp, _vpp, _vpp]),
    "sdfgpu_mesh_write_ply": (C.c_int, [_vp, C.c_char_p, C.c_char_p, _u64p]),
    "sdfgpu_ply_serialize": (C.c_int, [_vp, _u64, _vp, _u64, C.c_char_p, C.c_char_p, _u64p]),
    "sdfgpu_destroy": (None, [_vp]),
    "sdfgpu_last_error": (C.c_char_p, [_vp]),
    "sdfgpu_dims": (C.c_int, [_vp, _u32p]),
    "sdfgpu_slab": (C.c_int, [_vp, _u32p, _u32p, _u32p, _u32p]),
    "sdfgpu_set_tape": (C.c_int, [_vp, _vp, C.c_size_t]),
    "sdfgpu_tape_validate": (C.c_int, [_vp, C.c_size_t]),
    "sdfgpu_jit_check": (C.c_int, [_vp, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]),
    "sdfgpu_wasm_lower": (C.c_int, [_vp, C.c_size_t, _u32, _vp, C.c_size_t, C.POINTER(C.c_size_t), _fp, C.c_char_p, C.c_size_t]),
    "sdfgpu_wasm_lower_live": (C.c_int, [_vp, C.c_size_t, _vp, C.c_size_t, _u32, _vp, C.c_size_t, C.POINTER(C.c_size_t), _fp,
                                         C.c_char_p, C.c_size_t]),
    "sdfgpu_update": (C.c_int, [_vp, _fp, _u32, _u64p]),
    "sdfgpu_update_surface": (C.c_int, [_vp, C.POINTER(Surface), C.c_double, _u64p]),
    "sdfgpu_fill_all": (C.c_int, [_vp]),
    "sdfgpu_resample_box": (C.c_int, [_vp, _fp, _u64p]),
    "sdfgpu_cull_stats": (C.c_int, [_vp, _u64p, _u64p, _u64p, _u32p]),
    "sdfgpu_voxel_positions": (C.c_int, [_vp, _u64, _u64, _vp]),
    "sdfgpu_ingest_samples": (C.c_int, [_vp, _u64, _u64, _vp]),
    "sdfgpu_commit": (C.c_int, [_vp]),
    "sdfgpu_loading_state": (C.c_int, [_vp, _u64p, _u64p, _u32p, _u32p]),
    "sdfgpu_loading_create": (C.c_int, [_u32p, _u32, _vpp]),
    "sdfgpu_loading_destroy": (None, [_vp]),
    "sdfgpu_loading_reset": (None, [_vp, _u32]),
    "sdfgpu_loading_next": (C.c_int, [_vp, _u32p]),
    "sdfgpu_loading_next_run": (_u64, [_vp, _u64, _u32p, _u32p]),
    "sdfgpu_loading_len": (_u64, [_vp]),
    "sdfgpu_loading_total_iterations": (_u64, [_vp]),
    "sdfgpu_loading_passes_left": (_u32, [_vp]),
    "sdfgpu_reset": (C.c_int, [_vp, _u32]),
    "sdfgpu_download": (C.c_int, [_vp, _vp, _vp]),
    "sdfgpu_device_ptrs": (C.c_int, [_vp, _vpp, _vpp]),
    "sdfgpu_camera_default": (None, [C.POINTER(Camera), _u32, _u32]),
    "sdfgpu_look_at_rh": (None, [_fp, _fp, _fp, _fp]),
    "sdfgpu_perspective": (None, [C.c_float, C.c_float, C.c_float, C.c_float, _fp]),
    "sdfgpu_camera_rays": (C.c_int, [C.POINTER(Camera), _u32, _u32, C.POINTER(Rays)]),
    "sdfgpu_trace": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, _vp, _vp, _vp]),
    "sdfgpu_trace_rgba8": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, _vp, _vp]),
    "sdfgpu_trace_device": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, C.c_int, _vpp, _vpp, _vpp]),
    "sdfgpu_trace_params": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, C.c_int, _fp, _fp, _fp, _u32p]),
    "sdfgpu_trace_slab_keys": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, _vpp]),
    "sdfgpu_gl_register": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32]),
    "sdfgpu_gl_unregister": (C.c_int, [_vp]),
    "sdfgpu_trace_gl": (C.c_int, [_vp, C.POINTER(Camera)]),
    "sdfgpu_exact_trace_prepare": (C.c_int, [_vp, _vpp, _u64p, _u64p]),
    "sdfgpu_dist_volume_read": (C.c_int, [_vp, _u64, _u64, _vp]),
    "sdfgpu_dist_volume_write": (C.c_int, [_vp, _u64, _u64, _vp]),
    "sdfgpu_trace_exact_keys": (C.c_int, [_vp, C.POINTER(Camera), _u32, _u32, _vpp]),
    "sdfgpu_keys_download": (C.c_int, [_vp, _vp, _u32, _u32, _vp, _vp]),
    "sdfgpu_sync": (C.c_int, [_vp]),
    "sdfgpu_stream": (_vp, [_vp]),
    "sdfgpu_launch_count": (_u64, [_vp]),
    "sdfgpu_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "sdfgpu_get_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """Load libsdfgpu.so and attach the signatures.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  sdf-viewer_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != SDFGPU_OK:
        msg = load().sdfgpu_last_error(ctx)
        raise SdfGpuError(rc, msg.decode("utf-8", "replace") if msg else "")


def check_group(rc, group=None):
    if rc != SDFGPU_OK:
        msg = load().sdfgpu_group_last_error(group)
        raise SdfGpuError(rc, msg.decode("utf-8", "replace") if msg else "")

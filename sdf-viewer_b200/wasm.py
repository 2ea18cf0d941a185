"""WebAssembly SDFs on the GPU: `lower()` turns the `sample` export of a guest module (the ABI of
/root/reference/src/sdf/wasm/mod.rs:5-37, loaded by the reference through src/sdf/wasm/native.rs) into a tape
with `sdfgpu_wasm_lower`; `WasmSDF` is the `SDFSurface` over it.  The lowering itself is C++ in libsdfgpu.so
(csrc/wasm_lower.cu): nothing here interprets WebAssembly."""
import ctypes as C

from . import _lib
from .sdf import SDFSurface


class WasmLoweringError(_lib.SdfGpuError):
    """The module cannot be expressed as a tape (message says why); sample it on the host instead."""


def lower(wasm_bytes, sdf_id=0, memory=None):
    """(tape_bytes, bounding_box, summary) of the guest's SDF `sdf_id` (0 = root, src/sdf/wasm/mod.rs:5-37).
    `memory`: the linear memory of a live instance (bytes, whole pages) to lower the guest as it is NOW --
    after init() and any set_parameter calls -- instead of instantiating the module afresh."""
    lib = _lib.load()
    buf = (C.c_char * len(wasm_bytes)).from_buffer_copy(bytes(wasm_bytes))
    need = C.c_size_t()
    bb = (C.c_float * 6)()
    log = C.create_string_buffer(1024)
    if memory is None:
        def call(out, cap):
            return lib.sdfgpu_wasm_lower(buf, len(wasm_bytes), int(sdf_id), out, cap, C.byref(need), bb, log, len(log))
    else:
        mem = (C.c_char * len(memory)).from_buffer_copy(bytes(memory)) if len(memory) else None

        def call(out, cap):
            return lib.sdfgpu_wasm_lower_live(buf, len(wasm_bytes), mem, len(memory), int(sdf_id), out, cap, C.byref(need), bb,
                                              log, len(log))
    rc = call(None, 0)
    if rc != _lib.SDFGPU_OK:
        raise WasmLoweringError(rc, log.value.decode("utf-8", "replace"))
    tape = C.create_string_buffer(need.value)
    rc = call(tape, need.value)
    if rc != _lib.SDFGPU_OK:
        raise WasmLoweringError(rc, log.value.decode("utf-8", "replace"))
    b = list(bb)
    return tape.raw[:need.value], (tuple(b[:3]), tuple(b[3:])), log.value.decode()


class WasmSDF(SDFSurface):
    """An existing .wasm SDF as a surface with a tape: evaluated on the GPU by SDFViewer.update / update_surface."""

    def __init__(self, wasm_bytes, sdf_id=0, memory=None):
        self._wasm, self._id = bytes(wasm_bytes), int(sdf_id)
        self._tape, self._bb, self.summary = lower(self._wasm, self._id, memory)
        self._changed = False

    def relower(self, memory):
        """The host changed a parameter of the live guest (set_parameter): lower it again from its memory.  The
        next update() re-samples the whole bounding box, as the demo's own changed() asks for."""
        self._tape, self._bb, self.summary = lower(self._wasm, self._id, memory)
        self._changed = True

    def changed(self):
        if self._changed:
            self._changed = False
            return self._bb
        return None

    def bounding_box(self):
        return self._bb

    def tape(self):
        return self._tape

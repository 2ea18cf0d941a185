"""Host mirror of the `SDFSurface` contract (/root/reference/src/sdf/mod.rs:33-101) for the
hot path: `bounding_box`, `changed`, and -- new, optional -- `tape()`, the GPU-evaluable
lowering of `sample`.  The reference can only call `sample(p)` point by point through the WASM
sandbox (src/sdf/wasm/native.rs:188-217); an SDF that provides a tape is evaluated on the GPU."""
from . import tape as _tape


class SDFSurface:
    def bounding_box(self):  # src/sdf/mod.rs:37
        raise NotImplementedError

    def sample(self, points, distance_only=False):  # src/sdf/mod.rs:43, batched
        """`SDFSample`s (n, 7) float32 -- distance, r, g, b, metallic, roughness, occlusion
        (src/sdf/mod.rs:104-118) -- for points (n, 3).  Only surfaces without a tape need it."""
        raise NotImplementedError

    def changed(self):  # src/sdf/mod.rs:87 -> Option<[Vector3; 2]>
        return None

    def tape(self):
        """Tape bytes (include/sdfgpu_tape.h) equivalent to `sample(p, false)`; a surface that
        cannot be lowered leaves this unimplemented and is sampled on the host."""
        raise NotImplementedError


class SDFDemo(SDFSurface):
    """`SDFDemo` (src/sdf/demo/mod.rs): L-inf cube (brick) minus sphere (normal colours) with a
    seam material.  Parameter ids follow the demo's `parameters()` only loosely -- names are used."""

    def __init__(self, cube_half_side=0.95, sphere_radius=1.05, max_distance_custom_material=0.05,
                 disable_sphere=False, cube_material=_tape.MAT_BRICK, sphere_material=_tape.MAT_NORMAL):
        self.params = dict(cube_half_side=cube_half_side, sphere_radius=sphere_radius,
                           max_distance_custom_material=max_distance_custom_material,
                           disable_sphere=disable_sphere, cube_material=cube_material,
                           sphere_material=sphere_material)
        self._changed = False

    def bounding_box(self):  # demo/mod.rs:47-49
        return ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))

    def set_parameter(self, name, value):  # demo/mod.rs:117-132: any edit dirties the whole SDF
        if name not in self.params:
            raise KeyError(name)
        self.params[name] = value
        self._changed = True

    def changed(self):  # demo/mod.rs:136-145: reports the whole bounding box, once
        if self._changed:
            self._changed = False
            return self.bounding_box()
        return None

    def tape(self):
        return _tape.demo_tape(**self.params)


class TapeSDF(SDFSurface):
    """Any SDF given directly as tape bytes plus its bounding box."""

    def __init__(self, tape_bytes, bounding_box=((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))):
        self._tape = bytes(tape_bytes)
        self._bb = bounding_box

    def bounding_box(self):
        return self._bb

    def tape(self):
        return self._tape

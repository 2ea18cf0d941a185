"""sdf-viewer_b200: the B200-native grid-fill + sphere-trace hot path of Yeicor/sdf-viewer.

Layout: `csrc/` holds the sm_100a kernels and the C ABI (`include/sdfgpu.h`); this package is
the host-side mirror of the reference's `SDFViewer` / `SDFSurface` interface over that ABI.
Import as `sdf_viewer_b200` (shim at the repo root).  Nothing here imports `oracle/`.
"""
from . import tape, loading, sdf, wasm  # noqa: F401
from .sdf import SDFSurface, SDFDemo, TapeSDF  # noqa: F401
from .loading import LoadingManager, NativeLoadingManager  # noqa: F401
from .wasm import WasmSDF, WasmLoweringError  # noqa: F401
from .viewer import (  # noqa: F401
    SDFViewer, SDFViewerGroup, Camera, Rays, SdfGpuError, GBUF_FLOATS, dims_from_bb, default_camera, look_at_camera, camera_rays,
    jit_check, tape_validate, ply_serialize,
)

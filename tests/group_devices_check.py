"""Run as ONE process on a box with >= 2 GPUs (tests/test_sharded_gpu.py launches it): the single-process form of the
linked slabs -- sdfgpu_group_* with every rank on a device of its own, all driven by one host thread, which is how the
reference's scene runs (src/app/scene/mod.rs:22-31, 158-225).  The group must reproduce ONE handle holding the whole
grid: volumes bit for bit after every loading pass, frames bit for bit (RGBA8, depth, G-buffer) for cameras outside
and inside the box, dirty boxes, reset + fill_all -- with the trace as one streaming kernel per device (and in rounds),
halo slices filled by their holder (and pushed by the neighbours)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import sdf_viewer_b200 as S  # noqa: E402
from test_linked_gpu import BB, group_body  # noqa: E402


def main():
    n_dev = torch.cuda.device_count()
    assert n_dev >= 2, "needs at least 2 GPUs"
    n = min(n_dev, 4)
    w, h = 200, 150
    for dims, halo_push, trace_mode in (((48, 40, 36), False, 0), ((40, 36, 50), True, 0), ((36, 32, 33), False, 1),
                                        ((33, 31, 64), True, 2)):
        with S.SDFViewer.new_voxels(dims, BB, 3) as whole, \
                S.SDFViewerGroup.new_voxels(dims, BB, 3, list(range(n)), w, h, gbuf=True, halo_push=halo_push,
                                            trace_mode=trace_mode) as g:
            assert g.size == n
            stream = trace_mode != 1
            assert all(r.get_info("link_trace_stream") == int(stream) and r.get_info("link_halo_push") == int(halo_push) for r in g.ranks)
            group_body(S, whole, g, S.SDFDemo(), dims, w, h)
        print(f"group over {n} devices ok: dims {dims}, halo {'pushed' if halo_push else 'filled locally'}, trace "
              f"{'streamed' if stream else 'in rounds'}")
    print("group_devices_check ok")


if __name__ == "__main__":
    main()

"""Z-slab sharding of the grid across GPUs, one process per GPU.

The fill is a pure map over voxels (`sample` depends only on the position,
/root/reference/src/sdf/mod.rs:43), and the flat index is z-major
(src/app/scene/sdf/mod.rs:177), so rank g owns the contiguous slices
[g*D/G, (g+1)*D/G) and no collective is needed to fill them.  The one exchange step is the
halo: the trilinear taps of the tracer at a slab face read one slice of the neighbour.

Default (`linked=True`): the slab handles are LINKED through the C ABI (include/sdfgpu.h "linked slabs",
csrc/link.cu).  `torch.distributed` is used ONCE, at set-up, to gather the 320-byte link blobs (CUDA IPC handles);
after that every call below is a plain library call -- the fill is one launch whose boundary tiles go first and are
pushed into the neighbours' halo slices by the copy engines behind a flag, the trace hands rays from rank to rank
(exact: the frame is the single-volume frame bit for bit) and stores finished pixels straight into rank 0's frame
over NVLink.  No collective-library call and no host synchronisation between ranks in the frame loop.

Fallback (`linked=False`, or when a rank cannot map its peers): the round-1 path -- halo exchange by CUDA IPC pushes
ordered with a 4-byte NCCL all-reduce, or NCCL send/recv; sort-last trace composited with all-reduce(MIN) over 64-bit
(depth, RGBA8) keys (approximate at slab faces: every rank marches its own sub-box).

`torch` is used for the process group, streams and as a view on the library's device memory.
"""
import numpy as np

from .viewer import SDFViewer, SdfGpuError


def slab_range(depth, rank, world):
    """Owned z slices of `rank`: contiguous, balanced, covering [0, depth)."""
    return (rank * depth) // world, ((rank + 1) * depth) // world


def stored_range(depth, z_begin, z_end):
    """Stored slices = owned slices plus one halo slice on each interior face (include/sdfgpu.h)."""
    if z_begin == z_end:
        return z_begin, z_end
    return max(z_begin - 1, 0), min(z_end + 1, depth)


def halo_plan(depth, rank, world):
    """The send/recv list of one rank: (kind, peer, z) with kind in {"send", "recv"}; z is the global
    slice index moved.  Ranks with empty slabs take no part; neighbours are the nearest non-empty ranks."""
    zb, ze = slab_range(depth, rank, world)
    if zb == ze:
        return []
    ops = []
    lo = rank - 1
    while lo >= 0 and slab_range(depth, lo, world)[0] == slab_range(depth, lo, world)[1]:
        lo -= 1
    hi = rank + 1
    while hi < world and slab_range(depth, hi, world)[0] == slab_range(depth, hi, world)[1]:
        hi += 1
    if lo >= 0:
        ops += [("send", lo, zb), ("recv", lo, zb - 1)]
    if hi < world:
        ops += [("send", hi, ze - 1), ("recv", hi, ze)]
    return ops


def exchange_halos(dist, textures, dims, rank, world, group=None):
    """Exchange the boundary slices of every tensor in `textures` (each a flat float32 tensor holding
    the stored slices [z_lo, z_hi) of a W*H*4-float-per-slice volume).  Device agnostic (NCCL on GPU,
    gloo in the CPU tests)."""
    W, H, D = dims
    zb, ze = slab_range(D, rank, world)
    z_lo, _ = stored_range(D, zb, ze)
    n = W * H * 4
    ops = []
    for kind, peer, z in halo_plan(D, rank, world):
        for t in textures:
            view = t[(z - z_lo) * n:(z - z_lo + 1) * n]
            ops.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, view, peer, group))
    if not ops:
        return 0
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return len(ops)


def gather_slabs(dist, full, dims, world, floats_per_voxel=1, group=None):
    """All-gather of a z-sharded, z-major volume in place: `full` is a flat tensor covering the WHOLE grid
    (W*H*D*floats_per_voxel) in which every rank has filled its own slices; afterwards every rank holds all
    of them.  One broadcast per non-empty slab, so slabs may be uneven (D not divisible by the world size).
    Device agnostic (NCCL on GPU, gloo in the CPU tests).  Returns the number of broadcasts."""
    W, H, D = dims
    n = W * H * floats_per_voxel
    ops = 0
    for r in range(world):
        zb, ze = slab_range(D, r, world)
        if ze > zb:
            dist.broadcast(full[zb * n:ze * n], src=r, group=group)
            ops += 1
    return ops


class _DevMem:
    """A __cuda_array_interface__ view of library-owned device memory."""

    def __init__(self, ptr, n_items, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n_items),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class ShardedViewer:
    """One rank's part of a Z-sharded SDFViewer.  With world == 1 it is a plain SDFViewer."""

    def __init__(self, dims, bb, loading_passes, rank=0, world=1, device=0, group=None, fused=True, linked=True,
                 max_width=1920, max_height=1080, gbuf=False, halo_push=False, trace_mode=0):
        self.dims, self.bb, self.rank, self.world, self.device = tuple(dims), bb, rank, world, device
        self.halo_push, self.trace_mode = halo_push, trace_mode
        self.dist = group  # the torch.distributed module (None when world == 1)
        self.fused = False
        self.linked = False
        self._synced = True  # no rank can still be reading a halo slice (nothing has been traced yet)
        if world > 1:
            zr = slab_range(dims[2], rank, world)
            self.viewer = SDFViewer.new_voxels(dims, bb, loading_passes, device=device, z_range=zr)
            self.viewer.set_option("fill_halo", 0)
            import torch
            self._torch = torch
            self._stream = torch.cuda.ExternalStream(self.viewer.stream, device=torch.device("cuda", device))
            p0, p1 = self.viewer.device_ptrs()
            n = dims[0] * dims[1] * (self.viewer.z_hi - self.viewer.z_lo) * 4
            self._tex = [torch.as_tensor(_DevMem(p, n, "<f4"), device=torch.device("cuda", device)) for p in (p0, p1)]
            self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device))
            if linked:
                self._link(max_width, max_height, gbuf)
            if fused and not self.linked:
                self._attach_neighbours()
        else:
            self.viewer = SDFViewer.new_voxels(dims, bb, loading_passes, device=device)

    def close(self):
        if self.linked:  # every rank stops using its peers' memory before any of it is unmapped or freed
            self.viewer.sync()
            self.dist.barrier()
            self.viewer.link_detach()
            self.dist.barrier()
            self.linked = False
        self.viewer.close()

    def _link(self, max_width, max_height, gbuf):
        """Set-up only: gather every rank's link blob and attach.  All ranks link, or none does."""
        v, dist, t = self.viewer, self.dist, self._torch
        ok, blob = 1, None
        try:
            blob = v.link_export(self.rank, self.world, max_width, max_height, gbuf=gbuf, halo_push=self.halo_push,
                                 trace_mode=self.trace_mode)
        except SdfGpuError:
            ok = 0
        blobs = [None] * self.world
        dist.all_gather_object(blobs, blob)
        if ok and all(b is not None for b in blobs):
            try:
                v.link_attach(blobs)
            except SdfGpuError as e:
                self.link_error = str(e)
                ok = 0
        else:
            ok = 0
        flag = t.tensor([ok], dtype=t.int32, device=t.device("cuda", self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.linked = bool(flag.item())
        if not self.linked:
            v.link_detach()
            v.set_option("fill_halo", 0)

    def _attach_neighbours(self):
        """Exchange CUDA IPC handles of the volumes and map the two neighbours' (fused halo exchange).
        Falls back to NCCL send/recv on every rank if any rank cannot map its neighbours."""
        v, dist, t = self.viewer, self.dist, self._torch
        ok = 1
        try:
            mine = (v.ipc_export() if v.z_end > v.z_begin else None, v.z_lo, v.z_hi)
        except SdfGpuError:
            mine, ok = (None, v.z_lo, v.z_hi), 0
        infos = [None] * self.world
        dist.all_gather_object(infos, mine)
        if ok:
            try:
                for kind, peer, z in halo_plan(self.dims[2], self.rank, self.world):
                    if kind != "send":
                        continue
                    handles, plo, phi = infos[peer]
                    if handles is None:
                        raise SdfGpuError(-2, "neighbour exported no handles")
                    v.ipc_attach(0 if peer < self.rank else 1, handles, plo, phi)
            except SdfGpuError:
                ok = 0
        flag = t.tensor([ok], dtype=t.int32, device=t.device("cuda", self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.fused = bool(flag.item())
        if not self.fused:
            v.ipc_detach()

    def _barrier(self):
        """Orders the ranks on the library's stream without blocking the host: a 4-byte all-reduce."""
        with self._torch.cuda.stream(self._stream):
            self.dist.all_reduce(self._flag)

    def _before_fill(self):
        if self.linked:
            return  # ordered by the library's flags
        # a neighbour may still be reading the halo slices this fill overwrites -- unless the last thing
        # every rank did was a collective that came after its reads (the compositing all-reduce)
        if self.world > 1 and self.fused and not self._synced:
            self._barrier()

    def _after_fill(self):
        if self.world == 1 or self.linked:
            return
        if self.fused:
            self._barrier()  # every rank's fill, and with it every halo slice, is complete
            self._synced = True
        else:
            with self._torch.cuda.stream(self._stream):  # NCCL orders itself after / before this stream
                exchange_halos(self.dist, self._tex, self.dims, self.rank, self.world)

    def fill_all(self):
        self._before_fill()
        self.viewer.fill_all()
        self._after_fill()

    def update(self, sdf, max_passes=0):
        self._before_fill()
        it = self.viewer.update(sdf, max_passes)
        self._after_fill()
        return it

    def commit(self):
        self.viewer.commit()

    def _composite(self, cam, width, height):
        t = self._torch
        keys = self.viewer.trace_slab_keys(cam, width, height)
        kt = t.as_tensor(_DevMem(keys, width * height, "<i8"), device=t.device("cuda", self.device))
        with t.cuda.stream(self._stream):
            self.dist.all_reduce(kt, op=self.dist.ReduceOp.MIN)  # keys are < 2^62: signed MIN == unsigned MIN
        self._synced = True  # every rank's trace (its halo reads) precedes the completion of this all-reduce
        return keys

    def trace_device(self, cam, width, height):
        if self.world == 1:
            return self.viewer.trace_device(cam, width, height)
        if self.linked:
            return self.viewer.trace_linked_device(cam, width, height)
        return self._composite(cam, width, height)

    def resample_box(self, box, count=False):
        self._before_fill()
        n = self.viewer.resample_box(box, count)
        self._after_fill()
        return n

    def reset(self, loading_passes):
        self._before_fill()
        self.viewer.reset(loading_passes)
        if self.world > 1 and not self.linked:
            self._barrier()

    def gather_distance_volume(self):
        """Replicate the distance channel of the whole grid on every rank (4 bytes per voxel): each rank
        extracts its own slices, then one broadcast per slab over NCCL.  Needed again after every fill."""
        t = self._torch
        ptr, _first, _count = self.viewer.exact_trace_prepare()
        W, H, D = self.dims
        full = t.as_tensor(_DevMem(ptr, W * H * D, "<f4"), device=t.device("cuda", self.device))
        with t.cuda.stream(self._stream):
            gather_slabs(self.dist, full, self.dims, self.world)

    def trace_exact_device(self, cam, width, height, gather=True):
        """trace_exact_host without the download: returns the device pointer of the composited keys."""
        if self.world == 1:
            return self.viewer.trace_device(cam, width, height)
        t = self._torch
        if gather:
            self.gather_distance_volume()
        keys = self.viewer.trace_exact_keys(cam, width, height)
        kt = t.as_tensor(_DevMem(keys, width * height, "<i8"), device=t.device("cuda", self.device))
        with t.cuda.stream(self._stream):
            self.dist.all_reduce(kt, op=self.dist.ReduceOp.MIN)
        self._synced = True
        return keys

    def trace_exact_host(self, cam, width, height, gather=True):
        """The frame a single GPU holding the whole grid would trace, bit for bit (RGBA8 + depth): every
        rank marches every ray through the replicated distance volume, the owner of each hit shades it,
        an all-reduce(MIN) composites.  `gather=False` re-uses the distance volume of the previous call
        (the volume has not changed: camera motion only)."""
        if self.world == 1:
            return self.viewer.trace_rgba8(cam, width, height)
        keys = self.trace_exact_device(cam, width, height, gather)
        return self.viewer.keys_download(keys, width, height)

    def trace_host(self, cam, width, height, rgba_out=None, depth_out=None):
        """Frame to host memory.  One GPU: RGBA32F + depth (the shader's outputs).  Sharded: the
        composited RGBA8 + depth (what the keys carry); returned on every rank."""
        if self.world == 1:
            if rgba_out is not None and rgba_out.dtype == np.uint8:
                return self.viewer.trace_rgba8(cam, width, height, rgba_out, depth_out)
            out = {}
            if rgba_out is not None:
                out["rgba"] = rgba_out
            if depth_out is not None:
                out["depth"] = depth_out
            r, d, _ = self.viewer.trace(cam, width, height, out=out)
            return r, d
        if self.linked:  # the frame lands in rank 0's (pinned) buffers; the other ranks only take part
            r, d, _ = self.viewer.trace_linked(cam, width, height, rgba_out, depth_out, presenter=(self.rank == 0))
            return r, d
        keys = self._composite(cam, width, height)
        return self.viewer.keys_download(keys, width, height, rgba_out, depth_out)

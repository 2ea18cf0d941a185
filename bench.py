#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json): SDF samples/s of the grid fill
and primary rays/s of the sphere trace, demo_sdf at 512^3 / 1920x1080 on one B200; Z-sharded
weak scaling on N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one full fill of the grid (every voxel sampled once through the tape) + one trace of the
frame.  `value` is voxels / step time with everything resident in HBM (CUDA events on the library's
stream; the fill kernel alone is `fill_samples_per_sec` and the `roofline` object); `e2e` is the same through the host-buffer C-ABI calls (tape H2D, frame D2H inside the
timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
BYTES_PER_SAMPLE = 32  # tex0 + tex1 RGBA32F (scene/sdf/mod.rs:76,196-208)


def grid_for(n_gpus, side):
    """Weak scaling: every rank owns side^3 voxels.  1: s^3, 2: s x s x 2s ... 8: (2s)^3."""
    dims = [side, side, side]
    k, axis = n_gpus, 2
    while k > 1:
        dims[axis] *= 2
        k //= 2
        axis = (axis - 1) % 3
    return tuple(dims)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.06:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1]); power.append(float(f[2]))
            except Exception:
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def ncu_traffic():
    """dram bytes read + written per fill launch, from the committed `ncu --set full` capture."""
    try:
        rd = wr = None
        for line in open(os.path.join(ROOT, "profiles", "r01_fill_ncu_summary.txt")):
            k, _, v = line.partition(" = ")
            if k.startswith("dram__bytes_"):
                num, unit = v.split()[:2]
                val = float(num) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                if "read" in k: rd = val
                else: wr = val
        return rd + wr
    except Exception:
        return None


def trace_profile():
    """L1/TEX and L2 hit rates and DRAM traffic of the trace kernel from the committed ncu capture."""
    keys = {"l1tex__t_sector_hit_rate.pct": "l1tex_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
            "dram__bytes_read.sum": "dram_read", "gpu__time_duration.sum": "ncu_duration_us"}
    out = {}
    try:
        for line in open(os.path.join(ROOT, "profiles", "r01_trace_ncu_summary.txt")):
            k, _, v = line.partition(" = ")
            if k in keys:
                num, unit = (v.split() + [""])[:2]
                out[keys[k]] = float(num) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}.get(unit, 1)
        out["source"] = "profiles/r01_trace_ncu_summary.txt (ncu --set full; 512^3 volume, 1920x1080, default camera)"
        return out
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_fill_sample(orc, tape, dims, threads, target_s=6.0):
    """Oracle fill (OpenMP over z, all host threads) on a bounded sample of the same grid; samples/s."""
    v = orc.Viewer(BB, dims, 1)
    s = orc.Sampler(tape=tape)
    mid = dims[2] // 2
    probe = min(dims[2], max(1, 2 * threads))  # enough slices to occupy every thread
    z0 = max(0, mid - probe // 2)
    t = time.perf_counter(); v.fill_all(s, z0, z0 + probe, threads); per_slice = (time.perf_counter() - t) / probe
    nz = max(probe, min(dims[2], int(target_s / max(per_slice, 1e-9))))
    reps = max(1, min(16, int(target_s / max(per_slice * nz, 1e-9))))
    z0 = max(0, mid - nz // 2)
    t = time.perf_counter()
    n = sum(v.fill_all(s, z0, z0 + nz, threads) for _ in range(reps))
    dt = time.perf_counter() - t
    return n / dt, (f"{reps} x z slices [{z0},{z0 + nz}) of {dims[0]}x{dims[1]}x{dims[2]} "
                    f"({n} samples, {dt:.2f} s wall, {threads} threads)")


def cpu_reference_order_rate(orc, tape, side=96):
    """The reference's own configuration: ONE thread, LoadingManager visit order, 2 passes
    (scene/sdf/mod.rs:173-215 is single-threaded); a small grid bounds the run."""
    v = orc.Viewer(BB, (side, side, side), 2)
    s = orc.Sampler(tape=tape)
    t = time.perf_counter(); it = v.update(s); dt = time.perf_counter() - t
    return side ** 3 / dt, f"{side}^3 grid, 2 passes, {it} iterations, reference visit order, 1 thread, {dt:.2f} s"


def cpu_trace_sample(orc, S, tape, dims, W, H, threads, rows=48):
    """Oracle trace (the restated fragment shader) of a band of rows of the same frame; rays/s."""
    v = orc.Viewer(BB, dims, 1)
    v.fill_all(orc.Sampler(tape=tape), threads=threads)
    cam = S.default_camera(W, H)
    P = orc.trace_params(S.camera_rays(cam, W, H), BB, dims, lod=1.0, filter_linear=1)
    r0 = max(0, H // 2 - rows // 2)
    t = time.perf_counter()
    orc.trace(P, v.tex0, v.tex1, W, H, rows=(r0, r0 + rows), gbuf=False, threads=threads)
    dt = time.perf_counter() - t
    return W * rows / dt, f"rows [{r0},{r0 + rows}) of the {W}x{H} frame ({W * rows} rays, {dt:.2f} s, {threads} threads)"


def host_sampled_path(orc, S, threads, side=256):
    """SURVEY 8f row 1 measured: a surface WITHOUT a tape (what every existing .wasm SDF is) loaded through
    sdfgpu_update_surface -- the library walks the LoadingManager, calls the surface's sample_batch on the
    host cores (here the oracle's SDFDemo::sample standing in for the guest, C to C, no Python in the loop)
    and scatters the results on the GPU.  CPU-bound by construction; reported beside the CPU's own rate."""
    import ctypes as C
    from sdf_viewer_b200 import _lib
    P = orc.demo_params()
    surf = _lib.Surface()
    surf.self = C.cast(C.pointer(P), C.c_void_p)
    surf.sample_batch = C.cast(orc.lib().orc_demo_sample, _lib.SAMPLE_BATCH_FN)
    surf.sample_threads = threads
    with S.SDFViewer.new_voxels((side, side, side), BB, 2) as v:
        it = C.c_uint64()
        total = len(v.loading_mgr)
        t = time.perf_counter()
        while len(v.loading_mgr):
            S.viewer.check(v._lib.sdfgpu_update_surface(v._h, C.byref(surf), 0.030, C.byref(it)), v._h)
        v.sync()
        dt = time.perf_counter() - t
    return {"value": side ** 3 / dt, "unit": "samples/s", "threads": threads,
            "sample": f"{side}^3 grid, 2 passes ({total} iterations), 30 ms per update call as the scene does, {dt:.2f} s"}


def trace_modes_live(torch, v, stream, cam, W, H, reps=20):
    """Trace time for the other distance sources of the march (option trace_distance_volume; the step above uses 0,
    tex0.r in place): dense R32F copy, the same as a 3-D CUDA array through the TMU in point mode (both give the
    identical frame), and hardware LINEAR filtering (approximate, outside the 1e-5 bar).  The volume is not
    re-filled in between, which is the situation these modes are for (many frames per fill)."""
    out = {}
    try:
        for mode, name in ((0, "tex0_in_place_ms"), (1, "dense_r32f_ms"), (2, "tmu_point_ms"), (3, "tmu_hw_linear_approx_ms")):
            v.set_option("trace_distance_volume", mode)
            v.trace_device(cam, W, H)  # builds this mode's distance volume
            v.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                v.trace_device(cam, W, H)
            e1.record(stream)
            v.sync()
            torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / reps
    except Exception as e:  # an extra: never a reason to lose the bench line
        out["error"] = str(e)
    finally:
        try:
            v.set_option("trace_distance_volume", 0)
        except Exception:
            pass
    return out


def workload_name(workload, dims, W, H):
    return (f"{ {'demo': 'demo_sdf', 'csg': 'csg_1k', 'wasm': 'demo_sdf as a WebAssembly guest lowered to a scalar program'}[workload]} {dims[0]}x{dims[1]}x{dims[2]} "
            f"grid fill + {W}x{H} sphere trace, default scene camera")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Rust/wasmer build cannot
    be produced here (no cargo/rustc), so this is the C++ restatement in oracle/ (kind "port"), on all
    host cores (the reference loop itself is single-threaded, scene/sdf/mod.rs:174)."""
    if rank != 0:
        return
    import orc
    import sdf_viewer_b200.tape as T
    orc.build()
    dims = grid_for(args.gpus, args.grid)
    tape = T.demo_tape()
    threads = orc.lib().orc_max_threads()
    rates, sample = [], ""
    per_step = max(0.5, min(20.0, 90.0 / max(1, args.steps + args.warmup)))  # the whole run: about a minute and a half
    for i in range(args.warmup + args.steps):
        r, sample = cpu_fill_sample(orc, tape, dims, threads, target_s=per_step)
        if i >= args.warmup:
            rates.append(r)
    val = sum(rates) / len(rates)
    n_vox = dims[0] * dims[1] * dims[2]
    import sdf_viewer_b200 as S
    one_thread, one_thread_sample = cpu_reference_order_rate(orc, tape)
    cpu_dims = tuple(min(d, 256) for d in dims)  # the CPU trace sample marches a volume the host can hold
    rays, rays_sample = cpu_trace_sample(orc, S, tape, cpu_dims, args.width, args.height, threads)
    print(json.dumps({
        "impl": "reference", "metric": "sdf_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_vox / val, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name("demo", dims, args.width, args.height),
                   "step": "bounded sample of the grid fill on the host cores (C++ port of the reference loop, OpenMP over z)"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                         "reference_order_1_thread": one_thread, "reference_order_sample": one_thread_sample,
                         "rays_per_sec": rays, "rays_sample": rays_sample + f", {cpu_dims[0]}^3 volume"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="voxels per side owned by each GPU")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--workload", default="demo", choices=["demo", "csg", "wasm"],
                    help="demo: SDFDemo (the headline); csg: 1000 primitives; wasm: SDFDemo again, but as a WebAssembly guest "
                         "(hand-compiled in tests/test_wasm_lower.py) lowered to a scalar program by sdfgpu_wasm_lower")
    ap.add_argument("--vpt", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exact-trace", action="store_true",
                    help="N > 1: trace with the replicated distance volume + owner shading (bit-exact frame) instead of sort-last; "
                         "the gather of the distance channel after every fill is inside the step")
    ap.add_argument("--no-fused-halo", action="store_true", help="exchange halos with NCCL send/recv instead of in-kernel peer stores")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import sdf_viewer_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    dims = grid_for(n_gpus, args.grid)
    W, H = args.width, args.height
    if args.workload == "wasm":
        import test_wasm_lower  # test infrastructure: the guest module is assembled there (no WASM toolchain here)
        tape, _, lowering = S.wasm.lower(test_wasm_lower.guest_reference_demo().build())
    else:
        tape = S.tape.demo_tape() if args.workload == "demo" else S.tape.csg_tape()
    cam = S.default_camera(W, H)

    from sdf_viewer_b200.sharded import ShardedViewer
    sv = ShardedViewer(dims, BB, 2, rank=rank, world=world, device=local, group=dist, fused=not args.no_fused_halo)
    v = sv.viewer
    if args.vpt:
        v.set_option("fill_voxels_per_thread", args.vpt)
    if args.ctas:
        v.set_option("fill_ctas_per_sm", args.ctas)
    v.set_tape(tape)
    stream = torch.cuda.ExternalStream(v.stream, device=torch.device("cuda", local))
    own_voxels = dims[0] * dims[1] * (v.z_end - v.z_begin)
    total_voxels = dims[0] * dims[1] * dims[2]

    def step_device(ev=None):
        if ev: ev[0].record(stream)
        sv.fill_all()          # fill own slab (+ NCCL halo exchange when world > 1)
        sv.commit()            # SDFViewer::commit: lod = 1 -> LINEAR filtering (scene/sdf/mod.rs:226-238)
        if ev: ev[1].record(stream)
        if args.exact_trace and world > 1:
            sv.trace_exact_device(cam, W, H)
        else:
            sv.trace_device(cam, W, H)  # frame stays in HBM (+ MIN-composite over ranks when world > 1)
        if ev: ev[2].record(stream)

    def sync_all():
        v.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    l0 = v.launch_count
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    clk = ClockSampler(local) if rank == 0 else None
    wall0 = time.time()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_device(evs[i])
    sync_all()
    elapsed = time.perf_counter() - t0
    wall1 = time.time()
    launches = v.launch_count - l0
    fill_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    trace_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    step_ms = evs[0][0].elapsed_time(evs[-1][2]) / args.steps
    if dist:
        t = torch.tensor([step_ms, fill_ms, trace_ms, elapsed * 1e3 / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, fill_ms, trace_ms, wall_ms = t.tolist()
    else:
        wall_ms = elapsed * 1e3 / args.steps

    # ---- end to end through the host-buffer ABI: tape H2D + fill + trace + frame D2H
    rgba_h = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()  # the reference presents an RGBA8 framebuffer
    depth_h = torch.empty((H, W), dtype=torch.float32).pin_memory().numpy()
    e2e_steps = max(3, min(args.steps, 30))

    def step_e2e():
        v.set_tape(tape)
        sv.fill_all()
        sv.commit()
        return sv.trace_host(cam, W, H, rgba_h, depth_h)

    for _ in range(2):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    clocks = clk.stop(wall0, time.time()) if clk else None
    hit_frac = float((depth_h < 1.0).mean())
    trace_modes = trace_modes_live(torch, v, stream, cam, W, H) if n_gpus == 1 else None

    if rank != 0:
        if dist: dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    achieved = own_voxels * BYTES_PER_SAMPLE / (fill_ms * 1e-3) / 1e9
    out = {
        # value: voxels of the whole job / the time of the whole step (fill + commit + trace), the same
        # K-step interval ms_per_step reports; the fill kernel alone is fill_samples_per_sec / roofline
        "metric": "sdf_samples_per_sec", "value": total_voxels / (step_ms * 1e-3), "unit": "samples/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, dims, W, H),
                   "sharding": (f"z-slabs x{n_gpus}, halo exchange: " + ("fused: boundary slices pushed by DMA over NVLink (CUDA IPC) while the interior fills"
                                if sv.fused else "NCCL send/recv after the fill")) if n_gpus > 1 else "single GPU",
                   "l2": "volume (32 B/voxel) exceeds the 126 MB L2, no flush needed" if own_voxels * 32 > 2.5e8 else "volume fits L2",
                   "step": "fill_all + commit + trace (lod 1, LINEAR filter, fp32 trilinear)" +
                           (", exact sharded trace (distance channel gathered after the fill, owner shading)"
                            if args.exact_trace and n_gpus > 1 else "")},
        "fill_ms": fill_ms, "trace_ms": trace_ms, "fill_samples_per_sec": total_voxels / (fill_ms * 1e-3),
        "rays_per_sec": W * H / (trace_ms * 1e-3), "hit_fraction": hit_frac,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic() if (n_gpus == 1 and args.grid == 512 and args.workload == "demo") else None,
                     "traffic_source": "profiles/r01_fill_ncu_summary.txt (ncu --set full, same kernel and grid)", "peak_source": peak_src, "kernel": "fill_kernel",
                     "algorithmic_bytes_per_launch": own_voxels * BYTES_PER_SAMPLE},
        "e2e": {"value": total_voxels / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": len(tape) + 256, "d2h_bytes_per_step": W * H * 8,
                "what": "set_tape (H2D) + fill + commit + trace + frame RGBA8+depth D2H into pinned host memory"},
        "gpu_launches": int(launches), "clocks": clocks, "host_ms_per_step": wall_ms,
        "trace_modes": trace_modes,
        "trace_profile": trace_profile() if (n_gpus == 1 and args.grid == 512 and args.workload == "demo") else None,
    }
    if n_gpus > 1:
        out["e2e"]["d2h_bytes_per_step"] = W * H * 8
        out["e2e"]["what"] = ("set_tape (H2D) + fill + NCCL halo exchange + slab trace + all-reduce(MIN) composite + "
                              "frame RGBA8+depth D2H").replace("NCCL halo exchange", "fused halo exchange" if sv.fused else "NCCL halo exchange")
    if not args.no_cpu_baseline and n_gpus == 1:
        import orc
        orc.build()
        threads = orc.lib().orc_max_threads()
        val, sample = cpu_fill_sample(orc, tape, dims, threads, target_s=8.0)
        one_thread, one_thread_sample = cpu_reference_order_rate(orc, tape)
        cpu_dims = tuple(min(d, 256) for d in dims)
        rays, rays_sample = cpu_trace_sample(orc, S, tape, cpu_dims, W, H, threads)
        out["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                               "reference_order_1_thread": one_thread, "reference_order_sample": one_thread_sample,
                               "rays_per_sec": rays, "rays_sample": rays_sample + f", {cpu_dims[0]}^3 volume"}
        try:
            out["host_sampled_path"] = host_sampled_path(orc, S, threads)
        except Exception as e:  # an extra, never a reason to lose the bench line
            out["host_sampled_path"] = {"error": str(e)}
    print(json.dumps(out))
    if dist: dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json): SDF samples/s of the grid fill
and primary rays/s of the sphere trace, demo_sdf at 512^3 / 1920x1080 on one B200; Z-sharded
weak scaling on N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one full fill of the grid (every voxel sampled once through the tape) + one trace of the
frame.  `value` is voxels / step time with everything resident in HBM (CUDA events on the library's
stream; the fill kernel alone is `fill_samples_per_sec` and the `roofline` object); `e2e` is the same
through the host-buffer C-ABI calls (tape H2D, frame D2H into pinned memory inside the timed region).
N > 1: one process per GPU, slab handles linked through the C ABI (include/sdfgpu.h "linked slabs");
torch.distributed is used at set-up (link blobs) and for the barriers / max-over-ranks of the timing
contract only -- the step itself makes no collective-library call.  After the timed region every N > 1
run checks itself (`parity_check`).  Prints ONE JSON line on rank 0.

The other BASELINE.json configs ride along as extra keys (outside the timed region): `csg_1k_512` (C3),
`dirty_60hz` (C5), `trace_closeup`, `mesh` at N = 1; `c4_trace_2160p` at N > 1.
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
PROGRAM_KERNEL = {0: "fill_kernel<interpreter>", 1: "sdfgpu_fill_jit", 2: "fill_kernel<demo>"}
FILL_CAPTURE = "profiles/r02_final_fill_ncu_summary.txt"
TRACE_CAPTURE = "profiles/r02_final_trace_ncu_summary.txt"


def grid_for(n_gpus, side):
    """Weak scaling: every rank owns side^3 voxels.  1: s^3, 2: s x s x 2s ... 8: (2s)^3."""
    dims = [side, side, side]
    k, axis = n_gpus, 2
    while k > 1:
        dims[axis] *= 2
        k //= 2
        axis = (axis - 1) % 3
    return tuple(dims)


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1; the CPU legs pass this count to
    the oracle explicitly so that every N is measured against the same number of cores."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(workload, dims, W, H):
    return (f"{ {'demo': 'demo_sdf', 'csg': 'csg_1k', 'wasm': 'demo_sdf as a WebAssembly guest lowered to a scalar program'}[workload]} "
            f"{dims[0]}x{dims[1]}x{dims[2]} grid fill + {W}x{H} sphere trace, default scene camera")


def common_config(args, dims):
    """Identical in both arms."""
    return {"workload": workload_name(args.workload, dims, args.width, args.height), "grid": list(dims),
            "frame": [args.width, args.height], "voxels_per_gpu": args.grid ** 3, "n_gpus": args.gpus,
            "step": "fill of every voxel + trace of the frame"}


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


CAPTURE_COMMITS = {"profiles/r02_final_fill_ncu_summary.txt": "a405693 (round 2)", "profiles/r02_final_trace_ncu_summary.txt": "a405693 (round 2)",
                   "profiles/r02_final_fill_csg_ncu_summary.txt": "a405693 (round 2)"}


def capture_commit(path):
    """The commit that last touched a committed capture (so a stale capture is visible in the line); the GPU box
    has no .git, there the recorded hash is used."""
    if not os.path.isdir(os.path.join(ROOT, ".git")):
        return CAPTURE_COMMITS.get(path)
    try:
        r = subprocess.run(["git", "-C", ROOT, "log", "-1", "--format=%h %cs", "--", path], capture_output=True, text=True, timeout=10)
        return r.stdout.strip() or None
    except Exception:
        return None


def ncu_metrics(path, keys):
    out = {}
    try:
        for line in open(os.path.join(ROOT, path)):
            k, _, v = line.partition(" = ")
            if k in keys:
                num, unit = (v.split() + [""])[:2]
                out[keys[k]] = float(num) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    except Exception:
        return None
    return out


def ncu_traffic():
    """dram bytes read + written per fill launch, from the committed `ncu --set full` capture of this kernel."""
    m = ncu_metrics(FILL_CAPTURE, {"dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr"})
    return m["rd"] + m["wr"] if m and "rd" in m and "wr" in m else None


def trace_profile():
    """L1/TEX and L2 hit rates and DRAM traffic of the trace kernel from the committed ncu capture."""
    out = ncu_metrics(TRACE_CAPTURE, {"l1tex__t_sector_hit_rate.pct": "l1tex_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
                                      "dram__bytes_read.sum": "dram_read", "gpu__time_duration.sum": "ncu_duration_us"})
    if out:
        out["source"] = TRACE_CAPTURE + " (ncu --set full; 512^3 volume, 1920x1080, default camera)"
        out["capture_commit"] = capture_commit(TRACE_CAPTURE)
    return out


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ CPU legs (oracle)

def cpu_sampler(orc, workload, tape):
    """The closest stand-in for the reference's native SDFDemo::sample: the oracle's direct restatement
    (oracle/sdf_oracle.cpp demo_sample); other workloads have no native form and run the oracle's tape evaluator."""
    return orc.Sampler() if workload == "demo" else orc.Sampler(tape=tape)


def cpu_fill_sample(orc, sampler, dims, threads, target_s=6.0):
    """Oracle fill (OpenMP over z, `threads` host threads) on a bounded sample of the same grid; samples/s."""
    v = orc.Viewer(BB, dims, 1)
    s = sampler
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


def cpu_reference_order_rate(orc, sampler, side=96):
    """The reference's own configuration: ONE thread, LoadingManager visit order, 2 passes
    (scene/sdf/mod.rs:173-215 is single-threaded); a small grid bounds the run."""
    v = orc.Viewer(BB, (side, side, side), 2)
    t = time.perf_counter(); it = v.update(sampler); dt = time.perf_counter() - t
    return side ** 3 / dt, f"{side}^3 grid, 2 passes, {it} iterations, reference visit order, 1 thread, {dt:.2f} s"


def cpu_trace_sample(orc, sampler, dims, W, H, threads, rows=48):
    """Oracle trace (the restated fragment shader) of a band of rows of the same frame; rays/s."""
    v = orc.Viewer(BB, dims, 1)
    v.fill_all(sampler, threads=threads)
    P = orc.trace_params(orc.default_rays(W, H), BB, dims, lod=1.0, filter_linear=1)
    r0 = max(0, H // 2 - rows // 2)
    t = time.perf_counter()
    orc.trace(P, v.tex0, v.tex1, W, H, rows=(r0, r0 + rows), gbuf=False, threads=threads)
    dt = time.perf_counter() - t
    return W * rows / dt, f"rows [{r0},{r0 + rows}) of the {W}x{H} frame ({W * rows} rays, {dt:.2f} s, {threads} threads)"


def cpu_baseline_block(orc, workload, tape, dims, W, H, threads, target_s):
    sampler = cpu_sampler(orc, workload, tape)
    val, sample = cpu_fill_sample(orc, sampler, dims, threads, target_s=target_s)
    one_thread, one_thread_sample = cpu_reference_order_rate(orc, sampler)
    cpu_dims = tuple(min(d, 256) for d in dims)  # the CPU trace sample marches a volume the host can hold
    rays, rays_sample = cpu_trace_sample(orc, sampler, cpu_dims, W, H, threads)
    return {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
            "sampler": "direct restatement of SDFDemo::sample (oracle demo_sample)" if workload == "demo" else "oracle tape evaluator",
            "reference_order_1_thread": one_thread, "reference_order_sample": one_thread_sample,
            "rays_per_sec": rays, "rays_sample": rays_sample + f", {cpu_dims[0]}^3 volume"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Rust/wasmer build cannot
    be produced here (no cargo/rustc), so this is the C++ restatement in oracle/ (kind "port"), on all
    host cores (the reference loop itself is single-threaded, scene/sdf/mod.rs:174).  Loads oracle/liboracle.so
    only -- nothing of the product."""
    if rank != 0:
        return
    import orc
    orc.build()
    dims = grid_for(args.gpus, args.grid)
    threads = host_threads()
    tape = None
    if args.workload != "demo":
        raise SystemExit("--impl reference runs the demo workload (the reference's SDFDemo)")
    rates, sample = [], ""
    # the whole run: about a minute and a half (SDFGPU_BENCH_REF_SECONDS: another budget, for the CPU test of this arm)
    budget = float(os.environ.get("SDFGPU_BENCH_REF_SECONDS", "90"))
    per_step = max(0.5 if budget >= 30 else 0.02, min(20.0, budget / max(1, args.steps + args.warmup)))
    sampler = cpu_sampler(orc, "demo", tape)
    for i in range(args.warmup + args.steps):
        r, sample = cpu_fill_sample(orc, sampler, dims, threads, target_s=per_step)
        if i >= args.warmup:
            rates.append(r)
    val = sum(rates) / len(rates)
    n_vox = dims[0] * dims[1] * dims[2]
    one_thread, one_thread_sample = cpu_reference_order_rate(orc, sampler)
    cpu_dims = tuple(min(d, 256) for d in dims)
    rays, rays_sample = cpu_trace_sample(orc, sampler, cpu_dims, args.width, args.height, threads)
    print(json.dumps({
        "impl": "reference", "metric": "sdf_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_vox / val, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args, dims),
        "detail": {"step": "bounded sample of the grid fill on the host cores (C++ port of the reference loop, OpenMP over z); "
                           "ms_per_step extrapolates the sample's rate to the whole grid",
                   "omp_threads": threads, "omp_env": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                         "sampler": "direct restatement of SDFDemo::sample (oracle demo_sample)",
                         "reference_order_1_thread": one_thread, "reference_order_sample": one_thread_sample,
                         "rays_per_sec": rays, "rays_sample": rays_sample + f", {cpu_dims[0]}^3 volume"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU extras

def timed(torch, v, stream, fn, reps):
    fn(); v.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream); v.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


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
    identical frame), and hardware LINEAR filtering (approximate, outside the 1e-5 bar); and for the persistent-warp
    kernel (trace_variant 2).  The volume is not re-filled in between (many frames per fill)."""
    out = {}
    try:
        for mode, name in ((0, "tex0_in_place_ms"), (1, "dense_r32f_ms"), (2, "tmu_point_ms"), (3, "tmu_hw_linear_approx_ms")):
            v.set_option("trace_distance_volume", mode)
            out[name] = timed(torch, v, stream, lambda: v.trace_device(cam, W, H), reps)
        v.set_option("trace_distance_volume", 0)
        v.set_option("trace_variant", 2)
        out["persistent_warps_ms"] = timed(torch, v, stream, lambda: v.trace_device(cam, W, H), reps)
    except Exception as e:  # an extra: never a reason to lose the bench line
        out["error"] = str(e)
    finally:
        try:
            v.set_option("trace_distance_volume", 0)
            v.set_option("trace_variant", 0)
        except Exception:
            pass
    return out


def extras_single_gpu(torch, S, v, stream, W, H, side, peak):
    """BASELINE.json configs C3 and C5 and the close-up trace, on the handle of the bench (refilled afterwards by
    nobody: this runs after every timed region)."""
    out = {}
    try:  # ---- C3: CSG of 1000 random primitives, 512^3 (tape-interpreter stress)
        v.set_tape(S.tape.csg_tape())
        ms = timed(torch, v, stream, v.fill_all, 5)
        prog = v.get_info("last_fill_program")
        v.commit()
        cam = S.default_camera(W, H)
        tms = timed(torch, v, stream, lambda: v.trace_device(cam, W, H), 10)
        alu = ncu_metrics("profiles/r02_final_fill_csg_ncu_summary.txt",
                          {"smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
                           "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
                           "smsp__inst_executed.sum": "warp_instructions"})
        out["csg_1k_512"] = {
            "config": f"BASELINE.json configs[2]: union of 1000 random primitives, {side}^3, {W}x{H}",
            "fill_ms": ms, "samples_per_sec": side ** 3 / (ms * 1e-3), "kernel": PROGRAM_KERNEL.get(prog, str(prog)),
            "voxels_per_thread": v.get_info("last_fill_voxels_per_thread"), "tile_culling": bool(v.get_info("tape_culled")),
            "hbm_frac": side ** 3 * BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9 / peak, "trace_ms": tms,
            "rays_per_sec": W * H / (tms * 1e-3),
            "ncu": dict(alu or {}, source="profiles/r02_final_fill_csg_ncu_summary.txt", capture_commit=capture_commit("profiles/r02_final_fill_csg_ncu_summary.txt"))}
        try:
            st = v.cull_stats()
            out["csg_1k_512"]["cull_survivors_per_tile"] = st
        except Exception as e:
            out["csg_1k_512"]["cull_survivors_per_tile"] = {"error": str(e)}
    except Exception as e:
        out["csg_1k_512"] = {"error": str(e)}
    try:  # ---- C5: animated parameter, whole-bbox change (what the reference demo reports), 3-pass re-sample
        sdf = S.SDFDemo()
        v.reset(2)
        v.update(sdf)
        v.sync()

        def sweep():
            sdf.set_parameter("sphere_radius", 1.05 if sdf.params["sphere_radius"] < 1.05 else 1.04)
            v.update(sdf)

        ms = timed(torch, v, stream, sweep, 10)
        h = 128 / float(side - 1)
        box = (-h, -h, -h, h, h, h)
        n = v.resample_box(box, count=True)
        us = 1e3 * timed(torch, v, stream, lambda: v.resample_box(box), 50)
        out["dirty_60hz"] = {
            "config": f"BASELINE.json configs[4]: sphere_radius edited every frame (demo/sphere.rs:75-84), {side}^3",
            "whole_bbox_change_ms": ms, "whole_bbox_change_hz": 1e3 / ms,
            "whole_bbox_what": "set_tape + the 3-pass re-sample of scene/sdf/mod.rs:144-153 (passes 1, 2 conditional)",
            "frame_budget_fraction_at_60hz": ms / (1e3 / 60.0),
            "sub_block_voxels": int(n), "sub_block_us": us, "sub_block_samples_per_sec": n / (us * 1e-6),
            "sub_block_what": "sdfgpu_resample_box of a 128^3 dirty AABB (mod.rs:184-190 restricted to its index range)"}
    except Exception as e:
        out["dirty_60hz"] = {"error": str(e)}
    try:  # ---- the close-up camera (87 % of the rays enter the box)
        v.set_tape(S.tape.demo_tape()); v.fill_all(); v.commit()
        cam = S.look_at_camera((0.9, 1.1, 1.8), (0, 0, 0), W, H)
        res = {}
        for variant in (0, 2):
            v.set_option("trace_variant", variant)
            ms = timed(torch, v, stream, lambda: v.trace_device(cam, W, H), 20)
            res[f"variant{variant}_ms"] = ms
            res[f"variant{variant}_rays_per_sec"] = W * H / (ms * 1e-3)
        v.set_option("trace_variant", 0)
        out["trace_closeup"] = res
    except Exception as e:
        out["trace_closeup"] = {"error": str(e)}
    try:  # ---- isosurface mesh from the resident volume
        if hasattr(v, "mesh"):
            cells = (side - 1) ** 3
            v.mesh(download=False); v.sync()
            t = time.perf_counter()
            reps = 5
            for _ in range(reps):
                nv, nt = v.mesh(download=False)
            v.sync()
            dt = (time.perf_counter() - t) / reps
            out["mesh"] = {"vertices": int(nv), "triangles": int(nt), "ms": dt * 1e3, "triangles_per_sec": nt / dt,
                           "cells_per_sec": cells / dt, "what": "marching cubes over the resident distance volume, counts read back"}
    except Exception as e:
        out["mesh"] = {"error": str(e)}
    return out


def parity_check_multi(torch, dist, S, sv, tape, dims, cam, W, H, rank, world, local, rgba_h, depth_h):
    """After the timed region of an N > 1 run: (1) each halo slice equals the same slice computed locally,
    (2) random texels of the own slab equal the oracle's point samples, (3) the linked frame equals the frame
    of ONE handle holding the whole grid (rank 0, when the grid fits beside its slab).  Bit for bit."""
    import numpy as np
    from sdf_viewer_b200.sharded import _DevMem
    v = sv.viewer
    res = {}
    ok = True
    v.sync(); torch.cuda.synchronize(); dist.barrier()
    n = dims[0] * dims[1] * 4
    dev = torch.device("cuda", local)
    halos = 0
    for z in ([v.z_lo] if v.z_lo < v.z_begin else []) + ([v.z_hi - 1] if v.z_hi > v.z_end else []):
        with S.SDFViewer.new_voxels(dims, BB, 1, device=local, z_range=(z, z + 1)) as one:
            one.set_option("fill_halo", 0)
            one.set_tape(tape); one.fill_all(); one.sync()
            p0, p1 = one.device_ptrs()
            off = (z - one.z_lo) * n
            for p, mine in zip((p0, p1), sv._tex):
                t = torch.as_tensor(_DevMem(p, (one.z_hi - one.z_lo) * n, "<f4"), device=dev)[off:off + n]
                ok &= bool(torch.equal(t.view(torch.int32), mine[(z - v.z_lo) * n:(z - v.z_lo + 1) * n].view(torch.int32)))
            halos += 1
    res["halo_slices_checked"] = halos
    # (1b) each halo slice equals the neighbour's actual boundary slice (fetched with NCCL here, outside any timed region)
    ops, got = [], []
    for nb, z_send, z_halo in ((rank - 1, v.z_begin, v.z_lo), (rank + 1, v.z_end - 1, v.z_hi - 1)):
        if nb < 0 or nb >= world:
            continue
        for tex in sv._tex:
            buf = torch.empty(n, dtype=torch.float32, device=dev)
            ops += [dist.P2POp(dist.isend, tex[(z_send - v.z_lo) * n:(z_send - v.z_lo + 1) * n].clone(), nb), dist.P2POp(dist.irecv, buf, nb)]
            got.append((buf, tex[(z_halo - v.z_lo) * n:(z_halo - v.z_lo + 1) * n]))
    for r in dist.batch_isend_irecv(ops) if ops else []:
        r.wait()
    torch.cuda.synchronize()
    for buf, mine in got:
        ok &= bool(torch.equal(buf.view(torch.int32), mine.view(torch.int32)))
    res["halo_slices_vs_neighbour"] = len(got)
    try:
        import orc
        orc.build()
        rng = np.random.default_rng(1234 + rank)
        k = max(1000, 100000 // world)
        idx = np.stack([rng.integers(0, dims[0], k), rng.integers(0, dims[1], k), rng.integers(v.z_begin, v.z_end, k)], 1).astype(np.uint32)
        flat = ((idx[:, 2].astype(np.int64) - v.z_lo) * dims[1] + idx[:, 1]) * dims[0] + idx[:, 0]
        ft = torch.from_numpy(flat).to(dev)
        o0, o1 = orc.Viewer(BB, dims, 1, alloc=False).sample_voxels(orc.Sampler(tape=tape), idx, threads=1)
        for mine, want in zip(sv._tex, (o0, o1)):
            got = mine.view(-1, 4)[ft].cpu().numpy()
            ok &= bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
        res["texels_vs_oracle"] = k * world
    except Exception as e:
        res["texels_vs_oracle"] = f"skipped: {e}"
    frame = "skipped (the whole grid does not fit beside rank 0's slab)"
    sv.trace_host(cam, W, H, rgba_h, depth_h)
    if rank == 0 and dims[0] * dims[1] * dims[2] * BYTES_PER_SAMPLE < 100e9:
        with S.SDFViewer.new_voxels(dims, BB, 2, device=local) as whole:
            whole.set_tape(tape); whole.fill_all(); whole.commit()
            want8, want_d = whole.trace_rgba8(cam, W, H)
        same = bool(np.array_equal(want8, rgba_h) and np.array_equal(np.clip(want_d, 0, 1).view(np.uint32), depth_h.view(np.uint32)))
        ok &= same
        frame = "bit-exact vs one handle holding the whole grid" if same else "DIFFERS from the single-handle frame"
    res["frame"] = frame
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["status"] = "ok" if flag.item() else "FAILED"
    return res


def alt_modes_multi(torch, dist, ShardedViewer, dims, tape, cam, W, H, rank, world, local, steps=20):
    """N > 1, after the timed region: the same step with the halo slices PUSHED by the neighbours and the trace in
    ROUNDS (the two alternatives the default replaces), so that every scaling run carries the comparison."""
    out = {}
    for name, kw in (("halo_push_stream_trace", {"halo_push": True}), ("halo_local_rounds_trace", {"trace_mode": 1})):
        sv = ShardedViewer(dims, BB, 2, rank=rank, world=world, device=local, group=dist, max_width=W, max_height=H, **kw)
        try:
            if not sv.linked:
                out[name] = {"error": getattr(sv, "link_error", "not linked")}
                continue
            v = sv.viewer
            v.set_tape(tape)
            stream = torch.cuda.ExternalStream(v.stream, device=torch.device("cuda", local))
            for _ in range(3):
                sv.fill_all(); sv.commit(); sv.trace_device(cam, W, H)
            v.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
            for e in evs:
                e[0].record(stream)
                sv.fill_all(); sv.commit()
                e[1].record(stream)
                sv.trace_device(cam, W, H)
                e[2].record(stream)
            v.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([evs[0][0].elapsed_time(evs[-1][2]) / steps, sum(e[0].elapsed_time(e[1]) for e in evs) / steps,
                              sum(e[1].elapsed_time(e[2]) for e in evs) / steps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name] = {"ms_per_step": t[0].item(), "fill_ms": t[1].item(), "trace_ms": t[2].item(), "steps": steps}
        except Exception as e:  # an extra, never a reason to lose the bench line
            out[name] = {"error": str(e)}
        finally:
            sv.close()
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 / C5 / close-up / 2160p legs that follow the timed region")
    ap.add_argument("--no-linked", action="store_true",
                    help="N > 1: the round-1 path (IPC halo pushes ordered by NCCL, sort-last trace + all-reduce(MIN)) instead of linked slabs")
    ap.add_argument("--halo-push", action="store_true",
                    help="N > 1, linked: the neighbours push the halo slices over NVLink (default: every rank fills its own)")
    ap.add_argument("--trace-rounds", action="store_true",
                    help="N > 1, linked: trace in `world` rounds of one kernel (default: one streaming kernel per rank)")
    ap.add_argument("--exact-trace", action="store_true",
                    help="with --no-linked: trace with the replicated distance volume + owner shading instead of sort-last")
    ap.add_argument("--no-fused-halo", action="store_true", help="with --no-linked: exchange halos with NCCL send/recv")
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
    args.gpus = n_gpus
    dims = grid_for(n_gpus, args.grid)
    W, H = args.width, args.height
    if args.workload == "wasm":
        import test_wasm_lower  # test infrastructure: the guest module is assembled there (no WASM toolchain here)
        tape, _, lowering = S.wasm.lower(test_wasm_lower.guest_reference_demo().build())
    else:
        tape = S.tape.demo_tape() if args.workload == "demo" else S.tape.csg_tape()
    cam = S.default_camera(W, H)

    from sdf_viewer_b200.sharded import ShardedViewer
    want_c4 = world > 1 and not args.no_extras
    def make_viewer(trace_mode):
        sv = ShardedViewer(dims, BB, 2, rank=rank, world=world, device=local, group=dist, fused=not args.no_fused_halo,
                           linked=not args.no_linked, max_width=max(W, 3840 if want_c4 else 0), max_height=max(H, 2160 if want_c4 else 0),
                           halo_push=args.halo_push, trace_mode=trace_mode)
        v = sv.viewer
        if args.vpt:
            v.set_option("fill_voxels_per_thread", args.vpt)
        if args.ctas:
            v.set_option("fill_ctas_per_sm", args.ctas)
        v.set_tape(tape)
        return sv, v, torch.cuda.ExternalStream(v.stream, device=torch.device("cuda", local))

    sv, v, stream = make_viewer(1 if args.trace_rounds else 0)
    own_voxels = dims[0] * dims[1] * (v.z_end - v.z_begin)
    total_voxels = dims[0] * dims[1] * dims[2]

    def step_device(ev=None):
        if ev: ev[0].record(stream)
        sv.fill_all()          # fill own slab (+ halo exchange when world > 1)
        sv.commit()            # SDFViewer::commit: lod = 1 -> LINEAR filtering (scene/sdf/mod.rs:226-238)
        if ev: ev[1].record(stream)
        if args.exact_trace and world > 1 and not sv.linked:
            sv.trace_exact_device(cam, W, H)
        else:
            sv.trace_device(cam, W, H)  # frame stays in HBM (rank 0's when world > 1)
        if ev: ev[2].record(stream)

    def sync_all():
        v.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # The warm-up doubles as the check that the ranks' streaming trace kernels find each other (they wait for one another
    # inside the kernel and give up after link_timeout_ms).  Should any rank fail here, EVERY rank falls back to the
    # trace in rounds -- the same frame, no kernel that waits -- rather than lose the line; the line then says so.
    fallback = None
    ok, why = 1, ""
    try:
        for _ in range(args.warmup):
            step_device()
        v.sync()
        if os.environ.get("SDFGPU_BENCH_FORCE_FALLBACK") and world > 1 and rank == world - 1:
            raise RuntimeError("forced by SDFGPU_BENCH_FORCE_FALLBACK (test of the fallback)")
    except Exception as e:
        ok, why = 0, str(e)
    if dist:
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        all_ok = bool(flag.item())
    else:
        all_ok = bool(ok)
    if not all_ok:
        if not (world > 1 and sv.linked and not args.trace_rounds):
            raise SystemExit(f"bench.py: the warm-up failed: {why or 'on another rank'}")
        whys = [None] * world
        dist.all_gather_object(whys, why)
        fallback = "streaming trace failed in the warm-up (" + "; ".join(f"rank {r}: {w}" for r, w in enumerate(whys) if w) + "): trace in rounds"
        try:
            sv.close()
        except Exception:
            pass
        sv, v, stream = make_viewer(1)
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
    launches = v.launch_count - l0
    fill_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    trace_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    step_ms = evs[0][0].elapsed_time(evs[-1][2]) / args.steps
    if dist:
        t = torch.tensor([step_ms, fill_ms, trace_ms, elapsed * 1e3 / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, fill_ms, trace_ms, wall_ms = t.tolist()
        t = torch.tensor([float(launches)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = int(t.item())
    else:
        wall_ms = elapsed * 1e3 / args.steps

    # ---- end to end through the host-buffer ABI: tape H2D + fill + trace + frame D2H into pinned memory (rank 0)
    rgba_h = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()  # the reference presents an RGBA8 framebuffer
    depth_h = torch.empty((H, W), dtype=torch.float32).pin_memory().numpy()
    e2e_steps = max(3, min(args.steps, 30))

    def step_e2e():
        v.set_tape(tape)
        sv.fill_all()
        sv.commit()
        return sv.trace_host(cam, W, H, rgba_h, depth_h)

    for _ in range(3):
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
    program = v.get_info("last_fill_program")
    peak, peak_src = peaks()

    # ---- after the timed regions
    parity = None
    c4 = None
    if world > 1:
        if sv.linked:
            parity = parity_check_multi(torch, dist, S, sv, tape, dims, cam, W, H, rank, world, local, rgba_h, depth_h)
        if want_c4 and sv.linked:
            try:
                cam4 = S.default_camera(3840, 2160)
                for _ in range(3):
                    sv.trace_device(cam4, 3840, 2160)
                sync_all()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(20):
                    sv.trace_device(cam4, 3840, 2160)
                e1.record(stream)
                sync_all()
                t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                r4 = torch.empty((2160, 3840, 4), dtype=torch.uint8).pin_memory().numpy()
                d4 = torch.empty((2160, 3840), dtype=torch.float32).pin_memory().numpy()
                sv.trace_host(cam4, 3840, 2160, r4, d4)
                sync_all()
                t0 = time.perf_counter()
                for _ in range(10):
                    sv.trace_host(cam4, 3840, 2160, r4, d4)
                sync_all()
                c4 = {"config": f"BASELINE.json configs[3]: {dims[0]}x{dims[1]}x{dims[2]} Z-sharded over {world} GPUs, 3840x2160 raymarch",
                      "trace_ms": t.item(), "rays_per_sec": 3840 * 2160 / (t.item() * 1e-3),
                      "e2e_frame_ms": (time.perf_counter() - t0) * 1e3 / 10, "e2e_what": "trace + RGBA8+depth D2H (66 MB) on rank 0",
                      "hit_fraction": float((d4 < 1.0).mean())}
            except Exception as e:
                c4 = {"error": str(e)}
    alt = None
    if world > 1 and sv.linked and not args.no_extras and not args.halo_push and not args.trace_rounds and not fallback:
        alt = alt_modes_multi(torch, dist, ShardedViewer, dims, tape, cam, W, H, rank, world, local)
    trace_modes = extras = None
    if n_gpus == 1 and not args.no_extras:
        trace_modes = trace_modes_live(torch, v, stream, cam, W, H)
        if args.grid == 512:
            extras = extras_single_gpu(torch, S, v, stream, W, H, args.grid, peak)

    if rank != 0:
        sv.close()
        if dist: dist.destroy_process_group()
        return
    achieved = own_voxels * BYTES_PER_SAMPLE / (fill_ms * 1e-3) / 1e9
    headline_cfg = n_gpus == 1 and args.grid == 512 and args.workload == "demo"
    if n_gpus == 1:
        sharding = "single GPU"
    elif sv.linked:
        sharding = (f"z-slabs x{n_gpus}, linked through the C ABI: one fill launch per rank, "
                    + ("boundary tiles first, DMA halo push behind a flag" if v.get_info("link_halo_push")
                       else "halo slices filled by their holder (no exchange: sample() is a pure function of the position)")
                    + ", exact ray-hand-off trace ("
                    + ("one streaming kernel per rank, rays pushed into the neighbour's queue over NVLink" if v.get_info("link_trace_stream")
                       else f"{n_gpus} rounds")
                    + "), pixels stored into rank 0's frame over NVLink; flags: "
                    + ("stream memory operations" if v.get_info("link_memops") else "spin-wait kernels"))
    else:
        sharding = (f"z-slabs x{n_gpus}, round-1 path: " + ("IPC halo pushes ordered by a 4-byte NCCL all-reduce" if sv.fused else "NCCL send/recv halo exchange")
                    + ", " + ("exact trace on a replicated distance volume" if args.exact_trace else "sort-last trace + all-reduce(MIN)"))
    out = {
        # value: voxels of the whole job / the time of the whole step (fill + commit + trace), the same
        # K-step interval ms_per_step reports; the fill kernel alone is fill_samples_per_sec / roofline
        "metric": "sdf_samples_per_sec", "value": total_voxels / (step_ms * 1e-3), "unit": "samples/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args, dims),
        "detail": {"sharding": sharding,
                   "l2": "volume (32 B/voxel) exceeds the 126 MB L2, no flush needed" if own_voxels * 32 > 2.5e8 else "volume fits L2",
                   "step": "fill_all + commit + trace (lod 1, LINEAR filter, fp32 trilinear)"},
        "fill_ms": fill_ms, "trace_ms": trace_ms, "fill_samples_per_sec": total_voxels / (fill_ms * 1e-3),
        "rays_per_sec": W * H / (trace_ms * 1e-3), "hit_fraction": hit_frac,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic() if headline_cfg else None,
                     "traffic_source": f"{FILL_CAPTURE} (ncu --set full, same kernel and grid; capture committed at {capture_commit(FILL_CAPTURE)})"
                     if headline_cfg else None,
                     "peak_source": peak_src, "kernel": PROGRAM_KERNEL.get(program, str(program)),
                     "algorithmic_bytes_per_launch": own_voxels * BYTES_PER_SAMPLE},
        "e2e": {"value": total_voxels / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": (len(tape) + 256) * n_gpus, "d2h_bytes_per_step": W * H * 8,
                "what": "set_tape (H2D, every rank) + fill + commit + trace + frame RGBA8+depth D2H into pinned host memory (rank 0)"},
        "gpu_launches": int(launches), "clocks": clocks, "host_ms_per_step": wall_ms,
        "trace_modes": trace_modes,
        "trace_profile": trace_profile() if headline_cfg else None,
    }
    if fallback:
        out["fallback"] = fallback
    if parity is not None:
        out["parity_check"] = parity["status"]
        out["parity_detail"] = parity
    if c4 is not None:
        out["c4_trace_2160p"] = c4
    if alt is not None:
        out["alternatives"] = alt
    if extras:
        out.update(extras)
    if not args.no_cpu_baseline and n_gpus == 1:
        import orc
        orc.build()
        threads = host_threads()
        out["cpu_baseline"] = cpu_baseline_block(orc, args.workload, tape, dims, W, H, threads, target_s=8.0)
        try:
            out["host_sampled_path"] = host_sampled_path(orc, S, threads)
        except Exception as e:  # an extra, never a reason to lose the bench line
            out["host_sampled_path"] = {"error": str(e)}
    print(json.dumps(out))
    sv.close()
    if dist: dist.destroy_process_group()


if __name__ == "__main__":
    main()

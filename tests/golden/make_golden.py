"""Generates the golden fixtures in this directory from the CPU oracle (oracle/sdf_oracle.cpp).

The reference holds NO value-level vectors for this path and cannot be
built or run here (no cargo / wasm runtime / GL), so these vectors come from the oracle -- a
restatement, "parity unpinned" -- and serve to (a) freeze the oracle against regressions and
(b) give the GPU tests a target that does not depend on compiling the oracle on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import sdf_viewer_b200 as S  # noqa: E402

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def main():
    orc.build()
    rng = np.random.default_rng(20261017)
    # 1. point samples of SDFDemo::sample (demo/mod.rs:51-75): random + near-surface + axis points
    pts = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    shell = rng.normal(size=(600, 3)); shell = (shell / np.linalg.norm(shell, axis=1, keepdims=True) * rng.uniform(1.0, 1.1, (600, 1))).astype(np.float32)
    faces = rng.uniform(-1, 1, (400, 3)).astype(np.float32); faces[np.arange(400), rng.integers(0, 3, 400)] = rng.choice([-1, 1], 400) * rng.uniform(0.9, 1.0, 400).astype(np.float32)
    special = np.array([[1, 1, 1], [-1, -1, -1], [1, 0, 0], [0, 0, 0], [0.95, 0.95, 0.95], [0, 1.05, 0], [-0.0, 0.5, 1.0],
                        [0.015873075] * 3, [-0.96825397, -0.96825397, -0.96825397]], np.float32)
    pts = np.concatenate([pts, shell, faces, special]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "demo_samples.npz"), points=pts, samples=orc.demo_sample(pts),
                        samples_distance_only=orc.demo_sample(pts, distance_only=True))
    # 2. a fully loaded 16^3 demo volume (2 passes) and the iteration count
    v = orc.Viewer(BB, (16, 16, 16), 2)
    its = v.update(orc.Sampler(tape=S.tape.demo_tape()))
    np.savez_compressed(os.path.join(HERE, "demo_volume_16.npz"), tex0=v.tex0.copy(), tex1=v.tex1.copy(), iterations=its)
    # 3. CSG tape samples (40 primitives, seed 11)
    table = S.tape.csg_primitive_table(40, seed=11)
    tape = S.tape.csg_tape(table)
    cp = rng.uniform(-1, 1, (2048, 3)).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "csg_samples.npz"), table=table, points=cp, samples=orc.tape_sample(tape, cp))
    # 4. a traced frame: 32^3 demo volume, 2 passes committed (lod 1, LINEAR), default camera, 160x120
    v = orc.Viewer(BB, (32, 32, 32), 2)
    v.update(orc.Sampler(tape=S.tape.demo_tape()))
    w, h = 160, 120
    cam = S.default_camera(w, h)
    rays = S.camera_rays(cam, w, h)
    P = orc.trace_params(rays, BB, (32, 32, 32), lod=1.0, filter_linear=1)
    rgba, depth, gbuf = orc.trace(P, v.tex0, v.tex1, w, h)
    np.savez_compressed(os.path.join(HERE, "trace_32_160x120.npz"), rgba=rgba, depth=depth, gbuf=gbuf,
                        origin=np.array(rays.origin), base=np.array(rays.base), dx=np.array(rays.dx),
                        dy=np.array(rays.dy), bvp=np.array(rays.bvp))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

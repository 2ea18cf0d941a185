"""Render the demo SDF (and the CSG-1k workload) through the GPU path and write PNGs (no imaging library:
zlib + struct).  usage: python tools/render_png.py [out_dir] [grid_side]"""
import os, struct, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdf_viewer_b200 as S

BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))


def write_png(path, rgba8_bottom_up):
    img = np.ascontiguousarray(rgba8_bottom_up[::-1])  # row 0 of the frame is the bottom row (GL convention)
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    os.makedirs(out, exist_ok=True)
    w, h = 1280, 720
    for name, tape in (("demo", S.tape.demo_tape()), ("csg1k", S.tape.csg_tape())):
        with S.SDFViewer.from_bb(BB, side, 2) as v:
            v.set_tape(tape); v.fill_all(); v.commit()
            for cname, cam in (("default", S.default_camera(w, h)), ("closeup", S.look_at_camera((1.2, 1.4, 2.2), (0, 0, 0), w, h))):
                rgba8, depth = v.trace_rgba8(cam, w, h)
                bg = np.array([24, 24, 28, 255], np.uint8)  # composite over a dark background like the viewer's
                a = rgba8[..., 3:4].astype(np.float32) / 255.0
                comp = (rgba8.astype(np.float32) * a + bg * (1 - a)).astype(np.uint8)
                comp[..., 3] = 255
                path = os.path.join(out, f"frame_{name}_{side}_{cname}.png")
                write_png(path, comp)
                print(path, f"hit fraction {float((depth < 1).mean()):.3f}")


if __name__ == "__main__":
    main()

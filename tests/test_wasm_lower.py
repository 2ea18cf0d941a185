"""sdfgpu_wasm_lower: the `sample` export of a WebAssembly SDF (guest ABI: /root/reference/src/sdf/wasm/mod.rs:5-37,
host src/sdf/wasm/native.rs:29-98,163-217) lowered to a scalar-program tape.

No WASM toolchain exists in the build image, so the guests are assembled here (tests/wasm_asm.py) in the shapes a
compiler emits: results in a static buffer or on a shadow stack copied out with memory.copy, parameters loaded from
data segments, helper calls, call_indirect through a table (trait objects), loops with concrete trip counts,
br_if / if-else / early returns on values that depend on the position, integer work on truncated coordinates.
Each lowered tape is evaluated by the oracle's interpreter and compared bit for bit with the same formula written
in numpy float32; the specialiser's CUDA for it must compile (NVRTC, sm_100a).  Two of the guests also fill a grid on
the GPU (`-m gpu`, ran on a B200), as does the sweep over all of them."""
import os
import struct

import numpy as np
import pytest

from wasm_asm import F32, F64, I32, I64, Module

f32 = np.float32
OUT, BBP = 1024, 2048
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
SAMPLE_SIG = ([I32, F32, F32, F32, I32], [I32])
X, Y, Z = ("local.get", 1), ("local.get", 2), ("local.get", 3)


def base_module(bb=(-1, -1, -1, 1, 1, 1), pages=1):
    m = Module(pages=pages)
    m.data_at(BBP, struct.pack("<6f", *bb))
    m.func([I32], [I32], body=[("i32.const", BBP)], export="bounding_box")
    return m


def store_out(k, value_instrs, base=("i32.const", OUT)):
    return [base] + list(value_instrs) + [("f32.store", 4 * k)]


def points(n=300, seed=0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-1, 1, (n, 3)).astype(f32)
    p[:8] = [[0, 0, 0], [0.5, 0.5, 0.5], [-0.5, 0.5, 0], [1, -1, 1], [0.25, -0.125, 0.0], [-0.0, 0.0, 1.0], [0.3, 0.3, 0.3], [-1, -1, -1]]
    return p


def same(a, b):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def lowered(S, oracle, m, sdf_id=0):
    tape, bb, summary = S.wasm.lower(m.build(), sdf_id)
    assert S.jit_check(tape, 4)                       # the generated kernel compiles for sm_100a
    return tape, bb, summary


# ---------------------------------------------------------------------------------------------- guests

def guest_sphere_static():
    """Result in a static buffer; the radius lives in a data segment (a `static` in the guest)."""
    m = base_module()
    m.data_at(512, struct.pack("<f", 0.7))
    body = store_out(0, [X, X, "f32.mul", Y, Y, "f32.mul", "f32.add", Z, Z, "f32.mul", "f32.add", "f32.sqrt",
                         ("i32.const", 512), ("f32.load", 0), "f32.sub"])
    body += store_out(1, [X, "f32.abs"]) + store_out(2, [("f32.const", 0.25)]) + store_out(3, [Y, ("f32.const", 0.0), "f32.max"])
    body += store_out(4, [("f32.const", 0.0)]) + store_out(5, [("f32.const", 0.5)]) + store_out(6, [("f32.const", 1.0)])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    return m


def ref_sphere_static(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    o = np.zeros((len(p), 7), f32)
    o[:, 0] = np.sqrt((x * x + y * y) + z * z) - f32(0.7)
    o[:, 1] = np.abs(x)
    o[:, 2] = 0.25
    o[:, 3] = np.where(y > 0, y, f32(0.0)) + f32(0.0) * 0   # wasm max(y, +0): y for y > 0, +0 otherwise (incl. -0)
    o[:, 5], o[:, 6] = 0.5, 1.0
    return o


def guest_box_branchy():
    """max() written with br_if, a bump allocator in a global for the result, and an if / else with a result."""
    m = base_module()
    heap = m.global_(I32, 4096)
    D, P = ("local.get", 5), ("local.get", 6)
    body = [X, "f32.abs", ("local.set", 5)]
    for axis in (Y, Z):
        body += [("block", []), axis, "f32.abs", D, "f32.gt", "i32.eqz", ("br_if", 0), axis, "f32.abs", ("local.set", 5), "end"]
    body += [D, ("f32.const", 0.5), "f32.sub", ("local.set", 5)]
    body += [("global.get", heap), ("local.tee", 6), ("i32.const", 32), "i32.add", ("global.set", heap)]
    body += [P, D, ("f32.store", 0)]
    body += [P, D, ("f32.const", 0.1), "f32.gt", ("if", [F32]), ("f32.const", 0.0), "else",
             X, ("f32.const", 4.0), "f32.mul", "f32.floor", ("f32.const", 0.25), "f32.mul", "end", ("f32.store", 4)]
    for k, c in ((2, 0.2), (3, 0.3), (4, 0.0), (5, 0.9), (6, 1.0)):
        body += [P, ("f32.const", c), ("f32.store", 4 * k)]
    m.func(*SAMPLE_SIG, locals=[F32, I32], body=body + [P], export="sample")
    return m


def ref_box_branchy(p):
    a = np.abs(p)
    d = a[:, 0].copy()
    d = np.where(a[:, 1] > d, a[:, 1], d)
    d = np.where(a[:, 2] > d, a[:, 2], d)
    d = d - f32(0.5)
    o = np.zeros((len(p), 7), f32)
    o[:, 0] = d
    o[:, 1] = np.where(d > f32(0.1), f32(0.0), np.floor(p[:, 0] * f32(4.0)) * f32(0.25))
    o[:, 2], o[:, 3], o[:, 5], o[:, 6] = 0.2, 0.3, 0.9, 1.0
    return o


SPHERES = [(0.3, 0.2, -0.1, 0.35, 0.9), (-0.4, -0.3, 0.2, 0.3, 0.5), (0.0, 0.5, 0.5, 0.25, 0.2), (-0.2, 0.1, -0.6, 0.4, 0.7)]


def guest_csg_calls():
    """A union of four spheres read from a table in memory: a loop with a concrete trip count, a helper function,
    call_indirect through the function table (as a trait object's vtable), the result built on a shadow stack
    (global stack pointer) and copied to the output buffer with memory.copy; `init()` fills the table's count."""
    m = base_module()
    sp = m.global_(I32, 60000)
    TAB, COUNT = 4096, 4000
    m.data_at(TAB, b"".join(struct.pack("<5f", *s) for s in SPHERES))
    m.data_at(3000, struct.pack("<I", 1))  # which material function to call (index into the table)
    init = m.func([], [], body=[("i32.const", COUNT), ("i32.const", len(SPHERES)), ("i32.store", 0)], export="init")
    # dist(x, y, z, ptr) -> f32
    dist = m.func([F32, F32, F32, I32], [F32], locals=[F32, F32, F32], body=[
        ("local.get", 0), ("local.get", 3), ("f32.load", 0), "f32.sub", ("local.set", 4),
        ("local.get", 1), ("local.get", 3), ("f32.load", 4), "f32.sub", ("local.set", 5),
        ("local.get", 2), ("local.get", 3), ("f32.load", 8), "f32.sub", ("local.set", 6),
        ("local.get", 4), ("local.get", 4), "f32.mul", ("local.get", 5), ("local.get", 5), "f32.mul", "f32.add",
        ("local.get", 6), ("local.get", 6), "f32.mul", "f32.add", "f32.sqrt", ("local.get", 3), ("f32.load", 12), "f32.sub"])
    mat0 = m.func([F32], [F32], body=[("local.get", 0), ("f32.const", 0.5), "f32.mul"])
    mat1 = m.func([F32], [F32], body=[("local.get", 0), "f32.abs"])
    m.table([mat0, mat1])
    t_mat = m.type_index([F32], [F32])
    I, BEST, COL, DI, SP, PTR = (("local.get", k) for k in (5, 6, 7, 8, 9, 10))
    body = [("global.get", sp), ("i32.const", 32), "i32.sub", ("local.tee", 9), ("global.set", sp),
            ("f32.const", 1e9), ("local.set", 6), ("f32.const", 0.0), ("local.set", 7), ("i32.const", 0), ("local.set", 5),
            ("block", []), ("loop", []),
            I, ("i32.const", COUNT), ("i32.load", 0), "i32.ge_u", ("br_if", 1),
            ("i32.const", TAB), I, ("i32.const", 20), "i32.mul", "i32.add", ("local.set", 10),
            X, Y, Z, PTR, ("call", dist), ("local.set", 8),
            PTR, ("f32.load", 16), COL, DI, BEST, "f32.lt", "select", ("local.set", 7),
            DI, BEST, DI, BEST, "f32.lt", "select", ("local.set", 6),
            I, ("i32.const", 1), "i32.add", ("local.set", 5), ("br", 0), "end", "end",
            SP, BEST, ("f32.store", 0), SP, COL, ("f32.store", 4), SP, COL, COL, "f32.mul", ("f32.store", 8),
            SP, ("f32.const", 0.1), ("f32.store", 12),
            SP, BEST, ("i32.const", 3000), ("i32.load", 0), ("call_indirect", t_mat), ("f32.store", 16),
            SP, ("f32.const", 0.6), ("f32.store", 20), SP, ("f32.const", 1.0), ("f32.store", 24),
            ("i32.const", OUT), SP, ("i32.const", 28), ("memory.copy",),
            SP, ("i32.const", 32), "i32.add", ("global.set", sp), ("i32.const", OUT)]
    m.func(*SAMPLE_SIG, locals=[I32, F32, F32, F32, I32, I32], body=body, export="sample")
    assert init is not None
    return m


def ref_csg_calls(p):
    best = np.full(len(p), f32(1e9))
    col = np.zeros(len(p), f32)
    for cx, cy, cz, r, c in SPHERES:
        qx, qy, qz = p[:, 0] - f32(cx), p[:, 1] - f32(cy), p[:, 2] - f32(cz)
        d = np.sqrt((qx * qx + qy * qy) + qz * qz) - f32(r)
        col = np.where(d < best, f32(c), col)
        best = np.where(d < best, d, best)
    o = np.zeros((len(p), 7), f32)
    o[:, 0], o[:, 1], o[:, 2], o[:, 3], o[:, 4], o[:, 5], o[:, 6] = best, col, col * col, 0.1, np.abs(best), 0.6, 1.0
    return o


def guest_early_returns():
    """Nested branches on the position with early `return`s and a br out of two blocks carrying a value."""
    m = base_module()
    A, B = 1100, 1200
    m.data_at(A, struct.pack("<7f", 0.5, 1, 0, 0, 0.1, 0.2, 1.0))
    m.data_at(B, struct.pack("<7f", -0.25, 0, 1, 0, 0.3, 0.4, 0.5))
    body = [X, ("f32.const", 0.0), "f32.gt",
            ("if", []),
            Y, ("f32.const", 0.0), "f32.gt", ("if", []), ("i32.const", A), "return", "end",
            Z, ("f32.const", 0.5), "f32.lt", ("if", []), ("i32.const", B), "return", "end",
            "end",
            # d = block(result f32) { block { br_if 0 (x+y < 0); br 1 (x*y) }; -(x+y) }
            ("block", [F32]), ("block", []), X, Y, "f32.add", ("f32.const", 0.0), "f32.lt", ("br_if", 0), X, Y, "f32.mul", ("br", 1), "end",
            X, Y, "f32.add", "f32.neg", "end", ("local.set", 5)]
    body += store_out(0, [("local.get", 5)]) + store_out(1, [Z, "f32.nearest"]) + store_out(2, [X, Y, "f32.copysign"])
    body += store_out(3, [X, Y, "f32.min"]) + store_out(4, [X, "f32.ceil"]) + store_out(5, [Y, "f32.trunc"]) + store_out(6, [X, Y, "f32.div"])
    m.func(*SAMPLE_SIG, locals=[F32], body=body + [("i32.const", OUT)], export="sample")
    return m


def ref_early_returns(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    with np.errstate(all="ignore"):
        general = np.zeros((len(p), 7), f32)
        general[:, 0] = np.where(x + y < 0, -(x + y), x * y)
        general[:, 1] = np.rint(z)
        general[:, 2] = np.copysign(x, y)
        mn = np.where(x == y, (x.view(np.uint32) | y.view(np.uint32)).view(f32), np.minimum(x, y))
        general[:, 3] = mn
        general[:, 4] = np.ceil(x)
        general[:, 5] = np.trunc(y)
        general[:, 6] = x / y
    a = np.array([0.5, 1, 0, 0, 0.1, 0.2, 1.0], f32)
    b = np.array([-0.25, 0, 1, 0, 0.3, 0.4, 0.5], f32)
    out = general
    out = np.where(((x > 0) & ~(y > 0) & (z < f32(0.5)))[:, None], b[None, :], out)
    out = np.where(((x > 0) & (y > 0))[:, None], a[None, :], out)
    return out.astype(f32)


def guest_integer_checker():
    """Integer arithmetic on truncated coordinates (a checker pattern), shifts and conversions back to float."""
    m = base_module()
    cell = lambda axis: [axis, ("f32.const", 8.0), "f32.mul", "f32.floor", ("i32.trunc_sat_f32_s",)]  # noqa: E731
    body = cell(X) + cell(Y) + ["i32.add"] + cell(Z) + ["i32.add", ("local.tee", 5), ("i32.const", 1), "i32.and", ("local.set", 6)]
    body += store_out(0, [Y])
    body += store_out(1, [("f32.const", 0.9), ("f32.const", 0.1), ("local.get", 6), "select"])
    body += store_out(2, [("local.get", 5), "f32.convert_i32_s", ("f32.const", 0.0625), "f32.mul"])
    body += store_out(3, [("local.get", 5), ("i32.const", 2), "i32.shl", ("i32.const", 3), "i32.shr_s", ("i32.const", 7), "i32.xor", "f32.convert_i32_s"])
    body += store_out(4, [X, ("i32.trunc_f32_s",) if False else "i32.trunc_f32_s", "f32.convert_i32_s"])
    body += store_out(5, [X, "i32.reinterpret_f32", ("i32.const", 0x7FFFFFFF), "i32.and", "f32.reinterpret_i32"])
    body += store_out(6, [("local.get", 5), ("i32.const", 0), "i32.lt_s", ("if", [F32]), ("f32.const", 0.25), "else", ("f32.const", 1.0), "end"])
    m.func(*SAMPLE_SIG, locals=[I32, I32], body=body + [("i32.const", OUT)], export="sample")
    return m


def ref_integer_checker(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    c = (np.floor(x * f32(8)).astype(np.int32) + np.floor(y * f32(8)).astype(np.int32) + np.floor(z * f32(8)).astype(np.int32))
    o = np.zeros((len(p), 7), f32)
    o[:, 0] = y
    o[:, 1] = np.where((c & 1) != 0, f32(0.9), f32(0.1))
    o[:, 2] = c.astype(f32) * f32(0.0625)
    o[:, 3] = (((c << 2) >> 3) ^ 7).astype(f32)
    o[:, 4] = np.trunc(x).astype(np.int32).astype(f32)
    o[:, 5] = np.abs(x)
    o[:, 6] = np.where(c < 0, f32(0.25), f32(1.0))
    return o


def guest_registry():
    """What a Rust guest does before any arithmetic: look the sdf_id up in a registry (an i64 hash of the id picks a
    bucket in memory), dispatch on the entry's kind with br_table, grow the memory for the result on first use,
    narrow stores / sign-extending loads on the way.  All of it is concrete and must leave no trace in the tape."""
    m = base_module()
    BUCKETS = 8192
    # hash(id) = ((id * 0x9E3779B97F4A7C15) rotl 17) >> 61  -> bucket 0..7; bucket for id 0 holds kind 2, radius 0.6
    def hash_bucket(i):
        h = (i * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        h = ((h << 17) | (h >> 47)) & (2 ** 64 - 1)
        return h >> 61
    b0 = hash_bucket(0)
    m.data_at(BUCKETS + 8 * b0, struct.pack("<hbbf", -2, 2, 0, 0.6))   # i16 tag -2, i8 kind 2, pad, f32 radius
    ENTRY, PTR, KIND = ("local.get", 5), ("local.get", 6), ("local.get", 7)
    body = [("local.get", 0), "i64.extend_i32_u", ("i64.const", 0x9E3779B97F4A7C15), "i64.mul", ("i64.const", 17), "i64.rotl",
            ("i64.const", 61), "i64.shr_u", "i32.wrap_i64", ("i32.const", 8), "i32.mul", ("i32.const", BUCKETS), "i32.add", ("local.set", 5),
            # the tag must be -2 (sign-extending 16-bit load), else trap
            ENTRY, ("i32.load16_s", 0), ("i32.const", -2 & 0xFFFFFFFF), "i32.ne", ("if", []), "unreachable", "end",
            ENTRY, ("i32.load8_u", 2), ("local.set", 7),
            # result buffer: one fresh page
            ("i32.const", 1), ("memory.grow",), ("i32.const", 16), "i32.shl", ("local.set", 6),
            PTR, ("i32.const", 0x55), ("i32.store8", 3), PTR, ("i32.const", 0x1234), ("i32.store16", 0),   # scribbles overwritten below
            ("block", []), ("block", []), ("block", []), KIND, ("br_table", [0, 1, 2], 0), "end",
            PTR, ("f32.const", 111.0), ("f32.store", 0), ("br", 1), "end",
            PTR, ("f32.const", 222.0), ("f32.store", 0), ("br", 0), "end"]
    # kind 2 lands here with nothing stored yet: the sphere
    body += [PTR, X, X, "f32.mul", Y, Y, "f32.mul", "f32.add", Z, Z, "f32.mul", "f32.add", "f32.sqrt", ENTRY, ("f32.load", 4), "f32.sub",
             ("f32.store", 0)]
    for k in range(1, 7):
        body += [PTR, ("f32.const", 0.1 * k), ("f32.store", 4 * k)]
    m.func(*SAMPLE_SIG, locals=[I32, I32, I32], body=body + [PTR], export="sample")
    return m


def ref_registry(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    o = np.zeros((len(p), 7), f32)
    o[:, 0] = np.sqrt((x * x + y * y) + z * z) - f32(0.6)
    for k in range(1, 7):
        o[:, k] = f32(0.1 * k)
    return o


def add_musl_fmodf(m):
    """compiler-builtins / musl `fmodf` written out in WebAssembly: what a Rust guest's `a % b` on f32 calls.  Its main
    loop runs (exponent of x - exponent of y) times, so it cannot be unrolled for a symbolic x."""
    G = lambda k: ("local.get", k)      # noqa: E731   locals: 0 x, 1 y, 2 uxi, 3 uyi, 4 ex, 5 ey, 6 sx, 7 i
    S_ = lambda k: ("local.set", k)     # noqa: E731
    C = lambda v: ("i32.const", v)      # noqa: E731
    zero_times_x = [("f32.const", 0.0), G(0), "f32.mul", "return"]

    def normalize(u, e):
        return [G(e), "i32.eqz", ("if", []),
                G(u), C(9), "i32.shl", S_(7),
                ("block", []), ("loop", []), G(7), C(31), "i32.shr_u", ("br_if", 1),
                G(e), C(1), "i32.sub", S_(e), G(7), C(1), "i32.shl", S_(7), ("br", 0), "end", "end",
                G(u), C(1), G(e), "i32.sub", "i32.shl", S_(u),
                "else", G(u), C(0x007FFFFF), "i32.and", C(0x00800000), "i32.or", S_(u), "end"]

    def subtract_step():
        return [G(2), G(3), "i32.sub", S_(7), G(7), C(31), "i32.shr_u", "i32.eqz", ("if", []),
                G(7), "i32.eqz", ("if", [])] + zero_times_x + ["end", G(7), S_(2), "end"]

    body = [G(0), "i32.reinterpret_f32", S_(2), G(1), "i32.reinterpret_f32", S_(3),
            G(2), C(23), "i32.shr_u", C(255), "i32.and", S_(4), G(3), C(23), "i32.shr_u", C(255), "i32.and", S_(5),
            G(2), C(0x80000000), "i32.and", S_(6),
            G(3), C(1), "i32.shl", "i32.eqz", G(1), G(1), "f32.ne", "i32.or", G(4), C(255), "i32.eq", "i32.or", ("if", []),
            G(0), G(1), "f32.mul", G(0), G(1), "f32.mul", "f32.div", "return", "end",
            G(2), C(1), "i32.shl", G(3), C(1), "i32.shl", "i32.le_u", ("if", []),
            G(2), C(1), "i32.shl", G(3), C(1), "i32.shl", "i32.eq", ("if", [])] + zero_times_x + ["end", G(0), "return", "end"]
    body += normalize(2, 4) + normalize(3, 5)
    body += [("block", []), ("loop", []), G(4), G(5), "i32.le_s", ("br_if", 1)] + subtract_step() + \
            [G(2), C(1), "i32.shl", S_(2), G(4), C(1), "i32.sub", S_(4), ("br", 0), "end", "end"]
    body += subtract_step()
    body += [("block", []), ("loop", []), G(2), C(23), "i32.shr_u", ("br_if", 1),
             G(2), C(1), "i32.shl", S_(2), G(4), C(1), "i32.sub", S_(4), ("br", 0), "end", "end"]
    body += [G(4), C(0), "i32.gt_s", ("if", []),
             G(2), C(0x00800000), "i32.sub", G(4), C(23), "i32.shl", "i32.or", S_(2),
             "else", G(2), C(1), G(4), "i32.sub", "i32.shr_u", S_(2), "end",
             G(2), G(6), "i32.or", "f32.reinterpret_i32"]
    return m.func([F32, F32], [F32], locals=[I32] * 6, body=body)


def guest_brick_modulo():
    """`%` on floats, as the reference's brick texture uses it (src/sdf/demo/cube.rs:192), through the guest's own fmodf."""
    m = base_module()
    fmod = add_musl_fmodf(m)
    body = store_out(0, [X, ("f32.const", 0.3), "f32.add", "f32.abs", ("f32.const", 3.0), "f32.mul", ("f32.const", 0.5), ("call", fmod)])
    body += store_out(1, [Y, ("f32.const", 0.25), ("call", fmod)])
    body += store_out(2, [("f32.const", 5.625), ("f32.const", 0.5), ("call", fmod)])          # concrete arguments: simply executed
    body += store_out(3, [Z, ("f32.const", 100.0), "f32.mul", ("f32.const", 0.7), ("call", fmod)])
    body += store_out(4, [X, Y, ("call", fmod)])
    body += store_out(5, [Z, ("f32.const", 0.0), ("call", fmod)]) + store_out(6, [("f32.const", 1.0)])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    return m


def ref_brick_modulo(p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    o = np.zeros((len(p), 7), f32)
    with np.errstate(all="ignore"):
        o[:, 0] = np.fmod(np.abs(x + f32(0.3)) * f32(3.0), f32(0.5))
        o[:, 1] = np.fmod(y, f32(0.25))
        o[:, 2] = np.fmod(f32(5.625), f32(0.5))
        o[:, 3] = np.fmod(z * f32(100.0), f32(0.7))
        o[:, 4] = np.fmod(x, y)
        o[:, 5] = np.fmod(z, f32(0.0))
    o[:, 6] = 1.0
    return o


def guest_reference_demo(half=0.95, radius=1.05, seam=0.05):
    """The reference's own SDFDemo (src/sdf/demo/mod.rs:51-75: brick cube minus sphere with a seam material; cube.rs:79-89,
    164-222; sphere.rs:37-47,122-124) hand-compiled to WebAssembly the way the Rust source is structured: one function
    per `sample`, SDFSample structs in guest memory, the air early-outs, the tri-planar brick texture with `%` (through
    the guest's fmodf) and `floor`, `normalize`, the struct-valued `if` of the combinator done with memory.copy."""
    m = base_module()
    fmod = add_musl_fmodf(m)
    PAR, BOXS, SPHS = 256, 1200, 1300
    m.data_at(PAR, struct.pack("<3f", half, radius, seam))
    G = lambda k: ("local.get", k)      # noqa: E731
    St = lambda k: ("local.set", k)     # noqa: E731
    K = lambda v: ("f32.const", v)      # noqa: E731

    def store_fields(ptr, values):  # fields 1..6 of the SDFSample at `ptr` (instruction lists that leave an f32)
        out = []
        for k, v in enumerate(values, start=1):
            out += [ptr] + v + [("f32.store", 4 * k)]
        return out

    # brick_texture(px, py, pz, nx, ny, nz, dst)   locals: 7 u, 8 v, 9 bx, 10 by, 11 max_cement
    colour = lambda a, b, c: [[K(a), K(255.0), "f32.div"], [K(b), K(255.0), "f32.div"], [K(c), K(255.0), "f32.div"]]  # noqa: E731
    brick = m.func([F32] * 6 + [I32], [], locals=[F32] * 5, body=[
        G(3), "f32.abs", G(4), "f32.abs", "f32.gt", ("if", []),
        G(3), "f32.abs", G(5), "f32.abs", "f32.gt", ("if", []), G(2), St(7), G(1), St(8), "else", G(0), St(7), G(1), St(8), "end",
        "else",
        G(4), "f32.abs", G(5), "f32.abs", "f32.gt", ("if", []), G(2), St(7), G(0), St(8), "else", G(0), St(7), G(1), St(8), "end",
        "end",
        G(7), G(8), K(0.25), "f32.div", "f32.floor", K(4.0), "f32.div", "f32.add", "f32.abs", K(0.5), ("call", fmod), St(9),
        G(8), "f32.abs", K(0.25), ("call", fmod), St(10),
        K(0.2), K(2.0), "f32.div", K(0.25), "f32.mul", St(11),
        G(9), G(11), "f32.lt", G(9), K(0.5), G(11), "f32.sub", "f32.gt", "i32.or",
        G(10), G(11), "f32.lt", "i32.or", G(10), K(0.25), G(11), "f32.sub", "f32.gt", "i32.or",
        ("if", [])] + store_fields(G(6), colour(56.0, 70.0, 60.0) + [[K(0.4)], [K(0.5)], [K(1.0)]]) + ["else"] +
        store_fields(G(6), colour(150.0, 24.0, 10.0) + [[K(0.2)], [K(0.8)], [K(0.0)]]) + ["end"])
    zeros = [[K(0.0)]] * 6
    HALF = [("i32.const", PAR), ("f32.load", 0)]
    # cube_sample(x, y, z, dst)   locals: 4 d, 5 nx, 6 ny, 7 nz
    normal = lambda a, dst: [K(1.0), G(a), "f32.copysign", K(0.0), G(a), "f32.abs"] + HALF + ["f32.gt", "select", St(dst)]  # noqa: E731
    cube = m.func([F32] * 3 + [I32], [], locals=[F32] * 4, body=[
        G(0), "f32.abs", G(1), "f32.abs", "f32.max", G(2), "f32.abs", "f32.max"] + HALF + ["f32.sub", St(4),
        G(3), G(4), ("f32.store", 0),
        G(4), K(0.1), "f32.gt", ("if", [])] + store_fields(G(3), zeros) + ["else"] +
        normal(0, 5) + normal(1, 6) + normal(2, 7) + [G(0), G(1), G(2), G(5), G(6), G(7), G(3), ("call", brick), "end"])
    # sphere_sample(x, y, z, dst)   locals: 4 len, 5 d, 6 inv
    sphere = m.func([F32] * 3 + [I32], [], locals=[F32] * 3, body=[
        G(0), G(0), "f32.mul", G(1), G(1), "f32.mul", "f32.add", G(2), G(2), "f32.mul", "f32.add", "f32.sqrt", St(4),
        G(4), ("i32.const", PAR), ("f32.load", 4), "f32.sub", St(5), G(3), G(5), ("f32.store", 0),
        G(5), K(0.1), "f32.gt", ("if", [])] + store_fields(G(3), zeros) + ["else",
        K(1.0), G(4), "f32.div", St(6)] +
        store_fields(G(3), [[G(0), G(6), "f32.mul", "f32.abs"], [G(1), G(6), "f32.mul", "f32.abs"], [G(2), G(6), "f32.mul", "f32.abs"],
                            [K(0.0)], [K(0.0)], [K(0.0)]]) + ["end"])
    # sample   locals: 5 dist, 6 inter
    BD = [("i32.const", BOXS), ("f32.load", 0)]
    SD = [("i32.const", SPHS), ("f32.load", 0)]
    body = [X, Y, Z, ("i32.const", BOXS), ("call", cube), X, Y, Z, ("i32.const", SPHS), ("call", sphere)]
    body += BD + SD + ["f32.neg", "f32.max", St(5)] + BD + ["f32.abs"] + SD + ["f32.abs", "f32.sub", St(6)]
    body += [G(6), K(0.0), "f32.lt", ("if", []), ("i32.const", OUT), ("i32.const", BOXS), ("i32.const", 28), ("memory.copy",),
             "else", ("i32.const", OUT), ("i32.const", SPHS), ("i32.const", 28), ("memory.copy",), "end"]
    body += [G(6), "f32.abs", ("i32.const", PAR), ("f32.load", 8), "f32.le", ("if", [])] + \
        store_fields(("i32.const", OUT), [[K(0.5)], [K(0.6)], [K(0.7)], [K(0.5)], [K(0.0)], [K(0.0)]]) + ["end"]
    body += [("i32.const", OUT), G(5), ("f32.store", 0), ("i32.const", OUT)]
    m.func(*SAMPLE_SIG, locals=[F32, F32], body=body, export="sample")
    return m


GUESTS = {
    "sphere_static": (guest_sphere_static, ref_sphere_static),
    "box_branchy": (guest_box_branchy, ref_box_branchy),
    "csg_calls": (guest_csg_calls, ref_csg_calls),
    "early_returns": (guest_early_returns, ref_early_returns),
    "integer_checker": (guest_integer_checker, ref_integer_checker),
    "registry": (guest_registry, ref_registry),
    "brick_modulo": (guest_brick_modulo, ref_brick_modulo),
}
# guest_reference_demo is checked against the oracle's SDFDemo (test_the_reference_demo_as_a_wasm_guest), not a formula


@pytest.mark.parametrize("name", list(GUESTS))
def test_lowered_guest_equals_its_formula(S, oracle, name):
    make, ref = GUESTS[name]
    tape, bb, summary = lowered(S, oracle, make())
    assert bb == BB and summary.startswith("lowered:")
    p = points()
    got = oracle.tape_sample(tape, p)
    want = ref(p)
    bad = ~((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want)))
    assert not bad.any(), (name, np.argwhere(bad)[:5], got[bad][:5], want[bad][:5])


def test_fmodf_is_recognised_by_what_it_computes(S, oracle):
    """The guest's fmodf (an exponent loop that cannot be unrolled for a symbolic operand) is identified by probing
    it with concrete arguments and becomes one FMOD op per call; a function that differs from fmodf anywhere on the
    probe grid is not taken for it (it is inlined, and here fails because its loop depends on the position)."""
    tape, _, summary = lowered(S, oracle, guest_brick_modulo())
    assert "5 fmodf calls recognised" in summary and "0 symbolic branches" in summary
    # the same function with one constant changed (mantissa mask) is no fmodf
    m = base_module()
    impostor = add_musl_fmodf(m)
    t, locs, body = m.funcs[impostor - len(m.imports)]
    body[body.index(("i32.const", 0x007FFFFF))] = ("i32.const", 0x007FFFFE)
    m.func(*SAMPLE_SIG, body=store_out(0, [X, ("f32.const", 0.5), ("call", impostor)]) + [("i32.const", OUT)], export="sample")
    with pytest.raises(S.WasmLoweringError) as e:
        S.wasm.lower(m.build())
    assert e.value.code == -3 and "depend on the position" in str(e.value)


@pytest.mark.parametrize("effect", ["store", "global", "stack_frame"])
def test_fmodf_with_a_side_effect_is_not_replaced(S, oracle, effect):
    """A callee that returns fmodf on every probe but also leaves something behind -- a store to guest memory, a changed
    global -- is NOT replaced by the one-op fmod (the side effect would be lost); writing into its own shadow-stack frame
    (below the caller's stack pointer, global 0) is what compiled code does and is accepted."""
    m = base_module()
    sp = m.global_(I32, 60000)                    # __stack_pointer
    counter = m.global_(I32, 0)
    f = add_musl_fmodf(m)
    t, locs, body = m.funcs[f - len(m.imports)]
    if effect == "store":
        body[0:0] = [("i32.const", 512), ("i32.const", 512), ("i32.load", 0), ("i32.const", 1), "i32.add", ("i32.store", 0)]
    elif effect == "global":
        body[0:0] = [("global.get", counter), ("i32.const", 1), "i32.add", ("global.set", counter)]
    else:
        body[0:0] = [("global.get", sp), ("i32.const", 16), "i32.sub", ("local.get", 0), ("f32.store", 0)]  # a spill
    m.func(*SAMPLE_SIG, body=store_out(0, [X, ("f32.const", 0.5), ("call", f)]) + [("i32.const", OUT)], export="sample")
    if effect == "stack_frame":
        tape, _, summary = S.wasm.lower(m.build())
        assert "1 fmodf calls recognised" in summary
        p = points(200, seed=5)
        assert same(oracle.tape_sample(tape, p)[:, 0], np.fmod(p[:, 0], f32(0.5)))
    else:
        with pytest.raises(S.WasmLoweringError) as e:
            S.wasm.lower(m.build())
        assert e.value.code == -3


def test_the_reference_demo_as_a_wasm_guest(S, oracle):
    """SDFDemo hand-compiled to WebAssembly lowers to a scalar program that reproduces the oracle's direct restatement
    of `SDFDemo::sample` bit for bit -- on random points, on the voxel positions of a grid (where the seams, brick joints
    and air early-outs are hit exactly) and with other parameters through a live-memory re-lowering."""
    m = guest_reference_demo()
    wasm = m.build()
    tape, bb, summary = lowered(S, oracle, m)
    assert bb == BB and "fmodf calls recognised" in summary
    v = oracle.Viewer(BB, (24, 24, 24), 1)
    grid = np.array([v.voxel_pos(x, y, z) for z in range(24) for y in range(24) for x in range(24)], f32)
    pts = np.concatenate([points(4000, seed=3), grid, [[1.0, 0.0, 0.0], [0.95, 0.95, 0.95], [0.0, 0.0, 0.0], [-1.0, 1.0, -1.0]]]).astype(f32)
    got, want = oracle.tape_sample(tape, pts), oracle.demo_sample(pts)
    bad = ~((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want)))
    assert not bad.any(), (summary, np.argwhere(bad)[:5], pts[np.argwhere(bad)[:3, 0]], got[bad][:5], want[bad][:5])
    # every branch of the demo is exercised by those points
    assert (want[:, 4] == f32(0.4)).any() and (want[:, 4] == f32(0.2)).any() and (want[:, 4] == f32(0.5)).any() and (want[:, 1] == 0).any()
    # set_parameter on the live guest (radius 1.05 -> 0.9, cube 0.95 -> 0.85): same structure, new constants
    mem = bytearray(65536)
    for off, payload in m.data:
        mem[off:off + len(payload)] = payload
    mem[256:268] = struct.pack("<3f", 0.85, 0.9, 0.05)
    tape2 = S.wasm.lower(wasm, memory=bytes(mem))[0]
    assert len(tape2) == len(tape)
    want2 = oracle.demo_sample(pts, oracle.demo_params(cube_half_side=0.85, sphere_radius=0.9))
    assert same(oracle.tape_sample(tape2, pts), want2)
    # (equal constants are shared, so a parameter that lands exactly on another constant of the program -- a cube of
    # 0.8, the brick's roughness -- gives a program one constant shorter: still correct, compiled once more)
    mem[256:260] = struct.pack("<f", 0.8)
    tape3 = S.wasm.lower(wasm, memory=bytes(mem))[0]
    assert len(tape3) <= len(tape)
    assert same(oracle.tape_sample(tape3, pts), oracle.demo_sample(pts, oracle.demo_params(cube_half_side=0.8, sphere_radius=0.9)))


def test_parameter_change_keeps_the_structure(S, oracle):
    """Re-lowering after a guest parameter changed (here: the radius in the data segment) yields the same
    program with other constants -- the compiled kernel is re-used, as for the hand-written demo tape."""
    a = guest_sphere_static()
    b = guest_sphere_static()
    b.data[1] = (512, struct.pack("<f", 0.55))
    ta, tb = S.wasm.lower(a.build())[0], S.wasm.lower(b.build())[0]
    assert len(ta) == len(tb) and ta != tb
    n_consts = struct.unpack_from("<I", ta, 16)[0]
    off = 32 + 2 * 16
    assert ta[:off] == tb[:off] and ta[off + 4 * n_consts:] == tb[off + 4 * n_consts:]      # only constants differ
    p = points(50)
    assert same(oracle.tape_sample(tb, p)[:, 0], np.sqrt((p[:, 0] ** 2 + p[:, 1] ** 2) + p[:, 2] ** 2).astype(f32) - f32(0.55))


def test_lowering_a_live_instance(S, oracle):
    """sdfgpu_wasm_lower_live takes the linear memory of a running instance, so a parameter the host changed
    through the guest's set_parameter (which lands in guest memory) shows up in the tape.  The "live" memory is
    built here by applying the module's data segments and then poking the radius, as set_parameter would."""
    m = guest_sphere_static()
    wasm = m.build()
    mem = bytearray(65536)
    for off, payload in m.data:
        mem[off:off + len(payload)] = payload
    fresh = S.wasm.lower(wasm)[0]
    assert S.wasm.lower(wasm, memory=bytes(mem))[0] == fresh          # an untouched instance lowers like the module
    mem[512:516] = struct.pack("<f", 0.45)                            # set_parameter(radius = 0.45)
    tape, bb, _ = S.wasm.lower(wasm, memory=bytes(mem))
    p = points(60)
    assert bb == BB and len(tape) == len(fresh)
    assert same(oracle.tape_sample(tape, p)[:, 0], np.sqrt((p[:, 0] ** 2 + p[:, 1] ** 2) + p[:, 2] ** 2).astype(f32) - f32(0.45))
    # csg guest: init() already ran in the live instance (the sphere count is in memory), it must not be needed again
    g = guest_csg_calls()
    mem = bytearray(65536)
    for off, payload in g.data:
        mem[off:off + len(payload)] = payload
    mem[4000:4004] = struct.pack("<I", 2)                             # a live instance that holds only two spheres
    tape2 = S.wasm.lower(g.build(), memory=bytes(mem))[0]
    best = np.minimum(*[np.sqrt(((p[:, 0] - f32(cx)) ** 2 + (p[:, 1] - f32(cy)) ** 2) + (p[:, 2] - f32(cz)) ** 2).astype(f32) - f32(r)
                        for cx, cy, cz, r, _ in SPHERES[:2]])
    assert same(oracle.tape_sample(tape2, p)[:, 0], best)
    with pytest.raises(S.WasmLoweringError):
        S.wasm.lower(wasm, memory=b"\0" * 1000)                      # not whole pages


def test_wasm_sdf_surface(S):
    sdf = S.WasmSDF(guest_box_branchy().build())
    assert sdf.bounding_box() == BB and sdf.changed() is None
    assert sdf.tape()[:4] == b"SDFT"


def expect_failure(S, m, code, *fragments):
    with pytest.raises(S.WasmLoweringError) as e:
        S.wasm.lower(m.build() if isinstance(m, Module) else m)
    assert e.value.code == code, str(e.value)
    for f in fragments:
        assert f in str(e.value), str(e.value)


def test_what_cannot_be_lowered_says_why(S):
    # a loop whose exit depends on the position
    m = base_module()
    m.func(*SAMPLE_SIG, locals=[F32], body=[X, ("local.set", 5), ("block", []), ("loop", []), ("local.get", 5), ("f32.const", 1.0), "f32.ge", ("br_if", 1),
                                            ("local.get", 5), ("f32.const", 0.001), "f32.add", ("local.set", 5), ("br", 0), "end", "end",
                                            ("i32.const", OUT)], export="sample")
    expect_failure(S, m, -3, "depend on the position")
    # an address that depends on the position
    m = base_module()
    m.func(*SAMPLE_SIG, body=[("i32.const", OUT), X, ("f32.const", 4.0), "f32.mul", ("i32.trunc_sat_f32_s",), ("i32.const", 4), "i32.mul",
                              ("f32.load", 4096), ("f32.store", 0), ("i32.const", OUT)], export="sample")
    expect_failure(S, m, -3, "address that depends on the position")
    # a host import on the way
    m = Module()
    imp = m.import_func("env", "host_noise", [F32], [F32])
    m.data_at(BBP, struct.pack("<6f", -1, -1, -1, 1, 1, 1))
    m.func([I32], [I32], body=[("i32.const", BBP)], export="bounding_box")
    m.func(*SAMPLE_SIG, body=[("i32.const", OUT), X, ("call", imp), ("f32.store", 0), ("i32.const", OUT)], export="sample")
    expect_failure(S, m, -3, "env.host_noise")
    # ... but a host function that returns nothing (a logging hook) is skipped, in init() and in sample() alike
    m = Module()
    log = m.import_func("env", "log_f32", [F32], [])
    m.data_at(BBP, struct.pack("<6f", -1, -1, -1, 1, 1, 1))
    m.func([I32], [I32], body=[("i32.const", BBP)], export="bounding_box")
    m.func([], [], body=[("f32.const", 1.0), ("call", log)], export="init")
    body = [X, ("call", log)] + store_out(0, [X, Y, "f32.mul"])
    for k in range(1, 7):
        body += store_out(k, [("f32.const", 0.0)])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    tape = S.wasm.lower(m.build())[0]
    import orc
    p = points(10)
    assert same(orc.tape_sample(tape, p)[:, 0], p[:, 0] * p[:, 1])
    # WASI calls during set-up (a wasm32-wasi guest seeds its hasher with random_get) are answered with errno 0
    m = Module()
    rnd = m.import_func("wasi_snapshot_preview1", "random_get", [I32, I32], [I32])
    m.data_at(BBP, struct.pack("<6f", -1, -1, -1, 1, 1, 1))
    m.func([I32], [I32], body=[("i32.const", BBP)], export="bounding_box")
    # init: if random_get(5000, 16) != 0 { unreachable }; scale = 2.0 + f32(seed word, left at 0)
    m.func([], [], body=[("i32.const", 5000), ("i32.const", 16), ("call", rnd), ("if", []), "unreachable", "end",
                         ("i32.const", 5100), ("f32.const", 2.0), ("i32.const", 5000), ("i32.load", 0), "f32.convert_i32_u", "f32.add",
                         ("f32.store", 0)], export="init")
    body = store_out(0, [X, ("i32.const", 5100), ("f32.load", 0), "f32.mul"])
    for k in range(1, 7):
        body += store_out(k, [("f32.const", 0.0)])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    assert same(orc.tape_sample(S.wasm.lower(m.build())[0], p)[:, 0], p[:, 0] * f32(2.0))
    # f64 arithmetic on the position
    m = base_module()
    m.func(*SAMPLE_SIG, body=[("i32.const", OUT), X, "f64.promote_f32", "f64.sqrt", "f32.demote_f64", ("f32.store", 0), ("i32.const", OUT)], export="sample")
    expect_failure(S, m, -3, "no 32-bit scalar form")
    # a guest that traps for every position
    m = base_module()
    m.func(*SAMPLE_SIG, body=["unreachable"], export="sample")
    expect_failure(S, m, -3, "traps")
    # missing exports, wrong signature, garbage
    m = Module()
    m.func([I32], [I32], body=[("i32.const", BBP)], export="bounding_box")
    expect_failure(S, m, -1, "bounding_box and sample")
    m = base_module()
    m.func([I32, F32, F32, F32], [I32], body=[("i32.const", OUT)], export="sample")
    expect_failure(S, m, -1, "signatures")
    expect_failure(S, b"\0asm\x01\0\0\0\x01\xff\xff\xff\xff\x0f", -1)
    expect_failure(S, b"not wasm at all", -1, "WebAssembly")


def test_mutated_modules_never_crash_the_lowering(S, monkeypatch):
    """The module is untrusted input: random byte flips, truncations and splices of valid guests must end in
    SDFGPU_OK or an error code, never in a crash or a hang (the instruction budget bounds every run)."""
    monkeypatch.setenv("SDFGPU_WASM_BUDGET", "200000")
    rng = np.random.default_rng(2024)
    seeds = [make().build() for make, _ in GUESTS.values()]
    outcomes = {"ok": 0, "invalid": 0, "tape": 0}
    for it in range(1500):
        w = bytearray(seeds[it % len(seeds)])
        kind = it % 4
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                w[int(rng.integers(8, len(w)))] = int(rng.integers(0, 256))
        elif kind == 1:
            w = w[:int(rng.integers(8, len(w)))]
        elif kind == 2:
            i = int(rng.integers(8, len(w)))
            w[i] ^= 1 << int(rng.integers(0, 8))
        else:
            a, b = sorted(int(v) for v in rng.integers(8, len(w), 2))
            other = seeds[(it + 1) % len(seeds)]
            w = w[:a] + other[a:b] + w[b:]
        try:
            S.wasm.lower(bytes(w))
            outcomes["ok"] += 1
        except S.WasmLoweringError as e:
            outcomes["invalid" if e.code == -1 else "tape"] += 1
            assert e.code in (-1, -3) and str(e)
    assert outcomes["invalid"] > 100 and sum(outcomes.values()) == 1500


TRAP_SAMPLE = np.array([1.0, 0, 0, 0, 0, 0, 0], f32)  # SDFSample::new(1.0, zero), src/sdf/wasm/native.rs:196-203


def test_position_dependent_traps_give_the_reference_fallback_sample(S, oracle):
    """Where the guest would trap, the reference logs the error and uses SDFSample::new(1.0, zero) for that voxel
    (src/sdf/wasm/native.rs:196-203).  The lowering keeps a trap predicate (the branch conditions that lead to
    `unreachable`, zero divisors, truncations out of range) and selects that sample where it holds."""
    # (a) `if x < -0.25 { unreachable }` -- a bounds check the compiler left in
    m = base_module()
    body = [X, ("f32.const", -0.25), "f32.lt", ("if", []), "unreachable", "end"] + store_out(0, [X, Y, "f32.add"])
    for k in range(1, 7):
        body += store_out(k, [("f32.const", 0.125 * k)])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    tape, _, summary = lowered(S, oracle, m)
    p = points(200)
    got = oracle.tape_sample(tape, p)
    want = np.tile(np.array([0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75], f32), (len(p), 1))
    want[:, 0] = p[:, 0] + p[:, 1]
    trapped = p[:, 0] < f32(-0.25)
    assert 20 < trapped.sum() < 180
    want[trapped] = TRAP_SAMPLE
    assert same(got, want)
    assert "1 symbolic branches" in summary and "traps kept (1 branch sides, 0 div/trunc ops)" in summary

    # (b) a trap inside a nested branch, on the else side, with the other sides rejoining
    m = base_module()
    body = [X, ("f32.const", 0.0), "f32.gt", ("if", [F32]),
            Y, ("f32.const", 0.5), "f32.gt", ("if", [F32]), ("f32.const", 2.0), "else", "unreachable", "end",
            "else", ("f32.const", 3.0), "end", ("local.set", 5)]
    body += store_out(0, [("local.get", 5)])
    for k in range(1, 7):
        body += store_out(k, [Z])
    m.func(*SAMPLE_SIG, locals=[F32], body=body + [("i32.const", OUT)], export="sample")
    tape, _, summary = lowered(S, oracle, m)
    got = oracle.tape_sample(tape, p)
    want = np.repeat(p[:, 2:3], 7, axis=1).astype(f32)
    want[:, 0] = np.where(p[:, 0] > 0, f32(2.0), f32(3.0))
    want[(p[:, 0] > 0) & ~(p[:, 1] > f32(0.5))] = TRAP_SAMPLE
    assert same(got, want)

    # (c) integer division by a position-dependent divisor: traps where trunc(x * 4) == 0; and INT_MIN / -1
    m = base_module()
    q = [X, ("f32.const", 4.0), "f32.mul", ("i32.trunc_sat_f32_s",)]
    body = store_out(0, [("i32.const", 100)] + q + ["i32.div_s", "f32.convert_i32_s"])
    body += store_out(1, [("i32.const", 100)] + q + ["i32.rem_u", "f32.convert_i32_u"])
    body += store_out(2, [("i32.const", -0x80000000), Y, ("f32.const", 0.0), "f32.lt", ("if", [I32]), ("i32.const", -1), "else",
                          ("i32.const", 7), "end", "i32.div_s", "f32.convert_i32_s"])
    for k in range(3, 7):
        body += store_out(k, [Z])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    tape, _, summary = lowered(S, oracle, m)
    got = oracle.tape_sample(tape, p)
    d = np.trunc(p[:, 0] * f32(4)).astype(np.int64)
    want = np.repeat(p[:, 2:3], 7, axis=1).astype(f32)
    dz = np.where(d == 0, 1, d)
    want[:, 0] = (np.sign(dz) * (100 // np.abs(dz))).astype(f32)
    want[:, 1] = (100 % (dz & 0xFFFFFFFF)).astype(f32)
    want[:, 2] = f32(-0x80000000 // 7 + 1)  # truncating division
    want[(d == 0) | (p[:, 1] < 0)] = TRAP_SAMPLE
    assert (d == 0).sum() > 10
    assert same(got, want)
    assert "div/trunc ops" in summary

    # (d) the trapping truncation: i32.trunc_f32_s(x * 3e9) traps outside [-2^31, 2^31) and on NaN
    m = base_module()
    body = store_out(0, [X, ("f32.const", 3.0e9), "f32.mul", "i32.trunc_f32_s", "f32.convert_i32_s"])
    body += store_out(1, [Y, ("f32.const", 5.0e9), "f32.mul", "i32.trunc_f32_u", "f32.convert_i32_u"])
    for k in range(2, 7):
        body += store_out(k, [Z])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    tape, _, summary = lowered(S, oracle, m)
    got = oracle.tape_sample(tape, p)
    a, b = p[:, 0] * f32(3.0e9), p[:, 1] * f32(5.0e9)
    ok = (a >= f32(-2147483648.0)) & (a < f32(2147483648.0)) & (b > f32(-1.0)) & (b < f32(4294967296.0))
    want = np.repeat(p[:, 2:3], 7, axis=1).astype(f32)
    want[:, 0] = np.where(ok, np.trunc(np.where(ok, a, 0)).astype(np.int64), 0).astype(f32)
    want[:, 1] = np.where(ok, np.trunc(np.where(ok, b, 0)).astype(np.int64), 0).astype(f32)
    want[~ok] = TRAP_SAMPLE
    assert 10 < ok.sum() < len(p) - 10
    assert same(got, want)


def test_symbolic_words_in_memory_are_overwritten_whole_or_not_at_all(S, oracle):
    """A struct that held position-dependent values may be zeroed (memory.fill) or overwritten by wider stores and
    reused; a narrow store INTO such a word has no 32-bit form and is refused."""
    def guest(extra):
        m = base_module()
        body = [("i32.const", 3000), X, ("f32.store", 0), ("i32.const", 3004), Y, ("f32.store", 0)]   # spill x, y
        body += extra
        body += store_out(0, [("i32.const", 3000), ("f32.load", 0), ("i32.const", 3004), ("f32.load", 0), "f32.add"])
        for k in range(1, 7):
            body += store_out(k, [("f32.const", 0.5)])
        m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
        return m

    p = points(30)
    # zero both words with memory.fill, then store z over the first
    zeroed = guest([("i32.const", 3000), ("i32.const", 0), ("i32.const", 8), ("memory.fill",), ("i32.const", 3000), Z, ("f32.store", 0)])
    assert same(oracle.tape_sample(S.wasm.lower(zeroed.build())[0], p)[:, 0], p[:, 2] + f32(0.0))
    # an i64 store over both words
    wide = guest([("i32.const", 3000), ("i64.const", 0x3F8000003F000000), ("i64.store", 0)])
    assert same(oracle.tape_sample(S.wasm.lower(wide.build())[0], p)[:, 0], np.full(len(p), 1.5, f32))
    # one byte into a symbolic word
    expect_failure(S, guest([("i32.const", 3001), ("i32.const", 7), ("i32.store8", 0)]), -3, "part of a symbolic word")


def test_struct_copies_through_64_bit_moves(S, oracle):
    """Compilers copy a 28-byte SDFSample with three i64 load / store pairs and one i32: 64-bit values whose words
    depend on the position can be moved (memory, locals, select, a branch that picks one of two structs) though
    not computed with."""
    def guest(tail):
        m = base_module()
        A, B = 3000, 3100
        body = []
        for k, v in enumerate([[X, Y, "f32.add"], [X, "f32.abs"], [Y, "f32.neg"], [Z], [("f32.const", 0.25)], [X, Z, "f32.mul"], [("f32.const", 1.0)]]):
            body += [("i32.const", A)] + v + [("f32.store", 4 * k)]
        for k, v in enumerate([[Z, Y, "f32.sub"], [("f32.const", 0.5)], [Y], [X], [Z, "f32.sqrt"], [("f32.const", 0.0)], [Y, Y, "f32.mul"]]):
            body += [("i32.const", B)] + v + [("f32.store", 4 * k)]
        m.func(*SAMPLE_SIG, locals=[I64, I32], body=body + tail + [("i32.const", OUT)], export="sample")
        return m

    def copy(src, via_local=False):
        out = []
        for off in (0, 8, 16):
            if via_local:
                out += [("i32.const", src), ("i64.load", off), ("local.set", 5), ("i32.const", OUT), ("local.get", 5), ("i64.store", off)]
            else:
                out += [("i32.const", OUT), ("i32.const", src), ("i64.load", off), ("i64.store", off)]
        return out + [("i32.const", OUT), ("i32.const", src), ("i32.load", 24), ("i32.store", 24)]

    p = points(60)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    with np.errstate(all="ignore"):
        a = np.stack([x + y, np.abs(x), -y, z, np.full_like(x, 0.25), x * z, np.ones_like(x)], 1).astype(f32)
        b = np.stack([z - y, np.full_like(x, 0.5), y, x, np.sqrt(z), np.zeros_like(x), y * y], 1).astype(f32)
    assert same(oracle.tape_sample(S.wasm.lower(guest(copy(3000)).build())[0], p), a)
    assert same(oracle.tape_sample(S.wasm.lower(guest(copy(3100, via_local=True)).build())[0], p), b)
    # `if x > y { *out = a } else { *out = b }` with the copies in the arms
    picked = guest([X, Y, "f32.gt", ("if", [])] + copy(3000) + ["else"] + copy(3100, via_local=True) + ["end"])
    assert same(oracle.tape_sample(S.wasm.lower(picked.build())[0], p), np.where((x > y)[:, None], a, b))
    # select between two 64-bit halves of the structs
    sel = []
    for off in (0, 8, 16):
        sel += [("i32.const", OUT), ("i32.const", 3000), ("i64.load", off), ("i32.const", 3100), ("i64.load", off), X, Y, "f32.gt", "select", ("i64.store", off)]
    sel += [("i32.const", OUT), ("i32.const", 3000), ("i32.load", 24), ("i32.const", 3100), ("i32.load", 24), X, Y, "f32.gt", "select", ("i32.store", 24)]
    assert same(oracle.tape_sample(S.wasm.lower(guest(sel).build())[0], p), np.where((x > y)[:, None], a, b))
    # arithmetic on such a value is refused
    expect_failure(S, guest([("i32.const", OUT), ("i32.const", 3000), ("i64.load", 0), ("i64.const", 1), "i64.add", ("i64.store", 0)]), -3,
                   "64-bit arithmetic")


def test_sdf_id_is_forwarded(S, oracle):
    """Every export takes the SDF's id first (0 = root, src/sdf/wasm/mod.rs:8-10): a guest with two children."""
    m = Module()
    m.data_at(BBP, struct.pack("<6f", -1, -1, -1, 1, 1, 1) + struct.pack("<6f", -0.5, -0.5, -0.5, 0.5, 0.5, 0.5))
    m.func([I32], [I32], body=[("i32.const", BBP), ("local.get", 0), ("i32.const", 24), "i32.mul", "i32.add"], export="bounding_box")
    body = store_out(0, [X, ("local.get", 0), "f32.convert_i32_s", "f32.add"])
    for k in range(1, 7):
        body += store_out(k, [("f32.const", 0.0)])
    m.func(*SAMPLE_SIG, body=body + [("i32.const", OUT)], export="sample")
    p = points(20)
    for sdf_id, bb in ((0, BB), (1, ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)))):
        tape, got_bb, _ = lowered(S, oracle, m, sdf_id)
        assert got_bb == bb
        assert same(oracle.tape_sample(tape, p)[:, 0], p[:, 0] + f32(sdf_id))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["csg_calls", "early_returns"])
def test_lowered_guest_fills_on_gpu(S, oracle, name):
    """A lowered WebAssembly guest evaluated by the specialised fill kernel: both volumes equal the oracle's
    interpretation of the same tape bit for bit (first run on a B200: profiles/r01_scalar_wasm_first_gpu_run.log)."""
    dims = (40, 36, 32)
    tape = S.wasm.lower(GUESTS[name][0]().build())[0]
    o = oracle.Viewer(BB, dims, 2)
    o.update(oracle.Sampler(tape=tape))
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        v.set_tape(tape)
        v.update(None)
        t0, t1 = v.download()
        assert v.get_info("last_fill_program") == 1
    assert same(t0, o.tex0) and same(t1, o.tex1)


@pytest.mark.gpu
def test_lowered_guests_fill_on_gpu(S, oracle):
    """GPU run of every lowered guest (incl. the reference's SDFDemo hand-compiled to WebAssembly), bit-exact
    against the oracle's interpretation of the same tape."""
    dims = (40, 36, 32)
    makers = {name: make for name, (make, _) in GUESTS.items()}
    makers["reference_demo"] = guest_reference_demo
    for name, make in makers.items():
        sdf = S.WasmSDF(make().build())
        o = oracle.Viewer(BB, dims, 2)
        o.update(oracle.Sampler(tape=sdf.tape()))
        with S.SDFViewer.new_voxels(dims, BB, 2) as v:
            assert v.update_surface(sdf, 0.030) == o.total_iterations()
            t0, t1 = v.download()
        assert same(t0, o.tex0) and same(t1, o.tex1), name

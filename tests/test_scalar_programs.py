"""Scalar programs of the tape (include/sdfgpu_tape.h, `sdft_sop` / SDFT_OP_SCALAR): straight-line SSA arithmetic
with WebAssembly's numeric semantics -- what a `sample` function compiled to WASM lowers to.

CPU-side evidence (no GPU in the build container):
  * the oracle's interpreter against an independent numpy evaluator, op by op and on random programs;
  * the CUDA text the specialiser generates for a program, compiled FOR THE HOST behind a shim of the
    few intrinsics it uses, against the oracle -- this checks the code generator's wiring;
  * NVRTC compiles that text for sm_100a (sdfgpu_jit_check);
  * malformed programs are rejected.
GPU (`-m gpu`): `test_scalar_tape_sphere_fills_on_gpu` and the wider sweep `test_scalar_tape_fill_on_gpu` (random
programs x voxels-per-thread).  The sweep's first B200 run (round 2) caught the compiler folding `__float2uint_rz` of a
constant NaN word to 1 in the V = 1 kernel; the conversions now spell their special cases out (jit.cu)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
f32, u32, i32 = np.float32, np.uint32, np.int32


def w2f(w):
    return np.array([w], u32).view(f32)[0]


def f2w(f):
    return int(np.array([f], f32).view(u32)[0])


def ref_eval(ops, consts, p):
    """Independent evaluator (numpy scalars) of one program at one point; returns {channel: f32}."""
    v, out = [], {}
    with np.errstate(all="ignore"):
        for op, a, b, c in ops:
            name = NAMES[op]
            A = v[a] if a < len(v) else 0
            B = v[b] if b < len(v) else 0
            Cw = v[c] if c < len(v) else 0
            fa, fb = w2f(A), w2f(B)
            sa, sb = int(np.array([A], u32).view(i32)[0]), int(np.array([B], u32).view(i32)[0])
            if name in ("PX", "PY", "PZ"):
                r = f2w(p["XYZ".index(name[1])])
            elif name == "CONST":
                r = f2w(consts[a])
            elif name == "IMM":
                r = a
            elif name == "FNEG":
                r = A ^ 0x80000000
            elif name == "FABS":
                r = A & 0x7FFFFFFF
            elif name == "FSQRT":
                r = f2w(np.sqrt(fa))
            elif name == "FFLOOR":
                r = f2w(np.floor(fa))
            elif name == "FCEIL":
                r = f2w(np.ceil(fa))
            elif name == "FTRUNC":
                r = f2w(np.trunc(fa))
            elif name == "FNEAREST":
                r = f2w(np.rint(fa))
            elif name in ("FADD", "FSUB", "FMUL", "FDIV"):
                r = f2w({"FADD": fa + fb, "FSUB": fa - fb, "FMUL": fa * fb, "FDIV": np.divide(fa, fb)}[name])
            elif name in ("FMIN", "FMAX"):
                if np.isnan(fa) or np.isnan(fb):
                    r = 0x7FC00000
                elif fa == fb:
                    r = (A | B) if name == "FMIN" else (A & B)
                else:
                    r = f2w(min(fa, fb) if name == "FMIN" else max(fa, fb))
            elif name == "FCOPYSIGN":
                r = (A & 0x7FFFFFFF) | (B & 0x80000000)
            elif name == "FMOD":
                r = f2w(np.fmod(fa, fb))
            elif name in ("FEQ", "FNE", "FLT", "FGT", "FLE", "FGE"):
                r = int({"FEQ": fa == fb, "FNE": fa != fb, "FLT": fa < fb, "FGT": fa > fb, "FLE": fa <= fb, "FGE": fa >= fb}[name])
            elif name == "IADD":
                r = (A + B) & 0xFFFFFFFF
            elif name == "ISUB":
                r = (A - B) & 0xFFFFFFFF
            elif name == "IMUL":
                r = (A * B) & 0xFFFFFFFF
            elif name == "IAND":
                r = A & B
            elif name == "IOR":
                r = A | B
            elif name == "IXOR":
                r = A ^ B
            elif name == "ISHL":
                r = (A << (B & 31)) & 0xFFFFFFFF
            elif name == "ISHR_U":
                r = A >> (B & 31)
            elif name == "ISHR_S":
                r = (sa >> (B & 31)) & 0xFFFFFFFF
            elif name == "IDIV_S":
                r = 0 if (B == 0 or (A == 0x80000000 and B == 0xFFFFFFFF)) else int(abs(sa) // abs(sb) * (1 if (sa < 0) == (sb < 0) else -1)) & 0xFFFFFFFF
            elif name == "IDIV_U":
                r = 0 if B == 0 else A // B
            elif name == "IREM_S":
                r = 0 if (B == 0 or B == 0xFFFFFFFF) else int((abs(sa) % abs(sb)) * (1 if sa >= 0 else -1)) & 0xFFFFFFFF
            elif name == "IREM_U":
                r = 0 if B == 0 else A % B
            elif name in ("IEQ", "INE", "ILT_U", "IGT_U", "ILE_U", "IGE_U"):
                r = int({"IEQ": A == B, "INE": A != B, "ILT_U": A < B, "IGT_U": A > B, "ILE_U": A <= B, "IGE_U": A >= B}[name])
            elif name in ("ILT_S", "IGT_S", "ILE_S", "IGE_S"):
                r = int({"ILT_S": sa < sb, "IGT_S": sa > sb, "ILE_S": sa <= sb, "IGE_S": sa >= sb}[name])
            elif name == "IEQZ":
                r = int(A == 0)
            elif name == "SELECT":
                r = B if A != 0 else Cw
            elif name == "F_FROM_I_S":
                r = f2w(f32(sa))
            elif name == "F_FROM_I_U":
                r = f2w(f32(A))
            elif name == "I_FROM_F_S":
                r = 0 if np.isnan(fa) else int(np.clip(np.trunc(np.float64(fa)), -2 ** 31, 2 ** 31 - 1)) & 0xFFFFFFFF
            elif name == "I_FROM_F_U":
                r = 0 if np.isnan(fa) else int(np.clip(np.trunc(np.float64(fa)), 0, 2 ** 32 - 1))
            elif name == "OUT":
                out[b] = fa
                r = 0
            else:
                raise AssertionError(name)
            v.append(r)
    return out


NAMES = {}


@pytest.fixture(scope="module", autouse=True)
def _names(S):
    NAMES.update({code: name for name, code in S.tape.S.items()})


UNARY = ["FNEG", "FABS", "FSQRT", "FFLOOR", "FCEIL", "FTRUNC", "FNEAREST", "IEQZ", "F_FROM_I_S", "F_FROM_I_U", "I_FROM_F_S",
         "I_FROM_F_U"]
BINARY = ["FADD", "FSUB", "FMUL", "FDIV", "FMIN", "FMAX", "FCOPYSIGN", "FMOD", "FEQ", "FNE", "FLT", "FGT", "FLE", "FGE", "IADD", "ISUB",
          "IMUL", "IAND", "IOR", "IXOR", "ISHL", "ISHR_U", "ISHR_S", "IDIV_S", "IDIV_U", "IREM_S", "IREM_U", "IEQ", "INE", "ILT_S", "ILT_U", "IGT_S", "IGT_U", "ILE_S",
          "ILE_U", "IGE_S", "IGE_U"]
SPECIAL = [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1.5, 2.5, -2.5, 3.0e9, -3.0e9, 5.0e9, np.inf, -np.inf, np.nan, 1e-40, 16777217.0]


FLOAT_ARITH = {"FSQRT", "FFLOOR", "FCEIL", "FTRUNC", "FNEAREST", "FADD", "FSUB", "FMUL", "FDIV", "FMIN", "FMAX", "FMOD", "F_FROM_I_S",
               "F_FROM_I_U"}


def canon(p, r):
    """NaN payloads are not specified (WebAssembly leaves them nondeterministic; x86, numpy and the GPU differ), so
    the tests replace every NaN by the canonical one before its bits can reach an integer op or an output."""
    return p.op("SELECT", p.op("FNE", r, r), p.imm(0x7FC00000), r)


def random_program(T, rng, n_ops=60):
    p = T.ScalarProgram()
    vals = [p.px(), p.py(), p.pz()]
    for s in rng.choice(SPECIAL, 4):
        vals.append(p.const(s))
    for w in (0, 1, 31, 33, 0x80000000, 0xFFFFFFFF, 0x7FC00000, int(rng.integers(0, 2 ** 32))):
        vals.append(p.imm(w))
    for _ in range(n_ops):
        kind = rng.integers(0, 10)
        pick = lambda: int(rng.choice(vals))  # noqa: E731
        if kind < 3:
            name = str(rng.choice(UNARY))
            r = p.op(name, pick())
        elif kind < 9:
            name = str(rng.choice(BINARY))
            r = p.op(name, pick(), pick())
        else:
            name, r = "SELECT", p.op("SELECT", pick(), pick(), pick())
        vals.append(canon(p, r) if name in FLOAT_ARITH else r)
    for ch in range(7):
        p.out(ch, int(rng.choice(vals[-20:])))
    return p


def build_tape(T, p):
    t = T.TapeBuilder()
    t.scalar(p).emit(T.OP_END)
    return t, t.build()


def same_f32(a, b):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    return bool(np.all((a.view(u32) == b.view(u32)) | (np.isnan(a) & np.isnan(b))))


def test_every_op_against_the_numpy_evaluator(S, oracle):
    """One program per op over a grid of special operands (signed zeros, halves, infinities, NaN, values beyond
    the i32 range, a denormal, 2^24 + 1), evaluated by the oracle and by ref_eval."""
    T = S.tape
    words = [f2w(f32(s)) for s in SPECIAL] + [0, 1, 2, 31, 32, 33, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 12345, 0xFFFFFFF0]
    for name in UNARY + BINARY + ["SELECT"]:
        p = T.ScalarProgram()
        ins = [p.imm(w) for w in words]
        outs = []
        for i, a in enumerate(ins):
            for b in (ins if name in BINARY or name == "SELECT" else ins[:1]):
                r = p.op(name, a, b, ins[(i + 3) % len(ins)])
                outs.append(canon(p, r) if name in FLOAT_ARITH else r)
        # fold all results into 7 outputs through XOR so every one is observed
        acc = [p.imm(0) for _ in range(7)]
        for k, r in enumerate(outs):
            acc[k % 7] = p.op("IXOR", p.op("IADD", acc[k % 7], p.imm(k)), r)
        for ch in range(7):
            p.out(ch, acc[ch])
        t, tape = build_tape(T, p)
        got = oracle.tape_sample(tape, np.zeros((1, 3), f32))[0]
        want = ref_eval(t.sops, t.consts, (f32(0), f32(0), f32(0)))
        assert same_f32(got, [want[ch] for ch in range(7)]), name


@pytest.mark.parametrize("seed", range(8))
def test_random_programs_oracle_equals_numpy(S, oracle, seed):
    rng = np.random.default_rng(seed)
    T = S.tape
    t, tape = build_tape(T, random_program(T, rng))
    pts = np.concatenate([rng.uniform(-1, 1, (24, 3)), [[0, 0, 0], [1, -1, 0.5], [-0.0, 1e-30, 3.0e9]]]).astype(f32)
    got = oracle.tape_sample(tape, pts)
    for i, pt in enumerate(pts):
        want = ref_eval(t.sops, t.consts, pt)
        assert same_f32(got[i], [want[ch] for ch in range(7)]), (seed, i)


HOST_SHIM = r'''
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __int2float_rn(int i) { return (float)i; }
static inline float __uint2float_rn(unsigned u) { return (float)u; }
static inline int __float2int_rz(float f) {  // the PTX cvt.rzi.s32.f32: saturating, NaN -> 0
    if (f != f) return 0;
    if (f <= -2147483648.0f) return INT32_MIN;
    if (f >= 2147483648.0f) return INT32_MAX;
    return (int)f;
}
static inline unsigned __float2uint_rz(float f) {
    if (f != f || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (unsigned)f;
}
#define SDFGPU_STEP(op, pc)
%(helpers)s
%(body)s
struct Smp { float d, r, g, b, m, ro, o; };
template <int V> struct Machine { Smp A[V]; float qx[V], qy[V], qz[V]; };
struct Env { const float* consts; };
extern "C" void run(const float* consts, const float* pts, int n, float* out) {
    constexpr int V = %(V)d;
    Env E{consts};
    for (int i = 0; i < n; i += V) {
        Machine<V> M;
        std::memset(&M, 0, sizeof M);
        for (int v = 0; v < V; ++v) { M.qx[v] = pts[3 * (i + v)]; M.qy[v] = pts[3 * (i + v) + 1]; M.qz[v] = pts[3 * (i + v) + 2]; }
        SDFGPU_JIT_BODY
        for (int v = 0; v < V; ++v) std::memcpy(out + 7 * (i + v), &M.A[v], 28);
    }
}
'''


def host_build_of_generated_code(S, tape, V, tmp_path):
    src = S.jit_check(tape, V)            # also proves that NVRTC compiles it for sm_100a
    body = next(line for line in src.splitlines() if line.startswith("#define SDFGPU_JIT_BODY"))
    helpers = src[:src.index("#define SDFGPU_JIT_BODY")]
    cpp = tmp_path / f"gen_{V}.cpp"
    cpp.write_text(HOST_SHIM % {"helpers": helpers, "body": body, "V": V})
    so = tmp_path / f"gen_{V}.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", str(cpp), "-o", str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(so))
    lib.run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("seed,V", [(0, 1), (1, 2), (2, 4), (3, 8), (4, 4)])
def test_generated_cuda_text_on_the_host_equals_oracle(S, oracle, tmp_path, seed, V):
    """The specialiser's output for a scalar program (the text NVRTC compiles for the GPU), built for the host
    with the handful of intrinsics shimmed, gives the oracle's values bit for bit."""
    rng = np.random.default_rng(100 + seed)
    T = S.tape
    t, tape = build_tape(T, random_program(T, rng, n_ops=80))
    lib = host_build_of_generated_code(S, tape, V, tmp_path)
    pts = np.concatenate([rng.uniform(-1, 1, (29, 3)), [[0, 0, 0], [1, -1, 0.5], [-0.0, 1e-30, 3.0e9]]]).astype(f32)
    consts = np.asarray(t.consts, f32)
    out = np.zeros((len(pts), 7), f32)
    lib.run(consts.ctypes.data, pts.ctypes.data, len(pts), out.ctypes.data)
    assert same_f32(out, oracle.tape_sample(tape, pts))


def sphere_with_bands(T, radius=0.7):
    """A hand-written SDF as a scalar program: a sphere whose colour is banded along y (floor / select)."""
    p = T.ScalarProgram()
    x, y, z = p.px(), p.py(), p.pz()
    d = p.op("FSUB", p.op("FSQRT", p.op("FADD", p.op("FADD", p.op("FMUL", x, x), p.op("FMUL", y, y)), p.op("FMUL", z, z))),
             p.const(radius))
    band = p.op("IAND", p.op("I_FROM_F_S", p.op("FFLOOR", p.op("FMUL", y, p.const(8.0)))), p.imm(1))
    p.out(0, d)
    p.out(1, p.op("SELECT", band, p.const(0.9), p.const(0.1)))
    p.out(2, p.op("FABS", x))
    p.out(3, p.op("FMIN", p.op("FMAX", z, p.const(0.0)), p.const(1.0)))
    p.out(4, p.const(0.25))
    p.out(5, p.const(0.5))
    p.out(6, p.const(1.0))
    return p


def test_scalar_program_composes_with_the_rest_of_the_tape(S, oracle):
    """P_SUB moves the program's input; the result unions with a primitive."""
    T = S.tape
    t = T.TapeBuilder()
    c = t.const([0.25, 0.0, -0.25])
    box = t.prim(T.SHAPE_BOX_LINF, (0.5, 0.5, 0.5), 0.2, T.MAT_FLAT, color=(0.2, 0.3, 0.4), occlusion=1.0)
    t.emit(T.OP_P_SUB, c).scalar(sphere_with_bands(T)).emit(T.OP_P_RESET).emit(T.OP_UNION_PRIM, box).emit(T.OP_END)
    tape = t.build()
    pts = np.random.default_rng(5).uniform(-1, 1, (200, 3)).astype(f32)
    got = oracle.tape_sample(tape, pts)
    q = pts - np.array([0.25, 0.0, -0.25], f32)
    d_sphere = np.sqrt((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]).astype(f32) - f32(0.7)
    b = np.abs(pts - f32(0.5))
    d_box = np.maximum(np.maximum(b[:, 0], b[:, 1]), b[:, 2]) - f32(0.2)
    assert same_f32(got[:, 0], np.where(d_box < d_sphere, d_box, d_sphere))
    assert S.jit_check(tape, 4)          # the mixed structure compiles too


def test_malformed_scalar_programs_are_rejected(S):
    T = S.tape

    def rejected(mutate):
        p = sphere_with_bands(T)
        t = T.TapeBuilder()
        t.scalar(p).emit(T.OP_END)
        mutate(t)
        with pytest.raises(S.SdfGpuError) as e:
            S.jit_check(t.build(), 2)
        assert e.value.code == -3
        return str(e.value)

    assert "earlier value" in rejected(lambda t: t.sops.__setitem__(5, (T.S["FADD"], 5, 0, 0)))       # self reference
    assert "earlier value" in rejected(lambda t: t.sops.__setitem__(3, (T.S["FNEG"], 9, 0, 0)))       # forward reference
    assert "unknown op" in rejected(lambda t: t.sops.__setitem__(4, (7, 0, 0, 0)))
    assert "constant" in rejected(lambda t: t.sops.__setitem__(3, (T.S["CONST"], 999, 0, 0)))
    assert "channel" in rejected(lambda t: t.sops.__setitem__(len(t.sops) - 1, (T.S["OUT"], 0, 7, 0)))
    assert "out of range" in rejected(lambda t: t.instr.__setitem__(0, (T.OP_SCALAR, 1, len(t.sops), 0.0)))
    assert "out of range" in rejected(lambda t: t.instr.__setitem__(0, (T.OP_SCALAR, 0, 0, 0.0)))
    # a tape whose header promises more scalar ops than it carries
    t = T.TapeBuilder()
    t.scalar(sphere_with_bands(T)).emit(T.OP_END)
    with pytest.raises(S.SdfGpuError):
        S.jit_check(t.build()[:-16], 2)
    # an operand that names an OUT op (which yields no value): PX, OUT, FNEG v1, OUT -- the validator used to accept
    # it and the specialiser then emitted an undefined identifier
    p = T.ScalarProgram()
    x = p.px()
    o = p.out(0, x)
    p.out(1, p.op("FNEG", o))
    t = T.TapeBuilder()
    t.scalar(p).emit(T.OP_END)
    with pytest.raises(S.SdfGpuError) as e:
        S.tape_validate(t.build())
    assert e.value.code == -3 and "OUT" in str(e.value)
    # every single-operand mutation that points a later op at an OUT is rejected as well
    base = sphere_with_bands(T)
    outs = [i for i, o_ in enumerate(base.ops) if o_[0] == T.S["OUT"]]
    assert outs
    for i in outs[:-1]:
        t = T.TapeBuilder()
        t.scalar(base)
        t.sops.append((T.S["FNEG"], i, 0, 0))
        t.instr[0] = (T.OP_SCALAR, 0, len(t.sops), 0.0)
        t.emit(T.OP_END)
        with pytest.raises(S.SdfGpuError):
            S.tape_validate(t.build())


def _have_gpu(S):
    try:
        S.SDFViewer.new_voxels((2, 2, 2), BB, 1).close()
        return True
    except S.SdfGpuError:
        return False


@pytest.mark.gpu
def test_scalar_tape_sphere_fills_on_gpu(S, oracle):
    """A hand-written scalar program evaluated by the specialised fill kernel, bit-exact against the oracle
    (first run on a B200: profiles/r01_scalar_wasm_first_gpu_run.log)."""
    dims = (40, 36, 32)
    _, tape = build_tape(S.tape, sphere_with_bands(S.tape))
    o = oracle.Viewer(BB, dims, 2)
    o.update(oracle.Sampler(tape=tape))
    with S.SDFViewer.new_voxels(dims, BB, 2) as v:
        v.set_tape(tape)
        v.update(None)
        t0, t1 = v.download()
        assert v.get_info("last_fill_program") == 1
    assert same_f32(t0, o.tex0) and same_f32(t1, o.tex1)


@pytest.mark.gpu
def test_scalar_tape_fill_on_gpu(S, oracle):
    """GPU run of scalar-program tapes (a hand-written one and random ones, every voxels-per-thread variant),
    bit-exact against the oracle."""
    T = S.tape
    dims = (40, 36, 32)
    rng = np.random.default_rng(7)
    progs = [sphere_with_bands(T)] + [random_program(T, rng, n_ops=50) for _ in range(3)]
    for k, p in enumerate(progs):
        _, tape = build_tape(T, p)
        o = oracle.Viewer(BB, dims, 2)
        o.update(oracle.Sampler(tape=tape))
        for vpt in (0, 1, 2, 8):
            with S.SDFViewer.new_voxels(dims, BB, 2) as v:
                v.set_option("fill_voxels_per_thread", vpt)
                v.set_tape(tape)
                v.update(None)
                t0, t1 = v.download()
                assert v.get_info("last_fill_program") == 1
            assert same_f32(t0, o.tex0) and same_f32(t1, o.tex1), (k, vpt)
    with S.SDFViewer.new_voxels(dims, BB, 1) as v:       # the interpreter cannot run them: loud failure
        v.set_option("fill_program", 1)
        v.set_tape(build_tape(T, progs[0])[1])
        with pytest.raises(S.SdfGpuError):
            v.fill_all()


def test_disassembler_lists_every_section(S):
    T = S.tape
    _, tape = build_tape(T, sphere_with_bands(T))
    text = T.disassemble(tape)
    assert "OP_SCALAR" in text and "fsqrt v" in text and "A.d = v" in text and "const[0] = 0.7" in text and "imm 0x00000001" in text
    demo = T.disassemble(T.demo_tape())
    assert "OP_POP_DEMO_DIFF" in demo and "box centre" in demo and "material brick" in demo
    with pytest.raises(ValueError):
        T.disassemble(b"\0" * 64)


def test_every_op_compiles_for_sm100a(S):
    """One program that uses every scalar op: the specialiser's CUDA for it compiles with NVRTC (sm_100a)."""
    T = S.tape
    p = T.ScalarProgram()
    x, y, z = p.px(), p.py(), p.pz()
    k, i = p.const(0.75), p.imm(3)
    vals = [x, y, z, k, i]
    for name in UNARY:
        vals.append(p.op(name, vals[len(vals) % 5]))
    for name in BINARY:
        vals.append(p.op(name, vals[-1], vals[len(vals) % 7]))
    vals.append(p.op("SELECT", vals[-1], vals[-2], vals[-3]))
    for ch in range(7):
        p.out(ch, vals[-1 - ch])
    used = {op for op, _, _, _ in p.ops}
    assert used == set(T.S.values()), sorted(set(T.S.values()) - used)
    _, tape = build_tape(T, p)
    for V in (1, 8):
        src = S.jit_check(tape, V)
        assert "fmodf(" in src and "sdfw_fmin(" in src and "__float2uint_rz(" in src

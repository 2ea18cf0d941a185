"""Builders for the SDF instruction tape (include/sdfgpu_tape.h).

The reference has no op tree -- an SDF is an opaque `SDFSurface::sample(p)`
callback (/root/reference/src/sdf/mod.rs:33-43) -- so the tape is this build's
GPU-side stand-in for it.  `demo_tape` is the hand lowering of the reference's
built-in `SDFDemo` (src/sdf/demo/mod.rs:51-75, cube.rs:79-89, sphere.rs:37-47).
"""
import struct

import numpy as np

SDFT_MAGIC = 0x54464453
SDFT_VERSION = 1
SDFT_MAX_STACK = 8

SHAPE_SPHERE, SHAPE_BOX_LINF = 0, 1
MAT_FLAT, MAT_BRICK, MAT_NORMAL = 0, 1, 2

OP_END = 0
OP_PRIM, OP_UNION_PRIM, OP_INTER_PRIM, OP_UNION_RANGE = 1, 2, 3, 4
OP_PUSH, OP_POP_UNION, OP_POP_INTER, OP_POP_DEMO_DIFF = 8, 9, 10, 11
OP_D_NEG, OP_D_ABS, OP_D_ADD, OP_D_MUL, OP_D_MAX, OP_D_MIN = 16, 17, 18, 19, 20, 21
OP_M_SET = 24
OP_P_RESET, OP_P_SUB, OP_P_MUL, OP_P_ABS = 32, 33, 34, 35
OP_SCALAR = 40

# scalar-program ops (enum sdft_sop_op)
S = {name: code for name, code in dict(
    PX=0, PY=1, PZ=2, CONST=3, IMM=4,
    FNEG=8, FABS=9, FSQRT=10, FFLOOR=11, FCEIL=12, FTRUNC=13, FNEAREST=14,
    FADD=16, FSUB=17, FMUL=18, FDIV=19, FMIN=20, FMAX=21, FCOPYSIGN=22, FMOD=23,
    FEQ=24, FNE=25, FLT=26, FGT=27, FLE=28, FGE=29,
    IADD=32, ISUB=33, IMUL=34, IAND=35, IOR=36, IXOR=37, ISHL=38, ISHR_U=39, ISHR_S=40, IDIV_S=41, IDIV_U=42, IREM_S=43,
    IEQ=44, INE=45, ILT_S=46, ILT_U=47, IGT_S=48, IGT_U=49, ILE_S=50, ILE_U=51, IGE_S=52, IGE_U=53, IEQZ=54, IREM_U=55,
    SELECT=56, F_FROM_I_S=57, F_FROM_I_U=58, I_FROM_F_S=59, I_FROM_F_U=60, OUT=63).items()}

SOP_DTYPE = np.dtype([("op", "<u4"), ("a", "<u4"), ("b", "<u4"), ("c", "<u4")])
INSTR_DTYPE = np.dtype([("op", "<u4"), ("a", "<u4"), ("b", "<u4"), ("imm", "<f4")])
PRIM_DTYPE = np.dtype(
    [("center", "<f4", 3), ("size", "<f4"), ("color", "<f4", 3), ("metallic", "<f4"), ("roughness", "<f4"),
     ("occlusion", "<f4"), ("air_skip", "<f4"), ("kind", "<u4")]
)
assert INSTR_DTYPE.itemsize == 16 and PRIM_DTYPE.itemsize == 48


class TapeBuilder:
    """Accumulates instructions, primitives and constants; `build()` returns the tape bytes."""

    def __init__(self):
        self.instr = []
        self.prims = []
        self.consts = []
        self.sops = []

    # -- tables
    def prim(self, shape, center, size, material=MAT_FLAT, color=(0.0, 0.0, 0.0), metallic=0.0, roughness=0.0,
             occlusion=0.0, air_skip=float("inf")):
        self.prims.append((tuple(center), size, tuple(color), metallic, roughness, occlusion, air_skip,
                           shape | (material << 8)))
        return len(self.prims) - 1

    def prims_from_array(self, arr):
        """Append a PRIM_DTYPE array; returns the index of its first element."""
        first = len(self.prims)
        for r in np.asarray(arr, dtype=PRIM_DTYPE):
            self.prims.append((tuple(r["center"]), r["size"], tuple(r["color"]), r["metallic"], r["roughness"],
                               r["occlusion"], r["air_skip"], int(r["kind"])))
        return first

    def const(self, values):
        first = len(self.consts)
        self.consts.extend(float(v) for v in values)
        return first

    # -- instructions
    def emit(self, op, a=0, b=0, imm=0.0):
        self.instr.append((op, a, b, imm))
        return self

    def scalar(self, program):
        """Append a `ScalarProgram` to the scalar section and emit the SDFT_OP_SCALAR that runs it."""
        first = len(self.sops)
        for op, a, b, c in program.ops:
            if op == S["CONST"]:
                a = self.const([program.const_values[a]])
            self.sops.append((op, a, b, c))
        return self.emit(OP_SCALAR, first, len(program.ops))

    def build(self):
        ins = np.array(self.instr, dtype=INSTR_DTYPE) if self.instr else np.zeros(0, INSTR_DTYPE)
        prims = np.array(self.prims, dtype=PRIM_DTYPE) if self.prims else np.zeros(0, PRIM_DTYPE)
        consts = np.array(self.consts, dtype="<f4")
        sops = np.array(self.sops, dtype=SOP_DTYPE) if self.sops else np.zeros(0, SOP_DTYPE)
        hdr = struct.pack("<8I", SDFT_MAGIC, SDFT_VERSION, len(ins), len(prims), len(consts), len(sops), 0, 0)
        return hdr + ins.tobytes() + prims.tobytes() + consts.tobytes() + sops.tobytes()


class ScalarProgram:
    """A straight-line SSA program over untyped 32-bit words (include/sdfgpu_tape.h, `sdft_sop`): each method
    appends one op and returns the index of its value.  `TapeBuilder.scalar(prog)` places it in a tape."""

    def __init__(self):
        self.ops = []
        self.const_values = []

    def op(self, name, a=0, b=0, c=0):
        self.ops.append((S[name], int(a), int(b), int(c)))
        return len(self.ops) - 1

    def px(self):
        return self.op("PX")

    def py(self):
        return self.op("PY")

    def pz(self):
        return self.op("PZ")

    def const(self, value):
        """An f32 constant: runtime data of the tape (a new value keeps the compiled kernel)."""
        self.const_values.append(float(value))
        return self.op("CONST", len(self.const_values) - 1)

    def imm(self, word):
        return self.op("IMM", int(word) & 0xFFFFFFFF)

    def out(self, channel, value):
        """A[channel] = value; channel 0..6 = distance, r, g, b, metallic, roughness, occlusion."""
        return self.op("OUT", value, channel)


def demo_tape(cube_half_side=0.95, cube_material=MAT_BRICK, sphere_radius=1.05, sphere_material=MAT_NORMAL,
              max_distance_custom_material=0.05, disable_sphere=False):
    """`SDFDemo` (src/sdf/demo/mod.rs:20-32 for the defaults) lowered to the tape."""
    t = TapeBuilder()
    # "the air has no texture": material skipped when the distance exceeds 0.1 (cube.rs:83, sphere.rs:41)
    box = t.prim(SHAPE_BOX_LINF, (0.0, 0.0, 0.0), cube_half_side, cube_material, air_skip=0.1)
    t.emit(OP_PRIM, box)
    if not disable_sphere:  # demo/mod.rs:54-55
        sph = t.prim(SHAPE_SPHERE, (0.0, 0.0, 0.0), sphere_radius, sphere_material, air_skip=0.1)
        # seam threshold, then the forced seam material of demo/mod.rs:66-69
        c = t.const([max_distance_custom_material, 0.5, 0.6, 0.7, 0.5, 0.0, 0.0])
        t.emit(OP_PUSH).emit(OP_PRIM, sph).emit(OP_POP_DEMO_DIFF, c)
    t.emit(OP_END)
    return t.build()


def csg_primitive_table(n=1000, seed=1234):
    """The CSG stress workload of BASELINE.json configs[2]: n random spheres /
    L-inf boxes with flat materials.  Deterministic; the same table feeds the oracle and the kernel."""
    rng = np.random.default_rng(seed)
    arr = np.zeros(n, PRIM_DTYPE)
    shape = (rng.random(n) < 0.5).astype(np.uint32)  # 1 = box
    arr["center"] = rng.uniform(-0.8, 0.8, (n, 3)).astype(np.float32)
    arr["size"] = rng.uniform(0.03, 0.12, n).astype(np.float32)
    arr["color"] = rng.uniform(0.1, 1.0, (n, 3)).astype(np.float32)
    arr["metallic"] = rng.uniform(0.0, 1.0, n).astype(np.float32)
    arr["roughness"] = rng.uniform(0.0, 1.0, n).astype(np.float32)
    arr["occlusion"] = 1.0
    arr["air_skip"] = np.inf
    arr["kind"] = shape | (MAT_FLAT << 8)
    return arr


def csg_tape(table=None, clip_radius=0.98):
    """intersect(union(table[0..n)), sphere(0, clip_radius)); union / intersect carry the winning
    child's whole sample, ties keep the first (include/sdfgpu_tape.h)."""
    if table is None:
        table = csg_primitive_table()
    t = TapeBuilder()
    first = t.prims_from_array(table)
    clip = t.prim(SHAPE_SPHERE, (0.0, 0.0, 0.0), clip_radius, MAT_FLAT, color=(0.8, 0.8, 0.8), metallic=0.1,
                  roughness=0.6, occlusion=1.0)
    t.emit(OP_UNION_RANGE, first, len(table)).emit(OP_INTER_PRIM, clip).emit(OP_END)
    return t.build()


def disassemble(tape_bytes):
    """A readable listing of a tape (instructions, primitives, constants, scalar programs): a debugging aid for
    hand-written tapes and for what sdfgpu_wasm_lower produces."""
    magic, version, n_instr, n_prims, n_consts, n_sops = struct.unpack_from("<6I", tape_bytes, 0)
    if magic != SDFT_MAGIC:
        raise ValueError("not a tape")
    off = 32
    instr = np.frombuffer(tape_bytes, INSTR_DTYPE, n_instr, off); off += 16 * n_instr
    prims = np.frombuffer(tape_bytes, PRIM_DTYPE, n_prims, off); off += 48 * n_prims
    consts = np.frombuffer(tape_bytes, "<f4", n_consts, off); off += 4 * n_consts
    sops = np.frombuffer(tape_bytes, SOP_DTYPE, n_sops, off)
    op_names = {v: k for k, v in globals().items() if k.startswith("OP_") and isinstance(v, int)}
    s_names = {v: k for k, v in S.items()}
    lines = [f"tape v{version}: {n_instr} instructions, {n_prims} primitives, {n_consts} constants, {n_sops} scalar ops"]
    for pc, i in enumerate(instr):
        name = op_names.get(int(i["op"]), f"op{int(i['op'])}")
        lines.append(f"  {pc:3d}  {name:<18s} a={int(i['a'])} b={int(i['b'])} imm={float(i['imm']):g}")
        if int(i["op"]) == OP_SCALAR:
            first, count = int(i["a"]), int(i["b"])
            for k in range(count):
                o = sops[first + k]
                n, a, b, c = s_names.get(int(o["op"]), f"s{int(o['op'])}"), int(o["a"]), int(o["b"]), int(o["c"])
                if n in ("PX", "PY", "PZ"):
                    text = n.lower()
                elif n == "CONST":
                    text = f"const[{a}] = {float(consts[a]):g}" if a < n_consts else f"const[{a}] (out of range)"
                elif n == "IMM":
                    text = f"imm 0x{a:08x}"
                elif n == "OUT":
                    text = f"A.{('d', 'r', 'g', 'b', 'metallic', 'roughness', 'occlusion')[b] if b < 7 else b} = v{a}"
                elif n == "SELECT":
                    text = f"v{a} ? v{b} : v{c}"
                elif n in ("FNEG", "FABS", "FSQRT", "FFLOOR", "FCEIL", "FTRUNC", "FNEAREST", "IEQZ", "F_FROM_I_S", "F_FROM_I_U",
                           "I_FROM_F_S", "I_FROM_F_U"):
                    text = f"{n.lower()} v{a}"
                else:
                    text = f"{n.lower()} v{a}, v{b}"
                lines.append(f"         v{k:<4d} {text}")
    for k, pr in enumerate(prims):
        kind = int(pr["kind"])
        lines.append(f"  prim {k}: {'sphere' if (kind & 0xff) == SHAPE_SPHERE else 'box'} centre {tuple(float(v) for v in pr['center'])} "
                     f"size {float(pr['size']):g} material {('flat', 'brick', 'normal')[(kind >> 8) & 0xff] if (kind >> 8) & 0xff < 3 else kind >> 8}")
    return "\n".join(lines)

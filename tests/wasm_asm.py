"""A minimal WebAssembly binary assembler for the tests of sdfgpu_wasm_lower (no WASM toolchain exists in the
build image, so the guest modules are written here instruction by instruction).  TEST INFRASTRUCTURE.

    m = Module(pages=1)
    f = m.func([F32, F32], [F32], locals=[F32], body=[("local.get", 0), ("local.get", 1), "f32.add"], export="add")
    wasm = m.build()

An instruction is a string ("f32.add") or a tuple: ("local.get", i), ("i32.const", n), ("f32.const", x),
("block", result_types) / ("loop", ...) / ("if", ...), ("br", depth), ("br_if", depth), ("br_table", [..], default),
("call", func), ("call_indirect", type_index), ("f32.load", offset) / any load or store with its offset
(natural alignment), ("global.get", i), ("global.set", i), ("memory.copy",), ("memory.fill",)."""
import struct

I32, I64, F32, F64 = 0x7F, 0x7E, 0x7D, 0x7C

PLAIN = {
    "unreachable": 0x00, "nop": 0x01, "else": 0x05, "end": 0x0B, "return": 0x0F, "drop": 0x1A, "select": 0x1B,
    "i32.eqz": 0x45, "i32.eq": 0x46, "i32.ne": 0x47, "i32.lt_s": 0x48, "i32.lt_u": 0x49, "i32.gt_s": 0x4A, "i32.gt_u": 0x4B,
    "i32.le_s": 0x4C, "i32.le_u": 0x4D, "i32.ge_s": 0x4E, "i32.ge_u": 0x4F,
    "i64.eqz": 0x50, "i64.eq": 0x51, "i64.ne": 0x52, "i64.lt_u": 0x54,
    "f32.eq": 0x5B, "f32.ne": 0x5C, "f32.lt": 0x5D, "f32.gt": 0x5E, "f32.le": 0x5F, "f32.ge": 0x60,
    "i32.clz": 0x67, "i32.ctz": 0x68, "i32.popcnt": 0x69, "i32.add": 0x6A, "i32.sub": 0x6B, "i32.mul": 0x6C, "i32.div_s": 0x6D,
    "i32.div_u": 0x6E, "i32.rem_s": 0x6F, "i32.rem_u": 0x70, "i32.and": 0x71, "i32.or": 0x72, "i32.xor": 0x73, "i32.shl": 0x74,
    "i32.shr_s": 0x75, "i32.shr_u": 0x76, "i32.rotl": 0x77, "i32.rotr": 0x78,
    "i64.add": 0x7C, "i64.sub": 0x7D, "i64.mul": 0x7E, "i64.and": 0x83, "i64.or": 0x84, "i64.xor": 0x85, "i64.shl": 0x86,
    "i64.shr_u": 0x88, "i64.rotl": 0x89,
    "f32.abs": 0x8B, "f32.neg": 0x8C, "f32.ceil": 0x8D, "f32.floor": 0x8E, "f32.trunc": 0x8F, "f32.nearest": 0x90, "f32.sqrt": 0x91,
    "f32.add": 0x92, "f32.sub": 0x93, "f32.mul": 0x94, "f32.div": 0x95, "f32.min": 0x96, "f32.max": 0x97, "f32.copysign": 0x98,
    "f64.add": 0xA0, "f64.mul": 0xA2, "f64.sqrt": 0x9F,
    "i32.wrap_i64": 0xA7, "i32.trunc_f32_s": 0xA8, "i32.trunc_f32_u": 0xA9, "i64.extend_i32_s": 0xAC, "i64.extend_i32_u": 0xAD,
    "f32.convert_i32_s": 0xB2, "f32.convert_i32_u": 0xB3, "f32.demote_f64": 0xB6, "f64.promote_f32": 0xBB,
    "i32.reinterpret_f32": 0xBC, "f32.reinterpret_i32": 0xBE, "i32.extend8_s": 0xC0, "i32.extend16_s": 0xC1,
}
MEM = {  # name -> (opcode, log2 of the natural alignment)
    "i32.load": (0x28, 2), "i64.load": (0x29, 3), "f32.load": (0x2A, 2), "f64.load": (0x2B, 3), "i32.load8_s": (0x2C, 0),
    "i32.load8_u": (0x2D, 0), "i32.load16_s": (0x2E, 1), "i32.load16_u": (0x2F, 1),
    "i32.store": (0x36, 2), "i64.store": (0x37, 3), "f32.store": (0x38, 2), "f64.store": (0x39, 3), "i32.store8": (0x3A, 0),
    "i32.store16": (0x3B, 1),
}
INDEXED = {"br": 0x0C, "br_if": 0x0D, "call": 0x10, "local.get": 0x20, "local.set": 0x21, "local.tee": 0x22, "global.get": 0x23,
           "global.set": 0x24}
SAT = {"i32.trunc_sat_f32_s": 0, "i32.trunc_sat_f32_u": 1}


def uleb(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def sleb(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if (n == 0 and not (b & 0x40)) or (n == -1 and (b & 0x40)):
            out.append(b)
            return bytes(out)
        out.append(b | 0x80)


def vec(items):
    return uleb(len(items)) + b"".join(items)


def name(s):
    b = s.encode()
    return uleb(len(b)) + b


def encode(instr, module):
    if isinstance(instr, str):
        instr = (instr,)
    op, args = instr[0], instr[1:]
    if op in PLAIN:
        return bytes([PLAIN[op]])
    if op in INDEXED:
        return bytes([INDEXED[op]]) + uleb(args[0])
    if op in MEM:
        code, align = MEM[op]
        return bytes([code]) + uleb(align) + uleb(args[0] if args else 0)
    if op in ("block", "loop", "if"):
        code = {"block": 0x02, "loop": 0x03, "if": 0x04}[op]
        res = list(args[0]) if args else []
        if len(res) == 0:
            bt = bytes([0x40])
        elif len(res) == 1:
            bt = bytes([res[0]])
        else:
            bt = sleb(module.type_index([], res))
        return bytes([code]) + bt
    if op == "i32.const":
        v = args[0]
        if v >= 2 ** 31:
            v -= 2 ** 32
        return bytes([0x41]) + sleb(v)
    if op == "i64.const":
        v = args[0]
        if v >= 2 ** 63:
            v -= 2 ** 64
        return bytes([0x42]) + sleb(v)
    if op == "f32.const":
        return bytes([0x43]) + struct.pack("<f", args[0])
    if op == "f32.const_bits":
        return bytes([0x43]) + struct.pack("<I", args[0])
    if op == "f64.const":
        return bytes([0x44]) + struct.pack("<d", args[0])
    if op == "br_table":
        return bytes([0x0E]) + vec([uleb(t) for t in args[0]]) + uleb(args[1])
    if op == "call_indirect":
        return bytes([0x11]) + uleb(args[0]) + b"\x00"
    if op == "memory.size":
        return b"\x3f\x00"
    if op == "memory.grow":
        return b"\x40\x00"
    if op == "memory.copy":
        return b"\xfc\x0a\x00\x00"
    if op == "memory.fill":
        return b"\xfc\x0b\x00"
    if op in SAT:
        return b"\xfc" + uleb(SAT[op])
    if op == "raw":
        return bytes(args[0])
    raise ValueError(f"unknown instruction {op}")


class Module:
    def __init__(self, pages=1, export_memory=True):
        self.types, self.funcs, self.exports, self.globals, self.data, self.elems = [], [], [], [], [], []
        self.imports = []
        self.pages, self.table_size, self.start = pages, 0, None
        if export_memory:
            self.exports.append(("memory", 2, 0))

    def type_index(self, params, results):
        t = (tuple(params), tuple(results))
        if t not in self.types:
            self.types.append(t)
        return self.types.index(t)

    def import_func(self, module, field, params, results):
        assert not self.funcs, "imports come first in the function index space"
        self.imports.append((module, field, self.type_index(params, results)))
        return len(self.imports) - 1

    def func(self, params, results, locals=(), body=(), export=None):
        idx = len(self.imports) + len(self.funcs)
        self.funcs.append((self.type_index(params, results), list(locals), list(body)))
        if export:
            self.exports.append((export, 0, idx))
        return idx

    def global_(self, valtype, value, mutable=True):
        self.globals.append((valtype, mutable, value))
        return len(self.globals) - 1

    def data_at(self, offset, payload):
        self.data.append((offset, bytes(payload)))

    def table(self, entries, offset=0):
        self.elems.append((offset, list(entries)))
        self.table_size = max(self.table_size, offset + len(entries))

    def build(self):
        def section(sid, payload):
            return bytes([sid]) + uleb(len(payload)) + payload

        out = b"\0asm" + struct.pack("<I", 1)
        # encode the bodies first: multi-value block types may add entries to the type section
        bodies = []
        for _, locs, body in self.funcs:
            code = b"".join(encode(i, self) for i in body) + b"\x0b"
            bodies.append(vec([uleb(1) + bytes([t]) for t in locs]) + code)
        out += section(1, vec([b"\x60" + vec([bytes([p]) for p in ps]) + vec([bytes([r]) for r in rs]) for ps, rs in self.types]))
        if self.imports:
            out += section(2, vec([name(m) + name(f) + b"\x00" + uleb(t) for m, f, t in self.imports]))
        out += section(3, vec([uleb(t) for t, _, _ in self.funcs]))
        if self.table_size:
            out += section(4, vec([b"\x70\x00" + uleb(self.table_size)]))
        out += section(5, vec([b"\x00" + uleb(self.pages)]))
        if self.globals:
            g = []
            for vt, mut, val in self.globals:
                init = {I32: lambda v: b"\x41" + sleb(v), F32: lambda v: b"\x43" + struct.pack("<f", v)}[vt](val)
                g.append(bytes([vt, 1 if mut else 0]) + init + b"\x0b")
            out += section(6, vec(g))
        out += section(7, vec([name(n) + bytes([k]) + uleb(i) for n, k, i in self.exports]))
        if self.start is not None:
            out += section(8, uleb(self.start))
        if self.elems:
            out += section(9, vec([b"\x00\x41" + sleb(off) + b"\x0b" + vec([uleb(f) for f in fs]) for off, fs in self.elems]))
        out += section(10, vec([uleb(len(b)) + b for b in bodies]))
        if self.data:
            out += section(11, vec([b"\x00\x41" + sleb(off) + b"\x0b" + uleb(len(p)) + p for off, p in self.data]))
        return out

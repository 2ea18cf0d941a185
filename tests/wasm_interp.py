"""A direct interpreter of the instruction lists of tests/wasm_asm.py (not of the binary): the independent reference
for differential tests of sdfgpu_wasm_lower.  It executes a guest's `sample` at ONE concrete point with numpy
float32 arithmetic and WebAssembly's rules for min / max / nearest / saturating truncation.  TEST INFRASTRUCTURE:
covers the instructions the random guest generator (tests/test_wasm_fuzz.py) and the hand-written guests use."""
import struct

import numpy as np

f32 = np.float32
MASK32 = 0xFFFFFFFF


class Trap(Exception):
    pass


def _bits(x):
    return int(np.array([x], f32).view(np.uint32)[0])


def _from_bits(w):
    return np.array([w & MASK32], np.uint32).view(f32)[0]


def _s32(w):
    w &= MASK32
    return w - (1 << 32) if w & 0x80000000 else w


def wasm_min(a, b):
    if np.isnan(a) or np.isnan(b):
        return f32(np.nan)
    if a == b:
        return _from_bits(_bits(a) | _bits(b))
    return a if a < b else b


def wasm_max(a, b):
    if np.isnan(a) or np.isnan(b):
        return f32(np.nan)
    if a == b:
        return _from_bits(_bits(a) & _bits(b))
    return a if a > b else b


def trunc_sat(x, signed):
    if np.isnan(x):
        return 0
    t = float(np.trunc(np.float64(x)))
    lo, hi = (-(2 ** 31), 2 ** 31 - 1) if signed else (0, 2 ** 32 - 1)
    return int(min(max(t, lo), hi)) & MASK32


class Instance:
    def __init__(self, module):
        self.m = module
        self.mem = bytearray(module.pages * 65536)
        for off, payload in module.data:
            self.mem[off:off + len(payload)] = payload
        self.globals = [v for _, _, v in module.globals]
        self.table = {}
        for off, fs in module.elems:
            for k, f in enumerate(fs):
                self.table[off + k] = f
        self.exports = {n: i for n, k, i in module.exports if k == 0}
        self._match = {}

    # ---- structure: positions of the else / end that close each block / loop / if
    def _matches(self, body):
        key = id(body)
        if key in self._match:
            return self._match[key]
        out, open_ = {}, []
        for pc, ins in enumerate(body):
            op = ins if isinstance(ins, str) else ins[0]
            if op in ("block", "loop", "if"):
                open_.append(pc)
                out[pc] = [None, None]
            elif op == "else":
                out[open_[-1]][0] = pc
            elif op == "end":
                out[open_.pop()][1] = pc
        self._match[key] = out
        return out

    def call(self, func_index, args):
        n_imp = len(self.m.imports)
        if func_index < n_imp:
            raise Trap("host import")
        type_index, locs, body = self.m.funcs[func_index - n_imp]
        params, results = self.m.types[type_index]
        locals_ = list(args) + [f32(0) if t == 0x7D else 0 for t in locs]
        stack = []
        match = self._matches(body)
        labels = []          # (kind, start_pc, end_pc, height, arity)
        pc = 0

        def branch(depth):
            nonlocal pc, stack
            if depth >= len(labels):
                return True  # to the function body: return
            kind, start, end, height, arity = labels[len(labels) - 1 - depth]
            keep = stack[len(stack) - arity:] if arity else []
            del stack[height:]
            stack.extend(keep)
            if kind == "loop":
                del labels[len(labels) - depth:]
                pc = start + 1
            else:
                del labels[len(labels) - 1 - depth:]
                pc = end + 1
            return False

        while pc < len(body):
            ins = body[pc]
            op, a = (ins, ()) if isinstance(ins, str) else (ins[0], ins[1:])
            pc += 1
            if op in ("block", "loop"):
                res = len(a[0]) if a else 0
                labels.append((op, pc - 1, match[pc - 1][1], len(stack), 0 if op == "loop" else res))
            elif op == "if":
                res = len(a[0]) if a else 0
                c = stack.pop()
                els, end = match[pc - 1]
                if c & MASK32:
                    labels.append(("if", pc - 1, end, len(stack), res))
                elif els is not None:
                    labels.append(("if", pc - 1, end, len(stack), res))
                    pc = els + 1
                else:
                    pc = end + 1
            elif op == "else":
                kind, start, end, height, arity = labels.pop()
                pc = end + 1
            elif op == "end":
                if labels:
                    labels.pop()
            elif op == "br":
                if branch(a[0]):
                    break
            elif op == "br_if":
                if stack.pop() & MASK32:
                    if branch(a[0]):
                        break
            elif op == "br_table":
                k = stack.pop() & MASK32
                if branch(a[0][k] if k < len(a[0]) else a[1]):
                    break
            elif op == "return":
                break
            elif op == "unreachable":
                raise Trap("unreachable")
            elif op == "nop":
                pass
            elif op == "call":
                n_imp2 = len(self.m.imports)
                if a[0] < n_imp2:
                    raise Trap("host import")
                ps, rs = self.m.types[self.m.funcs[a[0] - n_imp2][0]]
                argv = stack[len(stack) - len(ps):] if ps else []
                del stack[len(stack) - len(ps):]
                stack.extend(self.call(a[0], argv))
            elif op == "call_indirect":
                k = stack.pop() & MASK32
                if k not in self.table:
                    raise Trap("undefined element")
                ps, rs = self.m.types[a[0]]
                argv = stack[len(stack) - len(ps):] if ps else []
                del stack[len(stack) - len(ps):]
                stack.extend(self.call(self.table[k], argv))
            elif op == "drop":
                stack.pop()
            elif op == "select":
                c, y, x = stack.pop(), stack.pop(), stack.pop()
                stack.append(x if c & MASK32 else y)
            elif op == "local.get":
                stack.append(locals_[a[0]])
            elif op == "local.set":
                locals_[a[0]] = stack.pop()
            elif op == "local.tee":
                locals_[a[0]] = stack[-1]
            elif op == "global.get":
                stack.append(self.globals[a[0]])
            elif op == "global.set":
                self.globals[a[0]] = stack.pop()
            elif op == "i32.const":
                stack.append(a[0] & MASK32)
            elif op == "f32.const":
                stack.append(f32(a[0]))
            elif op == "f32.load":
                addr = (stack.pop() + a[0]) & MASK32
                stack.append(f32(struct.unpack_from("<f", self.mem, addr)[0]))
            elif op == "i32.load":
                addr = (stack.pop() + a[0]) & MASK32
                stack.append(struct.unpack_from("<I", self.mem, addr)[0])
            elif op == "f32.store":
                v, addr = stack.pop(), (stack.pop() + a[0]) & MASK32
                self.mem[addr:addr + 4] = np.array([v], f32).tobytes()
            elif op == "i32.store":
                v, addr = stack.pop(), (stack.pop() + a[0]) & MASK32
                struct.pack_into("<I", self.mem, addr, v & MASK32)
            elif op == "memory.copy":
                n, s, d = stack.pop(), stack.pop(), stack.pop()
                self.mem[d:d + n] = bytes(self.mem[s:s + n])
            else:
                self._numeric(op, stack)
        n_res = len(results)
        return stack[len(stack) - n_res:] if n_res else []

    def _numeric(self, op, st):
        with np.errstate(all="ignore"):
            if op in ("f32.abs", "f32.neg", "f32.ceil", "f32.floor", "f32.trunc", "f32.nearest", "f32.sqrt"):
                x = st.pop()
                st.append({"f32.abs": lambda: _from_bits(_bits(x) & 0x7FFFFFFF), "f32.neg": lambda: _from_bits(_bits(x) ^ 0x80000000),
                           "f32.ceil": lambda: np.ceil(x), "f32.floor": lambda: np.floor(x), "f32.trunc": lambda: np.trunc(x),
                           "f32.nearest": lambda: np.rint(x), "f32.sqrt": lambda: np.sqrt(x)}[op]())
            elif op in ("f32.add", "f32.sub", "f32.mul", "f32.div", "f32.min", "f32.max", "f32.copysign"):
                y, x = st.pop(), st.pop()
                st.append({"f32.add": lambda: x + y, "f32.sub": lambda: x - y, "f32.mul": lambda: x * y, "f32.div": lambda: np.divide(x, y),
                           "f32.min": lambda: wasm_min(x, y), "f32.max": lambda: wasm_max(x, y),
                           "f32.copysign": lambda: _from_bits((_bits(x) & 0x7FFFFFFF) | (_bits(y) & 0x80000000))}[op]())
            elif op in ("f32.eq", "f32.ne", "f32.lt", "f32.gt", "f32.le", "f32.ge"):
                y, x = st.pop(), st.pop()
                st.append(int({"f32.eq": x == y, "f32.ne": x != y, "f32.lt": x < y, "f32.gt": x > y, "f32.le": x <= y, "f32.ge": x >= y}[op]))
            elif op == "i32.eqz":
                st.append(int((st.pop() & MASK32) == 0))
            elif op in ("i32.add", "i32.sub", "i32.mul", "i32.and", "i32.or", "i32.xor", "i32.shl", "i32.shr_u", "i32.shr_s"):
                y, x = st.pop() & MASK32, st.pop() & MASK32
                st.append({"i32.add": x + y, "i32.sub": x - y, "i32.mul": x * y, "i32.and": x & y, "i32.or": x | y, "i32.xor": x ^ y,
                           "i32.shl": x << (y & 31), "i32.shr_u": x >> (y & 31), "i32.shr_s": _s32(x) >> (y & 31)}[op] & MASK32)
            elif op in ("i32.div_s", "i32.div_u", "i32.rem_s", "i32.rem_u"):
                y, x = st.pop() & MASK32, st.pop() & MASK32
                if y == 0 or (op == "i32.div_s" and x == 0x80000000 and y == MASK32):
                    raise Trap("integer division")
                sx, sy = _s32(x), _s32(y)
                q = abs(sx) // abs(sy) * (1 if (sx < 0) == (sy < 0) else -1)
                st.append({"i32.div_s": q, "i32.div_u": x // y, "i32.rem_s": sx - q * sy, "i32.rem_u": x % y}[op] & MASK32)
            elif op in ("i32.eq", "i32.ne", "i32.lt_u", "i32.gt_u", "i32.le_u", "i32.ge_u"):
                y, x = st.pop() & MASK32, st.pop() & MASK32
                st.append(int({"i32.eq": x == y, "i32.ne": x != y, "i32.lt_u": x < y, "i32.gt_u": x > y, "i32.le_u": x <= y, "i32.ge_u": x >= y}[op]))
            elif op in ("i32.lt_s", "i32.gt_s", "i32.le_s", "i32.ge_s"):
                y, x = _s32(st.pop()), _s32(st.pop())
                st.append(int({"i32.lt_s": x < y, "i32.gt_s": x > y, "i32.le_s": x <= y, "i32.ge_s": x >= y}[op]))
            elif op == "f32.convert_i32_s":
                st.append(f32(_s32(st.pop())))
            elif op == "f32.convert_i32_u":
                st.append(f32(st.pop() & MASK32))
            elif op in ("i32.trunc_sat_f32_s", "i32.trunc_f32_s"):
                st.append(trunc_sat(st.pop(), True))
            elif op in ("i32.trunc_sat_f32_u", "i32.trunc_f32_u"):
                st.append(trunc_sat(st.pop(), False))
            elif op == "i32.reinterpret_f32":
                st.append(_bits(st.pop()))
            elif op == "f32.reinterpret_i32":
                st.append(_from_bits(st.pop()))
            else:
                raise NotImplementedError(op)


def sample(module, point, sdf_id=0):
    """The seven floats the guest's `sample(sdf_id, x, y, z, 0)` leaves behind the pointer it returns, on a fresh
    instance (after its optional init())."""
    inst = Instance(module)
    if "init" in inst.exports:
        inst.call(inst.exports["init"], [])
    (ptr,) = inst.call(inst.exports["sample"], [sdf_id, f32(point[0]), f32(point[1]), f32(point[2]), 0])
    return np.frombuffer(bytes(inst.mem[ptr:ptr + 28]), f32).copy()

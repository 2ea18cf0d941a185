"""Differential fuzzing of sdfgpu_wasm_lower: random guests (nested if / else with and without results, br_if out of
blocks, early returns that write their own result, counted loops, selects, integer work on truncated coordinates,
helper calls) are lowered to a tape and evaluated by the oracle, and executed directly by the reference interpreter
of tests/wasm_interp.py; the seven floats must agree bit for bit at every test point.  This exercises the part of
the lowering that formulas cannot: forking at branches that depend on the position and merging the paths' results.
(40 seeds run here; the same loop over 15 000 further seeds was run once offline without a mismatch, and swapping the
operands of either merge in wasm_lower.cu makes dozens of these tests fail.)"""
import numpy as np
import pytest

import wasm_interp
from test_wasm_lower import BBP, OUT, SAMPLE_SIG, base_module, same
from wasm_asm import F32, I32

f32 = np.float32
N_F, N_I = 5, 3                      # scratch locals: f32 at 5..9, i32 at 10..12
F_LOCALS = list(range(5, 5 + N_F))
I_LOCALS = list(range(5 + N_F, 5 + N_F + N_I))
CONSTS = [0.0, 0.25, 0.5, 1.0, -0.75, 2.0, 3.5, -0.125, 8.0, 0.1]
SCRATCH = [3000, 3004, 3008, 3012]   # guest memory the statements spill to and the expressions read back


class Gen:
    def __init__(self, rng, helper):
        self.rng = rng
        self.helper = helper
        self.forks = 0

    def pick(self, seq):
        return seq[int(self.rng.integers(0, len(seq)))]

    # ---- expressions
    def fexpr(self, d=0):
        r = self.rng.integers(0, 13 if d < 3 else 4)
        if r == 12:
            return [("i32.const", self.pick(SCRATCH)), ("f32.load", 0)]
        if r == 0:
            return [("local.get", self.pick([1, 2, 3]))]
        if r == 1:
            return [("local.get", self.pick(F_LOCALS))]
        if r == 2:
            return [("f32.const", self.pick(CONSTS))]
        if r == 3:
            return [("local.get", self.pick([1, 2, 3])), ("f32.const", self.pick(CONSTS)), self.pick(["f32.add", "f32.mul", "f32.sub"])]
        if r in (4, 5):
            return self.fexpr(d + 1) + [self.pick(["f32.abs", "f32.neg", "f32.sqrt", "f32.floor", "f32.ceil", "f32.trunc", "f32.nearest"])]
        if r in (6, 7, 8):
            return self.fexpr(d + 1) + self.fexpr(d + 1) + [self.pick(["f32.add", "f32.sub", "f32.mul", "f32.div", "f32.min", "f32.max"])]
        if r == 9:
            return self.fexpr(d + 1) + self.fexpr(d + 1) + self.cond(d + 1) + ["select"]
        if r == 10:
            return self.iexpr(d + 1) + [self.pick(["f32.convert_i32_s", "f32.convert_i32_u"])]
        return self.fexpr(d + 1) + self.fexpr(d + 1) + [("call", self.helper)]

    def iexpr(self, d=0):
        r = self.rng.integers(0, 9 if d < 3 else 3)
        if r == 8:  # division / remainder by 1..8 (a zero divisor would trap in the guest)
            return self.iexpr(d + 1) + self.iexpr(d + 1) + [("i32.const", 7), "i32.and", ("i32.const", 1), "i32.add",
                                                            self.pick(["i32.div_s", "i32.div_u", "i32.rem_s", "i32.rem_u"])]
        if r == 0:
            return [("i32.const", int(self.rng.integers(-4, 9)) & 0xFFFFFFFF)]
        if r == 1:
            return [("local.get", self.pick(I_LOCALS))]
        if r in (2, 3):
            return self.fexpr(d + 1) + [("f32.const", self.pick([1.0, 4.0, 8.0])), "f32.mul", (self.pick(["i32.trunc_sat_f32_s", "i32.trunc_sat_f32_u"]),)]
        if r in (4, 5):
            return self.iexpr(d + 1) + self.iexpr(d + 1) + [self.pick(["i32.add", "i32.sub", "i32.mul", "i32.and", "i32.or", "i32.xor", "i32.shl",
                                                                        "i32.shr_s", "i32.shr_u"])]
        if r == 6:
            return self.cond(d + 1)
        return self.iexpr(d + 1) + self.iexpr(d + 1) + self.cond(d + 1) + ["select"]

    def cond(self, d=0):
        r = self.rng.integers(0, 4)
        if r < 2:
            return self.fexpr(d + 1) + self.fexpr(d + 1) + [self.pick(["f32.lt", "f32.gt", "f32.le", "f32.ge", "f32.eq", "f32.ne"])]
        if r == 2:
            return self.iexpr(d + 1) + self.iexpr(d + 1) + [self.pick(["i32.lt_s", "i32.lt_u", "i32.gt_s", "i32.ge_u", "i32.eq", "i32.ne", "i32.le_s"])]
        return self.iexpr(d + 1) + ["i32.eqz"]

    # ---- statements
    def write_result(self):
        out = []
        for k in range(7):
            out += [("i32.const", OUT)] + self.fexpr(2) + [("f32.store", 4 * k)]
        return out

    def stmts(self, n, d):
        out = []
        for _ in range(n):
            out += self.stmt(d)
        return out

    def stmt(self, d):
        r = self.rng.integers(0, 11)
        can_fork = self.forks < 7 and d < 3
        if r == 10 and can_fork:  # a `match` on a position-dependent integer: br_table out of nested blocks
            self.forks += 2
            n = int(self.rng.integers(2, 4))
            out = [("block", [])] * (n + 1) + self.iexpr() + [("br_table", list(range(n)), n), "end"]
            for k in range(n):
                out += self.stmts(1, d + 1) + [("br", n - 1 - k), "end"]
            return out
        if r == 10:
            r = 1
        if r == 0:  # spill to guest memory (on this path only, when inside a branch)
            return [("i32.const", self.pick(SCRATCH))] + self.fexpr() + [("f32.store", 0)]
        if r < 3 or not can_fork and r >= 5:
            return self.fexpr() + [("local.set", self.pick(F_LOCALS))]
        if r == 3:
            return self.iexpr() + [("local.set", self.pick(I_LOCALS))]
        if r == 4:  # a counted loop (the trip count does not depend on the position)
            i = self.pick(I_LOCALS)
            n = int(self.rng.integers(1, 4))
            body = self.fexpr(1) + [("local.set", self.pick(F_LOCALS))]
            return [("i32.const", 0), ("local.set", i), ("block", []), ("loop", []), ("local.get", i), ("i32.const", n), "i32.ge_u", ("br_if", 1)] + \
                body + [("local.get", i), ("i32.const", 1), "i32.add", ("local.set", i), ("br", 0), "end", "end"]
        self.forks += 1
        if r == 5:
            return self.cond() + [("if", [])] + self.stmts(int(self.rng.integers(1, 3)), d + 1) + ["end"]
        if r == 6:
            return self.cond() + [("if", [])] + self.stmts(int(self.rng.integers(1, 3)), d + 1) + ["else"] + \
                self.stmts(int(self.rng.integers(1, 3)), d + 1) + ["end"]
        if r == 7:  # if with a result
            return self.cond() + [("if", [F32])] + self.fexpr(1) + ["else"] + self.fexpr(1) + ["end", ("local.set", self.pick(F_LOCALS))]
        if r == 8:  # leave a block early
            return [("block", [])] + self.stmts(1, d + 1) + self.cond() + [("br_if", 0)] + self.stmts(int(self.rng.integers(1, 3)), d + 1) + ["end"]
        # an early return that writes its own result
        return self.cond() + [("if", [])] + self.write_result() + [("i32.const", OUT), "return", "end"]


def random_guest(seed):
    rng = np.random.default_rng(seed)
    m = base_module()
    helper = m.func([F32, F32], [F32], locals=[F32], body=[
        ("local.get", 0), ("local.get", 1), "f32.mul", ("local.tee", 2), ("local.get", 0), "f32.gt",
        ("if", [F32]), ("local.get", 2), ("local.get", 1), "f32.sub", "else", ("local.get", 0), ("f32.const", 0.5), "f32.add", "end"])
    g = Gen(rng, helper)
    body = []
    for k, loc in enumerate(F_LOCALS):
        body += [("local.get", 1 + k % 3), ("f32.const", CONSTS[k]), "f32.add", ("local.set", loc)]
    body += g.stmts(int(rng.integers(4, 9)), 0) + g.write_result() + [("i32.const", OUT)]
    m.func(*SAMPLE_SIG, locals=[F32] * N_F + [I32] * N_I, body=body, export="sample")
    return m, g.forks


def fuzz_points(rng):
    p = rng.uniform(-1, 1, (40, 3)).astype(f32)
    p[:6] = [[0, 0, 0], [0.5, 0.5, 0.5], [-0.0, 0.25, -0.5], [1, -1, 1], [0.125, -0.125, 0.0], [-1, -1, -1]]
    return p


@pytest.mark.parametrize("seed", range(40))
def test_random_guest_lowers_to_what_it_computes(S, oracle, seed):
    m, forks = random_guest(1000 + seed)
    tape, bb, summary = S.wasm.lower(m.build())
    p = fuzz_points(np.random.default_rng(seed))
    got = oracle.tape_sample(tape, p)
    want = np.array([wasm_interp.sample(m, pt) for pt in p])
    bad = ~((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want)))
    assert not bad.any(), (seed, summary, np.argwhere(bad)[:4], got[bad][:4], want[bad][:4])


def test_random_guests_fork_and_compile(S):
    """The generator really produces position-dependent control flow, and the specialiser's CUDA for such a
    program compiles for sm_100a."""
    merged = []
    for seed in range(1000, 1012):
        m, forks = random_guest(seed)
        tape, _, summary = S.wasm.lower(m.build())
        merged.append(int(summary.split("constants, ")[1].split(" ")[0]))
    assert max(merged) >= 8 and sum(1 for n in merged if n > 0) >= 8
    m, _ = random_guest(1003)
    assert S.jit_check(S.wasm.lower(m.build())[0], 4)

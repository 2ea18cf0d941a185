// wasm_lower.cu -- host code only: lowers the `sample` export of a WebAssembly SDF module to a tape
// (the north star's "JIT-lowers the loaded SDF").
//
// The reference runs the guest's `sample(sdf_id, x, y, z, distance_only) -> *SDFSample` once per voxel inside a
// sandbox (/root/reference/src/sdf/wasm/mod.rs:5-37 is the ABI, src/sdf/wasm/native.rs:29-98,188-217 the host).
// Here the module is executed ONCE, partially: integers, addresses, globals and memory are concrete, the three
// coordinates are symbolic.  Every f32 / i32 operation that touches a symbolic word appends one op to a scalar
// program (include/sdfgpu_tape.h, `sdft_sop`: WebAssembly's own numeric semantics); everything else --
// allocator, registry lookups, vtable calls, loops with concrete trip counts -- simply runs and disappears.
// A branch on a symbolic condition forks the execution; the two sides are merged with selects where they meet
// again (if-conversion of stack, locals, globals and memory: merge_states), or, when a side leaves for good, at
// the end of the call (the seven floats each path leaves in guest memory: merge), so the tape is branch free.  What cannot be expressed makes the
// lowering fail with a message (symbolic addresses or loop bounds, i64 / f64 arithmetic on symbolic values,
// host imports, SIMD): the caller then samples that SDF on the host (sdfgpu_update_surface).  The guest's libm
// fmodf (what `%` on floats calls) is recognised by what it computes and becomes one op (behaves_like_fmodf).
//
// A small, self-contained interpreter of the WebAssembly MVP (+ sign extension, saturating truncation, bulk
// memory copy / fill) follows; it validates nothing beyond what it needs to run safely.
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "sdfgpu.h"
#include "sdfgpu_tape.h"

namespace {

enum : uint8_t { T_I32 = 0, T_I64 = 1, T_F32 = 2, T_F64 = 3 };

struct Val {
    uint8_t ty = T_I32;
    bool sym = false;
    uint64_t bits = 0;     // concrete value (i32 / f32 in the low word)
    uint32_t node = 0;     // symbolic: index of the SSA node (of the low word for a 64-bit value)
    uint32_t node_hi = 0;  // symbolic i64 / f64: the node of the high word.  Such a value can only be moved
                           // (loaded, stored, kept in locals, selected): compilers copy structs of floats that way
};

struct FuncType {
    std::vector<uint8_t> params, results;
};

struct BlockInfo {
    size_t else_pc = 0;  // position of the `else` opcode, 0 if none
    size_t end_pc = 0;   // position of the matching `end` opcode
};

struct Func {
    uint32_t type = 0;
    bool imported = false;
    std::string import_name;
    const uint8_t* code = nullptr;  // the expression (after the locals declaration)
    size_t code_len = 0;
    std::vector<uint8_t> local_types;  // declared locals (params excluded)
    bool scanned = false;
    std::unordered_map<size_t, BlockInfo> blocks;  // pc of block / loop / if opcode -> its else / end
};

struct Global {
    Val v;
    bool mut = false;
};

struct Module {
    std::vector<FuncType> types;
    std::vector<Func> funcs;
    std::vector<int64_t> table;  // function index or -1
    std::vector<Global> globals;
    uint32_t mem_pages = 0, mem_max = 65536;
    bool has_memory = false;
    std::vector<uint8_t> base;  // linear memory as instantiated (+ what the concrete set-up calls wrote)
    std::map<std::string, std::pair<uint8_t, uint32_t>> exports;  // name -> (kind, index)
    int64_t start = -1;
    std::vector<std::pair<const uint8_t*, size_t>> passive_data;
};

struct Label {
    bool is_loop = false;
    size_t cont_pc = 0;  // where a branch to this label continues
    size_t end_pc = 0;   // position after the construct's `end`
    size_t height = 0;   // value-stack height at entry (below the parameters)
    uint32_t arity = 0;  // values a branch carries
    uint32_t results = 0;
};

struct Frame {
    uint32_t func = 0;
    size_t pc = 0;
    std::vector<Val> locals;
    std::vector<Label> labels;  // labels[0] is the function body
};

struct Cell {  // one aligned 32-bit word of guest memory written since instantiation
    bool sym = false;
    uint32_t w = 0;  // concrete bytes (little endian) or the node
};

struct State {
    std::vector<Val> stack;
    std::vector<Frame> frames;
    std::vector<Val> globals;
    std::map<uint32_t, Cell> mem;  // overlay over Module::base, keyed by the word's address
    uint32_t pages = 0;
    // Trap predicate of this path: a node that is non-zero at the positions where the guest would have trapped
    // before reaching this point (a side of a position-dependent branch that ran into `unreachable` / a failed
    // bounds check, a zero divisor, a truncation out of range); NO_TRAP = never.  The reference answers a trapped
    // sample() call with SDFSample::new(1.0, 0) (src/sdf/wasm/native.rs:196-203): the outputs select that there.
    uint32_t trap = 0xffffffffu;
};
constexpr uint32_t NO_TRAP = 0xffffffffu;

struct Node {
    uint32_t op, a, b, c;
};

enum Status { ST_OK = 0, ST_TRAP = 1, ST_FAIL = 2, ST_REJOIN = 3 };  // ST_REJOIN: the path stopped at the join point it was given

struct Leaf {
    Status st = ST_OK;
    std::vector<Val> vals;
    uint32_t trap = NO_TRAP;  // State::trap of the path that produced the values
};

struct Lowerer {
    Module m;
    std::vector<Node> nodes;
    std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t>, uint32_t> cse;
    std::vector<uint32_t> const_bits;  // f32 constants referenced by SDFT_S_CONST nodes (index = node.a)
    std::map<uint32_t, uint32_t> const_index;
    std::string err;
    uint64_t budget = 400u * 1000u * 1000u;  // instructions, over all paths
    uint32_t leaves = 0;
    uint32_t fork_depth = 0;  // symbolic branches open on the current path (run() recurses once per branch)
    bool malformed = false;
    std::map<uint32_t, int> libm_class;  // function index -> 0 unknown, 1 behaves exactly like C fmodf
    uint32_t recognised_calls = 0;
    uint32_t skipped_imports = 0;  // calls of void host imports that were skipped
    uint32_t trap_sides = 0;       // sides of position-dependent branches that end in a trap (folded into the trap predicate)
    uint32_t trap_ops = 0;         // position-dependent divisions / truncations whose trap condition joined the predicate

    bool fail(const char* fmt, ...) {
        if (err.empty()) {
            char buf[512];
            va_list ap;
            va_start(ap, fmt);
            vsnprintf(buf, sizeof buf, fmt, ap);
            va_end(ap);
            err = buf;
        }
        return false;
    }

    // ------------------------------------------------------------------ SSA
    uint32_t node(uint32_t op, uint32_t a = 0, uint32_t b = 0, uint32_t c = 0) {
        const auto key = std::make_tuple(op, a, b, c);
        auto it = cse.find(key);
        if (it != cse.end()) return it->second;
        nodes.push_back(Node{op, a, b, c});
        return cse[key] = (uint32_t)nodes.size() - 1;
    }
    uint32_t node_of(const Val& v) {
        if (v.sym) return v.node;
        const uint32_t w = (uint32_t)v.bits;
        if (v.ty == T_F32) {
            auto it = const_index.find(w);
            uint32_t k;
            if (it == const_index.end()) {
                k = (uint32_t)const_bits.size();
                const_bits.push_back(w);
                const_index[w] = k;
            } else {
                k = it->second;
            }
            return node(SDFT_S_CONST, k);
        }
        return node(SDFT_S_IMM, w);
    }
    static Val symv(uint8_t ty, uint32_t n) {
        Val v;
        v.ty = ty; v.sym = true; v.node = n;
        return v;
    }
    static Val sym64(uint8_t ty, uint32_t lo, uint32_t hi) {
        Val v;
        v.ty = ty; v.sym = true; v.node = lo; v.node_hi = hi;
        return v;
    }
    static bool is64(uint8_t ty) { return ty == T_I64 || ty == T_F64; }
    // the node of one half of a 64-bit value (a concrete half is data of the tape, like any memory word)
    uint32_t half_node(const Val& v, int hi) {
        if (v.sym) return hi ? v.node_hi : v.node;
        return node_of(conc(T_F32, hi ? (v.bits >> 32) : (v.bits & 0xffffffffull)));
    }
    static Val conc(uint8_t ty, uint64_t bits) {
        Val v;
        v.ty = ty; v.bits = (ty == T_I32 || ty == T_F32) ? (bits & 0xffffffffull) : bits;
        return v;
    }

    // --------------------------------------------------------------- decoder
    struct Rd {
        const uint8_t* p;
        const uint8_t* end;
        bool ok = true;
        uint8_t u8() {
            if (p >= end) { ok = false; return 0; }
            return *p++;
        }
        uint64_t uleb() {
            uint64_t r = 0;
            for (int shift = 0; shift < 70; shift += 7) {
                const uint8_t b = u8();
                r |= (uint64_t)(b & 0x7f) << (shift < 64 ? shift : 63);
                if (!(b & 0x80)) return r;
                if (!ok) return 0;
            }
            ok = false;
            return 0;
        }
        int64_t sleb() {
            uint64_t r = 0;
            int shift = 0;
            uint8_t b;
            do {
                b = u8();
                if (shift < 64) r |= (uint64_t)(b & 0x7f) << shift;
                shift += 7;
            } while ((b & 0x80) && ok && shift < 77);
            if (shift < 64 && (b & 0x40)) r |= ~(uint64_t)0 << shift;
            return (int64_t)r;
        }
        uint32_t u32() { return (uint32_t)uleb(); }
        std::string name() {
            const uint32_t n = u32();
            if (!ok || (size_t)(end - p) < n) { ok = false; return ""; }
            std::string s((const char*)p, n);
            p += n;
            return s;
        }
    };

    static bool valtype(uint8_t b, uint8_t* out) {
        switch (b) {
            case 0x7f: *out = T_I32; return true;
            case 0x7e: *out = T_I64; return true;
            case 0x7d: *out = T_F32; return true;
            case 0x7c: *out = T_F64; return true;
            default: return false;
        }
    }

    bool const_expr(Rd& r, Val* out) {  // i32.const / i64.const / f32.const / f64.const / global.get, then end
        const uint8_t op = r.u8();
        switch (op) {
            case 0x41: *out = conc(T_I32, (uint64_t)r.sleb()); break;
            case 0x42: *out = conc(T_I64, (uint64_t)r.sleb()); break;
            case 0x43: { uint32_t w = 0; for (int i = 0; i < 4; ++i) w |= (uint32_t)r.u8() << (8 * i); *out = conc(T_F32, w); break; }
            case 0x44: { uint64_t w = 0; for (int i = 0; i < 8; ++i) w |= (uint64_t)r.u8() << (8 * i); *out = conc(T_F64, w); break; }
            case 0x23: { const uint32_t g = r.u32(); if (g >= m.globals.size()) return false; *out = m.globals[g].v; break; }
            default: return false;
        }
        return r.u8() == 0x0b && r.ok;
    }

    bool parse(const uint8_t* bytes, size_t len) {
        malformed = true;
        if (len < 8 || memcmp(bytes, "\0asm", 4) != 0 || bytes[4] != 1 || bytes[5] || bytes[6] || bytes[7])
            return fail("not a WebAssembly 1.0 binary module");
        Rd r{bytes + 8, bytes + len};
        std::vector<uint32_t> func_types;
        size_t n_imported_funcs = 0;
        while (r.p < r.end) {
            const uint8_t id = r.u8();
            const uint32_t size = r.u32();
            if (!r.ok || (size_t)(r.end - r.p) < size) return fail("truncated section %u", id);
            Rd s{r.p, r.p + size};
            r.p += size;
            switch (id) {
                case 1: {  // types
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        if (s.u8() != 0x60) return fail("unsupported type form");
                        FuncType ft;
                        uint32_t np = s.u32();
                        for (uint32_t k = 0; k < np && s.ok; ++k) { uint8_t t; if (!valtype(s.u8(), &t)) return fail("unsupported parameter type"); ft.params.push_back(t); }
                        uint32_t nr = s.u32();
                        for (uint32_t k = 0; k < nr && s.ok; ++k) { uint8_t t; if (!valtype(s.u8(), &t)) return fail("unsupported result type"); ft.results.push_back(t); }
                        m.types.push_back(ft);
                    }
                    break;
                }
                case 2: {  // imports
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        const std::string mod = s.name(), nm = s.name();
                        const uint8_t kind = s.u8();
                        if (kind == 0) {
                            Func f;
                            f.type = s.u32(); f.imported = true; f.import_name = mod + "." + nm;
                            m.funcs.push_back(f);
                            ++n_imported_funcs;
                        } else if (kind == 1) {
                            s.u8(); const uint8_t fl = s.u8(); s.u32(); if (fl & 1) s.u32();
                            return fail("imported tables are not supported");
                        } else if (kind == 2) {
                            return fail("imported memories are not supported");
                        } else if (kind == 3) {
                            return fail("imported globals are not supported");
                        } else {
                            return fail("unknown import kind %u", kind);
                        }
                    }
                    break;
                }
                case 3: {
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) func_types.push_back(s.u32());
                    break;
                }
                case 4: {  // tables
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        s.u8();
                        const uint8_t fl = s.u8();
                        const uint32_t mn = s.u32();
                        if (fl & 1) s.u32();
                        if (i == 0) { if (mn > (1u << 20)) return fail("table too large"); m.table.assign(mn, -1); }
                    }
                    break;
                }
                case 5: {  // memories
                    const uint32_t n = s.u32();
                    if (n > 1) return fail("multiple memories are not supported");
                    if (n == 1) {
                        const uint8_t fl = s.u8();
                        m.mem_pages = s.u32();
                        if (fl & 1) m.mem_max = s.u32();
                        if (fl & ~1u) return fail("shared / 64-bit memories are not supported");
                        if (m.mem_pages > 4096) return fail("initial memory above 256 MiB");
                        m.has_memory = true;
                    }
                    break;
                }
                case 6: {  // globals
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        uint8_t t;
                        if (!valtype(s.u8(), &t)) return fail("unsupported global type");
                        Global g;
                        g.mut = s.u8() != 0;
                        if (!const_expr(s, &g.v)) return fail("unsupported global initialiser");
                        g.v.ty = t;
                        m.globals.push_back(g);
                    }
                    break;
                }
                case 7: {
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        const std::string nm = s.name();
                        const uint8_t kind = s.u8();
                        const uint32_t idx = s.u32();
                        m.exports[nm] = std::make_pair(kind, idx);
                    }
                    break;
                }
                case 8: m.start = s.u32(); break;
                case 9: {  // elements: active segments of function indices only
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        const uint32_t flag = s.u32();
                        if (flag != 0 && flag != 2) return fail("element segment kind %u is not supported", flag);
                        if (flag == 2 && s.u32() != 0) return fail("only table 0 is supported");
                        Val off;
                        if (!const_expr(s, &off) || off.ty != T_I32) return fail("unsupported element offset");
                        if (flag == 2 && s.u8() != 0) return fail("unsupported element kind");
                        const uint32_t cnt = s.u32();
                        for (uint32_t k = 0; k < cnt && s.ok; ++k) {
                            const uint32_t f = s.u32();
                            const uint64_t at = (uint64_t)(uint32_t)off.bits + k;
                            if (at >= m.table.size()) return fail("element segment outside the table");
                            m.table[at] = f;
                        }
                    }
                    break;
                }
                case 10: {  // code
                    const uint32_t n = s.u32();
                    if (n != func_types.size()) return fail("function and code sections disagree");
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        const uint32_t body = s.u32();
                        if (!s.ok || (size_t)(s.end - s.p) < body) return fail("truncated function body");
                        Rd b{s.p, s.p + body};
                        s.p += body;
                        Func f;
                        f.type = func_types[i];
                        const uint32_t groups = b.u32();
                        for (uint32_t g = 0; g < groups && b.ok; ++g) {
                            const uint32_t cnt = b.u32();
                            uint8_t t;
                            if (!valtype(b.u8(), &t)) return fail("unsupported local type");
                            if (cnt > 50000 || f.local_types.size() + cnt > 50000) return fail("too many locals");
                            f.local_types.insert(f.local_types.end(), cnt, t);
                        }
                        if (!b.ok) return fail("truncated locals");
                        f.code = b.p;
                        f.code_len = (size_t)(b.end - b.p);
                        m.funcs.push_back(f);
                    }
                    break;
                }
                case 11: {  // data
                    const uint32_t n = s.u32();
                    for (uint32_t i = 0; i < n && s.ok; ++i) {
                        const uint32_t flag = s.u32();
                        if (flag == 1) {
                            const uint32_t cnt = s.u32();
                            if ((size_t)(s.end - s.p) < cnt) return fail("truncated data segment");
                            m.passive_data.push_back(std::make_pair(s.p, (size_t)cnt));
                            s.p += cnt;
                            continue;
                        }
                        if (flag == 2 && s.u32() != 0) return fail("only memory 0 is supported");
                        if (flag > 2) return fail("unknown data segment kind");
                        Val off;
                        if (!const_expr(s, &off) || off.ty != T_I32) return fail("unsupported data offset");
                        const uint32_t cnt = s.u32();
                        if (!s.ok || (size_t)(s.end - s.p) < cnt) return fail("truncated data segment");
                        const uint64_t at = (uint32_t)off.bits;
                        if (at + cnt > (uint64_t)m.mem_pages * 65536) return fail("data segment outside the memory");
                        if (m.base.size() < at + cnt) m.base.resize(at + cnt, 0);
                        memcpy(m.base.data() + at, s.p, cnt);
                        s.p += cnt;
                        m.passive_data.push_back(std::make_pair((const uint8_t*)nullptr, (size_t)0));
                    }
                    break;
                }
                default: break;  // custom (0), data count (12), anything newer: skipped
            }
            if (!s.ok) return fail("malformed section %u", id);
        }
        for (const Func& f : m.funcs)
            if (f.type >= m.types.size()) return fail("function type out of range");
        (void)n_imported_funcs;
        malformed = false;
        return true;
    }

    // length of the immediates of the instruction whose opcode was just read; false = unknown instruction
    bool skip_immediates(uint8_t op, Rd& r) {
        switch (op) {
            case 0x02: case 0x03: case 0x04: r.sleb(); return true;
            case 0x0c: case 0x0d: r.u32(); return true;
            case 0x0e: { const uint32_t n = r.u32(); for (uint32_t i = 0; i <= n && r.ok; ++i) r.u32(); return true; }
            case 0x10: r.u32(); return true;
            case 0x11: r.u32(); r.u32(); return true;
            case 0x1c: { const uint32_t n = r.u32(); for (uint32_t i = 0; i < n && r.ok; ++i) r.u8(); return true; }
            case 0x20: case 0x21: case 0x22: case 0x23: case 0x24: case 0x25: case 0x26: r.u32(); return true;
            case 0x3f: case 0x40: r.u8(); return true;
            case 0x41: r.sleb(); return true;
            case 0x42: r.sleb(); return true;
            case 0x43: for (int i = 0; i < 4; ++i) r.u8(); return true;
            case 0x44: for (int i = 0; i < 8; ++i) r.u8(); return true;
            case 0xd0: r.u8(); return true;
            case 0xd2: r.u32(); return true;
            case 0xfc: {
                const uint32_t sub = r.u32();
                switch (sub) {
                    case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: return true;
                    case 8: r.u32(); r.u8(); return true;
                    case 9: r.u32(); return true;
                    case 10: r.u8(); r.u8(); return true;
                    case 11: r.u8(); return true;
                    case 12: case 14: r.u32(); r.u32(); return true;
                    case 13: case 15: case 16: case 17: r.u32(); return true;
                    default: return false;
                }
            }
            case 0xfd: case 0xfe: return false;  // SIMD, threads
            default:
                if (op >= 0x28 && op <= 0x3e) { r.u32(); r.u32(); return true; }
                return true;
        }
    }

    bool scan(Func& f) {
        if (f.scanned) return true;
        Rd r{f.code, f.code + f.code_len};
        std::vector<size_t> open;
        while (r.p < r.end) {
            const size_t pc = (size_t)(r.p - f.code);
            const uint8_t op = r.u8();
            if (op == 0x02 || op == 0x03 || op == 0x04) {
                open.push_back(pc);
                f.blocks[pc] = BlockInfo();
            } else if (op == 0x05) {
                if (open.empty()) return fail("else without if");
                f.blocks[open.back()].else_pc = pc;
            } else if (op == 0x0b) {
                if (open.empty()) {  // the function body's end
                    if (r.p != r.end) return fail("code after the function's end");
                    break;
                }
                f.blocks[open.back()].end_pc = pc;
                open.pop_back();
            }
            if (!skip_immediates(op, r) || !r.ok) return fail("unsupported instruction 0x%02x", op);
        }
        if (!open.empty()) return fail("unterminated block");
        f.scanned = true;
        return true;
    }

    bool block_arity(int64_t bt, uint32_t* params, uint32_t* results) {
        if (bt == -64) { *params = 0; *results = 0; return true; }  // 0x40: empty
        if (bt < 0) { *params = 0; *results = 1; return true; }     // a value type
        if ((uint64_t)bt >= m.types.size()) return fail("block type out of range");
        *params = (uint32_t)m.types[bt].params.size();
        *results = (uint32_t)m.types[bt].results.size();
        return true;
    }

    // ---------------------------------------------------------------- memory
    bool in_bounds(const State& st, uint64_t addr, uint32_t n) { return addr + n <= (uint64_t)st.pages * 65536; }

    uint32_t base_word(uint32_t wa) const {
        uint32_t w = 0;
        for (int i = 0; i < 4; ++i)
            if ((size_t)wa + i < m.base.size()) w |= (uint32_t)m.base[wa + i] << (8 * i);
        return w;
    }
    // 0 ok, 1 the byte belongs to a symbolic word
    int read_byte(const State& st, uint32_t addr, uint8_t* out) {
        auto it = st.mem.find(addr & ~3u);
        if (it != st.mem.end()) {
            if (it->second.sym) return 1;
            *out = (uint8_t)(it->second.w >> (8 * (addr & 3u)));
            return 0;
        }
        *out = addr < m.base.size() ? m.base[addr] : 0;
        return 0;
    }
    // concrete bytes into guest memory.  A symbolic word may be overwritten as a whole (zeroing a struct, copying
    // over a spilled value); overwriting part of one has no 32-bit form and fails.
    bool write_bytes(State& st, uint32_t addr, const uint8_t* data, uint32_t n) {
        uint32_t i = 0;
        while (i < n) {
            const uint32_t a = addr + i, wa = a & ~3u;
            auto it = st.mem.find(wa);
            if (a == wa && n - i >= 4) {  // a whole word
                const uint32_t w = (uint32_t)data[i] | (uint32_t)data[i + 1] << 8 | (uint32_t)data[i + 2] << 16 | (uint32_t)data[i + 3] << 24;
                if (it == st.mem.end()) st.mem.insert(std::make_pair(wa, Cell{false, w}));
                else it->second = Cell{false, w};
                i += 4;
                continue;
            }
            if (it == st.mem.end()) it = st.mem.insert(std::make_pair(wa, Cell{false, base_word(wa)})).first;
            if (it->second.sym) return fail("a store overwrites part of a symbolic word at 0x%x", wa);
            const int sh = 8 * (int)(a & 3u);
            it->second.w = (it->second.w & ~(0xffu << sh)) | ((uint32_t)data[i] << sh);
            ++i;
        }
        return true;
    }
    Cell cell_at(const State& st, uint32_t wa) const {
        auto it = st.mem.find(wa);
        return it != st.mem.end() ? it->second : Cell{false, base_word(wa)};
    }
    // an aligned 64-bit load of which at least one word is symbolic: the two halves' nodes
    bool load_pair(State& st, uint64_t addr, uint32_t* lo, uint32_t* hi) {
        if (addr & 3u) return false;
        const Cell a = cell_at(st, (uint32_t)addr), b = cell_at(st, (uint32_t)addr + 4);
        if (!a.sym && !b.sym) return false;
        *lo = a.sym ? a.w : node_of(conc(T_F32, a.w));
        *hi = b.sym ? b.w : node_of(conc(T_F32, b.w));
        return true;
    }
    bool load(State& st, uint64_t addr, uint32_t nbytes, uint64_t* out, bool* is_sym, uint32_t* sym_node) {
        *is_sym = false;
        if (nbytes == 4 && (addr & 3u) == 0) {
            auto it = st.mem.find((uint32_t)addr);
            if (it != st.mem.end() && it->second.sym) { *is_sym = true; *sym_node = it->second.w; return true; }
        }
        uint64_t v = 0;
        for (uint32_t i = 0; i < nbytes; ++i) {
            uint8_t b;
            if (read_byte(st, (uint32_t)(addr + i), &b)) return fail("a load of %u bytes at 0x%x overlaps a symbolic word", nbytes, (uint32_t)addr);
            v |= (uint64_t)b << (8 * i);
        }
        *out = v;
        return true;
    }
    bool store(State& st, uint64_t addr, uint32_t nbytes, const Val& v) {
        if (v.sym && is64(v.ty) && nbytes == 8 && (addr & 3u) == 0) {  // a moved pair of words
            st.mem[(uint32_t)addr] = Cell{true, v.node};
            st.mem[(uint32_t)addr + 4] = Cell{true, v.node_hi};
            return true;
        }
        if (v.sym) {
            if (nbytes != 4 || (addr & 3u) || is64(v.ty))
                return fail("a symbolic value is stored with %u bytes at 0x%x (only aligned 32- and 64-bit stores are lowered)", nbytes, (uint32_t)addr);
            st.mem[(uint32_t)addr] = Cell{true, v.node};
            return true;
        }
        if (nbytes == 4 && (addr & 3u) == 0) {
            st.mem[(uint32_t)addr] = Cell{false, (uint32_t)v.bits};
            return true;
        }
        uint8_t bytes[8];
        for (uint32_t i = 0; i < nbytes && i < 8; ++i) bytes[i] = (uint8_t)(v.bits >> (8 * i));
        return write_bytes(st, (uint32_t)addr, bytes, nbytes);
    }
    void commit(State& st) {  // fold a finished CONCRETE set-up call into the instantiated memory
        for (auto& kv : st.mem) {
            if (kv.second.sym) continue;
            if (m.base.size() < (size_t)kv.first + 4) m.base.resize((size_t)kv.first + 4, 0);
            for (int i = 0; i < 4; ++i) m.base[kv.first + i] = (uint8_t)(kv.second.w >> (8 * i));
        }
        st.mem.clear();
        for (size_t i = 0; i < st.globals.size() && i < m.globals.size(); ++i) m.globals[i].v = st.globals[i];
        m.mem_pages = st.pages;
    }

    // ----------------------------------------------------------- numerics
    static float f32_of(uint64_t bits) { const uint32_t w = (uint32_t)bits; float f; memcpy(&f, &w, 4); return f; }
    static uint64_t bits_of(float f) { uint32_t w; memcpy(&w, &f, 4); return w; }
    static double f64_of(uint64_t bits) { double d; memcpy(&d, &bits, 8); return d; }
    static uint64_t bits_of(double d) { uint64_t w; memcpy(&w, &d, 8); return w; }
    static float wmin(float a, float b) {
        if (a != a || b != b) return f32_of(0x7fc00000u);
        if (a == b) return f32_of(bits_of(a) | bits_of(b));
        return a < b ? a : b;
    }
    static float wmax(float a, float b) {
        if (a != a || b != b) return f32_of(0x7fc00000u);
        if (a == b) return f32_of(bits_of(a) & bits_of(b));
        return a > b ? a : b;
    }
    static double wmin(double a, double b) {
        if (a != a || b != b) return f64_of(0x7ff8000000000000ull);
        if (a == b) return f64_of(bits_of(a) | bits_of(b));
        return a < b ? a : b;
    }
    static double wmax(double a, double b) {
        if (a != a || b != b) return f64_of(0x7ff8000000000000ull);
        if (a == b) return f64_of(bits_of(a) & bits_of(b));
        return a > b ? a : b;
    }
    static int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
    static int ctz32(uint32_t x) { return x ? __builtin_ctz(x) : 32; }
    static int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
    static int ctz64(uint64_t x) { return x ? __builtin_ctzll(x) : 64; }

    // trunc of a float to an integer range; sat = saturating form.  Returns false on a trap.
    template <class F>
    static bool trunc_to(F f, double lo, double hi_excl, bool sat, bool is_signed, int bits, uint64_t* out) {
        if (f != f) { if (!sat) return false; *out = 0; return true; }
        const double d = std::trunc((double)f);
        if (d < lo || d >= hi_excl) {
            if (!sat) return false;
            if (d < lo) *out = is_signed ? (bits == 32 ? (uint64_t)(uint32_t)INT32_MIN : (uint64_t)INT64_MIN) : 0;
            else *out = is_signed ? (bits == 32 ? (uint64_t)INT32_MAX : (uint64_t)INT64_MAX) : (bits == 32 ? 0xffffffffull : ~0ull);
            return true;
        }
        if (is_signed) *out = (uint64_t)(int64_t)d;
        else *out = (uint64_t)d;
        if (bits == 32) *out &= 0xffffffffull;
        return true;
    }

    // symbolic form of a numeric instruction, or 0xffffffff when it has none
    static uint32_t sym_unop(uint8_t op) {
        switch (op) {
            case 0x45: return SDFT_S_IEQZ;
            case 0x8b: return SDFT_S_FABS;
            case 0x8c: return SDFT_S_FNEG;
            case 0x8d: return SDFT_S_FCEIL;
            case 0x8e: return SDFT_S_FFLOOR;
            case 0x8f: return SDFT_S_FTRUNC;
            case 0x90: return SDFT_S_FNEAREST;
            case 0x91: return SDFT_S_FSQRT;
            case 0xb2: return SDFT_S_F_FROM_I_S;
            case 0xb3: return SDFT_S_F_FROM_I_U;
            default: return 0xffffffffu;
        }
    }
    static uint32_t sym_binop(uint8_t op) {
        switch (op) {
            case 0x46: return SDFT_S_IEQ;
            case 0x47: return SDFT_S_INE;
            case 0x48: return SDFT_S_ILT_S;
            case 0x49: return SDFT_S_ILT_U;
            case 0x4a: return SDFT_S_IGT_S;
            case 0x4b: return SDFT_S_IGT_U;
            case 0x4c: return SDFT_S_ILE_S;
            case 0x4d: return SDFT_S_ILE_U;
            case 0x4e: return SDFT_S_IGE_S;
            case 0x4f: return SDFT_S_IGE_U;
            case 0x5b: return SDFT_S_FEQ;
            case 0x5c: return SDFT_S_FNE;
            case 0x5d: return SDFT_S_FLT;
            case 0x5e: return SDFT_S_FGT;
            case 0x5f: return SDFT_S_FLE;
            case 0x60: return SDFT_S_FGE;
            case 0x6a: return SDFT_S_IADD;
            case 0x6b: return SDFT_S_ISUB;
            case 0x6c: return SDFT_S_IMUL;
            case 0x6d: return SDFT_S_IDIV_S;  // the guest would trap on a zero divisor / overflow: the op yields 0 there
            case 0x6e: return SDFT_S_IDIV_U;
            case 0x6f: return SDFT_S_IREM_S;
            case 0x70: return SDFT_S_IREM_U;
            case 0x71: return SDFT_S_IAND;
            case 0x72: return SDFT_S_IOR;
            case 0x73: return SDFT_S_IXOR;
            case 0x74: return SDFT_S_ISHL;
            case 0x75: return SDFT_S_ISHR_S;
            case 0x76: return SDFT_S_ISHR_U;
            case 0x92: return SDFT_S_FADD;
            case 0x93: return SDFT_S_FSUB;
            case 0x94: return SDFT_S_FMUL;
            case 0x95: return SDFT_S_FDIV;
            case 0x96: return SDFT_S_FMIN;
            case 0x97: return SDFT_S_FMAX;
            case 0x98: return SDFT_S_FCOPYSIGN;
            default: return 0xffffffffu;
        }
    }

    static int sop_inputs(uint32_t op) {
        if (op <= SDFT_S_IMM) return 0;
        if (op == SDFT_S_SELECT) return 3;
        if ((op >= SDFT_S_FNEG && op <= SDFT_S_FNEAREST) || op == SDFT_S_IEQZ || (op >= SDFT_S_F_FROM_I_S && op <= SDFT_S_I_FROM_F_U))
            return 1;
        return 2;
    }

    // result type of a numeric instruction (by opcode ranges of the MVP)
    static uint8_t result_type(uint8_t op) {
        if (op >= 0x45 && op <= 0x66) return T_I32;                        // tests and comparisons
        if (op >= 0x67 && op <= 0x78) return T_I32;
        if (op >= 0x79 && op <= 0x8a) return T_I64;
        if (op >= 0x8b && op <= 0x98) return T_F32;
        if (op >= 0x99 && op <= 0xa6) return T_F64;
        switch (op) {
            case 0xa7: case 0xa8: case 0xa9: case 0xaa: case 0xab: case 0xbc: case 0xc0: case 0xc1: return T_I32;
            case 0xac: case 0xad: case 0xae: case 0xaf: case 0xb0: case 0xb1: case 0xbd: case 0xc2: case 0xc3: case 0xc4: return T_I64;
            case 0xb2: case 0xb3: case 0xb4: case 0xb5: case 0xb6: case 0xbe: return T_F32;
            default: return T_F64;
        }
    }

    // concrete unary numeric instruction; returns ST_TRAP on a trapping conversion
    Status unop(uint8_t op, const Val& a, Val* out) {
        const uint32_t x = (uint32_t)a.bits;
        const uint64_t X = a.bits;
        const float f = f32_of(a.bits);
        const double d = f64_of(a.bits);
        uint64_t r = 0;
        switch (op) {
            case 0x45: r = x == 0; break;
            case 0x50: r = X == 0; break;
            case 0x67: r = (uint32_t)clz32(x); break;
            case 0x68: r = (uint32_t)ctz32(x); break;
            case 0x69: r = (uint32_t)__builtin_popcount(x); break;
            case 0x79: r = (uint64_t)clz64(X); break;
            case 0x7a: r = (uint64_t)ctz64(X); break;
            case 0x7b: r = (uint64_t)__builtin_popcountll(X); break;
            case 0x8b: r = x & 0x7fffffffu; break;
            case 0x8c: r = x ^ 0x80000000u; break;
            case 0x8d: r = bits_of(ceilf(f)); break;
            case 0x8e: r = bits_of(floorf(f)); break;
            case 0x8f: r = bits_of(truncf(f)); break;
            case 0x90: r = bits_of(nearbyintf(f)); break;
            case 0x91: r = bits_of(sqrtf(f)); break;
            case 0x99: r = X & 0x7fffffffffffffffull; break;
            case 0x9a: r = X ^ 0x8000000000000000ull; break;
            case 0x9b: r = bits_of(std::ceil(d)); break;
            case 0x9c: r = bits_of(std::floor(d)); break;
            case 0x9d: r = bits_of(std::trunc(d)); break;
            case 0x9e: r = bits_of(std::nearbyint(d)); break;
            case 0x9f: r = bits_of(std::sqrt(d)); break;
            case 0xa7: r = (uint32_t)X; break;                                                                 // i32.wrap_i64
            case 0xa8: if (!trunc_to(f, -2147483648.0, 2147483648.0, false, true, 32, &r)) return ST_TRAP; break;
            case 0xa9: if (!trunc_to(f, 0.0, 4294967296.0, false, false, 32, &r) || f <= -1.0f) return ST_TRAP; break;
            case 0xaa: if (!trunc_to(d, -2147483648.0, 2147483648.0, false, true, 32, &r)) return ST_TRAP; break;
            case 0xab: if (!trunc_to(d, 0.0, 4294967296.0, false, false, 32, &r) || d <= -1.0) return ST_TRAP; break;
            case 0xac: r = (uint64_t)(int64_t)(int32_t)x; break;                                               // i64.extend_i32_s
            case 0xad: r = (uint64_t)x; break;
            case 0xae: if (!trunc_to(f, -9223372036854775808.0, 9223372036854775808.0, false, true, 64, &r)) return ST_TRAP; break;
            case 0xaf: if (!trunc_to(f, 0.0, 18446744073709551616.0, false, false, 64, &r) || f <= -1.0f) return ST_TRAP; break;
            case 0xb0: if (!trunc_to(d, -9223372036854775808.0, 9223372036854775808.0, false, true, 64, &r)) return ST_TRAP; break;
            case 0xb1: if (!trunc_to(d, 0.0, 18446744073709551616.0, false, false, 64, &r) || d <= -1.0) return ST_TRAP; break;
            case 0xb2: r = bits_of((float)(int32_t)x); break;
            case 0xb3: r = bits_of((float)x); break;
            case 0xb4: r = bits_of((float)(int64_t)X); break;
            case 0xb5: r = bits_of((float)X); break;
            case 0xb6: r = bits_of((float)d); break;                                                           // f32.demote_f64
            case 0xb7: r = bits_of((double)(int32_t)x); break;
            case 0xb8: r = bits_of((double)x); break;
            case 0xb9: r = bits_of((double)(int64_t)X); break;
            case 0xba: r = bits_of((double)X); break;
            case 0xbb: r = bits_of((double)f); break;                                                          // f64.promote_f32
            case 0xbc: case 0xbe: r = x; break;                                                                // reinterpret 32
            case 0xbd: case 0xbf: r = X; break;                                                                // reinterpret 64
            case 0xc0: r = (uint32_t)(int32_t)(int8_t)x; break;
            case 0xc1: r = (uint32_t)(int32_t)(int16_t)x; break;
            case 0xc2: r = (uint64_t)(int64_t)(int8_t)X; break;
            case 0xc3: r = (uint64_t)(int64_t)(int16_t)X; break;
            case 0xc4: r = (uint64_t)(int64_t)(int32_t)X; break;
            default: return ST_FAIL;
        }
        *out = conc(result_type(op), r);
        return ST_OK;
    }

    Status binop(uint8_t op, const Val& a, const Val& b, Val* out) {
        const uint32_t x = (uint32_t)a.bits, y = (uint32_t)b.bits;
        const int32_t sx = (int32_t)x, sy = (int32_t)y;
        const uint64_t X = a.bits, Y = b.bits;
        const int64_t SX = (int64_t)X, SY = (int64_t)Y;
        const float f = f32_of(a.bits), g = f32_of(b.bits);
        const double d = f64_of(a.bits), e = f64_of(b.bits);
        uint64_t r = 0;
        switch (op) {
            case 0x46: r = x == y; break;
            case 0x47: r = x != y; break;
            case 0x48: r = sx < sy; break;
            case 0x49: r = x < y; break;
            case 0x4a: r = sx > sy; break;
            case 0x4b: r = x > y; break;
            case 0x4c: r = sx <= sy; break;
            case 0x4d: r = x <= y; break;
            case 0x4e: r = sx >= sy; break;
            case 0x4f: r = x >= y; break;
            case 0x51: r = X == Y; break;
            case 0x52: r = X != Y; break;
            case 0x53: r = SX < SY; break;
            case 0x54: r = X < Y; break;
            case 0x55: r = SX > SY; break;
            case 0x56: r = X > Y; break;
            case 0x57: r = SX <= SY; break;
            case 0x58: r = X <= Y; break;
            case 0x59: r = SX >= SY; break;
            case 0x5a: r = X >= Y; break;
            case 0x5b: r = f == g; break;
            case 0x5c: r = f != g; break;
            case 0x5d: r = f < g; break;
            case 0x5e: r = f > g; break;
            case 0x5f: r = f <= g; break;
            case 0x60: r = f >= g; break;
            case 0x61: r = d == e; break;
            case 0x62: r = d != e; break;
            case 0x63: r = d < e; break;
            case 0x64: r = d > e; break;
            case 0x65: r = d <= e; break;
            case 0x66: r = d >= e; break;
            case 0x6a: r = x + y; break;
            case 0x6b: r = x - y; break;
            case 0x6c: r = x * y; break;
            case 0x6d: if (y == 0 || (sx == INT32_MIN && sy == -1)) return ST_TRAP; r = (uint32_t)(sx / sy); break;
            case 0x6e: if (y == 0) return ST_TRAP; r = x / y; break;
            case 0x6f: if (y == 0) return ST_TRAP; r = (sy == -1) ? 0u : (uint32_t)(sx % sy); break;
            case 0x70: if (y == 0) return ST_TRAP; r = x % y; break;
            case 0x71: r = x & y; break;
            case 0x72: r = x | y; break;
            case 0x73: r = x ^ y; break;
            case 0x74: r = x << (y & 31); break;
            case 0x75: r = (uint32_t)(sx >> (y & 31)); break;
            case 0x76: r = x >> (y & 31); break;
            case 0x77: r = (x << (y & 31)) | (x >> ((32 - (y & 31)) & 31)); break;
            case 0x78: r = (x >> (y & 31)) | (x << ((32 - (y & 31)) & 31)); break;
            case 0x7c: r = X + Y; break;
            case 0x7d: r = X - Y; break;
            case 0x7e: r = X * Y; break;
            case 0x7f: if (Y == 0 || (SX == INT64_MIN && SY == -1)) return ST_TRAP; r = (uint64_t)(SX / SY); break;
            case 0x80: if (Y == 0) return ST_TRAP; r = X / Y; break;
            case 0x81: if (Y == 0) return ST_TRAP; r = (SY == -1) ? 0ull : (uint64_t)(SX % SY); break;
            case 0x82: if (Y == 0) return ST_TRAP; r = X % Y; break;
            case 0x83: r = X & Y; break;
            case 0x84: r = X | Y; break;
            case 0x85: r = X ^ Y; break;
            case 0x86: r = X << (Y & 63); break;
            case 0x87: r = (uint64_t)(SX >> (Y & 63)); break;
            case 0x88: r = X >> (Y & 63); break;
            case 0x89: r = (X << (Y & 63)) | (X >> ((64 - (Y & 63)) & 63)); break;
            case 0x8a: r = (X >> (Y & 63)) | (X << ((64 - (Y & 63)) & 63)); break;
            case 0x92: r = bits_of(f + g); break;
            case 0x93: r = bits_of(f - g); break;
            case 0x94: r = bits_of(f * g); break;
            case 0x95: r = bits_of(f / g); break;
            case 0x96: r = bits_of(wmin(f, g)); break;
            case 0x97: r = bits_of(wmax(f, g)); break;
            case 0x98: r = (x & 0x7fffffffu) | (y & 0x80000000u); break;
            case 0xa0: r = bits_of(d + e); break;
            case 0xa1: r = bits_of(d - e); break;
            case 0xa2: r = bits_of(d * e); break;
            case 0xa3: r = bits_of(d / e); break;
            case 0xa4: r = bits_of(wmin(d, e)); break;
            case 0xa5: r = bits_of(wmax(d, e)); break;
            case 0xa6: r = (X & 0x7fffffffffffffffull) | (Y & 0x8000000000000000ull); break;
            default: return ST_FAIL;
        }
        *out = conc(result_type(op), r);
        return ST_OK;
    }

    static bool is_unop(uint8_t op) {
        return op == 0x45 || op == 0x50 || (op >= 0x67 && op <= 0x69) || (op >= 0x79 && op <= 0x7b) ||
               (op >= 0x8b && op <= 0x91) || (op >= 0x99 && op <= 0x9f) || (op >= 0xa7 && op <= 0xc4);
    }

    // ------------------------------------------------------------ interpreter
    bool enter(State& st, uint32_t fi) {  // arguments are on the stack
        if (fi >= m.funcs.size()) return fail("call of function %u, which does not exist", fi);
        Func& f = m.funcs[fi];
        if (f.imported) {
            // a host function that returns nothing (logging, tracing hooks) cannot influence the result: skipped.
            // One that returns a value would have to be answered by the host: not lowerable, except WASI (below).
            const FuncType& it = m.types[f.type];
            // WASI calls (a wasm32-wasi guest seeds its HashMap with random_get, may query the environment or the
            // clock while it sets itself up): answered with errno 0 and untouched out-parameters -- "success, nothing
            // there".  The guest stays self-consistent (it inserts and looks up with the same seeds).
            const bool wasi = f.import_name.compare(0, 5, "wasi_") == 0 && f.import_name.find(".proc_exit") == std::string::npos;
            if (!it.results.empty() && !wasi)
                return fail("the guest calls the host import `%s` on the way to its result", f.import_name.c_str());
            if (st.stack.size() < it.params.size()) return fail("stack underflow at a call");
            st.stack.resize(st.stack.size() - it.params.size());
            for (uint8_t t : it.results) st.stack.push_back(conc(t, 0));
            ++skipped_imports;
            return true;
        }
        if (!scan(f)) return false;
        if (st.frames.size() > 2000) return fail("call depth above 2000");
        const FuncType& ft = m.types[f.type];
        if (st.stack.size() < ft.params.size()) return fail("stack underflow at a call");
        Frame fr;
        fr.func = fi;
        fr.locals.assign(st.stack.end() - ft.params.size(), st.stack.end());
        st.stack.resize(st.stack.size() - ft.params.size());
        for (uint8_t t : f.local_types) fr.locals.push_back(conc(t, 0));
        Label l;
        l.height = st.stack.size();
        l.arity = l.results = (uint32_t)ft.results.size();
        l.cont_pc = l.end_pc = f.code_len;  // branching to the body label returns
        fr.labels.push_back(l);
        st.frames.push_back(std::move(fr));
        return true;
    }

    // carry `arity` values to `height`; false (with the error set) when the stack does not hold them
    bool unwind(State& st, size_t height, uint32_t arity) {
        if (arity > st.stack.size() || height > st.stack.size() - arity) return fail("stack underflow at a branch or return");
        std::vector<Val> keep(st.stack.end() - arity, st.stack.end());
        st.stack.resize(height);
        st.stack.insert(st.stack.end(), keep.begin(), keep.end());
        return true;
    }

    // returns true when the outermost frame returned
    bool branch(State& st, uint32_t depth) {
        Frame& fr = st.frames.back();
        if (depth >= fr.labels.size()) { fail("branch depth out of range"); return false; }
        const size_t idx = fr.labels.size() - 1 - depth;
        const Label l = fr.labels[idx];
        if (idx == 0) return do_return(st);
        if (!unwind(st, l.height, l.arity)) return false;
        if (l.is_loop) {
            fr.labels.resize(idx + 1);
            fr.pc = l.cont_pc;
        } else {
            fr.labels.resize(idx);
            fr.pc = l.end_pc;
        }
        return false;
    }
    bool do_return(State& st) {
        Frame& fr = st.frames.back();
        const Label l = fr.labels[0];
        if (!unwind(st, l.height, l.results)) return false;
        st.frames.pop_back();
        return st.frames.empty();
    }

    typedef Leaf (*FinishFn)(Lowerer&, State&);

    // A guest reaches `%` on floats through its libm's fmodf: an integer loop over the operands' exponents whose
    // trip count depends on the position, so it cannot be unrolled.  But fmodf is an EXACT function, so a callee
    // can be recognised by what it computes: a (f32, f32) -> f32 function that returns C's fmodf bit for bit on a
    // grid of probes (signed zeros, denormals, huge ratios, infinities, NaN) is one, whatever its instructions
    // are, and a call of it with symbolic arguments becomes one SDFT_S_FMOD op.
    // It must also leave nothing behind: a call that is replaced by one op loses whatever else the callee did.  A probe
    // run may write below the caller's shadow-stack pointer (global 0 of a Rust / clang guest: the callee's own frame),
    // nowhere else; it may not change a global, and it may not call a host import.  SDFGPU_WASM_NO_LIBM=1 switches the
    // recognition off altogether (such a guest is then sampled on the host).
    bool behaves_like_fmodf(const State& at, uint32_t fi) {
        auto it = libm_class.find(fi);
        if (it != libm_class.end()) return it->second == 1;
        static const bool disabled = [] { const char* e = getenv("SDFGPU_WASM_NO_LIBM"); return e && *e && *e != '0'; }();
        if (disabled) { libm_class[fi] = 0; return false; }
        static const float xs[] = {0.0f, -0.0f, 1.0f, -1.0f, 0.3f, 5.5f, -7.25f, 1e-3f, 123456.7f, 1e30f, -1e-30f, 1e-40f, 0.75f, 2.5f,
                                   16777216.0f, -3.4e38f, 0.1f};
        static const float ys[] = {0.5f, 0.25f, 1.0f, 3.0f, -2.0f, 0.1f, 1e-40f, 1e30f, 0.0f, -0.0f, 7.0f, 1.17549435e-38f};
        std::vector<std::pair<float, float>> probes;
        for (float x : xs) for (float y : ys) probes.push_back(std::make_pair(x, y));
        const float inf = f32_of(0x7f800000u), nan = f32_of(0x7fc00000u);
        const std::pair<float, float> special[] = {{inf, 1.0f}, {-inf, 2.0f}, {1.0f, inf}, {-2.5f, -inf}, {nan, 1.0f}, {1.0f, nan}, {inf, inf}};
        for (const auto& pr : special) probes.push_back(pr);
        // the probes must not disturb the lowering in progress
        const std::string saved_err = err;
        const uint64_t saved_budget = budget;
        const uint32_t saved_leaves = leaves, saved_depth = fork_depth, saved_imports = skipped_imports;
        const bool have_sp = !at.globals.empty() && at.globals[0].ty == T_I32 && !at.globals[0].sym;
        const uint32_t sp0 = have_sp ? (uint32_t)at.globals[0].bits : 0u;
        bool same = true;
        for (size_t k = 0; k < probes.size() && same; ++k) {
            State t;
            t.globals = at.globals;
            t.mem = at.mem;
            t.pages = at.pages;
            t.stack.push_back(conc(T_F32, bits_of(probes[k].first)));
            t.stack.push_back(conc(T_F32, bits_of(probes[k].second)));
            err.clear();
            budget = 200000;  // a real fmodf needs a few thousand instructions at most
            fork_depth = 0;
            Leaf l;
            if (!enter(t, fi)) { same = false; break; }
            l = run(t, [](Lowerer&, State& s) { Leaf o; o.vals = s.stack; return o; });
            if (l.st != ST_OK || l.vals.size() != 1 || l.vals[0].sym) { same = false; break; }
            const float want = fmodf(probes[k].first, probes[k].second);
            const float got = f32_of(l.vals[0].bits);
            same = (got != got && want != want) || bits_of(got) == bits_of(want);
            // no side effects: host imports, globals, memory outside the callee's own stack frame
            if (skipped_imports != saved_imports || t.globals.size() != at.globals.size()) same = false;
            for (size_t g = 0; g < at.globals.size() && same; ++g)
                if (t.globals[g].sym != at.globals[g].sym || t.globals[g].bits != at.globals[g].bits || t.globals[g].node != at.globals[g].node) same = false;
            for (auto c = t.mem.begin(); c != t.mem.end() && same; ++c) {
                const auto before = at.mem.find(c->first);
                const bool changed = before == at.mem.end() || before->second.sym != c->second.sym || before->second.w != c->second.w;
                if (changed && !(have_sp && c->first < sp0 && sp0 - c->first <= 1024u)) same = false;  // a leaf's frame is small
            }
        }
        skipped_imports = saved_imports;
        err = saved_err;
        budget = saved_budget;
        leaves = saved_leaves;
        fork_depth = saved_depth;
        libm_class[fi] = same ? 1 : 0;
        return same;
    }

    // true when the call was replaced by one op (the two arguments are popped, the result pushed)
    bool call_as_known_function(State& st, uint32_t fi) {
        if (fi >= m.funcs.size() || m.funcs[fi].imported) return false;
        const FuncType& ft = m.types[m.funcs[fi].type];
        if (ft.params != std::vector<uint8_t>{T_F32, T_F32} || ft.results != std::vector<uint8_t>{T_F32} || st.stack.size() < 2) return false;
        const Val& b = st.stack[st.stack.size() - 1];
        const Val& a = st.stack[st.stack.size() - 2];
        if (!a.sym && !b.sym) return false;  // concrete arguments: just run it
        if (!behaves_like_fmodf(st, fi)) return false;
        const uint32_t n = node(SDFT_S_FMOD, node_of(a), node_of(b));
        st.stack.resize(st.stack.size() - 2);
        st.stack.push_back(symv(T_F32, n));
        ++recognised_calls;
        return true;
    }

    // trap predicates: any non-zero word means "trapped", so OR is the bitwise one
    uint32_t trap_or(uint32_t t, uint32_t cond) { return t == NO_TRAP ? cond : node(SDFT_S_IOR, t, cond); }
    uint32_t trap_select(uint32_t cond, uint32_t ta, uint32_t tb) {
        if (ta == tb) return ta;
        const uint32_t zero = node(SDFT_S_IMM, 0);
        return node(SDFT_S_SELECT, cond, ta == NO_TRAP ? zero : ta, tb == NO_TRAP ? zero : tb);
    }

    // `a` is the outcome where cond_node holds, `b` where it does not
    Leaf merge(uint32_t cond_node, const Leaf& a, const Leaf& b) {
        if (a.st == ST_FAIL || b.st == ST_FAIL) { Leaf l; l.st = ST_FAIL; return l; }
        if (a.st == ST_TRAP && b.st == ST_TRAP) return a;
        if (a.st == ST_TRAP) {  // the guest traps where the condition holds: the other side's values, and the predicate
            Leaf l = b;
            l.trap = trap_or(b.trap, cond_node);
            ++trap_sides;
            return l;
        }
        if (b.st == ST_TRAP) {
            Leaf l = a;
            l.trap = trap_or(a.trap, node(SDFT_S_IEQZ, cond_node));
            ++trap_sides;
            return l;
        }
        Leaf out;
        out.trap = trap_select(cond_node, a.trap, b.trap);
        out.vals.resize(a.vals.size());
        for (size_t i = 0; i < a.vals.size(); ++i) {
            const Val& x = a.vals[i];
            const Val& y = b.vals[i];
            if (!x.sym && !y.sym && x.bits == y.bits) { out.vals[i] = x; continue; }
            if (x.sym && y.sym && x.node == y.node) { out.vals[i] = x; continue; }
            out.vals[i] = symv(x.ty, node(SDFT_S_SELECT, cond_node, node_of(x), node_of(y)));
        }
        return out;
    }

    bool fork_ok() {
        if (fork_depth >= 200)
            return fail("more than 200 nested branches depend on the position (a loop whose exit depends on it?)");
        if (++leaves > 65536) return fail("more than 65536 paths");
        return true;
    }
    Leaf failed() { Leaf l; l.st = ST_FAIL; return l; }
    Leaf trapped() { Leaf l; l.st = ST_TRAP; return l; }

    // Where the two sides of a branch on a symbolic condition meet again: the instruction after the construct,
    // in the same frame, with the same labels open.
    struct Stop {
        size_t depth, pc, labels;
    };

    static bool same_val(const Val& x, const Val& y) {
        if (x.sym != y.sym || x.ty != y.ty) return false;
        if (!x.sym) return x.bits == y.bits;
        return x.node == y.node && (!is64(x.ty) || x.node_hi == y.node_hi);
    }
    bool merge_val(uint32_t cond, const Val& x, const Val& y, Val* out) {
        if (same_val(x, y)) { *out = x; return true; }
        if (x.ty != y.ty) return false;
        if (is64(x.ty)) {  // a 64-bit value is a pair of words: select each half
            *out = sym64(x.ty, node(SDFT_S_SELECT, cond, half_node(x, 0), half_node(y, 0)),
                         node(SDFT_S_SELECT, cond, half_node(x, 1), half_node(y, 1)));
            return true;
        }
        *out = symv(x.ty, node(SDFT_S_SELECT, cond, node_of(x), node_of(y)));
        return true;
    }
    // if-conversion: one state that is `a` where cond holds and `b` where it does not.  Both stopped at the same
    // join point; false when they differ in something a select cannot express (then the caller keeps them apart).
    bool merge_states(uint32_t cond, const State& a, const State& b, State* out) {
        if (a.frames.size() != b.frames.size() || a.stack.size() != b.stack.size() || a.pages != b.pages ||
            a.globals.size() != b.globals.size())
            return false;
        State mrg = a;
        for (size_t i = 0; i < a.stack.size(); ++i)
            if (!merge_val(cond, a.stack[i], b.stack[i], &mrg.stack[i])) return false;
        for (size_t i = 0; i < a.globals.size(); ++i)
            if (!merge_val(cond, a.globals[i], b.globals[i], &mrg.globals[i])) return false;
        for (size_t k = 0; k < a.frames.size(); ++k) {
            const Frame &fa = a.frames[k], &fb = b.frames[k];
            if (fa.func != fb.func || fa.pc != fb.pc || fa.locals.size() != fb.locals.size() || fa.labels.size() != fb.labels.size())
                return false;
            for (size_t i = 0; i < fa.labels.size(); ++i)
                if (fa.labels[i].height != fb.labels[i].height || fa.labels[i].end_pc != fb.labels[i].end_pc ||
                    fa.labels[i].cont_pc != fb.labels[i].cont_pc)
                    return false;
            for (size_t i = 0; i < fa.locals.size(); ++i)
                if (!merge_val(cond, fa.locals[i], fb.locals[i], &mrg.frames[k].locals[i])) return false;
        }
        // guest memory: words either side wrote since instantiation
        auto word_node = [&](const State& s, uint32_t addr) -> Cell {
            auto it = s.mem.find(addr);
            return it != s.mem.end() ? it->second : Cell{false, base_word(addr)};
        };
        for (int side = 0; side < 2; ++side) {
            const State& s = side ? b : a;
            for (const auto& kv : s.mem) {
                const Cell ca = word_node(a, kv.first), cb = word_node(b, kv.first);
                if (ca.sym == cb.sym && ca.w == cb.w) { mrg.mem[kv.first] = ca; continue; }
                // a concrete word of memory is data, not structure: a constant of the tape (its bits survive the
                // float load unchanged), so that guests which differ only in such values share a kernel
                const uint32_t na = ca.sym ? ca.w : node_of(conc(T_F32, ca.w)), nb = cb.sym ? cb.w : node_of(conc(T_F32, cb.w));
                mrg.mem[kv.first] = Cell{true, node(SDFT_S_SELECT, cond, na, nb)};
            }
        }
        mrg.trap = trap_select(cond, a.trap, b.trap);
        *out = std::move(mrg);
        return true;
    }

    // The two sides of a branch on the symbolic condition `cond` have run until they rejoined (ST_REJOIN, their
    // states in A / B), finished the whole call (a final leaf) or trapped.  Returns true when execution continues
    // from `*st` (merged, or the surviving side); otherwise *out is the merged final outcome.
    bool resolve_fork(uint32_t cond, State& A, Leaf la, State& B, Leaf lb, FinishFn finish, State* st, Leaf* out) {
        if (la.st == ST_FAIL || lb.st == ST_FAIL) { *out = failed(); return false; }
        if (la.st == ST_REJOIN && lb.st == ST_REJOIN) {
            State mrg;
            if (merge_states(cond, A, B, &mrg)) { *st = std::move(mrg); return true; }
        }
        // one side traps: execution continues with the other one, and the positions that took the trapping side
        // join the trap predicate (A is the side where cond holds)
        if (la.st == ST_REJOIN && lb.st == ST_TRAP) {
            A.trap = trap_or(A.trap, node(SDFT_S_IEQZ, cond));
            ++trap_sides;
            *st = std::move(A);
            return true;
        }
        if (la.st == ST_TRAP && lb.st == ST_REJOIN) {
            B.trap = trap_or(B.trap, cond);
            ++trap_sides;
            *st = std::move(B);
            return true;
        }
        // one side left the construct for good (return, branch further out, unmergeable state): run whatever
        // stopped at the join point to the end as well and merge the final outcomes
        ++fork_depth;
        if (la.st == ST_REJOIN) la = run(A, finish, nullptr);
        if (lb.st == ST_REJOIN && la.st != ST_FAIL) lb = run(B, finish, nullptr);
        --fork_depth;
        *out = merge(cond, la, lb);
        return false;
    }

    // Run `st` until its outermost frame returns, then hand the state to `finish`; with `stop`, also until the
    // path reaches that join point (ST_REJOIN, the state stays in `st`).
    Leaf run(State& st, FinishFn finish, const Stop* stop = nullptr) {
        for (;;) {
            if (st.frames.empty()) return finish(*this, st);
            if (stop && st.frames.size() == stop->depth && st.frames.back().pc == stop->pc && st.frames.back().labels.size() == stop->labels) {
                Leaf l;
                l.st = ST_REJOIN;
                return l;
            }
            if (budget == 0) { fail("instruction budget exhausted (a loop whose exit depends on the position?)"); return failed(); }
            --budget;
            Frame& fr = st.frames.back();
            Func& f = m.funcs[fr.func];
            if (fr.pc >= f.code_len) {  // fell off the end of the body
                if (do_return(st)) return finish(*this, st);
                if (!err.empty()) return failed();
                continue;
            }
            Rd r{f.code + fr.pc, f.code + f.code_len};
            const size_t at = fr.pc;
            const uint8_t op = r.u8();
#define NEED(n) do { if (st.stack.size() < (size_t)(n)) { fail("stack underflow at 0x%02x", op); return failed(); } } while (0)
#define ADVANCE() do { if (!r.ok) { fail("truncated instruction"); return failed(); } fr.pc = (size_t)(r.p - f.code); } while (0)
            switch (op) {
                case 0x00: return trapped();  // unreachable
                case 0x01: ADVANCE(); break;
                case 0x02: case 0x03: {  // block, loop
                    const int64_t bt = r.sleb();
                    uint32_t np, nr;
                    if (!block_arity(bt, &np, &nr)) return failed();
                    ADVANCE();
                    NEED(np);
                    const BlockInfo& bi = f.blocks[at];
                    Label l;
                    l.is_loop = op == 0x03;
                    l.height = st.stack.size() - np;
                    l.arity = l.is_loop ? np : nr;
                    l.results = nr;
                    l.cont_pc = l.is_loop ? fr.pc : bi.end_pc + 1;
                    l.end_pc = bi.end_pc + 1;
                    fr.labels.push_back(l);
                    break;
                }
                case 0x04: {  // if
                    const int64_t bt = r.sleb();
                    uint32_t np, nr;
                    if (!block_arity(bt, &np, &nr)) return failed();
                    ADVANCE();
                    NEED(1);
                    const Val c = st.stack.back();
                    st.stack.pop_back();
                    NEED(np);
                    const BlockInfo bi = f.blocks[at];
                    Label l;
                    l.height = st.stack.size() - np;
                    l.arity = l.results = nr;
                    l.cont_pc = l.end_pc = bi.end_pc + 1;
                    auto take = [&](State& s, bool truth) {
                        Frame& g = s.frames.back();
                        if (truth) { g.labels.push_back(l); }                       // pc is already at the then arm
                        else if (bi.else_pc) { g.labels.push_back(l); g.pc = bi.else_pc + 1; }
                        else { g.pc = bi.end_pc + 1; }
                    };
                    if (!c.sym) { take(st, (uint32_t)c.bits != 0); break; }
                    if (!fork_ok()) return failed();
                    const Stop join{st.frames.size(), bi.end_pc + 1, fr.labels.size()};
                    State A = st, B = st;
                    take(A, true);
                    take(B, false);
                    ++fork_depth;
                    const Leaf la = run(A, finish, &join);
                    const Leaf lb = la.st == ST_FAIL ? failed() : run(B, finish, &join);
                    --fork_depth;
                    Leaf out;
                    if (!resolve_fork(c.node, A, la, B, lb, finish, &st, &out)) return out;
                    break;  // `st` is the merged state at the join point
                }
                case 0x05: {  // else reached from the then arm: leave the construct
                    const Label l = fr.labels.back();
                    fr.labels.pop_back();
                    fr.pc = l.end_pc;
                    break;
                }
                case 0x0b: {  // end
                    ADVANCE();
                    if (fr.labels.size() > 1) fr.labels.pop_back();
                    else if (do_return(st)) return finish(*this, st);
                    else if (!err.empty()) return failed();
                    break;
                }
                case 0x0c: {
                    const uint32_t d = r.u32();
                    ADVANCE();
                    if (branch(st, d)) return finish(*this, st);
                    if (!err.empty()) return failed();
                    break;
                }
                case 0x0d: {
                    const uint32_t d = r.u32();
                    ADVANCE();
                    NEED(1);
                    const Val c = st.stack.back();
                    st.stack.pop_back();
                    if (!c.sym) {
                        if ((uint32_t)c.bits != 0) {
                            if (branch(st, d)) return finish(*this, st);
                            if (!err.empty()) return failed();
                        }
                        break;
                    }
                    if (!fork_ok()) return failed();
                    if (d >= fr.labels.size()) { fail("branch depth out of range"); return failed(); }
                    const size_t target = fr.labels.size() - 1 - d;
                    const bool forward = target != 0 && !fr.labels[target].is_loop;  // a block's end: the sides can meet there
                    State A = st, B = st;  // A takes the branch
                    Leaf la, lb;
                    Stop join{0, 0, 0};
                    ++fork_depth;
                    if (branch(A, d)) la = finish(*this, A);
                    else if (!err.empty()) la = failed();
                    else if (forward) {
                        la.st = ST_REJOIN;  // it is at the join point already
                        join = Stop{A.frames.size(), A.frames.back().pc, A.frames.back().labels.size()};
                    } else la = run(A, finish, nullptr);
                    lb = la.st == ST_FAIL ? failed() : run(B, finish, la.st == ST_REJOIN ? &join : nullptr);
                    --fork_depth;
                    Leaf out;
                    if (!resolve_fork(c.node, A, la, B, lb, finish, &st, &out)) return out;
                    break;
                }
                case 0x0e: {
                    const uint32_t n = r.u32();
                    std::vector<uint32_t> targets(n + 1);
                    for (uint32_t i = 0; i <= n && r.ok; ++i) targets[i] = r.u32();
                    ADVANCE();
                    NEED(1);
                    const Val c = st.stack.back();
                    st.stack.pop_back();
                    if (c.sym) {
                        // a `match` on a value that depends on the position: every arm is followed to the end of the
                        // call (the arms leave through different labels, there is no common join point to stop at) and
                        // the outcomes are chained: index == 0 ? arm 0 : index == 1 ? arm 1 : ... : default
                        if (n > 64) { fail("br_table with %u position-dependent arms", n); return failed(); }
                        std::vector<Leaf> arms(n + 1);
                        ++fork_depth;
                        for (uint32_t k = 0; k <= n; ++k) {
                            if (!fork_ok()) { --fork_depth; return failed(); }
                            State arm = st;
                            if (branch(arm, targets[k])) arms[k] = finish(*this, arm);
                            else if (!err.empty()) arms[k] = failed();
                            else arms[k] = run(arm, finish, nullptr);
                            if (arms[k].st == ST_FAIL) { --fork_depth; return failed(); }
                        }
                        --fork_depth;
                        Leaf out = arms[n];
                        for (uint32_t k = n; k-- > 0;) out = merge(node(SDFT_S_IEQ, c.node, node(SDFT_S_IMM, k)), arms[k], out);
                        return out;
                    }
                    const uint32_t k = (uint32_t)c.bits;
                    if (branch(st, targets[k < n ? k : n])) return finish(*this, st);
                    if (!err.empty()) return failed();
                    break;
                }
                case 0x0f:
                    ADVANCE();
                    if (do_return(st)) return finish(*this, st);
                    if (!err.empty()) return failed();
                    break;
                case 0x10: {
                    const uint32_t fi = r.u32();
                    ADVANCE();
                    if (call_as_known_function(st, fi)) break;
                    if (!enter(st, fi)) return failed();
                    break;
                }
                case 0x11: {
                    const uint32_t ti = r.u32();
                    r.u32();
                    ADVANCE();
                    NEED(1);
                    const Val c = st.stack.back();
                    st.stack.pop_back();
                    if (c.sym) { fail("call_indirect through an index that depends on the position"); return failed(); }
                    const uint32_t k = (uint32_t)c.bits;
                    if (k >= m.table.size() || m.table[k] < 0) return trapped();
                    const uint32_t fi = (uint32_t)m.table[k];
                    if (fi >= m.funcs.size() || ti >= m.types.size()) return trapped();
                    const FuncType &want = m.types[ti], &have = m.types[m.funcs[fi].type];
                    if (want.params != have.params || want.results != have.results) return trapped();
                    if (call_as_known_function(st, fi)) break;
                    if (!enter(st, fi)) return failed();
                    break;
                }
                case 0x1a: ADVANCE(); NEED(1); st.stack.pop_back(); break;
                case 0x1b: case 0x1c: {
                    if (op == 0x1c) { const uint32_t n = r.u32(); for (uint32_t i = 0; i < n && r.ok; ++i) r.u8(); }
                    ADVANCE();
                    NEED(3);
                    const Val c = st.stack.back(); st.stack.pop_back();
                    const Val y = st.stack.back(); st.stack.pop_back();
                    const Val x = st.stack.back(); st.stack.pop_back();
                    if (!c.sym) { st.stack.push_back((uint32_t)c.bits != 0 ? x : y); break; }
                    if (!x.sym && !y.sym && x.bits == y.bits) { st.stack.push_back(x); break; }
                    Val sel;
                    if (!merge_val(c.node, x, y, &sel)) { fail("select of values of different types"); return failed(); }
                    st.stack.push_back(sel);
                    break;
                }
                case 0x20: { const uint32_t i = r.u32(); ADVANCE(); if (i >= fr.locals.size()) { fail("local out of range"); return failed(); } st.stack.push_back(fr.locals[i]); break; }
                case 0x21: { const uint32_t i = r.u32(); ADVANCE(); NEED(1); if (i >= fr.locals.size()) { fail("local out of range"); return failed(); } fr.locals[i] = st.stack.back(); st.stack.pop_back(); break; }
                case 0x22: { const uint32_t i = r.u32(); ADVANCE(); NEED(1); if (i >= fr.locals.size()) { fail("local out of range"); return failed(); } fr.locals[i] = st.stack.back(); break; }
                case 0x23: { const uint32_t i = r.u32(); ADVANCE(); if (i >= st.globals.size()) { fail("global out of range"); return failed(); } st.stack.push_back(st.globals[i]); break; }
                case 0x24: { const uint32_t i = r.u32(); ADVANCE(); NEED(1); if (i >= st.globals.size()) { fail("global out of range"); return failed(); } st.globals[i] = st.stack.back(); st.stack.pop_back(); break; }
                case 0x3f: r.u8(); ADVANCE(); st.stack.push_back(conc(T_I32, st.pages)); break;
                case 0x40: {
                    r.u8();
                    ADVANCE();
                    NEED(1);
                    const Val n = st.stack.back();
                    st.stack.pop_back();
                    if (n.sym) { fail("memory.grow by an amount that depends on the position"); return failed(); }
                    const uint64_t want = (uint64_t)st.pages + (uint32_t)n.bits;
                    if (want > m.mem_max || want > 16384) st.stack.push_back(conc(T_I32, 0xffffffffu));
                    else { st.stack.push_back(conc(T_I32, st.pages)); st.pages = (uint32_t)want; }
                    break;
                }
                case 0x41: { const int64_t v = r.sleb(); ADVANCE(); st.stack.push_back(conc(T_I32, (uint64_t)v)); break; }
                case 0x42: { const int64_t v = r.sleb(); ADVANCE(); st.stack.push_back(conc(T_I64, (uint64_t)v)); break; }
                case 0x43: { uint32_t w = 0; for (int i = 0; i < 4; ++i) w |= (uint32_t)r.u8() << (8 * i); ADVANCE(); st.stack.push_back(conc(T_F32, w)); break; }
                case 0x44: { uint64_t w = 0; for (int i = 0; i < 8; ++i) w |= (uint64_t)r.u8() << (8 * i); ADVANCE(); st.stack.push_back(conc(T_F64, w)); break; }
                case 0xfc: {
                    const uint32_t sub = r.u32();
                    if (sub <= 7) {  // saturating truncations
                        ADVANCE();
                        NEED(1);
                        const Val a = st.stack.back();
                        st.stack.pop_back();
                        if (a.sym) {
                            if (sub == 0) { st.stack.push_back(symv(T_I32, node(SDFT_S_I_FROM_F_S, a.node))); break; }
                            if (sub == 1) { st.stack.push_back(symv(T_I32, node(SDFT_S_I_FROM_F_U, a.node))); break; }
                            fail("a 64-bit conversion of a value that depends on the position");
                            return failed();
                        }
                        uint64_t o = 0;
                        const float fv = f32_of(a.bits);
                        const double dv = f64_of(a.bits);
                        switch (sub) {
                            case 0: trunc_to(fv, -2147483648.0, 2147483648.0, true, true, 32, &o); break;
                            case 1: trunc_to(fv, 0.0, 4294967296.0, true, false, 32, &o); break;
                            case 2: trunc_to(dv, -2147483648.0, 2147483648.0, true, true, 32, &o); break;
                            case 3: trunc_to(dv, 0.0, 4294967296.0, true, false, 32, &o); break;
                            case 4: trunc_to(fv, -9223372036854775808.0, 9223372036854775808.0, true, true, 64, &o); break;
                            case 5: trunc_to(fv, 0.0, 18446744073709551616.0, true, false, 64, &o); break;
                            case 6: trunc_to(dv, -9223372036854775808.0, 9223372036854775808.0, true, true, 64, &o); break;
                            default: trunc_to(dv, 0.0, 18446744073709551616.0, true, false, 64, &o); break;
                        }
                        st.stack.push_back(conc(sub < 4 ? T_I32 : T_I64, o));
                        break;
                    }
                    if (sub == 10 || sub == 11) {  // memory.copy, memory.fill
                        r.u8();
                        if (sub == 10) r.u8();
                        ADVANCE();
                        NEED(3);
                        const Val n = st.stack.back(); st.stack.pop_back();
                        const Val s = st.stack.back(); st.stack.pop_back();
                        const Val d = st.stack.back(); st.stack.pop_back();
                        if (n.sym || d.sym || (sub == 10 && s.sym)) { fail("memory.copy / fill with operands that depend on the position"); return failed(); }
                        const uint32_t len = (uint32_t)n.bits, dst = (uint32_t)d.bits, src = (uint32_t)s.bits;
                        if (!in_bounds(st, dst, len) || (sub == 10 && !in_bounds(st, src, len))) return trapped();
                        if (len > (64u << 20)) { fail("memory.copy / fill of more than 64 MiB"); return failed(); }
                        if (sub == 11) {
                            if (s.sym) { fail("memory.fill with a symbolic byte"); return failed(); }
                            const std::vector<uint8_t> fill(len, (uint8_t)s.bits);
                            if (!write_bytes(st, dst, fill.data(), len)) return failed();
                        } else if (len && ((dst | src | len) & 3u) == 0) {  // word-wise: symbolic words move as they are
                            std::vector<Cell> tmp(len / 4);
                            for (uint32_t i = 0; i < len / 4; ++i) {
                                auto it = st.mem.find(src + 4 * i);
                                tmp[i] = it != st.mem.end() ? it->second : Cell{false, base_word(src + 4 * i)};
                            }
                            for (uint32_t i = 0; i < len / 4; ++i) st.mem[dst + 4 * i] = tmp[i];
                        } else {
                            std::vector<uint8_t> tmp(len);
                            for (uint32_t i = 0; i < len; ++i)
                                if (read_byte(st, src + i, &tmp[i])) { fail("an unaligned memory.copy moves a symbolic word"); return failed(); }
                            if (!write_bytes(st, dst, tmp.data(), len)) return failed();
                        }
                        break;
                    }
                    fail("unsupported instruction 0xfc %u", sub);
                    return failed();
                }
                default: {
                    if (op >= 0x28 && op <= 0x35) {  // loads
                        r.u32();
                        const uint32_t off = r.u32();
                        ADVANCE();
                        NEED(1);
                        const Val a = st.stack.back();
                        st.stack.pop_back();
                        if (a.sym) { fail("a load from an address that depends on the position"); return failed(); }
                        static const uint8_t nb[14] = {4, 8, 4, 8, 1, 1, 2, 2, 1, 1, 2, 2, 4, 4};
                        static const uint8_t ty[14] = {T_I32, T_I64, T_F32, T_F64, T_I32, T_I32, T_I32, T_I32, T_I64, T_I64, T_I64, T_I64, T_I64, T_I64};
                        static const uint8_t sg[14] = {0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0};
                        const int k = op - 0x28;
                        const uint64_t addr = (uint64_t)(uint32_t)a.bits + off;
                        if (!in_bounds(st, addr, nb[k])) return trapped();
                        uint64_t v = 0;
                        bool is_sym = false;
                        uint32_t sn = 0;
                        if (k == 1 || k == 3) {  // i64.load / f64.load of words that depend on the position: a moved pair
                            uint32_t lo, hi;
                            if (load_pair(st, addr, &lo, &hi)) { st.stack.push_back(sym64(ty[k], lo, hi)); break; }
                        }
                        if (!load(st, addr, nb[k], &v, &is_sym, &sn)) return failed();
                        if (is_sym) {
                            if (k != 0 && k != 2) { fail("a symbolic word is loaded as a 64-bit value"); return failed(); }
                            st.stack.push_back(symv(ty[k], sn));
                            break;
                        }
                        if (sg[k]) {
                            const int bits = 8 * nb[k];
                            if (v & (1ull << (bits - 1))) v |= ~0ull << bits;
                        }
                        st.stack.push_back(conc(ty[k], v));
                        break;
                    }
                    if (op >= 0x36 && op <= 0x3e) {  // stores
                        r.u32();
                        const uint32_t off = r.u32();
                        ADVANCE();
                        NEED(2);
                        const Val v = st.stack.back(); st.stack.pop_back();
                        const Val a = st.stack.back(); st.stack.pop_back();
                        if (a.sym) { fail("a store to an address that depends on the position"); return failed(); }
                        static const uint8_t nb[9] = {4, 8, 4, 8, 1, 2, 1, 2, 4};
                        const uint64_t addr = (uint64_t)(uint32_t)a.bits + off;
                        const uint32_t n = nb[op - 0x36];
                        if (!in_bounds(st, addr, n)) return trapped();
                        if (!store(st, addr, n, v)) return failed();
                        break;
                    }
                    if (op >= 0x45 && op <= 0xc4) {  // numeric
                        ADVANCE();
                        if (is_unop(op)) {
                            NEED(1);
                            const Val a = st.stack.back();
                            st.stack.pop_back();
                            if (a.sym && is64(a.ty)) {  // a moved pair: only what keeps it a pair, or takes its low word
                                if (op == 0xbd || op == 0xbf) { st.stack.push_back(sym64(result_type(op), a.node, a.node_hi)); break; }
                                if (op == 0xa7) { st.stack.push_back(symv(T_I32, a.node)); break; }  // i32.wrap_i64
                                fail("64-bit arithmetic (0x%02x) on a value that depends on the position", op);
                                return failed();
                            }
                            if (a.sym) {
                                if (op == 0xbc || op == 0xbe) { st.stack.push_back(symv(result_type(op), a.node)); break; }  // reinterpret
                                // the trapping truncations: identical to the saturating ones wherever the guest does not
                                // trap (a guest that traps has no defined sample; the reference host substitutes one)
                                if (op == 0xa8 || op == 0xa9) {
                                    // in range (and not NaN): -2^31 <= trunc(x) < 2^31, resp. -1 < x < 2^32
                                    const uint32_t lo = op == 0xa8 ? node(SDFT_S_FGE, a.node, node_of(conc(T_F32, 0xCF000000u)))   // -2147483648.0f
                                                                   : node(SDFT_S_FGT, a.node, node_of(conc(T_F32, 0xBF800000u)));  // -1.0f
                                    const uint32_t hi = node(SDFT_S_FLT, a.node, node_of(conc(T_F32, op == 0xa8 ? 0x4F000000u : 0x4F800000u)));
                                    st.trap = trap_or(st.trap, node(SDFT_S_IEQZ, node(SDFT_S_IAND, lo, hi)));
                                    ++trap_ops;
                                    st.stack.push_back(symv(T_I32, node(op == 0xa8 ? SDFT_S_I_FROM_F_S : SDFT_S_I_FROM_F_U, a.node)));
                                    break;
                                }
                                const uint32_t so = sym_unop(op);
                                if (so == 0xffffffffu) { fail("instruction 0x%02x on a value that depends on the position has no 32-bit scalar form", op); return failed(); }
                                st.stack.push_back(symv(result_type(op), node(so, a.node)));
                                break;
                            }
                            Val o;
                            const Status s = unop(op, a, &o);
                            if (s == ST_TRAP) return trapped();
                            if (s == ST_FAIL) { fail("unsupported instruction 0x%02x", op); return failed(); }
                            st.stack.push_back(o);
                        } else {
                            NEED(2);
                            const Val b = st.stack.back(); st.stack.pop_back();
                            const Val a = st.stack.back(); st.stack.pop_back();
                            if ((a.sym && is64(a.ty)) || (b.sym && is64(b.ty))) {
                                fail("64-bit arithmetic (0x%02x) on a value that depends on the position", op);
                                return failed();
                            }
                            if (a.sym || b.sym) {
                                const uint32_t so = sym_binop(op);
                                if (so == 0xffffffffu) { fail("instruction 0x%02x on a value that depends on the position has no 32-bit scalar form", op); return failed(); }
                                if (op >= 0x6d && op <= 0x70) {  // i32.div_s / div_u / rem_s / rem_u trap on a zero divisor, div_s on INT_MIN / -1
                                    const uint32_t nb_ = node_of(b);
                                    uint32_t t = node(SDFT_S_IEQZ, nb_);
                                    if (op == 0x6d)
                                        t = node(SDFT_S_IOR, t, node(SDFT_S_IAND, node(SDFT_S_IEQ, node_of(a), node(SDFT_S_IMM, 0x80000000u)),
                                                                     node(SDFT_S_IEQ, nb_, node(SDFT_S_IMM, 0xffffffffu))));
                                    if (!b.sym) {  // a concrete divisor: the condition is known now
                                        const uint32_t y = (uint32_t)b.bits;
                                        if (y == 0) return trapped();
                                        if (!(op == 0x6d && y == 0xffffffffu)) t = NO_TRAP;
                                    }
                                    if (t != NO_TRAP) { st.trap = trap_or(st.trap, t); ++trap_ops; }
                                }
                                st.stack.push_back(symv(result_type(op), node(so, node_of(a), node_of(b))));
                                break;
                            }
                            Val o;
                            const Status s = binop(op, a, b, &o);
                            if (s == ST_TRAP) return trapped();
                            if (s == ST_FAIL) { fail("unsupported instruction 0x%02x", op); return failed(); }
                            st.stack.push_back(o);
                        }
                        break;
                    }
                    fail("unsupported instruction 0x%02x", op);
                    return failed();
                }
            }
#undef NEED
#undef ADVANCE
        }
    }

    State fresh_state() {
        State st;
        for (const Global& g : m.globals) st.globals.push_back(g.v);
        st.pages = m.mem_pages;
        return st;
    }

    // call an export with concrete arguments and no symbolic inputs; its writes become part of the instance
    bool call_concrete(uint32_t fi, const std::vector<Val>& args, std::vector<Val>* results) {
        State st = fresh_state();
        st.stack = args;
        if (!enter(st, fi)) return false;
        const Leaf l = run(st, [](Lowerer& self, State& s) {
            Leaf out;
            out.vals = s.stack;
            self.commit(s);
            return out;
        });
        if (l.st == ST_TRAP) return fail("the guest trapped during its set-up calls");
        if (l.st != ST_OK) return false;
        if (results) *results = l.vals;
        return true;
    }
};

void put_log(char* log, size_t cap, const std::string& s) {
    if (log && cap) snprintf(log, cap, "%s", s.c_str());
}

}  // namespace

static int lower_impl(const void* wasm, size_t wasm_bytes, const void* memory, size_t memory_bytes, uint32_t sdf_id,
                      void* tape_out, size_t tape_cap, size_t* tape_len, float bb_out[6], char* log, size_t log_cap) {
    if (log && log_cap) log[0] = '\0';
    if (tape_len) *tape_len = 0;
    if (!wasm || !wasm_bytes) { put_log(log, log_cap, "wasm is NULL or empty"); return SDFGPU_ERR_INVALID; }
    Lowerer L;
    if (const char* e = getenv("SDFGPU_WASM_BUDGET")) {  // instructions the partial evaluation may execute (tests)
        const long long n = atoll(e);
        if (n > 0) L.budget = (uint64_t)n;
    }
    if (!L.parse((const uint8_t*)wasm, wasm_bytes)) { put_log(log, log_cap, L.err); return SDFGPU_ERR_INVALID; }
    auto find_func = [&](const char* name, uint32_t* fi) {
        auto it = L.m.exports.find(name);
        if (it == L.m.exports.end() || it->second.first != 0 || it->second.second >= L.m.funcs.size()) return false;
        *fi = it->second.second;
        return true;
    };
    // required exports: memory, bounding_box, sample (native.rs:59-63)
    uint32_t f_bb = 0, f_sample = 0, f_tmp = 0;
    if (!L.m.has_memory || !find_func("bounding_box", &f_bb) || !find_func("sample", &f_sample)) {
        put_log(log, log_cap, "the module does not export memory, bounding_box and sample (src/sdf/wasm/native.rs:59-63)");
        return SDFGPU_ERR_INVALID;
    }
    {
        const FuncType& ts = L.m.types[L.m.funcs[f_sample].type];
        const FuncType& tb = L.m.types[L.m.funcs[f_bb].type];
        const std::vector<uint8_t> want = {T_I32, T_F32, T_F32, T_F32, T_I32};
        if (ts.params != want || ts.results != std::vector<uint8_t>{T_I32} || tb.params != std::vector<uint8_t>{T_I32} ||
            tb.results != std::vector<uint8_t>{T_I32}) {
            put_log(log, log_cap, "sample / bounding_box do not have the signatures of src/sdf/wasm/mod.rs:5-37");
            return SDFGPU_ERR_INVALID;
        }
    }
    std::vector<Val> res;
    if (memory) {
        // a live instance's linear memory (after init() and whatever set_parameter calls the host made): it
        // replaces instantiation.  Globals keep their initial values -- the ones a guest mutates across calls
        // (the shadow stack pointer) are restored by every export before it returns.
        if (memory_bytes == 0 || memory_bytes % 65536 || memory_bytes / 65536 > 16384) {
            put_log(log, log_cap, "the memory snapshot is not a whole number of 64 KiB pages");
            return SDFGPU_ERR_INVALID;
        }
        L.m.base.assign((const uint8_t*)memory, (const uint8_t*)memory + memory_bytes);
        L.m.mem_pages = (uint32_t)(memory_bytes / 65536);
    } else {
        // instantiate: start function, then the optional init() (native.rs:52-56)
        if (L.m.start >= 0 && !L.call_concrete((uint32_t)L.m.start, {}, nullptr)) { put_log(log, log_cap, "start: " + L.err); return SDFGPU_ERR_TAPE; }
        if (find_func("init", &f_tmp) && L.m.types[L.m.funcs[f_tmp].type].params.empty() &&
            !L.call_concrete(f_tmp, {}, nullptr)) {
            put_log(log, log_cap, "init: " + L.err);
            return SDFGPU_ERR_TAPE;
        }
    }
    // bounding_box(sdf_id) -> *[f32; 6] (src/sdf/ffi.rs:42-50, native.rs:163-186)
    if (!L.call_concrete(f_bb, {Lowerer::conc(T_I32, sdf_id)}, &res) || res.size() != 1) {
        put_log(log, log_cap, "bounding_box: " + L.err);
        return SDFGPU_ERR_TAPE;
    }
    {
        const uint32_t p = (uint32_t)res[0].bits;
        State st = L.fresh_state();
        for (int i = 0; i < 6; ++i) {
            uint64_t v = 0;
            bool is_sym = false;
            uint32_t sn = 0;
            if (!L.in_bounds(st, (uint64_t)p + 4 * i, 4) || !L.load(st, (uint64_t)p + 4 * i, 4, &v, &is_sym, &sn)) {
                put_log(log, log_cap, "bounding_box returned a pointer outside the guest memory");
                return SDFGPU_ERR_TAPE;
            }
            if (bb_out) bb_out[i] = Lowerer::f32_of(v);
        }
        if (find_func("bounding_box_free", &f_tmp) && L.m.types[L.m.funcs[f_tmp].type].params == std::vector<uint8_t>{T_I32})
            (void)L.call_concrete(f_tmp, {Lowerer::conc(T_I32, p)}, nullptr);
        L.err.clear();
    }
    // sample(sdf_id, x, y, z, distance_only = 0) with symbolic coordinates (src/sdf/ffi.rs:57-65, native.rs:188-217)
    State st = L.fresh_state();
    st.stack.push_back(Lowerer::conc(T_I32, sdf_id));
    st.stack.push_back(Lowerer::symv(T_F32, L.node(SDFT_S_PX)));
    st.stack.push_back(Lowerer::symv(T_F32, L.node(SDFT_S_PY)));
    st.stack.push_back(Lowerer::symv(T_F32, L.node(SDFT_S_PZ)));
    st.stack.push_back(Lowerer::conc(T_I32, 0));
    if (!L.enter(st, f_sample)) { put_log(log, log_cap, "sample: " + L.err); return SDFGPU_ERR_TAPE; }
    const Leaf leaf = L.run(st, [](Lowerer& self, State& s) {
        Leaf out;
        if (s.stack.empty() || s.stack.back().sym) {
            self.fail("sample returns a pointer that depends on the position");
            out.st = ST_FAIL;
            return out;
        }
        const uint32_t p = (uint32_t)s.stack.back().bits;
        for (uint32_t i = 0; i < 7; ++i) {  // the 28 bytes of SDFSample (src/sdf/mod.rs:104-118)
            uint64_t v = 0;
            bool is_sym = false;
            uint32_t sn = 0;
            if (!self.in_bounds(s, (uint64_t)p + 4 * i, 4)) { out.st = ST_TRAP; return out; }
            if (!self.load(s, (uint64_t)p + 4 * i, 4, &v, &is_sym, &sn)) { out.st = ST_FAIL; return out; }
            out.vals.push_back(is_sym ? Lowerer::symv(T_F32, sn) : Lowerer::conc(T_F32, v));
        }
        out.trap = s.trap;
        return out;
    });
    if (leaf.st == ST_TRAP) { put_log(log, log_cap, "sample traps for every position"); return SDFGPU_ERR_TAPE; }
    if (leaf.st != ST_OK || leaf.vals.size() != 7) { put_log(log, log_cap, "sample: " + L.err); return SDFGPU_ERR_TAPE; }

    // ---- the program: nodes reachable from the seven outputs, in creation (= topological) order
    std::vector<uint32_t> outs(7);
    for (int i = 0; i < 7; ++i) {
        outs[i] = L.node_of(leaf.vals[i]);
        // where the guest would have trapped the reference substitutes SDFSample::new(1.0, zero)
        // (src/sdf/wasm/native.rs:196-203; src/sdf/mod.rs:121-125): distance 1, everything else 0
        if (leaf.trap != NO_TRAP)
            outs[i] = L.node(SDFT_S_SELECT, leaf.trap, L.node_of(Lowerer::conc(T_F32, i == 0 ? 0x3F800000u : 0u)), outs[i]);
    }
    std::vector<char> live(L.nodes.size(), 0);
    for (uint32_t o : outs) live[o] = 1;
    for (size_t i = L.nodes.size(); i-- > 0;) {
        if (!live[i]) continue;
        const Node& n = L.nodes[i];
        const int n_in = Lowerer::sop_inputs(n.op);
        if (n_in >= 1) live[n.a] = 1;
        if (n_in >= 2) live[n.b] = 1;
        if (n_in >= 3) live[n.c] = 1;
    }
    std::vector<uint32_t> remap(L.nodes.size(), 0), used_consts;
    std::map<uint32_t, uint32_t> const_remap;
    std::vector<sdft_sop> sops;
    for (size_t i = 0; i < L.nodes.size(); ++i) {
        if (!live[i]) continue;
        const Node& n = L.nodes[i];
        sdft_sop o;
        o.op = n.op; o.a = n.a; o.b = n.b; o.c = n.c;
        if (n.op == SDFT_S_CONST) {
            auto it = const_remap.find(n.a);
            if (it == const_remap.end()) {
                const_remap[n.a] = (uint32_t)used_consts.size();
                o.a = (uint32_t)used_consts.size();
                used_consts.push_back(L.const_bits[n.a]);
            } else {
                o.a = it->second;
            }
            o.b = o.c = 0;
        } else if (n.op > SDFT_S_IMM) {
            const int n_in = Lowerer::sop_inputs(n.op);
            o.a = remap[n.a];
            o.b = n_in >= 2 ? remap[n.b] : 0;
            o.c = n_in >= 3 ? remap[n.c] : 0;
        }
        remap[i] = (uint32_t)sops.size();
        sops.push_back(o);
    }
    for (uint32_t k = 0; k < 7; ++k) sops.push_back(sdft_sop{SDFT_S_OUT, remap[outs[k]], k, 0});
    if (sops.size() > SDFT_MAX_SOPS || used_consts.size() > SDFT_MAX_CONSTS) {
        char buf[160];
        snprintf(buf, sizeof buf, "the lowered program has %zu ops and %zu constants (limits %u, %u)", sops.size(), used_consts.size(),
                 SDFT_MAX_SOPS, SDFT_MAX_CONSTS);
        put_log(log, log_cap, buf);
        return SDFGPU_ERR_TAPE;
    }
    // ---- the tape: SCALAR over the whole program, END
    sdft_header h;
    memset(&h, 0, sizeof h);
    h.magic = SDFT_MAGIC; h.version = SDFT_VERSION;
    h.n_instr = 2; h.n_prims = 0; h.n_consts = (uint32_t)used_consts.size();
    h.reserved[0] = (uint32_t)sops.size();
    const sdft_instr ins[2] = {{SDFT_OP_SCALAR, 0, (uint32_t)sops.size(), 0.0f}, {SDFT_OP_END, 0, 0, 0.0f}};
    const size_t need = sizeof h + sizeof ins + used_consts.size() * 4 + sops.size() * sizeof(sdft_sop);
    if (tape_len) *tape_len = need;
    if (tape_out && tape_cap >= need) {
        unsigned char* p = (unsigned char*)tape_out;
        memcpy(p, &h, sizeof h); p += sizeof h;
        memcpy(p, ins, sizeof ins); p += sizeof ins;
        if (!used_consts.empty()) memcpy(p, used_consts.data(), used_consts.size() * 4);
        p += used_consts.size() * 4;
        memcpy(p, sops.data(), sops.size() * sizeof(sdft_sop));
    } else if (tape_out) {
        put_log(log, log_cap, "tape buffer too small");
        return SDFGPU_ERR_INVALID;
    }
    {
        char buf[320];
        int n = snprintf(buf, sizeof buf, "lowered: %zu scalar ops, %zu constants, %u symbolic branches merged", sops.size(),
                         used_consts.size(), L.leaves);
        if (L.recognised_calls) n += snprintf(buf + n, sizeof buf - n, ", %u fmodf calls recognised", L.recognised_calls);
        if (L.skipped_imports) n += snprintf(buf + n, sizeof buf - n, ", %u void host imports skipped", L.skipped_imports);
        if (leaf.trap != NO_TRAP)
            n += snprintf(buf + n, sizeof buf - n, ", position-dependent traps kept (%u branch sides, %u div/trunc ops): those voxels "
                          "get the reference's fallback sample (1.0, 0)", L.trap_sides, L.trap_ops);
        put_log(log, log_cap, buf);
    }
    return SDFGPU_OK;
}

extern "C" __attribute__((visibility("default"))) int sdfgpu_wasm_lower(const void* wasm, size_t wasm_bytes, uint32_t sdf_id,
                                                                          void* tape_out, size_t tape_cap, size_t* tape_len,
                                                                          float bb_out[6], char* log, size_t log_cap) {
    return lower_impl(wasm, wasm_bytes, nullptr, 0, sdf_id, tape_out, tape_cap, tape_len, bb_out, log, log_cap);
}

extern "C" __attribute__((visibility("default"))) int sdfgpu_wasm_lower_live(const void* wasm, size_t wasm_bytes, const void* memory,
                                                                               size_t memory_bytes, uint32_t sdf_id, void* tape_out,
                                                                               size_t tape_cap, size_t* tape_len, float bb_out[6],
                                                                               char* log, size_t log_cap) {
    if (!memory) {
        if (log && log_cap) snprintf(log, log_cap, "memory is NULL");
        return SDFGPU_ERR_INVALID;
    }
    return lower_impl(wasm, wasm_bytes, memory, memory_bytes, sdf_id, tape_out, tape_cap, tape_len, bb_out, log, log_cap);
}

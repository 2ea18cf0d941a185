// sdf_oracle.cpp -- CPU restatement of the sdf-viewer grid-fill + sphere-trace
// hot path.  TEST INFRASTRUCTURE ONLY: nothing under sdf-viewer_b200/ may
// import, link or execute this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker and
// as the timed CPU baseline.
//
// PARITY STATUS: "parity unpinned".  The reference (Rust + wasmer + GLSL) cannot
// be compiled or run in this environment (no cargo/rustc/wasm runtime/GL), and
// its own tests hold no value-level golden vectors for this path; the only
// reference tests are the five LoadingManager invariants
// (/root/reference/src/app/scene/sdf/loading.rs:117-171), which tests/ port
// verbatim.  Every function below cites the reference lines it restates.
// What stands in for reference outputs: known answers worked out on paper from the cited lines (fill: tests/test_oracle.py;
// trace: tests/trace_kats.py), and two restatements written a second time from the reference alone, in numpy, which this
// file has to equal -- tests/demo_numpy_ref.py (SDFDemo::sample, positions, store rules: bit for bit) and
// tests/frag_numpy_ref.py (material.frag + the GL sampling rules: same hit / miss class on every pixel).
// Arithmetic from the un-vendored crates three-d 0.18.2 / three-d-asset 0.9.2 /
// cgmath 0.18.0 (Cargo.lock:6596,6613,1054) is restated from their published
// algorithms and isolated in srgb_u8_to_linear(), calculate_lighting(),
// tone_mapping(), color_mapping().
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
// (no FMA contraction: Rust/WASM f32 arithmetic is never fused).
//
// Paths in comments are relative to /root/reference.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/sdfgpu_tape.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))
#define SDFGPU_GBUF_FLOATS_ORC 16  // same record as SDFGPU_GBUF_FLOATS in include/sdfgpu.h

namespace {

struct V3 {
    float x, y, z;
};
struct Sample {  // SDFSample, src/sdf/mod.rs:104-118
    float d, r, g, b, metallic, roughness, occlusion;
};

inline Sample sample_new(float d, float r, float g, float b) {  // SDFSample::new, src/sdf/mod.rs:120-126
    return Sample{d, r, g, b, 0.0f, 0.0f, 0.0f};
}

// ---------------------------------------------------------------- constants

// src/app/scene/sdf/mod.rs:42  `const AIR_DIST: f32 = 1e-1 + 0.001234;`
inline float air_dist() {
    volatile float a = 1e-1f, b = 0.001234f;
    return a + b;
}

// three-d-asset Srgba::to_linear_srgb (call site src/app/scene/sdf/mod.rs:201):
// c = u8/255; c < 0.04045 ? c/12.92 : ((c+0.055)/1.055)^2.4
inline float srgb_u8_to_linear(unsigned v) {
    float c = (float)v / 255.0f;
    if (c < 0.04045f) return c / 12.92f;
    return powf((c + 0.055f) / 1.055f, 2.4f);
}

// three-d-asset `Srgba::from(Vector3<f32>)`: `(c * 255.0) as u8`; Rust float->int
// casts saturate and map NaN to 0.
inline unsigned f32_to_u8_sat(float v) {
    float s = v * 255.0f;
    if (!(s == s)) return 0u;
    if (s <= 0.0f) return 0u;
    if (s >= 255.0f) return 255u;
    return (unsigned)s;  // truncation
}

inline float rust_max(float a, float b) { return fmaxf(a, b); }  // f32::max: NaN-ignoring
inline float rust_min(float a, float b) { return fminf(a, b); }
inline float rust_signum(float x) { return (x != x) ? x : copysignf(1.0f, x); }

// ------------------------------------------------------- demo SDF, direct form

struct DemoParams {
    float cube_half_side;                // cube.rs:17  default 0.95
    uint32_t cube_material;              // cube.rs:15  default brick (SDFT_MAT_BRICK)
    float sphere_radius;                 // sphere.rs:13 default 1.05
    uint32_t sphere_material;            // sphere.rs:11 default normal (SDFT_MAT_NORMAL)
    float max_distance_custom_material;  // demo/mod.rs:26 default 0.05
    uint32_t disable_sphere;             // demo/mod.rs:28 default false
};

// sample_brick_texture, src/sdf/demo/cube.rs:181-222
Sample brick_texture(V3 p, V3 n, float distance) {
    const float BRICK_R = 150.f / 255.f, BRICK_G = 24.f / 255.f, BRICK_B = 10.f / 255.f;
    const float BRICK_WIDTH = 0.5f, BRICK_HEIGHT = 0.25f;
    const float CEM_R = 56.f / 255.f, CEM_G = 70.f / 255.f, CEM_B = 60.f / 255.f;
    const float CEMENT_THICKNESS = 0.2f;
    float u, v;
    if (fabsf(n.x) > fabsf(n.y)) {         // :206
        if (fabsf(n.x) > fabsf(n.z)) { u = p.z; v = p.y; }  // :207-209
        else { u = p.x; v = p.y; }         // :210-212
    } else if (fabsf(n.y) > fabsf(n.z)) {  // :214
        u = p.z; v = p.x;
    } else {                               // :217
        u = p.x; v = p.y;
    }
    // compute_tex2d, :189-202
    float row_num = v / BRICK_HEIGHT;
    float brick_offset = floorf(row_num) / 4.0f;
    float bx = fmodf(fabsf(u + brick_offset), BRICK_WIDTH);
    float by = fmodf(fabsf(v), BRICK_HEIGHT);
    float max_cement = CEMENT_THICKNESS / 2.0f * BRICK_HEIGHT;
    Sample s;
    s.d = distance;
    if (bx < max_cement || bx > BRICK_WIDTH - max_cement || by < max_cement ||
        by > BRICK_HEIGHT - max_cement) {
        s.r = CEM_R; s.g = CEM_G; s.b = CEM_B; s.metallic = 0.4f; s.roughness = 0.5f; s.occlusion = 1.0f;
    } else {
        s.r = BRICK_R; s.g = BRICK_G; s.b = BRICK_B; s.metallic = 0.2f; s.roughness = 0.8f; s.occlusion = 0.0f;
    }
    return s;
}

// SDFDemoCube::normal, src/sdf/demo/cube.rs:164-177
V3 cube_normal(V3 p, float half) {
    V3 n{0.f, 0.f, 0.f};
    if (fabsf(p.x) > half) n.x = rust_signum(p.x);
    if (fabsf(p.y) > half) n.y = rust_signum(p.y);
    if (fabsf(p.z) > half) n.z = rust_signum(p.z);
    return n;
}

// cgmath InnerSpace::normalize = self * (1 / magnitude); magnitude = sqrt(x*x + y*y + z*z)
// (call site src/sdf/demo/sphere.rs:123)
V3 cg_normalize(V3 p) {
    float m = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    float inv = 1.0f / m;
    return V3{p.x * inv, p.y * inv, p.z * inv};
}

// Material::render, src/sdf/demo/cube.rs:50-59
Sample material_render(uint32_t mat, float dist, V3 p, V3 n) {
    if (mat == SDFT_MAT_BRICK) return brick_texture(p, n, dist);
    return sample_new(dist, fabsf(n.x), fabsf(n.y), fabsf(n.z));
}

// SDFDemoCube::sample, src/sdf/demo/cube.rs:79-89
Sample demo_cube_sample(const DemoParams& P, V3 p, bool distance_only) {
    float dist_box = rust_max(rust_max(fabsf(p.x), fabsf(p.y)), fabsf(p.z)) - P.cube_half_side;
    distance_only = distance_only || dist_box > 0.1f;
    if (distance_only) return sample_new(dist_box, 0.f, 0.f, 0.f);
    return material_render(P.cube_material, dist_box, p, cube_normal(p, P.cube_half_side));
}

// SDFDemoSphere::sample, src/sdf/demo/sphere.rs:37-47 (cgmath distance to the origin)
Sample demo_sphere_sample(const DemoParams& P, V3 p, bool distance_only) {
    float dist_sphere = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z) - P.sphere_radius;
    distance_only = distance_only || dist_sphere > 0.1f;
    if (distance_only) return sample_new(dist_sphere, 0.f, 0.f, 0.f);
    return material_render(P.sphere_material, dist_sphere, p, cg_normalize(p));
}

// The SDFDemo combinator, src/sdf/demo/mod.rs:58-73
Sample demo_diff(Sample box, Sample sph, float thresh, const float seam[6]) {
    float dist = rust_max(box.d, -sph.d);                      // :58
    float inter = fabsf(box.d) - fabsf(sph.d);                 // :60
    Sample s = (inter < 0.0f) ? box : sph;                     // :61
    if (fabsf(inter) <= thresh) {                              // :62
        s.r = seam[0]; s.g = seam[1]; s.b = seam[2];           // :66
        s.metallic = seam[3]; s.roughness = seam[4]; s.occlusion = seam[5];  // :67-69
    }
    s.d = dist;                                                // :72
    return s;
}

// SDFDemo::sample, src/sdf/demo/mod.rs:51-75
Sample demo_sample(const DemoParams& P, V3 p, bool distance_only) {
    Sample box = demo_cube_sample(P, p, distance_only);
    if (P.disable_sphere) return box;
    Sample sph = demo_sphere_sample(P, p, distance_only);
    const float seam[6] = {0.5f, 0.6f, 0.7f, 0.5f, 0.0f, 0.0f};
    return demo_diff(box, sph, P.max_distance_custom_material, seam);
}

// ---------------------------------------------------------- tape interpreter

struct Tape {
    const sdft_header* hdr;
    const sdft_instr* instr;
    const sdft_prim* prims;
    const float* consts;
    const sdft_sop* sops;  // optional scalar-program section, hdr->reserved[0] ops
};

bool tape_parse(const void* bytes, size_t len, Tape* t) {
    if (len < sizeof(sdft_header)) return false;
    const sdft_header* h = (const sdft_header*)bytes;
    if (h->magic != SDFT_MAGIC || h->version != SDFT_VERSION) return false;
    size_t need = sizeof(sdft_header) + (size_t)h->n_instr * sizeof(sdft_instr) +
                  (size_t)h->n_prims * sizeof(sdft_prim) + (size_t)h->n_consts * 4;
    need += (size_t)h->reserved[0] * sizeof(sdft_sop);
    if (need > len) return false;
    t->hdr = h;
    t->instr = (const sdft_instr*)((const char*)bytes + sizeof(sdft_header));
    t->prims = (const sdft_prim*)(t->instr + h->n_instr);
    t->consts = (const float*)(t->prims + h->n_prims);
    t->sops = (const sdft_sop*)(t->consts + h->n_consts);
    return true;
}

// ---- scalar programs (include/sdfgpu_tape.h): WebAssembly's numeric semantics on untyped 32-bit words
inline float w2f(uint32_t w) { float f; memcpy(&f, &w, 4); return f; }
inline uint32_t f2w(float f) { uint32_t w; memcpy(&w, &f, 4); return w; }
inline float wasm_fmin(float a, float b) {
    if (a != a || b != b) return w2f(0x7fc00000u);
    if (a == b) return w2f(f2w(a) | f2w(b));  // -0 < +0
    return a < b ? a : b;
}
inline float wasm_fmax(float a, float b) {
    if (a != a || b != b) return w2f(0x7fc00000u);
    if (a == b) return w2f(f2w(a) & f2w(b));
    return a > b ? a : b;
}
inline int32_t wasm_trunc_sat_s(float f) {
    if (f != f) return 0;
    if (f <= -2147483648.0f) return INT32_MIN;
    if (f >= 2147483648.0f) return INT32_MAX;
    return (int32_t)f;
}
inline uint32_t wasm_trunc_sat_u(float f) {
    if (f != f || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

void run_scalar_program(const Tape& t, uint32_t first, uint32_t count, V3 p, Sample& A) {
    std::vector<uint32_t> v(count);
    float* out[7] = {&A.d, &A.r, &A.g, &A.b, &A.metallic, &A.roughness, &A.occlusion};
    for (uint32_t i = 0; i < count; ++i) {
        const sdft_sop& o = t.sops[first + i];
        const uint32_t a = o.a < i ? v[o.a] : 0u, b = o.b < i ? v[o.b] : 0u, c = o.c < i ? v[o.c] : 0u;
        const float fa = w2f(a), fb = w2f(b);
        uint32_t r = 0;
        switch (o.op) {
            case SDFT_S_PX: r = f2w(p.x); break;
            case SDFT_S_PY: r = f2w(p.y); break;
            case SDFT_S_PZ: r = f2w(p.z); break;
            case SDFT_S_CONST: r = f2w(t.consts[o.a]); break;
            case SDFT_S_IMM: r = o.a; break;
            case SDFT_S_FNEG: r = a ^ 0x80000000u; break;
            case SDFT_S_FABS: r = a & 0x7fffffffu; break;
            case SDFT_S_FSQRT: r = f2w(sqrtf(fa)); break;
            case SDFT_S_FFLOOR: r = f2w(floorf(fa)); break;
            case SDFT_S_FCEIL: r = f2w(ceilf(fa)); break;
            case SDFT_S_FTRUNC: r = f2w(truncf(fa)); break;
            case SDFT_S_FNEAREST: r = f2w(nearbyintf(fa)); break;  // default rounding mode: half to even
            case SDFT_S_FADD: r = f2w(fa + fb); break;
            case SDFT_S_FSUB: r = f2w(fa - fb); break;
            case SDFT_S_FMUL: r = f2w(fa * fb); break;
            case SDFT_S_FDIV: r = f2w(fa / fb); break;
            case SDFT_S_FMIN: r = f2w(wasm_fmin(fa, fb)); break;
            case SDFT_S_FMAX: r = f2w(wasm_fmax(fa, fb)); break;
            case SDFT_S_FCOPYSIGN: r = (a & 0x7fffffffu) | (b & 0x80000000u); break;
            case SDFT_S_FMOD: r = f2w(fmodf(fa, fb)); break;  // exact by definition (cube.rs:192 `%`)
            case SDFT_S_FEQ: r = fa == fb; break;
            case SDFT_S_FNE: r = fa != fb; break;
            case SDFT_S_FLT: r = fa < fb; break;
            case SDFT_S_FGT: r = fa > fb; break;
            case SDFT_S_FLE: r = fa <= fb; break;
            case SDFT_S_FGE: r = fa >= fb; break;
            case SDFT_S_IADD: r = a + b; break;
            case SDFT_S_ISUB: r = a - b; break;
            case SDFT_S_IMUL: r = a * b; break;
            case SDFT_S_IAND: r = a & b; break;
            case SDFT_S_IOR: r = a | b; break;
            case SDFT_S_IXOR: r = a ^ b; break;
            case SDFT_S_ISHL: r = a << (b & 31u); break;
            case SDFT_S_ISHR_U: r = a >> (b & 31u); break;
            case SDFT_S_ISHR_S: r = (uint32_t)((int32_t)a >> (b & 31u)); break;
            case SDFT_S_IDIV_S: r = (b == 0u || (a == 0x80000000u && b == 0xffffffffu)) ? 0u : (uint32_t)((int32_t)a / (int32_t)b); break;
            case SDFT_S_IDIV_U: r = b == 0u ? 0u : a / b; break;
            case SDFT_S_IREM_S: r = (b == 0u || b == 0xffffffffu) ? 0u : (uint32_t)((int32_t)a % (int32_t)b); break;
            case SDFT_S_IREM_U: r = b == 0u ? 0u : a % b; break;
            case SDFT_S_IEQ: r = a == b; break;
            case SDFT_S_INE: r = a != b; break;
            case SDFT_S_ILT_S: r = (int32_t)a < (int32_t)b; break;
            case SDFT_S_ILT_U: r = a < b; break;
            case SDFT_S_IGT_S: r = (int32_t)a > (int32_t)b; break;
            case SDFT_S_IGT_U: r = a > b; break;
            case SDFT_S_ILE_S: r = (int32_t)a <= (int32_t)b; break;
            case SDFT_S_ILE_U: r = a <= b; break;
            case SDFT_S_IGE_S: r = (int32_t)a >= (int32_t)b; break;
            case SDFT_S_IGE_U: r = a >= b; break;
            case SDFT_S_IEQZ: r = a == 0u; break;
            case SDFT_S_SELECT: r = a != 0u ? b : c; break;
            case SDFT_S_F_FROM_I_S: r = f2w((float)(int32_t)a); break;
            case SDFT_S_F_FROM_I_U: r = f2w((float)a); break;
            case SDFT_S_I_FROM_F_S: r = (uint32_t)wasm_trunc_sat_s(fa); break;
            case SDFT_S_I_FROM_F_U: r = wasm_trunc_sat_u(fa); break;
            case SDFT_S_OUT: if (o.b < 7) *out[o.b] = fa; break;
            default: break;
        }
        v[i] = r;
    }
}

inline float prim_distance(const sdft_prim& pr, V3 q) {
    if ((pr.kind & 0xffu) == SDFT_SHAPE_SPHERE)
        return sqrtf(q.x * q.x + q.y * q.y + q.z * q.z) - pr.size;          // sphere.rs:39
    return rust_max(rust_max(fabsf(q.x), fabsf(q.y)), fabsf(q.z)) - pr.size;  // cube.rs:81
}

Sample prim_sample(const sdft_prim& pr, V3 p) {
    V3 q{p.x - pr.center[0], p.y - pr.center[1], p.z - pr.center[2]};
    float d = prim_distance(pr, q);
    if (d > pr.air_skip) return sample_new(d, 0.f, 0.f, 0.f);  // cube.rs:83-85, sphere.rs:41-43
    uint32_t mat = (pr.kind >> 8) & 0xffu;
    if (mat == SDFT_MAT_FLAT) {
        return Sample{d, pr.color[0], pr.color[1], pr.color[2], pr.metallic, pr.roughness, pr.occlusion};
    }
    V3 n = ((pr.kind & 0xffu) == SDFT_SHAPE_SPHERE) ? cg_normalize(q) : cube_normal(q, pr.size);
    return material_render(mat, d, q, n);
}

inline Sample op_union(Sample x, Sample y) { return (y.d < x.d) ? y : x; }
inline Sample op_inter(Sample x, Sample y) { return (y.d > x.d) ? y : x; }

Sample tape_sample(const Tape& t, V3 p0) {
    Sample A = sample_new(0.f, 0.f, 0.f, 0.f);
    Sample S[SDFT_MAX_STACK];
    unsigned sp = 0;
    V3 p = p0;
    for (uint32_t pc = 0; pc < t.hdr->n_instr; ++pc) {
        const sdft_instr& I = t.instr[pc];
        switch (I.op) {
            case SDFT_OP_END: return A;
            case SDFT_OP_PRIM: A = prim_sample(t.prims[I.a], p); break;
            case SDFT_OP_UNION_PRIM: A = op_union(A, prim_sample(t.prims[I.a], p)); break;
            case SDFT_OP_INTER_PRIM: A = op_inter(A, prim_sample(t.prims[I.a], p)); break;
            case SDFT_OP_UNION_RANGE: {
                A = prim_sample(t.prims[I.a], p);
                for (uint32_t k = 1; k < I.b; ++k) A = op_union(A, prim_sample(t.prims[I.a + k], p));
                break;
            }
            case SDFT_OP_PUSH: S[sp++] = A; break;
            case SDFT_OP_POP_UNION: A = op_union(S[--sp], A); break;
            case SDFT_OP_POP_INTER: A = op_inter(S[--sp], A); break;
            case SDFT_OP_POP_DEMO_DIFF: {
                const float* c = t.consts + I.a;
                A = demo_diff(S[--sp], A, c[0], c + 1);
                break;
            }
            case SDFT_OP_D_NEG: A.d = -A.d; break;
            case SDFT_OP_D_ABS: A.d = fabsf(A.d); break;
            case SDFT_OP_D_ADD: A.d = A.d + I.imm; break;
            case SDFT_OP_D_MUL: A.d = A.d * I.imm; break;
            case SDFT_OP_D_MAX: A.d = rust_max(A.d, I.imm); break;
            case SDFT_OP_D_MIN: A.d = rust_min(A.d, I.imm); break;
            case SDFT_OP_M_SET: {
                const float* c = t.consts + I.a;
                A.r = c[0]; A.g = c[1]; A.b = c[2]; A.metallic = c[3]; A.roughness = c[4]; A.occlusion = c[5];
                break;
            }
            case SDFT_OP_P_RESET: p = p0; break;
            case SDFT_OP_P_SUB: {
                const float* c = t.consts + I.a;
                p.x = p.x - c[0]; p.y = p.y - c[1]; p.z = p.z - c[2];
                break;
            }
            case SDFT_OP_P_MUL: p.x = p.x * I.imm; p.y = p.y * I.imm; p.z = p.z * I.imm; break;
            case SDFT_OP_P_ABS:
                if (I.a & 1u) p.x = fabsf(p.x);
                if (I.a & 2u) p.y = fabsf(p.y);
                if (I.a & 4u) p.z = fabsf(p.z);
                break;
            case SDFT_OP_SCALAR:
                if ((uint64_t)I.a + I.b <= t.hdr->reserved[0]) run_scalar_program(t, I.a, I.b, p, A);
                break;
            default: break;
        }
    }
    return A;
}

// --------------------------------------------------------- LoadingManager

// prev_power_of_2, src/app/scene/sdf/loading.rs:108-115
uint32_t prev_power_of_2(uint32_t x) {
    x = x | (x >> 1);
    x = x | (x >> 2);
    x = x | (x >> 4);
    x = x | (x >> 8);
    x = x | (x >> 16);
    return x - (x >> 1);
}

struct LoadingManager {  // src/app/scene/sdf/loading.rs:5-19
    uint64_t limits[3];
    uint64_t passes;
    uint64_t step_size;
    uint64_t next_index[3];
    uint64_t iterations;
    uint64_t total_iterations;

    void reset(uint64_t p) {  // :37-43
        passes = p;
        uint32_t e = (uint32_t)(p > 1 ? p : 1) - 1;
        step_size = (uint64_t)1 << e;
        next_index[0] = next_index[1] = next_index[2] = 0;
        iterations = 0;
        total_iterations = 0;
    }
    bool next(uint64_t out[3]) {  // :50-76
        if (step_size == 0) return false;
        iterations += 1;
        total_iterations += 1;
        out[0] = next_index[0]; out[1] = next_index[1]; out[2] = next_index[2];
        next_index[0] += step_size;
        if (next_index[0] >= limits[0]) {
            next_index[0] = 0;
            next_index[1] += step_size;
            if (next_index[1] >= limits[1]) {
                next_index[1] = 0;
                next_index[2] += step_size;
                if (next_index[2] >= limits[2]) {
                    step_size = prev_power_of_2((uint32_t)(step_size - 1));
                    next_index[0] = next_index[1] = next_index[2] = 0;
                    iterations = 0;
                }
            }
        }
        return true;
    }
    uint64_t len() const {  // :80-89
        uint64_t s = step_size, it = 0;
        while (s > 0) {
            uint64_t a = (limits[0] + s - 1) / s, b = (limits[1] + s - 1) / s, c = (limits[2] + s - 1) / s;
            it += a * b * c;
            s = prev_power_of_2((uint32_t)(s - 1));
        }
        return it - iterations;
    }
    uint64_t passes_left() const {  // :99-105
        if (step_size == 0) return 0;
        return (uint64_t)log2f((float)step_size) + 1;
    }
};

// ----------------------------------------------------------------- SDFViewer

struct Sampler {
    int kind;  // 0 = demo direct, 1 = tape
    DemoParams demo;
    std::vector<unsigned char> tape_bytes;
    Tape tape;
    Sample sample(V3 p) const { return kind == 0 ? demo_sample(demo, p, false) : tape_sample(tape, p); }
};

struct Viewer {  // SDFViewer, src/app/scene/sdf/mod.rs:21-38
    uint32_t dims[3];
    float bb[6];
    std::vector<float> tex0, tex1;  // [f32;4] texels
    LoadingManager lm;
    bool has_changed_box = false;
    float changed_box[6];
    bool changed_box_while_loading = false;
    float lut[256];
};

// body of the loop, src/app/scene/sdf/mod.rs:196-208
inline void store_sample(const Viewer& v, Sample s, float* t0, float* t1) {
    float dd = 1e-1f + s.d;                                   // :196
    if (dd < 0.0f) dd = 0.0f; else if (dd > 1.0f) dd = 1.0f;  // f32::clamp
    t0[0] = dd;
    if (s.r == 0.f && s.g == 0.f && s.b == 0.f) { s.r = 0.5f; s.g = 0.5f; s.b = 0.5f; }  // :197-200
    t0[1] = v.lut[f32_to_u8_sat(s.r)];                        // :201-204
    t0[2] = v.lut[f32_to_u8_sat(s.g)];
    t0[3] = v.lut[f32_to_u8_sat(s.b)];
    t1[0] = s.metallic;                                       // :205
    t1[1] = s.roughness;                                      // :206
    t1[2] = (s.occlusion <= 0.0f) ? 1.0f : s.occlusion;       // :208
}

// voxel position, src/app/scene/sdf/mod.rs:179-182: three separately rounded ops
inline V3 voxel_pos(const Viewer& v, uint64_t x, uint64_t y, uint64_t z) {
    float sx = (float)v.dims[0] - 1.f, sy = (float)v.dims[1] - 1.f, sz = (float)v.dims[2] - 1.f;  // :161
    float bx = v.bb[3] - v.bb[0], by = v.bb[4] - v.bb[1], bz = v.bb[5] - v.bb[2];                // :160
    V3 p{(float)x, (float)y, (float)z};
    p.x = p.x / sx; p.y = p.y / sy; p.z = p.z / sz;  // :180
    p.x = p.x * bx; p.y = p.y * by; p.z = p.z * bz;  // :181
    p.x = p.x + v.bb[0]; p.y = p.y + v.bb[1]; p.z = p.z + v.bb[2];  // :182
    return p;
}

// ------------------------------------------------------------------- tracer

struct TraceParams {
    float origin[3], base[3], dx[3], dy[3], bvp[16];
    float bmin[3], bmax[3];
    uint32_t dims[3];
    float lod;
    uint32_t filter_linear;  // GL filter state: 0 NEAREST (until the first commit at lod 1), 1 LINEAR
    float tint[4];
    uint32_t tone_mapping, color_mapping;
    float gamma;
    float ambient[3];
    // multi-GPU slab: stored z range [z_lo, z_hi) of the arrays passed in, and the
    // sub-box this trace is clipped to (== bmin/bmax for a whole volume)
    uint32_t z_lo, z_hi;
    float clip_min[3], clip_max[3];
};

struct Vol {
    const float* tex;
    uint32_t W, H, D, z_lo, z_hi;
};

// GL MIRRORED_REPEAT (src/app/scene/sdf/mod.rs:113-115) on an integer texel index
inline int64_t mirror_idx(int64_t i, int64_t n) {
    int64_t m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m < n ? m : 2 * n - 1 - m;
}

inline const float* texel(const Vol& v, int64_t x, int64_t y, int64_t z) {
    x = mirror_idx(x, v.W); y = mirror_idx(y, v.H); z = mirror_idx(z, v.D);
    // slab storage: clamp into the stored range (the caller guarantees taps stay within the halo)
    if (z < (int64_t)v.z_lo) z = v.z_lo;
    if (z >= (int64_t)v.z_hi) z = v.z_hi - 1;
    return v.tex + 4 * (((uint64_t)(z - v.z_lo) * v.H + (uint64_t)y) * v.W + (uint64_t)x);
}

inline float lerp1(float a, float b, float f) { return a + f * (b - a); }

// texture(sampler3D, p01) with GL_LINEAR: texel centres at (i+0.5)/N, exact fp32 weights.
// (hardware uses ~8-bit fixed-point weights; this build defines the filter as exact fp32.)
void tex_linear(const Vol& v, const float p01[3], float out[4]) {
    float ux = p01[0] * (float)v.W - 0.5f, uy = p01[1] * (float)v.H - 0.5f, uz = p01[2] * (float)v.D - 0.5f;
    float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
    float fx = ux - fx0, fy = uy - fy0, fz = uz - fz0;
    int64_t x0 = (int64_t)fx0, y0 = (int64_t)fy0, z0 = (int64_t)fz0;
    const float* c000 = texel(v, x0, y0, z0);
    const float* c100 = texel(v, x0 + 1, y0, z0);
    const float* c010 = texel(v, x0, y0 + 1, z0);
    const float* c110 = texel(v, x0 + 1, y0 + 1, z0);
    const float* c001 = texel(v, x0, y0, z0 + 1);
    const float* c101 = texel(v, x0 + 1, y0, z0 + 1);
    const float* c011 = texel(v, x0, y0 + 1, z0 + 1);
    const float* c111 = texel(v, x0 + 1, y0 + 1, z0 + 1);
    for (int c = 0; c < 4; ++c) {
        float a = lerp1(c000[c], c100[c], fx), b = lerp1(c010[c], c110[c], fx);
        float e = lerp1(c001[c], c101[c], fx), f = lerp1(c011[c], c111[c], fx);
        out[c] = lerp1(lerp1(a, b, fy), lerp1(e, f, fy), fz);
    }
}

// texture() with GL_NEAREST: texel floor(p01*N), mirrored
void tex_nearest(const Vol& v, const float p01[3], float out[4]) {
    int64_t x = (int64_t)floorf(p01[0] * (float)v.W), y = (int64_t)floorf(p01[1] * (float)v.H),
            z = (int64_t)floorf(p01[2] * (float)v.D);
    const float* t = texel(v, x, y, z);
    for (int c = 0; c < 4; ++c) out[c] = t[c];
}

// sdfSampleRawInterp / sdfSampleRawNearest, src/app/scene/sdf/material.frag:27-53.  While loading
// (lod != 1) the coordinate is rounded to the lod lattice first (:33-34); texture() then applies
// the texture's current GL filter: NEAREST at creation (scene/sdf/mod.rs:110-111), switched to
// LINEAR by the first commit at lod == 1 and never switched back (:227-238, :241-250).
void sdf_sample_raw_interp(const TraceParams& P, const Vol& v, V3 p, float out[4]) {
    float p01[3] = {(p.x - P.bmin[0]) / (P.bmax[0] - P.bmin[0]), (p.y - P.bmin[1]) / (P.bmax[1] - P.bmin[1]),
                    (p.z - P.bmin[2]) / (P.bmax[2] - P.bmin[2])};  // :30,:44
    if (P.lod != 1.0f) {                                           // :43
        float rs[3] = {(float)P.dims[0] / P.lod, (float)P.dims[1] / P.lod, (float)P.dims[2] / P.lod};  // :33
        for (int i = 0; i < 3; ++i) p01[i] = floorf(p01[i] * rs[i] + 0.5f) / rs[i];  // :34 round(); ties up
    }
    if (P.filter_linear) tex_linear(v, p01, out);
    else tex_nearest(v, p01, out);
}

// sdfOutOfBoundsDist, material.frag:83-88 (against the clip box: the whole bbox on one GPU)
inline float oob_dist(const float bmin[3], const float bmax[3], V3 p) {
    float ox = fmaxf(bmin[0] - p.x, p.x - bmax[0]);
    float oy = fmaxf(bmin[1] - p.y, p.y - bmax[1]);
    float oz = fmaxf(bmin[2] - p.z, p.z - bmax[2]);
    return fmaxf(ox, fmaxf(oy, oz));
}

// three-d AmbientLight without environment map:
//   occlusion * ambientColor * mix(surface_color, vec3(0), metallic)
inline void calculate_lighting(const TraceParams& P, const float albedo[3], float metallic, float occlusion,
                               float out[3]) {
    for (int c = 0; c < 3; ++c) {
        float mixv = albedo[c] * (1.0f - metallic) + 0.0f * metallic;
        out[c] = 0.0f + (occlusion * P.ambient[c]) * mixv;
    }
}

// three-d ToneMapping::fragment_shader_source (0 none, 1 Reinhard, 2 ACES (Narkowicz fit), 3 filmic)
inline float tone_mapping(uint32_t type, float c) {
    if (type == 1u) {
        c = c / (c + 1.0f);
    } else if (type == 2u) {
        c = c * (2.51f * c + 0.03f) / (c * (2.43f * c + 0.59f) + 0.14f);
    } else if (type == 3u) {
        c = fmaxf(0.0f, c - 0.004f);
        c = (c * (6.2f * c + 0.5f)) / (c * (6.2f * c + 1.7f) + 0.06f);
        c = powf(c, 2.2f);
    } else {
        return c;  // ToneMapping::None
    }
    return fminf(fmaxf(c, 0.0f), 1.0f);
}

// three-d ColorMapping::fragment_shader_source (1 = compute to sRGB)
inline float color_mapping(uint32_t type, float c) {
    if (type == 1u) {
        float lo = c * 12.92f;
        float hi = 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
        return (c < 0.0031308f) ? lo : hi;  // mix(lo, hi, step(0.0031308, c))
    }
    return c;
}

void trace_pixel(const TraceParams& P, const Vol& v0, const Vol& v1, uint32_t i, uint32_t j, float* rgba,
                 float* depth, float* gbuf) {
    float fx = (float)i + 0.5f, fy = (float)j + 0.5f;
    V3 cam{P.origin[0], P.origin[1], P.origin[2]};
    V3 du{(P.base[0] + P.dx[0] * fx) + P.dy[0] * fy, (P.base[1] + P.dx[1] * fx) + P.dy[1] * fy,
          (P.base[2] + P.dx[2] * fx) + P.dy[2] * fy};
    float g[SDFGPU_GBUF_FLOATS_ORC] = {0};
#define MISS(code)                                              \
    do {                                                        \
        if (rgba) rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0.0f; \
        if (depth) *depth = 1.0f;                               \
        g[3] = (code);                                          \
        if (gbuf) memcpy(gbuf, g, sizeof(g));                   \
        return;                                                 \
    } while (0)
    // Fragment coverage: the rasterised cube (src/app/scene/sdf/mod.rs:254-282) == ray/AABB slab test.
    float ix = 1.0f / du.x, iy = 1.0f / du.y, iz = 1.0f / du.z;
    float t1x = (P.clip_min[0] - cam.x) * ix, t2x = (P.clip_max[0] - cam.x) * ix;
    float t1y = (P.clip_min[1] - cam.y) * iy, t2y = (P.clip_max[1] - cam.y) * iy;
    float t1z = (P.clip_min[2] - cam.z) * iz, t2z = (P.clip_max[2] - cam.z) * iz;
    float tmin = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fminf(t1z, t2z));
    float tmax = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fmaxf(t1z, t2z));
    if (!(tmax >= fmaxf(tmin, 0.0f))) MISS(-3.0f);
    // `pos`: front-face entry point, or the back-face exit point when the camera is inside.
    float tf = (tmin < 0.0f) ? tmax : tmin;
    V3 pos{cam.x + du.x * tf, cam.y + du.y * tf, cam.z + du.z * tf};
    // material.frag:133-139
    V3 rd{pos.x - cam.x, pos.y - cam.y, pos.z - cam.z};
    float inv = 1.0f / sqrtf(rd.x * rd.x + rd.y * rd.y + rd.z * rd.z);
    rd.x *= inv; rd.y *= inv; rd.z *= inv;
    V3 ro = pos;
    V3 probe{ro.x + rd.x * 0.2f, ro.y + rd.y * 0.2f, ro.z + rd.z * 0.2f};
    if (oob_dist(P.clip_min, P.clip_max, probe) > 0.0f) {
        ro = V3{cam.x + rd.x * 0.2f, cam.y + rd.y * 0.2f, cam.z + rd.z * 0.2f};
    }
    // sdfRaycast, material.frag:92-128 (maxSteps = 256, :142)
    V3 p = ro;
    float t = 0.0f;
    float s0[4] = {0, 0, 0, 0};
    float hit_w = -1.0f;
    int steps = 0;
    for (int it = 0; it < 256; ++it) {
        steps = it;
        if (it >= 255) { hit_w = -1.0f; break; }                                  // :99-102
        if (oob_dist(P.clip_min, P.clip_max, p) > 1e-4f) { hit_w = -2.0f; break; }  // :106-109
        sdf_sample_raw_interp(P, v0, p, s0);                                      // :112
        float dist = s0[0] - 1e-1f;                                               // :59
        if (dist < 1e-5f) { hit_w = t; break; }                                   // :117-121
        t += dist;                                                                // :124
        p.x += rd.x * dist; p.y += rd.y * dist; p.z += rd.z * dist;               // :125
    }
    g[0] = p.x; g[1] = p.y; g[2] = p.z; g[15] = (float)steps;
    if (hit_w < 0.0f) MISS(hit_w);  // :145-149
#undef MISS
    float s1[4];
    sdf_sample_raw_interp(P, v1, p, s1);  // :154
    g[3] = hit_w;
    memcpy(g + 4, s0, 16);
    memcpy(g + 8, s1, 16);
    if (gbuf) {
        // sdfNormal, material.frag:73-80
        float l[3] = {(float)P.dims[0] / P.lod, (float)P.dims[1] / P.lod, (float)P.dims[2] / P.lod};
        float h = 1.0f / sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
        const float k[4][3] = {{1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {1, 1, 1}};
        float n[3] = {0, 0, 0};
        for (int q = 0; q < 4; ++q) {
            float s[4];
            sdf_sample_raw_interp(P, v0, V3{p.x + k[q][0] * h, p.y + k[q][1] * h, p.z + k[q][2] * h}, s);
            float dq = s[0] - 1e-1f;
            n[0] += k[q][0] * dq; n[1] += k[q][1] * dq; n[2] += k[q][2] * dq;
        }
        float ninv = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        g[12] = n[0] * ninv; g[13] = n[1] * ninv; g[14] = n[2] * ninv;
        memcpy(gbuf, g, sizeof(g));
    }
    if (rgba) {
        float albedo[3] = {s0[1] * P.tint[0], s0[2] * P.tint[1], s0[3] * P.tint[2]};  // :158-159
        float col[3];
        calculate_lighting(P, albedo, s1[0], s1[2], col);                              // :163
        for (int c = 0; c < 3; ++c) {
            float x = tone_mapping(P.tone_mapping, col[c]);                            // :167
            x = color_mapping(P.color_mapping, x);                                     // :168
            if (P.gamma != 0.0f) x = powf(x, P.gamma);                                 // :171-173
            rgba[c] = x;
        }
        rgba[3] = P.tint[3];                                                           // :169
    }
    if (depth) {  // :180-181
        const float* m = P.bvp;
        float z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
        float w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15];
        *depth = z / w;
    }
}

}  // namespace

// =============================================================== C interface

ORC_API float orc_air_dist(void) { return air_dist(); }

ORC_API void orc_srgb_lut(float out[256]) {
    for (unsigned i = 0; i < 256; ++i) out[i] = srgb_u8_to_linear(i);
}

ORC_API uint32_t orc_f32_to_u8(float v) { return f32_to_u8_sat(v); }

ORC_API void orc_demo_params_default(DemoParams* P) {
    P->cube_half_side = 0.95f;
    P->cube_material = SDFT_MAT_BRICK;
    P->sphere_radius = 1.05f;
    P->sphere_material = SDFT_MAT_NORMAL;
    P->max_distance_custom_material = 0.05f;
    P->disable_sphere = 0;
}

// SDFDemo::sample at n points; out = n x 7 floats in SDFSample order
ORC_API void orc_demo_sample(const DemoParams* P, const float* pts, uint64_t n, int distance_only, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        Sample s = demo_sample(*P, V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, distance_only != 0);
        memcpy(out + 7 * i, &s, 28);
    }
}

ORC_API int orc_tape_sample(const void* tape, uint64_t len, const float* pts, uint64_t n, float* out) {
    Tape t;
    if (!tape_parse(tape, len, &t)) return -1;
    for (uint64_t i = 0; i < n; ++i) {
        Sample s = tape_sample(t, V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
        memcpy(out + 7 * i, &s, 28);
    }
    return 0;
}

// ---- LoadingManager
ORC_API void* orc_lm_new(const uint64_t limits[3], uint64_t passes) {
    LoadingManager* lm = new LoadingManager();
    lm->limits[0] = limits[0]; lm->limits[1] = limits[1]; lm->limits[2] = limits[2];
    lm->reset(passes);
    return lm;
}
ORC_API void orc_lm_free(void* h) { delete (LoadingManager*)h; }
ORC_API int orc_lm_next(void* h, uint64_t out[3]) { return ((LoadingManager*)h)->next(out) ? 1 : 0; }
ORC_API uint64_t orc_lm_len(void* h) { return ((LoadingManager*)h)->len(); }
ORC_API uint64_t orc_lm_total_iterations(void* h) { return ((LoadingManager*)h)->total_iterations; }
ORC_API uint64_t orc_lm_passes_left(void* h) { return ((LoadingManager*)h)->passes_left(); }
ORC_API uint32_t orc_prev_power_of_2(uint32_t x) { return prev_power_of_2(x); }

// ---- SDFViewer
// from_bb, src/app/scene/sdf/mod.rs:46-72
ORC_API void orc_dims_from_bb(const float bb[6], uint32_t max_voxels_side, uint32_t out[3]) {
    float sz[3] = {bb[3] - bb[0], bb[4] - bb[1], bb[5] - bb[2]};
    int max_dim = 0;  // Iterator::max_by returns the LAST maximum
    for (int i = 1; i < 3; ++i)
        if (sz[i] >= sz[max_dim]) max_dim = i;
    for (int i = 0; i < 3; ++i) {
        if (i == max_dim) out[i] = max_voxels_side;
        else {
            float f = (float)max_voxels_side * sz[i] / sz[max_dim];
            out[i] = (f != f || f <= 0.f) ? 0u : (f >= 4294967296.f ? 0xffffffffu : (uint32_t)f);
        }
    }
}

ORC_API void* orc_viewer_new(const float bb[6], const uint32_t dims[3], uint64_t passes) {
    Viewer* v = new Viewer();
    memcpy(v->dims, dims, 12);
    memcpy(v->bb, bb, 24);
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    v->tex0.assign(n * 4, air_dist());  // new_voxels, :76-77
    v->tex1.assign(n * 4, air_dist());
    v->lm.limits[0] = dims[0]; v->lm.limits[1] = dims[1]; v->lm.limits[2] = dims[2];
    v->lm.reset(passes);
    for (unsigned i = 0; i < 256; ++i) v->lut[i] = srgb_u8_to_linear(i);
    return v;
}
// Geometry only (no volumes): for point-wise checks of grids too large to hold twice on the host.
ORC_API void* orc_viewer_new_noalloc(const float bb[6], const uint32_t dims[3], uint64_t passes) {
    Viewer* v = new Viewer();
    memcpy(v->dims, dims, 12);
    memcpy(v->bb, bb, 24);
    v->lm.limits[0] = dims[0]; v->lm.limits[1] = dims[1]; v->lm.limits[2] = dims[2];
    v->lm.reset(passes);
    for (unsigned i = 0; i < 256; ++i) v->lut[i] = srgb_u8_to_linear(i);
    return v;
}
ORC_API void orc_viewer_free(void* h) { delete (Viewer*)h; }
ORC_API float* orc_viewer_tex0(void* h) { return ((Viewer*)h)->tex0.data(); }
ORC_API float* orc_viewer_tex1(void* h) { return ((Viewer*)h)->tex1.data(); }
ORC_API uint64_t orc_viewer_len(void* h) { return ((Viewer*)h)->lm.len(); }
ORC_API uint64_t orc_viewer_total_iterations(void* h) { return ((Viewer*)h)->lm.total_iterations; }
ORC_API uint64_t orc_viewer_passes_left(void* h) { return ((Viewer*)h)->lm.passes_left(); }
ORC_API void orc_voxel_pos(void* h, uint64_t x, uint64_t y, uint64_t z, float out[3]) {
    V3 p = voxel_pos(*(Viewer*)h, x, y, z);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}

ORC_API void* orc_sampler_demo(const DemoParams* P) {
    Sampler* s = new Sampler();
    s->kind = 0;
    s->demo = *P;
    return s;
}
ORC_API void* orc_sampler_tape(const void* tape, uint64_t len) {
    Sampler* s = new Sampler();
    s->kind = 1;
    s->tape_bytes.assign((const unsigned char*)tape, (const unsigned char*)tape + len);
    if (!tape_parse(s->tape_bytes.data(), len, &s->tape)) { delete s; return nullptr; }
    return s;
}
ORC_API void orc_sampler_free(void* s) { delete (Sampler*)s; }

// SDFViewer::update, src/app/scene/sdf/mod.rs:128-217.  `max_iterations` replaces
// max_delta_time (0 = until the LoadingManager is exhausted); at least one iteration runs.
ORC_API uint64_t orc_viewer_update(void* h, void* sampler, const float* changed, uint64_t max_iterations) {
    Viewer& v = *(Viewer*)h;
    const Sampler& sdf = *(const Sampler*)sampler;
    bool just_changed_box = false;
    if (changed) {  // :131-139
        if (v.has_changed_box) {  // merge_bounding_boxes, src/sdf/defaults.rs:59-72
            for (int i = 0; i < 3; ++i) {
                v.changed_box[i] = rust_min(v.changed_box[i], changed[i]);
                v.changed_box[3 + i] = rust_max(v.changed_box[3 + i], changed[3 + i]);
            }
        } else {
            memcpy(v.changed_box, changed, 24);
            v.has_changed_box = true;
        }
        v.changed_box_while_loading = v.lm.len() > 0 || v.changed_box_while_loading;
        just_changed_box = true;
    }
    if (v.has_changed_box) {  // :144-154
        if (v.lm.len() == 0) {
            v.lm.reset(3);
            if (!just_changed_box) {
                if (!v.changed_box_while_loading) v.has_changed_box = false;
                v.changed_box_while_loading = false;
            }
        }
    }
    uint64_t start_iter = v.lm.total_iterations;
    const float AIR = air_dist();
    bool first = true;
    uint64_t done = 0;
    while (first || max_iterations == 0 || done < max_iterations) {  // :173
        first = false;
        uint64_t idx[3];
        if (!v.lm.next(idx)) break;  // :175, :212-214
        done++;
        uint64_t flat = (idx[2] * v.dims[1] + idx[1]) * v.dims[0] + idx[0];  // :177
        V3 pos = voxel_pos(v, idx[0], idx[1], idx[2]);                        // :179-182
        bool update_required = v.tex0[4 * flat] == AIR;                       // :184
        if (v.has_changed_box) {                                              // :185-190
            const float* c = v.changed_box;
            update_required = update_required || (pos.x >= c[0] && pos.x <= c[3] && pos.y >= c[1] &&
                                                  pos.y <= c[4] && pos.z >= c[2] && pos.z <= c[5]);
        }
        if (update_required) store_sample(v, sdf.sample(pos), &v.tex0[4 * flat], &v.tex1[4 * flat]);  // :191-211
    }
    return v.lm.total_iterations - start_iter;  // :216
}

// The end state of a fresh viewer after its LoadingManager is exhausted: every voxel sampled
// once (pure map, so the visit order is immaterial).  OpenMP over z: the "all host cores" CPU
// baseline.  z range [z0, z1) lets the caller bound the sample.  Returns voxels written.
ORC_API uint64_t orc_viewer_fill_all(void* h, void* sampler, uint32_t z0, uint32_t z1, int threads) {
    Viewer& v = *(Viewer*)h;
    const Sampler& sdf = *(const Sampler*)sampler;
    if (z1 > v.dims[2]) z1 = v.dims[2];
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t z = z0; z < (int64_t)z1; ++z)
        for (uint64_t y = 0; y < v.dims[1]; ++y)
            for (uint64_t x = 0; x < v.dims[0]; ++x) {
                uint64_t flat = ((uint64_t)z * v.dims[1] + y) * v.dims[0] + x;
                store_sample(v, sdf.sample(voxel_pos(v, x, y, (uint64_t)z)), &v.tex0[4 * flat], &v.tex1[4 * flat]);
            }
    return (uint64_t)(z1 > z0 ? z1 - z0 : 0) * v.dims[1] * v.dims[0];
}

// What update() stores for the given voxels (idx = n x 3 voxel indices): out0 / out1 = n x 4 floats.
// Lets the full-size GPU tests check a random subset of a 512^3 / 1024^3 volume point by point.
ORC_API void orc_viewer_sample_voxels(void* h, void* sampler, const uint32_t* idx, uint64_t n, float* out0,
                                      float* out1, int threads) {
    Viewer& v = *(Viewer*)h;
    const Sampler& sdf = *(const Sampler*)sampler;
    const float AIR = air_dist();
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        float* t0 = out0 + 4 * i;
        float* t1 = out1 + 4 * i;
        for (int c = 0; c < 4; ++c) t0[c] = t1[c] = AIR;  // new_voxels, :76-77
        store_sample(v, sdf.sample(voxel_pos(v, idx[3 * i], idx[3 * i + 1], idx[3 * i + 2])), t0, t1);
    }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- camera (cgmath 0.18 Matrix4::look_at_rh / perspective, call site src/app/scene/mod.rs:82-95)
ORC_API void orc_look_at_rh(const float eye[3], const float center[3], const float up[3], float m[16]) {
    V3 f = cg_normalize(V3{center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]});
    V3 s{f.y * up[2] - f.z * up[1], f.z * up[0] - f.x * up[2], f.x * up[1] - f.y * up[0]};
    s = cg_normalize(s);
    V3 u{s.y * f.z - s.z * f.y, s.z * f.x - s.x * f.z, s.x * f.y - s.y * f.x};
    float es = eye[0] * s.x + eye[1] * s.y + eye[2] * s.z;
    float eu = eye[0] * u.x + eye[1] * u.y + eye[2] * u.z;
    float ef = eye[0] * f.x + eye[1] * f.y + eye[2] * f.z;
    float r[16] = {s.x, u.x, -f.x, 0.f, s.y, u.y, -f.y, 0.f, s.z, u.z, -f.z, 0.f, -es, -eu, ef, 1.f};
    memcpy(m, r, 64);
}
ORC_API void orc_perspective(float fovy_rad, float aspect, float near, float far, float m[16]) {
    float f = 1.0f / tanf(fovy_rad / 2.0f);  // cgmath: Rad::cot(fovy / 2)
    float r[16] = {f / aspect, 0, 0, 0, 0, f, 0, 0, 0, 0, (far + near) / (near - far), -1.f,
                   0, 0, (2.f * far * near) / (near - far), 0};
    memcpy(m, r, 64);
}

// ---- tracer
ORC_API uint64_t orc_trace_params_size(void) { return sizeof(TraceParams); }

// Rows [row0,row1) of a width x height frame (row 0 = bottom).  Output pointers address the
// full frame.  tex0/tex1 hold z slices [P->z_lo, P->z_hi).
ORC_API void orc_trace(const TraceParams* P, const float* tex0, const float* tex1, uint32_t width,
                       uint32_t height, uint32_t row0, uint32_t row1, float* rgba, float* depth, float* gbuf,
                       int threads) {
    Vol v0{tex0, P->dims[0], P->dims[1], P->dims[2], P->z_lo, P->z_hi};
    Vol v1{tex1, P->dims[0], P->dims[1], P->dims[2], P->z_lo, P->z_hi};
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t j = row0; j < (int64_t)row1; ++j)
        for (uint32_t i = 0; i < width; ++i) {
            size_t px = (size_t)j * width + i;
            trace_pixel(*P, v0, v1, i, (uint32_t)j, rgba ? rgba + 4 * px : nullptr, depth ? depth + px : nullptr,
                        gbuf ? gbuf + SDFGPU_GBUF_FLOATS_ORC * px : nullptr);
        }
}

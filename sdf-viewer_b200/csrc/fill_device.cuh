// fill_device.cuh -- device code of the grid-fill kernel: SDFViewer::update's inner loop
// (/root/reference/src/app/scene/sdf/mod.rs:173-215) for a whole LoadingManager pass
// (src/app/scene/sdf/loading.rs:50-76) in one launch.
//
// Compiled twice from this one source: ahead of time by nvcc (fill.cu: the tape interpreter and
// the built-in demo program) and at set_tape time by NVRTC (api.cu: a straight-line kernel
// specialised for the structure of the loaded tape).  Keep it free of host / libc includes.
//
// Shape: persistent CTAs (grid = SMs x resident CTAs), 256 threads = 8 warps.  A tile is
// 32 (x) x 8 (y) x V (z) lattice points: a warp owns one x-row of 32 consecutive voxels, so each
// of its two stores per voxel (tex0, tex1: one float4 each) is one 512-byte contiguous burst =
// four full 128-byte lines.  The tape image (lowered instructions, primitive table, constants,
// sRGB LUT, per-axis position tables) is staged into shared memory once per CTA with a TMA bulk
// copy (cp.async.bulk + mbarrier).  Every warp walks the tape in lock step (the tape is the same
// for all voxels, so instruction fetch never diverges).
//
// Arithmetic: IEEE binary32, one rounding per operation: -fmad=false -prec-div=true
// -prec-sqrt=true -ftz=false (the reference is Rust/WASM f32, which never fuses a multiply-add).
//
// Algorithmic bytes: 32 B written per sampled voxel (16 B tex0 + 16 B tex1); conditional passes
// also read the 4-byte tex0.r of each visited voxel.
#pragma once

#include "sdfgpu_device_types.h"

namespace sdfgpu {
namespace dev {

struct Smp {  // SDFSample, src/sdf/mod.rs:104-118
    float d, r, g, b, m, ro, o;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------ TMA bulk copy
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------- exact f32 helpers
// fmodf(x, 0.5) and fmodf(x, 0.25) for x >= 0 (`%` in cube.rs:192).  Exact: x*2 and
// trunc()*0.5 are exact scalings and the final subtraction yields fmod's (always
// representable) result.  Floats >= 2^24 are multiples of 0.5; inf/NaN give NaN.
__device__ __forceinline__ float fmod_p5(float x) {
    if (!(x < 16777216.0f)) return x * 0.0f;
    return x - truncf(x * 2.0f) * 0.5f;
}
__device__ __forceinline__ float fmod_p25(float x) {
    if (!(x < 8388608.0f)) return x * 0.0f;
    return x - truncf(x * 4.0f) * 0.25f;
}
// f32::signum (cube.rs:168) where it is reached: behind `x.abs() > side`, which is false for NaN,
// so only the +-1-with-the-sign-of-x case remains
__device__ __forceinline__ float signum_not_nan(float x) { return copysignf(1.0f, x); }

// sample_brick_texture, src/sdf/demo/cube.rs:181-222
__device__ __forceinline__ void brick_texture(float px, float py, float pz, float nx, float ny, float nz, Smp& s) {
    float u, v;
    const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
    if (ax > ay) {  // :206
        if (ax > az) { u = pz; v = py; } else { u = px; v = py; }
    } else if (ay > az) {  // :214
        u = pz; v = px;
    } else {
        u = px; v = py;
    }
    const float row_num = v * 4.0f;                      // v / BRICK_HEIGHT (0.25): exact either way
    const float brick_offset = floorf(row_num) * 0.25f;  // / 4.
    const float bx = fmod_p5(fabsf(u + brick_offset));
    const float by = fmod_p25(fabsf(v));
    const float max_cement = 0.2f / 2.0f * 0.25f;
    const bool cement = bx < max_cement || bx > 0.5f - max_cement || by < max_cement || by > 0.25f - max_cement;
    s.r = cement ? 56.f / 255.f : 150.f / 255.f;
    s.g = cement ? 70.f / 255.f : 24.f / 255.f;
    s.b = cement ? 60.f / 255.f : 10.f / 255.f;
    s.m = cement ? 0.4f : 0.2f;
    s.ro = cement ? 0.5f : 0.8f;
    s.o = cement ? 1.0f : 0.0f;
}

// distance of one primitive; q = p - centre.  sphere.rs:39 (cgmath magnitude: x*x + y*y + z*z
// summed left to right), cube.rs:81 (f32::max chain)
__device__ __forceinline__ float prim_distance(uint32_t shape, float qx, float qy, float qz, float size) {
    if (shape == SDFT_SHAPE_SPHERE) return sqrtf(qx * qx + qy * qy + qz * qz) - size;
    return fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) - size;
}

// One primitive with its material: SDFDemoCube::sample / SDFDemoSphere::sample
// (cube.rs:79-89, sphere.rs:37-47) generalised by the record's fields; shape and material are
// compile-time here (the lowered opcode carries them).
template <int SHAPE, int MAT>
__device__ __forceinline__ Smp prim_sample(const float4 geom, const float4 m0, const float4 m1, float px, float py,
                                           float pz) {
    const float qx = px - geom.x, qy = py - geom.y, qz = pz - geom.z;
    Smp s;
    float len = 0.0f;
    if (SHAPE == SDFT_SHAPE_SPHERE) {
        len = sqrtf(qx * qx + qy * qy + qz * qz);
        s.d = len - geom.w;
    } else {
        s.d = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) - geom.w;
    }
    // material evaluated unconditionally and dropped by selects when the sample is "air"
    // (the reference's distance_only shortcut, cube.rs:83-85): no divergent branch in the warp
    if (MAT == SDFT_MAT_FLAT) {
        s.r = m0.x; s.g = m0.y; s.b = m0.z; s.m = m0.w; s.ro = m1.x; s.o = m1.y;
    } else {
        float nx, ny, nz;
        if (SHAPE == SDFT_SHAPE_SPHERE) {  // cgmath normalize = v * (1 / |v|), sphere.rs:123
            const float inv = __frcp_rn(len);  // correctly rounded 1 / len == IEEE 1.0f / len
            nx = qx * inv; ny = qy * inv; nz = qz * inv;
        } else {  // cube.rs:164-177
            nx = fabsf(qx) > geom.w ? signum_not_nan(qx) : 0.0f;
            ny = fabsf(qy) > geom.w ? signum_not_nan(qy) : 0.0f;
            nz = fabsf(qz) > geom.w ? signum_not_nan(qz) : 0.0f;
        }
        if (MAT == SDFT_MAT_BRICK) {
            brick_texture(qx, qy, qz, nx, ny, nz, s);
        } else {  // Material::Normal, cube.rs:56
            s.r = fabsf(nx); s.g = fabsf(ny); s.b = fabsf(nz);
            s.m = s.ro = s.o = 0.0f;
        }
    }
    if (s.d > m1.z) s.r = s.g = s.b = s.m = s.ro = s.o = 0.0f;
    return s;
}

// runtime (shape, material): used once per voxel after a UNION_RANGE fold
__device__ __forceinline__ Smp prim_sample_rt(const float4 g, const float4 m0, const float4 m1, float px, float py,
                                              float pz) {
    const uint32_t kind = __float_as_uint(m1.w);
    switch (((kind & 0xffu) ? 3u : 0u) + ((kind >> 8) & 0xffu)) {
        case 0: return prim_sample<SDFT_SHAPE_SPHERE, SDFT_MAT_FLAT>(g, m0, m1, px, py, pz);
        case 1: return prim_sample<SDFT_SHAPE_SPHERE, SDFT_MAT_BRICK>(g, m0, m1, px, py, pz);
        case 2: return prim_sample<SDFT_SHAPE_SPHERE, SDFT_MAT_NORMAL>(g, m0, m1, px, py, pz);
        case 3: return prim_sample<SDFT_SHAPE_BOX_LINF, SDFT_MAT_FLAT>(g, m0, m1, px, py, pz);
        case 4: return prim_sample<SDFT_SHAPE_BOX_LINF, SDFT_MAT_BRICK>(g, m0, m1, px, py, pz);
        default: return prim_sample<SDFT_SHAPE_BOX_LINF, SDFT_MAT_NORMAL>(g, m0, m1, px, py, pz);
    }
}

// `Srgba::from(Vector3<f32>)`: (c * 255.0) as u8 -- saturating, NaN -> 0 (scene/sdf/mod.rs:201)
// cvt.rzi.u8.f32 truncates, saturates to [0, 255] and maps NaN to 0 -- Rust's `as u8` -- in one instruction
__device__ __forceinline__ uint32_t f32_to_u8_sat(float c) {
    uint32_t r;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(c * 255.0f));
    return r;
}

// st.global.cs: the volume is written once and is larger than L2 at the sizes that matter
__device__ __forceinline__ void store_texel(float4* p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ void stack_store(float* st, const Smp& s) {
    constexpr int NT = FILL_THREADS;
    st[0 * NT] = s.d; st[1 * NT] = s.r; st[2 * NT] = s.g; st[3 * NT] = s.b;
    st[4 * NT] = s.m; st[5 * NT] = s.ro; st[6 * NT] = s.o;
}
__device__ __forceinline__ void stack_load(const float* st, Smp& s) {
    constexpr int NT = FILL_THREADS;
    s.d = st[0 * NT]; s.r = st[1 * NT]; s.g = st[2 * NT]; s.b = st[3 * NT];
    s.m = st[4 * NT]; s.ro = st[5 * NT]; s.o = st[6 * NT];
}

// The store rules of scene/sdf/mod.rs:196-208 for one sample: tex0 = (clamp(0.1 + d, 0, 1),
// linear rgb of the u8-quantised colour, grey if the colour is exactly black), tex1 = (metallic,
// roughness, occlusion or 1 if <= 0, AIR_DIST -- never written by the reference, :76).
__device__ __forceinline__ void store_rules(Smp s, const float* lut, float air_dist, float4& t0, float4& t1) {
    t0.x = fminf(fmaxf(1e-1f + s.d, 0.0f), 1.0f);  // f32::clamp; NaN stays NaN below
    if (s.d != s.d) t0.x = s.d;
    if (s.r == 0.0f && s.g == 0.0f && s.b == 0.0f) { s.r = 0.5f; s.g = 0.5f; s.b = 0.5f; }
    t0.y = lut[f32_to_u8_sat(s.r)];
    t0.z = lut[f32_to_u8_sat(s.g)];
    t0.w = lut[f32_to_u8_sat(s.b)];
    t1.x = s.m;
    t1.y = s.ro;
    t1.z = (s.o <= 0.0f) ? 1.0f : s.o;
    t1.w = air_dist;
}

// ------------------------------------------------------------ tape machine
// Machine model of include/sdfgpu_tape.h for V voxels at once: A accumulator, T top of the sample
// stack (registers; deeper levels in shared memory), q the position register.
template <int V>
struct Machine {
    Smp A[V], T[V];
    float qx[V], qy[V], qz[V];
    float posx, posy, posz[V];
};

struct Env {
    const float4* geom;
    const float4* mat0;
    const float4* mat1;
    const float* consts;
    float* stack;          // this thread's column of the shared-memory stack
    const uint32_t* list;  // per-tile survivors of the culled UNION_RANGE
    uint32_t list_n;
    bool culled;
};

// One lowered instruction.  OPC is a compile-time DeviceOp: the interpreter reaches this through
// its jump table, the specialised kernels call it in sequence.
template <int V, int OPC>
__device__ __forceinline__ void exec_op(const uint4 I, Machine<V>& M, const Env& E) {
    constexpr int NT = FILL_THREADS;
    if constexpr (OPC >= DOP_PRIM && OPC < DOP_PRIM + 18) {
        constexpr int MODE = (OPC - DOP_PRIM) / 6, SHAPE = ((OPC - DOP_PRIM) % 6) / 3, MAT = (OPC - DOP_PRIM) % 3;
        const float4 g = E.geom[I.y], m0 = E.mat0[I.y], m1 = E.mat1[I.y];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const Smp y = prim_sample<SHAPE, MAT>(g, m0, m1, M.qx[v], M.qy[v], M.qz[v]);
            if (MODE == 0) M.A[v] = y;                                  // A = y
            else if (MODE == 1) { if (y.d < M.A[v].d) M.A[v] = y; }     // union(A, y): ties keep A
            else { if (y.d > M.A[v].d) M.A[v] = y; }                    // intersect(A, y)
        }
    } else if constexpr (OPC == DOP_UNION_RANGE) {
        // the left-to-right fold == argmin with ties to the lowest index, then one material evaluation
        const uint32_t n = E.culled ? E.list_n : I.z;
        float best_d[V];
        uint32_t best_k[V];
        {
            const uint32_t k = E.culled ? E.list[0] : I.y;
            const float4 g = E.geom[k];
            const uint32_t shape = __float_as_uint(E.mat1[k].w) & 0xffu;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                best_d[v] = prim_distance(shape, M.qx[v] - g.x, M.qy[v] - g.y, M.qz[v] - g.z, g.w);
                best_k[v] = k;
            }
        }
        for (uint32_t j = 1; j < n; ++j) {
            const uint32_t k = E.culled ? E.list[j] : I.y + j;
            const float4 g = E.geom[k];
            const uint32_t shape = __float_as_uint(E.mat1[k].w) & 0xffu;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float d = prim_distance(shape, M.qx[v] - g.x, M.qy[v] - g.y, M.qz[v] - g.z, g.w);
                if (d < best_d[v]) { best_d[v] = d; best_k[v] = k; }
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t k = best_k[v];
            M.A[v] = prim_sample_rt(E.geom[k], E.mat0[k], E.mat1[k], M.qx[v], M.qy[v], M.qz[v]);
        }
    } else if constexpr (OPC == DOP_PUSH_REG || OPC == DOP_PUSH_MEM) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (OPC == DOP_PUSH_MEM) stack_store(E.stack + (size_t)((I.z * V + v) * 7) * NT, M.T[v]);
            M.T[v] = M.A[v];
        }
    } else if constexpr (OPC >= DOP_POP_UNION && OPC <= DOP_POP_DEMO_DIFF_MEM) {
        constexpr int KIND = (OPC - DOP_POP_UNION) % 3;
        constexpr bool RELOAD = OPC >= DOP_POP_UNION_MEM;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            Smp& A = M.A[v];
            const Smp B = M.T[v];  // pushed first
            if (KIND == 0) {
                if (!(A.d < B.d)) A = B;  // union(B, A): ties keep B
            } else if (KIND == 1) {
                if (!(A.d > B.d)) A = B;  // intersect(B, A)
            } else {  // SDFDemo::sample, demo/mod.rs:58-73; B = box, A = sphere
                const float* c = E.consts + I.y;
                const float dist = fmaxf(B.d, -A.d);          // :58
                const float inter = fabsf(B.d) - fabsf(A.d);  // :60
                Smp s = (inter < 0.0f) ? B : A;               // :61
                if (fabsf(inter) <= c[0]) {                   // :62
                    s.r = c[1]; s.g = c[2]; s.b = c[3]; s.m = c[4]; s.ro = c[5]; s.o = c[6];
                }
                s.d = dist;                                   // :72
                A = s;
            }
            if (RELOAD) stack_load(E.stack + (size_t)((I.z * V + v) * 7) * NT, M.T[v]);
        }
    } else if constexpr (OPC >= DOP_D_NEG && OPC <= DOP_D_MIN) {
        const float imm = __uint_as_float(I.w);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float& d = M.A[v].d;
            if (OPC == DOP_D_NEG) d = -d;
            else if (OPC == DOP_D_ABS) d = fabsf(d);
            else if (OPC == DOP_D_ADD) d = d + imm;
            else if (OPC == DOP_D_MUL) d = d * imm;
            else if (OPC == DOP_D_MAX) d = fmaxf(d, imm);
            else d = fminf(d, imm);
        }
    } else if constexpr (OPC == DOP_M_SET) {
        const float* c = E.consts + I.y;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            Smp& A = M.A[v];
            A.r = c[0]; A.g = c[1]; A.b = c[2]; A.m = c[3]; A.ro = c[4]; A.o = c[5];
        }
    } else if constexpr (OPC == DOP_P_RESET) {
#pragma unroll
        for (int v = 0; v < V; ++v) { M.qx[v] = M.posx; M.qy[v] = M.posy; M.qz[v] = M.posz[v]; }
    } else if constexpr (OPC == DOP_P_SUB) {
        const float* c = E.consts + I.y;
#pragma unroll
        for (int v = 0; v < V; ++v) { M.qx[v] = M.qx[v] - c[0]; M.qy[v] = M.qy[v] - c[1]; M.qz[v] = M.qz[v] - c[2]; }
    } else if constexpr (OPC == DOP_P_MUL) {
        const float imm = __uint_as_float(I.w);
#pragma unroll
        for (int v = 0; v < V; ++v) { M.qx[v] = M.qx[v] * imm; M.qy[v] = M.qy[v] * imm; M.qz[v] = M.qz[v] * imm; }
    } else if constexpr (OPC == DOP_P_ABS) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (I.y & 1u) M.qx[v] = fabsf(M.qx[v]);
            if (I.y & 2u) M.qy[v] = fabsf(M.qy[v]);
            if (I.y & 4u) M.qz[v] = fabsf(M.qz[v]);
        }
    }
}

#define SDFGPU_STEP(OPC, PC) exec_op<V, (OPC)>(s_instr[(PC)], M, E);
#ifndef SDFGPU_JIT_BODY
#define SDFGPU_JIT_BODY
#endif
#define SDFGPU_DEMO_BODY                                                                                     \
    SDFGPU_STEP(DOP_PRIM + 0 * 6 + SDFT_SHAPE_BOX_LINF * 3 + SDFT_MAT_BRICK, 0)                              \
    SDFGPU_STEP(DOP_PUSH_REG, 1)                                                                             \
    SDFGPU_STEP(DOP_PRIM + 0 * 6 + SDFT_SHAPE_SPHERE * 3 + SDFT_MAT_NORMAL, 2)                               \
    SDFGPU_STEP(DOP_POP_DEMO_DIFF, 3)

// The lowered tape for the V positions in M.pos*: the specialised straight-line body, the built-in demo body, or
// the fetch / dispatch loop.
template <int V, int PROG>
__device__ __forceinline__ void run_tape(Machine<V>& M, const Env& E, const uint4* s_instr) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
        M.A[v].d = M.A[v].r = M.A[v].g = M.A[v].b = M.A[v].m = M.A[v].ro = M.A[v].o = 0.0f;
        M.T[v] = M.A[v];
        M.qx[v] = M.posx; M.qy[v] = M.posy; M.qz[v] = M.posz[v];
    }
    if constexpr (PROG == PROG_JIT) {
        SDFGPU_JIT_BODY
    } else if constexpr (PROG == PROG_DEMO) {
        SDFGPU_DEMO_BODY
    } else {
        for (uint32_t pc = 0;; ++pc) {
            const uint4 I = s_instr[pc];
            if (I.x == DOP_END) break;
#define C(n) case n: exec_op<V, n>(I, M, E); break;
            switch (I.x) {
                C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16) C(17) C(18)
                C(19) C(20) C(21) C(22) C(23) C(24) C(25) C(26) C(27) C(28) C(29) C(30) C(31) C(32) C(33) C(34)
                C(35) C(36) C(37) C(38)
                default: break;
            }
#undef C
        }
    }
}

// Point mode (FillParams::points): one arbitrary position per thread, the raw sample out.
template <int PROG>
__device__ __forceinline__ void point_body(const FillParams& P, const Env& E, const uint4* s_instr, uint32_t tile) {
    const uint32_t i = tile * FILL_THREADS + threadIdx.x;
    const bool ok = i < P.n_points;
    const uint32_t j = ok ? i : P.n_points - 1u;  // idle lanes of the last tile repeat the last point
    Machine<1> M;
    M.posx = P.points[3u * j]; M.posy = P.points[3u * j + 1u]; M.posz[0] = P.points[3u * j + 2u];
    run_tape<1, PROG>(M, E, s_instr);
    if (ok) {
        float* o = P.points_out + 7u * (size_t)i;
        const Smp& a = M.A[0];
        o[0] = a.d; o[1] = a.r; o[2] = a.g; o[3] = a.b; o[4] = a.m; o[5] = a.ro; o[6] = a.o;
    }
}

// One tile for one thread: V voxels at (lx, ly, lz0 .. lz0 + V - 1) of the lattice.  FULL: the tile
// lies entirely inside the lattice and the pass is unconditional -- no per-voxel predicates at all.
template <int V, int PROG, bool FULL>
__device__ __forceinline__ void tile_body(const FillParams& P, const Env& E, const uint4* s_instr, const float* s_lut,
                                          const float* s_px, const float* s_py, const float* s_pz, uint32_t lx,
                                          uint32_t ly, uint32_t lz0, size_t slice, size_t vstride,
                                          uint32_t& touched_local) {
    const bool row_ok = FULL || (lx < P.nx && ly < P.ny);
    const uint32_t gx = P.rx0 + lx * P.step, gy = P.ry0 + ly * P.step, gz0 = P.rz0 + lz0 * P.step;
    const size_t flat0 = row_ok ? (size_t)(gz0 - P.z_lo) * slice + (size_t)gy * P.W + gx : 0;

    Machine<V> M;
    M.posx = s_px[FULL ? gx : min(gx, P.W - 1u)];
    M.posy = s_py[FULL ? gy : min(gy, P.H - 1u)];
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        act[v] = FULL || (row_ok && lz0 + v < P.nz);
        M.posz[v] = s_pz[FULL ? gz0 + v * P.step : min(gz0 + v * P.step, P.D - 1u)];
    }
    if (!FULL && P.conditional) {  // scene/sdf/mod.rs:184-190
        const bool in_xy = P.has_box && M.posx >= P.box[0] && M.posx <= P.box[3] && M.posy >= P.box[1] && M.posy <= P.box[4];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (act[v]) {
                bool need;
                if (P.conditional == FILL_SKIP_KNOWN) {  // sampled so far: exactly the multiples of known_step
                    need = ((gx | gy | (gz0 + v * P.step)) & (P.known_step - 1u)) != 0u;
                } else {
                    need = P.conditional == FILL_READ &&
                           __ldg(reinterpret_cast<const float*>(P.tex0 + flat0 + v * vstride)) == P.air_dist;
                    need = need || (in_xy && M.posz[v] >= P.box[2] && M.posz[v] <= P.box[5]);
                }
                act[v] = need;
            }
        }
        // nothing to sample in this warp's row (passes over an already loaded region)
        bool any = false;
#pragma unroll
        for (int v = 0; v < V; ++v) { any = any || act[v]; touched_local += act[v] ? 1u : 0u; }
        if (!__any_sync(0xffffffffu, any)) return;
    }

    run_tape<V, PROG>(M, E, s_instr);

    // ---- the stores of scene/sdf/mod.rs:196-208
    float4* const out0 = P.tex0 + flat0;
    float4* const out1 = P.tex1 + flat0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (FULL || act[v]) {
            float4 t0, t1;
            store_rules(M.A[v], s_lut, P.air_dist, t0, t1);
            store_texel(out0 + v * vstride, t0);
            store_texel(out1 + v * vstride, t1);
        }
    }
}

// The exact-safe cull of a UNION_RANGE against the position box [lo, hi], by all FILL_THREADS threads of a CTA: of
// the n_src candidates -- primitive src[k], or first + k when src is null, ascending -- primitive k is dropped only if
// its lower distance bound over the box (minus a margin) exceeds some candidate's upper bound (plus a margin), so the
// argmin over the survivors is the argmin over the candidates at every point of the box; their order is kept (ordered
// ballot compaction), which keeps the lowest-index tie rule.  Survivors to out[], their number returned (CTA-uniform).
// The caller has made sure that nobody still reads out[], s_red, s_cnt.
__device__ __forceinline__ uint32_t cull_box(const float4* geom, const float4* mat1, const uint32_t* src, uint32_t first,
                                             uint32_t n_src, float lox, float hix, float loy, float hiy, float loz,
                                             float hiz, float* s_red, uint32_t* s_cnt, uint32_t* out) {
    constexpr int NT = FILL_THREADS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float EPS = 1e-5f;
    // pass 1: U = min over primitives of the upper bound
    float umin = __int_as_float(0x7f800000);
    for (uint32_t k = threadIdx.x; k < n_src; k += NT) {
        const uint32_t pk = src ? src[k] : first + k;
        const float4 g = geom[pk];
        const uint32_t shape = __float_as_uint(mat1[pk].w) & 0xffu;
        const float fx = fmaxf(fabsf(lox - g.x), fabsf(hix - g.x));
        const float fy = fmaxf(fabsf(loy - g.y), fabsf(hiy - g.y));
        const float fz = fmaxf(fabsf(loz - g.z), fabsf(hiz - g.z));
        float ub = (shape == SDFT_SHAPE_SPHERE) ? sqrtf(fx * fx + fy * fy + fz * fz) : fmaxf(fmaxf(fx, fy), fz);
        ub = ub - g.w;
        ub = ub + EPS * (1.0f + fabsf(ub));
        umin = fminf(umin, ub);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o));
    if (lane == 0) s_red[warp] = umin;
    __syncthreads();
    float U = s_red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) U = fminf(U, s_red[w]);
    // pass 2: ordered compaction of the survivors (index order keeps the tie rule)
    uint32_t base = 0;
    for (uint32_t k0 = 0; k0 < n_src; k0 += NT) {
        const uint32_t k = k0 + threadIdx.x;
        bool keep = false;
        uint32_t pk = 0;
        if (k < n_src) {
            pk = src ? src[k] : first + k;
            const float4 g = geom[pk];
            const uint32_t shape = __float_as_uint(mat1[pk].w) & 0xffu;
            const float nx_ = fmaxf(fmaxf(lox - g.x, g.x - hix), 0.0f);
            const float ny_ = fmaxf(fmaxf(loy - g.y, g.y - hiy), 0.0f);
            const float nz_ = fmaxf(fmaxf(loz - g.z, g.z - hiz), 0.0f);
            float lb = (shape == SDFT_SHAPE_SPHERE) ? sqrtf(nx_ * nx_ + ny_ * ny_ + nz_ * nz_)
                                                    : fmaxf(fmaxf(nx_, ny_), nz_);
            lb = lb - g.w;
            lb = lb - EPS * (1.0f + fabsf(lb));
            keep = !(lb > U);  // NaN bounds keep the primitive
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t wbase = base, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            const uint32_t c = s_cnt[w];
            if (w < warp) wbase += c;
            total += c;
        }
        if (keep) out[wbase + __popc(bal & ((1u << lane) - 1u))] = pk;
        base += total;
        __syncthreads();
    }
    return base;
}

// V = voxels per thread along z (lattice units).
template <int V, int PROG>
__device__ __forceinline__ void fill_body(const FillParams& P) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NT = FILL_THREADS;

    // ---- stage the tape image: one elected thread issues the bulk copy
    const uint32_t img_bytes = P.tape_img_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + img_bytes);  // img_bytes is a multiple of 16
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, img_bytes);
        for (uint32_t off = 0; off < img_bytes; off += 32768u) {
            const uint32_t n = min(32768u, img_bytes - off);
            bulk_g2s(smem + off, P.tape_img + off, n, bar);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const TapeImageHeader* hdr = reinterpret_cast<const TapeImageHeader*>(smem);
    const uint4* s_instr = reinterpret_cast<const uint4*>(smem + hdr->off_instr);
    const float* s_lut = reinterpret_cast<const float*>(smem + hdr->off_lut);
    const float* s_px = reinterpret_cast<const float*>(smem + hdr->off_px);
    const float* s_py = reinterpret_cast<const float*>(smem + hdr->off_py);
    const float* s_pz = reinterpret_cast<const float*>(smem + hdr->off_pz);

    // scratch after the image: [16 B bar][32 B reduce][32 B counts][cull list u32 x n_cull][stack floats]
    float* s_red = reinterpret_cast<float*>(smem + img_bytes + 16);
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem + img_bytes + 16 + 32);
    uint32_t* s_list = reinterpret_cast<uint32_t*>(smem + img_bytes + 16 + 64);
    const uint32_t n_cull = (hdr->flags & TAPE_FLAG_CULL) ? hdr->cull_count : 0u;
    const uint32_t cull_first = hdr->cull_first;

    Env E;
    E.geom = reinterpret_cast<const float4*>(smem + hdr->off_geom);
    E.mat0 = reinterpret_cast<const float4*>(smem + hdr->off_mat0);
    E.mat1 = reinterpret_cast<const float4*>(smem + hdr->off_mat1);
    E.consts = reinterpret_cast<const float*>(smem + hdr->off_consts);
    E.stack = reinterpret_cast<float*>(smem + img_bytes + 16 + 64 + ((n_cull * 4u + 15u) & ~15u)) + threadIdx.x;
    E.list = s_list;
    E.list_n = 0;
    E.culled = n_cull != 0 && P.points == nullptr;

    if constexpr (V == 1) {
        if (P.points) {  // kernel-uniform
            for (uint32_t tile = blockIdx.x; tile < P.tiles_z; tile += gridDim.x) point_body<PROG>(P, E, s_instr, tile);
            return;
        }
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = P.tiles_x * P.tiles_y * P.tiles_z;
    const size_t slice = (size_t)P.W * P.H;
    const size_t vstride = (size_t)P.step * slice;
    uint32_t touched_local = 0;

    // tile -> (tx, ty, tz), x fastest; advanced incrementally by gridDim.x per iteration.  tz is the tile's place in
    // the z ORDER: with n_boundary_tiles (multi-GPU fill of a whole slab) the order is first z tile, last z tile, then
    // the interior, so that the slices the neighbours need are complete early (FillParams)
    uint32_t tx = blockIdx.x % P.tiles_x, ty = (blockIdx.x / P.tiles_x) % P.tiles_y,
             tz = blockIdx.x / (P.tiles_x * P.tiles_y);
    const uint32_t sx = gridDim.x % P.tiles_x, sy = (gridDim.x / P.tiles_x) % P.tiles_y,
                   sz = gridDim.x / (P.tiles_x * P.tiles_y);

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- per-tile culling of the UNION_RANGE (exact-safe: a primitive is dropped only if
        // its lower bound over the tile exceeds some other primitive's upper bound)
        const uint32_t tzz = P.n_boundary_tiles ? (tz == 0u ? 0u : (tz == 1u ? P.tiles_z - 1u : tz - 1u)) : tz;
        if (n_cull) {
            __syncthreads();  // previous tile's readers of s_list are done
            const uint32_t lx0 = tx * FILL_TILE_X, ly0 = ty * FILL_TILE_Y, lz0 = tzz * V;
            const uint32_t lx1 = min(lx0 + FILL_TILE_X, P.nx) - 1, ly1 = min(ly0 + FILL_TILE_Y, P.ny) - 1,
                           lz1 = min(lz0 + V, P.nz) - 1;
            const float ax = s_px[P.rx0 + lx0 * P.step], bx = s_px[P.rx0 + lx1 * P.step];
            const float ay = s_py[P.ry0 + ly0 * P.step], by = s_py[P.ry0 + ly1 * P.step];
            const float az = s_pz[P.rz0 + lz0 * P.step], bz = s_pz[P.rz0 + lz1 * P.step];
            const float lox = fminf(ax, bx), hix = fmaxf(ax, bx);
            const float loy = fminf(ay, by), hiy = fmaxf(ay, by);
            const float loz = fminf(az, bz), hiz = fmaxf(az, bz);
            // a tile inside one cell of the coarse pre-cull starts from that cell's survivors, not from the whole range
            const uint32_t* src = nullptr;
            uint32_t n_src = n_cull;
            if (P.cell_lists) {
                const uint32_t cx0 = (P.rx0 + lx0 * P.step) >> CULL_CELL_SHIFT, cx1 = (P.rx0 + lx1 * P.step) >> CULL_CELL_SHIFT;
                const uint32_t cy0 = (P.ry0 + ly0 * P.step) >> CULL_CELL_SHIFT, cy1 = (P.ry0 + ly1 * P.step) >> CULL_CELL_SHIFT;
                const uint32_t cz0 = (P.rz0 + lz0 * P.step) >> CULL_CELL_SHIFT, cz1 = (P.rz0 + lz1 * P.step) >> CULL_CELL_SHIFT;
                if (cx0 == cx1 && cy0 == cy1 && cz0 == cz1) {
                    const uint32_t cell = (cz0 * P.cells_y + cy0) * P.cells_x + cx0;
                    src = P.cell_lists + (size_t)cell * n_cull;
                    n_src = P.cell_counts[cell];
                }
            }
            const uint32_t base = cull_box(E.geom, E.mat1, src, cull_first, n_src, lox, hix, loy, hiy, loz, hiz, s_red, s_cnt, s_list);
            E.list_n = base;
            if (P.cull_stats && threadIdx.x == 0) {
                atomicAdd(P.cull_stats, (unsigned long long)base);
                atomicAdd(P.cull_stats + 1, 1ull);
                atomicMax(P.cull_stats + 2, (unsigned long long)base);
            }
        }

        const uint32_t lx = tx * FILL_TILE_X + lane;
        const uint32_t ly = ty * FILL_TILE_Y + warp;
        const uint32_t lz0 = tzz * V;
        // advance to this CTA's next tile (mixed-radix add with carries)
        tx += sx; ty += sy; tz += sz;
        if (tx >= P.tiles_x) { tx -= P.tiles_x; ++ty; }
        if (ty >= P.tiles_y) { ty -= P.tiles_y; ++tz; }

        // a tile that lies entirely inside the lattice needs no per-voxel predicates (CTA-uniform test)
        const bool full = !P.conditional && (lx - lane) + FILL_TILE_X <= P.nx && (ly - warp) + FILL_TILE_Y <= P.ny &&
                          lz0 + V <= P.nz;
        if (full) tile_body<V, PROG, true>(P, E, s_instr, s_lut, s_px, s_py, s_pz, lx, ly, lz0, slice, vstride, touched_local);
        else tile_body<V, PROG, false>(P, E, s_instr, s_lut, s_px, s_py, s_pz, lx, ly, lz0, slice, vstride, touched_local);

        if (tile < P.n_boundary_tiles) {  // CTA-uniform
            // the copy engines and the stream's semaphore are outside this GPU's SMs: system-scope fence
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0 && atomicAdd(P.boundary_count, 1u) + 1u == P.n_boundary_tiles) {
                *P.boundary_count = 0u;  // every boundary tile has been counted: ready for the next launch
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t*>(P.boundary_flag) = P.boundary_epoch;
            }
        }
    }

    if (P.touched) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) touched_local += __shfl_xor_sync(0xffffffffu, touched_local, o);
        if (lane == 0 && touched_local) atomicAdd(P.touched, (unsigned long long)touched_local);
    }
}

}  // namespace dev
}  // namespace sdfgpu

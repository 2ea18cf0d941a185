/*
 * sdfgpu_tape.h -- binary format of the SDF instruction tape.
 *
 * The reference has NO op tree: an SDF is an opaque `SDFSurface::sample(p)`
 * callback (/root/reference/src/sdf/mod.rs:33-43) reached point by point
 * through a WASM sandbox (src/sdf/wasm/native.rs:188-217).  A GPU cannot call
 * that callback, so this build defines a flat tape that an SDF author (or the
 * built-in demo, src/sdf/demo/) lowers to once, and that both the CUDA fill
 * kernel and the CPU oracle interpret.  The tape is the GPU-side stand-in for
 * the `sdf: impl SDFSurface` argument of `SDFViewer::update`
 * (src/app/scene/sdf/mod.rs:128).
 *
 * Machine model (per voxel):
 *   P   : position register, 3 x f32, starts at the voxel position
 *   A   : accumulator, one SDF sample = 7 x f32 in SDFSample order
 *         (distance, r, g, b, metallic, roughness, occlusion; src/sdf/mod.rs:104-118)
 *   S[] : small stack of samples (depth <= SDFT_MAX_STACK)
 * The value of A after the last instruction is the result of sample(p, false).
 *
 * All arithmetic is IEEE-754 binary32, one rounding per operation, never
 * fused (the reference is Rust/WASM f32, which never contracts).
 *
 * Layout (little endian):
 *   sdft_header | sdft_instr[n_instr] | sdft_prim[n_prims] | float consts[n_consts] | sdft_sop[n_sops]
 * (the last section is optional: n_sops = header.reserved[0], 0 in tapes without scalar programs)
 */
#ifndef SDFGPU_TAPE_H
#define SDFGPU_TAPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFT_MAGIC 0x54464453u /* "SDFT" */
#define SDFT_VERSION 1u
#define SDFT_MAX_STACK 8u
#define SDFT_MAX_INSTR 4096u
#define SDFT_MAX_PRIMS 4096u
#define SDFT_MAX_CONSTS 4096u
#define SDFT_MAX_SOPS 8192u

typedef struct sdft_header {
    uint32_t magic;    /* SDFT_MAGIC */
    uint32_t version;  /* SDFT_VERSION */
    uint32_t n_instr;
    uint32_t n_prims;
    uint32_t n_consts;
    uint32_t reserved[3]; /* [0] = n_sops (scalar-program section, below); [1], [2] = 0 */
} sdft_header; /* 32 bytes */

/* One instruction: 16 bytes so the interpreter fetches it with one 128-bit load. */
typedef struct sdft_instr {
    uint32_t op;  /* sdft_op */
    uint32_t a;   /* first operand: primitive index / constant index / axis mask */
    uint32_t b;   /* second operand: count */
    float imm;    /* immediate for D_* ops */
} sdft_instr;

/*
 * One primitive: 48 bytes = three float4 rows.
 *   row0: centre xyz, size (sphere radius | box half side)
 *   row1: colour rgb, metallic               (FLAT material only)
 *   row2: roughness, occlusion, air_skip, kind bits (uint32 reinterpreted)
 * `air_skip`: when the primitive's distance is > air_skip the material is not
 * evaluated and the sample is SDFSample::new(d, 0) -- the reference demo's
 * "the air has no texture" shortcut (src/sdf/demo/cube.rs:83-85,
 * sphere.rs:41-43; value 0.1 there).  +inf disables it.
 */
typedef struct sdft_prim {
    float center[3];
    float size;
    float color[3];
    float metallic;
    float roughness;
    float occlusion;
    float air_skip;
    uint32_t kind; /* shape | (material << 8) */
} sdft_prim;

/*
 * Scalar programs: arbitrary straight-line arithmetic for SDFs that are not made of the primitives above
 * (what a `sample` function compiled to WASM lowers to, sdfgpu_wasm_lower).  One op: 16 bytes.  A program is
 * the run sops[a .. a+b) named by an SDFT_OP_SCALAR instruction; value i of the program is the result of its
 * i-th op, an untyped 32-bit word (f32 or i32, as the consuming op reads it); operands name EARLIER values
 * of the same program by their index relative to the program's first op.  Semantics are WebAssembly's
 * (f32 ops IEEE-754 with one rounding each; min / max propagate NaN and order -0 < +0; nearest rounds half
 * to even; float -> int conversions saturate, NaN -> 0; shifts use the count modulo 32).
 */
typedef struct sdft_sop {
    uint32_t op; /* sdft_sop_op */
    uint32_t a, b, c;
} sdft_sop;

enum sdft_sop_op {
    SDFT_S_PX = 0, SDFT_S_PY = 1, SDFT_S_PZ = 2, /* the position register P                        */
    SDFT_S_CONST = 3,      /* bits of consts[a] (ABSOLUTE constant index): runtime data, a parameter edit keeps the kernel */
    SDFT_S_IMM = 4,        /* the word a itself: part of the program's structure              */
    /* f32 -> f32 */
    SDFT_S_FNEG = 8, SDFT_S_FABS = 9, SDFT_S_FSQRT = 10, SDFT_S_FFLOOR = 11, SDFT_S_FCEIL = 12, SDFT_S_FTRUNC = 13,
    SDFT_S_FNEAREST = 14,
    SDFT_S_FADD = 16, SDFT_S_FSUB = 17, SDFT_S_FMUL = 18, SDFT_S_FDIV = 19, SDFT_S_FMIN = 20, SDFT_S_FMAX = 21,
    SDFT_S_FCOPYSIGN = 22,
    SDFT_S_FMOD = 23,      /* C fmodf: the exact remainder of a / b with the sign of a (Rust's `%`, which a WASM guest reaches
                              through its libm; not a WebAssembly instruction)                                              */
    /* f32 x f32 -> 0 / 1 */
    SDFT_S_FEQ = 24, SDFT_S_FNE = 25, SDFT_S_FLT = 26, SDFT_S_FGT = 27, SDFT_S_FLE = 28, SDFT_S_FGE = 29,
    /* i32 */
    SDFT_S_IADD = 32, SDFT_S_ISUB = 33, SDFT_S_IMUL = 34, SDFT_S_IAND = 35, SDFT_S_IOR = 36, SDFT_S_IXOR = 37,
    SDFT_S_ISHL = 38, SDFT_S_ISHR_U = 39, SDFT_S_ISHR_S = 40,
    SDFT_S_IDIV_S = 41, SDFT_S_IDIV_U = 42, SDFT_S_IREM_S = 43, /* where WebAssembly traps (divisor 0, INT_MIN / -1) the result is 0 */
    SDFT_S_IEQ = 44, SDFT_S_INE = 45, SDFT_S_ILT_S = 46, SDFT_S_ILT_U = 47, SDFT_S_IGT_S = 48, SDFT_S_IGT_U = 49,
    SDFT_S_ILE_S = 50, SDFT_S_ILE_U = 51, SDFT_S_IGE_S = 52, SDFT_S_IGE_U = 53, SDFT_S_IEQZ = 54,
    SDFT_S_IREM_U = 55,
    /* mixed */
    SDFT_S_SELECT = 56,    /* v[a] != 0 ? v[b] : v[c]                                        */
    SDFT_S_F_FROM_I_S = 57, SDFT_S_F_FROM_I_U = 58, /* f32.convert_i32_s / _u              */
    SDFT_S_I_FROM_F_S = 59, SDFT_S_I_FROM_F_U = 60, /* i32.trunc_sat_f32_s / _u            */
    SDFT_S_OUT = 63        /* A[b] = v[a] as f32; b = 0..6: distance, r, g, b, metallic, roughness, occlusion.
                            * Yields no value: an operand that names an OUT op is rejected (SDFGPU_ERR_TAPE) */
};

enum sdft_shape {
    SDFT_SHAPE_SPHERE = 0,   /* sqrt(dx^2+dy^2+dz^2) - size   (sphere.rs:39) */
    SDFT_SHAPE_BOX_LINF = 1  /* max(|dx|,|dy|,|dz|) - size    (cube.rs:81)   */
};

enum sdft_material {
    SDFT_MAT_FLAT = 0,   /* colour/metallic/roughness/occlusion from the record */
    SDFT_MAT_BRICK = 1,  /* procedural brick, tri-planar (cube.rs:181-222)      */
    SDFT_MAT_NORMAL = 2  /* colour = |normal|, m=r=o=0 (cube.rs:56)             */
};

enum sdft_op {
    SDFT_OP_END = 0,
    /* ---- primitives (a = primitive index) ---- */
    SDFT_OP_PRIM = 1,         /* A = prim[a](P)                                   */
    SDFT_OP_UNION_PRIM = 2,   /* A = union(A, prim[a](P))                         */
    SDFT_OP_INTER_PRIM = 3,   /* A = intersect(A, prim[a](P))                     */
    SDFT_OP_UNION_RANGE = 4,  /* A = prim[a] U prim[a+1] U ... U prim[a+b-1], b>=1 */
    /* ---- sample stack ---- */
    SDFT_OP_PUSH = 8,         /* S[sp++] = A                                      */
    SDFT_OP_POP_UNION = 9,    /* B = S[--sp]; A = union(B, A)                     */
    SDFT_OP_POP_INTER = 10,   /* B = S[--sp]; A = intersect(B, A)                 */
    SDFT_OP_POP_DEMO_DIFF = 11, /* B = S[--sp]; A = demo_diff(B, A, consts[a..a+6]),
                                   the SDFDemo combinator (demo/mod.rs:58-73):
                                   consts = seam threshold, colour rgb, m, r, o   */
    /* ---- distance channel of A ---- */
    SDFT_OP_D_NEG = 16,       /* A.d = -A.d                                       */
    SDFT_OP_D_ABS = 17,       /* A.d = |A.d|                                      */
    SDFT_OP_D_ADD = 18,       /* A.d = A.d + imm                                  */
    SDFT_OP_D_MUL = 19,       /* A.d = A.d * imm                                  */
    SDFT_OP_D_MAX = 20,       /* A.d = max(A.d, imm)                              */
    SDFT_OP_D_MIN = 21,       /* A.d = min(A.d, imm)                              */
    /* ---- material channels of A ---- */
    SDFT_OP_M_SET = 24,       /* A.rgb,m,r,o = consts[a..a+5]                     */
    /* ---- position register ---- */
    SDFT_OP_P_RESET = 32,     /* P = voxel position                               */
    SDFT_OP_P_SUB = 33,       /* P = P - consts[a..a+2]                           */
    SDFT_OP_P_MUL = 34,       /* P = P * imm                                      */
    SDFT_OP_P_ABS = 35,       /* P.i = |P.i| for each axis bit set in a           */
    /* ---- scalar program ---- */
    SDFT_OP_SCALAR = 40       /* run sops[a .. a+b): reads P, its SDFT_S_OUT ops write A (channels it does not write keep
                                 their value).  Evaluated by the kernel specialised for the tape (NVRTC); a box without
                                 NVRTC rejects such a tape instead of falling back */
};

/*
 * union(X, Y)     = (Y.d < X.d) ? Y : X      -- whole sample; ties keep X (first)
 * intersect(X, Y) = (Y.d > X.d) ? Y : X
 * UNION_RANGE folds left to right with the same rule, so the winner is the
 * lowest-index primitive among those with the minimum distance.
 */

#ifdef __cplusplus
}
#endif
#endif /* SDFGPU_TAPE_H */

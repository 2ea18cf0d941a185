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
 *   sdft_header | sdft_instr[n_instr] | sdft_prim[n_prims] | float consts[n_consts]
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

typedef struct sdft_header {
    uint32_t magic;    /* SDFT_MAGIC */
    uint32_t version;  /* SDFT_VERSION */
    uint32_t n_instr;
    uint32_t n_prims;
    uint32_t n_consts;
    uint32_t reserved[3];
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
    SDFT_OP_P_ABS = 35        /* P.i = |P.i| for each axis bit set in a           */
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

// sdfgpu_internal.h -- structs shared by the host side (api.cu) and the kernels.
// Paths in comments are relative to /root/reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdfgpu.h"
#include "../../include/sdfgpu_tape.h"

namespace sdfgpu {

// ---- shared-memory image of a tape, built once per set_tape on the host and
// bulk-copied (cp.async.bulk, TMA) into every CTA of the fill kernel.
//   [0]                TapeImageHeader (64 B)
//   [off_instr]        sdft_instr[n_instr]                16 B each
//   [off_geom]         float4 geom[n_prims]   centre.xyz, size
//   [off_mat0]         float4 mat0[n_prims]   colour.rgb, metallic
//   [off_mat1]         float4 mat1[n_prims]   roughness, occlusion, air_skip, kind bits
//   [off_consts]       float consts[n_consts]
//   [off_lut]          float lut[256]         sRGB u8 -> linear (scene/sdf/mod.rs:201)
//   [off_px/py/pz]     float pos tables       voxel position per index and axis (scene/sdf/mod.rs:179-182)
// every offset is a multiple of 16 bytes; total is a multiple of 16 bytes.
struct TapeImageHeader {
    uint32_t n_instr, n_prims, n_consts, flags;
    uint32_t off_instr, off_geom, off_mat0, off_mat1;
    uint32_t off_consts, off_lut, off_px, off_py;
    uint32_t off_pz, max_stack, cull_first, cull_count;  // UNION_RANGE [cull_first, cull_first+cull_count) may be culled per tile
};
static_assert(sizeof(TapeImageHeader) == 64, "header is four float4 rows");

enum : uint32_t {
    TAPE_FLAG_CULL = 1u  // the tape holds exactly one UNION_RANGE, reached with P == voxel position
};

// ---- lowered (device) opcodes.  sdfgpu_set_tape translates the public tape (sdfgpu_tape.h) into
// this dense set so the interpreter dispatches through one jump table and each primitive op is
// already specialised by shape and material:
//   * PRIM / UNION_PRIM / INTER_PRIM a  ->  DOP_PRIM + mode*6 + shape*3 + material
//   * the top of the sample stack lives in registers; PUSH / POP_* carry in their opcode whether
//     a deeper level has to be spilled to / reloaded from shared memory (depth is static).
enum DeviceOp : uint32_t {
    DOP_END = 0,
    DOP_PRIM = 1,  // 18 variants: mode (0 set, 1 union, 2 intersect) * 6 + shape * 3 + material
    DOP_UNION_RANGE = 19,
    DOP_PUSH_REG = 20,   // T = A                      (stack was empty)
    DOP_PUSH_MEM = 21,   // spill T to level b, T = A   (b = depth before the push - 1)
    DOP_POP_UNION = 22,  // +0 union, +1 intersect, +2 demo_diff; B = T
    DOP_POP_INTER = 23,
    DOP_POP_DEMO_DIFF = 24,
    DOP_POP_UNION_MEM = 25,  // same, then reload T from level b (b = depth after the pop - 1)
    DOP_POP_INTER_MEM = 26,
    DOP_POP_DEMO_DIFF_MEM = 27,
    DOP_D_NEG = 28,
    DOP_D_ABS = 29,
    DOP_D_ADD = 30,
    DOP_D_MUL = 31,
    DOP_D_MAX = 32,
    DOP_D_MIN = 33,
    DOP_M_SET = 34,
    DOP_P_RESET = 35,
    DOP_P_SUB = 36,
    DOP_P_MUL = 37,
    DOP_P_ABS = 38,
    DOP_COUNT = 39
};

constexpr int FILL_THREADS = 256;  // 8 warps: a tile is 32 (x) x 8 (y) x V (z) lattice points
constexpr int FILL_TILE_X = 32;
constexpr int FILL_TILE_Y = 8;

struct FillParams {
    float4* tex0;  // stored slab, slice z_lo first
    float4* tex1;
    const unsigned char* tape_img;  // global copy of the shared-memory image (16 B aligned)
    uint32_t tape_img_bytes;        // multiple of 16
    uint32_t W, H, D;               // global grid
    uint32_t z_lo;                  // first stored slice
    // lattice region visited: index = r0 + i*step for i in [0, n), per axis
    uint32_t rx0, ry0, rz0;
    uint32_t nx, ny, nz;
    uint32_t step;
    uint32_t tiles_x, tiles_y, tiles_z;  // in lattice units
    uint32_t conditional;  // 1: sample iff tex0.r == AIR_DIST or position in box (scene/sdf/mod.rs:184-190)
    uint32_t has_box;
    float box[6];          // pending changed box
    float air_dist;
    uint32_t streaming_stores;
    uint32_t stack_floats;  // shared-memory floats reserved for the sample stack
    unsigned long long* touched;  // optional counter of voxels sampled (may be null)
};

struct TraceParams {
    const float4* tex0;
    const float4* tex1;
    float origin[3], base[3], dx[3], dy[3], bvp[16];
    float bmin[3], bmax[3];          // sdfBoundsMin/Max
    float clip_min[3], clip_max[3];  // == bounds on one GPU; the slab's sub-box for sort-last
    uint32_t W, H, D;
    uint32_t z_lo, z_hi;  // stored slices
    float lod;
    uint32_t filter_linear;  // GL filter of both textures: 0 NEAREST, 1 LINEAR (scene/sdf/mod.rs:110-111,241-250)
    float tint[4];
    uint32_t tone_mapping, color_mapping;
    float gamma;
    float ambient[3];
    uint32_t width, height;
    float4* rgba;              // may be null
    float* depth;              // may be null
    float* gbuf;               // may be null
    unsigned long long* keys;  // may be null (sort-last compositing keys)
};

// launchers (fill.cu / trace.cu)
cudaError_t launch_fill(const FillParams& p, int voxels_per_thread, int grid_ctas, size_t smem_bytes,
                        cudaStream_t s);
size_t fill_smem_bytes(uint32_t tape_img_bytes, uint32_t n_cull, uint32_t max_stack, int voxels_per_thread,
                       uint32_t* stack_floats);
int fill_max_ctas_per_sm(int voxels_per_thread, size_t smem_bytes);
cudaError_t fill_prepare(size_t smem_bytes);  // opt in to > 48 KB dynamic shared memory
cudaError_t launch_set_const(float4* dst, size_t n_texels, float v, int grid_ctas, cudaStream_t s);
cudaError_t launch_trace(const TraceParams& p, int variant, cudaStream_t s);
cudaError_t launch_keys_unpack(const unsigned long long* keys, uint32_t n, uint8_t* rgba8, float* depth,
                               cudaStream_t s);

}  // namespace sdfgpu

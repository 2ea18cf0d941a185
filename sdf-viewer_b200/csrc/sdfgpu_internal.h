// sdfgpu_internal.h -- structs shared by the host side (api.cu) and the kernels.
// Paths in comments are relative to /root/reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdfgpu.h"
#include "sdfgpu_tape.h"

#include "sdfgpu_device_types.h"

namespace sdfgpu {

constexpr uint32_t TRACE_MAX_BANDS = 32;
constexpr uint32_t TRACE_OUTSIDE_RUN = 8;  // tiles without a ray per CTA of trace_tiles_kernel

struct TraceParams {
    const float4* tex0;
    const float4* tex1;
    const float* dist;  // dist_mode 1: tex0.r of every stored texel as a dense array
    unsigned long long dist_tex;  // dist_mode 2 / 3: cudaTextureObject_t over an R32F 3-D array (point / linear filter)
    uint32_t full_dist;           // exact multi-GPU trace: `dist` covers the WHOLE grid (replicated on every rank)
    uint32_t own_z0, own_z1;      //   and this rank shades only the hits whose lower z tap lies in [own_z0, own_z1)
    uint32_t dist_mode;           // where the LINEAR march reads distances: 0 tex0.r, 1 dense array, 2 TMU point, 3 TMU linear
    float origin[3], base[3], dx[3], dy[3], bvp[16];
    float bmin[3], bmax[3];          // sdfBoundsMin/Max
    float clip_min[3], clip_max[3];  // == bounds on one GPU; the slab's sub-box for sort-last
    float inv_size[3];               // exact 1 / (bmax - bmin), valid when size_pow2
    uint32_t size_pow2;              // all three box sizes are powers of two: x / size == x * inv_size bit for bit
    uint32_t W, H, D;
    uint32_t z_lo, z_hi;  // stored slices
    float lod;
    uint32_t filter_linear;  // GL filter of both textures: 0 NEAREST, 1 LINEAR (scene/sdf/mod.rs:110-111,241-250)
    float tint[4];
    uint32_t tone_mapping, color_mapping;
    float gamma;
    float ambient[3];
    uint32_t width, height;
    uint32_t max_steps;         // sdfRaycast's maxSteps: 256 (material.frag:142)
    uint32_t tiles_x, tiles_y;  // 8 x 8 pixel tiles
    uint32_t rect[4];           // tile rectangle [x0, y0, x1, y1) that contains every pixel whose ray can hit the clip box
    // trace_tiles_kernel can announce the frame band by band (sdfgpu_trace_rgba8: the rows of a finished band cross
    // PCIe while the others are still being traced): the CTAs come band by band -- band_rows tile rows each, in the
    // order of band_order -- and the last CTA of a band to finish stores band_epoch into band_flags[band]
    uint32_t n_bands, band_rows, band_epoch;  // n_bands == 0: one band, no flags
    uint8_t band_order[TRACE_MAX_BANDS];
    uint32_t* band_done;                      // n_bands counters (zero between frames)
    uint32_t* band_flags;                     // n_bands flags, awaited by stream memory operations
    // trace_tiles_kernel: the order of the tiles inside the rectangle (tile ids ty * tiles_x + tx, band by band; null:
    // row by row) and where it records this frame's longest march per tile (null: nowhere) -- see tile_order_kernel
    const uint32_t* tile_order;
    uint32_t* tile_cost;
    float4* rgba;              // may be null
    float* depth;              // may be null
    float* gbuf;               // may be null
    unsigned long long* keys;  // may be null (sort-last compositing keys)
    uint32_t* rgba8;           // may be null: the frame as an RGBA8 framebuffer would hold it (round(clamp(c) * 255))
};

// What the round kernel of the tracer (trace.cu: trace_rounds_kernel) needs beyond TraceParams.  A frame is traced
// in `world` rounds.  Round 0 starts every ray whose start position this rank owns; a ray marches while the lower z
// tap of its texture fetch lies in this rank's own slices [own_z0, own_z1) (the upper tap is then an own or a halo
// slice), and is handed to the neighbour -- position, t and step count, 24 bytes -- when it leaves them; rounds
// k >= 1 continue the rays the neighbours handed over in round k - 1.  The step sequence of every ray is the one a
// single GPU holding the whole grid runs, bit for bit.  With linked == 0 the kernel is the plain single-volume trace.
// Hand-over is PUSH: a rank stores the entries into its neighbour's in-queue and publishes the count there, so every
// kernel reads its work from local memory.
struct LinkParams {
    uint32_t linked;        // 0: single volume, outputs of TraceParams; 1: sharded, finished pixels go to the presenter
    uint32_t first;         // round 0 of a frame: the work units are 8 x 4 pixel tiles; else 32-entry runs of the in-queues
    uint32_t is_presenter;  // writes the pixels whose ray never enters the box
    uint32_t own_z0, own_z1;
    uint32_t max_pixels;    // capacity of every queue buffer
    uint32_t* work_head;    // local: next work unit (reset by the last CTA)
    uint32_t* ctas_done;    // local: CTAs that have finished (reset by the last CTA)
    // in-queues (OWN memory: the neighbours PUSH): what they appended in the previous round, [0] the neighbour below
    // (rays travelling up), [1] the neighbour above (rays travelling down); counts published by their last CTA
    const uint32_t* in_count[2];
    const float4* in_pos[2];
    const uint2* in_id[2];
    // out-queues of this round: [0] rays leaving downwards, [1] upwards -- each in the in-queue of that neighbour
    // (peer memory: stores over NVLink, no load ever crosses it); null without a neighbour
    uint32_t* out_count;      // own memory: two reservation counters
    float4* out_pos[2];       // position.xyz, t
    uint2* out_id[2];         // pixel index, step count
    uint32_t* out_publish[2]; // the neighbours' in_count words for this round (written by the last CTA)
    uint32_t* reset_count;    // the counter pair this kernel's last CTA zeroes (ring of 4: the pair of round + 2)
    unsigned long long* frame_keys;  // presenter's (depth, RGBA8) key frame -- peer memory on the other ranks
    float* frame_gbuf;               // presenter's G-buffer frame, or null
    uint32_t* sig_round[2];          // the neighbours' "round done" flags (peer memory), or null
    uint32_t sig_round_value;
    uint32_t* sig_frame;             // the presenter's "rank r finished the frame" flag (last round only), or null
    uint32_t sig_frame_value;
    // ---- stream mode (trace.cu: trace_stream_kernel): ONE launch per rank and frame instead of `world` rounds.  The
    // warps first work through the tiles, then poll the in-queues, so a handed-over ray continues as soon as it has
    // arrived instead of at the next round.  A queue entry is two float4 -- (x, y, z, tag), (t, pixel, steps, tag) --
    // and is complete when both tags equal `epoch`: no fence, no count between producer and consumer.  A rank closes
    // an out-queue with its final count once nothing can enter it any more: its tiles are done and the in-queue on
    // the other side (rays keep their z direction) is closed and worked off.
    uint32_t epoch;                         // tag of this frame's entries (frame number + 1: never 0, the memset value)
    uint32_t timeout_ms;                    // a rank that waits longer for its neighbours gives up (sets *timed_out)
    uint32_t* tiles_done;                   // local counters, reset by the last CTA: tile units finished,
    uint32_t* in_head;                      //   [2] in-queue units claimed,
    uint32_t* in_done;                      //   [2] in-queue units finished,
    uint32_t* sent_final;                   //   [2] out-queue closed
    uint32_t* timed_out;                    // mapped host memory
    const unsigned long long* in_final[2];  // own memory, written by the neighbour: epoch << 32 | entries of its out-queue
    unsigned long long* out_final[2];       // the same word of the neighbours (peer memory)
    const float4* in_q[2];                  // entries, own memory; null without a neighbour on that side
    float4* out_q[2];                       // peer memory
    // presenter only: the tiles outside the screen rectangle of the box (no ray: the presenter alone writes them) come
    // FIRST and go straight into the RGBA8 / depth frame instead of the key frame; the last of their units raises
    // outside_flag, on which the presenter's stream waits to copy the rows that hold nothing else to the host while
    // the frame is still being traced.  Null: the tiles outside come last and go into the key frame like every pixel.
    uint32_t* out_rgba8;
    float* out_depth;
    uint32_t* outside_done;                 // local counter, reset by the last CTA
    uint32_t* outside_flag;                 // local: epoch of the last frame whose outside tiles are all written
};

// launchers (fill.cu / trace.cu).  `program`: dev::PROG_INTERPRET or dev::PROG_DEMO (built in)
cudaError_t launch_fill(const FillParams& p, int voxels_per_thread, int program, int grid_ctas, size_t smem_bytes,
                        cudaStream_t s);
size_t fill_smem_bytes(uint32_t tape_img_bytes, uint32_t n_cull, uint32_t max_stack, int voxels_per_thread,
                       uint32_t* stack_floats);
int fill_max_ctas_per_sm(int voxels_per_thread, int program, size_t smem_bytes);
cudaError_t fill_prepare(int voxels_per_thread, int program, size_t smem_bytes);  // opt in to > 48 KB dynamic smem

// jit.cu: straight-line kernels specialised for a tape structure (NVRTC + driver API, both dlopen'ed)
}  // namespace sdfgpu
#include <string>
#include <vector>
namespace sdfgpu {
std::string jit_source(const std::vector<uint32_t>& opcodes, int voxels_per_thread);
bool jit_compile(const std::vector<uint32_t>& opcodes, int voxels_per_thread, int cc_major, int cc_minor,
                 std::vector<char>* cubin, std::string* err);
bool jit_available(std::string* why);
bool jit_get(int device, int cc_major, int cc_minor, const std::vector<uint32_t>& opcodes, int voxels_per_thread,
             size_t smem_bytes, void** fn_out, int* max_ctas_per_sm, std::string* err);
bool jit_launch(void* fn, const FillParams& p, int grid, size_t smem, cudaStream_t s, std::string* err);

cudaError_t launch_cull_cells(const unsigned char* img_dev, const uint32_t dims[3], uint32_t cells_x, uint32_t cells_y,
                              uint32_t cells_z, uint32_t* lists, uint32_t* counts, cudaStream_t s);
cudaError_t launch_ingest(float4* tex0, float4* tex1, const float* samples_dev, size_t first, size_t n,
                          const float* lut_dev, float air_dist, int grid, cudaStream_t s);
cudaError_t launch_ingest_scatter(float4* tex0, float4* tex1, const float* samples_dev, const uint32_t* idx_dev,
                                  size_t n, const float* lut_dev, float air_dist, int grid, cudaStream_t s);
cudaError_t launch_gather_dist(const float4* tex0, const uint32_t* idx_dev, size_t n, float* out_dev, int grid,
                               cudaStream_t s);
cudaError_t launch_set_const(float4* dst, size_t n_texels, float v, int grid_ctas, cudaStream_t s);
cudaError_t launch_trace(const TraceParams& p, int variant, cudaStream_t s);
cudaError_t launch_tile_order(const TraceParams& p, const uint32_t* prev_cost, uint32_t* order, cudaStream_t s);
cudaError_t launch_trace_rounds(const TraceParams& p, const LinkParams& l, int grid_ctas, cudaStream_t s);
int trace_rounds_max_ctas_per_sm(const TraceParams& p);
cudaError_t launch_trace_stream(const TraceParams& p, const LinkParams& l, int grid_ctas, cudaStream_t s);
int trace_stream_max_ctas_per_sm(const TraceParams& p);
cudaError_t launch_signal(uint32_t* const* flags, const uint32_t* values, int n, cudaStream_t s);  // fence.sys + stores (peer flags)
cudaError_t launch_spin_wait(const uint32_t* flag, uint32_t value, uint32_t* timed_out, cudaStream_t s);  // fallback for cuStreamWaitValue32
cudaError_t launch_extract_dist(const float4* tex0, float* dist, size_t n, int grid, cudaStream_t s);
cudaError_t launch_extract_dist_array(const float4* tex0, unsigned long long surf, uint32_t W, uint32_t H,
                                      uint32_t stored_slices, int grid, cudaStream_t s);
cudaError_t launch_keys_unpack(const unsigned long long* keys, uint32_t n, uint8_t* rgba8, float* depth,
                               cudaStream_t s);

}  // namespace sdfgpu

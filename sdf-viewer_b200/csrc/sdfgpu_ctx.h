// sdfgpu_ctx.h -- the handle behind the C ABI (include/sdfgpu.h) and the host-side helpers shared by api.cu and
// link.cu.  Internal: not installed.  Paths in comments are relative to /root/reference.
#pragma once

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sdfgpu_internal.h"

namespace sdfgpu {

// src/app/scene/sdf/loading.rs:108-115
inline uint32_t prev_power_of_2(uint32_t x) {
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return x - (x >> 1);
}

struct LoadingState {  // LoadingManager at pass granularity, loading.rs:5-19
    uint64_t limits[3] = {0, 0, 0};
    uint64_t passes = 0;
    uint64_t step_size = 0;
    uint64_t next[3] = {0, 0, 0};  // next_index: only the host-sampled path stops inside a pass
    uint64_t iterations = 0;       // iterations done in the current pass (0 between passes)
    uint64_t total_iterations = 0;

    void reset(uint64_t p) {  // :37-43
        passes = p;
        const uint32_t e = (uint32_t)(p > 1 ? p : 1) - 1;
        step_size = e < 63 ? (uint64_t)1 << e : (uint64_t)1 << 62;
        next[0] = next[1] = next[2] = 0;
        iterations = 0;
        total_iterations = 0;
    }
    uint64_t pass_items(uint64_t s) const {
        return ((limits[0] + s - 1) / s) * ((limits[1] + s - 1) / s) * ((limits[2] + s - 1) / s);
    }
    uint64_t len() const {  // :80-89
        uint64_t s = step_size, it = 0;
        while (s > 0) {
            it += pass_items(s);
            s = prev_power_of_2((uint32_t)(s - 1));
        }
        return it - iterations;
    }
    uint32_t passes_left() const {  // :99-105
        if (step_size == 0) return 0;
        return (uint32_t)log2f((float)step_size) + 1;
    }
    // Up to `max_iters` consecutive iterations of next() (:50-76) that lie in one x row: they visit
    // (x0 + i * step, y, z) for i < take.  Returns take (0 when loading is done) and advances the
    // cursor and the counters exactly as `take` calls of next() would; *pass_end is set when the last
    // of them ended the pass (step_size is then already the next pass's).
    uint64_t next_row(uint64_t max_iters, uint64_t* x0, uint64_t* y, uint64_t* z, uint64_t* step, bool* pass_end) {
        *pass_end = false;
        if (step_size == 0 || max_iters == 0) return 0;
        const uint64_t s = step_size;
        *x0 = next[0]; *y = next[1]; *z = next[2]; *step = s;
        // next() yields next_index even when it lies outside the limits (an empty axis): one iteration
        const uint64_t row_left = next[0] < limits[0] ? (limits[0] - next[0] + s - 1) / s : 1;
        const uint64_t take = row_left < max_iters ? row_left : max_iters;
        iterations += take;
        total_iterations += take;
        next[0] += take * s;
        if (next[0] >= limits[0]) {
            next[0] = 0;
            next[1] += s;
            if (next[1] >= limits[1]) {
                next[1] = 0;
                next[2] += s;
                if (next[2] >= limits[2]) {
                    step_size = prev_power_of_2((uint32_t)(s - 1));
                    next[0] = next[1] = next[2] = 0;
                    iterations = 0;
                    *pass_end = true;
                }
            }
        }
        return take;
    }
    void finish_pass() {  // the rest of the current pass at once, then the tail of next(), :67-71
        total_iterations += pass_items(step_size) - iterations;
        step_size = prev_power_of_2((uint32_t)(step_size - 1));
        next[0] = next[1] = next[2] = 0;
        iterations = 0;
    }
};


constexpr uint32_t LINK_MAX_WORLD = 16;

// Multi-GPU link of a slab handle (link.cu): neighbours' volumes and every rank's arena mapped into this
// process (CUDA IPC, or plain peer pointers when the other handle lives in this process), the epochs of the
// flag protocol, and what the round kernels of the sharded trace need.
struct LinkState {
    bool on = false;
    uint32_t rank = 0, world = 1;
    uint32_t max_pixels = 0;
    bool want_gbuf = false;
    unsigned char* arena = nullptr;  // own arena (device memory, exported)
    size_t arena_bytes = 0;
    unsigned char* peer_arena[LINK_MAX_WORLD] = {};  // every rank's arena as seen from this device ([rank] = own)
    bool peer_arena_ipc[LINK_MAX_WORLD] = {};        // opened with cudaIpcOpenMemHandle (closed on detach)
    bool peer_tex_ipc[2] = {false, false};
    int nb[2] = {-1, -1};            // neighbour ranks below / above (-1: none)
    uint32_t fill_epoch = 0;         // fills (of any kind) this handle has signalled
    uint32_t frame_epoch = 0;        // linked traces started
    uint32_t round_epoch = 0;        // global trace rounds started (world rounds per frame)
    bool memops = false;             // cuStreamWaitValue32 usable: waits are stream memory operations, else a spin kernel
    // halo slices: false (default) -- every rank fills its two halo slices itself (sample() is a pure function of the
    // position, src/sdf/mod.rs:43, so they equal the neighbours' boundary slices bit for bit and the fill needs no
    // exchange at all); true (SDFGPU_LINK_HALO_PUSH) -- the neighbours push them over NVLink
    bool halo_push = false;
    // trace: true -- one streaming kernel per rank and frame (needs every rank on a device of its own: the kernels
    // wait for each other); false -- `world` rounds of the round kernel
    bool stream = false;
    // presenter (rank 0): what follows its trace kernel -- waiting for the other ranks' pixels, unpacking the key
    // frame, the copies to the host, telling the ranks -- runs on a stream of its own, so that the handle's stream goes
    // straight on to the next fill
    cudaStream_t present_stream = nullptr;
    cudaEvent_t ev_traced = nullptr, ev_presented = nullptr;
    uint32_t* timed_out_host = nullptr;  // mapped host word the stream kernel sets when it gives up waiting
    uint32_t timeout_ms = 8000;
    // frame in flight (begin / round / end are separate so that a single-process group can interleave ranks)
    uint32_t cur_w = 0, cur_h = 0, cur_round = 0;
    bool cur_gbuf = false;
    bool cur_outside_first = false;  // this frame's presenter kernel writes the tiles outside the rectangle first, into the frame
    bool in_frame = false;
    // SDFGPU_LINK_TIMING=1 (development): events around the waits and the kernel of every round, printed to stderr
    // by a synchronising sdfgpu_trace_linked
    std::vector<cudaEvent_t> timing_events;
    size_t timing_used = 0;
};

// The mesher's device buffers (mesh.cu); `points` / `samples` also stage sdfgpu_sample_points.
struct MeshState {
    uint32_t* vert_base = nullptr;      size_t vert_base_cap = 0;      // per lattice point
    uint32_t* block_counts = nullptr;   size_t block_counts_cap = 0;
    uint32_t* block_offsets = nullptr;  size_t block_offsets_cap = 0;
    float* vertices = nullptr;          size_t vertices_cap = 0;       // 12 floats per vertex (mesh.rs Vertex)
    uint32_t* indices = nullptr;        size_t indices_cap = 0;        // 3 per triangle
    float* points = nullptr;            size_t points_cap = 0;
    float* samples = nullptr;           size_t samples_cap = 0;
    uint64_t n_vertices = 0, n_triangles = 0;
    bool valid = false;
};

}  // namespace sdfgpu

struct sdfgpu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    float bb[6];
    uint32_t dims[3];
    uint32_t z_begin = 0, z_end = 0, z_lo = 0, z_hi = 0;
    float4* tex0 = nullptr;
    float4* tex1 = nullptr;
    size_t stored_texels = 0;
    sdfgpu::LoadingState lm;
    bool has_changed_box = false;
    float changed_box[6];
    bool changed_box_while_loading = false;
    // What is known about the volume without reading it: 0 = every stored voxel holds AIR_DIST (fresh /
    // reset); s > 0 = exactly the lattice points of step s (power of two) have been sampled, the rest hold
    // AIR_DIST; 1 = every voxel has been sampled; -1 = unknown (the conditional passes read tex0.r).
    int64_t known_step = 0;
    float lod = 1.0f;            // SDFViewerMaterial::lod_dist_between_samples, material.rs:27
    bool filter_linear = false;  // GL filter state: NEAREST until a commit at lod == 1 (mod.rs:110-111,227-238)
    // tape
    bool has_tape = false;
    std::vector<unsigned char> img_host;
    unsigned char* img_dev = nullptr;
    size_t img_dev_cap = 0;
    // coarse pre-cull of the current tape's culled UNION_RANGE (FillParams::cell_lists), rebuilt by every set_tape
    uint32_t* cell_lists = nullptr;   // cells x cull_count
    uint32_t* cell_counts = nullptr;  // follows the lists in the same allocation
    size_t cell_lists_cap = 0;        // in u32
    uint32_t cells[3] = {0, 0, 0};
    bool cell_lists_valid = false;
    int opt_cull_cells = 1;           // 0: every tile culls the whole range (measurement)
    sdfgpu::TapeImageHeader hdr;
    std::vector<uint32_t> opcodes;  // lowered opcode sequence (+ scalar programs' text) = the tape's structure (JIT cache key)
    uint32_t n_top_ops = 0;         // opcodes up to and including DOP_END
    bool has_scalar = false;        // the tape runs scalar programs: only the specialised (NVRTC) kernel evaluates them
    bool structure_is_demo = false; // matches the built-in PROG_DEMO kernel
    std::vector<float> px, py, pz;  // host copies of the position tables
    float lut[256];
    // frame
    uint32_t fw = 0, fh = 0;
    float4* rgba_dev = nullptr;
    float* depth_dev = nullptr;
    float* gbuf_dev = nullptr;
    unsigned long long* keys_dev = nullptr;
    uint32_t* rgba8_dev = nullptr;
    float* ingest_dev = nullptr;  // staging for sdfgpu_ingest_samples: records, then the LUT
    size_t ingest_cap = 0;
    // sdfgpu_update_surface, host-sampled path: two pinned staging buffers (records + texel indices) with
    // their device copies, so that sampling chunk i+1 overlaps the transfer and scatter of chunk i
    struct HostStage {
        float* rec = nullptr; uint32_t* idx = nullptr;          // pinned host
        float* rec_dev = nullptr; uint32_t* idx_dev = nullptr;  // device
        size_t cap = 0;                                         // voxels
        cudaEvent_t done = nullptr;
        bool in_flight = false;
    } stage[2];
    uint64_t stage_turn = 0;
    float* gather_host = nullptr;  // pinned: tex0.r of a chunk's candidates when the state is unknown
    float* gather_dev = nullptr;
    size_t gather_cap = 0;
    int64_t pass_known = 0;        // known_step when the current (partially walked) host pass began
    double host_rate = 1.0e6;      // LoadingManager iterations per second the host-sampled path last ran at
    bool host_rate_known = false;
    std::vector<unsigned char> tape_bytes;  // the public tape last given to sdfgpu_set_tape (change detection)
    float* lut_dev = nullptr;
    float* dist_dev = nullptr;  // optional distance-only volume for the tracer (option trace_distance_volume = 1)
    cudaArray_t dist_arr = nullptr;  // the same as an R32F 3-D CUDA array (options 2, 3): written through dist_surf,
    cudaSurfaceObject_t dist_surf = 0;  // read through dist_tex[0] (point filter) or dist_tex[1] (linear filter)
    cudaTextureObject_t dist_tex[2] = {0, 0};
    bool dist_valid = false;    // the distance volume the current option uses mirrors tex0.r
    // GL textures of the presenter registered for interop (sdfgpu_gl_register): RGBA8 colour, optional R32F depth
    cudaGraphicsResource* gl_res[2] = {nullptr, nullptr};
    uint32_t gl_w = 0, gl_h = 0;
    float* dist_full = nullptr;         // exact multi-GPU trace: tex0.r of the WHOLE grid (W*H*D floats), replicated
    bool dist_full_own_valid = false;   //   this handle's own slices of it mirror tex0.r
    bool peers_ever = false;  // a neighbour may hold an IPC mapping of this handle's volumes
    int opt_dist_volume = 0;
    unsigned long long* touched_dev = nullptr;
    // neighbours' volumes opened with cudaIpcOpenMemHandle (fused halo exchange)
    float4* peer_tex0[2] = {nullptr, nullptr};
    float4* peer_tex1[2] = {nullptr, nullptr};
    uint32_t peer_z_lo[2] = {0, 0};
    unsigned long long* cull_stats_dev = nullptr;  // set while sdfgpu_cull_stats runs its fill
    cudaStream_t halo_stream = nullptr;  // DMA pushes of the boundary slices, overlapped with the interior fill
    cudaStream_t copy_stream = nullptr;  // frame rows to the host while the next band of the frame is traced
    cudaStream_t copy_stream2 = nullptr; //   ... colour on the first, depth on the second
    // per 8x8-pixel tile: the longest march of the frame being traced / of the previous frame; the tile order made of it
    uint32_t* tile_cost[2] = {nullptr, nullptr};
    uint32_t* tile_order = nullptr;
    size_t tile_cap = 0;
    int cost_cur = 0;
    bool cost_valid = false;
    uint32_t cost_w = 0, cost_h = 0;
    int opt_tile_order = 1;              // 0: the tiles inside the rectangle row by row; 1: auto; 2: always by cost
    uint32_t* band_counters = nullptr;   // 64 per-band CTA counters + 64 per-band flags (trace_tiles_kernel)
    uint32_t band_epoch = 0;
    int opt_trace_bands = 6;             // sdfgpu_trace_rgba8: bands per frame (1: trace the frame, then copy it)
    cudaEvent_t ev_boundary = nullptr, ev_pushed = nullptr;
    // options
    int opt_vpt = 0;        // voxels per thread (0 = default)
    int opt_ctas = 0;       // CTAs per SM (0 = as many as fit)
    int opt_fill_halo = 1;  // compute the halo slices locally (0: the host exchanges them)
    int opt_max_steps = 256;    // maxSteps of sdfRaycast (material.frag:142); other values are for experiments only
    int opt_trace_variant = 0;  // 0: 8x8 tiles, heavy-first 1-D grid; 1: plain 2-D grid of 32x8 CTAs
    int opt_program = 0;    // 0 auto (JIT, else built-in, else interpreter), 1 interpreter, 2 built-in, 3 JIT or fail
    int cc_major = 0, cc_minor = 0;
    int last_program = -1, last_ctas = 0, last_vpt = 0;
    std::string jit_note;   // why the JIT was not used (if it was not)
    size_t smem_prepared[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // [V][interpreter | demo]
    uint64_t launches = 0;
    std::string err;
    sdfgpu::LinkState link;
    sdfgpu::MeshState mesh;
    sdfgpu::TraceParams link_tp;  // trace parameters of the linked frame in flight
    bool fill_boundary_first = false;  // the next run_fill is the whole-slab launch of a linked fill_all
    int opt_link_wait = 0;             // 0: stream memory operations when the driver has them, 1: spin-wait kernels
    int opt_link_halo_push = 0;        // before sdfgpu_link_export: 1 = SDFGPU_LINK_HALO_PUSH
    int opt_link_trace_mode = 0;       // before sdfgpu_link_export: 0 auto, 1 rounds (SDFGPU_LINK_ROUNDS), 2 stream
    int opt_link_timeout_ms = 8000;
    uint32_t* trace_counters = nullptr;  // work_head, ctas_done of the round kernel (un-linked handles)
};


namespace sdfgpu {

int fail(sdfgpu_ctx* ctx, int code, const char* fmt, ...);

#define CK(ctx, call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            (void)cudaGetLastError();                                                                  \
            return sdfgpu::fail((ctx), SDFGPU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
        }                                                                                              \
    } while (0)

// api.cu
float air_dist_value();
void set_device(sdfgpu_ctx* ctx);
bool has_peers(const sdfgpu_ctx* ctx);
int push_halos(sdfgpu_ctx* ctx, cudaStream_t s, bool lo = true, bool hi = true);
int ensure_frame(sdfgpu_ctx* ctx, uint32_t w, uint32_t h, bool want_gbuf, bool want_keys);
int fill_trace_params(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t w, uint32_t h, bool slab_clip, TraceParams* tp);
int dispatch_fill(sdfgpu_ctx* ctx, FillParams& p, int V);  // program choice + launch of a prepared fill

// mesh.cu
int sample_points_device(sdfgpu_ctx* ctx, const float* points_dev, uint32_t n, float* out_dev);  // the tape at n positions
void mesh_free(sdfgpu_ctx* ctx);

// link.cu
int link_after_fill(sdfgpu_ctx* ctx, bool touched_lo, bool touched_hi);  // push the boundary slices that changed, signal the neighbours
int link_fill_all_fused(sdfgpu_ctx* ctx, FillParams* p);                  // fills in the boundary-first fields of a full-slab launch
int link_fill_all_pushed(sdfgpu_ctx* ctx);                                // after that launch: flag-ordered DMA push + signal
int link_trace_begin(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t w, uint32_t h, bool want_gbuf);
bool stream_wait_value_available();  // cuStreamWaitValue32 resolved
bool stream_wait_value(cudaStream_t s, const uint32_t* flag, uint32_t value);  // stream continues once *flag >= value
int link_trace_round(sdfgpu_ctx* ctx);
int link_trace_stream(sdfgpu_ctx* ctx);  // LinkState::stream: the whole frame of this rank in one launch
int link_trace_issue(sdfgpu_ctx* ctx);   // every round, or the one streaming launch
int link_trace_end(sdfgpu_ctx* ctx, uint8_t* rgba8, float* depth, float* gbuf, bool sync);
void link_free(sdfgpu_ctx* ctx);

}  // namespace sdfgpu

// api.cu -- host side of libsdfgpu.so: the C ABI of include/sdfgpu.h.
//
// Holds what `SDFViewer` holds in the reference
// (/root/reference/src/app/scene/sdf/mod.rs:21-38): the two volumes (here in
// HBM only), the LoadingManager counters (src/app/scene/sdf/loading.rs:5-19),
// the pending changed box and the lod latched by commit().  A LoadingManager
// pass is one launch of the fill kernel (fill.cu); the tracer is trace.cu.
// There is no CPU fallback: every compute entry point fails with
// SDFGPU_ERR_CUDA when no device is usable.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "sdfgpu_ctx.h"

using namespace sdfgpu;

// CUDA <-> OpenGL interop lives in cudart; its header wants <GL/gl.h>, which this image does not have
// (GLuint and GLenum are unsigned int)
extern "C" cudaError_t cudaGraphicsGLRegisterImage(cudaGraphicsResource** resource, unsigned int image, unsigned int target,
                                                   unsigned int flags);

#define SDFGPU_API extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_thread_error;

}  // namespace

int sdfgpu::fail(sdfgpu_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_thread_error = buf;
    return code;
}

float sdfgpu::air_dist_value() {  // src/app/scene/sdf/mod.rs:42
    volatile float a = 1e-1f, b = 0.001234f;
    return a + b;
}

void sdfgpu::set_device(sdfgpu_ctx* ctx) { (void)cudaSetDevice(ctx->device); }

namespace {

// three-d-asset Srgba::to_linear_srgb on one u8 channel (call site scene/sdf/mod.rs:201)
float srgb_u8_to_linear(unsigned v) {
    volatile float c = (float)v / 255.0f;
    if (c < 0.04045f) return c / 12.92f;
    volatile float t = (c + 0.055f) / 1.055f;
    return powf(t, 2.4f);
}

void build_pos_table(std::vector<float>& t, uint32_t n, float lo, float hi) {  // scene/sdf/mod.rs:160-161,179-182
    t.resize(n);
    volatile float size = hi - lo;
    volatile float nm1 = (float)n - 1.0f;
    for (uint32_t i = 0; i < n; ++i) {
        volatile float p = (float)i;
        p = p / nm1;
        p = p * size;
        p = p + lo;
        t[i] = p;
    }
}

uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// voxels per thread: 8 for the straight-line kernels (specialised / built in), 4 for the interpreter
// (measured on B200, tools/sweep_minb.py)
int default_vpt(const sdfgpu_ctx* ctx) {
    if (ctx->opt_vpt) return ctx->opt_vpt;
    if (ctx->opt_program == 1) return 4;
    if (ctx->has_scalar) return 4;  // long straight-line programs: fewer voxels per thread keep the registers in check
    if (ctx->opt_program != 2 && ctx->n_top_ops <= 96 && jit_available(nullptr)) return 8;
    return ctx->structure_is_demo ? 8 : 4;
}

}  // namespace

// picks the program for a prepared launch (a kernel specialised for this tape structure (NVRTC), the built-in demo
// program, or the interpreter) and launches it
int sdfgpu::dispatch_fill(sdfgpu_ctx* ctx, FillParams& p, int V) {
    const uint32_t n_cull = (ctx->hdr.flags & TAPE_FLAG_CULL) ? ctx->hdr.cull_count : 0u;
    const size_t smem = fill_smem_bytes(p.tape_img_bytes, n_cull, ctx->hdr.max_stack, V, &p.stack_floats);
    if (smem > 227u * 1024u)
        return fail(ctx, SDFGPU_ERR_TAPE, "tape needs %zu bytes of shared memory per CTA (limit 232448)", smem);
    // ---- pick the program: a kernel specialised for this tape structure (NVRTC), the built-in
    // demo program, or the interpreter
    const uint64_t n_tiles = (uint64_t)p.tiles_x * p.tiles_y * p.tiles_z;
    if (n_tiles > 0xffffffffull) return fail(ctx, SDFGPU_ERR_INVALID, "grid too large for one launch");
    void* jit_fn = nullptr;
    int per_sm = 0, program = dev::PROG_INTERPRET;
    if (ctx->opt_program == 0 || ctx->opt_program == 3) {
        std::string why;
        if (ctx->n_top_ops > 96) why = "tape longer than 96 instructions";
        else if (jit_get(ctx->device, ctx->cc_major, ctx->cc_minor, ctx->opcodes, V, smem, &jit_fn, &per_sm, &why))
            program = dev::PROG_JIT;
        if (program != dev::PROG_JIT) {
            ctx->jit_note = why;
            if (ctx->opt_program == 3) return fail(ctx, SDFGPU_ERR_CUDA, "JIT kernel unavailable: %s", why.c_str());
        }
    }
    if (program != dev::PROG_JIT && ctx->has_scalar)
        return fail(ctx, SDFGPU_ERR_CUDA, "the tape runs scalar programs, which only the specialised kernel evaluates: %s",
                    ctx->opt_program == 1 || ctx->opt_program == 2 ? "fill_program forces the interpreter" : ctx->jit_note.c_str());
    if (program != dev::PROG_JIT) {
        if ((ctx->opt_program == 0 || ctx->opt_program == 2) && ctx->structure_is_demo) program = dev::PROG_DEMO;
        const size_t prepared_key = smem;
        if (prepared_key > ctx->smem_prepared[V == 1 ? 0 : V == 2 ? 1 : V == 4 ? 2 : 3][program == dev::PROG_DEMO]) {
            CK(ctx, fill_prepare(V, program, smem));
            ctx->smem_prepared[V == 1 ? 0 : V == 2 ? 1 : V == 4 ? 2 : 3][program == dev::PROG_DEMO] = smem;
        }
        per_sm = fill_max_ctas_per_sm(V, program, smem);
        if (per_sm < 1) return fail(ctx, SDFGPU_ERR_CUDA, "fill kernel does not fit on an SM (smem %zu)", smem);
    }
    if (ctx->opt_ctas > 0 && ctx->opt_ctas < per_sm) per_sm = ctx->opt_ctas;
    uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (program == dev::PROG_JIT) {
        std::string why;
        if (!jit_launch(jit_fn, p, (int)grid, smem, ctx->stream, &why)) return fail(ctx, SDFGPU_ERR_CUDA, "%s", why.c_str());
    } else {
        CK(ctx, launch_fill(p, V, program, (int)grid, smem, ctx->stream));
    }
    ctx->last_program = program; ctx->last_ctas = per_sm; ctx->last_vpt = V;
    ctx->launches++;
    return SDFGPU_OK;
}

namespace {

// one fill launch over lattice {r0 + i*step} restricted to the index box [lo, hi) per axis
int run_fill(sdfgpu_ctx* ctx, uint32_t step, const uint32_t lo[3], const uint32_t hi[3], uint32_t conditional,
             unsigned long long* touched, uint32_t known_step = 0) {
    if (!ctx->has_tape) return fail(ctx, SDFGPU_ERR_STATE, "no tape set (call sdfgpu_set_tape first)");
    FillParams p;
    memset(&p, 0, sizeof p);
    uint32_t r0[3], n[3];
    for (int a = 0; a < 3; ++a) {
        if (hi[a] <= lo[a]) return SDFGPU_OK;
        r0[a] = ((lo[a] + step - 1) / step) * step;
        if (r0[a] >= hi[a]) return SDFGPU_OK;
        n[a] = (hi[a] - r0[a] + step - 1) / step;
    }
    int V = default_vpt(ctx);
    if (!ctx->opt_vpt)  // thin launches (a boundary slice, a dirty box): do not pad the z extent of a tile with idle voxels
        while (V > 1 && (uint32_t)V > n[2]) V >>= 1;
    p.tex0 = ctx->tex0; p.tex1 = ctx->tex1;
    p.tape_img = ctx->img_dev; p.tape_img_bytes = (uint32_t)ctx->img_host.size();
    p.W = ctx->dims[0]; p.H = ctx->dims[1]; p.D = ctx->dims[2];
    p.z_lo = ctx->z_lo;
    p.rx0 = r0[0]; p.ry0 = r0[1]; p.rz0 = r0[2];
    p.nx = n[0]; p.ny = n[1]; p.nz = n[2];
    p.step = step;
    p.tiles_x = (n[0] + FILL_TILE_X - 1) / FILL_TILE_X;
    p.tiles_y = (n[1] + FILL_TILE_Y - 1) / FILL_TILE_Y;
    p.tiles_z = (n[2] + V - 1) / V;
    p.conditional = conditional;
    p.known_step = known_step;
    p.has_box = ctx->has_changed_box ? 1u : 0u;
    if (ctx->has_changed_box) memcpy(p.box, ctx->changed_box, sizeof p.box);
    p.air_dist = air_dist_value();
    p.touched = touched;
    p.cull_stats = ctx->cull_stats_dev;  // null unless sdfgpu_cull_stats is running
    if (ctx->cell_lists_valid && ctx->opt_cull_cells) {
        p.cell_lists = ctx->cell_lists; p.cell_counts = ctx->cell_counts;
        p.cells_x = ctx->cells[0]; p.cells_y = ctx->cells[1];
    }
    if (ctx->fill_boundary_first) (void)link_fill_all_fused(ctx, &p);
    const int rc = dispatch_fill(ctx, p, V);
    if (rc != SDFGPU_OK) return rc;
    ctx->dist_valid = false; ctx->dist_full_own_valid = false;
    return SDFGPU_OK;
}

}  // namespace

bool sdfgpu::has_peers(const sdfgpu_ctx* ctx) { return ctx->peer_tex0[0] || ctx->peer_tex0[1]; }

// Copy this handle's first / last owned slice (both volumes) into the neighbours' halo slices:
// device-to-device copies into the mapped peer volumes, i.e. NVLink DMA by the copy engines.
int sdfgpu::push_halos(sdfgpu_ctx* ctx, cudaStream_t s, bool lo, bool hi) {
    const size_t slice = (size_t)ctx->dims[0] * ctx->dims[1];
    for (int side = 0; side < 2; ++side) {
        if (!ctx->peer_tex0[side] || !(side == 0 ? lo : hi)) continue;
        const uint32_t z = side == 0 ? ctx->z_begin : ctx->z_end - 1;
        const size_t src = (size_t)(z - ctx->z_lo) * slice, dst = (size_t)(z - ctx->peer_z_lo[side]) * slice;
        CK(ctx, cudaMemcpyAsync(ctx->peer_tex0[side] + dst, ctx->tex0 + src, slice * sizeof(float4), cudaMemcpyDeviceToDevice, s));
        CK(ctx, cudaMemcpyAsync(ctx->peer_tex1[side] + dst, ctx->tex1 + src, slice * sizeof(float4), cudaMemcpyDeviceToDevice, s));
    }
    return SDFGPU_OK;
}

namespace {

void fill_z_range(const sdfgpu_ctx* ctx, uint32_t* za, uint32_t* zb) {
    if (ctx->opt_fill_halo) { *za = ctx->z_lo; *zb = ctx->z_hi; }
    else { *za = ctx->z_begin; *zb = ctx->z_end; }
}

int alloc_volumes(sdfgpu_ctx* ctx) {
    ctx->stored_texels = (size_t)ctx->dims[0] * ctx->dims[1] * (ctx->z_hi - ctx->z_lo);
    if (ctx->stored_texels) {
        CK(ctx, cudaMalloc(&ctx->tex0, ctx->stored_texels * sizeof(float4)));
        CK(ctx, cudaMalloc(&ctx->tex1, ctx->stored_texels * sizeof(float4)));
    }
    CK(ctx, cudaMalloc(&ctx->touched_dev, sizeof(unsigned long long)));
    CK(ctx, cudaMalloc(&ctx->trace_counters, 2 * sizeof(uint32_t)));
    CK(ctx, cudaMemsetAsync(ctx->trace_counters, 0, 2 * sizeof(uint32_t), ctx->stream));
    return SDFGPU_OK;
}

int reset_volumes(sdfgpu_ctx* ctx) {  // new_voxels, scene/sdf/mod.rs:76-77: AIR_DIST in all 4 channels of both
    ctx->dist_valid = false; ctx->dist_full_own_valid = false;
    const int grid = ctx->sm_count * 8;
    CK(ctx, launch_set_const(ctx->tex0, ctx->stored_texels, air_dist_value(), grid, ctx->stream));
    CK(ctx, launch_set_const(ctx->tex1, ctx->stored_texels, air_dist_value(), grid, ctx->stream));
    if (ctx->stored_texels) ctx->launches += 2;
    return SDFGPU_OK;
}

int create_common(const float bb[6], const uint32_t voxels[3], uint32_t passes, int device, uint32_t z_begin,
                  uint32_t z_end, sdfgpu_ctx** out) {
    if (!out) return fail(nullptr, SDFGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!bb || !voxels) return fail(nullptr, SDFGPU_ERR_INVALID, "bb / voxels is NULL");
    if (z_begin > z_end || z_end > voxels[2]) return fail(nullptr, SDFGPU_ERR_INVALID, "bad slab range");
    if (voxels[0] > 65535u || voxels[1] > 65535u || voxels[2] > 65535u)
        return fail(nullptr, SDFGPU_ERR_INVALID, "more than 65535 voxels on a side");
    {
        const uint32_t zl = z_begin > 0 ? z_begin - 1 : 0, zh = z_end < voxels[2] ? z_end + 1 : voxels[2];
        if ((uint64_t)voxels[0] * voxels[1] * (zh - zl) >= (1ull << 31))
            return fail(nullptr, SDFGPU_ERR_INVALID, "a handle stores at most 2^31 texels per volume (32 GiB); shard along z");
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, SDFGPU_ERR_CUDA, "no CUDA device (%s); libsdfgpu has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n_dev) return fail(nullptr, SDFGPU_ERR_INVALID, "device %d out of range", device);
    sdfgpu_ctx* ctx = new (std::nothrow) sdfgpu_ctx();
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "out of host memory");
    ctx->device = device;
    memcpy(ctx->bb, bb, sizeof ctx->bb);
    memcpy(ctx->dims, voxels, sizeof ctx->dims);
    ctx->z_begin = z_begin; ctx->z_end = z_end;
    ctx->z_lo = z_begin > 0 ? z_begin - 1 : 0;
    ctx->z_hi = z_end < voxels[2] ? z_end + 1 : voxels[2];
    if (z_begin == z_end) { ctx->z_lo = z_begin; ctx->z_hi = z_end; }
    ctx->lm.limits[0] = voxels[0]; ctx->lm.limits[1] = voxels[1]; ctx->lm.limits[2] = voxels[2];
    ctx->lm.reset(passes);
    for (unsigned i = 0; i < 256; ++i) ctx->lut[i] = srgb_u8_to_linear(i);
    build_pos_table(ctx->px, voxels[0], bb[0], bb[3]);
    build_pos_table(ctx->py, voxels[1], bb[1], bb[4]);
    build_pos_table(ctx->pz, voxels[2], bb[2], bb[5]);
    int rc = SDFGPU_OK;
    do {
        cudaError_t ce;
        if ((ce = cudaSetDevice(device)) != cudaSuccess ||
            (ce = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
            (ce = cudaDeviceGetAttribute(&ctx->cc_major, cudaDevAttrComputeCapabilityMajor, device)) != cudaSuccess ||
            (ce = cudaDeviceGetAttribute(&ctx->cc_minor, cudaDevAttrComputeCapabilityMinor, device)) != cudaSuccess ||
            (ce = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            (void)cudaGetLastError();
            rc = fail(nullptr, SDFGPU_ERR_CUDA, "device setup failed: %s", cudaGetErrorString(ce));
            break;
        }
        if ((rc = alloc_volumes(ctx)) != SDFGPU_OK) break;
        if ((rc = reset_volumes(ctx)) != SDFGPU_OK) break;
    } while (0);
    if (rc != SDFGPU_OK) {
        g_thread_error = ctx->err.empty() ? g_thread_error : ctx->err;
        sdfgpu_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SDFGPU_OK;
}

// Rust `as usize` on f32: saturating, NaN -> 0
uint32_t f32_as_usize(float f) {
    if (!(f == f) || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

}  // namespace

// ------------------------------------------------------------------ create

SDFGPU_API float sdfgpu_air_dist(void) { return air_dist_value(); }

SDFGPU_API int sdfgpu_dims_from_bb(const float bb[6], uint32_t max_voxels_side, uint32_t out[3]) {
    if (!bb || !out) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    // SDFViewer::from_bb, scene/sdf/mod.rs:47-68.  Iterator::max_by keeps the LAST maximum.
    volatile float sz[3] = {bb[3] - bb[0], bb[4] - bb[1], bb[5] - bb[2]};
    int max_dim = 0;
    for (int i = 1; i < 3; ++i)
        if (sz[i] >= sz[max_dim]) max_dim = i;
    for (int i = 0; i < 3; ++i) {
        if (i == max_dim) {
            out[i] = max_voxels_side;
        } else {
            volatile float f = (float)max_voxels_side * sz[i];
            f = f / sz[max_dim];
            out[i] = f32_as_usize(f);
        }
    }
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_create(const float bb[6], uint32_t max_voxels_side, uint32_t loading_passes, int device,
                             sdfgpu_ctx** out) {
    uint32_t dims[3];
    int rc = sdfgpu_dims_from_bb(bb, max_voxels_side, dims);
    if (rc != SDFGPU_OK) return rc;
    return create_common(bb, dims, loading_passes, device, 0, dims[2], out);
}

SDFGPU_API int sdfgpu_create_voxels(const float bb[6], const uint32_t voxels[3], uint32_t loading_passes, int device,
                                    sdfgpu_ctx** out) {
    if (!voxels) return fail(nullptr, SDFGPU_ERR_INVALID, "voxels is NULL");
    return create_common(bb, voxels, loading_passes, device, 0, voxels[2], out);
}

SDFGPU_API int sdfgpu_create_slab(const float bb[6], const uint32_t voxels[3], uint32_t loading_passes, int device,
                                  uint32_t z_begin, uint32_t z_end, sdfgpu_ctx** out) {
    return create_common(bb, voxels, loading_passes, device, z_begin, z_end, out);
}

SDFGPU_API int sdfgpu_ipc_detach(sdfgpu_ctx* ctx) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    if (ctx->stream) (void)cudaStreamSynchronize(ctx->stream);
    if (ctx->halo_stream) (void)cudaStreamSynchronize(ctx->halo_stream);
    if (ctx->copy_stream) (void)cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->copy_stream2) (void)cudaStreamSynchronize(ctx->copy_stream2);
    for (int side = 0; side < 2; ++side) {
        if (ctx->peer_tex0[side]) (void)cudaIpcCloseMemHandle(ctx->peer_tex0[side]);
        if (ctx->peer_tex1[side]) (void)cudaIpcCloseMemHandle(ctx->peer_tex1[side]);
        ctx->peer_tex0[side] = ctx->peer_tex1[side] = nullptr;
    }
    (void)cudaGetLastError();
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_ipc_export(sdfgpu_ctx* ctx, void* handles, size_t handles_bytes) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!handles || handles_bytes < 2 * sizeof(cudaIpcMemHandle_t))
        return fail(ctx, SDFGPU_ERR_INVALID, "handles buffer must hold 2 x %zu bytes", sizeof(cudaIpcMemHandle_t));
    if (!ctx->tex0 || !ctx->tex1) return fail(ctx, SDFGPU_ERR_STATE, "handle stores no voxels");
    set_device(ctx);
    cudaIpcMemHandle_t h[2];
    CK(ctx, cudaIpcGetMemHandle(&h[0], ctx->tex0));
    CK(ctx, cudaIpcGetMemHandle(&h[1], ctx->tex1));
    memcpy(handles, h, sizeof h);
    ctx->peers_ever = true;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_ipc_attach(sdfgpu_ctx* ctx, int side, const void* handles, size_t handles_bytes,
                                 uint32_t peer_z_lo, uint32_t peer_z_hi) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (side != 0 && side != 1) return fail(ctx, SDFGPU_ERR_INVALID, "side must be 0 (lower) or 1 (upper)");
    if (!handles || handles_bytes < 2 * sizeof(cudaIpcMemHandle_t)) return fail(ctx, SDFGPU_ERR_INVALID, "bad handles buffer");
    if (ctx->z_begin == ctx->z_end) return fail(ctx, SDFGPU_ERR_STATE, "this handle owns no slices");
    // the slice this rank mirrors must be one the neighbour stores (its halo)
    const uint32_t mirrored = side == 0 ? ctx->z_begin : ctx->z_end - 1;
    if (!(mirrored >= peer_z_lo && mirrored < peer_z_hi))
        return fail(ctx, SDFGPU_ERR_INVALID, "slice %u is not stored by the neighbour [%u,%u)", mirrored, peer_z_lo, peer_z_hi);
    if (ctx->peer_tex0[side]) return fail(ctx, SDFGPU_ERR_STATE, "side %d already attached", side);
    set_device(ctx);
    cudaIpcMemHandle_t h[2];
    memcpy(h, handles, sizeof h);
    void *p0 = nullptr, *p1 = nullptr;
    CK(ctx, cudaIpcOpenMemHandle(&p0, h[0], cudaIpcMemLazyEnablePeerAccess));
    cudaError_t e = cudaIpcOpenMemHandle(&p1, h[1], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        (void)cudaIpcCloseMemHandle(p0);
        (void)cudaGetLastError();
        return fail(ctx, SDFGPU_ERR_CUDA, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    }
    ctx->peer_tex0[side] = (float4*)p0;
    ctx->peer_tex1[side] = (float4*)p1;
    ctx->peer_z_lo[side] = peer_z_lo;
    return SDFGPU_OK;
}

SDFGPU_API void sdfgpu_destroy(sdfgpu_ctx* ctx) {
    if (!ctx) return;
    set_device(ctx);
    if (ctx->stream) (void)cudaStreamSynchronize(ctx->stream);
    link_free(ctx);
    mesh_free(ctx);
    (void)sdfgpu_ipc_detach(ctx);
    (void)cudaFree(ctx->trace_counters);
    (void)cudaFree(ctx->tex0); (void)cudaFree(ctx->tex1); (void)cudaFree(ctx->img_dev); (void)cudaFree(ctx->cell_lists);
    (void)cudaFree(ctx->rgba_dev); (void)cudaFree(ctx->depth_dev); (void)cudaFree(ctx->gbuf_dev);
    (void)cudaFree(ctx->keys_dev); (void)cudaFree(ctx->rgba8_dev); (void)cudaFree(ctx->touched_dev);
    (void)cudaFree(ctx->ingest_dev); (void)cudaFree(ctx->lut_dev); (void)cudaFree(ctx->dist_dev);
    (void)cudaFree(ctx->dist_full);
    (void)sdfgpu_gl_unregister(ctx);
    if (ctx->dist_tex[0]) (void)cudaDestroyTextureObject(ctx->dist_tex[0]);
    if (ctx->dist_tex[1]) (void)cudaDestroyTextureObject(ctx->dist_tex[1]);
    if (ctx->dist_surf) (void)cudaDestroySurfaceObject(ctx->dist_surf);
    if (ctx->dist_arr) (void)cudaFreeArray(ctx->dist_arr);
    for (auto& st : ctx->stage) {
        (void)cudaFreeHost(st.rec); (void)cudaFreeHost(st.idx); (void)cudaFree(st.rec_dev); (void)cudaFree(st.idx_dev);
        if (st.done) (void)cudaEventDestroy(st.done);
    }
    (void)cudaFreeHost(ctx->gather_host); (void)cudaFree(ctx->gather_dev);
    if (ctx->halo_stream) (void)cudaStreamDestroy(ctx->halo_stream);
    if (ctx->copy_stream) (void)cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_stream2) (void)cudaStreamDestroy(ctx->copy_stream2);
    (void)cudaFree(ctx->band_counters);
    (void)cudaFree(ctx->tile_cost[0]); (void)cudaFree(ctx->tile_cost[1]); (void)cudaFree(ctx->tile_order);
    if (ctx->ev_boundary) (void)cudaEventDestroy(ctx->ev_boundary);
    if (ctx->ev_pushed) (void)cudaEventDestroy(ctx->ev_pushed);
    if (ctx->stream) (void)cudaStreamDestroy(ctx->stream);
    (void)cudaGetLastError();
    delete ctx;
}

SDFGPU_API const char* sdfgpu_last_error(const sdfgpu_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_thread_error.c_str();
}

SDFGPU_API int sdfgpu_dims(const sdfgpu_ctx* ctx, uint32_t out_voxels[3]) {
    if (!ctx || !out_voxels) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    memcpy(out_voxels, ctx->dims, sizeof ctx->dims);
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_slab(const sdfgpu_ctx* ctx, uint32_t* z_begin, uint32_t* z_end, uint32_t* z_lo, uint32_t* z_hi) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (z_begin) *z_begin = ctx->z_begin;
    if (z_end) *z_end = ctx->z_end;
    if (z_lo) *z_lo = ctx->z_lo;
    if (z_hi) *z_hi = ctx->z_hi;
    return SDFGPU_OK;
}

// -------------------------------------------------------------------- tape

namespace {

struct ParsedTape {
    sdft_header h;
    std::vector<sdft_instr> low;  // lowered instructions (DeviceOp), operands kept
    std::vector<sdft_prim> prims;
    std::vector<float> consts;
    std::vector<sdft_sop> sops;     // scalar-program section
    // the structure: lowered opcodes up to and including DOP_END; when the tape has scalar programs their
    // text follows: n_programs, {pc, first, count} per program, n_sops, {op, a, b, c} per op
    std::vector<uint32_t> opcodes;
    uint32_t n_top_ops = 0;         // opcodes up to and including DOP_END
    bool has_scalar = false;
    uint32_t max_stack = 0;
    bool cull = false;
    uint32_t cull_first = 0, cull_count = 0;
};

// Validate a public tape (sdfgpu_tape.h) and lower it to the device opcodes (sdfgpu_device_types.h
// DeviceOp): primitive ops specialised by shape and material, stack ops by their static depth.
int parse_and_lower(sdfgpu_ctx* ctx, const void* tape, size_t tape_bytes, ParsedTape* out) {
    if (!tape) return fail(ctx, SDFGPU_ERR_INVALID, "tape is NULL");
    if (tape_bytes < sizeof(sdft_header)) return fail(ctx, SDFGPU_ERR_TAPE, "tape shorter than its header");
    sdft_header& h = out->h;
    memcpy(&h, tape, sizeof h);
    if (h.magic != SDFT_MAGIC) return fail(ctx, SDFGPU_ERR_TAPE, "bad tape magic 0x%08x", h.magic);
    if (h.version != SDFT_VERSION) return fail(ctx, SDFGPU_ERR_TAPE, "unsupported tape version %u", h.version);
    if (h.n_instr > SDFT_MAX_INSTR || h.n_prims > SDFT_MAX_PRIMS || h.n_consts > SDFT_MAX_CONSTS)
        return fail(ctx, SDFGPU_ERR_TAPE, "tape exceeds limits (%u instr, %u prims, %u consts)", h.n_instr, h.n_prims,
                    h.n_consts);
    const uint32_t n_sops = h.reserved[0];
    if (n_sops > SDFT_MAX_SOPS) return fail(ctx, SDFGPU_ERR_TAPE, "tape exceeds limits (%u scalar ops)", n_sops);
    if (h.reserved[1] || h.reserved[2]) return fail(ctx, SDFGPU_ERR_TAPE, "reserved header words must be 0");
    const size_t need = sizeof(sdft_header) + (size_t)h.n_instr * sizeof(sdft_instr) +
                        (size_t)h.n_prims * sizeof(sdft_prim) + (size_t)h.n_consts * 4 + (size_t)n_sops * sizeof(sdft_sop);
    if (need > tape_bytes) return fail(ctx, SDFGPU_ERR_TAPE, "tape truncated: %zu bytes needed, %zu given", need, tape_bytes);
    std::vector<sdft_instr> instr(h.n_instr);
    out->prims.resize(h.n_prims);
    out->consts.resize(h.n_consts);
    const std::vector<sdft_prim>& prims = out->prims;
    const unsigned char* b = (const unsigned char*)tape + sizeof h;
    if (h.n_instr) memcpy(instr.data(), b, instr.size() * sizeof(sdft_instr));
    b += instr.size() * sizeof(sdft_instr);
    if (h.n_prims) memcpy(out->prims.data(), b, prims.size() * sizeof(sdft_prim));
    b += prims.size() * sizeof(sdft_prim);
    if (h.n_consts) memcpy(out->consts.data(), b, out->consts.size() * 4);
    b += out->consts.size() * 4;
    out->sops.resize(n_sops);
    if (n_sops) memcpy(out->sops.data(), b, (size_t)n_sops * sizeof(sdft_sop));
    std::vector<uint32_t> programs;  // {pc, first, count} per SDFT_OP_SCALAR

    for (uint32_t k = 0; k < h.n_prims; ++k) {
        const uint32_t shape = prims[k].kind & 0xffu, mat = (prims[k].kind >> 8) & 0xffu;
        if (shape > SDFT_SHAPE_BOX_LINF || mat > SDFT_MAT_NORMAL || (prims[k].kind >> 16) != 0)
            return fail(ctx, SDFGPU_ERR_TAPE, "primitive %u: unknown kind 0x%x", k, prims[k].kind);
    }

    // ---- validate (operands in range, stack balanced) and lower in one walk
    out->low = instr;
    uint32_t depth = 0, max_depth = 0, n_ranges = 0;
    bool p_clean = true, cull_ok = false, ended = false;
    for (uint32_t pc = 0; pc < h.n_instr && !ended; ++pc) {
        const sdft_instr& S = instr[pc];
        sdft_instr& I = out->low[pc];
        switch (S.op) {
            case SDFT_OP_END: I.op = DOP_END; ended = true; break;
            case SDFT_OP_PRIM: case SDFT_OP_UNION_PRIM: case SDFT_OP_INTER_PRIM: {
                if (S.a >= h.n_prims) return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: primitive %u out of range", pc, S.a);
                const uint32_t mode = S.op - SDFT_OP_PRIM;
                const uint32_t shape = prims[S.a].kind & 0xffu, mat = (prims[S.a].kind >> 8) & 0xffu;
                I.op = DOP_PRIM + mode * 6 + shape * 3 + mat;
                break;
            }
            case SDFT_OP_UNION_RANGE:
                if (S.b < 1 || (uint64_t)S.a + S.b > h.n_prims)
                    return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: primitive range [%u,+%u) out of range", pc, S.a, S.b);
                I.op = DOP_UNION_RANGE;
                ++n_ranges;
                out->cull_first = S.a; out->cull_count = S.b; cull_ok = p_clean;
                break;
            case SDFT_OP_PUSH:
                if (depth >= SDFT_MAX_STACK) return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: stack overflow", pc);
                if (depth == 0) { I.op = DOP_PUSH_REG; I.b = 0; }
                else { I.op = DOP_PUSH_MEM; I.b = depth - 1; }
                ++depth;
                if (depth > max_depth) max_depth = depth;
                break;
            case SDFT_OP_POP_UNION: case SDFT_OP_POP_INTER: case SDFT_OP_POP_DEMO_DIFF: {
                if (depth == 0) return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: stack underflow", pc);
                if (S.op == SDFT_OP_POP_DEMO_DIFF && (uint64_t)S.a + 7 > h.n_consts)
                    return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: constants out of range", pc);
                const uint32_t kind = S.op - SDFT_OP_POP_UNION;
                --depth;
                if (depth == 0) { I.op = DOP_POP_UNION + kind; I.b = 0; }
                else { I.op = DOP_POP_UNION_MEM + kind; I.b = depth - 1; }
                break;
            }
            case SDFT_OP_D_NEG: I.op = DOP_D_NEG; break;
            case SDFT_OP_D_ABS: I.op = DOP_D_ABS; break;
            case SDFT_OP_D_ADD: I.op = DOP_D_ADD; break;
            case SDFT_OP_D_MUL: I.op = DOP_D_MUL; break;
            case SDFT_OP_D_MAX: I.op = DOP_D_MAX; break;
            case SDFT_OP_D_MIN: I.op = DOP_D_MIN; break;
            case SDFT_OP_M_SET:
                if ((uint64_t)S.a + 6 > h.n_consts) return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: constants out of range", pc);
                I.op = DOP_M_SET;
                break;
            case SDFT_OP_P_RESET: I.op = DOP_P_RESET; p_clean = true; break;
            case SDFT_OP_P_SUB:
                if ((uint64_t)S.a + 3 > h.n_consts) return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: constants out of range", pc);
                I.op = DOP_P_SUB; p_clean = false;
                break;
            case SDFT_OP_P_MUL: I.op = DOP_P_MUL; p_clean = false; break;
            case SDFT_OP_P_ABS: I.op = DOP_P_ABS; p_clean = false; break;
            case SDFT_OP_SCALAR: {
                if (S.b < 1 || (uint64_t)S.a + S.b > n_sops)
                    return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: scalar program [%u,+%u) out of range", pc, S.a, S.b);
                for (uint32_t i = 0; i < S.b; ++i) {  // operands name earlier values of the same program
                    const sdft_sop& o = out->sops[S.a + i];
                    int n_in;
                    switch (o.op) {
                        case SDFT_S_PX: case SDFT_S_PY: case SDFT_S_PZ: case SDFT_S_IMM: n_in = 0; break;
                        case SDFT_S_CONST:
                            if (o.a >= h.n_consts) return fail(ctx, SDFGPU_ERR_TAPE, "scalar op %u: constant %u out of range", S.a + i, o.a);
                            n_in = 0;
                            break;
                        case SDFT_S_FNEG: case SDFT_S_FABS: case SDFT_S_FSQRT: case SDFT_S_FFLOOR: case SDFT_S_FCEIL:
                        case SDFT_S_FTRUNC: case SDFT_S_FNEAREST: case SDFT_S_IEQZ: case SDFT_S_F_FROM_I_S:
                        case SDFT_S_F_FROM_I_U: case SDFT_S_I_FROM_F_S: case SDFT_S_I_FROM_F_U: n_in = 1; break;
                        case SDFT_S_OUT:
                            if (o.b > 6) return fail(ctx, SDFGPU_ERR_TAPE, "scalar op %u: output channel %u out of range", S.a + i, o.b);
                            n_in = 1;
                            break;
                        case SDFT_S_SELECT: n_in = 3; break;
                        default:
                            if ((o.op >= SDFT_S_FADD && o.op <= SDFT_S_FMOD) || (o.op >= SDFT_S_FEQ && o.op <= SDFT_S_FGE) ||
                                (o.op >= SDFT_S_IADD && o.op <= SDFT_S_IREM_S) || (o.op >= SDFT_S_IEQ && o.op <= SDFT_S_IGE_U) ||
                                o.op == SDFT_S_IREM_U) {
                                n_in = 2;
                                break;
                            }
                            return fail(ctx, SDFGPU_ERR_TAPE, "scalar op %u: unknown op %u", S.a + i, o.op);
                    }
                    if ((n_in >= 1 && o.a >= i) || (n_in >= 2 && o.b >= i) || (n_in >= 3 && o.c >= i))
                        return fail(ctx, SDFGPU_ERR_TAPE, "scalar op %u: operand does not name an earlier value", S.a + i);
                    // SDFT_S_OUT writes a channel and yields no value: naming it as an operand has no meaning
                    // (the specialiser emits no register for it)
                    const uint32_t ins[3] = {o.a, o.b, o.c};
                    for (int q = 0; q < n_in; ++q)
                        if (out->sops[S.a + ins[q]].op == SDFT_S_OUT)
                            return fail(ctx, SDFGPU_ERR_TAPE, "scalar op %u: operand %d names an OUT op, which yields no value",
                                        S.a + i, q);
                }
                I.op = DOP_SCALAR;
                programs.push_back(pc); programs.push_back(S.a); programs.push_back(S.b);
                break;
            }
            default: return fail(ctx, SDFGPU_ERR_TAPE, "instr %u: unknown op %u", pc, S.op);
        }
        out->opcodes.push_back(I.op);
    }
    if (!ended) out->opcodes.push_back(DOP_END);
    out->n_top_ops = (uint32_t)out->opcodes.size();
    out->has_scalar = !programs.empty();
    if (out->has_scalar) {  // the programs' text is part of the structure the kernel is specialised for
        out->opcodes.push_back((uint32_t)programs.size() / 3);
        out->opcodes.insert(out->opcodes.end(), programs.begin(), programs.end());
        out->opcodes.push_back(n_sops);
        for (const sdft_sop& o : out->sops) {
            out->opcodes.push_back(o.op); out->opcodes.push_back(o.a); out->opcodes.push_back(o.b); out->opcodes.push_back(o.c);
        }
    }
    out->max_stack = max_depth;
    out->cull = n_ranges == 1 && cull_ok && out->cull_count >= 16;
    return SDFGPU_OK;
}

}  // namespace

SDFGPU_API int sdfgpu_set_tape(sdfgpu_ctx* ctx, const void* tape, size_t tape_bytes) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    ParsedTape pt;
    const int prc = parse_and_lower(ctx, tape, tape_bytes, &pt);
    if (prc != SDFGPU_OK) return prc;
    const sdft_header& h = pt.h;

    // ---- build the shared-memory image
    TapeImageHeader ih;
    memset(&ih, 0, sizeof ih);
    ih.n_instr = h.n_instr; ih.n_prims = h.n_prims; ih.n_consts = h.n_consts;
    ih.max_stack = pt.max_stack;
    if (pt.cull) {
        ih.flags |= TAPE_FLAG_CULL;
        ih.cull_first = pt.cull_first; ih.cull_count = pt.cull_count;
    }
    uint32_t off = sizeof(TapeImageHeader);
    ih.off_instr = off; off += (h.n_instr + 1u) * 16u;  // + a terminating DOP_END (zero-initialised slot)
    ih.off_geom = off; off += h.n_prims * 16u;
    ih.off_mat0 = off; off += h.n_prims * 16u;
    ih.off_mat1 = off; off += h.n_prims * 16u;
    ih.off_consts = off; off += align16(h.n_consts * 4u);
    ih.off_lut = off; off += 1024u;
    ih.off_px = off; off += align16(ctx->dims[0] * 4u);
    ih.off_py = off; off += align16(ctx->dims[1] * 4u);
    ih.off_pz = off; off += align16(ctx->dims[2] * 4u);
    if (off > 200u * 1024u)
        return fail(ctx, SDFGPU_ERR_TAPE, "tape image is %u bytes; the fill kernel stages at most 204800 in shared memory", off);
    std::vector<unsigned char> img(off, 0);
    memcpy(img.data(), &ih, sizeof ih);
    if (h.n_instr) memcpy(img.data() + ih.off_instr, pt.low.data(), h.n_instr * 16u);
    for (uint32_t k = 0; k < h.n_prims; ++k) {
        const sdft_prim& pr = pt.prims[k];
        const float g[4] = {pr.center[0], pr.center[1], pr.center[2], pr.size};
        const float m0[4] = {pr.color[0], pr.color[1], pr.color[2], pr.metallic};
        float m1[4] = {pr.roughness, pr.occlusion, pr.air_skip, 0.0f};
        memcpy(&m1[3], &pr.kind, 4);
        memcpy(img.data() + ih.off_geom + 16u * k, g, 16);
        memcpy(img.data() + ih.off_mat0 + 16u * k, m0, 16);
        memcpy(img.data() + ih.off_mat1 + 16u * k, m1, 16);
    }
    if (h.n_consts) memcpy(img.data() + ih.off_consts, pt.consts.data(), h.n_consts * 4u);
    memcpy(img.data() + ih.off_lut, ctx->lut, 1024);
    if (ctx->dims[0]) memcpy(img.data() + ih.off_px, ctx->px.data(), ctx->dims[0] * 4u);
    if (ctx->dims[1]) memcpy(img.data() + ih.off_py, ctx->py.data(), ctx->dims[1] * 4u);
    if (ctx->dims[2]) memcpy(img.data() + ih.off_pz, ctx->pz.data(), ctx->dims[2] * 4u);

    set_device(ctx);
    if (img.size() > ctx->img_dev_cap) {
        // the previous image may still be read by a fill in flight
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        (void)cudaFree(ctx->img_dev);
        ctx->img_dev = nullptr; ctx->img_dev_cap = 0;
        CK(ctx, cudaMalloc(&ctx->img_dev, img.size()));
        ctx->img_dev_cap = img.size();
    }
    // stream-ordered after earlier fills; the source is pageable, so this call returns once it is staged
    CK(ctx, cudaMemcpyAsync(ctx->img_dev, img.data(), img.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->cell_lists_valid = false;
    if ((ih.flags & TAPE_FLAG_CULL) && ctx->dims[0] && ctx->dims[1] && ctx->dims[2]) {
        // coarse pre-cull: the survivors of every 64^3-voxel cell, so that a tile culls tens of candidates, not the range
        const uint32_t C = 1u << CULL_CELL_SHIFT;
        for (int a = 0; a < 3; ++a) ctx->cells[a] = (ctx->dims[a] + C - 1) / C;
        const size_t n_cells = (size_t)ctx->cells[0] * ctx->cells[1] * ctx->cells[2];
        const size_t need = n_cells * ih.cull_count + n_cells;
        if (need <= (size_t)1 << 28) {  // at most 1 GiB of lists: beyond that the tiles cull the whole range
            if (need > ctx->cell_lists_cap) {
                (void)cudaFree(ctx->cell_lists);
                ctx->cell_lists = nullptr; ctx->cell_lists_cap = 0;
                CK(ctx, cudaMalloc(&ctx->cell_lists, need * sizeof(uint32_t)));
                ctx->cell_lists_cap = need;
            }
            ctx->cell_counts = ctx->cell_lists + n_cells * ih.cull_count;
            CK(ctx, launch_cull_cells(ctx->img_dev, ctx->dims, ctx->cells[0], ctx->cells[1], ctx->cells[2], ctx->cell_lists,
                                      ctx->cell_counts, ctx->stream));
            ctx->launches++;
            ctx->cell_lists_valid = true;
        }
    }
    ctx->img_host.swap(img);
    ctx->hdr = ih;
    ctx->opcodes.swap(pt.opcodes);
    ctx->n_top_ops = pt.n_top_ops;
    ctx->has_scalar = pt.has_scalar;
    const uint32_t demo_ops[] = {DOP_PRIM + 0 * 6 + SDFT_SHAPE_BOX_LINF * 3 + SDFT_MAT_BRICK, DOP_PUSH_REG,
                                 DOP_PRIM + 0 * 6 + SDFT_SHAPE_SPHERE * 3 + SDFT_MAT_NORMAL, DOP_POP_DEMO_DIFF, DOP_END};
    ctx->structure_is_demo = ctx->opcodes.size() == 5 && !memcmp(ctx->opcodes.data(), demo_ops, sizeof demo_ops);
    ctx->has_tape = true;
    ctx->tape_bytes.assign((const unsigned char*)tape, (const unsigned char*)tape + tape_bytes);
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_tape_validate(const void* tape, size_t tape_bytes) {
    ParsedTape pt;
    return parse_and_lower(nullptr, tape, tape_bytes, &pt);
}

SDFGPU_API int sdfgpu_jit_check(const void* tape, size_t tape_bytes, int voxels_per_thread, char* log, size_t log_cap) {
    if (log && log_cap) log[0] = '\0';
    ParsedTape pt;
    const int prc = parse_and_lower(nullptr, tape, tape_bytes, &pt);
    if (prc != SDFGPU_OK) return prc;
    if (voxels_per_thread != 1 && voxels_per_thread != 2 && voxels_per_thread != 4 && voxels_per_thread != 8)
        return fail(nullptr, SDFGPU_ERR_INVALID, "voxels_per_thread must be 1, 2, 4 or 8");
    std::vector<char> cubin;
    std::string err;
    if (!jit_compile(pt.opcodes, voxels_per_thread, 10, 0, &cubin, &err)) {
        if (log && log_cap) snprintf(log, log_cap, "%s", err.c_str());
        return fail(nullptr, SDFGPU_ERR_CUDA, "%s", err.c_str());
    }
    if (log && log_cap) snprintf(log, log_cap, "%s", jit_source(pt.opcodes, voxels_per_thread).c_str());
    return (int)cubin.size() > 0 ? SDFGPU_OK : SDFGPU_ERR_CUDA;
}

// -------------------------------------------------------------------- fill

namespace {

// The head of SDFViewer::update: merge the reported box into the pending one (scene/sdf/mod.rs:131-139;
// merge_bounding_boxes, src/sdf/defaults.rs:59-72) and, when nothing is loading, start the 3-pass
// re-sample (:144-154).
void changed_box_state_machine(sdfgpu_ctx* ctx, const float* changed_box) {
    bool just_changed_box = false;
    if (changed_box) {
        if (ctx->has_changed_box) {
            for (int i = 0; i < 3; ++i) {
                ctx->changed_box[i] = fminf(ctx->changed_box[i], changed_box[i]);
                ctx->changed_box[3 + i] = fmaxf(ctx->changed_box[3 + i], changed_box[3 + i]);
            }
        } else {
            memcpy(ctx->changed_box, changed_box, sizeof ctx->changed_box);
            ctx->has_changed_box = true;
        }
        ctx->changed_box_while_loading = ctx->lm.len() > 0 || ctx->changed_box_while_loading;
        just_changed_box = true;
    }
    if (ctx->has_changed_box && ctx->lm.len() == 0) {
        ctx->lm.reset(3);
        if (!just_changed_box) {
            if (!ctx->changed_box_while_loading) ctx->has_changed_box = false;
            ctx->changed_box_while_loading = false;
        }
    }
}

// index range [first, last] of table entries inside [lo, hi] (closed, float compare, :187-189)
bool index_range_in(const std::vector<float>& t, float lo, float hi, uint32_t* first, uint32_t* last) {
    uint32_t f = 0xffffffffu, l = 0;
    for (uint32_t i = 0; i < t.size(); ++i)
        if (t[i] >= lo && t[i] <= hi) {
            if (f == 0xffffffffu) f = i;
            l = i;
        }
    if (f == 0xffffffffu) return false;
    *first = f; *last = l;
    return true;
}

// The loop of SDFViewer::update (:173-215) for a surface that has a tape: one LoadingManager pass per
// kernel launch, until nothing is pending or `max_passes` (0 = no limit) have run.  The launches
// are asynchronous and a pass takes well under a millisecond, so max_delta_time has nothing to bound.
int gpu_passes(sdfgpu_ctx* ctx, uint32_t max_passes, uint64_t* iterations) {
    const uint64_t start_iter = ctx->lm.total_iterations;
    uint32_t za, zb;
    fill_z_range(ctx, &za, &zb);
    uint32_t done = 0;
    while (ctx->lm.step_size != 0 && (max_passes == 0 || done < max_passes)) {
        const uint32_t step = (uint32_t)ctx->lm.step_size;
        uint32_t lo[3] = {0, 0, za}, hi[3] = {ctx->dims[0], ctx->dims[1], zb};
        // The rule "sample iff tex0.r == AIR_DIST or position in box" (:184-190) needs no read when the
        // host knows which voxels hold AIR_DIST (re-sampling a voxel whose stored value merely equals
        // AIR_DIST is idempotent: sample() is pure, src/sdf/mod.rs:43).
        int rc = SDFGPU_OK;
        // a pass the host-sampled path left half way: its voxels are a mix, read tex0.r
        const int64_t k = ctx->lm.iterations ? -1 : ctx->known_step;
        if (!ctx->has_changed_box && k == 0) {
            rc = run_fill(ctx, step, lo, hi, FILL_ALL, nullptr);                       // everything is AIR_DIST
            ctx->known_step = step;
        } else if (!ctx->has_changed_box && k == 1) {
            // nothing holds AIR_DIST: the pass samples nothing
        } else if (!ctx->has_changed_box && k > 1 && k % step == 0) {
            rc = run_fill(ctx, step, lo, hi, FILL_SKIP_KNOWN, nullptr, (uint32_t)k);   // skip the coarser lattice
            ctx->known_step = step;
        } else if (ctx->has_changed_box && k == 1) {
            // only the voxels inside the box: restrict the launch to its index AABB
            bool empty = false;
            const std::vector<float>* tab[3] = {&ctx->px, &ctx->py, &ctx->pz};
            for (int a = 0; a < 3 && !empty; ++a) {
                uint32_t first, last;
                if (!index_range_in(*tab[a], ctx->changed_box[a], ctx->changed_box[3 + a], &first, &last)) { empty = true; break; }
                if (first > lo[a]) lo[a] = first;
                if (last + 1 < hi[a]) hi[a] = last + 1;
            }
            if (!empty) rc = run_fill(ctx, step, lo, hi, FILL_BOX_ONLY, nullptr);
        } else {
            rc = run_fill(ctx, step, lo, hi, FILL_READ, nullptr);
            ctx->known_step = step == 1 ? 1 : -1;  // after a full step-1 pass no voxel holds AIR_DIST
        }
        if (rc != SDFGPU_OK) return rc;
        ctx->lm.finish_pass();
        ++done;
    }
    if (ctx->link.on) {  // collective: every rank signals, whether or not anything was pending
        const int rc = link_after_fill(ctx, done != 0, done != 0);
        if (rc != SDFGPU_OK) return rc;
    } else if (done && has_peers(ctx)) {
        const int rc = push_halos(ctx, ctx->stream);
        if (rc != SDFGPU_OK) return rc;
    }
    if (iterations) *iterations = ctx->lm.total_iterations - start_iter;  // :216
    return SDFGPU_OK;
}

int ensure_lut_dev(sdfgpu_ctx* ctx) {
    if (!ctx->lut_dev) {
        CK(ctx, cudaMalloc(&ctx->lut_dev, sizeof ctx->lut));
        CK(ctx, cudaMemcpyAsync(ctx->lut_dev, ctx->lut, sizeof ctx->lut, cudaMemcpyHostToDevice, ctx->stream));
    }
    return SDFGPU_OK;
}

constexpr size_t kMaxHostChunk = 1u << 20;  // LoadingManager iterations walked per chunk at most

int ensure_stage(sdfgpu_ctx* ctx, sdfgpu_ctx::HostStage& st, size_t voxels) {
    if (!st.done) CK(ctx, cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    if (st.in_flight) {  // the previous chunk that used this buffer must have been consumed
        CK(ctx, cudaEventSynchronize(st.done));
        st.in_flight = false;
    }
    if (voxels <= st.cap) return SDFGPU_OK;
    // pinned allocations are slow: size the buffer once for the largest chunk this handle can produce
    size_t cap = ctx->stored_texels < kMaxHostChunk ? ctx->stored_texels : kMaxHostChunk;
    if (cap < 4096) cap = 4096;
    while (cap < voxels) cap *= 2;
    (void)cudaFreeHost(st.rec); (void)cudaFreeHost(st.idx); (void)cudaFree(st.rec_dev); (void)cudaFree(st.idx_dev);
    st.rec = nullptr; st.idx = nullptr; st.rec_dev = nullptr; st.idx_dev = nullptr; st.cap = 0;
    CK(ctx, cudaMallocHost(&st.rec, cap * 7 * sizeof(float)));
    CK(ctx, cudaMallocHost(&st.idx, cap * sizeof(uint32_t)));
    CK(ctx, cudaMalloc(&st.rec_dev, cap * 7 * sizeof(float)));
    CK(ctx, cudaMalloc(&st.idx_dev, cap * sizeof(uint32_t)));
    st.cap = cap;
    return SDFGPU_OK;
}

// records + texel indices of a filled staging buffer -> device -> ingest_scatter_kernel
int scatter_stage(sdfgpu_ctx* ctx, sdfgpu_ctx::HostStage& st, size_t n) {
    CK(ctx, cudaMemcpyAsync(st.idx_dev, st.idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(st.rec_dev, st.rec, n * 7 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, launch_ingest_scatter(ctx->tex0, ctx->tex1, st.rec_dev, st.idx_dev, n, ctx->lut_dev, air_dist_value(),
                                  ctx->sm_count * 8, ctx->stream));
    CK(ctx, cudaEventRecord(st.done, ctx->stream));
    st.in_flight = true;
    ctx->launches++;
    ctx->dist_valid = false; ctx->dist_full_own_valid = false;
    return SDFGPU_OK;
}

// run fn(t) for t in [0, threads) on that many host threads (the calling thread takes t = 0)
template <class F>
void parallel_for_threads(size_t threads, F&& fn) {
    if (threads <= 1) { fn((size_t)0); return; }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (size_t t = 1; t < threads; ++t) pool.emplace_back([&fn, t] { fn(t); });
    fn((size_t)0);
    for (auto& th : pool) th.join();
}

// sdf.sample(pos, false) for n positions (:193)
void sample_range(const sdfgpu_surface* sdf, const float* xyz, size_t n, float* out) {
    if (!n) return;
    if (sdf->sample_batch) {
        sdf->sample_batch(sdf->self, xyz, (uint64_t)n, 0, out);
    } else {
        for (size_t i = 0; i < n; ++i) sdf->sample(sdf->self, xyz + 3 * i, 0, out + 7 * i);
    }
}

struct RowRun {  // `take` consecutive LoadingManager iterations inside one x row: x0 + i * step, y, z
    uint32_t x0, y, z, take;
};

// The loop of SDFViewer::update (:173-215) for a surface WITHOUT a tape: the LoadingManager is walked
// on the host in the reference's order, in chunks; per chunk the voxels that need an update are
// sampled through the callbacks and scattered into the volumes by ingest_scatter_kernel.  Deciding
// `update_required`, building positions and sampling all run on the surface's sample_threads.
int host_sampled_passes(sdfgpu_ctx* ctx, const sdfgpu_surface* sdf, double max_seconds, uint64_t* iterations) {
    LoadingState& lm = ctx->lm;
    const uint64_t start_iter = lm.total_iterations;
    const auto t_start = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    const uint64_t W = ctx->dims[0], H = ctx->dims[1], D = ctx->dims[2];
    const uint64_t slice = W * H;
    uint32_t za, zb;
    fill_z_range(ctx, &za, &zb);
    const float air = air_dist_value();
    int rc;
    if ((rc = ensure_lut_dev(ctx)) != SDFGPU_OK) return rc;
    if (W * H * D == 0) {  // an empty grid has nothing to visit
        while (lm.step_size != 0) lm.finish_pass();
        return SDFGPU_OK;
    }
    const size_t max_threads = sdf->sample_threads > 1 ? sdf->sample_threads : 1;
    std::vector<RowRun> runs;
    std::vector<size_t> block_first, block_count;  // per thread: first run / candidates
    std::vector<float> xyz;
    std::vector<uint32_t> cand;
    std::vector<unsigned char> cand_in_box;
    bool first = true, pass_finished = false;
    // iterations walked before the clock is read again: what the rate seen so far (kept across calls)
    // fits into three quarters of what is left of the budget, at least one iteration when it is zero (:173)
    auto chunk_for = [&](double seconds_left) -> uint64_t {
        if (!(seconds_left > 0.0)) return 1;
        const double want = ctx->host_rate * seconds_left * 0.75;
        return want < 256.0 ? 256 : want > (double)kMaxHostChunk ? kMaxHostChunk : (uint64_t)want;
    };
    uint64_t chunk = chunk_for(max_seconds);
    while (lm.step_size != 0 && (first || elapsed() < max_seconds)) {
        first = false;
        const double t_chunk = elapsed();
        const uint64_t step = lm.step_size;
        if (lm.iterations == 0) ctx->pass_known = ctx->known_step;
        // while a pass is half done the volume is a mix the other entry points cannot describe
        const int64_t k = ctx->pass_known;
        ctx->known_step = -1;
        const bool has_box = ctx->has_changed_box;
        const float* box = ctx->changed_box;
        // ---- walk up to `chunk` iterations of this pass from the cursor (loading.rs:50-76) as row runs
        runs.clear();
        uint64_t walked = 0;
        bool pass_end = false;
        while (walked < chunk && !pass_end) {
            uint64_t x0, y, z, st_;
            const uint64_t take = lm.next_row(chunk - walked, &x0, &y, &z, &st_, &pass_end);
            if (!take) break;
            walked += take;
            if (x0 >= W || y >= H || z >= D) continue;  // cannot happen on a non-empty grid
            if (z < za || z >= zb) continue;            // a slab handle skips the other ranks' slices
            runs.push_back(RowRun{(uint32_t)x0, (uint32_t)y, (uint32_t)z, (uint32_t)take});
        }
        // ---- update_required (:184-190) per visited voxel.  Without a read: k == 0 everything is AIR_DIST;
        // k == 1 nothing is (only the box matters); k = 2^j exactly the multiples of k have been sampled;
        // k == -1 unknown: every visited voxel is a candidate and tex0.r is read below.
        auto for_each_candidate = [&](const RowRun& r, auto&& emit) {
            const bool row_in_box = has_box && ctx->py[r.y] >= box[1] && ctx->py[r.y] <= box[4] && ctx->pz[r.z] >= box[2] &&
                                    ctx->pz[r.z] <= box[5];
            if (k == 1 && !row_in_box) return;
            for (uint32_t i = 0; i < r.take; ++i) {
                const uint64_t x = r.x0 + (uint64_t)i * step;
                const bool in_box = row_in_box && ctx->px[x] >= box[0] && ctx->px[x] <= box[3];
                const bool is_air = k == 0 || (k > 1 && ((x | r.y | r.z) & (uint64_t)(k - 1)) != 0);
                if (k == -1 || in_box || is_air) emit(x, in_box);
            }
        };
        // contiguous blocks of runs per thread, balanced by iterations
        size_t threads = max_threads;
        if (threads > walked / 4096 + 1) threads = (size_t)(walked / 4096 + 1);
        block_first.assign(threads + 1, runs.size());
        {
            uint64_t in_runs = 0;
            for (const RowRun& r : runs) in_runs += r.take;
            uint64_t acc = 0;
            size_t t = 0;
            for (size_t i = 0; i < runs.size(); ++i) {
                while (t < threads && acc >= in_runs * t / threads) block_first[t++] = i;
                acc += runs[i].take;
            }
            block_first[0] = 0;
        }
        block_count.assign(threads + 1, 0);
        parallel_for_threads(threads, [&](size_t t) {
            size_t n = 0;
            for (size_t i = block_first[t]; i < block_first[t + 1]; ++i) for_each_candidate(runs[i], [&](uint64_t, bool) { ++n; });
            block_count[t + 1] = n;
        });
        for (size_t t = 0; t < threads; ++t) block_count[t + 1] += block_count[t];  // -> offsets
        size_t n = block_count[threads];
        auto fill_block = [&](size_t t, uint32_t* idx_out, float* xyz_out, unsigned char* in_box_out) {
            size_t o = block_count[t];
            for (size_t i = block_first[t]; i < block_first[t + 1]; ++i) {
                const RowRun& r = runs[i];
                const uint32_t row = (uint32_t)((r.z - ctx->z_lo) * slice + (uint64_t)r.y * W);
                for_each_candidate(r, [&](uint64_t x, bool in_box) {
                    idx_out[o] = row + (uint32_t)x;
                    xyz_out[3 * o] = ctx->px[x]; xyz_out[3 * o + 1] = ctx->py[r.y]; xyz_out[3 * o + 2] = ctx->pz[r.z];
                    if (in_box_out) in_box_out[o] = in_box ? 1 : 0;
                    ++o;
                });
            }
        };
        if (n && k != -1) {
            // ---- known state: positions, samples and texel indices go straight into the pinned staging buffer
            sdfgpu_ctx::HostStage& st = ctx->stage[ctx->stage_turn & 1];
            ++ctx->stage_turn;
            if ((rc = ensure_stage(ctx, st, n)) != SDFGPU_OK) return rc;
            xyz.resize(3 * n);
            parallel_for_threads(threads, [&](size_t t) {
                fill_block(t, st.idx, xyz.data(), nullptr);
                const size_t a = block_count[t], b = block_count[t + 1];
                sample_range(sdf, xyz.data() + 3 * a, b - a, st.rec + 7 * a);
            });
            if ((rc = scatter_stage(ctx, st, n)) != SDFGPU_OK) return rc;
        } else if (n) {
            // ---- unknown state: read tex0.r of the candidates, keep those that hold AIR_DIST or lie in the box
            cand.resize(n); xyz.resize(3 * n); cand_in_box.resize(n);
            parallel_for_threads(threads, [&](size_t t) { fill_block(t, cand.data(), xyz.data(), cand_in_box.data()); });
            if (n > ctx->gather_cap) {
                CK(ctx, cudaStreamSynchronize(ctx->stream));
                (void)cudaFreeHost(ctx->gather_host); (void)cudaFree(ctx->gather_dev);
                ctx->gather_host = nullptr; ctx->gather_dev = nullptr; ctx->gather_cap = 0;
                size_t cap = ctx->stored_texels < kMaxHostChunk ? ctx->stored_texels : kMaxHostChunk;
                if (cap < 4096) cap = 4096;
                while (cap < n) cap *= 2;
                CK(ctx, cudaMallocHost(&ctx->gather_host, cap * sizeof(float)));
                CK(ctx, cudaMalloc(&ctx->gather_dev, cap * sizeof(float)));
                ctx->gather_cap = cap;
            }
            sdfgpu_ctx::HostStage& st = ctx->stage[ctx->stage_turn & 1];
            ++ctx->stage_turn;
            if ((rc = ensure_stage(ctx, st, n)) != SDFGPU_OK) return rc;
            memcpy(st.idx, cand.data(), n * sizeof(uint32_t));
            CK(ctx, cudaMemcpyAsync(st.idx_dev, st.idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
            CK(ctx, launch_gather_dist(ctx->tex0, st.idx_dev, n, ctx->gather_dev, ctx->sm_count * 8, ctx->stream));
            ctx->launches++;
            CK(ctx, cudaMemcpyAsync(ctx->gather_host, ctx->gather_dev, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            size_t m = 0;
            for (size_t i = 0; i < n; ++i)
                if (ctx->gather_host[i] == air || cand_in_box[i]) {
                    st.idx[m] = cand[i];
                    xyz[3 * m] = xyz[3 * i]; xyz[3 * m + 1] = xyz[3 * i + 1]; xyz[3 * m + 2] = xyz[3 * i + 2];
                    ++m;
                }
            n = m;
            if (n) {
                size_t th = max_threads;
                if (th > n / 4096 + 1) th = n / 4096 + 1;
                const size_t per = (n + th - 1) / th;
                parallel_for_threads(th, [&](size_t t) {
                    const size_t a = t * per < n ? t * per : n, b = (t + 1) * per < n ? (t + 1) * per : n;
                    sample_range(sdf, xyz.data() + 3 * a, b - a, st.rec + 7 * a);
                });
                if ((rc = scatter_stage(ctx, st, n)) != SDFGPU_OK) return rc;
            }
        }
        if (pass_end) {
            // sampled so far: what was known before this pass plus the lattice of `step`
            if (k == -1) ctx->known_step = step == 1 ? 1 : -1;
            else if (k == 0) ctx->known_step = (int64_t)step;
            else ctx->known_step = k < (int64_t)step ? k : (int64_t)step;
            pass_finished = true;
        }
        // the rate this chunk ran at (smoothed, kept for the next call) sizes the next chunk
        const double el = elapsed(), dt = el - t_chunk;
        if (dt > 0.0 && walked >= 256) {
            const double r = (double)walked / dt;
            ctx->host_rate = ctx->host_rate_known ? 0.5 * ctx->host_rate + 0.5 * r : r;
            ctx->host_rate_known = true;
        }
        chunk = chunk_for(max_seconds - el);
        if (chunk < 256 && max_seconds > 0.0) chunk = 256;
    }
    if (ctx->link.on) {
        // ranks stop at different places of a time-budgeted pass: every call pushes and signals, so the epochs agree
        if ((rc = link_after_fill(ctx, true, true)) != SDFGPU_OK) return rc;
    } else if (pass_finished && has_peers(ctx)) {
        if ((rc = push_halos(ctx, ctx->stream)) != SDFGPU_OK) return rc;
    }
    if (iterations) *iterations = lm.total_iterations - start_iter;  // :216
    return SDFGPU_OK;
}

}  // namespace

SDFGPU_API int sdfgpu_update(sdfgpu_ctx* ctx, const float* changed_box, uint32_t max_passes, uint64_t* iterations) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (iterations) *iterations = 0;
    if (!ctx->has_tape) return fail(ctx, SDFGPU_ERR_STATE, "no tape set (call sdfgpu_set_tape first)");
    set_device(ctx);
    changed_box_state_machine(ctx, changed_box);
    return gpu_passes(ctx, max_passes, iterations);
}

SDFGPU_API int sdfgpu_update_surface(sdfgpu_ctx* ctx, const sdfgpu_surface* sdf, double max_delta_seconds,
                                     uint64_t* iterations) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (iterations) *iterations = 0;
    if (!sdf) return fail(ctx, SDFGPU_ERR_INVALID, "sdf is NULL");
    if (max_delta_seconds != max_delta_seconds) return fail(ctx, SDFGPU_ERR_INVALID, "max_delta_seconds is NaN");
    const void* tape = nullptr;
    size_t tape_len = 0;
    const bool has_tape = sdf->tape && sdf->tape(sdf->self, &tape, &tape_len) && tape && tape_len;
    if (!has_tape && !sdf->sample && !sdf->sample_batch)
        return fail(ctx, SDFGPU_ERR_INVALID, "the surface has neither a tape nor a sample callback");
    set_device(ctx);
    float box[6];
    const bool changed = sdf->changed && sdf->changed(sdf->self, box);  // :130
    if (has_tape && (!ctx->has_tape || ctx->tape_bytes.size() != tape_len || memcmp(ctx->tape_bytes.data(), tape, tape_len))) {
        const int rc = sdfgpu_set_tape(ctx, tape, tape_len);
        if (rc != SDFGPU_OK) return rc;
    }
    changed_box_state_machine(ctx, changed ? box : nullptr);
    if (has_tape) return gpu_passes(ctx, 0, iterations);
    return host_sampled_passes(ctx, sdf, max_delta_seconds, iterations);
}

SDFGPU_API int sdfgpu_fill_all(sdfgpu_ctx* ctx) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    uint32_t za, zb;
    fill_z_range(ctx, &za, &zb);
    int rc;
    if (ctx->link.on && ctx->link.halo_push) {
        // ONE launch: the tiles of the two boundary slices first; the copy engines push them into the neighbours'
        // halo slices, behind the kernel's boundary flag, while the interior is being filled (link.cu)
        const uint32_t lo[3] = {0, 0, ctx->z_begin}, hi[3] = {ctx->dims[0], ctx->dims[1], ctx->z_end};
        ctx->fill_boundary_first = true;
        rc = run_fill(ctx, 1, lo, hi, FILL_ALL, nullptr);
        ctx->fill_boundary_first = false;
        if (rc != SDFGPU_OK) return rc;
        if ((rc = link_fill_all_pushed(ctx)) != SDFGPU_OK) return rc;
    } else if (!ctx->link.on && has_peers(ctx) && ctx->z_end - ctx->z_begin >= 3) {
        // fused halo exchange: fill the boundary slices first, then let the copy engines push them into
        // the neighbours' halo slices over NVLink WHILE the interior is being filled
        if (!ctx->halo_stream) {
            CK(ctx, cudaStreamCreateWithFlags(&ctx->halo_stream, cudaStreamNonBlocking));
            CK(ctx, cudaEventCreateWithFlags(&ctx->ev_boundary, cudaEventDisableTiming));
            CK(ctx, cudaEventCreateWithFlags(&ctx->ev_pushed, cudaEventDisableTiming));
        }
        const uint32_t lo0[3] = {0, 0, ctx->z_begin}, hi0[3] = {ctx->dims[0], ctx->dims[1], ctx->z_begin + 1};
        const uint32_t lo1[3] = {0, 0, ctx->z_end - 1}, hi1[3] = {ctx->dims[0], ctx->dims[1], ctx->z_end};
        if ((rc = run_fill(ctx, 1, lo0, hi0, FILL_ALL, nullptr)) != SDFGPU_OK) return rc;
        if ((rc = run_fill(ctx, 1, lo1, hi1, FILL_ALL, nullptr)) != SDFGPU_OK) return rc;
        CK(ctx, cudaEventRecord(ctx->ev_boundary, ctx->stream));
        CK(ctx, cudaStreamWaitEvent(ctx->halo_stream, ctx->ev_boundary, 0));
        if ((rc = push_halos(ctx, ctx->halo_stream)) != SDFGPU_OK) return rc;
        CK(ctx, cudaEventRecord(ctx->ev_pushed, ctx->halo_stream));
        const uint32_t lo[3] = {0, 0, ctx->z_begin + 1}, hi[3] = {ctx->dims[0], ctx->dims[1], ctx->z_end - 1};
        if ((rc = run_fill(ctx, 1, lo, hi, FILL_ALL, nullptr)) != SDFGPU_OK) return rc;
        CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pushed, 0));
    } else {
        const uint32_t lo[3] = {0, 0, za}, hi[3] = {ctx->dims[0], ctx->dims[1], zb};
        if ((rc = run_fill(ctx, 1, lo, hi, FILL_ALL, nullptr)) != SDFGPU_OK) return rc;
        if (ctx->link.on) rc = link_after_fill(ctx, true, true);  // halo slices filled above: counts the fill, nothing to exchange
        else if (has_peers(ctx)) rc = push_halos(ctx, ctx->stream);
        if (rc != SDFGPU_OK) return rc;
    }
    ctx->known_step = 1;
    while (ctx->lm.step_size != 0) ctx->lm.finish_pass();
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_cull_stats(sdfgpu_ctx* ctx, uint64_t* tiles, uint64_t* survivors_sum, uint64_t* survivors_max,
                                 uint32_t* primitives) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!ctx->has_tape) return fail(ctx, SDFGPU_ERR_STATE, "no tape set (call sdfgpu_set_tape first)");
    set_device(ctx);
    if (primitives) *primitives = (ctx->hdr.flags & TAPE_FLAG_CULL) ? ctx->hdr.cull_count : 0u;
    unsigned long long st[3] = {0, 0, 0};
    if (ctx->hdr.flags & TAPE_FLAG_CULL) {
        unsigned long long* dev = nullptr;
        CK(ctx, cudaMalloc(&dev, sizeof st));
        cudaError_t e = cudaMemsetAsync(dev, 0, sizeof st, ctx->stream);
        int rc = e == cudaSuccess ? SDFGPU_OK : fail(ctx, SDFGPU_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        if (rc == SDFGPU_OK) {
            ctx->cull_stats_dev = dev;
            rc = sdfgpu_fill_all(ctx);  // the statistics of one fill of every voxel (idempotent: sample() is pure)
            ctx->cull_stats_dev = nullptr;
        }
        if (rc == SDFGPU_OK) {
            e = cudaMemcpyAsync(st, dev, sizeof st, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = fail(ctx, SDFGPU_ERR_CUDA, "reading the statistics failed: %s", cudaGetErrorString(e));
        }
        (void)cudaFree(dev);
        if (rc != SDFGPU_OK) return rc;
    }
    if (survivors_sum) *survivors_sum = st[0];
    if (tiles) *tiles = st[1];
    if (survivors_max) *survivors_max = st[2];
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_resample_box(sdfgpu_ctx* ctx, const float box[6], uint64_t* voxels_touched) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (voxels_touched) *voxels_touched = 0;
    if (!box) return fail(ctx, SDFGPU_ERR_INVALID, "box is NULL");
    if (!ctx->has_tape) return fail(ctx, SDFGPU_ERR_STATE, "no tape set (call sdfgpu_set_tape first)");
    set_device(ctx);
    // index AABB of the voxels whose position lies in the closed box (float compare, :187-189)
    uint32_t lo[3], hi[3];
    const std::vector<float>* tab[3] = {&ctx->px, &ctx->py, &ctx->pz};
    bool empty = false;
    for (int a = 0; a < 3; ++a) {
        uint32_t first, last;
        if (!index_range_in(*tab[a], box[a], box[3 + a], &first, &last)) { empty = true; break; }  // nothing inside
        lo[a] = first; hi[a] = last + 1;
    }
    uint32_t za, zb;
    fill_z_range(ctx, &za, &zb);
    if (!empty) {
        if (lo[2] < za) lo[2] = za;
        if (hi[2] > zb) hi[2] = zb;
        if (lo[2] >= hi[2]) empty = true;  // the box lies in other ranks' slices
    }
    int rc = SDFGPU_OK;
    if (!empty) {
        // temporarily present `box` as the pending box of a conditional pass
        const bool saved_has = ctx->has_changed_box;
        float saved[6];
        memcpy(saved, ctx->changed_box, sizeof saved);
        ctx->has_changed_box = true;
        memcpy(ctx->changed_box, box, sizeof saved);
        if (voxels_touched) {
            cudaError_t e = cudaMemsetAsync(ctx->touched_dev, 0, sizeof(unsigned long long), ctx->stream);
            if (e != cudaSuccess) rc = fail(ctx, SDFGPU_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        }
        // voxels_touched counts the slab's OWN voxels: halo slices that are filled here too (fill_halo) go into launches
        // of their own, without the counter
        const uint32_t mode = ctx->known_step == 1 ? FILL_BOX_ONLY : FILL_READ;
        const uint32_t own_lo = lo[2] > ctx->z_begin ? lo[2] : ctx->z_begin, own_hi = hi[2] < ctx->z_end ? hi[2] : ctx->z_end;
        if (rc == SDFGPU_OK && voxels_touched && (lo[2] < own_lo || hi[2] > own_hi)) {
            const uint32_t parts[3][2] = {{lo[2], own_lo < hi[2] ? own_lo : hi[2]}, {own_lo, own_hi}, {own_hi > lo[2] ? own_hi : lo[2], hi[2]}};
            for (int k = 0; k < 3 && rc == SDFGPU_OK; ++k) {
                if (parts[k][0] >= parts[k][1]) continue;
                const uint32_t plo[3] = {lo[0], lo[1], parts[k][0]}, phi[3] = {hi[0], hi[1], parts[k][1]};
                rc = run_fill(ctx, 1, plo, phi, mode, k == 1 ? ctx->touched_dev : nullptr);
            }
        } else if (rc == SDFGPU_OK) {
            rc = run_fill(ctx, 1, lo, hi, mode, voxels_touched ? ctx->touched_dev : nullptr);
        }
        if (ctx->known_step != 1) ctx->known_step = -1;
        ctx->has_changed_box = saved_has;
        memcpy(ctx->changed_box, saved, sizeof saved);
    }
    // a neighbour's halo slice changes only when the box reaches this handle's first / last own slice
    const bool touch_lo = !empty && lo[2] <= ctx->z_begin && hi[2] > ctx->z_begin;
    const bool touch_hi = !empty && lo[2] < ctx->z_end && hi[2] >= ctx->z_end;
    if (rc == SDFGPU_OK && ctx->link.on) rc = link_after_fill(ctx, touch_lo, touch_hi);  // collective: always signals
    else if (rc == SDFGPU_OK && has_peers(ctx) && (touch_lo || touch_hi)) rc = push_halos(ctx, ctx->stream, touch_lo, touch_hi);
    if (rc != SDFGPU_OK) return rc;
    if (voxels_touched && !empty) {
        unsigned long long n = 0;
        CK(ctx, cudaMemcpyAsync(&n, ctx->touched_dev, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        *voxels_touched = n;
    }
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_voxel_positions(const sdfgpu_ctx* ctx, uint64_t first_flat, uint64_t count, float* xyz) {
    if (!ctx || (!xyz && count)) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    const uint64_t W = ctx->dims[0], H = ctx->dims[1], D = ctx->dims[2];
    if (first_flat + count > W * H * D || first_flat + count < first_flat)
        return fail(nullptr, SDFGPU_ERR_INVALID, "voxel range out of the grid");
    for (uint64_t i = 0; i < count; ++i) {  // flat = (z*H + y)*W + x, scene/sdf/mod.rs:177; position :179-182
        const uint64_t f = first_flat + i;
        xyz[3 * i + 0] = ctx->px[f % W];
        xyz[3 * i + 1] = ctx->py[(f / W) % H];
        xyz[3 * i + 2] = ctx->pz[f / (W * H)];
    }
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_ingest_samples(sdfgpu_ctx* ctx, uint64_t first_flat, uint64_t count, const void* samples) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (count == 0) return SDFGPU_OK;
    if (!samples) return fail(ctx, SDFGPU_ERR_INVALID, "samples is NULL");
    const uint64_t slice = (uint64_t)ctx->dims[0] * ctx->dims[1];
    const uint64_t lo = (uint64_t)ctx->z_lo * slice, hi = (uint64_t)ctx->z_hi * slice;
    if (first_flat < lo || first_flat + count > hi || first_flat + count < first_flat)
        return fail(ctx, SDFGPU_ERR_INVALID, "voxel range [%llu,+%llu) is outside the slices this handle stores",
                    (unsigned long long)first_flat, (unsigned long long)count);
    set_device(ctx);
    {
        const int rc = ensure_lut_dev(ctx);
        if (rc != SDFGPU_OK) return rc;
    }
    const size_t bytes = (size_t)count * 7 * sizeof(float);
    if (bytes > ctx->ingest_cap) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        (void)cudaFree(ctx->ingest_dev);
        ctx->ingest_dev = nullptr; ctx->ingest_cap = 0;
        CK(ctx, cudaMalloc(&ctx->ingest_dev, bytes));
        ctx->ingest_cap = bytes;
    }
    CK(ctx, cudaMemcpyAsync(ctx->ingest_dev, samples, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, launch_ingest(ctx->tex0, ctx->tex1, ctx->ingest_dev, (size_t)(first_flat - lo), (size_t)count, ctx->lut_dev,
                          air_dist_value(), ctx->sm_count * 8, ctx->stream));
    ctx->launches++;
    ctx->dist_valid = false; ctx->dist_full_own_valid = false;
    if (ctx->known_step != 1) ctx->known_step = -1;
    // the staging buffer is reused by the next call: wait until the kernel has consumed it
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_commit(sdfgpu_ctx* ctx) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    ctx->lod = exp2f((float)(uint8_t)ctx->lm.passes_left());  // scene/sdf/mod.rs:226
    if (ctx->lod == 1.0f) ctx->filter_linear = true;         // :227-230, never switched back
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_loading_state(const sdfgpu_ctx* ctx, uint64_t* len, uint64_t* total_iterations,
                                    uint32_t* passes_left, uint32_t* passes) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (len) *len = ctx->lm.len();
    if (total_iterations) *total_iterations = ctx->lm.total_iterations;
    if (passes_left) *passes_left = ctx->lm.passes_left();
    if (passes) *passes = (uint32_t)ctx->lm.passes;
    return SDFGPU_OK;
}

// ---- the LoadingManager as a device-free object (same LoadingState the handles use)
struct sdfgpu_loading {
    LoadingState lm;
};

SDFGPU_API int sdfgpu_loading_create(const uint32_t limits[3], uint32_t passes, sdfgpu_loading** out) {
    if (!out) return fail(nullptr, SDFGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!limits) return fail(nullptr, SDFGPU_ERR_INVALID, "limits is NULL");
    sdfgpu_loading* l = new (std::nothrow) sdfgpu_loading();
    if (!l) return fail(nullptr, SDFGPU_ERR_INVALID, "out of host memory");
    for (int a = 0; a < 3; ++a) l->lm.limits[a] = limits[a];
    l->lm.reset(passes);
    *out = l;
    return SDFGPU_OK;
}

SDFGPU_API void sdfgpu_loading_destroy(sdfgpu_loading* l) { delete l; }

SDFGPU_API void sdfgpu_loading_reset(sdfgpu_loading* l, uint32_t passes) {
    if (l) l->lm.reset(passes);
}

SDFGPU_API uint64_t sdfgpu_loading_next_run(sdfgpu_loading* l, uint64_t max_iters, uint32_t first[3], uint32_t* step) {
    if (!l) return 0;
    uint64_t x0 = 0, y = 0, z = 0, s = 0;
    bool pass_end;
    const uint64_t n = l->lm.next_row(max_iters, &x0, &y, &z, &s, &pass_end);
    if (n && first) { first[0] = (uint32_t)x0; first[1] = (uint32_t)y; first[2] = (uint32_t)z; }
    if (n && step) *step = (uint32_t)s;
    return n;
}

SDFGPU_API int sdfgpu_loading_next(sdfgpu_loading* l, uint32_t out_index[3]) {
    uint32_t step;
    return sdfgpu_loading_next_run(l, 1, out_index, &step) ? 1 : 0;
}

SDFGPU_API uint64_t sdfgpu_loading_len(const sdfgpu_loading* l) { return l ? l->lm.len() : 0; }
SDFGPU_API uint64_t sdfgpu_loading_total_iterations(const sdfgpu_loading* l) { return l ? l->lm.total_iterations : 0; }
SDFGPU_API uint32_t sdfgpu_loading_passes_left(const sdfgpu_loading* l) { return l ? l->lm.passes_left() : 0; }

SDFGPU_API int sdfgpu_reset(sdfgpu_ctx* ctx, uint32_t loading_passes) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    int rc;
    if (ctx->link.on && ctx->link.halo_push) {
        // a linked handle's halo slices are written by its neighbours only: reset the own slices and push the two
        // boundary slices like any fill does (collective: the neighbours do the same to this handle's halo slices).
        // Resetting the halo slices here would race with a neighbour that has already reset and filled again.
        ctx->dist_valid = false; ctx->dist_full_own_valid = false;
        const size_t slice = (size_t)ctx->dims[0] * ctx->dims[1];
        const size_t off = (size_t)(ctx->z_begin - ctx->z_lo) * slice, n = (size_t)(ctx->z_end - ctx->z_begin) * slice;
        const int grid = ctx->sm_count * 8;
        CK(ctx, launch_set_const(ctx->tex0 + off, n, air_dist_value(), grid, ctx->stream));
        CK(ctx, launch_set_const(ctx->tex1 + off, n, air_dist_value(), grid, ctx->stream));
        ctx->launches += 2;
        rc = link_after_fill(ctx, true, true);
    } else {
        rc = reset_volumes(ctx);
    }
    if (rc != SDFGPU_OK) return rc;
    ctx->lm.reset(loading_passes);
    ctx->known_step = 0;
    ctx->has_changed_box = false;
    ctx->changed_box_while_loading = false;
    ctx->lod = 1.0f;
    ctx->filter_linear = false;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_download(sdfgpu_ctx* ctx, float* tex0, float* tex1) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    const size_t slice = (size_t)ctx->dims[0] * ctx->dims[1];
    const size_t off = (size_t)(ctx->z_begin - ctx->z_lo) * slice;
    const size_t n = (size_t)(ctx->z_end - ctx->z_begin) * slice;
    if (n) {
        if (tex0) CK(ctx, cudaMemcpyAsync(tex0, ctx->tex0 + off, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        if (tex1) CK(ctx, cudaMemcpyAsync(tex1, ctx->tex1 + off, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_device_ptrs(sdfgpu_ctx* ctx, void** tex0, void** tex1) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (tex0) *tex0 = ctx->tex0;
    if (tex1) *tex1 = ctx->tex1;
    return SDFGPU_OK;
}

// ------------------------------------------------------------------- trace

namespace {

void mat4_mul(const float* a, const float* b, float* out) {  // column-major out = a * b
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) {
            volatile float s = 0.0f;
            for (int k = 0; k < 4; ++k) {
                volatile float t = a[k * 4 + rr] * b[c * 4 + k];
                s = s + t;
            }
            r[c * 4 + rr] = s;
        }
    memcpy(out, r, sizeof r);
}

void normalize3(float v[3]) {  // cgmath: v * (1 / |v|)
    volatile float m = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    volatile float inv = 1.0f / m;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

}  // namespace

int sdfgpu::ensure_frame(sdfgpu_ctx* ctx, uint32_t w, uint32_t h, bool want_gbuf, bool want_keys) {
    const size_t n = (size_t)w * h;
    if (w != ctx->fw || h != ctx->fh) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        // ... and whoever else may still read the old frame: a linked presenter's unpack and copies, the banded copies
        if (ctx->link.present_stream) CK(ctx, cudaStreamSynchronize(ctx->link.present_stream));
        if (ctx->copy_stream) CK(ctx, cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->copy_stream2) CK(ctx, cudaStreamSynchronize(ctx->copy_stream2));
        (void)cudaFree(ctx->rgba_dev); (void)cudaFree(ctx->depth_dev); (void)cudaFree(ctx->gbuf_dev);
        (void)cudaFree(ctx->keys_dev); (void)cudaFree(ctx->rgba8_dev);
        ctx->rgba_dev = nullptr; ctx->depth_dev = nullptr; ctx->gbuf_dev = nullptr; ctx->keys_dev = nullptr;
        ctx->rgba8_dev = nullptr;
        ctx->fw = ctx->fh = 0;
        if (n) {
            CK(ctx, cudaMalloc(&ctx->rgba_dev, n * sizeof(float4)));
            CK(ctx, cudaMalloc(&ctx->depth_dev, n * sizeof(float)));
        }
        ctx->fw = w; ctx->fh = h;
    }
    if (want_gbuf && !ctx->gbuf_dev && n) CK(ctx, cudaMalloc(&ctx->gbuf_dev, n * SDFGPU_GBUF_FLOATS * sizeof(float)));
    if (want_keys && !ctx->keys_dev && n) CK(ctx, cudaMalloc(&ctx->keys_dev, n * sizeof(unsigned long long)));
    return SDFGPU_OK;
}

int sdfgpu::fill_trace_params(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t w, uint32_t h, bool slab_clip, TraceParams* tp) {
    sdfgpu_rays rays;
    int rc = sdfgpu_camera_rays(cam, w, h, &rays);
    if (rc != SDFGPU_OK) return fail(ctx, rc, "%s", g_thread_error.c_str());
    memset(tp, 0, sizeof *tp);
    tp->tex0 = ctx->tex0; tp->tex1 = ctx->tex1;
    memcpy(tp->origin, rays.origin, 12); memcpy(tp->base, rays.base, 12);
    memcpy(tp->dx, rays.dx, 12); memcpy(tp->dy, rays.dy, 12); memcpy(tp->bvp, rays.bvp, 64);
    for (int a = 0; a < 3; ++a) {
        tp->bmin[a] = tp->clip_min[a] = ctx->bb[a];
        tp->bmax[a] = tp->clip_max[a] = ctx->bb[3 + a];
    }
    if (slab_clip) {  // texel slices [z_begin, z_end) <=> p01.z * D in [z_begin, z_end)
        volatile float size = ctx->bb[5] - ctx->bb[2];
        volatile float a0 = (float)ctx->z_begin / (float)ctx->dims[2], a1 = (float)ctx->z_end / (float)ctx->dims[2];
        if (ctx->z_begin > 0) { volatile float t = a0 * size; tp->clip_min[2] = ctx->bb[2] + t; }
        if (ctx->z_end < ctx->dims[2]) { volatile float t = a1 * size; tp->clip_max[2] = ctx->bb[2] + t; }
    }
    tp->size_pow2 = 1;
    for (int a = 0; a < 3; ++a) {
        volatile float size = tp->bmax[a] - tp->bmin[a];
        int e = 0;
        const float m = frexpf(size, &e);  // size = m * 2^e, m in [0.5, 1): a power of two has m == 0.5
        // the reciprocal must be a normal float too (|e| small) for x / size == x * (1 / size) to hold bit for bit
        if (!(m == 0.5f) || e < -60 || e > 60) tp->size_pow2 = 0;
        tp->inv_size[a] = 1.0f / size;
    }
    tp->W = ctx->dims[0]; tp->H = ctx->dims[1]; tp->D = ctx->dims[2];
    tp->z_lo = ctx->z_lo; tp->z_hi = ctx->z_hi;
    tp->lod = ctx->lod;
    tp->filter_linear = ctx->filter_linear ? 1u : 0u;
    memcpy(tp->tint, cam->tint, 16);
    tp->tone_mapping = cam->tone_mapping; tp->color_mapping = cam->color_mapping;
    tp->gamma = cam->gamma;
    memcpy(tp->ambient, cam->ambient, 12);
    // optional distance-only volume (4 B per voxel) for the march, rebuilt here after any change of tex0
    // (not with IPC neighbours, whose halo pushes this handle cannot observe): 1 = dense linear array,
    // 2 / 3 = R32F 3-D CUDA array behind a texture object with point / hardware-linear filtering
    if (ctx->opt_dist_volume && ctx->opt_trace_variant == 0 && ctx->stored_texels && !has_peers(ctx) && !ctx->peers_ever) {
        if (ctx->opt_dist_volume == 1) {
            if (!ctx->dist_dev) CK(ctx, cudaMalloc(&ctx->dist_dev, ctx->stored_texels * sizeof(float)));
            if (!ctx->dist_valid) {
                CK(ctx, launch_extract_dist(ctx->tex0, ctx->dist_dev, ctx->stored_texels, ctx->sm_count * 8, ctx->stream));
                ctx->launches++;
                ctx->dist_valid = true;
            }
            tp->dist = ctx->dist_dev;
        } else {
            const uint32_t Ds = ctx->z_hi - ctx->z_lo;
            if (!ctx->dist_arr) {
                const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float>();
                CK(ctx, cudaMalloc3DArray(&ctx->dist_arr, &fmt, make_cudaExtent(ctx->dims[0], ctx->dims[1], Ds),
                                          cudaArraySurfaceLoadStore));
                cudaResourceDesc res;
                memset(&res, 0, sizeof res);
                res.resType = cudaResourceTypeArray;
                res.res.array.array = ctx->dist_arr;
                CK(ctx, cudaCreateSurfaceObject(&ctx->dist_surf, &res));
                for (int lin = 0; lin < 2; ++lin) {
                    cudaTextureDesc td;
                    memset(&td, 0, sizeof td);
                    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
                    td.filterMode = lin ? cudaFilterModeLinear : cudaFilterModePoint;
                    td.readMode = cudaReadModeElementType;
                    td.normalizedCoords = 0;
                    CK(ctx, cudaCreateTextureObject(&ctx->dist_tex[lin], &res, &td, nullptr));
                }
            }
            if (!ctx->dist_valid) {
                CK(ctx, launch_extract_dist_array(ctx->tex0, (unsigned long long)ctx->dist_surf, ctx->dims[0], ctx->dims[1], Ds,
                                                  ctx->sm_count * 8, ctx->stream));
                ctx->launches++;
                ctx->dist_valid = true;
            }
            tp->dist_tex = (unsigned long long)ctx->dist_tex[ctx->opt_dist_volume == 3 ? 1 : 0];
        }
        tp->dist_mode = (uint32_t)ctx->opt_dist_volume;
    }
    tp->width = w; tp->height = h;
    tp->max_steps = (uint32_t)ctx->opt_max_steps;
    // screen rectangle of the projected clip box (a convex box projects inside the bounding rectangle of
    // its projected corners); +-2 pixels of slack; the whole frame if a corner is not in front of the camera
    tp->tiles_x = (w + 7) / 8; tp->tiles_y = (h + 7) / 8;
    tp->n_bands = 0; tp->band_rows = tp->tiles_y; tp->band_epoch = 0; tp->band_done = nullptr; tp->band_flags = nullptr;
    double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
    bool full = false;
    for (int c = 0; c < 8 && !full; ++c) {
        const double p[3] = {(c & 1) ? tp->clip_max[0] : tp->clip_min[0], (c & 2) ? tp->clip_max[1] : tp->clip_min[1],
                             (c & 4) ? tp->clip_max[2] : tp->clip_min[2]};
        const float* m = tp->bvp;
        const double X = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
        const double Y = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
        const double Wc = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
        if (!(Wc > 1e-4)) { full = true; break; }
        const double sx = X / Wc * w, sy = Y / Wc * h;
        if (!(sx == sx) || !(sy == sy)) { full = true; break; }
        x0 = sx < x0 ? sx : x0; x1 = sx > x1 ? sx : x1;
        y0 = sy < y0 ? sy : y0; y1 = sy > y1 ? sy : y1;
    }
    if (ctx->stored_texels == 0) {
        // an empty volume (a bounding box with a zero-size axis gives 0 voxels, scene/sdf/mod.rs:54-64):
        // nothing can be sampled, every pixel is a miss
        tp->rect[0] = tp->rect[1] = tp->rect[2] = tp->rect[3] = 0;
    } else if (full) {
        tp->rect[0] = tp->rect[1] = 0; tp->rect[2] = tp->tiles_x; tp->rect[3] = tp->tiles_y;
    } else {
        auto clampi = [](double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); };
        const uint32_t px0 = (uint32_t)clampi(floor(x0) - 2, 0, w), px1 = (uint32_t)clampi(ceil(x1) + 2, 0, w);
        const uint32_t py0 = (uint32_t)clampi(floor(y0) - 2, 0, h), py1 = (uint32_t)clampi(ceil(y1) + 2, 0, h);
        tp->rect[0] = px0 / 8; tp->rect[1] = py0 / 8;
        tp->rect[2] = (px1 + 7) / 8; tp->rect[3] = (py1 + 7) / 8;
        if (tp->rect[2] <= tp->rect[0] || tp->rect[3] <= tp->rect[1]) tp->rect[0] = tp->rect[1] = tp->rect[2] = tp->rect[3] = 0;
    }
    return SDFGPU_OK;
}

namespace {

// the tracer of an un-linked handle: variant 0 / 1 the per-tile kernels, variant 2 the persistent round kernel
// (one round, no hand-off; trace.cu trace_rounds_kernel) -- same frame, bit for bit
int launch_trace_any(sdfgpu_ctx* ctx, const TraceParams& tp_in) {
    const int variant = ctx->stored_texels ? ctx->opt_trace_variant : 0;
    TraceParams tp = tp_in;
    // Frame-to-frame coherence for the tile grid: every frame records the longest march of each tile inside the box's
    // rectangle, and the next frame of the same size starts those tiles longest-first (tile_order_kernel), so that the
    // frame does not end on a long march that happened to start late (ncu: the SMs idled for half of the kernel).
    // It pays when the box fills the screen (close-up at 512^3 / 1080p: 0.320 -> 0.242 ms): then the frame is bound by
    // throughput and ends on whichever long tiles started last.  When the box covers a minority of the frame (the
    // scene's default camera), the frame takes as long as its longest march, which no order shortens, and the 12 us of
    // sorting would be lost: "auto" (1) orders only when the rectangle holds at least half of the tiles; 2 = always.
    const uint32_t n_heavy = (tp.rect[2] - tp.rect[0]) * (tp.rect[3] - tp.rect[1]);
    // (Not for a frame in bands, sdfgpu_trace_rgba8: there the copy to the host is what the frame waits for, and the
    // sort only adds to it -- measured 0.555 -> 0.582 ms close-up.)
    const bool wanted = ctx->opt_tile_order == 2 ||
                        (ctx->opt_tile_order == 1 && tp.n_bands == 0 && 2u * n_heavy >= tp.tiles_x * tp.tiles_y);
    bool record = false;
    if (variant != 1 && !(variant == 2 && tp.dist_mode == 0 && !tp.full_dist) && wanted && n_heavy >= 256u) {
        const size_t n_tiles = (size_t)tp.tiles_x * tp.tiles_y;
        if (n_tiles > ctx->tile_cap) {
            (void)cudaFree(ctx->tile_cost[0]); (void)cudaFree(ctx->tile_cost[1]); (void)cudaFree(ctx->tile_order);
            ctx->tile_cost[0] = ctx->tile_cost[1] = ctx->tile_order = nullptr;
            ctx->tile_cap = 0; ctx->cost_valid = false;
            CK(ctx, cudaMalloc(&ctx->tile_cost[0], n_tiles * sizeof(uint32_t)));
            CK(ctx, cudaMalloc(&ctx->tile_cost[1], n_tiles * sizeof(uint32_t)));
            CK(ctx, cudaMalloc(&ctx->tile_order, n_tiles * sizeof(uint32_t)));
            ctx->tile_cap = n_tiles;
        }
        const int cur = ctx->cost_cur, prev = 1 - cur;
        CK(ctx, cudaMemsetAsync(ctx->tile_cost[cur], 0, n_tiles * sizeof(uint32_t), ctx->stream));
        if (ctx->cost_valid && ctx->cost_w == tp.width && ctx->cost_h == tp.height) {
            CK(ctx, launch_tile_order(tp, ctx->tile_cost[prev], ctx->tile_order, ctx->stream));
            ctx->launches++;
            tp.tile_order = ctx->tile_order;
        }
        tp.tile_cost = ctx->tile_cost[cur];
        record = true;
    }
    if (variant == 2 && tp.dist_mode == 0 && !tp.full_dist) {
        LinkParams lp;
        memset(&lp, 0, sizeof lp);
        lp.first = 1u;
        lp.own_z0 = 0; lp.own_z1 = ctx->dims[2];
        lp.work_head = ctx->trace_counters;
        lp.ctas_done = ctx->trace_counters + 1;
        const int per_sm = trace_rounds_max_ctas_per_sm(tp);
        if (per_sm < 1) return fail(ctx, SDFGPU_ERR_CUDA, "the trace kernel does not fit on an SM");
        CK(ctx, launch_trace_rounds(tp, lp, ctx->sm_count * per_sm, ctx->stream));
    } else {
        CK(ctx, launch_trace(tp, variant == 2 ? 0 : variant, ctx->stream));
    }
    ctx->launches++;
    if (record) {
        ctx->cost_cur = 1 - ctx->cost_cur;
        ctx->cost_valid = true; ctx->cost_w = tp.width; ctx->cost_h = tp.height;
    } else {
        ctx->cost_valid = false;  // the costs on record are not the previous frame's any more
    }
    return SDFGPU_OK;
}

}  // namespace

SDFGPU_API void sdfgpu_look_at_rh(const float eye[3], const float center[3], const float up[3], float m[16]) {
    // cgmath 0.18 Matrix4::look_to_rh (three-d Camera::set_view; call site scene/mod.rs:82-95)
    float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
    normalize3(f);
    float s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    normalize3(s);
    const float u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    const float es = eye[0] * s[0] + eye[1] * s[1] + eye[2] * s[2];
    const float eu = eye[0] * u[0] + eye[1] * u[1] + eye[2] * u[2];
    const float ef = eye[0] * f[0] + eye[1] * f[1] + eye[2] * f[2];
    const float r[16] = {s[0], u[0], -f[0], 0.f, s[1], u[1], -f[1], 0.f, s[2], u[2], -f[2], 0.f, -es, -eu, ef, 1.f};
    memcpy(m, r, sizeof r);
}

SDFGPU_API void sdfgpu_perspective(float fovy_rad, float aspect, float z_near, float z_far, float m[16]) {
    const float f = 1.0f / tanf(fovy_rad / 2.0f);  // cgmath PerspectiveFov -> Matrix4
    const float r[16] = {f / aspect, 0, 0, 0, 0, f, 0, 0, 0, 0, (z_far + z_near) / (z_near - z_far), -1.f,
                         0, 0, (2.f * z_far * z_near) / (z_near - z_far), 0};
    memcpy(m, r, sizeof r);
}

SDFGPU_API void sdfgpu_camera_default(sdfgpu_camera* cam, uint32_t width, uint32_t height) {
    if (!cam) return;
    memset(cam, 0, sizeof *cam);
    const float eye[3] = {2.5f, 3.0f, 5.0f}, center[3] = {0.f, 0.f, 0.f}, up[3] = {0.f, 1.f, 0.f};  // scene/mod.rs:89-91
    memcpy(cam->position, eye, 12);
    sdfgpu_look_at_rh(eye, center, up, cam->view);
    const float aspect = height ? (float)width / (float)height : 1.0f;
    sdfgpu_perspective(45.0f * 3.14159265358979323846f / 180.0f, aspect, 0.1f, 1000.0f, cam->projection);  // :92-94
    cam->tint[0] = cam->tint[1] = cam->tint[2] = cam->tint[3] = 1.0f;  // Srgba::WHITE, material.rs:29
    cam->tone_mapping = 2;   // three-d Camera default: ToneMapping::Aces
    cam->color_mapping = 1;  // ColorMapping::ComputeToSrgb
    cam->gamma = 0.0f;       // env "gamma" unset, material.rs:39
    cam->ambient[0] = cam->ambient[1] = cam->ambient[2] = 1.0f;  // AmbientLight(1.0, WHITE), scene/mod.rs:106
}

SDFGPU_API int sdfgpu_camera_rays(const sdfgpu_camera* cam, uint32_t width, uint32_t height, sdfgpu_rays* out) {
    if (!cam || !out) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    if (width == 0 || height == 0) return fail(nullptr, SDFGPU_ERR_INVALID, "empty frame");
    const float* V = cam->view;
    const float* P = cam->projection;
    if (P[0] == 0.0f || P[5] == 0.0f) return fail(nullptr, SDFGPU_ERR_INVALID, "singular projection");
    // rows of the view rotation: side, up, -forward
    const float s[3] = {V[0], V[4], V[8]}, u[3] = {V[1], V[5], V[9]}, f[3] = {-V[2], -V[6], -V[10]};
    const float sx = 1.0f / P[0], sy = 1.0f / P[5];
    for (int a = 0; a < 3; ++a) {
        out->origin[a] = cam->position[a];
        out->base[a] = f[a] - s[a] * sx - u[a] * sy;
        out->dx[a] = s[a] * (2.0f * sx / (float)width);
        out->dy[a] = u[a] * (2.0f * sy / (float)height);
    }
    const float bias[16] = {0.5f, 0, 0, 0, 0, 0.5f, 0, 0, 0, 0, 0.5f, 0, 0.5f, 0.5f, 0.5f, 1.0f};  // material.rs:90-95
    float pv[16];
    mat4_mul(P, V, pv);
    mat4_mul(bias, pv, out->bvp);
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_device(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                                   int want_gbuf, void** rgba_dev, void** depth_dev, void** gbuf_dev) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!cam) return fail(ctx, SDFGPU_ERR_INVALID, "cam is NULL");
    if (width == 0 || height == 0) return fail(ctx, SDFGPU_ERR_INVALID, "empty frame");
    if (ctx->link.on)
        return fail(ctx, SDFGPU_ERR_STATE, "a linked handle traces collectively: sdfgpu_trace_rgba8 / sdfgpu_trace_linked "
                                           "(the frame is composed on rank 0)");
    set_device(ctx);
    int rc = ensure_frame(ctx, width, height, want_gbuf != 0, false);
    if (rc != SDFGPU_OK) return rc;
    TraceParams tp;
    if ((rc = fill_trace_params(ctx, cam, width, height, false, &tp)) != SDFGPU_OK) return rc;
    tp.rgba = ctx->rgba_dev; tp.depth = ctx->depth_dev;
    tp.gbuf = want_gbuf ? ctx->gbuf_dev : nullptr;
    if ((rc = launch_trace_any(ctx, tp)) != SDFGPU_OK) return rc;
    if (rgba_dev) *rgba_dev = ctx->rgba_dev;
    if (depth_dev) *depth_dev = ctx->depth_dev;
    if (gbuf_dev) *gbuf_dev = want_gbuf ? ctx->gbuf_dev : nullptr;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height, float* rgba,
                            float* depth, float* gbuf) {
    void *r = nullptr, *d = nullptr, *g = nullptr;
    const int rc = sdfgpu_trace_device(ctx, cam, width, height, gbuf != nullptr, &r, &d, &g);
    if (rc != SDFGPU_OK) return rc;
    const size_t n = (size_t)width * height;
    if (rgba) CK(ctx, cudaMemcpyAsync(rgba, r, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (depth) CK(ctx, cudaMemcpyAsync(depth, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (gbuf) CK(ctx, cudaMemcpyAsync(gbuf, g, n * SDFGPU_GBUF_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_rgba8(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                                  uint8_t* rgba8, float* depth) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!cam) return fail(ctx, SDFGPU_ERR_INVALID, "cam is NULL");
    if (width == 0 || height == 0) return fail(ctx, SDFGPU_ERR_INVALID, "empty frame");
    if (ctx->link.on) return sdfgpu_trace_linked(ctx, cam, width, height, 0, rgba8, depth, nullptr);
    set_device(ctx);
    int rc = ensure_frame(ctx, width, height, false, false);
    if (rc != SDFGPU_OK) return rc;
    const size_t n = (size_t)width * height;
    if (!ctx->rgba8_dev) CK(ctx, cudaMalloc(&ctx->rgba8_dev, n * sizeof(uint32_t)));
    TraceParams tp;
    if ((rc = fill_trace_params(ctx, cam, width, height, false, &tp)) != SDFGPU_OK) return rc;
    tp.rgba8 = ctx->rgba8_dev; tp.depth = ctx->depth_dev;
    const int variant = ctx->stored_texels ? ctx->opt_trace_variant : 0;
    uint32_t bands = (uint32_t)ctx->opt_trace_bands;
    if (bands > tp.tiles_y) bands = tp.tiles_y;
    if (bands > TRACE_MAX_BANDS) bands = TRACE_MAX_BANDS;
    if (variant != 0 || bands < 2 || (!rgba8 && !depth) || !stream_wait_value_available()) {
        if ((rc = launch_trace_any(ctx, tp)) != SDFGPU_OK) return rc;
        if (rgba8) CK(ctx, cudaMemcpyAsync(rgba8, ctx->rgba8_dev, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (depth) CK(ctx, cudaMemcpyAsync(depth, ctx->depth_dev, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        return SDFGPU_OK;
    }
    // The frame in bands of tile rows, ONE launch: the last CTA of a band to finish raises the band's flag; the copy
    // engine, waiting on the flags in a second stream, takes the rows of a finished band to the host while the others
    // are still being traced -- the 8 bytes per pixel cross PCIe behind the march instead of after it.  Bands that lie
    // outside the screen rectangle of the box hold no ray: they are traced and copied first.
    // (One launch PER band was measured too: every launch pays the latency of its longest ray again, 0.47 -> 0.63 ms
    // at 12 bands.)  Same kernel, same pixels.
    if (!ctx->copy_stream) CK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_stream2) CK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
    if (!ctx->band_counters) {
        CK(ctx, cudaMalloc(&ctx->band_counters, 2 * 64 * sizeof(uint32_t)));
        CK(ctx, cudaMemset(ctx->band_counters, 0, 2 * 64 * sizeof(uint32_t)));  // before the copy stream first looks at a flag
    }
    tp.band_rows = (tp.tiles_y + bands - 1) / bands;
    tp.n_bands = (tp.tiles_y + tp.band_rows - 1) / tp.band_rows;
    tp.band_epoch = ++ctx->band_epoch;
    tp.band_done = ctx->band_counters; tp.band_flags = ctx->band_counters + 64;
    {   // bands without a ray first: traced (stores only) and on their way to the host while the others march
        uint32_t m = 0;
        for (uint32_t pass = 0; pass < 2; ++pass)
            for (uint32_t k = 0; k < tp.n_bands; ++k) {
                const bool holds_rays = k * tp.band_rows < tp.rect[3] && (k + 1) * tp.band_rows > tp.rect[1] && tp.rect[2] > tp.rect[0];
                if (holds_rays == (pass == 1)) tp.band_order[m++] = (uint8_t)k;
            }
    }
    if ((rc = launch_trace_any(ctx, tp)) != SDFGPU_OK) return rc;
    for (uint32_t i = 0; i < tp.n_bands; ++i) {
        const uint32_t k = tp.band_order[i];
        if (!stream_wait_value(ctx->copy_stream, tp.band_flags + k, tp.band_epoch)) {
            // nothing may still write the caller's buffers when this returns
            (void)cudaStreamSynchronize(ctx->stream); (void)cudaStreamSynchronize(ctx->copy_stream); (void)cudaStreamSynchronize(ctx->copy_stream2);
            return fail(ctx, SDFGPU_ERR_CUDA, "cuStreamWaitValue32 failed");
        }
        const uint32_t y0 = k * tp.band_rows * 8u, y1 = (k + 1) * tp.band_rows * 8u < height ? (k + 1) * tp.band_rows * 8u : height;
        const size_t off = (size_t)y0 * width, cnt = (size_t)(y1 - y0) * width;
        if (rgba8) CK(ctx, cudaMemcpyAsync(rgba8 + off * 4, ctx->rgba8_dev + off, cnt * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (depth) {  // colour and depth on a stream each: the fixed cost of the small copies overlaps
            if (!stream_wait_value(ctx->copy_stream2, tp.band_flags + k, tp.band_epoch)) {
                (void)cudaStreamSynchronize(ctx->stream); (void)cudaStreamSynchronize(ctx->copy_stream); (void)cudaStreamSynchronize(ctx->copy_stream2);
                return fail(ctx, SDFGPU_ERR_CUDA, "cuStreamWaitValue32 failed");
            }
            CK(ctx, cudaMemcpyAsync(depth + off, ctx->depth_dev + off, cnt * 4, cudaMemcpyDeviceToHost, ctx->copy_stream2));
        }
    }
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream2));
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream));  // the last copy follows the last band
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_params(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                                   int slab_clip, float clip_min[3], float clip_max[3], float* lod,
                                   uint32_t* filter_linear) {
    if (!ctx || !cam) return fail(ctx, SDFGPU_ERR_INVALID, "NULL argument");
    TraceParams tp;
    const int rc = fill_trace_params(ctx, cam, width, height, slab_clip != 0, &tp);
    if (rc != SDFGPU_OK) return rc;
    if (clip_min) memcpy(clip_min, tp.clip_min, 12);
    if (clip_max) memcpy(clip_max, tp.clip_max, 12);
    if (lod) *lod = tp.lod;
    if (filter_linear) *filter_linear = tp.filter_linear;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_slab_keys(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                                      void** keys_dev) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!cam || !keys_dev) return fail(ctx, SDFGPU_ERR_INVALID, "NULL argument");
    if (width == 0 || height == 0) return fail(ctx, SDFGPU_ERR_INVALID, "empty frame");
    set_device(ctx);
    int rc = ensure_frame(ctx, width, height, false, true);
    if (rc != SDFGPU_OK) return rc;
    TraceParams tp;
    if ((rc = fill_trace_params(ctx, cam, width, height, true, &tp)) != SDFGPU_OK) return rc;
    tp.keys = ctx->keys_dev;
    if ((rc = launch_trace_any(ctx, tp)) != SDFGPU_OK) return rc;
    *keys_dev = ctx->keys_dev;
    return SDFGPU_OK;
}

// ---- CUDA <-> GL interop presenter

SDFGPU_API int sdfgpu_gl_unregister(sdfgpu_ctx* ctx) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    for (auto& r : ctx->gl_res)
        if (r) {
            if (ctx->stream) (void)cudaStreamSynchronize(ctx->stream);
            (void)cudaGraphicsUnregisterResource(r);
            r = nullptr;
        }
    (void)cudaGetLastError();
    ctx->gl_w = ctx->gl_h = 0;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_gl_register(sdfgpu_ctx* ctx, uint32_t color_texture, uint32_t depth_texture, uint32_t gl_target,
                                  uint32_t width, uint32_t height) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (width == 0 || height == 0) return fail(ctx, SDFGPU_ERR_INVALID, "empty frame");
    if (color_texture == 0) return fail(ctx, SDFGPU_ERR_INVALID, "color_texture is 0");
    (void)sdfgpu_gl_unregister(ctx);
    set_device(ctx);
    const uint32_t tex[2] = {color_texture, depth_texture};
    for (int i = 0; i < 2; ++i) {
        if (!tex[i]) continue;
        const cudaError_t e = cudaGraphicsGLRegisterImage(&ctx->gl_res[i], tex[i], gl_target, cudaGraphicsRegisterFlagsWriteDiscard);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ctx->gl_res[i] = nullptr;
            (void)sdfgpu_gl_unregister(ctx);
            return fail(ctx, SDFGPU_ERR_CUDA,
                        "cudaGraphicsGLRegisterImage(texture %u) failed: %s (is the GL context that owns it current on this thread, "
                        "on this device?)", tex[i], cudaGetErrorString(e));
        }
    }
    ctx->gl_w = width; ctx->gl_h = height;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_gl(sdfgpu_ctx* ctx, const sdfgpu_camera* cam) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!cam) return fail(ctx, SDFGPU_ERR_INVALID, "cam is NULL");
    if (!ctx->gl_res[0]) return fail(ctx, SDFGPU_ERR_STATE, "no GL texture registered (call sdfgpu_gl_register first)");
    set_device(ctx);
    const uint32_t w = ctx->gl_w, h = ctx->gl_h;
    int rc = ensure_frame(ctx, w, h, false, false);
    if (rc != SDFGPU_OK) return rc;
    if (!ctx->rgba8_dev) CK(ctx, cudaMalloc(&ctx->rgba8_dev, (size_t)w * h * sizeof(uint32_t)));
    TraceParams tp;
    if ((rc = fill_trace_params(ctx, cam, w, h, false, &tp)) != SDFGPU_OK) return rc;
    tp.rgba8 = ctx->rgba8_dev; tp.depth = ctx->depth_dev;
    if ((rc = launch_trace_any(ctx, tp)) != SDFGPU_OK) return rc;
    // map, copy the frame into the textures' arrays on the device (row 0 = bottom row = GL's origin), unmap:
    // the frame never visits the host
    const int n = ctx->gl_res[1] ? 2 : 1;
    CK(ctx, cudaGraphicsMapResources(n, ctx->gl_res, ctx->stream));
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
        cudaArray_t arr = nullptr;
        e = cudaGraphicsSubResourceGetMappedArray(&arr, ctx->gl_res[i], 0, 0);
        if (e == cudaSuccess)
            e = cudaMemcpy2DToArrayAsync(arr, 0, 0, i == 0 ? (const void*)ctx->rgba8_dev : (const void*)ctx->depth_dev, (size_t)w * 4,
                                         (size_t)w * 4, h, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    const cudaError_t eu = cudaGraphicsUnmapResources(n, ctx->gl_res, ctx->stream);  // GL may use the textures after this
    if (e != cudaSuccess || eu != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ctx, SDFGPU_ERR_CUDA, "copy into the GL textures failed: %s", cudaGetErrorString(e != cudaSuccess ? e : eu));
    }
    return SDFGPU_OK;
}

// ---- exact multi-GPU trace: replicated full-grid distance volume + owner shading (include/sdfgpu.h)

SDFGPU_API int sdfgpu_exact_trace_prepare(sdfgpu_ctx* ctx, void** dist_dev, uint64_t* own_first, uint64_t* own_count) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    const size_t slice = (size_t)ctx->dims[0] * ctx->dims[1];
    const size_t total = slice * ctx->dims[2];
    if (!ctx->dist_full && total) CK(ctx, cudaMalloc(&ctx->dist_full, total * sizeof(float)));
    const size_t own = slice * (ctx->z_end - ctx->z_begin);
    if (own && !ctx->dist_full_own_valid) {
        CK(ctx, launch_extract_dist(ctx->tex0 + (size_t)(ctx->z_begin - ctx->z_lo) * slice,
                                    ctx->dist_full + (size_t)ctx->z_begin * slice, own, ctx->sm_count * 8, ctx->stream));
        ctx->launches++;
    }
    ctx->dist_full_own_valid = true;
    if (dist_dev) *dist_dev = ctx->dist_full;
    if (own_first) *own_first = (uint64_t)ctx->z_begin * slice;
    if (own_count) *own_count = own;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_dist_volume_read(sdfgpu_ctx* ctx, uint64_t first, uint64_t count, float* host_dst) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    const uint64_t total = (uint64_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2];
    if (!ctx->dist_full) return fail(ctx, SDFGPU_ERR_STATE, "no distance volume (call sdfgpu_exact_trace_prepare first)");
    if (first + count > total || first + count < first || (!host_dst && count)) return fail(ctx, SDFGPU_ERR_INVALID, "range out of the grid");
    set_device(ctx);
    if (count) CK(ctx, cudaMemcpyAsync(host_dst, ctx->dist_full + first, count * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_dist_volume_write(sdfgpu_ctx* ctx, uint64_t first, uint64_t count, const float* host_src) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    const uint64_t total = (uint64_t)ctx->dims[0] * ctx->dims[1] * ctx->dims[2];
    if (!ctx->dist_full) return fail(ctx, SDFGPU_ERR_STATE, "no distance volume (call sdfgpu_exact_trace_prepare first)");
    if (first + count > total || first + count < first || (!host_src && count)) return fail(ctx, SDFGPU_ERR_INVALID, "range out of the grid");
    set_device(ctx);
    if (count) CK(ctx, cudaMemcpyAsync(ctx->dist_full + first, host_src, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));  // the source is the caller's: it may be reused on return
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_exact_keys(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                                       void** keys_dev) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!cam || !keys_dev) return fail(ctx, SDFGPU_ERR_INVALID, "NULL argument");
    if (width == 0 || height == 0) return fail(ctx, SDFGPU_ERR_INVALID, "empty frame");
    if (!ctx->dist_full || !ctx->dist_full_own_valid)
        return fail(ctx, SDFGPU_ERR_STATE, "the volume changed since sdfgpu_exact_trace_prepare (prepare, gather, then trace)");
    set_device(ctx);
    int rc = ensure_frame(ctx, width, height, false, true);
    if (rc != SDFGPU_OK) return rc;
    TraceParams tp;
    if ((rc = fill_trace_params(ctx, cam, width, height, false, &tp)) != SDFGPU_OK) return rc;
    tp.dist = ctx->dist_full; tp.dist_mode = 0; tp.dist_tex = 0;
    tp.full_dist = 1;
    tp.own_z0 = ctx->z_begin; tp.own_z1 = ctx->z_end;
    tp.keys = ctx->keys_dev;
    CK(ctx, launch_trace(tp, 0, ctx->stream));
    ctx->launches++;
    *keys_dev = ctx->keys_dev;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_keys_download(sdfgpu_ctx* ctx, const void* keys_dev, uint32_t width, uint32_t height,
                                    uint8_t* rgba8, float* depth) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!keys_dev) return fail(ctx, SDFGPU_ERR_INVALID, "keys_dev is NULL");
    set_device(ctx);
    const size_t n = (size_t)width * height;
    if (n == 0 || n > 0xffffffffull) return fail(ctx, SDFGPU_ERR_INVALID, "bad frame size");
    uint8_t* r_dev = nullptr;
    float* d_dev = nullptr;
    int rc = SDFGPU_OK;
    cudaError_t e = cudaSuccess;
    do {
        if (rgba8 && (e = cudaMalloc(&r_dev, n * 4)) != cudaSuccess) break;
        if (depth && (e = cudaMalloc(&d_dev, n * 4)) != cudaSuccess) break;
        if ((e = launch_keys_unpack((const unsigned long long*)keys_dev, (uint32_t)n, r_dev, d_dev, ctx->stream)) != cudaSuccess) break;
        ctx->launches++;
        if (rgba8 && (e = cudaMemcpyAsync(rgba8, r_dev, n * 4, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
        if (depth && (e = cudaMemcpyAsync(depth, d_dev, n * 4, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
        e = cudaStreamSynchronize(ctx->stream);
    } while (0);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        rc = fail(ctx, SDFGPU_ERR_CUDA, "keys download failed: %s", cudaGetErrorString(e));
    }
    (void)cudaFree(r_dev); (void)cudaFree(d_dev);
    return rc;
}

// ------------------------------------------------------------------ stream

SDFGPU_API int sdfgpu_sync(sdfgpu_ctx* ctx) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->link.present_stream) CK(ctx, cudaStreamSynchronize(ctx->link.present_stream));  // a presenter's unpack and copies
    return SDFGPU_OK;
}

SDFGPU_API void* sdfgpu_stream(sdfgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

SDFGPU_API uint64_t sdfgpu_launch_count(const sdfgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

SDFGPU_API int sdfgpu_set_option(sdfgpu_ctx* ctx, const char* key, int64_t value) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!key) return fail(ctx, SDFGPU_ERR_INVALID, "key is NULL");
    if (!strcmp(key, "fill_voxels_per_thread")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8)
            return fail(ctx, SDFGPU_ERR_INVALID, "fill_voxels_per_thread must be 0, 1, 2, 4 or 8");
        ctx->opt_vpt = (int)value;
    } else if (!strcmp(key, "fill_ctas_per_sm")) {
        if (value < 0 || value > 32) return fail(ctx, SDFGPU_ERR_INVALID, "fill_ctas_per_sm out of range");
        ctx->opt_ctas = (int)value;
    } else if (!strcmp(key, "fill_halo")) {
        if (ctx->opt_fill_halo != (value != 0) && ctx->known_step != 0) ctx->known_step = -1;  // the filled z range changes
        ctx->opt_fill_halo = value != 0;
    } else if (!strcmp(key, "trace_distance_volume")) {
        if (value < 0 || value > 3) return fail(ctx, SDFGPU_ERR_INVALID, "trace_distance_volume must be 0..3");
        if (ctx->opt_dist_volume != (int)value) ctx->dist_valid = false; ctx->dist_full_own_valid = false;  // each form is rebuilt on its first use
        ctx->opt_dist_volume = (int)value;
    } else if (!strcmp(key, "trace_max_steps")) {
        if (value < 2 || value > 65536) return fail(ctx, SDFGPU_ERR_INVALID, "trace_max_steps out of range");
        ctx->opt_max_steps = (int)value;
    } else if (!strcmp(key, "trace_variant")) {
        if (value < 0 || value > 2) return fail(ctx, SDFGPU_ERR_INVALID, "trace_variant must be 0, 1 or 2");
        ctx->opt_trace_variant = (int)value;
    } else if (!strcmp(key, "fill_cull_cells")) {
        ctx->opt_cull_cells = value != 0;
    } else if (!strcmp(key, "trace_tile_order")) {
        if (value < 0 || value > 2) return fail(ctx, SDFGPU_ERR_INVALID, "trace_tile_order must be 0, 1 (auto) or 2 (always)");
        ctx->opt_tile_order = (int)value;
        ctx->cost_valid = false;
    } else if (!strcmp(key, "trace_bands")) {
        if (value < 1 || value > (int64_t)TRACE_MAX_BANDS) return fail(ctx, SDFGPU_ERR_INVALID, "trace_bands must be 1..%u", TRACE_MAX_BANDS);
        ctx->opt_trace_bands = (int)value;
    } else if (!strcmp(key, "link_wait_mode")) {
        if (value < 0 || value > 1) return fail(ctx, SDFGPU_ERR_INVALID, "link_wait_mode must be 0 or 1");
        if (ctx->link.on) return fail(ctx, SDFGPU_ERR_STATE, "set link_wait_mode before sdfgpu_link_attach");
        ctx->opt_link_wait = (int)value;
    } else if (!strcmp(key, "link_halo_push")) {
        if (value < 0 || value > 1) return fail(ctx, SDFGPU_ERR_INVALID, "link_halo_push must be 0 or 1");
        if (ctx->link.arena) return fail(ctx, SDFGPU_ERR_STATE, "set link_halo_push before sdfgpu_link_export");
        ctx->opt_link_halo_push = (int)value;
    } else if (!strcmp(key, "link_trace_mode")) {
        if (value < 0 || value > 2) return fail(ctx, SDFGPU_ERR_INVALID, "link_trace_mode must be 0 (auto), 1 (rounds) or 2 (stream)");
        if (ctx->link.arena) return fail(ctx, SDFGPU_ERR_STATE, "set link_trace_mode before sdfgpu_link_export");
        ctx->opt_link_trace_mode = (int)value;
    } else if (!strcmp(key, "link_timeout_ms")) {
        if (value < 1 || value > 3600000) return fail(ctx, SDFGPU_ERR_INVALID, "link_timeout_ms out of range");
        ctx->opt_link_timeout_ms = (int)value;
        ctx->link.timeout_ms = (uint32_t)value;
    } else if (!strcmp(key, "fill_program")) {
        if (value < 0 || value > 3) return fail(ctx, SDFGPU_ERR_INVALID, "fill_program must be 0..3");
        ctx->opt_program = (int)value;
    } else {
        return fail(ctx, SDFGPU_ERR_INVALID, "unknown option '%s'", key);
    }
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_get_info(const sdfgpu_ctx* ctx, const char* key, int64_t* value) {
    if (!ctx || !key || !value) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    if (!strcmp(key, "last_fill_program")) *value = ctx->last_program;
    else if (!strcmp(key, "last_fill_ctas_per_sm")) *value = ctx->last_ctas;
    else if (!strcmp(key, "last_fill_voxels_per_thread")) *value = ctx->last_vpt;
    else if (!strcmp(key, "sm_count")) *value = ctx->sm_count;
    else if (!strcmp(key, "device")) *value = ctx->device;
    else if (!strcmp(key, "tape_image_bytes")) *value = (int64_t)ctx->img_host.size();
    else if (!strcmp(key, "tape_culled")) *value = (ctx->hdr.flags & TAPE_FLAG_CULL) ? 1 : 0;
    else if (!strcmp(key, "jit_available")) { std::string why; *value = jit_available(&why) ? 1 : 0; }
    else if (!strcmp(key, "linked")) *value = ctx->link.on ? 1 : 0;
    else if (!strcmp(key, "link_memops")) *value = ctx->link.on && ctx->link.memops ? 1 : 0;
    else if (!strcmp(key, "link_halo_push")) *value = ctx->link.on && ctx->link.halo_push ? 1 : 0;
    else if (!strcmp(key, "link_trace_stream")) *value = ctx->link.on && ctx->link.stream ? 1 : 0;
    else if (!strcmp(key, "link_fill_epoch")) *value = ctx->link.fill_epoch;
    else if (!strcmp(key, "link_round_epoch")) *value = ctx->link.round_epoch;
    else return fail(nullptr, SDFGPU_ERR_INVALID, "unknown info key '%s'", key);
    return SDFGPU_OK;
}

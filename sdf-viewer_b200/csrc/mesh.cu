// mesh.cu -- isosurface mesh of the RESIDENT distance volume: marching cubes on the GPU, per-vertex material and
// normal through the tape, ASCII PLY.  SURVEY.md section 8f row 4.
//
// What it stands for in the reference (paths relative to /root/reference):
//   src/sdf/meshers/isosurface.rs:16-66   mesh(): MarchingCubes::<Signed>::new(N).extract(..) of the `isosurface`
//                                         crate samples the SDF on (N+1)^3 points of the unit cube mapped onto the
//                                         bounding box (:94-98) -- the lattice of an SDFViewer with N+1 voxels per
//                                         axis (scene/sdf/mod.rs:179-182), which the fill kernel has ALREADY sampled:
//                                         here nothing is re-sampled, the cells are classified from tex0.r
//   src/sdf/meshers/isosurface.rs:86-91   per-vertex normal = sdf.normal(p) -> defaults.rs:49-56 (4 tetrahedral taps)
//   src/sdf/meshers/mesh.rs:22-33         Mesh::postproc: one sample(p, false) per vertex for colour / metallic /
//                                         roughness / occlusion
//   src/sdf/meshers/mesh.rs:38-129        serialize_ply: ASCII PLY, the property list reproduced below
// The `isosurface` crate is un-vendored (Cargo.toml:91, git 185a0eb): its case table and vertex order cannot be
// restated, so vertex ORDER and triangulation of ambiguous cases are this build's ("parity unpinned"); the table is
// derived in mc_table.py.  Positions are in the SDF's own coordinates (what postproc samples at).
//
// Kernels (one thread per lattice column chunk: 32 x 8 columns per CTA, MESH_ZC points along z each):
//   count      per CTA: number of vertices (sign-changing edges owned by its lattice points) and triangles
//   scan       exclusive scan of the per-CTA counts (one CTA; a few thousand to a few hundred thousand entries)
//   vertices   positions (linear interpolation of tex0.r - 0.1 along the edge), the vertex id base of every lattice
//              point that owns a vertex, and the 5 sample positions of each vertex (itself + 4 normal taps)
//   [the tape, point mode: fill_device.cuh point_body]
//   finish     normal from the 4 tap distances, material from the vertex sample -> 12-float vertex records
//   triangles  index triples through the owners' vertex id bases
// Vertex and triangle order are deterministic (CTA, thread, z).
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>

#include "sdfgpu_ctx.h"
#include "mc_table.inc"

using namespace sdfgpu;

#define SDFGPU_API extern "C" __attribute__((visibility("default")))

namespace sdfgpu {
namespace {

constexpr int MESH_ZC = 8;        // lattice points along z per thread
constexpr int MESH_THREADS = 256;  // 32 (x) x 8 (y) columns
constexpr uint32_t VB_MASK = (1u << 29) - 1u;

__constant__ unsigned char c_mc_count[256];
__constant__ unsigned char c_mc_tris[256][3 * MC_MAX_TRIS];
__constant__ unsigned char c_mc_edge_owner[12][4];

struct MeshParams {
    const float4* tex0;
    uint32_t W, H, D;
    uint32_t tiles_x, tiles_y, tiles_z;
    const float* px;  // voxel positions per axis (the tape image's tables, scene/sdf/mod.rs:179-182)
    const float* py;
    const float* pz;
    uint32_t* block_counts;   // [2 * n_blocks]: vertices, triangles
    uint32_t* block_offsets;  // [2 * n_blocks] exclusive; totals at [2 * n_blocks], [2 * n_blocks + 1]
    uint32_t n_blocks;
    uint32_t* vert_base;      // per lattice point: vertex id of its first owned edge | crossing flags << 29
    float* vertices;          // 12 floats per vertex
    float* points;            // 5 positions per vertex
    const float* samples;     // 5 x 7 floats per vertex
    uint32_t* indices;        // 3 per triangle
    uint32_t n_vertices;
};

// distance as the shader decodes it (material.frag:56-60); lattice points beyond the grid never take part
__device__ __forceinline__ float ld_dist(const MeshParams& P, uint32_t x, uint32_t y, uint32_t z) {
    return __ldg(reinterpret_cast<const float*>(P.tex0 + ((size_t)z * P.H + y) * P.W + x)) - 1e-1f;
}

struct Column {
    uint32_t x, y, z0, nz;  // this thread's lattice column and its z range [z0, z0 + nz)
    bool in;                // column inside the grid
    bool has_x, has_y;      // the +x / +y neighbours exist
};

__device__ __forceinline__ Column column_of(const MeshParams& P) {
    Column c;
    const uint32_t b = blockIdx.x;
    const uint32_t tx = b % P.tiles_x, ty = (b / P.tiles_x) % P.tiles_y, tz = b / (P.tiles_x * P.tiles_y);
    c.x = tx * 32u + (threadIdx.x & 31u);
    c.y = ty * 8u + (threadIdx.x >> 5);
    c.z0 = tz * MESH_ZC;
    c.in = c.x < P.W && c.y < P.H;
    c.nz = c.in ? min((uint32_t)MESH_ZC, P.D - c.z0) : 0u;
    c.has_x = c.x + 1u < P.W;
    c.has_y = c.y + 1u < P.H;
    return c;
}

// Walks the column: f(z, d[8], flags, case, cell) for every lattice point.  d[c] is the distance of cell corner c
// (valid where the corner exists), flags bit a = the edge from this point along axis a changes sign, `cell` = the
// cell with this point as corner 0 exists.
template <typename F>
__device__ __forceinline__ void walk(const MeshParams& P, const Column& c, F&& f) {
    if (c.nz == 0u) return;
    float lo[4], hi[4];  // corners (x,y) (x+1,y) (x,y+1) (x+1,y+1) at z and z + 1
    const uint32_t x1 = c.has_x ? c.x + 1u : c.x, y1 = c.has_y ? c.y + 1u : c.y;
    lo[0] = ld_dist(P, c.x, c.y, c.z0); lo[1] = ld_dist(P, x1, c.y, c.z0);
    lo[2] = ld_dist(P, c.x, y1, c.z0); lo[3] = ld_dist(P, x1, y1, c.z0);
    for (uint32_t k = 0; k < c.nz; ++k) {
        const uint32_t z = c.z0 + k;
        const bool has_z = z + 1u < P.D;
        const uint32_t z1 = has_z ? z + 1u : z;
        hi[0] = ld_dist(P, c.x, c.y, z1); hi[1] = ld_dist(P, x1, c.y, z1);
        hi[2] = ld_dist(P, c.x, y1, z1); hi[3] = ld_dist(P, x1, y1, z1);
        const float d[8] = {lo[0], lo[1], lo[2], lo[3], hi[0], hi[1], hi[2], hi[3]};
        const bool in0 = d[0] < 0.0f;
        uint32_t flags = 0u;
        if (c.has_x && in0 != (d[1] < 0.0f)) flags |= 1u;
        if (c.has_y && in0 != (d[2] < 0.0f)) flags |= 2u;
        if (has_z && in0 != (d[4] < 0.0f)) flags |= 4u;
        const bool cell = c.has_x && c.has_y && has_z;
        uint32_t cs = 0u;
        if (cell) {
#pragma unroll
            for (int q = 0; q < 8; ++q) cs |= (d[q] < 0.0f ? 1u : 0u) << q;
        }
        f(z, d, flags, cs, cell);
#pragma unroll
        for (int q = 0; q < 4; ++q) lo[q] = hi[q];
    }
}

// exclusive scan of one value per thread over the CTA (thread order); *total = the CTA's sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t s_w[MESH_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    __syncthreads();  // s_w may still be read by a previous call
    if (lane == 31u) s_w[warp] = inc;
    __syncthreads();
    uint32_t base = 0u, sum = 0u;
#pragma unroll
    for (int w = 0; w < MESH_THREADS / 32; ++w) {
        const uint32_t c = s_w[w];
        if ((uint32_t)w < warp) base += c;
        sum += c;
    }
    *total = sum;
    return base + inc - v;
}

__global__ void __launch_bounds__(MESH_THREADS) mesh_count_kernel(const __grid_constant__ MeshParams P) {
    const Column c = column_of(P);
    uint32_t nv = 0u, nt = 0u;
    walk(P, c, [&](uint32_t, const float*, uint32_t flags, uint32_t cs, bool cell) {
        nv += __popc(flags);
        if (cell) nt += c_mc_count[cs];
    });
    uint32_t tv, tt;
    (void)block_exclusive_scan(nv, &tv);
    (void)block_exclusive_scan(nt, &tt);
    if (threadIdx.x == 0) {
        P.block_counts[2u * blockIdx.x] = tv;
        P.block_counts[2u * blockIdx.x + 1u] = tt;
    }
}

// one CTA of 1024 threads: exclusive scan of the two interleaved count arrays; totals behind the last entry
__global__ void __launch_bounds__(1024) mesh_scan_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                                         uint32_t n) {
    __shared__ uint32_t s_w[2][32];
    __shared__ uint32_t s_carry[2];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) s_carry[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        uint32_t v[2] = {i < n ? counts[2u * i] : 0u, i < n ? counts[2u * i + 1u] : 0u};
        uint32_t inc[2] = {v[0], v[1]};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, inc[0], o), b = __shfl_up_sync(0xffffffffu, inc[1], o);
            if (lane >= (uint32_t)o) { inc[0] += a; inc[1] += b; }
        }
        if (lane == 31u) { s_w[0][warp] = inc[0]; s_w[1][warp] = inc[1]; }
        __syncthreads();
        uint32_t wb[2] = {0u, 0u}, sum[2] = {0u, 0u};
        for (int w = 0; w < 32; ++w) {
            if ((uint32_t)w < warp) { wb[0] += s_w[0][w]; wb[1] += s_w[1][w]; }
            sum[0] += s_w[0][w]; sum[1] += s_w[1][w];
        }
        const uint32_t c0 = s_carry[0], c1 = s_carry[1];
        if (i < n) {
            offsets[2u * i] = c0 + wb[0] + inc[0] - v[0];
            offsets[2u * i + 1u] = c1 + wb[1] + inc[1] - v[1];
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] = c0 + sum[0]; s_carry[1] = c1 + sum[1]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[2u * n] = s_carry[0]; offsets[2u * n + 1u] = s_carry[1]; }
}

__global__ void __launch_bounds__(MESH_THREADS) mesh_vertices_kernel(const __grid_constant__ MeshParams P) {
    const Column c = column_of(P);
    uint32_t nv = 0u;
    walk(P, c, [&](uint32_t, const float*, uint32_t flags, uint32_t, bool) { nv += __popc(flags); });
    uint32_t total;
    uint32_t vid = P.block_offsets[2u * blockIdx.x] + block_exclusive_scan(nv, &total);
    if (nv == 0u) return;
    const float eps = 0.001f;  // SDFSurface::normal(p, None), src/sdf/defaults.rs:50
    const float kx[4] = {1.f, -1.f, -1.f, 1.f}, ky[4] = {-1.f, 1.f, -1.f, 1.f}, kz[4] = {-1.f, -1.f, 1.f, 1.f};  // :52-55
    walk(P, c, [&](uint32_t z, const float* d, uint32_t flags, uint32_t, bool) {
        if (!flags) return;
        P.vert_base[((size_t)z * P.H + c.y) * P.W + c.x] = vid | (flags << 29);
        const float bx = P.px[c.x], by = P.py[c.y], bz = P.pz[z];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!(flags & (1u << a))) continue;
            const float da = d[0], db = d[a == 0 ? 1 : a == 1 ? 2 : 4];
            const float t = (0.0f - da) / (db - da);
            float p[3] = {bx, by, bz};
            if (a == 0) p[0] = bx + t * (P.px[c.x + 1u] - bx);
            else if (a == 1) p[1] = by + t * (P.py[c.y + 1u] - by);
            else p[2] = bz + t * (P.pz[z + 1u] - bz);
            float* v = P.vertices + (size_t)vid * 12u;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
            float* q = P.points + (size_t)vid * 15u;
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                q[3 + 3 * k] = p[0] + kx[k] * eps; q[4 + 3 * k] = p[1] + ky[k] * eps; q[5 + 3 * k] = p[2] + kz[k] * eps;
            }
            ++vid;
        }
    });
}

// vertex records: position (written above), normal = normalize(sum k_i * d_i) (defaults.rs:52-55, cgmath normalize =
// v * (1 / |v|)), colour / metallic / roughness / occlusion of the vertex's own sample (mesh.rs:24-31)
__global__ void __launch_bounds__(256) mesh_finish_kernel(const __grid_constant__ MeshParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_vertices) return;
    const float* s = P.samples + (size_t)i * 35u;
    const float kx[4] = {1.f, -1.f, -1.f, 1.f}, ky[4] = {-1.f, 1.f, -1.f, 1.f}, kz[4] = {-1.f, -1.f, 1.f, 1.f};
    float nx = 0.f, ny = 0.f, nz = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float dk = s[7 * (k + 1)];
        nx = nx + kx[k] * dk; ny = ny + ky[k] * dk; nz = nz + kz[k] * dk;
    }
    const float inv = 1.0f / sqrtf((nx * nx + ny * ny) + nz * nz);
    float* v = P.vertices + (size_t)i * 12u;
    v[3] = nx * inv; v[4] = ny * inv; v[5] = nz * inv;
#pragma unroll
    for (int k = 0; k < 6; ++k) v[6 + k] = s[1 + k];
}

__global__ void __launch_bounds__(MESH_THREADS) mesh_triangles_kernel(const __grid_constant__ MeshParams P) {
    const Column c = column_of(P);
    uint32_t nt = 0u;
    walk(P, c, [&](uint32_t, const float*, uint32_t, uint32_t cs, bool cell) { if (cell) nt += c_mc_count[cs]; });
    uint32_t total;
    uint32_t tid = P.block_offsets[2u * blockIdx.x + 1u] + block_exclusive_scan(nt, &total);
    if (nt == 0u) return;
    walk(P, c, [&](uint32_t z, const float*, uint32_t, uint32_t cs, bool cell) {
        if (!cell) return;
        const uint32_t n = c_mc_count[cs];
        for (uint32_t k = 0; k < 3u * n; ++k) {
            const uint32_t e = c_mc_tris[cs][k];
            const uint32_t ox = c.x + c_mc_edge_owner[e][0], oy = c.y + c_mc_edge_owner[e][1], oz = z + c_mc_edge_owner[e][2];
            const uint32_t axis = c_mc_edge_owner[e][3];
            const uint32_t vb = P.vert_base[((size_t)oz * P.H + oy) * P.W + ox];
            // the owner's vertices are numbered in axis order: skip the crossing edges before this one
            P.indices[(size_t)tid * 3u + k] = (vb & VB_MASK) + __popc((vb >> 29) & ((1u << axis) - 1u));
        }
        tid += n;
    });
}

bool g_tables_loaded[64] = {};

int load_tables(sdfgpu_ctx* ctx) {
    if (ctx->device >= 0 && ctx->device < 64 && g_tables_loaded[ctx->device]) return SDFGPU_OK;
    CK(ctx, cudaMemcpyToSymbol(c_mc_count, k_mc_count, sizeof k_mc_count));
    CK(ctx, cudaMemcpyToSymbol(c_mc_tris, k_mc_tris, sizeof k_mc_tris));
    CK(ctx, cudaMemcpyToSymbol(c_mc_edge_owner, k_mc_edge_owner, sizeof k_mc_edge_owner));
    if (ctx->device >= 0 && ctx->device < 64) g_tables_loaded[ctx->device] = true;
    return SDFGPU_OK;
}

template <typename T>
int grow(sdfgpu_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return SDFGPU_OK;
    (void)cudaFree(*p);
    *p = nullptr; *cap = 0;
    const size_t n = need + need / 8 + 1024;
    CK(ctx, cudaMalloc(p, n * sizeof(T)));
    *cap = n;
    return SDFGPU_OK;
}

// Rust's `{}` of an f32 (what ply-rs writes): the shortest digits that round-trip, never an exponent
size_t fmt_f32(char* out, float v) {
    if (v != v) { memcpy(out, "NaN", 3); return 3; }
    char* o = out;
    if (std::signbit(v)) { *o++ = '-'; v = -v; }
    if (std::isinf(v)) { memcpy(o, "inf", 3); return (size_t)(o - out) + 3; }
    if (v == 0.0f) { *o++ = '0'; return (size_t)(o - out); }
    char sci[48];
    const auto r = std::to_chars(sci, sci + sizeof sci - 1, v, std::chars_format::scientific);  // d[.ddd]e[+-]XX, shortest
    *r.ptr = 0;
    const char* e = sci;
    while (*e != 'e') ++e;
    const int exp10 = atoi(e + 1);
    char digits[24];
    int nd = 0;
    for (const char* p = sci; p < e; ++p)
        if (*p != '.') digits[nd++] = *p;
    // value = 0.d1 d2 ... x 10^(exp10 + 1)
    const int point = exp10 + 1;  // digits before the decimal point
    if (point <= 0) {
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -point; ++i) *o++ = '0';
        for (int i = 0; i < nd; ++i) *o++ = digits[i];
    } else {
        for (int i = 0; i < point; ++i) *o++ = i < nd ? digits[i] : '0';
        if (nd > point) {
            *o++ = '.';
            for (int i = point; i < nd; ++i) *o++ = digits[i];
        }
    }
    return (size_t)(o - out);
}

// (c * 255.9999) as u8, mesh.rs:103-105 (saturating float-to-int cast, NaN -> 0)
unsigned color_u8(float c) {
    volatile float x = c * 255.9999f;
    if (!(x == x) || x <= 0.0f) return 0u;
    if (x >= 255.0f) return 255u;
    return (unsigned)x;
}

}  // namespace

int sample_points_device(sdfgpu_ctx* ctx, const float* points_dev, uint32_t n, float* out_dev) {
    if (!ctx->has_tape) return fail(ctx, SDFGPU_ERR_STATE, "no tape set (call sdfgpu_set_tape first)");
    if (n == 0) return SDFGPU_OK;
    FillParams p;
    memset(&p, 0, sizeof p);
    p.tape_img = ctx->img_dev; p.tape_img_bytes = (uint32_t)ctx->img_host.size();
    p.W = ctx->dims[0]; p.H = ctx->dims[1]; p.D = ctx->dims[2];
    p.nx = FILL_TILE_X; p.ny = FILL_TILE_Y; p.nz = (n + FILL_THREADS - 1) / FILL_THREADS;
    p.step = 1;
    p.tiles_x = 1; p.tiles_y = 1; p.tiles_z = p.nz;
    p.air_dist = air_dist_value();
    p.points = points_dev; p.points_out = out_dev; p.n_points = n;
    return dispatch_fill(ctx, p, 1);
}

}  // namespace sdfgpu

SDFGPU_API int sdfgpu_sample_points(sdfgpu_ctx* ctx, const float* xyz, uint64_t n, float* out) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (n == 0) return SDFGPU_OK;
    if (!xyz || !out) return fail(ctx, SDFGPU_ERR_INVALID, "NULL buffer");
    if (n > 0x7fffffffull) return fail(ctx, SDFGPU_ERR_INVALID, "at most 2^31 - 1 points per call");
    set_device(ctx);
    MeshState& M = ctx->mesh;
    int rc;
    if ((rc = grow(ctx, &M.points, &M.points_cap, (size_t)n * 3)) != SDFGPU_OK) return rc;
    if ((rc = grow(ctx, &M.samples, &M.samples_cap, (size_t)n * 7)) != SDFGPU_OK) return rc;
    M.valid = false;  // the staging buffers are shared with the mesher
    CK(ctx, cudaMemcpyAsync(M.points, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = sample_points_device(ctx, M.points, (uint32_t)n, M.samples)) != SDFGPU_OK) return rc;
    CK(ctx, cudaMemcpyAsync(out, M.samples, (size_t)n * 28, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_mesh(sdfgpu_ctx* ctx, uint64_t* n_vertices, uint64_t* n_triangles) {
    if (n_vertices) *n_vertices = 0;
    if (n_triangles) *n_triangles = 0;
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (ctx->z_lo != 0 || ctx->z_hi != ctx->dims[2] || ctx->z_begin != 0 || ctx->z_end != ctx->dims[2])
        return fail(ctx, SDFGPU_ERR_STATE, "the mesher needs a handle that holds the whole grid (not a slab)");
    if (!ctx->has_tape) return fail(ctx, SDFGPU_ERR_STATE, "no tape set: the vertex materials and normals are sampled through it");
    const uint32_t W = ctx->dims[0], H = ctx->dims[1], D = ctx->dims[2];
    if ((uint64_t)W * H * D == 0) return SDFGPU_OK;
    set_device(ctx);
    int rc;
    if ((rc = load_tables(ctx)) != SDFGPU_OK) return rc;
    MeshState& M = ctx->mesh;
    M.valid = false;
    MeshParams p;
    memset(&p, 0, sizeof p);
    p.tex0 = ctx->tex0;
    p.W = W; p.H = H; p.D = D;
    p.tiles_x = (W + 31u) / 32u; p.tiles_y = (H + 7u) / 8u; p.tiles_z = (D + MESH_ZC - 1u) / MESH_ZC;
    const uint64_t nb = (uint64_t)p.tiles_x * p.tiles_y * p.tiles_z;
    if (nb > 0x7fffffffull) return fail(ctx, SDFGPU_ERR_INVALID, "grid too large for the mesher");
    p.n_blocks = (uint32_t)nb;
    p.px = reinterpret_cast<const float*>(ctx->img_dev + ctx->hdr.off_px);
    p.py = reinterpret_cast<const float*>(ctx->img_dev + ctx->hdr.off_py);
    p.pz = reinterpret_cast<const float*>(ctx->img_dev + ctx->hdr.off_pz);
    if ((rc = grow(ctx, &M.block_counts, &M.block_counts_cap, (size_t)nb * 2)) != SDFGPU_OK) return rc;
    if ((rc = grow(ctx, &M.block_offsets, &M.block_offsets_cap, (size_t)nb * 2 + 2)) != SDFGPU_OK) return rc;
    p.block_counts = M.block_counts; p.block_offsets = M.block_offsets;
    mesh_count_kernel<<<p.n_blocks, MESH_THREADS, 0, ctx->stream>>>(p);
    mesh_scan_kernel<<<1, 1024, 0, ctx->stream>>>(M.block_counts, M.block_offsets, p.n_blocks);
    CK(ctx, cudaGetLastError());
    ctx->launches += 2;
    uint32_t totals[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(totals, M.block_offsets + (size_t)nb * 2, sizeof totals, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t nv = totals[0], nt = totals[1];
    if (nv > VB_MASK) return fail(ctx, SDFGPU_ERR_INVALID, "more than 2^29 vertices");
    M.n_vertices = nv; M.n_triangles = nt;
    if (nv && nt) {
        if ((rc = grow(ctx, &M.vert_base, &M.vert_base_cap, (size_t)W * H * D)) != SDFGPU_OK) return rc;
        if ((rc = grow(ctx, &M.vertices, &M.vertices_cap, (size_t)nv * 12)) != SDFGPU_OK) return rc;
        if ((rc = grow(ctx, &M.points, &M.points_cap, (size_t)nv * 15)) != SDFGPU_OK) return rc;
        if ((rc = grow(ctx, &M.samples, &M.samples_cap, (size_t)nv * 35)) != SDFGPU_OK) return rc;
        if ((rc = grow(ctx, &M.indices, &M.indices_cap, (size_t)nt * 3)) != SDFGPU_OK) return rc;
        p.vert_base = M.vert_base; p.vertices = M.vertices; p.points = M.points; p.samples = M.samples; p.indices = M.indices;
        p.n_vertices = (uint32_t)nv;
        mesh_vertices_kernel<<<p.n_blocks, MESH_THREADS, 0, ctx->stream>>>(p);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
        // Mesh::postproc + the normals: the tape at 5 positions per vertex (point mode of the fill kernel)
        const uint64_t total_pts = nv * 5;
        for (uint64_t first = 0; first < total_pts;) {
            const uint32_t n = (uint32_t)std::min<uint64_t>(total_pts - first, 1u << 30);
            if ((rc = sample_points_device(ctx, M.points + first * 3, n, M.samples + first * 7)) != SDFGPU_OK) return rc;
            first += n;
        }
        mesh_finish_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(p);
        mesh_triangles_kernel<<<p.n_blocks, MESH_THREADS, 0, ctx->stream>>>(p);
        CK(ctx, cudaGetLastError());
        ctx->launches += 2;
    } else {
        M.n_vertices = M.n_triangles = 0;
    }
    M.valid = true;
    if (n_vertices) *n_vertices = M.n_vertices;
    if (n_triangles) *n_triangles = M.n_triangles;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_mesh_download(sdfgpu_ctx* ctx, float* vertices, uint32_t* indices) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    MeshState& M = ctx->mesh;
    if (!M.valid) return fail(ctx, SDFGPU_ERR_STATE, "no mesh: call sdfgpu_mesh first");
    set_device(ctx);
    if (vertices && M.n_vertices)
        CK(ctx, cudaMemcpyAsync(vertices, M.vertices, M.n_vertices * 48, cudaMemcpyDeviceToHost, ctx->stream));
    if (indices && M.n_triangles)
        CK(ctx, cudaMemcpyAsync(indices, M.indices, M.n_triangles * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_mesh_device_ptrs(sdfgpu_ctx* ctx, const float** vertices_dev, const uint32_t** indices_dev) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    MeshState& M = ctx->mesh;
    if (!M.valid) return fail(ctx, SDFGPU_ERR_STATE, "no mesh: call sdfgpu_mesh first");
    if (vertices_dev) *vertices_dev = M.n_vertices ? M.vertices : nullptr;
    if (indices_dev) *indices_dev = M.n_triangles ? M.indices : nullptr;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_ply_serialize(const float* vertices, uint64_t n_vertices, const uint32_t* indices, uint64_t n_triangles,
                                    const char* comment, const char* path, uint64_t* bytes_written) {
    if (bytes_written) *bytes_written = 0;
    if ((n_vertices && !vertices) || (n_triangles && !indices) || !path) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(nullptr, SDFGPU_ERR_INVALID, "cannot open '%s' for writing", path);
    // header, mesh.rs:45-94 (element and property order as defined there)
    std::string head = "ply\nformat ascii 1.0\n";
    if (comment && *comment) head += std::string("comment ") + comment + "\n";
    head += "element vertex " + std::to_string(n_vertices) + "\n";
    for (const char* pr : {"x", "y", "z", "nx", "ny", "nz"}) head += std::string("property float ") + pr + "\n";
    for (const char* pr : {"red", "green", "blue"}) head += std::string("property uchar ") + pr + "\n";
    for (const char* pr : {"metallic", "roughness", "occlusion"}) head += std::string("property float ") + pr + "\n";
    head += "element face " + std::to_string(n_triangles) + "\nproperty list uchar int vertex_index\nend_header\n";
    uint64_t total = 0;
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
    total += head.size();
    // body: formatted in parallel, written in order
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 1;
    if (nthreads > 32) nthreads = 32;
    const uint64_t CHUNK = 1u << 16;
    auto emit = [&](uint64_t n_items, auto&& fmt_item) {
        for (uint64_t base = 0; base < n_items && ok; base += CHUNK * nthreads) {
            std::vector<std::string> parts(nthreads);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; ++t) {
                const uint64_t a = base + (uint64_t)t * CHUNK, b = std::min<uint64_t>(a + CHUNK, n_items);
                if (a >= b) break;
                th.emplace_back([&, t, a, b] {
                    std::string& s = parts[t];
                    s.reserve((size_t)(b - a) * 96);
                    char buf[1024];
                    for (uint64_t i = a; i < b; ++i) s.append(buf, fmt_item(buf, i));
                });
            }
            for (auto& x : th) x.join();
            for (unsigned t = 0; t < nthreads && ok; ++t) {
                ok = fwrite(parts[t].data(), 1, parts[t].size(), f) == parts[t].size();
                total += parts[t].size();
            }
        }
    };
    emit(n_vertices, [&](char* buf, uint64_t i) -> size_t {  // mesh.rs:97-113
        const float* v = vertices + i * 12;
        char* o = buf;
        for (int k = 0; k < 6; ++k) { o += fmt_f32(o, v[k]); *o++ = ' '; }
        for (int k = 6; k < 9; ++k) { o += (size_t)snprintf(o, 8, "%u", color_u8(v[k])); *o++ = ' '; }
        for (int k = 9; k < 12; ++k) { o += fmt_f32(o, v[k]); *o++ = k == 11 ? '\n' : ' '; }
        return (size_t)(o - buf);
    });
    emit(n_triangles, [&](char* buf, uint64_t i) -> size_t {  // mesh.rs:115-121
        const uint32_t* t = indices + i * 3;
        return (size_t)snprintf(buf, 64, "3 %d %d %d\n", (int)t[0], (int)t[1], (int)t[2]);
    });
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(nullptr, SDFGPU_ERR_INVALID, "writing '%s' failed", path);
    if (bytes_written) *bytes_written = total;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_mesh_write_ply(sdfgpu_ctx* ctx, const char* path, const char* comment, uint64_t* bytes_written) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    MeshState& M = ctx->mesh;
    if (!M.valid) return fail(ctx, SDFGPU_ERR_STATE, "no mesh: call sdfgpu_mesh first");
    std::vector<float> v((size_t)M.n_vertices * 12);
    std::vector<uint32_t> idx((size_t)M.n_triangles * 3);
    const int rc = sdfgpu_mesh_download(ctx, v.data(), idx.data());
    if (rc != SDFGPU_OK) return rc;
    const int rc2 = sdfgpu_ply_serialize(v.data(), M.n_vertices, idx.data(), M.n_triangles, comment, path, bytes_written);
    if (rc2 != SDFGPU_OK) return fail(ctx, rc2, "%s", sdfgpu_last_error(nullptr));
    return SDFGPU_OK;
}

void sdfgpu::mesh_free(sdfgpu_ctx* ctx) {
    MeshState& M = ctx->mesh;
    (void)cudaFree(M.vert_base); (void)cudaFree(M.block_counts); (void)cudaFree(M.block_offsets);
    (void)cudaFree(M.vertices); (void)cudaFree(M.indices); (void)cudaFree(M.points); (void)cudaFree(M.samples);
    (void)cudaGetLastError();
    M = MeshState();
}

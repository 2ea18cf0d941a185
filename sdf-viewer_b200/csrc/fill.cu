// fill.cu -- ahead-of-time instances of the grid-fill kernel (device code: fill_device.cuh):
// the tape interpreter (any tape) and the built-in straight-line program for the structure of the
// reference's own SDFDemo (/root/reference/src/sdf/demo/mod.rs:51-75).  Kernels specialised for
// other tape structures are compiled from the same device source at set_tape time (jit.cu).
#include "fill_device.cuh"
#include "sdfgpu_internal.h"

namespace sdfgpu {
namespace {

// MINB = CTAs per SM the register allocation aims at (V voxels per thread need ~ 30 + 25 V registers)
template <int V, int PROG, int MINB>
__global__ void __launch_bounds__(FILL_THREADS, MINB) fill_kernel(const FillParams P) {
    dev::fill_body<V, PROG>(P);
}

__global__ void __launch_bounds__(256) set_const_kernel(float4* __restrict__ dst, size_t n, float v) {
    const float4 val = make_float4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = val;
}

// Batched ingest of samples computed elsewhere (a CPU / WASM SDFSurface::sample):
// `samples` holds n SDFSample records (7 floats, src/sdf/mod.rs:104-118) for the n voxels that
// start at flat index `first` of the stored slab; applies the store rules of scene/sdf/mod.rs:196-208.
__global__ void __launch_bounds__(256) ingest_kernel(float4* __restrict__ tex0, float4* __restrict__ tex1,
                                                     const float* __restrict__ samples, size_t first, size_t n,
                                                     const float* __restrict__ lut, float air_dist) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float* r = samples + 7 * i;
        dev::Smp s;
        s.d = r[0]; s.r = r[1]; s.g = r[2]; s.b = r[3]; s.m = r[4]; s.ro = r[5]; s.o = r[6];
        float4 t0, t1;
        dev::store_rules(s, s_lut, air_dist, t0, t1);
        tex0[first + i] = t0;
        tex1[first + i] = t1;
    }
}

// Scattered form of the ingest for the reference's interlaced visit order (loading.rs:50-76): record
// i belongs to the stored texel idx[i].  Records are read 28 contiguous bytes per thread; the two
// 16-byte stores per voxel are strided by the pass's step, which is what the visit order dictates.
__global__ void __launch_bounds__(256) ingest_scatter_kernel(float4* __restrict__ tex0, float4* __restrict__ tex1,
                                                             const float* __restrict__ samples,
                                                             const uint32_t* __restrict__ idx, size_t n,
                                                             const float* __restrict__ lut, float air_dist) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float* r = samples + 7 * i;
        dev::Smp s;
        s.d = r[0]; s.r = r[1]; s.g = r[2]; s.b = r[3]; s.m = r[4]; s.ro = r[5]; s.o = r[6];
        float4 t0, t1;
        dev::store_rules(s, s_lut, air_dist, t0, t1);
        const uint32_t t = idx[i];
        tex0[t] = t0;
        tex1[t] = t1;
    }
}

// tex0.r of the stored texels idx[0..n): the `tex0[flat][0] == AIR_DIST` test of scene/sdf/mod.rs:184
// when the host does not know which voxels were sampled
__global__ void __launch_bounds__(256) gather_dist_kernel(const float4* __restrict__ tex0,
                                                          const uint32_t* __restrict__ idx, size_t n,
                                                          float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = reinterpret_cast<const float*>(tex0 + idx[i])[0];
}

// Coarse pre-cull of a culled UNION_RANGE (FillParams::cell_lists): one CTA per cell of CULL_CELL^3 voxels runs the
// fill kernel's own cull (dev::cull_box) against the positions of the whole cell, reading the tape image in global
// memory, and writes the survivors -- ascending -- to the cell's cull_count slots.
__global__ void __launch_bounds__(FILL_THREADS) cull_cells_kernel(const unsigned char* __restrict__ img, uint32_t W, uint32_t H,
                                                                  uint32_t D, uint32_t cells_x, uint32_t cells_y,
                                                                  uint32_t* __restrict__ lists, uint32_t* __restrict__ counts) {
    __shared__ float s_red[FILL_THREADS / 32];
    __shared__ uint32_t s_cnt[FILL_THREADS / 32];
    const TapeImageHeader* hdr = reinterpret_cast<const TapeImageHeader*>(img);
    const float4* geom = reinterpret_cast<const float4*>(img + hdr->off_geom);
    const float4* mat1 = reinterpret_cast<const float4*>(img + hdr->off_mat1);
    const float* px = reinterpret_cast<const float*>(img + hdr->off_px);
    const float* py = reinterpret_cast<const float*>(img + hdr->off_py);
    const float* pz = reinterpret_cast<const float*>(img + hdr->off_pz);
    const uint32_t cell = blockIdx.x, cx = cell % cells_x, cy = (cell / cells_x) % cells_y, cz = cell / (cells_x * cells_y);
    const uint32_t C = 1u << CULL_CELL_SHIFT;
    const float ax = px[cx * C], bx = px[min(cx * C + C, W) - 1u];
    const float ay = py[cy * C], by = py[min(cy * C + C, H) - 1u];
    const float az = pz[cz * C], bz = pz[min(cz * C + C, D) - 1u];
    const uint32_t n = dev::cull_box(geom, mat1, nullptr, hdr->cull_first, hdr->cull_count, fminf(ax, bx), fmaxf(ax, bx),
                                     fminf(ay, by), fmaxf(ay, by), fminf(az, bz), fmaxf(az, bz), s_red, s_cnt,
                                     lists + (size_t)cell * hdr->cull_count);
    if (threadIdx.x == 0) counts[cell] = n;
}

typedef void (*fill_fn)(const FillParams);

fill_fn pick(int V, int program) {
    const bool demo = program == dev::PROG_DEMO;
    switch (V) {
        case 1: return demo ? fill_kernel<1, dev::PROG_DEMO, 4> : fill_kernel<1, dev::PROG_INTERPRET, 4>;
        case 2: return demo ? fill_kernel<2, dev::PROG_DEMO, 3> : fill_kernel<2, dev::PROG_INTERPRET, 3>;
        case 4: return demo ? fill_kernel<4, dev::PROG_DEMO, 2> : fill_kernel<4, dev::PROG_INTERPRET, 2>;
        case 8: return demo ? fill_kernel<8, dev::PROG_DEMO, 2> : fill_kernel<8, dev::PROG_INTERPRET, 1>;
        default: return nullptr;
    }
}

}  // namespace

size_t fill_smem_bytes(uint32_t tape_img_bytes, uint32_t n_cull, uint32_t max_stack, int V, uint32_t* stack_floats) {
    // the top stack level lives in registers
    const uint32_t sf = (max_stack > 1 ? max_stack - 1 : 0) * 7u * (uint32_t)V * FILL_THREADS;
    if (stack_floats) *stack_floats = sf;
    return (size_t)tape_img_bytes + 16 + 64 + ((n_cull * 4u + 15u) & ~15u) + (size_t)sf * 4u;
}

cudaError_t fill_prepare(int V, int program, size_t smem_bytes) {
    fill_fn f = pick(V, program);
    if (!f) return cudaErrorInvalidValue;
    return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

int fill_max_ctas_per_sm(int V, int program, size_t smem_bytes) {
    fill_fn f = pick(V, program);
    int n = 0;
    if (!f || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, f, FILL_THREADS, smem_bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

cudaError_t launch_fill(const FillParams& p, int V, int program, int grid, size_t smem, cudaStream_t s) {
    fill_fn f = pick(V, program);
    if (!f) return cudaErrorInvalidValue;
    f<<<grid, FILL_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_cull_cells(const unsigned char* img_dev, const uint32_t dims[3], uint32_t cells_x, uint32_t cells_y,
                              uint32_t cells_z, uint32_t* lists, uint32_t* counts, cudaStream_t s) {
    cull_cells_kernel<<<cells_x * cells_y * cells_z, FILL_THREADS, 0, s>>>(img_dev, dims[0], dims[1], dims[2], cells_x, cells_y,
                                                                         lists, counts);
    return cudaGetLastError();
}

cudaError_t launch_ingest(float4* tex0, float4* tex1, const float* samples_dev, size_t first, size_t n,
                          const float* lut_dev, float air_dist, int grid, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    ingest_kernel<<<grid, 256, 0, s>>>(tex0, tex1, samples_dev, first, n, lut_dev, air_dist);
    return cudaGetLastError();
}

cudaError_t launch_ingest_scatter(float4* tex0, float4* tex1, const float* samples_dev, const uint32_t* idx_dev,
                                  size_t n, const float* lut_dev, float air_dist, int grid, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const size_t blocks = (n + 255) / 256;
    ingest_scatter_kernel<<<(int)(blocks < (size_t)grid ? blocks : (size_t)grid), 256, 0, s>>>(tex0, tex1, samples_dev,
                                                                                            idx_dev, n, lut_dev, air_dist);
    return cudaGetLastError();
}

cudaError_t launch_gather_dist(const float4* tex0, const uint32_t* idx_dev, size_t n, float* out_dev, int grid,
                               cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const size_t blocks = (n + 255) / 256;
    gather_dist_kernel<<<(int)(blocks < (size_t)grid ? blocks : (size_t)grid), 256, 0, s>>>(tex0, idx_dev, n, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_set_const(float4* dst, size_t n, float v, int grid, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    set_const_kernel<<<grid, 256, 0, s>>>(dst, n, v);
    return cudaGetLastError();
}

}  // namespace sdfgpu

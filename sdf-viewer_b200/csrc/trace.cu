// trace.cu -- the sphere tracer: material.frag main() + sdfRaycast
// (/root/reference/src/app/scene/sdf/material.frag:92-182), one ray per thread.
//
// The rasterised bounding cube that produces the fragments in the reference
// (src/app/scene/sdf/mod.rs:254-282, Cull::None material.rs:75-81) is replaced
// by a ray / AABB slab test per pixel.  The two RGBA32F volumes are read as
// linear [f32;4] arrays (x fastest) through the read-only L1/TEX path
// (ld.global.nc); filtering is point fetch + exact fp32 trilinear with GL
// texel-centre addressing and MIRRORED_REPEAT (scene/sdf/mod.rs:113-115) --
// hardware LINEAR filtering has ~8-bit weights and cannot meet 1e-5.  While
// marching only the distance lane (.r, 4 B) of each texel is fetched; the full
// texel is fetched once at the hit.
//
// Where the march reads its distances from (TraceParams::dist_mode, option trace_distance_volume):
//   0  the .r lane of tex0 in place (default)            1  a dense R32F copy of tex0.r in linear memory
//   2  an R32F 3-D CUDA array read through a texture object in POINT mode: the TMU and its block-linear
//      layout serve the 8 taps, the trilinear blend stays exact fp32 -- same values, same frame
//   3  the same array with hardware LINEAR filtering: one fetch per step, 8-bit weights -- a fast mode
//      that is NOT within the 1e-5 bar and is reported separately (hard part 2)
//
// A warp owns an 8 x 4 pixel tile so neighbouring rays share texels; finished
// lanes drop out and the warp leaves the loop as soon as its ballot is empty.
// CTAs are 8 x 8 pixel tiles in a 1-D grid ordered heavy-first (tiles inside the
// screen rectangle of the projected box), see trace_tiles_kernel.
//
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true (GLSL highp f32
// without contraction is what the oracle restates).
#include "sdfgpu_internal.h"

namespace sdfgpu {
namespace {

struct Vol {
    const float4* tex;
    int W, H, D, z_lo, z_hi;
};

// GL MIRRORED_REPEAT on an integer texel index
__device__ __forceinline__ int mirror_idx(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    int m = i % (2 * n);
    if (m < 0) m += 2 * n;
    return m < n ? m : 2 * n - 1 - m;
}

// The two taps x0, x0 + 1 of a LINEAR fetch.  For x0 in [-1, n - 1] -- every position inside the
// box -- MIRRORED_REPEAT of (x0, x0 + 1) is (max(x0, 0), min(x0 + 1, n - 1)); anything else takes
// the general path.
__device__ __forceinline__ void mirror_pair(int x0, int n, int& a, int& b) {
    if ((unsigned)(x0 + 1) <= (unsigned)n) {
        a = max(x0, 0); b = min(x0 + 1, n - 1);
    } else {
        a = mirror_idx(x0, n); b = mirror_idx(x0 + 1, n);
    }
}

__device__ __forceinline__ size_t texel_index(const Vol& v, int x, int y, int z) {
    x = mirror_idx(x, v.W); y = mirror_idx(y, v.H); z = mirror_idx(z, v.D);
    z = max(z, v.z_lo); z = min(z, v.z_hi - 1);  // slab storage: taps stay within the halo
    return ((size_t)(z - v.z_lo) * v.H + (size_t)y) * v.W + (size_t)x;
}

__device__ __forceinline__ float lerp1(float a, float b, float f) { return a + f * (b - a); }

struct Taps {  // texel indices fit 32 bits: sdfgpu_create rejects volumes of 2^32 texels or more per handle
    uint32_t i000, i100, i010, i110, i001, i101, i011, i111;
    float fx, fy, fz;
};

// the texel coordinates of the two taps per axis of a LINEAR fetch whose lower corner is (x0, y0, z0);
// z relative to the first stored slice
__device__ __forceinline__ void tap_coords(const Vol& v, int x0, int y0, int z0, int& xa, int& xb, int& ya, int& yb,
                                           int& za, int& zb) {
    mirror_pair(x0, v.W, xa, xb);
    mirror_pair(y0, v.H, ya, yb);
    mirror_pair(z0, v.D, za, zb);
    za = min(max(za, v.z_lo), v.z_hi - 1) - v.z_lo;
    zb = min(max(zb, v.z_lo), v.z_hi - 1) - v.z_lo;
}

// texture(sampler3D, p01) with GL_LINEAR: texel centres at (i + 0.5) / N
__device__ __forceinline__ Taps linear_taps(const Vol& v, float ax, float ay, float az) {
    const float ux = ax * (float)v.W - 0.5f, uy = ay * (float)v.H - 0.5f, uz = az * (float)v.D - 0.5f;
    const float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
    Taps t;
    t.fx = ux - fx0; t.fy = uy - fy0; t.fz = uz - fz0;
    int xa, xb, ya, yb, za, zb;
    tap_coords(v, (int)fx0, (int)fy0, (int)fz0, xa, xb, ya, yb, za, zb);
    const uint32_t ra = ((uint32_t)za * v.H + ya) * v.W, rb = ((uint32_t)za * v.H + yb) * v.W;
    const uint32_t rc = ((uint32_t)zb * v.H + ya) * v.W, rd = ((uint32_t)zb * v.H + yb) * v.W;
    t.i000 = ra + xa; t.i100 = ra + xb; t.i010 = rb + xa; t.i110 = rb + xb;
    t.i001 = rc + xa; t.i101 = rc + xb; t.i011 = rd + xa; t.i111 = rd + xb;
    return t;
}

__device__ __forceinline__ float trilerp(float c000, float c100, float c010, float c110, float c001, float c101,
                                         float c011, float c111, float fx, float fy, float fz) {
    const float a = lerp1(c000, c100, fx), b = lerp1(c010, c110, fx);
    const float e = lerp1(c001, c101, fx), f = lerp1(c011, c111, fx);
    return lerp1(lerp1(a, b, fy), lerp1(e, f, fy), fz);
}

__device__ __forceinline__ float ldx(const float4* t, uint32_t i) { return __ldg(reinterpret_cast<const float*>(t + i)); }

// sdfSampleRawInterp / sdfSampleRawNearest (material.frag:27-53): while loading (lod != 1) the
// coordinate is first rounded to the lod lattice (:33-34); the fetch then uses whatever GL filter
// the texture currently has (NEAREST until the first commit at lod == 1, LINEAR afterwards).
template <bool SNAP>
__device__ __forceinline__ void tex_coord(const TraceParams& P, const Vol& v, float px, float py, float pz, float& ax,
                                          float& ay, float& az) {
    // (p - min) / (max - min), :30, :44.  When the size is a power of two the host passes its exact
    // reciprocal and the division becomes a multiplication with the same bits.
    if (P.size_pow2) {
        ax = (px - P.bmin[0]) * P.inv_size[0]; ay = (py - P.bmin[1]) * P.inv_size[1]; az = (pz - P.bmin[2]) * P.inv_size[2];
    } else {
        ax = (px - P.bmin[0]) / (P.bmax[0] - P.bmin[0]);
        ay = (py - P.bmin[1]) / (P.bmax[1] - P.bmin[1]);
        az = (pz - P.bmin[2]) / (P.bmax[2] - P.bmin[2]);
    }
    if (SNAP) {
        const float rx = (float)v.W / P.lod, ry = (float)v.H / P.lod, rz = (float)v.D / P.lod;  // :33
        ax = floorf(ax * rx + 0.5f) / rx; ay = floorf(ay * ry + 0.5f) / ry; az = floorf(az * rz + 0.5f) / rz;  // :34
    }
}

// The distance lanes of the 8 texels around the last LINEAR fetch.  Near the surface a ray takes
// many tiny steps inside one cell; those steps reuse the registers instead of re-fetching
// (same values, same arithmetic: only the loads are skipped).
struct CellCache {
    float x0, y0, z0;  // floor() of the texel coordinates of the cached cell
    float c000, c100, c010, c110, c001, c101, c011, c111;
};

// point fetch of texel (x, y, z) of the R32F 3-D array: unnormalised coordinates, texel centre
__device__ __forceinline__ float tex_point(cudaTextureObject_t t, int x, int y, int z) {
    return tex3D<float>(t, (float)x + 0.5f, (float)y + 0.5f, (float)z + 0.5f);
}

template <bool SNAP, int DIST>
__device__ __forceinline__ float sample_dist_cached(const TraceParams& P, const Vol& v, float px, float py, float pz,
                                                    CellCache& cc) {
    float ax, ay, az;
    tex_coord<SNAP>(P, v, px, py, pz, ax, ay, az);
    if (DIST == 3) {
        // hardware trilinear: with unnormalised coordinates the unit fetches around x - 0.5, which is GL's
        // u * N - 0.5; clamp addressing equals MIRRORED_REPEAT for every position inside the box
        return tex3D<float>((cudaTextureObject_t)P.dist_tex, ax * (float)v.W, ay * (float)v.H,
                            az * (float)v.D - (float)v.z_lo);
    }
    const float ux = ax * (float)v.W - 0.5f, uy = ay * (float)v.H - 0.5f, uz = az * (float)v.D - 0.5f;
    const float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
    if (fx0 != cc.x0 || fy0 != cc.y0 || fz0 != cc.z0) {
        if (DIST == 2) {  // the same 8 values through the texture unit (point mode, block-linear array)
            int xa, xb, ya, yb, za, zb;
            tap_coords(v, (int)fx0, (int)fy0, (int)fz0, xa, xb, ya, yb, za, zb);
            const cudaTextureObject_t t = (cudaTextureObject_t)P.dist_tex;
            cc.c000 = tex_point(t, xa, ya, za); cc.c100 = tex_point(t, xb, ya, za);
            cc.c010 = tex_point(t, xa, yb, za); cc.c110 = tex_point(t, xb, yb, za);
            cc.c001 = tex_point(t, xa, ya, zb); cc.c101 = tex_point(t, xb, ya, zb);
            cc.c011 = tex_point(t, xa, yb, zb); cc.c111 = tex_point(t, xb, yb, zb);
            cc.x0 = fx0; cc.y0 = fy0; cc.z0 = fz0;
            return trilerp(cc.c000, cc.c100, cc.c010, cc.c110, cc.c001, cc.c101, cc.c011, cc.c111, ux - fx0, uy - fy0,
                           uz - fz0);
        }
        const Taps t = linear_taps(v, ax, ay, az);
        if (DIST == 1) {  // optional distance-only copy of tex0.r: 4 B per voxel instead of 16, same values
            cc.c000 = __ldg(P.dist + t.i000); cc.c100 = __ldg(P.dist + t.i100); cc.c010 = __ldg(P.dist + t.i010);
            cc.c110 = __ldg(P.dist + t.i110); cc.c001 = __ldg(P.dist + t.i001); cc.c101 = __ldg(P.dist + t.i101);
            cc.c011 = __ldg(P.dist + t.i011); cc.c111 = __ldg(P.dist + t.i111);
        } else {
            cc.c000 = ldx(v.tex, t.i000); cc.c100 = ldx(v.tex, t.i100); cc.c010 = ldx(v.tex, t.i010);
            cc.c110 = ldx(v.tex, t.i110); cc.c001 = ldx(v.tex, t.i001); cc.c101 = ldx(v.tex, t.i101);
            cc.c011 = ldx(v.tex, t.i011); cc.c111 = ldx(v.tex, t.i111);
        }
        cc.x0 = fx0; cc.y0 = fy0; cc.z0 = fz0;
    }
    return trilerp(cc.c000, cc.c100, cc.c010, cc.c110, cc.c001, cc.c101, cc.c011, cc.c111, ux - fx0, uy - fy0,
                   uz - fz0);
}

// distance lane only
template <bool SNAP, bool LINEAR>
__device__ __forceinline__ float sample_dist(const TraceParams& P, const Vol& v, float px, float py, float pz) {
    float ax, ay, az;
    tex_coord<SNAP>(P, v, px, py, pz, ax, ay, az);
    if (LINEAR) {
        const Taps t = linear_taps(v, ax, ay, az);
        return trilerp(ldx(v.tex, t.i000), ldx(v.tex, t.i100), ldx(v.tex, t.i010), ldx(v.tex, t.i110),
                       ldx(v.tex, t.i001), ldx(v.tex, t.i101), ldx(v.tex, t.i011), ldx(v.tex, t.i111), t.fx, t.fy,
                       t.fz);
    } else {  // GL_NEAREST: texel floor(p01 * N)
        const int x = (int)floorf(ax * (float)v.W), y = (int)floorf(ay * (float)v.H), z = (int)floorf(az * (float)v.D);
        return ldx(v.tex, texel_index(v, x, y, z));
    }
}

// distance lane from a dense R32F array that covers the Vol `v` (the replicated full-grid distance
// volume of the exact multi-GPU trace): same addressing, same blend as sample_dist
template <bool SNAP, bool LINEAR>
__device__ __forceinline__ float sample_dist_array(const TraceParams& P, const Vol& v, float px, float py, float pz) {
    float ax, ay, az;
    tex_coord<SNAP>(P, v, px, py, pz, ax, ay, az);
    if (LINEAR) {
        const Taps t = linear_taps(v, ax, ay, az);
        return trilerp(__ldg(P.dist + t.i000), __ldg(P.dist + t.i100), __ldg(P.dist + t.i010), __ldg(P.dist + t.i110),
                       __ldg(P.dist + t.i001), __ldg(P.dist + t.i101), __ldg(P.dist + t.i011), __ldg(P.dist + t.i111),
                       t.fx, t.fy, t.fz);
    } else {
        const int x = (int)floorf(ax * (float)v.W), y = (int)floorf(ay * (float)v.H), z = (int)floorf(az * (float)v.D);
        return __ldg(P.dist + texel_index(v, x, y, z));
    }
}

// Exact multi-GPU trace: every rank marches every ray through the replicated distance volume, and the
// rank that OWNS the hit shades it from its slab.  The owner is the rank whose own slices hold the
// lower z tap of the hit's fetch (its upper tap is then an own or a halo slice), so exactly one rank
// shades each hit and it reads the same texels a single GPU would.
template <bool SNAP, bool LINEAR>
__device__ __forceinline__ bool owns_hit(const TraceParams& P, const Vol& vfull, float px, float py, float pz) {
    float ax, ay, az;
    tex_coord<SNAP>(P, vfull, px, py, pz, ax, ay, az);
    int z;
    if (LINEAR) {
        int zb;
        mirror_pair((int)floorf(az * (float)vfull.D - 0.5f), vfull.D, z, zb);
    } else {
        z = mirror_idx((int)floorf(az * (float)vfull.D), vfull.D);
    }
    return (uint32_t)z >= P.own_z0 && (uint32_t)z < P.own_z1;
}

template <bool SNAP, bool LINEAR>
__device__ __forceinline__ float4 sample_full(const TraceParams& P, const Vol& v, float px, float py, float pz) {
    float ax, ay, az;
    tex_coord<SNAP>(P, v, px, py, pz, ax, ay, az);
    if (LINEAR) {
        const Taps t = linear_taps(v, ax, ay, az);
        const float4 c000 = __ldg(v.tex + t.i000), c100 = __ldg(v.tex + t.i100), c010 = __ldg(v.tex + t.i010),
                     c110 = __ldg(v.tex + t.i110), c001 = __ldg(v.tex + t.i001), c101 = __ldg(v.tex + t.i101),
                     c011 = __ldg(v.tex + t.i011), c111 = __ldg(v.tex + t.i111);
        float4 r;
        r.x = trilerp(c000.x, c100.x, c010.x, c110.x, c001.x, c101.x, c011.x, c111.x, t.fx, t.fy, t.fz);
        r.y = trilerp(c000.y, c100.y, c010.y, c110.y, c001.y, c101.y, c011.y, c111.y, t.fx, t.fy, t.fz);
        r.z = trilerp(c000.z, c100.z, c010.z, c110.z, c001.z, c101.z, c011.z, c111.z, t.fx, t.fy, t.fz);
        r.w = trilerp(c000.w, c100.w, c010.w, c110.w, c001.w, c101.w, c011.w, c111.w, t.fx, t.fy, t.fz);
        return r;
    } else {
        const int x = (int)floorf(ax * (float)v.W), y = (int)floorf(ay * (float)v.H), z = (int)floorf(az * (float)v.D);
        return __ldg(v.tex + texel_index(v, x, y, z));
    }
}

// sdfOutOfBoundsDist, material.frag:83-88
__device__ __forceinline__ float oob_dist(const float* bmin, const float* bmax, float x, float y, float z) {
    const float ox = fmaxf(bmin[0] - x, x - bmax[0]);
    const float oy = fmaxf(bmin[1] - y, y - bmax[1]);
    const float oz = fmaxf(bmin[2] - z, z - bmax[2]);
    return fmaxf(ox, fmaxf(oy, oz));
}

// three-d ToneMapping / ColorMapping shader chunks (material.rs:37-38); see the oracle header
__device__ __forceinline__ float tone_mapping(uint32_t type, float c) {
    if (type == 1u) {
        c = c / (c + 1.0f);
    } else if (type == 2u) {
        c = c * (2.51f * c + 0.03f) / (c * (2.43f * c + 0.59f) + 0.14f);
    } else if (type == 3u) {
        c = fmaxf(0.0f, c - 0.004f);
        c = (c * (6.2f * c + 0.5f)) / (c * (6.2f * c + 1.7f) + 0.06f);
        c = powf(c, 2.2f);
    } else {
        return c;
    }
    return fminf(fmaxf(c, 0.0f), 1.0f);
}
__device__ __forceinline__ float color_mapping(uint32_t type, float c) {
    if (type == 1u) {
        const float lo = c * 12.92f;
        const float hi = 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
        return (c < 0.0031308f) ? lo : hi;
    }
    return c;
}

__device__ __forceinline__ unsigned long long pack_key(float depth, float r, float g, float b, float a) {
    const uint32_t R = min(__float2uint_rn(fminf(fmaxf(r, 0.f), 1.f) * 255.0f), 255u);
    const uint32_t G = min(__float2uint_rn(fminf(fmaxf(g, 0.f), 1.f) * 255.0f), 255u);
    const uint32_t B = min(__float2uint_rn(fminf(fmaxf(b, 0.f), 1.f) * 255.0f), 255u);
    const uint32_t A = min(__float2uint_rn(fminf(fmaxf(a, 0.f), 1.f) * 255.0f), 255u);
    const uint32_t rgba = R | (G << 8) | (B << 16) | (A << 24);
    const float dc = fminf(fmaxf(depth, 0.0f), 1.0f);
    return ((unsigned long long)__float_as_uint(dc) << 32) | rgba;
}

// A pixel the host proved to lie outside the screen rectangle of the projected clip box: its ray
// misses the box (code -3), no ray arithmetic needed.
__device__ __forceinline__ void write_outside(const TraceParams& P, size_t px) {
    if (P.rgba) P.rgba[px] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P.depth) P.depth[px] = 1.0f;
    if (P.gbuf) {
        float4* gp = reinterpret_cast<float4*>(P.gbuf + px * SDFGPU_GBUF_FLOATS);
        gp[0] = make_float4(0.f, 0.f, 0.f, -3.0f);
        gp[1] = gp[2] = gp[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (P.keys) P.keys[px] = pack_key(1.0f, 0.f, 0.f, 0.f, 0.f);
    if (P.rgba8) P.rgba8[px] = 0u;
}

template <bool SNAP, bool LINEAR, int DIST = 0, bool FULL = false>
__device__ __forceinline__ int trace_pixel(const TraceParams& P, uint32_t i, uint32_t j) {  // returns the ray's steps
    const size_t px = (size_t)j * P.width + i;

    const Vol v0{P.tex0, (int)P.W, (int)P.H, (int)P.D, (int)P.z_lo, (int)P.z_hi};
    const Vol v1{P.tex1, (int)P.W, (int)P.H, (int)P.D, (int)P.z_lo, (int)P.z_hi};
    const Vol vfull{nullptr, (int)P.W, (int)P.H, (int)P.D, 0, (int)P.D};  // FULL: P.dist covers the whole grid

    float g[SDFGPU_GBUF_FLOATS];
#pragma unroll
    for (int k = 0; k < SDFGPU_GBUF_FLOATS; ++k) g[k] = 0.0f;
    float code;
    float hx = 0.f, hy = 0.f, hz = 0.f;
    int steps = 0;
    float s0x = 0.f;
    bool hit = false;
    float rdx = 0.f, rdy = 0.f, rdz = 0.f;

    const float fx = (float)i + 0.5f, fy = (float)j + 0.5f;
    const float camx = P.origin[0], camy = P.origin[1], camz = P.origin[2];
    const float dux = (P.base[0] + P.dx[0] * fx) + P.dy[0] * fy;
    const float duy = (P.base[1] + P.dx[1] * fx) + P.dy[1] * fy;
    const float duz = (P.base[2] + P.dx[2] * fx) + P.dy[2] * fy;
    // fragment coverage: ray / AABB slab test
    const float ix = 1.0f / dux, iy = 1.0f / duy, iz = 1.0f / duz;
    const float t1x = (P.clip_min[0] - camx) * ix, t2x = (P.clip_max[0] - camx) * ix;
    const float t1y = (P.clip_min[1] - camy) * iy, t2y = (P.clip_max[1] - camy) * iy;
    const float t1z = (P.clip_min[2] - camz) * iz, t2z = (P.clip_max[2] - camz) * iz;
    const float tmin = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fminf(t1z, t2z));
    const float tmax = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fmaxf(t1z, t2z));
    if (!(tmax >= fmaxf(tmin, 0.0f))) {
        code = -3.0f;
    } else {
        const float tf = (tmin < 0.0f) ? tmax : tmin;
        const float posx = camx + dux * tf, posy = camy + duy * tf, posz = camz + duz * tf;
        // material.frag:133-139
        rdx = posx - camx; rdy = posy - camy; rdz = posz - camz;
        const float inv = 1.0f / sqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
        rdx *= inv; rdy *= inv; rdz *= inv;
        float rox = posx, roy = posy, roz = posz;
        if (oob_dist(P.clip_min, P.clip_max, rox + rdx * 0.2f, roy + rdy * 0.2f, roz + rdz * 0.2f) > 0.0f) {
            rox = camx + rdx * 0.2f; roy = camy + rdy * 0.2f; roz = camz + rdz * 0.2f;
        }
        // sdfRaycast, material.frag:92-128, maxSteps = 256 (:142)
        float t = 0.0f;
        hx = rox; hy = roy; hz = roz;
        code = -1.0f;
        const int max_steps = (int)P.max_steps;  // 256 in the reference (material.frag:142)
        CellCache cell;
        cell.x0 = cell.y0 = cell.z0 = __int_as_float(0x7fc00000);  // NaN: never equal, first step always fetches
        cell.c000 = cell.c100 = cell.c010 = cell.c110 = cell.c001 = cell.c101 = cell.c011 = cell.c111 = 0.0f;
        for (int it = 0; it < max_steps; ++it) {
            steps = it;
            if (it >= max_steps - 1) { code = -1.0f; break; }                                          // :99-102
            if (oob_dist(P.clip_min, P.clip_max, hx, hy, hz) > 1e-4f) { code = -2.0f; break; }  // :106-109
            if (FULL) s0x = LINEAR ? sample_dist_cached<SNAP, 1>(P, vfull, hx, hy, hz, cell)       // :112
                                   : sample_dist_array<SNAP, LINEAR>(P, vfull, hx, hy, hz);
            else if (LINEAR) s0x = sample_dist_cached<SNAP, DIST>(P, v0, hx, hy, hz, cell);
            else s0x = sample_dist<SNAP, LINEAR>(P, v0, hx, hy, hz);
            const float dist = s0x - 1e-1f;                                                  // :59
            if (dist < 1e-5f) { code = t; hit = true; break; }                               // :117-121
            t += dist;                                                                       // :124
            hx += rdx * dist; hy += rdy * dist; hz += rdz * dist;                            // :125
        }
    }

    if (FULL && hit && !owns_hit<SNAP, LINEAR>(P, vfull, hx, hy, hz)) {
        hit = false;    // another rank shades this pixel; here it looks like a miss to the MIN composite
        code = -4.0f;
    }
    g[0] = hx; g[1] = hy; g[2] = hz; g[3] = code; g[15] = (float)steps;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    float depth = 1.0f;
    if (hit) {
        const float4 s0 = sample_full<SNAP, LINEAR>(P, v0, hx, hy, hz);
        const float4 s1 = sample_full<SNAP, LINEAR>(P, v1, hx, hy, hz);  // :154
        g[4] = s0.x; g[5] = s0.y; g[6] = s0.z; g[7] = s0.w;
        g[8] = s1.x; g[9] = s1.y; g[10] = s1.z; g[11] = s1.w;
        if (P.gbuf) {  // sdfNormal, :73-80 (dead for the ambient-only light set, kept for the G-buffer)
            const float lx = (float)P.W / P.lod, ly = (float)P.H / P.lod, lz = (float)P.D / P.lod;
            const float h = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
            const float kx[4] = {1.f, -1.f, -1.f, 1.f}, ky[4] = {-1.f, -1.f, 1.f, 1.f}, kz[4] = {-1.f, 1.f, -1.f, 1.f};
            float nx = 0.f, ny = 0.f, nz = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float pxq = hx + kx[q] * h, pyq = hy + ky[q] * h, pzq = hz + kz[q] * h;
                const float dq = (FULL ? sample_dist_array<SNAP, LINEAR>(P, vfull, pxq, pyq, pzq)
                                       : sample_dist<SNAP, LINEAR>(P, v0, pxq, pyq, pzq)) - 1e-1f;
                nx += kx[q] * dq; ny += ky[q] * dq; nz += kz[q] * dq;
            }
            const float ninv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
            g[12] = nx * ninv; g[13] = ny * ninv; g[14] = nz * ninv;
        }
        // :158-173
        const float al[3] = {s0.y * P.tint[0], s0.z * P.tint[1], s0.w * P.tint[2]};
        float col[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // calculate_lighting with one ambient light: occlusion * ambient * mix(albedo, 0, metallic)
            const float mixv = al[c] * (1.0f - s1.x) + 0.0f * s1.x;
            float x = 0.0f + (s1.z * P.ambient[c]) * mixv;
            x = tone_mapping(P.tone_mapping, x);
            x = color_mapping(P.color_mapping, x);
            if (P.gamma != 0.0f) x = powf(x, P.gamma);
            col[c] = x;
        }
        out = make_float4(col[0], col[1], col[2], P.tint[3]);
        // :180-181
        const float* m = P.bvp;
        const float z = ((m[2] * hx + m[6] * hy) + m[10] * hz) + m[14];
        const float w = ((m[3] * hx + m[7] * hy) + m[11] * hz) + m[15];
        depth = z / w;
    }
    if (P.rgba) P.rgba[px] = out;
    if (P.depth) P.depth[px] = depth;
    if (P.gbuf) {
        float4* gp = reinterpret_cast<float4*>(P.gbuf + px * SDFGPU_GBUF_FLOATS);
        gp[0] = make_float4(g[0], g[1], g[2], g[3]);
        gp[1] = make_float4(g[4], g[5], g[6], g[7]);
        gp[2] = make_float4(g[8], g[9], g[10], g[11]);
        gp[3] = make_float4(g[12], g[13], g[14], g[15]);
    }
    if (P.keys) P.keys[px] = pack_key(depth, out.x, out.y, out.z, out.w);
    if (P.rgba8) P.rgba8[px] = (uint32_t)(pack_key(depth, out.x, out.y, out.z, out.w) & 0xffffffffull);
    return steps;
}

// Variant 0 (default): 1-D grid of 8 x 8 pixel tiles (2 warps of 8 x 4), ordered so that the tiles
// inside the screen rectangle of the projected clip box come first: the long marches start at
// once and the cheap outside tiles fill in behind them.  Small CTAs release their SM slot as soon
// as their own rays end instead of waiting for the slowest of 8 warps.
template <bool SNAP, bool LINEAR, int DIST = 0, bool FULL = false>
__global__ void __launch_bounds__(64) trace_tiles_kernel(const __grid_constant__ TraceParams P) {
    // the CTAs come band by band in the order of band_order -- a band is the tile rows [ty0, ty1) -- and within a band
    // first one CTA per tile inside the rectangle (clipped to the band), then one per run of TRACE_OUTSIDE_RUN tiles
    // outside it (no ray, stores only: a CTA per tile costs more to schedule than its 64 pixels take to write)
    uint32_t b = blockIdx.x, band = 0, ty0 = 0, ty1 = P.tiles_y, ry0, ry1, rw, n_heavy, n_out, n_ctas, heavy_off = 0;
    for (uint32_t k = 0;; ++k) {
        if (P.n_bands) {
            band = P.band_order[k];
            ty0 = band * P.band_rows; ty1 = min(ty0 + P.band_rows, P.tiles_y);
        }
        ry0 = min(max(P.rect[1], ty0), ty1); ry1 = min(max(P.rect[3], ty0), ty1);
        rw = P.rect[2] - P.rect[0];
        n_heavy = rw * (ry1 - ry0);
        n_out = P.tiles_x * (ty1 - ty0) - n_heavy;
        n_ctas = n_heavy + (n_out + TRACE_OUTSIDE_RUN - 1u) / TRACE_OUTSIDE_RUN;
        if (b < n_ctas || k + 1u >= P.n_bands) break;
        b -= n_ctas;
        heavy_off += n_heavy;
    }
    const uint32_t rh = ry1 - ry0;
    if (b < n_heavy) {
        // the tiles inside the rectangle: row by row, or -- tile_order -- the ones whose marches were longest in the
        // previous frame first (tile_order_kernel), so that the frame does not end on a long march that started late
        uint32_t tx = P.rect[0] + b % rw, ty = ry0 + b / rw;
        if (P.tile_order) {
            const uint32_t id = P.tile_order[heavy_off + b];
            tx = id % P.tiles_x; ty = id / P.tiles_x;
        }
        const uint32_t i = tx * 8u + (threadIdx.x & 7u), j = ty * 8u + (threadIdx.x >> 3);
        int steps = 0;
        if (i < P.width && j < P.height) steps = trace_pixel<SNAP, LINEAR, DIST, FULL>(P, i, j);
        if (P.tile_cost) {  // this frame's cost of the tile: its longest march
            steps = __reduce_max_sync(0xffffffffu, steps);
            if ((threadIdx.x & 31u) == 0u) atomicMax(P.tile_cost + ty * P.tiles_x + tx, (uint32_t)steps);
        }
    } else {
        const uint32_t o0 = (b - n_heavy) * TRACE_OUTSIDE_RUN, o1 = min(o0 + TRACE_OUTSIDE_RUN, n_out);
        const uint32_t n_top = (ry0 - ty0) * P.tiles_x, side = P.tiles_x - rw;
        for (uint32_t o = o0; o < o1; ++o) {
            uint32_t q = o, tx, ty;
            if (q < n_top) {
                tx = q % P.tiles_x; ty = ty0 + q / P.tiles_x;
            } else if (q - n_top < rh * side) {
                q -= n_top;
                const uint32_t k = q % side;
                ty = ry0 + q / side; tx = k < P.rect[0] ? k : k + rw;
            } else {
                q -= n_top + rh * side;
                tx = q % P.tiles_x; ty = ry1 + q / P.tiles_x;
            }
            const uint32_t i = tx * 8u + (threadIdx.x & 7u), j = ty * 8u + (threadIdx.x >> 3);
            if (i < P.width && j < P.height) write_outside(P, (size_t)j * P.width + i);
        }
    }
    if (P.n_bands) {  // the last CTA of a band of tile rows to finish: every pixel of those rows is in memory
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(P.band_done + band, 1u) + 1u == n_ctas) {
                P.band_done[band] = 0u;
                __threadfence();
                *reinterpret_cast<volatile uint32_t*>(P.band_flags + band) = P.band_epoch;
            }
        }
    }
}

// Variant 1: plain 2-D grid, 8 warps per CTA, each an 8 x 4 pixel tile; CTA tile = 32 x 8 pixels
template <bool SNAP, bool LINEAR>
__global__ void __launch_bounds__(256) trace_kernel(const __grid_constant__ TraceParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * 32u + (warp & 3) * 8u + (lane & 7);
    const uint32_t j = blockIdx.y * 8u + (warp >> 2) * 4u + (lane >> 3);
    if (i >= P.width || j >= P.height) return;
    trace_pixel<SNAP, LINEAR>(P, i, j);
}

// ---------------------------------------------------------------------------------------------------------
// Round kernel: persistent warps that pull their work from a queue.  One kernel serves the single-volume
// trace (linked == 0: one round, every ray runs to its end) and the Z-sharded trace (LinkParams,
// sdfgpu_internal.h): a ray marches while this rank owns the lower z tap of its fetch and is handed to the
// neighbour otherwise, so the step sequence -- and with it hit point, step count and pixel -- is the single
// volume's, bit for bit, with a one-slice halo and no replicated data.
//
// Work units are 32 rays: in the first round an 8 x 4 pixel tile (tiles inside the screen rectangle of the
// projected box first: those hold every long march), later 32 consecutive entries of a neighbour's out-queue.
// A warp takes the next unit with one atomicAdd as soon as ITS rays are done -- a 255-step grazing ray holds
// one warp slot, not a CTA -- and leaves the march loop on an explicit warp ballot.

struct Ray {
    float hx, hy, hz, t;
    int it;
};

// ray through pixel (i, j): the fragment the rasterised bounding cube would produce (material.frag:133-139).
// false: the ray does not enter the clip box (no fragment).
__device__ __forceinline__ bool ray_setup(const TraceParams& P, uint32_t i, uint32_t j, float& rdx, float& rdy, float& rdz,
                                          float& rox, float& roy, float& roz) {
    const float fx = (float)i + 0.5f, fy = (float)j + 0.5f;
    const float camx = P.origin[0], camy = P.origin[1], camz = P.origin[2];
    const float dux = (P.base[0] + P.dx[0] * fx) + P.dy[0] * fy;
    const float duy = (P.base[1] + P.dx[1] * fx) + P.dy[1] * fy;
    const float duz = (P.base[2] + P.dx[2] * fx) + P.dy[2] * fy;
    const float ix = 1.0f / dux, iy = 1.0f / duy, iz = 1.0f / duz;
    const float t1x = (P.clip_min[0] - camx) * ix, t2x = (P.clip_max[0] - camx) * ix;
    const float t1y = (P.clip_min[1] - camy) * iy, t2y = (P.clip_max[1] - camy) * iy;
    const float t1z = (P.clip_min[2] - camz) * iz, t2z = (P.clip_max[2] - camz) * iz;
    const float tmin = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fminf(t1z, t2z));
    const float tmax = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fmaxf(t1z, t2z));
    rdx = rdy = rdz = 0.0f; rox = roy = roz = 0.0f;
    if (!(tmax >= fmaxf(tmin, 0.0f))) return false;
    const float tf = (tmin < 0.0f) ? tmax : tmin;
    const float posx = camx + dux * tf, posy = camy + duy * tf, posz = camz + duz * tf;
    rdx = posx - camx; rdy = posy - camy; rdz = posz - camz;
    const float inv = 1.0f / sqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
    rdx *= inv; rdy *= inv; rdz *= inv;
    rox = posx; roy = posy; roz = posz;
    if (oob_dist(P.clip_min, P.clip_max, rox + rdx * 0.2f, roy + rdy * 0.2f, roz + rdz * 0.2f) > 0.0f) {
        rox = camx + rdx * 0.2f; roy = camy + rdy * 0.2f; roz = camz + rdz * 0.2f;
    }
    return true;
}

// the lower z tap of the fetch at normalised coordinate az (same arithmetic as linear_taps / the NEAREST fetch)
template <bool LINEAR>
__device__ __forceinline__ int lower_z_tap(int D, float az) {
    if (LINEAR) {
        int a, b;
        mirror_pair((int)floorf(az * (float)D - 0.5f), D, a, b);
        return a;
    }
    return mirror_idx((int)floorf(az * (float)D), D);
}

// distance lane at normalised coordinates (what sample_dist_cached / sample_dist compute after tex_coord)
template <bool LINEAR>
__device__ __forceinline__ float dist_at(const Vol& v, float ax, float ay, float az, CellCache& cc) {
    if (LINEAR) {
        const float ux = ax * (float)v.W - 0.5f, uy = ay * (float)v.H - 0.5f, uz = az * (float)v.D - 0.5f;
        const float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
        if (fx0 != cc.x0 || fy0 != cc.y0 || fz0 != cc.z0) {
            const Taps t = linear_taps(v, ax, ay, az);
            cc.c000 = ldx(v.tex, t.i000); cc.c100 = ldx(v.tex, t.i100); cc.c010 = ldx(v.tex, t.i010);
            cc.c110 = ldx(v.tex, t.i110); cc.c001 = ldx(v.tex, t.i001); cc.c101 = ldx(v.tex, t.i101);
            cc.c011 = ldx(v.tex, t.i011); cc.c111 = ldx(v.tex, t.i111);
            cc.x0 = fx0; cc.y0 = fy0; cc.z0 = fz0;
        }
        return trilerp(cc.c000, cc.c100, cc.c010, cc.c110, cc.c001, cc.c101, cc.c011, cc.c111, ux - fx0, uy - fy0,
                       uz - fz0);
    }
    const int x = (int)floorf(ax * (float)v.W), y = (int)floorf(ay * (float)v.H), z = (int)floorf(az * (float)v.D);
    return ldx(v.tex, texel_index(v, x, y, z));
}

enum : int { RS_HIT = 0, RS_MISS = 1, RS_DOWN = 2, RS_UP = 3, RS_NONE = 4 };

// shading of a finished ray (material.frag:145-181) and its outputs.  Sharded: the (depth, RGBA8) key and the optional
// G-buffer record go to the presenter's frame; single volume: the buffers of TraceParams.
template <bool SNAP, bool LINEAR>
__device__ __forceinline__ void finish_pixel(const TraceParams& P, const LinkParams& L, const Vol& v0, const Vol& v1,
                                             uint32_t px, bool hit, float code, const Ray& r) {
    float g[SDFGPU_GBUF_FLOATS];
#pragma unroll
    for (int k = 0; k < SDFGPU_GBUF_FLOATS; ++k) g[k] = 0.0f;
    const float hx = r.hx, hy = r.hy, hz = r.hz;
    g[0] = hx; g[1] = hy; g[2] = hz; g[3] = code; g[15] = (float)r.it;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    float depth = 1.0f;
    float* const gbuf = L.linked ? L.frame_gbuf : P.gbuf;
    if (hit) {
        const float4 s0 = sample_full<SNAP, LINEAR>(P, v0, hx, hy, hz);
        const float4 s1 = sample_full<SNAP, LINEAR>(P, v1, hx, hy, hz);  // :154
        g[4] = s0.x; g[5] = s0.y; g[6] = s0.z; g[7] = s0.w;
        g[8] = s1.x; g[9] = s1.y; g[10] = s1.z; g[11] = s1.w;
        if (gbuf) {  // sdfNormal, :73-80 (dead for the ambient-only light set, kept for the G-buffer)
            const float lx = (float)P.W / P.lod, ly = (float)P.H / P.lod, lz = (float)P.D / P.lod;
            const float h = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
            const float kx[4] = {1.f, -1.f, -1.f, 1.f}, ky[4] = {-1.f, -1.f, 1.f, 1.f}, kz[4] = {-1.f, 1.f, -1.f, 1.f};
            float nx = 0.f, ny = 0.f, nz = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float dq = sample_dist<SNAP, LINEAR>(P, v0, hx + kx[q] * h, hy + ky[q] * h, hz + kz[q] * h) - 1e-1f;
                nx += kx[q] * dq; ny += ky[q] * dq; nz += kz[q] * dq;
            }
            const float ninv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
            g[12] = nx * ninv; g[13] = ny * ninv; g[14] = nz * ninv;
        }
        const float al[3] = {s0.y * P.tint[0], s0.z * P.tint[1], s0.w * P.tint[2]};  // :158-173
        float col[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mixv = al[c] * (1.0f - s1.x) + 0.0f * s1.x;
            float x = 0.0f + (s1.z * P.ambient[c]) * mixv;
            x = tone_mapping(P.tone_mapping, x);
            x = color_mapping(P.color_mapping, x);
            if (P.gamma != 0.0f) x = powf(x, P.gamma);
            col[c] = x;
        }
        out = make_float4(col[0], col[1], col[2], P.tint[3]);
        const float* m = P.bvp;  // :180-181
        const float z = ((m[2] * hx + m[6] * hy) + m[10] * hz) + m[14];
        const float w = ((m[3] * hx + m[7] * hy) + m[11] * hz) + m[15];
        depth = z / w;
    }
    if (gbuf) {
        float4* gp = reinterpret_cast<float4*>(gbuf + (size_t)px * SDFGPU_GBUF_FLOATS);
        gp[0] = make_float4(g[0], g[1], g[2], g[3]);
        gp[1] = make_float4(g[4], g[5], g[6], g[7]);
        gp[2] = make_float4(g[8], g[9], g[10], g[11]);
        gp[3] = make_float4(g[12], g[13], g[14], g[15]);
    }
    if (L.linked) {
        L.frame_keys[px] = pack_key(depth, out.x, out.y, out.z, out.w);
        return;
    }
    if (P.rgba) P.rgba[px] = out;
    if (P.depth) P.depth[px] = depth;
    if (P.keys) P.keys[px] = pack_key(depth, out.x, out.y, out.z, out.w);
    if (P.rgba8) P.rgba8[px] = (uint32_t)(pack_key(depth, out.x, out.y, out.z, out.w) & 0xffffffffull);
}

template <bool SNAP, bool LINEAR>
__global__ void __launch_bounds__(256) trace_rounds_kernel(const __grid_constant__ TraceParams P,
                                                           const __grid_constant__ LinkParams L) {
    const uint32_t lane = threadIdx.x & 31u;
    const Vol v0{P.tex0, (int)P.W, (int)P.H, (int)P.D, (int)P.z_lo, (int)P.z_hi};
    const Vol v1{P.tex1, (int)P.W, (int)P.H, (int)P.D, (int)P.z_lo, (int)P.z_hi};

    // ---- the work of this round.  First round: one unit per 8 x 4 tile inside the screen rectangle of the box (all
    // the marching is there), then units of OUTSIDE_RUN tiles outside it (stores only: one queue access per 512 pixels)
    constexpr uint32_t OUTSIDE_RUN = 16u;
    const uint32_t tiles_y4 = (P.height + 3u) / 4u;                              // rows of 8 x 4 tiles
    const uint32_t rw = P.rect[2] - P.rect[0], ry0 = min(P.rect[1] * 2u, tiles_y4);
    const uint32_t rh4 = min(P.rect[3] * 2u, tiles_y4) - ry0;
    const uint32_t n_heavy = rw * rh4;
    const uint32_t n_outside = P.tiles_x * tiles_y4 - n_heavy;
    uint32_t n_units, units0 = 0, cnt0 = 0, cnt1 = 0;
    if (L.first) {
        // tiles outside the rectangle hold no ray that enters the box: only the presenter has to write them
        n_units = n_heavy + ((!L.linked || L.is_presenter) ? (n_outside + OUTSIDE_RUN - 1u) / OUTSIDE_RUN : 0u);
    } else {
        if (L.in_count[0]) cnt0 = *reinterpret_cast<const volatile uint32_t*>(L.in_count[0]);
        if (L.in_count[1]) cnt1 = *reinterpret_cast<const volatile uint32_t*>(L.in_count[1]);
        units0 = (cnt0 + 31u) / 32u;
        n_units = units0 + (cnt1 + 31u) / 32u;
    }

    // every warp's first unit is its own index -- a round with little or no work costs no queue access at all -- the
    // rest are handed out by the queue
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (;;) {
        if (unit >= n_units) break;

        if (L.first && unit >= n_heavy) {  // a run of tiles no ray of which enters the box: miss code -3, no arithmetic
            const uint32_t b0 = (unit - n_heavy) * OUTSIDE_RUN, b1 = min(b0 + OUTSIDE_RUN, n_outside);
            const uint32_t n_top = ry0 * P.tiles_x, side = P.tiles_x - rw;
            Ray none;
            none.hx = none.hy = none.hz = none.t = 0.0f; none.it = 0;
            for (uint32_t bb = b0; bb < b1; ++bb) {
                uint32_t b = bb, tx, ty;
                if (b < n_top) {
                    tx = b % P.tiles_x; ty = b / P.tiles_x;
                } else if (b - n_top < rh4 * side) {
                    b -= n_top;
                    const uint32_t k = b % side;
                    ty = ry0 + b / side; tx = k < P.rect[0] ? k : k + rw;
                } else {
                    b -= n_top + rh4 * side;
                    tx = b % P.tiles_x; ty = ry0 + rh4 + b / P.tiles_x;
                }
                const uint32_t i = tx * 8u + (lane & 7u), j = ty * 4u + (lane >> 3);
                if (i < P.width && j < P.height) finish_pixel<SNAP, LINEAR>(P, L, v0, v1, j * P.width + i, false, -3.0f, none);
            }
            if (lane == 0) unit = n_warps + atomicAdd(L.work_head, 1u);
            unit = __shfl_sync(0xffffffffu, unit, 0);
            continue;
        }

        // ---- this lane's ray
        Ray r;
        r.hx = r.hy = r.hz = r.t = 0.0f; r.it = 0;
        uint32_t px = 0;
        float rdx = 0.f, rdy = 0.f, rdz = 0.f;
        int status = RS_NONE;   // RS_NONE: nothing (left) to do for this lane
        bool marching = false;
        float code = -3.0f;
        if (L.first) {
            const uint32_t tx = P.rect[0] + unit % rw, ty = ry0 + unit / rw;
            const uint32_t i = tx * 8u + (lane & 7u), j = ty * 4u + (lane >> 3);
            if (i < P.width && j < P.height) {
                px = j * P.width + i;
                float rox, roy, roz;
                if (!ray_setup(P, i, j, rdx, rdy, rdz, rox, roy, roz)) {
                    // no fragment: miss code -3, written by the presenter (every rank sees the same test)
                    if (!L.linked || L.is_presenter) status = RS_MISS;
                } else {
                    r.hx = rox; r.hy = roy; r.hz = roz;
                    marching = true;
                }
            }
        } else {
            const uint32_t q = unit < units0 ? 0u : 1u;
            const uint32_t e = (unit - (q ? units0 : 0u)) * 32u + lane;
            if (e < (q ? cnt1 : cnt0)) {
                const float4 pos = L.in_pos[q][e];
                const uint2 id = L.in_id[q][e];
                r.hx = pos.x; r.hy = pos.y; r.hz = pos.z; r.t = pos.w;
                px = id.x; r.it = (int)id.y;
                float rox, roy, roz;
                (void)ray_setup(P, px % P.width, px / P.width, rdx, rdy, rdz, rox, roy, roz);  // the direction only
                marching = true;
            }
        }

        // ---- sdfRaycast, material.frag:92-128, maxSteps = 256 (:142).  The warp leaves on an empty ballot.
        const int max_steps = (int)P.max_steps;
        CellCache cell;
        cell.x0 = cell.y0 = cell.z0 = __int_as_float(0x7fc00000);  // NaN: never equal, the first step always fetches
        cell.c000 = cell.c100 = cell.c010 = cell.c110 = cell.c001 = cell.c101 = cell.c011 = cell.c111 = 0.0f;
        const bool started_here = L.first != 0u;
        while (__ballot_sync(0xffffffffu, marching)) {
            if (marching) {
                float ax, ay, az;
                tex_coord<SNAP>(P, v0, r.hx, r.hy, r.hz, ax, ay, az);
                // who owns this sample: the rank whose own slices hold the lower z tap.  The two termination tests
                // need no texel, so whichever rank holds the ray applies them first: a position outside the box (whose
                // taps mirror back to slices of some other rank) ends here instead of travelling on, and inside the box
                // the lower tap is monotone along the ray -- at most world - 1 hand-offs
                int go = 0;
                if (L.linked) {
                    const int z = lower_z_tap<LINEAR>((int)P.D, az);
                    go = z < (int)L.own_z0 ? RS_DOWN : (z >= (int)L.own_z1 ? RS_UP : 0);
                }
                if (go && started_here && r.it == 0) {
                    status = RS_NONE; marching = false;   // a ray that STARTS elsewhere is started by its owner
                } else if (r.it >= max_steps - 1) {                                                  // :99-102
                    code = -1.0f; status = RS_MISS; marching = false;
                } else if (oob_dist(P.clip_min, P.clip_max, r.hx, r.hy, r.hz) > 1e-4f) {             // :106-109
                    code = -2.0f; status = RS_MISS; marching = false;
                } else if (go) {
                    status = go; marching = false;
                } else {
                    const float dist = dist_at<LINEAR>(v0, ax, ay, az, cell) - 1e-1f;                // :112, :59
                    if (dist < 1e-5f) {                                                              // :117-121
                        code = r.t; status = RS_HIT; marching = false;
                    } else {
                        r.t += dist;                                                                 // :124
                        r.hx += rdx * dist; r.hy += rdy * dist; r.hz += rdz * dist;                  // :125
                        ++r.it;
                    }
                }
            }
        }

        // ---- finished rays: shade and write; rays that left this rank's slices: append to the out-queue
        if (status == RS_HIT || status == RS_MISS) finish_pixel<SNAP, LINEAR>(P, L, v0, v1, px, status == RS_HIT, code, r);
        if (L.linked) {
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
                const bool go = status == (dir ? RS_UP : RS_DOWN);
                const uint32_t m = __ballot_sync(0xffffffffu, go);
                if (m == 0u) continue;
                uint32_t base = 0;
                if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(L.out_count + dir, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (go) {
                    const uint32_t e = base + __popc(m & ((1u << lane) - 1u));
                    L.out_pos[dir][e] = make_float4(r.hx, r.hy, r.hz, r.t);
                    L.out_id[dir][e] = make_uint2(px, (uint32_t)r.it);
                }
            }
        }
        if (lane == 0) unit = n_warps + atomicAdd(L.work_head, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
    }

    // ---- the last CTA to finish resets the counters and tells the neighbours (and, after the last round, the
    // presenter) that this round's queue entries and pixels are in place
    // out-queue entries and the peer stores of finished pixels: the CTA's writes are ordered before thread 0's fence by
    // the barrier (fences are cumulative), so one system-scope fence per CTA is enough
    __syncthreads();
    if (threadIdx.x == 0) {
        if (L.linked) __threadfence_system();
        if (atomicAdd(L.ctas_done, 1u) + 1u == gridDim.x) {
            *L.work_head = 0u;
            *L.ctas_done = 0u;
            if (L.linked) {
                // every CTA has finished: the reservation counters are final -- publish them in the neighbours' arenas
                if (L.out_publish[0]) *reinterpret_cast<volatile uint32_t*>(L.out_publish[0]) = L.out_count[0];
                if (L.out_publish[1]) *reinterpret_cast<volatile uint32_t*>(L.out_publish[1]) = L.out_count[1];
                L.reset_count[0] = 0u; L.reset_count[1] = 0u;
                __threadfence_system();
                if (L.sig_round[0]) *reinterpret_cast<volatile uint32_t*>(L.sig_round[0]) = L.sig_round_value;
                if (L.sig_round[1]) *reinterpret_cast<volatile uint32_t*>(L.sig_round[1]) = L.sig_round_value;
                if (L.sig_frame) *reinterpret_cast<volatile uint32_t*>(L.sig_frame) = L.sig_frame_value;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stream kernel: the Z-sharded trace in ONE launch per rank.  Same ownership rule, same arithmetic and the same
// frame as the round kernel, but no round structure: the persistent warps work through this rank's tiles and then
// poll the in-queues, so a ray that a neighbour hands over continues as soon as its entry has arrived -- the ranks
// form a pipeline along z instead of taking turns.  Producer and consumer share no counter: an entry is two 16-byte
// stores, each carrying the frame's tag, and the consumer works on a 32-entry unit when all its tags are there (16-byte
// stores are not torn on NVLink).  A queue is closed by its producer's final count (LinkParams::out_final).

__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_volatile_f4(float4* p, float x, float y, float z, float w) {
    asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// lane 0 of any warp, whenever something completed: close the out-queue of direction `dir` (0 down, 1 up) if nothing
// can enter it any more.  Rays keep the sign of their z direction, so what leaves downwards started here or came
// from above (in-queue 1), and the other way round.
__device__ __forceinline__ void stream_maybe_close(const LinkParams& L, int dir, uint32_t n_tile_units) {
    if (!L.out_final[dir]) return;
    if (ld_volatile_u32(L.sent_final + dir)) return;
    if (ld_volatile_u32(L.tiles_done) != n_tile_units) return;
    const int src = 1 - dir;
    if (L.in_q[src]) {
        const unsigned long long f = ld_volatile_u64(L.in_final[src]);
        if ((uint32_t)(f >> 32) != L.epoch) return;
        if (ld_volatile_u32(L.in_done + src) != ((uint32_t)f + 31u) / 32u) return;
    }
    // (the counters above were incremented after the units' reservations in out_count had returned)
    if (atomicExch(L.sent_final + dir, 1u) != 0u) return;
    const uint32_t c = ld_volatile_u32(L.out_count + dir);
    *reinterpret_cast<volatile unsigned long long*>(L.out_final[dir]) = ((unsigned long long)L.epoch << 32) | c;
}

template <bool SNAP, bool LINEAR>
__global__ void __launch_bounds__(256, 4) trace_stream_kernel(const __grid_constant__ TraceParams P,
                                                           const __grid_constant__ LinkParams L) {
    const uint32_t lane = threadIdx.x & 31u;
    const Vol v0{P.tex0, (int)P.W, (int)P.H, (int)P.D, (int)P.z_lo, (int)P.z_hi};
    const Vol v1{P.tex1, (int)P.W, (int)P.H, (int)P.D, (int)P.z_lo, (int)P.z_hi};

    // ---- tile units, as in round 0 of the round kernel: 8 x 4 tiles inside the screen rectangle of the box, then
    // (presenter only) runs of OUTSIDE_RUN tiles outside it
    constexpr uint32_t OUTSIDE_RUN = 16u;
    const uint32_t tiles_y4 = (P.height + 3u) / 4u;
    const uint32_t rw = P.rect[2] - P.rect[0], ry0 = min(P.rect[1] * 2u, tiles_y4);
    const uint32_t rh4 = min(P.rect[3] * 2u, tiles_y4) - ry0;
    const uint32_t n_heavy = rw * rh4;
    const uint32_t n_outside = P.tiles_x * tiles_y4 - n_heavy;
    const uint32_t n_out_units = L.is_presenter ? (n_outside + OUTSIDE_RUN - 1u) / OUTSIDE_RUN : 0u;
    const uint32_t n_units = n_heavy + n_out_units;
    // unit numbers: the tiles inside the rectangle, then the runs outside it -- or, with out_rgba8, the other way round
    const uint32_t out_first = L.out_rgba8 ? 0u : n_heavy, heavy_first = L.out_rgba8 ? n_out_units : 0u;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t capacity = (L.max_pixels + 31u) & ~31u;  // entries per queue buffer
    const uint32_t tag = L.epoch;
    constexpr uint32_t NO_TICKET = 0xffffffffu;

    uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // tickets held -- a unit of the in-queue from below / from above -- with the lanes of the unit already traced, and
    // whether the unit has been seen incomplete before
    uint32_t pend0 = NO_TICKET, pend1 = NO_TICKET, mask0 = 0u, mask1 = 0u;
    bool seen0 = false, seen1 = false;
    uint32_t sleep_ns = 32u;
    unsigned long long idle_since = 0ull;

    for (;;) {
        Ray r;
        r.hx = r.hy = r.hz = r.t = 0.0f; r.it = 0;
        uint32_t px = 0;
        float rdx = 0.f, rdy = 0.f, rdz = 0.f;
        int status = RS_NONE;
        bool marching = false;
        float code = -3.0f;
        int src = -1;  // -1: a tile unit; 0 / 1: a unit of the in-queue from below / above
        uint32_t src_valid = 0xffffffffu;  // lanes of that unit that exist (all, unless the queue is closed within it)

        // ---- the in-queues first: a ray that has already travelled is on some other rank's critical path.  Units are
        // handed out by TICKET, one atomicAdd (verifying a unit first and claiming it with a compare-and-swap hands out
        // one unit per memory round trip: measured, 12 ms for a frame whose every ray crosses a slab face).  A warp takes
        // a ticket where the head unit has begun to arrive, or where the queue is closed and not handed out yet.  It
        // holds at most one ticket per queue and looks at both, so a ticket for entries that are still far away keeps
        // it neither from the other queue nor from the tiles: every entry that has arrived belongs to a warp that is
        // polling or busy with a finite unit.  A complete unit is traced at once; an incomplete one the second time
        // the warp finds it so (its stragglers must not wait for the queue to close: they are the frame's tail), the
        // rest of it in a later pass.  A ticket beyond the final count is void.  Lane q looks after queue q.
        bool got = false, open = false;
        const bool tile_next = unit < n_units;
        const bool outside_next = tile_next && unit >= out_first && unit - out_first < n_out_units;
        if (!outside_next) {  // (not between the presenter's store-only runs: they are short)
            uint32_t cnt = 0, fin = 0, tk = lane ? pend1 : pend0, msk = lane ? mask1 : mask0, op = 0u, pre = 0u, valid = 0xffffffffu;
            if (lane < 2u && L.in_q[lane]) {
                const uint32_t q = lane;
                op = 1u;
                const unsigned long long f = ld_volatile_u64(L.in_final[q]);
                fin = (uint32_t)(f >> 32) == tag ? 1u : 0u;
                cnt = (uint32_t)f;
                if (tk == NO_TICKET) {
                    const uint32_t h = ld_volatile_u32(L.in_head + q);
                    bool take;
                    if (fin) take = h * 32u < cnt;
                    else take = h * 32u < capacity && __float_as_uint(ld_volatile_f4(L.in_q[q] + 2u * (size_t)(h * 32u)).w) == tag;
                    if (take) { tk = atomicAdd(L.in_head + q, 1u); msk = 0u; }
                    else if (fin) op = 0u;  // closed and handed out
                }
                if (tk != NO_TICKET && fin) {
                    const uint32_t base = tk * 32u;
                    valid = base >= cnt ? 0u : (cnt - base >= 32u ? 0xffffffffu : (1u << (cnt - base)) - 1u);
                    if (valid == 0u) {  // void: the head is beyond the count
                        tk = NO_TICKET; op = 0u;
                    } else if ((msk & valid) == valid) {  // the queue was closed inside this unit, after its last entry was traced
                        if (atomicAdd(L.in_done + q, 1u) + 1u == (cnt + 31u) / 32u) stream_maybe_close(L, 1 - (int)q, n_units);
                        tk = NO_TICKET; msk = 0u;
                    }
                }
                if (tk != NO_TICKET && tk * 32u < capacity)  // the first entry not traced yet: has it arrived?
                    pre = __float_as_uint(ld_volatile_f4(L.in_q[q] + 2u * (size_t)(tk * 32u + (uint32_t)__ffs((int)~msk) - 1u)).w) == tag ? 1u : 0u;
            }
            pend0 = __shfl_sync(0xffffffffu, tk, 0); pend1 = __shfl_sync(0xffffffffu, tk, 1);
            mask0 = __shfl_sync(0xffffffffu, msk, 0); mask1 = __shfl_sync(0xffffffffu, msk, 1);
            open = __any_sync(0xffffffffu, op != 0u);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (got || !__shfl_sync(0xffffffffu, pre, q)) continue;
                const uint32_t t_q = q ? pend1 : pend0, m_q = q ? mask1 : mask0;
                const uint32_t valid_q = __shfl_sync(0xffffffffu, valid, q);
                const uint32_t e = t_q * 32u + lane;
                const bool want = ((valid_q & ~m_q) >> lane) & 1u;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                bool ready = false;
                if (want) {
                    a = ld_volatile_f4(L.in_q[q] + 2u * (size_t)e);
                    b = ld_volatile_f4(L.in_q[q] + 2u * (size_t)e + 1u);
                    ready = __float_as_uint(a.w) == tag && __float_as_uint(b.w) == tag;
                }
                const uint32_t ready_lanes = __ballot_sync(0xffffffffu, ready);
                const bool complete = __ballot_sync(0xffffffffu, want && !ready) == 0u;
                const bool seen = q ? seen1 : seen0;
                if (ready_lanes == 0u) continue;
                if (!complete && (!seen || (!tile_next && sleep_ns < 512u))) {
                    // give the rest of the unit until the next look (a warp without tiles: a few looks, about 5 us)
                    if (q) seen1 = true; else seen0 = true;
                    continue;
                }
                got = true; src = q; src_valid = valid_q;
                if (q) { mask1 = m_q | ready_lanes; seen1 = false; } else { mask0 = m_q | ready_lanes; seen0 = false; }
                if (ready) {
                    r.hx = a.x; r.hy = a.y; r.hz = a.z; r.t = b.x;
                    px = __float_as_uint(b.y); r.it = (int)__float_as_uint(b.z);
                    float rox, roy, roz;
                    (void)ray_setup(P, px % P.width, px / P.width, rdx, rdy, rdz, rox, roy, roz);  // the direction only
                    marching = true;
                }
            }
        }

        if (got) {
            sleep_ns = 32u; idle_since = 0ull;
        } else if (tile_next) {
            if (outside_next) {  // a run of tiles no ray of which enters the box: miss code -3, no arithmetic
                const uint32_t b0 = (unit - out_first) * OUTSIDE_RUN, b1 = min(b0 + OUTSIDE_RUN, n_outside);
                const uint32_t n_top = ry0 * P.tiles_x, side = P.tiles_x - rw;
                for (uint32_t bb = b0; bb < b1; ++bb) {
                    uint32_t b = bb, tx, ty;
                    if (b < n_top) {
                        tx = b % P.tiles_x; ty = b / P.tiles_x;
                    } else if (b - n_top < rh4 * side) {
                        b -= n_top;
                        const uint32_t k = b % side;
                        ty = ry0 + b / side; tx = k < P.rect[0] ? k : k + rw;
                    } else {
                        b -= n_top + rh4 * side;
                        tx = b % P.tiles_x; ty = ry0 + rh4 + b / P.tiles_x;
                    }
                    const uint32_t i = tx * 8u + (lane & 7u), j = ty * 4u + (lane >> 3);
                    if (i < P.width && j < P.height) {
                        const uint32_t opx = j * P.width + i;
                        // rows of tiles above / below the rectangle go into the frame itself -- depth 1, colour 0, what
                        // the key of a miss unpacks to --; tiles beside it share their rows with rays: key frame
                        if (L.out_rgba8 && (ty < ry0 || ty >= ry0 + rh4)) {
                            L.out_rgba8[opx] = 0u;
                            L.out_depth[opx] = 1.0f;
                            if (L.frame_gbuf) {
                                float4* gp = reinterpret_cast<float4*>(L.frame_gbuf + (size_t)opx * SDFGPU_GBUF_FLOATS);
                                gp[0] = make_float4(0.f, 0.f, 0.f, -3.0f);
                                gp[1] = gp[2] = gp[3] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        } else {
                            finish_pixel<SNAP, LINEAR>(P, L, v0, v1, opx, false, -3.0f, r);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (L.out_rgba8) {
                        __threadfence();  // the run's pixels before the count that may raise the flag
                        if (atomicAdd(L.outside_done, 1u) + 1u == n_out_units)
                            *reinterpret_cast<volatile uint32_t*>(L.outside_flag) = L.epoch;
                    }
                    if (atomicAdd(L.tiles_done, 1u) + 1u == n_units) { stream_maybe_close(L, 0, n_units); stream_maybe_close(L, 1, n_units); }
                    unit = n_warps + atomicAdd(L.work_head, 1u);
                }
                unit = __shfl_sync(0xffffffffu, unit, 0);
                continue;
            }
            const uint32_t hu = unit - heavy_first;
            const uint32_t tx = P.rect[0] + hu % rw, ty = ry0 + hu / rw;
            const uint32_t i = tx * 8u + (lane & 7u), j = ty * 4u + (lane >> 3);
            if (i < P.width && j < P.height) {
                px = j * P.width + i;
                float rox, roy, roz;
                if (!ray_setup(P, i, j, rdx, rdy, rdz, rox, roy, roz)) {
                    if (L.is_presenter) status = RS_MISS;  // no fragment: every rank sees the same test, one writes
                } else {
                    r.hx = rox; r.hy = roy; r.hz = roz;
                    marching = true;
                }
            }
        } else {
            // nothing to do right now.  Every way out of the loop passes the close test: the last warp to learn that an
            // in-queue is closed (and empty) may be the one that has to close the out-queue behind it
            if (lane == 0) { stream_maybe_close(L, 0, n_units); stream_maybe_close(L, 1, n_units); }
            if (!open) break;  // both in-queues are closed and handed out, no ticket pending: this warp is done
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (idle_since == 0ull) idle_since = now;
            if (now - idle_since > (unsigned long long)L.timeout_ms * 1000000ull) {
                if (lane == 0) *reinterpret_cast<volatile uint32_t*>(L.timed_out) = 1u;
                break;
            }
            __nanosleep(sleep_ns);
            if (sleep_ns < 1024u) sleep_ns *= 2u;
            continue;
        }

        // ---- sdfRaycast, material.frag:92-128 (the round kernel's loop)
        const int max_steps = (int)P.max_steps;
        CellCache cell;
        cell.x0 = cell.y0 = cell.z0 = __int_as_float(0x7fc00000);
        cell.c000 = cell.c100 = cell.c010 = cell.c110 = cell.c001 = cell.c101 = cell.c011 = cell.c111 = 0.0f;
        const bool started_here = src < 0;
        while (__ballot_sync(0xffffffffu, marching)) {
            if (marching) {
                float ax, ay, az;
                tex_coord<SNAP>(P, v0, r.hx, r.hy, r.hz, ax, ay, az);
                const int z = lower_z_tap<LINEAR>((int)P.D, az);
                const int go = z < (int)L.own_z0 ? RS_DOWN : (z >= (int)L.own_z1 ? RS_UP : 0);
                if (go && started_here && r.it == 0) {
                    status = RS_NONE; marching = false;   // a ray that STARTS elsewhere is started by its owner
                } else if (r.it >= max_steps - 1) {                                                  // :99-102
                    code = -1.0f; status = RS_MISS; marching = false;
                } else if (oob_dist(P.clip_min, P.clip_max, r.hx, r.hy, r.hz) > 1e-4f) {             // :106-109
                    code = -2.0f; status = RS_MISS; marching = false;
                } else if (go) {
                    status = go; marching = false;
                } else {
                    const float dist = dist_at<LINEAR>(v0, ax, ay, az, cell) - 1e-1f;                // :112, :59
                    if (dist < 1e-5f) {                                                              // :117-121
                        code = r.t; status = RS_HIT; marching = false;
                    } else {
                        r.t += dist;                                                                 // :124
                        r.hx += rdx * dist; r.hy += rdy * dist; r.hz += rdz * dist;                  // :125
                        ++r.it;
                    }
                }
            }
        }

        if (status == RS_HIT || status == RS_MISS) finish_pixel<SNAP, LINEAR>(P, L, v0, v1, px, status == RS_HIT, code, r);
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            const bool go = status == (dir ? RS_UP : RS_DOWN);
            const uint32_t m = __ballot_sync(0xffffffffu, go);
            if (m == 0u) continue;
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(L.out_count + dir, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (go && L.out_q[dir]) {
                const uint32_t e = base + __popc(m & ((1u << lane) - 1u));
                float4* dst = L.out_q[dir] + 2u * (size_t)e;
                st_volatile_f4(dst, r.hx, r.hy, r.hz, __uint_as_float(tag));
                st_volatile_f4(dst + 1, r.t, __uint_as_float(px), __uint_as_float((uint32_t)r.it), __uint_as_float(tag));
            }
        }
        __syncwarp();
        if (lane == 0) {
            // (the unit's reservations in out_count have been performed: their return values were used above.  No fence
            // here -- it would wait for the unit's pixel stores to cross NVLink)
            if (src < 0) {
                if (atomicAdd(L.tiles_done, 1u) + 1u == n_units) { stream_maybe_close(L, 0, n_units); stream_maybe_close(L, 1, n_units); }
                unit = n_warps + atomicAdd(L.work_head, 1u);
            } else if ((((src ? mask1 : mask0) & src_valid) == src_valid)) {
                // every entry of the unit has been traced; the unit that completes a closed in-queue closes the out-queue
                // its rays travel on
                const uint32_t done = atomicAdd(L.in_done + src, 1u) + 1u;
                const unsigned long long f = ld_volatile_u64(L.in_final[src]);
                if ((uint32_t)(f >> 32) == tag && done == ((uint32_t)f + 31u) / 32u) stream_maybe_close(L, 1 - src, n_units);
            }
        }
        if (src < 0) {
            unit = __shfl_sync(0xffffffffu, unit, 0);
        } else if (((src ? mask1 : mask0) & src_valid) == src_valid) {  // the ticket is used up
            if (src) { pend1 = NO_TICKET; mask1 = 0u; } else { pend0 = NO_TICKET; mask0 = 0u; }
        }
    }

    // ---- the last CTA to finish resets the counters and tells the neighbours and the presenter.  The CTA's peer
    // stores of finished pixels are ordered before thread 0's system fence by the barrier.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(L.ctas_done, 1u) + 1u == gridDim.x) {
            *L.work_head = 0u; *L.ctas_done = 0u; *L.tiles_done = 0u;
            if (L.outside_done) *L.outside_done = 0u;
            L.in_head[0] = L.in_head[1] = 0u; L.in_done[0] = L.in_done[1] = 0u;
            L.sent_final[0] = L.sent_final[1] = 0u; L.out_count[0] = L.out_count[1] = 0u;
            __threadfence_system();
            if (L.sig_round[0]) *reinterpret_cast<volatile uint32_t*>(L.sig_round[0]) = L.sig_round_value;
            if (L.sig_round[1]) *reinterpret_cast<volatile uint32_t*>(L.sig_round[1]) = L.sig_round_value;
            if (L.sig_frame) *reinterpret_cast<volatile uint32_t*>(L.sig_frame) = L.sig_frame_value;
        }
    }
}

// flags in (peer) memory: everything this stream did before is visible system-wide first
__global__ void signal_kernel(uint32_t* f0, uint32_t v0, uint32_t* f1, uint32_t v1, uint32_t* f2, uint32_t v2, uint32_t* f3,
                              uint32_t v3) {
    __threadfence_system();
    if (f0) *reinterpret_cast<volatile uint32_t*>(f0) = v0;
    if (f1) *reinterpret_cast<volatile uint32_t*>(f1) = v1;
    if (f2) *reinterpret_cast<volatile uint32_t*>(f2) = v2;
    if (f3) *reinterpret_cast<volatile uint32_t*>(f3) = v3;
}

// fallback for cuStreamWaitValue32: one thread spins until (int)(*flag - value) >= 0, at most ~2 s
__global__ void spin_wait_kernel(const uint32_t* flag, uint32_t value, uint32_t* timed_out) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const uint32_t v = *reinterpret_cast<const volatile uint32_t*>(flag);
        if ((int)(v - value) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) { if (timed_out) *timed_out = 1u; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// tex0.r of every stored texel as a dense float array (the optional distance volume of the tracer)
__global__ void __launch_bounds__(256) extract_dist_kernel(const float4* __restrict__ tex0, float* __restrict__ dist,
                                                           size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dist[i] = __ldg(reinterpret_cast<const float*>(tex0 + i));
}

// the same into an R32F 3-D CUDA array (block-linear, what the texture unit reads) through a surface
__global__ void __launch_bounds__(256) extract_dist_surf_kernel(const float4* __restrict__ tex0, cudaSurfaceObject_t surf,
                                                                uint32_t W, uint32_t H, uint32_t Ds) {
    const size_t n = (size_t)W * H * Ds;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % W), y = (uint32_t)((i / W) % H), z = (uint32_t)(i / ((size_t)W * H));
        surf3Dwrite(__ldg(reinterpret_cast<const float*>(tex0 + i)), surf, (int)(x * sizeof(float)), (int)y, (int)z);
    }
}

__global__ void keys_unpack_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint8_t* __restrict__ rgba8,
                                   float* __restrict__ depth) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    if (rgba8) reinterpret_cast<uint32_t*>(rgba8)[i] = (uint32_t)(k & 0xffffffffull);
    if (depth) depth[i] = __uint_as_float((uint32_t)(k >> 32));
}

}  // namespace

// The tiles inside the screen rectangle of the box, sorted for trace_tiles_kernel: by band (in the order the kernel walks
// the bands), within a band by the longest march the tile had in the PREVIOUS frame, longest first.  A counting sort
// by one CTA (the keys are 32 bands x 256 step counts).  Always a permutation of the CURRENT rectangle's tiles -- the old
// costs only decide the order, so a camera that has moved costs some of the gain, never a pixel.
namespace {
__global__ void __launch_bounds__(1024) tile_order_kernel(const __grid_constant__ TraceParams P, const uint32_t* __restrict__ prev,
                                                          uint32_t* __restrict__ order) {
    constexpr uint32_t NB = TRACE_MAX_BANDS * 256u;
    __shared__ uint32_t bins[NB];
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t band_pos[TRACE_MAX_BANDS];
    const uint32_t tid = threadIdx.x, rw = P.rect[2] - P.rect[0], rh = P.rect[3] - P.rect[1], n = rw * rh;
    for (uint32_t k = tid; k < NB; k += 1024u) bins[k] = 0u;
    if (tid < TRACE_MAX_BANDS) band_pos[tid] = 0u;
    __syncthreads();
    if (tid < P.n_bands) band_pos[P.band_order[tid]] = tid;
    __syncthreads();
    auto key_of = [&](uint32_t i, uint32_t& id) {
        const uint32_t tx = P.rect[0] + i % rw, ty = P.rect[1] + i / rw;
        id = ty * P.tiles_x + tx;
        const uint32_t cost = min(prev[id], 255u);
        const uint32_t pos = P.n_bands ? band_pos[min(ty / P.band_rows, TRACE_MAX_BANDS - 1u)] : 0u;
        return pos * 256u + (255u - cost);
    };
    for (uint32_t i = tid; i < n; i += 1024u) {
        uint32_t id;
        atomicAdd(&bins[key_of(i, id)], 1u);
    }
    __syncthreads();
    // exclusive scan of the bins: 8 per thread, then the 1024 partial sums
    constexpr uint32_t PER = NB / 1024u;
    uint32_t local[PER], sum = 0u;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { local[k] = bins[tid * PER + k]; sum += local[k]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31u) >= (uint32_t)o) incl += v;
    }
    if ((tid & 31u) == 31u) warp_sums[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32u) {
        uint32_t w = warp_sums[tid], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= (uint32_t)o) wi += v;
        }
        warp_sums[tid] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sums[tid >> 5] + incl - sum;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) { bins[tid * PER + k] = run; run += local[k]; }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 1024u) {
        uint32_t id;
        const uint32_t key = key_of(i, id);
        order[atomicAdd(&bins[key], 1u)] = id;
    }
}
}  // namespace

cudaError_t launch_tile_order(const TraceParams& p, const uint32_t* prev_cost, uint32_t* order, cudaStream_t s) {
    tile_order_kernel<<<1, 1024, 0, s>>>(p, prev_cost, order);
    return cudaGetLastError();
}

cudaError_t launch_trace(const TraceParams& p, int variant, cudaStream_t s) {
    if (p.width == 0 || p.height == 0) return cudaSuccess;
    const bool snap = p.lod != 1.0f, lin = p.filter_linear != 0;
    if (variant == 0) {
        unsigned grid = 0;  // per band: a CTA per tile inside the rectangle, a CTA per run of tiles outside it
        for (uint32_t k = 0; k < (p.n_bands ? p.n_bands : 1u); ++k) {
            const uint32_t ty0 = p.n_bands ? k * p.band_rows : 0u;
            const uint32_t ty1 = p.n_bands ? (ty0 + p.band_rows < p.tiles_y ? ty0 + p.band_rows : p.tiles_y) : p.tiles_y;
            const uint32_t ry0 = p.rect[1] < ty0 ? ty0 : (p.rect[1] > ty1 ? ty1 : p.rect[1]);
            const uint32_t ry1 = p.rect[3] < ty0 ? ty0 : (p.rect[3] > ty1 ? ty1 : p.rect[3]);
            const uint32_t n_heavy = (p.rect[2] - p.rect[0]) * (ry1 - ry0), n_out = p.tiles_x * (ty1 - ty0) - n_heavy;
            grid += n_heavy + (n_out + TRACE_OUTSIDE_RUN - 1u) / TRACE_OUTSIDE_RUN;
        }
        const uint32_t mode = lin ? p.dist_mode : 0u;  // the distance volumes serve the LINEAR march only
        if (p.full_dist) {  // exact multi-GPU trace: replicated full-grid distance volume, hits shaded by their owner
            if (!snap && lin) trace_tiles_kernel<false, true, 1, true><<<grid, 64, 0, s>>>(p);
            else if (!snap && !lin) trace_tiles_kernel<false, false, 0, true><<<grid, 64, 0, s>>>(p);
            else if (snap && lin) trace_tiles_kernel<true, true, 1, true><<<grid, 64, 0, s>>>(p);
            else trace_tiles_kernel<true, false, 0, true><<<grid, 64, 0, s>>>(p);
        } else if (!snap && lin && mode == 1) trace_tiles_kernel<false, true, 1><<<grid, 64, 0, s>>>(p);
        else if (!snap && lin && mode == 2) trace_tiles_kernel<false, true, 2><<<grid, 64, 0, s>>>(p);
        else if (!snap && lin && mode == 3) trace_tiles_kernel<false, true, 3><<<grid, 64, 0, s>>>(p);
        else if (!snap && lin) trace_tiles_kernel<false, true><<<grid, 64, 0, s>>>(p);
        else if (!snap && !lin) trace_tiles_kernel<false, false><<<grid, 64, 0, s>>>(p);
        else if (snap && lin && mode == 1) trace_tiles_kernel<true, true, 1><<<grid, 64, 0, s>>>(p);
        else if (snap && lin && mode == 2) trace_tiles_kernel<true, true, 2><<<grid, 64, 0, s>>>(p);
        else if (snap && lin && mode == 3) trace_tiles_kernel<true, true, 3><<<grid, 64, 0, s>>>(p);
        else if (snap && lin) trace_tiles_kernel<true, true><<<grid, 64, 0, s>>>(p);
        else trace_tiles_kernel<true, false><<<grid, 64, 0, s>>>(p);
    } else {
        const dim3 grid((p.width + 31) / 32, (p.height + 7) / 8);
        if (!snap && lin) trace_kernel<false, true><<<grid, 256, 0, s>>>(p);
        else if (!snap && !lin) trace_kernel<false, false><<<grid, 256, 0, s>>>(p);
        else if (snap && lin) trace_kernel<true, true><<<grid, 256, 0, s>>>(p);
        else trace_kernel<true, false><<<grid, 256, 0, s>>>(p);
    }
    return cudaGetLastError();
}

typedef void (*rounds_fn)(const TraceParams, const LinkParams);
static rounds_fn pick_rounds(const TraceParams& p) {
    const bool snap = p.lod != 1.0f, lin = p.filter_linear != 0;
    return !snap && lin ? trace_rounds_kernel<false, true> : !snap ? trace_rounds_kernel<false, false>
           : lin ? trace_rounds_kernel<true, true> : trace_rounds_kernel<true, false>;
}

int trace_rounds_max_ctas_per_sm(const TraceParams& p) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pick_rounds(p), 256, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

cudaError_t launch_trace_rounds(const TraceParams& p, const LinkParams& l, int grid, cudaStream_t s) {
    if (p.width == 0 || p.height == 0 || grid <= 0) return cudaSuccess;
    pick_rounds(p)<<<grid, 256, 0, s>>>(p, l);
    return cudaGetLastError();
}

static rounds_fn pick_stream(const TraceParams& p) {
    const bool snap = p.lod != 1.0f, lin = p.filter_linear != 0;
    return !snap && lin ? trace_stream_kernel<false, true> : !snap ? trace_stream_kernel<false, false>
           : lin ? trace_stream_kernel<true, true> : trace_stream_kernel<true, false>;
}

int trace_stream_max_ctas_per_sm(const TraceParams& p) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pick_stream(p), 256, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

cudaError_t launch_trace_stream(const TraceParams& p, const LinkParams& l, int grid, cudaStream_t s) {
    if (p.width == 0 || p.height == 0 || grid <= 0) return cudaSuccess;
    pick_stream(p)<<<grid, 256, 0, s>>>(p, l);
    return cudaGetLastError();
}

cudaError_t launch_signal(uint32_t* const* flags, const uint32_t* values, int n, cudaStream_t s) {
    for (int i = 0; i < n; i += 4) {
        uint32_t* f[4] = {nullptr, nullptr, nullptr, nullptr};
        uint32_t v[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4 && i + k < n; ++k) { f[k] = flags[i + k]; v[k] = values[i + k]; }
        signal_kernel<<<1, 1, 0, s>>>(f[0], v[0], f[1], v[1], f[2], v[2], f[3], v[3]);
    }
    return cudaGetLastError();
}

cudaError_t launch_spin_wait(const uint32_t* flag, uint32_t value, uint32_t* timed_out, cudaStream_t s) {
    spin_wait_kernel<<<1, 1, 0, s>>>(flag, value, timed_out);
    return cudaGetLastError();
}

cudaError_t launch_extract_dist(const float4* tex0, float* dist, size_t n, int grid, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    extract_dist_kernel<<<grid, 256, 0, s>>>(tex0, dist, n);
    return cudaGetLastError();
}

cudaError_t launch_extract_dist_array(const float4* tex0, unsigned long long surf, uint32_t W, uint32_t H, uint32_t Ds,
                                      int grid, cudaStream_t s) {
    if ((size_t)W * H * Ds == 0) return cudaSuccess;
    extract_dist_surf_kernel<<<grid, 256, 0, s>>>(tex0, (cudaSurfaceObject_t)surf, W, H, Ds);
    return cudaGetLastError();
}

cudaError_t launch_keys_unpack(const unsigned long long* keys, uint32_t n, uint8_t* rgba8, float* depth,
                               cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    // small CTAs: the kernel shares the SMs with the next fill (a few hundred free registers per SM are enough)
    keys_unpack_kernel<<<(n + 63) / 64, 64, 0, s>>>(keys, n, rgba8, depth);
    return cudaGetLastError();
}

}  // namespace sdfgpu

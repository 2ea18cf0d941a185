// sdfgpu_viewer.hpp -- C++17 host side above the C ABI of sdfgpu.h (header only).
//
// The reference's host language is Rust, which this image cannot build; the next compiled
// language is C++, so this header is the host-side mirror a Rust maintainer would write with
// `extern "C"` (INTEGRATION.md has that Rust).  Same names, argument meaning and error behaviour
// as the reference for the hot path; paths below are relative to /root/reference.
//
//   sdfgpu::SDFSample       src/sdf/mod.rs:104-126      (#[repr(C)], 7 x f32 = 28 bytes)
//   sdfgpu::SDFSurface      src/sdf/mod.rs:33-101       (the trait; + the optional tape() capability)
//   sdfgpu::LoadingManager  src/app/scene/sdf/loading.rs:5-115
//   sdfgpu::SDFViewer       src/app/scene/sdf/mod.rs:21-251  (from_bb, new_voxels, update, commit)
//   sdfgpu::TapeBuilder / sdfgpu::SDFDemo   the tape of sdfgpu_tape.h; src/sdf/demo/mod.rs:20-75
//
// Nothing here computes on the CPU: update() hands the surface to sdfgpu_update_surface, which
// evaluates a tape on the GPU, or -- for a surface without one -- calls the surface's own
// sample() (user code, e.g. a WASM guest) and lets the GPU apply the store rules.
#ifndef SDFGPU_VIEWER_HPP
#define SDFGPU_VIEWER_HPP

#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <initializer_list>
#include <limits>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sdfgpu.h"
#include "sdfgpu_tape.h"

namespace sdfgpu {

struct Vector3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
};
using BoundingBox = std::array<Vector3, 2>;  // [min, max], src/sdf/mod.rs:37

// src/sdf/mod.rs:104-118; the exact bytes the WASM host reads (src/sdf/wasm/native.rs:204-216)
struct SDFSample {
    float distance = 0.0f;
    float color[3] = {0.0f, 0.0f, 0.0f};
    float metallic = 0.0f;
    float roughness = 0.0f;
    float occlusion = 0.0f;
    SDFSample() = default;
    SDFSample(float d, float r, float g, float b) : distance(d), color{r, g, b} {}  // SDFSample::new, :122-125
};
static_assert(sizeof(SDFSample) == 28, "SDFSample must be 7 packed floats");

// A failing call of the C ABI.  The reference never panics on a failing guest (it logs and returns a
// benign value, src/sdf/wasm/native.rs:196-203); errors here are host-side misuse or CUDA failures.
class Error : public std::runtime_error {
   public:
    Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }

   private:
    int code_;
};

inline void check(int rc, const sdfgpu_ctx* ctx = nullptr) {
    if (rc != SDFGPU_OK) {
        const char* m = sdfgpu_last_error(ctx);
        throw Error(rc, m ? m : "");
    }
}

// trait SDFSurface, src/sdf/mod.rs:33-101 (the part the hot path uses)
class SDFSurface {
   public:
    virtual ~SDFSurface() = default;
    virtual BoundingBox bounding_box() const = 0;                       // :37
    virtual SDFSample sample(Vector3 p, bool distance_only) const = 0;  // :43
    // "TODO: Batched sampling" (:39): override when the implementation can do better than a loop
    virtual void sample_batch(const float* xyz, uint64_t n, bool distance_only, SDFSample* out) const {
        for (uint64_t i = 0; i < n; ++i) out[i] = sample(Vector3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, distance_only);
    }
    virtual std::optional<BoundingBox> changed() const { return std::nullopt; }  // :87
    // Not in the reference: bytes of a tape (sdfgpu_tape.h) equivalent to sample(p, false).  A surface
    // that returns one is evaluated on the GPU and sample() is never called by the viewer.
    virtual std::optional<std::vector<unsigned char>> tape() const { return std::nullopt; }
    // How many host threads may call sample() at once (the reference: one, scene/sdf/mod.rs:174)
    virtual unsigned sample_threads() const { return 1; }
};

// LoadingManager, src/app/scene/sdf/loading.rs (device-free; the same state a viewer keeps)
class LoadingManager {
   public:
    LoadingManager(std::array<uint32_t, 3> limits, uint32_t passes) : limits(limits), passes(passes) {
        check(sdfgpu_loading_create(limits.data(), passes, &h_));
    }
    ~LoadingManager() { sdfgpu_loading_destroy(h_); }
    LoadingManager(const LoadingManager&) = delete;
    LoadingManager& operator=(const LoadingManager&) = delete;
    void reset(uint32_t p) { passes = p; sdfgpu_loading_reset(h_, p); }       // :37-43
    std::optional<std::array<uint32_t, 3>> next() {                           // :50-76
        std::array<uint32_t, 3> idx{};
        if (!sdfgpu_loading_next(h_, idx.data())) return std::nullopt;
        return idx;
    }
    uint64_t len() const { return sdfgpu_loading_len(h_); }                              // :80-89
    uint64_t total_iterations() const { return sdfgpu_loading_total_iterations(h_); }    // :94-96
    uint32_t passes_left() const { return sdfgpu_loading_passes_left(h_); }              // :99-105

    std::array<uint32_t, 3> limits;
    uint32_t passes;

   private:
    sdfgpu_loading* h_ = nullptr;
};

struct Frame {  // what `volume.render(&camera, lights)` leaves in the framebuffer (scene/mod.rs:213-215)
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba8;  // row 0 = bottom row (GL window coordinates)
    std::vector<float> depth;    // gl_FragDepth, 1.0 where no fragment was written
};

// SDFViewer, src/app/scene/sdf/mod.rs:21-251
class SDFViewer {
   public:
    // `viewer.loading_mgr.{len(), total_iterations(), passes_left(), passes, limits}` as the scene
    // reads them (scene/mod.rs:153,229-239)
    struct LoadingView {
        const sdfgpu_ctx* ctx;
        std::array<uint32_t, 3> limits;
        uint32_t passes;
        uint64_t len() const { uint64_t v = 0; check(sdfgpu_loading_state(ctx, &v, nullptr, nullptr, nullptr), ctx); return v; }
        uint64_t total_iterations() const { uint64_t v = 0; check(sdfgpu_loading_state(ctx, nullptr, &v, nullptr, nullptr), ctx); return v; }
        uint32_t passes_left() const { uint32_t v = 0; check(sdfgpu_loading_state(ctx, nullptr, nullptr, &v, nullptr), ctx); return v; }
    };

    // from_bb(ctx, bb, max_voxels_side, loading_passes), :46-72 (the three-d Context becomes a device index)
    static SDFViewer from_bb(const BoundingBox& bb, uint32_t max_voxels_side, uint32_t loading_passes, int device = 0) {
        const std::array<float, 6> b = flat(bb);
        sdfgpu_ctx* h = nullptr;
        check(sdfgpu_create(b.data(), max_voxels_side, loading_passes, device, &h));
        return SDFViewer(h, bb);
    }
    // new_voxels(ctx, voxels, bb, loading_passes), :75-101
    static SDFViewer new_voxels(std::array<uint32_t, 3> voxels, const BoundingBox& bb, uint32_t loading_passes, int device = 0) {
        const std::array<float, 6> b = flat(bb);
        sdfgpu_ctx* h = nullptr;
        check(sdfgpu_create_voxels(b.data(), voxels.data(), loading_passes, device, &h));
        return SDFViewer(h, bb);
    }
    SDFViewer(SDFViewer&& o) noexcept : bounding_box(o.bounding_box), h_(o.h_) { o.h_ = nullptr; }
    SDFViewer& operator=(SDFViewer&& o) noexcept {
        if (this != &o) { sdfgpu_destroy(h_); h_ = o.h_; bounding_box = o.bounding_box; o.h_ = nullptr; }
        return *this;
    }
    SDFViewer(const SDFViewer&) = delete;
    SDFViewer& operator=(const SDFViewer&) = delete;
    ~SDFViewer() { sdfgpu_destroy(h_); }

    // update(&mut self, sdf: impl SDFSurface, max_delta_time) -> usize, :128-217.  Exceptions thrown by
    // the surface's callbacks do not cross the C ABI: the first one is rethrown here after the call, and
    // the failing samples take the reference's benign value (distance 1.0, native.rs:202).
    template <class Rep, class Period>
    size_t update(const SDFSurface& sdf, std::chrono::duration<Rep, Period> max_delta_time) {
        Trampoline t{&sdf, {}, nullptr};
        sdfgpu_surface s;
        std::memset(&s, 0, sizeof s);
        s.self = &t;
        s.bounding_box = &Trampoline::bounding_box;
        s.sample = &Trampoline::sample;
        s.sample_batch = &Trampoline::sample_batch;
        s.changed = &Trampoline::changed;
        s.tape = &Trampoline::tape;
        s.sample_threads = sdf.sample_threads();
        uint64_t iterations = 0;
        const int rc = sdfgpu_update_surface(h_, &s, std::chrono::duration<double>(max_delta_time).count(), &iterations);
        if (t.error) std::rethrow_exception(t.error);
        check(rc, h_);
        return (size_t)iterations;
    }

    void commit() { check(sdfgpu_commit(h_), h_); }  // :220-239

    LoadingView loading_mgr() const {
        uint32_t p = 0;
        check(sdfgpu_loading_state(h_, nullptr, nullptr, nullptr, &p), h_);
        return LoadingView{h_, voxels(), p};
    }

    // tex0.width / height / depth, read by scene/mod.rs:149-150
    std::array<uint32_t, 3> voxels() const {
        std::array<uint32_t, 3> d{};
        check(sdfgpu_dims(h_, d.data()), h_);
        return d;
    }
    uint32_t width() const { return voxels()[0]; }
    uint32_t height() const { return voxels()[1]; }
    uint32_t depth() const { return voxels()[2]; }

    // The CPU-side Vec<[f32;4]> volumes (:23-25), x fastest, then y, then z
    void download(std::vector<float>* tex0, std::vector<float>* tex1) {
        const auto d = voxels();
        const size_t n = (size_t)d[0] * d[1] * d[2] * 4;
        if (tex0) tex0->resize(n);
        if (tex1) tex1->resize(n);
        check(sdfgpu_download(h_, tex0 ? tex0->data() : nullptr, tex1 ? tex1->data() : nullptr), h_);
    }

    // Stands where the scene calls `volume.render(&camera, lights)` (scene/mod.rs:213-215): the
    // SDFViewerMaterial fragment shader (material.frag:130-182) for every pixel.
    Frame render(const sdfgpu_camera& camera, uint32_t width, uint32_t height) {
        Frame f;
        f.width = width; f.height = height;
        f.rgba8.resize((size_t)width * height * 4);
        f.depth.resize((size_t)width * height);
        check(sdfgpu_trace_rgba8(h_, &camera, width, height, f.rgba8.data(), f.depth.data()), h_);
        return f;
    }

    static sdfgpu_camera default_camera(uint32_t width, uint32_t height) {  // scene/mod.rs:82-95
        sdfgpu_camera c;
        sdfgpu_camera_default(&c, width, height);
        return c;
    }

    sdfgpu_ctx* handle() { return h_; }
    BoundingBox bounding_box;

   private:
    friend class SDFViewerGroup;
    SDFViewer(sdfgpu_ctx* h, const BoundingBox& bb) : bounding_box(bb), h_(h) {}
    static std::array<float, 6> flat(const BoundingBox& bb) {
        return {bb[0].x, bb[0].y, bb[0].z, bb[1].x, bb[1].y, bb[1].z};
    }

    // C callbacks of sdfgpu_surface -> virtual calls; exceptions are parked, not propagated through C
    struct Trampoline {
        const SDFSurface* sdf;
        std::vector<unsigned char> tape_bytes;
        std::exception_ptr error;

        static void bounding_box(void* self, float out[6]) {
            auto* t = static_cast<Trampoline*>(self);
            try {
                const auto b = flat(t->sdf->bounding_box());
                std::memcpy(out, b.data(), sizeof(float) * 6);
            } catch (...) { t->park(); }
        }
        static void benign(float* out, uint64_t n) {  // native.rs:202: SDFSample::new(1.0, zero)
            for (uint64_t i = 0; i < n; ++i) {
                std::memset(out + 7 * i, 0, 28);
                out[7 * i] = 1.0f;
            }
        }
        static void sample(void* self, const float p[3], int distance_only, float out[7]) {
            auto* t = static_cast<Trampoline*>(self);
            try {
                const SDFSample s = t->sdf->sample(Vector3{p[0], p[1], p[2]}, distance_only != 0);
                std::memcpy(out, &s, 28);
            } catch (...) { t->park(); benign(out, 1); }
        }
        static void sample_batch(void* self, const float* xyz, uint64_t n, int distance_only, float* out) {
            auto* t = static_cast<Trampoline*>(self);
            try {
                t->sdf->sample_batch(xyz, n, distance_only != 0, reinterpret_cast<SDFSample*>(out));
            } catch (...) { t->park(); benign(out, n); }
        }
        static int changed(void* self, float out[6]) {
            auto* t = static_cast<Trampoline*>(self);
            try {
                const auto c = t->sdf->changed();
                if (!c) return 0;
                const auto b = flat(*c);
                std::memcpy(out, b.data(), sizeof(float) * 6);
                return 1;
            } catch (...) { t->park(); return 0; }
        }
        static int tape(void* self, const void** bytes, size_t* len) {
            auto* t = static_cast<Trampoline*>(self);
            try {
                auto b = t->sdf->tape();
                if (!b || b->empty()) return 0;
                t->tape_bytes = std::move(*b);
                *bytes = t->tape_bytes.data();
                *len = t->tape_bytes.size();
                return 1;
            } catch (...) { t->park(); return 0; }
        }
        void park() {  // may run on several sampling threads: keep the first exception
            static std::mutex m;
            std::lock_guard<std::mutex> g(m);
            if (!error) error = std::current_exception();
        }
    };

    sdfgpu_ctx* h_ = nullptr;
};

// The same viewer over several GPUs, driven by ONE thread like the reference's scene (scene/mod.rs:22-31,158-225):
// the grid is sharded along z over the devices, every call runs on all of them, and render() returns the frame of
// one SDFViewer holding the whole grid, bit for bit (include/sdfgpu.h "linked slabs", sdfgpu_group_*).  Not in the
// reference.
class SDFViewerGroup {
   public:
    // SDFViewer::from_bb over the devices of `device_mask` (bit d = CUDA device d); frames of at most max_width x max_height
    static SDFViewerGroup from_bb(const BoundingBox& bb, uint32_t max_voxels_side, uint32_t loading_passes, uint32_t device_mask,
                                  uint32_t max_width = 1920, uint32_t max_height = 1080) {
        const std::array<float, 6> b = SDFViewer::flat(bb);
        sdfgpu_group* g = nullptr;
        check(sdfgpu_group_create_mask(b.data(), max_voxels_side, loading_passes, device_mask, max_width, max_height, &g));
        return SDFViewerGroup(g, bb);
    }
    // SDFViewer::new_voxels over an explicit device list (a device may appear more than once)
    static SDFViewerGroup new_voxels(std::array<uint32_t, 3> voxels, const BoundingBox& bb, uint32_t loading_passes,
                                     const std::vector<int>& devices, uint32_t max_width = 1920, uint32_t max_height = 1080) {
        const std::array<float, 6> b = SDFViewer::flat(bb);
        sdfgpu_group* g = nullptr;
        check(sdfgpu_group_create(b.data(), voxels.data(), loading_passes, devices.data(), (uint32_t)devices.size(), max_width,
                                  max_height, 0, &g));
        return SDFViewerGroup(g, bb);
    }
    SDFViewerGroup(SDFViewerGroup&& o) noexcept : bounding_box(o.bounding_box), g_(o.g_) { o.g_ = nullptr; }
    SDFViewerGroup(const SDFViewerGroup&) = delete;
    SDFViewerGroup& operator=(const SDFViewerGroup&) = delete;
    ~SDFViewerGroup() { sdfgpu_group_destroy(g_); }

    uint32_t size() const { return sdfgpu_group_size(g_); }

    template <class Rep, class Period>
    size_t update(const SDFSurface& sdf, std::chrono::duration<Rep, Period> max_delta_time) {  // SDFViewer::update
        SDFViewer::Trampoline t{&sdf, {}, nullptr};
        sdfgpu_surface s;
        std::memset(&s, 0, sizeof s);
        s.self = &t;
        s.bounding_box = &SDFViewer::Trampoline::bounding_box;
        s.sample = &SDFViewer::Trampoline::sample;
        s.sample_batch = &SDFViewer::Trampoline::sample_batch;
        s.changed = &SDFViewer::Trampoline::changed;
        s.tape = &SDFViewer::Trampoline::tape;
        s.sample_threads = sdf.sample_threads();
        uint64_t iterations = 0;
        const int rc = sdfgpu_group_update_surface(g_, &s, std::chrono::duration<double>(max_delta_time).count(), &iterations);
        if (t.error) std::rethrow_exception(t.error);
        gcheck(rc);
        return (size_t)iterations;
    }
    void commit() { gcheck(sdfgpu_group_commit(g_)); }
    uint64_t loading_len() const {
        uint64_t v = 0;
        gcheck(sdfgpu_group_loading_state(g_, &v, nullptr, nullptr, nullptr));
        return v;
    }
    std::array<uint32_t, 3> voxels() const {
        std::array<uint32_t, 3> d{};
        check(sdfgpu_dims(sdfgpu_group_rank(g_, 0), d.data()));
        return d;
    }
    void download(std::vector<float>* tex0, std::vector<float>* tex1) {  // the whole grid
        const auto d = voxels();
        const size_t n = (size_t)d[0] * d[1] * d[2] * 4;
        if (tex0) tex0->resize(n);
        if (tex1) tex1->resize(n);
        gcheck(sdfgpu_group_download(g_, tex0 ? tex0->data() : nullptr, tex1 ? tex1->data() : nullptr));
    }
    Frame render(const sdfgpu_camera& camera, uint32_t width, uint32_t height) {  // volume.render(&camera, lights)
        Frame f;
        f.width = width; f.height = height;
        f.rgba8.resize((size_t)width * height * 4);
        f.depth.resize((size_t)width * height);
        gcheck(sdfgpu_group_trace_rgba8(g_, &camera, width, height, f.rgba8.data(), f.depth.data()));
        return f;
    }
    sdfgpu_group* handle() { return g_; }
    BoundingBox bounding_box;

   private:
    SDFViewerGroup(sdfgpu_group* g, const BoundingBox& bb) : bounding_box(bb), g_(g) {}
    void gcheck(int rc) const {
        if (rc != SDFGPU_OK) throw Error(rc, sdfgpu_group_last_error(g_));
    }
    sdfgpu_group* g_ = nullptr;
};

// ---------------------------------------------------------------- tape (include/sdfgpu_tape.h)

// A scalar program (sdft_sop run, include/sdfgpu_tape.h): straight-line SSA arithmetic with WebAssembly's
// numerics.  Each call appends one op and returns the index of its value; TapeBuilder::scalar places it.
class ScalarProgram {
   public:
    uint32_t op(uint32_t code, uint32_t a = 0, uint32_t b = 0, uint32_t c = 0) {
        ops_.push_back(sdft_sop{code, a, b, c});
        return (uint32_t)ops_.size() - 1;
    }
    uint32_t px() { return op(SDFT_S_PX); }
    uint32_t py() { return op(SDFT_S_PY); }
    uint32_t pz() { return op(SDFT_S_PZ); }
    uint32_t constant(float v) {  // runtime data of the tape: a new value keeps the compiled kernel
        consts_.push_back(v);
        return op(SDFT_S_CONST, (uint32_t)consts_.size() - 1);
    }
    uint32_t imm(uint32_t word) { return op(SDFT_S_IMM, word); }
    uint32_t out(uint32_t channel, uint32_t value) { return op(SDFT_S_OUT, value, channel); }

   private:
    friend class TapeBuilder;
    std::vector<sdft_sop> ops_;
    std::vector<float> consts_;
};

class TapeBuilder {
   public:
    uint32_t prim(uint32_t shape, Vector3 center, float size, uint32_t material = SDFT_MAT_FLAT,
                  std::array<float, 3> color = {0.0f, 0.0f, 0.0f}, float metallic = 0.0f, float roughness = 0.0f,
                  float occlusion = 0.0f, float air_skip = std::numeric_limits<float>::infinity()) {
        sdft_prim p;
        p.center[0] = center.x; p.center[1] = center.y; p.center[2] = center.z;
        p.size = size;
        p.color[0] = color[0]; p.color[1] = color[1]; p.color[2] = color[2];
        p.metallic = metallic; p.roughness = roughness; p.occlusion = occlusion; p.air_skip = air_skip;
        p.kind = shape | (material << 8);
        prims_.push_back(p);
        return (uint32_t)prims_.size() - 1;
    }
    uint32_t constants(std::initializer_list<float> values) {
        const uint32_t first = (uint32_t)consts_.size();
        consts_.insert(consts_.end(), values.begin(), values.end());
        return first;
    }
    TapeBuilder& emit(uint32_t op, uint32_t a = 0, uint32_t b = 0, float imm = 0.0f) {
        instr_.push_back(sdft_instr{op, a, b, imm});
        return *this;
    }
    TapeBuilder& scalar(const ScalarProgram& p) {  // append the program and the SDFT_OP_SCALAR that runs it
        const uint32_t first = (uint32_t)sops_.size();
        for (sdft_sop o : p.ops_) {
            if (o.op == SDFT_S_CONST) {
                consts_.push_back(p.consts_[o.a]);
                o.a = (uint32_t)consts_.size() - 1;
            }
            sops_.push_back(o);
        }
        return emit(SDFT_OP_SCALAR, first, (uint32_t)p.ops_.size());
    }
    std::vector<unsigned char> build() const {
        sdft_header h;
        std::memset(&h, 0, sizeof h);
        h.magic = SDFT_MAGIC; h.version = SDFT_VERSION;
        h.n_instr = (uint32_t)instr_.size(); h.n_prims = (uint32_t)prims_.size(); h.n_consts = (uint32_t)consts_.size();
        h.reserved[0] = (uint32_t)sops_.size();
        std::vector<unsigned char> out(sizeof h + instr_.size() * sizeof(sdft_instr) + prims_.size() * sizeof(sdft_prim) +
                                       consts_.size() * sizeof(float) + sops_.size() * sizeof(sdft_sop));
        unsigned char* p = out.data();
        std::memcpy(p, &h, sizeof h); p += sizeof h;
        if (!instr_.empty()) std::memcpy(p, instr_.data(), instr_.size() * sizeof(sdft_instr));
        p += instr_.size() * sizeof(sdft_instr);
        if (!prims_.empty()) std::memcpy(p, prims_.data(), prims_.size() * sizeof(sdft_prim));
        p += prims_.size() * sizeof(sdft_prim);
        if (!consts_.empty()) std::memcpy(p, consts_.data(), consts_.size() * sizeof(float));
        p += consts_.size() * sizeof(float);
        if (!sops_.empty()) std::memcpy(p, sops_.data(), sops_.size() * sizeof(sdft_sop));
        return out;
    }

   private:
    std::vector<sdft_instr> instr_;
    std::vector<sdft_prim> prims_;
    std::vector<float> consts_;
    std::vector<sdft_sop> sops_;
};

// SDFDemo (src/sdf/demo/mod.rs:20-75) as a surface that lowers itself to a tape: an L-inf cube with a
// brick texture minus a sphere coloured by its normal, with a seam material.  It is evaluated on the
// GPU; sample() is deliberately not a second CPU implementation.
class SDFDemo : public SDFSurface {
   public:
    float cube_half_side = 0.95f;                 // cube.rs:17
    uint32_t cube_material = SDFT_MAT_BRICK;      // cube.rs:16
    float sphere_radius = 1.05f;                  // sphere.rs:13
    uint32_t sphere_material = SDFT_MAT_NORMAL;   // sphere.rs:12
    float max_distance_custom_material = 0.05f;   // demo/mod.rs:26
    bool disable_sphere = false;                  // demo/mod.rs:25

    BoundingBox bounding_box() const override {  // demo/mod.rs:47-49
        return {Vector3{-1.0f, -1.0f, -1.0f}, Vector3{1.0f, 1.0f, 1.0f}};
    }
    SDFSample sample(Vector3, bool) const override {
        throw std::logic_error("sdfgpu::SDFDemo is evaluated on the GPU through tape(); it has no host sample()");
    }
    // set_parameter (demo/mod.rs:117-132): any edit dirties the whole SDF; changed() reports it once (:136-145)
    void mark_changed() { changed_ = true; }
    std::optional<BoundingBox> changed() const override {
        if (!changed_) return std::nullopt;
        changed_ = false;
        return bounding_box();
    }
    std::optional<std::vector<unsigned char>> tape() const override {
        TapeBuilder t;
        // "the air has no texture": the material is skipped when the distance exceeds 0.1 (cube.rs:83, sphere.rs:41)
        const uint32_t box = t.prim(SDFT_SHAPE_BOX_LINF, Vector3{}, cube_half_side, cube_material, {0, 0, 0}, 0, 0, 0, 0.1f);
        t.emit(SDFT_OP_PRIM, box);
        if (!disable_sphere) {  // demo/mod.rs:54-55
            const uint32_t sph = t.prim(SDFT_SHAPE_SPHERE, Vector3{}, sphere_radius, sphere_material, {0, 0, 0}, 0, 0, 0, 0.1f);
            // seam threshold, then the forced seam material of demo/mod.rs:66-69
            const uint32_t c = t.constants({max_distance_custom_material, 0.5f, 0.6f, 0.7f, 0.5f, 0.0f, 0.0f});
            t.emit(SDFT_OP_PUSH).emit(SDFT_OP_PRIM, sph).emit(SDFT_OP_POP_DEMO_DIFF, c);
        }
        t.emit(SDFT_OP_END);
        return t.build();
    }

   private:
    mutable bool changed_ = false;  // the trait takes &self: interior mutability, as in the reference
};

// An existing .wasm SDF (the guest ABI of src/sdf/wasm/mod.rs:5-37, loaded by the reference with
// src/sdf/wasm/native.rs) as a surface with a tape: `sdfgpu_wasm_lower` executes the module once with symbolic
// coordinates and the GPU evaluates the result.  Throws sdfgpu::Error with the reason when the guest cannot be
// lowered; the host then wraps its own WASM runtime in an SDFSurface whose sample() calls the guest, and the
// viewer samples that on the host (the reference's own path, batched).
class WasmSDF : public SDFSurface {
   public:
    WasmSDF(const void* wasm, size_t wasm_bytes, uint32_t sdf_id = 0) {
        char log[1024];
        size_t need = 0;
        float bb[6];
        int rc = sdfgpu_wasm_lower(wasm, wasm_bytes, sdf_id, nullptr, 0, &need, bb, log, sizeof log);
        if (rc != SDFGPU_OK) throw Error(rc, log);
        tape_.resize(need);
        rc = sdfgpu_wasm_lower(wasm, wasm_bytes, sdf_id, tape_.data(), tape_.size(), &need, bb, log, sizeof log);
        if (rc != SDFGPU_OK) throw Error(rc, log);
        bb_ = {Vector3{bb[0], bb[1], bb[2]}, Vector3{bb[3], bb[4], bb[5]}};
        summary = log;
    }
    BoundingBox bounding_box() const override { return bb_; }
    SDFSample sample(Vector3, bool) const override {
        throw std::logic_error("sdfgpu::WasmSDF is evaluated on the GPU through tape(); it has no host sample()");
    }
    std::optional<std::vector<unsigned char>> tape() const override { return tape_; }
    std::string summary;  // "lowered: N scalar ops, ..."

   private:
    std::vector<unsigned char> tape_;
    BoundingBox bb_;
};

}  // namespace sdfgpu

#endif  // SDFGPU_VIEWER_HPP

// host_viewer.cpp -- the C++ host side (include/sdfgpu_viewer.hpp) driven the way the reference's
// scene drives SDFViewer (/root/reference/src/app/scene/mod.rs:154-215), plus the reference's own
// LoadingManager tests (src/app/scene/sdf/loading.rs:117-171) against sdfgpu::LoadingManager.
//
//   host_viewer loading          CPU: the five loading.rs cases
//   host_viewer tape             CPU: hex dump of sdfgpu::SDFDemo's tape (compared with tape.py's)
//   host_viewer scalar           CPU: hex dump of a scalar-program tape built with sdfgpu::ScalarProgram
//   host_viewer wasm <file>      CPU: sdfgpu::WasmSDF lowers a WebAssembly SDF; prints the summary, bounding box, tape
//   host_viewer gpu <out_dir>    GPU: (1) the demo surface with a tape, (2) a surface WITHOUT a tape whose
//                                sample() is the oracle's SDFDemo::sample; volumes and a frame are
//                                written to <out_dir> for the Python test to compare with the oracle
//
// TEST CODE: links liboracle.so as the stand-in for a user's CPU SDF.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sdfgpu_viewer.hpp"

extern "C" {  // oracle/sdf_oracle.cpp
struct OrcDemoParams {
    float cube_half_side; uint32_t cube_material; float sphere_radius; uint32_t sphere_material;
    float max_distance_custom_material; uint32_t disable_sphere;
};
void orc_demo_params_default(OrcDemoParams* p);
void orc_demo_sample(const OrcDemoParams* p, const float* pts, uint64_t n, int distance_only, float* out);
}

#define REQUIRE(c)                                                                  \
    do {                                                                            \
        if (!(c)) { std::fprintf(stderr, "%s:%d: REQUIRE(%s) failed\n", __FILE__, __LINE__, #c); std::exit(1); } \
    } while (0)

// loading.rs:122-146 `test_loading_manager`
static void loading_case(std::array<uint32_t, 3> limits) {
    const uint32_t num_passes = 3;
    sdfgpu::LoadingManager lm(limits, num_passes);
    std::vector<int> hits((size_t)limits[0] * limits[1] * limits[2], 0);
    uint64_t iterations = 0;
    const uint64_t total = iterations + lm.len();
    while (auto v = lm.next()) {
        const size_t flat = (*v)[0] + (size_t)(*v)[1] * limits[0] + (size_t)(*v)[2] * limits[0] * limits[1];
        hits[flat] += 1;
        REQUIRE(hits[flat] <= (int)num_passes);
        iterations += 1;
        REQUIRE(total == iterations + lm.len());
    }
    for (int h : hits) REQUIRE(h >= 1);
    REQUIRE(lm.passes_left() == 0 && lm.total_iterations() == iterations);
}

// a user's SDFSurface without a tape: sample() is CPU code the library knows nothing about
class HostDemo : public sdfgpu::SDFSurface {
   public:
    HostDemo() { orc_demo_params_default(&params); }
    sdfgpu::BoundingBox bounding_box() const override { return {sdfgpu::Vector3{-1, -1, -1}, sdfgpu::Vector3{1, 1, 1}}; }
    sdfgpu::SDFSample sample(sdfgpu::Vector3 p, bool distance_only) const override {
        sdfgpu::SDFSample s;
        const float xyz[3] = {p.x, p.y, p.z};
        orc_demo_sample(&params, xyz, 1, distance_only, reinterpret_cast<float*>(&s));
        return s;
    }
    unsigned sample_threads() const override { return threads; }
    OrcDemoParams params;
    unsigned threads = 1;
};

static void write_file(const std::string& path, const void* data, size_t bytes) {
    FILE* f = std::fopen(path.c_str(), "wb");
    REQUIRE(f != nullptr);
    REQUIRE(std::fwrite(data, 1, bytes, f) == bytes);
    std::fclose(f);
}

// SDFViewerAppScene::set_sdf + the load loop of render() (scene/mod.rs:154-155,168-199)
static size_t load(sdfgpu::SDFViewer& viewer, const sdfgpu::SDFSurface& sdf, std::chrono::milliseconds per_frame,
                   size_t* frames) {
    size_t updates = 0;
    *frames = 0;
    for (;;) {
        const size_t cpu_updates = viewer.update(sdf, per_frame);
        ++*frames;
        updates += cpu_updates;
        if (cpu_updates == 0) break;
        viewer.commit();
    }
    viewer.commit();
    return updates;
}

static int run_gpu(const std::string& out) {
    // (1) a surface with a tape: every pass is one kernel
    {
        sdfgpu::SDFDemo sdf;
        auto viewer = sdfgpu::SDFViewer::from_bb(sdf.bounding_box(), 32, 2);
        REQUIRE(viewer.width() == 32 && viewer.height() == 32 && viewer.depth() == 32);
        const uint64_t total = viewer.loading_mgr().len();
        REQUIRE(total == 32 * 32 * 32 + 16 * 16 * 16);
        size_t frames = 0;
        REQUIRE(load(viewer, sdf, std::chrono::milliseconds(30), &frames) == total);
        REQUIRE(viewer.loading_mgr().len() == 0 && viewer.loading_mgr().passes_left() == 0);
        REQUIRE(viewer.loading_mgr().passes == 2 && viewer.loading_mgr().total_iterations() == total);
        std::vector<float> t0, t1;
        viewer.download(&t0, &t1);
        write_file(out + "/tape_tex0.bin", t0.data(), t0.size() * 4);
        write_file(out + "/tape_tex1.bin", t1.data(), t1.size() * 4);
        const sdfgpu::Frame f = viewer.render(sdfgpu::SDFViewer::default_camera(160, 120), 160, 120);
        write_file(out + "/tape_rgba8.bin", f.rgba8.data(), f.rgba8.size());
        write_file(out + "/tape_depth.bin", f.depth.data(), f.depth.size() * 4);
        // a parameter edit: changed() reports the whole box once, the viewer re-samples in 3 passes
        sdf.sphere_radius = 0.9f;
        sdf.mark_changed();
        REQUIRE(load(viewer, sdf, std::chrono::milliseconds(30), &frames) > 0);
        viewer.download(&t0, &t1);
        write_file(out + "/tape_changed_tex0.bin", t0.data(), t0.size() * 4);
        write_file(out + "/tape_changed_tex1.bin", t1.data(), t1.size() * 4);
    }
    // (2) a surface without a tape: sampled on the host in the reference's order, a few ms per frame
    {
        HostDemo sdf;
        sdf.threads = 4;
        auto viewer = sdfgpu::SDFViewer::new_voxels({24, 20, 16}, sdf.bounding_box(), 3);
        const uint64_t total = viewer.loading_mgr().len();
        size_t frames = 0;
        REQUIRE(load(viewer, sdf, std::chrono::milliseconds(1), &frames) == total);
        REQUIRE(frames >= 2);
        std::vector<float> t0, t1;
        viewer.download(&t0, &t1);
        write_file(out + "/host_tex0.bin", t0.data(), t0.size() * 4);
        write_file(out + "/host_tex1.bin", t1.data(), t1.size() * 4);
        std::printf("host-sampled load: %zu iterations in %zu frames\n", (size_t)total, frames);
    }
    // (3) error behaviour: an exception in the user's sample() surfaces from update(), nothing aborts
    {
        struct Broken : HostDemo {
            sdfgpu::SDFSample sample(sdfgpu::Vector3, bool) const override { throw std::runtime_error("guest trapped"); }
        } sdf;
        auto viewer = sdfgpu::SDFViewer::new_voxels({4, 4, 4}, sdf.bounding_box(), 1);
        bool thrown = false;
        try {
            viewer.update(sdf, std::chrono::seconds(10));
        } catch (const std::runtime_error& e) {
            thrown = std::string(e.what()) == "guest trapped";
        }
        REQUIRE(thrown);
        bool rejected = false;
        try {
            sdfgpu::SDFViewer::from_bb(sdf.bounding_box(), 8, 1, 4096);  // no such device
        } catch (const sdfgpu::Error& e) {
            rejected = e.code() == SDFGPU_ERR_INVALID;
        }
        REQUIRE(rejected);
    }
    // (4) the same viewer as a GROUP of slabs (here three on device 0; one per GPU on a multi-GPU host): one thread drives
    // all of them, and volumes and frame equal the single viewer's bit for bit
    {
        sdfgpu::SDFDemo sdf, sdf2;
        auto one = sdfgpu::SDFViewer::new_voxels({40, 36, 30}, sdf.bounding_box(), 2);
        auto group = sdfgpu::SDFViewerGroup::new_voxels({40, 36, 30}, sdf.bounding_box(), 2, {0, 0, 0}, 160, 120);
        REQUIRE(group.size() == 3 && group.voxels() == one.voxels());
        size_t frames = 0;
        const size_t total = load(one, sdf, std::chrono::milliseconds(30), &frames);
        size_t got = 0;
        while (group.loading_len() > 0) {
            got += group.update(sdf2, std::chrono::milliseconds(30));
            group.commit();
        }
        REQUIRE(got == total);
        std::vector<float> a0, a1, b0, b1;
        one.download(&a0, &a1);
        group.download(&b0, &b1);
        REQUIRE(a0.size() == b0.size() && std::memcmp(a0.data(), b0.data(), a0.size() * 4) == 0);
        REQUIRE(std::memcmp(a1.data(), b1.data(), a1.size() * 4) == 0);
        const sdfgpu_camera cam = sdfgpu::SDFViewer::default_camera(160, 120);
        const sdfgpu::Frame fa = one.render(cam, 160, 120), fb = group.render(cam, 160, 120);
        REQUIRE(fa.rgba8 == fb.rgba8);
        REQUIRE(std::memcmp(fa.depth.data(), fb.depth.data(), fa.depth.size() * 4) == 0);
        auto solo = sdfgpu::SDFViewerGroup::from_bb(sdf.bounding_box(), 16, 1, 1u, 64, 48);  // a mask of one device
        REQUIRE(solo.size() == 1);
        std::printf("group of %u slabs == one viewer\n", group.size());
    }
    std::printf("gpu ok\n");
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "loading") {
        loading_case({2, 2, 2});     // loading.rs:147-170
        loading_case({8, 8, 8});
        loading_case({64, 64, 64});
        loading_case({11, 11, 11});
        loading_case({8, 11, 17});
        std::printf("loading ok\n");
        return 0;
    }
    if (mode == "tape") {
        sdfgpu::SDFDemo sdf;
        if (argc > 2) sdf.disable_sphere = true;
        const auto t = *sdf.tape();
        for (unsigned char c : t) std::printf("%02x", c);
        std::printf("\n");
        return 0;
    }
    if (mode == "gpu" && argc > 2) {
        try {
            return run_gpu(argv[2]);
        } catch (const std::exception& e) {
            std::fprintf(stderr, "exception: %s\n", e.what());
            return 2;
        }
    }
    if (mode == "scalar") {  // CPU: the banded sphere of tests/test_scalar_programs.py written with sdfgpu::ScalarProgram
        sdfgpu::ScalarProgram p;
        const uint32_t x = p.px(), y = p.py(), z = p.pz();
        // one statement per op: C++ does not order the evaluation of nested call arguments
        const uint32_t xx = p.op(SDFT_S_FMUL, x, x), yy = p.op(SDFT_S_FMUL, y, y);
        const uint32_t s2 = p.op(SDFT_S_FADD, xx, yy);
        const uint32_t zz = p.op(SDFT_S_FMUL, z, z);
        const uint32_t len = p.op(SDFT_S_FSQRT, p.op(SDFT_S_FADD, s2, zz));
        const uint32_t radius = p.constant(0.7f);
        const uint32_t d = p.op(SDFT_S_FSUB, len, radius);
        const uint32_t eight = p.constant(8.0f);
        const uint32_t cell = p.op(SDFT_S_I_FROM_F_S, p.op(SDFT_S_FFLOOR, p.op(SDFT_S_FMUL, y, eight)));
        const uint32_t one = p.imm(1);
        const uint32_t band = p.op(SDFT_S_IAND, cell, one);
        p.out(0, d);
        const uint32_t hi = p.constant(0.9f), lo = p.constant(0.1f);
        p.out(1, p.op(SDFT_S_SELECT, band, hi, lo));
        p.out(2, p.op(SDFT_S_FABS, x));
        const uint32_t zero = p.constant(0.0f);
        const uint32_t zc = p.op(SDFT_S_FMAX, z, zero);
        const uint32_t fone = p.constant(1.0f);
        p.out(3, p.op(SDFT_S_FMIN, zc, fone));
        p.out(4, p.constant(0.25f));
        p.out(5, p.constant(0.5f));
        p.out(6, p.constant(1.0f));
        sdfgpu::TapeBuilder t;
        t.scalar(p).emit(SDFT_OP_END);
        const std::vector<unsigned char> tape = t.build();
        REQUIRE(sdfgpu_tape_validate(tape.data(), tape.size()) == SDFGPU_OK);
        for (unsigned char c : tape) std::printf("%02x", c);
        std::printf("\n");
        return 0;
    }
    if (mode == "wasm" && argc > 2) {  // CPU: lower a .wasm SDF through sdfgpu::WasmSDF, print its tape as hex
        FILE* f = std::fopen(argv[2], "rb");
        REQUIRE(f != nullptr);
        std::vector<unsigned char> bytes;
        for (int c; (c = std::fgetc(f)) != EOF;) bytes.push_back((unsigned char)c);
        std::fclose(f);
        try {
            sdfgpu::WasmSDF sdf(bytes.data(), bytes.size());
            const auto bb = sdf.bounding_box();
            std::printf("%s\n%g %g %g %g %g %g\n", sdf.summary.c_str(), bb[0].x, bb[0].y, bb[0].z, bb[1].x, bb[1].y, bb[1].z);
            const std::vector<unsigned char> tape = *sdf.tape();
            for (unsigned char c : tape) std::printf("%02x", c);
            std::printf("\n");
            return 0;
        } catch (const sdfgpu::Error& e) {
            std::printf("cannot lower (%d): %s\n", e.code(), e.what());
            return 3;
        }
    }
    std::fprintf(stderr, "usage: host_viewer loading | tape [nosphere] | scalar | wasm <file.wasm> | gpu <out_dir>\n");
    return 64;
}

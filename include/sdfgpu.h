/*
 * sdfgpu.h -- C ABI of libsdfgpu.so: the B200 (sm_100a) grid-fill and
 * sphere-trace hot path of sdf-viewer, behind plain pointers and sizes.
 *
 * Each entry point cites the reference interface it replaces; paths are
 * relative to /root/reference.  The library plugs in at the seam between
 * `SDFViewerAppScene::render` (src/app/scene/mod.rs:158-225) and
 * `SDFViewer` / `SDFViewerMaterial` (src/app/scene/sdf/{mod,material}.rs).
 *
 * Conventions
 *  - every function returns SDFGPU_OK (0) or a negative sdfgpu_status; it
 *    never aborts or throws (reference: a failing guest is logged and a benign
 *    value returned, src/sdf/wasm/native.rs:196-203).  The message is kept per
 *    handle (sdfgpu_last_error) or per thread when there is no handle.
 *  - volumes are `[f32;4]` texels, x fastest, then y, then z
 *    (flat = (z*H + y)*W + x, src/app/scene/sdf/mod.rs:177).
 *  - compute entry points enqueue on the handle's CUDA stream and return;
 *    entry points that fill HOST memory synchronise that stream first.
 *  - a handle is used from one thread at a time; distinct handles are independent.
 *  - there is NO CPU fallback: without a CUDA device every call fails with
 *    SDFGPU_ERR_CUDA.
 */
#ifndef SDFGPU_H
#define SDFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdfgpu_ctx sdfgpu_ctx;
struct sdfgpu_camera;  /* defined below (trace) */
struct sdfgpu_surface; /* defined below (update from an SDFSurface) */

typedef enum sdfgpu_status {
    SDFGPU_OK = 0,
    SDFGPU_ERR_INVALID = -1, /* bad argument                                  */
    SDFGPU_ERR_CUDA = -2,    /* CUDA runtime/driver error, or no device       */
    SDFGPU_ERR_TAPE = -3,    /* malformed tape                                */
    SDFGPU_ERR_STATE = -4    /* call order (e.g. update before set_tape)      */
} sdfgpu_status;

/* Uncomputed-voxel marker, f32(0.1) + f32(0.001234) (src/app/scene/sdf/mod.rs:42). */
float sdfgpu_air_dist(void);

/* ------------------------------------------------------------------ create */

/* SDFViewer::from_bb (src/app/scene/sdf/mod.rs:46-72): the longest bbox axis
 * gets `max_voxels_side` voxels, the others `(max * size_i / size_max) as usize`.
 * bb = {min.x,min.y,min.z,max.x,max.y,max.z}.  Both volumes start as AIR_DIST
 * in all four channels (:76-77). */
int sdfgpu_create(const float bb[6], uint32_t max_voxels_side, uint32_t loading_passes,
                  int device, sdfgpu_ctx** out);

/* The voxel-count rule of from_bb alone (src/app/scene/sdf/mod.rs:47-68); no device needed. */
int sdfgpu_dims_from_bb(const float bb[6], uint32_t max_voxels_side, uint32_t out_voxels[3]);

/* SDFViewer::new_voxels (src/app/scene/sdf/mod.rs:75-101). */
int sdfgpu_create_voxels(const float bb[6], const uint32_t voxels[3], uint32_t loading_passes,
                         int device, sdfgpu_ctx** out);

/* Multi-GPU: this handle owns only z in [z_begin, z_end) of the global grid
 * `voxels`, plus one halo slice on each interior face.  Not in the reference
 * (single process, src/app/scene/sdf/mod.rs:174 "TODO: parallel iteration"). */
int sdfgpu_create_slab(const float bb[6], const uint32_t voxels[3], uint32_t loading_passes,
                       int device, uint32_t z_begin, uint32_t z_end, sdfgpu_ctx** out);

/* (Earlier, lower-level form of the halo exchange, kept for hosts that order the ranks themselves; linked handles
 * -- below -- supersede it.)  Fused halo exchange between slab handles that live in different processes on one node (one
 * process per GPU): a handle exports CUDA IPC handles of its two volumes (2 x 64 bytes:
 * cudaIpcMemHandle_t of tex0, then of tex1); its neighbours attach them (side 0 = the neighbour
 * below this handle, 1 = above; [peer_z_lo, peer_z_hi) = the neighbour's stored slices).  From then
 * on sdfgpu_fill_all fills the two boundary slices first and lets the copy engines push them
 * into the neighbours' halo slices over NVLink while the interior is being filled; update /
 * resample_box push them after their last pass.  The caller orders fills and traces across ranks
 * (a stream-ordered barrier).  Not in the reference. */
int sdfgpu_ipc_export(sdfgpu_ctx* ctx, void* handles, size_t handles_bytes);
int sdfgpu_ipc_attach(sdfgpu_ctx* ctx, int side, const void* handles, size_t handles_bytes,
                      uint32_t peer_z_lo, uint32_t peer_z_hi);
int sdfgpu_ipc_detach(sdfgpu_ctx* ctx);

/* ------------------------------------------------------- multi-GPU: linked slabs
 *
 * Z-sharded operation of slab handles with no collective library and no host synchronisation between ranks.  Not in
 * the reference (one process, one thread: src/app/scene/mod.rs:22-31,158-225).  Two ways in:
 *
 *  (1) ONE PROCESS drives every GPU (what a drop-in for the reference's single-threaded scene needs): sdfgpu_group_*
 *      below -- sdfgpu_group_create_mask(bb, max_voxels_side, passes, device_mask, ...) is SDFViewer::from_bb over the
 *      devices of the mask.
 *  (2) ONE PROCESS PER GPU: every rank creates its slab (sdfgpu_create_slab: the slabs tile [0, D) in rank order, none
 *      empty), calls sdfgpu_link_export, the SDFGPU_LINK_BLOB_BYTES blobs of all ranks are gathered by any means (they
 *      hold CUDA IPC handles: same node only), and every rank calls sdfgpu_link_attach with all of them.
 *
 * Linking maps every rank's arena (flags, ray queues, the presenter's frame) -- and, with SDFGPU_LINK_HALO_PUSH, the
 * neighbours' volumes -- into this handle's address space.  From then on these entry points are COLLECTIVE -- every rank calls them in the same order
 * with the same arguments: sdfgpu_fill_all, sdfgpu_update, sdfgpu_update_surface, sdfgpu_resample_box, sdfgpu_reset,
 * sdfgpu_commit, sdfgpu_trace_rgba8 / sdfgpu_trace_linked.  Options read by sdfgpu_link_export: "link_halo_push",
 * "link_trace_mode" (0 auto, 1 rounds, 2 stream), "link_timeout_ms" (how long a streaming trace waits for a neighbour).
 *   fill:  no exchange.  Every rank fills its stored slices, the one halo slice per interior face included:
 *          SDFSurface::sample is a pure function of the position between changed() events (src/sdf/mod.rs:43), so the
 *          halo slices equal the neighbours' boundary slices bit for bit and cost 2 of D/world slices of extra work.
 *          With SDFGPU_LINK_HALO_PUSH the neighbours push them instead: one launch, the tiles of the first and last own
 *          slice go first, the kernel releases a flag when they are complete, and the copy engines push them into the
 *          neighbours' halo slices over NVLink while the interior is still being filled; update / resample_box push
 *          after their last pass, and only the faces a dirty box touches.  (Measured: the pushes are slow while both
 *          HBMs are saturated by the fills -- DESIGN.md section 5.)
 *   trace: exact.  A ray marches on the rank that owns the lower z tap of its texture fetch and is handed to the
 *          neighbour (position, t, step count) when it leaves that rank's slices.  With every rank on a device of its
 *          own this is ONE kernel per rank and frame: its persistent warps trace the rank's own rays, then poll the
 *          in-queue the neighbours store into over NVLink, so the ranks form a pipeline along z.  Ranks that share a
 *          device (tests), or SDFGPU_LINK_ROUNDS, take turns instead: `world` rounds of one kernel.  Either way the
 *          frame equals sdfgpu_trace_rgba8 of ONE handle holding the whole grid bit for bit (RGBA8, depth, and the
 *          G-buffer but for the normals of hits next to a slab face).  Finished pixels are stored straight into the frame
 *          of rank 0 (the presenter), which alone receives rgba8 / depth / gbuf; the other ranks pass NULL.
 * A streaming trace whose neighbour does not take part within "link_timeout_ms" fails with SDFGPU_ERR_STATE on every
 * rank that waited; the link's queues are then in an undefined state: detach and link again.  sdfgpu_cull_stats runs a
 * fill and is therefore collective on a linked handle, too.
 * Link a handle right after creating it (volumes still AIR_DIST), use frames of at most max_width * max_height pixels,
 * and detach every rank before destroying any of them.  SDFGPU_LINK_GBUF reserves a G-buffer frame (tests). */
#define SDFGPU_LINK_BLOB_BYTES 320
#define SDFGPU_LINK_GBUF 1u
#define SDFGPU_LINK_HALO_PUSH 2u /* halo slices pushed by the neighbours (below) instead of filled by their holder */
#define SDFGPU_LINK_ROUNDS 4u    /* trace in `world` rounds even when every rank has a device of its own */
#define SDFGPU_LINK_STREAM 8u    /* insist on the one-launch trace: attach fails when ranks share a device */
int sdfgpu_link_export(sdfgpu_ctx* ctx, uint32_t rank, uint32_t world, uint32_t max_width, uint32_t max_height,
                       uint32_t flags, void* blob, size_t blob_bytes);
int sdfgpu_link_attach(sdfgpu_ctx* ctx, const void* blobs, uint32_t world);
int sdfgpu_link_detach(sdfgpu_ctx* ctx);
/* The collective trace of a linked handle (sdfgpu_trace_rgba8 forwards here with want_gbuf = 0).  want_gbuf must be
 * the same on every rank; outputs are written on rank 0 only (any may be NULL). */
int sdfgpu_trace_linked(sdfgpu_ctx* ctx, const struct sdfgpu_camera* cam, uint32_t width, uint32_t height, int want_gbuf,
                        uint8_t* rgba8, float* depth, float* gbuf);
/* The same, enqueued only (no host synchronisation, the counterpart of sdfgpu_trace_device): the frame stays in the
 * presenter's HBM; rank 0 receives device pointers valid until its next trace, the other ranks NULL. */
int sdfgpu_trace_linked_device(sdfgpu_ctx* ctx, const struct sdfgpu_camera* cam, uint32_t width, uint32_t height,
                               int want_gbuf, uint8_t** rgba8_dev, float** depth_dev, float** gbuf_dev);

/* The single-process form: a group owns one linked slab handle per device and runs every collective call on all of
 * them (round by round for the trace).  `devices` may name a device more than once (the tests shard a grid over
 * several handles of one GPU).  Mirrors the SDFViewer surface: from_bb / new_voxels, update, commit, the volumes,
 * and the frame of `volume.render` (scene/mod.rs:168-215). */
typedef struct sdfgpu_group sdfgpu_group;
int sdfgpu_group_create(const float bb[6], const uint32_t voxels[3], uint32_t loading_passes, const int* devices,
                        uint32_t n_devices, uint32_t max_width, uint32_t max_height, uint32_t flags, sdfgpu_group** out);
int sdfgpu_group_create_mask(const float bb[6], uint32_t max_voxels_side, uint32_t loading_passes, uint32_t device_mask,
                             uint32_t max_width, uint32_t max_height, sdfgpu_group** out);
void sdfgpu_group_destroy(sdfgpu_group* g);
uint32_t sdfgpu_group_size(const sdfgpu_group* g);
sdfgpu_ctx* sdfgpu_group_rank(sdfgpu_group* g, uint32_t rank); /* a rank's handle: options, info, its own slices */
const char* sdfgpu_group_last_error(const sdfgpu_group* g);
int sdfgpu_group_set_tape(sdfgpu_group* g, const void* tape, size_t tape_bytes);
int sdfgpu_group_update(sdfgpu_group* g, const float* changed_box, uint32_t max_passes, uint64_t* iterations);
int sdfgpu_group_update_surface(sdfgpu_group* g, const struct sdfgpu_surface* sdf, double max_delta_seconds,
                                uint64_t* iterations);
int sdfgpu_group_fill_all(sdfgpu_group* g);
int sdfgpu_group_resample_box(sdfgpu_group* g, const float box[6], uint64_t* voxels_touched);
int sdfgpu_group_commit(sdfgpu_group* g);
int sdfgpu_group_reset(sdfgpu_group* g, uint32_t loading_passes);
int sdfgpu_group_set_option(sdfgpu_group* g, const char* key, int64_t value);
int sdfgpu_group_sync(sdfgpu_group* g);
int sdfgpu_group_loading_state(const sdfgpu_group* g, uint64_t* len, uint64_t* total_iterations,
                               uint32_t* passes_left, uint32_t* passes);
int sdfgpu_group_download(sdfgpu_group* g, float* tex0, float* tex1); /* the whole grid, W*H*D texels each */
int sdfgpu_group_trace_rgba8(sdfgpu_group* g, const struct sdfgpu_camera* cam, uint32_t width, uint32_t height,
                             uint8_t* rgba8, float* depth);
int sdfgpu_group_trace(sdfgpu_group* g, const struct sdfgpu_camera* cam, uint32_t width, uint32_t height,
                       uint8_t* rgba8, float* depth, float* gbuf); /* gbuf needs SDFGPU_LINK_GBUF at creation */

/* ------------------------------------------------------- sampling at arbitrary positions, mesh export
 *
 * sdfgpu_sample_points: SDFSurface::sample(p, false) (src/sdf/mod.rs:43) of the current tape for n positions
 * (xyz: n x 3 floats) -> out: n x 7 floats (distance, r, g, b, metallic, roughness, occlusion: the SDFSample of
 * src/sdf/mod.rs:104-118, raw -- none of the volume's store rules applied).
 *
 * sdfgpu_mesh: the `mesh` subcommand's pipeline (src/sdf/meshers/mod.rs:66-87) on the GPU, from the volume that is
 * ALREADY resident instead of re-sampling the SDF: marching cubes over the cells of the lattice (the sign of
 * tex0.r - 0.1, material.frag:56-60; an SDFViewer of N + 1 voxels per axis holds exactly the (N + 1)^3 samples the
 * reference's MarchingCubes::new(N) takes, meshers/isosurface.rs:26-29,94-98), one vertex per sign-changing lattice
 * edge (shared by the triangles around it), then Mesh::postproc (meshers/mesh.rs:22-33) through the tape: colour,
 * metallic, roughness, occlusion = sample(vertex, false); normal = SDFSurface::normal(vertex, None)
 * (src/sdf/defaults.rs:49-56).  Vertex records are 12 floats -- position, normal, colour, metallic, roughness,
 * occlusion (mesh.rs Vertex) -- in the SDF's coordinates; triangles are counter-clockwise seen from outside.  The
 * volume must be fully loaded and the handle must hold the whole grid.  The `isosurface` crate is not vendored with
 * the reference: vertex order and the triangulation of ambiguous cells are this library's own (watertight by
 * construction, see mc_table.py).
 *
 * sdfgpu_ply_serialize / sdfgpu_mesh_write_ply: Mesh::serialize_ply (meshers/mesh.rs:38-129): ASCII PLY with the
 * reference's element / property list, floats in Rust's shortest round-trip `{}` form, colours (c * 255.9999) as u8;
 * `comment` is the header's comment line (the reference writes "Created with <version>"). */
int sdfgpu_sample_points(sdfgpu_ctx* ctx, const float* xyz, uint64_t n, float* out);
int sdfgpu_mesh(sdfgpu_ctx* ctx, uint64_t* n_vertices, uint64_t* n_triangles);
int sdfgpu_mesh_download(sdfgpu_ctx* ctx, float* vertices, uint32_t* indices);
int sdfgpu_mesh_device_ptrs(sdfgpu_ctx* ctx, const float** vertices_dev, const uint32_t** indices_dev);
int sdfgpu_mesh_write_ply(sdfgpu_ctx* ctx, const char* path, const char* comment, uint64_t* bytes_written);
int sdfgpu_ply_serialize(const float* vertices, uint64_t n_vertices, const uint32_t* indices, uint64_t n_triangles,
                         const char* comment, const char* path, uint64_t* bytes_written);

/* Dropping the SDFViewer (scene/mod.rs:154-155 rebuilds it on every set_sdf). */
void sdfgpu_destroy(sdfgpu_ctx* ctx);

const char* sdfgpu_last_error(const sdfgpu_ctx* ctx); /* ctx may be NULL: per-thread message */

/* tex0.width/height/depth (read by scene/mod.rs:149-150). */
int sdfgpu_dims(const sdfgpu_ctx* ctx, uint32_t out_voxels[3]);
/* Stored z range of this handle including halo slices: [z_lo, z_hi). */
int sdfgpu_slab(const sdfgpu_ctx* ctx, uint32_t* z_begin, uint32_t* z_end,
                uint32_t* z_lo, uint32_t* z_hi);

/* -------------------------------------------------------------------- fill */

/* Replaces the `sdf: impl SDFSurface` argument of SDFViewer::update
 * (src/app/scene/sdf/mod.rs:128): the SDF as a tape (sdfgpu_tape.h).  Copies
 * the bytes; may be called again when a parameter changes
 * (SDFSurface::set_parameter, src/sdf/mod.rs:73). */
int sdfgpu_set_tape(sdfgpu_ctx* ctx, const void* tape, size_t tape_bytes);

/* Device-free validation of a tape (what sdfgpu_set_tape checks before it touches the device): header, section
 * sizes, operand ranges, stack balance, scalar programs.  SDFGPU_OK or SDFGPU_ERR_TAPE / _INVALID with the
 * reason in sdfgpu_last_error(NULL). */
int sdfgpu_tape_validate(const void* tape, size_t tape_bytes);

/* Device-free check of the specialiser: validates `tape`, lowers it and compiles the straight-line
 * fill kernel for its structure with NVRTC for sm_100a (what sdfgpu_set_tape + the first fill do on
 * a GPU box).  On success `log` receives the generated translation unit, on failure the compiler
 * log.  Returns SDFGPU_ERR_CUDA when NVRTC is not installed. */
int sdfgpu_jit_check(const void* tape, size_t tape_bytes, int voxels_per_thread, char* log, size_t log_cap);

/* Lowers the `sample` export of a WebAssembly SDF module -- the guest ABI of src/sdf/wasm/mod.rs:5-37 that
 * src/sdf/wasm/native.rs loads -- to a tape, so that an existing .wasm SDF runs on the GPU instead of being
 * called once per voxel.  Device free.  The module is instantiated (start function, optional `init()`,
 * native.rs:52-56), `bounding_box(sdf_id)` is executed and its six floats returned in bb_out, and
 * `sample(sdf_id, x, y, z, 0)` is executed once with SYMBOLIC coordinates: integer work, addresses, allocator
 * and registry look-ups run concretely and vanish, every f32 / i32 operation on a value that depends on the
 * position becomes one op of a scalar program (sdfgpu_tape.h), and branches on such values are merged into
 * selects.  The result is the tape {SDFT_OP_SCALAR, SDFT_OP_END}; its constants are runtime data, so lowering
 * the module again after a parameter change re-uses the compiled kernel.
 *   tape_out / tape_cap : receives the tape; *tape_len = bytes needed (call with tape_out = NULL to size it)
 *   log                 : on success a one-line summary, on failure why this module cannot be lowered
 * Returns SDFGPU_OK; SDFGPU_ERR_INVALID for a malformed module or one without the required exports
 * (native.rs:59-63); SDFGPU_ERR_TAPE when the guest does something that has no tape form (an address, loop
 * bound or call target that depends on the position; 64-bit arithmetic on such values; a host import that returns a
 * value -- imports that return nothing, e.g. logging hooks, are skipped and WASI calls answered with errno 0; SIMD) --
 * such an SDF is sampled on the host through sdfgpu_update_surface instead. */
int sdfgpu_wasm_lower(const void* wasm, size_t wasm_bytes, uint32_t sdf_id, void* tape_out, size_t tape_cap,
                      size_t* tape_len, float bb_out[6], char* log, size_t log_cap);

/* The same for a LIVE instance: `memory` is a copy of the guest's linear memory as the host's runtime holds it
 * now (wasmer: `memory.view(&store)`), i.e. after `init()` and after whatever `set_parameter` calls were made
 * (src/sdf/wasm/native.rs:386-460 forwards them to the guest, which keeps its parameters in its own memory).
 * It replaces instantiation, so the tape reflects the current parameters; after the guest's `changed()`
 * reports a box (native.rs:463-483), lower again and hand the new tape to the viewer -- same structure, new
 * constants, same compiled kernel.  memory_bytes must be a multiple of 65536. */
int sdfgpu_wasm_lower_live(const void* wasm, size_t wasm_bytes, const void* memory, size_t memory_bytes,
                           uint32_t sdf_id, void* tape_out, size_t tape_cap, size_t* tape_len, float bb_out[6],
                           char* log, size_t log_cap);

/* SDFViewer::update (src/app/scene/sdf/mod.rs:128-217).
 *   changed_box : result of sdf.changed() this frame ({min,max}), or NULL for None;
 *                 merged into the pending box as :131-139 does.
 *   max_passes  : how many LoadingManager passes to run at most (0 = all that
 *                 are pending); replaces max_delta_time -- a pass is one kernel.
 *   iterations  : out, LoadingManager iterations performed (the return value of update).
 * State machine (pending box, 3-pass re-sample, changed_box_while_loading)
 * follows :141-154.  A voxel is (re)sampled iff tex0.r == AIR_DIST or its
 * position lies in the pending box (:184-190); stored values follow :196-208. */
int sdfgpu_update(sdfgpu_ctx* ctx, const float* changed_box, uint32_t max_passes,
                  uint64_t* iterations);

/* One-kernel full fill: the state the reference reaches once its
 * LoadingManager is exhausted on a fresh SDFViewer, without the per-pass
 * AIR_DIST reads.  Marks loading as finished. */
int sdfgpu_fill_all(sdfgpu_ctx* ctx);

/* Re-sample only the voxels whose position lies in `box` (closed interval,
 * float compare, src/app/scene/sdf/mod.rs:187-189): the 60 Hz parameter-sweep
 * path.  Launches over the index AABB of the box only. */
int sdfgpu_resample_box(sdfgpu_ctx* ctx, const float box[6], uint64_t* voxels_touched);

/* Batched ingest for SDFs that have no tape (any existing .wasm SDFSurface; the reference's own
 * "TODO: Batched sampling", src/sdf/mod.rs:39): the host evaluates sample(p, false) itself for a run
 * of voxels and hands the raw results over; the GPU applies the store rules of
 * src/app/scene/sdf/mod.rs:196-208 and writes both volumes.
 *   sdfgpu_voxel_positions: positions of the voxels [first_flat, first_flat + count) in flat order
 *     (flat = (z*H + y)*W + x, :177), computed exactly as :179-182 does; xyz = count * 3 floats.
 *   sdfgpu_ingest_samples: `samples` = count SDFSample records (7 floats = 28 bytes each:
 *     distance, r, g, b, metallic, roughness, occlusion; src/sdf/mod.rs:104-118 -- the bytes the host
 *     reads out of WASM memory, src/sdf/wasm/native.rs:204-216) for the same voxels. */
int sdfgpu_voxel_positions(const sdfgpu_ctx* ctx, uint64_t first_flat, uint64_t count, float* xyz);
int sdfgpu_ingest_samples(sdfgpu_ctx* ctx, uint64_t first_flat, uint64_t count, const void* samples);

/* ------------------------------------------------ update from an SDFSurface */

/* The `SDFSurface` trait (src/sdf/mod.rs:33-101) as a table of C callbacks -- the same four
 * functions the WASM ABI exposes (src/sdf/wasm/mod.rs:5-37; host side src/sdf/wasm/native.rs:163-217,
 * 463-483), so a host that already holds a `Box<dyn SDFSurface>` (a WasmerSDF, the built-in demo,
 * anything else) passes thin trampolines.  `self` is handed back to every callback.
 *   bounding_box : SDFSurface::bounding_box (:37), {min.xyz, max.xyz}; may be NULL (update never calls it)
 *   sample       : SDFSurface::sample (:43) for one point; out = the 7 floats of SDFSample
 *                  (:104-118: distance, r, g, b, metallic, roughness, occlusion)
 *   sample_batch : optional batched form (the reference's "TODO: Batched sampling", :39): n points
 *                  xyz[3n] -> out[7n]; used instead of `sample` when not NULL
 *   changed      : SDFSurface::changed (:87): returns non-zero and fills out_box when Some(box); may be NULL
 *   tape         : optional GPU capability, not in the reference: bytes of a tape (sdfgpu_tape.h)
 *                  equivalent to sample(p, false), valid until the next call on this surface; returns
 *                  non-zero when the surface has one.  NULL or zero => the surface is sampled on the
 *                  host through sample / sample_batch and the results are ingested.
 *   sample_threads : how many host threads may call sample / sample_batch concurrently (0 or 1 =
 *                  only the calling thread, which is what the reference does, scene/sdf/mod.rs:174;
 *                  WasmerSDF serialises on a Mutex anyway, native.rs:189). */
typedef struct sdfgpu_surface {
    void* self;
    void (*bounding_box)(void* self, float out_bb[6]);
    void (*sample)(void* self, const float p[3], int distance_only, float out_sample[7]);
    void (*sample_batch)(void* self, const float* xyz, uint64_t n, int distance_only, float* out_samples);
    int (*changed)(void* self, float out_box[6]);
    int (*tape)(void* self, const void** bytes, size_t* len);
    uint32_t sample_threads;
} sdfgpu_surface;

/* SDFViewer::update(&mut self, sdf: impl SDFSurface, max_delta_time) -> usize
 * (src/app/scene/sdf/mod.rs:128-217) with the trait object itself: polls sdf.changed(), runs the
 * changed-box state machine (:131-154) and then
 *  - a surface WITH a tape: re-sends the tape when it is new or reported a change and runs every
 *    pending LoadingManager pass on the GPU (sdfgpu_update; a pass is far below max_delta_time);
 *  - a surface WITHOUT a tape (any existing .wasm SDF): walks the LoadingManager in the reference's
 *    order (loading.rs:50-76) for at most `max_delta_seconds` (at least one iteration, :173), decides
 *    `update_required` per voxel (:184-190), calls sample for the required voxels on the host,
 *    and scatters the results into the volumes with the store rules of :196-208 applied on the GPU.
 *    A pass may stop half way and continue in the next call, exactly like the reference.
 * `iterations` receives the LoadingManager iterations performed (the return value of update). */
int sdfgpu_update_surface(sdfgpu_ctx* ctx, const sdfgpu_surface* sdf, double max_delta_seconds,
                          uint64_t* iterations);

/* SDFViewer::commit (src/app/scene/sdf/mod.rs:220-239): no upload is needed
 * (the volumes already live in HBM); latches
 * lod_dist_between_samples = 2^passes_left for the tracer (:226). */
int sdfgpu_commit(sdfgpu_ctx* ctx);

/* loading_mgr.{len(), total_iterations(), passes_left(), passes}
 * (src/app/scene/sdf/loading.rs:80-105; read by scene/mod.rs:153,229-239). */
int sdfgpu_loading_state(const sdfgpu_ctx* ctx, uint64_t* len, uint64_t* total_iterations,
                         uint32_t* passes_left, uint32_t* passes);

/* The LoadingManager itself (src/app/scene/sdf/loading.rs:5-115) as a device-free object: the same
 * counters and cursor a handle keeps internally, for hosts that want the visit order (progress bars,
 * their own batching) and for the tests ported from loading.rs:117-171.
 *   next     : Iterator::next (:50-76): returns 1 and the index, or 0 when loading is done
 *   next_run : up to max_iters consecutive next() results that share one x row: they are
 *              (first[0] + i * *step, first[1], first[2]) for i < the returned count (0 = done)
 *   len / total_iterations / passes_left : :80-105 */
typedef struct sdfgpu_loading sdfgpu_loading;
int sdfgpu_loading_create(const uint32_t limits[3], uint32_t passes, sdfgpu_loading** out);
void sdfgpu_loading_destroy(sdfgpu_loading* lm);
void sdfgpu_loading_reset(sdfgpu_loading* lm, uint32_t passes);
int sdfgpu_loading_next(sdfgpu_loading* lm, uint32_t out_index[3]);
uint64_t sdfgpu_loading_next_run(sdfgpu_loading* lm, uint64_t max_iters, uint32_t first[3], uint32_t* step);
uint64_t sdfgpu_loading_len(const sdfgpu_loading* lm);
uint64_t sdfgpu_loading_total_iterations(const sdfgpu_loading* lm);
uint32_t sdfgpu_loading_passes_left(const sdfgpu_loading* lm);

/* Reset both volumes to AIR_DIST and restart loading with `loading_passes`
 * (what set_sdf does by rebuilding the viewer, scene/mod.rs:154-155). */
int sdfgpu_reset(sdfgpu_ctx* ctx, uint32_t loading_passes);

/* The CPU-side `tex0` / `tex1` Vec<[f32;4]> (src/app/scene/sdf/mod.rs:23-25).
 * Each buffer holds W*H*(z_end-z_begin) texels of 4 floats (owned slices, no
 * halo); either pointer may be NULL. */
int sdfgpu_download(sdfgpu_ctx* ctx, float* tex0, float* tex1);

/* Device pointers of the stored slab (including halo slices, z_lo first), for
 * CUDA/GL interop and for the NCCL halo exchange done by the host side. */
int sdfgpu_device_ptrs(sdfgpu_ctx* ctx, void** tex0, void** tex1);

/* ------------------------------------------------------------------- trace */

/* Uniform contract of the tracer: SDFViewerMaterial::use_uniforms
 * (src/app/scene/sdf/material.rs:50-73) plus the camera of scene/mod.rs:82-95.
 * Matrices are column-major (cgmath / GLSL). */
typedef struct sdfgpu_camera {
    float position[3];     /* cameraPosition                                       */
    float view[16];        /* camera.view()                                        */
    float projection[16];  /* camera.projection()                                  */
    float tint[4];         /* surfaceColorTint, linear (material.rs:64); white     */
    uint32_t tone_mapping; /* three-d ToneMapping: 0 none 1 reinhard 2 aces 3 filmic */
    uint32_t color_mapping;/* three-d ColorMapping: 0 none 1 compute-to-sRGB       */
    float gamma;           /* GAMMA_CORRECTION define (material.rs:39-41); 0 = off */
    float ambient[3];      /* AmbientLight colour * intensity (scene/mod.rs:106)   */
} sdfgpu_camera;

/* The camera of SDFViewerAppScene::new (src/app/scene/mod.rs:82-95): eye (2.5,3,5) looking at the
 * origin, up +Y, fovy 45 deg, near 0.1, far 1000; three-d defaults ACES tone mapping + sRGB colour
 * mapping; one white ambient light of intensity 1 (scene/mod.rs:106). */
void sdfgpu_camera_default(sdfgpu_camera* cam, uint32_t width, uint32_t height);

/* cgmath 0.18 Matrix4::look_at_rh / perspective, column-major (what three-d's Camera stores). */
void sdfgpu_look_at_rh(const float eye[3], const float center[3], const float up[3], float m[16]);
void sdfgpu_perspective(float fovy_rad, float aspect, float z_near, float z_far, float m[16]);

/* CUDA <-> OpenGL interop presenter: the frame goes from the trace kernel into the presenter's GL textures
 * on the device, with no copy through the host (the reference renders straight into the egui/three-d
 * framebuffer, src/app/frameinput.rs:12-67, src/app/scene/mod.rs:202-224).
 *   sdfgpu_gl_register : color_texture = a GL_RGBA8 texture of width x height; depth_texture = a GL_R32F
 *       texture of the same size that receives gl_FragDepth, or 0; gl_target = GL_TEXTURE_2D (0x0DE1).
 *       The GL context that owns them must be current on the calling thread and live on this handle's
 *       device.  Replaces an earlier registration.
 *   sdfgpu_trace_gl    : material.frag main() for every pixel, written into the registered textures (row 0 =
 *       bottom row, GL's origin); stream-ordered, GL may sample them when the call returns.  The host
 *       then draws one full-screen quad that copies colour and writes gl_FragDepth under the blend /
 *       depth state of src/app/scene/sdf/material.rs:75-81.
 * Without a current GL context registration fails with SDFGPU_ERR_CUDA and the host keeps
 * sdfgpu_trace_rgba8. */
int sdfgpu_gl_register(sdfgpu_ctx* ctx, uint32_t color_texture, uint32_t depth_texture, uint32_t gl_target,
                       uint32_t width, uint32_t height);
int sdfgpu_gl_unregister(sdfgpu_ctx* ctx);
int sdfgpu_trace_gl(sdfgpu_ctx* ctx, const sdfgpu_camera* cam);

/* Lower-level ray description derived from the camera (kept public so the
 * parity tests can hand identical numbers to the oracle):
 *   unnormalised direction of pixel (i,j), j = 0 at the BOTTOM row (GL window
 *   coordinates):  base + (i + 0.5) * dx + (j + 0.5) * dy
 *   bvp = bias * projection * view (material.rs:89-97), column-major. */
typedef struct sdfgpu_rays {
    float origin[3];
    float base[3];
    float dx[3];
    float dy[3];
    float bvp[16];
} sdfgpu_rays;

int sdfgpu_camera_rays(const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                       sdfgpu_rays* out);

/* G-buffer record per pixel, SDFGPU_GBUF_FLOATS floats:
 *  [0..2] hit position  [3] t (distance from ray origin; -1 out of steps,
 *  -2 out of bounds, -3 ray misses the bounding box => no fragment,
 *  -4 exact multi-GPU trace only: a hit that another rank owns and shades)
 *  [4..7] raw tex0 at hit  [8..11] raw tex1 at hit  [12..14] normal  [15] steps */
#define SDFGPU_GBUF_FLOATS 16

/* material.frag main() (src/app/scene/sdf/material.frag:130-182) for every
 * pixel whose ray enters the bounding box.  Outputs (any may be NULL):
 *   rgba  : width*height*4 floats, outColor as the shader writes it
 *   depth : width*height floats, gl_FragDepth (1.0 on miss)
 *   gbuf  : width*height*SDFGPU_GBUF_FLOATS floats
 * Host pointers; the frame is traced on the device and copied back. */
int sdfgpu_trace(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                 float* rgba, float* depth, float* gbuf);

/* The frame as the reference's RGBA8 framebuffer receives it: rgba8 = width*height*4 bytes,
 * round(clamp(outColor, 0, 1) * 255) per channel, plus gl_FragDepth; 8 bytes per pixel over PCIe
 * instead of 20.  Either pointer may be NULL. */
int sdfgpu_trace_rgba8(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                       uint8_t* rgba8, float* depth);

/* Same, but leaves the frame in device memory owned by the handle (valid until
 * the next trace with a different size, or destroy) and does not synchronise. */
int sdfgpu_trace_device(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width,
                        uint32_t height, int want_gbuf, void** rgba_dev, void** depth_dev,
                        void** gbuf_dev);

/* What the next trace would use: the clip box (the bounding box, or with slab_clip != 0 this
 * handle's slab sub-box), the lod latched by sdfgpu_commit and the texture filter state
 * (0 NEAREST until the first commit at lod == 1, then LINEAR; scene/sdf/mod.rs:110-111,226-238).
 * Any output may be NULL.  For tests that hand the same numbers to the oracle. */
int sdfgpu_trace_params(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                        int slab_clip, float clip_min[3], float clip_max[3], float* lod,
                        uint32_t* filter_linear);

/* Trace only this handle's slab (multi-GPU sort-last): rays are clipped to the
 * slab's own sub-box; per pixel a 64-bit key = (depth bits << 32) | RGBA8 is
 * written so that an element-wise MIN over ranks composites the frame. */
int sdfgpu_trace_slab_keys(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width,
                           uint32_t height, void** keys_dev);

/* Exact multi-GPU trace.  The sort-last trace above restarts every ray at its slab's faces, which moves
 * hit points by O(1e-5).  The exact form keeps the single-GPU step sequence: every rank holds a copy of the
 * distance channel of the WHOLE grid (4 bytes per voxel), marches every ray through it, and the rank
 * that owns a hit -- the one whose own slices hold the lower z tap of the hit's texture fetch -- shades it
 * from its slab; the other ranks write the miss key, so the same element-wise MIN composites a frame that
 * equals sdfgpu_trace_rgba8 of one handle holding the whole grid, bit for bit.
 *   sdfgpu_exact_trace_prepare : (after every change of this handle's volume) allocates the full-grid
 *       distance volume on first use and extracts this handle's own slices into their place.  Returns the
 *       device pointer of the volume (W*H*D floats, flat order of :177) and this handle's own range in
 *       floats, so the host can all-gather in place (NCCL), or stage through
 *   sdfgpu_dist_volume_read / _write : host copies of a float range of that volume (hosts without NCCL).
 *   sdfgpu_trace_exact_keys : the trace; the caller has gathered every rank's range beforehand.
 * Not in the reference (single process). */
int sdfgpu_exact_trace_prepare(sdfgpu_ctx* ctx, void** dist_dev, uint64_t* own_first, uint64_t* own_count);
int sdfgpu_dist_volume_read(sdfgpu_ctx* ctx, uint64_t first, uint64_t count, float* host_dst);
int sdfgpu_dist_volume_write(sdfgpu_ctx* ctx, uint64_t first, uint64_t count, const float* host_src);
int sdfgpu_trace_exact_keys(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width,
                            uint32_t height, void** keys_dev);

/* Pack / unpack helpers for the composited keys (device -> host RGBA8 + depth). */
int sdfgpu_keys_download(sdfgpu_ctx* ctx, const void* keys_dev, uint32_t width,
                         uint32_t height, uint8_t* rgba8, float* depth);

/* Statistics of the per-tile culling the fill applies to a UNION_RANGE of >= 16 primitives (exact-safe: a primitive is
 * dropped from a tile's list only if its lower distance bound over the tile exceeds another primitive's upper bound;
 * DESIGN.md section 4.1).  Runs ONE fill of every voxel with the current tape (as sdfgpu_fill_all; idempotent) and
 * returns the number of tiles, the sum and the maximum of the survivors per tile, and the primitives of the range
 * (0 everywhere when the tape has no culled range).  Not in the reference (no CSG workload there). */
int sdfgpu_cull_stats(sdfgpu_ctx* ctx, uint64_t* tiles, uint64_t* survivors_sum, uint64_t* survivors_max,
                      uint32_t* primitives);

/* ------------------------------------------------------------------ stream */

int sdfgpu_sync(sdfgpu_ctx* ctx);
/* cudaStream_t of the handle, as void* (for CUDA events on the launching stream). */
void* sdfgpu_stream(sdfgpu_ctx* ctx);
/* Kernels launched by this handle so far (bench.py's gpu_launches claim). */
uint64_t sdfgpu_launch_count(const sdfgpu_ctx* ctx);
/* Fill-kernel knobs (0 = library default); used by the benchmarks and tests:
 *   "fill_voxels_per_thread" 0|1|2|4|8   "fill_ctas_per_sm" 0..32
 *   "fill_halo" 0|1 (compute halo slices locally; 0 when the host exchanges them)
 *   "fill_cull_cells" 0|1 (default 1): tiles of a culled UNION_RANGE start from the survivors of their 64^3-voxel cell
 *                  (computed once per set_tape) instead of the whole range; same volume either way
 *   "trace_distance_volume" 0..3: where the LINEAR march reads its distances.  0 (default) tex0.r in place;
 *                  1 a distance-only copy of tex0.r in linear memory (4 B / voxel, rebuilt after every change):
 *                  same values, a quarter of the bytes per fetch; 2 the copy as an R32F 3-D CUDA array read
 *                  through a texture object in point mode (TMU + block-linear layout, exact fp32 blend: same
 *                  frame); 3 the array with hardware LINEAR filtering (one fetch per step, 8-bit weights:
 *                  a fast mode OUTSIDE the 1e-5 parity bar).  1-3 pay off when many frames are traced per fill
 *   "trace_variant" 0 heavy-first 8x8 tiles | 1 plain 2-D grid | 2 persistent warps pulling 8x4 tiles from a queue,
 *                  explicit warp-ballot exit (the kernel linked handles always use); same frame bit for bit
 *   "trace_tile_order" 0 | 1 auto (default) | 2 always: the tile grid starts the tiles whose marches were longest in the
 *                  previous frame of the same size first, so that a frame does not end on a long march that started
 *                  late; auto = when the box's screen rectangle holds at least half of the frame's tiles.  Same frame
 *   "trace_bands" 1..32 (default 6): sdfgpu_trace_rgba8 traces the frame in this many bands of tile rows and copies each
 *                  band's rows to the host while the next is traced (1: trace, then copy); same pixels
 *   "link_wait_mode" 0 cuStreamWaitValue32 when the driver has it (default) | 1 spin-wait kernels; before link_attach
 *   "link_halo_push" 0|1 and "link_trace_mode" 0 auto | 1 rounds | 2 stream: the SDFGPU_LINK_* flags as options, read by
 *                  sdfgpu_link_export;  "link_timeout_ms" (default 8000): how long a streaming trace kernel waits
 *                  for rays of a neighbour that never arrives before the frame fails with SDFGPU_ERR_STATE
 *   "fill_program" 0 auto (kernel specialised for the tape structure, else built-in demo program,
 *                  else interpreter) | 1 interpreter | 2 built-in or interpreter | 3 specialised or fail */
int sdfgpu_set_option(sdfgpu_ctx* ctx, const char* key, int64_t value);
/* Introspection: "last_fill_program" (0 interpreter, 1 specialised/JIT, 2 built-in demo),
 * "last_fill_ctas_per_sm", "last_fill_voxels_per_thread", "sm_count", "tape_image_bytes",
 * "tape_culled", "jit_available", "device", "linked", "link_memops", "link_halo_push", "link_trace_stream". */
int sdfgpu_get_info(const sdfgpu_ctx* ctx, const char* key, int64_t* value);

#ifdef __cplusplus
}
#endif
#endif /* SDFGPU_H */

// link.cu -- Z-sharded multi-GPU operation of slab handles behind the C ABI (include/sdfgpu.h, "multi-GPU").
//
// The reference is one process on one thread (/root/reference/src/app/scene/mod.rs:22-31, 158-225) and has no
// multi-GPU path (src/app/scene/sdf/mod.rs:174 "TODO: parallel iteration"); everything here is this build's.
//
// A slab handle owns the z slices [z_begin, z_end) of the grid plus one halo slice per interior face.  LINKING
// it to the handles of the other ranks -- in other processes (one process per GPU: CUDA IPC) or in this process
// (peer pointers) -- maps the neighbours' volumes and every rank's ARENA: a block of device memory with the
// flags, ray queues and the presenter's frame.  After that the ordinary entry points are collective:
//
//   fill    one launch per rank; the tiles holding the two boundary slices go first, the last of them releases
//           a flag, and the copy engines -- waiting on that flag in a second stream -- push the slices into the
//           neighbours' halo slices over NVLink while the grid fills the interior (fill_device.cuh, FillParams).
//   trace   `world` rounds of one persistent kernel (trace.cu, trace_rounds_kernel): a ray marches on the rank
//           that owns the lower z tap of its fetch and is handed to the neighbour (24 bytes) when it leaves, so
//           the frame equals the single-volume frame bit for bit; finished pixels are stored straight into the
//           presenter's frame (rank 0) over NVLink.  No collective library call, no host synchronisation
//           between ranks: ordering is by epoch flags in the arenas -- written by the producing kernel's last
//           CTA (or a one-thread kernel after DMA), awaited by stream memory operations (cuStreamWaitValue32).
//
// Flags (all u32 epochs, compared cyclically):
//   halo_in[side]     the neighbour on that side has pushed the boundary slice of its fill number `value`
//   round_done[side]  that neighbour has finished global trace round value - 1: its out-queue of that round is
//                     complete, and it no longer reads the halo slices / in-queues of earlier rounds
//   frame_done[r]     (presenter) rank r has stored all its pixels of frame value - 1
//   consumed          (from the presenter) frames 0 .. value - 1 have been unpacked: their key frame may be reused
#include <unistd.h>

#include <cstddef>
#include <cstdlib>
#include <new>

#include "sdfgpu_ctx.h"

using namespace sdfgpu;

#define SDFGPU_API extern "C" __attribute__((visibility("default")))

namespace {

// ---- arena layout
struct ArenaHeader {
    uint32_t halo_in[2];
    uint32_t round_done[2];
    uint32_t consumed;
    uint32_t boundary_flag;
    uint32_t boundary_count;
    uint32_t work_head;
    uint32_t ctas_done;
    uint32_t timed_out;
    uint32_t tiles_done;      // stream trace: tile units finished
    uint32_t outside_done;    // stream trace, presenter: units of tiles outside the box's screen rectangle finished
    uint32_t in_head[2];      // stream trace: in-queue units claimed [0 from below | 1 from above]
    uint32_t in_done[2];      //   ... and finished
    uint32_t frame_done[LINK_MAX_WORLD];
    uint32_t out_count[4][2];
    uint32_t in_count[2][2];  // [round parity][0 from below | 1 from above], written by the neighbours
    uint32_t sent_final[2];   // stream trace: this rank has closed its out-queue [0 down | 1 up]
    unsigned long long in_final[2][2];  // stream trace, [frame parity][from below | above], written by the neighbours:
                                        // frame tag << 32 | entries in that in-queue (the queue is closed)
    uint32_t outside_flag;    // stream trace, presenter: frame tag of the last frame whose outside tiles are all written
    uint32_t pad2[9];
};
static_assert(sizeof(ArenaHeader) == 256, "arena header is 256 bytes");
static_assert(offsetof(ArenaHeader, in_final) % 8 == 0, "64-bit words are aligned");

struct ArenaLayout {
    // in-queues [round / frame parity][0 from below | 1 from above], 32 bytes per entry: the round kernel keeps
    // float4 positions at pos and uint2 ids at id, the stream kernel two tagged float4 per entry from pos on
    size_t pos[2][2], id[2][2], keys[2], gbuf, total;
};
ArenaLayout arena_layout(uint32_t max_pixels, bool want_gbuf) {
    ArenaLayout a;
    size_t off = sizeof(ArenaHeader);
    const size_t n = ((size_t)max_pixels + 31u) & ~(size_t)31u;
    for (int p = 0; p < 2; ++p)
        for (int q = 0; q < 2; ++q) {
            a.pos[p][q] = off; off += n * sizeof(float4);
            a.id[p][q] = off; off += n * sizeof(float4);
        }
    for (int p = 0; p < 2; ++p) { a.keys[p] = off; off += n * sizeof(unsigned long long); }
    a.gbuf = want_gbuf ? off : 0;
    if (want_gbuf) off += n * SDFGPU_GBUF_FLOATS * sizeof(float);
    a.total = off;
    return a;
}

struct LinkBlob {  // what a rank publishes (SDFGPU_LINK_BLOB_BYTES)
    uint32_t magic, rank, world, flags;
    uint64_t pid;
    uint32_t device, z_begin, z_end, z_lo, z_hi, max_pixels;
    uint32_t dims[3], pad;
    uint64_t arena_bytes;
    uint64_t p_tex0, p_tex1, p_arena;
    cudaIpcMemHandle_t h_tex0, h_tex1, h_arena;
    unsigned char uuid[16];  // of the device: ranks that share a GPU cannot run the stream trace
};
static_assert(sizeof(LinkBlob) <= SDFGPU_LINK_BLOB_BYTES, "blob fits");
constexpr uint32_t LINK_MAGIC = 0x4b4c4453u;  // "SDLK"

// ---- stream memory operations (driver API, resolved through the runtime)
typedef int (*cuStreamWaitValue32_t)(cudaStream_t, unsigned long long addr, uint32_t value, unsigned flags);
cuStreamWaitValue32_t g_wait32 = nullptr;
bool g_wait32_looked = false;

cuStreamWaitValue32_t wait32() {
    if (!g_wait32_looked) {
        g_wait32_looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            g_wait32 = reinterpret_cast<cuStreamWaitValue32_t>(fn);
        (void)cudaGetLastError();
    }
    return g_wait32;
}

ArenaHeader* hdr_of(unsigned char* arena) { return reinterpret_cast<ArenaHeader*>(arena); }

// stream `s` continues once the u32 at `flag` (own arena) has reached `value`
int wait_flag(sdfgpu_ctx* ctx, cudaStream_t s, uint32_t* flag, uint32_t value) {
    if (value == 0) return SDFGPU_OK;  // flags start at 0
    if (ctx->link.memops) {
        const int r = wait32()(s, (unsigned long long)(uintptr_t)flag, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (r != 0) return fail(ctx, SDFGPU_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", r);
        return SDFGPU_OK;
    }
    CK(ctx, launch_spin_wait(flag, value, &hdr_of(ctx->link.arena)->timed_out, s));
    ctx->launches++;
    return SDFGPU_OK;
}

int signal_flags(sdfgpu_ctx* ctx, cudaStream_t s, uint32_t* const* flags, const uint32_t* values, int n) {
    if (n == 0) return SDFGPU_OK;
    CK(ctx, launch_signal(flags, values, n, s));
    ctx->launches += (uint64_t)(n + 3) / 4;
    return SDFGPU_OK;
}

ArenaHeader* peer_hdr(sdfgpu_ctx* ctx, int rank) { return hdr_of(ctx->link.peer_arena[rank]); }

bool link_timing() {
    static const bool on = [] { const char* e = getenv("SDFGPU_LINK_TIMING"); return e && *e && *e != '0'; }();
    return on;
}

void timing_mark(sdfgpu_ctx* ctx, cudaStream_t s = nullptr) {
    if (!link_timing()) return;
    LinkState& L = ctx->link;
    if (L.timing_used == L.timing_events.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        L.timing_events.push_back(e);
    }
    (void)cudaEventRecord(L.timing_events[L.timing_used++], s ? s : ctx->stream);
}

// tell both neighbours that fill number `epoch` of this rank is in their halo slices
int signal_halo_in(sdfgpu_ctx* ctx, cudaStream_t s, uint32_t epoch) {
    uint32_t* flags[2];
    uint32_t values[2];
    int n = 0;
    for (int side = 0; side < 2; ++side) {
        const int nb = ctx->link.nb[side];
        if (nb < 0) continue;
        flags[n] = &peer_hdr(ctx, nb)->halo_in[1 - side];  // I am the neighbour on ITS other side
        values[n++] = epoch;
    }
    return signal_flags(ctx, s, flags, values, n);
}

// the neighbours have finished every trace round issued so far: their halo slices may be overwritten
int wait_neighbours_idle(sdfgpu_ctx* ctx, cudaStream_t s, bool lo, bool hi) {
    for (int side = 0; side < 2; ++side) {
        if (ctx->link.nb[side] < 0 || !(side == 0 ? lo : hi)) continue;
        const int rc = wait_flag(ctx, s, &hdr_of(ctx->link.arena)->round_done[side], ctx->link.round_epoch);
        if (rc != SDFGPU_OK) return rc;
    }
    return SDFGPU_OK;
}

int ensure_halo_stream(sdfgpu_ctx* ctx) {
    if (!ctx->halo_stream) {
        CK(ctx, cudaStreamCreateWithFlags(&ctx->halo_stream, cudaStreamNonBlocking));
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_boundary, cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_pushed, cudaEventDisableTiming));
    }
    return SDFGPU_OK;
}

}  // namespace

bool sdfgpu::stream_wait_value_available() { return wait32() != nullptr; }

bool sdfgpu::stream_wait_value(cudaStream_t s, const uint32_t* flag, uint32_t value) {
    return wait32() && wait32()(s, (unsigned long long)(uintptr_t)flag, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */) == 0;
}

// ------------------------------------------------------------------------------------------------- fill

int sdfgpu::link_after_fill(sdfgpu_ctx* ctx, bool touched_lo, bool touched_hi) {
    LinkState& L = ctx->link;
    if (!L.on) return SDFGPU_OK;
    int rc;
    const uint32_t f = ++L.fill_epoch;
    if (!L.halo_push) return SDFGPU_OK;  // the halo slices were filled here like the own ones
    if ((rc = wait_neighbours_idle(ctx, ctx->stream, touched_lo, touched_hi)) != SDFGPU_OK) return rc;
    if ((rc = push_halos(ctx, ctx->stream, touched_lo, touched_hi)) != SDFGPU_OK) return rc;
    return signal_halo_in(ctx, ctx->stream, f);
}

int sdfgpu::link_fill_all_fused(sdfgpu_ctx* ctx, FillParams* p) {
    LinkState& L = ctx->link;
    if (!L.on || !L.halo_push || (L.nb[0] < 0 && L.nb[1] < 0)) return SDFGPU_OK;
    ArenaHeader* h = hdr_of(L.arena);
    const uint32_t bz = p->tiles_z < 2u ? p->tiles_z : 2u;
    p->n_boundary_tiles = p->tiles_x * p->tiles_y * bz;
    p->boundary_epoch = L.fill_epoch + 1u;
    p->boundary_count = &h->boundary_count;
    p->boundary_flag = &h->boundary_flag;
    return SDFGPU_OK;
}

int sdfgpu::link_fill_all_pushed(sdfgpu_ctx* ctx) {
    LinkState& L = ctx->link;
    if (!L.on) return SDFGPU_OK;
    const uint32_t f = ++L.fill_epoch;
    if (!L.halo_push || (L.nb[0] < 0 && L.nb[1] < 0)) return SDFGPU_OK;
    int rc;
    if ((rc = ensure_halo_stream(ctx)) != SDFGPU_OK) return rc;
    // the copy engines wait for the fill kernel's boundary flag (epochs are unique: no event needed), not for the kernel
    if ((rc = wait_flag(ctx, ctx->halo_stream, &hdr_of(L.arena)->boundary_flag, f)) != SDFGPU_OK) return rc;
    if ((rc = wait_neighbours_idle(ctx, ctx->halo_stream, true, true)) != SDFGPU_OK) return rc;
    if ((rc = push_halos(ctx, ctx->halo_stream)) != SDFGPU_OK) return rc;
    if ((rc = signal_halo_in(ctx, ctx->halo_stream, f)) != SDFGPU_OK) return rc;
    CK(ctx, cudaEventRecord(ctx->ev_pushed, ctx->halo_stream));
    // whatever follows the fill on the main stream (the next fill above all) comes after the push has read the slices
    CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pushed, 0));
    return SDFGPU_OK;
}

// ------------------------------------------------------------------------------------------------ trace

int sdfgpu::link_trace_begin(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t w, uint32_t h, bool want_gbuf) {
    LinkState& L = ctx->link;
    if (!L.on) return fail(ctx, SDFGPU_ERR_STATE, "handle is not linked");
    if (L.in_frame) return fail(ctx, SDFGPU_ERR_STATE, "a linked frame is already in flight");
    if (!cam) return fail(ctx, SDFGPU_ERR_INVALID, "cam is NULL");
    if (w == 0 || h == 0 || (uint64_t)w * h > L.max_pixels)
        return fail(ctx, SDFGPU_ERR_INVALID, "frame %ux%u exceeds the %u pixels the link was created for", w, h, L.max_pixels);
    if (want_gbuf && !L.want_gbuf) return fail(ctx, SDFGPU_ERR_STATE, "the link was exported without a G-buffer frame");
    set_device(ctx);
    int rc;
    if (L.rank == 0) {
        if ((rc = ensure_frame(ctx, w, h, false, false)) != SDFGPU_OK) return rc;
        if (!ctx->rgba8_dev) CK(ctx, cudaMalloc(&ctx->rgba8_dev, (size_t)w * h * sizeof(uint32_t)));
    }
    if ((rc = fill_trace_params(ctx, cam, w, h, false, &ctx->link_tp)) != SDFGPU_OK) return rc;
    ctx->link_tp.dist = nullptr; ctx->link_tp.dist_mode = 0; ctx->link_tp.dist_tex = 0;
    const uint32_t t = L.frame_epoch++;
    ArenaHeader* hd = hdr_of(L.arena);
    for (int side = 0; side < 2 && L.halo_push; ++side)  // the neighbours' boundary slices of the last fill are in my halo slices
        if (L.nb[side] >= 0 && (rc = wait_flag(ctx, ctx->stream, &hd->halo_in[side], L.fill_epoch)) != SDFGPU_OK) return rc;
    if (L.rank != 0) {
        // the presenter's key frame of this parity was last used by frame t - 2 (the single G-buffer frame by t - 1)
        const uint32_t need = want_gbuf ? t : (t >= 1u ? t - 1u : 0u);
        if ((rc = wait_flag(ctx, ctx->stream, &hd->consumed, need)) != SDFGPU_OK) return rc;
    } else if (L.present_stream && t >= 1u) {
        // The presenter's trace kernel comes after everything its presenter stream still has to do for the previous
        // frame.  Not for the key frame's sake (that is the frame before): a streaming trace kernel fills every SM and
        // WAITS for the other ranks, which wait for the `consumed` signal that follows the unpack kernel -- which would
        // then find no SM to run on.  Between two frames there normally is a fill, during which that work is done.
        CK(ctx, cudaStreamWaitEvent(ctx->stream, L.ev_presented, 0));
    }
    L.cur_w = w; L.cur_h = h; L.cur_round = 0;
    L.cur_outside_first = false;
    L.cur_gbuf = want_gbuf;
    L.in_frame = true;
    L.timing_used = 0;
    timing_mark(ctx);
    return SDFGPU_OK;
}

int sdfgpu::link_trace_round(sdfgpu_ctx* ctx) {
    LinkState& L = ctx->link;
    if (!L.on || !L.in_frame) return fail(ctx, SDFGPU_ERR_STATE, "no linked frame in flight");
    if (L.cur_round >= L.world) return fail(ctx, SDFGPU_ERR_STATE, "all %u rounds of the frame have been issued", L.world);
    set_device(ctx);
    const uint32_t k = L.cur_round++;
    const uint32_t g = L.round_epoch++;
    const uint32_t t = L.frame_epoch - 1u;
    ArenaHeader* hd = hdr_of(L.arena);
    int rc;
    for (int side = 0; side < 2; ++side)  // the neighbour has finished round g - 1
        if (L.nb[side] >= 0 && (rc = wait_flag(ctx, ctx->stream, &hd->round_done[side], g)) != SDFGPU_OK) return rc;
    timing_mark(ctx);
    const ArenaLayout lay = arena_layout(L.max_pixels, L.want_gbuf);
    LinkParams lp;
    memset(&lp, 0, sizeof lp);
    lp.linked = 1u;
    lp.first = k == 0 ? 1u : 0u;
    lp.is_presenter = L.rank == 0 ? 1u : 0u;
    lp.own_z0 = ctx->z_begin; lp.own_z1 = ctx->z_end;
    lp.max_pixels = L.max_pixels;
    lp.work_head = &hd->work_head;
    lp.ctas_done = &hd->ctas_done;
    if (k > 0) {
        const uint32_t pg = g - 1u;  // the round whose out-queues are this round's in-queues
        for (int q = 0; q < 2; ++q) {
            if (L.nb[q] < 0) continue;
            lp.in_count[q] = &hd->in_count[pg & 1u][q];
            lp.in_pos[q] = reinterpret_cast<const float4*>(L.arena + lay.pos[pg & 1u][q]);
            lp.in_id[q] = reinterpret_cast<const uint2*>(L.arena + lay.id[pg & 1u][q]);
        }
    }
    lp.out_count = hd->out_count[g & 3u];
    for (int dir = 0; dir < 2; ++dir) {  // dir 0: down, into the lower neighbour's "from above" queue; 1: up
        const int nb = L.nb[dir];
        if (nb < 0) continue;
        unsigned char* a = L.peer_arena[nb];
        const int q = 1 - dir;
        lp.out_pos[dir] = reinterpret_cast<float4*>(a + lay.pos[g & 1u][q]);
        lp.out_id[dir] = reinterpret_cast<uint2*>(a + lay.id[g & 1u][q]);
        lp.out_publish[dir] = &hdr_of(a)->in_count[g & 1u][q];
    }
    lp.reset_count = hd->out_count[(g + 2u) & 3u];
    unsigned char* pa = L.peer_arena[0];
    lp.frame_keys = reinterpret_cast<unsigned long long*>(pa + lay.keys[t & 1u]);
    lp.frame_gbuf = L.cur_gbuf ? reinterpret_cast<float*>(pa + lay.gbuf) : nullptr;
    for (int side = 0; side < 2; ++side)
        if (L.nb[side] >= 0) lp.sig_round[side] = &peer_hdr(ctx, L.nb[side])->round_done[1 - side];
    lp.sig_round_value = g + 1u;
    if (k + 1u == L.world && L.rank != 0) {
        lp.sig_frame = &hdr_of(pa)->frame_done[L.rank];
        lp.sig_frame_value = t + 1u;
    }
    const int per_sm = trace_rounds_max_ctas_per_sm(ctx->link_tp);
    if (per_sm < 1) return fail(ctx, SDFGPU_ERR_CUDA, "the trace kernel does not fit on an SM");
    CK(ctx, launch_trace_rounds(ctx->link_tp, lp, ctx->sm_count * per_sm, ctx->stream));
    ctx->launches++;
    timing_mark(ctx);
    return SDFGPU_OK;
}

// The whole frame of this rank in one launch (LinkState::stream): the kernels of the ranks run side by side and feed
// each other's in-queues; nothing here waits for a neighbour -- the kernel does.
int sdfgpu::link_trace_stream(sdfgpu_ctx* ctx) {
    LinkState& L = ctx->link;
    if (!L.on || !L.in_frame) return fail(ctx, SDFGPU_ERR_STATE, "no linked frame in flight");
    if (!L.stream) return fail(ctx, SDFGPU_ERR_STATE, "the link traces in rounds");
    if (L.cur_round != 0) return fail(ctx, SDFGPU_ERR_STATE, "the frame has been issued");
    set_device(ctx);
    L.cur_round = L.world;
    const uint32_t g = L.round_epoch++;
    const uint32_t t = L.frame_epoch - 1u;
    ArenaHeader* hd = hdr_of(L.arena);
    const ArenaLayout lay = arena_layout(L.max_pixels, L.want_gbuf);
    LinkParams lp;
    memset(&lp, 0, sizeof lp);
    lp.linked = 1u;
    lp.first = 1u;
    lp.is_presenter = L.rank == 0 ? 1u : 0u;
    lp.own_z0 = ctx->z_begin; lp.own_z1 = ctx->z_end;
    lp.max_pixels = L.max_pixels;
    lp.work_head = &hd->work_head;
    lp.ctas_done = &hd->ctas_done;
    lp.epoch = t + 1u;
    lp.timeout_ms = L.timeout_ms;
    lp.tiles_done = &hd->tiles_done;
    lp.in_head = hd->in_head;
    lp.in_done = hd->in_done;
    lp.sent_final = hd->sent_final;
    lp.timed_out = L.timed_out_host;
    lp.out_count = hd->out_count[0];
    const uint32_t par = t & 1u;
    for (int q = 0; q < 2; ++q) {  // q 0: the neighbour below feeds my "from below" queue; dir 0: I feed ITS "from above" queue
        const int nb = L.nb[q];
        if (nb < 0) continue;
        lp.in_q[q] = reinterpret_cast<const float4*>(L.arena + lay.pos[par][q]);
        lp.in_final[q] = &hd->in_final[par][q];
        unsigned char* a = L.peer_arena[nb];
        lp.out_q[q] = reinterpret_cast<float4*>(a + lay.pos[par][1 - q]);
        lp.out_final[q] = &hdr_of(a)->in_final[par][1 - q];
    }
    unsigned char* pa = L.peer_arena[0];
    lp.frame_keys = reinterpret_cast<unsigned long long*>(pa + lay.keys[par]);
    lp.frame_gbuf = L.cur_gbuf ? reinterpret_cast<float*>(pa + lay.gbuf) : nullptr;
    for (int side = 0; side < 2; ++side)
        if (L.nb[side] >= 0) lp.sig_round[side] = &peer_hdr(ctx, L.nb[side])->round_done[1 - side];
    lp.sig_round_value = g + 1u;
    if (L.rank != 0) {
        lp.sig_frame = &hdr_of(pa)->frame_done[L.rank];
        lp.sig_frame_value = t + 1u;
    } else if (L.memops) {
        // the presenter's own tiles outside the box's rectangle: first, and straight into the frame (link_trace_end
        // copies the rows that hold nothing else to the host while the rest is traced)
        lp.out_rgba8 = ctx->rgba8_dev; lp.out_depth = ctx->depth_dev;
        lp.outside_done = &hd->outside_done; lp.outside_flag = &hd->outside_flag;
    }
    L.cur_outside_first = lp.out_rgba8 != nullptr;
    const int per_sm = trace_stream_max_ctas_per_sm(ctx->link_tp);
    if (per_sm < 1) return fail(ctx, SDFGPU_ERR_CUDA, "the trace kernel does not fit on an SM");
    timing_mark(ctx);
    CK(ctx, launch_trace_stream(ctx->link_tp, lp, ctx->sm_count * per_sm, ctx->stream));
    ctx->launches++;
    timing_mark(ctx);
    return SDFGPU_OK;
}

// every round of the frame (or its one streaming launch)
int sdfgpu::link_trace_issue(sdfgpu_ctx* ctx) {
    if (ctx->link.stream) return link_trace_stream(ctx);
    for (uint32_t k = 0; k < ctx->link.world; ++k) {
        const int rc = link_trace_round(ctx);
        if (rc != SDFGPU_OK) return rc;
    }
    return SDFGPU_OK;
}

int sdfgpu::link_trace_end(sdfgpu_ctx* ctx, uint8_t* rgba8, float* depth, float* gbuf, bool sync) {
    LinkState& L = ctx->link;
    if (!L.on || !L.in_frame) return fail(ctx, SDFGPU_ERR_STATE, "no linked frame in flight");
    if (L.cur_round != L.world) return fail(ctx, SDFGPU_ERR_STATE, "%u of %u rounds issued", L.cur_round, L.world);
    set_device(ctx);
    L.in_frame = false;
    const uint32_t t = L.frame_epoch - 1u;
    int rc;
    if (L.rank == 0) {
        if (!L.present_stream) {
            CK(ctx, cudaStreamCreateWithFlags(&L.present_stream, cudaStreamNonBlocking));
            CK(ctx, cudaEventCreateWithFlags(&L.ev_traced, cudaEventDisableTiming));
            CK(ctx, cudaEventCreateWithFlags(&L.ev_presented, cudaEventDisableTiming));
        }
        cudaStream_t ps = L.present_stream;
        ArenaHeader* hd = hdr_of(L.arena);
        const ArenaLayout lay = arena_layout(L.max_pixels, L.want_gbuf);
        const size_t n = (size_t)L.cur_w * L.cur_h;
        // pixel rows [y0, y1) hold the tiles inside the screen rectangle of the box; the rows above and below hold only
        // pixels the presenter's kernel wrote first, straight into the frame (LinkParams::out_rgba8): they cross PCIe
        // while the frame is still being traced, behind the kernel's outside_flag
        size_t y0 = 0, y1 = L.cur_h;
        if (L.cur_outside_first) {
            const TraceParams& tp = ctx->link_tp;
            const uint32_t tiles_y4 = (tp.height + 3u) / 4u;
            const uint32_t ry0 = tp.rect[1] * 2u < tiles_y4 ? tp.rect[1] * 2u : tiles_y4;
            const uint32_t ry1 = tp.rect[3] * 2u < tiles_y4 ? tp.rect[3] * 2u : tiles_y4;
            y0 = (size_t)ry0 * 4u;
            y1 = (size_t)ry1 * 4u < L.cur_h ? (size_t)ry1 * 4u : L.cur_h;
            if (tp.rect[2] <= tp.rect[0] || y1 <= y0) y0 = y1 = 0;  // no tile inside: every row is an outside row
            if ((rgba8 || depth) && (y0 > 0 || y1 < L.cur_h)) {
                if ((rc = wait_flag(ctx, ps, &hd->outside_flag, t + 1u)) != SDFGPU_OK) return rc;
                const size_t w = L.cur_w, lo = y0 * w, hi = y1 * w;
                if (rgba8 && lo) CK(ctx, cudaMemcpyAsync(rgba8, ctx->rgba8_dev, lo * 4, cudaMemcpyDeviceToHost, ps));
                if (depth && lo) CK(ctx, cudaMemcpyAsync(depth, ctx->depth_dev, lo * 4, cudaMemcpyDeviceToHost, ps));
                if (rgba8 && hi < n) CK(ctx, cudaMemcpyAsync(rgba8 + hi * 4, ctx->rgba8_dev + hi, (n - hi) * 4, cudaMemcpyDeviceToHost, ps));
                if (depth && hi < n) CK(ctx, cudaMemcpyAsync(depth + hi, ctx->depth_dev + hi, (n - hi) * 4, cudaMemcpyDeviceToHost, ps));
            }
        }
        CK(ctx, cudaEventRecord(L.ev_traced, ctx->stream));
        CK(ctx, cudaStreamWaitEvent(ps, L.ev_traced, 0));
        for (uint32_t r = 1; r < L.world; ++r)
            if ((rc = wait_flag(ctx, ps, &hd->frame_done[r], t + 1u)) != SDFGPU_OK) return rc;
        timing_mark(ctx, ps);
        const size_t lo = y0 * L.cur_w, cnt = (y1 - y0) * L.cur_w;  // the rows that came through the key frame
        if (cnt) {
            CK(ctx, launch_keys_unpack(reinterpret_cast<const unsigned long long*>(L.arena + lay.keys[t & 1u]) + lo, (uint32_t)cnt,
                                       reinterpret_cast<uint8_t*>(ctx->rgba8_dev + lo), ctx->depth_dev + lo, ps));
            ctx->launches++;
            if (rgba8) CK(ctx, cudaMemcpyAsync(rgba8 + lo * 4, ctx->rgba8_dev + lo, cnt * 4, cudaMemcpyDeviceToHost, ps));
            if (depth) CK(ctx, cudaMemcpyAsync(depth + lo, ctx->depth_dev + lo, cnt * 4, cudaMemcpyDeviceToHost, ps));
        }
        if (gbuf && L.cur_gbuf)
            CK(ctx, cudaMemcpyAsync(gbuf, L.arena + lay.gbuf, n * SDFGPU_GBUF_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, ps));
        uint32_t* flags[LINK_MAX_WORLD];
        uint32_t values[LINK_MAX_WORLD];
        int m = 0;
        for (uint32_t r = 1; r < L.world; ++r) { flags[m] = &peer_hdr(ctx, (int)r)->consumed; values[m++] = t + 1u; }
        if ((rc = signal_flags(ctx, ps, flags, values, m)) != SDFGPU_OK) return rc;
        CK(ctx, cudaEventRecord(L.ev_presented, ps));
        timing_mark(ctx, ps);
    } else {
        timing_mark(ctx);
    }
    if (sync) {
        if (L.present_stream) CK(ctx, cudaStreamSynchronize(L.present_stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (link_timing() && L.timing_used >= 2) {
            std::string line = "[sdfgpu link timing] rank " + std::to_string(L.rank) + " frame " + std::to_string(t) + " us:";
            for (size_t i = 1; i < L.timing_used; ++i) {
                float ms = 0.0f;
                (void)cudaEventElapsedTime(&ms, L.timing_events[i - 1], L.timing_events[i]);
                char b[32];
                snprintf(b, sizeof b, " %.1f", ms * 1e3f);
                line += b;
            }
            fprintf(stderr, "%s  (per round: wait, kernel; presenter then: wait frame_done, unpack+copies)\n", line.c_str());
        }
        if (!L.memops) {
            uint32_t to = 0;
            CK(ctx, cudaMemcpy(&to, &hdr_of(L.arena)->timed_out, 4, cudaMemcpyDeviceToHost));
            if (to) return fail(ctx, SDFGPU_ERR_STATE, "a wait on another rank's flag timed out (ranks out of step?)");
        }
        if (L.stream && L.timed_out_host && *reinterpret_cast<volatile uint32_t*>(L.timed_out_host)) {
            *L.timed_out_host = 0u;
            return fail(ctx, SDFGPU_ERR_STATE, "the trace kernel waited %u ms for its neighbours' rays and gave up (ranks out of step?)", L.timeout_ms);
        }
    }
    return SDFGPU_OK;
}

// ------------------------------------------------------------------------------------------- link set-up

void sdfgpu::link_free(sdfgpu_ctx* ctx) {
    LinkState& L = ctx->link;
    if (ctx->stream) (void)cudaStreamSynchronize(ctx->stream);
    if (ctx->halo_stream) (void)cudaStreamSynchronize(ctx->halo_stream);
    for (uint32_t r = 0; r < LINK_MAX_WORLD; ++r) {
        if (L.peer_arena[r] && L.peer_arena_ipc[r]) (void)cudaIpcCloseMemHandle(L.peer_arena[r]);
        L.peer_arena[r] = nullptr; L.peer_arena_ipc[r] = false;
    }
    if (L.on || L.arena) {
        for (int side = 0; side < 2; ++side) {
            if (L.peer_tex_ipc[side]) {
                if (ctx->peer_tex0[side]) (void)cudaIpcCloseMemHandle(ctx->peer_tex0[side]);
                if (ctx->peer_tex1[side]) (void)cudaIpcCloseMemHandle(ctx->peer_tex1[side]);
            }
            if (L.on) ctx->peer_tex0[side] = ctx->peer_tex1[side] = nullptr;
            L.peer_tex_ipc[side] = false;
        }
    }
    (void)cudaFree(L.arena);
    if (L.present_stream) { (void)cudaStreamSynchronize(L.present_stream); (void)cudaStreamDestroy(L.present_stream); }
    for (cudaEvent_t e : {L.ev_traced, L.ev_presented})
        if (e) (void)cudaEventDestroy(e);
    if (L.timed_out_host) (void)cudaFreeHost(L.timed_out_host);
    for (cudaEvent_t e : L.timing_events) (void)cudaEventDestroy(e);
    (void)cudaGetLastError();
    L = LinkState();
}

SDFGPU_API int sdfgpu_link_export(sdfgpu_ctx* ctx, uint32_t rank, uint32_t world, uint32_t max_width, uint32_t max_height,
                                  uint32_t flags, void* blob, size_t blob_bytes) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (!blob || blob_bytes < SDFGPU_LINK_BLOB_BYTES) return fail(ctx, SDFGPU_ERR_INVALID, "blob must hold %d bytes", SDFGPU_LINK_BLOB_BYTES);
    if (world < 1 || world > LINK_MAX_WORLD || rank >= world) return fail(ctx, SDFGPU_ERR_INVALID, "bad rank / world (at most %u ranks)", LINK_MAX_WORLD);
    if ((uint64_t)max_width * max_height == 0 || (uint64_t)max_width * max_height > (1u << 28))
        return fail(ctx, SDFGPU_ERR_INVALID, "bad maximum frame size");
    if (ctx->z_begin == ctx->z_end || ctx->stored_texels == 0) return fail(ctx, SDFGPU_ERR_STATE, "a linked handle must own at least one slice");
    if (ctx->link.on || ctx->link.arena) return fail(ctx, SDFGPU_ERR_STATE, "already exported (sdfgpu_link_detach first)");
    if (has_peers(ctx)) return fail(ctx, SDFGPU_ERR_STATE, "neighbours already attached with sdfgpu_ipc_attach");
    set_device(ctx);
    LinkState& L = ctx->link;
    if (flags & ~(SDFGPU_LINK_GBUF | SDFGPU_LINK_HALO_PUSH | SDFGPU_LINK_ROUNDS | SDFGPU_LINK_STREAM))
        return fail(ctx, SDFGPU_ERR_INVALID, "unknown link flags 0x%x", flags);
    if (ctx->opt_link_halo_push) flags |= SDFGPU_LINK_HALO_PUSH;
    if (ctx->opt_link_trace_mode == 1) flags |= SDFGPU_LINK_ROUNDS;
    if (ctx->opt_link_trace_mode == 2) flags |= SDFGPU_LINK_STREAM;
    if ((flags & SDFGPU_LINK_ROUNDS) && (flags & SDFGPU_LINK_STREAM)) return fail(ctx, SDFGPU_ERR_INVALID, "SDFGPU_LINK_ROUNDS and SDFGPU_LINK_STREAM exclude each other");
    L.rank = rank; L.world = world;
    L.max_pixels = max_width * max_height;
    L.want_gbuf = (flags & SDFGPU_LINK_GBUF) != 0;
    L.halo_push = (flags & SDFGPU_LINK_HALO_PUSH) != 0;
    L.timeout_ms = (uint32_t)ctx->opt_link_timeout_ms;
    const ArenaLayout lay = arena_layout(L.max_pixels, L.want_gbuf);
    CK(ctx, cudaMalloc(&L.arena, lay.total));
    L.arena_bytes = lay.total;
    // header and ray queues: the stream trace recognises an entry by its tag, which is never 0
    CK(ctx, cudaMemsetAsync(L.arena, 0, lay.keys[0], ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaHostAlloc(reinterpret_cast<void**>(&L.timed_out_host), sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
    *L.timed_out_host = 0u;
    LinkBlob b;
    memset(&b, 0, sizeof b);
    b.magic = LINK_MAGIC; b.rank = rank; b.world = world; b.flags = flags;
    {
        cudaDeviceProp prop;
        CK(ctx, cudaGetDeviceProperties(&prop, ctx->device));
        static_assert(sizeof prop.uuid == sizeof b.uuid, "device uuid is 16 bytes");
        memcpy(b.uuid, &prop.uuid, sizeof b.uuid);
    }
    b.pid = (uint64_t)getpid();
    b.device = (uint32_t)ctx->device;
    b.z_begin = ctx->z_begin; b.z_end = ctx->z_end; b.z_lo = ctx->z_lo; b.z_hi = ctx->z_hi;
    b.max_pixels = L.max_pixels;
    memcpy(b.dims, ctx->dims, sizeof b.dims);
    b.arena_bytes = lay.total;
    b.p_tex0 = (uint64_t)(uintptr_t)ctx->tex0; b.p_tex1 = (uint64_t)(uintptr_t)ctx->tex1; b.p_arena = (uint64_t)(uintptr_t)L.arena;
    // IPC handles are only needed by other processes; a failure here (e.g. a platform without IPC) surfaces at attach
    if (cudaIpcGetMemHandle(&b.h_tex0, ctx->tex0) != cudaSuccess || cudaIpcGetMemHandle(&b.h_tex1, ctx->tex1) != cudaSuccess ||
        cudaIpcGetMemHandle(&b.h_arena, L.arena) != cudaSuccess) {
        (void)cudaGetLastError();
        b.flags |= 0x80000000u;  // no IPC handles
    }
    memset(blob, 0, SDFGPU_LINK_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    ctx->peers_ever = true;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_link_attach(sdfgpu_ctx* ctx, const void* blobs, uint32_t world) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    LinkState& L = ctx->link;
    if (!L.arena) return fail(ctx, SDFGPU_ERR_STATE, "call sdfgpu_link_export first");
    if (L.on) return fail(ctx, SDFGPU_ERR_STATE, "already attached");
    if (!blobs || world != L.world) return fail(ctx, SDFGPU_ERR_INVALID, "blobs of all %u ranks expected", L.world);
    set_device(ctx);
    std::vector<LinkBlob> bs(world);
    uint32_t z = 0;
    for (uint32_t r = 0; r < world; ++r) {
        memcpy(&bs[r], (const unsigned char*)blobs + (size_t)r * SDFGPU_LINK_BLOB_BYTES, sizeof(LinkBlob));
        const LinkBlob& b = bs[r];
        if (b.magic != LINK_MAGIC || b.rank != r || b.world != world) return fail(ctx, SDFGPU_ERR_INVALID, "blob %u is not rank %u's", r, r);
        if (b.max_pixels != L.max_pixels || ((b.flags ^ (L.want_gbuf ? SDFGPU_LINK_GBUF : 0u)) & SDFGPU_LINK_GBUF))
            return fail(ctx, SDFGPU_ERR_INVALID, "rank %u was exported with other frame parameters", r);
        if (((b.flags & SDFGPU_LINK_HALO_PUSH) != 0) != L.halo_push)
            return fail(ctx, SDFGPU_ERR_INVALID, "rank %u was exported with another halo mode (SDFGPU_LINK_HALO_PUSH)", r);
        if (memcmp(b.dims, ctx->dims, sizeof b.dims)) return fail(ctx, SDFGPU_ERR_INVALID, "rank %u has another grid", r);
        if (b.z_begin != z || b.z_end <= b.z_begin) return fail(ctx, SDFGPU_ERR_INVALID, "the slabs must tile the grid in rank order, none empty (rank %u owns [%u,%u))", r, b.z_begin, b.z_end);
        z = b.z_end;
    }
    if (z != ctx->dims[2]) return fail(ctx, SDFGPU_ERR_INVALID, "the slabs end at slice %u of %u", z, ctx->dims[2]);
    if (bs[L.rank].p_arena != (uint64_t)(uintptr_t)L.arena) return fail(ctx, SDFGPU_ERR_INVALID, "blob %u is not this handle's", L.rank);
    // the stream trace's kernels wait for each other: every rank needs a device of its own
    bool distinct = true, any_rounds = false, any_stream = false;
    for (uint32_t r = 0; r < world; ++r) {
        any_rounds |= (bs[r].flags & SDFGPU_LINK_ROUNDS) != 0;
        any_stream |= (bs[r].flags & SDFGPU_LINK_STREAM) != 0;
        for (uint32_t q = 0; q < r; ++q) distinct &= memcmp(bs[r].uuid, bs[q].uuid, sizeof bs[r].uuid) != 0;
    }
    if (any_stream && (!distinct || any_rounds))
        return fail(ctx, SDFGPU_ERR_INVALID, "SDFGPU_LINK_STREAM needs every rank on a device of its own, and no rank asking for SDFGPU_LINK_ROUNDS");
    const uint64_t pid = (uint64_t)getpid();
    auto map = [&](const LinkBlob& b, uint64_t raw, const cudaIpcMemHandle_t& h, void** out, bool* ipc) -> int {
        *ipc = false;
        if (b.pid == pid) {  // the other handle lives in this process: its pointer is valid here (unified addressing)
            if ((int)b.device != ctx->device) {
                int can = 0;
                CK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, (int)b.device));
                if (!can) return fail(ctx, SDFGPU_ERR_CUDA, "device %d cannot access device %u", ctx->device, b.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess((int)b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    (void)cudaGetLastError();
                    return fail(ctx, SDFGPU_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
                }
                (void)cudaGetLastError();
            }
            *out = (void*)(uintptr_t)raw;
            return SDFGPU_OK;
        }
        if (b.flags & 0x80000000u) return fail(ctx, SDFGPU_ERR_CUDA, "rank %u could not export CUDA IPC handles", b.rank);
        CK(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
        *ipc = true;
        return SDFGPU_OK;
    };
    int rc = SDFGPU_OK;
    L.nb[0] = L.rank > 0 ? (int)L.rank - 1 : -1;
    L.nb[1] = L.rank + 1 < world ? (int)L.rank + 1 : -1;
    for (uint32_t r = 0; r < world && rc == SDFGPU_OK; ++r) {
        if (r == L.rank) { L.peer_arena[r] = L.arena; continue; }
        void* p = nullptr;
        rc = map(bs[r], bs[r].p_arena, bs[r].h_arena, &p, &L.peer_arena_ipc[r]);
        L.peer_arena[r] = (unsigned char*)p;
    }
    for (int side = 0; side < 2 && rc == SDFGPU_OK && L.halo_push; ++side) {  // the neighbours' volumes: targets of the halo push
        if (L.nb[side] < 0) continue;
        const LinkBlob& b = bs[L.nb[side]];
        void *p0 = nullptr, *p1 = nullptr;
        bool i0 = false, i1 = false;
        if ((rc = map(b, b.p_tex0, b.h_tex0, &p0, &i0)) != SDFGPU_OK) break;
        ctx->peer_tex0[side] = (float4*)p0;
        L.peer_tex_ipc[side] = i0;
        if ((rc = map(b, b.p_tex1, b.h_tex1, &p1, &i1)) != SDFGPU_OK) break;
        ctx->peer_tex1[side] = (float4*)p1;
        ctx->peer_z_lo[side] = b.z_lo;
    }
    if (rc != SDFGPU_OK) {
        const std::string why = ctx->err;
        const uint32_t rank = L.rank, w = L.world, mp = L.max_pixels, to = L.timeout_ms;
        const bool gb = L.want_gbuf, hp = L.halo_push;
        unsigned char* arena = L.arena;
        const size_t ab = L.arena_bytes;
        uint32_t* toh = L.timed_out_host;
        L.arena = nullptr;  // keep the exported arena: the caller may retry or detach
        L.timed_out_host = nullptr;
        L.on = true;        // so that link_free clears the neighbour pointers
        link_free(ctx);
        L.arena = arena; L.arena_bytes = ab; L.rank = rank; L.world = w; L.max_pixels = mp; L.want_gbuf = gb; L.halo_push = hp;
        L.timed_out_host = toh; L.timeout_ms = to;
        ctx->err = why;
        return rc;
    }
    L.memops = wait32() != nullptr && ctx->opt_link_wait != 1;
    L.stream = distinct && !any_rounds && world > 1;
    L.on = true;
    // push mode: the slab is filled as [z_begin, z_end) and the halo slices come from the neighbours; otherwise every
    // rank fills its stored slices [z_lo, z_hi), halo slices included
    const int fill_halo = L.halo_push ? 0 : 1;
    if (ctx->opt_fill_halo != fill_halo && ctx->known_step != 0) ctx->known_step = -1;
    ctx->opt_fill_halo = fill_halo;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_link_detach(sdfgpu_ctx* ctx) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    set_device(ctx);
    link_free(ctx);
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_trace_linked(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height, int want_gbuf,
                                   uint8_t* rgba8, float* depth, float* gbuf) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    int rc = link_trace_begin(ctx, cam, width, height, want_gbuf != 0);
    if (rc != SDFGPU_OK) return rc;
    if ((rc = link_trace_issue(ctx)) != SDFGPU_OK) { ctx->link.in_frame = false; return rc; }
    return link_trace_end(ctx, rgba8, depth, gbuf, true);
}

SDFGPU_API int sdfgpu_trace_linked_device(sdfgpu_ctx* ctx, const sdfgpu_camera* cam, uint32_t width, uint32_t height,
                                          int want_gbuf, uint8_t** rgba8_dev, float** depth_dev, float** gbuf_dev) {
    if (!ctx) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL ctx");
    if (rgba8_dev) *rgba8_dev = nullptr;
    if (depth_dev) *depth_dev = nullptr;
    if (gbuf_dev) *gbuf_dev = nullptr;
    int rc = link_trace_begin(ctx, cam, width, height, want_gbuf != 0);
    if (rc != SDFGPU_OK) return rc;
    if ((rc = link_trace_issue(ctx)) != SDFGPU_OK) { ctx->link.in_frame = false; return rc; }
    if ((rc = link_trace_end(ctx, nullptr, nullptr, nullptr, false)) != SDFGPU_OK) return rc;
    if (ctx->link.rank == 0) {
        if (rgba8_dev) *rgba8_dev = reinterpret_cast<uint8_t*>(ctx->rgba8_dev);
        if (depth_dev) *depth_dev = ctx->depth_dev;
        if (gbuf_dev && want_gbuf)
            *gbuf_dev = reinterpret_cast<float*>(ctx->link.arena + arena_layout(ctx->link.max_pixels, ctx->link.want_gbuf).gbuf);
    }
    return SDFGPU_OK;
}

// ------------------------------------------------------------------------ single-process group of slabs

struct sdfgpu_group {
    std::vector<sdfgpu_ctx*> ranks;
    std::string err;
};

namespace {

int gfail(sdfgpu_group* g, sdfgpu_ctx* c, int rc) {
    if (g) g->err = c ? sdfgpu_last_error(c) : sdfgpu_last_error(nullptr);
    return rc;
}

}  // namespace

SDFGPU_API int sdfgpu_group_create(const float bb[6], const uint32_t voxels[3], uint32_t loading_passes, const int* devices,
                                   uint32_t n_devices, uint32_t max_width, uint32_t max_height, uint32_t flags, sdfgpu_group** out) {
    if (!out) return fail(nullptr, SDFGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!bb || !voxels || !devices) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL argument");
    if (n_devices < 1 || n_devices > LINK_MAX_WORLD) return fail(nullptr, SDFGPU_ERR_INVALID, "1 to %u devices", LINK_MAX_WORLD);
    if (voxels[2] < n_devices) return fail(nullptr, SDFGPU_ERR_INVALID, "fewer z slices (%u) than devices (%u)", voxels[2], n_devices);
    sdfgpu_group* g = new (std::nothrow) sdfgpu_group();
    if (!g) return fail(nullptr, SDFGPU_ERR_INVALID, "out of host memory");
    int rc = SDFGPU_OK;
    std::vector<unsigned char> blobs((size_t)n_devices * SDFGPU_LINK_BLOB_BYTES);
    for (uint32_t r = 0; r < n_devices && rc == SDFGPU_OK; ++r) {
        const uint32_t zb = (uint32_t)(((uint64_t)r * voxels[2]) / n_devices), ze = (uint32_t)(((uint64_t)(r + 1) * voxels[2]) / n_devices);
        sdfgpu_ctx* c = nullptr;
        rc = sdfgpu_create_slab(bb, voxels, loading_passes, devices[r], zb, ze, &c);
        if (rc != SDFGPU_OK) break;
        g->ranks.push_back(c);
        if (n_devices > 1)
            rc = sdfgpu_link_export(c, r, n_devices, max_width, max_height, flags, blobs.data() + (size_t)r * SDFGPU_LINK_BLOB_BYTES,
                                    SDFGPU_LINK_BLOB_BYTES);
        if (rc != SDFGPU_OK) fail(nullptr, rc, "%s", sdfgpu_last_error(c));
    }
    for (uint32_t r = 0; r < n_devices && rc == SDFGPU_OK && n_devices > 1; ++r) {
        rc = sdfgpu_link_attach(g->ranks[r], blobs.data(), n_devices);
        if (rc != SDFGPU_OK) fail(nullptr, rc, "%s", sdfgpu_last_error(g->ranks[r]));
    }
    if (rc != SDFGPU_OK) {
        const std::string why = sdfgpu_last_error(nullptr);
        sdfgpu_group_destroy(g);
        return fail(nullptr, rc, "%s", why.c_str());
    }
    *out = g;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_create_mask(const float bb[6], uint32_t max_voxels_side, uint32_t loading_passes, uint32_t device_mask,
                                        uint32_t max_width, uint32_t max_height, sdfgpu_group** out) {
    uint32_t dims[3];
    const int rc = sdfgpu_dims_from_bb(bb, max_voxels_side, dims);
    if (rc != SDFGPU_OK) return rc;
    int devices[32];
    uint32_t n = 0;
    for (int d = 0; d < 32; ++d)
        if (device_mask & (1u << d)) devices[n++] = d;
    if (n == 0) return fail(nullptr, SDFGPU_ERR_INVALID, "device_mask selects no device");
    return sdfgpu_group_create(bb, dims, loading_passes, devices, n, max_width, max_height, 0, out);
}

SDFGPU_API void sdfgpu_group_destroy(sdfgpu_group* g) {
    if (!g) return;
    // every rank stops using its peers' memory before any of it is freed
    for (sdfgpu_ctx* c : g->ranks) (void)sdfgpu_sync(c);
    for (sdfgpu_ctx* c : g->ranks) (void)sdfgpu_link_detach(c);
    for (sdfgpu_ctx* c : g->ranks) sdfgpu_destroy(c);
    delete g;
}

SDFGPU_API uint32_t sdfgpu_group_size(const sdfgpu_group* g) { return g ? (uint32_t)g->ranks.size() : 0; }
SDFGPU_API sdfgpu_ctx* sdfgpu_group_rank(sdfgpu_group* g, uint32_t rank) { return g && rank < g->ranks.size() ? g->ranks[rank] : nullptr; }
SDFGPU_API const char* sdfgpu_group_last_error(const sdfgpu_group* g) { return g ? g->err.c_str() : sdfgpu_last_error(nullptr); }

#define GROUP_EACH(call)                                              \
    do {                                                              \
        if (!g) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL group"); \
        for (sdfgpu_ctx* c : g->ranks) {                              \
            const int rc_ = (call);                                   \
            if (rc_ != SDFGPU_OK) return gfail(g, c, rc_);            \
        }                                                             \
    } while (0)

SDFGPU_API int sdfgpu_group_set_tape(sdfgpu_group* g, const void* tape, size_t tape_bytes) {
    GROUP_EACH(sdfgpu_set_tape(c, tape, tape_bytes));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_fill_all(sdfgpu_group* g) {
    GROUP_EACH(sdfgpu_fill_all(c));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_update(sdfgpu_group* g, const float* changed_box, uint32_t max_passes, uint64_t* iterations) {
    if (iterations) *iterations = 0;
    uint64_t it = 0;
    GROUP_EACH(sdfgpu_update(c, changed_box, max_passes, &it));
    if (iterations) *iterations = it;  // every rank walks the same LoadingManager
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_update_surface(sdfgpu_group* g, const sdfgpu_surface* sdf, double max_delta_seconds, uint64_t* iterations) {
    if (iterations) *iterations = 0;
    if (!g) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL group");
    if (!sdf) return fail(nullptr, SDFGPU_ERR_INVALID, "sdf is NULL");
    // sdf.changed() reports a box ONCE (src/sdf/mod.rs:87): poll it here and hand the same answer to every rank
    float box[6];
    struct Once { const sdfgpu_surface* s; int has; float box[6]; } once{sdf, 0, {0, 0, 0, 0, 0, 0}};
    once.has = sdf->changed ? sdf->changed(sdf->self, box) : 0;
    if (once.has) memcpy(once.box, box, sizeof box);
    sdfgpu_surface proxy = *sdf;
    proxy.self = &once;
    proxy.changed = [](void* self, float out_box[6]) -> int {
        Once* o = (Once*)self;
        if (o->has) memcpy(out_box, o->box, sizeof o->box);
        return o->has;
    };
    // the other callbacks get the surface's own `self` back through trampolines
    proxy.bounding_box = sdf->bounding_box ? +[](void* self, float out_bb[6]) { Once* o = (Once*)self; o->s->bounding_box(o->s->self, out_bb); } : nullptr;
    proxy.sample = sdf->sample ? +[](void* self, const float p[3], int d, float out[7]) { Once* o = (Once*)self; o->s->sample(o->s->self, p, d, out); } : nullptr;
    proxy.sample_batch = sdf->sample_batch ? +[](void* self, const float* xyz, uint64_t n, int d, float* out) { Once* o = (Once*)self; o->s->sample_batch(o->s->self, xyz, n, d, out); } : nullptr;
    proxy.tape = sdf->tape ? +[](void* self, const void** bytes, size_t* len) -> int { Once* o = (Once*)self; return o->s->tape(o->s->self, bytes, len); } : nullptr;
    uint64_t it = 0;
    const double each = max_delta_seconds / (double)g->ranks.size();  // the host samples for one rank after the other
    GROUP_EACH(sdfgpu_update_surface(c, &proxy, each, &it));
    if (iterations) *iterations = it;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_resample_box(sdfgpu_group* g, const float box[6], uint64_t* voxels_touched) {
    if (voxels_touched) *voxels_touched = 0;
    uint64_t total = 0, n = 0;
    GROUP_EACH((n = 0, sdfgpu_resample_box(c, box, voxels_touched ? &n : nullptr) == SDFGPU_OK ? (total += n, SDFGPU_OK) : SDFGPU_ERR_CUDA));
    if (voxels_touched) *voxels_touched = total;
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_commit(sdfgpu_group* g) {
    GROUP_EACH(sdfgpu_commit(c));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_reset(sdfgpu_group* g, uint32_t loading_passes) {
    GROUP_EACH(sdfgpu_reset(c, loading_passes));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_set_option(sdfgpu_group* g, const char* key, int64_t value) {
    GROUP_EACH(sdfgpu_set_option(c, key, value));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_sync(sdfgpu_group* g) {
    GROUP_EACH(sdfgpu_sync(c));
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_loading_state(const sdfgpu_group* g, uint64_t* len, uint64_t* total_iterations, uint32_t* passes_left,
                                          uint32_t* passes) {
    if (!g || g->ranks.empty()) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL group");
    return sdfgpu_loading_state(g->ranks[0], len, total_iterations, passes_left, passes);
}

SDFGPU_API int sdfgpu_group_download(sdfgpu_group* g, float* tex0, float* tex1) {
    if (!g) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL group");
    for (sdfgpu_ctx* c : g->ranks) {
        const size_t off = (size_t)c->z_begin * c->dims[0] * c->dims[1] * 4;
        const int rc = sdfgpu_download(c, tex0 ? tex0 + off : nullptr, tex1 ? tex1 + off : nullptr);
        if (rc != SDFGPU_OK) return gfail(g, c, rc);
    }
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_trace(sdfgpu_group* g, const sdfgpu_camera* cam, uint32_t width, uint32_t height, uint8_t* rgba8,
                                  float* depth, float* gbuf) {
    if (!g || g->ranks.empty()) return fail(nullptr, SDFGPU_ERR_INVALID, "NULL group");
    if (g->ranks.size() == 1) {  // one device: the plain handle
        sdfgpu_ctx* c = g->ranks[0];
        int rc = SDFGPU_OK;
        if (gbuf) rc = sdfgpu_trace(c, cam, width, height, nullptr, nullptr, gbuf);
        if (rc == SDFGPU_OK) rc = sdfgpu_trace_rgba8(c, cam, width, height, rgba8, depth);
        return rc == SDFGPU_OK ? rc : gfail(g, c, rc);
    }
    // round-major order: what a rank waits for has always been enqueued before (streams of handles that share a
    // device may share a hardware queue)
    for (sdfgpu_ctx* c : g->ranks) {
        const int rc = link_trace_begin(c, cam, width, height, gbuf != nullptr);
        if (rc != SDFGPU_OK) {
            for (sdfgpu_ctx* d : g->ranks) d->link.in_frame = false;
            return gfail(g, c, rc);
        }
    }
    if (g->ranks[0]->link.stream) {  // every rank on its own device: one launch each, side by side
        for (sdfgpu_ctx* c : g->ranks) {
            const int rc = link_trace_stream(c);
            if (rc != SDFGPU_OK) return gfail(g, c, rc);
        }
    } else {
        for (size_t k = 0; k < g->ranks.size(); ++k)
            for (sdfgpu_ctx* c : g->ranks) {
                const int rc = link_trace_round(c);
                if (rc != SDFGPU_OK) return gfail(g, c, rc);
            }
    }
    for (size_t r = g->ranks.size(); r-- > 0;) {  // the presenter (rank 0) last: its end synchronises
        sdfgpu_ctx* c = g->ranks[r];
        const int rc = link_trace_end(c, rgba8, depth, gbuf, r == 0);
        if (rc != SDFGPU_OK) return gfail(g, c, rc);
    }
    return SDFGPU_OK;
}

SDFGPU_API int sdfgpu_group_trace_rgba8(sdfgpu_group* g, const sdfgpu_camera* cam, uint32_t width, uint32_t height, uint8_t* rgba8,
                                        float* depth) {
    return sdfgpu_group_trace(g, cam, width, height, rgba8, depth, nullptr);
}

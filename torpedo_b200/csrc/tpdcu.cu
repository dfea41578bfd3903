// tpdcu.cu — the C ABI of libtpdcu.so (include/tpdcu.h): context, device buffers, frame orchestration.
//
// This is the CUDA re-statement of the *orchestration half* of tpd::GaussianEngine
// (torpedo/volumetric/src/GaussianEngine.cpp): compile() -> upload_gaussians, rasterFrame() -> raster,
// draw() -> read_frame / external-memory output. What differs by design:
//   * no mid-frame GPU->CPU read-back of tilesRendered (GaussianEngine.cpp:662-674): P stays on the device,
//     launches are sized by the grow-only pair capacity and an overflowing frame is re-rendered lazily;
//   * a two-level sort: the visible Gaussians by depth, duplication in that order, then the pairs by tile only (sort.cu)
//     instead of one 23-pass sort of every pair (GaussianEngine.cpp:822-841);
//   * 14 launches per frame at 1080p (12 of them replayed from a CUDA graph) instead of 2 + 1 + 92 + 2 dispatches with 95 pipeline
//     barriers (GaussianEngine.cpp:777-863);
//   * several frames in flight, like the reference's Frame objects (GaussianEngine.h:104-117, SurfaceRenderer.h:66): every
//     frame slot owns its splat arrays, pair buffers and a private stream, so the memory-bound front of frame k+1
//     (preprocess, sort) overlaps the SM-bound blend of frame k. Only the blend is ordered after the caller's stream
//     (it is what writes the target); the caller's stream then waits for the frame.
#include "../../include/tpdcu.h"
#include "common.cuh"

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace tpdcu;

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CK(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            const int code_ = (e_ == cudaErrorMemoryAllocation) ? TPDCU_ERR_OOM : TPDCU_ERR_CUDA;                 \
            cudaGetLastError();                                                                                   \
            return fail(code_, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                                         \
    } while (0)

struct PlanHead {  // head of a SortPlan as the frame left it
    uint32_t n, num_passes, final_sel, passes_run, bias, total_bits;
};
static_assert(offsetof(SortPlan, total_bits) == 5 * sizeof(uint32_t), "PlanHead mirrors the head of SortPlan");
struct FrameStatus {  // pinned host mirror of what a frame reports back
    uint32_t pairs_total;
    uint32_t visible;
    unsigned long long pairs64;
    PlanHead tile, depth;
};

constexpr int MAX_SLOTS = 4;
constexpr uint32_t TICKET_RING = 256;  // frames that may be enqueued between two host-side checks (tpdcu_finish & co.)

// What it takes to render a frame again: kept for every frame enqueued since the last check, because P is only looked at
// lazily (no mid-frame read-back) and an overflowing frame has to be repeated after its buffers have grown.
struct FrameTicket {
    float ubo[TPDCU_CAMERA_FLOATS];
    uint32_t sh_degree;
    uint8_t* out;
    size_t pitch;
    cudaStream_t user_stream;
    uint32_t ran_capacity;  // pair capacity of the slot it ran on
    uint32_t status;        // index into the pinned status ring
    int slot;
};

struct GraphSig {  // everything a captured frame bakes in
    const void* zero_region; size_t zero_bytes; const void* keys0; const void* keys1; const void* geo; const void* posop;
    uint32_t n, capacity, width, height, sh_degree, entity_count;
    bool operator==(const GraphSig& o) const { return memcmp(this, &o, sizeof(GraphSig)) == 0; }
};

// One frame in flight: the per-frame twin of GaussianEngine::Frame (GaussianEngine.h:104-117).
struct FrameSlot {
    // per-Gaussian outputs of the preprocess stage
    SplatGeo* geo = nullptr;
    float4* color = nullptr;
    float2* depth_radius = nullptr;
    uint2* rect = nullptr;
    uint32_t* offsets = nullptr;
    uint64_t* depth_words[2] = { nullptr, nullptr };  // visible Gaussians: depth << 32 | index, ping-pong of the depth sort
    uint32_t n_alloc = 0;
    // camera-derived constants and model matrices
    FrameCam* cam = nullptr;
    float* models = nullptr;
    float* vm = nullptr;
    float* pm = nullptr;
    uint32_t entity_alloc = 0;
    uint64_t models_version = 0;
    // sort / raster state
    SortPlan* plan = nullptr;        // tile sort (and the standalone pair sort)
    SortPlan* depth_plan = nullptr;  // depth sort
    uint8_t* zero_region = nullptr;
    size_t zero_bytes = 0, off_scan_desc = 0, off_emit_desc = 0, off_ranges = 0, off_lb_depth = 0, off_lb_tile = 0;
    uint32_t zero_n = 0, zero_capacity = 0, zero_tiles = 0;
    size_t zero_lb_tile_bytes = 0;
    uint32_t capacity = 0;
    uint64_t* keys[2] = { nullptr, nullptr };  // pair words tile << 32 | index, ping-pong of the tile sort
    uint32_t* vals[2] = { nullptr, nullptr };  // standalone pair sort only
    uint32_t vals_capacity = 0;
    // execution
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, done = nullptr, read_done = nullptr;
    bool read_pending = false;  // an asynchronous read of `target` has been enqueued since the slot last rendered into it
    bool graph_valid = false;
    GraphSig graph_sig;
    cudaGraphExec_t graph_exec = nullptr;
    uint8_t* target = nullptr;  // internal render target of this slot (used when the caller has not bound one)
    size_t target_bytes = 0;
    bool busy = false;  // something has been enqueued on `stream` since it was last waited for
};

struct tpdcu_ctx {
    int device = 0;
    int sm_count = 0;
    std::string device_name;

    // scene (shared by the frame slots, read-only while rendering)
    uint32_t n = 0, entity_count = 0;
    float4* posop = nullptr;
    float4* cov_a = nullptr;
    float2* cov_b = nullptr;
    float4* sh = nullptr;
    uint32_t* entity = nullptr;
    std::vector<float> models_host;
    uint64_t models_version = 1;

    FrameSlot slots[MAX_SLOTS];
    int frames_in_flight = 3;
    int next_slot = 0;

    bool use_graph = true;
    uint32_t graph_launches = 0, graph_captures = 0;
    cudaStream_t capture_stream = nullptr;

    // target
    uint32_t* tile_cost = nullptr;       // per tile: splats the blend consumed in an earlier frame (scheduling hint only)
    uint32_t tile_cost_tiles = 0;
    uint32_t width = 0, height = 0;
    uint8_t* bound_out = nullptr;
    size_t bound_pitch = 0;
    cudaExternalMemory_t ext_mem = nullptr;

    FrameStatus* status = nullptr;       // pinned ring, TICKET_RING entries
    std::vector<FrameTicket> unchecked;  // frames enqueued since the last host-side check, oldest first
    uint32_t next_status = 0;
    FrameTicket newest{};                // the most recent frame: what finish / read_* / introspection refer to
    bool have_newest = false;
    uint32_t frames_repeated = 0;        // frames rendered again because they overflowed (grow-only buffers: warm-up only)

    bool timing = false;
    bool stop_after_emit = false;        // introspection (tpdcu_read_emitted): frames end after the duplication stage
    cudaEvent_t ev[9] = {};
    float stage_ms[TPDCU_NUM_STAGES] = {};

    // standalone sort
    cudaEvent_t sort_ev[2] = {};
    cudaStream_t sort_stream = nullptr;
    bool sort_done = false;
};

static uint32_t bit_length(uint32_t v) {
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}
static uint32_t tiles_of(const tpdcu_ctx* c) {
    return ((c->width + TILE_PX - 1) / TILE_PX) * ((c->height + TILE_PX - 1) / TILE_PX);
}
static uint32_t tile_bits_of(const tpdcu_ctx* c) {
    const uint32_t tiles = tiles_of(c);
    return tiles > 1 ? bit_length(tiles - 1) : 0u;
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void drop_graph(FrameSlot& f) {
    if (f.graph_exec) cudaGraphExecDestroy(f.graph_exec);
    f.graph_exec = nullptr;
    f.graph_valid = false;
}

static void free_slot_scene(FrameSlot& f) {
    cudaFree(f.geo); cudaFree(f.color); cudaFree(f.depth_radius); cudaFree(f.rect); cudaFree(f.offsets);
    cudaFree(f.depth_words[0]); cudaFree(f.depth_words[1]);
    cudaFree(f.models); cudaFree(f.vm); cudaFree(f.pm);
    f.geo = nullptr; f.color = nullptr; f.depth_radius = nullptr; f.rect = nullptr; f.offsets = nullptr;
    f.depth_words[0] = f.depth_words[1] = nullptr;
    f.models = f.vm = f.pm = nullptr;
    f.n_alloc = 0; f.entity_alloc = 0; f.models_version = 0;
    drop_graph(f);
}

static void free_slot_pairs(FrameSlot& f) {
    for (int i = 0; i < 2; ++i) { cudaFree(f.keys[i]); cudaFree(f.vals[i]); f.keys[i] = nullptr; f.vals[i] = nullptr; }
    f.capacity = 0;
    f.vals_capacity = 0;
    drop_graph(f);
}

static void free_scene(tpdcu_ctx* c) {
    cudaFree(c->posop); cudaFree(c->cov_a); cudaFree(c->cov_b); cudaFree(c->sh); cudaFree(c->entity);
    c->posop = c->cov_a = nullptr; c->cov_b = nullptr; c->sh = nullptr; c->entity = nullptr;
    c->n = 0; c->entity_count = 0;
    for (auto& f : c->slots) free_slot_scene(f);
}

static int ensure_slot_scene(tpdcu_ctx* c, FrameSlot& f) {
    if (f.n_alloc != c->n || f.entity_alloc != c->entity_count) {
        free_slot_scene(f);
        CK(cudaMalloc(&f.geo, (size_t)c->n * sizeof(SplatGeo)));
        CK(cudaMalloc(&f.color, (size_t)c->n * sizeof(float4)));
        CK(cudaMalloc(&f.depth_radius, (size_t)c->n * sizeof(float2)));
        CK(cudaMalloc(&f.rect, (size_t)c->n * sizeof(uint2)));
        // whole sort tiles: the onesweep passes prefetch and pad by tile
        for (int i = 0; i < 2; ++i) CK(cudaMalloc(&f.depth_words[i], (align_up(c->n, SORT_TILE) + 8) * sizeof(uint64_t)));  // + 8: a tile's bulk copy may read one word past its last
        CK(cudaMalloc(&f.offsets, ((size_t)c->n + 1) * sizeof(uint32_t)));
        CK(cudaMalloc(&f.models, (size_t)c->entity_count * 16 * sizeof(float)));
        CK(cudaMalloc(&f.vm, (size_t)c->entity_count * 16 * sizeof(float)));
        CK(cudaMalloc(&f.pm, (size_t)c->entity_count * 16 * sizeof(float)));
        f.n_alloc = c->n;
        f.entity_alloc = c->entity_count;
    }
    return TPDCU_OK;
}

static int ensure_pairs(FrameSlot& f, uint32_t want) {
    if (want <= f.capacity) return TPDCU_OK;
    // grow-only (GaussianEngine.cpp:671-674,793-804), rounded to whole sort tiles
    const uint64_t cap64 = align_up(want, SORT_TILE);
    if (cap64 > 0xffffffffull - SORT_TILE) return fail(TPDCU_ERR_INVALID, "pair capacity exceeds 2^32");
    free_slot_pairs(f);
    const uint32_t cap = (uint32_t)cap64;
    for (int i = 0; i < 2; ++i) CK(cudaMalloc(&f.keys[i], ((size_t)cap + 8) * sizeof(uint64_t)));  // + 8: a tile's bulk copy may read one word past its last
    f.capacity = cap;
    return TPDCU_OK;
}

static int ensure_vals(FrameSlot& f) {  // value buffers of the standalone pair sort
    if (f.vals_capacity == f.capacity && f.vals[0]) return TPDCU_OK;
    for (int i = 0; i < 2; ++i) { cudaFree(f.vals[i]); f.vals[i] = nullptr; }
    for (int i = 0; i < 2; ++i) CK(cudaMalloc(&f.vals[i], (size_t)f.capacity * sizeof(uint32_t)));
    f.vals_capacity = f.capacity;
    return TPDCU_OK;
}

// The per-frame zeroed region: FrameCtl | preprocess scan descriptors | duplication scan descriptors | tile ranges + blend order |
// look-back arrays of the depth sort | look-back arrays of the tile sort (last: the standalone sort may need more passes)
static size_t lookback_bytes(const FrameSlot& f, uint32_t passes, uint32_t kind) {
    return align_up((size_t)std::max(passes, 1u) * sort_parts(f.capacity, kind) * SORT_BINS * sizeof(uint32_t), 256);
}
static int ensure_zero_region(tpdcu_ctx* c, FrameSlot& f, size_t lb_tile_bytes) {
    const uint32_t tiles = tiles_of(c);
    if (f.zero_region && f.zero_n == c->n && f.zero_capacity == f.capacity && f.zero_tiles == tiles && f.zero_lb_tile_bytes >= lb_tile_bytes)
        return TPDCU_OK;
    lb_tile_bytes = std::max(lb_tile_bytes, f.zero_capacity == f.capacity ? f.zero_lb_tile_bytes : (size_t)0);
    cudaFree(f.zero_region);
    f.zero_region = nullptr;
    drop_graph(f);
    const uint32_t pre_parts = (c->n + PRE_PART - 1) / PRE_PART;
    size_t off = align_up(sizeof(FrameCtl), 256);
    f.off_scan_desc = off; off = align_up(off + (size_t)pre_parts * sizeof(uint64_t), 256);
    f.off_emit_desc = off; off = align_up(off + (size_t)emit_parts(c->n) * sizeof(uint64_t), 256);
    f.off_ranges = off;    off = align_up(off + (size_t)tiles * 4 * sizeof(uint32_t), 256);  // ranges (x2) | blend order | its buckets
    f.off_lb_depth = off;  off = align_up(off + (size_t)sort_passes_for(32) * sort_parts(c->n, SORT_KIND_DEPTH) * SORT_BINS * sizeof(uint32_t), 256);
    f.off_lb_tile = off;   off += lb_tile_bytes;
    CK(cudaMalloc(&f.zero_region, off));
    f.zero_bytes = off;
    f.zero_n = c->n; f.zero_capacity = f.capacity; f.zero_tiles = tiles; f.zero_lb_tile_bytes = lb_tile_bytes;
    return TPDCU_OK;
}
// bytes of the zeroed region a launch whose pair-buffer look-back arrays take `lb_tile_bytes` touches
static size_t zero_bytes_for(const FrameSlot& f, size_t lb_tile_bytes) { return std::min(f.zero_bytes, f.off_lb_tile + lb_tile_bytes); }

static int ensure_status(tpdcu_ctx* c) {
    if (c->status) return TPDCU_OK;
    CK(cudaMallocHost(&c->status, sizeof(FrameStatus) * TICKET_RING));
    memset(c->status, 0, sizeof(FrameStatus) * TICKET_RING);
    return TPDCU_OK;
}

static int ensure_target(tpdcu_ctx* c, FrameSlot& f) {
    const size_t need = (size_t)c->width * c->height * 4;
    if (need <= f.target_bytes && f.target) return TPDCU_OK;
    cudaFree(f.target);
    f.target = nullptr;
    CK(cudaMalloc(&f.target, std::max<size_t>(need, 4)));
    f.target_bytes = need;
    return TPDCU_OK;
}

static int sync_slots(tpdcu_ctx* c) {
    for (auto& f : c->slots)
        if (f.stream) CK(cudaStreamSynchronize(f.stream));
    return TPDCU_OK;
}

// Blend cost hints, one per tile, shared by the frame slots (they only steer the dispatch order). A new framebuffer size
// starts without hints; ensure_zero_region drops the slots' graphs in that case, so no captured launch keeps the old pointer.
static int ensure_tile_cost(tpdcu_ctx* c) {
    const uint32_t tiles = tiles_of(c);
    if (c->tile_cost && c->tile_cost_tiles == tiles) return TPDCU_OK;
    if (int r = sync_slots(c)) return r;
    cudaFree(c->tile_cost);
    c->tile_cost = nullptr;
    CK(cudaMalloc(&c->tile_cost, (size_t)std::max(tiles, 1u) * sizeof(uint32_t)));
    CK(cudaMemset(c->tile_cost, 0, (size_t)std::max(tiles, 1u) * sizeof(uint32_t)));
    CK(cudaDeviceSynchronize());  // the slots' streams do not synchronise with the stream the memset ran on
    c->tile_cost_tiles = tiles;
    for (auto& f : c->slots) drop_graph(f);
    return TPDCU_OK;
}

struct FrameLaunch {
    PreprocessLaunch pre;
    SortLaunch depth_sort;
    EmitLaunch emit;
    SortLaunch tile_sort;
    RasterLaunch raster;
    size_t zero_bytes;
};

// The part of a frame whose launch parameters do not change from frame to frame: everything between the camera setup and
// the blend. Either enqueued directly or captured once into a CUDA graph and replayed.
static int enqueue_middle(tpdcu_ctx* c, FrameSlot& f, const FrameLaunch& l, cudaStream_t s, bool timing, bool stop_after_emit = false) {
    CK(cudaMemsetAsync(f.zero_region, 0, l.zero_bytes, s));
    if (timing) CK(cudaEventRecord(c->ev[1], s));
    CK(launch_preprocess(l.pre, s));   // the SH colour is evaluated by the blend, for the splats it stages
    if (timing) CK(cudaEventRecord(c->ev[2], s));
    CK(launch_sort(l.depth_sort, 0, s, nullptr));
    if (timing) CK(cudaEventRecord(c->ev[3], s));
    CK(launch_emit(l.emit, s));
    if (timing) CK(cudaEventRecord(c->ev[4], s));
    if (stop_after_emit) return TPDCU_OK;  // tpdcu_read_emitted: the pair words as the duplication stage left them
    CK(launch_sort(l.tile_sort, 0, s, timing ? c->ev[8] : nullptr));
    if (timing) CK(cudaEventRecord(c->ev[5], s));
    CK(launch_ranges(l.raster, f.capacity, s));
    if (timing) CK(cudaEventRecord(c->ev[6], s));
    return TPDCU_OK;
}

// Enqueue the frame described by the ticket on slot `f`. No host synchronisation. With stage timing everything runs on the
// caller's stream; otherwise the front of the frame runs on the slot's stream unordered with the caller's stream, the blend
// waits for the caller's stream (it writes the target) and the caller's stream waits for the frame.
static int enqueue_frame(tpdcu_ctx* c, FrameSlot& f, FrameTicket& tk) {
    if (int r = ensure_slot_scene(c, f)) return r;
    if (f.capacity == 0)
        if (int r = ensure_pairs(f, SORT_TILE)) return r;
    const uint32_t tile_bits = tile_bits_of(c);
    const uint32_t tile_passes = sort_passes_for(tile_bits);
    const size_t lb_tile_bytes = lookback_bytes(f, tile_passes, SORT_KIND_TILE);
    if (int r = ensure_zero_region(c, f, lb_tile_bytes)) return r;
    if (int r = ensure_tile_cost(c)) return r;
    const bool t = c->timing;
    cudaStream_t user = tk.user_stream;
    cudaStream_t s = t ? user : f.stream;

    if (f.models_version != c->models_version) {
        CK(cudaMemcpyAsync(f.models, c->models_host.data(), sizeof(float) * 16 * c->entity_count, cudaMemcpyHostToDevice, s));
        f.models_version = c->models_version;
    }

    FrameCtl* ctl = reinterpret_cast<FrameCtl*>(f.zero_region);
    FrameLaunch l{};
    l.zero_bytes = zero_bytes_for(f, lb_tile_bytes);
    PreprocessLaunch& p = l.pre;
    p.scene = SceneArrays{ c->posop, c->cov_a, c->cov_b, c->sh, c->entity_count > 1 ? c->entity : nullptr, c->n, c->entity_count };
    p.models = f.models; p.cam = f.cam; p.vm = f.vm; p.pm = f.pm;
    p.ctl = ctl;
    p.scan_desc = reinterpret_cast<uint64_t*>(f.zero_region + f.off_scan_desc);
    p.out = SplatArrays{ f.geo, f.color, f.depth_radius, f.rect, f.offsets };
    p.depth_words = f.depth_words[0];
    p.width = c->width; p.height = c->height; p.sh_degree = std::min(tk.sh_degree, 3u);  // GaussianEngine.cpp:366-370

    SortLaunch& ds = l.depth_sort;
    ds.keys[0] = f.depth_words[0]; ds.keys[1] = f.depth_words[1];
    ds.frame = ctl; ds.ctl = &ctl->depth_sort; ds.plan = f.depth_plan;
    ds.lookback = reinterpret_cast<uint32_t*>(f.zero_region + f.off_lb_depth);
    ds.kind = SORT_KIND_DEPTH; ds.capacity = c->n; ds.end_bit = 32; ds.tile_bits = tile_bits; ds.sm_count = c->sm_count;

    EmitLaunch& em = l.emit;
    em.depth_words[0] = f.depth_words[0]; em.depth_words[1] = f.depth_words[1];
    em.depth_plan = f.depth_plan; em.rect = f.rect; em.ctl = ctl;
    em.scan_desc = reinterpret_cast<uint64_t*>(f.zero_region + f.off_emit_desc);
    em.keys = f.keys[0]; em.n = c->n; em.capacity = f.capacity; em.width = c->width; em.tile_bits = tile_bits;

    SortLaunch& ts = l.tile_sort;
    ts.keys[0] = f.keys[0]; ts.keys[1] = f.keys[1];
    ts.frame = ctl; ts.ctl = &ctl->tile_sort; ts.plan = f.plan;
    ts.lookback = reinterpret_cast<uint32_t*>(f.zero_region + f.off_lb_tile);
    // whole digits: the top depth bits that do not fill a digit of the depth sort ride in the pair key (DepthSplit)
    ts.kind = SORT_KIND_TILE; ts.capacity = f.capacity; ts.end_bit = tile_passes * SORT_RADIX_BITS; ts.tile_bits = tile_bits; ts.sm_count = c->sm_count;

    RasterLaunch& ra = l.raster;
    ra.keys[0] = f.keys[0]; ra.keys[1] = f.keys[1];
    ra.plan = f.plan; ra.geo = f.geo; ra.color = f.color; ra.depth_radius = f.depth_radius;
    ra.ranges = reinterpret_cast<uint32_t*>(f.zero_region + f.off_ranges);
    ra.order = ra.ranges + 2 * (size_t)tiles_of(c);
    ra.tile_cost = c->tile_cost;
    ra.posop = c->posop; ra.sh = c->sh; ra.cam = f.cam; ra.sh_degree = l.pre.sh_degree;
    ra.out = tk.out; ra.pitch = tk.pitch; ra.width = c->width; ra.height = c->height; ra.sm_count = c->sm_count;

    if (t) CK(cudaEventRecord(c->ev[0], s));
    CameraUbo cu;
    memcpy(cu.f, tk.ubo, sizeof(cu.f));  // by-value kernel argument: no per-frame H2D copy (updateCameraBuffer, :764-775)
    CK(launch_setup(p, cu, s));

    bool replayed = false;
    const bool stop = c->stop_after_emit;
    if (c->use_graph && !t && !stop) {
        GraphSig sig;
        memset(&sig, 0, sizeof(sig));
        sig.zero_region = f.zero_region; sig.zero_bytes = l.zero_bytes; sig.keys0 = f.keys[0]; sig.keys1 = f.keys[1];
        sig.geo = f.geo; sig.posop = c->posop; sig.n = c->n; sig.capacity = f.capacity; sig.width = c->width; sig.height = c->height;
        sig.sh_degree = p.sh_degree; sig.entity_count = c->entity_count;
        if (!(f.graph_valid && f.graph_sig == sig)) {
            // (re)capture on a private stream: nothing executes during capture
            drop_graph(f);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(c->capture_stream, cudaStreamCaptureModeRelaxed);
            int rc = TPDCU_OK;
            if (e == cudaSuccess) {
                rc = enqueue_middle(c, f, l, c->capture_stream, false);
                e = cudaStreamEndCapture(c->capture_stream, &graph);
            }
            if (e == cudaSuccess && rc == TPDCU_OK && graph) e = cudaGraphInstantiate(&f.graph_exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e == cudaSuccess && rc == TPDCU_OK && f.graph_exec) {
                f.graph_valid = true;
                f.graph_sig = sig;
                ++c->graph_captures;
            } else {
                cudaGetLastError();
                drop_graph(f);
                c->use_graph = false;  // capture is unavailable in this process: keep launching directly
            }
        }
        if (f.graph_valid) {
            CK(cudaGraphLaunch(f.graph_exec, s));
            ++c->graph_launches;
            replayed = true;
        }
    }
    if (!replayed)
        if (int r = enqueue_middle(c, f, l, s, t, stop)) return r;

    if (!t && !stop) {
        if (tk.out == f.target) {
            // the slot's own target: the only other user is an asynchronous read of the frame it held before
            // (tpdcu_read_frame_async); the blend does not have to wait for the newer frames' copies on the caller's stream
            if (f.read_pending) CK(cudaStreamWaitEvent(s, f.read_done, 0));
            f.read_pending = false;
        } else {  // the blend writes the caller's memory: order it after whatever the caller's stream did with it
            CK(cudaEventRecord(f.fork, user));
            CK(cudaStreamWaitEvent(s, f.fork, 0));
        }
    }
    if (!stop) CK(launch_blend(ra, s));
    if (t) CK(cudaEventRecord(c->ev[7], s));

    FrameStatus* st = &c->status[tk.status];
    CK(cudaMemcpyAsync(&st->pairs_total, &ctl->pairs_total, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&st->pairs64, &ctl->pairs64, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&st->tile, f.plan, sizeof(PlanHead), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&st->depth, f.depth_plan, sizeof(PlanHead), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(f.done, s));
    if (!t) CK(cudaStreamWaitEvent(user, f.done, 0));
    tk.ran_capacity = f.capacity;
    f.busy = true;
    return TPDCU_OK;
}

static uint32_t grown_capacity(uint32_t pairs) {
    const uint64_t want = (uint64_t)pairs + pairs / 8 + SORT_TILE;  // 12.5 % head-room against view changes
    return (uint32_t)std::min<uint64_t>(want, 0xffffffffull - 2 * SORT_TILE);
}

static int check_ready(tpdcu_ctx* c) {
    if (!c) return fail(TPDCU_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    return TPDCU_OK;
}

static int wait_slots(tpdcu_ctx* c) {
    for (auto& f : c->slots)
        if (f.busy) {
            CK(cudaEventSynchronize(f.done));
            f.busy = false;
        }
    return TPDCU_OK;
}

static int finish_internal(tpdcu_ctx* c);

// Enqueue one frame on the next slot and remember how to repeat it.
static int raster_one(tpdcu_ctx* c, const float* ubo, uint32_t sh_degree, cudaStream_t user, uint8_t* out, size_t pitch) {
    if (c->unchecked.size() >= TICKET_RING - 2)  // the status ring is full: look at what is pending before going on
        if (int r = finish_internal(c)) return r;
    const int depth = c->timing ? 1 : c->frames_in_flight;
    if (c->next_slot >= depth) c->next_slot = 0;
    const int slot = c->next_slot;
    c->next_slot = (c->next_slot + 1) % depth;
    FrameTicket tk{};
    memcpy(tk.ubo, ubo, sizeof(tk.ubo));
    tk.sh_degree = sh_degree;
    if (!out) {  // no caller-owned target: every frame slot renders into its own internal one
        if (int r = ensure_target(c, c->slots[slot])) return r;
        out = c->slots[slot].target;
        pitch = (size_t)c->width * 4;
    }
    tk.out = out;
    tk.pitch = pitch;
    tk.user_stream = user;
    tk.slot = slot;
    tk.status = c->next_status;
    c->next_status = (c->next_status + 1) % (TICKET_RING - 1);  // the last entry belongs to the standalone sort
    if (int r = enqueue_frame(c, c->slots[slot], tk)) return r;
    c->unchecked.push_back(tk);
    c->newest = tk;
    c->have_newest = true;
    return TPDCU_OK;
}

// Host-side check of everything enqueued since the last check: wait for the frames; if one overflowed its pair buffers,
// grow them and render that frame AND EVERY LATER ONE again, in their original order — unless a later frame went to the same
// target, in which case the frame has been superseded and repeating it would clobber the newer image. Repeating the later
// frames too keeps two promises: the newest frame stays the newest (finish / read_* / introspection refer to it), and its
// intermediate buffers are the ones of the (possibly re-allocated) slot it last ran on.
static int finish_internal(tpdcu_ctx* c) {
    if (!c->have_newest) return fail(TPDCU_ERR_STATE, "no frame has been rendered");
    for (int attempt = 0; !c->unchecked.empty(); ++attempt) {
        if (attempt >= 6) return fail(TPDCU_ERR_STATE, "frames kept overflowing their buffers");
        if (int r = wait_slots(c)) return r;
        std::vector<FrameTicket> pending;
        pending.swap(c->unchecked);
        std::vector<FrameTicket> redo;
        uint32_t max_pairs = 0;
        size_t first_overflow = pending.size();
        for (size_t i = 0; i < pending.size(); ++i) {
            const FrameStatus& st = c->status[pending[i].status];
            if (st.pairs64 >= 0xffffffffull - 2ull * SORT_TILE)
                return fail(TPDCU_ERR_INVALID, "a frame produced " + std::to_string(st.pairs64) + " (tile, Gaussian) pairs: the pair buffers hold fewer than 2^32");
            if (st.pairs_total <= pending[i].ran_capacity) continue;
            max_pairs = std::max(max_pairs, st.pairs_total);
            first_overflow = std::min(first_overflow, i);
        }
        for (size_t i = first_overflow; i < pending.size(); ++i) {
            bool superseded = false;
            for (size_t j = i + 1; j < pending.size() && !superseded; ++j) superseded = pending[j].out == pending[i].out;
            if (!superseded) redo.push_back(pending[i]);
        }
        if (max_pairs)
            for (int k = 0; k < MAX_SLOTS; ++k)
                if (c->slots[k].capacity || k < c->frames_in_flight)
                    if (int r = ensure_pairs(c->slots[k], grown_capacity(max_pairs))) return r;
        c->frames_repeated += (uint32_t)redo.size();
        for (const FrameTicket& tk : redo) {
            CK(cudaStreamSynchronize(tk.user_stream));
            c->next_slot = tk.slot;  // every slot is idle here: the frame runs again where it ran (its internal target lives there)
            if (int r = raster_one(c, tk.ubo, tk.sh_degree, tk.user_stream, tk.out, tk.pitch)) return r;
        }
    }
    if (int r = wait_slots(c)) return r;
    if (c->timing) {
        CK(cudaEventSynchronize(c->ev[7]));
        float ms;
        // ev: 0 start | 1 after clear+setup | 2 after preprocess+colour | 3 after the depth sort | 4 after duplication |
        //     5 after the tile sort | 6 after ranges | 7 after blend
        for (int k = 0; k < 7; ++k) { CK(cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1])); c->stage_ms[k] = ms; }
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[7]));
        c->stage_ms[7] = ms;
        c->stage_ms[8] = (float)c->status[c->newest.status].depth.passes_run;
        c->stage_ms[9] = (float)c->status[c->newest.status].tile.passes_run;
        CK(cudaEventElapsedTime(&ms, c->ev[4], c->ev[8]));
        c->stage_ms[10] = ms;
    }
    return TPDCU_OK;
}

static FrameSlot& last(tpdcu_ctx* c) { return c->slots[c->newest.slot]; }
static const FrameStatus& last_status(tpdcu_ctx* c) { return c->status[c->newest.status]; }

static void forget_frames(tpdcu_ctx* c) {
    c->unchecked.clear();
    c->have_newest = false;
    c->next_slot = 0;
}

// ================================================================================================
// C ABI
// ================================================================================================

extern "C" {

const char* tpdcu_last_error(void) { return g_last_error.c_str(); }

int tpdcu_create(int device, tpdcu_ctx** out) {
    if (!out) return fail(TPDCU_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(TPDCU_ERR_INVALID, "no such CUDA device");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TPDCU_ERR_CUDA, std::string("libtpdcu is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
                                        std::to_string(prop.minor) + " (" + prop.name + "); there is no fallback path");
    CK(cudaSetDevice(device));
    tpdcu_ctx* c = new tpdcu_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->device_name = prop.name;
    auto bail = [&](int r) { tpdcu_destroy(c); return r; };
#define CKB(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return bail(fail(TPDCU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
    } while (0)
    for (auto& f : c->slots) {
        memset(&f.graph_sig, 0, sizeof(f.graph_sig));
        CKB(cudaMalloc(&f.cam, sizeof(FrameCam)));
        CKB(cudaMalloc(&f.plan, sizeof(SortPlan)));
        CKB(cudaMemset(f.plan, 0, sizeof(SortPlan)));
        CKB(cudaMalloc(&f.depth_plan, sizeof(SortPlan)));
        CKB(cudaMemset(f.depth_plan, 0, sizeof(SortPlan)));
        CKB(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking));
        CKB(cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming));
        CKB(cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
        CKB(cudaEventCreateWithFlags(&f.read_done, cudaEventDisableTiming));
    }
    CKB(cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev) CKB(cudaEventCreate(&e));
    for (auto& e : c->sort_ev) CKB(cudaEventCreate(&e));
#undef CKB
    if (cudaError_t e = init_sort_attributes()) return bail(fail(TPDCU_ERR_CUDA, std::string("init_sort_attributes: ") + cudaGetErrorString(e)));
    if (int r = ensure_status(c)) return bail(r);
    {   // the onesweep ranking needs lane-ordered returning shared-memory atomics (sort.cu: rank_selftest_kernel)
        uint32_t mismatches = 0;
        if (cudaError_t e = sort_rank_selftest(&mismatches)) return bail(fail(TPDCU_ERR_CUDA, std::string("sort_rank_selftest: ") + cudaGetErrorString(e)));
        if (mismatches != 0)
            return bail(fail(TPDCU_ERR_CUDA, "this device's shared-memory atomics do not return lane-ordered values (" + std::to_string(mismatches) +
                                                 " lanes differ): the stable ranking of the radix sort cannot run on it"));
    }
    *out = c;
    return TPDCU_OK;
}

void tpdcu_destroy(tpdcu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_scene(c);
    for (auto& f : c->slots) {
        free_slot_pairs(f);
        cudaFree(f.cam); cudaFree(f.plan); cudaFree(f.depth_plan); cudaFree(f.zero_region); cudaFree(f.target);
        if (f.stream) cudaStreamDestroy(f.stream);
        if (f.fork) cudaEventDestroy(f.fork);
        if (f.done) cudaEventDestroy(f.done);
        if (f.read_done) cudaEventDestroy(f.read_done);
    }
    cudaFree(c->tile_cost);
    if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
    if (c->ext_mem) cudaDestroyExternalMemory(c->ext_mem);
    if (c->status) cudaFreeHost(c->status);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->sort_ev) if (e) cudaEventDestroy(e);
    delete c;
}

int tpdcu_device_info(tpdcu_ctx* c, char* buf, size_t buf_bytes, int* sm_count) {
    if (int r = check_ready(c)) return r;
    if (buf && buf_bytes) snprintf(buf, buf_bytes, "sm_100a / %s / %d SMs", c->device_name.c_str(), c->sm_count);
    if (sm_count) *sm_count = c->sm_count;
    return TPDCU_OK;
}

static int upload_common(tpdcu_ctx* c, const void* d_recs, uint32_t n, const uint32_t* d_entity, uint32_t entity_count,
                         cudaStream_t s) {
    // The new scene is allocated BESIDE the old one and only then swapped in, so that a failed compile() leaves the engine
    // with the scene it had; if the two do not fit together the old one is dropped first (and a failure then leaves none).
    float4 *posop = nullptr, *cov_a = nullptr, *sh = nullptr;
    float2* cov_b = nullptr;
    uint32_t* entity = nullptr;
    auto alloc_all = [&]() -> cudaError_t {
        cudaError_t e = cudaMalloc(&posop, (size_t)n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&cov_a, (size_t)n * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&cov_b, (size_t)n * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc(&sh, (size_t)n * SH_PLANES * sizeof(float4));
        if (e == cudaSuccess && entity_count > 1) e = cudaMalloc(&entity, (size_t)n * sizeof(uint32_t));
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(posop); cudaFree(cov_a); cudaFree(cov_b); cudaFree(sh); cudaFree(entity);
            posop = cov_a = sh = nullptr; cov_b = nullptr; entity = nullptr;
        }
        return e;
    };
    cudaError_t e = alloc_all();
    if (e == cudaErrorMemoryAllocation && c->n != 0) {
        free_scene(c);
        forget_frames(c);
        e = alloc_all();
    }
    CK(e);
    free_scene(c);
    forget_frames(c);
    c->posop = posop; c->cov_a = cov_a; c->cov_b = cov_b; c->sh = sh; c->entity = entity;
    if (entity_count > 1) {
        if (d_entity) CK(cudaMemcpyAsync(c->entity, d_entity, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        else CK(cudaMemsetAsync(c->entity, 0, (size_t)n * sizeof(uint32_t), s));
    }
    c->n = n;
    c->entity_count = entity_count;
    c->models_host.assign((size_t)entity_count * 16, 0.0f);  // identity (createBindlessTransformBuffer)
    for (uint32_t e2 = 0; e2 < entity_count; ++e2)
        for (int k = 0; k < 4; ++k) c->models_host[(size_t)e2 * 16 + k * 5] = 1.0f;
    ++c->models_version;
    CompileLaunch cl{ reinterpret_cast<const float*>(d_recs), c->posop, c->cov_a, c->cov_b, c->sh, n };
    CK(launch_compile_scene(cl, s));
    return TPDCU_OK;
}

int tpdcu_upload_gaussians(tpdcu_ctx* c, const void* recs240, uint32_t n, const uint32_t* entity_idx, uint32_t entity_count) {
    if (int r = check_ready(c)) return r;
    if (n == 0) return TPDCU_OK;  // GaussianEngine.cpp:362-365: warn and return
    if (!recs240) return fail(TPDCU_ERR_INVALID, "recs240 is null");
    if (entity_count == 0) entity_count = 1;
    if (entity_idx)
        for (uint32_t i = 0; i < n; ++i)
            if (entity_idx[i] >= entity_count) return fail(TPDCU_ERR_INVALID, "entity index out of range");
    CK(cudaDeviceSynchronize());
    void* staging = nullptr;
    uint32_t* d_entity = nullptr;
    CK(cudaMalloc(&staging, (size_t)n * TPDCU_GAUSSIAN_BYTES));
    cudaError_t e = cudaMemcpy(staging, recs240, (size_t)n * TPDCU_GAUSSIAN_BYTES, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && entity_idx && entity_count > 1) {
        e = cudaMalloc(&d_entity, (size_t)n * 4);
        if (e == cudaSuccess) e = cudaMemcpy(d_entity, entity_idx, (size_t)n * 4, cudaMemcpyHostToDevice);
    }
    int r = TPDCU_OK;
    if (e != cudaSuccess) r = fail(TPDCU_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
    if (r == TPDCU_OK) r = upload_common(c, staging, n, d_entity, entity_count, nullptr);
    cudaDeviceSynchronize();
    cudaFree(staging);
    cudaFree(d_entity);
    return r;
}

int tpdcu_upload_gaussians_device(tpdcu_ctx* c, const void* d_recs240, uint32_t n, const uint32_t* d_entity_idx,
                                  uint32_t entity_count, void* stream) {
    if (int r = check_ready(c)) return r;
    if (n == 0) return TPDCU_OK;
    if (!d_recs240) return fail(TPDCU_ERR_INVALID, "d_recs240 is null");
    if (entity_count == 0) entity_count = 1;
    CK(cudaDeviceSynchronize());
    const int r = upload_common(c, d_recs240, n, d_entity_idx, entity_count, (cudaStream_t)stream);
    // the frame slots run on private streams: the compiled scene must be complete before any of them starts
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return r;
}

int tpdcu_set_transform(tpdcu_ctx* c, uint32_t entity, const float m[16]) {
    if (int r = check_ready(c)) return r;
    if (c->n == 0) return fail(TPDCU_ERR_STATE, "no scene compiled");
    if (entity >= c->entity_count || !m) return fail(TPDCU_ERR_INVALID, "bad entity or matrix");
    memcpy(&c->models_host[(size_t)entity * 16], m, sizeof(float) * 16);
    ++c->models_version;
    return TPDCU_OK;
}

int tpdcu_resize(tpdcu_ctx* c, uint32_t width, uint32_t height) {
    if (int r = check_ready(c)) return r;
    if (width == 0 || height == 0 || width > 65535u * TILE_PX || height > 65535u * TILE_PX)
        return fail(TPDCU_ERR_INVALID, "bad framebuffer size");
    if (int r = sync_slots(c)) return r;
    forget_frames(c);
    c->width = width;
    c->height = height;
    return TPDCU_OK;
}

int tpdcu_bind_output_device_ptr(tpdcu_ctx* c, void* d_rgba8, size_t pitch_bytes) {
    if (int r = check_ready(c)) return r;
    if (d_rgba8 && pitch_bytes < (size_t)c->width * 4) return fail(TPDCU_ERR_INVALID, "pitch smaller than a row");
    if (d_rgba8 && (((uintptr_t)d_rgba8 | pitch_bytes) & 3u)) return fail(TPDCU_ERR_INVALID, "output must be 4-byte aligned");
    c->bound_out = reinterpret_cast<uint8_t*>(d_rgba8);
    c->bound_pitch = pitch_bytes;
    return TPDCU_OK;
}

int tpdcu_bind_output_fd(tpdcu_ctx* c, int fd, size_t bytes) {
    if (int r = check_ready(c)) return r;
    if (fd < 0 || bytes < (size_t)c->width * c->height * 4) return fail(TPDCU_ERR_INVALID, "bad fd or allocation too small");
    cudaExternalMemoryHandleDesc hd{};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = bytes;
    hd.flags = cudaExternalMemoryDedicated;  // torpedo allocates targets with VMA's DEDICATED_MEMORY_BIT (VmaUsage.cpp:28-42)
    cudaExternalMemory_t mem;
    cudaError_t ie = cudaImportExternalMemory(&mem, &hd);
    if (ie != cudaSuccess) {  // an exporter whose allocation is not a dedicated one (a failed import leaves the fd with the caller)
        cudaGetLastError();
        hd.flags = 0;
        ie = cudaImportExternalMemory(&mem, &hd);
    }
    CK(ie);
    cudaExternalMemoryBufferDesc bd{};
    bd.offset = 0;
    bd.size = bytes;
    void* ptr = nullptr;
    cudaError_t e = cudaExternalMemoryGetMappedBuffer(&ptr, mem, &bd);
    if (e != cudaSuccess) {
        cudaDestroyExternalMemory(mem);
        return fail(TPDCU_ERR_CUDA, std::string("cudaExternalMemoryGetMappedBuffer: ") + cudaGetErrorString(e));
    }
    if (int r = sync_slots(c)) return r;
    if (c->ext_mem) cudaDestroyExternalMemory(c->ext_mem);
    c->ext_mem = mem;
    c->bound_out = reinterpret_cast<uint8_t*>(ptr);
    c->bound_pitch = (size_t)c->width * 4;
    return TPDCU_OK;
}

// ---- one frame array in the collecting GPU's memory, mapped into every process of the box (tpdcu.h) ------------------
static_assert(sizeof(cudaIpcMemHandle_t) == TPDCU_IPC_HANDLE_BYTES, "the ABI passes the CUDA IPC handle as 64 opaque bytes");

int tpdcu_ipc_frames_create(int device, size_t bytes, void** d_frames, unsigned char handle[TPDCU_IPC_HANDLE_BYTES]) {
    if (!d_frames || !handle || bytes == 0) return fail(TPDCU_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(device));
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));   // a whole allocation of its own: the handle names the allocation, not an offset into one
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        return fail(TPDCU_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    *d_frames = p;
    return TPDCU_OK;
}

int tpdcu_ipc_frames_open(int device, const unsigned char handle[TPDCU_IPC_HANDLE_BYTES], void** d_frames) {
    if (!d_frames || !handle) return fail(TPDCU_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_frames = p;
    return TPDCU_OK;
}

int tpdcu_ipc_frames_close(int device, void* d_frames) {
    if (!d_frames) return TPDCU_OK;
    CK(cudaSetDevice(device));
    CK(cudaDeviceSynchronize());   // no kernel of this process may still be storing into the mapping
    CK(cudaIpcCloseMemHandle(d_frames));
    return TPDCU_OK;
}

int tpdcu_ipc_frames_destroy(int device, void* d_frames) {
    if (!d_frames) return TPDCU_OK;
    CK(cudaSetDevice(device));
    CK(cudaFree(d_frames));
    return TPDCU_OK;
}

int tpdcu_raster(tpdcu_ctx* c, const float camera_ubo[TPDCU_CAMERA_FLOATS], uint32_t sh_degree, void* stream) {
    if (int r = check_ready(c)) return r;
    if (!camera_ubo) return fail(TPDCU_ERR_INVALID, "camera_ubo is null");
    if (c->n == 0) return fail(TPDCU_ERR_STATE, "no scene compiled");
    if (c->width == 0) return fail(TPDCU_ERR_STATE, "tpdcu_resize has not been called");
    return raster_one(c, camera_ubo, sh_degree, (cudaStream_t)stream, c->bound_out, c->bound_pitch);
}

int tpdcu_raster_views(tpdcu_ctx* c, const float* camera_ubos, uint32_t n_views, uint32_t sh_degree, void* d_frames,
                       size_t frame_stride_bytes, void* stream) {
    if (int r = check_ready(c)) return r;
    if (!camera_ubos || !d_frames || n_views == 0) return fail(TPDCU_ERR_INVALID, "bad batch arguments");
    if (c->n == 0) return fail(TPDCU_ERR_STATE, "no scene compiled");
    if (c->width == 0) return fail(TPDCU_ERR_STATE, "tpdcu_resize has not been called");
    const size_t pitch = (size_t)c->width * 4;
    if (frame_stride_bytes < pitch * c->height) return fail(TPDCU_ERR_INVALID, "frame stride smaller than a frame");
    const bool timing = c->timing;
    if (timing) CK(cudaDeviceSynchronize());
    c->timing = false;  // per-stage events describe single frames only
    int rc = TPDCU_OK;
    for (uint32_t v = 0; v < n_views && rc == TPDCU_OK; ++v)
        rc = raster_one(c, camera_ubos + (size_t)v * TPDCU_CAMERA_FLOATS, sh_degree, (cudaStream_t)stream,
                        reinterpret_cast<uint8_t*>(d_frames) + (size_t)v * frame_stride_bytes, pitch);
    if (rc == TPDCU_OK) rc = finish_internal(c);  // every view has its own target: each one that overflowed is rendered again
    if (rc == TPDCU_OK) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) rc = fail(TPDCU_ERR_CUDA, std::string("batch sync: ") + cudaGetErrorString(e));
    }
    c->timing = timing;
    return rc;
}

int tpdcu_finish(tpdcu_ctx* c, uint32_t* pairs) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    if (pairs) *pairs = last_status(c).pairs_total;
    return TPDCU_OK;
}

int tpdcu_read_frame(tpdcu_ctx* c, void* host_rgba8, size_t host_pitch_bytes) {
    if (int r = check_ready(c)) return r;
    if (!host_rgba8 || host_pitch_bytes < (size_t)c->width * 4) return fail(TPDCU_ERR_INVALID, "bad host buffer");
    if (int r = finish_internal(c)) return r;
    FrameSlot& f = last(c);
    CK(cudaMemcpy2DAsync(host_rgba8, host_pitch_bytes, c->newest.out, c->newest.pitch, (size_t)c->width * 4, c->height, cudaMemcpyDeviceToHost, f.stream));
    CK(cudaStreamSynchronize(f.stream));
    return TPDCU_OK;
}

int tpdcu_read_frame_async(tpdcu_ctx* c, void* host_rgba8, size_t host_pitch_bytes, void* stream) {
    if (int r = check_ready(c)) return r;
    if (!host_rgba8 || host_pitch_bytes < (size_t)c->width * 4) return fail(TPDCU_ERR_INVALID, "bad host buffer");
    if (!c->have_newest) return fail(TPDCU_ERR_STATE, "no frame has been rendered");
    // `stream` must be the stream the frame was rastered with (it already waits for the frame); no host synchronisation here
    // cudaMemcpyDefault: the destination may be host memory or device memory — in particular another GPU's frame array
    // (tpdcu_ipc_frames_open), which a copy engine fills over NVLink while the SMs render the next frame
    const size_t row = (size_t)c->width * 4;
    if (host_pitch_bytes == row && c->newest.pitch == row)
        CK(cudaMemcpyAsync(host_rgba8, c->newest.out, row * c->height, cudaMemcpyDefault, (cudaStream_t)stream));
    else
        CK(cudaMemcpy2DAsync(host_rgba8, host_pitch_bytes, c->newest.out, c->newest.pitch, row, c->height, cudaMemcpyDefault,
                             (cudaStream_t)stream));
    FrameSlot& f = last(c);
    if (c->newest.out == f.target) {  // the next frame of this slot must not overwrite the target before the copy has read it
        CK(cudaEventRecord(f.read_done, (cudaStream_t)stream));
        f.read_pending = true;
    }
    return TPDCU_OK;
}

int tpdcu_frames_repeated(tpdcu_ctx* c, uint32_t* count) {
    if (int r = check_ready(c)) return r;
    if (c->have_newest)
        if (int r = finish_internal(c)) return r;
    if (count) *count = c->frames_repeated;
    return TPDCU_OK;
}

int tpdcu_get_counts(tpdcu_ctx* c, uint32_t* pairs, uint32_t* visible) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    if (pairs) *pairs = last_status(c).pairs_total;
    if (visible) *visible = last_status(c).visible;
    return TPDCU_OK;
}

int tpdcu_read_splats(tpdcu_ctx* c, void* host_splats48, uint32_t n) {
    if (int r = check_ready(c)) return r;
    if (!host_splats48 || n > c->n) return fail(TPDCU_ERR_INVALID, "bad splat read");
    if (int r = finish_internal(c)) return r;
    if (n == 0) return TPDCU_OK;
    FrameSlot& f = last(c);
    {   // The frame itself computes colours on demand inside the blend; the reference's Splat records carry the colour of
        // every visible Gaussian, so the introspection export evaluates them all first (same arithmetic: sh_basis / sh_accumulate of common.cuh).
        PreprocessLaunch p{};
        p.scene = SceneArrays{ c->posop, c->cov_a, c->cov_b, c->sh, c->entity_count > 1 ? c->entity : nullptr, c->n, c->entity_count };
        p.cam = f.cam;
        p.out = SplatArrays{ f.geo, f.color, f.depth_radius, f.rect, f.offsets };
        p.width = c->width; p.height = c->height; p.sh_degree = std::min(c->newest.sh_degree, 3u);
        CK(launch_color(p, f.stream));
    }
    void* tmp = nullptr;
    CK(cudaMalloc(&tmp, (size_t)n * TPDCU_SPLAT_BYTES));
    cudaError_t e = launch_export_splats(SplatArrays{ f.geo, f.color, f.depth_radius, f.rect, f.offsets }, n, tmp, f.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_splats48, tmp, (size_t)n * TPDCU_SPLAT_BYTES, cudaMemcpyDeviceToHost, f.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f.stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(TPDCU_ERR_CUDA, std::string("read_splats: ") + cudaGetErrorString(e));
    return TPDCU_OK;
}

static int read_sorted(tpdcu_ctx* c, void* host, uint32_t count, bool want_keys) {
    if (int r = check_ready(c)) return r;
    if (!host) return fail(TPDCU_ERR_INVALID, "host buffer is null");
    if (int r = finish_internal(c)) return r;
    const FrameStatus& st = last_status(c);
    FrameSlot& f = last(c);
    if (count > st.pairs_total) return fail(TPDCU_ERR_INVALID, "count exceeds the frame's pair count");
    if (count == 0) return TPDCU_OK;
    // the frame's result is one array of words tile << 32 | index: expand it into the reference's (key, value) arrays
    RasterLaunch ra{};
    ra.keys[0] = f.keys[0]; ra.keys[1] = f.keys[1]; ra.plan = f.plan; ra.depth_radius = f.depth_radius;
    uint64_t* uk = f.keys[(st.tile.final_sel & 1u) ^ 1u];  // the ping-pong buffer the sort no longer needs
    uint32_t* uv = nullptr;
    CK(cudaMalloc(&uv, (size_t)std::max(st.tile.n, 1u) * 4));
    cudaError_t e = launch_sort_unpack(ra, uk, uv, c->sm_count, f.stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(host, want_keys ? (const void*)uk : (const void*)uv, (size_t)count * (want_keys ? 8 : 4), cudaMemcpyDeviceToHost, f.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f.stream);
    cudaFree(uv);
    if (e != cudaSuccess) return fail(TPDCU_ERR_CUDA, std::string("read_sorted: ") + cudaGetErrorString(e));
    return TPDCU_OK;
}

int tpdcu_read_keys(tpdcu_ctx* c, uint64_t* host_keys, uint32_t count) { return read_sorted(c, host_keys, count, true); }
int tpdcu_read_values(tpdcu_ctx* c, uint32_t* host_vals, uint32_t count) { return read_sorted(c, host_vals, count, false); }

int tpdcu_read_ranges(tpdcu_ctx* c, uint32_t* host_ranges2, uint32_t tile_count) {
    if (int r = check_ready(c)) return r;
    if (!host_ranges2 || tile_count > tiles_of(c)) return fail(TPDCU_ERR_INVALID, "bad range read");
    if (int r = finish_internal(c)) return r;
    FrameSlot& f = last(c);
    CK(cudaMemcpyAsync(host_ranges2, f.zero_region + f.off_ranges, (size_t)tile_count * 8, cudaMemcpyDeviceToHost, f.stream));
    CK(cudaStreamSynchronize(f.stream));
    return TPDCU_OK;
}

int tpdcu_read_unsorted(tpdcu_ctx* c, uint64_t* host_keys, uint32_t* host_vals, uint32_t count) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    FrameSlot& f = last(c);
    if (count > last_status(c).pairs_total) return fail(TPDCU_ERR_INVALID, "count exceeds the frame's pair count");
    if (count == 0) return TPDCU_OK;
    // The frame itself never materialises the reference's unsorted buffers (its duplication runs in depth order): rebuild
    // them from the per-Gaussian records, in the reference's order.
    uint64_t* dk = nullptr;
    uint32_t* dv = nullptr;
    cudaError_t e = cudaMalloc(&dk, (size_t)count * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dv, (size_t)count * 4);
    if (e == cudaSuccess)
        e = launch_export_unsorted(SplatArrays{ f.geo, f.color, f.depth_radius, f.rect, f.offsets }, c->n, c->width, dk, dv, count, f.stream);
    if (e == cudaSuccess && host_keys) e = cudaMemcpyAsync(host_keys, dk, (size_t)count * 8, cudaMemcpyDeviceToHost, f.stream);
    if (e == cudaSuccess && host_vals) e = cudaMemcpyAsync(host_vals, dv, (size_t)count * 4, cudaMemcpyDeviceToHost, f.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f.stream);
    cudaFree(dk);
    cudaFree(dv);
    if (e != cudaSuccess) return fail(TPDCU_ERR_CUDA, std::string("read_unsorted: ") + cudaGetErrorString(e));
    return TPDCU_OK;
}

int tpdcu_read_emitted(tpdcu_ctx* c, uint64_t* host_words, uint32_t count) {
    if (int r = check_ready(c)) return r;
    if (!host_words && count) return fail(TPDCU_ERR_INVALID, "host buffer is null");
    if (int r = finish_internal(c)) return r;
    if (count > last_status(c).pairs_total) return fail(TPDCU_ERR_INVALID, "count exceeds the frame's pair count");
    // The tile sort ping-pongs over the buffer the duplication stage wrote, so its output is gone when a frame ends: run the
    // newest frame again on its slot up to and including emit_kernel, copy the words, then render it once more in full so
    // that every other read still describes a complete frame.
    const FrameTicket tk = c->newest;
    c->stop_after_emit = true;
    c->next_slot = tk.slot;
    int rc = raster_one(c, tk.ubo, tk.sh_degree, tk.user_stream, tk.out, tk.pitch);
    c->stop_after_emit = false;
    if (rc == TPDCU_OK) rc = finish_internal(c);
    if (rc == TPDCU_OK && count) {
        FrameSlot& f = c->slots[tk.slot];
        cudaError_t e = cudaMemcpyAsync(host_words, f.keys[0], (size_t)count * 8, cudaMemcpyDeviceToHost, f.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(f.stream);
        if (e != cudaSuccess) rc = fail(TPDCU_ERR_CUDA, std::string("read_emitted: ") + cudaGetErrorString(e));
    }
    c->next_slot = tk.slot;
    const int rc2 = raster_one(c, tk.ubo, tk.sh_degree, tk.user_stream, tk.out, tk.pitch);
    if (rc == TPDCU_OK) rc = rc2;
    if (rc == TPDCU_OK) rc = finish_internal(c);
    return rc;
}

int tpdcu_enable_stage_timing(tpdcu_ctx* c, int enable) {
    if (int r = check_ready(c)) return r;
    CK(cudaDeviceSynchronize());  // timed frames run on the caller's stream, the others on the slots' streams: do not mix in flight
    c->timing = enable != 0;
    c->next_slot = 0;
    return TPDCU_OK;
}

int tpdcu_stage_times_ms(tpdcu_ctx* c, float times_ms[TPDCU_NUM_STAGES]) {
    if (int r = check_ready(c)) return r;
    if (!times_ms) return fail(TPDCU_ERR_INVALID, "times_ms is null");
    if (!c->timing) return fail(TPDCU_ERR_STATE, "stage timing is not enabled");
    if (int r = finish_internal(c)) return r;
    // 0 clear+setup | 1 preprocess+colour | 2 depth sort | 3 duplication | 4 tile sort | 5 ranges | 6 blend | 7 frame |
    // 8 depth-sort passes run | 9 tile-sort passes run | 10 histogram+plan share of the tile sort
    memcpy(times_ms, c->stage_ms, sizeof(float) * TPDCU_NUM_STAGES);
    return TPDCU_OK;
}

int tpdcu_get_sort_info(tpdcu_ctx* c, uint32_t* depth_bits, uint32_t* depth_passes, uint32_t* tile_bits, uint32_t* tile_passes) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    const FrameStatus& st = last_status(c);
    if (depth_bits) *depth_bits = st.depth.total_bits;
    if (depth_passes) *depth_passes = st.depth.passes_run;
    if (tile_bits) *tile_bits = st.tile.total_bits;
    if (tile_passes) *tile_passes = st.tile.passes_run;
    return TPDCU_OK;
}

int tpdcu_set_graph_replay(tpdcu_ctx* c, int enable, uint32_t* captures, uint32_t* launches) {
    if (int r = check_ready(c)) return r;
    if (enable >= 0) {
        c->use_graph = enable != 0;
        if (!c->use_graph) {
            if (int r = sync_slots(c)) return r;
            for (auto& f : c->slots) drop_graph(f);
        }
    }
    if (captures) *captures = c->graph_captures;
    if (launches) *launches = c->graph_launches;
    return TPDCU_OK;
}

int tpdcu_set_frames_in_flight(tpdcu_ctx* c, int frames) {
    if (int r = check_ready(c)) return r;
    if (frames < 1 || frames > MAX_SLOTS) return fail(TPDCU_ERR_INVALID, "frames in flight must be in [1, 4]");
    if (c->have_newest)
        if (int r = finish_internal(c)) return r;
    if (int r = sync_slots(c)) return r;
    c->frames_in_flight = frames;
    c->next_slot = 0;
    return TPDCU_OK;
}

int tpdcu_get_capacity(tpdcu_ctx* c, uint32_t* capacity_pairs) {
    if (int r = check_ready(c)) return r;
    uint32_t cap = 0;
    for (const auto& f : c->slots) cap = std::max(cap, f.capacity);  // every frame slot grows on its own; report the largest
    if (capacity_pairs) *capacity_pairs = cap;
    return TPDCU_OK;
}

int tpdcu_reserve_pairs(tpdcu_ctx* c, uint32_t capacity_pairs) {
    if (int r = check_ready(c)) return r;
    if (c->have_newest)
        if (int r = finish_internal(c)) return r;
    if (int r = sync_slots(c)) return r;
    CK(cudaDeviceSynchronize());
    const bool regrow_newest = c->have_newest && capacity_pairs > c->slots[c->newest.slot].capacity;
    for (int k = 0; k < c->frames_in_flight; ++k)
        if (int r = ensure_pairs(c->slots[k], capacity_pairs)) return r;
    if (regrow_newest) {
        // the newest frame's pair buffers have just been replaced: render it again so that introspection still describes it
        const FrameTicket tk = c->newest;
        c->next_slot = tk.slot;
        if (int r = raster_one(c, tk.ubo, tk.sh_degree, tk.user_stream, tk.out, tk.pitch)) return r;
        if (int r = finish_internal(c)) return r;
    }
    return TPDCU_OK;
}

int tpdcu_sort_pairs_device(tpdcu_ctx* c, uint64_t* d_keys, uint32_t* d_vals, uint32_t n, uint32_t end_bit, void* stream) {
    if (int r = check_ready(c)) return r;
    if ((!d_keys || !d_vals) && n) return fail(TPDCU_ERR_INVALID, "null device buffers");
    if (end_bit > 64) return fail(TPDCU_ERR_INVALID, "end_bit > 64");
    if (n >= (1u << 30)) return fail(TPDCU_ERR_INVALID, "n must be below 2^30");
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = sync_slots(c)) return r;
    forget_frames(c);  // slot 0's pair buffers are about to be reused
    FrameSlot& f = c->slots[0];
    if (n > f.capacity) {
        CK(cudaDeviceSynchronize());
        if (int r = ensure_pairs(f, n)) return r;
    }
    if (f.capacity == 0)
        if (int r = ensure_pairs(f, SORT_TILE)) return r;
    if (int r = ensure_vals(f)) return r;
    const size_t lb_bytes = lookback_bytes(f, sort_passes_for(end_bit), SORT_KIND_PAIRS);
    if (int r = ensure_zero_region(c, f, lb_bytes)) return r;
    CK(cudaMemsetAsync(f.zero_region, 0, zero_bytes_for(f, lb_bytes), s));
    if (n) {
        CK(cudaMemcpyAsync(f.keys[0], d_keys, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(f.vals[0], d_vals, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    }
    FrameCtl* ctl = reinterpret_cast<FrameCtl*>(f.zero_region);
    SortLaunch so{};
    so.keys[0] = f.keys[0]; so.keys[1] = f.keys[1]; so.vals[0] = f.vals[0]; so.vals[1] = f.vals[1];
    so.frame = ctl; so.ctl = &ctl->tile_sort; so.plan = f.plan;
    so.lookback = reinterpret_cast<uint32_t*>(f.zero_region + f.off_lb_tile);
    so.kind = SORT_KIND_PAIRS; so.capacity = f.capacity; so.end_bit = end_bit; so.sm_count = c->sm_count;
    CK(cudaEventRecord(c->sort_ev[0], s));
    CK(launch_sort(so, n, s, nullptr));
    CK(cudaEventRecord(c->sort_ev[1], s));
    CK(launch_sort_copy_result(so, d_keys, d_vals, n, s));
    CK(cudaMemcpyAsync(&c->status[TICKET_RING - 1].tile, f.plan, sizeof(PlanHead), cudaMemcpyDeviceToHost, s));
    c->sort_stream = s;
    c->sort_done = true;
    return TPDCU_OK;
}

int tpdcu_sort_last_ms(tpdcu_ctx* c, float* ms, uint32_t* passes_run) {
    if (int r = check_ready(c)) return r;
    if (!c->sort_done) return fail(TPDCU_ERR_STATE, "tpdcu_sort_pairs_device has not been called");
    CK(cudaEventSynchronize(c->sort_ev[1]));
    CK(cudaStreamSynchronize(c->sort_stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, c->sort_ev[0], c->sort_ev[1]));
    if (ms) *ms = t;
    if (passes_run) *passes_run = c->status[TICKET_RING - 1].tile.passes_run;
    return TPDCU_OK;
}

}  // extern "C"

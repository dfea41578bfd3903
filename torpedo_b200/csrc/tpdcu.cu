// tpdcu.cu — the C ABI of libtpdcu.so (include/tpdcu.h): context, device buffers, frame orchestration.
//
// This is the CUDA re-statement of the *orchestration half* of tpd::GaussianEngine
// (torpedo/volumetric/src/GaussianEngine.cpp): compile() -> upload_gaussians, rasterFrame() -> raster,
// draw() -> read_frame / external-memory output. What differs by design:
//   * no mid-frame GPU->CPU read-back of tilesRendered (GaussianEngine.cpp:662-674): P stays on the device,
//     launches are sized by the grow-only pair capacity and an overflowing frame is re-rendered lazily;
//   * 1 fused preprocess launch + (2 + <=6) sort launches + 2 raster launches per frame instead of
//     2 + 1 + 92 + 2 dispatches with 95 pipeline barriers (GaussianEngine.cpp:777-863).
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

struct FrameStatus {  // pinned host mirror of what a frame reports back
    uint32_t pairs_total;
    uint32_t visible;
    // head of the SortPlan as the frame left it
    uint32_t n, num_passes, final_sel, passes_run, bias, depth_bits, total_bits, idx_bits, packed, packed_overflow;
};
static_assert(offsetof(SortPlan, packed_overflow) == 9 * sizeof(uint32_t), "FrameStatus mirrors the head of SortPlan");

struct tpdcu_ctx {
    int device = 0;
    int sm_count = 0;
    std::string device_name;

    // scene
    uint32_t n = 0, entity_count = 0;
    float4* posop = nullptr;
    float4* cov_a = nullptr;
    float2* cov_b = nullptr;
    float4* sh = nullptr;
    uint32_t* entity = nullptr;
    SplatGeo* geo = nullptr;
    float4* color = nullptr;
    float2* depth_radius = nullptr;
    uint32_t* offsets = nullptr;
    float* models = nullptr;
    float* vm = nullptr;
    float* pm = nullptr;
    std::vector<float> models_host;
    bool models_dirty = false;

    // per-frame state
    FrameCam* cam = nullptr;
    SortPlan* plan = nullptr;
    uint8_t* zero_region = nullptr;
    size_t zero_bytes = 0;
    size_t off_scan_desc = 0, off_ranges = 0, off_lookback = 0;
    uint32_t zero_n = 0, zero_capacity = 0, zero_tiles = 0;
    uint32_t capacity = 0;
    uint64_t* keys[2] = { nullptr, nullptr };
    uint32_t* vals[2] = { nullptr, nullptr };
    uint32_t packed_word_bits = 64;
    bool packed_disabled = false;  // a frame whose depth range did not fit packed sort words switches the context to pair mode
    bool keep_unsorted = false;
    uint64_t* unsorted_keys = nullptr;
    uint32_t* unsorted_vals = nullptr;
    uint32_t unsorted_capacity = 0;

    // target
    uint32_t width = 0, height = 0;
    uint8_t* target = nullptr;  // internal
    size_t target_bytes = 0;
    uint8_t* bound_out = nullptr;
    size_t bound_pitch = 0;
    cudaExternalMemory_t ext_mem = nullptr;

    // last frame
    bool frame_pending = false, frame_valid = false;
    float last_ubo[TPDCU_CAMERA_FLOATS];
    uint32_t last_sh_degree = 3;
    cudaStream_t last_stream = nullptr;
    FrameStatus* status = nullptr;  // pinned, status_slots entries
    uint32_t status_slots = 0;
    cudaEvent_t frame_done = nullptr;

    bool timing = false;
    cudaEvent_t ev[7] = {};
    float stage_ms[TPDCU_NUM_STAGES] = {};

    // CUDA graph of the frame's parameter-invariant middle section (memset .. ranges); re-captured when anything it bakes in changes
    struct GraphSig {
        const void* zero_region; size_t zero_bytes; const void* keys0; const void* keys1; const void* geo; const void* posop;
        uint32_t n, capacity, width, height, sh_degree, packed_idx_bits, packed_word_bits, entity_count;
        bool keep_unsorted;
        bool operator==(const GraphSig& o) const { return memcmp(this, &o, sizeof(GraphSig)) == 0; }
    };
    bool use_graph = true;
    bool graph_valid = false;
    GraphSig graph_sig{};
    cudaGraphExec_t graph_exec = nullptr;
    cudaStream_t capture_stream = nullptr;
    uint32_t graph_launches = 0, graph_captures = 0;

    // standalone sort
    cudaEvent_t sort_ev[2] = {};
    float sort_ms = 0.f;
    uint32_t sort_passes = 0;
};

static uint32_t bit_length(uint32_t v) {
    uint32_t b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}
static uint32_t tiles_of(const tpdcu_ctx* c) {
    return ((c->width + TILE_PX - 1) / TILE_PX) * ((c->height + TILE_PX - 1) / TILE_PX);
}
static uint32_t frame_end_bit(const tpdcu_ctx* c) {
    const uint32_t tiles = tiles_of(c);
    return 32u + (tiles > 1 ? bit_length(tiles - 1) : 0u);
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// Packed sort words hold tile | depth - bias | index (sort.cu). Used when the index and tile bits leave at least 24 bits for
// the depth range of a frame (27 bits cover the default near/far planes); verified per frame on the device.
static uint32_t packed_idx_bits(const tpdcu_ctx* c) {
    if (c->packed_disabled || c->n == 0) return 0;
    const uint32_t idx_bits = std::max(1u, bit_length(c->n - 1));
    const uint32_t tile_bits = frame_end_bit(c) - 32u;
    return (idx_bits + tile_bits <= 40u) ? idx_bits : 0u;
}

static void free_scene(tpdcu_ctx* c) {
    cudaFree(c->posop); cudaFree(c->cov_a); cudaFree(c->cov_b); cudaFree(c->sh); cudaFree(c->entity);
    cudaFree(c->geo); cudaFree(c->color); cudaFree(c->depth_radius); cudaFree(c->offsets); cudaFree(c->models); cudaFree(c->vm); cudaFree(c->pm);
    c->posop = c->cov_a = nullptr; c->cov_b = nullptr; c->sh = nullptr; c->entity = nullptr; c->geo = nullptr; c->color = nullptr; c->depth_radius = nullptr;
    c->offsets = nullptr; c->models = c->vm = c->pm = nullptr;
    c->n = 0; c->entity_count = 0;
}

static void free_pairs(tpdcu_ctx* c) {
    for (int i = 0; i < 2; ++i) { cudaFree(c->keys[i]); cudaFree(c->vals[i]); c->keys[i] = nullptr; c->vals[i] = nullptr; }
    c->capacity = 0;
}

static int ensure_pairs(tpdcu_ctx* c, uint32_t want) {
    if (want <= c->capacity) return TPDCU_OK;
    // grow-only (GaussianEngine.cpp:671-674,793-804), rounded to whole sort tiles
    uint64_t cap64 = align_up(want, SORT_TILE);
    if (cap64 > 0xffffffffull - SORT_TILE) return fail(TPDCU_ERR_INVALID, "pair capacity exceeds 2^32");
    free_pairs(c);
    const uint32_t cap = (uint32_t)cap64;
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&c->keys[i], (size_t)cap * sizeof(uint64_t)));
        CK(cudaMalloc(&c->vals[i], (size_t)cap * sizeof(uint32_t)));
    }
    c->capacity = cap;
    return TPDCU_OK;
}

// The per-frame zeroed region: FrameCtl | scan descriptors | tile ranges | onesweep look-back arrays
static int ensure_zero_region(tpdcu_ctx* c, uint32_t passes) {
    const uint32_t tiles = tiles_of(c);
    if (c->zero_region && c->zero_n == c->n && c->zero_capacity == c->capacity && c->zero_tiles == tiles) return TPDCU_OK;
    cudaFree(c->zero_region);
    c->zero_region = nullptr;
    const uint32_t pre_parts = (c->n + PRE_PART - 1) / PRE_PART;
    size_t off = align_up(sizeof(FrameCtl), 256);
    c->off_scan_desc = off; off = align_up(off + (size_t)pre_parts * sizeof(uint64_t), 256);
    c->off_ranges = off;    off = align_up(off + (size_t)tiles * 2 * sizeof(uint32_t), 256);
    c->off_lookback = off;  off = align_up(off + (size_t)SORT_MAX_PASSES * sort_parts(c->capacity) * SORT_BINS * sizeof(uint32_t), 256);
    (void)passes;
    CK(cudaMalloc(&c->zero_region, off));
    c->zero_bytes = off;
    c->zero_n = c->n; c->zero_capacity = c->capacity; c->zero_tiles = tiles;
    return TPDCU_OK;
}

static int ensure_status(tpdcu_ctx* c, uint32_t slots) {
    if (slots <= c->status_slots) return TPDCU_OK;
    if (c->status) cudaFreeHost(c->status);
    c->status = nullptr;
    CK(cudaMallocHost(&c->status, sizeof(FrameStatus) * slots));
    memset(c->status, 0, sizeof(FrameStatus) * slots);
    c->status_slots = slots;
    return TPDCU_OK;
}

static int ensure_target(tpdcu_ctx* c) {
    const size_t need = (size_t)c->width * c->height * 4;
    if (need <= c->target_bytes && c->target) return TPDCU_OK;
    cudaFree(c->target);
    c->target = nullptr;
    CK(cudaMalloc(&c->target, std::max<size_t>(need, 4)));
    c->target_bytes = need;
    return TPDCU_OK;
}

struct FrameLaunch {
    PreprocessLaunch pre;
    SortLaunch sort;
    RasterLaunch raster;
};

// The part of a frame whose launch parameters do not change from frame to frame: everything between the camera setup and
// the blend. Either enqueued directly or captured once into a CUDA graph and replayed (13 launches -> 1 graph launch).
static int enqueue_middle(tpdcu_ctx* c, const FrameLaunch& f, cudaStream_t s, bool timing) {
    CK(cudaMemsetAsync(c->zero_region, 0, c->zero_bytes, s));
    if (timing) CK(cudaEventRecord(c->ev[1], s));
    CK(launch_preprocess(f.pre, s));
    CK(launch_color(f.pre, s));
    if (c->keep_unsorted && c->capacity) {
        CK(cudaMemcpyAsync(c->unsorted_keys, c->keys[0], (size_t)c->capacity * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(c->unsorted_vals, c->vals[0], (size_t)c->capacity * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (timing) CK(cudaEventRecord(c->ev[2], s));
    CK(launch_sort(f.sort, UINT32_MAX, s, timing ? c->ev[3] : nullptr));
    if (timing) CK(cudaEventRecord(c->ev[4], s));
    CK(launch_ranges(f.raster, s));
    if (timing) CK(cudaEventRecord(c->ev[5], s));
    return TPDCU_OK;
}

static void drop_graph(tpdcu_ctx* c) {
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    c->graph_exec = nullptr;
    c->graph_valid = false;
}

// Enqueue one frame. No host synchronisation.
static int enqueue_frame(tpdcu_ctx* c, const float* ubo, uint32_t sh_degree, cudaStream_t s, uint8_t* out, size_t pitch,
                         uint32_t slot) {
    const uint32_t end_bit = frame_end_bit(c);
    const uint32_t passes = (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    if (int r = ensure_zero_region(c, passes)) return r;
    if (c->models_dirty) {
        CK(cudaMemcpyAsync(c->models, c->models_host.data(), sizeof(float) * 16 * c->entity_count, cudaMemcpyHostToDevice, s));
        c->models_dirty = false;
    }
    if (c->keep_unsorted && c->capacity && c->unsorted_capacity < c->capacity) {
        cudaFree(c->unsorted_keys); cudaFree(c->unsorted_vals);
        c->unsorted_keys = nullptr; c->unsorted_vals = nullptr; c->unsorted_capacity = 0;
        CK(cudaMalloc(&c->unsorted_keys, (size_t)c->capacity * 8));
        CK(cudaMalloc(&c->unsorted_vals, (size_t)c->capacity * 4));
        c->unsorted_capacity = c->capacity;
    }
    const bool t = c->timing;

    FrameCtl* ctl = reinterpret_cast<FrameCtl*>(c->zero_region);
    FrameLaunch f{};
    PreprocessLaunch& p = f.pre;
    p.scene = SceneArrays{ c->posop, c->cov_a, c->cov_b, c->sh, c->entity_count > 1 ? c->entity : nullptr, c->n, c->entity_count };
    p.models = c->models; p.cam = c->cam; p.vm = c->vm; p.pm = c->pm;
    p.ctl = ctl;
    p.scan_desc = reinterpret_cast<uint64_t*>(c->zero_region + c->off_scan_desc);
    p.out = SplatArrays{ c->geo, c->color, c->depth_radius, c->offsets };
    p.keys = c->keys[0]; p.vals = c->vals[0];
    p.capacity = c->capacity;
    p.width = c->width; p.height = c->height; p.sh_degree = std::min(sh_degree, 3u);  // GaussianEngine.cpp:366-370

    SortLaunch& so = f.sort;
    so.keys[0] = c->keys[0]; so.keys[1] = c->keys[1]; so.vals[0] = c->vals[0]; so.vals[1] = c->vals[1];
    so.ctl = ctl; so.plan = c->plan;
    so.lookback = reinterpret_cast<uint32_t*>(c->zero_region + c->off_lookback);
    so.capacity = c->capacity; so.end_bit = end_bit; so.sm_count = c->sm_count;
    so.packed_idx_bits = packed_idx_bits(c);
    so.packed_word_bits = c->packed_word_bits;

    RasterLaunch& ra = f.raster;
    ra.keys[0] = c->keys[0]; ra.keys[1] = c->keys[1]; ra.vals[0] = c->vals[0]; ra.vals[1] = c->vals[1];
    ra.plan = c->plan; ra.geo = c->geo; ra.color = c->color;
    ra.ranges = reinterpret_cast<uint32_t*>(c->zero_region + c->off_ranges);
    ra.out = out; ra.pitch = pitch; ra.capacity = c->capacity; ra.width = c->width; ra.height = c->height;

    if (t) CK(cudaEventRecord(c->ev[0], s));
    CameraUbo cu;
    memcpy(cu.f, ubo, sizeof(cu.f));  // by-value kernel argument: no per-frame H2D copy (updateCameraBuffer, :764-775)
    CK(launch_setup(p, cu, s));

    bool replayed = false;
    if (c->use_graph && !t) {
        tpdcu_ctx::GraphSig sig;
        memset(&sig, 0, sizeof(sig));
        sig.zero_region = c->zero_region; sig.zero_bytes = c->zero_bytes; sig.keys0 = c->keys[0]; sig.keys1 = c->keys[1];
        sig.geo = c->geo; sig.posop = c->posop; sig.n = c->n; sig.capacity = c->capacity; sig.width = c->width; sig.height = c->height;
        sig.sh_degree = p.sh_degree; sig.packed_idx_bits = so.packed_idx_bits; sig.packed_word_bits = so.packed_word_bits;
        sig.entity_count = c->entity_count; sig.keep_unsorted = c->keep_unsorted;
        if (!(c->graph_valid && c->graph_sig == sig)) {
            // (re)capture on a private stream: nothing executes during capture, and the legacy default stream cannot capture
            drop_graph(c);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(c->capture_stream, cudaStreamCaptureModeRelaxed);
            int rc = TPDCU_OK;
            if (e == cudaSuccess) {
                rc = enqueue_middle(c, f, c->capture_stream, false);
                e = cudaStreamEndCapture(c->capture_stream, &graph);
            }
            if (e == cudaSuccess && rc == TPDCU_OK && graph) e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e == cudaSuccess && rc == TPDCU_OK && c->graph_exec) {
                c->graph_valid = true;
                c->graph_sig = sig;
                ++c->graph_captures;
            } else {
                cudaGetLastError();
                drop_graph(c);
                c->use_graph = false;  // capture is unavailable in this process: keep launching directly
            }
        }
        if (c->graph_valid) {
            CK(cudaGraphLaunch(c->graph_exec, s));
            ++c->graph_launches;
            replayed = true;
        }
    }
    if (!replayed)
        if (int r = enqueue_middle(c, f, s, t)) return r;

    CK(launch_blend(ra, s));
    if (t) CK(cudaEventRecord(c->ev[6], s));

    CK(cudaMemcpyAsync(&c->status[slot].pairs_total, &ctl->pairs_total, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&c->status[slot].n, c->plan, 10 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    return TPDCU_OK;
}

// Does a finished frame have to be rendered again (buffers too small / packed words too narrow)? Adjusts the context.
static int frame_needs_rerender(tpdcu_ctx* c, const FrameStatus& st, bool* again) {
    *again = false;
    if (st.pairs_total > c->capacity) {
        *again = true;
        return TPDCU_OK;
    }
    if (st.packed_overflow) {
        c->packed_disabled = true;
        *again = true;
    }
    return TPDCU_OK;
}

static uint32_t grown_capacity(uint32_t pairs) {
    const uint64_t want = (uint64_t)pairs + pairs / 8 + SORT_TILE;  // 12.5 % head-room against view changes
    return (uint32_t)std::min<uint64_t>(want, 0xffffffffull - 2 * SORT_TILE);
}

static uint8_t* out_ptr(tpdcu_ctx* c, size_t* pitch) {
    if (c->bound_out) { *pitch = c->bound_pitch; return c->bound_out; }
    *pitch = (size_t)c->width * 4;
    return c->target;
}

static int check_ready(tpdcu_ctx* c) {
    if (!c) return fail(TPDCU_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    return TPDCU_OK;
}

static int finish_internal(tpdcu_ctx* c) {
    if (!c->frame_valid && !c->frame_pending) return fail(TPDCU_ERR_STATE, "no frame has been rendered");
    for (int attempt = 0; c->frame_pending; ++attempt) {
        CK(cudaEventSynchronize(c->frame_done));
        const uint32_t pairs = c->status[0].pairs_total;
        bool again = false;
        if (int r = frame_needs_rerender(c, c->status[0], &again)) return r;
        if (!again) { c->frame_pending = false; c->frame_valid = true; break; }
        if (attempt >= 4) return fail(TPDCU_ERR_STATE, "pair buffer kept overflowing");
        CK(cudaStreamSynchronize(c->last_stream));
        if (pairs > c->capacity)
            if (int r = ensure_pairs(c, grown_capacity(pairs))) return r;
        size_t pitch; uint8_t* out = out_ptr(c, &pitch);
        if (int r = enqueue_frame(c, c->last_ubo, c->last_sh_degree, c->last_stream, out, pitch, 0)) return r;
        CK(cudaEventRecord(c->frame_done, c->last_stream));
    }
    if (c->timing) {
        CK(cudaEventSynchronize(c->ev[6]));
        float ms;
        for (int k = 0; k < 6; ++k) { CK(cudaEventElapsedTime(&ms, c->ev[k], c->ev[k + 1])); c->stage_ms[k] = ms; }
        // ev: 0 start | 1 after setup | 2 after preprocess | 3 after hist+plan | 4 after passes | 5 after ranges | 6 after blend
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[6]));
        c->stage_ms[6] = ms;
        c->stage_ms[7] = (float)c->status[0].passes_run;
    }
    return TPDCU_OK;
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
    CKB(cudaMalloc(&c->cam, sizeof(FrameCam)));
    CKB(cudaMalloc(&c->plan, sizeof(SortPlan)));
    CKB(cudaMemset(c->plan, 0, sizeof(SortPlan)));
    CKB(cudaEventCreateWithFlags(&c->frame_done, cudaEventDisableTiming));
    CKB(cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev) CKB(cudaEventCreate(&e));
    for (auto& e : c->sort_ev) CKB(cudaEventCreate(&e));
#undef CKB
    memset(&c->graph_sig, 0, sizeof(c->graph_sig));
    if (cudaError_t e = init_sort_attributes()) return bail(fail(TPDCU_ERR_CUDA, std::string("init_sort_attributes: ") + cudaGetErrorString(e)));
    if (int r = ensure_status(c, 1)) return bail(r);
    *out = c;
    return TPDCU_OK;
}

void tpdcu_destroy(tpdcu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_scene(c);
    free_pairs(c);
    cudaFree(c->cam); cudaFree(c->plan); cudaFree(c->zero_region); cudaFree(c->target);
    cudaFree(c->unsorted_keys); cudaFree(c->unsorted_vals);
    drop_graph(c);
    if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
    if (c->ext_mem) cudaDestroyExternalMemory(c->ext_mem);
    if (c->status) cudaFreeHost(c->status);
    if (c->frame_done) cudaEventDestroy(c->frame_done);
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
    // scene arrays
    free_scene(c);
    c->frame_valid = false;
    CK(cudaMalloc(&c->posop, (size_t)n * sizeof(float4)));
    CK(cudaMalloc(&c->cov_a, (size_t)n * sizeof(float4)));
    CK(cudaMalloc(&c->cov_b, (size_t)n * sizeof(float2)));
    CK(cudaMalloc(&c->sh, (size_t)n * SH_PLANES * sizeof(float4)));
    CK(cudaMalloc(&c->geo, (size_t)n * sizeof(SplatGeo)));
    CK(cudaMalloc(&c->color, (size_t)n * sizeof(float4)));
    CK(cudaMalloc(&c->depth_radius, (size_t)n * sizeof(float2)));
    CK(cudaMalloc(&c->offsets, ((size_t)n + 1) * sizeof(uint32_t)));
    CK(cudaMalloc(&c->models, (size_t)entity_count * 16 * sizeof(float)));
    CK(cudaMalloc(&c->vm, (size_t)entity_count * 16 * sizeof(float)));
    CK(cudaMalloc(&c->pm, (size_t)entity_count * 16 * sizeof(float)));
    if (entity_count > 1) {
        CK(cudaMalloc(&c->entity, (size_t)n * sizeof(uint32_t)));
        if (d_entity) CK(cudaMemcpyAsync(c->entity, d_entity, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        else CK(cudaMemsetAsync(c->entity, 0, (size_t)n * sizeof(uint32_t), s));
    }
    c->n = n;
    c->entity_count = entity_count;
    c->models_host.assign((size_t)entity_count * 16, 0.0f);  // identity (createBindlessTransformBuffer)
    for (uint32_t e = 0; e < entity_count; ++e)
        for (int k = 0; k < 4; ++k) c->models_host[(size_t)e * 16 + k * 5] = 1.0f;
    c->models_dirty = true;
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
    return upload_common(c, d_recs240, n, d_entity_idx, entity_count, (cudaStream_t)stream);
}

int tpdcu_set_transform(tpdcu_ctx* c, uint32_t entity, const float m[16]) {
    if (int r = check_ready(c)) return r;
    if (c->n == 0) return fail(TPDCU_ERR_STATE, "no scene compiled");
    if (entity >= c->entity_count || !m) return fail(TPDCU_ERR_INVALID, "bad entity or matrix");
    memcpy(&c->models_host[(size_t)entity * 16], m, sizeof(float) * 16);
    c->models_dirty = true;
    return TPDCU_OK;
}

int tpdcu_resize(tpdcu_ctx* c, uint32_t width, uint32_t height) {
    if (int r = check_ready(c)) return r;
    if (width == 0 || height == 0 || width > 65535u * TILE_PX || height > 65535u * TILE_PX)
        return fail(TPDCU_ERR_INVALID, "bad framebuffer size");
    if (c->frame_pending) CK(cudaStreamSynchronize(c->last_stream));
    c->frame_pending = false;
    c->frame_valid = false;
    c->width = width;
    c->height = height;
    return ensure_target(c);
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
    CK(cudaImportExternalMemory(&mem, &hd));
    cudaExternalMemoryBufferDesc bd{};
    bd.offset = 0;
    bd.size = bytes;
    void* ptr = nullptr;
    cudaError_t e = cudaExternalMemoryGetMappedBuffer(&ptr, mem, &bd);
    if (e != cudaSuccess) {
        cudaDestroyExternalMemory(mem);
        return fail(TPDCU_ERR_CUDA, std::string("cudaExternalMemoryGetMappedBuffer: ") + cudaGetErrorString(e));
    }
    drop_graph(c);
    if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
    if (c->ext_mem) cudaDestroyExternalMemory(c->ext_mem);
    c->ext_mem = mem;
    c->bound_out = reinterpret_cast<uint8_t*>(ptr);
    c->bound_pitch = (size_t)c->width * 4;
    return TPDCU_OK;
}

int tpdcu_raster(tpdcu_ctx* c, const float camera_ubo[TPDCU_CAMERA_FLOATS], uint32_t sh_degree, void* stream) {
    if (int r = check_ready(c)) return r;
    if (!camera_ubo) return fail(TPDCU_ERR_INVALID, "camera_ubo is null");
    if (c->n == 0) return fail(TPDCU_ERR_STATE, "no scene compiled");
    if (c->width == 0) return fail(TPDCU_ERR_STATE, "tpdcu_resize has not been called");
    cudaStream_t s = (cudaStream_t)stream;
    // A still-pending frame that overflowed is simply superseded by this one; its P still tells us how
    // far to grow before we start, if it has already landed.
    if (c->frame_pending && cudaEventQuery(c->frame_done) == cudaSuccess && c->status[0].pairs_total > c->capacity) {
        CK(cudaStreamSynchronize(c->last_stream));
        if (int r = ensure_pairs(c, grown_capacity(c->status[0].pairs_total))) return r;
    }
    cudaGetLastError();
    if (c->capacity == 0)
        if (int r = ensure_pairs(c, SORT_TILE)) return r;
    size_t pitch; uint8_t* out = out_ptr(c, &pitch);
    memcpy(c->last_ubo, camera_ubo, sizeof(c->last_ubo));
    c->last_sh_degree = sh_degree;
    c->last_stream = s;
    if (int r = enqueue_frame(c, camera_ubo, sh_degree, s, out, pitch, 0)) return r;
    CK(cudaEventRecord(c->frame_done, s));
    c->frame_pending = true;
    return TPDCU_OK;
}

int tpdcu_raster_views(tpdcu_ctx* c, const float* camera_ubos, uint32_t n_views, uint32_t sh_degree, void* d_frames,
                       size_t frame_stride_bytes, void* stream) {
    if (int r = check_ready(c)) return r;
    if (!camera_ubos || !d_frames || n_views == 0) return fail(TPDCU_ERR_INVALID, "bad batch arguments");
    if (c->n == 0) return fail(TPDCU_ERR_STATE, "no scene compiled");
    if (c->width == 0) return fail(TPDCU_ERR_STATE, "tpdcu_resize has not been called");
    const size_t pitch = (size_t)c->width * 4;
    if (frame_stride_bytes < pitch * c->height) return fail(TPDCU_ERR_INVALID, "frame stride smaller than a frame");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->frame_pending) CK(cudaStreamSynchronize(c->last_stream));
    c->frame_pending = false;
    if (int r = ensure_status(c, n_views)) return r;
    if (c->capacity == 0)
        if (int r = ensure_pairs(c, SORT_TILE)) return r;
    std::vector<uint32_t> todo(n_views);
    for (uint32_t v = 0; v < n_views; ++v) todo[v] = v;
    const bool timing = c->timing;
    c->timing = false;  // per-stage events describe single frames only
    int rc = TPDCU_OK;
    for (int attempt = 0; !todo.empty() && rc == TPDCU_OK; ++attempt) {
        if (attempt >= 5) { rc = fail(TPDCU_ERR_STATE, "pair buffer kept overflowing"); break; }
        for (uint32_t v : todo) {
            rc = enqueue_frame(c, camera_ubos + (size_t)v * TPDCU_CAMERA_FLOATS, sh_degree, s,
                               reinterpret_cast<uint8_t*>(d_frames) + (size_t)v * frame_stride_bytes, pitch, v);
            if (rc != TPDCU_OK) break;
        }
        if (rc != TPDCU_OK) break;
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { rc = fail(TPDCU_ERR_CUDA, std::string("batch sync: ") + cudaGetErrorString(e)); break; }
        std::vector<uint32_t> again;
        uint32_t max_pairs = 0;
        for (uint32_t v : todo) {
            bool redo = false;
            frame_needs_rerender(c, c->status[v], &redo);
            if (redo) { again.push_back(v); max_pairs = std::max(max_pairs, c->status[v].pairs_total); }
        }
        if (!again.empty() && max_pairs > c->capacity) rc = ensure_pairs(c, grown_capacity(max_pairs));
        todo.swap(again);
    }
    c->timing = timing;
    if (rc != TPDCU_OK) return rc;
    // the last view rendered is what introspection sees
    c->status[0] = c->status[n_views - 1];
    memcpy(c->last_ubo, camera_ubos + (size_t)(n_views - 1) * TPDCU_CAMERA_FLOATS, sizeof(c->last_ubo));
    c->last_sh_degree = sh_degree;
    c->last_stream = s;
    c->frame_valid = true;
    return TPDCU_OK;
}

int tpdcu_finish(tpdcu_ctx* c, uint32_t* pairs) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    if (pairs) *pairs = c->status[0].pairs_total;
    return TPDCU_OK;
}

int tpdcu_read_frame(tpdcu_ctx* c, void* host_rgba8, size_t host_pitch_bytes) {
    if (int r = check_ready(c)) return r;
    if (!host_rgba8 || host_pitch_bytes < (size_t)c->width * 4) return fail(TPDCU_ERR_INVALID, "bad host buffer");
    if (int r = finish_internal(c)) return r;
    size_t pitch; uint8_t* out = out_ptr(c, &pitch);
    CK(cudaMemcpy2DAsync(host_rgba8, host_pitch_bytes, out, pitch, (size_t)c->width * 4, c->height, cudaMemcpyDeviceToHost, c->last_stream));
    CK(cudaStreamSynchronize(c->last_stream));
    return TPDCU_OK;
}

int tpdcu_get_counts(tpdcu_ctx* c, uint32_t* pairs, uint32_t* visible) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    if (pairs) *pairs = c->status[0].pairs_total;
    if (visible) *visible = c->status[0].visible;
    return TPDCU_OK;
}

int tpdcu_read_splats(tpdcu_ctx* c, void* host_splats48, uint32_t n) {
    if (int r = check_ready(c)) return r;
    if (!host_splats48 || n > c->n) return fail(TPDCU_ERR_INVALID, "bad splat read");
    if (int r = finish_internal(c)) return r;
    if (n == 0) return TPDCU_OK;
    void* tmp = nullptr;
    CK(cudaMalloc(&tmp, (size_t)n * TPDCU_SPLAT_BYTES));
    cudaError_t e = launch_export_splats(SplatArrays{ c->geo, c->color, c->depth_radius, c->offsets }, n, tmp, c->last_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_splats48, tmp, (size_t)n * TPDCU_SPLAT_BYTES, cudaMemcpyDeviceToHost, c->last_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->last_stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(TPDCU_ERR_CUDA, std::string("read_splats: ") + cudaGetErrorString(e));
    return TPDCU_OK;
}

static int read_sorted(tpdcu_ctx* c, void* host, uint32_t count, bool want_keys) {
    if (int r = check_ready(c)) return r;
    if (!host) return fail(TPDCU_ERR_INVALID, "host buffer is null");
    if (int r = finish_internal(c)) return r;
    if (count > c->status[0].pairs_total) return fail(TPDCU_ERR_INVALID, "count exceeds the frame's pair count");
    if (count == 0) return TPDCU_OK;
    const FrameStatus& st = c->status[0];
    const void* src;
    if (st.packed) {
        // the frame's result is one array of packed words: expand it into the reference's (key, value) arrays in the
        // buffers the sort no longer needs
        SortLaunch so{};
        so.keys[0] = c->keys[0]; so.keys[1] = c->keys[1]; so.plan = c->plan; so.sm_count = c->sm_count;
        uint64_t* uk = c->keys[(st.final_sel & 1u) ^ 1u];
        uint32_t* uv = c->vals[1];
        CK(launch_sort_unpack(so, uk, uv, c->last_stream));
        src = want_keys ? (const void*)uk : (const void*)uv;
    } else {
        src = want_keys ? (const void*)c->keys[st.final_sel & 1u] : (const void*)c->vals[st.final_sel & 1u];
    }
    CK(cudaMemcpyAsync(host, src, (size_t)count * (want_keys ? 8 : 4), cudaMemcpyDeviceToHost, c->last_stream));
    CK(cudaStreamSynchronize(c->last_stream));
    return TPDCU_OK;
}

int tpdcu_read_keys(tpdcu_ctx* c, uint64_t* host_keys, uint32_t count) { return read_sorted(c, host_keys, count, true); }
int tpdcu_read_values(tpdcu_ctx* c, uint32_t* host_vals, uint32_t count) { return read_sorted(c, host_vals, count, false); }

int tpdcu_read_ranges(tpdcu_ctx* c, uint32_t* host_ranges2, uint32_t tile_count) {
    if (int r = check_ready(c)) return r;
    if (!host_ranges2 || tile_count > tiles_of(c)) return fail(TPDCU_ERR_INVALID, "bad range read");
    if (int r = finish_internal(c)) return r;
    CK(cudaMemcpyAsync(host_ranges2, c->zero_region + c->off_ranges, (size_t)tile_count * 8, cudaMemcpyDeviceToHost, c->last_stream));
    CK(cudaStreamSynchronize(c->last_stream));
    return TPDCU_OK;
}

int tpdcu_keep_unsorted(tpdcu_ctx* c, int enable) {
    if (int r = check_ready(c)) return r;
    c->keep_unsorted = enable != 0;
    return TPDCU_OK;
}

int tpdcu_read_unsorted(tpdcu_ctx* c, uint64_t* host_keys, uint32_t* host_vals, uint32_t count) {
    if (int r = check_ready(c)) return r;
    if (!c->keep_unsorted || !c->unsorted_keys) return fail(TPDCU_ERR_STATE, "tpdcu_keep_unsorted was not enabled for the last frame");
    if (int r = finish_internal(c)) return r;
    if (count > c->status[0].pairs_total || count > c->unsorted_capacity) return fail(TPDCU_ERR_INVALID, "count exceeds the frame's pair count");
    if (host_keys) CK(cudaMemcpyAsync(host_keys, c->unsorted_keys, (size_t)count * 8, cudaMemcpyDeviceToHost, c->last_stream));
    if (host_vals) CK(cudaMemcpyAsync(host_vals, c->unsorted_vals, (size_t)count * 4, cudaMemcpyDeviceToHost, c->last_stream));
    CK(cudaStreamSynchronize(c->last_stream));
    return TPDCU_OK;
}

int tpdcu_enable_stage_timing(tpdcu_ctx* c, int enable) {
    if (int r = check_ready(c)) return r;
    c->timing = enable != 0;
    return TPDCU_OK;
}

int tpdcu_stage_times_ms(tpdcu_ctx* c, float times_ms[TPDCU_NUM_STAGES]) {
    if (int r = check_ready(c)) return r;
    if (!times_ms) return fail(TPDCU_ERR_INVALID, "times_ms is null");
    if (!c->timing) return fail(TPDCU_ERR_STATE, "stage timing is not enabled");
    if (int r = finish_internal(c)) return r;
    // stage_ms[k] = ev[k+1]-ev[k]: 0 clear+setup | 1 preprocess | 2 hist+plan | 3 passes | 4 ranges | 5 blend
    memcpy(times_ms, c->stage_ms, sizeof(float) * TPDCU_NUM_STAGES);
    return TPDCU_OK;
}

int tpdcu_get_sort_info(tpdcu_ctx* c, uint32_t* packed, uint32_t* depth_bits, uint32_t* idx_bits, uint32_t* total_bits) {
    if (int r = check_ready(c)) return r;
    if (int r = finish_internal(c)) return r;
    const FrameStatus& st = c->status[0];
    if (packed) *packed = st.packed;
    if (depth_bits) *depth_bits = st.depth_bits;
    if (idx_bits) *idx_bits = st.idx_bits;
    if (total_bits) *total_bits = st.total_bits;
    return TPDCU_OK;
}

int tpdcu_set_packed_word_bits(tpdcu_ctx* c, uint32_t bits) {
    if (int r = check_ready(c)) return r;
    if (bits < 1 || bits > 64) return fail(TPDCU_ERR_INVALID, "bits must be in [1, 64]");
    c->packed_word_bits = bits;
    return TPDCU_OK;
}

int tpdcu_set_graph_replay(tpdcu_ctx* c, int enable, uint32_t* captures, uint32_t* launches) {
    if (int r = check_ready(c)) return r;
    if (enable >= 0) {
        c->use_graph = enable != 0;
        if (!c->use_graph) {
            if (c->frame_pending) CK(cudaStreamSynchronize(c->last_stream));
            drop_graph(c);
        }
    }
    if (captures) *captures = c->graph_captures;
    if (launches) *launches = c->graph_launches;
    return TPDCU_OK;
}

int tpdcu_get_capacity(tpdcu_ctx* c, uint32_t* capacity_pairs) {
    if (int r = check_ready(c)) return r;
    if (capacity_pairs) *capacity_pairs = c->capacity;
    return TPDCU_OK;
}

int tpdcu_reserve_pairs(tpdcu_ctx* c, uint32_t capacity_pairs) {
    if (int r = check_ready(c)) return r;
    if (c->frame_pending) { if (int r = finish_internal(c)) return r; }
    CK(cudaDeviceSynchronize());
    return ensure_pairs(c, capacity_pairs);
}

int tpdcu_sort_pairs_device(tpdcu_ctx* c, uint64_t* d_keys, uint32_t* d_vals, uint32_t n, uint32_t end_bit, void* stream) {
    if (int r = check_ready(c)) return r;
    if ((!d_keys || !d_vals) && n) return fail(TPDCU_ERR_INVALID, "null device buffers");
    if (end_bit > 64) return fail(TPDCU_ERR_INVALID, "end_bit > 64");
    if (n >= (1u << 30)) return fail(TPDCU_ERR_INVALID, "n must be below 2^30");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->frame_pending) { if (int r = finish_internal(c)) return r; }
    c->frame_valid = false;  // the pair buffers are about to be reused
    if (n > c->capacity) {
        CK(cudaDeviceSynchronize());
        if (int r = ensure_pairs(c, n)) return r;
    }
    if (c->capacity == 0)
        if (int r = ensure_pairs(c, SORT_TILE)) return r;
    if (int r = ensure_zero_region(c, SORT_MAX_PASSES)) return r;
    CK(cudaMemsetAsync(c->zero_region, 0, c->zero_bytes, s));
    if (n) {
        CK(cudaMemcpyAsync(c->keys[0], d_keys, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(c->vals[0], d_vals, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    }
    SortLaunch so{};
    so.keys[0] = c->keys[0]; so.keys[1] = c->keys[1]; so.vals[0] = c->vals[0]; so.vals[1] = c->vals[1];
    so.ctl = reinterpret_cast<FrameCtl*>(c->zero_region); so.plan = c->plan;
    so.lookback = reinterpret_cast<uint32_t*>(c->zero_region + c->off_lookback);
    so.capacity = c->capacity; so.end_bit = end_bit; so.sm_count = c->sm_count;
    CK(cudaEventRecord(c->sort_ev[0], s));
    CK(launch_sort(so, n, s, nullptr));
    CK(cudaEventRecord(c->sort_ev[1], s));
    CK(launch_sort_copy_result(so, d_keys, d_vals, n, s));
    CK(cudaMemcpyAsync(&c->status[0].n, c->plan, 10 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    c->last_stream = s;
    return TPDCU_OK;
}

int tpdcu_sort_last_ms(tpdcu_ctx* c, float* ms, uint32_t* passes_run) {
    if (int r = check_ready(c)) return r;
    CK(cudaEventSynchronize(c->sort_ev[1]));
    CK(cudaStreamSynchronize(c->last_stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, c->sort_ev[0], c->sort_ev[1]));
    if (ms) *ms = t;
    if (passes_run) *passes_run = c->status[0].passes_run;
    return TPDCU_OK;
}

}  // extern "C"

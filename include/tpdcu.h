/* tpdcu.h — C ABI of libtpdcu.so: the B200 (sm_100a) CUDA implementation of torpedo's
 * Gaussian-splatting forward rasterizer.
 *
 * The reference (ndming/torpedo) has no FFI for this path: the boundary is the public C++ surface of
 * tpd::GaussianEngine (torpedo/volumetric/include/torpedo/volumetric/GaussianEngine.h:15-29) whose
 * implementation records nine Slang compute shaders into Vulkan command buffers
 * (torpedo/volumetric/src/GaussianEngine.cpp:621-712,777-863). Each entry point below replaces one
 * piece of that implementation; the reference line it stands in for is cited on the declaration.
 * The Vulkan-free C++ drop-in that calls this ABI is include/torpedo_b200/GaussianEngine.hpp;
 * INTEGRATION.md shows the binding a torpedo maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns TPDCU_OK (0) or a negative
 * error code and never throws; tpdcu_last_error() returns a thread-local message; one context per
 * GPU; a context is NOT thread-safe (the reference API is main-thread only). `stream` arguments are
 * CUstream / cudaStream_t handles passed as void* (NULL = the legacy default stream).
 * There is no CPU fallback: every entry point fails with TPDCU_ERR_CUDA when no sm_100 device exists.
 */
#ifndef TPDCU_H
#define TPDCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPDCU_OK 0
#define TPDCU_ERR_INVALID (-1) /* bad argument / call order */
#define TPDCU_ERR_CUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define TPDCU_ERR_OOM (-3)     /* device allocation failed */
#define TPDCU_ERR_STATE (-4)   /* nothing compiled / nothing rendered yet */

#define TPDCU_GAUSSIAN_BYTES 240 /* tpd::GaussianPoint, GaussianGeometry.h:10-18 == splat.slang:24-31 */
#define TPDCU_SPLAT_BYTES 48     /* splat.slang:33-39, GaussianEngine.h:124 */
#define TPDCU_CAMERA_FLOATS 34   /* splat.slang:18-22: view(16) | proj*view(16) | focalNDC(2) */
#define TPDCU_TILE 16            /* BLOCK_X == BLOCK_Y, GaussianEngine.h:122-123 */
#define TPDCU_NUM_STAGES 11

typedef struct tpdcu_ctx tpdcu_ctx;

/* ---- lifetime -------------------------------------------------------------------------------- */

/* Engine::init + GaussianEngine::onInitialized (rendering/src/Engine.cpp:7-40,
 * volumetric/src/GaussianEngine.cpp:75-170): bind to CUDA device `device`, create internal state. */
int tpdcu_create(int device, tpdcu_ctx** out);
/* GaussianEngine::destroy (GaussianEngine.cpp:877-953) */
void tpdcu_destroy(tpdcu_ctx* ctx);
const char* tpdcu_last_error(void);
/* "sm_100a / <device name> / <SM count>" of the bound device; for logs and bench JSON. */
int tpdcu_device_info(tpdcu_ctx* ctx, char* buf, size_t buf_bytes, int* sm_count);

/* ---- scene upload: GaussianEngine::compile (GaussianEngine.cpp:359-399) ----------------------- */

/* Snapshot `n` 240-byte GaussianPoint records (host memory, borrowed for the call) plus the per-Gaussian
 * entity (transform) index (`entity_idx` may be NULL => all 0) — createGaussianBuffer / createTransformIndexBuffer
 * (:401-410, :458-469). Model matrices start as identity (createBindlessTransformBuffer, :471-486).
 * The records are re-laid out on the device (see DESIGN.md "data layout"). n == 0 is a no-op that
 * leaves the previous scene in place, like the reference's warn-and-return (:362-365). */
int tpdcu_upload_gaussians(tpdcu_ctx* ctx, const void* recs240, uint32_t n, const uint32_t* entity_idx,
                           uint32_t entity_count);
/* Same, with `d_recs240` / `d_entity_idx` already in device memory on this GPU (used after the NCCL
 * broadcast of a replicated scene). Runs on `stream`; the inputs may be freed once the
 * stream has passed this point. */
int tpdcu_upload_gaussians_device(tpdcu_ctx* ctx, const void* d_recs240, uint32_t n, const uint32_t* d_entity_idx,
                                  uint32_t entity_count, void* stream);
/* TransformHost::transform (rendering/src/TransformHost.cpp:3-11): row-major 4x4 model matrix of one
 * entity, effective from the next raster call. */
int tpdcu_set_transform(tpdcu_ctx* ctx, uint32_t entity, const float m[16]);

/* ---- render target: createRenderTargets / onFramebufferResize (GaussianEngine.cpp:206-220,315-333) */

int tpdcu_resize(tpdcu_ctx* ctx, uint32_t width, uint32_t height);
/* Render into caller-owned device memory (R8G8B8A8_UNORM, rows `pitch_bytes` apart, pitch >= 4*width).
 * NULL restores the internal targets (one per frame in flight). */
int tpdcu_bind_output_device_ptr(tpdcu_ctx* ctx, void* d_rgba8, size_t pitch_bytes);
/* Render into a Vulkan allocation exported with VK_KHR_external_memory_fd (opaque fd): imported with
 * cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer; replaces Target + recordTargetCopy's
 * source (GaussianEngine.cpp:865-875). The fd is consumed on success. Linear RGBA8, tightly packed. */
int tpdcu_bind_output_fd(tpdcu_ctx* ctx, int fd, size_t bytes);

/* ---- frames of several GPUs collected in one GPU's memory (no reference counterpart: the reference is
 * one GPU, one target, GaussianEngine.cpp:315-333) -------------------------------------------------------------------
 * Views of a batch are sharded over one process per GPU; instead of rendering locally and gathering afterwards, every
 * process binds (tpdcu_bind_output_device_ptr) its views' slots of ONE frame array that lives in the collecting GPU's
 * HBM and is mapped into all of them with CUDA IPC: the blend kernel's pixel stores travel over NVLink / NVSwitch as
 * they are produced and no copy or collective follows the frame. The creator owns the allocation; the other processes
 * of the box open the 64-byte handle (peer access is enabled on first use). A process cannot open its own handle.
 * Completion across processes is the callers' business (a stream-ordered collective or a host barrier after
 * tpdcu_finish). */
#define TPDCU_IPC_HANDLE_BYTES 64
int tpdcu_ipc_frames_create(int device, size_t bytes, void** d_frames, unsigned char handle[TPDCU_IPC_HANDLE_BYTES]);
int tpdcu_ipc_frames_open(int device, const unsigned char handle[TPDCU_IPC_HANDLE_BYTES], void** d_frames);
int tpdcu_ipc_frames_close(int device, void* d_frames);   /* a mapping made by tpdcu_ipc_frames_open */
int tpdcu_ipc_frames_destroy(int device, void* d_frames); /* an allocation made by tpdcu_ipc_frames_create */

/* ---- per-frame hot path: GaussianEngine::rasterFrame (GaussianEngine.cpp:621-712) -------------- */

/* One frame: camera upload (updateCameraBuffer :764-775) -> geometry + scan (project.slang, prefix.slang)
 * -> depth sort of the visible Gaussians -> duplication in depth order (keygen.slang) -> sort of the pairs
 * by tile (together: radix-*.slang x23) -> tile ranges (range.slang) -> blending (blend.slang), which also
 * evaluates the SH colour (splat/common.slang:35-80) of the splats it stages. `camera_ubo` is the reference's 136-byte Camera block. Enqueued on `stream`
 * without any host synchronisation: unlike the reference (:662-674) the pair count P is never read back
 * mid-frame. If P turns out to exceed the pair-buffer capacity the frame is re-rendered after growing
 * the buffers inside the next tpdcu_finish()/tpdcu_read_*() call. */
int tpdcu_raster(tpdcu_ctx* ctx, const float camera_ubo[TPDCU_CAMERA_FLOATS], uint32_t sh_degree, void* stream);
/* A batch of independent views of the same scene: view v is rendered with
 * camera_ubos[34*v..] into d_frames + v*frame_stride_bytes (tightly packed RGBA8 rows). All views are
 * verified (and re-rendered after buffer growth if needed) before the call returns control of `stream`
 * to the caller; the call synchronises the host with `stream` once per batch. */
int tpdcu_raster_views(tpdcu_ctx* ctx, const float* camera_ubos, uint32_t n_views, uint32_t sh_degree,
                       void* d_frames, size_t frame_stride_bytes, void* stream);
/* Wait for the last frame; grow + re-render if it overflowed. Returns the frame's pair count in *pairs. */
int tpdcu_finish(tpdcu_ctx* ctx, uint32_t* pairs);
/* GaussianEngine::draw's copy (GaussianEngine.cpp:714-762) for callers without Vulkan: copy the
 * finished frame to host memory (rows `host_pitch_bytes` apart). Implies tpdcu_finish. */
int tpdcu_read_frame(tpdcu_ctx* ctx, void* host_rgba8, size_t host_pitch_bytes);

/* Same copy, enqueued on `stream` (the stream the frame was rastered with) without waiting: the caller synchronises with the
 * stream or an event of its own. With several frames in flight this lets frame k+1 start while frame k travels to the host —
 * the role the swap-chain fences play in the reference's loop (SurfaceRenderer.cpp:254-319). P is not looked at here: a frame
 * that overflowed its buffers is only repeated by the next tpdcu_finish; tpdcu_frames_repeated tells whether any was.
 * The destination may also be device memory, e.g. a slot of another GPU's frame array (tpdcu_ipc_frames_open): the copy
 * engine then pushes the frame over NVLink while the next one renders. */
int tpdcu_read_frame_async(tpdcu_ctx* ctx, void* host_rgba8, size_t host_pitch_bytes, void* stream);
/* Checks everything enqueued so far (like tpdcu_finish) and returns how many frames had to be rendered again since
 * tpdcu_create because they overflowed the grow-only pair buffers (warm-up frames, in practice). */
int tpdcu_frames_repeated(tpdcu_ctx* ctx, uint32_t* count);

/* ---- introspection (parity tests, benchmarks); all imply tpdcu_finish -------------------------- */

/* P = tilesRendered (prefix.slang:132); visible = Gaussians with tiles > 0. */
int tpdcu_get_counts(tpdcu_ctx* ctx, uint32_t* pairs, uint32_t* visible);
/* Splat buffer in the REFERENCE layout (48 B each, `tiles` holding the exclusive offset as after
 * prefix.slang). Fields other than radius/tiles of culled Gaussians are zero (the reference leaves
 * them stale, project.slang:34-35). A frame only evaluates the colour of the splats its blend stages;
 * this call evaluates it for every visible Gaussian first (same arithmetic). */
int tpdcu_read_splats(tpdcu_ctx* ctx, void* host_splats48, uint32_t n);
/* Sorted keys / values (first P entries) and per-tile (start,end) ranges. */
int tpdcu_read_keys(tpdcu_ctx* ctx, uint64_t* host_keys, uint32_t count);
int tpdcu_read_values(tpdcu_ctx* ctx, uint32_t* host_vals, uint32_t count);
int tpdcu_read_ranges(tpdcu_ctx* ctx, uint32_t* host_ranges2, uint32_t tile_count);
/* The reference's UNSORTED key/value buffers (keygen.slang:47-53: Gaussian i writes its tiles row-major at its prefix
 * offset). The frame never materialises them — its duplication stage runs over the depth-sorted Gaussians — so this
 * call rebuilds them from the last frame's per-Gaussian records, in the reference's order. */
int tpdcu_read_unsorted(tpdcu_ctx* ctx, uint64_t* host_keys, uint32_t* host_vals, uint32_t count);
/* What the duplication stage (emit_kernel, the stage that replaces keygen.slang:21-53) REALLY wrote for the newest frame: the
 * first `count` pair words (tile << extra | top depth bits) << 32 | Gaussian index, in depth order of the Gaussians and
 * row-major tile order within a Gaussian (`extra` = tile_bits reported by tpdcu_get_sort_info minus ceil(log2(tiles))).
 * The tile sort consumes that buffer, so the call renders the frame again up to the duplication stage, copies the words
 * and then renders it once more in full; parity tests use it to check the real emission, not a rebuilt one. */
int tpdcu_read_emitted(tpdcu_ctx* ctx, uint64_t* host_words, uint32_t count);
/* When enabled, CUDA events bracket every stage of each frame. times_ms (last finished frame):
 * [0] clear+camera setup [1] preprocess (projection, scan, visible compaction) + SH colour [2] depth sort of the visible
 * Gaussians [3] duplication [4] tile sort of the pairs [5] ranges [6] blend [7] whole frame
 * [8] / [9] number of onesweep passes the depth sort / the tile sort actually ran [10] histogram+plan share of [4]. */
int tpdcu_enable_stage_timing(tpdcu_ctx* ctx, int enable);
int tpdcu_stage_times_ms(tpdcu_ctx* ctx, float times_ms[TPDCU_NUM_STAGES]);
/* How the last finished frame was sorted (two levels, see csrc/sort.cu): the visible Gaussians by `depth_bits` low bits of
 * (depth - frame minimum) (words depth << 32 | index, 16 B moved per Gaussian and pass), then the pairs by `tile_bits` bits:
 * the tile id above whatever top depth bits did not fill a whole 8-bit digit of the depth sort (words key << 32 | index,
 * 16 B per pair and pass); *_passes = onesweep passes that ran. */
int tpdcu_get_sort_info(tpdcu_ctx* ctx, uint32_t* depth_bits, uint32_t* depth_passes, uint32_t* tile_bits, uint32_t* tile_passes);
/* The frame's launches between the camera setup and the blend do not change from frame to frame; they are captured once
 * into a CUDA graph and replayed (the reference re-records two command buffers every frame, GaussianEngine.cpp:637-697).
 * enable: 1/0 to switch replay on/off, -1 to only query. captures/launches (nullable): counters since tpdcu_create. */
int tpdcu_set_graph_replay(tpdcu_ctx* ctx, int enable, uint32_t* captures, uint32_t* launches);
/* Frames in flight (1..4, default 3), the counterpart of the reference's per-frame Frame objects (GaussianEngine.h:104-117,
 * SurfaceRenderer.h:66, which keeps 2): consecutive tpdcu_raster calls rotate through that many sets of per-frame buffers on
 * private streams so that the memory-bound front of the next frames overlaps the SM-bound blend of the current one. Ordering seen by the caller
 * is unchanged: the blend waits for `stream`, `stream` waits for the frame. tpdcu_finish/read_* refer to the newest frame;
 * an older frame that overflowed is re-rendered only if it went to a different target. */
int tpdcu_set_frames_in_flight(tpdcu_ctx* ctx, int frames);
/* Current pair-buffer capacity (grow-only, like GaussianEngine::reallocateBuffers :793-804) */
int tpdcu_get_capacity(tpdcu_ctx* ctx, uint32_t* capacity_pairs);
int tpdcu_reserve_pairs(tpdcu_ctx* ctx, uint32_t capacity_pairs);

/* ---- the sort on its own (replaces the radix loop, GaussianEngine.cpp:822-841) ------------------ */

/* Stable LSD onesweep sort of `n` (u64 key, u32 value) pairs already in device memory, on key bits
 * [0, end_bit). Result is written back into d_keys/d_vals. Scratch is owned by the context. */
int tpdcu_sort_pairs_device(tpdcu_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, uint32_t n, uint32_t end_bit,
                            void* stream);
/* Time (ms, CUDA events on `stream`) of the histogram+passes of the last tpdcu_sort_pairs_device call
 * excluding the copies in/out of the internal ping-pong buffers; and how many passes ran. */
int tpdcu_sort_last_ms(tpdcu_ctx* ctx, float* ms, uint32_t* passes_run);

#ifdef __cplusplus
}
#endif
#endif /* TPDCU_H */

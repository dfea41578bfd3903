// common.cuh — shared device/host declarations of libtpdcu (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tpdcu {

constexpr uint32_t TILE_PX = 16;          // BLOCK_X == BLOCK_Y (reference GaussianEngine.h:122-123)
constexpr uint32_t PRE_THREADS = 256;
#ifndef TPDCU_PRE_ITEMS
#define TPDCU_PRE_ITEMS 4
#endif
constexpr uint32_t PRE_ITEMS = TPDCU_PRE_ITEMS;  // Gaussians per thread
constexpr uint32_t PRE_BATCH = 4;         // of which this many are in flight at a time
constexpr uint32_t PRE_PART = PRE_THREADS * PRE_ITEMS;  // Gaussians per preprocess partition (one look-back each)
constexpr uint32_t SH_PLANES = 12;        // 48 SH floats = 12 float4 per Gaussian
constexpr uint32_t SORT_RADIX_BITS = 8;
constexpr uint32_t SORT_BINS = 1u << SORT_RADIX_BITS;
constexpr uint32_t SORT_MAX_PASSES = 8;   // 64-bit keys
#ifndef TPDCU_PDL
#define TPDCU_PDL 1
#endif
#ifndef TPDCU_SORT_KPT
#define TPDCU_SORT_KPT 32
#endif
#ifndef TPDCU_SORT_MINB
#define TPDCU_SORT_MINB 2
#endif
constexpr uint32_t SORT_THREADS = 256;
// keys per thread: 32 for the single-word sorts of a frame (8192-key tiles, two CTAs per SM: fewer look-backs and more
// independent work per thread beat occupancy), 16 for the standalone (key, value) pair sort, whose values double the state
constexpr uint32_t SORT_KPT_WORDS = TPDCU_SORT_KPT;
constexpr uint32_t SORT_KPT_PAIRS = 16;
constexpr uint32_t SORT_TILE_WORDS = SORT_THREADS * SORT_KPT_WORDS;
constexpr uint32_t SORT_TILE_PAIRS = SORT_THREADS * SORT_KPT_PAIRS;
constexpr uint32_t SORT_TILE = SORT_TILE_WORDS > SORT_TILE_PAIRS ? SORT_TILE_WORDS : SORT_TILE_PAIRS;  // buffer granularity
constexpr uint32_t SORT_WARPS = SORT_THREADS / 32;

// Look-back chains of a pass (sort.cu): the input of a pass is cut into SORT_CHAINS contiguous segments whose digit
// histograms are known up front, so every segment runs its own, eight times shorter, decoupled look-back. Segments of the
// first pass are equal ranges of positions; segments of pass p > 0 are the runs of 256 / SORT_CHAINS consecutive bins that
// pass p - 1 wrote (so a key's segment is a function of its previous digit and the histogram kernel can count it).
#ifndef TPDCU_SORT_CHAINS
#define TPDCU_SORT_CHAINS 4
#endif
constexpr uint32_t SORT_CHAINS = TPDCU_SORT_CHAINS;                    // words sorts; the standalone pair sort runs one chain
constexpr uint32_t SORT_CHAIN_BINS = SORT_BINS / SORT_CHAINS;          // previous-pass bins per segment
// Tile sort: every segment's last SORT_HALF_LAST full tiles are cut into half tiles, so that the end of a pass — where 1936
// tiles over 296 resident CTAs leave the CTAs finishing a whole tile-life apart — is made of half-size work (measured on the
// headline frame: 0 -> 63.0 us per pass, 9 or 18 -> 61.7-62.5, 37 -> 63.5; the depth sort, two tiles per CTA, loses with any).
#ifndef TPDCU_SORT_HALF_LAST
#define TPDCU_SORT_HALF_LAST (296 / 4 / TPDCU_SORT_CHAINS)
#endif
constexpr uint32_t SORT_HALF_LAST = TPDCU_SORT_HALF_LAST;
constexpr uint32_t SORT_WORD_PASSES = 4;                               // words sorts order at most 32 key bits
constexpr uint32_t SORT_CHAIN_ROWS = SORT_WORD_PASSES * SORT_CHAINS > SORT_MAX_PASSES ? SORT_WORD_PASSES * SORT_CHAINS : SORT_MAX_PASSES;
static_assert((SORT_CHAINS & (SORT_CHAINS - 1)) == 0 && SORT_CHAINS <= 32, "SORT_CHAINS is a power of two");

// Tickets and digit histograms of one radix sort.
struct SortCtl {
    uint32_t hist_done;                   // histogram CTAs that have added their counts (the last one makes the plan)
    uint32_t pad[3];
    uint32_t ticket[SORT_MAX_PASSES];
    uint32_t hist[SORT_MAX_PASSES][SORT_BINS];  // exclusive digit offsets of every pass (made by the plan)
    // row pass * chains + c: keys of segment c per digit of that pass (histogram kernel); the plan turns a pass's rows into
    // the exclusive sum over the segments before c: the offset of segment c's first key inside every bin's output run
    uint32_t chain_hist[SORT_CHAIN_ROWS][SORT_BINS];
};

// Per-frame control block; lives at the head of the per-frame zeroed region.
struct FrameCtl {
    uint32_t scan_ticket;                 // preprocess partition tickets
    uint32_t emit_ticket;                 // duplication partition tickets
    uint32_t pairs_total;                 // P (tilesRendered), may exceed capacity
    uint32_t visible;                     // Gaussians with tiles > 0
    uint32_t depth_max;                   // max float bits of viewZ over the visible Gaussians (atomicMax)
    uint32_t inv_depth_min;               // ~min float bits (atomicMax on the complement, so zero-init works)
    unsigned long long pairs64;           // P again, summed in 64 bits per partition: the packed scan carries at 2^32 (host-side check)
    SortCtl depth_sort;                   // visible Gaussians by view depth
    SortCtl tile_sort;                    // (tile, Gaussian) pairs by tile; also the standalone pair sort
};

// Which sort a launch performs: where n and the key geometry come from.
//   SORT_KIND_PAIRS  standalone API: (u64 key, u32 value) pairs, bits [0, end_bit), n from the host
//   SORT_KIND_DEPTH  words  float_bits(viewZ) << 32 | Gaussian index, n = FrameCtl::visible, sorted on the LOW bits of
//                    (depth - frame minimum), see DepthSplit
//   SORT_KIND_TILE   words  (tile << extra | top depth bits) << 32 | Gaussian index, n = min(FrameCtl::pairs_total, capacity)
enum : uint32_t { SORT_KIND_PAIRS = 0, SORT_KIND_DEPTH = 1, SORT_KIND_TILE = 2 };

// How a frame's sort key  tile | depth - minimum  is cut between the two sorts. Only the bits the frame's depth range
// occupies matter (26 at 1080p with the default planes). Both sorts work in 8-bit digits: when the depth bits leave a
// partial digit (26 = 3 x 8 + 2) and the tile sort has room in its last digit (13 tile bits use 2 x 8), those top depth bits
// ride below the tile id in the pair key and the depth sort drops a whole pass.
struct DepthSplit {
    uint32_t bias;        // frame minimum of float_bits(viewZ)
    uint32_t low_bits;    // bits of (depth - bias) sorted by the depth sort
    uint32_t extra;       // top bits of (depth - bias) carried in the pair key, below the tile id
};
__device__ __forceinline__ DepthSplit depth_split(const FrameCtl* fr, uint32_t tile_bits);

// Written by the plan kernel, read by every sort pass and by the consumers of the sorted result.
struct SortPlan {
    uint32_t n;                           // number of elements to sort
    uint32_t num_passes;
    uint32_t final_sel;                   // which ping-pong buffer holds the result
    uint32_t passes_run;
    uint32_t bias;                        // words: subtracted from the key (high 32 bits) before digit extraction
    uint32_t total_bits;                  // key bits the passes sort on
    uint32_t tile_shift;                  // tile sort: DepthSplit::extra, the tile id sits above that many depth bits
    uint32_t pad;
    uint32_t skip[SORT_MAX_PASSES];       // pass is an identity permutation (single occupied bin)
    uint32_t src_sel[SORT_MAX_PASSES];    // ping-pong buffer the pass reads from
    uint32_t chains;                      // look-back chains per pass: SORT_CHAINS (words) or 1 (pairs)
    uint32_t half_last;                   // per segment, the last this many full tiles' worth of elements is cut into half tiles
    uint32_t seg_start[SORT_MAX_PASSES][SORT_CHAINS + 1];   // first input position of every segment of the pass (+ n)
    uint32_t seg_tiles[SORT_MAX_PASSES][SORT_CHAINS + 1];   // tiles of the segments before c: descriptor row of its tile 0 (+ total)
};

// Camera-derived constants, produced once per frame by the setup kernel.
struct FrameCam {
    float V[16];                          // view matrix, row-major
    float focal[2];                       // 0.5 * size * focalNDC (project.slang:55)
    float cam_pos[3];                     // getCameraWorldPosition (splat/common.slang:88-92)
    float pad[3];
};

// Internal per-Gaussian outputs of the preprocess stage (the reference's 48-byte Splat, splat.slang:33-39, re-cut by
// consumer): `SplatGeo` is everything the blend stage needs to decide whether a pixel is touched (one 32-byte sector per
// gather), `color` is only fetched for splats that survive the per-tile cull, `depth_radius` is introspection-only.
struct __align__(32) SplatGeo {
    float px, py, conic_a, conic_b;         // float4 #0
    float conic_c, opacity, ext_x, ext_y;   // float4 #1: ext = half-extents of the alpha >= 1/255 region (conservative)
};
static_assert(sizeof(SplatGeo) == 32, "SplatGeo must be one 32-byte sector");

struct SplatArrays {
    SplatGeo* geo;                        // written for visible Gaussians only
    float4* color;                        // rgb (+pad), written for visible Gaussians only
    float2* depth_radius;                 // (viewZ, radius), visible only; read by the introspection exports
    uint2* rect;                          // tile rectangle (x0 | y0 << 16, w | h << 16), visible only
    uint32_t* offsets;                    // exclusive pair offset per Gaussian, n + 1 entries (prefix.slang semantics)
};

struct SceneArrays {
    const float4* posop;                  // xyz + opacity
    const float4* cov_a;                  // S00 S01 S02 S11
    const float2* cov_b;                  // S12 S22
    const float4* sh;                     // [n][SH_PLANES]: one 192-byte row per Gaussian, float index f = 3*coef + channel
    const uint32_t* entity;               // nullptr when the scene has a single entity
    uint32_t n;
    uint32_t entity_count;
};

// Lookback descriptor states (both the preprocess scan and the onesweep passes)
constexpr uint32_t FLAG_INVALID = 0u;
constexpr uint32_t FLAG_AGGREGATE = 1u;
constexpr uint32_t FLAG_PREFIX = 2u;

// Programmatic dependent launch (PDL). A kernel launched with pdl_launch() may become resident while its predecessor in the
// stream is still draining: pdl_wait() — the first statement of every kernel — holds it until the predecessor has completed and
// its writes are visible, and is a no-op for a kernel that was launched the ordinary way. pdl_release() lets the NEXT kernel of
// the stream be scheduled as soon as this grid's CTAs have all started: its blocks then take the SM slots this grid's last
// CTAs free one by one, instead of paying a launch and a ramp-up after the last of them has gone (a few microseconds per kernel
// boundary, fourteen boundaries per frame).
__device__ __forceinline__ void pdl_wait() {
#if TPDCU_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_release() {
#if TPDCU_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ DepthSplit depth_split(const FrameCtl* fr, uint32_t tile_bits) {
    const uint32_t dmin = ~fr->inv_depth_min, dmax = fr->depth_max;
    DepthSplit x;
    x.bias = dmax >= dmin ? dmin : 0u;
    const uint32_t depth_bits = dmax >= dmin ? 32u - __clz(dmax - dmin) : 0u;  // 0 when every Gaussian carries the same depth
    const uint32_t rem = depth_bits % SORT_RADIX_BITS;
    const uint32_t spare = (tile_bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS * SORT_RADIX_BITS - tile_bits;
    x.extra = (rem != 0 && rem <= spare) ? rem : 0u;
    x.low_bits = depth_bits - x.extra;
    return x;
}

// 256-bit global accesses (sm_100: LDG/STG.ENL2.256): a thread that owns 32 contiguous bytes moves a whole sector with one
// instruction instead of two half-sector ones. `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const uint64_t* p, uint64_t (&v)[4]) {
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
__device__ __forceinline__ void stg256(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(float4* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- spherical-harmonic colour (splat/common.slang:35-80, project.slang:82-83) -----------------
//
// Streaming form: the 48 coefficients of a Gaussian are consumed in memory order (float index f = 3 * coefficient + channel,
// four per 16-byte plane), so only one plane has to be live at a time. Every operation is an explicitly rounded
// __fmul_rn / __fadd_rn in the order of the reference's expression (terms added left to right, each term
// (constant * basis) * coefficient), so the result does not depend on the translation unit's FMA contraction setting.
// Band 3 keeps the reference's quirk: term 12 is "+ C3[3]*z*(2zz-3xx-3yy) + coefficient" (splat/common.slang:69).
struct ShBasis {
    float b[16];   // basis_k including its constant and sign: term_k = b[k] * coefficient_k  (k = 12: see above)
};
__device__ __forceinline__ ShBasis sh_basis(float x, float y, float z, int degree) {
    constexpr float C0 = 0.28209479177387814f;
    constexpr float C1 = 0.4886025119029199f;
    constexpr float C2[5] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                              0.5462742152960396f };
    constexpr float C3[7] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                              -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f };
    ShBasis s;
#pragma unroll
    for (int k = 0; k < 16; ++k) s.b[k] = 0.0f;
    s.b[0] = C0;
    if (degree > 0) {
        s.b[1] = __fmul_rn(C1, y);   // subtracted
        s.b[2] = __fmul_rn(C1, z);
        s.b[3] = __fmul_rn(C1, x);   // subtracted
        if (degree > 1) {
            const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), zx = __fmul_rn(z, x);
            s.b[4] = __fmul_rn(C2[0], xy);
            s.b[5] = __fmul_rn(C2[1], yz);
            s.b[6] = __fmul_rn(C2[2], __fsub_rn(__fsub_rn(__fmul_rn(2.0f, zz), xx), yy));
            s.b[7] = __fmul_rn(C2[3], zx);
            s.b[8] = __fmul_rn(C2[4], __fsub_rn(xx, yy));
            if (degree > 2) {
                s.b[9] = __fmul_rn(__fmul_rn(C3[0], y), __fsub_rn(__fmul_rn(3.0f, xx), yy));
                s.b[10] = __fmul_rn(__fmul_rn(C3[1], xy), z);
                s.b[11] = __fmul_rn(__fmul_rn(C3[2], y), __fsub_rn(__fsub_rn(__fmul_rn(4.0f, zz), xx), yy));
                s.b[12] = __fmul_rn(__fmul_rn(C3[3], z), __fsub_rn(__fsub_rn(__fmul_rn(2.0f, zz), __fmul_rn(3.0f, xx)), __fmul_rn(3.0f, yy)));
                s.b[13] = __fmul_rn(__fmul_rn(C3[4], x), __fsub_rn(__fsub_rn(__fmul_rn(4.0f, zz), xx), yy));
                s.b[14] = __fmul_rn(__fmul_rn(C3[5], z), __fsub_rn(xx, yy));
                s.b[15] = __fmul_rn(__fmul_rn(C3[6], x), __fsub_rn(xx, __fmul_rn(3.0f, yy)));
            }
        }
    }
    return s;
}
// one coefficient (float index f) into its channel's running sum
__device__ __forceinline__ void sh_accumulate(float (&acc)[3], const ShBasis& s, int f, float coef) {
    const int k = f / 3, c = f % 3;
    if (k == 0) acc[c] = __fmul_rn(s.b[0], coef);
    else if (k == 1 || k == 3) acc[c] = __fsub_rn(acc[c], __fmul_rn(s.b[k], coef));
    else if (k == 12) acc[c] = __fadd_rn(__fadd_rn(acc[c], s.b[12]), coef);
    else acc[c] = __fadd_rn(acc[c], __fmul_rn(s.b[k], coef));
}
__host__ __device__ constexpr int sh_planes(int degree) { return degree <= 0 ? 1 : degree == 1 ? 3 : degree == 2 ? 7 : 12; }

// view direction of the colour: normalize(mean - camera position) with IEEE sqrt and division (project.slang:82)
__device__ __forceinline__ float3 sh_direction(float px, float py, float pz, const float* cam_pos) {
    float dx = __fsub_rn(px, cam_pos[0]), dy = __fsub_rn(py, cam_pos[1]), dz = __fsub_rn(pz, cam_pos[2]);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    return make_float3(__fdiv_rn(dx, len), __fdiv_rn(dy, len), __fdiv_rn(dz, len));
}
// two consecutive planes (eight coefficients from float index 4 * p) into the running sums
__device__ __forceinline__ void sh_accumulate_planes(float (&acc)[3], const ShBasis& s, int p, int coefs, const float4& v, const float4& w) {
    if (4 * p + 0 < coefs) sh_accumulate(acc, s, 4 * p + 0, v.x);
    if (4 * p + 1 < coefs) sh_accumulate(acc, s, 4 * p + 1, v.y);
    if (4 * p + 2 < coefs) sh_accumulate(acc, s, 4 * p + 2, v.z);
    if (4 * p + 3 < coefs) sh_accumulate(acc, s, 4 * p + 3, v.w);
    if (4 * p + 4 < coefs) sh_accumulate(acc, s, 4 * p + 4, w.x);
    if (4 * p + 5 < coefs) sh_accumulate(acc, s, 4 * p + 5, w.y);
    if (4 * p + 6 < coefs) sh_accumulate(acc, s, 4 * p + 6, w.z);
    if (4 * p + 7 < coefs) sh_accumulate(acc, s, 4 * p + 7, w.w);
}
__host__ __device__ constexpr int sh_coefs(int degree) { return 3 * (degree <= 0 ? 1 : (degree + 1) * (degree + 1)); }
__device__ __forceinline__ float3 sh_finish(const float (&acc)[3]) {  // + 0.5, max 0 (splat/common.slang:79-80)
    return make_float3(fmaxf(__fadd_rn(acc[0], 0.5f), 0.0f), fmaxf(__fadd_rn(acc[1], 0.5f), 0.0f), fmaxf(__fadd_rn(acc[2], 0.5f), 0.0f));
}

// the same direction with one reciprocal square root (2^-22 relative error: the colour is tolerance-checked, 1/255) instead of
// an IEEE square root and three IEEE divisions: ~50 instructions less per staged splat of the blend
__device__ __forceinline__ float3 sh_direction_fast(float px, float py, float pz, const float* cam_pos) {
    const float dx = px - cam_pos[0], dy = py - cam_pos[1], dz = pz - cam_pos[2];
    const float inv = rsqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    return make_float3(dx * inv, dy * inv, dz * inv);
}
// colour of a Gaussian seen from the camera, its SH row read from global memory
template <bool FAST_DIRECTION = false>
__device__ __forceinline__ float3 sh_color(const float4* __restrict__ sh_row, float px, float py, float pz, const float* cam_pos,
                                           int degree) {
    const float3 d = FAST_DIRECTION ? sh_direction_fast(px, py, pz, cam_pos) : sh_direction(px, py, pz, cam_pos);
    const ShBasis s = sh_basis(d.x, d.y, d.z, degree);
    const int coefs = sh_coefs(degree);
    float acc[3] = { 0.0f, 0.0f, 0.0f };
    // Two planes (one 32-byte sector) per load: a thread that walks its own row costs the L1 one tag look-up per
    // request, and these rows are gathered (blend staging: 12 x 128-bit requests per splat made the L1 the bottleneck).
#pragma unroll
    for (int p = 0; p < (int)SH_PLANES; p += 2) {
        if (p < sh_planes(degree)) {
            float4 v, w;
            ldg256(sh_row + p, v, w);  // p + 1 <= 11: inside the 192-byte row whatever the degree
            sh_accumulate_planes(acc, s, p, coefs, v, w);
        }
    }
    return sh_finish(acc);
}

// ---- host-side launchers (one per translation unit) ---------------------------------------------

struct PreprocessLaunch {
    SceneArrays scene;
    const float* models;                  // device, entity_count x 16
    FrameCam* cam;                        // device scratch
    float* vm;                            // device scratch, entity_count x 16
    float* pm;                            // device scratch, entity_count x 16
    FrameCtl* ctl;
    uint64_t* scan_desc;                  // zeroed, one per partition
    SplatArrays out;
    uint64_t* depth_words;                // out: float_bits(viewZ) << 32 | index of every visible Gaussian, index order
    uint32_t width, height, sh_degree;
};
struct CameraUbo { float f[34]; };          // the reference's 136-byte Camera block, passed by value
cudaError_t launch_setup(const PreprocessLaunch& a, const CameraUbo& ubo, cudaStream_t s);
cudaError_t launch_preprocess(const PreprocessLaunch& a, cudaStream_t s);   // geometry + scan + visible compaction
cudaError_t launch_color(const PreprocessLaunch& a, cudaStream_t s);        // SH colour of the visible Gaussians

struct CompileLaunch {
    const float* recs240;                 // device, n x 60 floats
    float4* posop;
    float4* cov_a;
    float2* cov_b;
    float4* sh;
    uint32_t n;
};
cudaError_t launch_compile_scene(const CompileLaunch& a, cudaStream_t s);

// Duplication (keygen.slang) over the depth-sorted visible Gaussians: pairs come out ordered by (depth, index), so that the
// stable sort by tile that follows yields the reference's (tile, depth) order.
struct EmitLaunch {
    const uint64_t* depth_words[2];       // ping-pong buffers of the depth sort
    const SortPlan* depth_plan;           // which of them holds the result, and how many Gaussians are visible
    const uint2* rect;
    FrameCtl* ctl;
    uint64_t* scan_desc;                  // zeroed, one per partition
    uint64_t* keys;                       // out: (tile << extra | top depth bits) << 32 | index
    uint32_t n;                           // scene size (launch bound)
    uint32_t capacity;
    uint32_t width;
    uint32_t tile_bits;
};
cudaError_t launch_emit(const EmitLaunch& a, cudaStream_t s);
uint32_t emit_parts(uint32_t n);  // duplication partitions (one scan descriptor each)

// introspection: the reference's unsorted (key, value) buffers, in the reference's (index) order
cudaError_t launch_export_unsorted(const SplatArrays& a, uint32_t n, uint32_t width, uint64_t* keys, uint32_t* vals, uint32_t capacity,
                                   cudaStream_t s);

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute (see pdl_wait): the kernel may be scheduled before its
// predecessor in the stream has finished
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = TPDCU_PDL ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct SortLaunch {
    uint64_t* keys[2];                    // pair keys, or words
    uint32_t* vals[2];                    // SORT_KIND_PAIRS only
    FrameCtl* frame;                      // n and depth range (SORT_KIND_DEPTH / SORT_KIND_TILE)
    SortCtl* ctl;
    SortPlan* plan;
    uint32_t* lookback;                   // zeroed, [num_passes][parts_cap][SORT_BINS]
    uint32_t kind;
    uint32_t capacity;                    // launch bound for grids
    uint32_t end_bit;                     // pairs: end bit of the key; depth: 32; tile: tile bits rounded up to whole digits
    uint32_t tile_bits;                   // depth / tile: bits of the largest tile id (DepthSplit)
    int sm_count;
};
// n_host is used by SORT_KIND_PAIRS only.
cudaError_t launch_sort(const SortLaunch& a, uint32_t n_host, cudaStream_t s, cudaEvent_t ev_after_plan);
uint32_t sort_parts(uint32_t capacity, uint32_t kind);
uint32_t sort_passes_for(uint32_t end_bit);
cudaError_t init_sort_attributes();
cudaError_t sort_rank_selftest(uint32_t* mismatches_host);   // lanes whose atomic rank differs from the defined one (must be 0)

struct RasterLaunch {
    const uint64_t* keys[2];              // words tile << 32 | index, sorted
    const SortPlan* plan;                 // of the tile sort
    const SplatGeo* geo;
    const float4* color;                  // introspection only: the frame evaluates colours inside the blend
    const float2* depth_radius;
    uint32_t* ranges;                     // zeroed, tiles x 2
    uint32_t* order;                      // 2 x tiles: tile ids by decreasing expected cost (the blend's dispatch order) | scratch
    uint32_t* tile_cost;                  // tiles: splats the blend consumed per tile in an earlier frame (0: unknown); a hint
    const float4* posop;                  // scene: xyz + opacity (view direction of the SH colour)
    const float4* sh;                     // scene: [n][SH_PLANES]
    const FrameCam* cam;                  // camera position of this frame
    uint32_t sh_degree;
    uint8_t* out;
    size_t pitch;
    uint32_t width, height;
    int sm_count;
};
cudaError_t launch_ranges(const RasterLaunch& a, uint32_t capacity, cudaStream_t s);
cudaError_t launch_blend(const RasterLaunch& a, cudaStream_t s);

// introspection: sorted words -> the reference's (tile << 32 | depth bits, index) arrays
cudaError_t launch_sort_unpack(const RasterLaunch& a, uint64_t* out_keys, uint32_t* out_vals, int sm_count, cudaStream_t s);
cudaError_t launch_sort_copy_result(const SortLaunch& a, uint64_t* out_keys, uint32_t* out_vals, uint32_t n, cudaStream_t s);

cudaError_t launch_export_splats(const SplatArrays& a, uint32_t n, void* out48, cudaStream_t s);

}  // namespace tpdcu

// raster.cu — tile-range identification and per-tile front-to-back alpha blending.
//
// Replaces range.slang:16-34 (+ the fillBuffer of GaussianEngine.cpp:844) and blend.slang:22-104
// (+ the R8G8B8A8_UNORM image store of the Vulkan target, GaussianEngine.cpp:316-319).
// Colour is tolerance-checked (1/255), so FMA contraction and ex2.approx are allowed here; ranges
// are integer work and bit-exact.
#include "common.cuh"

namespace tpdcu {

// ---------------------------------------------------------------------------------------------------
// ranges: ranges[tile] = (first, last+1) over the sorted words; empty tiles stay (0,0)
// ---------------------------------------------------------------------------------------------------

// The words are sorted, so the tile id is non-decreasing: an interval of RANGE_STRIDE words holds a boundary only if its first
// word and the first word of the next interval differ. A thread looks at ONE word per interval (its head; the neighbour's head
// comes by shuffle); a warp whose 32 intervals hold boundaries in at most RANGE_SPARSE_MAX of them reads just those together, one
// after the other, and in a warp over a stretch of short lists (the sparse tiles at the picture's border sit next to each other
// in the sorted array) every lane scans its own interval. 35 MB of DRAM traffic for 15.86 M words at 1080p
// instead of the 127 MB of reading them all.
#ifndef TPDCU_RANGES_SAMPLED
#define TPDCU_RANGES_SAMPLED 1
#endif
#ifndef TPDCU_RANGE_STRIDE
#define TPDCU_RANGE_STRIDE 64
#endif
#ifndef TPDCU_RANGE_SPARSE_MAX
#define TPDCU_RANGE_SPARSE_MAX 2
#endif
#ifndef TPDCU_RANGES_CTAS_PER_SM
#define TPDCU_RANGES_CTAS_PER_SM 8   // all CTAs resident at once: the kernel is a few dependent memory latencies long
#endif
#ifndef TPDCU_RANGE_LANE_BATCH
#define TPDCU_RANGE_LANE_BATCH 2
#endif
constexpr uint32_t RANGE_STRIDE = TPDCU_RANGE_STRIDE;
constexpr uint32_t RANGE_SPARSE_MAX = TPDCU_RANGE_SPARSE_MAX;
static_assert(RANGE_STRIDE % 32 == 0 && RANGE_STRIDE >= 32, "a warp reads an interval in whole rows of 32 words");

#if TPDCU_RANGES_SAMPLED
__global__ void __launch_bounds__(256, TPDCU_RANGES_CTAS_PER_SM) ranges_kernel(RasterLaunch a) {
    pdl_wait();
    pdl_release();
    const uint32_t n = a.plan->n;
    if (n == 0) return;
    const uint64_t* __restrict__ keys = a.plan->final_sel ? a.keys[1] : a.keys[0];
    const uint32_t tshift = 32u + a.plan->tile_shift;  // words are (tile << shift | top depth bits) << 32 | Gaussian index
    uint2* ranges = reinterpret_cast<uint2*>(a.ranges);
    constexpr uint32_t NONE = 0xffffffffu, FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5), warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t intervals = (n + RANGE_STRIDE - 1) / RANGE_STRIDE;
    if (warp_id == 0 && lane == 0) {   // the two ends of the array
        ranges[(uint32_t)(keys[0] >> tshift)].x = 0;
        ranges[(uint32_t)(keys[n - 1] >> tshift)].y = n;
    }
    for (uint32_t i0 = warp_id * 32u; i0 < intervals; i0 += warps * 32u) {   // warp-uniform
        const uint32_t i = i0 + lane;
        const uint32_t head = i < intervals ? (uint32_t)(__ldg(keys + (size_t)i * RANGE_STRIDE) >> tshift) : NONE;
        uint32_t next = __shfl_down_sync(FULL, head, 1);
        if (lane == 31u) next = i + 1u < intervals ? (uint32_t)(__ldg(keys + (size_t)(i + 1u) * RANGE_STRIDE) >> tshift) : NONE;
        // the last interval has no successor (NONE differs from every tile): it is scanned, which is harmless
        uint32_t todo = __ballot_sync(FULL, i < intervals && head != next);
        if (__popc(todo) > RANGE_SPARSE_MAX) {
            // a stretch of short lists: every lane scans its OWN interval (whole 32-byte sectors, RANGE_LANE_BATCH loads in flight),
            // so the warp's time does not grow with the number of intervals that hold a boundary
            if ((todo >> lane) & 1u) {
                const uint32_t base = i * RANGE_STRIDE;
                uint32_t prev = head;
                constexpr uint32_t GROUPS = RANGE_STRIDE / 4, BATCH = TPDCU_RANGE_LANE_BATCH;
                for (uint32_t g0 = 0; g0 < GROUPS; g0 += BATCH) {
                    uint64_t k[BATCH][4];
#pragma unroll
                    for (uint32_t g = 0; g < BATCH; ++g) {
                        const uint32_t idx = base + 4u * (g0 + g);
                        if (idx + 4u <= n) {
                            ldg256(keys + idx, k[g]);
                        } else {
#pragma unroll
                            for (uint32_t q = 0; q < 4; ++q) k[g][q] = idx + q < n ? __ldg(keys + idx + q) : ~0ull;
                        }
                    }
#pragma unroll
                    for (uint32_t g = 0; g < BATCH; ++g) {
#pragma unroll
                        for (uint32_t q = 0; q < 4; ++q) {
                            const uint32_t idx = base + 4u * (g0 + g) + q;
                            const uint32_t t = (uint32_t)(k[g][q] >> tshift);
                            if (idx < n && t != prev) {   // never at idx == base: prev starts as that word's tile
                                ranges[prev].y = idx;
                                ranges[t].x = idx;
                            }
                            prev = idx < n ? t : prev;
                        }
                    }
                }
                if (next != NONE && next != prev) {   // the boundary at the head of the next interval
                    ranges[prev].y = base + RANGE_STRIDE;
                    ranges[next].x = base + RANGE_STRIDE;
                }
            }
            continue;
        }
        while (todo) {
            const uint32_t src = __ffs(todo) - 1u;
            todo &= todo - 1u;
            const uint32_t base = (i0 + src) * RANGE_STRIDE;
            uint32_t carry = __shfl_sync(FULL, head, src);   // tile of word `base`
            // words base+1 .. base+RANGE_STRIDE (the last one is the next interval's head: its boundary is reported here)
            uint32_t cur[RANGE_STRIDE / 32];
#pragma unroll
            for (uint32_t h = 0; h < RANGE_STRIDE / 32; ++h) {
                const uint32_t idx = base + 1u + h * 32u + lane;
                cur[h] = idx < n ? (uint32_t)(__ldg(keys + idx) >> tshift) : NONE;
            }
#pragma unroll
            for (uint32_t h = 0; h < RANGE_STRIDE / 32; ++h) {
                const uint32_t idx = base + 1u + h * 32u + lane;
                uint32_t prev = __shfl_up_sync(FULL, cur[h], 1);
                if (lane == 0u) prev = carry;
                carry = __shfl_sync(FULL, cur[h], 31);
                if (idx < n && cur[h] != prev) {
                    ranges[prev].y = idx;
                    ranges[cur[h]].x = idx;
                }
            }
        }
    }
}
#else
__global__ void __launch_bounds__(256) ranges_kernel(RasterLaunch a) {
    pdl_wait();
    pdl_release();
    const uint32_t n = a.plan->n;
    const uint64_t* __restrict__ keys = a.plan->final_sel ? a.keys[1] : a.keys[0];
    const uint32_t tshift = 32u + a.plan->tile_shift;  // words are (tile << shift | top depth bits) << 32 | Gaussian index
    uint2* ranges = reinterpret_cast<uint2*>(a.ranges);
    // four words per thread with one 32-byte load; the word before the group comes from the neighbour's sector (L1 hit)
    const uint32_t groups = (n + 3) / 4;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < groups; t += gridDim.x * blockDim.x) {
        const uint32_t idx = 4 * t;
        uint64_t k[4];
        if (idx + 4 <= n) {
            ldg256(keys + idx, k);
        } else {
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) k[q] = idx + q < n ? keys[idx + q] : 0ull;
        }
        uint32_t prev = idx == 0 ? 0xffffffffu : (uint32_t)(keys[idx - 1] >> tshift);
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            if (idx + q < n) {
                const uint32_t tq = (uint32_t)(k[q] >> tshift);
                if (tq != prev) {
                    if (idx + q != 0) ranges[prev].y = idx + q;
                    ranges[tq].x = idx + q;
                }
                if (idx + q == n - 1) ranges[tq].y = n;
                prev = tq;
            }
        }
    }
}
#endif

// ---------------------------------------------------------------------------------------------------
// blend order: tiles by decreasing expected cost (longest processing time first)
//
// The blend's cost per tile is the number of splats it consumes before its last pixel saturates (mean 320, max 4.2 k on
// the headline scene, lists: mean 1.9 k, max 20.8 k), the heaviest tile alone costs about what an average CTA slot
// processes in the whole kernel, and the hardware dispatches CTAs in blockIdx order: in image order the heavy tiles of
// the picture's centre start half-way through and the kernel ends on their tail (ncu: issue slots 65 % busy while an SM
// has work, 46 % over the kernel). A counting sort of the tiles into 256 log-scale buckets (8 per octave) gives the
// dispatch order; one CTA, ~8 us. Expected cost: what the blend consumed on this tile in an earlier frame of this context
// (+50 %, a renderer's consecutive views are close), capped by the list length; the list length alone before that.
// The order changes when a tile is processed, never what is computed for it.
// ---------------------------------------------------------------------------------------------------

// Splats of a tile's list the blend is expected to consume: what it consumed on that tile in an earlier frame of this
// context + 50 % + 256, capped by the list; `unknown` entries when there is no such frame.
__device__ __forceinline__ uint32_t expected_front(uint32_t len, uint32_t seen, uint32_t unknown) {
    return seen ? min(len, seen + seen / 2u + 256u) : min(len, unknown);
}

constexpr uint32_t ORDER_THREADS = 1024;
constexpr uint32_t ORDER_BUCKETS = 256;
constexpr uint32_t ORDER_UNROLL = 8;

__device__ __forceinline__ uint32_t order_bucket(uint32_t len) {
    const uint32_t w = 32u - __clz(len);                                 // bit length, 0 for an empty tile
    const uint32_t sub = w >= 4u ? (len >> (w - 4u)) & 7u : (len << (4u - w)) & 7u;  // the three bits below the leading one
    return ORDER_BUCKETS - 1u - min(w * 8u + sub, ORDER_BUCKETS - 1u);  // bucket 0 = longest lists
}

__global__ void __launch_bounds__(ORDER_THREADS) tile_order_kernel(RasterLaunch a, uint32_t tiles) {
    pdl_wait();
    __shared__ uint32_t s_off[ORDER_BUCKETS];
    __shared__ uint32_t s_warp[ORDER_BUCKETS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint2* __restrict__ ranges = reinterpret_cast<const uint2*>(a.ranges);
    if (tid < ORDER_BUCKETS) s_off[tid] = 0;
    __syncthreads();
    // The hints are written by the blends of other frames in flight while this kernel runs: read each one ONCE. (Counting
    // with one value and scattering with another would leave `order` no permutation — a tile rendered twice, one never.)
    uint32_t* bucket_of = a.order + tiles;
    // one CTA, so the global loads are the critical path: ORDER_UNROLL tiles per thread in flight at a time
    for (uint32_t t0 = tid; t0 < tiles; t0 += ORDER_THREADS * ORDER_UNROLL) {
        uint2 r[ORDER_UNROLL];
        uint32_t seen[ORDER_UNROLL];
#pragma unroll
        for (uint32_t k = 0; k < ORDER_UNROLL; ++k) {
            const uint32_t t = t0 + k * ORDER_THREADS;
            r[k] = t < tiles ? ranges[t] : make_uint2(0u, 0u);
            seen[k] = t < tiles ? a.tile_cost[t] : 0u;
        }
#pragma unroll
        for (uint32_t k = 0; k < ORDER_UNROLL; ++k) {
            const uint32_t t = t0 + k * ORDER_THREADS;
            if (t < tiles) {
                const uint32_t b = order_bucket(expected_front(r[k].y - r[k].x, seen[k], 0xffffffffu));
                bucket_of[t] = b;
                atomicAdd(&s_off[b], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t c = 0, incl = 0;
    if (tid < ORDER_BUCKETS) {
        c = s_off[tid];
        incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
    }
    __syncthreads();
    if (tid < ORDER_BUCKETS) {
        uint32_t before = 0;
        for (uint32_t w2 = 0; w2 < warp; ++w2) before += s_warp[w2];
        s_off[tid] = before + incl - c;
    }
    __syncthreads();
    for (uint32_t t = tid; t < tiles; t += ORDER_THREADS)
        a.order[atomicAdd(&s_off[bucket_of[t]], 1u)] = t;   // order inside a bucket is irrelevant; same thread wrote bucket_of[t]
}

cudaError_t launch_ranges(const RasterLaunch& a, uint32_t capacity, cudaStream_t s) {
    const uint32_t tiles = ((a.width + TILE_PX - 1) / TILE_PX) * ((a.height + TILE_PX - 1) / TILE_PX);
    if (capacity != 0) {
#if TPDCU_RANGES_SAMPLED
        uint32_t grid = (capacity / RANGE_STRIDE + 256) / 256;   // one thread per interval of RANGE_STRIDE words
#else
        uint32_t grid = (capacity / 4 + 256) / 256;
#endif
#if TPDCU_RANGES_SAMPLED
        const uint32_t cap = (uint32_t)(a.sm_count > 0 ? a.sm_count : 148) * TPDCU_RANGES_CTAS_PER_SM;
#else
        const uint32_t cap = (uint32_t)(a.sm_count > 0 ? a.sm_count : 148) * 8u;
#endif
        if (grid > cap) grid = cap;
        pdl_launch(ranges_kernel, grid, 256, 0, s, a);
    }
    if (tiles != 0) pdl_launch(tile_order_kernel, 1, ORDER_THREADS, 0, s, a, tiles);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// blend: one CTA per 16x16 tile, four warps = four 8x8 pixel quadrants, each lane owning two horizontally adjacent pixels.
//
// Splats are staged through shared memory in sorted order. While staging, the conservative bounding box of each splat's
// alpha >= 1/255 region (SplatGeo::ext_*, written by the preprocess kernel) is tested against the tile and against each
// quadrant: a splat that cannot reach a quadrant is never seen by that quadrant's warp (every pixel would `continue` on
// it, blend.slang:85,89, so the result is unchanged). On the headline scene the warps iterate over ~35 % of the
// (splat, pixel) pairs the reference evaluates. Each warp walks its own in-order list of queue entries; the conic is
// pre-scaled by -0.5*log2(e) at staging so that alpha = opacity * ex2(p).
// ---------------------------------------------------------------------------------------------------

constexpr uint32_t BLEND_THREADS = 128;
constexpr uint32_t BLEND_WARPS = BLEND_THREADS / 32;
// One staging pass per fill: with the colour evaluated at staging, every splat staged beyond the point where the tile
// saturates costs an SH evaluation (a 256-entry queue, two passes per fill, measured 3 % slower).
constexpr uint32_t BLEND_QUEUE = 128;
constexpr float LOG2E = 1.4426950408889634f;
static_assert(BLEND_WARPS == 4, "one warp per 8x8 quadrant of the 16x16 tile");

__device__ __forceinline__ uint32_t unorm8(float c) {
    // clamp to [0,1] (NaN -> 0), x255, round to nearest even: the R8G8B8A8_UNORM image store
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint32_t)__float2int_rn(c * 255.0f);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct BlendEntry {
    float4 g0;   // px, py, A', B'
    float4 g1;   // C', opacity, power2 threshold, -
    float4 col;  // r, g, b, -
};
static_assert(sizeof(BlendEntry) * BLEND_QUEUE <= 65536, "the per-quadrant lists hold 16-bit byte offsets into ent[]");
struct BlendSmem {
    BlendEntry ent[BLEND_QUEUE];
    uint16_t list[BLEND_WARPS][BLEND_QUEUE];     // per-quadrant byte offsets of queue entries, in sorted order
    uint32_t cnt[BLEND_WARPS][BLEND_WARPS + 1];  // [staging warp][tile, quadrant 0..3] survivors of the current round
};

// Shared-memory loads by 32-bit shared-window address: the drain loop below is instruction-bound, and these keep the
// address arithmetic down to one add per splat.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

#ifndef TPDCU_BLEND_FAST_DIRECTION
#define TPDCU_BLEND_FAST_DIRECTION true
#endif
#ifndef TPDCU_BLEND_GROUPS
#define TPDCU_BLEND_GROUPS 1
#endif
#ifndef TPDCU_BLEND_PIN_SMEM_BASES
#define TPDCU_BLEND_PIN_SMEM_BASES 1
#endif
#ifndef TPDCU_BLEND_MINB
#define TPDCU_BLEND_MINB 7
#endif
__global__ void __launch_bounds__(BLEND_THREADS, TPDCU_BLEND_MINB) blend_kernel(RasterLaunch a) {
    __shared__ BlendSmem sm;

    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX;
    const uint32_t tile = a.order[blockIdx.x];  // longest lists first (tile_order_kernel)
    const uint32_t tile_x0 = (tile % gx) * TILE_PX, tile_y0 = (tile / gx) * TILE_PX;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // quadrant `warp`: origin (8*(warp&1), 8*(warp>>1)); lane -> pixel pair at (2*(lane&3), lane>>2) inside it
    const uint32_t x0 = tile_x0 + 8u * (warp & 1u) + 2u * (lane & 3u), y0 = tile_y0 + 8u * (warp >> 1) + (lane >> 2);
    float fx0 = (float)x0, fy0 = (float)y0;
    asm volatile("" : "+f"(fx0), "+f"(fy0));  // keep the pixel coordinates in registers: do not re-convert them per splat
    const float tile_fx0 = (float)tile_x0, tile_fy0 = (float)tile_y0;

    const uint64_t* __restrict__ words = a.plan->final_sel ? a.keys[1] : a.keys[0];
    const float cam_pos[3] = { a.cam->cam_pos[0], a.cam->cam_pos[1], a.cam->cam_pos[2] };
    const uint2 range = reinterpret_cast<const uint2*>(a.ranges)[tile];
    const float4* __restrict__ geo4 = reinterpret_cast<const float4*>(a.geo);

    uint32_t inside = 0;  // bit k: pixel (x0 + k, y0) is inside the image
    if (y0 < a.height) {
        if (x0 < a.width) inside |= 1u;
        if (x0 + 1u < a.width) inside |= 2u;
    }
    // Upper bound of the accepted power per pixel: 0 while the pixel is accumulating (power > 0 is skipped,
    // blend.slang:85), -inf once it is done or if it lies outside the image, so that one compare covers both tests.
    const float NEG_INF = __int_as_float(0xff800000);
    float u0 = (inside & 1u) ? 0.0f : NEG_INF, u1 = (inside & 2u) ? 0.0f : NEG_INF;
    uint32_t ent_s = (uint32_t)__cvta_generic_to_shared(&sm.ent[0]);
    uint32_t list_s = (uint32_t)__cvta_generic_to_shared(&sm.list[warp][0]);
    // ptxas would otherwise re-derive these three loop invariants inside the drain loop (5 instructions per 2 splats)
    // to save registers; a value that went through a shuffle cannot be rematerialised.
    ent_s = __shfl_sync(0xffffffffu, ent_s, lane);
#if TPDCU_BLEND_PIN_SMEM_BASES
    // ... and the shuffle of a uniform value by the own lane IS folded away: ptxas re-derived both shared-window addresses
    // inside the drain loop (S2R SR_CgaCtaId, MOV, LEA, IMAD, IADD3 per two splats). An empty asm with read-write operands is opaque.
    asm volatile("" : "+r"(ent_s), "+r"(list_s));
#endif
    fx0 = __shfl_sync(0xffffffffu, fx0, lane);
    fy0 = __shfl_sync(0xffffffffu, fy0, lane);
    float T0 = 1.0f, T1 = 1.0f, r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;

    uint32_t in = range.x;
    // Software pipeline of the staging loads. A round's candidates are reached through two dependent gathers (sorted word ->
    // SplatGeo record) before the cull, and the survivors through a third (position + SH row) before the colour; with all
    // three issued inside the round every fill exposed three memory latencies (ncu: 36 % of the kernel's stall samples on
    // them, 16 % more on the barriers behind them). The words are fetched two rounds ahead and the SplatGeo records one round
    // ahead — both right after the cull of the current round, so they travel during its colour evaluation and its drain —
    // and a round only waits for the gather that depends on the cull: the survivors' rows. (Measured and discarded for
    // those: one cp.async.bulk per survivor into shared memory behind a per-warp mbarrier — 128 small TMA copies per round
    // are no faster than the loads and cost a CTA per SM in shared memory (0.349 ms); the SH degree as a template parameter
    // (0.353 ms); prefetch.global.L2 of the survivors' rows at cull time (0.373 ms against 0.342 ms); four fill warps + four drain
    // warps per tile with a double-buffered queue behind mbarriers (profiles/experiments/blend_ws_kernel.cuh: 0.343-0.416 ms).)
    uint32_t g_cur = 0, g_next = 0;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
    if (in + tid < range.y) g_cur = (uint32_t)__ldg(words + in + tid);
    if (in + BLEND_THREADS + tid < range.y) g_next = (uint32_t)__ldg(words + in + BLEND_THREADS + tid);
    if (in + tid < range.y) ldg256(geo4 + (size_t)g_cur * 2, ra, rb);  // the 32-byte SplatGeo record: one sector, one load
    while (true) {
        // ---- fill: append the next splats that can touch this tile / each quadrant, in order --------------------------
        uint32_t qn = 0;                                // queue entries
        uint32_t ln[BLEND_WARPS] = { 0u, 0u, 0u, 0u };  // list length of every quadrant (uniform across the CTA)
        while (qn + BLEND_THREADS <= BLEND_QUEUE && in < range.y) {
            const uint32_t idx = in + tid;
            uint32_t keep = 0;  // bit 0: tile, bits 1..4: quadrants 0..3
            const uint32_t g = g_cur;
            if (idx < range.y) {
                const float xlo = ra.x - rb.z - tile_fx0, xhi = ra.x + rb.z - tile_fx0;  // bbox relative to the tile origin
                const float ylo = ra.y - rb.w - tile_fy0, yhi = ra.y + rb.w - tile_fy0;
                const bool left = xhi >= 0.0f && xlo <= 7.0f, right = xhi >= 8.0f && xlo <= 15.0f;
                const bool top = yhi >= 0.0f && ylo <= 7.0f, bottom = yhi >= 8.0f && ylo <= 15.0f;
                keep = ((left && top) ? 2u : 0u) | ((right && top) ? 4u : 0u) | ((left && bottom) ? 8u : 0u) |
                       ((right && bottom) ? 16u : 0u);
                if (keep) keep |= 1u;
            }
            uint32_t ballot[BLEND_WARPS + 1];
#pragma unroll
            for (uint32_t q = 0; q <= BLEND_WARPS; ++q) {
                ballot[q] = __ballot_sync(0xffffffffu, (keep >> q) & 1u);
                if (lane == q) sm.cnt[warp][q] = __popc(ballot[q]);
            }
            // the next round's SplatGeo record (its word arrived a round ago) and the word of the round after it
            float4 ra_n = make_float4(0.f, 0.f, 0.f, 0.f), rb_n = ra_n;
            uint32_t g_next2 = 0;
            if (idx + BLEND_THREADS < range.y) ldg256(geo4 + (size_t)g_next * 2, ra_n, rb_n);
            if (idx + 2 * BLEND_THREADS < range.y) g_next2 = (uint32_t)__ldg(words + idx + 2 * BLEND_THREADS);
            __syncthreads();
            uint32_t before[BLEND_WARPS + 1], total[BLEND_WARPS + 1];
#pragma unroll
            for (uint32_t q = 0; q <= BLEND_WARPS; ++q) {
                before[q] = 0;
                total[q] = 0;
#pragma unroll
                for (uint32_t w = 0; w < BLEND_WARPS; ++w) {
                    const uint32_t c = sm.cnt[w][q];
                    if (w < warp) before[q] += c;
                    total[q] += c;
                }
            }
            if (keep) {
                const uint32_t pos = qn + before[0] + __popc(ballot[0] & lanemask_lt());
                // Colour on demand: the blend only ever stages the front of every list before its pixels saturate (19 % of
                // the visible Gaussians on the headline scene), so the SH colour (project.slang:82-83) is evaluated here,
                // for the splats that reach a queue, instead of by a kernel over all visible Gaussians (that kernel ran at
                // the HBM roofline: 1.2 GB per frame, 0.19 ms; here 0.56 GB are gathered inside a kernel that is not
                // memory-bound). Same arithmetic as the introspection kernel (common.cuh: sh_basis / sh_accumulate).
                const float4 po = __ldg(a.posop + g);
                sm.ent[pos].g0 = make_float4(ra.x, ra.y, (-0.5f * LOG2E) * ra.z, -LOG2E * ra.w);
                sm.ent[pos].g1 = make_float4((-0.5f * LOG2E) * rb.x, rb.y, -__log2f(255.0f * rb.y) - 0.01f, 0.0f);
#pragma unroll
                for (uint32_t q = 0; q < BLEND_WARPS; ++q)
                    if (keep & (2u << q)) sm.list[q][ln[q] + before[q + 1] + __popc(ballot[q + 1] & lanemask_lt())] = (uint16_t)(pos * sizeof(BlendEntry));
                const float3 c3 = sh_color<TPDCU_BLEND_FAST_DIRECTION>(a.sh + (size_t)g * SH_PLANES, po.x, po.y, po.z, cam_pos, (int)a.sh_degree);
                sm.ent[pos].col = make_float4(c3.x, c3.y, c3.z, 0.0f);
            }
            qn += total[0];
#pragma unroll
            for (uint32_t q = 0; q < BLEND_WARPS; ++q) ln[q] += total[q + 1];
            in += BLEND_THREADS;
            g_cur = g_next; g_next = g_next2; ra = ra_n; rb = rb_n;
            __syncthreads();
        }

        // ---- drain: front-to-back compositing (blend.slang:77-100) over this quadrant's list --------------------------
        const uint32_t my_ln = warp == 0 ? ln[0] : warp == 1 ? ln[1] : warp == 2 ? ln[2] : ln[3];
        // one splat of the quadrant's list against this lane's two pixels; `off` = byte offset of its queue entry
        auto composite = [&](uint32_t off) {
            const uint32_t e = ent_s + off;
            const float4 q0 = lds_f4(e);
            const float4 q1 = lds_f4(e + 16u);
            const float dx0 = q0.x - fx0, dy = q0.y - fy0;
            const float dx1 = dx0 - 1.0f;
            const float by = q0.w * dy;
            const float cy = q1.x * dy * dy;
            const float p0 = fmaf(dx0, fmaf(q0.z, dx0, by), cy);   // A'dx^2 + B'dx dy + C'dy^2 = log2(e) * power
            const float p1 = fmaf(dx1, fmaf(q0.z, dx1, by), cy);
            // Straight-line, predicated update of both pixels: only ~1/3 of the lanes get here for a typical splat, so
            // branch (re)convergence would cost more than the arithmetic it skips.
            const bool h0 = p0 <= u0 && p0 >= q1.z;    // below the threshold alpha < 1/255 (blend.slang:89)
            const bool h1 = p1 <= u1 && p1 >= q1.z;
            if (!(h0 || h1)) return;
            const float4 c = lds_f4(e + 32u);
            const float a0 = fminf(0.99f, q1.y * ex2_approx(p0)), a1 = fminf(0.99f, q1.y * ex2_approx(p1));
            const float w0 = a0 * T0, w1 = a1 * T1;
            const float t0 = T0 - w0, t1 = T1 - w1;    // T (1 - alpha)
            const bool v0 = h0 && a0 >= 1.0f / 255.0f, v1 = h1 && a1 >= 1.0f / 255.0f;
            const bool s0 = v0 && t0 >= 0.0001f, s1 = v1 && t1 >= 0.0001f;  // the splat is added (blend.slang:92-98) ...
            u0 = (v0 && !s0) ? NEG_INF : u0;                                // ... else the pixel is done and it is NOT
            u1 = (v1 && !s1) ? NEG_INF : u1;
            const float m0 = s0 ? w0 : 0.0f, m1 = s1 ? w1 : 0.0f;
            r0 = fmaf(c.x, m0, r0); g0 = fmaf(c.y, m0, g0); b0 = fmaf(c.z, m0, b0);
            r1 = fmaf(c.x, m1, r1); g1 = fmaf(c.y, m1, g1); b1 = fmaf(c.z, m1, b1);
            T0 = s0 ? t0 : T0;
            T1 = s1 ? t1 : T1;
        };
        // Groups of eight splats: one vote whether the whole quadrant is done (uniform branch), ONE 16-byte load of the eight list
        // entries, eight bodies back to back with no loop control between them (the rolled loop spent 5.5 of its ~26 common-path
        // instructions per splat on the counter, the every-eighth test and the 16-bit list load); the last < 8 splats one by one.
        uint32_t j = 0;
        bool quadrant_done = false;
#if TPDCU_BLEND_GROUPS
        for (; j + 8u <= my_ln; j += 8u) {
            if (__all_sync(0xffffffffu, u0 < 0.0f && u1 < 0.0f)) { quadrant_done = true; break; }
            const uint4 offs = lds_u4(list_s + 2u * j);
            composite(offs.x & 0xffffu); composite(offs.x >> 16);
            composite(offs.y & 0xffffu); composite(offs.y >> 16);
            composite(offs.z & 0xffffu); composite(offs.z >> 16);
            composite(offs.w & 0xffffu); composite(offs.w >> 16);
        }
#endif
        if (!quadrant_done && j < my_ln) {
#pragma unroll 2
            for (; j < my_ln; ++j) {
                if ((j & 7u) == 0u && __all_sync(0xffffffffu, u0 < 0.0f && u1 < 0.0f)) break;
                composite(lds_u16(list_s + 2u * j));
            }
        }
        // block vote (blend.slang:56-63); also the barrier that lets the queue be refilled
        const bool finished = in >= range.y;
        if (__syncthreads_and(u0 < 0.0f && u1 < 0.0f) || finished) break;
    }

    if (tid == 0) a.tile_cost[tile] = max(min(in, range.y) - range.x, 1u);  // hint for the next frames' dispatch order
    if (inside & 1u)
        *reinterpret_cast<uint32_t*>(a.out + (size_t)y0 * a.pitch + (size_t)x0 * 4) =
            unorm8(r0) | (unorm8(g0) << 8) | (unorm8(b0) << 16) | 0xff000000u;
    if (inside & 2u)
        *reinterpret_cast<uint32_t*>(a.out + (size_t)y0 * a.pitch + (size_t)(x0 + 1u) * 4) =
            unorm8(r1) | (unorm8(g1) << 8) | (unorm8(b1) << 16) | 0xff000000u;
}

cudaError_t launch_blend(const RasterLaunch& a, cudaStream_t s) {
    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX, gy = (a.height + TILE_PX - 1) / TILE_PX;
    if (gx * gy == 0) return cudaSuccess;
    blend_kernel<<<gx * gy, BLEND_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tpdcu

// sort.cu — stable LSD radix sort in the onesweep style (one up-front multi-digit histogram, then ONE read + ONE write
// of every element per 8-bit digit, with a chained decoupled look-back across tiles instead of a separate scan per pass).
//
// Replaces the reference's 4-way radix sorter: radix-shuffle.slang:37-150, radix-prefixA.slang:37-169,
// radix-prefixB.slang:37-168, radix-mapping.slang:38-116 driven by GaussianEngine.cpp:822-841 (23 passes x 4 dispatches
// over every (tile | depth, index) pair at 1080p, ~1.1 KB of DRAM traffic per pair).
//
// A frame sorts in two levels (tpdcu.cu): the VISIBLE GAUSSIANS by depth (words depth << 32 | index, <= 4 passes over
// ~n elements), then — after the duplication stage has emitted the pairs in that order — the PAIRS by tile (words
// tile << 32 | index, ceil(tile_bits / 8) = 2 passes at 1080p). Both sorts are stable and the index rides in the low half
// of the word, so the result is exactly the reference's stable sort of (tile << 32 | depth, index) pairs, for
// 8 + 2 * 16 B of traffic per pair instead of 8 + 5 * 16 + 4.
//
// Element counts live in device memory (FrameCtl): the host never learns them inside a frame; the pass kernel is persistent
// (as many CTAs as stay resident, tiles drawn by ticket) and a pass the frame does not need exits at once. The standalone API
// sorts caller-provided (u64, u32) pairs with the same kernels.
//
// What makes a pass fast here (DESIGN.md §4 has the measurements): ranking by returning shared-memory atomics (one
// instruction per key), bank-swizzled per-warp counters, look-back CHAINS — the input of a pass is cut into SORT_CHAINS segments
// whose digit counts the histogram kernel provides up front, so every segment runs its own shallow decoupled look-back —
// persistent CTAs that load their next tile while they write the current one out, and half-size tiles for the end of a pass.
#include "common.cuh"

#include <algorithm>
#include <type_traits>

namespace tpdcu {

constexpr uint32_t HIST_THREADS = 512;
constexpr uint32_t HIST_KPT = 8;
constexpr uint32_t LOOKBACK_VALUE_MASK = (1u << 30) - 1u;
#ifndef TPDCU_LOOKBACK_BATCH
#define TPDCU_LOOKBACK_BATCH 4
#endif
constexpr uint32_t LOOKBACK_BATCH = TPDCU_LOOKBACK_BATCH;
#ifndef TPDCU_HIST_CTAS_TILE
#define TPDCU_HIST_CTAS_TILE 2           // histogram CTAs per SM (each adds passes x segments x 256 counters to the global ones)
#endif
#ifndef TPDCU_HIST_CTAS_DEPTH
#define TPDCU_HIST_CTAS_DEPTH 2
#endif
#ifndef TPDCU_SORT_SWIZZLE
#define TPDCU_SORT_SWIZZLE 1             // bank swizzle of the per-warp digit counters (hist_slot)
#endif
#ifndef TPDCU_SORT_GRID_FACTOR
#define TPDCU_SORT_GRID_FACTOR 1         // CTAs launched per resident slot of the persistent pass kernel
#endif
#ifndef TPDCU_SORT_MINB_WORDS
#define TPDCU_SORT_MINB_WORDS 2
#endif

// What a sort launch works on (see SORT_KIND_* in common.cuh); derived on the device because n and the depth range are.
struct SortSpec {
    uint32_t n, bias, total_bits;
};
__device__ __forceinline__ SortSpec sort_spec(const FrameCtl* fr, uint32_t kind, uint32_t n_host, uint32_t capacity, uint32_t end_bit,
                                              uint32_t tile_bits) {
    SortSpec x;
    if (kind == SORT_KIND_PAIRS) {
        x.n = n_host; x.bias = 0; x.total_bits = end_bit;
    } else {
        const DepthSplit ds = depth_split(fr, tile_bits);
        if (kind == SORT_KIND_DEPTH) {
            x.n = min(fr->visible, capacity); x.bias = ds.bias; x.total_bits = ds.low_bits;
        } else {
            x.n = min(fr->pairs_total, capacity); x.bias = 0; x.total_bits = tile_bits + ds.extra;
        }
    }
    return x;
}
// the 64-bit quantity whose bits [0, total_bits) are sorted
template <bool WORDS>
__device__ __forceinline__ uint64_t sort_key(uint64_t k, uint32_t bias) {
    return WORDS ? (uint64_t)((uint32_t)(k >> 32) - bias) : k;
}

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint32_t* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// where bin d's counter lives inside a per-warp histogram row (see the digit packing in onesweep_kernel); an involution
__device__ __forceinline__ uint32_t hist_slot(uint32_t d) {
#if TPDCU_SORT_SWIZZLE
    return d ^ ((d >> 5) & 3u);
#else
    return d;
#endif
}
__device__ __forceinline__ uint32_t pass_mask(uint32_t pass, uint32_t total_bits) {
    const uint32_t left = total_bits > pass * SORT_RADIX_BITS ? total_bits - pass * SORT_RADIX_BITS : 0u;
    return left >= SORT_RADIX_BITS ? (SORT_BINS - 1u) : ((1u << left) - 1u);
}
__host__ __device__ __forceinline__ uint32_t passes_needed(uint32_t total_bits) { return (total_bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS; }

// ---------------------------------------------------------------------------------------------------
// up-front histogram of every digit (one read of the keys)
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void make_plan(const FrameCtl* frame, SortCtl* ctl, SortPlan* plan, uint32_t kind, uint32_t n_host,
                                          uint32_t capacity, uint32_t end_bit, uint32_t tile_bits);

// Tiles of a segment of `len` elements: full tiles of `tile` elements, except that the segment's last `half_last` full tiles'
// worth of elements is cut into half tiles (words sorts: the last tickets of a pass go to half tiles, so that the resident CTAs
// finish half a tile-life apart instead of a whole one). Shared by the plan and the pass kernel.
struct SegTiles { uint32_t full, total; };
__host__ __device__ __forceinline__ SegTiles seg_tile_counts(uint32_t len, uint32_t tile, uint32_t half_last) {
    const uint32_t whole = (len + tile - 1) / tile;
    SegTiles t;
    t.full = whole > half_last ? whole - half_last : 0u;
    const uint32_t rest = len - min(len, t.full * tile);
    t.total = t.full + (rest + tile / 2 - 1) / (tile / 2);
    return t;
}

__host__ __device__ __forceinline__ uint32_t chains_of(uint32_t kind) { return kind == SORT_KIND_PAIRS ? 1u : SORT_CHAINS; }
__host__ __device__ __forceinline__ uint32_t tile_of(uint32_t kind) { return kind == SORT_KIND_PAIRS ? SORT_TILE_PAIRS : SORT_TILE_WORDS; }
// length of a first-pass segment: whole tiles, SORT_CHAINS segments cover n
__device__ __forceinline__ uint32_t first_pass_segment(uint32_t n, uint32_t kind) {
    const uint32_t tile = tile_of(kind), chains = chains_of(kind);
    const uint32_t tiles = (n + tile - 1) / tile;
    return max((tiles + chains - 1) / chains, 1u) * tile;
}

// Every CTA counts a contiguous slice of the keys: for every pass p and every segment c of that pass's input (see
// SORT_CHAINS in common.cuh) the digit histogram of the keys that belong to it. Segment of a key: its position / segment
// length for the first pass, its previous digit / SORT_CHAIN_BINS for the later ones.
template <bool WORDS>
__global__ void __launch_bounds__(HIST_THREADS) sort_hist_kernel(const uint64_t* __restrict__ keys, const FrameCtl* frame, SortCtl* ctl,
                                                                  SortPlan* plan, uint32_t kind, uint32_t n_host, uint32_t capacity,
                                                                  uint32_t end_bit, uint32_t tile_bits) {
    __shared__ uint32_t h[SORT_CHAIN_ROWS * SORT_BINS];
    pdl_wait();
    pdl_release();
    const SortSpec sp = sort_spec(frame, kind, n_host, capacity, end_bit, tile_bits);
    const uint32_t n = sp.n, num_passes = passes_needed(sp.total_bits);
    const uint32_t chains = chains_of(kind), rows = num_passes * chains;
    const uint32_t seg0 = first_pass_segment(n, kind);
    for (uint32_t k = threadIdx.x; k < rows * SORT_BINS; k += HIST_THREADS) h[k] = 0;
    __syncthreads();
    // Each thread takes HIST_KPT CONSECUTIVE elements (two 32-byte loads): the pairs of one Gaussian are adjacent in the
    // unsorted buffer and usually share the upper tile bits, so run-length encoding the digits in registers removes most
    // shared-memory atomics and nearly all same-address conflicts. The loop is instruction-bound (ncu: issue slots 66 % busy,
    // 75 instructions per key before this form), so a words sort works on the 32-bit key, the per-pass constants are hoisted,
    // and all but the last chunk of the buffer take a path without per-element bounds.
    using Key = typename std::conditional<WORDS, uint32_t, uint64_t>::type;
    const uint32_t h_s = (uint32_t)__cvta_generic_to_shared(&h[0]);
    const uint32_t chunk = HIST_THREADS * HIST_KPT;
    const uint32_t chunks = (n + chunk - 1) / chunk, per_cta = (chunks + gridDim.x - 1) / gridDim.x;
    const uint32_t slice_end = min((uint64_t)n, (uint64_t)(blockIdx.x + 1) * per_cta * chunk);
    const uint32_t chain_shift = 31u - __clz(SORT_BINS / chains);   // previous digit -> segment: d >> chain_shift
    for (uint64_t base64 = (uint64_t)blockIdx.x * per_cta * chunk; base64 < slice_end; base64 += chunk) {
        const uint32_t first = (uint32_t)base64 + threadIdx.x * HIST_KPT;
        if (first >= n) continue;
        Key k[HIST_KPT];
        const uint32_t valid = min(HIST_KPT, n - first);
        if (valid == HIST_KPT) {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT / 4; ++j) {
                uint64_t v[4];
                ldg256(keys + first + 4 * j, v);
#pragma unroll
                for (uint32_t r = 0; r < 4; ++r) k[4 * j + r] = (Key)sort_key<WORDS>(v[r], sp.bias);
            }
        } else {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT; ++j) k[j] = (Key)sort_key<WORDS>(keys[min(first + j, n - 1u)], sp.bias);  // the tail repeats the last key: run-length absorbs it, `valid` below cuts it off
        }
        // first-pass segment of the thread's elements: changes at most once inside its HIST_KPT consecutive positions
        const uint32_t c_first = first / seg0;
        const uint32_t c_change = chains > 1 ? (c_first + 1u) * seg0 - first : HIST_KPT;   // elements j >= c_change sit in the next segment
        constexpr uint32_t MAX_P = WORDS ? SORT_WORD_PASSES : SORT_MAX_PASSES;
#pragma unroll
        for (uint32_t p = 0; p < MAX_P; ++p) {   // unrolled: shifts and the first-pass case become immediates / straight code
            if (p >= num_passes) break;
            const uint32_t shift = p * SORT_RADIX_BITS, mask = pass_mask(p, sp.total_bits);
            const uint32_t row_base = p * chains * SORT_BINS;
            // slot of element j: row (pass, segment) * 256 + swizzled digit. Segment: position for the first pass, the previous
            // digit for the others — for those, the bits above chain_shift of the previous digit, moved to bit 8.
            const uint32_t seg_shift = (p - 1u) * SORT_RADIX_BITS + chain_shift;
            const uint32_t seg_mask = p && chains > 1 ? (pass_mask(p - 1u, sp.total_bits) >> chain_shift) : 0u;
            auto slot = [&](uint32_t j) {
                const uint32_t d = (uint32_t)(k[j] >> shift) & mask;
                const uint32_t c = p == 0 ? (chains > 1 ? c_first + (j >= c_change ? 1u : 0u) : 0u) : ((uint32_t)(k[j] >> seg_shift) & seg_mask);
                return row_base + (c << SORT_RADIX_BITS) + hist_slot(d);   // bank swizzle: the tile sort's low digits sit at stride 4
            };
            if (valid == HIST_KPT && (kind == SORT_KIND_DEPTH || (p == 0 && kind == SORT_KIND_TILE))) {
                // depth bits of neighbours in index order are unrelated, and the pairs' low tile bits change with every element
                // (a Gaussian's tiles are emitted row by row): no runs to encode, one counting atomic per element is cheaper
                // than looking for them
#pragma unroll
                for (uint32_t j = 0; j < HIST_KPT; ++j)
                    asm volatile("red.shared.add.u32 [%0], 1;" : : "r"(h_s + 4u * slot(j)) : "memory");
            } else if (valid == HIST_KPT) {
                // run-length encode the eight slots; a run is flushed by ONE predicated shared-memory reduction (no branch:
                // the compiler's divergent-branch form of `if (changed) atomicAdd` cost 4 of 22 instructions per key and pass)
                uint32_t run_slot = slot(0), run = 1;
#pragma unroll
                for (uint32_t j = 1; j < HIST_KPT; ++j) {
                    const uint32_t sl = slot(j);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\t@p red.shared.add.u32 [%2], %3;\n\t}"
                                 : : "r"(sl), "r"(run_slot), "r"(h_s + 4u * run_slot), "r"(run) : "memory");
                    run = (sl != run_slot ? 0u : run) + 1u;
                    run_slot = sl;
                }
                asm volatile("red.shared.add.u32 [%0], %1;" : : "r"(h_s + 4u * run_slot), "r"(run) : "memory");
            } else {
                for (uint32_t j = 0; j < valid; ++j) atomicAdd(&h[slot(j)], 1u);   // only the buffer's last thread
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < rows * SORT_BINS; k += HIST_THREADS) {
        const uint32_t c = h[(k & ~(SORT_BINS - 1u)) | hist_slot(k & (SORT_BINS - 1u))];   // counter of bin k & 255 of row k >> 8
        if (c) atomicAdd(&ctl->chain_hist[0][0] + k, c);
    }
    // The last CTA to get here turns the histograms into the plan (exclusive digit offsets, segment offsets, passes to skip,
    // ping-pong schedule): one launch and one kernel boundary less per sort than a plan kernel of its own.
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->hist_done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        make_plan(frame, ctl, plan, kind, n_host, capacity, end_bit, tile_bits);
    }
}

// ---------------------------------------------------------------------------------------------------
// plan: exclusive digit offsets, segment offsets, identity-pass detection, ping-pong schedule
// ---------------------------------------------------------------------------------------------------

// Runs in ONE CTA of at least SORT_BINS threads; every thread of the CTA must call it (block barriers inside), the first
// SORT_BINS threads own one bin each. The histograms are read with gpu-scope loads: other CTAs produced them.
__device__ __forceinline__ void make_plan(const FrameCtl* frame, SortCtl* ctl, SortPlan* plan, uint32_t kind, uint32_t n_host,
                                          uint32_t capacity, uint32_t end_bit, uint32_t tile_bits) {
    __shared__ uint32_t s_warp[SORT_BINS / 32];
    __shared__ uint32_t s_skip[SORT_MAX_PASSES];
    const SortSpec sp = sort_spec(frame, kind, n_host, capacity, end_bit, tile_bits);
    const uint32_t n = sp.n, num_passes = passes_needed(sp.total_bits);
    const uint32_t chains = chains_of(kind);
    const uint32_t b = threadIdx.x, lane = b & 31u, warp = b >> 5;
    const bool owner = b < SORT_BINS;
    if (b < SORT_MAX_PASSES) s_skip[b] = 0;
    __syncthreads();
    // per bin: the segments' counts become the exclusive sum over the segments before them; their total is the bin's count.
    // All loads first (independent, one round trip), then the sums: the plan is a serial tail behind the histogram kernel.
    uint32_t seg_count[SORT_CHAIN_ROWS];
    const uint32_t rows = num_passes * chains;
#pragma unroll
    for (uint32_t r = 0; r < SORT_CHAIN_ROWS; ++r) seg_count[r] = (owner && r < rows) ? ld_relaxed_u32(&ctl->chain_hist[r][b]) : 0u;
    __shared__ uint32_t s_bin_start[SORT_MAX_PASSES][SORT_CHAINS + 1];   // exclusive offset of the first bin of every segment
    for (uint32_t p = 0; p < num_passes; ++p) {
        uint32_t c = 0;
#pragma unroll
        for (uint32_t r = 0; r < SORT_CHAIN_ROWS; ++r) {
            if (owner && r >= p * chains && r < (p + 1) * chains) {
                ctl->chain_hist[r][b] = c;
                c += seg_count[r];
            }
        }
        if (owner && c == n) s_skip[p] = 1;  // every key falls in this bin (also true for n == 0)
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        if (owner && lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (owner) {
            uint32_t wex = 0;
            for (uint32_t w = 0; w < warp; ++w) wex += s_warp[w];
            const uint32_t start = wex + incl - c;
            ctl->hist[p][b] = start;
            if (b % (SORT_BINS / chains) == 0) s_bin_start[p][b / (SORT_BINS / chains)] = start;
        }
        __syncthreads();
    }
    if (b == 0) {
        uint32_t sel = 0, run = 0;
        for (uint32_t p = 0; p < SORT_MAX_PASSES; ++p) {
            const uint32_t skip = p < num_passes ? s_skip[p] : 1u;
            plan->skip[p] = skip;
            plan->src_sel[p] = sel;
            if (!skip) { sel ^= 1u; ++run; }
        }
        plan->n = n;
        plan->num_passes = num_passes;
        plan->final_sel = sel;
        plan->passes_run = run;
        plan->bias = sp.bias;
        plan->total_bits = sp.total_bits;
        plan->tile_shift = kind == SORT_KIND_TILE ? depth_split(frame, tile_bits).extra : 0u;
        // segments of every pass's input and the descriptor rows of their tiles. A pass that follows an identity pass finds
        // every key in one previous bin: one segment covers [0, n) and the others are empty, which is what its keys' counts say.
        plan->chains = chains;
        const uint32_t half_last = kind == SORT_KIND_TILE ? SORT_HALF_LAST : 0u;   // the depth sort has two tiles per CTA: more, smaller tiles cost it more than its tail
        plan->half_last = half_last;
        const uint32_t tile = tile_of(kind), seg0 = first_pass_segment(n, kind);
        for (uint32_t p = 0; p < num_passes; ++p) {
            uint32_t rows = 0;
            for (uint32_t ch = 0; ch <= SORT_CHAINS; ++ch) {
                uint32_t start = n;
                if (ch < chains) start = p == 0 ? (uint32_t)min((uint64_t)n, (uint64_t)ch * seg0) : s_bin_start[p - 1][ch];
                plan->seg_start[p][ch] = start;
            }
            for (uint32_t ch = 0; ch <= SORT_CHAINS; ++ch) {
                plan->seg_tiles[p][ch] = rows;
                if (ch < SORT_CHAINS) rows += seg_tile_counts(plan->seg_start[p][ch + 1] - plan->seg_start[p][ch], tile, half_last).total;
            }
        }
    }
}

// stand-alone form: sorts with nothing to count (no elements, no key bits)
__global__ void __launch_bounds__(SORT_BINS) sort_plan_kernel(const FrameCtl* frame, SortCtl* ctl, SortPlan* plan, uint32_t kind,
                                                               uint32_t n_host, uint32_t capacity, uint32_t end_bit, uint32_t tile_bits) {
    make_plan(frame, ctl, plan, kind, n_host, capacity, end_bit, tile_bits);
}

// ---------------------------------------------------------------------------------------------------
// one onesweep pass
// ---------------------------------------------------------------------------------------------------

// Optional per-tile trace (profiles/micro/ws_trace.cu defines TPDCU_WS_TRACE and the buffer): globaltimer stamps of the phases.
#ifdef TPDCU_WS_TRACE
__device__ unsigned long long g_ws_trace[TPDCU_WS_TRACE][12];
__device__ __forceinline__ unsigned long long ws_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define WS_STAMP(part, slot) do { if ((part) < TPDCU_WS_TRACE) g_ws_trace[(part)][(slot)] = ws_now(); } while (0)
#define WS_NOTE(part, slot, v) do { if ((part) < TPDCU_WS_TRACE) g_ws_trace[(part)][(slot)] = (v); } while (0)
__device__ unsigned long long g_ws_lb[TPDCU_WS_TRACE][16];   // look-back: per batch, time the loads were issued / the batch was consumed
#define WS_LB_STAMP(part, slot) do { if ((part) < TPDCU_WS_TRACE && (slot) < 16) g_ws_lb[(part)][(slot)] = ws_now(); } while (0)
#else
#define WS_LB_STAMP(part, slot) do { } while (0)
#define WS_STAMP(part, slot) do { } while (0)
#define WS_NOTE(part, slot, v) do { } while (0)
#endif

// Sort modes. PAIRS: (u64 key, u32 value) in and out (standalone API). WORDS: single 64-bit words whose high half is the
// key and whose low half (the Gaussian index) is payload: a pass moves 16 B per element instead of 24 B and the value
// scatter through shared memory disappears.
enum : int { MODE_PAIRS = 0, MODE_WORDS = 1 };

template <int MODE>
struct OnesweepSmem {
    static constexpr bool WITH_VALS = MODE == MODE_PAIRS;
    static constexpr uint32_t TILE = MODE == MODE_PAIRS ? SORT_TILE_PAIRS : SORT_TILE_WORDS;
    uint64_t keys[TILE];
    alignas(16) uint32_t warp_hist[SORT_WARPS][SORT_BINS];  // zeroed with 16-byte stores
    alignas(16) uint32_t global_base[SORT_BINS];
    uint32_t scan[SORT_BINS / 32];
    uint32_t part;
    uint32_t vals[WITH_VALS ? TILE : 1];
};
static_assert(SORT_THREADS == SORT_BINS, "one thread per bin in the per-bin phases");

// A resident CTA works through tiles of TILE (or TILE / 2) elements, one ticket at a time. Per tile (block barriers in between):
//   digits + counting atomics (keys already in registers) | per bin: prefix over the warps, publish the aggregate, scan the
//   bins; draw the NEXT ticket | ranking atomics + scatter to smem | look-back per bin | the next tile's key loads are issued |
//   coalesced write-out (+ value scatter / write-out) | next tile.
// Between a CTA's end and its successor's first load lie a block launch, a ticket round trip and the plan loads — 2 us of a
// 10 us tile life with one tile per CTA (per-tile trace, profiles/micro/ws_trace.cu); the loop pays them once per CTA, and
// the next tile's keys travel while this tile's write-out runs. The next ticket is drawn at a fixed phase of every CTA's
// cycle: tickets are handed out in the order the tiles will really be started, which keeps the look-back from waiting on a
// tile whose CTA is still busy with the previous one.
template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS, MODE == MODE_WORDS ? TPDCU_SORT_MINB_WORDS : TPDCU_SORT_MINB)
onesweep_kernel(uint64_t* keys0, uint64_t* keys1, uint32_t* vals0, uint32_t* vals1, SortCtl* ctl,
                const SortPlan* __restrict__ plan, uint32_t* lookback_pass, uint32_t pass) {
    constexpr bool IN_PAIRS = MODE == MODE_PAIRS, OUT_PAIRS = MODE == MODE_PAIRS, WORDS = MODE == MODE_WORDS;
    using Smem = OnesweepSmem<MODE>;
    constexpr uint32_t SORT_KPT = MODE == MODE_PAIRS ? SORT_KPT_PAIRS : SORT_KPT_WORDS;  // shadows nothing: per-mode tile shape
    constexpr uint32_t HALF_KPT = SORT_KPT / 2;
    constexpr uint32_t SORT_TILE = Smem::TILE;
    static_assert(SORT_KPT % 8 == 0, "half tiles pack whole digit registers");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    pdl_wait();
    pdl_release();
    if (plan->skip[pass]) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) sm.part = atomicAdd(&ctl->ticket[pass], 1u);
    // the pass's segments (look-back chains): fetched while the first ticket is on its way
    uint32_t seg_start[SORT_CHAINS + 1], seg_tiles[SORT_CHAINS + 1];
#pragma unroll
    for (uint32_t c = 0; c <= SORT_CHAINS; ++c) {
        seg_start[c] = plan->seg_start[pass][c];
        seg_tiles[c] = plan->seg_tiles[pass][c];
    }
    const uint32_t chains = plan->chains, half_last = plan->half_last;
    const uint32_t total_tiles = seg_tiles[SORT_CHAINS];
    const uint32_t src = plan->src_sel[pass];
    const uint64_t* __restrict__ src_keys = src ? keys1 : keys0;
    const uint32_t* __restrict__ src_vals = src ? vals1 : vals0;
    uint64_t* __restrict__ dst_keys = src ? keys0 : keys1;
    uint32_t* __restrict__ dst_vals = src ? vals0 : vals1;
    const uint32_t bias = plan->bias, total_bits = plan->total_bits;
    const uint32_t shift = pass * SORT_RADIX_BITS, mask = pass_mask(pass, total_bits);
    auto digit_in = [&](uint64_t k) { return (uint32_t)(sort_key<WORDS>(k, bias) >> shift) & mask; };
    auto digit_out = digit_in;  // elements sit in shared memory in their input format
    const uint32_t my_slot = hist_slot(tid);
    const uint32_t hist_row = (uint32_t)__cvta_generic_to_shared(&sm.warp_hist[warp][0]);

    // Ticket -> (segment c, tile k of that segment), round-robin over the segments that still have tiles: consecutive tickets
    // go to different chains, so a chain's consecutive tiles start SORT_CHAINS tickets apart and its look-back is that much
    // shallower. F(k) = sum_c min(tiles_c, k) tickets precede round k; k is the last round that starts at or before the ticket.
    // A tile is `rows` 32-key rows per thread: SORT_KPT, or half of that for a segment's last tiles (seg_tile_counts).
    struct TileId { uint32_t chain, k, row0, base, end, rows; };   // segment, tile of it, its first descriptor row, elements [base, end) at most
    uint32_t min_tiles = 0xffffffffu;
#pragma unroll
    for (uint32_t c = 0; c < SORT_CHAINS; ++c) min_tiles = min(min_tiles, seg_tiles[c + 1] - seg_tiles[c]);
    auto locate = [&](uint32_t ticket) {
        TileId t{ 0u, ticket, 0u, 0u, 0u, SORT_KPT };
        if (chains > 1 && ticket < SORT_CHAINS * min_tiles) {   // every segment still has tiles: plain round-robin
            t.chain = ticket % SORT_CHAINS;
            t.k = ticket / SORT_CHAINS;
        } else if (chains > 1) {
            auto before_round = [&](uint32_t k) {
                uint32_t f = 0;
#pragma unroll
                for (uint32_t c = 0; c < SORT_CHAINS; ++c) f += min(seg_tiles[c + 1] - seg_tiles[c], k);
                return f;
            };
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (uint32_t c = 0; c < SORT_CHAINS; ++c) hi = max(hi, seg_tiles[c + 1] - seg_tiles[c]);
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (before_round(mid) <= ticket) lo = mid;
                else hi = mid - 1;
            }
            t.k = lo;
            uint32_t idx = ticket - before_round(lo);   // idx-th segment, in order, that has a tile k
            bool found = false;
#pragma unroll
            for (uint32_t c = 0; c < SORT_CHAINS; ++c) {
                const bool has = seg_tiles[c + 1] - seg_tiles[c] > t.k;
                if (!found && has) {
                    if (idx == 0) { t.chain = c; found = true; }
                    else --idx;
                }
            }
        }
        uint32_t begin = 0;
#pragma unroll
        for (uint32_t c = 0; c < SORT_CHAINS; ++c)
            if (c == t.chain) { begin = seg_start[c]; t.end = seg_start[c + 1]; t.row0 = seg_tiles[c]; }
        const uint32_t full = seg_tile_counts(t.end - begin, SORT_TILE, half_last).full;
        if (t.k < full) {
            t.base = begin + t.k * SORT_TILE;
        } else {
            t.base = begin + full * SORT_TILE + (t.k - full) * (SORT_TILE / 2);
            t.rows = HALF_KPT;
        }
        return t;
    };
    // key loads of a tile, warp-striped: item k of lane l is element warp * 32 * rows + 32 k + l of the tile; elements at or
    // beyond the segment's end are padding that sorts last
    uint64_t key[SORT_KPT];
    auto load_keys = [&](const TileId& t) {
        const uint32_t at = t.base + warp * (32u * t.rows) + lane;
        if (t.end - t.base >= t.rows * SORT_THREADS) {
#pragma unroll
            for (uint32_t k = 0; k < HALF_KPT; ++k) key[k] = src_keys[at + k * 32u];
            if (t.rows > HALF_KPT) {
#pragma unroll
                for (uint32_t k = HALF_KPT; k < SORT_KPT; ++k) key[k] = src_keys[at + k * 32u];
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < SORT_KPT; ++k) key[k] = (k < t.rows && (at + k * 32u) < t.end) ? src_keys[at + k * 32u] : ~0ull;
        }
    };

    __syncthreads();
    uint32_t part = sm.part;
    if (part >= total_tiles) return;
    TileId cur = locate(part);
    load_keys(cur);
    {   // this warp's counters start at zero (a warp only ever counts into its own row)
        uint4* z = reinterpret_cast<uint4*>(&sm.warp_hist[warp][0]);
        z[lane] = make_uint4(0, 0, 0, 0);
        z[lane + 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }

    for (;;) {
        const uint32_t rows = cur.rows;                   // 32-key rows per thread in this tile (uniform)
        const bool whole = rows > HALF_KPT;               // a full-size tile: the second half of every per-key loop runs
        const uint32_t tile_keys = rows * SORT_THREADS;
        const uint32_t tile_base = cur.base;
        const uint32_t n = cur.end;                       // elements at or beyond the segment's end are padding
        const uint32_t n_valid = min(tile_keys, cur.end - tile_base);
        const uint32_t row = cur.row0 + cur.k;            // this tile's descriptor row; its chain's rows are row0 .. row
        const uint32_t chain = cur.chain, k_tile = cur.k, row0 = cur.row0;
        const uint32_t warp_base = tile_base + warp * (32u * rows) + lane;
        const bool full = n_valid == tile_keys;
        if (tid == 0) WS_STAMP(part, 0);

        // ---- digits, computed once and packed four to a register; padding (only in a segment's last tile) goes to the top bin
        // What is packed is the digit's COUNTER SLOT, hist_slot(d) = d ^ ((d >> 5) & 3): the per-warp counters are only ever
        // indexed by it. The tile sort's low digit is (tile & 63) << 2 | top depth bits, and neighbours in depth order share
        // those depth bits, so the digits of a row sit at stride 4: unswizzled they fall on 8 of the 32 banks (ncu: 7.7
        // wavefronts per ranking atomic, 5.6 per counting atomic). XOR-ing bits 5-6 into bits 0-1 spreads a stride-4 run over
        // all banks and still maps 32 consecutive bins (a warp of the thread == bin phases) onto 32 distinct banks.
        uint32_t dpack[SORT_KPT / 4];
        auto pack_digits = [&](uint32_t q) {
            uint32_t w = 0;
#pragma unroll
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t k = q * 4 + r;
                const bool valid = full || (warp_base + k * 32u) < n;
                w |= hist_slot(valid ? digit_in(key[k]) : mask) << (8u * r);
            }
            dpack[q] = w;
        };
#pragma unroll
        for (uint32_t q = 0; q < HALF_KPT / 4; ++q) pack_digits(q);
        if (whole) {
#pragma unroll
            for (uint32_t q = HALF_KPT / 4; q < SORT_KPT / 4; ++q) pack_digits(q);
        }
        auto digit_at = [&](uint32_t k) { return (dpack[k >> 2] >> (8u * (k & 3u))) & 0xffu; };

        // ---- early counts: per-warp digit histograms ----------------------------------------------------
#pragma unroll
        for (uint32_t k = 0; k < HALF_KPT; ++k) atomicAdd(&sm.warp_hist[warp][digit_at(k)], 1u);
        if (whole) {
#pragma unroll
            for (uint32_t k = HALF_KPT; k < SORT_KPT; ++k) atomicAdd(&sm.warp_hist[warp][digit_at(k)], 1u);
        }

        // values (pair input; always whole tiles): issue the loads now, their latency hides behind the per-bin phases
        uint32_t val[IN_PAIRS ? SORT_KPT : 1];
        if (IN_PAIRS) {
            if (full) {
#pragma unroll
                for (uint32_t k = 0; k < SORT_KPT; ++k) val[k] = src_vals[warp_base + k * 32u];
            } else {
#pragma unroll
                for (uint32_t k = 0; k < SORT_KPT; ++k) val[k] = (warp_base + k * 32u) < n ? src_vals[warp_base + k * 32u] : 0u;
            }
        }
        __syncthreads();

        // ---- per-bin (thread == bin): exclusive prefix over warps, publish the tile aggregate, scan the bins -------------
        uint32_t* lb = lookback_pass + (size_t)row * SORT_BINS;
        uint32_t bin_count = 0;
#pragma unroll
        for (uint32_t w = 0; w < SORT_WARPS; ++w) {
            const uint32_t c = sm.warp_hist[w][my_slot];
            sm.warp_hist[w][my_slot] = bin_count;
            bin_count += c;
        }
        const uint32_t bin_count_valid = (tid == mask) ? bin_count - (tile_keys - n_valid) : bin_count;  // padding lives in the top bin
        if (tid == 0) { WS_STAMP(part, 1); WS_STAMP(part, 2); }
        st_relaxed_u32(lb + tid, ((k_tile == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) << 30) | bin_count_valid);
        uint32_t incl = bin_count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        if (lane == 31) sm.scan[warp] = incl;
        uint32_t bin_base = incl - bin_count;
        __syncthreads();
#pragma unroll
        for (uint32_t w = 0; w < SORT_BINS / 32; ++w)
            if (w < warp) bin_base += sm.scan[w];
#pragma unroll
        for (uint32_t w = 0; w < SORT_WARPS; ++w) sm.warp_hist[w][my_slot] += bin_base;
        uint32_t next_ticket = 0;
        if (tid == 0) next_ticket = atomicAdd(&ctl->ticket[pass], 1u);   // the next ticket travels while this tile is ranked
        __syncthreads();

        // ---- stable ranking: one returning shared-memory atomic per key; elements go straight to smem ---------------------
        // ATOMS with a destination register hands every lane the counter's value plus the number of LOWER lanes of the same
        // instruction that hit the same counter, i.e. the stable rank, and a warp's atomics execute in program order, so rows
        // stay ordered too (profiles/micro/atoms_rank.cu: 7-20 cycles per row and SM against 30-42 for eight ballots + bit logic;
        // lane order held on all 1.2e8 rows checked, and tpdcu_create re-checks it on the device: rank_selftest_kernel). Every
        // atomic is followed at once by its key's scatter (batches of 8 in flight: 69 us per tile-sort pass, 4: 67, 1: 66).
        uint32_t rank[OUT_PAIRS ? SORT_KPT : 1];
        auto rank_one = [&](uint32_t k) {
            uint32_t r;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(hist_row + 4u * digit_at(k)) : "memory");
            if (OUT_PAIRS) rank[k] = r;
            sm.keys[r] = key[k];
        };
#pragma unroll
        for (uint32_t k = 0; k < HALF_KPT; ++k) rank_one(k);
        if (whole) {
#pragma unroll
            for (uint32_t k = HALF_KPT; k < SORT_KPT; ++k) rank_one(k);
        }
        {   // this warp's counters are free again: zero them for the next tile
            __syncwarp();
            uint4* z = reinterpret_cast<uint4*>(&sm.warp_hist[warp][0]);
            z[lane] = make_uint4(0, 0, 0, 0);
            z[lane + 32] = make_uint4(0, 0, 0, 0);
        }
        if (tid == 0) { WS_STAMP(part, 3); WS_STAMP(part, 5); }

        // ---- decoupled look-back, one thread per bin -----------------------------------------------------
        uint32_t excl = 0, trace_rows = 0;
        (void)trace_rows;
        if (k_tile > 0) {
            // Tiles in flight publish their aggregate well before their prefix, so the walk back to the nearest PREFIX is
            // several tiles deep: read LOOKBACK_BATCH descriptors per round trip, consume them in order. The walk stays inside
            // the tile's own chain: rows row0 .. row - 1, and row0 (the segment's first tile) always carries a PREFIX.
            int look = (int)row - 1;
            const int first_row = (int)row0;
            bool done = false;
            while (!done) {
                uint32_t v[LOOKBACK_BATCH];
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j)
                    v[j] = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, first_row) * SORT_BINS + tid);
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j) {
                    if (!done) {
                        uint32_t x = v[j];
                        while ((x >> 30) == FLAG_INVALID) x = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, first_row) * SORT_BINS + tid);
                        excl += x & LOOKBACK_VALUE_MASK;
                        done = (x >> 30) == FLAG_PREFIX;  // reached at row0 at the latest, so look - j never goes below it unconsumed
                        ++trace_rows;
                    }
                }
                look -= (int)LOOKBACK_BATCH;
            }
            st_relaxed_u32(lb + tid, (FLAG_PREFIX << 30) | (excl + bin_count_valid));
        }
        // bin's output run + keys of the earlier segments in that bin + keys of this segment's earlier tiles - tile-local offset
        sm.global_base[tid] = ctl->hist[pass][tid] + ctl->chain_hist[pass * chains + chain][tid] + excl - bin_base;
        if (tid == 0) { WS_STAMP(part, 4); WS_STAMP(part, 6); WS_NOTE(part, 8, (unsigned long long)trace_rows); WS_NOTE(part, 9, 0ull); WS_NOTE(part, 10, (unsigned long long)chain); }
        if (tid == 0) sm.part = next_ticket;
        __syncthreads();

        // ---- the next tile: its ticket has arrived; its keys travel during the write-out ------------------------------------
        const uint32_t next_part = sm.part;
        const bool more = next_part < total_tiles;
        TileId nxt = cur;
        if (more) {
            nxt = locate(next_part);
            load_keys(nxt);   // this tile's keys sit in shared memory by now: the registers are free
        }

        // ---- write-out: position i of the locally sorted tile goes to global_base[digit] + i (contiguous per bin) ---------
        uint32_t pos[OUT_PAIRS ? SORT_KPT : 1];
        auto write_one = [&](uint32_t k, bool guarded) {
            const uint32_t i = tid + k * SORT_THREADS;
            if (!guarded || i < n_valid) {
                const uint64_t kk = sm.keys[i];
                const uint32_t p = sm.global_base[digit_out(kk)] + i;
                if (OUT_PAIRS) pos[k] = p;
                dst_keys[p] = kk;
            }
        };
        if (full) {  // every tile but a segment's last: no per-key bounds branch
#pragma unroll
            for (uint32_t k = 0; k < HALF_KPT; ++k) write_one(k, false);
            if (whole) {
#pragma unroll
                for (uint32_t k = HALF_KPT; k < SORT_KPT; ++k) write_one(k, false);
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < SORT_KPT; ++k) write_one(k, true);
        }
        if (tid == 0) WS_STAMP(part, 7);
        if (OUT_PAIRS) {
#pragma unroll
            for (uint32_t k = 0; k < SORT_KPT; ++k) sm.vals[rank[k]] = val[k];
            __syncthreads();
#pragma unroll
            for (uint32_t k = 0; k < SORT_KPT; ++k) {
                const uint32_t i = tid + k * SORT_THREADS;
                if (i < n_valid) dst_vals[pos[k]] = sm.vals[i];
            }
        }
        if (!more) break;
        // the next tile's scatter and look-back overwrite sm.keys / sm.vals / sm.global_base: they come after its two
        // barriers, which no thread passes before it has finished this write-out; sm.part is rewritten after them too
        cur = nxt;
        part = next_part;
    }
}

// introspection: sorted words -> the reference's (tile << 32 | depth bits, index) arrays
__global__ void sort_unpack_kernel(RasterLaunch a, uint64_t* out_keys, uint32_t* out_vals) {
    const uint64_t* w = a.plan->final_sel ? a.keys[1] : a.keys[0];
    const uint32_t n = a.plan->n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t x = w[i];
        const uint32_t g = (uint32_t)x;
        out_keys[i] = ((x >> (32u + a.plan->tile_shift)) << 32) | __float_as_uint(a.depth_radius[g].x);
        out_vals[i] = g;
    }
}

__global__ void sort_copy_result_kernel(const uint64_t* keys1, const uint32_t* vals1, uint64_t* out_keys, uint32_t* out_vals,
                                        const uint64_t* keys0, const uint32_t* vals0, const SortPlan* plan) {
    const uint64_t* sk = plan->final_sel ? keys1 : keys0;
    const uint32_t* sv = plan->final_sel ? vals1 : vals0;
    const uint32_t n = plan->n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        out_keys[i] = sk[i];
        out_vals[i] = sv[i];
    }
}

// upper bound of the tiles (= look-back descriptor rows, = CTAs) of a pass: every segment may end in a partial tile
uint32_t sort_parts(uint32_t capacity, uint32_t kind) {
    const uint32_t tile = tile_of(kind), chains = chains_of(kind);
    // + per segment: a partial last tile, and SORT_HALF_LAST full tiles turned into twice as many half tiles (+ 1 for rounding)
    return (capacity + tile - 1) / tile + chains * (kind == SORT_KIND_PAIRS ? 1u : SORT_HALF_LAST + 2u);
}
uint32_t sort_passes_for(uint32_t end_bit) { return passes_needed(end_bit); }

template <int MODE>
static cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(onesweep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(OnesweepSmem<MODE>));
}

// The ranking of onesweep_kernel relies on a property of the hardware that PTX does not promise: a warp's returning
// shared-memory atomics on one address hand out their values in lane order. This self-test (run once per context by
// tpdcu_create) ranks 2048 rows of adversarial digit patterns both ways — returning atomics, and match.any + popc, whose
// result is defined — and counts the rows that differ; a device that fails it is refused rather than sorted unstably.
__global__ void __launch_bounds__(256) rank_selftest_kernel(uint32_t* mismatches) {
    __shared__ uint32_t cnt[8][SORT_BINS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = lane; i < SORT_BINS; i += 32u) cnt[warp][i] = 0;
    __syncwarp();
    uint32_t bad = 0;
    for (uint32_t row = 0; row < 32u; ++row) {
        const uint32_t seed = ((blockIdx.x * 8u + warp) * 32u + row) * 32u + lane;
        uint32_t h = seed * 0x9E3779B1u; h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
        const uint32_t pattern = (blockIdx.x + row) & 3u;   // random 8-bit | stride 4 | two values | all equal
        const uint32_t d = pattern == 0 ? (h & 255u) : pattern == 1 ? ((h & 63u) << 2) : pattern == 2 ? (h & 1u) * 77u : (row * 7u) & 255u;
        const uint32_t before = cnt[warp][d];
        __syncwarp();
        uint32_t r;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(&cnt[warp][d])) : "memory");
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (r != before + __popc(peers & lanemask_lt())) ++bad;
        __syncwarp();
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t sort_rank_selftest(uint32_t* mismatches_host) {
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    e = cudaMemset(d, 0, sizeof(uint32_t));
    if (e == cudaSuccess) {
        rank_selftest_kernel<<<8, 256>>>(d);
        e = cudaMemcpy(mismatches_host, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    return e;
}

// opt in to > 48 KB of dynamic shared memory; called once per context, outside any stream capture
cudaError_t init_sort_attributes() {
    cudaError_t e = set_smem_attr<MODE_PAIRS>();
    if (e == cudaSuccess) e = set_smem_attr<MODE_WORDS>();
    return e;
}

cudaError_t launch_sort(const SortLaunch& a, uint32_t n_host, cudaStream_t s, cudaEvent_t ev_after_plan) {
    const bool words = a.kind != SORT_KIND_PAIRS;
    const uint32_t num_passes = passes_needed(a.end_bit);  // upper bound; the plan kernel marks the passes a frame does not need
    const uint32_t bound = words ? a.capacity : n_host;
    if (bound == 0 || num_passes == 0) {
        sort_plan_kernel<<<1, SORT_BINS, 0, s>>>(a.frame, a.ctl, a.plan, bound == 0 ? (uint32_t)SORT_KIND_PAIRS : a.kind, 0u, a.capacity, 0u, a.tile_bits);
        if (ev_after_plan) cudaEventRecord(ev_after_plan, s);
        return cudaGetLastError();
    }
    const uint32_t chunk = HIST_THREADS * HIST_KPT;
    uint32_t hist_grid = (bound + chunk - 1) / chunk;
    // every CTA adds up to passes x segments x 256 counters to the global histograms: few, fat CTAs
    const uint32_t hist_max = (uint32_t)a.sm_count * (a.kind == SORT_KIND_DEPTH ? TPDCU_HIST_CTAS_DEPTH : TPDCU_HIST_CTAS_TILE);
    if (hist_grid > hist_max) hist_grid = hist_max;
    if (words) pdl_launch(sort_hist_kernel<true>, hist_grid, HIST_THREADS, 0, s, (const uint64_t*)a.keys[0], (const FrameCtl*)a.frame, a.ctl, a.plan, a.kind, n_host, a.capacity, a.end_bit, a.tile_bits);
    else pdl_launch(sort_hist_kernel<false>, hist_grid, HIST_THREADS, 0, s, (const uint64_t*)a.keys[0], (const FrameCtl*)a.frame, a.ctl, a.plan, a.kind, n_host, a.capacity, a.end_bit, a.tile_bits);
    if (ev_after_plan) cudaEventRecord(ev_after_plan, s);
    const uint32_t parts = sort_parts(bound, a.kind);
    const uint32_t parts_cap = sort_parts(a.capacity, a.kind);
    // the pass kernel is persistent: as many CTAs as stay resident, each drawing tiles by ticket until none is left
    const uint32_t resident = (uint32_t)a.sm_count * (words ? TPDCU_SORT_MINB_WORDS : TPDCU_SORT_MINB) * TPDCU_SORT_GRID_FACTOR;
    for (uint32_t p = 0; p < num_passes; ++p) {
        uint32_t* lb = a.lookback + (size_t)p * parts_cap * SORT_BINS;
        if (words)
            pdl_launch(onesweep_kernel<MODE_WORDS>, std::min(parts, resident), SORT_THREADS, sizeof(OnesweepSmem<MODE_WORDS>), s, a.keys[0], a.keys[1], (uint32_t*)nullptr, (uint32_t*)nullptr, a.ctl, (const SortPlan*)a.plan, lb, p);
        else
            pdl_launch(onesweep_kernel<MODE_PAIRS>, std::min(parts, resident), SORT_THREADS, sizeof(OnesweepSmem<MODE_PAIRS>), s, a.keys[0], a.keys[1], a.vals[0], a.vals[1], a.ctl, (const SortPlan*)a.plan, lb, p);
    }
    return cudaGetLastError();
}

cudaError_t launch_sort_unpack(const RasterLaunch& a, uint64_t* out_keys, uint32_t* out_vals, int sm_count, cudaStream_t s) {
    sort_unpack_kernel<<<(uint32_t)sm_count * 8u, 256, 0, s>>>(a, out_keys, out_vals);
    return cudaGetLastError();
}

cudaError_t launch_sort_copy_result(const SortLaunch& a, uint64_t* out_keys, uint32_t* out_vals, uint32_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    uint32_t grid = (n + 255) / 256;
    if (grid > (uint32_t)a.sm_count * 8u) grid = (uint32_t)a.sm_count * 8u;
    sort_copy_result_kernel<<<grid, 256, 0, s>>>(a.keys[1], a.vals[1], out_keys, out_vals, a.keys[0], a.vals[0], a.plan);
    return cudaGetLastError();
}

}  // namespace tpdcu

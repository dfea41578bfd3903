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
// Element counts live in device memory (FrameCtl): the host never learns them inside a frame, grids are sized by the
// buffer capacities and surplus CTAs exit on their ticket. The standalone API sorts caller-provided (u64, u32) pairs.
#include "common.cuh"

#include <algorithm>

namespace tpdcu {

constexpr uint32_t HIST_THREADS = 512;
constexpr uint32_t HIST_KPT = 8;
constexpr uint32_t LOOKBACK_VALUE_MASK = (1u << 30) - 1u;
#ifndef TPDCU_LOOKBACK_BATCH
#define TPDCU_LOOKBACK_BATCH 8
#endif
constexpr uint32_t LOOKBACK_BATCH = TPDCU_LOOKBACK_BATCH;
#ifndef TPDCU_SORT_PREFETCH_TILES
#define TPDCU_SORT_PREFETCH_TILES 296
#endif
#ifndef TPDCU_SORT_RANK_ATOMS
#define TPDCU_SORT_RANK_ATOMS 1
#endif
#ifndef TPDCU_SORT_SWIZZLE
#define TPDCU_SORT_SWIZZLE 1
#endif
#ifndef TPDCU_SORT_LOOKBACK_VEC
#define TPDCU_SORT_LOOKBACK_VEC 0        // per-tile kernel: 64 threads walk four bins each with 16-byte descriptor loads
#endif
#ifndef TPDCU_SORT_LOOKBACK_EARLY
#define TPDCU_SORT_LOOKBACK_EARLY 0
#endif
#ifndef TPDCU_SORT_RANK_BATCH
#define TPDCU_SORT_RANK_BATCH 8
#endif
constexpr uint32_t SORT_RANK_BATCH = TPDCU_SORT_RANK_BATCH;
#ifndef TPDCU_SORT_MINB_WORDS
#define TPDCU_SORT_MINB_WORDS 2
#endif
constexpr uint32_t SORT_PREFETCH_TILES = TPDCU_SORT_PREFETCH_TILES;  // 148 SMs x 3 resident CTAs

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

template <bool WORDS>
__global__ void __launch_bounds__(HIST_THREADS) sort_hist_kernel(const uint64_t* __restrict__ keys, const FrameCtl* frame, SortCtl* ctl,
                                                                  SortPlan* plan, uint32_t kind, uint32_t n_host, uint32_t capacity,
                                                                  uint32_t end_bit, uint32_t tile_bits) {
    __shared__ uint32_t h[SORT_MAX_PASSES][SORT_BINS];
    const SortSpec sp = sort_spec(frame, kind, n_host, capacity, end_bit, tile_bits);
    const uint32_t n = sp.n, num_passes = passes_needed(sp.total_bits);
    for (uint32_t k = threadIdx.x; k < num_passes * SORT_BINS; k += HIST_THREADS) (&h[0][0])[k] = 0;
    __syncthreads();
    // Each thread takes HIST_KPT CONSECUTIVE elements (two 32-byte loads): the pairs of one Gaussian are adjacent in the
    // unsorted buffer and usually share the upper tile bits, so run-length encoding the digits in registers removes most
    // shared-memory atomics and nearly all same-address conflicts.
    const uint32_t chunk = HIST_THREADS * HIST_KPT;
    for (uint32_t base = blockIdx.x * chunk; base < n; base += gridDim.x * chunk) {
        const uint32_t first = base + threadIdx.x * HIST_KPT;
        uint64_t k[HIST_KPT];
        if (first + HIST_KPT <= n) {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT / 4; ++j) {
                uint64_t v[4];
                ldg256(keys + first + 4 * j, v);
                k[4 * j] = v[0]; k[4 * j + 1] = v[1]; k[4 * j + 2] = v[2]; k[4 * j + 3] = v[3];
            }
        } else {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT; ++j) k[j] = first + j < n ? keys[first + j] : 0ull;
        }
        const uint32_t valid = first >= n ? 0u : min(HIST_KPT, n - first);
        if (valid) {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT; ++j) k[j] = sort_key<WORDS>(k[j], sp.bias);
            for (uint32_t p = 0; p < num_passes; ++p) {
                const uint32_t shift = p * SORT_RADIX_BITS, mask = pass_mask(p, sp.total_bits);
                uint32_t run_digit = (uint32_t)(k[0] >> shift) & mask, run = 1;
#pragma unroll
                for (uint32_t j = 1; j < HIST_KPT; ++j) {
                    if (j < valid) {
                        const uint32_t d = (uint32_t)(k[j] >> shift) & mask;
                        if (d != run_digit) {
                            atomicAdd(&h[p][run_digit], run);
                            run_digit = d;
                            run = 0;
                        }
                        ++run;
                    }
                }
                atomicAdd(&h[p][run_digit], run);
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < num_passes * SORT_BINS; k += HIST_THREADS) {
        const uint32_t c = (&h[0][0])[k];
        if (c) atomicAdd(&ctl->hist[0][0] + k, c);
    }
    // The last CTA to get here turns the histograms into the plan (exclusive digit offsets, passes to skip, ping-pong
    // schedule): one launch and one kernel boundary less per sort than a plan kernel of its own.
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
// plan: exclusive digit offsets, identity-pass detection, ping-pong schedule
// ---------------------------------------------------------------------------------------------------

// Runs in ONE CTA of at least SORT_BINS threads; every thread of the CTA must call it (block barriers inside), the first
// SORT_BINS threads own one bin each. The histograms are read with gpu-scope loads: other CTAs produced them.
__device__ __forceinline__ void make_plan(const FrameCtl* frame, SortCtl* ctl, SortPlan* plan, uint32_t kind, uint32_t n_host,
                                          uint32_t capacity, uint32_t end_bit, uint32_t tile_bits) {
    __shared__ uint32_t s_warp[SORT_BINS / 32];
    __shared__ uint32_t s_skip[SORT_MAX_PASSES];
    const SortSpec sp = sort_spec(frame, kind, n_host, capacity, end_bit, tile_bits);
    const uint32_t n = sp.n, num_passes = passes_needed(sp.total_bits);
    const uint32_t b = threadIdx.x, lane = b & 31u, warp = b >> 5;
    const bool owner = b < SORT_BINS;
    if (b < SORT_MAX_PASSES) s_skip[b] = 0;
    __syncthreads();
    for (uint32_t p = 0; p < num_passes; ++p) {
        const uint32_t c = owner ? ld_relaxed_u32(&ctl->hist[p][b]) : 0u;
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
            ctl->hist[p][b] = wex + incl - c;
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
    alignas(16) uint32_t bin_count[SORT_BINS];   // per bin: valid keys of this tile / tile-local offset of the bin's run
    alignas(16) uint32_t bin_base[SORT_BINS];    // (handed from the thread == bin phase to the vectorised look-back)
    uint32_t scan[SORT_BINS / 32];
    uint32_t part;
    uint32_t vals[WITH_VALS ? TILE : 1];
};
static_assert(SORT_THREADS == SORT_BINS, "one thread per bin in the per-bin phases");

// One CTA = one tile of TILE elements. Phases (block barriers in between):
//   ticket + zero per-warp histograms | load keys, early counts | per-bin: warp prefix, publish aggregate, bin scan |
//   stable ranking (ballots) + scatter to smem | look-back per bin | coalesced write-out (+ value scatter / write-out)
template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS, MODE == MODE_WORDS ? TPDCU_SORT_MINB_WORDS : TPDCU_SORT_MINB)
onesweep_kernel(uint64_t* keys0, uint64_t* keys1, uint32_t* vals0, uint32_t* vals1, SortCtl* ctl,
                const SortPlan* __restrict__ plan, uint32_t* lookback_pass, uint32_t pass) {
    constexpr bool IN_PAIRS = MODE == MODE_PAIRS, OUT_PAIRS = MODE == MODE_PAIRS, WORDS = MODE == MODE_WORDS;
    using Smem = OnesweepSmem<MODE>;
    constexpr uint32_t SORT_KPT = MODE == MODE_PAIRS ? SORT_KPT_PAIRS : SORT_KPT_WORDS;  // shadows nothing: per-mode tile shape
    constexpr uint32_t SORT_TILE = Smem::TILE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    if (plan->skip[pass]) return;
    const uint32_t n = plan->n;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) sm.part = atomicAdd(&ctl->ticket[pass], 1u);
    {
        uint4* z = reinterpret_cast<uint4*>(&sm.warp_hist[0][0]);
#pragma unroll
        for (uint32_t k = 0; k < SORT_WARPS * SORT_BINS / 4 / SORT_THREADS; ++k) z[tid + k * SORT_THREADS] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t part = sm.part;
    const uint64_t tile_base64 = (uint64_t)part * SORT_TILE;
    if (tile_base64 >= n) return;
    if (tid == 0) WS_STAMP(part, 0);
    const uint32_t tile_base = (uint32_t)tile_base64;
    const uint32_t n_valid = min(SORT_TILE, n - tile_base);

    const uint32_t src = plan->src_sel[pass];
    const uint64_t* __restrict__ src_keys = src ? keys1 : keys0;
    const uint32_t* __restrict__ src_vals = src ? vals1 : vals0;
    uint64_t* __restrict__ dst_keys = src ? keys0 : keys1;
    uint32_t* __restrict__ dst_vals = src ? vals0 : vals1;

    // Tiles run in ticket order; the tile SORT_PREFETCH_TILES tickets ahead starts roughly when this one retires. One TMA
    // bulk prefetch per array pulls it into L2 now, so that its loads are L2 hits then.
    if (tid == 0) {
        const uint64_t ahead = (uint64_t)(part + SORT_PREFETCH_TILES) * SORT_TILE;
        if (ahead + SORT_TILE <= n) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_keys + ahead), "r"((uint32_t)(SORT_TILE * sizeof(uint64_t))) : "memory");
            if (IN_PAIRS)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_vals + ahead), "r"((uint32_t)(SORT_TILE * sizeof(uint32_t))) : "memory");
        }
    }
    const uint32_t bias = plan->bias, total_bits = plan->total_bits;
    const uint32_t shift = pass * SORT_RADIX_BITS, mask = pass_mask(pass, total_bits);
    auto digit_in = [&](uint64_t k) { return (uint32_t)(sort_key<WORDS>(k, bias) >> shift) & mask; };
    auto digit_out = digit_in;  // elements sit in shared memory in their input format

    // ---- load (warp-striped: item k of lane l sits at warp_base + 32k + l) --------------------------
    uint64_t key[SORT_KPT];
    const uint32_t warp_base = tile_base + warp * (32u * SORT_KPT) + lane;
    const bool full = n_valid == SORT_TILE;
    if (full) {
#pragma unroll
        for (uint32_t k = 0; k < SORT_KPT; ++k) key[k] = src_keys[warp_base + k * 32u];
    } else {
#pragma unroll
        for (uint32_t k = 0; k < SORT_KPT; ++k) key[k] = (warp_base + k * 32u) < n ? src_keys[warp_base + k * 32u] : ~0ull;  // padding sorts last
    }

    // ---- digits, computed once and packed four to a register; padding (only in the last tile) goes to the top bin -------
    // What is packed is the digit's COUNTER SLOT, hist_slot(d) = d ^ ((d >> 5) & 3): the per-warp counters are only ever
    // indexed by it. The tile sort's low digit is (tile & 63) << 2 | top depth bits, and neighbours in depth order share
    // those depth bits, so the digits of a row sit at stride 4: unswizzled they fall on 8 of the 32 banks (ncu: 7.7
    // wavefronts per ranking atomic, 5.6 per counting atomic). XOR-ing bits 5-6 into bits 0-1 spreads a stride-4 run over
    // all banks and still maps 32 consecutive bins (a warp of the thread == bin phases) onto 32 distinct banks.
    uint32_t dpack[SORT_KPT / 4];
#pragma unroll
    for (uint32_t q = 0; q < SORT_KPT / 4; ++q) {
        uint32_t w = 0;
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
            const uint32_t k = q * 4 + r;
            const bool valid = full || (warp_base + k * 32u) < n;
            w |= hist_slot(valid ? digit_in(key[k]) : mask) << (8u * r);
        }
        dpack[q] = w;
    }
    auto digit_at = [&](uint32_t k) { return (dpack[k >> 2] >> (8u * (k & 3u))) & 0xffu; };

    // ---- early counts: per-warp digit histograms ----------------------------------------------------
#pragma unroll
    for (uint32_t k = 0; k < SORT_KPT; ++k) atomicAdd(&sm.warp_hist[warp][digit_at(k)], 1u);

    // values (pair input): issue the loads now, their latency hides behind the per-bin phases
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
    uint32_t* lb = lookback_pass + (size_t)part * SORT_BINS;
    uint32_t bin_count = 0;
    const uint32_t my_slot = hist_slot(tid);
#pragma unroll
    for (uint32_t w = 0; w < SORT_WARPS; ++w) {
        const uint32_t c = sm.warp_hist[w][my_slot];
        sm.warp_hist[w][my_slot] = bin_count;
        bin_count += c;
    }
    const uint32_t bin_count_valid = (tid == mask) ? bin_count - (SORT_TILE - n_valid) : bin_count;  // padding lives in the top bin
    if (tid == 0) { WS_STAMP(part, 1); WS_STAMP(part, 2); }
    st_relaxed_u32(lb + tid, ((part == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) << 30) | bin_count_valid);
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
#if TPDCU_SORT_LOOKBACK_VEC
    sm.bin_count[tid] = bin_count_valid;
    sm.bin_base[tid] = bin_base;
#endif
    __syncthreads();

    // ---- stable ranking: peers with the same digit inside a 32-key row, rows in order; elements go straight to smem ---
#if TPDCU_SORT_LOOKBACK_EARLY
    // The first round trip of the look-back is issued now and consumed after the ranking: the predecessors published their
    // aggregates about when this tile did, and the ranking below does not depend on them.
    uint32_t v_early[LOOKBACK_BATCH];
#pragma unroll
    for (int j = 0; j < (int)LOOKBACK_BATCH; ++j)
        v_early[j] = part > 0 ? ld_relaxed_u32(lookback_pass + (size_t)max((int)part - 1 - j, 0) * SORT_BINS + tid) : 0u;
#endif
    uint32_t rank[OUT_PAIRS ? SORT_KPT : 1];
#if TPDCU_SORT_RANK_ATOMS
    // One shared-memory atomic per key: ATOMS.POPC.INC with a destination register hands every lane the counter's value
    // plus the number of LOWER lanes of the same instruction that hit the same counter, i.e. the stable rank, and a warp's
    // atomics execute in program order, so rows stay ordered too (profiles/micro/atoms_rank.cu: 7-20 cycles per row and SM
    // against 30-42 for eight ballots + bit logic; lane order held on all 1.2e8 rows checked, and tpdcu_create re-checks it
    // on the device it runs on). Batches: the atomics of a batch are in flight together, then their keys are scattered.
    {
        const uint32_t hist_base = (uint32_t)__cvta_generic_to_shared(&sm.warp_hist[warp][0]);
#pragma unroll
        for (uint32_t k0 = 0; k0 < SORT_KPT; k0 += SORT_RANK_BATCH) {
            uint32_t r[SORT_RANK_BATCH];
#pragma unroll
            for (uint32_t j = 0; j < SORT_RANK_BATCH; ++j)
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r[j]) : "r"(hist_base + 4u * digit_at(k0 + j)) : "memory");
#pragma unroll
            for (uint32_t j = 0; j < SORT_RANK_BATCH; ++j) {
                if (OUT_PAIRS) rank[k0 + j] = r[j];
                sm.keys[r[j]] = key[k0 + j];
            }
        }
    }
#else
#pragma unroll
    for (uint32_t k = 0; k < SORT_KPT; ++k) {
        const uint32_t d = digit_at(k);
        // peers = lanes of this row holding the same digit: eight ballots + bit logic on the ALU pipe (match.any executes on
        // the address-divergence unit: ADU 39 %, LSU 57 % — the ballot form was 20 % faster). Kept as the variant for a device
        // whose shared-memory atomics do not return lane-ordered values.
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (uint32_t bit = 0; bit < SORT_RADIX_BITS; ++bit)
            asm("{\n\t.reg .pred p;\n\t.reg .b32 b;\n\t"
                "and.b32 b, %1, %2;\n\tsetp.ne.u32 p, b, 0;\n\t"
                "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
                "@!p not.b32 b, b;\n\tand.b32 %0, %0, b;\n\t}"
                : "+r"(peers) : "r"(d), "r"(1u << bit));
        const uint32_t lower = __popc(peers & lanemask_lt());
        const uint32_t base = sm.warp_hist[warp][d];
        __syncwarp();
        if (lower == 0) sm.warp_hist[warp][d] = base + __popc(peers);
        __syncwarp();
        const uint32_t r = base + lower;
        if (OUT_PAIRS) rank[k] = r;
        sm.keys[r] = key[k];
    }
#endif

#if TPDCU_SORT_LOOKBACK_VEC
    // ---- decoupled look-back, four bins per thread: 64 threads, one 16-byte descriptor load per row and thread -------------
    // The walk costs what its loads cost (strong loads served by L2, ~0.9 us per batch of eight rows under load, a third of a
    // tile's life): a quarter of the requests for the same rows. Words are summed with their flags; the flags' share
    // (rows x flag << 30, modulo 2^32 like the sums) is taken out at the end. A row is consumed once its four words carry the
    // same valid flag (a tile publishes its descriptors together), else it is fetched again.
    if (tid == 0) { WS_STAMP(part, 3); WS_STAMP(part, 5); }
    if (tid < SORT_BINS / 4) {
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, trace_rows = 0;
        (void)trace_rows;
        if (part > 0) {
            int look = (int)part - 1;
            uint32_t agg_rows = 0;
            bool done = false;
            while (!done) {
                uint4 v[LOOKBACK_BATCH];
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j) v[j] = ld_relaxed_v4(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid * 4u);
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j) {
                    if (!done) {
                        uint4 x = v[j];
                        uint32_t all_and;
                        for (;;) {
                            all_and = x.x & x.y & x.z & x.w;
                            const uint32_t all_or = x.x | x.y | x.z | x.w;
                            if (((all_and ^ all_or) >> 30) == 0u && (all_and >> 30) != FLAG_INVALID) break;
                            x = ld_relaxed_v4(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid * 4u);
                        }
                        e0 += x.x; e1 += x.y; e2 += x.z; e3 += x.w;
                        if ((all_and >> 30) == FLAG_PREFIX) done = true;  // tile 0 always carries a PREFIX
                        else ++agg_rows;
                        ++trace_rows;
                    }
                }
                look -= (int)LOOKBACK_BATCH;
            }
            const uint32_t flags = (agg_rows * FLAG_AGGREGATE + FLAG_PREFIX) << 30;
            e0 -= flags; e1 -= flags; e2 -= flags; e3 -= flags;
            const uint4 c = *reinterpret_cast<const uint4*>(&sm.bin_count[tid * 4u]);
            st_relaxed_v4(lb + tid * 4u, make_uint4((FLAG_PREFIX << 30) | (e0 + c.x), (FLAG_PREFIX << 30) | (e1 + c.y),
                                                    (FLAG_PREFIX << 30) | (e2 + c.z), (FLAG_PREFIX << 30) | (e3 + c.w)));
        }
        const uint4 bb = *reinterpret_cast<const uint4*>(&sm.bin_base[tid * 4u]);
        const uint4 h = *reinterpret_cast<const uint4*>(&ctl->hist[pass][tid * 4u]);
        *reinterpret_cast<uint4*>(&sm.global_base[tid * 4u]) = make_uint4(h.x + e0 - bb.x, h.y + e1 - bb.y, h.z + e2 - bb.z, h.w + e3 - bb.w);
        if (tid == 0) { WS_STAMP(part, 4); WS_STAMP(part, 6); WS_NOTE(part, 8, (unsigned long long)trace_rows); WS_NOTE(part, 9, 0ull); WS_NOTE(part, 10, (unsigned long long)blockIdx.x); }
    }
#else
    // ---- decoupled look-back, one thread per bin -----------------------------------------------------
    {
        if (tid == 0) { WS_STAMP(part, 3); WS_STAMP(part, 5); }
        uint32_t excl = 0, trace_rows = 0;
        (void)trace_rows;
        if (part > 0) {
            // Tiles in flight publish their aggregate well before their prefix, so the walk back to the nearest PREFIX is
            // several tiles deep: read LOOKBACK_BATCH descriptors per round trip, consume them in order.
            int look = (int)part - 1;
            bool done = false;
#if TPDCU_SORT_LOOKBACK_EARLY
            bool first = true;
#endif
            while (!done) {
                uint32_t v[LOOKBACK_BATCH];
#if TPDCU_SORT_LOOKBACK_EARLY
                if (first) {
#pragma unroll
                    for (int j = 0; j < (int)LOOKBACK_BATCH; ++j) v[j] = v_early[j];
                    first = false;
                } else
#endif
                {
#pragma unroll
                    for (int j = 0; j < (int)LOOKBACK_BATCH; ++j)
                        v[j] = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid);
                }
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j) {
                    if (!done) {
                        uint32_t x = v[j];
                        while ((x >> 30) == FLAG_INVALID) x = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid);
                        excl += x & LOOKBACK_VALUE_MASK;
                        done = (x >> 30) == FLAG_PREFIX;  // tile 0 always carries a PREFIX, so look - j never goes below 0 unconsumed
                        ++trace_rows;
                    }
                }
                look -= (int)LOOKBACK_BATCH;
            }
            st_relaxed_u32(lb + tid, (FLAG_PREFIX << 30) | (excl + bin_count_valid));
        }
        sm.global_base[tid] = ctl->hist[pass][tid] + excl - bin_base;
        if (tid == 0) { WS_STAMP(part, 4); WS_STAMP(part, 6); WS_NOTE(part, 8, (unsigned long long)trace_rows); WS_NOTE(part, 9, 0ull); WS_NOTE(part, 10, (unsigned long long)blockIdx.x); }
    }
#endif
    __syncthreads();

    // ---- write-out: position i of the locally sorted tile goes to global_base[digit] + i (contiguous per bin) ---------
    uint32_t pos[OUT_PAIRS ? SORT_KPT : 1];
    if (full) {  // every tile but the last: no per-key bounds branch
#pragma unroll
        for (uint32_t k = 0; k < SORT_KPT; ++k) {
            const uint32_t i = tid + k * SORT_THREADS;
            const uint64_t kk = sm.keys[i];
            const uint32_t p = sm.global_base[digit_out(kk)] + i;
            if (OUT_PAIRS) pos[k] = p;
            dst_keys[p] = kk;
        }
    } else {
#pragma unroll
        for (uint32_t k = 0; k < SORT_KPT; ++k) {
            const uint32_t i = tid + k * SORT_THREADS;
            if (i < n_valid) {
                const uint64_t kk = sm.keys[i];
                const uint32_t p = sm.global_base[digit_out(kk)] + i;
                if (OUT_PAIRS) pos[k] = p;
                dst_keys[p] = kk;
            }
        }
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
}

#if TPDCU_SORT_WS
// ---------------------------------------------------------------------------------------------------
// one onesweep pass over single words, warp-specialised and persistent (the frame's two sorts)
// ---------------------------------------------------------------------------------------------------
//
// The per-tile kernel above spends two thirds of its warp time waiting: for its ticket, for its keys (a third of all stall
// samples) and in the look-back (a fifth), with two CTAs per SM to cover for each other (ncu, profiles/r2_onesweep_*.txt).
// Here ONE CTA per SM stays resident and runs two independent consumer groups of eight warps, each with a helper warp:
//   helper   draws the group's next ticket and streams that tile into the group's spare key buffer with TMA bulk copies
//            (cp.async.bulk -> mbarrier) while the group still works on the current tile; then resolves the current tile's
//            decoupled look-back — eight bins per lane, 128-bit descriptor loads — while the group ranks its keys.
//   group    keys shared -> registers, counting atomics, per-bin prefix + aggregate publication, ranking atomics + scatter
//            into the buffer the keys came from, (wait for the helper's bases), coalesced write-out.
// Tickets are drawn when a buffer frees up, not in lock-step, so tiles stay staggered across the SMs and the look-back
// stays shallow (persistent CTAs with a static tile assignment walked 67 descriptors deep).
constexpr uint32_t WS_GROUPS = 2;
constexpr uint32_t WS_GROUP_THREADS = SORT_THREADS;                 // one thread per bin in the per-bin phases
constexpr uint32_t WS_GROUP_WARPS = WS_GROUP_THREADS / 32;
constexpr uint32_t WS_THREADS = WS_GROUPS * (WS_GROUP_THREADS + 32);  // consumer warps first, then one helper warp per group
constexpr uint32_t WS_KPT = SORT_KPT_WORDS;
constexpr uint32_t WS_TILE = SORT_TILE_WORDS;
constexpr uint32_t WS_END = 0xffffffffu;
#ifndef TPDCU_WS_LOOKBACK_BATCH
#define TPDCU_WS_LOOKBACK_BATCH 8
#endif
constexpr int WS_LB_BATCH = TPDCU_WS_LOOKBACK_BATCH;
#ifndef TPDCU_WS_HELPER_LOOKBACK
#define TPDCU_WS_HELPER_LOOKBACK 0       // 1: the helper warp resolves the look-back (eight bins per lane); 0: the group does, one thread per bin
#endif
#ifndef TPDCU_WS_LOOKBACK_EARLY
#define TPDCU_WS_LOOKBACK_EARLY 1        // group look-back: first batch of descriptor loads issued before the ranking
#endif
#ifndef TPDCU_WS_PREFETCH_TILES
#define TPDCU_WS_PREFETCH_TILES 296
#endif
constexpr uint32_t WS_PREFETCH_TILES = TPDCU_WS_PREFETCH_TILES;      // 148 SMs x 2 groups
constexpr uint32_t WS_TMA_CHUNKS = 8;                                // bulk copies per tile (one per helper lane)
static_assert((WS_TILE * sizeof(uint64_t)) % (WS_TMA_CHUNKS * 16) == 0, "TMA chunks are multiples of 16 bytes");

struct WsGroupSmem {
    alignas(128) uint64_t keys[2][WS_TILE];                          // raw tile -> locally sorted tile, double-buffered
    alignas(16) uint32_t warp_hist[WS_GROUP_WARPS][SORT_BINS];
    alignas(16) uint32_t global_base[SORT_BINS];                     // helper -> group: where bin b's run of this tile starts, minus its tile-local offset
    alignas(16) uint32_t bin_count[SORT_BINS];                       // group -> helper: valid keys of this tile per bin
    alignas(16) uint32_t bin_base[SORT_BINS];                        // group -> helper: tile-local offset of the bin's run
    uint32_t scan[SORT_BINS / 32];
    uint32_t part[2];                                                // ticket of the tile in keys[i & 1], WS_END when there is none
    alignas(8) uint64_t raw_full[2], raw_empty[2], agg_ready, lb_done;
};
struct WsSmem { WsGroupSmem g[WS_GROUPS]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ready = 0;
    while (!ready)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ready) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void group_sync(uint32_t group) {  // named barrier of one consumer group (barrier 0 is __syncthreads)
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1u), "r"(WS_GROUP_THREADS) : "memory");
}
// every lane of the warp holds the same flag value
__device__ __forceinline__ bool __match_all_flags(uint32_t flag) {
    return __all_sync(0xffffffffu, flag == __shfl_sync(0xffffffffu, flag, 0));
}

__global__ void __launch_bounds__(WS_THREADS, 1)
onesweep_ws_kernel(uint64_t* keys0, uint64_t* keys1, SortCtl* ctl, const SortPlan* __restrict__ plan, uint32_t* lookback_pass, uint32_t pass) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WsSmem& smem = *reinterpret_cast<WsSmem*>(smem_raw);
    if (plan->skip[pass]) return;
    const uint32_t n = plan->n;
    const uint32_t src = plan->src_sel[pass];
    const uint64_t* __restrict__ src_keys = src ? keys1 : keys0;
    uint64_t* __restrict__ dst_keys = src ? keys0 : keys1;
    const uint32_t bias = plan->bias, total_bits = plan->total_bits;
    const uint32_t shift = pass * SORT_RADIX_BITS, mask = pass_mask(pass, total_bits);
    auto digit_of = [&](uint64_t k) { return (uint32_t)(sort_key<true>(k, bias) >> shift) & mask; };

    const uint32_t warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const bool helper = warp_id >= WS_GROUPS * WS_GROUP_WARPS;
    const uint32_t group = helper ? warp_id - WS_GROUPS * WS_GROUP_WARPS : warp_id / WS_GROUP_WARPS;
    WsGroupSmem& sm = smem.g[group];
    if (threadIdx.x == 0) {
        for (uint32_t g = 0; g < WS_GROUPS; ++g) {
            WsGroupSmem& x = smem.g[g];
            mbar_init(&x.raw_full[0], 1); mbar_init(&x.raw_full[1], 1);
            mbar_init(&x.raw_empty[0], WS_GROUP_WARPS); mbar_init(&x.raw_empty[1], WS_GROUP_WARPS);
            mbar_init(&x.agg_ready, WS_GROUP_WARPS); mbar_init(&x.lb_done, TPDCU_WS_HELPER_LOOKBACK ? 1 : WS_GROUP_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (helper) {
        // ---------------- helper warp: ticket + TMA of the next tile, look-back of the tile in flight ----------------
        // The next ticket is drawn when the look-back of the current tile has completed. Look-backs complete in (roughly)
        // ticket order, so tickets are handed out in the order the groups will really start their tiles and nobody spins
        // on the aggregate of a tile whose group is still busy with another one (tickets drawn a whole tile ahead were
        // uncorrelated with that order: every generation of tiles waited for its slowest member, 240 us per pass).
        auto draw = [&]() {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&ctl->ticket[pass], 1u);
            return __shfl_sync(0xffffffffu, t, 0);
        };
        uint32_t part = draw();
        for (uint32_t it = 0;; ++it) {
            const uint32_t b = it & 1u;
            // keys[b] last held tile it - 2: the group releases the buffer when it has written that tile out
            if (it >= 2) mbar_wait(&sm.raw_empty[b], ((it - 2) >> 1) & 1u);
            const bool more = (uint64_t)part * WS_TILE < n;
            if (lane == 0 && more) { WS_STAMP(part, 0); WS_NOTE(part, 10, (unsigned long long)(blockIdx.x * WS_GROUPS + group)); }
            if (lane == 0) {
                sm.part[b] = more ? part : WS_END;
                if (more) mbar_expect_tx(&sm.raw_full[b], (uint32_t)(WS_TILE * sizeof(uint64_t)));
                else mbar_arrive(&sm.raw_full[b]);
            }
            __syncwarp();
            if (!more) break;
            if (lane < WS_TMA_CHUNKS) {
                // the key buffers are allocated in whole tiles: the last tile is copied whole, its tail is masked by the group
                constexpr uint32_t chunk = (uint32_t)(WS_TILE * sizeof(uint64_t)) / WS_TMA_CHUNKS;
                const unsigned char* g = reinterpret_cast<const unsigned char*>(src_keys + (size_t)part * WS_TILE) + lane * chunk;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(reinterpret_cast<unsigned char*>(sm.keys[b]) + lane * chunk)), "l"(g), "r"(chunk), "r"(smem_u32(&sm.raw_full[b])) : "memory");
            } else if (lane == WS_TMA_CHUNKS) {
                // whoever draws the ticket one round of groups ahead finds its tile in L2
                const uint64_t ahead = (uint64_t)(part + WS_PREFETCH_TILES) * WS_TILE;
                if (ahead + WS_TILE <= n)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_keys + ahead), "r"((uint32_t)(WS_TILE * sizeof(uint64_t))) : "memory");
            }
#if TPDCU_WS_HELPER_LOOKBACK
            // look-back of this tile while the group ranks it
            mbar_wait(&sm.agg_ready, it & 1u);
            if (lane == 0) WS_STAMP(part, 3);
            // This lane owns bins 4 lane .. 4 lane + 3 and 128 + 4 lane .. 128 + 4 lane + 3: two 16-byte loads per descriptor row,
            // each a contiguous 512 bytes across the warp.
            uint32_t excl[8], trace_rows = 0, trace_retries = 0;
            (void)trace_rows; (void)trace_retries;
#pragma unroll
            for (int q = 0; q < 8; ++q) excl[q] = 0;
            if (part > 0) {
                // One warp walks all 256 bins, so the walk is kept warp-uniform and cheap: a tile's descriptors carry the same
                // flag in every bin (they are published together), a row is only consumed once it is uniform (re-fetched while
                // it has missing or mixed entries), the raw words are summed with their flags and the flags' contribution
                // (rows x flag << 30, modulo 2^32 like the sums) is taken out at the end.
                int look = (int)part - 1;
                uint32_t agg_rows = 0, batch_no = 0;
                (void)batch_no;
                bool done = false;
                while (!done) {
                    if (lane == 0) WS_LB_STAMP(part, 2 * batch_no);
                    uint4 v[WS_LB_BATCH][2];
#pragma unroll
                    for (int j = 0; j < WS_LB_BATCH; ++j) {
                        const uint32_t* d = lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + lane * 4u;
                        v[j][0] = ld_relaxed_v4(d);
                        v[j][1] = ld_relaxed_v4(d + 128);
                    }
#pragma unroll
                    for (int j = 0; j < WS_LB_BATCH; ++j) {
                        if (!done) {
                            const uint32_t* d = lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + lane * 4u;
                            uint32_t all_and, all_or;
                            for (;;) {
                                all_and = v[j][0].x & v[j][0].y & v[j][0].z & v[j][0].w & v[j][1].x & v[j][1].y & v[j][1].z & v[j][1].w;
                                all_or = v[j][0].x | v[j][0].y | v[j][0].z | v[j][0].w | v[j][1].x | v[j][1].y | v[j][1].z | v[j][1].w;
                                // uniform row: the two flag bits agree in every word of every lane, and are not INVALID
                                const bool uniform = ((all_and ^ all_or) >> 30) == 0u && (all_and >> 30) != FLAG_INVALID;
                                if (__all_sync(0xffffffffu, uniform) && __match_all_flags(all_and >> 30)) break;
                                v[j][0] = ld_relaxed_v4(d);
                                v[j][1] = ld_relaxed_v4(d + 128);
                                ++trace_retries;
                            }
                            ++trace_rows;
                            excl[0] += v[j][0].x; excl[1] += v[j][0].y; excl[2] += v[j][0].z; excl[3] += v[j][0].w;
                            excl[4] += v[j][1].x; excl[5] += v[j][1].y; excl[6] += v[j][1].z; excl[7] += v[j][1].w;
                            if ((all_and >> 30) == FLAG_PREFIX) done = true;   // tile 0 always carries a PREFIX
                            else ++agg_rows;
                        }
                    }
                    look -= WS_LB_BATCH;
                    if (lane == 0) WS_LB_STAMP(part, 2 * batch_no + 1);
                    ++batch_no;
                }
                const uint32_t flags = (agg_rows * FLAG_AGGREGATE + FLAG_PREFIX) << 30;   // modulo 2^32, like the sums
#pragma unroll
                for (int q = 0; q < 8; ++q) excl[q] -= flags;
                uint32_t* lb = lookback_pass + (size_t)part * SORT_BINS + lane * 4u;
                const uint4 c0 = *reinterpret_cast<const uint4*>(&sm.bin_count[lane * 4u]), c1 = *reinterpret_cast<const uint4*>(&sm.bin_count[128u + lane * 4u]);
                st_relaxed_v4(lb, make_uint4((FLAG_PREFIX << 30) | (excl[0] + c0.x), (FLAG_PREFIX << 30) | (excl[1] + c0.y),
                                             (FLAG_PREFIX << 30) | (excl[2] + c0.z), (FLAG_PREFIX << 30) | (excl[3] + c0.w)));
                st_relaxed_v4(lb + 128, make_uint4((FLAG_PREFIX << 30) | (excl[4] + c1.x), (FLAG_PREFIX << 30) | (excl[5] + c1.y),
                                                   (FLAG_PREFIX << 30) | (excl[6] + c1.z), (FLAG_PREFIX << 30) | (excl[7] + c1.w)));
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t bin = (q < 4 ? 0u : 128u) + lane * 4u + (uint32_t)(q & 3);
                sm.global_base[bin] = ctl->hist[pass][bin] + excl[q] - sm.bin_base[bin];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.lb_done);
            if (lane == 0) { WS_STAMP(part, 4); WS_NOTE(part, 8, (unsigned long long)trace_rows); WS_NOTE(part, 9, (unsigned long long)trace_retries); }
#else
            mbar_wait(&sm.lb_done, it & 1u);   // the group has resolved this tile's look-back: tickets follow that order
#endif
            part = draw();
        }
        return;
    }

    // ---------------- consumer group ----------------
    const uint32_t tid = threadIdx.x - group * WS_GROUP_THREADS, warp = tid >> 5;
    const uint32_t my_slot = hist_slot(tid);
    const uint32_t hist_row = smem_u32(&sm.warp_hist[warp][0]);
    for (uint32_t it = 0;; ++it) {
        const uint32_t b = it & 1u;
        mbar_wait(&sm.raw_full[b], (it >> 1) & 1u);
        const uint32_t part = sm.part[b];
        if (part == WS_END) break;
        if (tid == 0) WS_STAMP(part, 1);
        const uint32_t tile_base = part * WS_TILE;
        const uint32_t n_valid = min(WS_TILE, n - tile_base);
        const bool full = n_valid == WS_TILE;
        uint64_t* tile = sm.keys[b];

        // ---- keys: shared -> registers (warp-striped: item k of lane l is element warp * 32 * KPT + 32 k + l of the tile) ----
        uint64_t key[WS_KPT];
        const uint32_t local = warp * (32u * WS_KPT) + lane;
#pragma unroll
        for (uint32_t k = 0; k < WS_KPT; ++k) key[k] = tile[local + k * 32u];
        if (!full) {
#pragma unroll
            for (uint32_t k = 0; k < WS_KPT; ++k)
                if (local + k * 32u >= n_valid) key[k] = ~0ull;
        }
        {   // this warp's counters start at zero
            uint4* z = reinterpret_cast<uint4*>(&sm.warp_hist[warp][0]);
            z[lane] = make_uint4(0, 0, 0, 0);
            z[lane + 32] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        // ---- counter slots of the digits, four to a register; padding (last tile only) goes to the top bin ----
        uint32_t dpack[WS_KPT / 4];
#pragma unroll
        for (uint32_t q = 0; q < WS_KPT / 4; ++q) {
            uint32_t w = 0;
#pragma unroll
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t k = q * 4 + r;
                const bool valid = full || (local + k * 32u) < n_valid;
                w |= hist_slot(valid ? digit_of(key[k]) : mask) << (8u * r);
            }
            dpack[q] = w;
        }
        auto slot_at = [&](uint32_t k) { return (dpack[k >> 2] >> (8u * (k & 3u))) & 0xffu; };
#pragma unroll
        for (uint32_t k = 0; k < WS_KPT; ++k) atomicAdd(&sm.warp_hist[warp][slot_at(k)], 1u);
        group_sync(group);   // every key of the tile is in registers and counted

        // ---- per bin (thread == bin): prefix over the warps, publish the tile aggregate, scan the bins ----
        uint32_t bin_count = 0;
#pragma unroll
        for (uint32_t w = 0; w < WS_GROUP_WARPS; ++w) {
            const uint32_t c = sm.warp_hist[w][my_slot];
            sm.warp_hist[w][my_slot] = bin_count;
            bin_count += c;
        }
        const uint32_t bin_count_valid = (tid == mask) ? bin_count - (WS_TILE - n_valid) : bin_count;
        st_relaxed_u32(lookback_pass + (size_t)part * SORT_BINS + tid, ((part == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) << 30) | bin_count_valid);
        uint32_t incl = bin_count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        if (lane == 31) sm.scan[warp] = incl;
        uint32_t bin_base = incl - bin_count;
        group_sync(group);
#pragma unroll
        for (uint32_t w = 0; w < SORT_BINS / 32; ++w)
            if (w < warp) bin_base += sm.scan[w];
        sm.bin_count[tid] = bin_count_valid;
        sm.bin_base[tid] = bin_base;
#pragma unroll
        for (uint32_t w = 0; w < WS_GROUP_WARPS; ++w) sm.warp_hist[w][my_slot] += bin_base;
        __syncwarp();
#if TPDCU_WS_HELPER_LOOKBACK
        if (lane == 0) mbar_arrive(&sm.agg_ready);   // the helper may resolve this tile's look-back now
#endif
        if (tid == 0) WS_STAMP(part, 2);
        group_sync(group);
#if !TPDCU_WS_HELPER_LOOKBACK && TPDCU_WS_LOOKBACK_EARLY
        // first round trip of the look-back: issued now, consumed after the ranking
        uint32_t v_early[WS_LB_BATCH];
#pragma unroll
        for (int j = 0; j < WS_LB_BATCH; ++j)
            v_early[j] = part > 0 ? ld_relaxed_u32(lookback_pass + (size_t)max((int)part - 1 - j, 0) * SORT_BINS + tid) : 0u;
#endif

        // ---- stable ranking (one returning shared-memory atomic per key) + scatter into the buffer the keys came from ----
#pragma unroll
        for (uint32_t k0 = 0; k0 < WS_KPT; k0 += SORT_RANK_BATCH) {
            uint32_t r[SORT_RANK_BATCH];
#pragma unroll
            for (uint32_t j = 0; j < SORT_RANK_BATCH; ++j)
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r[j]) : "r"(hist_row + 4u * slot_at(k0 + j)) : "memory");
#pragma unroll
            for (uint32_t j = 0; j < SORT_RANK_BATCH; ++j) tile[r[j]] = key[k0 + j];
        }
        if (tid == 0) WS_STAMP(part, 5);
#if TPDCU_WS_HELPER_LOOKBACK
        group_sync(group);
        mbar_wait(&sm.lb_done, it & 1u);
#else
        {   // ---- decoupled look-back, one thread per bin ----
            if (tid == 0) WS_STAMP(part, 3);
            uint32_t excl = 0, trace_rows = 0;
            (void)trace_rows;
            if (part > 0) {
                int look = (int)part - 1;
                bool done = false;
#if TPDCU_WS_LOOKBACK_EARLY
                bool first = true;
#endif
                while (!done) {
                    uint32_t v[WS_LB_BATCH];
#if TPDCU_WS_LOOKBACK_EARLY
                    if (first) {
#pragma unroll
                        for (int j = 0; j < WS_LB_BATCH; ++j) v[j] = v_early[j];
                        first = false;
                    } else
#endif
                    {
#pragma unroll
                        for (int j = 0; j < WS_LB_BATCH; ++j)
                            v[j] = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid);
                    }
#pragma unroll
                    for (int j = 0; j < WS_LB_BATCH; ++j) {
                        if (!done) {
                            uint32_t x = v[j];
                            while ((x >> 30) == FLAG_INVALID) x = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid);
                            excl += x & LOOKBACK_VALUE_MASK;
                            done = (x >> 30) == FLAG_PREFIX;  // tile 0 always carries a PREFIX
                            ++trace_rows;
                        }
                    }
                    look -= WS_LB_BATCH;
                }
                st_relaxed_u32(lookback_pass + (size_t)part * SORT_BINS + tid, (FLAG_PREFIX << 30) | (excl + bin_count_valid));
            }
            sm.global_base[tid] = ctl->hist[pass][tid] + excl - bin_base;
            if (tid == 0) { WS_STAMP(part, 4); WS_NOTE(part, 8, (unsigned long long)trace_rows); WS_NOTE(part, 9, 0ull); }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.lb_done);   // the helper draws the group's next ticket now
        }
        group_sync(group);
#endif
        if (tid == 0) WS_STAMP(part, 6);

        // ---- write-out: position i of the locally sorted tile goes to global_base[digit] + i ----
        if (full) {
#pragma unroll
            for (uint32_t k = 0; k < WS_KPT; ++k) {
                const uint32_t i = tid + k * WS_GROUP_THREADS;
                const uint64_t kk = tile[i];
                dst_keys[sm.global_base[digit_of(kk)] + i] = kk;
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < WS_KPT; ++k) {
                const uint32_t i = tid + k * WS_GROUP_THREADS;
                if (i < n_valid) {
                    const uint64_t kk = tile[i];
                    dst_keys[sm.global_base[digit_of(kk)] + i] = kk;
                }
            }
        }
        // the buffer was written through the generic proxy (the scatter) and is about to be written by TMA (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.raw_empty[b]);   // this buffer may receive the tile after next
        if (tid == 0) WS_STAMP(part, 7);
    }
}

#endif  // TPDCU_SORT_WS

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

uint32_t sort_parts(uint32_t capacity, uint32_t kind) {
    const uint32_t tile = kind == SORT_KIND_PAIRS ? SORT_TILE_PAIRS : SORT_TILE_WORDS;
    return (capacity + tile - 1) / tile;
}
uint32_t sort_passes_for(uint32_t end_bit) { return passes_needed(end_bit); }

template <int MODE>
static cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(onesweep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(OnesweepSmem<MODE>));
}

// opt in to > 48 KB of dynamic shared memory; called once per context, outside any stream capture
cudaError_t init_sort_attributes() {
    cudaError_t e = set_smem_attr<MODE_PAIRS>();
    if (e == cudaSuccess) e = set_smem_attr<MODE_WORDS>();
#if TPDCU_SORT_WS
    if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem));
#endif
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
    const uint32_t hist_max = (uint32_t)a.sm_count * 4u;
    if (hist_grid > hist_max) hist_grid = hist_max;
    if (words) sort_hist_kernel<true><<<hist_grid, HIST_THREADS, 0, s>>>(a.keys[0], a.frame, a.ctl, a.plan, a.kind, n_host, a.capacity, a.end_bit, a.tile_bits);
    else sort_hist_kernel<false><<<hist_grid, HIST_THREADS, 0, s>>>(a.keys[0], a.frame, a.ctl, a.plan, a.kind, n_host, a.capacity, a.end_bit, a.tile_bits);
    if (ev_after_plan) cudaEventRecord(ev_after_plan, s);
    const uint32_t parts = sort_parts(bound, a.kind);
    const uint32_t parts_cap = sort_parts(a.capacity, a.kind);
    for (uint32_t p = 0; p < num_passes; ++p) {
        uint32_t* lb = a.lookback + (size_t)p * parts_cap * SORT_BINS;
#if TPDCU_SORT_WS
        if (words)  // persistent: one CTA per SM (fewer when the buffer cannot hold that many tiles), tiles drawn by ticket
            onesweep_ws_kernel<<<std::min<uint32_t>((uint32_t)a.sm_count, (parts + WS_GROUPS - 1) / WS_GROUPS), WS_THREADS, sizeof(WsSmem), s>>>(a.keys[0], a.keys[1], a.ctl, a.plan, lb, p);
#else
        if (words)
            onesweep_kernel<MODE_WORDS><<<parts, SORT_THREADS, sizeof(OnesweepSmem<MODE_WORDS>), s>>>(a.keys[0], a.keys[1], nullptr, nullptr, a.ctl, a.plan, lb, p);
#endif
        else
            onesweep_kernel<MODE_PAIRS><<<parts, SORT_THREADS, sizeof(OnesweepSmem<MODE_PAIRS>), s>>>(a.keys[0], a.keys[1], a.vals[0], a.vals[1], a.ctl, a.plan, lb, p);
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

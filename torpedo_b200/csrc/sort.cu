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
    uint32_t global_base[SORT_BINS];
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    uint32_t dpack[SORT_KPT / 4];
#pragma unroll
    for (uint32_t q = 0; q < SORT_KPT / 4; ++q) {
        uint32_t w = 0;
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
            const uint32_t k = q * 4 + r;
            const bool valid = full || (warp_base + k * 32u) < n;
            w |= (valid ? digit_in(key[k]) : mask) << (8u * r);
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
#pragma unroll
    for (uint32_t w = 0; w < SORT_WARPS; ++w) {
        const uint32_t c = sm.warp_hist[w][tid];
        sm.warp_hist[w][tid] = bin_count;
        bin_count += c;
    }
    const uint32_t bin_count_valid = (tid == mask) ? bin_count - (SORT_TILE - n_valid) : bin_count;  // padding lives in the top bin
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
    for (uint32_t w = 0; w < SORT_WARPS; ++w) sm.warp_hist[w][tid] += bin_base;
    __syncthreads();

    // ---- stable ranking: peers with the same digit inside a 32-key row, rows in order; elements go straight to smem ---
    uint32_t rank[OUT_PAIRS ? SORT_KPT : 1];
#pragma unroll
    for (uint32_t k = 0; k < SORT_KPT; ++k) {
        const uint32_t d = digit_at(k);
        // peers = lanes of this row holding the same digit. Eight ballots + bit logic on the ALU pipe instead of one
        // match.any: match.any executes on the address-divergence unit, which this kernel's shared-memory traffic already
        // keeps busy (ncu: ADU 39 %, LSU 57 % with match.any) — the ballot form made every pass ~20 % faster.
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

    // ---- decoupled look-back, one thread per bin -----------------------------------------------------
    {
        uint32_t excl = 0;
        if (part > 0) {
            // Tiles in flight publish their aggregate well before their prefix, so the walk back to the nearest PREFIX is
            // several tiles deep: read LOOKBACK_BATCH descriptors per round trip, consume them in order.
            int look = (int)part - 1;
            bool done = false;
            while (!done) {
                uint32_t v[LOOKBACK_BATCH];
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j)
                    v[j] = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid);
#pragma unroll
                for (int j = 0; j < (int)LOOKBACK_BATCH; ++j) {
                    if (!done) {
                        uint32_t x = v[j];
                        while ((x >> 30) == FLAG_INVALID) x = ld_relaxed_u32(lookback_pass + (size_t)max(look - j, 0) * SORT_BINS + tid);
                        excl += x & LOOKBACK_VALUE_MASK;
                        done = (x >> 30) == FLAG_PREFIX;  // tile 0 always carries a PREFIX, so look - j never goes below 0 unconsumed
                    }
                }
                look -= (int)LOOKBACK_BATCH;
            }
            st_relaxed_u32(lb + tid, (FLAG_PREFIX << 30) | (excl + bin_count_valid));
        }
        sm.global_base[tid] = ctl->hist[pass][tid] + excl - bin_base;
    }
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
        if (words)
            onesweep_kernel<MODE_WORDS><<<parts, SORT_THREADS, sizeof(OnesweepSmem<MODE_WORDS>), s>>>(a.keys[0], a.keys[1], nullptr, nullptr, a.ctl, a.plan, lb, p);
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

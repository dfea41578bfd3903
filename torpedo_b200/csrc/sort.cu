// sort.cu — stable 64-bit-key / 32-bit-value LSD radix sort in the onesweep style
// (one up-front multi-digit histogram, then ONE read + ONE write of every pair per 8-bit digit,
// with a chained decoupled look-back across tiles instead of a separate scan per pass).
//
// Replaces the reference's 4-way radix sorter: radix-shuffle.slang:37-150, radix-prefixA.slang:37-169,
// radix-prefixB.slang:37-168, radix-mapping.slang:38-116 driven by GaussianEngine.cpp:822-841
// (23 passes x 4 dispatches at 1080p, ~1.1 KB of DRAM traffic per pair) with
// ceil(end_bit/8) passes (6 at 1080p) of 24 B per pair each, + 8 B per pair for the histogram.
//
// The number of pairs lives in device memory (FrameCtl::pairs_total): the host never learns P inside a
// frame, grids are sized by the buffer capacity and surplus CTAs exit on their ticket.
#include "common.cuh"

namespace tpdcu {

constexpr uint32_t HIST_THREADS = 512;
constexpr uint32_t HIST_KPT = 8;
constexpr uint32_t LOOKBACK_VALUE_MASK = (1u << 30) - 1u;
constexpr uint32_t LOOKBACK_BATCH = 8;
#ifndef TPDCU_SORT_PREFETCH_TILES
#define TPDCU_SORT_PREFETCH_TILES 296
#endif
#ifndef TPDCU_SORT_MINB_PACKED
#define TPDCU_SORT_MINB_PACKED 4
#endif
constexpr uint32_t SORT_PREFETCH_TILES = TPDCU_SORT_PREFETCH_TILES;  // 148 SMs x 3 resident CTAs

// Order-preserving key compaction (frame path): keys are tile << 32 | float_bits(viewZ) with viewZ confined to
// [min, max] of the frame, so the passes sort on  tile << depth_bits | (depth - min)  instead — at 1080p with the default
// near/far planes that is 27 + 13 = 40 bits = 5 passes instead of 6. The stored keys are never modified.
struct KeyXform {
    uint32_t bias, depth_bits, total_bits;
};
__device__ __forceinline__ KeyXform make_xform(uint32_t depth_min, uint32_t depth_max, uint32_t end_bit, bool frame_keys) {
    KeyXform x;
    if (!frame_keys) { x.bias = 0; x.depth_bits = min(end_bit, 32u); x.total_bits = end_bit; return x; }
    x.bias = depth_max >= depth_min ? depth_min : 0u;
    const uint32_t span = depth_max >= depth_min ? depth_max - depth_min : 0u;
    x.depth_bits = 32u - __clz(span);           // 0 when every key carries the same depth
    x.total_bits = x.depth_bits + (end_bit - 32u);
    return x;
}
__device__ __forceinline__ uint32_t digit_of(uint64_t key, const KeyXform& x, uint32_t shift, uint32_t mask) {
    const uint32_t lo = (uint32_t)key - x.bias, hi = (uint32_t)(key >> 32);
    const uint64_t packed = x.depth_bits >= 32u ? (((uint64_t)hi << 32) | lo) : (((uint64_t)hi << x.depth_bits) | lo);
    return (uint32_t)(packed >> shift) & mask;
}
__device__ __forceinline__ uint64_t pack_word(uint64_t key, uint32_t val, const KeyXform& x, uint32_t idx_bits) {
    const uint32_t lo = (uint32_t)key - x.bias, hi = (uint32_t)(key >> 32);
    const uint64_t packed = x.depth_bits >= 32u ? (((uint64_t)hi << 32) | lo) : (((uint64_t)hi << x.depth_bits) | lo);
    return (packed << idx_bits) | val;
}

__device__ __forceinline__ uint32_t pass_mask(uint32_t pass, uint32_t total_bits) {
    const uint32_t left = total_bits > pass * SORT_RADIX_BITS ? total_bits - pass * SORT_RADIX_BITS : 0u;
    return left >= SORT_RADIX_BITS ? (SORT_BINS - 1u) : ((1u << left) - 1u);
}
__device__ __forceinline__ uint32_t passes_needed(uint32_t total_bits) { return (total_bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS; }

// ---------------------------------------------------------------------------------------------------
// up-front histogram of every digit (one read of the keys)
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(HIST_THREADS) sort_hist_kernel(const uint64_t* __restrict__ keys, FrameCtl* ctl,
                                                                  uint32_t n_host, uint32_t capacity, uint32_t num_passes,
                                                                  uint32_t end_bit) {
    __shared__ uint32_t h[SORT_MAX_PASSES][SORT_BINS];
    const uint32_t n = n_host == UINT32_MAX ? min(ctl->pairs_total, capacity) : n_host;
    const KeyXform xf = make_xform(~ctl->inv_depth_min, ctl->depth_max, end_bit, n_host == UINT32_MAX);
    num_passes = min(num_passes, passes_needed(xf.total_bits));  // passes above the packed key width see a single bin
    for (uint32_t k = threadIdx.x; k < num_passes * SORT_BINS; k += HIST_THREADS) (&h[0][0])[k] = 0;
    __syncthreads();
    // Each thread takes HIST_KPT CONSECUTIVE keys (four 16-byte loads): the pairs of one Gaussian are adjacent in the
    // unsorted buffer and share their depth bits (and usually the upper tile bits), so run-length encoding the digits in
    // registers removes most shared-memory atomics and nearly all same-address conflicts.
    const uint32_t chunk = HIST_THREADS * HIST_KPT;
    for (uint32_t base = blockIdx.x * chunk; base < n; base += gridDim.x * chunk) {
        const uint32_t first = base + threadIdx.x * HIST_KPT;
        uint64_t k[HIST_KPT];
        if (first + HIST_KPT <= n) {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT / 2; ++j) {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(keys + first) + j);
                k[2 * j] = v.x;
                k[2 * j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT; ++j) k[j] = first + j < n ? keys[first + j] : 0ull;
        }
        const uint32_t valid = first >= n ? 0u : min(HIST_KPT, n - first);
        if (valid) {
#pragma unroll
            for (uint32_t j = 0; j < HIST_KPT; ++j) k[j] = pack_word(k[j], 0u, xf, 0u);  // tile << depth_bits | depth - bias, once per key
            for (uint32_t p = 0; p < num_passes; ++p) {
                const uint32_t shift = p * SORT_RADIX_BITS, mask = pass_mask(p, xf.total_bits);
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
}

// ---------------------------------------------------------------------------------------------------
// plan: exclusive digit offsets, identity-pass detection, ping-pong schedule
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(SORT_BINS) sort_plan_kernel(FrameCtl* ctl, SortPlan* plan, uint32_t n_host, uint32_t capacity,
                                                               uint32_t num_passes, uint32_t end_bit, uint32_t packed_idx_bits, uint32_t packed_word_bits) {
    __shared__ uint32_t s_warp[SORT_BINS / 32];
    __shared__ uint32_t s_skip[SORT_MAX_PASSES];
    const uint32_t n = n_host == UINT32_MAX ? min(ctl->pairs_total, capacity) : n_host;
    const uint32_t b = threadIdx.x, lane = b & 31u, warp = b >> 5;
    const KeyXform xf = make_xform(~ctl->inv_depth_min, ctl->depth_max, end_bit, n_host == UINT32_MAX);
    num_passes = min(num_passes, passes_needed(xf.total_bits));
    if (b < SORT_MAX_PASSES) s_skip[b] = 0;
    __syncthreads();
    for (uint32_t p = 0; p < num_passes; ++p) {
        const uint32_t c = ctl->hist[p][b];
        if (c == n) s_skip[p] = 1;  // every key falls in this bin (also true for n == 0)
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wex = 0;
        for (uint32_t w = 0; w < warp; ++w) wex += s_warp[w];
        ctl->hist[p][b] = wex + incl - c;
        __syncthreads();
    }
    if (b == 0) {
        uint32_t sel = 0, run = 0;
        const bool packed = packed_idx_bits != 0u;
        for (uint32_t p = 0; p < SORT_MAX_PASSES; ++p) {
            uint32_t skip = p < num_passes ? s_skip[p] : 1u;
            if (packed && p == 0 && n > 0) skip = 0;  // the first pass also converts pairs to packed words: never skipped
            plan->skip[p] = skip;
            plan->src_sel[p] = sel;
            if (!skip) { sel ^= 1u; ++run; }
        }
        plan->n = n;
        plan->num_passes = num_passes;
        plan->final_sel = sel;
        plan->passes_run = run;
        plan->bias = xf.bias;
        plan->depth_bits = xf.depth_bits;
        plan->total_bits = xf.total_bits;
        plan->idx_bits = packed ? packed_idx_bits : 0u;
        plan->packed = packed ? 1u : 0u;
        // a packed word must hold tile | depth - bias | index; if this frame's depth range is too wide the host re-renders in pair mode
        plan->packed_overflow = (packed && xf.total_bits + packed_idx_bits > packed_word_bits) ? 1u : 0u;
    }
}

// ---------------------------------------------------------------------------------------------------
// one onesweep pass
// ---------------------------------------------------------------------------------------------------

// Sort modes. PAIRS: (u64 key, u32 value) in and out (standalone API, fallback). PACK: pairs in, single 64-bit words out
// (first pass of a frame). PACKED: words in and out. A word is  tile << (depth_bits+idx_bits) | (depth-bias) << idx_bits | index:
// the Gaussian index rides in the low bits of the key, so a pass moves 16 B per pair instead of 24 B, the value scatter
// through shared memory disappears, and — pairs being emitted in ascending index order — sorting the bits above idx_bits
// stably is exactly the reference's stable sort of (key, value) pairs.
enum : int { MODE_PAIRS = 0, MODE_PACK = 1, MODE_PACKED = 2 };

template <bool WITH_VALS>
struct OnesweepSmem {
    uint64_t keys[SORT_TILE];
    alignas(16) uint32_t warp_hist[SORT_WARPS][SORT_BINS];  // zeroed with 16-byte stores
    uint32_t global_base[SORT_BINS];
    uint32_t scan[SORT_BINS / 32];
    uint32_t part;
    uint32_t vals[WITH_VALS ? SORT_TILE : 1];
};
static_assert(SORT_THREADS == SORT_BINS, "one thread per bin in the per-bin phases");

// One CTA = one tile of SORT_TILE pairs. Phases (block barriers in between):
//   ticket + zero per-warp histograms | load keys, early counts | per-bin: warp prefix, publish aggregate, bin scan |
//   stable ranking (match.any) + scatter to smem | look-back per bin | coalesced write-out (+ value scatter / write-out)
template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS, MODE == MODE_PACKED ? TPDCU_SORT_MINB_PACKED : TPDCU_SORT_MINB)
onesweep_kernel(uint64_t* keys0, uint64_t* keys1, uint32_t* vals0, uint32_t* vals1, FrameCtl* ctl,
                const SortPlan* __restrict__ plan, uint32_t* lookback_pass, uint32_t pass) {
    constexpr bool IN_PAIRS = MODE != MODE_PACKED, OUT_PAIRS = MODE == MODE_PAIRS;
    using Smem = OnesweepSmem<OUT_PAIRS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    if (plan->skip[pass]) return;
    const uint32_t n = plan->n;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) sm.part = atomicAdd(&ctl->sort_ticket[pass], 1u);
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
    const KeyXform xf{ plan->bias, plan->depth_bits, plan->total_bits };
    const uint32_t idx_bits = plan->idx_bits;
    const uint32_t shift = pass * SORT_RADIX_BITS, mask = pass_mask(pass, xf.total_bits);
    // digit of an INPUT element / of an element as it sits in shared memory (= output format)
    auto digit_in = [&](uint64_t k) { return IN_PAIRS ? digit_of(k, xf, shift, mask) : ((uint32_t)(k >> (idx_bits + shift)) & mask); };
    auto digit_out = [&](uint64_t k) { return OUT_PAIRS ? digit_of(k, xf, shift, mask) : ((uint32_t)(k >> (idx_bits + shift)) & mask); };

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
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t lower = __popc(peers & lanemask_lt());
        const uint32_t base = sm.warp_hist[warp][d];
        __syncwarp();
        if (lower == 0) sm.warp_hist[warp][d] = base + __popc(peers);
        __syncwarp();
        const uint32_t r = base + lower;
        if (OUT_PAIRS) rank[k] = r;
        sm.keys[r] = MODE == MODE_PACK ? pack_word(key[k], val[k], xf, idx_bits) : key[k];
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

// introspection: packed words -> the reference's (key, value) arrays
__global__ void sort_unpack_kernel(const uint64_t* keys0, const uint64_t* keys1, const SortPlan* plan, uint64_t* out_keys, uint32_t* out_vals) {
    const uint64_t* w = plan->final_sel ? keys1 : keys0;
    const uint32_t n = plan->n, ib = plan->idx_bits, db = plan->depth_bits, bias = plan->bias;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t x = w[i];
        const uint64_t kp = x >> ib;
        const uint32_t depth = (db >= 32u ? (uint32_t)kp : (uint32_t)(kp & ((1ull << db) - 1ull))) + bias;
        const uint32_t tile = db >= 32u ? (uint32_t)(kp >> 32) : (uint32_t)(kp >> db);
        out_keys[i] = ((uint64_t)tile << 32) | depth;
        out_vals[i] = (uint32_t)(x & ((1ull << ib) - 1ull));
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

uint32_t sort_parts(uint32_t capacity) { return (capacity + SORT_TILE - 1) / SORT_TILE; }

template <int MODE>
static cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(onesweep_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(OnesweepSmem<MODE == MODE_PAIRS>));
}

// opt in to > 48 KB of dynamic shared memory; called once per context, outside any stream capture
cudaError_t init_sort_attributes() {
    cudaError_t e = set_smem_attr<MODE_PAIRS>();
    if (e == cudaSuccess) e = set_smem_attr<MODE_PACK>();
    if (e == cudaSuccess) e = set_smem_attr<MODE_PACKED>();
    return e;
}

cudaError_t launch_sort(const SortLaunch& a, uint32_t n_host, cudaStream_t s, cudaEvent_t ev_after_plan) {
    const uint32_t num_passes = (a.end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    const uint32_t bound = n_host == UINT32_MAX ? a.capacity : n_host;
    const uint32_t pib = a.packed_idx_bits;
    if (bound == 0 || num_passes == 0) {
        sort_plan_kernel<<<1, SORT_BINS, 0, s>>>(a.ctl, a.plan, n_host == UINT32_MAX ? UINT32_MAX : 0u, a.capacity, 0, a.end_bit, pib, a.packed_word_bits ? a.packed_word_bits : 64u);
        if (ev_after_plan) cudaEventRecord(ev_after_plan, s);
        return cudaGetLastError();
    }
    const uint32_t chunk = HIST_THREADS * HIST_KPT;
    uint32_t hist_grid = (bound + chunk - 1) / chunk;
    const uint32_t hist_max = (uint32_t)a.sm_count * 4u;
    if (hist_grid > hist_max) hist_grid = hist_max;
    sort_hist_kernel<<<hist_grid, HIST_THREADS, 0, s>>>(a.keys[0], a.ctl, n_host, a.capacity, num_passes, a.end_bit);
    sort_plan_kernel<<<1, SORT_BINS, 0, s>>>(a.ctl, a.plan, n_host, a.capacity, num_passes, a.end_bit, pib, a.packed_word_bits ? a.packed_word_bits : 64u);
    if (ev_after_plan) cudaEventRecord(ev_after_plan, s);
    const uint32_t parts = sort_parts(bound);
    const uint32_t parts_cap = sort_parts(a.capacity);
    for (uint32_t p = 0; p < num_passes; ++p) {
        uint32_t* lb = a.lookback + (size_t)p * parts_cap * SORT_BINS;
        if (pib == 0)
            onesweep_kernel<MODE_PAIRS><<<parts, SORT_THREADS, sizeof(OnesweepSmem<true>), s>>>(a.keys[0], a.keys[1], a.vals[0], a.vals[1], a.ctl, a.plan, lb, p);
        else if (p == 0)
            onesweep_kernel<MODE_PACK><<<parts, SORT_THREADS, sizeof(OnesweepSmem<false>), s>>>(a.keys[0], a.keys[1], a.vals[0], a.vals[1], a.ctl, a.plan, lb, p);
        else
            onesweep_kernel<MODE_PACKED><<<parts, SORT_THREADS, sizeof(OnesweepSmem<false>), s>>>(a.keys[0], a.keys[1], a.vals[0], a.vals[1], a.ctl, a.plan, lb, p);
    }
    return cudaGetLastError();
}

cudaError_t launch_sort_unpack(const SortLaunch& a, uint64_t* out_keys, uint32_t* out_vals, cudaStream_t s) {
    sort_unpack_kernel<<<(uint32_t)a.sm_count * 8u, 256, 0, s>>>(a.keys[0], a.keys[1], a.plan, out_keys, out_vals);
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

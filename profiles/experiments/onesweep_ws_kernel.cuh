// EXPERIMENT, not compiled into the product (round 2): a warp-specialised persistent onesweep pass — one CTA per SM, two
// consumer groups of eight warps, one helper warp per group (ticket + TMA bulk load of the next tile; optionally the look-back).
// Result on B200 (6 M-Gaussian headline frame, tile-sort pass over 15.86 M words): 82-134 us per pass against 74 us for the
// per-tile kernel it was meant to replace. What the per-tile trace (WS_STAMP) showed: the look-back, not the key loads, is
// what a tile waits for — a strong 16-byte descriptor load takes ~1.5 us per batch from a single helper warp competing with
// sixteen consumer warps for the LSU, and tickets drawn a tile ahead of their processing order every generation of tiles
// behind its slowest member. The lessons that stayed in sort.cu: returning shared-memory atomics for the ranking, the
// bank-swizzled counters, look-back chains (SORT_CHAINS) and the persistent per-tile loop that prefetches its next tile.
// Kept as it was when it last ran (it expects the helpers of sort.cu around it).
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


// EXPERIMENT, not compiled into the product (round 2): a warp-specialised blend — per tile four DRAIN warps (one per 8x8
// quadrant) and four FILL warps, the splat queue double-buffered behind full / empty mbarriers, no block barrier between staging
// and compositing, the block vote replaced by a counter of finished quadrants. Bit-identical images (the parity suite passes).
// Result on B200 (6 M-Gaussian headline frame): 0.416 ms at 3 CTAs per SM (72 registers), 0.365 at 4 (64), 0.343 at 5 (48, spills),
// against 0.332 ms for blend_kernel at 7 CTAs x 4 warps: the staging latency it hides is already hidden by seven co-resident
// CTAs, while half of its warps (the fill warps) are idle most of the time and the drain warps per SM drop from 28 to 12-20.
// Kept as it was when it last ran (it expects the helpers of raster.cu around it).
// ---------------------------------------------------------------------------------------------------
// blend, warp-specialised: four DRAIN warps (one per 8x8 quadrant) and four FILL warps per tile
//
// The kernel above alternates: 128 threads stage a round (cull, compaction, SH colour of the survivors), then the four
// quadrant warps drain it, then all vote. ncu: a quarter of its stall samples wait for the survivors' rows, a fifth sit on
// the barriers between the two halves. Here the two halves are different warps of the CTA and the queue is double-buffered:
// the fill warps stage round r + 1 (its gathers, its colour evaluation) while the drain warps composite round r; full / empty
// mbarriers per buffer are the only coupling, so a quadrant that is ahead runs up to one buffer ahead of the slowest, and no
// drain warp ever executes staging code. The fill runs at most one round past the one in which the tile saturates (it looks
// at the drain's stop flag right after every cull, before the rows are gathered). The arithmetic per pixel and splat is the
// one of blend_kernel: same order of splats, same operations, same bytes out.
// ---------------------------------------------------------------------------------------------------

constexpr uint32_t WSB_THREADS = 2 * BLEND_THREADS;   // warps 0-3 drain, warps 4-7 fill
struct WsbBuf {
    BlendEntry ent[BLEND_QUEUE];
    uint16_t list[BLEND_WARPS][BLEND_QUEUE];
    uint32_t ln[BLEND_WARPS];   // list length per quadrant
    uint32_t last;              // no round follows this one
};
struct WsbSmem {
    WsbBuf buf[2];
    uint32_t cnt[BLEND_WARPS][BLEND_WARPS + 1];
    uint32_t stop;              // every pixel of the tile is done (set by the drain warp that notices)
    uint32_t fstop;             // the fill warps' uniform copy of `stop` for the current round
    uint32_t done_quadrants, rounds_used;
    alignas(8) uint64_t full[2], empty[2];
};
#ifndef TPDCU_WSB_MINB
#define TPDCU_WSB_MINB 3
#endif

__device__ __forceinline__ void wsb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ready = 0;
    while (!ready)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ready) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void wsb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

__global__ void __launch_bounds__(WSB_THREADS, TPDCU_WSB_MINB) blend_ws_kernel(RasterLaunch a) {
    __shared__ WsbSmem sm;
    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX;
    const uint32_t tile = a.order[blockIdx.x];  // longest lists first (tile_order_kernel)
    const uint32_t tile_x0 = (tile % gx) * TILE_PX, tile_y0 = (tile / gx) * TILE_PX;
    const uint32_t lane = threadIdx.x & 31u;
    const bool fill = threadIdx.x >= BLEND_THREADS;
    const uint32_t tid = fill ? threadIdx.x - BLEND_THREADS : threadIdx.x, warp = tid >> 5;   // index inside the role
    const uint2 range = reinterpret_cast<const uint2*>(a.ranges)[tile];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.full[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.empty[b])), "r"(BLEND_WARPS) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.stop = 0; sm.fstop = 0; sm.done_quadrants = 0; sm.rounds_used = 0;
    }
    __syncthreads();

    if (fill) {
        // =============================== fill warps ===============================
        const uint64_t* __restrict__ words = a.plan->final_sel ? a.keys[1] : a.keys[0];
        const float cam_pos[3] = { a.cam->cam_pos[0], a.cam->cam_pos[1], a.cam->cam_pos[2] };
        const float4* __restrict__ geo4 = reinterpret_cast<const float4*>(a.geo);
        const float tile_fx0 = (float)tile_x0, tile_fy0 = (float)tile_y0;
        auto fill_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(BLEND_THREADS) : "memory"); };
        uint32_t in = range.x;
        uint32_t g_cur = 0, g_next = 0;
        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
        if (in + tid < range.y) g_cur = (uint32_t)__ldg(words + in + tid);
        if (in + BLEND_THREADS + tid < range.y) g_next = (uint32_t)__ldg(words + in + BLEND_THREADS + tid);
        if (in + tid < range.y) ldg256(geo4 + (size_t)g_cur * 2, ra, rb);
        for (uint32_t r = 0;; ++r) {
            WsbBuf& q = sm.buf[r & 1u];
            if (r >= 2) wsb_wait(&sm.empty[r & 1u], ((r >> 1) - 1u) & 1u);   // the drain warps have left this buffer
            const bool exhausted = in >= range.y;
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
            for (uint32_t k = 0; k <= BLEND_WARPS; ++k) {
                ballot[k] = __ballot_sync(0xffffffffu, (keep >> k) & 1u);
                if (lane == k) sm.cnt[warp][k] = __popc(ballot[k]);
            }
            if (tid == 0) sm.fstop = *reinterpret_cast<volatile uint32_t*>(&sm.stop);   // one reading for all four fill warps
            // the next round's SplatGeo record (its word arrived a round ago) and the word of the round after it
            float4 ra_n = make_float4(0.f, 0.f, 0.f, 0.f), rb_n = ra_n;
            uint32_t g_next2 = 0;
            if (idx + BLEND_THREADS < range.y) ldg256(geo4 + (size_t)g_next * 2, ra_n, rb_n);
            if (idx + 2 * BLEND_THREADS < range.y) g_next2 = (uint32_t)__ldg(words + idx + 2 * BLEND_THREADS);
            fill_sync();
            const bool stop = sm.fstop != 0u;
            uint32_t before[BLEND_WARPS + 1], total[BLEND_WARPS + 1];
#pragma unroll
            for (uint32_t k = 0; k <= BLEND_WARPS; ++k) {
                before[k] = 0;
                total[k] = 0;
#pragma unroll
                for (uint32_t w = 0; w < BLEND_WARPS; ++w) {
                    const uint32_t c = sm.cnt[w][k];
                    if (w < warp) before[k] += c;
                    total[k] += c;
                }
            }
            if (keep && !stop) {
                const uint32_t pos = before[0] + __popc(ballot[0] & lanemask_lt());
                const float4 po = __ldg(a.posop + g);
                q.ent[pos].g0 = make_float4(ra.x, ra.y, (-0.5f * LOG2E) * ra.z, -LOG2E * ra.w);
                q.ent[pos].g1 = make_float4((-0.5f * LOG2E) * rb.x, rb.y, -__log2f(255.0f * rb.y) - 0.01f, 0.0f);
#pragma unroll
                for (uint32_t k = 0; k < BLEND_WARPS; ++k)
                    if (keep & (2u << k)) q.list[k][before[k + 1] + __popc(ballot[k + 1] & lanemask_lt())] = (uint16_t)(pos * sizeof(BlendEntry));
                const float3 c3 = sh_color<TPDCU_BLEND_FAST_DIRECTION>(a.sh + (size_t)g * SH_PLANES, po.x, po.y, po.z, cam_pos, (int)a.sh_degree);
                q.ent[pos].col = make_float4(c3.x, c3.y, c3.z, 0.0f);
            }
            in += BLEND_THREADS;
            const bool last = stop || exhausted || in >= range.y;
            if (tid < BLEND_WARPS) q.ln[tid] = stop ? 0u : total[tid + 1];
            if (tid == 0) q.last = last ? 1u : 0u;
            g_cur = g_next; g_next = g_next2; ra = ra_n; rb = rb_n;
            fill_sync();                                   // every entry of the round is written; sm.cnt may be rewritten
            if (tid == 0) wsb_arrive(&sm.full[r & 1u]);    // release: the drain warps may read the buffer
            if (last) break;
        }
        return;
    }

    // =============================== drain warps ===============================
    // quadrant `warp`: origin (8*(warp&1), 8*(warp>>1)); lane -> pixel pair at (2*(lane&3), lane>>2) inside it
    const uint32_t x0 = tile_x0 + 8u * (warp & 1u) + 2u * (lane & 3u), y0 = tile_y0 + 8u * (warp >> 1) + (lane >> 2);
    float fx0 = (float)x0, fy0 = (float)y0;
    uint32_t inside = 0;  // bit k: pixel (x0 + k, y0) is inside the image
    if (y0 < a.height) {
        if (x0 < a.width) inside |= 1u;
        if (x0 + 1u < a.width) inside |= 2u;
    }
    const float NEG_INF = __int_as_float(0xff800000);
    float u0 = (inside & 1u) ? 0.0f : NEG_INF, u1 = (inside & 2u) ? 0.0f : NEG_INF;
    float T0 = 1.0f, T1 = 1.0f, r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;
    bool reported = false;   // this quadrant has been counted as done
    for (uint32_t r = 0;; ++r) {
        WsbBuf& q = sm.buf[r & 1u];
        wsb_wait(&sm.full[r & 1u], (r >> 1) & 1u);
        const uint32_t my_ln = q.ln[warp];
        const bool last = q.last != 0u;
        const uint32_t ent_s = (uint32_t)__cvta_generic_to_shared(&q.ent[0]);
        const uint32_t list_s = (uint32_t)__cvta_generic_to_shared(&q.list[warp][0]);
#pragma unroll 2
        for (uint32_t j = 0; j < my_ln; ++j) {
            if ((j & 7u) == 0u && __all_sync(0xffffffffu, u0 < 0.0f && u1 < 0.0f)) break;  // the whole quadrant is done
            const uint32_t e = ent_s + lds_u16(list_s + 2u * j);
            const float4 q0 = lds_f4(e);
            const float4 q1 = lds_f4(e + 16u);
            const float dx0 = q0.x - fx0, dy = q0.y - fy0;
            const float dx1 = dx0 - 1.0f;
            const float by = q0.w * dy;
            const float cy = q1.x * dy * dy;
            const float p0 = fmaf(dx0, fmaf(q0.z, dx0, by), cy);   // A'dx^2 + B'dx dy + C'dy^2 = log2(e) * power
            const float p1 = fmaf(dx1, fmaf(q0.z, dx1, by), cy);
            const bool h0 = p0 <= u0 && p0 >= q1.z;    // below the threshold alpha < 1/255 (blend.slang:89)
            const bool h1 = p1 <= u1 && p1 >= q1.z;
            if (!(h0 || h1)) continue;
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
        }
        // block vote without a block barrier (blend.slang:56-63): the quadrant that completes the count raises the stop flag
        if (!reported && __all_sync(0xffffffffu, u0 < 0.0f && u1 < 0.0f)) {
            reported = true;
            if (lane == 0 && atomicAdd(&sm.done_quadrants, 1u) == BLEND_WARPS - 1u) {
                sm.rounds_used = r + 1u;
                *reinterpret_cast<volatile uint32_t*>(&sm.stop) = 1u;
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (last && sm.rounds_used == 0u) atomicMax(&sm.rounds_used, r + 1u);
            wsb_arrive(&sm.empty[r & 1u]);   // release: the fill warps may refill this buffer
        }
        if (last) break;
    }
    if (tid == 0) {   // hint for the next frames' dispatch order: the splats of the rounds the tile needed
        const uint32_t rounds = max(*reinterpret_cast<volatile uint32_t*>(&sm.rounds_used), 1u);
        a.tile_cost[tile] = max(min(rounds * BLEND_THREADS, range.y - range.x), 1u);
    }
    if (inside & 1u)
        *reinterpret_cast<uint32_t*>(a.out + (size_t)y0 * a.pitch + (size_t)x0 * 4) =
            unorm8(r0) | (unorm8(g0) << 8) | (unorm8(b0) << 16) | 0xff000000u;
    if (inside & 2u)
        *reinterpret_cast<uint32_t*>(a.out + (size_t)y0 * a.pitch + (size_t)(x0 + 1u) * 4) =
            unorm8(r1) | (unorm8(g1) << 8) | (unorm8(b1) << 16) | 0xff000000u;
}


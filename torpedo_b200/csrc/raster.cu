// raster.cu — tile-range identification and per-tile front-to-back alpha blending.
//
// Replaces range.slang:16-34 (+ the fillBuffer of GaussianEngine.cpp:844) and blend.slang:22-104
// (+ the R8G8B8A8_UNORM image store of the Vulkan target, GaussianEngine.cpp:316-319).
// Colour is tolerance-checked (1/255), so FMA contraction and ex2.approx are allowed here; ranges
// are integer work and bit-exact.
#include "common.cuh"

namespace tpdcu {

// ---------------------------------------------------------------------------------------------------
// ranges: ranges[tile] = (first, last+1) over the sorted keys; empty tiles stay (0,0)
// ---------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) ranges_kernel(RasterLaunch a) {
    const uint32_t n = a.plan->n;
    const uint64_t* __restrict__ keys = a.plan->final_sel ? a.keys[1] : a.keys[0];
    uint2* ranges = reinterpret_cast<uint2*>(a.ranges);
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const uint32_t curr = (uint32_t)(keys[idx] >> 32);
        if (idx == 0) {
            ranges[curr].x = 0;
        } else {
            const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
            if (curr != prev) {
                ranges[prev].y = idx;
                ranges[curr].x = idx;
            }
        }
        if (idx == n - 1) ranges[curr].y = idx + 1;
    }
}

cudaError_t launch_ranges(const RasterLaunch& a, cudaStream_t s) {
    if (a.capacity == 0) return cudaSuccess;
    uint32_t grid = (a.capacity + 255) / 256;
    if (grid > 148u * 16u) grid = 148u * 16u;
    ranges_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// blend: one CTA per 16x16 tile; 64 threads, each owning a 2x2 pixel block, so the per-splat shared-memory reads, loop
// overhead and the dx/dy terms of the quadratic form are shared by four pixels. Splats are staged through shared memory in
// sorted order; while staging, splats whose alpha >= 1/255 region (SplatGeo::ext_*) cannot reach the tile are dropped —
// every pixel would `continue` on them (blend.slang:85,89), so dropping them leaves the result unchanged while removing
// about a quarter of the (pixel, splat) evaluations. The conic is pre-scaled by -0.5*log2(e) at staging so that
// alpha = opacity * ex2(power2).
// ---------------------------------------------------------------------------------------------------

constexpr uint32_t BLEND_THREADS = 64;
constexpr uint32_t BLEND_QUEUE = 256;
constexpr float LOG2E = 1.4426950408889634f;

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

__global__ void __launch_bounds__(BLEND_THREADS) blend_kernel(RasterLaunch a) {
    __shared__ float4 s_g0[BLEND_QUEUE];   // px, py, A', B'
    __shared__ float4 s_g1[BLEND_QUEUE];   // C', opacity, power2 threshold, -
    __shared__ float4 s_col[BLEND_QUEUE];  // r, g, b
    __shared__ uint32_t s_cnt[BLEND_THREADS / 32];

    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX;
    const uint32_t tile = blockIdx.x;
    const uint32_t tile_x0 = (tile % gx) * TILE_PX, tile_y0 = (tile / gx) * TILE_PX;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t x0 = tile_x0 + 2u * (tid & 7u), y0 = tile_y0 + 2u * (tid >> 3);
    const float fx0 = (float)x0, fy0 = (float)y0;
    const float tile_fx0 = (float)tile_x0, tile_fy0 = (float)tile_y0;

    const uint32_t* __restrict__ vals = a.plan->final_sel ? a.vals[1] : a.vals[0];
    const uint2 range = reinterpret_cast<const uint2*>(a.ranges)[tile];
    const float4* __restrict__ geo4 = reinterpret_cast<const float4*>(a.geo);

    // pixel k = (x0 + (k & 1), y0 + (k >> 1)); `live` bit k: still accumulating
    uint32_t live = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k)
        if (x0 + (k & 1u) < a.width && y0 + (k >> 1) < a.height) live |= 1u << k;
    const uint32_t inside = live;
    float T[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
    float cr[4] = { 0.f, 0.f, 0.f, 0.f }, cg[4] = { 0.f, 0.f, 0.f, 0.f }, cb[4] = { 0.f, 0.f, 0.f, 0.f };

    uint32_t in = range.x;
    while (true) {
        // ---- fill: append the next splats that can touch this tile, in order -----------------------------------------
        uint32_t qn = 0;
        while (qn + BLEND_THREADS <= BLEND_QUEUE && in < range.y) {
            const uint32_t idx = in + tid;
            bool keep = false;
            uint32_t g = 0;
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
            if (idx < range.y) {
                g = __ldg(vals + idx);
                r0 = __ldg(geo4 + (size_t)g * 2);
                r1 = __ldg(geo4 + (size_t)g * 2 + 1);
                keep = (r0.x + r1.z >= tile_fx0) && (r0.x - r1.z <= tile_fx0 + 15.0f) && (r0.y + r1.w >= tile_fy0) &&
                       (r0.y - r1.w <= tile_fy0 + 15.0f);
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_cnt[warp] = __popc(ballot);
            __syncthreads();
            const uint32_t c0 = s_cnt[0], c1 = s_cnt[1];
            if (keep) {
                const uint32_t pos = qn + (warp ? c0 : 0u) + __popc(ballot & lanemask_lt());
                const float4 col = __ldg(a.color + g);
                s_g0[pos] = make_float4(r0.x, r0.y, (-0.5f * LOG2E) * r0.z, -LOG2E * r0.w);
                s_g1[pos] = make_float4((-0.5f * LOG2E) * r1.x, r1.y, -__log2f(255.0f * r1.y) - 0.01f, 0.0f);
                s_col[pos] = col;
            }
            qn += c0 + c1;
            in += BLEND_THREADS;
            __syncthreads();
        }

        // ---- drain: front-to-back compositing (blend.slang:77-100) for the four pixels of this thread ----------------
        for (uint32_t j = 0; j < qn && live; ++j) {
            const float4 g0 = s_g0[j];
            const float4 g1 = s_g1[j];
            const float dx0 = g0.x - fx0, dy0 = g0.y - fy0;
            const float dx1 = dx0 - 1.0f, dy1 = dy0 - 1.0f;
            const float ax0 = g0.z * dx0, ax1 = g0.z * dx1;
            const float by0 = g0.w * dy0, by1 = g0.w * dy1;
            const float cy0 = g1.x * dy0 * dy0, cy1 = g1.x * dy1 * dy1;
            float p[4];
            p[0] = fmaf(dx0, ax0 + by0, cy0);
            p[1] = fmaf(dx1, ax1 + by0, cy0);
            p[2] = fmaf(dx0, ax0 + by1, cy1);
            p[3] = fmaf(dx1, ax1 + by1, cy1);
            uint32_t hit = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                if (p[k] <= 0.0f && p[k] >= g1.z) hit |= 1u << k;
            hit &= live;
            if (hit == 0) continue;
            const float4 c = s_col[j];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                if (hit & (1u << k)) {
                    const float alpha = fminf(0.99f, g1.y * ex2_approx(p[k]));
                    if (alpha >= 1.0f / 255.0f) {
                        const float test_T = T[k] * (1.0f - alpha);
                        if (test_T < 0.0001f) {
                            live &= ~(1u << k);  // done; this splat is NOT added (blend.slang:92-95)
                        } else {
                            const float w = alpha * T[k];
                            cr[k] = fmaf(c.x, w, cr[k]);
                            cg[k] = fmaf(c.y, w, cg[k]);
                            cb[k] = fmaf(c.z, w, cb[k]);
                            T[k] = test_T;
                        }
                    }
                }
            }
        }
        // block vote (blend.slang:56-63); also the barrier that lets the queue be refilled
        const bool finished = in >= range.y;
        if (__syncthreads_and(live == 0) || finished) break;
    }

#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        if (inside & (1u << k)) {
            const uint32_t rgba = unorm8(cr[k]) | (unorm8(cg[k]) << 8) | (unorm8(cb[k]) << 16) | 0xff000000u;
            *reinterpret_cast<uint32_t*>(a.out + (size_t)(y0 + (k >> 1)) * a.pitch + (size_t)(x0 + (k & 1u)) * 4) = rgba;
        }
    }
}

cudaError_t launch_blend(const RasterLaunch& a, cudaStream_t s) {
    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX, gy = (a.height + TILE_PX - 1) / TILE_PX;
    if (gx * gy == 0) return cudaSuccess;
    blend_kernel<<<gx * gy, BLEND_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tpdcu

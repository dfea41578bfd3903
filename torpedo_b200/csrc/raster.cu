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
// blend: one CTA per 16x16 tile, one pixel per thread, 256-splat batches staged in shared memory
// ---------------------------------------------------------------------------------------------------

constexpr uint32_t BLEND_THREADS = TILE_PX * TILE_PX;

__device__ __forceinline__ uint32_t unorm8(float c) {
    // clamp to [0,1] (NaN -> 0), x255, round to nearest even: cvt.rni.sat.u8.f32 semantics
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint32_t)__float2int_rn(c * 255.0f);
}

__global__ void __launch_bounds__(BLEND_THREADS) blend_kernel(RasterLaunch a) {
    __shared__ float4 s_geo0[BLEND_THREADS];  // px, py, conic_a, conic_b
    __shared__ float2 s_geo1[BLEND_THREADS];  // conic_c, opacity
    __shared__ float4 s_col[BLEND_THREADS];   // r, g, b

    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX;
    const uint32_t tile = blockIdx.x;
    const uint32_t tile_x = tile % gx, tile_y = tile / gx;
    const uint32_t tid = threadIdx.x;
    const uint32_t pix_x = tile_x * TILE_PX + (tid & (TILE_PX - 1)), pix_y = tile_y * TILE_PX + (tid / TILE_PX);
    const bool inside = pix_x < a.width && pix_y < a.height;
    const float fpx = (float)pix_x, fpy = (float)pix_y;

    const uint32_t* __restrict__ vals = a.plan->final_sel ? a.vals[1] : a.vals[0];
    const uint2 range = reinterpret_cast<const uint2*>(a.ranges)[tile];
    const float4* __restrict__ recs = reinterpret_cast<const float4*>(a.recs);

    bool done = !inside;
    float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;

    for (uint32_t start = range.x; start < range.y; start += BLEND_THREADS) {
        if (__syncthreads_and(done)) break;  // blend.slang:56-63; also fences the previous batch's reads
        const uint32_t limit = min(BLEND_THREADS, range.y - start);
        if (tid < limit) {
            const uint32_t g = __ldg(vals + start + tid);
            const float4 r0 = __ldg(recs + (size_t)g * 3 + 0);
            const float4 r1 = __ldg(recs + (size_t)g * 3 + 1);
            const float4 r2 = __ldg(recs + (size_t)g * 3 + 2);
            s_geo0[tid] = r0;
            s_geo1[tid] = make_float2(r1.x, r1.y);
            s_col[tid] = r2;
        }
        __syncthreads();
        for (uint32_t j = 0; !done && j < limit; ++j) {
            const float4 g0 = s_geo0[j];
            const float2 g1 = s_geo1[j];
            const float dx = g0.x - fpx, dy = g0.y - fpy;
            const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, g1.y * __expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 c = s_col[j];
            const float w = alpha * T;
            cr += c.x * w; cg += c.y * w; cb += c.z * w;
            T = test_T;
        }
    }
    if (inside) {
        const uint32_t rgba = unorm8(cr) | (unorm8(cg) << 8) | (unorm8(cb) << 16) | 0xff000000u;
        *reinterpret_cast<uint32_t*>(a.out + (size_t)pix_y * a.pitch + (size_t)pix_x * 4) = rgba;
    }
}

cudaError_t launch_blend(const RasterLaunch& a, cudaStream_t s) {
    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX, gy = (a.height + TILE_PX - 1) / TILE_PX;
    if (gx * gy == 0) return cudaSuccess;
    blend_kernel<<<gx * gy, BLEND_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tpdcu

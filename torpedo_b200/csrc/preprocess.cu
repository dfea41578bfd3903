// preprocess.cu — scene compilation, per-frame camera setup, and the fused
// project + exclusive-scan + tile-key duplication kernel.
//
// COMPILED WITH --fmad=false: everything that feeds a key, a tile count or a range must be evaluated in
// fp32, in the written order, with no contraction, exactly like oracle/tpd_oracle.c (SURVEY.md §8a
// "bit-critical chain"). sqrt, division and ceil are IEEE-correct with nvcc's defaults.
//
// Replaces, in one launch per frame and with no host read-back of P:
//   project.slang:33-106  (cull, view/clip transform, EWA covariance, conic, radius, rect, SH colour)
//   prefix.slang:38-159   (exclusive scan of the tile counts; decoupled look-back)
//   keygen.slang:21-53    (one (tile|depth, index) pair per overlapped tile)
//   GaussianEngine.cpp:662-674 (the mid-frame fence wait + read of tilesRendered is gone)
#include "common.cuh"

#include <algorithm>

namespace tpdcu {

// ---------------------------------------------------------------------------------------------------
// scene compilation: 240-byte AoS records -> planar arrays + camera-independent 3D covariance
// (splat/volume.slang:20-41 computeCovariance, evaluated once instead of once per frame)
// ---------------------------------------------------------------------------------------------------

constexpr uint32_t COMPILE_THREADS = 128;

__global__ void __launch_bounds__(COMPILE_THREADS) compile_scene_kernel(CompileLaunch a) {
    __shared__ float4 stage[COMPILE_THREADS * 15];  // 128 records x 240 B
    const uint32_t base = blockIdx.x * COMPILE_THREADS;
    const uint32_t count = min(COMPILE_THREADS, a.n - base);
    const float4* src = reinterpret_cast<const float4*>(a.recs240) + (size_t)base * 15;
    for (uint32_t k = threadIdx.x; k < count * 15; k += COMPILE_THREADS) stage[k] = src[k];
    __syncthreads();
    if (threadIdx.x >= count) return;
    const float* g = reinterpret_cast<const float*>(stage) + threadIdx.x * 60;
    const uint32_t i = base + threadIdx.x;

    const float qx = g[4], qy = g[5], qz = g[6], qw = g[7];
    const float sx = g[8] * g[11], sy = g[9] * g[11], sz = g[10] * g[11];
    const float R[9] = {
        1.0f - 2.0f * (qy * qy + qz * qz), 2.0f * (qx * qy - qw * qz),        2.0f * (qx * qz + qw * qy),
        2.0f * (qx * qy + qw * qz),        1.0f - 2.0f * (qx * qx + qz * qz), 2.0f * (qy * qz - qw * qx),
        2.0f * (qx * qz - qw * qy),        2.0f * (qy * qz + qw * qx),        1.0f - 2.0f * (qx * qx + qy * qy),
    };
    float sg[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        sg[r * 3 + 0] = R[r * 3 + 0] * sx;
        sg[r * 3 + 1] = R[r * 3 + 1] * sy;
        sg[r * 3 + 2] = R[r * 3 + 2] * sz;
    }
    auto cov = [&](int r, int c) {
        return (sg[r * 3 + 0] * sg[c * 3 + 0] + sg[r * 3 + 1] * sg[c * 3 + 1]) + sg[r * 3 + 2] * sg[c * 3 + 2];
    };
    a.posop[i] = make_float4(g[0], g[1], g[2], g[3]);
    a.cov_a[i] = make_float4(cov(0, 0), cov(0, 1), cov(0, 2), cov(1, 1));
    a.cov_b[i] = make_float2(cov(1, 2), cov(2, 2));

    // SH: reference layout is DC rgb then 15 R, 15 G, 15 B (splat/common.slang:25-31); internal layout is one 192-byte
    // row per Gaussian of coefficient-major rgb triplets, so that degree d touches the first ceil(3*(d+1)^2/4) float4 of
    // the row only and a row is fetched with whole 32-byte sectors whatever its neighbours' visibility.
    const float* sh = g + 12;
#pragma unroll
    for (uint32_t p = 0; p < SH_PLANES; ++p) {
        float v[4];
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t f = p * 4 + q, k = f / 3, c = f % 3;
            v[q] = (k == 0) ? sh[c] : sh[3 + c * 15 + (k - 1)];
        }
        a.sh[(size_t)i * SH_PLANES + p] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

cudaError_t launch_compile_scene(const CompileLaunch& a, cudaStream_t s) {
    if (a.n == 0) return cudaSuccess;
    compile_scene_kernel<<<(a.n + COMPILE_THREADS - 1) / COMPILE_THREADS, COMPILE_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// per-frame setup: mul(V, M), mul(P, M) per entity (splat/common.slang:102,106), camera position,
// focal length in pixels. Row-major, every dot product accumulated left to right.
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void mat4_mul(const float* a, const float* b, float* r) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            r[i * 4 + j] = ((a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j]) + a[i * 4 + 2] * b[2 * 4 + j]) +
                           a[i * 4 + 3] * b[3 * 4 + j];
}

__global__ void setup_kernel(PreprocessLaunch a, const __grid_constant__ CameraUbo ubo) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const float* V = ubo.f;
    const float* P = ubo.f + 16;
    if (e < a.scene.entity_count) {
        float m[16], v[16], p[16], r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { m[k] = a.models[e * 16 + k]; v[k] = V[k]; p[k] = P[k]; }
        mat4_mul(v, m, r);
#pragma unroll
        for (int k = 0; k < 16; ++k) a.vm[e * 16 + k] = r[k];
        mat4_mul(p, m, r);
#pragma unroll
        for (int k = 0; k < 16; ++k) a.pm[e * 16 + k] = r[k];
    }
    if (e == 0) {
        for (int k = 0; k < 16; ++k) a.cam->V[k] = V[k];
        for (int i = 0; i < 3; ++i)
            a.cam->cam_pos[i] = -((V[0 * 4 + i] * V[0 * 4 + 3] + V[1 * 4 + i] * V[1 * 4 + 3]) + V[2 * 4 + i] * V[2 * 4 + 3]);
        a.cam->focal[0] = 0.5f * (float)a.width * ubo.f[32];
        a.cam->focal[1] = 0.5f * (float)a.height * ubo.f[33];
    }
}

cudaError_t launch_setup(const PreprocessLaunch& a, const CameraUbo& ubo, cudaStream_t s) {
    const uint32_t threads = 64;
    setup_kernel<<<(a.scene.entity_count + threads - 1) / threads, threads, 0, s>>>(a, ubo);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// geometry + scan + visible compaction kernel (no colour: the blend evaluates it on demand, raster.cu)
// ---------------------------------------------------------------------------------------------------

// Scan descriptor: [63:62] state | [61:32] visible-Gaussian count | [31:0] pair count.
__device__ __forceinline__ uint64_t desc_pack(uint32_t flag, uint32_t visible, uint32_t pairs) {
    return ((uint64_t)flag << 62) | ((uint64_t)visible << 32) | (uint64_t)pairs;
}
constexpr uint64_t DESC_VALUE_MASK = (1ull << 62) - 1ull;

// Decoupled look-back of a chained scan (one warp): exclusive prefix of partition `part` (> 0) over descriptors that carry
// [63:62] state | [61:0] value. 32 predecessors are inspected per round.
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t* desc, uint32_t part, uint32_t lane) {
    uint64_t exclusive = 0;
    int look = (int)part - 1;
    while (true) {
        const int idx = look - (int)lane;
        uint64_t d = ((uint64_t)FLAG_PREFIX << 62);  // virtual partition -1: inclusive prefix 0
        if (idx >= 0) {
            do { d = ld_relaxed_u64(desc + idx); } while ((d >> 62) == FLAG_INVALID);
        }
        const uint32_t prefix_mask = __ballot_sync(0xffffffffu, (d >> 62) == FLAG_PREFIX);
        const uint32_t first = prefix_mask ? (uint32_t)__ffs((int)prefix_mask) - 1u : 32u;
        uint64_t v = (lane <= first) ? (d & DESC_VALUE_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        exclusive += v;
        if (prefix_mask) break;
        look -= 32;
    }
    return exclusive;
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Projected {
    uint32_t count, depth_bits;
};

// project.slang:33-90 for one Gaussian (colour excluded). Writes SplatGeo / depth_radius / rect when visible.
__device__ __forceinline__ Projected project_one(const PreprocessLaunch& a, uint32_t i, const float4 po, const float4 ca, const float2 cb,
                                                 const float* VM, const float* PM, const float* V, const float* focal, uint32_t gx,
                                                 uint32_t gy) {
    Projected out{ 0u, 0u };
    // splat/common.slang:98-119 passFrustumClipping
    auto row_point = [&](const float* m, int r) {
        return ((m[r * 4 + 0] * po.x + m[r * 4 + 1] * po.y) + m[r * 4 + 2] * po.z) + m[r * 4 + 3];
    };
    const float vx = row_point(VM, 0), vy = row_point(VM, 1), vz = row_point(VM, 2);
    if (vz <= 0.0f) return out;
    const float cx4 = row_point(PM, 0), cy4 = row_point(PM, 1), cz4 = row_point(PM, 2), cw4 = row_point(PM, 3);
    if (cx4 < -1.3f * cw4 || cx4 > 1.3f * cw4) return out;
    if (cy4 < -1.3f * cw4 || cy4 > 1.3f * cw4) return out;
    if (cz4 < 0.0f || cz4 > cw4) return out;
    const float inv_w = 1.0f / cw4;
    const float proj_x = cx4 * inv_w, proj_y = cy4 * inv_w;
    const float cov3[9] = { ca.x, ca.y, ca.z, ca.y, ca.w, cb.x, ca.z, cb.x, cb.y };

    // splat/volume.slang:44-63
    const float fx = focal[0] / vz, fy = focal[1] / vz;
    const float tx = vx / vz, ty = vy / vz;
    const float j02 = -fx * tx, j12 = -fy * ty;
    float T[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        T[0 * 3 + j] = fx * V[0 * 4 + j] + j02 * V[2 * 4 + j];
        T[1 * 3 + j] = fy * V[1 * 4 + j] + j12 * V[2 * 4 + j];
    }
    float M[6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            M[r * 2 + j] = (cov3[r * 3 + 0] * T[j * 3 + 0] + cov3[r * 3 + 1] * T[j * 3 + 1]) + cov3[r * 3 + 2] * T[j * 3 + 2];
    const float c00 = (T[0] * M[0] + T[1] * M[2]) + T[2] * M[4];
    const float c10 = (T[3] * M[0] + T[4] * M[2]) + T[5] * M[4];
    const float c11 = (T[3] * M[1] + T[4] * M[3]) + T[5] * M[5];

    // project.slang:60-72
    const float cvx = c00 + 0.3f, cvy = c10, cvz = c11 + 0.3f;
    const float det = cvx * cvz - cvy * cvy;
    if (det == 0.0f) return out;
    const float det_inv = 1.0f / det;
    const float mid = 0.5f * (cvx + cvz);
    const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda_1 = mid + sq, lambda_2 = mid - sq;
    const float radius = ceilf(3.0f * sqrtf(fmaxf(lambda_1, lambda_2)));
    if (!(fabsf(radius) <= 3.402823466e+38f)) return out;  // canonical: a non-finite radius is culled

    // splat/volume.slang:3-17
    const float px = ((proj_x + 1.0f) * (float)a.width - 1.0f) * 0.5f;
    const float py = ((proj_y + 1.0f) * (float)a.height - 1.0f) * 0.5f;
    const int x0 = min((int)gx, max(0, __float2int_rz((px - radius) / 16.0f)));
    const int y0 = min((int)gy, max(0, __float2int_rz((py - radius) / 16.0f)));
    const int x1 = min((int)gx, max(0, __float2int_rz((((px + radius) + 16.0f) - 1.0f) / 16.0f)));
    const int y1 = min((int)gy, max(0, __float2int_rz((((py + radius) + 16.0f) - 1.0f) / 16.0f)));
    out.count = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
    if (out.count == 0) return out;
    out.depth_bits = __float_as_uint(vz);
    a.out.rect[i] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)(x1 - x0) | ((uint32_t)(y1 - y0) << 16));

    // Conservative half-extents of the region where alpha = opacity*exp(power) can reach 1/255 (blend.slang:88-89):
    // power >= -t, t = ln(255*opacity), is the ellipse d^T conic d <= 2t whose bounding box is sqrt(2t*cov). Used ONLY to
    // skip splats that cannot touch a tile; the margins cover fp32 rounding of power/exp in the blend. Ill-conditioned
    // covariances are never culled.
    // (approximate log and square roots: their errors, ~1e-6 relative, are far inside the margins)
    float ext_x = 3.0e38f, ext_y = 3.0e38f;
    const float t = __logf(255.0f * po.w);
    if (!(t >= 0.0f)) {
        ext_x = ext_y = -1.0e30f;  // opacity < 1/255: alpha < 1/255 at every pixel
    } else if (det > 1e-3f * (cvx * cvz) && cvx > 0.0f && cvz > 0.0f) {
        const float tt = 2.0f * (t + 0.02f);
        ext_x = sqrt_approx(tt * cvx) * 1.0002f + 0.02f;
        ext_y = sqrt_approx(tt * cvz) * 1.0002f + 0.02f;
    }
    stg256(reinterpret_cast<float4*>(a.out.geo + i), make_float4(px, py, cvz * det_inv, -cvy * det_inv),
           make_float4(cvx * det_inv, po.w, ext_x, ext_y));  // the 32-byte SplatGeo record, one sector, one store
    a.out.depth_radius[i] = make_float2(vz, radius);
    return out;
}

// One CTA = one partition of PRE_PART consecutive Gaussians (PRE_ITEMS per thread, striped so that loads coalesce).
// Besides the per-Gaussian records it compacts the visible Gaussians, in index order, into `depth_words`
// (float_bits(viewZ) << 32 | index): the input of the depth sort. The (visible, pairs) scan also yields the reference's
// pair offsets (prefix.slang) and P.
#ifndef TPDCU_PRE_MINB
#define TPDCU_PRE_MINB 4
#endif
// A resident CTA works through partitions of PRE_PART consecutive Gaussians (PRE_ITEMS per thread, striped so that loads
// coalesce), one ticket at a time. Besides the per-Gaussian records it compacts the visible Gaussians, in index order, into
// `depth_words` (float_bits(viewZ) << 32 | index): the input of the depth sort. The (visible, pairs) scan also yields the
// reference's pair offsets (prefix.slang) and P.
//
// The look-back of a partition is DEFERRED behind the geometry of the CTA's next partition. A partition's geometry takes
// 2-10 us depending on how many of its Gaussians survive the culls, and with one partition per CTA every CTA sat on its SM
// slot until the slowest of its 32 predecessors had published an aggregate (ncu: 35 % of all stall samples behind that
// barrier, 121 polls of the descriptor window per partition). Now the aggregate is published as soon as it is known, the CTA
// goes on with the next partition, and only then resolves the previous one's prefix: every predecessor has published by then,
// and as every CTA defers alike the distance to the nearest PREFIX stays what it was.
__global__ void __launch_bounds__(PRE_THREADS, TPDCU_PRE_MINB) preprocess_kernel(PreprocessLaunch a) {
    pdl_wait();      // (its predecessor in the stream is a memset: a no-op) ...
    pdl_release();   // ... but the depth sort's histogram kernel may be scheduled as this grid's CTAs leave
    __shared__ uint32_t s_part;
    __shared__ uint64_t s_base;                 // exclusive (visible, pairs) prefix of the partition being completed
    __shared__ float s_vm[16], s_pm[16], s_v[12], s_focal[2];
    __shared__ uint64_t s_warp_tot[2][PRE_ITEMS][PRE_THREADS / 32];   // [parity of the partition's turn in this CTA]
    __shared__ uint64_t s_total[2];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_part = atomicAdd(&a.ctl->scan_ticket, 1u);
    const bool single_entity = a.scene.entity == nullptr;
    if (tid < 16) {
        if (single_entity) { s_vm[tid] = a.vm[tid]; s_pm[tid] = a.pm[tid]; }
        if (tid < 12) s_v[tid] = a.cam->V[tid];
        if (tid < 2) s_focal[tid] = a.cam->focal[tid];
    }
    __syncthreads();
    const uint32_t n = a.scene.n;
    const uint32_t num_parts = (n + PRE_PART - 1) / PRE_PART;
    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX, gy = (a.height + TILE_PX - 1) / TILE_PX;
    uint32_t part = s_part;
    __syncthreads();   // thread 0 rewrites s_part (the next ticket) inside the first turn, before that turn's first barrier

    // what a partition leaves behind for its completion, one turn later: per Gaussian its tile count, its depth bits and the
    // exclusive (visible, pairs) prefix inside its warp's group — parked in shared memory, the registers belong to the geometry
    __shared__ uint4 s_stash[2][PRE_ITEMS][PRE_THREADS];
    uint32_t prev_part = 0xffffffffu, prev_par = 0;

    // completion of partition `prev_part` (its turn had parity `par`): look-back, then the prefix-dependent stores
    auto complete = [&](uint32_t par) {
        if (warp == 0) {
            const uint64_t total = s_total[par];
            uint64_t exclusive = 0;
            if (prev_part != 0) {
                exclusive = lookback_exclusive(a.scan_desc, prev_part, lane);
                if (lane == 0) st_relaxed_u64(a.scan_desc + prev_part, ((uint64_t)FLAG_PREFIX << 62) | (exclusive + total));
            }
            if (lane == 0) {
                s_base = exclusive;
                if (prev_part == num_parts - 1) {
                    const uint64_t all = exclusive + total;
                    a.ctl->pairs_total = (uint32_t)all;
                    a.ctl->visible = (uint32_t)(all >> 32);
                    a.out.offsets[n] = (uint32_t)all;
                }
            }
        }
        __syncthreads();
        const uint64_t base = s_base;
        const uint32_t first = prev_part * PRE_PART + tid;
#pragma unroll
        for (uint32_t k = 0; k < PRE_ITEMS; ++k) {
            const uint32_t i = first + k * PRE_THREADS;
            if (i < n) {
                const uint4 st = s_stash[par][k][tid];
                const uint64_t at = base + s_warp_tot[par][k][warp] + (((uint64_t)st.w << 32) | st.z);
                a.out.offsets[i] = (uint32_t)at;
                if (st.x != 0) a.depth_words[(uint32_t)(at >> 32)] = ((uint64_t)st.y << 32) | i;
            }
        }
    };

    for (uint32_t turn = 0; part < num_parts; ++turn) {
        const uint32_t par = turn & 1u;
        const uint32_t first = part * PRE_PART + tid;

        // ---- geometry, in batches of PRE_BATCH Gaussians per thread: all loads of a batch first, then its arithmetic. Only
        // the tile count and the depth bits of a Gaussian stay in registers (project_one stores the records).
        Projected pr[PRE_ITEMS];
#pragma unroll
        for (uint32_t b0 = 0; b0 < PRE_ITEMS; b0 += PRE_BATCH) {
            float4 po[PRE_BATCH], ca[PRE_BATCH];
            float2 cb[PRE_BATCH];
#pragma unroll
            for (uint32_t k = 0; k < PRE_BATCH; ++k) {
                const uint32_t i = first + (b0 + k) * PRE_THREADS;
                if (i < n) {
                    po[k] = __ldg(a.scene.posop + i);
                    ca[k] = __ldg(a.scene.cov_a + i);
                    cb[k] = __ldg(a.scene.cov_b + i);
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < PRE_BATCH; ++k) {
                const uint32_t i = first + (b0 + k) * PRE_THREADS;
                pr[b0 + k] = Projected{ 0u, 0u };
                if (i < n) {
                    if (single_entity) {
                        pr[b0 + k] = project_one(a, i, po[k], ca[k], cb[k], s_vm, s_pm, s_v, s_focal, gx, gy);
                    } else {
                        float vm_l[16], pm_l[16];
                        // an index past the entity table (only possible through tpdcu_upload_gaussians_device, whose indices are
                        // not validated on the host) must not read outside the matrices
                        const uint32_t e = min(__ldg(a.scene.entity + i), a.scene.entity_count - 1u);
#pragma unroll
                        for (int q = 0; q < 16; ++q) { vm_l[q] = __ldg(a.vm + e * 16 + q); pm_l[q] = __ldg(a.pm + e * 16 + q); }
                        pr[b0 + k] = project_one(a, i, po[k], ca[k], cb[k], vm_l, pm_l, s_v, s_focal, gx, gy);
                    }
                }
            }
        }

        // ---- depth range of the frame (the depth sort only sorts the bits this range occupies, sort.cu sort_spec) ---------
        {
            uint32_t dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
            for (uint32_t k = 0; k < PRE_ITEMS; ++k)
                if (pr[k].count != 0) { dmin = min(dmin, pr[k].depth_bits); dmax = max(dmax, pr[k].depth_bits); }
            dmin = __reduce_min_sync(0xffffffffu, dmin);
            dmax = __reduce_max_sync(0xffffffffu, dmax);
            if (lane == 0 && dmax >= dmin) {  // after the first partitions the range rarely widens: test before the atomic
                if (dmax > ld_relaxed_u32(&a.ctl->depth_max)) atomicMax(&a.ctl->depth_max, dmax);
                if (~dmin > ld_relaxed_u32(&a.ctl->inv_depth_min)) atomicMax(&a.ctl->inv_depth_min, ~dmin);
            }
        }

        // ---- partition-local exclusive scan of (visible, pairs); group k = Gaussians [k*256, (k+1)*256) of the partition ----
#pragma unroll
        for (uint32_t k = 0; k < PRE_ITEMS; ++k) {
            const uint64_t mine = ((uint64_t)(pr[k].count != 0) << 32) | pr[k].count;
            uint64_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += up;
            }
            if (lane == 31) s_warp_tot[par][k][warp] = incl;
            const uint64_t excl = incl - mine;
            s_stash[par][k][tid] = make_uint4(pr[k].count, pr[k].depth_bits, (uint32_t)excl, (uint32_t)(excl >> 32));
        }
        if (tid == 0) s_part = atomicAdd(&a.ctl->scan_ticket, 1u);   // the next ticket travels during the rest of this turn
        __syncthreads();

        // ---- warp 0: exclusive scan of the PRE_ITEMS x 8 (group, warp) totals (group-major: the order of the Gaussians), and
        // the partition's aggregate goes out at once ------------------------------------------------------------------------
        if (warp == 0) {
            static_assert(PRE_ITEMS * (PRE_THREADS / 32) == 32, "one lane per (group, warp) total");
            const uint64_t mine_tot = (&s_warp_tot[par][0][0])[lane];
            uint64_t incl_tot = mine_tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t up = __shfl_up_sync(0xffffffffu, incl_tot, d);
                if (lane >= (uint32_t)d) incl_tot += up;
            }
            (&s_warp_tot[par][0][0])[lane] = incl_tot - mine_tot;  // exclusive prefix of this (group, warp) inside the partition
            const uint64_t total = __shfl_sync(0xffffffffu, incl_tot, 31);
            if (lane == 0) {
                s_total[par] = total;
                st_relaxed_u64(a.scan_desc + part, ((uint64_t)(part == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) << 62) | total);
                // the packed (visible | pairs) scan carries into the visible count once P reaches 2^32: keep an exact 64-bit P
                // beside it, so that the host can refuse such a frame instead of trusting a wrapped count
                atomicAdd(&a.ctl->pairs64, (unsigned long long)(uint32_t)total);
            }
        }

        // ---- the previous partition of this CTA: every predecessor of it has long published ------------------------------
        if (prev_part != 0xffffffffu) complete(par ^ 1u);   // (its barrier also orders warp 0's writes above before the reads below)
        else __syncthreads();
        prev_part = part;
        prev_par = par;
        part = s_part;
        __syncthreads();   // s_part is rewritten in the next turn, s_base by the next completion
    }
    if (prev_part != 0xffffffffu) complete(prev_par);
}

cudaError_t launch_preprocess(const PreprocessLaunch& a, cudaStream_t s) {
    if (a.scene.n == 0) return cudaSuccess;
    // persistent: as many CTAs as stay resident, each drawing partitions by ticket
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count = 148;
    }
    const uint32_t parts = (a.scene.n + PRE_PART - 1) / PRE_PART;
    preprocess_kernel<<<std::min<uint32_t>(parts, (uint32_t)sm_count * TPDCU_PRE_MINB), PRE_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// duplication (keygen.slang:47-53) over the DEPTH-SORTED visible Gaussians
// ---------------------------------------------------------------------------------------------------
//
// The reference emits the pairs of Gaussian i at offsets[i] (index order) and sorts them by tile << 32 | depth. Here the
// visible Gaussians arrive sorted by (depth, index); emitting their pairs in that order makes the remaining work a stable
// sort by tile only. One CTA = EMIT_PART consecutive sorted Gaussians: gather their rectangles, scan the tile counts
// (chained look-back across CTAs for the global offset), then emit the CTA's pairs cooperatively.

constexpr uint32_t EMIT_THREADS = 256;
constexpr uint32_t EMIT_ITEMS = 4;
constexpr uint32_t EMIT_PART = EMIT_THREADS * EMIT_ITEMS;

#ifndef TPDCU_EMIT_GROUP
#define TPDCU_EMIT_GROUP 4
#endif
#ifndef TPDCU_EMIT_MINB
#define TPDCU_EMIT_MINB 4
#endif
// Persistent, like preprocess_kernel: a CTA gathers and scans its next partition, publishes that aggregate, and only then
// resolves the look-back of its previous partition and emits that partition's pairs — by then every predecessor has published
// (ncu before: barrier stall 7.9 warps per issue behind the one-warp look-back).
__global__ void __launch_bounds__(EMIT_THREADS, TPDCU_EMIT_MINB) emit_kernel(EmitLaunch a) {
    pdl_wait();
    pdl_release();
    __shared__ uint32_t s_part;
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_off[2][EMIT_PART];    // exclusive pair offsets inside the partition     [parity of the CTA's turn]
    __shared__ uint32_t s_xy[2][EMIT_PART];     // rect origin: x0 | y0 << 16
    __shared__ uint32_t s_w[2][EMIT_PART];      // rect width | top depth bits << 16 (DepthSplit::extra of them)
    __shared__ uint32_t s_id[2][EMIT_PART];     // Gaussian index
    __shared__ uint32_t s_warp_tot[EMIT_ITEMS][EMIT_THREADS / 32];
    __shared__ uint32_t s_total[2];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_part = atomicAdd(&a.ctl->emit_ticket, 1u);
    __syncthreads();
    uint32_t part = s_part;
    __syncthreads();   // thread 0 rewrites s_part (the next ticket) inside the first turn, before that turn's first barrier
    const uint32_t visible = a.depth_plan->n;
    const uint32_t num_parts = (visible + EMIT_PART - 1) / EMIT_PART;
    const uint64_t* __restrict__ sorted = a.depth_plan->final_sel ? a.depth_words[1] : a.depth_words[0];
    const uint32_t gx = (a.width + TILE_PX - 1) / TILE_PX;
    const DepthSplit ds = depth_split(a.ctl, a.tile_bits);
    uint32_t prev_part = 0xffffffffu, prev_par = 0;

    // look-back + emission of partition `prev_part`, whose arrays sit in buffer `par`
    auto complete = [&](uint32_t par) {
        const uint32_t total = s_total[par];
        if (warp == 0) {
            uint64_t exclusive = 0;
            if (prev_part != 0) {
                exclusive = lookback_exclusive(a.scan_desc, prev_part, lane);
                if (lane == 0) st_relaxed_u64(a.scan_desc + prev_part, ((uint64_t)FLAG_PREFIX << 62) | (exclusive + total));
            }
            if (lane == 0) s_base = exclusive;
        }
        __syncthreads();
        const uint32_t base = (uint32_t)s_base;
        const uint32_t* off = s_off[par];
        const uint32_t* sw = s_w[par];
        const uint32_t* sxy = s_xy[par];
        const uint32_t* sid = s_id[par];
        // ---- emission: each thread takes groups of four CONSECUTIVE global slots (aligned to 4, so a full group is one
        // 32-byte store): one binary search finds the Gaussian owning the first slot, the next slots walk forward — the next
        // tile of the same rectangle (x+1, wrapping to the next row) or the first tile of the next Gaussian. The reference
        // loops serially per Gaussian; here big and small splats cost the same per pair.
        const uint32_t slot_end = base + total;
        constexpr uint32_t GROUP = TPDCU_EMIT_GROUP;   // consecutive slots per thread and search: whole 32-byte stores
        for (uint32_t G0 = (base & ~(GROUP - 1u)) + GROUP * tid; G0 < slot_end; G0 += GROUP * EMIT_THREADS) {
            const uint32_t lo = max(G0, base), hi = min(G0 + GROUP, slot_end);  // valid global slots of this group: [lo, hi)
            const uint32_t j = lo - base;
            uint32_t g = 0;
#pragma unroll
            for (uint32_t step = EMIT_PART / 2; step >= 1; step >>= 1)
                if (off[g + step] <= j) g += step;
            uint32_t wd = sw[g], xy = sxy[g], id = sid[g];
            uint32_t w = wd & 0xffffu;
            const uint32_t r = j - off[g];
            // r / w without the integer-division sequence: the quotient is a row of the rectangle (< 65536), so the float
            // quotient is off by at most one
            uint32_t ry = (uint32_t)__fdividef((float)r, (float)w);
            if (ry * w > r) --ry;
            else if ((ry + 1u) * w <= r) ++ry;
            uint32_t rx = r - ry * w;
            uint32_t row = ((xy >> 16) + ry) * gx + (xy & 0xffffu);  // tile id of the rectangle's column 0 in the current row
            uint32_t next_off = g + 1 < EMIT_PART ? off[g + 1] : 0xffffffffu;
            uint64_t key[GROUP];
#pragma unroll
            for (uint32_t q = 0; q < GROUP; ++q) {
                const uint32_t G = G0 + q;
                if (G >= lo && G < hi) {
                    if (G > lo) {
                        const uint32_t jq = G - base;
                        if (jq >= next_off) {  // first tile of the next Gaussian (padding slots of the last partition share their offset)
                            do {
                                ++g;
                                next_off = g + 1 < EMIT_PART ? off[g + 1] : 0xffffffffu;
                            } while (jq >= next_off);
                            wd = sw[g]; xy = sxy[g]; id = sid[g];
                            w = wd & 0xffffu;
                            rx = 0;
                            row = (xy >> 16) * gx + (xy & 0xffffu);
                        } else if (++rx == w) {
                            rx = 0;
                            row += gx;
                        }
                    }
                    key[q] = ((uint64_t)(((row + rx) << ds.extra) | (wd >> 16)) << 32) | id;
                }
            }
            if (lo == G0 && hi == G0 + GROUP && G0 + GROUP <= a.capacity) {
#pragma unroll
                for (uint32_t q = 0; q < GROUP; q += 4) stg256(a.keys + G0 + q, key[q], key[q + 1], key[q + 2], key[q + 3]);
            } else {
#pragma unroll
                for (uint32_t q = 0; q < GROUP; ++q) {
                    const uint32_t G = G0 + q;
                    if (G >= lo && G < hi && G < a.capacity) a.keys[G] = key[q];
                }
            }
        }
    };

    for (uint32_t turn = 0; part < num_parts; ++turn) {
        const uint32_t par = turn & 1u;
        // ---- gather: striped loads (coalesced), partition order = slot order ---------------------------------------------
        uint32_t cnt[EMIT_ITEMS];
#pragma unroll
        for (uint32_t k = 0; k < EMIT_ITEMS; ++k) {
            const uint32_t slot = k * EMIT_THREADS + tid;
            const uint32_t r = part * EMIT_PART + slot;
            uint32_t id = 0, dtop = 0;
            uint2 rc = make_uint2(0u, 0u);
            if (r < visible) {
                const uint64_t word = __ldg(sorted + r);
                id = (uint32_t)word;
                if (ds.extra) dtop = ((uint32_t)(word >> 32) - ds.bias) >> ds.low_bits;
                rc = __ldg(a.rect + id);
            }
            cnt[k] = (rc.y & 0xffffu) * (rc.y >> 16);
            s_xy[par][slot] = rc.x;
            s_w[par][slot] = (rc.y & 0xffffu) | (dtop << 16);
            s_id[par][slot] = id;
        }

        // ---- partition-local exclusive scan of the tile counts; group k = slots [k*256, (k+1)*256) -------------------------
        uint32_t incl[EMIT_ITEMS];
#pragma unroll
        for (uint32_t k = 0; k < EMIT_ITEMS; ++k) {
            incl[k] = cnt[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl[k], d);
                if (lane >= (uint32_t)d) incl[k] += up;
            }
            if (lane == 31) s_warp_tot[k][warp] = incl[k];
        }
        if (tid == 0) s_part = atomicAdd(&a.ctl->emit_ticket, 1u);   // the next ticket travels during the rest of this turn
        __syncthreads();
        uint32_t total = 0;
#pragma unroll
        for (uint32_t k = 0; k < EMIT_ITEMS; ++k) {
            uint32_t warp_excl = 0, group_total = 0;
#pragma unroll
            for (uint32_t w = 0; w < EMIT_THREADS / 32; ++w) {
                const uint32_t t = s_warp_tot[k][w];
                if (w < warp) warp_excl += t;
                group_total += t;
            }
            s_off[par][k * EMIT_THREADS + tid] = total + warp_excl + incl[k] - cnt[k];
            total += group_total;
        }
        if (tid == 0) {   // the partition's aggregate goes out at once
            s_total[par] = total;
            st_relaxed_u64(a.scan_desc + part, ((uint64_t)(part == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) << 62) | total);
        }

        // ---- the previous partition of this CTA: every predecessor of it has long published ------------------------------
        __syncthreads();   // s_total / s_off of this turn are complete; s_warp_tot may be rewritten
        if (prev_part != 0xffffffffu) complete(par ^ 1u);
        prev_part = part;
        prev_par = par;
        part = s_part;
        __syncthreads();   // s_part is rewritten in the next turn, s_base by the next completion, buffer par ^ 1 by the next gather
    }
    if (prev_part != 0xffffffffu) complete(prev_par);
}

uint32_t emit_parts(uint32_t n) { return (n + EMIT_PART - 1) / EMIT_PART; }

cudaError_t launch_emit(const EmitLaunch& a, cudaStream_t s) {
    if (a.n == 0) return cudaSuccess;
    // persistent: as many CTAs as stay resident, each drawing partitions by ticket (the visible count lives on the device)
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count = 148;
    }
    pdl_launch(emit_kernel, std::min<uint32_t>(emit_parts(a.n), (uint32_t)sm_count * TPDCU_EMIT_MINB), EMIT_THREADS, 0, s, a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// colour of EVERY visible Gaussian (project.slang:82-83, splat/common.slang:35-80): introspection only (tpdcu_read_splats).
// A frame evaluates the colour inside the blend, for the splats it stages (raster.cu); this kernel was the frame's colour
// stage before that and ran at the HBM roofline: a pure stream of 16 B position + up to 192 B of SH in, 16 B out, every
// visible Gaussian's SH row pulled into shared memory by ONE bulk-copy (TMA) instruction issued by its own thread — exact
// 32-byte sectors, no bytes fetched for culled neighbours.
// ---------------------------------------------------------------------------------------------------

constexpr uint32_t COLOR_THREADS = 128;
constexpr uint32_t COLOR_ROW_F4 = SH_PLANES + 1;  // odd stride in 16-byte units: the per-thread 128-bit row reads are conflict-free

template <int DEG>
__global__ void __launch_bounds__(COLOR_THREADS) color_kernel(PreprocessLaunch a) {
    constexpr int PLANES = DEG == 0 ? 1 : DEG == 1 ? 3 : DEG == 2 ? 7 : 12;
    __shared__ __align__(128) float4 rows[COLOR_THREADS * COLOR_ROW_F4];
    __shared__ __align__(8) uint64_t mbar[COLOR_THREADS / 32];
    const uint32_t n = a.scene.n;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t i = blockIdx.x * COLOR_THREADS + tid;
    const bool vis = i < n && a.out.offsets[i + 1] != a.out.offsets[i];  // culled: the reference leaves its colour stale
    const uint32_t vmask = __ballot_sync(0xffffffffu, vis);
    if (vmask == 0) return;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)__popc(vmask) * (uint32_t)(PLANES * 16)) : "memory");
    float4* row = rows + tid * COLOR_ROW_F4;
    if (vis)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(row)),
                     "l"(a.scene.sh + (size_t)i * SH_PLANES), "r"((uint32_t)(PLANES * 16)), "r"(bar)
                     : "memory");
    float3 dir = make_float3(0.f, 0.f, 0.f);
    if (vis) {
        const float4 po = __ldg(a.scene.posop + i);
        const float cam_pos[3] = { a.cam->cam_pos[0], a.cam->cam_pos[1], a.cam->cam_pos[2] };
        dir = sh_direction(po.x, po.y, po.z, cam_pos);
    }
    {
        uint32_t ready = 0;
        while (!ready)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ready) : "r"(bar) : "memory");
    }
    if (!vis) return;
    // the arithmetic of the blend's on-demand colour (sh_color), the row coming from shared memory
    const ShBasis s = sh_basis(dir.x, dir.y, dir.z, DEG);
    float acc[3] = { 0.0f, 0.0f, 0.0f };
#pragma unroll
    for (int p = 0; p < PLANES; p += 2) {
        const float4 v = row[p];
        const float4 w = p + 1 < PLANES ? row[p + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
        sh_accumulate_planes(acc, s, p, sh_coefs(DEG), v, w);
    }
    const float3 col = sh_finish(acc);
    a.out.color[i] = make_float4(col.x, col.y, col.z, 0.0f);
}

cudaError_t launch_color(const PreprocessLaunch& a, cudaStream_t s) {
    if (a.scene.n == 0) return cudaSuccess;
    const uint32_t grid = (a.scene.n + COLOR_THREADS - 1) / COLOR_THREADS;
    switch (a.sh_degree) {
        case 0: color_kernel<0><<<grid, COLOR_THREADS, 0, s>>>(a); break;
        case 1: color_kernel<1><<<grid, COLOR_THREADS, 0, s>>>(a); break;
        case 2: color_kernel<2><<<grid, COLOR_THREADS, 0, s>>>(a); break;
        default: color_kernel<3><<<grid, COLOR_THREADS, 0, s>>>(a); break;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// introspection: internal records -> reference Splat layout (splat.slang:33-39)
// ---------------------------------------------------------------------------------------------------

__global__ void export_splats_kernel(SplatArrays a, uint32_t n, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t off = a.offsets[i], cnt = a.offsets[i + 1] - off;
    uint32_t* o = out + (size_t)i * 12;
    if (cnt == 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = 0;
        o[3] = off;
        return;
    }
    const SplatGeo r = a.geo[i];
    const float4 c = a.color[i];
    const float2 zr = a.depth_radius[i];
    o[0] = __float_as_uint(c.x); o[1] = __float_as_uint(c.y); o[2] = __float_as_uint(c.z); o[3] = off;
    o[4] = __float_as_uint(r.px); o[5] = __float_as_uint(r.py); o[6] = __float_as_uint(zr.x); o[7] = __float_as_uint(zr.y);
    o[8] = __float_as_uint(r.conic_a); o[9] = __float_as_uint(r.conic_b); o[10] = __float_as_uint(r.conic_c);
    o[11] = __float_as_uint(r.opacity);
}

// The reference's unsorted key/value buffers (keygen.slang:47-53): Gaussian i writes its tiles row-major at offsets[i].
__global__ void export_unsorted_kernel(SplatArrays a, uint32_t n, uint32_t gx, uint64_t* keys, uint32_t* vals, uint32_t capacity) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t off = a.offsets[i];
    if (a.offsets[i + 1] == off) return;
    const uint2 rc = a.rect[i];
    const uint32_t depth = __float_as_uint(a.depth_radius[i].x);
    const uint32_t x0 = rc.x & 0xffffu, y0 = rc.x >> 16, w = rc.y & 0xffffu, h = rc.y >> 16;
    for (uint32_t y = y0; y < y0 + h; ++y)
        for (uint32_t x = x0; x < x0 + w; ++x, ++off)
            if (off < capacity) {
                keys[off] = ((uint64_t)(y * gx + x) << 32) | depth;
                vals[off] = i;
            }
}

cudaError_t launch_export_unsorted(const SplatArrays& a, uint32_t n, uint32_t width, uint64_t* keys, uint32_t* vals, uint32_t capacity,
                                   cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    export_unsorted_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, n, (width + TILE_PX - 1) / TILE_PX, keys, vals, capacity);
    return cudaGetLastError();
}

cudaError_t launch_export_splats(const SplatArrays& a, uint32_t n, void* out48, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    export_splats_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, n, reinterpret_cast<uint32_t*>(out48));
    return cudaGetLastError();
}

}  // namespace tpdcu

/* tpd_oracle.c — CPU restatement of torpedo's Gaussian rasterizer shaders. TEST INFRASTRUCTURE ONLY.
 * See tpd_oracle.h for scope, the "parity unpinned" note and the canonical fp32 evaluation rules.
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * All paths below are relative to /root/reference/torpedo/volumetric/assets/gaussian/ unless absolute. */
#include "tpd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ */
/* helpers                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* float -> int32, truncating, saturating, NaN -> 0 (NVIDIA F2I.TRUNC; SPIR-V leaves out-of-range undefined) */
static inline int32_t f2i_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

/* row-major 4x4 product, each element a left-to-right dot product (HLSL mul(A,B)) */
static void mat4_mul(const float* a, const float* b, float* r) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[i * 4 + j] = ((a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j]) + a[i * 4 + 2] * b[2 * 4 + j]) +
                           a[i * 4 + 3] * b[3 * 4 + j];
}

/* (M · [p,1])[row], left to right; the w term is m[row][3] * 1.0f == m[row][3] exactly */
static inline float mat4_row_point(const float* m, int row, const float* p) {
    return ((m[row * 4 + 0] * p[0] + m[row * 4 + 1] * p[1]) + m[row * 4 + 2] * p[2]) + m[row * 4 + 3];
}

static const float IDENTITY[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };

/* splat/common.slang:83-85 getComputeGrid */
static inline void compute_grid(uint32_t w, uint32_t h, uint32_t* gx, uint32_t* gy) {
    *gx = (w + TPDO_BLOCK_X - 1) / TPDO_BLOCK_X;
    *gy = (h + TPDO_BLOCK_Y - 1) / TPDO_BLOCK_Y;
}

/* splat/volume.slang:12-17 getBoundingRect: all-float, left to right, C truncation, clamp to [0,grid] */
static inline void bounding_rect(float px, float py, float radius, uint32_t gx, uint32_t gy, uint32_t* x0,
                                 uint32_t* y0, uint32_t* x1, uint32_t* y1) {
    *x0 = (uint32_t)imin((int32_t)gx, imax(0, f2i_sat((px - radius) / 16.0f)));
    *y0 = (uint32_t)imin((int32_t)gy, imax(0, f2i_sat((py - radius) / 16.0f)));
    *x1 = (uint32_t)imin((int32_t)gx, imax(0, f2i_sat((((px + radius) + 16.0f) - 1.0f) / 16.0f)));
    *y1 = (uint32_t)imin((int32_t)gy, imax(0, f2i_sat((((py + radius) + 16.0f) - 1.0f) / 16.0f)));
}

/* ------------------------------------------------------------------------------------------------ */
/* pass count: GaussianEngine.h:234-245, GaussianEngine.cpp:351-357                                  */
/* ------------------------------------------------------------------------------------------------ */

uint32_t tpdo_higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step;
        else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

uint32_t tpdo_radix_pass_count(uint32_t width, uint32_t height) {
    uint32_t gx, gy;
    compute_grid(width, height, &gx, &gy);
    const uint32_t bits = tpdo_higher_msb(gx * gy) + 32;
    return (bits + 1) / 2;
}

/* ------------------------------------------------------------------------------------------------ */
/* spherical harmonics: splat/common.slang:3-80                                                     */
/* ------------------------------------------------------------------------------------------------ */

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f };
static const float SH_C3[7] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f };

/* splat/common.slang:25-31 getFeature: channel c of basis idx (1..15) at sh[3 + c*15 + (idx-1)] */
static inline float feat(const float* sh, uint32_t idx, uint32_t c) { return sh[(idx - 1) + c * 15 + 3]; }

/* splat/common.slang:35-80 evaluateSphericalHarmonics, including the band-3 term 12 written as
 * "SH_C3[3]*z*(2zz-3xx-3yy) + getFeature(sh,12)" (:69 — a '+', not a '*'; replicated on purpose). */
static void eval_sh(const float* sh, const float dir[3], uint32_t degree, float out[3]) {
    const float x = dir[0], y = dir[1], z = dir[2];
    for (uint32_t c = 0; c < 3; ++c) {
        float result = SH_C0 * sh[c];
        if (degree > 0) {
            result = ((result - SH_C1 * y * feat(sh, 1, c)) + SH_C1 * z * feat(sh, 2, c)) - SH_C1 * x * feat(sh, 3, c);
            if (degree > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, zx = z * x;
                result = ((((result + SH_C2[0] * xy * feat(sh, 4, c)) + SH_C2[1] * yz * feat(sh, 5, c)) +
                           SH_C2[2] * (2.0f * zz - xx - yy) * feat(sh, 6, c)) +
                          SH_C2[3] * zx * feat(sh, 7, c)) +
                         SH_C2[4] * (xx - yy) * feat(sh, 8, c);
                if (degree > 2) {
                    result = (((((((result + SH_C3[0] * y * (3.0f * xx - yy) * feat(sh, 9, c)) +
                                   SH_C3[1] * xy * z * feat(sh, 10, c)) +
                                  SH_C3[2] * y * (4.0f * zz - xx - yy) * feat(sh, 11, c)) +
                                 SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy)) + /* scalar broadcast */
                                feat(sh, 12, c)) +                                        /* unscaled feature */
                               SH_C3[4] * x * (4.0f * zz - xx - yy) * feat(sh, 13, c)) +
                              SH_C3[5] * z * (xx - yy) * feat(sh, 14, c)) +
                             SH_C3[6] * x * (xx - 3.0f * yy) * feat(sh, 15, c);
                }
            }
        }
        result += 0.5f;
        out[c] = fmaxf(result, 0.0f);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* project: project.slang:33-90                                                                     */
/* ------------------------------------------------------------------------------------------------ */

static void project_one(const float* g, const float* VM, const float* PM, const float* V, const float cam_pos[3],
                        const float focal[2], uint32_t width, uint32_t height, uint32_t gx, uint32_t gy,
                        uint32_t sh_degree, uint32_t* s) {
    s[7] = f2u(0.0f); /* project.slang:34 reset radius */
    s[3] = 0;         /* project.slang:35 reset tile count */

    const float* mean = g; /* project.slang:38 */

    /* splat/common.slang:98-119 passFrustumClipping */
    float view_pos[3];
    for (int i = 0; i < 3; ++i) view_pos[i] = mat4_row_point(VM, i, mean); /* :102 mul(mul(V,M), float4(p,1)).xyz */
    if (view_pos[2] <= 0.0f) return;                                         /* :103 */
    float clip[4];
    for (int i = 0; i < 4; ++i) clip[i] = mat4_row_point(PM, i, mean); /* :106 */
    if (clip[0] < -1.3f * clip[3] || clip[0] > 1.3f * clip[3]) return; /* :109 */
    if (clip[1] < -1.3f * clip[3] || clip[1] > 1.3f * clip[3]) return; /* :110 */
    if (clip[2] < 0.0f || clip[2] > clip[3]) return;                   /* :114 */
    const float inv_w = 1.0f / clip[3];                                /* :116 */
    const float proj_x = clip[0] * inv_w, proj_y = clip[1] * inv_w;    /* :117 */

    /* splat/volume.slang:20-41 computeCovariance */
    const float qx = g[4], qy = g[5], qz = g[6], qw = g[7];
    const float sx = g[8] * g[11], sy = g[9] * g[11], sz = g[10] * g[11];
    const float R[9] = {
        1.0f - 2.0f * (qy * qy + qz * qz), 2.0f * (qx * qy - qw * qz),        2.0f * (qx * qz + qw * qy),
        2.0f * (qx * qy + qw * qz),        1.0f - 2.0f * (qx * qx + qz * qz), 2.0f * (qy * qz - qw * qx),
        2.0f * (qx * qz - qw * qy),        2.0f * (qy * qz + qw * qx),        1.0f - 2.0f * (qx * qx + qy * qy),
    };
    float sg[9]; /* sigma = mul(R, S), S diagonal: structural zeros skipped */
    for (int i = 0; i < 3; ++i) {
        sg[i * 3 + 0] = R[i * 3 + 0] * sx;
        sg[i * 3 + 1] = R[i * 3 + 1] * sy;
        sg[i * 3 + 2] = R[i * 3 + 2] * sz;
    }
    float cov3[9]; /* mul(sigma, transpose(sigma)) */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            cov3[i * 3 + j] = (sg[i * 3 + 0] * sg[j * 3 + 0] + sg[i * 3 + 1] * sg[j * 3 + 1]) + sg[i * 3 + 2] * sg[j * 3 + 2];

    /* splat/volume.slang:44-63 projectCovariance / computeJacobian */
    const float fx = focal[0] / view_pos[2], fy = focal[1] / view_pos[2];
    const float tx = view_pos[0] / view_pos[2], ty = view_pos[1] / view_pos[2];
    const float j02 = -fx * tx, j12 = -fy * ty;
    float T[6]; /* rows 0,1 of mul(J, W), W = (float3x3)V; J's zeros skipped; row 2 of J is zero */
    for (int j = 0; j < 3; ++j) {
        T[0 * 3 + j] = fx * V[0 * 4 + j] + j02 * V[2 * 4 + j];
        T[1 * 3 + j] = fy * V[1 * 4 + j] + j12 * V[2 * 4 + j];
    }
    float M[6]; /* columns 0,1 of mul(cov3D, transpose(T)): M[i][j] = cov3 row i · T row j */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j)
            M[i * 2 + j] = (cov3[i * 3 + 0] * T[j * 3 + 0] + cov3[i * 3 + 1] * T[j * 3 + 1]) + cov3[i * 3 + 2] * T[j * 3 + 2];
    const float c00 = (T[0] * M[0 * 2 + 0] + T[1] * M[1 * 2 + 0]) + T[2] * M[2 * 2 + 0];
    const float c10 = (T[3] * M[0 * 2 + 0] + T[4] * M[1 * 2 + 0]) + T[5] * M[2 * 2 + 0];
    const float c11 = (T[3] * M[0 * 2 + 1] + T[4] * M[1 * 2 + 1]) + T[5] * M[2 * 2 + 1];

    /* project.slang:60-72 */
    const float cx = c00 + 0.3f, cy = c10, cz = c11 + 0.3f;
    const float det = cx * cz - cy * cy;
    if (det == 0.0f) return;
    const float det_inv = 1.0f / det;
    const float conic_a = cz * det_inv, conic_b = -cy * det_inv, conic_c = cx * det_inv;
    const float mid = 0.5f * (cx + cz);
    const float lambda_1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda_2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float radius = ceilf(3.0f * sqrtf(fmaxf(lambda_1, lambda_2)));
    if (!(fabsf(radius) <= 3.402823466e+38f)) return; /* canonical: non-finite radius is culled */

    /* splat/volume.slang:3-8 ndc2pix; project.slang:75-79 */
    const float px = ((proj_x + 1.0f) * (float)width - 1.0f) * 0.5f;
    const float py = ((proj_y + 1.0f) * (float)height - 1.0f) * 0.5f;
    uint32_t x0, y0, x1, y1;
    bounding_rect(px, py, radius, gx, gy, &x0, &y0, &x1, &y1);
    const uint32_t touched = (x1 - x0) * (y1 - y0);
    if (touched == 0) return;

    /* project.slang:82-83 colour; direction uses the untransformed mean */
    float dir[3] = { mean[0] - cam_pos[0], mean[1] - cam_pos[1], mean[2] - cam_pos[2] };
    const float len = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
    dir[0] /= len; dir[1] /= len; dir[2] /= len;
    float color[3];
    eval_sh(g + 12, dir, sh_degree, color);

    /* project.slang:86-89 */
    s[0] = f2u(color[0]); s[1] = f2u(color[1]); s[2] = f2u(color[2]);
    s[4] = f2u(px); s[5] = f2u(py); s[6] = f2u(view_pos[2]); s[7] = f2u(radius);
    s[8] = f2u(conic_a); s[9] = f2u(conic_b); s[10] = f2u(conic_c); s[11] = f2u(g[3]);
    s[3] = touched;
}

void tpdo_project(const float* gaussians, uint32_t n, const uint32_t* entity_idx, const float* models,
                  uint32_t entity_count, const float* camera34, uint32_t width, uint32_t height,
                  uint32_t sh_degree, uint32_t* splats) {
    const float* V = camera34;
    const float* P = camera34 + 16;
    if (models == NULL) { models = IDENTITY; entity_count = 1; entity_idx = NULL; }

    /* per-entity mul(V, M) and mul(P, M) (splat/common.slang:102,106) */
    float* VM = (float*)malloc(sizeof(float) * 16 * entity_count);
    float* PM = (float*)malloc(sizeof(float) * 16 * entity_count);
    for (uint32_t e = 0; e < entity_count; ++e) {
        mat4_mul(V, models + 16 * e, VM + 16 * e);
        mat4_mul(P, models + 16 * e, PM + 16 * e);
    }

    /* splat/common.slang:88-92 getCameraWorldPosition: -mul(transpose(float3x3(V)), V[:,3]) */
    float cam_pos[3];
    for (int i = 0; i < 3; ++i) cam_pos[i] = -((V[0 * 4 + i] * V[0 * 4 + 3] + V[1 * 4 + i] * V[1 * 4 + 3]) + V[2 * 4 + i] * V[2 * 4 + 3]);

    /* project.slang:55 focal = 0.5 * imageSize * camera.focalNDC */
    const float focal[2] = { 0.5f * (float)width * camera34[32], 0.5f * (float)height * camera34[33] };
    uint32_t gx, gy;
    compute_grid(width, height, &gx, &gy);
    if (sh_degree > 3) sh_degree = 3; /* GaussianEngine.cpp:366-370 */

#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const uint32_t e = entity_idx ? entity_idx[i] : 0;
        project_one(gaussians + (size_t)i * TPDO_GAUSSIAN_FLOATS, VM + 16 * e, PM + 16 * e, V, cam_pos, focal, width,
                    height, gx, gy, sh_degree, splats + (size_t)i * TPDO_SPLAT_FLOATS);
    }
    free(VM);
    free(PM);
}

/* ------------------------------------------------------------------------------------------------ */
/* prefix: prefix.slang:38-159 (semantics)                                                          */
/* ------------------------------------------------------------------------------------------------ */

uint32_t tpdo_prefix(uint32_t* splats, uint32_t n) {
    uint32_t running = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t t = splats[(size_t)i * TPDO_SPLAT_FLOATS + 3];
        splats[(size_t)i * TPDO_SPLAT_FLOATS + 3] = running;
        running += t;
    }
    return running;
}

/* ------------------------------------------------------------------------------------------------ */
/* keygen: keygen.slang:21-53                                                                       */
/* ------------------------------------------------------------------------------------------------ */

void tpdo_keygen(const uint32_t* splats, uint32_t n, uint32_t width, uint32_t height, uint64_t* keys, uint32_t* vals) {
    uint32_t gx, gy;
    compute_grid(width, height, &gx, &gy);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const uint32_t* s = splats + (size_t)i * TPDO_SPLAT_FLOATS;
        const float radius = u2f(s[7]);
        if (radius <= 0.0f) continue; /* :29-30 */
        uint32_t x0, y0, x1, y1;
        bounding_rect(u2f(s[4]), u2f(s[5]), radius, gx, gy, &x0, &y0, &x1, &y1);
        uint32_t offset = s[3]; /* :45 exclusive prefix sits at the Gaussian's own index */
        for (uint32_t y = y0; y < y1; ++y)
            for (uint32_t x = x0; x < x1; ++x) {
                keys[offset] = ((uint64_t)(y * gx + x) << 32) | s[6]; /* :49 tile | bits(viewZ) */
                vals[offset] = (uint32_t)i;                            /* :50 */
                ++offset;
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* sort: radix-*.slang + GaussianEngine.cpp:822-841 (semantics: stable, bits [0, 2*pass_count))      */
/* ------------------------------------------------------------------------------------------------ */

void tpdo_sort(uint64_t* keys, uint32_t* vals, uint32_t p, uint64_t* tmp_keys, uint32_t* tmp_vals, uint32_t pass_count) {
    const uint32_t bits = 2 * pass_count > 64 ? 64 : 2 * pass_count;
    if (p == 0 || bits == 0) return;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if ((uint32_t)nt > p) nt = 1;
    uint32_t* hist = (uint32_t*)malloc(sizeof(uint32_t) * 256 * (size_t)nt);
    uint64_t* src_k = keys; uint32_t* src_v = vals;
    uint64_t* dst_k = tmp_keys; uint32_t* dst_v = tmp_vals;
    for (uint32_t shift = 0; shift < bits; shift += 8) {
        const uint32_t width = bits - shift < 8 ? bits - shift : 8;
        const uint32_t mask = (1u << width) - 1u;
#pragma omp parallel num_threads(nt)
        {
#ifdef _OPENMP
            const int t = omp_get_thread_num();
#else
            const int t = 0;
#endif
            const uint64_t lo = (uint64_t)p * t / nt, hi = (uint64_t)p * (t + 1) / nt;
            uint32_t* h = hist + 256 * (size_t)t;
            memset(h, 0, sizeof(uint32_t) * 256);
            for (uint64_t i = lo; i < hi; ++i) h[(src_k[i] >> shift) & mask]++;
#pragma omp barrier
#pragma omp single
            {
                uint32_t running = 0; /* digit-major, thread-minor: keeps chunk order => stable */
                for (uint32_t d = 0; d < 256; ++d)
                    for (int tt = 0; tt < nt; ++tt) {
                        const uint32_t c = hist[256 * (size_t)tt + d];
                        hist[256 * (size_t)tt + d] = running;
                        running += c;
                    }
            }
            for (uint64_t i = lo; i < hi; ++i) {
                const uint32_t pos = h[(src_k[i] >> shift) & mask]++;
                dst_k[pos] = src_k[i];
                dst_v[pos] = src_v[i];
            }
        }
        uint64_t* tk = src_k; src_k = dst_k; dst_k = tk;
        uint32_t* tv = src_v; src_v = dst_v; dst_v = tv;
    }
    if (src_k != keys) {
        memcpy(keys, src_k, sizeof(uint64_t) * p);
        memcpy(vals, src_v, sizeof(uint32_t) * p);
    }
    free(hist);
}

/* ------------------------------------------------------------------------------------------------ */
/* range: range.slang:16-34, zero fill GaussianEngine.cpp:844                                       */
/* ------------------------------------------------------------------------------------------------ */

void tpdo_range(const uint64_t* keys, uint32_t p, uint32_t* ranges, uint32_t tile_count) {
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)tile_count);
    for (uint32_t idx = 0; idx < p; ++idx) {
        const uint32_t curr = (uint32_t)(keys[idx] >> 32);
        if (idx == 0) {
            ranges[2 * curr + 0] = 0;
        } else {
            const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
            if (curr != prev) {
                ranges[2 * prev + 1] = idx;
                ranges[2 * curr + 0] = idx;
            }
        }
        if (idx == p - 1) ranges[2 * curr + 1] = idx + 1;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* blend: blend.slang:22-104                                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* R8G8B8A8_UNORM image store: clamp to [0,1] (NaN -> 0), scale by 255, round to nearest even */
static inline uint8_t unorm8(float c) {
    if (!(c > 0.0f)) return 0;
    if (c > 1.0f) c = 1.0f;
    return (uint8_t)lrintf(c * 255.0f);
}

void tpdo_blend(const uint32_t* splats, const uint32_t* vals, const uint32_t* ranges, uint32_t width, uint32_t height,
                uint8_t* rgba8, float* rgbf, uint32_t* evals) {
    uint32_t gx, gy;
    compute_grid(width, height, &gx, &gy);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t tile = 0; tile < (int64_t)gx * gy; ++tile) {
        const uint32_t tile_x = (uint32_t)(tile % gx), tile_y = (uint32_t)(tile / gx);
        const uint32_t start = ranges[2 * tile + 0], end = ranges[2 * tile + 1]; /* :41 */
        for (uint32_t ly = 0; ly < TPDO_BLOCK_Y; ++ly)
            for (uint32_t lx = 0; lx < TPDO_BLOCK_X; ++lx) {
                const uint32_t pix_x = tile_x * TPDO_BLOCK_X + lx, pix_y = tile_y * TPDO_BLOCK_Y + ly; /* :33 */
                if (!(pix_x < width && pix_y < height)) continue; /* :34,104: computed but never stored */
                float T = 1.0f, color[3] = { 0.0f, 0.0f, 0.0f };
                uint32_t iterated = 0;
                for (uint32_t k = start; k < end; ++k) { /* :77, rounds of 256 flattened; per-pixel `done` */
                    ++iterated;
                    const uint32_t* s = splats + (size_t)vals[k] * TPDO_SPLAT_FLOATS;
                    const float dx = u2f(s[4]) - (float)pix_x, dy = u2f(s[5]) - (float)pix_y; /* :83 */
                    const float ca = u2f(s[8]), cb = u2f(s[9]), cc = u2f(s[10]), opacity = u2f(s[11]);
                    const float power = -0.5f * (ca * dx * dx + cc * dy * dy) - cb * dx * dy; /* :84 */
                    if (power > 0.0f) continue;                                               /* :85 */
                    const float alpha = fminf(0.99f, opacity * expf(power));                  /* :88 */
                    if (alpha < 1.0f / 255.0f) continue;                                      /* :89 */
                    if (T * (1.0f - alpha) < 0.0001f) break; /* :92-95 done; this splat is NOT added */
                    for (int c = 0; c < 3; ++c) color[c] += u2f(s[c]) * alpha * T; /* :98 */
                    T *= (1.0f - alpha);                                           /* :99 */
                }
                const size_t o = (size_t)pix_y * width + pix_x;
                rgba8[4 * o + 0] = unorm8(color[0]);
                rgba8[4 * o + 1] = unorm8(color[1]);
                rgba8[4 * o + 2] = unorm8(color[2]);
                rgba8[4 * o + 3] = 255; /* :104 alpha 1.0 */
                if (rgbf) { rgbf[3 * o + 0] = color[0]; rgbf[3 * o + 1] = color[1]; rgbf[3 * o + 2] = color[2]; }
                if (evals) evals[o] = iterated;
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* whole frame: GaussianEngine.cpp:621-712 rasterFrame (project, prefix | keygen, sort, range, blend) */
/* ------------------------------------------------------------------------------------------------ */

static double now_ms(void) {
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.0;
#endif
}

uint32_t tpdo_frame(const float* gaussians, uint32_t n, const uint32_t* entity_idx, const float* models,
                    uint32_t entity_count, const float* camera34, uint32_t width, uint32_t height,
                    uint32_t sh_degree, uint32_t* splats, uint64_t* keys, uint32_t* vals, uint64_t* tmp_keys,
                    uint32_t* tmp_vals, uint32_t capacity, uint32_t* ranges, uint8_t* rgba8, double stage_ms[7]) {
    uint32_t gx, gy;
    compute_grid(width, height, &gx, &gy);
    double t0 = now_ms(), t = t0, u;
    tpdo_project(gaussians, n, entity_idx, models, entity_count, camera34, width, height, sh_degree, splats);
    u = now_ms(); stage_ms[0] = u - t; t = u;
    const uint32_t p = tpdo_prefix(splats, n);
    u = now_ms(); stage_ms[1] = u - t; t = u;
    if (p > capacity) return UINT32_MAX;
    tpdo_keygen(splats, n, width, height, keys, vals);
    u = now_ms(); stage_ms[2] = u - t; t = u;
    tpdo_sort(keys, vals, p, tmp_keys, tmp_vals, tpdo_radix_pass_count(width, height));
    u = now_ms(); stage_ms[3] = u - t; t = u;
    tpdo_range(keys, p, ranges, gx * gy);
    u = now_ms(); stage_ms[4] = u - t; t = u;
    tpdo_blend(splats, vals, ranges, width, height, rgba8, NULL, NULL);
    u = now_ms(); stage_ms[5] = u - t;
    stage_ms[6] = u - t0;
    return p;
}

/* ------------------------------------------------------------------------------------------------ */
/* PLY field transforms: /root/reference/torpedo/volumetric/src/GaussianGeometry.cpp:110-117         */
/* ------------------------------------------------------------------------------------------------ */

/* math/include/torpedo/math/common.h:31-40, :19-28 and vec4.h:248-265 (compensated dot) */
static inline void comp_mul(float a, float b, float* m, float* e) { *m = a * b; *e = fmaf(a, b, -*m); }
static inline void comp_sum(float a, float b, float* s, float* e) {
    *s = a + b;
    const float z = *s - a;
    *e = a - (*s - z) + (b - z);
}
static float comp_dot4(const float* a, const float* b) {
    float xx, e0, yy, e1, xy, e3, zz, e2, xz, e5, ww, e7, d, e;
    comp_mul(a[0], b[0], &xx, &e0);
    comp_mul(a[1], b[1], &yy, &e1);
    comp_sum(xx, yy, &xy, &e3);
    const float e4 = e0 + (e3 + e1);
    comp_mul(a[2], b[2], &zz, &e2);
    comp_sum(xy, zz, &xz, &e5);
    const float e6 = e4 + (e5 + e2);
    comp_mul(a[3], b[3], &ww, &e7);
    comp_sum(xz, ww, &d, &e);
    const float c = e6 + (e + e7);
    return d + c;
}

void tpdo_from_model_fields(const float* raw59, uint32_t n, float* gaussians) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* r = raw59 + (size_t)i * 59;
        float* g = gaussians + (size_t)i * TPDO_GAUSSIAN_FLOATS;
        g[0] = r[0]; g[1] = r[1]; g[2] = r[2];
        g[3] = 1.f / (1.f + expf(-r[10]));
        float q[4] = { r[4], r[5], r[6], r[3] }; /* (rot_1, rot_2, rot_3, rot_0) */
        const float inv = 1.0f / sqrtf(comp_dot4(q, q)); /* math/vec4.h:226-231: vec / scalar multiplies by the reciprocal */
        for (int k = 0; k < 4; ++k) g[4 + k] = q[k] * inv;
        g[8] = expf(r[7]); g[9] = expf(r[8]); g[10] = expf(r[9]); g[11] = 1.0f;
        memcpy(g + 12, r + 11, sizeof(float) * 48);
    }
}

int tpdo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm under torchrun inherits OMP_NUM_THREADS=1: it asks for every core it may run on instead */
void tpdo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

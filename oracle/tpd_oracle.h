/* tpd_oracle.h — CPU oracle for torpedo's Gaussian-splatting forward rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (torpedo_b200/, include/) never includes, links or dlopens it.
 *
 * PARITY UNPINNED BY THE REFERENCE: ndming/torpedo ships no tests, golden vectors or fixtures for this
 * path, and its Slang shaders cannot be compiled or run in this image (no slangc, no Vulkan ICD). This
 * file is a stage-by-stage restatement of the shaders under
 *   /root/reference/torpedo/volumetric/assets/gaussian/
 * with every function citing the file:line it follows. What IS pinned against real reference code:
 * the camera/projection matrices and the 240-byte input record, via oracle/_ref (the reference's own
 * Camera.cpp / PerspectiveCamera.cpp / math headers compiled in place, see oracle/Makefile) and the
 * fixtures generated from it under tests/golden/.
 *
 * Canonical floating-point evaluation (the reference's SPIR-V leaves it driver-defined):
 * fp32 everywhere, the operation order as written in the .slang source, row·column dot products
 * accumulated left to right, no FMA contraction (built with -ffp-contract=off), structural zeros of
 * sparse matrices skipped, float→int conversion truncating and saturating (NaN → 0) as NVIDIA's
 * F2I does, fmaxf/fminf NaN semantics. A Gaussian whose radius is not finite is culled.
 */
#ifndef TPD_ORACLE_H
#define TPD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPDO_GAUSSIAN_FLOATS 60 /* 240-byte record: splat.slang:24-31 */
#define TPDO_SPLAT_FLOATS 12    /* 48-byte record:  splat.slang:33-39 */
#define TPDO_CAMERA_FLOATS 34   /* 136-byte UBO:    splat.slang:18-22 */
#define TPDO_BLOCK_X 16
#define TPDO_BLOCK_Y 16

/* Reference Splat layout (splat.slang:33-39), viewed as 12 x 32-bit words:
 *  [0..2] color rgb  [3] tiles (uint32)  [4] px [5] py [6] viewZ [7] radius  [8..10] conic A,B,C  [11] opacity */

/* GaussianEngine.h:234-245 (getHigherMSB) + GaussianEngine.cpp:351-357 (updateRadixPassCount) */
uint32_t tpdo_higher_msb(uint32_t n);
uint32_t tpdo_radix_pass_count(uint32_t width, uint32_t height);

/* project.slang:33-90 with splat/common.slang + splat/volume.slang.
 * gaussians: n x 60 floats. entity_idx: n transform indices or NULL (all 0). models: entity_count x 16
 * floats row-major, or NULL (identity). splats: n x 12 words, read-modify-write exactly like the shader:
 * radius and tiles are reset for every Gaussian, the other fields are written only for visible ones
 * (culled Gaussians keep whatever the buffer held before). */
void tpdo_project(const float* gaussians, uint32_t n, const uint32_t* entity_idx, const float* models,
                  uint32_t entity_count, const float* camera34, uint32_t width, uint32_t height,
                  uint32_t sh_degree, uint32_t* splats);

/* prefix.slang:38-159 — semantics only: in-place EXCLUSIVE scan of splats[i].tiles; returns the total
 * (what the shader writes to tilesRendered[0], prefix.slang:132). */
uint32_t tpdo_prefix(uint32_t* splats, uint32_t n);

/* keygen.slang:21-53 */
void tpdo_keygen(const uint32_t* splats, uint32_t n, uint32_t width, uint32_t height, uint64_t* keys,
                 uint32_t* vals);

/* radix-{shuffle,prefixA,prefixB,mapping}.slang + GaussianEngine.cpp:822-841 — semantics only:
 * a STABLE sort of (key,val) pairs on key bits [0, 2*pass_count). tmp_* are scratch of p elements. */
void tpdo_sort(uint64_t* keys, uint32_t* vals, uint32_t p, uint64_t* tmp_keys, uint32_t* tmp_vals,
               uint32_t pass_count);

/* range.slang:16-34 after the zero fill of GaussianEngine.cpp:844. ranges: tiles x 2 uint32 (start,end). */
void tpdo_range(const uint64_t* keys, uint32_t p, uint32_t* ranges, uint32_t tile_count);

/* blend.slang:22-104 + the R8G8B8A8_UNORM store (GaussianEngine.cpp:316-319).
 * rgba8: width*height*4 bytes. rgbf (optional): width*height*3 floats, the colour before the UNORM
 * conversion. evals (optional): width*height uint32, number of splats each pixel iterated over before
 * it was done (the blend stage's algorithmic work). */
void tpdo_blend(const uint32_t* splats, const uint32_t* vals, const uint32_t* ranges, uint32_t width,
                uint32_t height, uint8_t* rgba8, float* rgbf, uint32_t* evals);

/* Whole frame, stage by stage, with per-stage wall times (ms): project, prefix, keygen, sort, range,
 * blend, total. Buffers as above; keys/vals/tmp must hold `capacity` pairs. Returns P, or
 * UINT32_MAX if P > capacity (nothing after prefix is executed then). */
uint32_t tpdo_frame(const float* gaussians, uint32_t n, const uint32_t* entity_idx, const float* models,
                    uint32_t entity_count, const float* camera34, uint32_t width, uint32_t height,
                    uint32_t sh_degree, uint32_t* splats, uint64_t* keys, uint32_t* vals,
                    uint64_t* tmp_keys, uint32_t* tmp_vals, uint32_t capacity, uint32_t* ranges,
                    uint8_t* rgba8, double stage_ms[7]);

/* 3DGS-PLY → GaussianPoint field transforms (GaussianGeometry.cpp:110-117): opacity=sigmoid(raw),
 * quaternion=(rot_1,rot_2,rot_3,rot_0) normalised (compensated dot, math/vec4.h:248-265),
 * scale=exp(raw) with modifier 1, SH copied. raw: n x (3 pos,4 rot,3 scale,1 opacity,48 sh)=59 floats. */
void tpdo_from_model_fields(const float* raw59, uint32_t n, float* gaussians);

int tpdo_num_threads(void);
/* OpenMP threads of the following calls (bench.py's reference arm: all cores, whatever OMP_NUM_THREADS torchrun exported) */
void tpdo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

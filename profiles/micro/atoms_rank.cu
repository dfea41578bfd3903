// Micro-benchmark (B200, sm_100a): what does a stable in-warp ranking cost per 32-key row?
//   add1   r = atom.shared.add(&cnt[d], 1)   (SASS ATOMS.ADD with a return value) — and: is the returned value ordered by lane?
//   popc   atomicAdd(&cnt[d], 1) with the result unused (SASS ATOMS.POPC.INC): counting only
//   or     atom.shared.or(&mask[d], 1 << lane) + ld.shared           (defined semantics)
//   match  match.any.sync.b32
//   ballot eight vote.ballot + bit logic (what onesweep_kernel does today)
// Digit patterns: 0 random 8-bit, 1 stride 4 (the tile sort's pass 0: two constant low bits), 2 all equal, 3 all distinct.
// Prints cycles per row per SM (all warps of the SM together) and the number of rows whose POPC.INC ranks were NOT in lane order.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 32, ITERS = 64;

__device__ __forceinline__ uint32_t lt_mask() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ __forceinline__ uint32_t digit_of(uint32_t pattern, uint32_t seed, uint32_t lane) {
    const uint32_t h = hash32(seed * 32u + lane);
    switch (pattern) {
        case 0: return h & 255u;
        case 1: return ((h & 63u) << 2) | (seed & 3u);                 // two constant low bits
        case 2: return seed & 255u;                                    // one bin
        case 3: return (lane * 8u + (seed & 7u)) & 255u;               // all distinct
        case 4: return ((seed + lane) & 63u) << 2;                     // consecutive tiles (emission order)
        default: return (h & 127u);                                    // 7-bit
    }
}

template <int MODE>
__global__ void __launch_bounds__(1024) rank_kernel(uint32_t pattern, uint32_t* out, unsigned long long* cycles, uint32_t* violations) {
    extern __shared__ uint32_t sm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = blockDim.x >> 5;
    uint32_t* cnt = sm + warp * 512u;        // [256] counters
    uint32_t* msk = cnt + 256u;              // [256] match masks
    for (uint32_t i = lane; i < 512u; i += 32u) cnt[i] = 0;
    __syncthreads();
    uint32_t acc = 0, bad = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll 8
        for (int row = 0; row < ROWS; ++row) {
            const uint32_t seed = (blockIdx.x * 64u + warp) * 4096u + it * ROWS + row;
            const uint32_t d = digit_of(pattern, seed, lane);
            uint32_t r;
            if (MODE == 0) {          // ATOMS.POPC.INC with return
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(cnt + d)) : "memory");
            } else if (MODE == 1) {   // POPC.INC + verification of lane order against match.any
                const uint32_t before = cnt[d];
                __syncwarp();
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(cnt + d)) : "memory");
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                const uint32_t want = before + __popc(peers & lt_mask());
                if (__any_sync(0xffffffffu, r != want)) ++bad;
                __syncwarp();
            } else if (MODE == 2) {   // atomic OR of the lane bit, read back, leader clears + bumps the counter
                atomicOr(msk + d, 1u << lane);
                __syncwarp();
                const uint32_t peers = msk[d];
                const uint32_t base = cnt[d];
                __syncwarp();
                const uint32_t lower = __popc(peers & lt_mask());
                if (lower == 0) { msk[d] = 0; cnt[d] = base + __popc(peers); }
                __syncwarp();
                r = base + lower;
            } else if (MODE == 3) {   // match.any + counter
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                const uint32_t base = cnt[d];
                __syncwarp();
                const uint32_t lower = __popc(peers & lt_mask());
                if (lower == 0) cnt[d] = base + __popc(peers);
                __syncwarp();
                r = base + lower;
            } else if (MODE == 4) {   // eight ballots + counter (today's kernel)
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (uint32_t bit = 0; bit < 8; ++bit) {
                    const bool p = (d >> bit) & 1u;
                    const uint32_t b = __ballot_sync(0xffffffffu, p);
                    peers &= p ? b : ~b;
                }
                const uint32_t base = cnt[d];
                __syncwarp();
                const uint32_t lower = __popc(peers & lt_mask());
                if (lower == 0) cnt[d] = base + __popc(peers);
                __syncwarp();
                r = base + lower;
            } else if (MODE == 5) {   // atomic add of a lane-specific value
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(cnt + d)), "r"(lane + 1u) : "memory");
            } else {                   // MODE 6: count only, result unused (SASS ATOMS.POPC.INC) — today's early counts
                atomicAdd(cnt + d, 1u);
                r = d;
            }
            acc += r;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (lane == 0 && bad) atomicAdd(violations, bad);
    (void)nwarps;
}

template <int MODE>
static void run(const char* name, int threads) {
    uint32_t* out; unsigned long long* cyc; uint32_t* viol;
    const int blocks = 148;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8); cudaMalloc(&viol, 4);
    const size_t smem = (threads / 32) * 512 * 4;
    cudaFuncSetAttribute(rank_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (uint32_t pattern = 0; pattern < 6; ++pattern) {
        cudaMemset(viol, 0, 4);
        rank_kernel<MODE><<<blocks, threads, smem>>>(pattern, out, cyc, viol);
        rank_kernel<MODE><<<blocks, threads, smem>>>(pattern, out, cyc, viol);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        unsigned long long h[148]; uint32_t v;
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&v, viol, 4, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)h[i]; mean /= blocks;
        const double rows_per_sm = (double)(threads / 32) * ROWS * ITERS;
        printf("{\"mode\": \"%s\", \"warps_per_sm\": %d, \"pattern\": %u, \"cycles_per_row_per_sm\": %.2f, \"rows_checked\": %.0f, \"order_violations\": %u}\n",
               name, threads / 32, pattern, mean / rows_per_sm, MODE == 1 ? rows_per_sm * blocks * 2 : 0.0, v);
    }
    cudaFree(out); cudaFree(cyc); cudaFree(viol);
}

int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("atoms_add1_ret", threads);
        run<1>("atoms_add1_ret_checked", threads);
        run<2>("atomic_or", threads);
        run<3>("match_any", threads);
        run<4>("ballot8", threads);
        run<5>("atoms_add_value", threads);
        run<6>("popc_inc_noret", threads);
    }
    return 0;
}

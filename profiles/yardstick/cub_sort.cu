// Yardstick: cub::DeviceRadixSort::SortKeys from CUDA 12.9 on words shaped like a frame's tile-sort
// input — (tile << 2 | depth bits) << 32 | Gaussian index, 15 significant key bits = two 8-bit onesweep passes — and the
// 24-bit depth sort of the visible Gaussians. Prints ms per sort (CUDA events, median of 20), to set beside
// stages_ms.tile_sort / depth_sort of bench.py. Library code: measured, never linked into libtpdcu.
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void fill(uint64_t* k, uint32_t n, uint32_t key_bits, uint64_t seed) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        k[i] = ((z >> 40) & ((1ull << key_bits) - 1)) << 32 | (z & 0x7fffffu);
    }
}

static float time_sort(uint64_t* in, uint64_t* out, uint32_t n, int begin_bit, int end_bit) {
    void* tmp = nullptr;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(tmp, bytes, in, out, n, begin_bit, end_bit);
    cudaMalloc(&tmp, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    std::vector<float> t;
    for (int r = 0; r < 23; ++r) {
        cudaEventRecord(a);
        cub::DeviceRadixSort::SortKeys(tmp, bytes, in, out, n, begin_bit, end_bit);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r >= 3) t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    cudaFree(tmp);
    return t[t.size() / 2];
}

int main(int argc, char** argv) {
    const uint32_t pairs = argc > 1 ? (uint32_t)atoll(argv[1]) : 15856111u;
    const uint32_t visible = argc > 2 ? (uint32_t)atoll(argv[2]) : 4797287u;
    uint64_t *in, *out;
    cudaMalloc(&in, (size_t)pairs * 8);
    cudaMalloc(&out, (size_t)pairs * 8);
    fill<<<1184, 256>>>(in, pairs, 15, 1);
    const float tile_ms = time_sort(in, out, pairs, 32, 47);
    fill<<<1184, 256>>>(in, visible, 24, 2);
    const float depth_ms = time_sort(in, out, visible, 32, 56);
    fill<<<1184, 256>>>(in, pairs, 32, 3);
    const float full_ms = time_sort(in, out, pairs, 0, 64);
    printf("{\"cub_version\": %d, \"pairs\": %u, \"visible\": %u, \"tile_sort_15bit_ms\": %.4f, \"depth_sort_24bit_ms\": %.4f, "
           "\"full_64bit_sort_ms\": %.4f, \"tile_sort_gbs_model\": %.1f}\n",
           CUB_VERSION, pairs, visible, tile_ms, depth_ms, full_ms, (double)pairs * (8 + 2 * 16) / tile_ms / 1e6);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}

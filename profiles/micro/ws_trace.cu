// Per-tile timeline of the warp-specialised onesweep pass (sort.cu, onesweep_ws_kernel) on tile-sort-like input:
// globaltimer stamps of ticket draw, keys in shared memory, aggregate published, look-back start/end, ranking done,
// bases received, write-out done, plus the look-back depth and the number of row re-fetches per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTPDCU_WS_TRACE=4096 -I torpedo_b200/csrc -o profiles/micro/bin/ws_trace profiles/micro/ws_trace.cu
#include "../../torpedo_b200/csrc/sort.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace tpdcu;

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 15856111u;
    const uint32_t tile_bits = 13, grid_x = 120, grid_y = 68;
    std::vector<uint64_t> h(n);
    uint64_t rng = 88172645463325252ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    // emission order: Gaussians in depth order, each a small rectangle of tiles, row-major; centre-heavy like the garden scene
    uint32_t i = 0, g = 0;
    while (i < n) {
        const uint32_t w = 1 + next() % 3, hh = 1 + next() % 3;
        auto centred = [&](uint32_t lim) { const uint64_t a = next() % lim, b = next() % lim, c = next() % lim; return (uint32_t)((a + b + c) / 3); };
        const uint32_t x0 = std::min(centred(grid_x), grid_x - w), y0 = std::min(centred(grid_y), grid_y - hh);
        const uint32_t d2 = (uint32_t)(((uint64_t)i * 4) / n);
        for (uint32_t y = 0; y < hh && i < n; ++y)
            for (uint32_t x = 0; x < w && i < n; ++x) h[i++] = ((uint64_t)((((y0 + y) * grid_x + x0 + x) << 2) | d2) << 32) | g;
        ++g;
    }
    const uint32_t cap = (n + SORT_TILE - 1) / SORT_TILE * SORT_TILE;
    uint64_t* keys[2];
    cudaMalloc(&keys[0], (size_t)cap * 8); cudaMalloc(&keys[1], (size_t)cap * 8);
    cudaMemcpy(keys[0], h.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
    FrameCtl fc{};
    fc.pairs_total = n; fc.visible = g;
    fc.inv_depth_min = ~0x3f000000u; fc.depth_max = 0x3f000000u + (1u << 25) + 5u;   // 26 depth bits: two ride in the pair key
    FrameCtl* d_fc; cudaMalloc(&d_fc, sizeof(FrameCtl) + 4096);
    SortPlan* plan; cudaMalloc(&plan, sizeof(SortPlan)); cudaMemset(plan, 0, sizeof(SortPlan));
    const uint32_t passes = 2, parts = sort_parts(cap, SORT_KIND_TILE);
    uint32_t* lb; cudaMalloc(&lb, (size_t)passes * parts * SORT_BINS * 4);
    init_sort_attributes();
    SortLaunch so{};
    so.keys[0] = keys[0]; so.keys[1] = keys[1]; so.frame = d_fc; so.ctl = &d_fc->tile_sort; so.plan = plan; so.lookback = lb;
    so.kind = SORT_KIND_TILE; so.capacity = cap; so.end_bit = 16; so.tile_bits = tile_bits; so.sm_count = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(keys[0], h.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(d_fc, &fc, sizeof(fc), cudaMemcpyHostToDevice);
        cudaMemset(lb, 0, (size_t)passes * parts * SORT_BINS * 4);
        cudaEventRecord(e0);
        launch_sort(so, 0, 0, nullptr);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    // verify: stable sort by the high 15 bits
    std::vector<uint64_t> out(n);
    SortPlan hp; cudaMemcpy(&hp, plan, sizeof(hp), cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), keys[hp.final_sel], (size_t)n * 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    for (uint32_t k = 1; k < n && ok; ++k) ok = (out[k - 1] >> 32) < (out[k] >> 32) || ((out[k - 1] >> 32) == (out[k] >> 32) && (uint32_t)out[k - 1] <= (uint32_t)out[k]);
    printf("{\"n\": %u, \"sort_ms\": %.4f, \"sorted\": %s, \"passes_run\": %u}\n", n, ms, ok ? "true" : "false", hp.passes_run);
#ifdef TPDCU_WS_TRACE
    static unsigned long long tr[TPDCU_WS_TRACE][12];
    cudaMemcpyFromSymbol(tr, g_ws_trace, sizeof(tr));
    const uint32_t tiles = std::min<uint32_t>((n + SORT_TILE_WORDS - 1) / SORT_TILE_WORDS, TPDCU_WS_TRACE);
    unsigned long long t0 = ~0ull;
    for (uint32_t t = 0; t < tiles; ++t) t0 = std::min(t0, tr[t][0]);
    printf("# last pass; ns since the first ticket: tile group draw keys_in agg lb_start lb_done rank_done bases_in write_done | depth retries\n");
    for (uint32_t t = 0; t < tiles; ++t) {
        printf("%u %llu", t, tr[t][10]);
        for (int k = 0; k < 8; ++k) printf(" %lld", (long long)(tr[t][k] - t0));
        printf(" | %llu %llu\n", tr[t][8], tr[t][9]);
    }
    static unsigned long long lbt[TPDCU_WS_TRACE][16];
    cudaMemcpyFromSymbol(lbt, g_ws_lb, sizeof(lbt));
    printf("# look-back batches of every 97th tile: ns since lb_start: issue/consumed pairs\n");
    for (uint32_t t = 5; t < tiles; t += 97) {
        printf("LB %u:", t);
        for (int k = 0; k < 16 && lbt[t][k] >= tr[t][3]; ++k) printf(" %lld", (long long)(lbt[t][k] - tr[t][3]));
        printf("\n");
    }
#endif
    return 0;
}

// Stand-alone timing harness for the pair kernel (tuning sweeps; not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I cyclistsocialforce_b200/csrc
//        [-DCSF_PAIR_TPT=.. ...] tools/pairbench.cu -o gpurun_out/pb_x
#include "../cyclistsocialforce_b200/csrc/csf_pair.cu"
#include "../cyclistsocialforce_b200/csrc/csf_pair_tiled.cu"
#include <algorithm>
#include <numeric>
#include <vector>
#include <random>
int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 65536;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const double q = 1.0 / (1 << 18);
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0.0, 4.0 * sqrt((double)n)), A(-3.14159, 3.14159);
    std::vector<Xycs<float>> h(n);
    for (auto& e : h) { double a = A(rng); e.xq = (int)(U(rng) / q); e.yq = (int)(U(rng) / q); e.c = cosf(a); e.s = sinf(a); }
    Xycs<float>* d; float* f; void* ws;
    cudaMalloc(&d, n * sizeof(Xycs<float>)); cudaMalloc(&f, n * 8);
    cudaMemcpy(d, h.data(), n * sizeof(Xycs<float>), cudaMemcpyHostToDevice);
    CsfFieldParams fp = {7.0, 0.995, 0.7, 0.5, 5.0, 0.3, 4.9, 2.0943951023931953, q, 0, 0, 0, 0, 0, 0};
    if (getenv("CSF_PB_CUTOFF_LOG2")) fp.cutoff_log2 = atof(getenv("CSF_PB_CUTOFF_LOG2"));   // default 40
    size_t wsb = csf_pair_workspace_bytes(n, n, 4);
    cudaMalloc(&ws, wsb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const bool nodense = argc > 4;   // 4th argument: skip the dense kernel (large n)
    for (int i = 0; i < (nodense ? 0 : 3); ++i) csf_pair_forces_f32(d, n, d, n, &fp, f, 0, ws, wsb, 0);
    cudaDeviceSynchronize();
    float best = 1e30f, tot = 0;
    for (int i = 0; i < (nodense ? 0 : reps); ++i) {
        cudaEventRecord(e0); csf_pair_forces_f32(d, n, d, n, &fp, f, 0, ws, wsb, 0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); tot += ms;
    }
    if (argc > 3) {  // tiled variant
        int64_t* keys; cudaMalloc(&keys, n * 8);
        csf_morton_keys_f32(d, n, 0, 0, (4.0 * sqrt((double)n) / q) / 65535.0, keys, 0);   // key domain = bounding box
        std::vector<int64_t> hk(n), perm(n);
        cudaMemcpy(hk.data(), keys, n * 8, cudaMemcpyDeviceToHost);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return hk[a] < hk[b]; });
        int64_t* dperm; cudaMalloc(&dperm, n * 8); cudaMemcpy(dperm, perm.data(), n * 8, cudaMemcpyHostToDevice);
        void *sorted, *tiles, *ws2; unsigned long long* stats;
        cudaMalloc(&sorted, csf_tiled_padded_sources(n) * 16); cudaMalloc(&tiles, csf_tiled_num_tiles(n) * 16);
        cudaMalloc(&stats, 8); cudaMemset(stats, 0, 8);
        size_t wsb2 = csf_pair_tiled_workspace_bytes(n, n, 4); cudaMalloc(&ws2, wsb2);
        float* f2; cudaMalloc(&f2, n * 8);
        if (argc > 5) {   // 5th argument: divisor -- time one shard (the first n/div targets of the spatial order)
            const int64_t nt = n / atoi(argv[5]);
            for (int i = 0; i < 3; ++i) {
                csf_tile_sources_f32(d, n, dperm, sorted, tiles, 0);
                csf_pair_forces_tiled_f32(sorted, tiles, n, d, dperm, nt, &fp, f2, 0, ws2, wsb2, nullptr, 0);
            }
            cudaDeviceSynchronize();
            float bt = 1e30f;
            for (int i = 0; i < reps; ++i) {
                cudaEventRecord(e0); csf_pair_forces_tiled_f32(sorted, tiles, n, d, dperm, nt, &fp, f2, 0, ws2, wsb2, nullptr, 0);
                cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); bt = fminf(bt, ms);
            }
            printf("SHARD n_src=%lld n_tgt=%lld best %.4f ms (x%d = %.4f ms) err=%s\n", (long long)n, (long long)nt, bt,
                   atoi(argv[5]), bt * atoi(argv[5]), cudaGetErrorString(cudaGetLastError()));
            return 0;
        }
        csf_tile_sources_f32(d, n, dperm, sorted, tiles, 0);
        csf_pair_forces_tiled_f32(sorted, tiles, n, d, dperm, n, &fp, f2, 0, ws2, wsb2, stats, 0);
        cudaDeviceSynchronize();
        unsigned long long ne = 0; cudaMemcpy(&ne, stats, 8, cudaMemcpyDeviceToHost);
        float bt = 1e30f, tt = 0, ttile = 0;
        for (int i = 0; i < reps; ++i) {
            cudaEventRecord(e0); csf_tile_sources_f32(d, n, dperm, sorted, tiles, 0); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms0; cudaEventElapsedTime(&ms0, e0, e1); ttile += ms0;
            cudaEventRecord(e0); csf_pair_forces_tiled_f32(sorted, tiles, n, d, dperm, n, &fp, f2, 0, ws2, wsb2, nullptr, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
            bt = fminf(bt, ms); tt += ms;
        }
        std::vector<float> a(2 * n), b(2 * n);
        cudaMemcpy(a.data(), f, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), f2, n * 8, cudaMemcpyDeviceToHost);
        double md = 0, cs2 = 0; for (int64_t i = 0; i < 2 * n; ++i) { md = fmax(md, fabs((double)a[i] - b[i])); cs2 += fabs(b[i]); }
        printf("TILED n=%lld ctas/sm=%d mean %.3f ms best %.3f ms (tile build %.3f ms)  %.1f Gpair/s dense-equivalent; evaluated fraction %.3f -> %.2f TFLOP/s@76 executed; max|dense-tiled| %.3e checksum %.6e err=%s\n",
               (long long)n, g_tiled_ctas[0], tt / reps, bt, ttile / reps, (double)n * (n - 1) / bt * 1e-6,
               (double)ne / ((double)n * n), (double)ne * 76 / bt * 1e-9, md, cs2, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<float> out(2 * n); cudaMemcpy(out.data(), f, n * 8, cudaMemcpyDeviceToHost);
    double cs = 0; for (float v : out) cs += fabs(v);
    const double pairs = (double)n * (n - 1);
    printf("n=%lld ctas/sm=%d mean %.3f ms best %.3f ms  %.1f Gpair/s (best)  %.2f TFLOP/s@76  checksum %.6e  err=%s\n",
           (long long)n, g_pair_ctas_per_sm[0], tot / reps, best, pairs / best * 1e-6, pairs * 76 / best * 1e-9, cs,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Single-thread latency (cycles) of the pieces of the keyed-draw contract on B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --fmad=false -o latency_probe latency_probe.cu && ./latency_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/seir_draws.h"

__global__ void probe(double *out, long long *cyc, const double *tab, unsigned long long *gmem)
{
    double acc = 0.0;
    uint32_t c = 1;
    long long t0, t1;
    // Philox chain
    t0 = clock64();
    for (int k = 0; k < 64; ++k) { sd_block b = sd_draw(7, c, 3, 5, k); c = b.w[0]; }
    t1 = clock64(); cyc[0] = (t1 - t0) / 64; acc += c;
    double x = 0.3 + 1e-9 * c;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) x = 0.5 + 0.25 * (1.0 + sd_log(x) * 0.1);
    t1 = clock64(); cyc[1] = (t1 - t0) / 64; acc += x;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) x = 0.4 + 0.01 * sd_exp(x);
    t1 = clock64(); cyc[2] = (t1 - t0) / 64; acc += x;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) x = SD_DIV(1.0, 1.5 + x);
    t1 = clock64(); cyc[3] = (t1 - t0) / 64; acc += x;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) x = SD_SQRT(1.5 + x);
    t1 = clock64(); cyc[4] = (t1 - t0) / 64; acc += x;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) x = SD_FMA(x, 0.999, 0.001);
    t1 = clock64(); cyc[5] = (t1 - t0) / 64; acc += x;
    t0 = clock64();
    for (int k = 0; k < 16; ++k) x = 1e-3 * sd_lognormal_days(11, (uint32_t)(x * 1e6) + k, 2, 13, 0, 1.621, 0.418);
    t1 = clock64(); cyc[6] = (t1 - t0) / 16; acc += x;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) x = 0.001 * sd_exp_days(0.1 + 0.5 * (x - floor(x)), 4.6) + 0.3;
    t1 = clock64(); cyc[7] = (t1 - t0) / 64; acc += x;
    // dependent global loads (L2-resident small table / atomics)
    unsigned long long idx = 0;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) idx = __ldcg(gmem + (idx & 1023));
    t1 = clock64(); cyc[8] = (t1 - t0) / 64; acc += idx;
    t0 = clock64();
    for (int k = 0; k < 64; ++k) idx = atomicMax(gmem + 2048 + (idx & 1023), (unsigned long long)k);
    t1 = clock64(); cyc[9] = (t1 - t0) / 64; acc += idx;
    t0 = clock64();
    for (int k = 0; k < 16; ++k) { __threadfence(); gmem[4096 + k] = k; }
    t1 = clock64(); cyc[10] = (t1 - t0) / 16;
    // DRAM-latency chain: 256 MB stride walk
    t0 = clock64();
    for (int k = 0; k < 32; ++k) idx = __ldcg(gmem + 8192 + ((idx + k * 1048583ull * 8) & ((1ull << 25) - 1)));
    t1 = clock64(); cyc[11] = (t1 - t0) / 32; acc += idx;
    out[0] = acc;
}

int main()
{
    double *out; long long *cyc; unsigned long long *g;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&g, (8192 + (1ull << 25)) * 8 + 1024);
    cudaMemset(g, 0, (8192 + (1ull << 25)) * 8);
    probe<<<1, 1>>>(out, cyc, nullptr, g);
    probe<<<1, 1>>>(out, cyc, nullptr, g);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[] = {"philox4x32-10", "sd_log", "sd_exp", "ddiv", "dsqrt", "dfma", "sd_lognormal_days", "sd_exp_days",
                           "ldcg L2 hit", "atomicMax u64 (L2)", "threadfence+store", "ldcg DRAM"};
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int k = 0; k < 12; ++k) printf("%-22s %6lld cycles  (%.3f us at %d MHz)\n", names[k], h[k], h[k] / (clk / 1e3), clk / 1000);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

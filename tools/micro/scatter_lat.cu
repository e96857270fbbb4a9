// Microbenchmark: dependent scattered byte loads from a buffer of a given size, per-SM cycles per load.
// Question: do SMs of one die see much slower cache-missing scattered loads (k_forward_line's access pattern)?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void k_chase(const unsigned char* __restrict__ buf, size_t n, int iters, unsigned long long* cyc, unsigned long long* cnt, unsigned* sink) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long x = (blockIdx.x * 1315423911ull + threadIdx.x * 2654435761ull) | 1ull;
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x = x * 6364136223846793005ull + 1442695040888963407ull + acc;
        const size_t idx = (size_t)((x >> 20) % n);
        acc += buf[idx];                      // next address depends on the loaded value (acc): one load in flight per thread
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { atomicAdd(cyc + smid, (unsigned long long)(t1 - t0)); atomicAdd(cnt + smid, 1ull); }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}
int main(int argc, char** argv) {
    const int iters = 256;
    unsigned long long *cyc, *cnt; unsigned* sink;
    cudaMalloc(&cyc, 256 * 8); cudaMalloc(&cnt, 256 * 8); cudaMalloc(&sink, 4);
    for (size_t mb : {8, 32, 64, 128, 192, 256, 512, 1024}) {
        const size_t n = mb << 20;
        unsigned char* buf; if (cudaMalloc(&buf, n) != cudaSuccess) { printf("alloc fail\n"); return 1; }
        cudaMemset(buf, 1, n);
        for (int warps_per_block : {4}) {
            cudaMemset(cyc, 0, 256 * 8); cudaMemset(cnt, 0, 256 * 8);
            k_chase<<<148 * 12 * 4, 32 * warps_per_block>>>(buf, n, iters, cyc, cnt, sink);
            cudaDeviceSynchronize();
            std::vector<unsigned long long> hc(256), hn(256);
            cudaMemcpy(hc.data(), cyc, 256 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hn.data(), cnt, 256 * 8, cudaMemcpyDeviceToHost);
            std::vector<double> lat; std::vector<unsigned long long> nw;
            for (int s = 0; s < 256; s++) if (hn[s]) { lat.push_back((double)hc[s] / hn[s] / iters); nw.push_back(hn[s]); }
            std::vector<double> sl = lat; std::sort(sl.begin(), sl.end());
            printf("%5zu MB: SMs %zu  cycles/load min %.0f p25 %.0f med %.0f p75 %.0f max %.0f | warps/SM min %llu max %llu\n", mb, sl.size(), sl[0], sl[sl.size() / 4], sl[sl.size() / 2],
                   sl[3 * sl.size() / 4], sl.back(), *std::min_element(nw.begin(), nw.end()), *std::max_element(nw.begin(), nw.end()));
            if (mb == 128) { for (size_t s = 0; s < lat.size(); s++) printf("%.0f%s", lat[s], (s % 16 == 15) ? "\n" : " "); printf("\n"); }
        }
        cudaFree(buf);
    }
    return 0;
}

// latency microbenchmarks (one warp): L2 hit, L1 hit, prefetch.global.L1 effect, cp.async, smem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const uint32_t *chain, int n, uint32_t *out, int mode, int delay) {
    if (threadIdx.x != 0) return;
    uint32_t idx = 0;
    long long t0 = clock64();
    long long acc = 0;
    for (int i = 0; i < n; i++) {
        if (mode == 1) {  // prefetch next-next? no: prefetch target then spin `delay` then load
            asm volatile("prefetch.global.L1 [%0];" ::"l"(chain + idx));
            long long s = clock64();
            while (clock64() - s < delay) {}
        } else if (mode == 2) {  // touch the line by a real load first, then spin, then load again (L1 hit?)
            uint32_t tmp;
            asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(tmp) : "l"(chain + idx + 1));
            acc += tmp;
            long long s = clock64();
            while (clock64() - s < delay) {}
        } else if (mode == 3) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(chain + idx));
            long long s = clock64();
            while (clock64() - s < delay) {}
        }
        long long a = clock64();
        uint32_t nx = chain[idx];
        if (nx == 0xffffffffu) break;
        long long b = clock64();
        acc += 0;
        out[2 + (i & 1023)] = (uint32_t)(b - a);
        idx = nx;
    }
    long long t1 = clock64();
    out[0] = (uint32_t)(t1 - t0);
    out[1] = idx + (uint32_t)acc;
}
int main() {
    const int N = 1 << 19;  // 2 MB of uint32, stride chain with 32-element (128 B) granularity
    uint32_t *h = new uint32_t[N];
    // random permutation cycle over lines
    int lines = N / 32;
    int *perm = new int[lines];
    for (int i = 0; i < lines; i++) perm[i] = i;
    srand(1);
    for (int i = lines - 1; i > 0; i--) { int j = rand() % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (int i = 0; i < N; i++) h[i] = 0;
    for (int i = 0; i < lines; i++) h[perm[i] * 32] = perm[(i + 1) % lines] * 32;
    uint32_t *d, *o;
    cudaMalloc(&d, N * 4); cudaMalloc(&o, 4096 * 4);
    cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice);
    uint32_t ho[1100];
    for (int mode = 0; mode < 4; mode++)
        for (int delay : {0, 600, 1500}) {
            if (mode == 0 && delay) continue;
            k<<<1, 32>>>(d, 2000, o, mode, delay);
            cudaDeviceSynchronize();
            k<<<1, 32>>>(d, 2000, o, mode, delay);
            cudaDeviceSynchronize();
            cudaMemcpy(ho, o, 1100 * 4, cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 1000; i++) s += ho[2 + i];
            printf("mode %d delay %4d: avg load latency %.1f cycles (total %u)\n", mode, delay, s / 1000, ho[0]);
        }
    return 0;
}

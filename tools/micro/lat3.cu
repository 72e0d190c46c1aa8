// does a store to ANOTHER line disturb a line that was pulled into L1?  (one thread, 2 MB pointer chase)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t *chain, uint32_t *other, int n, uint32_t *out, int mode) {
    if (threadIdx.x != 0) return;
    uint32_t idx = 0;
    long long acc = 0;
    for (int i = 0; i < n; i++) {
        uint32_t first = chain[idx + 1];  // pull the line
        acc += first;
        if (mode == 1) other[(i * 64) & 0xffff] = (uint32_t)i;                 // store to a different line (L2 resident)
        if (mode == 2) { acc += other[(i * 64 + 4096) & 0xffff]; other[(i * 64) & 0xffff] = (uint32_t)acc; }  // load + store elsewhere
        if (mode == 3) { other[(i * 64) & 0xffff] = (uint32_t)i; __syncwarp(); }
        long long s = clock64();
        while (clock64() - s < 400) {}
        long long a = clock64();
        uint32_t nx = chain[idx];
        if (nx == 0xffffffffu) break;
        long long b = clock64();
        out[2 + (i & 1023)] = (uint32_t)(b - a);
        idx = nx;
    }
    out[1] = idx + (uint32_t)acc;
}
int main() {
    const int N = 1 << 19;
    uint32_t *h = new uint32_t[N];
    int lines = N / 32;
    int *perm = new int[lines];
    for (int i = 0; i < lines; i++) perm[i] = i;
    srand(1);
    for (int i = lines - 1; i > 0; i--) { int j = rand() % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (int i = 0; i < N; i++) h[i] = 0;
    for (int i = 0; i < lines; i++) h[perm[i] * 32] = perm[(i + 1) % lines] * 32;
    uint32_t *d, *o, *oth;
    cudaMalloc(&d, N * 4); cudaMalloc(&o, 4096 * 4); cudaMalloc(&oth, 65536 * 4);
    cudaMemset(oth, 0, 65536 * 4);
    uint32_t ho[1100];
    for (int mode = 0; mode < 4; mode++) {
        cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(d, oth, 2000, o, mode);
        cudaDeviceSynchronize();
        cudaMemcpy(ho, o, 1100 * 4, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 1000; i++) s += ho[2 + i];
        printf("mode %d (0 nothing, 1 store elsewhere, 2 load+store elsewhere, 3 store+syncwarp): reload latency %.1f\n", mode, s / 1000);
    }
    return 0;
}

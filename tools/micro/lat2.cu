// does a plain load allocate in L1?  does a store to the line keep / update / invalidate it?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t *chain, int n, uint32_t *out, int mode) {
    if (threadIdx.x != 0) return;
    uint32_t idx = 0;
    long long acc = 0;
    for (int i = 0; i < n; i++) {
        volatile uint32_t *vp = chain;
        uint32_t first = chain[idx + 1];            // plain load: first touch of the line (L2)
        acc += first;
        if (mode == 1) chain[idx + 2] = (uint32_t)i;  // store into the same line
        if (mode == 2) { chain[idx + 2] = (uint32_t)i; __threadfence_block(); }
        if (mode == 3) atomicAdd(&chain[idx + 3], 1u);
        long long s = clock64();
        while (clock64() - s < 400) {}
        long long a = clock64();
        uint32_t nx = chain[idx];                    // second touch
        if (nx == 0xffffffffu) break;
        long long b = clock64();
        uint32_t again = chain[idx + 2];
        acc += again;
        long long c2 = clock64();
        out[2 + (i & 1023)] = (uint32_t)(b - a);
        out[1100 + (i & 1023)] = (uint32_t)(c2 - b);
        idx = nx;
        (void)vp;
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
    uint32_t *d, *o;
    cudaMalloc(&d, N * 4); cudaMalloc(&o, 4096 * 4);
    uint32_t ho[2200];
    for (int mode = 0; mode < 4; mode++) {
        cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(d, 2000, o, mode);
        cudaDeviceSynchronize();
        cudaMemcpy(ho, o, 2200 * 4, cudaMemcpyDeviceToHost);
        double s = 0, s2 = 0; for (int i = 0; i < 1000; i++) { s += ho[2 + i]; s2 += ho[1100 + i]; }
        printf("mode %d (0 none,1 store,2 store+fence,3 atomic): 2nd-touch latency %.1f, store-word reload %.1f\n", mode, s / 1000, s2 / 1000);
    }
    return 0;
}

// tools/ubench3.cu -- where does the block scheduler put the first wave of CTAs when two fit per SM?
// Launches `grid` equal CTAs (256 threads, ~17 KB smem, long spin) and prints CTAs per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k(int *smid, long long *t0, long long *t1, long long spin)
{
    extern __shared__ float dyn[];
    unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    long long a; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a));
    const long long c0 = clock64();
    float x = threadIdx.x;
    while (clock64() - c0 < spin) x = x * 1.0001f + 1e-3f;
    dyn[threadIdx.x] = x;
    long long b; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(b));
    if (threadIdx.x == 0) { smid[blockIdx.x] = id; t0[blockIdx.x] = a; t1[blockIdx.x] = b; }
}
int main(int argc, char **argv)
{
    int *smid; long long *t0, *t1;
    cudaMalloc(&smid, 4096 * 4); cudaMalloc(&t0, 4096 * 8); cudaMalloc(&t1, 4096 * 8);
    static int h[4096]; static long long a[4096], b[4096];
    const int grids[] = {144, 148, 152, 192, 296, 300};
    for (int grid : grids) {
        k<<<grid, 256, 17 * 1024>>>(smid, t0, t1, 2000000);
        cudaDeviceSynchronize();
        cudaMemcpy(h, smid, grid * 4, cudaMemcpyDeviceToHost); cudaMemcpy(a, t0, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b, t1, grid * 8, cudaMemcpyDeviceToHost);
        int cnt[256] = {0}; long long mn = a[0], mx = b[0];
        int late = 0;
        for (int i = 0; i < grid; ++i) { cnt[h[i]]++; if (a[i] < mn) mn = a[i]; if (b[i] > mx) mx = b[i]; }
        for (int i = 0; i < grid; ++i) if (a[i] - mn > 500000) late++;    // started > 0.5 ms after the first
        int hist[8] = {0}; int used = 0;
        for (int s = 0; s < 256; ++s) { if (cnt[s]) used++; hist[cnt[s] > 7 ? 7 : cnt[s]]++; }
        printf("grid %4d: SMs used %3d  | SMs with 1 CTA: %3d, 2: %3d, 3: %3d, 4: %3d | CTAs started late: %3d | span %.3f ms\n", grid, used, hist[1], hist[2], hist[3], hist[4], late, (mx - mn) * 1e-6);
        printf("   first 16 bid->smid:");
        for (int i = 0; i < 16; ++i) printf(" %d", h[i]);
        printf("\n");
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// oracle/ref_cuda_kernel_harness.cu -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Times the reference's only CUDA kernel, `nbody` (ver5_all/programming_models/cuda/Compute.cu:31-66),
// ALONE: CUDA events around the launch exactly as the reference issues it (:159-162; grid
// (n+bs-1)/bs, block 1024 as in :137-143), no PCIe copies, no host update.  The kernel is taken
// from the UNMODIFIED reference file by #including it from where it lies under /root/reference
// (oracle/Makefile passes the include path); this file adds only a main().  The end-to-end figure of
// the same backend (its own timer, copies + host update included) comes from ver5_all_cuda/nbody.x.
//
//   ref_cuda_kernel_only <nPart> <launches> [block]   ->  one JSON line on stdout
//
// Inputs are the reference's uniform-cube ICs in spirit (positions U(0,1), mass n*U(0,1)) from a
// fixed LCG: the kernel is branch-free in the data, so its time does not depend on the values.
#include "Compute.cu"   // the reference TU, unmodified (also defines GSimulation::start(), unused here)

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv)
{
    const int n = argc > 1 ? std::atoi(argv[1]) : 131072;
    const int launches = argc > 2 ? std::atoi(argv[2]) : 10;
    const int bs = argc > 3 ? std::atoi(argv[3]) : 1024;
    std::vector<float> h((size_t)n * 4);
    unsigned long long x = 42;
    for (size_t i = 0; i < h.size(); ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        h[i] = (float)((x >> 40) & 0xffffff) / 16777216.0f;
    }
    for (int i = 0; i < n; ++i) h[(size_t)3 * n + i] *= (float)n;
    float *d[7];
    for (int k = 0; k < 7; ++k)
        if (cudaMalloc(&d[k], (size_t)n * sizeof(float)) != cudaSuccess) { std::fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemcpy(d[0], h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d[1], h.data() + n, (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d[2], h.data() + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d[6], h.data() + 3 * (size_t)n, (size_t)n * 4, cudaMemcpyHostToDevice);
    const int grid = (n + bs - 1) / bs;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) nbody<<<grid, bs>>>(d[0], d[1], d[2], d[3], d[4], d[5], d[6], n);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int l = 0; l < launches; ++l) nbody<<<grid, bs>>>(d[0], d[1], d[2], d[3], d[4], d[5], d[6], n);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { std::fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    float ax0 = 0.f;
    cudaMemcpy(&ax0, d[3], 4, cudaMemcpyDeviceToHost);
    const double per = ms / launches;
    std::printf("{\"n\": %d, \"launches\": %d, \"block\": %d, \"ms_per_launch\": %.6f, \"gpairs_per_s\": %.3f, \"acc_x0\": %.9g}\n",
                n, launches, bs, per, (double)n * n / (per * 1e-3) / 1e9, (double)ax0);
    return 0;
}

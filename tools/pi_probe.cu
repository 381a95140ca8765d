// tools/pi_probe.cu -- experiment: pack the f32x2 lanes over two i-bodies ("P-i": j-body operands
// become 32-bit broadcast operands) instead of over two j-bodies ("P-j", the shipped kernel).
// Forces only, N bodies, same register blocking (4 i-bodies per thread, 256 threads, 2 CTAs/SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/pi_probe tools/pi_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../nbody-demo-2023_b200/csrc/nbx_kernels.cuh"

__device__ __forceinline__ float rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int UNROLL, int STAGE>
__global__ void __launch_bounds__(256, 2) pi_kernel(const float4 *__restrict__ body, float4 *__restrict__ acc, int n, int jsplits, float eps2)
{
    constexpr int TJ = 256, R2 = 2;
    __shared__ float4 tile[2][TJ];
    const int tid = threadIdx.x;
    const int itile = blockIdx.x / jsplits, split = blockIdx.x % jsplits;
    const int jb = (int)((long long)n * split / jsplits), je = (int)((long long)n * (split + 1) / jsplits);
    float2 nx[R2], ny[R2], nz[R2], ax[R2], ay[R2], az[R2];
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        const int i0 = 2 * (itile * 256 * R2 + k * 256 + tid);
        const float4 a = body[i0], b = body[i0 + 1];
        nx[k] = make_float2(-a.x, -b.x); ny[k] = make_float2(-a.y, -b.y); nz[k] = make_float2(-a.z, -b.z);
        ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
    }
    const float2 e2 = make_float2(eps2, eps2);
    const int ntiles = (je - jb) / TJ;
    tile[0][tid] = body[jb + tid];
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) tile[(t + 1) & 1][tid] = body[jb + (t + 1) * TJ + tid];
        const float4 *q = tile[t & 1];
#pragma unroll 1
        for (int j = 0; j < TJ; j += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float4 b = q[j + u];
                const float2 xj = make_float2(b.x, b.x), yj = make_float2(b.y, b.y), zj = make_float2(b.z, b.z), mj = make_float2(b.w, b.w);
                if (STAGE) {
                    float2 dx[R2], dy[R2], dz[R2], s[R2];
#pragma unroll
                    for (int k = 0; k < R2; ++k) { dx[k] = __fadd2_rn(nx[k], xj); dy[k] = __fadd2_rn(ny[k], yj); dz[k] = __fadd2_rn(nz[k], zj); }
#pragma unroll
                    for (int k = 0; k < R2; ++k) s[k] = __ffma2_rn(dx[k], dx[k], e2);
#pragma unroll
                    for (int k = 0; k < R2; ++k) s[k] = __ffma2_rn(dy[k], dy[k], s[k]);
#pragma unroll
                    for (int k = 0; k < R2; ++k) s[k] = __ffma2_rn(dz[k], dz[k], s[k]);
#pragma unroll
                    for (int k = 0; k < R2; ++k) s[k] = make_float2(rsq(s[k].x), rsq(s[k].y));
#pragma unroll
                    for (int k = 0; k < R2; ++k) { const float2 i2 = __fmul2_rn(s[k], s[k]); const float2 mi = __fmul2_rn(s[k], mj); s[k] = __fmul2_rn(i2, mi); }
#pragma unroll
                    for (int k = 0; k < R2; ++k) { ax[k] = __ffma2_rn(dx[k], s[k], ax[k]); ay[k] = __ffma2_rn(dy[k], s[k], ay[k]); az[k] = __ffma2_rn(dz[k], s[k], az[k]); }
                } else {
#pragma unroll
                    for (int k = 0; k < R2; ++k) {
                        const float2 dx = __fadd2_rn(nx[k], xj), dy = __fadd2_rn(ny[k], yj), dz = __fadd2_rn(nz[k], zj);
                        float2 r2 = __ffma2_rn(dx, dx, e2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dz, dz, r2);
                        const float2 inv = make_float2(rsq(r2.x), rsq(r2.y));
                        const float2 i2 = __fmul2_rn(inv, inv), mi = __fmul2_rn(inv, mj), s = __fmul2_rn(i2, mi);
                        ax[k] = __ffma2_rn(dx, s, ax[k]); ay[k] = __ffma2_rn(dy, s, ay[k]); az[k] = __ffma2_rn(dz, s, az[k]);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        const int i0 = 2 * (itile * 256 * R2 + k * 256 + tid);
        float4 *dst = acc + (size_t)split * n;
        dst[i0] = make_float4(ax[k].x, ay[k].x, az[k].x, 0.f);
        dst[i0 + 1] = make_float4(ax[k].y, ay[k].y, az[k].y, 0.f);
    }
}

template <int U, int ST>
double time_pi(const float4 *body, float4 *acc, int n, int splits, int reps)
{
    const int grid = n / 1024 * splits;
    pi_kernel<U, ST><<<grid, 256>>>(body, acc, n, splits, 1e-3f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) pi_kernel<U, ST><<<grid, 256>>>(body, acc, n, splits, 1e-3f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 262144, splits = 37, reps = 5;
    std::vector<float4> h(n);
    srand(1);
    for (auto &b : h) b = make_float4(rand() / (float)RAND_MAX, rand() / (float)RAND_MAX, rand() / (float)RAND_MAX, 1e-4f * rand() / (float)RAND_MAX);
    float4 *body, *acc; cudaMalloc(&body, n * sizeof(float4)); cudaMalloc(&acc, (size_t)splits * n * sizeof(float4));
    cudaMemcpy(body, h.data(), n * sizeof(float4), cudaMemcpyHostToDevice);
    // correctness spot check of P-i against a double sum on the host for 8 bodies
    pi_kernel<4, 1><<<n / 1024 * splits, 256>>>(body, acc, n, splits, 1e-3f);
    std::vector<float4> a((size_t)splits * n);
    cudaMemcpy(a.data(), acc, a.size() * sizeof(float4), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < n; i += n / 8) {
        double sx = 0, gx = 0;
        for (int j = 0; j < n; ++j) {
            double dx = (double)h[j].x - h[i].x, dy = (double)h[j].y - h[i].y, dz = (double)h[j].z - h[i].z;
            double r2 = dx * dx + dy * dy + dz * dz + 1e-3; sx += dx * h[j].w / (r2 * sqrt(r2));
        }
        for (int s = 0; s < splits; ++s) gx += a[(size_t)s * n + i].x;
        worst = fmax(worst, fabs(gx - sx) / fabs(sx));
    }
    printf("P-i spot check: worst rel err of ax vs fp64 = %.2e\n", worst);
    const double pairs = (double)n * n;
    double t;
    t = time_pi<2, 0>(body, acc, n, splits, reps); printf("P-i u2 body-major : %8.3f ms %8.1f Gpairs/s %5.1f%%\n", t, pairs / t / 1e6, pairs / t / 1e6 * 20e-3 / 74.45 * 100);
    t = time_pi<4, 0>(body, acc, n, splits, reps); printf("P-i u4 body-major : %8.3f ms %8.1f Gpairs/s %5.1f%%\n", t, pairs / t / 1e6, pairs / t / 1e6 * 20e-3 / 74.45 * 100);
    t = time_pi<2, 1>(body, acc, n, splits, reps); printf("P-i u2 stage-major: %8.3f ms %8.1f Gpairs/s %5.1f%%\n", t, pairs / t / 1e6, pairs / t / 1e6 * 20e-3 / 74.45 * 100);
    t = time_pi<4, 1>(body, acc, n, splits, reps); printf("P-i u4 stage-major: %8.3f ms %8.1f Gpairs/s %5.1f%%\n", t, pairs / t / 1e6, pairs / t / 1e6 * 20e-3 / 74.45 * 100);
    t = time_pi<8, 1>(body, acc, n, splits, reps); printf("P-i u8 stage-major: %8.3f ms %8.1f Gpairs/s %5.1f%%\n", t, pairs / t / 1e6, pairs / t / 1e6 * 20e-3 / 74.45 * 100);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

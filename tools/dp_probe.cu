// tools/dp_probe.cu -- experiment: run part of the pairs on the FP64 pipe, next to the packed-FP32
// stream.  Per thread: R32 i-bodies in packed FP32 (j-packed lanes, as the shipped kernel) plus R64
// i-bodies in FP64 (DADD/DFMA/DMUL + rsqrt.approx.ftz.f64 = MUFU.RSQ64H).  Forces only; j tiles
// staged by plain loads (double-buffered); the FP64 bodies read a double copy of the tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/dp_probe tools/dp_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double rsqd(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

// body layout (global): float records {x0,x1,y0,y1 | z0,z1,m0,m1} per pair, and double4 {x,y,z,m} per body
template <int THREADS, int R32, int R64, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dp_kernel(const float4 *__restrict__ rec, const double4 *__restrict__ dbl,
                                                          const float4 *__restrict__ plain, float4 *__restrict__ acc, int n, int jsplits, float eps2)
{
    constexpr int TJ = 128;                                  // bodies per tile
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *ft = reinterpret_cast<float4 *>(smem);                       // [2][TJ] float4 (TJ/2 records)
    double4 *dt = reinterpret_cast<double4 *>(smem + 2 * TJ * 16);       // [2][TJ] double4
    const int tid = threadIdx.x;
    constexpr int BI = THREADS * (R32 + R64);
    const int itile = blockIdx.x / jsplits, split = blockIdx.x % jsplits;
    const int jb = (int)((long long)(n / TJ) * split / jsplits) * TJ, je = (int)((long long)(n / TJ) * (split + 1) / jsplits) * TJ;
    const int ibase = itile * BI;
    float2 nx[R32 ? R32 : 1], ny[R32 ? R32 : 1], nz[R32 ? R32 : 1], ax[R32 ? R32 : 1], ay[R32 ? R32 : 1], az[R32 ? R32 : 1];
    double dxi[R64 ? R64 : 1], dyi[R64 ? R64 : 1], dzi[R64 ? R64 : 1], dax[R64 ? R64 : 1], day[R64 ? R64 : 1], daz[R64 ? R64 : 1];
#pragma unroll
    for (int b = 0; b < R32; ++b) {
        const float4 p = plain[ibase + b * THREADS + tid];
        nx[b] = make_float2(-p.x, -p.x); ny[b] = make_float2(-p.y, -p.y); nz[b] = make_float2(-p.z, -p.z);
        ax[b] = ay[b] = az[b] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int b = 0; b < R64; ++b) {
        const float4 p = plain[ibase + (R32 + b) * THREADS + tid];
        dxi[b] = p.x; dyi[b] = p.y; dzi[b] = p.z; dax[b] = day[b] = daz[b] = 0.0;
    }
    const float2 e2 = make_float2(eps2, eps2);
    const double de2 = eps2;
    auto load_tile = [&](int t, int buf) {
        for (int k = tid; k < TJ; k += THREADS) {
            ft[buf * TJ + k] = rec[jb + t * TJ + k];
            if (R64) dt[buf * TJ + k] = dbl[jb + t * TJ + k];
        }
    };
    const int ntiles = (je - jb) / TJ;
    load_tile(0, 0);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) load_tile(t + 1, (t + 1) & 1);
        const float4 *fr = ft + (t & 1) * TJ;
        const double4 *dr = dt + (t & 1) * TJ;
#pragma unroll 1
        for (int jr = 0; jr < TJ / 2; jr += 2) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (R32) {
                    const float4 q0 = fr[2 * (jr + u)], q1 = fr[2 * (jr + u) + 1];
                    const float2 xj = make_float2(q0.x, q0.y), yj = make_float2(q0.z, q0.w), zj = make_float2(q1.x, q1.y), mj = make_float2(q1.z, q1.w);
                    float2 dx[R32 ? R32 : 1], dy[R32 ? R32 : 1], dz[R32 ? R32 : 1], s[R32 ? R32 : 1];
#pragma unroll
                    for (int b = 0; b < R32; ++b) { dx[b] = __fadd2_rn(xj, nx[b]); dy[b] = __fadd2_rn(yj, ny[b]); dz[b] = __fadd2_rn(zj, nz[b]); }
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = __ffma2_rn(dx[b], dx[b], e2);
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = __ffma2_rn(dy[b], dy[b], s[b]);
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = __ffma2_rn(dz[b], dz[b], s[b]);
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = make_float2(rsq(s[b].x), rsq(s[b].y));
#pragma unroll
                    for (int b = 0; b < R32; ++b) { const float2 i2 = __fmul2_rn(s[b], s[b]); const float2 mi = __fmul2_rn(mj, s[b]); s[b] = __fmul2_rn(i2, mi); }
#pragma unroll
                    for (int b = 0; b < R32; ++b) { ax[b] = __ffma2_rn(dx[b], s[b], ax[b]); ay[b] = __ffma2_rn(dy[b], s[b], ay[b]); az[b] = __ffma2_rn(dz[b], s[b], az[b]); }
                }
                if (R64) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double4 q = dr[2 * (jr + u) + h];
#pragma unroll
                        for (int b = 0; b < R64; ++b) {
                            const double ex = q.x - dxi[b], ey = q.y - dyi[b], ez = q.z - dzi[b];
                            double r2 = fma(ex, ex, de2); r2 = fma(ey, ey, r2); r2 = fma(ez, ez, r2);
                            const double inv = rsqd(r2);
                            const double sd = (inv * inv) * (q.w * inv);
                            dax[b] = fma(ex, sd, dax[b]); day[b] = fma(ey, sd, day[b]); daz[b] = fma(ez, sd, daz[b]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    float4 *dst = acc + (size_t)split * n;
#pragma unroll
    for (int b = 0; b < R32; ++b) dst[ibase + b * THREADS + tid] = make_float4(ax[b].x + ax[b].y, ay[b].x + ay[b].y, az[b].x + az[b].y, 0.f);
#pragma unroll
    for (int b = 0; b < R64; ++b) dst[ibase + (R32 + b) * THREADS + tid] = make_float4((float)dax[b], (float)day[b], (float)daz[b], 0.f);
}

static float4 *g_rec, *g_plain, *g_acc; static double4 *g_dbl;
template <int THREADS, int R32, int R64, int MINB>
void run(const char *name, int n, int splits, std::vector<float4> &h, bool check)
{
    constexpr int BI = THREADS * (R32 + R64);
    if (n % BI) { printf("%-40s skipped (n %% %d)\n", name, BI); return; }
    const int grid = n / BI * splits, smem = 2 * 128 * 16 + 2 * 128 * 32;
    cudaFuncSetAttribute(dp_kernel<THREADS, R32, R64, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dp_kernel<THREADS, R32, R64, MINB><<<grid, THREADS, smem>>>(g_rec, g_dbl, g_plain, g_acc, n, splits, 1e-3f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    const int reps = 4;
    for (int r = 0; r < reps; ++r) dp_kernel<THREADS, R32, R64, MINB><<<grid, THREADS, smem>>>(g_rec, g_dbl, g_plain, g_acc, n, splits, 1e-3f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dp_kernel<THREADS, R32, R64, MINB>, THREADS, smem);
    const double rate = (double)n * n / ms / 1e6;
    double worst = 0;
    if (check) {
        std::vector<float4> acc((size_t)splits * n);
        cudaMemcpy(acc.data(), g_acc, acc.size() * sizeof(float4), cudaMemcpyDeviceToHost);
        for (int i = 7; i < n; i += n / 16) {
            double sx = 0, sy = 0, sz = 0, gx = 0, gy = 0, gz = 0;
            for (int j = 0; j < n; ++j) {
                double dx = (double)h[j].x - h[i].x, dy = (double)h[j].y - h[i].y, dz = (double)h[j].z - h[i].z;
                double r2 = dx * dx + dy * dy + dz * dz + 1e-3, w = h[j].w / (r2 * sqrt(r2));
                sx += dx * w; sy += dy * w; sz += dz * w;
            }
            for (int s = 0; s < splits; ++s) { gx += acc[(size_t)s * n + i].x; gy += acc[(size_t)s * n + i].y; gz += acc[(size_t)s * n + i].z; }
            worst = fmax(worst, sqrt((gx - sx) * (gx - sx) + (gy - sy) * (gy - sy) + (gz - sz) * (gz - sz)) / sqrt(sx * sx + sy * sy + sz * sz));
        }
    }
    printf("%-40s occ=%d %8.3f ms %8.1f Gpairs/s %5.1f%% of FP32 peak   worst rel force err vs fp64 %.1e\n", name, occ, ms, rate, rate * 20e-3 / 74.45 * 100, worst);
}


template <int T32, int T64, int R32, int R64>
__global__ void __launch_bounds__(T32 + T64, 1) ws_kernel(const float4 *__restrict__ rec, const double4 *__restrict__ dbl,
                                                         const float4 *__restrict__ plain, float4 *__restrict__ acc, int n, int jsplits, float eps2)
{
    constexpr int TJ = 128, THREADS = T32 + T64;
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *ft = reinterpret_cast<float4 *>(smem);
    double4 *dt = reinterpret_cast<double4 *>(smem + 2 * TJ * 16);
    const int tid = threadIdx.x;
    constexpr int BI = T32 * R32 + T64 * R64;
    const int itile = blockIdx.x / jsplits, split = blockIdx.x % jsplits;
    const int jb = (int)((long long)(n / TJ) * split / jsplits) * TJ, je = (int)((long long)(n / TJ) * (split + 1) / jsplits) * TJ;
    const int ibase = itile * BI;
    const bool is32 = tid < T32;
    float2 nx[R32], ny[R32], nz[R32], ax[R32], ay[R32], az[R32];
    double dxi[R64], dyi[R64], dzi[R64], dax[R64], day[R64], daz[R64];
    if (is32) {
#pragma unroll
        for (int b = 0; b < R32; ++b) {
            const float4 p = plain[ibase + b * T32 + tid];
            nx[b] = make_float2(-p.x, -p.x); ny[b] = make_float2(-p.y, -p.y); nz[b] = make_float2(-p.z, -p.z);
            ax[b] = ay[b] = az[b] = make_float2(0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int b = 0; b < R64; ++b) {
            const float4 p = plain[ibase + T32 * R32 + b * T64 + (tid - T32)];
            dxi[b] = p.x; dyi[b] = p.y; dzi[b] = p.z; dax[b] = day[b] = daz[b] = 0.0;
        }
    }
    const float2 e2 = make_float2(eps2, eps2);
    const double de2 = eps2;
    auto load_tile = [&](int t, int buf) {
        for (int k = tid; k < TJ; k += THREADS) { ft[buf * TJ + k] = rec[jb + t * TJ + k]; dt[buf * TJ + k] = dbl[jb + t * TJ + k]; }
    };
    const int ntiles = (je - jb) / TJ;
    load_tile(0, 0);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) load_tile(t + 1, (t + 1) & 1);
        const float4 *fr = ft + (t & 1) * TJ;
        const double4 *dr = dt + (t & 1) * TJ;
        if (is32) {
#pragma unroll 1
            for (int jr = 0; jr < TJ / 2; jr += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float4 q0 = fr[2 * (jr + u)], q1 = fr[2 * (jr + u) + 1];
                    const float2 xj = make_float2(q0.x, q0.y), yj = make_float2(q0.z, q0.w), zj = make_float2(q1.x, q1.y), mj = make_float2(q1.z, q1.w);
                    float2 dx[R32], dy[R32], dz[R32], s[R32];
#pragma unroll
                    for (int b = 0; b < R32; ++b) { dx[b] = __fadd2_rn(xj, nx[b]); dy[b] = __fadd2_rn(yj, ny[b]); dz[b] = __fadd2_rn(zj, nz[b]); }
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = __ffma2_rn(dx[b], dx[b], e2);
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = __ffma2_rn(dy[b], dy[b], s[b]);
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = __ffma2_rn(dz[b], dz[b], s[b]);
#pragma unroll
                    for (int b = 0; b < R32; ++b) s[b] = make_float2(rsq(s[b].x), rsq(s[b].y));
#pragma unroll
                    for (int b = 0; b < R32; ++b) { const float2 i2 = __fmul2_rn(s[b], s[b]); const float2 mi = __fmul2_rn(mj, s[b]); s[b] = __fmul2_rn(i2, mi); }
#pragma unroll
                    for (int b = 0; b < R32; ++b) { ax[b] = __ffma2_rn(dx[b], s[b], ax[b]); ay[b] = __ffma2_rn(dy[b], s[b], ay[b]); az[b] = __ffma2_rn(dz[b], s[b], az[b]); }
                }
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < TJ; j += 2) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double4 q = dr[j + h];
#pragma unroll
                    for (int b = 0; b < R64; ++b) {
                        const double ex = q.x - dxi[b], ey = q.y - dyi[b], ez = q.z - dzi[b];
                        double r2 = fma(ex, ex, de2); r2 = fma(ey, ey, r2); r2 = fma(ez, ez, r2);
                        const double inv = rsqd(r2);
                        const double sd = (inv * inv) * (q.w * inv);
                        dax[b] = fma(ex, sd, dax[b]); day[b] = fma(ey, sd, day[b]); daz[b] = fma(ez, sd, daz[b]);
                    }
                }
            }
        }
        __syncthreads();
    }
    float4 *dst = acc + (size_t)split * n;
    if (is32) {
#pragma unroll
        for (int b = 0; b < R32; ++b) dst[ibase + b * T32 + tid] = make_float4(ax[b].x + ax[b].y, ay[b].x + ay[b].y, az[b].x + az[b].y, 0.f);
    } else {
#pragma unroll
        for (int b = 0; b < R64; ++b) dst[ibase + T32 * R32 + b * T64 + (tid - T32)] = make_float4((float)dax[b], (float)day[b], (float)daz[b], 0.f);
    }
}

template <int T32, int T64, int R32, int R64>
void run_ws(const char *name, int n, int splits)
{
    constexpr int BI = T32 * R32 + T64 * R64;
    if (n % BI) { printf("%-44s skipped (n %% %d)\n", name, BI); return; }
    const int grid = n / BI * splits, smem = 2 * 128 * 16 + 2 * 128 * 32;
    cudaFuncSetAttribute(ws_kernel<T32, T64, R32, R64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ws_kernel<T32, T64, R32, R64><<<grid, T32 + T64, smem>>>(g_rec, g_dbl, g_plain, g_acc, n, splits, 1e-3f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    const int reps = 4;
    for (int r = 0; r < reps; ++r) ws_kernel<T32, T64, R32, R64><<<grid, T32 + T64, smem>>>(g_rec, g_dbl, g_plain, g_acc, n, splits, 1e-3f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ws_kernel<T32, T64, R32, R64>, T32 + T64, smem);
    const double rate = (double)n * n / ms / 1e6;
    printf("%-44s occ=%d %8.3f ms %8.1f Gpairs/s %5.1f%% of FP32 peak (%s)\n", name, occ, ms, rate, rate * 20e-3 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 645120;   // divisible by every tile size below      // divisible by 256*{4,5,6}, 384*{4,5,6}, 128
    std::vector<float4> h(n), rec(n); std::vector<double4> dbl(n);
    srand(1);
    for (auto &b : h) b = make_float4(rand() / (float)RAND_MAX, rand() / (float)RAND_MAX, rand() / (float)RAND_MAX, 1e-4f * rand() / (float)RAND_MAX);
    for (int i = 0; i < n; i += 2) {
        rec[i] = make_float4(h[i].x, h[i + 1].x, h[i].y, h[i + 1].y);
        rec[i + 1] = make_float4(h[i].z, h[i + 1].z, h[i].w, h[i + 1].w);
    }
    for (int i = 0; i < n; ++i) dbl[i] = make_double4(h[i].x, h[i].y, h[i].z, h[i].w);
    const int splits = 30;
    cudaMalloc(&g_rec, n * sizeof(float4)); cudaMalloc(&g_plain, n * sizeof(float4)); cudaMalloc(&g_dbl, n * sizeof(double4)); cudaMalloc(&g_acc, (size_t)splits * n * sizeof(float4));
    cudaMemcpy(g_rec, rec.data(), n * sizeof(float4), cudaMemcpyHostToDevice);
    cudaMemcpy(g_plain, h.data(), n * sizeof(float4), cudaMemcpyHostToDevice);
    cudaMemcpy(g_dbl, dbl.data(), n * sizeof(double4), cudaMemcpyHostToDevice);
    run<256, 4, 0, 2>("FP32 only  t256 R32=4 (2 CTAs/SM)", n, splits, h, true);
    run<256, 0, 2, 2>("FP64 only  t256 R64=2", n, splits, h, true);
    run<256, 4, 1, 1>("mixed      t256 R32=4 R64=1 (1 CTA/SM)", n, splits, h, true);
    run<256, 4, 2, 1>("mixed      t256 R32=4 R64=2 (1 CTA/SM)", n, splits, h, true);
    run<384, 4, 1, 1>("mixed      t384 R32=4 R64=1 (1 CTA/SM)", n, splits, h, false);
    run<384, 4, 2, 1>("mixed      t384 R32=4 R64=2 (1 CTA/SM)", n, splits, h, false);
    run<512, 4, 1, 1>("mixed      t512 R32=4 R64=1 (1 CTA/SM)", n, splits, h, false);
    run<256, 2, 1, 2>("mixed      t256 R32=2 R64=1 (2 CTAs/SM)", n, splits, h, false);
    run<128, 4, 2, 3>("mixed      t128 R32=4 R64=2 (3 CTAs/SM)", n, splits, h, false);
    run<128, 4, 1, 3>("mixed      t128 R32=4 R64=1 (3 CTAs/SM)", n, splits, h, false);
    run_ws<256, 128, 4, 2>("warp-spec 8 FP32 warps (R4) + 4 FP64 warps (R2)", n, splits);
    run_ws<256, 256, 4, 2>("warp-spec 8 FP32 warps (R4) + 8 FP64 warps (R2)", n, splits);
    run_ws<256, 256, 4, 1>("warp-spec 8 FP32 warps (R4) + 8 FP64 warps (R1)", n, splits);
    run_ws<256, 128, 4, 4>("warp-spec 8 FP32 warps (R4) + 4 FP64 warps (R4)", n, splits);
    run_ws<384, 128, 4, 2>("warp-spec 12 FP32 warps (R4) + 4 FP64 warps (R2)", n, splits);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

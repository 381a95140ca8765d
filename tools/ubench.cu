// tools/ubench.cu -- what can the sm_100a FP32 pipe sustain with packed (f32x2) instructions,
// and what does a MUFU.RSQ / LDS / non-FMA instruction cost it?  Answers the ceiling question for
// the step kernel's inner loop (12 packed FP32 + 2 MUFU per (i, j-pair)).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ACC 12

__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: pure FFMA2            MODE 1: 6 FFMA2 : 1 MUFU      MODE 2: 12 FFMA2 : 1 MUFU
// MODE 3: pure scalar FFMA      MODE 4: 12 FFMA : 1 MUFU       MODE 5: 24 FFMA2 : 1 LDS.128
// MODE 6: FADD2:FFMA2:FMUL2 1:2:1 pure                          MODE 7: 6 FFMA2 : 1 MUFU, 3-operand distinct
// MODE 8: 6 FFMA2 : 1 FMNMX (ALU pipe)   MODE 9: 3 FFMA2 : 1 MUFU
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float *out, int iters, float a0, float b0)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a0, b0, a0, b0);
    __syncthreads();
    float2 acc[ACC];
    float sacc[2 * ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); sacc[2 * i] = acc[i].x; sacc[2 * i + 1] = acc[i].y; }
    float2 a = make_float2(a0, a0 * 1.0001f), b = make_float2(b0, b0 * 0.999f);
    float m0 = 1.5f + threadIdx.x, m1 = 2.5f, m2 = 3.5f, m3 = 4.5f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 5 || MODE == 8 || MODE == 9) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < ACC; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
                if (MODE == 1) { m0 = rsq(m0); m1 = rsq(m1); }            // 12 : 2
                if (MODE == 2 && r == 0) { m0 = rsq(m0); }                // 24 : 2 -> with r: 24:... (one per 2 rounds => 24:1?) keep 12:1 below
                if (MODE == 2 && r == 1) { m1 = rsq(m1); }
                if (MODE == 9) { m0 = rsq(m0); m1 = rsq(m1); m2 = rsq(m2); m3 = rsq(m3); }   // 12 : 4
                if (MODE == 8) { m0 = fmaxf(m0, m1 + 0.f); m1 = fminf(m1, m2); }
            }
            if (MODE == 5) { float4 q = sm[(it & 63)]; a.x += q.x * 1e-30f; }
        } else if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 2 * ACC; ++i) sacc[i] = fmaf(sacc[i], a.x, b.x);
            if (MODE == 4) { m0 = rsq(m0); m1 = rsq(m1); }               // 24 scalar : 2
        } else if (MODE == 6) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < ACC; i += 4) {
                    acc[i] = __fadd2_rn(acc[i], a);
                    acc[i + 1] = __ffma2_rn(acc[i + 1], a, b);
                    acc[i + 2] = __ffma2_rn(acc[i + 2], b, a);
                    acc[i + 3] = __fmul2_rn(acc[i + 3], a);
                }
            }
        } else if (MODE == 7) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < ACC; ++i) acc[i] = __ffma2_rn(acc[(i + 1) % ACC], acc[(i + 5) % ACC], acc[i]);
                m0 = rsq(m0); m1 = rsq(m1);
            }
        }
    }
    float s = m0 + m1 + m2 + m3 + a.x;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i].x + acc[i].y + sacc[2 * i] + sacc[2 * i + 1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double fma_lane_ops_per_iter_per_thread, float *out, int sms, double ghz)
{
    const int iters = 20000;
    const int grid = sms * 2;
    k<MODE><<<grid, 256>>>(out, 100, 1.0001f, 1e-7f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 1e-7f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = fma_lane_ops_per_iter_per_thread * iters * 256.0 * grid;
    const double per_clk_sm = lane_ops / (ms * 1e-3 * ghz * 1e9) / sms;
    printf("%-44s %8.3f ms  %7.2f FP32 lane-ops/clk/SM  (%5.1f%% of 128)\n", name, ms, per_clk_sm, per_clk_sm / 128 * 100);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
    float *out; cudaMalloc(&out, sizeof(float) * 256 * sms * 2);
    run<0>("pure FFMA2", 2.0 * 2 * ACC, out, sms, ghz);
    run<1>("FFMA2 : MUFU = 6 : 1", 2.0 * 2 * ACC, out, sms, ghz);
    run<2>("FFMA2 : MUFU = 12 : 1", 2.0 * 2 * ACC, out, sms, ghz);
    run<9>("FFMA2 : MUFU = 3 : 1", 2.0 * 2 * ACC, out, sms, ghz);
    run<3>("pure scalar FFMA", 2.0 * ACC, out, sms, ghz);
    run<4>("scalar FFMA : MUFU = 12 : 1", 2.0 * ACC, out, sms, ghz);
    run<5>("FFMA2 : LDS.128 = 24 : 1", 2.0 * 2 * ACC, out, sms, ghz);
    run<6>("FADD2:FFMA2:FMUL2 = 1:2:1", 2.0 * 2 * ACC, out, sms, ghz);
    run<7>("FFMA2 (3 distinct regs) : MUFU = 6 : 1", 2.0 * 2 * ACC, out, sms, ghz);
    run<8>("FFMA2 : FMNMX(alu) = 6 : 1", 2.0 * 2 * ACC, out, sms, ghz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

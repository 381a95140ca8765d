// nbx_kernels.cuh -- sm_100a kernels for the O(N^2) force + Euler + kinetic-energy step.
//
// What the reference does per step (ver2/GSimulation.cpp:130-173, float):
//   a_i  = sum_j G m_j (r_j - r_i) (|r_j - r_i|^2 + eps2)^(-3/2)     (all j, self term = 0)
//   v_i += a_i dt ; r_i += v_i dt ; kenergy = 0.5 sum_i m_i |v_i|^2
// and its only CUDA kernel (ver5_all/programming_models/cuda/Compute.cu:31-66) is one
// thread per i streaming j from global memory, with the update done on the host.
//
// This file is a different design, for B200:
//   * bodies live in HBM "pair-packed": one 32-byte record per two bodies,
//     {x0,x1,y0,y1 | z0,z1,Gm0,Gm1}.  A record is at once the TMA unit, the operand
//     layout of Blackwell's packed FP32 instructions (FFMA2/FADD2/FMUL2: two lanes of
//     a 64-bit register pair) and the multi-GPU exchange unit;
//   * each CTA streams j-records through a ring of shared-memory stages filled by
//     1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 4 stages deep;
//   * each thread register-blocks R = 2*R2 i-bodies and evaluates every (i, j-pair)
//     with 12 packed FP32 instructions + 2 MUFU.RSQ, i.e. 12 FP32-pipe lane-ops and
//     7 issue slots per pair instead of 13 -- the FP32 pipe, not issue, is the limit;
//   * from 65 536 bodies on the default is the q-scaled pair (the QS block in step_kernel; shapes
//     "_qi"): j-records pre-multiplied by (G m_j)^(-1/2) -- rewritten once per step by qscale_kernel --
//     make the subtract an FMA and drop the multiply by G m: 11 instructions, and with the packed
//     lanes over the two i-bodies of a record the j data are scalars, so only the three
//     accumulates read three distinct 64-bit registers (the register file has two banks: an
//     instruction holds its slot for max(pipe cycles, distinct even sources, distinct odd sources));
//   * float sums are kept short: the lane accumulators are folded into a second float per
//     (body, component) in shared memory every 64 j tiles (two-level accumulation, the default;
//     a single accumulator over 5e5 terms is biased low by 2e-5 .. 7e-5 at N = 1 M);
//   * the Euler update, the kinetic-energy reduction (deterministic: per-CTA partial,
//     last CTA sums in tile order) and, on several GPUs, the NVLink stores of the
//     updated records into every peer's replica -- one NVSwitch multicast store per half
//     record where a multicast team exists -- all run in the same kernel's epilogue; the
//     wait for the peers' previous step at the top of the kernel is bounded by %globaltimer;
//   * a j-split with a last-arriver combine in fixed split order fills the 148 SMs: the first
//     `whole_tiles` i-tiles (whole rounds of the SM count, when there are at least three) run
//     unsplit, the remaining tail tiles are cut `j_splits` ways along j.  (A persistent stream-K
//     grid -- equal spans of the tile-major (i-tile, j-chunk) work, one CTA per SM -- measured
//     within +-1% of this at N = 2000 ... 131 072 and was dropped.)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nbx {

constexpr int kMaxWorld = 8;

struct StepParams {
    const float4 *pos_in;   // pair-packed records, n_pad bodies = n_pad float4s
    float4 *pos_out;        // same layout, the other half of the ping-pong
    float4 *vel;            // shard-local: (vx, vy, vz, m) per body
    float4 *part;           // [j_splits][split_bodies] partial accelerations of the split tiles
    int *tile_ticket;       // [i_tiles - whole_tiles] arrival counters (self-resetting)
    double *ke_part;        // [i_tiles]
    int *ke_ticket;         // [1]
    double *ke_out;         // [steps] kinetic energy, slot = *dev_step
    int *dev_step;          // step index inside the current run (reset by the host per run)
    int ke_cap;             // slots in ke_out
    int *dev_epoch;         // steps completed since create (never reset; P2P flag value)
    int *dev_err;           // device error word (0 = fine): peer timeout, host abort, debug-build check
    float4 *acc_out;        // non-null: store accelerations, do not update (nbx_accelerations)
    int n_pad;
    int i_begin, i_count;
    int i_tiles;            // CTAs' worth of i-bodies in this shard
    int whole_tiles;        // tiles [0, whole_tiles) are not split along j
    int j_splits;           // split count of tiles [whole_tiles, i_tiles) in THIS launch
    int split_bodies;       // i-bodies covered by split tiles (stride of `part`)
    int discard_partials;   // 1: the combining CTA drops the consumed partial lines from L2 (large tails only)
    // A launch may cover only a window of the j-bodies (NCCL-overlap mode runs a step as two
    // launches: the rank's own j-shard while the all-gather is in flight, then the rest):
    int j_org, j_len;       // j window = [j_org, j_org + j_len) modulo n_pad (multiples of 8)
    int split_base;         // partial slot of this launch's split 0
    int split_total;        // contributors per split tile over all launches of the step
    float dt, eps2;
    // P2P exchange (world > 1 and exchange == P2P): peers' replicas and completion flags
    int world, rank, p2p;
    float4 *peer_pos_out[kMaxWorld];  // [g] = rank g's pos_out (self entry unused)
    float4 *mc_pos_out;               // non-null: NVSwitch multicast mapping of pos_out (one store reaches every GPU)
    int *peer_flags[kMaxWorld];       // [g] = rank g's flags[kMaxWorld]; we write slot [rank]
    const int *my_flags;              // this rank's flags[kMaxWorld]
    unsigned long long peer_wait_ns;  // how long a step may wait for a peer's previous step
    // Trace build (-DNBX_TRACE, libnbx_trace.so) only: per-CTA %globaltimer stamps, kTraceWords per CTA per step
    unsigned long long *trace;
    int trace_steps;                  // steps the buffer holds (later steps are not recorded)
    // q-scaled shapes (MATH bit kMathQScale) only: the j-records in the form the 11-instruction pair needs,
    // 48 bytes per body pair, rewritten from pos_in by qscale_kernel before every step launch
    float4 *qrec;
};
constexpr int kMathQScale = 1 << 22;
constexpr int kTraceWords = 6;        // start, first tile landed, sweep done, exit, smid, last-arriver flag

// Values of *dev_err (low byte; the rest carries detail: peer rank << 8, or source line << 8).
enum { kDevErrPeerTimeout = 1, kDevErrHostAbort = 2, kDevErrDebugCheck = 3 };

// ------------------------------------------------------------------------------
//  PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.RSQ; x >= eps2 > 0 always
    return y;
}
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// NVSwitch multicast stores (SASS: the store carries the multimem qualifier): one instruction, every GPU
// of the multicast team receives the data in its own copy of the buffer.
__device__ __forceinline__ void multimem_st_v4(float4 *mc_ptr, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void multimem_st_v2(float2 *mc_ptr, float2 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc_ptr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_volatile(const int *p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dev_fail(int *err, int code) { atomicCAS(err, 0, code); }

// Debug build (make DEBUG=1 -> libnbx_debug.so): every index the kernel stores through and every
// ticket value is range-checked; a failed check records its source line in *dev_err and nbx_run
// returns NBX_ERR_DEBUG.  (compute-sanitizer is closed on the GPU pool this was developed on.)
#ifdef NBX_DEBUG
#define NBX_CHECK(cond)                                                                  \
    do {                                                                                 \
        if (!(cond)) nbx::dev_fail(p.dev_err, nbx::kDevErrDebugCheck | (__LINE__ << 8)); \
    } while (0)
#else
#define NBX_CHECK(cond) ((void)0)
#endif

// Packed (FADD2/FMUL2/FFMA2) or scalar form of a two-lane FP32 op.
template <bool SCALAR> __device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    if (SCALAR) return make_float2(a.x + b.x, a.y + b.y);
    return __fadd2_rn(a, b);
}
template <bool SCALAR> __device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    if (SCALAR) return make_float2(a.x * b.x, a.y * b.y);
    return __fmul2_rn(a, b);
}
template <bool SCALAR> __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    if (SCALAR) return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
    return __ffma2_rn(a, b, c);
}

// Shared-memory footprint of one CTA (host uses the same formula).  MATH bit 32 ("acc64") adds
// one double per (i-body, component) per thread: the second accumulation level; q-scaled shapes
// stream 24 instead of 16 bytes per j-body.
template <int THREADS, int TJ, int STAGES, int R2 = 0, int MATH = 0>
__host__ __device__ constexpr int step_smem_bytes()
{
    return STAGES * TJ * ((MATH & kMathQScale) ? 24 : 16) + 2 * STAGES * 8 + (THREADS / 32) * 8 + 16 +
           ((MATH & 32) ? 3 * 2 * R2 * THREADS * 8 : 0) +
           ((MATH & 256) ? 3 * 2 * R2 * THREADS * 4 : 0);
}

// ------------------------------------------------------------------------------
//  The step kernel.
//    R2      i-body PAIRS register-blocked per thread (R = 2*R2 bodies)
//    THREADS CTA size;  BI = THREADS*R i-bodies per CTA
//    TJ      j-bodies per TMA stage (multiple of 8)
//    STAGES  ring depth (>= 3: one being read, one landed, one in flight)
//    UNROLL  j-records per inner-loop trip (1, 2 or 4)
//    MINB    __launch_bounds__ min CTAs per SM
//  grid = whole_tiles + (i_tiles - whole_tiles) * j_splits CTAs, whole tiles first
// ------------------------------------------------------------------------------
template <int R2, int THREADS, int TJ, int STAGES, int UNROLL, int MINB, int MATH = 0>
__global__ void __launch_bounds__(THREADS, MINB) step_kernel(const __grid_constant__ StepParams p)
{
    constexpr int R = 2 * R2;
    constexpr int WARPS = THREADS / 32;
    static_assert(TJ % 8 == 0 && STAGES >= 3 && (UNROLL == 1 || UNROLL == 2 || UNROLL == 4), "shape");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
    constexpr bool QS = (MATH & kMathQScale) != 0;
    constexpr int JB = QS ? 24 : 16;                         // bytes per j-body in the ring
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TJ * JB);
    uint64_t *empty = full + STAGES;
    double *red = reinterpret_cast<double *>(empty + STAGES);
    int *s_flag = reinterpret_cast<int *>(red + WARPS);
    // acc64 (MATH & 32): accuracy option.  Each thread folds its float lane sums into
    // a double per (body, component) after every 4th j tile, so no float sum is longer than 2*TJ terms:
    // the large-N float summation error (1e-4 at 1 M, 1e-3 at 4 M for the reference's single
    // accumulator) drops to the 1e-6 level for ~1% time.  hi[q * THREADS + tid]: conflict-free.
    double *hi = reinterpret_cast<double *>(smem_raw + step_smem_bytes<THREADS, TJ, STAGES, 0, MATH & kMathQScale>());
    // f2 (MATH & 256): two-level FLOAT accumulation, the default.  A single float accumulator per lane is
    // not just noisy at large N, it is BIASED: summed over 5e5 terms the force magnitude comes out
    // systematically low (N = 1 M: -2e-5 here, -7e-5 for the reference's own single-accumulator float
    // loop, measured against fp64 -- tests/golden/truth_*), and the kinetic energy inherits twice that.
    // Folding the lane sums into a second float per (body, component) in shared memory after every 4th
    // j tile keeps every float sum <= 1024 terms long; cost: 12 LDS/FADD/STS per 24 576 FP32 instructions.
    float *hif = reinterpret_cast<float *>(smem_raw + step_smem_bytes<THREADS, TJ, STAGES, 0, MATH & kMathQScale>());

    const int tid = threadIdx.x;
#ifdef NBX_TRACE
    unsigned long long *tr = nullptr;
    if (tid == 0 && p.trace != nullptr) {
        const int step = ld_volatile(p.dev_step);
        if (step < p.trace_steps) {
            tr = p.trace + ((size_t)step * gridDim.x + blockIdx.x) * kTraceWords;
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            tr[0] = globaltimer_ns(); tr[1] = tr[2] = tr[3] = 0; tr[4] = smid; tr[5] = 0;
        }
    }
#define NBX_STAMP(k) do { if (tid == 0 && tr) tr[k] = globaltimer_ns(); } while (0)
#define NBX_MARK(k, v) do { if (tid == 0 && tr) tr[k] = (v); } while (0)
#else
#define NBX_STAMP(k) ((void)0)
#define NBX_MARK(k, v) ((void)0)
#endif
    int tile = blockIdx.x, split = 0, nsplit = 1, contributors = 1;
    if (tile >= p.whole_tiles) {
        const int r = tile - p.whole_tiles;
        tile = p.whole_tiles + r / p.j_splits;
        split = r % p.j_splits;
        nsplit = p.j_splits;
        contributors = p.split_total;
    }

    // ---- j range of this CTA inside the launch's window, in 8-body chunks so every TMA copy
    //      is 128-byte granular
    const int chunks = p.j_len >> 3;
    const int jb = (int)(((long long)chunks * split) / nsplit) << 3;
    const int je = (int)(((long long)chunks * (split + 1)) / nsplit) << 3;
    const int ntiles = (je - jb + TJ - 1) / TJ;

    // Programmatic dependent launch: let the next step's grid be scheduled as soon as SM resources
    // free up, and do our own set-up (barrier init) before waiting for the previous step's grid --
    // only what follows griddepcontrol.wait may read what that grid wrote.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], WARPS);
        }
        mbar_fence_init();
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
        // A poisoned context (peer timeout, host abort, failed debug check) drains: every CTA of
        // every queued step returns at once.
#ifdef NBX_DEBUG
        int bad = ld_volatile(p.dev_err);
#else
        int bad = p.p2p ? ld_volatile(p.dev_err) : 0;   // single-GPU runs never set it: skip the load
#endif
        // P2P exchange: every rank must have finished the previous step (its epilogue wrote
        // into OUR pos_in) before we read it.  Peers run on other GPUs; no kernel on this
        // GPU is waited on.  The wait is bounded: a dead or stuck peer turns into an error
        // word that nbx_run reports as NBX_ERR_PEER instead of a hang.
        if (p.p2p && !bad) {
            const int epoch = *p.dev_epoch;
            for (int g = 0; g < p.world && !bad; ++g) {
                if (g == p.rank || ld_acquire_sys(&p.my_flags[g]) >= epoch) continue;
                const unsigned long long t0 = globaltimer_ns();
                while (ld_acquire_sys(&p.my_flags[g]) < epoch) {
                    if ((bad = ld_volatile(p.dev_err)) != 0) break;
                    if (globaltimer_ns() - t0 > p.peer_wait_ns) {
                        dev_fail(p.dev_err, kDevErrPeerTimeout | (g << 8));
                        bad = 1;
                        break;
                    }
                    __nanosleep(200);
                }
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        *s_flag = bad;
    }
    __syncthreads();
    if (*s_flag) return;

    auto issue_tile = [&](int t) {
        const int cnt = min(TJ, je - (jb + t * TJ));
        int j0 = p.j_org + jb + t * TJ;                    // window may wrap around n_pad
        if (j0 >= p.n_pad) j0 -= p.n_pad;
        const int head = min(cnt, p.n_pad - j0);
        const int st = t % STAGES;
        NBX_CHECK(cnt > 0 && (cnt & 7) == 0 && j0 >= 0 && j0 + head <= p.n_pad && (j0 & 7) == 0);
        if (QS) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(p.qrec);
            unsigned char *dst = smem_raw + st * (TJ * 24);
            mbar_expect_tx(&full[st], (uint32_t)cnt * 24u);
            tma_load_1d(dst, src + (size_t)j0 * 24, (uint32_t)head * 24u, &full[st]);
            if (head < cnt) tma_load_1d(dst + head * 24, src, (uint32_t)(cnt - head) * 24u, &full[st]);
            return;
        }
        mbar_expect_tx(&full[st], (uint32_t)cnt * 16u);
        tma_load_1d(tiles + st * TJ, p.pos_in + j0, (uint32_t)head * 16u, &full[st]);
        if (head < cnt) tma_load_1d(tiles + st * TJ + head, p.pos_in, (uint32_t)(cnt - head) * 16u, &full[st]);
    };
    if (tid == 0) {
#pragma unroll
        for (int t = 0; t < STAGES - 2; ++t)
            if (t < ntiles) issue_tile(t);
    }

    // ---- my i-bodies: R2 records, negated and duplicated for the packed subtract
    float2 nx[R], ny[R], nz[R];
    float2 ax[R], ay[R], az[R];
    const int pair_base = tile * (THREADS * R2) + tid;   // shard-local record index of k = 0
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        int ip = pair_base + k * THREADS;
        ip = min(ip, (p.i_count >> 1) - 1);               // clamp: tail threads redo the last record
        const size_t gp = (size_t)(p.i_begin >> 1) + ip;
        // plain (L2) loads, not ld.global.nc: under programmatic dependent launch this kernel's
        // lifetime overlaps the previous step's grid, which was still writing this buffer
        NBX_CHECK(ip >= 0 && 2 * gp + 1 < (size_t)p.n_pad);
        const float4 q0 = __ldcg(&p.pos_in[2 * gp]);
        const float4 q1 = __ldcg(&p.pos_in[2 * gp + 1]);
        if (QS) {   // lanes = the two bodies of the record
            nx[k] = make_float2(-q0.x, -q0.y); ny[k] = make_float2(-q0.z, -q0.w); nz[k] = make_float2(-q1.x, -q1.y);
            continue;
        }
        nx[2 * k] = make_float2(-q0.x, -q0.x); nx[2 * k + 1] = make_float2(-q0.y, -q0.y);
        ny[2 * k] = make_float2(-q0.z, -q0.z); ny[2 * k + 1] = make_float2(-q0.w, -q0.w);
        nz[2 * k] = make_float2(-q1.x, -q1.x); nz[2 * k + 1] = make_float2(-q1.y, -q1.y);
    }
#pragma unroll
    for (int b = 0; b < R; ++b) ax[b] = ay[b] = az[b] = make_float2(0.f, 0.f);
    float2 as[(MATH & 128) ? R : 1];
#pragma unroll
    for (int b = 0; b < ((MATH & 128) ? R : 1); ++b) as[b] = make_float2(0.f, 0.f);
    const float2 eps2v = make_float2(p.eps2, p.eps2);

    // ---- sweep the j tiles
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0) {
            const int u = t + STAGES - 2;                  // refill two tiles behind the reader
            if (u < ntiles) {
                if (u >= STAGES) mbar_wait(&empty[u % STAGES], ((u / STAGES) - 1) & 1);
                issue_tile(u);
            }
        }
        const int st = t % STAGES;
        mbar_wait(&full[st], (t / STAGES) & 1);
        if (t == 0) NBX_STAMP(1);
        const float4 *rec = QS ? reinterpret_cast<const float4 *>(smem_raw + st * (TJ * 24)) : tiles + st * TJ;
        const int nrec = min(TJ, je - (jb + t * TJ)) >> 1;  // records in this tile (multiple of 4)
#pragma unroll 1
        for (int jr = 0; jr < nrec; jr += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (QS) {
                    // q-scaled pair, lanes packed over i-bodies (11 packed FP32 instructions instead of 12, and only
                    // the three accumulates read three distinct 64-bit registers).  With q_j = (G m_j)^(-1/2) the
                    // j-record holds the scalars q x_j, q, q y_j, q^2 eps, q z_j; per (i-pair, j):
                    //     e = q (r_j - r_i) = fma(-r_i, q, q r_j)                3 FFMA2: a pair and two 32-bit scalars of
                    //                                                            opposite register parity = 2 + 2 bank reads
                    //     w = e.e + q^2 eps = q^2 (|d|^2 + eps)                  3 FFMA2
                    //     u = rsqrt(w)^3    = (G m)^(3/2) (|d|^2 + eps)^(-3/2)   2 MUFU + 2 FMUL2 (no "times G m")
                    //     a += u e          = G m (|d|^2 + eps)^(-3/2) d         3 FFMA2
                    // Record pair = 48 bytes = {qx0,q0,qy0,qe0 | qz0,-,qx1,q1 | qy1,qe1,qz1,-}: in every LDS.128 the
                    // q of a body sits in an odd register and its coordinates in even ones.
                    const float4 *r3 = rec + 3 * (jr + u);
                    const float4 c0 = r3[0], c1 = r3[1], c2 = r3[2];
                    constexpr int QP = (MATH >> 12) & 1023;     // source-order bits
                    constexpr int NU = 2 * R2;                  // (i-pair, j-body) units per record
                    float2 ex[NU], ey[NU], ez[NU], sv[NU];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float sq = h ? c1.w : c0.y, sx = h ? c1.z : c0.x, sy = h ? c2.x : c0.z, sz = h ? c2.z : c1.x;
                        const float2 qq = make_float2(sq, sq), qx = make_float2(sx, sx), qy = make_float2(sy, sy), qz = make_float2(sz, sz);
#pragma unroll
                        for (int k = 0; k < R2; ++k) {
                            const int n = (QP & 1) ? k * 2 + h : h * R2 + k;
                            ex[n] = __ffma2_rn(nx[k], qq, qx); ey[n] = __ffma2_rn(ny[k], qq, qy); ez[n] = __ffma2_rn(nz[k], qq, qz);
                        }
                    }
                    auto body_of = [](int n) { return (QP & 1) ? (n & 1) : n / R2; };
                    auto pair_of = [](int n) { return (QP & 1) ? (n >> 1) : n % R2; };
                    auto ord = [](int bit, int n) { return (QP & bit) ? NU - 1 - n : n; };
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        const int n = ord(32, i);
                        const float se = body_of(n) ? c2.y : c0.w;
                        sv[n] = __ffma2_rn(ex[n], ex[n], make_float2(se, se));
                    }
#pragma unroll
                    for (int i = 0; i < NU; ++i) sv[ord(32, i)] = __ffma2_rn(ey[ord(32, i)], ey[ord(32, i)], sv[ord(32, i)]);
#pragma unroll
                    for (int i = 0; i < NU; ++i) sv[ord(32, i)] = __ffma2_rn(ez[ord(32, i)], ez[ord(32, i)], sv[ord(32, i)]);
#pragma unroll
                    for (int i = 0; i < NU; ++i) sv[ord(64, i)] = make_float2(rsqrt_approx(sv[ord(64, i)].x), rsqrt_approx(sv[ord(64, i)].y));
#pragma unroll
                    for (int i = 0; i < NU; ++i) sv[ord(128, i)] = __fmul2_rn(__fmul2_rn(sv[ord(128, i)], sv[ord(128, i)]), sv[ord(128, i)]);
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        const int n = ord(8, i), k = pair_of(n);
                        if (QP & 256) {
                            ax[k] = __ffma2_rn(sv[n], ex[n], ax[k]); ay[k] = __ffma2_rn(sv[n], ey[n], ay[k]); az[k] = __ffma2_rn(sv[n], ez[n], az[k]);
                        } else {
                            ax[k] = __ffma2_rn(ex[n], sv[n], ax[k]); ay[k] = __ffma2_rn(ey[n], sv[n], ay[k]); az[k] = __ffma2_rn(ez[n], sv[n], az[k]);
                        }
                    }
                    continue;
                }
                const float4 q0 = rec[2 * (jr + u)];
                const float4 q1 = rec[2 * (jr + u) + 1];
                const float2 xj = make_float2(q0.x, q0.y), yj = make_float2(q0.z, q0.w);
                const float2 zj = make_float2(q1.x, q1.y), mj = make_float2(q1.z, q1.w);
                if (MATH & 16) {
                    // stage-major source order (all subtracts, then all r^2, ...): same arithmetic
                    float2 dx[R], dy[R], dz[R], sv[R];
#pragma unroll
                    for (int bb = 0; bb < R; ++bb) {
                        const int b = ((MATH >> 12) & 16) ? R - 1 - bb : bb;
                        dx[b] = __fadd2_rn(xj, nx[b]); dy[b] = __fadd2_rn(yj, ny[b]); dz[b] = __fadd2_rn(zj, nz[b]);
                    }
                    // PERM (MATH bits 12-21): semantically equivalent source orders.  ptxas's register assignment -- and
                    // with it the operand-bank behaviour of the three-operand FFMA2s -- depends on the source order at
                    // the 1-5 % level; the default (504: bodies walked in reverse in every stage, accumulate written
                    // as s*d + acc) is the fastest of 58 orders A/B-ed on the final source (profiles/r02_ab_perm_*.log).
                    constexpr int PERM = (MATH >> 12) & 1023;
                    auto rv = [](int bit, int b) { return (PERM & bit) ? R - 1 - b : b; };   // stage-wise reversed body order
                    if (PERM & 2) {
#pragma unroll
                        for (int b = 0; b < R; ++b) sv[b] = __ffma2_rn(dz[b], dz[b], eps2v);
#pragma unroll
                        for (int b = 0; b < R; ++b) sv[b] = __ffma2_rn(dy[b], dy[b], sv[b]);
#pragma unroll
                        for (int b = 0; b < R; ++b) sv[b] = __ffma2_rn(dx[b], dx[b], sv[b]);
                    } else {
#pragma unroll
                        for (int b = 0; b < R; ++b) sv[rv(32, b)] = __ffma2_rn(dx[rv(32, b)], dx[rv(32, b)], eps2v);
#pragma unroll
                        for (int b = 0; b < R; ++b) sv[rv(32, b)] = __ffma2_rn(dy[rv(32, b)], dy[rv(32, b)], sv[rv(32, b)]);
#pragma unroll
                        for (int b = 0; b < R; ++b) sv[rv(32, b)] = __ffma2_rn(dz[rv(32, b)], dz[rv(32, b)], sv[rv(32, b)]);
                    }
#pragma unroll
                    for (int b = 0; b < R; ++b) sv[rv(64, b)] = make_float2(rsqrt_approx(sv[rv(64, b)].x), rsqrt_approx(sv[rv(64, b)].y));
#pragma unroll
                    for (int bb = 0; bb < R; ++bb) {
                        const int b = rv(128, bb);
                        if (PERM & 1) {
                            const float2 t = __fmul2_rn(mj, sv[b]);
                            sv[b] = __fmul2_rn(__fmul2_rn(t, sv[b]), sv[b]);
                        } else {
                            const float2 inv2 = __fmul2_rn(sv[b], sv[b]);
                            const float2 mi = (PERM & 512) ? __fmul2_rn(sv[b], mj) : __fmul2_rn(mj, sv[b]);
                            sv[b] = (PERM & 512) ? __fmul2_rn(mi, inv2) : __fmul2_rn(inv2, mi);
                        }
                    }
                    if (MATH & 128) {
                        // ablation: accumulate sum s*r_j and sum s (the j operand is shared by the R bodies, so the
                        // accumulate reads two fresh 64-bit registers instead of three); a_i = sum s r_j - r_i sum s
                        // is formed once at the end.  One more FP32 instruction per pair (13 instead of 12).
#pragma unroll
                        for (int b = 0; b < R; ++b) ax[b] = __ffma2_rn(xj, sv[b], ax[b]);
#pragma unroll
                        for (int b = 0; b < R; ++b) ay[b] = __ffma2_rn(yj, sv[b], ay[b]);
#pragma unroll
                        for (int b = 0; b < R; ++b) az[b] = __ffma2_rn(zj, sv[b], az[b]);
#pragma unroll
                        for (int b = 0; b < R; ++b) as[b] = __fadd2_rn(as[b], sv[b]);
                    } else if (PERM & 4) {
#pragma unroll
                        for (int b = 0; b < R; ++b) ax[(PERM & 8) ? R - 1 - b : b] = __ffma2_rn(dx[(PERM & 8) ? R - 1 - b : b], sv[(PERM & 8) ? R - 1 - b : b], ax[(PERM & 8) ? R - 1 - b : b]);
#pragma unroll
                        for (int b = 0; b < R; ++b) ay[(PERM & 8) ? R - 1 - b : b] = __ffma2_rn(dy[(PERM & 8) ? R - 1 - b : b], sv[(PERM & 8) ? R - 1 - b : b], ay[(PERM & 8) ? R - 1 - b : b]);
#pragma unroll
                        for (int b = 0; b < R; ++b) az[(PERM & 8) ? R - 1 - b : b] = __ffma2_rn(dz[(PERM & 8) ? R - 1 - b : b], sv[(PERM & 8) ? R - 1 - b : b], az[(PERM & 8) ? R - 1 - b : b]);
                    } else {
#pragma unroll
                        for (int bb = 0; bb < R; ++bb) {
                            const int b = (PERM & 8) ? R - 1 - bb : bb;
                            if (PERM & 256) {      // multiplicands swapped: the other operand slots of the FFMA2
                                ax[b] = __ffma2_rn(sv[b], dx[b], ax[b]);
                                ay[b] = __ffma2_rn(sv[b], dy[b], ay[b]);
                                az[b] = __ffma2_rn(sv[b], dz[b], az[b]);
                            } else {
                                ax[b] = __ffma2_rn(dx[b], sv[b], ax[b]);
                                ay[b] = __ffma2_rn(dy[b], sv[b], ay[b]);
                                az[b] = __ffma2_rn(dz[b], sv[b], az[b]);
                            }
                        }
                    }
                } else
#pragma unroll
                for (int b = 0; b < R; ++b) {
                    // MATH bit set = that op group is issued as scalar instead of packed instructions
                    // (8: subtract, 4: r^2 chain, 2: Gm*inv^3, 1: accumulate).  0 = all packed is
                    // the default and the fastest of the 16 mixes: 71.1% of FP32 peak against 62.5%
                    // all-scalar, every scalarised group costs 2.6-4.6%
                    // (profiles/r01_ab_packed_vs_scalar_n262144.log).
                    const float2 dx = add2<(MATH & 8) != 0>(xj, nx[b]);
                    const float2 dy = add2<(MATH & 8) != 0>(yj, ny[b]);
                    const float2 dz = add2<(MATH & 8) != 0>(zj, nz[b]);
                    float2 r2 = fma2<(MATH & 4) != 0>(dx, dx, eps2v);
                    r2 = fma2<(MATH & 4) != 0>(dy, dy, r2);
                    r2 = fma2<(MATH & 4) != 0>(dz, dz, r2);
                    const float2 inv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                    const float2 inv2 = mul2<(MATH & 2) != 0>(inv, inv);
                    const float2 mi = mul2<(MATH & 2) != 0>(mj, inv);
                    const float2 s = mul2<(MATH & 2) != 0>(inv2, mi);
                    ax[b] = fma2<(MATH & 1) != 0>(dx, s, ax[b]);
                    ay[b] = fma2<(MATH & 1) != 0>(dy, s, ay[b]);
                    az[b] = fma2<(MATH & 1) != 0>(dz, s, az[b]);
                }
            }
        }
        if ((MATH & 32) && ((t & 3) == 3 || t == ntiles - 1)) {   // every 4th tile and the last one
#pragma unroll
            for (int b = 0; b < R; ++b) {
                double *h = hi + (3 * b) * THREADS + tid;
                const bool first = (t <= 3);
                h[0] = (first ? 0.0 : h[0]) + ((double)ax[b].x + (double)ax[b].y);
                h[THREADS] = (first ? 0.0 : h[THREADS]) + ((double)ay[b].x + (double)ay[b].y);
                h[2 * THREADS] = (first ? 0.0 : h[2 * THREADS]) + ((double)az[b].x + (double)az[b].y);
                ax[b] = ay[b] = az[b] = make_float2(0.f, 0.f);
            }
        }
        constexpr int FOLD = 4 << (2 * ((MATH >> 9) & 3));     // fold period in j tiles: 4 (default), 16, 64, 256
        if (QS && (MATH & 256) && ((t & (FOLD - 1)) == FOLD - 1 || t == ntiles - 1)) {
            // lanes are bodies: accumulator pair k holds bodies 2k (x lane) and 2k + 1 (y lane)
#pragma unroll
            for (int k = 0; k < R2; ++k) {
                float *h0 = hif + (3 * (2 * k)) * THREADS + tid, *h1 = hif + (3 * (2 * k + 1)) * THREADS + tid;
                const bool first = (t < FOLD);
                h0[0] = (first ? 0.f : h0[0]) + ax[k].x; h0[THREADS] = (first ? 0.f : h0[THREADS]) + ay[k].x;
                h0[2 * THREADS] = (first ? 0.f : h0[2 * THREADS]) + az[k].x;
                h1[0] = (first ? 0.f : h1[0]) + ax[k].y; h1[THREADS] = (first ? 0.f : h1[THREADS]) + ay[k].y;
                h1[2 * THREADS] = (first ? 0.f : h1[2 * THREADS]) + az[k].y;
                ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
            }
        }
        if (!QS && (MATH & 256) && ((t & (FOLD - 1)) == FOLD - 1 || t == ntiles - 1)) {
#pragma unroll
            for (int b = 0; b < R; ++b) {
                float *h = hif + (3 * b) * THREADS + tid;
                const bool first = (t < FOLD);
                h[0] = (first ? 0.f : h[0]) + (ax[b].x + ax[b].y);
                h[THREADS] = (first ? 0.f : h[THREADS]) + (ay[b].x + ay[b].y);
                h[2 * THREADS] = (first ? 0.f : h[2 * THREADS]) + (az[b].x + az[b].y);
                ax[b] = ay[b] = az[b] = make_float2(0.f, 0.f);
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[st]);
    }

    NBX_STAMP(2);
    // ---- fold the two j lanes
    float fx[R], fy[R], fz[R];
    if (QS) {
        static_assert(!QS || !(MATH & (32 | 128)), "q-scaled shapes: float accumulation only");
        // back to the per-body view the split combine and the epilogue use
#pragma unroll
        for (int k = R2 - 1; k >= 0; --k) {
            const float2 tx = nx[k], ty = ny[k], tz = nz[k];
            nx[2 * k] = make_float2(tx.x, tx.x); nx[2 * k + 1] = make_float2(tx.y, tx.y);
            ny[2 * k] = make_float2(ty.x, ty.x); ny[2 * k + 1] = make_float2(ty.y, ty.y);
            nz[2 * k] = make_float2(tz.x, tz.x); nz[2 * k + 1] = make_float2(tz.y, tz.y);
        }
        if (!(MATH & 256)) {
#pragma unroll
            for (int k = R2 - 1; k >= 0; --k) {
                const float2 tx = ax[k], ty = ay[k], tz = az[k];
                ax[2 * k] = make_float2(tx.x, 0.f); ax[2 * k + 1] = make_float2(tx.y, 0.f);
                ay[2 * k] = make_float2(ty.x, 0.f); ay[2 * k + 1] = make_float2(ty.y, 0.f);
                az[2 * k] = make_float2(tz.x, 0.f); az[2 * k + 1] = make_float2(tz.y, 0.f);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < R; ++b) {
        if (MATH & 32) {
            const double *h = hi + (3 * b) * THREADS + tid;
            fx[b] = ntiles > 0 ? (float)h[0] : 0.f;
            fy[b] = ntiles > 0 ? (float)h[THREADS] : 0.f;
            fz[b] = ntiles > 0 ? (float)h[2 * THREADS] : 0.f;
        } else if (MATH & 256) {
            const float *h = hif + (3 * b) * THREADS + tid;
            fx[b] = ntiles > 0 ? h[0] : 0.f;
            fy[b] = ntiles > 0 ? h[THREADS] : 0.f;
            fz[b] = ntiles > 0 ? h[2 * THREADS] : 0.f;
        } else if (MATH & 128) {
            const float ssum = as[b].x + as[b].y;
            fx[b] = fmaf(nx[b].x, ssum, ax[b].x + ax[b].y);     // nx = -x_i
            fy[b] = fmaf(ny[b].x, ssum, ay[b].x + ay[b].y);
            fz[b] = fmaf(nz[b].x, ssum, az[b].x + az[b].y);
        } else {
            fx[b] = ax[b].x + ax[b].y;
            fy[b] = ay[b].x + ay[b].y;
            fz[b] = az[b].x + az[b].y;
        }
    }

    // ---- j-split: park partials, the last CTA of this i-tile adds them in split order
    if (contributors > 1) {
        constexpr int BI = THREADS * R;
        const int tb0 = p.whole_tiles * BI;                    // first body held in `part`
        float4 *mine = p.part + (size_t)(p.split_base + split) * p.split_bodies - tb0;
#pragma unroll
        for (int k = 0; k < R2; ++k) {
            const int ip = pair_base + k * THREADS;
            if (2 * ip < p.i_count) {
                NBX_CHECK(p.part != nullptr && 2 * ip - tb0 >= 0 && 2 * ip + 1 - tb0 < p.split_bodies &&
                          p.split_base + split < p.split_total);
                mine[2 * ip] = make_float4(fx[2 * k], fy[2 * k], fz[2 * k], 0.f);
                mine[2 * ip + 1] = make_float4(fx[2 * k + 1], fy[2 * k + 1], fz[2 * k + 1], 0.f);
            }
        }
        __threadfence();
        __syncthreads();
        int *ticket = &p.tile_ticket[tile - p.whole_tiles];
        if (tid == 0) {
            const int arrived = atomicAdd(ticket, 1);
            NBX_CHECK(arrived >= 0 && arrived < contributors && tile - p.whole_tiles < p.i_tiles - p.whole_tiles);
            *s_flag = (arrived == contributors - 1);
        }
        __syncthreads();
        if (!*s_flag) { NBX_STAMP(3); return; }
        NBX_MARK(5, 1);
        __threadfence();
        if (tid == 0) *ticket = 0;
        // The adds stay in split order (bit-reproducible), but the L2 loads of all R bodies and CH
        // consecutive splits are issued together: the combine is a chain of L2 round trips on the
        // step's critical path (traced at N = 16 384: 8 dependent rounds = 4.5 us of a 117 us step),
        // this makes it ceil(S / CH) rounds (CH sized to stay inside the register budget the j loop sets).
        {
            constexpr int CH = (R <= 2) ? 4 : (R <= 4) ? 3 : 2;
            const float4 *src[R];
            bool live[R];
#pragma unroll
            for (int b = 0; b < R; ++b) {
                const int ip = pair_base + (b >> 1) * THREADS;
                live[b] = 2 * ip < p.i_count;
                const int body = live[b] ? 2 * ip + (b & 1) - tb0 : 0;
                src[b] = p.part + body;
                fx[b] = fy[b] = fz[b] = 0.f;
            }
            for (int s0 = 0; s0 < contributors; s0 += CH) {
                float4 v[CH][R];
#pragma unroll
                for (int c = 0; c < CH; ++c)
#pragma unroll
                    for (int b = 0; b < R; ++b)
                        v[c][b] = (s0 + c < contributors && live[b]) ? __ldcg(src[b] + (size_t)(s0 + c) * p.split_bodies)
                                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < CH; ++c)
#pragma unroll
                    for (int b = 0; b < R; ++b) { fx[b] += v[c][b].x; fy[b] += v[c][b].y; fz[b] += v[c][b].z; }
            }
        }
        if (p.discard_partials) {
            // The partials of this tile are dead now.  Large split tails (N = 1 M: 55 MB per step) would still be
            // written back from L2 to HBM as dirty lines nobody reads again: drop the lines instead
            // (discard.global.L2 = invalidate without write-back).  Every thread has finished reading first.
            __syncthreads();
            constexpr int LINES = BI / 8;                       // 128-byte lines per (tile, split) slab
            const int first = tile * BI - tb0;                  // first body of this tile inside `part`
            for (int l = tid; l < contributors * LINES; l += THREADS) {
                const int s = l / LINES, body = first + (l % LINES) * 8;
                if (body + 8 <= p.split_bodies) {
                    const float4 *line = p.part + (size_t)s * p.split_bodies + body;
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
                }
            }
        }
    }

    // ---- epilogue: Euler update (ver2:153-165), energy term (ver2:167-170), exchange
    double e = 0.0;
#pragma unroll
    for (int k = 0; k < R2; ++k) {
        const int ip = pair_base + k * THREADS;
        if (2 * ip >= p.i_count) continue;
        NBX_CHECK(ip >= 0 && 2 * ip + 1 < p.i_count);
        if (p.acc_out != nullptr) {
            p.acc_out[2 * ip] = make_float4(fx[2 * k], fy[2 * k], fz[2 * k], 0.f);
            p.acc_out[2 * ip + 1] = make_float4(fx[2 * k + 1], fy[2 * k + 1], fz[2 * k + 1], 0.f);
            continue;
        }
        float4 v0 = p.vel[2 * ip], v1 = p.vel[2 * ip + 1];
        v0.x = fmaf(fx[2 * k], p.dt, v0.x); v0.y = fmaf(fy[2 * k], p.dt, v0.y); v0.z = fmaf(fz[2 * k], p.dt, v0.z);
        v1.x = fmaf(fx[2 * k + 1], p.dt, v1.x); v1.y = fmaf(fy[2 * k + 1], p.dt, v1.y); v1.z = fmaf(fz[2 * k + 1], p.dt, v1.z);
        p.vel[2 * ip] = v0;
        p.vel[2 * ip + 1] = v1;
        const float4 r0 = make_float4(fmaf(v0.x, p.dt, -nx[2 * k].x), fmaf(v1.x, p.dt, -nx[2 * k + 1].x),
                                      fmaf(v0.y, p.dt, -ny[2 * k].x), fmaf(v1.y, p.dt, -ny[2 * k + 1].x));
        const float2 r1 = make_float2(fmaf(v0.z, p.dt, -nz[2 * k].x), fmaf(v1.z, p.dt, -nz[2 * k + 1].x));
        const size_t gp = (size_t)(p.i_begin >> 1) + ip;
        NBX_CHECK(2 * gp + 1 < (size_t)p.n_pad);
        p.pos_out[2 * gp] = r0;
        *reinterpret_cast<float2 *>(&p.pos_out[2 * gp + 1]) = r1;      // Gm0,Gm1 never change
        if (p.p2p) {
            if (p.mc_pos_out != nullptr) {
                // one multicast store per half record: the switch replicates it into every replica
                multimem_st_v4(&p.mc_pos_out[2 * gp], r0);
                multimem_st_v2(reinterpret_cast<float2 *>(&p.mc_pos_out[2 * gp + 1]), r1);
            } else {
                for (int g = 0; g < p.world; ++g) {
                    if (g == p.rank) continue;
                    float4 *dst = p.peer_pos_out[g];
                    NBX_CHECK(dst != nullptr && dst != p.pos_out);
                    dst[2 * gp] = r0;                                       // NVLink store
                    *reinterpret_cast<float2 *>(&dst[2 * gp + 1]) = r1;
                }
            }
        }
        e += (double)(v0.w * (v0.x * v0.x + v0.y * v0.y + v0.z * v0.z));
        e += (double)(v1.w * (v1.x * v1.x + v1.y * v1.y + v1.z * v1.z));
    }
    if (p.acc_out != nullptr) return;
    NBX_STAMP(3);

    // ---- kinetic energy: warp shuffle -> smem -> per-tile partial -> last CTA sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((tid & 31) == 0) red[tid >> 5] = e;
    if (p.p2p) __threadfence_system();   // my peer stores are visible before my ticket is
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < WARPS; ++w) s += red[w];
        NBX_CHECK(tile >= 0 && tile < p.i_tiles);
        p.ke_part[tile] = s;
        __threadfence();
        const int arrived = atomicAdd(p.ke_ticket, 1);
        NBX_CHECK(arrived >= 0 && arrived < p.i_tiles);
        *s_flag = (arrived == p.i_tiles - 1);
    }
    __syncthreads();
    if (!*s_flag) return;
    __threadfence();
    double s = 0.0;
    for (int i = tid; i < p.i_tiles; i += THREADS) s += __ldcg(&p.ke_part[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < WARPS; ++w) tot += red[w];
        const int step = *p.dev_step;
        NBX_CHECK(step >= 0 && step < p.ke_cap);
        p.ke_out[step] = 0.5 * tot;
        *p.dev_step = step + 1;
        *p.ke_ticket = 0;
        const int epoch = *p.dev_epoch + 1;
        *p.dev_epoch = epoch;
        if (p.p2p) {
            __threadfence_system();
            for (int g = 0; g < p.world; ++g)
                if (g != p.rank) st_release_sys(&p.peer_flags[g][p.rank], epoch);
        }
    }
}

// ------------------------------------------------------------------------------
//  q-scaled shapes: derive the j-records of the 11-instruction pair from the state records.
//  One thread per body pair of the window [rec_org, rec_org + rec_len) (modulo n_pad / 2):
//     {x0,x1,y0,y1 | z0,z1,Gm0,Gm1}  ->  {q0 x0,q0,q0 y0,q0^2 eps | q0 z0,0,q1 x1,q1 | q1 y1,q1^2 eps,q1 z1,0},   q = Gm^(-1/2)
//  Runs on the step's stream right before step_kernel (which still takes its i-bodies, the Euler update
//  and the exchange from the state records): 40 bytes per body of HBM traffic, 20 us at N = 4 M against a
//  step of 0.8 - 6.5 s.  Zero-mass bodies (padding) get Gm = 1e-30: q = 1e15, their pair term underflows to 0.
//  In P2P mode this is the first kernel of a step to read what the peers stored, so it carries the same
//  bounded wait for their flags as step_kernel (which then finds them set).
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qscale_kernel(const __grid_constant__ StepParams p, int rec_org, int rec_len)
{
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        int bad = p.p2p ? ld_volatile(p.dev_err) : 0;
        if (p.p2p && !bad) {
            const int epoch = *p.dev_epoch;
            for (int g = 0; g < p.world && !bad; ++g) {
                if (g == p.rank || ld_acquire_sys(&p.my_flags[g]) >= epoch) continue;
                const unsigned long long t0 = globaltimer_ns();
                while (ld_acquire_sys(&p.my_flags[g]) < epoch) {
                    if ((bad = ld_volatile(p.dev_err)) != 0) break;
                    if (globaltimer_ns() - t0 > p.peer_wait_ns) {
                        dev_fail(p.dev_err, kDevErrPeerTimeout | (g << 8));
                        bad = 1;
                        break;
                    }
                    __nanosleep(200);
                }
            }
        }
        s_bad = bad;
    }
    __syncthreads();
    if (s_bad) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rec_len) return;
    const int nrec = p.n_pad >> 1;
    int rec = rec_org + r;
    if (rec >= nrec) rec -= nrec;
    NBX_CHECK(p.qrec != nullptr && rec >= 0 && rec < nrec);
    const float4 q0 = __ldcg(&p.pos_in[2 * rec]);
    const float4 q1 = __ldcg(&p.pos_in[2 * rec + 1]);
    const float s0 = 1.0f / sqrtf(fmaxf(q1.z, 1e-30f));
    const float s1 = 1.0f / sqrtf(fmaxf(q1.w, 1e-30f));
    p.qrec[3 * rec] = make_float4(s0 * q0.x, s0, s0 * q0.z, (s0 * s0) * p.eps2);
    p.qrec[3 * rec + 1] = make_float4(s0 * q1.x, 0.f, s1 * q0.y, s1);
    p.qrec[3 * rec + 2] = make_float4(s1 * q0.w, (s1 * s1) * p.eps2, s1 * q1.y, 0.f);
}

// ------------------------------------------------------------------------------
//  Layout kernels: host SoA staging <-> pair-packed records.
// ------------------------------------------------------------------------------
// stage = 7 arrays of `stride` floats (px py pz vx vy vz mass) holding bodies [first, first + stride)
// of the caller's arrays.  Packs records [rec_begin, rec_begin + rec_count); bodies >= n are
// zero-mass padding.  pos_b may be null (sharded upload: the second replica is copied afterwards).
__global__ void pack_kernel(const float *__restrict__ stage, int stride, int first, int n, int rec_begin,
                            int rec_count, int i_begin, int i_count, float G, float4 *__restrict__ pos_a,
                            float4 *__restrict__ pos_b, float4 *__restrict__ vel)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rec_count) return;
    const int rec = rec_begin + r;                           // record = body pair
    float x[2], y[2], z[2], m[2], vx[2], vy[2], vz[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = 2 * rec + h;
        const int k = i - first;
        const bool live = i < n && k >= 0 && k < stride;
        x[h] = live ? stage[k] : 0.f;
        y[h] = live ? stage[(size_t)stride + k] : 0.f;
        z[h] = live ? stage[2 * (size_t)stride + k] : 0.f;
        vx[h] = live ? stage[3 * (size_t)stride + k] : 0.f;
        vy[h] = live ? stage[4 * (size_t)stride + k] : 0.f;
        vz[h] = live ? stage[5 * (size_t)stride + k] : 0.f;
        m[h] = live ? stage[6 * (size_t)stride + k] : 0.f;
    }
    const float4 q0 = make_float4(x[0], x[1], y[0], y[1]);
    const float4 q1 = make_float4(z[0], z[1], G * m[0], G * m[1]);
    pos_a[2 * rec] = q0; pos_a[2 * rec + 1] = q1;
    if (pos_b != nullptr) { pos_b[2 * rec] = q0; pos_b[2 * rec + 1] = q1; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int li = 2 * rec + h - i_begin;
        if (li >= 0 && li < i_count) vel[li] = make_float4(vx[h], vy[h], vz[h], m[h]);
    }
}

// stage = 6 arrays of `stride` floats: px py pz vx vy vz of bodies [first, first + count), count <=
// stride; velocities are written only for bodies of this shard ([i_begin, i_begin + i_count)).
__global__ void unpack_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel, int stride, int first,
                              int count, int i_begin, int i_count, float *__restrict__ stage)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = first + k;
    const float *rec = reinterpret_cast<const float *>(pos + 2 * (size_t)(i >> 1));
    const int h = i & 1;
    stage[k] = rec[h];
    stage[(size_t)stride + k] = rec[2 + h];
    stage[2 * (size_t)stride + k] = rec[4 + h];
    const int li = i - i_begin;
    if (li >= 0 && li < i_count) {
        const float4 v = vel[li];
        stage[3 * (size_t)stride + k] = v.x;
        stage[4 * (size_t)stride + k] = v.y;
        stage[5 * (size_t)stride + k] = v.z;
    }
}

// float4 accelerations of a shard -> 3 SoA arrays of `count` floats (nbx_accelerations).
__global__ void acc_soa_kernel(const float4 *__restrict__ acc, int count, float *__restrict__ stage)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float4 a = acc[k];
    stage[k] = a.x;
    stage[(size_t)count + k] = a.y;
    stage[2 * (size_t)count + k] = a.z;
}

}  // namespace nbx

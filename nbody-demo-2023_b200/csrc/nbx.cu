// nbx.cu -- the C ABI of libnbx.so (include/nbx.h): context, device-resident state,
// step launcher, multi-GPU exchange.  Host side of the kernels in nbx_kernels.cuh.
//
// Reference boundary this implements: the body of a backend's GSimulation::start()
// (ver5_all/programming_models/cuda/Compute.cu:69-232) minus printing -- allocation
// (:76-108), upload (:115-123), the per-step kernel + host update (:150-194).  Unlike
// that backend nothing crosses PCIe per step and every CUDA/NCCL call is checked.
#include "../../include/nbx.h"
#include "ic.hpp"
#include "nbx_kernels.cuh"
#include "nbx_multicast.hpp"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using nbx::StepParams;

// ------------------------------------------------------------------------------
//  errors
// ------------------------------------------------------------------------------
static thread_local std::string g_err = "";

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(NBX_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------
//  NCCL, loaded on first use so single-GPU users need no libnccl at all
// ------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.handle) return NBX_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(NBX_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                            \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                     \
    if (!g_nccl.field) return fail(NBX_ERR_NCCL, "libnccl lacks symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return NBX_OK;
}

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(NBX_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// ------------------------------------------------------------------------------
//  kernel-shape table
// ------------------------------------------------------------------------------
struct Variant {
    const char *name;
    int r2, threads, tj, stages, unroll, minb;
    int smem;
    const void *fn;
    bool qscale;   // j-records come from the q-scaled array (qscale_kernel runs before every step launch)
};

template <int R2, int THREADS, int TJ, int STAGES, int UNROLL, int MINB, int MATH = 0>
static Variant make_variant(const char *name)
{
    return Variant{name, R2, THREADS, TJ, STAGES, UNROLL, MINB, nbx::step_smem_bytes<THREADS, TJ, STAGES, R2, MATH>(),
                   (const void *)nbx::step_kernel<R2, THREADS, TJ, STAGES, UNROLL, MINB, MATH>, (MATH & nbx::kMathQScale) != 0};
}

static const std::vector<Variant> &variants()
{
    static const std::vector<Variant> v = {
        // shapes: r<i-bodies per thread>_t<threads>_u<j-records per trip>; "_stage" = stage-major source
        // order of the inner loop (+1% over body-major in same-box A/B, profiles/r01_ab_*.log);
        // "_f2" = two-level float accumulation (lane sums folded into a second float in shared memory every
        // 64 j tiles): removes the systematic low bias of long float sums for 0.4% of throughput
        // (profiles/r02d_*: N = 1 M kinetic energy 2e-7 from the fp64 truth instead of 1.8e-4)
        // (504 << 12): the four bodies are walked in reverse order in every stage of the loop body and the accumulate
        // takes its multiplicands as (s, d) -- the fastest of 58 semantically equivalent source orders on the final
        // source (ptxas register assignment / operand slots; profiles/r02_ab_perm_*)
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (504 << 12)>("r4_t256_u4_stage_f2"),   // [0] default for shards >= 8192 bodies of N < 65 536 (kLargeVariant)
        make_variant<1, 128, 256, 4, 4, 6>("r2_t128_u4"),                            // [1] default for small shards (kSmallVariant)
        make_variant<2, 256, 256, 4, 4, 2, 48>("r4_t256_u4_stage_acc64"),            // [2] accuracy option (kAccurateVariant)
        make_variant<2, 256, 256, 4, 4, 2, 16>("r4_t256_u4_stage"),                  // [3] one float accumulator per lane, like the reference's loops
        // CTA sizes the ver5_all-style CLI can ask for (argv[5] = thread_dim0, cuda/Compute.cu:137-145)
        make_variant<2, 128, 256, 4, 2, 4>("r4_t128_u2"),
        make_variant<2, 512, 512, 4, 2, 1>("r4_t512_u2"),
        make_variant<2, 64, 128, 4, 2, 8>("r4_t64_u2"),
        // "_qi": the q-scaled pair with lanes packed over i-bodies -- 11 packed FP32 instructions instead of 12, and the
        // subtract reads a register pair and two scalars of opposite parity (QS block in nbx_kernels.cuh); source order 1
        // (units numbered pair-major) is the fastest of the 7 orders A/B-ed (profiles/r02_ab_qscaled_ipacked_*.log)
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (1 << 12) | nbx::kMathQScale>("r4_t256_u4_stage_f2_qi"),   // [7] default from 65 536 bodies on (kQScaleVariant)
#ifdef NBX_ABLATION
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | nbx::kMathQScale>("r4_t256_u4_stage_f2_qi_p0"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (256 << 12) | nbx::kMathQScale>("r4_t256_u4_stage_f2_qi_p256"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (257 << 12) | nbx::kMathQScale>("r4_t256_u4_stage_f2_qi_p257"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (488 << 12) | nbx::kMathQScale>("r4_t256_u4_stage_f2_qi_p488"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (489 << 12) | nbx::kMathQScale>("r4_t256_u4_stage_f2_qi_p489"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024 | nbx::kMathQScale>("r4_t256_u2_stage_f2_qi_p0"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024 | (256 << 12) | nbx::kMathQScale>("r4_t256_u2_stage_f2_qi_p256"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | nbx::kMathQScale>("r4_t256_u4_stage_qi_p0"),                      // one float accumulator per body
        make_variant<4, 128, 256, 4, 2, 4, 16 | 256 | 1024 | nbx::kMathQScale>("r8_t128_u2_stage_f2_qi_p0"),      // 8 bodies per thread: same speed
        make_variant<4, 128, 256, 4, 2, 4, 16 | 256 | 1024 | (256 << 12) | nbx::kMathQScale>("r8_t128_u2_stage_f2_qi_p256"),
        make_variant<4, 256, 256, 4, 2, 1, 16 | 256 | 1024 | nbx::kMathQScale>("r8_t256_u2_stage_f2_qi_p0"),
        // Shapes kept only for the tuning tools (tools/sweep.py, tools/ab.py): `make ablation`
        // builds libnbx_ablation.so with them; the product library does not carry them.
        make_variant<2, 256, 256, 4, 2, 2>("r4_t256_u2"),
        make_variant<2, 256, 256, 4, 2, 2, 16>("r4_t256_u2_stage"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024>("r4_t256_u2_stage_f2"),
        make_variant<2, 256, 256, 4, 4, 2, 256 | 1024>("r4_t256_u4_f2"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256>("r4_t256_u4_stage_f2p4"),           // fold every 4 tiles: -3%
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 512>("r4_t256_u4_stage_f2p16"),    // every 16: -0.5%
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1536>("r4_t256_u4_stage_f2p256"),
        // source-order permutations (MATH bits 12-19; see PERM in nbx_kernels.cuh): same arithmetic, 2569 .. 2699 G pairs/s at N = 1 M
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024>("r4_t256_u4_stage_f2_perm0"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (1 << 12)>("r4_t256_u4_stage_f2_perm1"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (3 << 12)>("r4_t256_u4_stage_f2_perm3"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (8 << 12)>("r4_t256_u4_stage_f2_perm8"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (16 << 12)>("r4_t256_u4_stage_f2_perm16"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (184 << 12)>("r4_t256_u4_stage_f2_perm184"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024 | (8 << 12)>("r4_t256_u2_stage_f2_perm8"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024 | (248 << 12)>("r4_t256_u2_stage_f2_perm248"),
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (248 << 12)>("r4_t256_u4_stage_f2_perm248"),    // 504 without the swapped accumulate multiplicands
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (760 << 12)>("r4_t256_u4_stage_f2_perm760"),    // 248 + swapped multiplies
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (1016 << 12)>("r4_t256_u4_stage_f2_perm1016"),  // both
        make_variant<2, 256, 256, 4, 4, 2, 16 | 256 | 1024 | (256 << 12)>("r4_t256_u4_stage_f2_perm256"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024 | (504 << 12)>("r4_t256_u2_stage_f2_perm504"),
        make_variant<2, 256, 256, 4, 2, 2, 16 | 256 | 1024 | (264 << 12)>("r4_t256_u2_stage_f2_perm264"),
        // more, smaller CTAs per SM (independent instruction streams): 4 x 128 threads, 8 x 64 threads
        make_variant<2, 128, 256, 4, 4, 4, 16 | 256 | 1024 | (504 << 12)>("r4_t128_u4_stage_f2"),
        make_variant<2, 128, 256, 4, 4, 4, 16 | 256 | 1024 | (248 << 12)>("r4_t128_u4_stage_f2_perm248"),
        make_variant<2, 128, 256, 4, 4, 4, 16 | 256 | 1024>("r4_t128_u4_stage_f2_perm0"),
        make_variant<2, 64, 256, 4, 4, 8, 16 | 256 | 1024 | (504 << 12)>("r4_t64_u4_stage_f2"),
        make_variant<2, 256, 256, 4, 4, 2>("r4_t256_u4"),
        make_variant<2, 256, 256, 4, 1, 2>("r4_t256_u1"),
        make_variant<3, 256, 256, 4, 2, 2>("r6_t256_u2"),
        make_variant<4, 256, 256, 4, 1, 1>("r8_t256_u1"),
        make_variant<1, 256, 256, 4, 4, 3>("r2_t256_u4"),
        make_variant<2, 256, 256, 4, 2, 2, 1>("r4_t256_u2_sacc"),                        // scalar accumulate
        make_variant<2, 256, 256, 4, 2, 2, 15>("r4_t256_u2_scalar"),                     // no packed FP32 at all
        make_variant<2, 256, 256, 4, 2, 3, 16>("r4_t256_u2_stage_occ3"),                 // 3 CTAs/SM x 78 registers: -3%
        make_variant<2, 256, 256, 4, 4, 2, 16 | 128>("r4_t256_u4_stage_xjacc"),          // accumulate sum s*r_j, sum s: 13 ops/pair, -7%, 4e-3 force error at 1 M
        make_variant<4, 256, 256, 4, 2, 1, 16>("r8_t256_u2_stage"),                      // 8 bodies per thread, 8 warps/SM: no better at any N
        make_variant<3, 256, 256, 4, 2, 1, 16>("r6_t256_u2_stage"),
        make_variant<2, 256, 256, 8, 4, 2, 16>("r4_t256_u4_stage_s8"),                   // deeper TMA ring: no change
        make_variant<2, 256, 512, 4, 4, 2, 16>("r4_t256_u4_stage_tj512"),                // larger TMA tiles: no change
#endif
    };
    return v;
}

constexpr int kLargeVariant = 0, kSmallVariant = 1, kAccurateVariant = 2, kQScaleVariant = 7;
// From this many bodies on, the 11-instruction q-scaled pair is the default: +2.9 % at N = 1 M, +1.5 % at 65 536, slower at
// 16 384 (its record rewrite is one more launch per step); its one extra rounding per body is invisible in a sum over
// >= 65 536 pairs and 2e-7 instead of 2e-8 of the force at N = 2000 (tests/qscale_emulation.py).
constexpr int kQScaleMinBodies = 65536;
constexpr int kSmallShardBodies = 8192;   // below this the 256-body CTAs of kSmallVariant fill the SMs better

// ------------------------------------------------------------------------------
//  context
// ------------------------------------------------------------------------------
struct nbx_ctx {
    int n = 0, n_pad = 0, device = 0, rank = 0, world = 1, i_begin = 0, i_count = 0;
    float dt = 0.1f, G = 6.67259e-11f, eps2 = 1e-3f;
    int sm_count = 0, sm_clock_khz = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    float4 *pos[2] = {nullptr, nullptr};
    int cur = 0;
    float4 *vel = nullptr;
    float4 *qrec = nullptr;    // q-scaled shapes only: 24 B per body, rewritten before every step launch
    float4 *part = nullptr;
    float4 *acc = nullptr;
    int *tile_ticket = nullptr;
    double *ke_part = nullptr;
    int *counters = nullptr;   // [0]=ke_ticket [1]=dev_step [2]=dev_epoch [3]=dev_err [4]=barrier word
    int *flags = nullptr;      // [kMaxWorld] peers publish their epoch here
    double *ke_dev = nullptr;
    int ke_cap = 0;
    float *stage = nullptr;
    size_t stage_floats = 0;
    bool uploaded = false;

    // configuration
    int variant = 0, opt_variant = -1, opt_splits = 0, opt_graph = -1, opt_pdl = -1, opt_accurate = 0;
    int exchange = NBX_EXCHANGE_P2P;   // the one default (include/nbx.h); callers fall back to NCCL together
    int peer_timeout_ms = 30000;
    int device_error = 0;              // last value read from counters[3]; non-zero = poisoned
    bool debug_fault = false;          // debug build only
    int smem_pad = 0;                  // tuning knob ("smem_pad_kb")
    int auto_pad = 0;                  // single-wave grids: pad so that one CTA owns an SM (see resolve)
    bool single_wave = false;
    unsigned long long *trace = nullptr;   // trace build only
    int trace_steps = 0, trace_ctas = 0;
    bool resolved = false;
    int i_tiles = 0, whole_tiles = 0, j_splits = 1, split_bodies = 0, ctas_per_sm = 0, use_graph = 0;

    // graphs (2-step and 16-step replay units; both start and end on pos[0])
    cudaGraphExec_t graph2 = nullptr, graph16 = nullptr;

    // multi-GPU
    cudaStream_t comm_stream = nullptr;          // NCCL-overlap mode: all-gathers run here
    cudaEvent_t ev_step = nullptr, ev_gather = nullptr;
    bool gather_pending = false;
    int s_local = 1, s_remote = 1;               // NCCL-overlap mode: j-splits of the two launches
    ncclComm_t comm = nullptr;
    bool p2p_ready = false;
    float4 *peer_pos[2][nbx::kMaxWorld] = {};
    int *peer_flags[nbx::kMaxWorld] = {};
    std::vector<void *> ipc_opened;
    nbx_mc::Buffer mcbuf[2];                     // NVSwitch multicast mappings of pos[0], pos[1] (when mc_active)
    bool mc_active = false;
    std::string mc_note = "not attempted";
    float4 *retired_pos[2] = {nullptr, nullptr};   // cudaMalloc replicas superseded by multicast memory while peers hold IPC mappings of them
    int opt_multicast = -1;                      // -1 auto (try it inside one process), 0 off, 1 required

    long long kernel_launches = 0, aux_launches = 0;
    double last_run_seconds = 0.0, kernel_seconds_total = 0.0;
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// j-split count for `tiles` equal i-tiles sweeping `j_len` j-bodies on `sms` SMs (cost model below).
static int pick_splits(int tiles, int j_len, int sms)
{
    const int smax = std::max(1, std::min(64, j_len / 128));
    double best = 1e300;
    int splits = 1;
    for (int s = 1; s <= smax; ++s) {
        const double rounds = std::ceil((double)tiles * s / sms);
        const double cost = rounds * (48.0 + (double)j_len / s) + 2.0 * s;
        if (cost < best * 0.998) { best = cost; splits = s; }
    }
    return splits;
}

// The launch plan of one shard: kernel shape, i-tiles, which tiles run unsplit, j-split counts.
// Pure host arithmetic (no CUDA calls) so that nbx_plan() can expose it to CPU-only tests.
struct Plan {
    int variant, i_tiles, whole_tiles, j_splits, split_bodies, s_local, s_remote, use_graph;
};

static Plan make_plan(int n_pad, int i_count, int world, int sm_count, int exchange, int opt_variant,
                      int opt_accurate, int opt_splits, int opt_graph)
{
    Plan p{};
    const bool overlap = world > 1 && exchange == NBX_EXCHANGE_NCCL_OVERLAP;
    // The q-scaled shape is not picked by itself in the NCCL-overlap mode (a step is two launches there, each with its own
    // record rewrite: that combination has only been run with the shape forced through the "variant" option).
    p.variant = opt_variant >= 0 ? opt_variant
                : opt_accurate   ? kAccurateVariant
                : i_count < kSmallShardBodies ? kSmallVariant
                : (n_pad >= kQScaleMinBodies && !overlap) ? kQScaleVariant
                                                          : kLargeVariant;
    const Variant &v = variants()[p.variant];
    const int bi = v.threads * v.r2 * 2;
    p.i_tiles = (i_count + bi - 1) / bi;

    // j-split.  CTAs are dealt round-robin to the SMs, so a group of t equal tiles cut S ways costs
    //   ceil(t*S / SMs) rounds x (fixed cost per CTA + n_pad/S j-bodies) + S partials to combine.
    // Fitted to same-box A/B runs (tools/ab.py; profiles/r01_ab_*.log): the fixed cost (prologue,
    // first TMA round trip, epilogue) is worth ~48 j-bodies of a 1024-body CTA, a split ~2.
    // Tiles that fill whole rounds of the SM count run unsplit (no partial-force traffic); only
    // the tail -- everything, when there are fewer tiles than SMs -- is split.  A forced
    // "j_splits" applies to every tile (that is what makes results shard-count independent).
    int splits = opt_splits;
    p.whole_tiles = 0;
    p.s_local = p.s_remote = 1;
    if (overlap) {
        // a step is two launches: the own j-shard (no remote data needed), then the other shards
        // once their all-gather has landed; every tile is split, partials of both launches meet
        // in the last-arriver combine of the second one.
        p.s_local = splits > 0 ? splits : pick_splits(p.i_tiles, i_count, sm_count);
        p.s_remote = splits > 0 ? splits : pick_splits(p.i_tiles, n_pad - i_count, sm_count);
        splits = p.s_local + p.s_remote;
    } else if (splits <= 0) {
        // Unsplit tiles only when they fill >= 3 whole rounds of the SMs.  The first wave is not
        // dealt evenly (traced at 152 tiles: 6 of 148 SMs received two whole-tile CTAs, 6 none) and
        // co-resident CTAs do not share an SM fairly (the older one runs, the newer one starves),
        // so with 1-2 rounds of long CTAs the step took 2x; from 3 rounds on the short tail CTAs
        // even it out (profiles/r01_hybrid_probe.log).
        const int rounds = p.i_tiles / sm_count;
        p.whole_tiles = rounds >= 3 ? rounds * sm_count : 0;
        const int tail = p.i_tiles - p.whole_tiles;
        splits = tail > 0 ? pick_splits(tail, n_pad, sm_count) : 1;
    }
    splits = std::max(1, std::min(splits, std::max(1, n_pad / 8)));
    p.j_splits = splits;
    if (splits == 1) p.whole_tiles = p.i_tiles;
    p.split_bodies = std::max(0, i_count - p.whole_tiles * bi);

    if (opt_graph >= 0)
        p.use_graph = opt_graph;
    else   // launch latency matters below ~1 ms per step
        p.use_graph = ((double)n_pad * (double)i_count < 2.5e9) ? 1 : 0;
    if (world > 1 && exchange != NBX_EXCHANGE_P2P) p.use_graph = 0;
    return p;
}

static int resolve(nbx_ctx *c)
{
    if (c->resolved) return NBX_OK;
    const Plan pl = make_plan(c->n_pad, c->i_count, c->world, c->sm_count, c->exchange, c->opt_variant,
                              c->opt_accurate, c->opt_splits, c->opt_graph);
    c->variant = pl.variant;
    c->i_tiles = pl.i_tiles; c->whole_tiles = pl.whole_tiles; c->j_splits = pl.j_splits;
    c->split_bodies = pl.split_bodies; c->s_local = pl.s_local; c->s_remote = pl.s_remote;
    c->use_graph = pl.use_graph;
    const int splits = pl.j_splits;
    const Variant &v = variants()[c->variant];
    CU(cudaSetDevice(c->device));
    // A grid that fits the SMs once (small N: 144 CTAs at N = 16 384) runs one CTA per SM.  Consecutive
    // steps then overlap best with programmatic dependent launch -- but only if a dependent CTA cannot
    // squeeze in NEXT to a running one (traced: it did, two CTAs shared 72 SMs and the step took 0.20 ms
    // instead of 0.114).  Padding the dynamic shared memory past half an SM's worth makes residency 1;
    // the dependents then start exactly as SMs drain (profiles/r02_trace_c1.log: 117.0 -> 115.8 us).
    const int ctas_total = c->whole_tiles + (c->i_tiles - c->whole_tiles) * c->j_splits;
    c->single_wave = ctas_total <= c->sm_count && c->opt_pdl != 0 && c->smem_pad == 0 &&
                     !(c->world > 1 && c->exchange == NBX_EXCHANGE_NCCL_OVERLAP);
    c->auto_pad = c->single_wave ? std::max(0, 117 * 1024 - v.smem) : 0;
    CU(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem + c->smem_pad + c->auto_pad));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, v.threads, v.smem + c->smem_pad + c->auto_pad));
    if (occ < 1) return fail(NBX_ERR_CUDA, "kernel variant %s cannot be resident", v.name);
    c->ctas_per_sm = occ;

    if (c->part) { CU(cudaFree(c->part)); c->part = nullptr; }
    if (c->qrec) { CU(cudaFree(c->qrec)); c->qrec = nullptr; }
    if (v.qscale) CU(cudaMalloc(&c->qrec, (size_t)c->n_pad * 24));
    if (c->tile_ticket) { CU(cudaFree(c->tile_ticket)); c->tile_ticket = nullptr; }
    if (c->ke_part) { CU(cudaFree(c->ke_part)); c->ke_part = nullptr; }
    if (splits > 1) CU(cudaMalloc(&c->part, (size_t)splits * c->split_bodies * sizeof(float4)));
    CU(cudaMalloc(&c->tile_ticket, (size_t)(c->i_tiles + 1) * sizeof(int)));
    CU(cudaMemset(c->tile_ticket, 0, (size_t)(c->i_tiles + 1) * sizeof(int)));
    CU(cudaMalloc(&c->ke_part, (size_t)c->i_tiles * sizeof(double)));
    if (c->graph2) { cudaGraphExecDestroy(c->graph2); c->graph2 = nullptr; }
    if (c->graph16) { cudaGraphExecDestroy(c->graph16); c->graph16 = nullptr; }
    if (c->trace) { CU(cudaFree(c->trace)); c->trace = nullptr; }
    c->trace_ctas = c->whole_tiles + (c->i_tiles - c->whole_tiles) * c->j_splits;
    if (c->trace_steps > 0) {
        const size_t bytes = (size_t)c->trace_steps * c->trace_ctas * nbx::kTraceWords * sizeof(unsigned long long);
        CU(cudaMalloc(&c->trace, bytes));
        CU(cudaMemset(c->trace, 0, bytes));
    }
    c->resolved = true;
    return NBX_OK;
}

enum { PHASE_WHOLE = 0, PHASE_LOCAL = 1, PHASE_REMOTE = 2 };

static void fill_params(const nbx_ctx *c, StepParams &p, int in_buf, float4 *acc_out, int phase = PHASE_WHOLE)
{
    std::memset(&p, 0, sizeof p);
    p.pos_in = c->pos[in_buf];
    p.pos_out = c->pos[in_buf ^ 1];
    p.qrec = c->qrec;
    p.vel = c->vel;
    p.part = c->part;
    p.tile_ticket = c->tile_ticket;
    p.ke_part = c->ke_part;
    p.ke_ticket = c->counters + 0;
    p.dev_step = c->counters + 1;
    p.dev_epoch = c->counters + 2;
    p.dev_err = c->counters + 3;
    p.ke_out = c->ke_dev;
    p.ke_cap = c->debug_fault ? 0 : c->ke_cap;
    p.peer_wait_ns = (unsigned long long)c->peer_timeout_ms * 1000000ull;
    p.trace = c->trace;
    p.trace_steps = c->trace_steps;
    p.acc_out = acc_out;
    p.n_pad = c->n_pad;
    p.i_begin = c->i_begin;
    p.i_count = c->i_count;
    p.i_tiles = c->i_tiles;
    p.whole_tiles = c->whole_tiles;
    p.j_splits = c->j_splits;
    p.split_bodies = c->split_bodies;
    // partial scratch beyond ~16 MB does not stay in L2 between its producer and its consumer anyway
    p.discard_partials = ((size_t)c->j_splits * c->split_bodies * sizeof(float4) > ((size_t)16 << 20)) ? 1 : 0;
    p.j_org = 0;
    p.j_len = c->n_pad;
    p.split_base = 0;
    p.split_total = c->j_splits;
    const bool overlap = c->world > 1 && c->exchange == NBX_EXCHANGE_NCCL_OVERLAP;
    if (overlap && phase == PHASE_WHOLE) {           // nbx_accelerations: one launch, s_remote slots
        p.j_splits = p.split_total = c->s_remote;
    } else if (phase == PHASE_LOCAL) {
        p.j_org = c->i_begin; p.j_len = c->i_count;
        p.j_splits = c->s_local; p.split_base = 0;
    } else if (phase == PHASE_REMOTE) {
        p.j_org = (c->i_begin + c->i_count) % c->n_pad; p.j_len = c->n_pad - c->i_count;
        p.j_splits = c->s_remote; p.split_base = c->s_local;
    }
    p.dt = c->dt;
    p.eps2 = c->eps2;
    p.world = c->world;
    p.rank = c->rank;
    p.p2p = (acc_out == nullptr && c->world > 1 && c->exchange == NBX_EXCHANGE_P2P) ? 1 : 0;
    if (p.p2p) {
        for (int g = 0; g < c->world; ++g) {
            p.peer_pos_out[g] = c->peer_pos[in_buf ^ 1][g];
            p.peer_flags[g] = c->peer_flags[g];
        }
        p.my_flags = c->flags;
        p.mc_pos_out = c->mc_active ? reinterpret_cast<float4 *>(c->mcbuf[in_buf ^ 1].mcva) : nullptr;
    }
}

static int launch_step(nbx_ctx *c, int in_buf, float4 *acc_out = nullptr, int phase = PHASE_WHOLE)
{
    const Variant &v = variants()[c->variant];
    StepParams p;
    fill_params(c, p, in_buf, acc_out, phase);
    void *args[] = {&p};
    if (v.qscale) {
        // the launch's j window, in records; a plain launch: it starts when the previous step has completed
        const int rec_org = p.j_org >> 1, rec_len = p.j_len >> 1;
        nbx::qscale_kernel<<<(rec_len + 255) / 256, 256, 0, c->stream>>>(p, rec_org, rec_len);
        CU(cudaGetLastError());
        c->aux_launches++;
    }
    const int ctas = c->whole_tiles + (c->i_tiles - c->whole_tiles) * p.j_splits;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(v.threads);
    cfg.dynamicSmemBytes = v.smem + c->smem_pad + c->auto_pad;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap our prologue with
    attr[0].val.programmaticStreamSerializationAllowed = 1;            // the previous step's tail
    cfg.attrs = attr;
    // Auto: grids of many waves, and single-wave grids made one-CTA-per-SM by resolve().  In between
    // (a wave or two, two CTAs per SM) early-scheduled dependents take the free slots unevenly.
    const bool pdl = c->opt_pdl >= 0 ? c->opt_pdl != 0 : (ctas >= 4 * c->sm_count || c->single_wave);
    cfg.numAttrs = pdl ? 1 : 0;
    CU(cudaLaunchKernelExC(&cfg, v.fn, args));
    c->kernel_launches++;
    return NBX_OK;
}

static int build_graph(nbx_ctx *c, int steps, cudaGraphExec_t *out)
{
    cudaGraph_t g = nullptr;
    const long long before = c->kernel_launches, before_aux = c->aux_launches;
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = NBX_OK;
    for (int s = 0; s < steps && rc == NBX_OK; ++s) rc = launch_step(c, s & 1);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->kernel_launches = before;   // capturing is not launching
    c->aux_launches = before_aux;
    if (rc != NBX_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(NBX_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(NBX_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
    return NBX_OK;
}

static int ensure_ke(nbx_ctx *c, int nsteps)
{
    if (nsteps <= c->ke_cap) return NBX_OK;
    // the graphs bake ke_dev's address in: drop them with the old buffer
    if (c->graph2) { cudaGraphExecDestroy(c->graph2); c->graph2 = nullptr; }
    if (c->graph16) { cudaGraphExecDestroy(c->graph16); c->graph16 = nullptr; }
    if (c->ke_dev) CU(cudaFree(c->ke_dev));
    c->ke_dev = nullptr;
    const int cap = std::max(nsteps, 1024);
    CU(cudaMalloc(&c->ke_dev, (size_t)cap * sizeof(double)));
    c->ke_cap = cap;
    return NBX_OK;
}

// One step's kernel launch(es) on c->stream; flips the ping-pong.
static int step_compute(nbx_ctx *c)
{
    int rc;
    if (c->world > 1 && c->exchange == NBX_EXCHANGE_NCCL_OVERLAP) {
        if ((rc = launch_step(c, c->cur, nullptr, PHASE_LOCAL))) return rc;      // needs no remote data
        if (c->gather_pending) CU(cudaStreamWaitEvent(c->stream, c->ev_gather, 0));
        if ((rc = launch_step(c, c->cur, nullptr, PHASE_REMOTE))) return rc;     // + combine + epilogue
    } else {
        if ((rc = launch_step(c, c->cur))) return rc;
    }
    c->cur ^= 1;
    return NBX_OK;
}

// The exchange that follows a step (inside ncclGroupStart/End when one thread drives several GPUs).
static int step_exchange(nbx_ctx *c)
{
    if (c->world < 2 || c->exchange == NBX_EXCHANGE_P2P) return NBX_OK;   // P2P: done by the epilogue
    float4 *buf = c->pos[c->cur];                                          // freshly written shard, in place
    cudaStream_t st = c->stream;
    if (c->exchange == NBX_EXCHANGE_NCCL_OVERLAP) {
        CU(cudaEventRecord(c->ev_step, c->stream));
        CU(cudaStreamWaitEvent(c->comm_stream, c->ev_step, 0));
        st = c->comm_stream;
    }
    NC(g_nccl.AllGather(buf + c->i_begin, buf, (size_t)c->i_count * 4, ncclFloat, c->comm, st));
    return NBX_OK;
}

// After the collective has really been enqueued (i.e. after ncclGroupEnd when one thread drives
// several GPUs -- inside a group the call above is only recorded): mark the point the next step's
// remote-shard launch has to wait for.
static int step_exchange_mark(nbx_ctx *c)
{
    if (c->world > 1 && c->exchange == NBX_EXCHANGE_NCCL_OVERLAP) {
        CU(cudaEventRecord(c->ev_gather, c->comm_stream));
        c->gather_pending = true;
    }
    return NBX_OK;
}

// Kinetic energies of all shards; leaves c->stream ordered after every outstanding exchange.
static int finish_run(nbx_ctx *c, int nsteps, bool reduce)
{
    if (c->world < 2) return NBX_OK;
    cudaStream_t st = c->stream;
    const bool overlap = c->exchange == NBX_EXCHANGE_NCCL_OVERLAP;
    if (overlap) {
        CU(cudaEventRecord(c->ev_step, c->stream));
        CU(cudaStreamWaitEvent(c->comm_stream, c->ev_step, 0));
        st = c->comm_stream;
    }
    if (reduce && nsteps > 0)
        NC(g_nccl.AllReduce(c->ke_dev, c->ke_dev, (size_t)nsteps, ncclDouble, ncclSum, c->comm, st));
    if (overlap) {
        CU(cudaEventRecord(c->ev_gather, c->comm_stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_gather, 0));
    }
    return NBX_OK;
}

// Enqueue nsteps steps (and their exchanges) on c->stream; no host sync.
static int enqueue_steps(nbx_ctx *c, int nsteps)
{
    const bool nccl_x = c->world > 1 && c->exchange != NBX_EXCHANGE_P2P;
    int left = nsteps;
    if (c->use_graph && !nccl_x) {
        if (c->cur == 1 && left > 0) {   // graphs start on pos[0]
            int rc = launch_step(c, 1);
            if (rc) return rc;
            c->cur = 0;
            --left;
        }
        if (left >= 16 && !c->graph16) { int rc = build_graph(c, 16, &c->graph16); if (rc) return rc; }
        while (left >= 16) {
            CU(cudaGraphLaunch(c->graph16, c->stream));
            c->kernel_launches += 16;
            if (variants()[c->variant].qscale) c->aux_launches += 16;
            left -= 16;
        }
        if (left >= 2 && !c->graph2) { int rc = build_graph(c, 2, &c->graph2); if (rc) return rc; }
        while (left >= 2) {
            CU(cudaGraphLaunch(c->graph2, c->stream));
            c->kernel_launches += 2;
            if (variants()[c->variant].qscale) c->aux_launches += 2;
            left -= 2;
        }
    }
    while (left > 0) {
        int rc = step_compute(c);
        if (rc) return rc;
        if ((rc = step_exchange(c))) return rc;
        if ((rc = step_exchange_mark(c))) return rc;
        --left;
    }
    return NBX_OK;
}

// ------------------------------------------------------------------------------
//  C ABI
// ------------------------------------------------------------------------------
extern "C" {

int nbx_abi_version(void) { return NBX_ABI_VERSION; }
const char *nbx_last_error(void) { return g_err.c_str(); }

int nbx_device_count(int *count)
{
    if (!count) return fail(NBX_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return NBX_OK;
}

int nbx_plan(int n, int rank, int world, int sm_count, int exchange, long long variant, long long j_splits,
             nbx_info *out)
{
    if (!out) return fail(NBX_ERR_ARG, "out is NULL");
    if (n < 1 || world < 1 || world > nbx::kMaxWorld || rank < 0 || rank >= world || sm_count < 1)
        return fail(NBX_ERR_ARG, "bad n/rank/world/sm_count");
    if (variant < -1 || variant >= (long long)variants().size()) return fail(NBX_ERR_ARG, "variant out of range");
    if (exchange != NBX_EXCHANGE_NCCL && exchange != NBX_EXCHANGE_P2P && exchange != NBX_EXCHANGE_NCCL_OVERLAP)
        return fail(NBX_ERR_ARG, "unknown exchange %d", exchange);
    if (j_splits < 0 || j_splits > 4096) return fail(NBX_ERR_ARG, "j_splits out of range");
    const int n_pad = round_up(n, 8 * world), i_count = n_pad / world;
    const Plan pl = make_plan(n_pad, i_count, world, sm_count, exchange, (int)variant, 0, (int)j_splits, -1);
    const Variant &v = variants()[pl.variant];
    std::memset(out, 0, sizeof *out);
    out->abi_version = NBX_ABI_VERSION;
    out->sm_count = sm_count; out->n = n; out->n_pad = n_pad; out->rank = rank; out->world = world;
    out->i_begin = rank * i_count; out->i_count = i_count;
    out->threads = v.threads; out->bodies_per_thread = 2 * v.r2; out->tile_bodies = v.tj; out->stages = v.stages;
    out->i_tiles = pl.i_tiles; out->whole_tiles = pl.whole_tiles; out->j_splits = pl.j_splits;
    out->use_graph = pl.use_graph; out->exchange = exchange; out->variant = pl.variant;
    return NBX_OK;
}

int nbx_variant_count(void) { return (int)variants().size(); }
const char *nbx_variant_name(int idx)
{
    if (idx < 0 || idx >= (int)variants().size()) return "";
    return variants()[idx].name;
}

int nbx_create(nbx_ctx **out, int n, int device, int rank, int world, float dt, float G, float eps2)
{
    if (!out) return fail(NBX_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n < 1) return fail(NBX_ERR_ARG, "n must be >= 1 (got %d)", n);
    if (world < 1 || world > nbx::kMaxWorld || rank < 0 || rank >= world)
        return fail(NBX_ERR_ARG, "bad rank/world %d/%d (world <= %d)", rank, world, nbx::kMaxWorld);
    if (!(eps2 > 0.f)) return fail(NBX_ERR_ARG, "eps2 must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(NBX_ERR_NODEVICE, "no CUDA device visible (libnbx has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(NBX_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(NBX_ERR_NODEVICE, "device %d is sm_%d%d; libnbx is built for sm_100a only", device, prop.major, prop.minor);

    nbx_ctx *c = new nbx_ctx();
    c->n = n; c->device = device; c->rank = rank; c->world = world;
    c->dt = dt; c->G = G; c->eps2 = eps2;
    c->n_pad = round_up(n, 8 * world);          // shards and TMA chunks stay 128-byte granular
    c->i_count = c->n_pad / world;
    c->i_begin = rank * c->i_count;
    c->sm_count = prop.multiProcessorCount;
    c->sm_clock_khz = prop.clockRate;
    auto bail = [&](int rc) { nbx_destroy(c); return rc; };
#define CUB(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return bail(fail(NBX_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_))); \
    } while (0)
    CUB(cudaSetDevice(device));
    CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB(cudaEventCreate(&c->ev0));
    CUB(cudaEventCreate(&c->ev1));
    CUB(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&c->ev_step, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_gather, cudaEventDisableTiming));
    CUB(cudaMalloc(&c->pos[0], (size_t)c->n_pad * sizeof(float4)));
    CUB(cudaMalloc(&c->pos[1], (size_t)c->n_pad * sizeof(float4)));
    CUB(cudaMalloc(&c->vel, (size_t)c->i_count * sizeof(float4)));
    CUB(cudaMalloc(&c->counters, 8 * sizeof(int)));
    CUB(cudaMemset(c->counters, 0, 8 * sizeof(int)));
    CUB(cudaMalloc(&c->flags, nbx::kMaxWorld * sizeof(int)));
    CUB(cudaMemset(c->flags, 0, nbx::kMaxWorld * sizeof(int)));
#undef CUB
    *out = c;
    return NBX_OK;
}

void nbx_destroy(nbx_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->graph2) cudaGraphExecDestroy(c->graph2);
    if (c->graph16) cudaGraphExecDestroy(c->graph16);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->mc_active) {            // the replicas live in driver-API allocations bound to the multicast team
        nbx_mc::release(c->mcbuf[0]); nbx_mc::release(c->mcbuf[1]);
    } else {
        cudaFree(c->pos[0]); cudaFree(c->pos[1]);
    }
    cudaFree(c->retired_pos[0]); cudaFree(c->retired_pos[1]);
    cudaFree(c->vel); cudaFree(c->qrec); cudaFree(c->part); cudaFree(c->acc);
    cudaFree(c->tile_ticket); cudaFree(c->ke_part); cudaFree(c->counters); cudaFree(c->flags);
    cudaFree(c->ke_dev); cudaFree(c->stage); cudaFree(c->trace);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_step) cudaEventDestroy(c->ev_step);
    if (c->ev_gather) cudaEventDestroy(c->ev_gather);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int nbx_set_option(nbx_ctx *c, const char *key, long long value)
{
    if (!c || !key) return fail(NBX_ERR_ARG, "NULL argument");
    const std::string k(key);
    if (k == "j_splits") {
        if (value < 0 || value > 4096) return fail(NBX_ERR_ARG, "j_splits out of range");
        c->opt_splits = (int)value;
    } else if (k == "graph") {
        c->opt_graph = value < 0 ? -1 : (value ? 1 : 0);
    } else if (k == "accurate") {
        c->opt_accurate = value ? 1 : 0;
    } else if (k == "pdl") {
        c->opt_pdl = value < 0 ? -1 : (value ? 1 : 0);
    } else if (k == "exchange") {
        if (value != NBX_EXCHANGE_NCCL && value != NBX_EXCHANGE_P2P && value != NBX_EXCHANGE_NCCL_OVERLAP)
            return fail(NBX_ERR_ARG, "unknown exchange %lld", value);
        c->exchange = (int)value;
#ifdef NBX_TRACE
    } else if (k == "trace_steps") {   // record per-CTA timestamps of the first `value` steps of each run
        if (value < 0 || value > 4096) return fail(NBX_ERR_ARG, "trace_steps out of range");
        c->trace_steps = (int)value;
#endif
    } else if (k == "multicast") {     // NVSwitch multicast for the P2P exchange: -1 auto, 0 never, 1 fail if unavailable
        c->opt_multicast = value < 0 ? -1 : (value ? 1 : 0);
        return NBX_OK;
    } else if (k == "smem_pad_kb") {   // extra dynamic shared memory per CTA: lowers the resident CTAs per SM
        if (value < 0 || value > 200) return fail(NBX_ERR_ARG, "smem_pad_kb out of range");
        c->smem_pad = (int)value * 1024;
    } else if (k == "peer_timeout_ms") {
        if (value < 1 || value > 3600000) return fail(NBX_ERR_ARG, "peer_timeout_ms out of range");
        c->peer_timeout_ms = (int)value;
        if (c->graph2) { cudaGraphExecDestroy(c->graph2); c->graph2 = nullptr; }     // baked into graph nodes
        if (c->graph16) { cudaGraphExecDestroy(c->graph16); c->graph16 = nullptr; }
        return NBX_OK;
#ifdef NBX_DEBUG
    } else if (k == "debug_fault") {   // fault injection for tests/test_gpu_debug_build.py: the kernel is told
        c->debug_fault = value != 0;   // it has no energy slots, so its slot check fires (the store stays valid)
#endif
    } else if (k == "variant") {
        if (value < -1 || value >= (long long)variants().size()) return fail(NBX_ERR_ARG, "variant out of range");
        c->opt_variant = (int)value;
    } else {
        return fail(NBX_ERR_ARG, "unknown option '%s'", key);
    }
    c->resolved = false;
    return NBX_OK;
}

int nbx_get_info(const nbx_ctx *c, nbx_info *o)
{
    if (!c || !o) return fail(NBX_ERR_ARG, "NULL argument");
    const Variant &v = variants()[c->variant];
    std::memset(o, 0, sizeof *o);
    o->abi_version = NBX_ABI_VERSION;
    o->device = c->device; o->sm_count = c->sm_count; o->sm_clock_khz = c->sm_clock_khz;
    o->n = c->n; o->n_pad = c->n_pad; o->rank = c->rank; o->world = c->world;
    o->i_begin = c->i_begin; o->i_count = c->i_count;
    o->threads = v.threads; o->bodies_per_thread = 2 * v.r2; o->tile_bodies = v.tj; o->stages = v.stages;
    o->i_tiles = c->i_tiles; o->whole_tiles = c->whole_tiles; o->j_splits = c->j_splits; o->ctas_per_sm = c->ctas_per_sm;
    o->use_graph = c->use_graph; o->exchange = c->exchange; o->variant = c->variant;
    o->kernel_launches = c->kernel_launches; o->aux_launches = c->aux_launches;
    o->last_run_seconds = c->last_run_seconds; o->kernel_seconds_total = c->kernel_seconds_total;
    o->device_error = c->device_error; o->peer_timeout_ms = c->peer_timeout_ms;
    o->multicast = c->mc_active ? 1 : 0;
    return NBX_OK;
}

static int ensure_stage(nbx_ctx *c, size_t floats)
{
    if (floats <= c->stage_floats) return NBX_OK;
    if (c->stage) CU(cudaFree(c->stage));
    c->stage = nullptr;
    CU(cudaMalloc(&c->stage, floats * sizeof(float)));
    c->stage_floats = floats;
    return NBX_OK;
}

// H2D of bodies [lo, hi) of the caller's seven arrays into the staging buffer + pack into pos[0]
// (and pos[1] when `both`) and vel.  Leaves the work on c->stream (no sync).
static int stage_and_pack(nbx_ctx *c, const float *const src[7], int lo, int hi, int rec_begin, int rec_count, bool both)
{
    CU(cudaSetDevice(c->device));
    const size_t cnt = (size_t)std::max(hi - lo, 0);
    int rc = ensure_stage(c, 7 * std::max(cnt, (size_t)1));
    if (rc) return rc;
    if (cnt)
        for (int k = 0; k < 7; ++k)
            CU(cudaMemcpyAsync(c->stage + k * cnt, src[k] + lo, cnt * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (rec_count > 0) {
        nbx::pack_kernel<<<(rec_count + 255) / 256, 256, 0, c->stream>>>(c->stage, (int)cnt, lo, c->n, rec_begin, rec_count,
                                                                         c->i_begin, c->i_count, c->G, c->pos[0],
                                                                         both ? c->pos[1] : nullptr, c->vel);
        CU(cudaGetLastError());
        c->aux_launches++;
    }
    return NBX_OK;
}

static void mark_uploaded(nbx_ctx *c)
{
    c->cur = 0;
    c->gather_pending = false;
    c->uploaded = true;
}

int nbx_upload(nbx_ctx *c, const float *px, const float *py, const float *pz,
               const float *vx, const float *vy, const float *vz, const float *mass)
{
    if (!c || !px || !py || !pz || !vx || !vy || !vz || !mass) return fail(NBX_ERR_ARG, "NULL argument");
    const float *src[7] = {px, py, pz, vx, vy, vz, mass};
    int rc = stage_and_pack(c, src, 0, c->n, 0, c->n_pad / 2, true);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    mark_uploaded(c);
    return NBX_OK;
}

// This shard's slice of the caller's arrays and its records (the padding tail belongs to the last shard).
static void shard_span(const nbx_ctx *c, int *lo, int *hi)
{
    *lo = std::min(c->i_begin, c->n);
    *hi = std::min(c->i_begin + c->i_count, c->n);
}

int nbx_upload_sharded(nbx_ctx *c, const float *px, const float *py, const float *pz,
                       const float *vx, const float *vy, const float *vz, const float *mass)
{
    if (!c || !px || !py || !pz || !vx || !vy || !vz || !mass) return fail(NBX_ERR_ARG, "NULL argument");
    if (c->world == 1) return nbx_upload(c, px, py, pz, vx, vy, vz, mass);
    if (!c->comm) return fail(NBX_ERR_STATE, "nbx_upload_sharded needs nbx_comm_init first (it all-gathers the packed shards)");
    const float *src[7] = {px, py, pz, vx, vy, vz, mass};
    int lo, hi;
    shard_span(c, &lo, &hi);
    int rc = stage_and_pack(c, src, lo, hi, c->i_begin / 2, c->i_count / 2, false);
    if (rc) return rc;
    // packed records of every shard, GPU to GPU (16 B/body over NVLink instead of 28 B/body over PCIe)
    NC(g_nccl.AllGather(c->pos[0] + c->i_begin, c->pos[0], (size_t)c->i_count * 4, ncclFloat, c->comm, c->stream));
    CU(cudaMemcpyAsync(c->pos[1], c->pos[0], (size_t)c->n_pad * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    mark_uploaded(c);
    return NBX_OK;
}

int nbx_upload_group(nbx_ctx **ctxs, int count, const float *px, const float *py, const float *pz,
                     const float *vx, const float *vy, const float *vz, const float *mass)
{
    if (!ctxs || count < 1 || !px || !py || !pz || !vx || !vy || !vz || !mass) return fail(NBX_ERR_ARG, "NULL argument");
    for (int g = 0; g < count; ++g)
        if (!ctxs[g] || ctxs[g]->world != count || ctxs[g]->rank != g || ctxs[g]->n != ctxs[0]->n)
            return fail(NBX_ERR_ARG, "ctxs[%d] is not rank %d of %d", g, g, count);
    if (count == 1) return nbx_upload(ctxs[0], px, py, pz, vx, vy, vz, mass);
    const float *src[7] = {px, py, pz, vx, vy, vz, mass};
    int rc;
    // peer copies need peer access; where a pair cannot be enabled every context uploads in full
    bool peers_ok = true;
    for (int g = 0; g < count && peers_ok; ++g)
        for (int h = 0; h < count && peers_ok; ++h) {
            if (ctxs[g]->device == ctxs[h]->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctxs[g]->device, ctxs[h]->device) != cudaSuccess || !can) peers_ok = false;
        }
    cudaGetLastError();
    if (!peers_ok) {
        for (int g = 0; g < count; ++g)
            if ((rc = nbx_upload(ctxs[g], px, py, pz, vx, vy, vz, mass))) return rc;
        return NBX_OK;
    }
    for (int g = 0; g < count; ++g) {
        int lo, hi;
        shard_span(ctxs[g], &lo, &hi);
        if ((rc = stage_and_pack(ctxs[g], src, lo, hi, ctxs[g]->i_begin / 2, ctxs[g]->i_count / 2, false))) return rc;
    }
    for (int g = 0; g < count; ++g) {          // owner g pushes its packed shard into every other replica
        nbx_ctx *c = ctxs[g];
        CU(cudaSetDevice(c->device));
        for (int h = 0; h < count; ++h) {
            if (h == g) continue;
            CU(cudaMemcpyPeerAsync(ctxs[h]->pos[0] + c->i_begin, ctxs[h]->device, c->pos[0] + c->i_begin, c->device,
                                   (size_t)c->i_count * sizeof(float4), c->stream));
        }
    }
    for (int g = 0; g < count; ++g) {
        CU(cudaSetDevice(ctxs[g]->device));
        CU(cudaStreamSynchronize(ctxs[g]->stream));
    }
    for (int g = 0; g < count; ++g) {
        nbx_ctx *c = ctxs[g];
        CU(cudaSetDevice(c->device));
        CU(cudaMemcpyAsync(c->pos[1], c->pos[0], (size_t)c->n_pad * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    }
    for (int g = 0; g < count; ++g) {
        CU(cudaSetDevice(ctxs[g]->device));
        CU(cudaStreamSynchronize(ctxs[g]->stream));
        mark_uploaded(ctxs[g]);
    }
    return NBX_OK;
}

// positions of bodies [plo, phi) and velocities of [vlo, vhi) -> the caller's arrays (same offsets)
static int unpack_and_copy(nbx_ctx *c, int plo, int phi, float *px, float *py, float *pz, float *vx, float *vy, float *vz)
{
    if (!c->uploaded) return fail(NBX_ERR_STATE, "download before upload");
    CU(cudaSetDevice(c->device));
    const size_t cnt = (size_t)std::max(phi - plo, 0);
    if (cnt == 0) return NBX_OK;
    int rc = ensure_stage(c, 7 * cnt);
    if (rc) return rc;
    nbx::unpack_kernel<<<((int)cnt + 255) / 256, 256, 0, c->stream>>>(c->pos[c->cur], c->vel, (int)cnt, plo, (int)cnt,
                                                                      c->i_begin, c->i_count, c->stage);
    CU(cudaGetLastError());
    c->aux_launches++;
    float *dst[3] = {px, py, pz};
    for (int k = 0; k < 3; ++k)
        if (dst[k]) CU(cudaMemcpyAsync(dst[k] + plo, c->stage + k * cnt, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    const size_t lo = (size_t)std::max(std::min(c->i_begin, c->n), plo);
    const size_t hi = (size_t)std::min(std::min(c->i_begin + c->i_count, c->n), phi);
    float *vdst[3] = {vx, vy, vz};
    for (int k = 0; k < 3; ++k)
        if (vdst[k] && hi > lo)
            CU(cudaMemcpyAsync(vdst[k] + lo, c->stage + (3 + k) * cnt + (lo - plo), (hi - lo) * sizeof(float),
                               cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_download(nbx_ctx *c, float *px, float *py, float *pz, float *vx, float *vy, float *vz)
{
    if (!c) return fail(NBX_ERR_ARG, "NULL context");
    return unpack_and_copy(c, 0, c->n, px, py, pz, vx, vy, vz);
}

int nbx_download_shard(nbx_ctx *c, float *px, float *py, float *pz, float *vx, float *vy, float *vz)
{
    if (!c) return fail(NBX_ERR_ARG, "NULL context");
    int lo, hi;
    shard_span(c, &lo, &hi);
    return unpack_and_copy(c, lo, hi, px, py, pz, vx, vy, vz);
}

static int device_error_code(nbx_ctx *c, int word)
{
    c->device_error = word;
    const int kind = word & 0xff, detail = word >> 8;
    if (kind == nbx::kDevErrPeerTimeout)
        return fail(NBX_ERR_PEER, "rank %d: peer rank %d did not finish its step within %d ms (P2P exchange); context poisoned",
                    c->rank, detail, c->peer_timeout_ms);
    if (kind == nbx::kDevErrHostAbort) return fail(NBX_ERR_PEER, "rank %d: run aborted; context poisoned", c->rank);
    if (kind == nbx::kDevErrDebugCheck)
        return fail(NBX_ERR_DEBUG, "device-side check failed at nbx_kernels.cuh:%d; context poisoned", detail);
    return fail(NBX_ERR_CUDA, "unknown device error word %d", word);
}

static int check_runnable(nbx_ctx *c, int nsteps, bool group)
{
    if (!c) return fail(NBX_ERR_ARG, "NULL context");
    if (nsteps < 0) return fail(NBX_ERR_ARG, "nsteps must be >= 0");
    if (c->device_error) return device_error_code(c, c->device_error);
    if (!c->uploaded) return fail(NBX_ERR_STATE, "run before upload");
    if (c->world > 1) {
        if (c->exchange == NBX_EXCHANGE_P2P) {
            if (!c->p2p_ready && !group)
                return fail(NBX_ERR_STATE, "P2P exchange (the default) needs nbx_p2p_attach first; or set \"exchange\" to NBX_EXCHANGE_NCCL");
        } else if (!c->comm) {
            return fail(NBX_ERR_STATE, "NCCL exchange needs nbx_comm_init / nbx_comm_init_all first");
        }
    }
    return NBX_OK;
}

// Host-side abort: make every queued or spinning step of this context return (the kernels poll the
// error word), then drain its streams.  Used when a multi-GPU enqueue fails half way.
static void poison(nbx_ctx *c)
{
    if (!c || !c->counters) return;
    cudaSetDevice(c->device);
    static const int word = nbx::kDevErrHostAbort;
    cudaMemcpyAsync(c->counters + 3, &word, sizeof(int), cudaMemcpyHostToDevice, c->comm_stream);
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamSynchronize(c->stream);
    if (!c->device_error) c->device_error = word;
    cudaGetLastError();
}

// After the steps have been synchronised: did the device report a peer timeout / failed check?
static int read_device_error(nbx_ctx *c)
{
#ifndef NBX_DEBUG
    if (!(c->world > 1 && c->exchange == NBX_EXCHANGE_P2P)) return NBX_OK;   // nothing can set it
#endif
    int word = 0;
    CU(cudaMemcpy(&word, c->counters + 3, sizeof(int), cudaMemcpyDeviceToHost));
    return word ? device_error_code(c, word) : NBX_OK;
}

int nbx_run(nbx_ctx *c, int nsteps, double *kenergy_out, double *seconds_out)
{
    int rc = check_runnable(c, nsteps, false);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if ((rc = resolve(c))) return rc;
    if ((rc = ensure_ke(c, nsteps))) return rc;
    const bool p2p = c->world > 1 && c->exchange == NBX_EXCHANGE_P2P;
    if (p2p && c->comm) {
        // In-stream barrier: no rank starts stepping before every rank has enqueued this run (and so has
        // finished its nbx_upload: a peer's epilogue stores into OUR replica) -- and the in-kernel
        // peer wait then only ever covers the skew between GPUs that ARE stepping.
        CU(cudaMemsetAsync(c->counters + 4, 0, sizeof(int), c->stream));
        NC(g_nccl.AllReduce(c->counters + 4, c->counters + 4, 1, ncclInt, ncclSum, c->comm, c->stream));
    }
    CU(cudaMemsetAsync(c->counters + 1, 0, sizeof(int), c->stream));   // dev_step = 0
    CU(cudaEventRecord(c->ev0, c->stream));
    if ((rc = enqueue_steps(c, nsteps)) || (rc = finish_run(c, nsteps, c->comm != nullptr))) {
        const std::string why = g_err;
        poison(c);
        g_err = why;
        return rc;
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    if (kenergy_out && nsteps > 0)
        CU(cudaMemcpyAsync(kenergy_out, c->ke_dev, (size_t)nsteps * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if ((rc = read_device_error(c))) return rc;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_run_seconds = ms * 1e-3;
    c->kernel_seconds_total += c->last_run_seconds;
    if (seconds_out) *seconds_out = c->last_run_seconds;
    return NBX_OK;
}

// One process, several GPUs on an NVSwitch: move the position replicas into allocations bound to a
// multicast team, so the epilogue exchange is one multimem.st per record instead of world-1 stores.
// Not being able to is not an error (unless "multicast" = 1): the unicast NVLink stores remain.
static int try_multicast_group(nbx_ctx **ctxs, int count)
{
    if (ctxs[0]->opt_multicast == 0 || ctxs[0]->mc_active) return NBX_OK;
    std::vector<int> devices((size_t)count);
    for (int g = 0; g < count; ++g) devices[g] = ctxs[g]->device;
    bool distinct = true;
    for (int g = 0; g < count; ++g)
        for (int h = g + 1; h < count; ++h) distinct = distinct && devices[g] != devices[h];
    const size_t bytes = (size_t)ctxs[0]->n_pad * sizeof(float4);
    std::vector<nbx_mc::Buffer> team[2];
    std::string why = distinct ? "" : "two shards share a GPU";
    for (int b = 0; b < 2 && why.empty(); ++b) why = nbx_mc::create_team(devices, bytes, team[b]);
    if (!why.empty()) {
        for (auto &t : team)
            for (auto &buf : t) nbx_mc::release(buf);
        cudaGetLastError();
        if (ctxs[0]->opt_multicast == 1) return fail(NBX_ERR_CUDA, "NVSwitch multicast required but unavailable: %s", why.c_str());
        if (std::getenv("NBX_VERBOSE"))
            std::fprintf(stderr, "nbx: NVSwitch multicast unavailable (%s); the exchange uses unicast NVLink stores\n", why.c_str());
        for (int g = 0; g < count; ++g) ctxs[g]->mc_note = why;
        return NBX_OK;
    }
    for (int g = 0; g < count; ++g) {
        nbx_ctx *c = ctxs[g];
        CU(cudaSetDevice(c->device));
        CU(cudaStreamSynchronize(c->stream));
        for (int b = 0; b < 2; ++b) {
            float4 *fresh = reinterpret_cast<float4 *>(team[b][g].uc);
            CU(cudaMemcpy(fresh, c->pos[b], bytes, cudaMemcpyDeviceToDevice));
            CU(cudaFree(c->pos[b]));
            c->pos[b] = fresh;
            c->mcbuf[b] = team[b][g];
        }
        c->mc_active = true;
        c->mc_note = "active";
        c->resolved = false;            // graphs bake the old addresses in
    }
    return NBX_OK;
}

// One process, several GPUs: map every context's replicas and flags into every other context.
static int attach_group(nbx_ctx **ctxs, int count)
{
    int mrc = try_multicast_group(ctxs, count);
    if (mrc) return mrc;
    std::vector<unsigned char> blobs((size_t)count * NBX_P2P_BLOB_BYTES);
    int rc;
    for (int g = 0; g < count; ++g)
        if ((rc = nbx_p2p_export(ctxs[g], blobs.data() + (size_t)g * NBX_P2P_BLOB_BYTES))) return rc;
    for (int g = 0; g < count; ++g)
        if ((rc = nbx_p2p_attach(ctxs[g], blobs.data()))) return rc;
    return NBX_OK;
}

int nbx_p2p_attach_group(nbx_ctx **ctxs, int count)
{
    if (!ctxs || count < 1 || count > nbx::kMaxWorld) return fail(NBX_ERR_ARG, "bad context array");
    for (int g = 0; g < count; ++g)
        if (!ctxs[g] || ctxs[g]->world != count || ctxs[g]->rank != g) return fail(NBX_ERR_ARG, "ctxs[%d] is not rank %d of %d", g, g, count);
    return count > 1 ? attach_group(ctxs, count) : NBX_OK;
}

int nbx_run_group(nbx_ctx **ctxs, int count, int nsteps, double *kenergy_out, double *seconds_out)
{
    if (!ctxs || count < 1) return fail(NBX_ERR_ARG, "bad context array");
    int rc;
    for (int g = 0; g < count; ++g) {
        if ((rc = check_runnable(ctxs[g], nsteps, true))) return rc;
        const nbx_ctx *a = ctxs[g], *b = ctxs[0];
        if (a->world != count || a->rank != g) return fail(NBX_ERR_ARG, "ctxs[%d] is not rank %d of %d", g, g, count);
        if (a->n != b->n || a->exchange != b->exchange || a->opt_splits != b->opt_splits || a->opt_variant != b->opt_variant ||
            a->opt_accurate != b->opt_accurate || a->dt != b->dt || a->G != b->G || a->eps2 != b->eps2)
            return fail(NBX_ERR_ARG, "ctxs[%d] differs from ctxs[0] in n / exchange / j_splits / variant / accurate / constants", g);
    }
    if (count > 1 && ctxs[0]->exchange == NBX_EXCHANGE_P2P) {
        bool ready = true;
        for (int g = 0; g < count; ++g) ready = ready && ctxs[g]->p2p_ready;
        if (!ready && attach_group(ctxs, count) != NBX_OK) {
            // no peer access between some pair: every context takes the NCCL all-gather instead
            const std::string why = g_err;
            for (int g = 0; g < count; ++g) {
                if (!ctxs[g]->comm)
                    return fail(NBX_ERR_STATE, "P2P exchange unavailable (%s) and no communicator for the NCCL fallback: call nbx_comm_init_all",
                                why.c_str());
                ctxs[g]->exchange = NBX_EXCHANGE_NCCL;
                ctxs[g]->resolved = false;
            }
        }
    }
    for (int g = 0; g < count; ++g) {
        CU(cudaSetDevice(ctxs[g]->device));
        if ((rc = resolve(ctxs[g]))) return rc;
        if ((rc = ensure_ke(ctxs[g], nsteps))) return rc;
        CU(cudaMemsetAsync(ctxs[g]->counters + 1, 0, sizeof(int), ctxs[g]->stream));
        CU(cudaEventRecord(ctxs[g]->ev0, ctxs[g]->stream));
    }
    const bool nccl_x = count > 1 && ctxs[0]->exchange != NBX_EXCHANGE_P2P;
    auto enqueue_all = [&]() -> int {
        if (count == 1) return enqueue_steps(ctxs[0], nsteps);
        // step-major order: with the P2P exchange a GPU's step s+1 waits until every peer has
        // finished step s, so no GPU may be queued far ahead of the others by this one thread.
        for (int s = 0; s < nsteps; ++s) {
            for (int g = 0; g < count; ++g) {
                CU(cudaSetDevice(ctxs[g]->device));
                if ((rc = step_compute(ctxs[g]))) return rc;
            }
            if (nccl_x) {
                NC(g_nccl.GroupStart());
                for (int g = 0; g < count; ++g) {
                    CU(cudaSetDevice(ctxs[g]->device));
                    if ((rc = step_exchange(ctxs[g]))) { g_nccl.GroupEnd(); return rc; }
                }
                NC(g_nccl.GroupEnd());
                for (int g = 0; g < count; ++g) {
                    CU(cudaSetDevice(ctxs[g]->device));
                    if ((rc = step_exchange_mark(ctxs[g]))) return rc;
                }
            }
        }
        for (int g = 0; g < count; ++g) {            // kinetic energy is summed on the host below
            CU(cudaSetDevice(ctxs[g]->device));
            if ((rc = finish_run(ctxs[g], nsteps, false))) return rc;
        }
        return NBX_OK;
    };
    if ((rc = enqueue_all())) {
        // work may be queued on some GPUs and not on others: abort all of it, leave no kernel spinning
        const std::string why = g_err;
        for (int g = 0; g < count; ++g) poison(ctxs[g]);
        g_err = why;
        return rc;
    }
    std::vector<double> tmp((size_t)std::max(nsteps, 1));
    if (kenergy_out) std::fill(kenergy_out, kenergy_out + nsteps, 0.0);
    double tmax = 0.0;
    int first_err = NBX_OK;
    std::string first_why;
    for (int g = 0; g < count; ++g) {
        nbx_ctx *c = ctxs[g];
        CU(cudaSetDevice(c->device));
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if ((rc = read_device_error(c)) && first_err == NBX_OK) { first_err = rc; first_why = g_err; }
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->last_run_seconds = ms * 1e-3;
        c->kernel_seconds_total += c->last_run_seconds;
        tmax = std::max(tmax, c->last_run_seconds);
        if (kenergy_out && nsteps > 0) {
            CU(cudaMemcpy(tmp.data(), c->ke_dev, (size_t)nsteps * sizeof(double), cudaMemcpyDeviceToHost));
            for (int s = 0; s < nsteps; ++s) kenergy_out[s] += tmp[s];   // rank order: deterministic
        }
    }
    if (first_err) { g_err = first_why; return first_err; }
    if (seconds_out) *seconds_out = tmax;
    return NBX_OK;
}

int nbx_accelerations(nbx_ctx *c, float *ax, float *ay, float *az)
{
    if (!c || !ax || !ay || !az) return fail(NBX_ERR_ARG, "NULL argument");
    if (!c->uploaded) return fail(NBX_ERR_STATE, "accelerations before upload");
    CU(cudaSetDevice(c->device));
    int rc = resolve(c);
    if (rc) return rc;
    if (!c->acc) CU(cudaMalloc(&c->acc, (size_t)c->i_count * sizeof(float4)));
    if ((rc = ensure_ke(c, 1))) return rc;
    if ((rc = launch_step(c, c->cur, c->acc))) return rc;
    const size_t cnt = (size_t)c->i_count;
    if ((rc = ensure_stage(c, 3 * cnt))) return rc;
    nbx::acc_soa_kernel<<<((int)cnt + 255) / 256, 256, 0, c->stream>>>(c->acc, (int)cnt, c->stage);
    CU(cudaGetLastError());
    c->aux_launches++;
    float *dst[3] = {ax, ay, az};
    for (int k = 0; k < 3; ++k)
        CU(cudaMemcpyAsync(dst[k], c->stage + k * cnt, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_simulate(int n, int nsteps, float dt, float G, float eps2,
                 float *px, float *py, float *pz, float *vx, float *vy, float *vz,
                 const float *mass, double *kenergy_out, double *seconds_out)
{
    nbx_ctx *c = nullptr;
    int rc = nbx_create(&c, n, 0, 0, 1, dt, G, eps2);
    if (rc) return rc;
    if ((rc = nbx_upload(c, px, py, pz, vx, vy, vz, mass)) == NBX_OK)
        if ((rc = nbx_run(c, nsteps, kenergy_out, seconds_out)) == NBX_OK)
            rc = nbx_download(c, px, py, pz, vx, vy, vz);
    nbx_destroy(c);
    return rc;
}

int nbx_trace_read(nbx_ctx *c, unsigned long long *out, size_t capacity_words, int *steps, int *ctas)
{
    if (!c || !steps || !ctas) return fail(NBX_ERR_ARG, "NULL argument");
#ifndef NBX_TRACE
    (void)out; (void)capacity_words;
    return fail(NBX_ERR_STATE, "this library was built without -DNBX_TRACE (make trace -> libnbx_trace.so)");
#else
    *steps = c->trace_steps; *ctas = c->trace_ctas;
    const size_t words = (size_t)c->trace_steps * c->trace_ctas * nbx::kTraceWords;
    if (!c->trace || words == 0) return fail(NBX_ERR_STATE, "no trace recorded: set option trace_steps before the run");
    if (!out || capacity_words < words) return fail(NBX_ERR_ARG, "trace needs %zu words", words);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpy(out, c->trace, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return NBX_OK;
#endif
}

// ---- multi-GPU plumbing --------------------------------------------------------
int nbx_comm_unique_id(void *id_out)
{
    if (!id_out) return fail(NBX_ERR_ARG, "id_out is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == NBX_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return NBX_OK;
}

int nbx_comm_init(nbx_ctx *c, const void *id_bytes)
{
    if (!c || !id_bytes) return fail(NBX_ERR_ARG, "NULL argument");
    int rc = nccl_load();
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    if (c->comm) { g_nccl.CommDestroy(c->comm); c->comm = nullptr; }   // re-initialisation replaces the old one
    NC(g_nccl.CommInitRank(&c->comm, c->world, id, c->rank));
    return NBX_OK;
}

int nbx_comm_init_all(nbx_ctx **ctxs, int count)
{
    if (!ctxs || count < 1 || count > nbx::kMaxWorld) return fail(NBX_ERR_ARG, "bad context array");
    int rc = nccl_load();
    if (rc) return rc;
    std::vector<int> devs(count);
    std::vector<ncclComm_t> comms(count);
    for (int g = 0; g < count; ++g) {
        if (!ctxs[g] || ctxs[g]->world != count || ctxs[g]->rank != g) return fail(NBX_ERR_ARG, "ctxs[%d] is not rank %d of %d", g, g, count);
        devs[g] = ctxs[g]->device;
    }
    for (int g = 0; g < count; ++g)
        if (ctxs[g]->comm) { g_nccl.CommDestroy(ctxs[g]->comm); ctxs[g]->comm = nullptr; }
    NC(g_nccl.CommInitAll(comms.data(), count, devs.data()));
    for (int g = 0; g < count; ++g) ctxs[g]->comm = comms[g];
    return NBX_OK;
}

struct P2PBlob {
    uint32_t magic;
    int32_t rank;
    int64_t pid;
    void *raw[3];                    // pos[0], pos[1], flags (valid inside the exporting process)
    int32_t device;
    int32_t pad;
    cudaIpcMemHandle_t h[3];
};
static_assert(sizeof(P2PBlob) <= NBX_P2P_BLOB_BYTES, "P2P blob size");

int nbx_p2p_export(nbx_ctx *c, void *blob_out)
{
    if (!c || !blob_out) return fail(NBX_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    P2PBlob b;
    std::memset(&b, 0, sizeof b);
    b.magic = 0x4e425850u;
    b.rank = c->rank;
    b.pid = (int64_t)getpid();
    b.device = c->device;
    b.raw[0] = c->pos[0]; b.raw[1] = c->pos[1]; b.raw[2] = c->flags;
    // IPC handles serve peers in OTHER processes; inside one process the raw pointers are used, so a
    // platform without CUDA IPC can still run the one-process group
    b.pad = 1;
    void *bufs[3] = {c->pos[0], c->pos[1], c->flags};
    if (c->mc_active) {
        b.pad = 0;       // multicast-bound replicas are driver-API (cuMemCreate) memory: legacy CUDA IPC does not apply
    } else {
        for (int k = 0; k < 3; ++k)
            if (cudaIpcGetMemHandle(&b.h[k], bufs[k]) != cudaSuccess) { cudaGetLastError(); b.pad = 0; }
    }
    std::memset(blob_out, 0, NBX_P2P_BLOB_BYTES);
    std::memcpy(blob_out, &b, sizeof b);
    return NBX_OK;
}

// min over the ranks of a small integer (NCCL on the context's stream): the consensus step of collective set-up
static int consensus_min(nbx_ctx *c, int mine, int *all)
{
    int *word = c->counters + 4;
    CU(cudaMemcpyAsync(word, &mine, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(word, word, 1, ncclInt, ncclMin, c->comm, c->stream));
    CU(cudaMemcpyAsync(all, word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

// One process per GPU: put the replicas behind an NVSwitch multicast team shared ACROSS the processes
// (nbx_multicast.hpp, second half).  Collective over the communicator; every phase ends in a consensus so that
// either all ranks switch to the multicast stores or none does.  Unavailable is not an error (unless "multicast" = 1).
static int try_multicast_procs(nbx_ctx *c, long long pid0, const void *tag0)
{
    if (c->opt_multicast == 0 || c->mc_active || !c->comm) return NBX_OK;
    char name[96];
    std::snprintf(name, sizeof name, "nbx-mc-%lld-%p", pid0, tag0);
    const size_t bytes = (size_t)c->n_pad * sizeof(float4);
    nbx_mc::Buffer bufs[2];
    std::string why;
    int all = 0, rc;
    auto give_up = [&](const std::string &w) {
        nbx_mc::release(bufs[0]); nbx_mc::release(bufs[1]);
        cudaGetLastError();
        c->mc_note = w;
        if (c->opt_multicast == 1) return fail(NBX_ERR_CUDA, "NVSwitch multicast required but unavailable: %s", w.c_str());
        if (std::getenv("NBX_VERBOSE"))
            std::fprintf(stderr, "nbx: rank %d: NVSwitch multicast unavailable (%s); the exchange uses unicast NVLink stores\n", c->rank, w.c_str());
        return (int)NBX_OK;
    };
    const char *phase[3] = {"open", "join", "bind"};
    for (int ph = 0; ph < 3; ++ph) {
        why = ph == 0 ? nbx_mc::mp_open_team(c->rank, c->world, c->device, bytes, name, bufs, 2)
              : ph == 1 ? nbx_mc::mp_join(bufs, 2) : nbx_mc::mp_bind_and_map(bufs, 2);
        if ((rc = consensus_min(c, why.empty() ? 1 : 0, &all))) { give_up("consensus failed"); return rc; }
        if (!all) return give_up(why.empty() ? std::string("another rank failed in phase ") + phase[ph] : why);
    }
    CU(cudaSetDevice(c->device));
    for (int b = 0; b < 2; ++b) {
        float4 *fresh = reinterpret_cast<float4 *>(bufs[b].uc);
        CU(cudaMemcpy(fresh, c->pos[b], bytes, cudaMemcpyDeviceToDevice));
        c->retired_pos[b] = c->pos[b];          // peers hold CUDA IPC mappings of it: free it with the context
        c->pos[b] = fresh;
        c->mcbuf[b] = bufs[b];
    }
    c->mc_active = true;
    c->mc_note = "active (across processes)";
    c->resolved = false;
    return NBX_OK;
}

static int map_peers(nbx_ctx *c, const unsigned char *base)
{
    for (int g = 0; g < c->world; ++g) {
        P2PBlob b;
        std::memcpy(&b, base + (size_t)g * NBX_P2P_BLOB_BYTES, sizeof b);
        if (b.magic != 0x4e425850u || b.rank != g) return fail(NBX_ERR_ARG, "blob %d is malformed", g);
        if (g == c->rank) {
            c->peer_pos[0][g] = c->pos[0]; c->peer_pos[1][g] = c->pos[1]; c->peer_flags[g] = c->flags;
            continue;
        }
        void *ptr[3];
        if (b.pid == (int64_t)getpid() && b.device == c->device) {
            for (int k = 0; k < 3; ++k) ptr[k] = b.raw[k];        // two shards on one GPU: plain pointers
        } else if (b.pid == (int64_t)getpid()) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, c->device, b.device));
            if (!can) return fail(NBX_ERR_CUDA, "device %d cannot access peer %d", c->device, b.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(NBX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e));
            cudaGetLastError();
            for (int k = 0; k < 3; ++k) ptr[k] = b.raw[k];
        } else {
            if (!b.pad) return fail(NBX_ERR_CUDA, "rank %d could not export CUDA IPC handles", g);
            for (int k = 0; k < 3; ++k) {
                CU(cudaIpcOpenMemHandle(&ptr[k], b.h[k], cudaIpcMemLazyEnablePeerAccess));
                c->ipc_opened.push_back(ptr[k]);
            }
        }
        c->peer_pos[0][g] = static_cast<float4 *>(ptr[0]);
        c->peer_pos[1][g] = static_cast<float4 *>(ptr[1]);
        c->peer_flags[g] = static_cast<int *>(ptr[2]);
    }
    return NBX_OK;
}

int nbx_p2p_attach(nbx_ctx *c, const void *blobs)
{
    if (!c || !blobs) return fail(NBX_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    const unsigned char *base = static_cast<const unsigned char *>(blobs);
    const int map_rc = map_peers(c, base);
    P2PBlob b0, bl;
    std::memcpy(&b0, base, sizeof b0);
    std::memcpy(&bl, base + (size_t)(c->world - 1) * NBX_P2P_BLOB_BYTES, sizeof bl);
    const bool across_processes = c->world > 1 && b0.magic == 0x4e425850u && bl.magic == 0x4e425850u && b0.pid != bl.pid;
    if (across_processes && c->comm && c->opt_multicast != 0) {
        // the multicast set-up below is collective: first agree that every rank got this far
        const std::string why = g_err;
        int all = 0, rc;
        if ((rc = consensus_min(c, map_rc == NBX_OK ? 1 : 0, &all))) return rc;
        if (!all) {
            if (map_rc) { g_err = why; return map_rc; }
            return fail(NBX_ERR_CUDA, "another rank could not map its peers' buffers");
        }
        c->p2p_ready = true;
        return try_multicast_procs(c, (long long)b0.pid, b0.raw[0]);
    }
    if (map_rc) return map_rc;
    c->p2p_ready = true;
    return NBX_OK;
}

// ---- host helpers ----------------------------------------------------------------
void nbx_ic_uniform(int n, float *px, float *py, float *pz, float *vx, float *vy, float *vz, float *mass)
{
    nbx_ic::uniform_pos(n, px, py, pz);
    nbx_ic::uniform_vel(n, vx, vy, vz);
    nbx_ic::uniform_mass(n, mass);
}

void nbx_ic_plummer(int n, float *px, float *py, float *pz, float *vx, float *vy, float *vz, float *mass)
{
    nbx_ic::plummer_pos(n, px, py, pz);
    nbx_ic::uniform_vel(n, vx, vy, vz);
    nbx_ic::uniform_mass(n, mass);
}

double nbx_gflop_per_step(int n)
{
    const double nd = (double)n;
    return 1e-9 * ((11. + 18.) * nd * nd + nd * 19.);
}

int nbx_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(NBX_ERR_ARG, "ptr is NULL");
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return NBX_OK;
}

int nbx_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return NBX_OK;
}

}  // extern "C"

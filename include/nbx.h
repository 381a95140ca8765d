/*
 * include/nbx.h -- C ABI of libnbx.so, the B200-native backend for the O(N^2)
 * force / explicit-Euler / kinetic-energy path of NTHU-SC/nbody-demo-2023.
 *
 * The reference has no FFI: its backends are chosen at LINK time by which
 * translation unit defines `void GSimulation::start()`
 * (ver5_all/Makefile:1-104, ver5_all/CMakeLists.txt:17-49; CUDA backend:
 * ver5_all/programming_models/cuda/Compute.cu:69-232).  This header is the
 * boundary such a backend TU calls: plain pointers and sizes, no C++ or torch
 * types, every function returns 0 on success and a non-zero code otherwise
 * (nbx_last_error() gives the text).  No exception crosses it.  There is no CPU
 * fallback behind it: without a CUDA device every compute entry point fails.
 *
 * Each entry point cites the reference lines it replaces.
 */
#ifndef NBX_H
#define NBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBX_ABI_VERSION 2

#if defined(__GNUC__)
#define NBX_API __attribute__((visibility("default")))
#else
#define NBX_API
#endif

typedef struct nbx_ctx nbx_ctx;

enum {
    NBX_OK = 0,
    NBX_ERR_ARG = 1,      /* bad argument                                   */
    NBX_ERR_CUDA = 2,     /* a CUDA runtime/driver call failed              */
    NBX_ERR_NCCL = 3,     /* an NCCL call failed / NCCL not available       */
    NBX_ERR_STATE = 4,    /* call sequence error (e.g. run before upload)   */
    NBX_ERR_NODEVICE = 5, /* no usable sm_100 device                        */
    NBX_ERR_PEER = 6,     /* multi-GPU: a peer GPU did not finish its step within "peer_timeout_ms",
                             or the run was aborted; the context is poisoned -- destroy it */
    NBX_ERR_DEBUG = 7     /* debug build (libnbx_debug.so) only: a device-side index/ticket
                             check failed; nbx_last_error() names the source line          */
};

/* How updated positions reach the other GPUs after each step (multi-GPU only).
 * ONE default everywhere (library, nbody.x, bench.py): NBX_EXCHANGE_P2P.  It needs the peers'
 * buffers mapped (nbx_p2p_attach; nbx_run_group does it by itself inside one process); where
 * that is impossible (no peer access / no CUDA IPC) the callers in this repo fall back to
 * NBX_EXCHANGE_NCCL on all ranks together (dist.make_sharded_context, GSimulation::start). */
enum {
    NBX_EXCHANGE_NCCL = 0, /* ncclAllGather of the updated shard (stream ordered)          */
    NBX_EXCHANGE_P2P = 1,  /* DEFAULT: the force kernel's epilogue stores each updated body straight
                              into every peer's replica over NVLink; flags replace the collective */
    NBX_EXCHANGE_NCCL_OVERLAP = 2 /* ncclAllGather on a side stream, hidden behind the next step's
                              force work on the rank's own j-shard (a step = two launches)   */
};

typedef struct nbx_info {
    int abi_version;
    int device;             /* CUDA ordinal                                           */
    int sm_count;
    int sm_clock_khz;       /* cudaDevAttrClockRate                                   */
    int n;                  /* bodies asked for                                       */
    int n_pad;              /* bodies stored (zero-mass padding to the shard grain)   */
    int rank, world;
    int i_begin, i_count;   /* this context's i-shard [i_begin, i_begin+i_count)      */
    int threads;            /* CTA size of the force kernel                           */
    int bodies_per_thread;  /* register-blocked i-bodies per thread                   */
    int tile_bodies;        /* j-bodies per TMA stage                                 */
    int stages;             /* TMA ring depth                                         */
    int i_tiles;            /* CTAs along i                                           */
    int whole_tiles;        /* leading i-tiles that run unsplit                       */
    int j_splits;           /* j-split count of the remaining i-tiles (1 = none)      */
    int ctas_per_sm;        /* resident CTAs per SM (occupancy query)                 */
    int use_graph;          /* steps replayed from a CUDA graph                       */
    int exchange;           /* NBX_EXCHANGE_*                                         */
    int variant;            /* kernel-shape index in use (after auto-selection)       */
    long long kernel_launches;   /* force-kernel launches since create                */
    long long aux_launches;      /* pack/unpack launches, and the per-step record rewrite of the q-scaled shapes */
    double last_run_seconds;     /* device time of the last nbx_run (CUDA events)     */
    double kernel_seconds_total; /* device time summed over every nbx_run             */
    int device_error;       /* last device error word (0 = none): low byte 1 = peer timeout
                               (peer rank << 8), 2 = aborted, 3 = debug check (source line << 8) */
    int peer_timeout_ms;    /* bound on the in-kernel wait for a peer's previous step */
    int multicast;          /* 1: the P2P exchange goes through an NVSwitch multicast mapping (one
                               multimem.st per record instead of world-1 NVLink stores) */
    int reserved;
} nbx_info;

/* ---- library ---------------------------------------------------------------- */
NBX_API int nbx_abi_version(void);
NBX_API const char *nbx_last_error(void);      /* thread-local, never NULL */
NBX_API int nbx_device_count(int *count);      /* number of visible CUDA devices (0 is not an error) */

/* ---- context ----------------------------------------------------------------
 * Replaces the allocation block of GSimulation::start()
 * (ver5/GSimulation.cpp:102-114; cuda/Compute.cu:76-108) and the hard-coded
 * constants (ver0/GSimulation.cpp:30 dt, :114 softening, :116 G).
 * `n` bodies in total; this context owns the i-shard `rank` of `world`
 * (cpu/Compute.cpp:47-58 is the reference's i-range precedent) on CUDA device
 * `device`.  Single GPU: rank 0, world 1. */
NBX_API int nbx_create(nbx_ctx **out, int n, int device, int rank, int world,
               float dt, float G, float eps2);
NBX_API void nbx_destroy(nbx_ctx *ctx);

/* Tuning / mode knobs; all optional, call before the first nbx_run.
 *   "j_splits"  >=1 force a j-split count, 0 = auto (fills the SMs at small N)
 *   "graph"     1/0 replay steps from a CUDA graph (auto: on for small N)
 *   "accurate"  1/0 fold the float lane sums into DOUBLE accumulators every 4 j tiles (forces 3e-8 from fp64 at
 *               N = 1 M, -3 % throughput).  The default kernel already keeps float sums short (two-level FLOAT
 *               accumulation, 3e-7); variant 3 is the single-accumulator scheme of the reference's loops.
 *   "pdl"       1/0 programmatic dependent launch between consecutive steps (-1 = auto: many-wave grids, and
 *               grids that fit the SMs once -- those are then run one CTA per SM)
 *   "multicast" P2P exchange: -1 auto = use NVSwitch multicast (cuMulticast* + multimem.st) when the driver
 *               offers it, 0 never, 1 fail if unavailable; NBX_VERBOSE=1 prints why it was not used.  Works inside
 *               one process (nbx_run_group / nbx_p2p_attach_group) and across processes (nbx_p2p_attach, when a
 *               communicator exists: the set-up is collective and must be called with the same option on every rank)
 *   "smem_pad_kb"  extra dynamic shared memory per CTA in KiB (tuning: caps the resident CTAs per SM)
 *   "exchange"  NBX_EXCHANGE_* (default NBX_EXCHANGE_P2P)
 *   "peer_timeout_ms"  P2P exchange: how long a step may wait inside the kernel for a peer GPU to
 *               finish the previous step before the run fails with NBX_ERR_PEER (default 30000)
 *   "variant"   index into the compiled kernel-shape table (see nbx_variant_name), -1 = auto */
NBX_API int nbx_set_option(nbx_ctx *ctx, const char *key, long long value);
NBX_API int nbx_get_info(const nbx_ctx *ctx, nbx_info *out);
/* The launch plan nbx_run would use for shard `rank` of `world` on a GPU with `sm_count` SMs --
 * padding, shard, kernel shape, unsplit tiles, j-split count -- without touching a device
 * (host arithmetic only; fills the planning fields of nbx_info).  variant/j_splits: -1/0 = auto. */
NBX_API int nbx_plan(int n, int rank, int world, int sm_count, int exchange, long long variant,
                     long long j_splits, nbx_info *out);
/* Trace build only (make trace -> libnbx_trace.so; the product library returns NBX_ERR_STATE): per-CTA
 * %globaltimer stamps of the first "trace_steps" (option) steps of the last run, 6 words per CTA per step:
 * start, first j tile landed, j sweep done, exit, SM id, 1 if this CTA ran the split-combine + update.
 * tools/trace_steps.py turns them into the launch-gap / prologue / sweep / combine split of a step. */
NBX_API int nbx_trace_read(nbx_ctx *ctx, unsigned long long *out, size_t capacity_words, int *steps, int *ctas);
NBX_API int nbx_variant_count(void);
NBX_API const char *nbx_variant_name(int idx);

/* ---- state ------------------------------------------------------------------
 * Host SoA arrays of n floats each, caller-owned (the reference's ParticleSoA,
 * ver3/Particle.hpp:43-58).  upload replaces the one-off H2D copies at
 * cuda/Compute.cu:115-123 (and makes the per-step ones at :152-154 unnecessary:
 * state stays on the device).  Every rank passes the FULL arrays; a context keeps
 * all positions and masses and the velocities of its own shard.
 * download replaces the D2H at cuda/Compute.cu:164-166: positions of all n bodies
 * and velocities of this context's shard only ([i_begin, i_begin+i_count) of
 * vx/vy/vz are written, the rest is left untouched).
 * In a multi-process job every rank must have returned from nbx_upload before any rank's
 * nbx_run starts stepping: nbx_run enforces it with an in-stream NCCL barrier when a communicator
 * exists; without one (P2P exchange, nbx_comm_init never called) the host must barrier itself. */
NBX_API int nbx_upload(nbx_ctx *ctx, const float *px, const float *py, const float *pz,
               const float *vx, const float *vy, const float *vz, const float *mass);
NBX_API int nbx_download(nbx_ctx *ctx, float *px, float *py, float *pz,
                 float *vx, float *vy, float *vz);
/* Sharded variants for multi-GPU jobs: the same FULL caller-owned arrays are passed, but each
 * context moves only its own i-shard over PCIe (28 B/body of the shard up, 24 B/body down) --
 * the reference's MPI code ships every array to every rank each step
 * (ver5_all/GSimulation.cpp:170-189).  upload_sharded is COLLECTIVE (every rank calls it; the
 * packed records are then all-gathered GPU-to-GPU with ncclAllGather, so it needs nbx_comm_init
 * first); upload_group is its one-process form (peer copies between the contexts' GPUs, NCCL
 * not needed).  download_shard writes only [i_begin, i_begin+i_count) of all six arrays. */
NBX_API int nbx_upload_sharded(nbx_ctx *ctx, const float *px, const float *py, const float *pz,
               const float *vx, const float *vy, const float *vz, const float *mass);
NBX_API int nbx_upload_group(nbx_ctx **ctxs, int count, const float *px, const float *py, const float *pz,
               const float *vx, const float *vy, const float *vz, const float *mass);
NBX_API int nbx_download_shard(nbx_ctx *ctx, float *px, float *py, float *pz,
                 float *vx, float *vy, float *vz);

/* ---- the hot path -------------------------------------------------------------
 * Advance nsteps explicit-Euler steps on the device: the step loop of
 * GSimulation::start() (ver0/GSimulation.cpp:127-173; force :130-150, update
 * :153-165, energy :151,167-173).  kenergy_out[s] = 0.5 * sum m v^2 after step s
 * (all shards: the multi-GPU sum is taken inside), may be NULL.  seconds_out =
 * device time of the loop (CUDA events on the launching stream), may be NULL.
 * Blocks until the steps are done.  In a multi-process job every rank calls it
 * with the same nsteps.  P2P exchange without a communicator (nbx_comm_init never called):
 * kenergy_out holds THIS shard's part of the sum and the caller adds the ranks' values.
 * A peer that stops stepping makes the call fail with NBX_ERR_PEER after "peer_timeout_ms"
 * instead of hanging; the context is then poisoned (every later call fails the same way). */
NBX_API int nbx_run(nbx_ctx *ctx, int nsteps, double *kenergy_out, double *seconds_out);

/* Accelerations only (no update): a_i for this context's shard from the current
 * positions, for the sampled-fp64 checks.  ax/ay/az: host, i_count floats each. */
NBX_API int nbx_accelerations(nbx_ctx *ctx, float *ax, float *ay, float *az);

/* One-call convenience = create + upload + run + download + destroy on device 0
 * with host buffers; what a backend's start() does end to end. */
NBX_API int nbx_simulate(int n, int nsteps, float dt, float G, float eps2,
                 float *px, float *py, float *pz, float *vx, float *vy, float *vz,
                 const float *mass, double *kenergy_out, double *seconds_out);

/* ---- multi-GPU plumbing ---------------------------------------------------------
 * The reference's multi-device precedent is MPI: replicate positions, shard i
 * (ver5_all/GSimulation.cpp:170-214, cpu/Compute.cpp:47-58,95-97).  Here:
 *  - one process per GPU: rank 0 calls nbx_comm_unique_id, the host program ships
 *    the 128 bytes to every rank (torch.distributed / MPI / a file), every rank
 *    calls nbx_comm_init;
 *  - one process, several GPUs: nbx_comm_init_all over an array of contexts, then
 *    nbx_run_group drives them together.
 * P2P exchange needs the peers' buffers mapped: nbx_p2p_export gives an opaque
 * blob (NBX_P2P_BLOB_BYTES), the host all-gathers the blobs, nbx_p2p_attach maps
 * them (cudaIpc* across processes, direct peer access inside one process).
 * nbx_run_group with the (default) P2P exchange attaches its contexts by itself and needs no
 * communicator at all; if a GPU cannot reach a peer it switches every context to the NCCL
 * all-gather (then nbx_comm_init_all must have been called). */
#define NBX_UNIQUE_ID_BYTES 128
#define NBX_P2P_BLOB_BYTES 256
NBX_API int nbx_comm_unique_id(void *id_out);
NBX_API int nbx_comm_init(nbx_ctx *ctx, const void *id);
NBX_API int nbx_comm_init_all(nbx_ctx **ctxs, int count);
NBX_API int nbx_run_group(nbx_ctx **ctxs, int count, int nsteps, double *kenergy_out, double *seconds_out);
NBX_API int nbx_p2p_export(nbx_ctx *ctx, void *blob_out);
NBX_API int nbx_p2p_attach(nbx_ctx *ctx, const void *blobs /* world * NBX_P2P_BLOB_BYTES */);
/* One process: export + attach for every context of the group in one call, after trying to put the
 * replicas behind an NVSwitch multicast team (option "multicast").  What nbx_run_group does by itself;
 * public so that a host can see the failure and choose the NCCL exchange instead. */
NBX_API int nbx_p2p_attach_group(nbx_ctx **ctxs, int count);

/* ---- host helpers ------------------------------------------------------------------
 * Initial conditions with the reference's own RNG call sequence
 * (init_pos/init_vel/init_mass, ver0/GSimulation.cpp:44-93): three
 * std::mt19937(42) streams through std::uniform_real_distribution<float>.
 * Plummer positions (BASELINE config 3) are new: r = a/sqrt(u^(-2/3)-1), a = 1,
 * r <= 10a, isotropic, std::mt19937_64(20231); velocities and masses as above. */
NBX_API void nbx_ic_uniform(int n, float *px, float *py, float *pz,
                    float *vx, float *vy, float *vz, float *mass);
NBX_API void nbx_ic_plummer(int n, float *px, float *py, float *pz,
                    float *vx, float *vy, float *vz, float *mass);
/* ver0/GSimulation.cpp:122: the CLI's GFlop-per-step convention, (11+18) n^2 + 19 n. */
NBX_API double nbx_gflop_per_step(int n);
/* Pinned host memory for callers that want async-speed copies. */
NBX_API int nbx_host_alloc(void **ptr, size_t bytes);
NBX_API int nbx_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* NBX_H */

// integration/b200/Compute.cpp -- the B200 backend as the reference's own build would take it.
//
// The reference chooses its backend at LINK time: ver5_all/GSimulation.cpp defines every member of
// GSimulation except start(), and exactly one programming_models/<x>/Compute.* supplies
// `void GSimulation::start()` (ver5_all/Makefile:1-104 picks SOURCES per ARCH and :104 appends
// main.cpp GSimulation.cpp).  This file is that one translation unit for ARCH=b200.  It is
// compiled against the reference's UNMODIFIED GSimulation.hpp / GSimulation.cpp / main.cpp:
//
//   g++ -std=c++14 -O2 -I$REF/ver5_all -I$NBX/include
//       $NBX/integration/b200/Compute.cpp $REF/ver5_all/main.cpp $REF/ver5_all/GSimulation.cpp
//       -L$NBX/nbody-demo-2023_b200 -lnbx -Wl,-rpath,$NBX/nbody-demo-2023_b200 -o nbody.x
//
// (tests/test_integration_link.py does exactly this and checks the table the binary prints.)
// It follows the shape of the reference's CUDA backend (programming_models/cuda/Compute.cu:69-232):
// host SoA allocation, init_*(), header, step loop with the per-window row, summary -- but the
// whole step (force, Euler update, kinetic energy) runs on the device behind include/nbx.h and
// nothing crosses PCIe per step.  NBODY_GPUS=G shards the bodies over G GPUs of the node.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "GSimulation.hpp"
#include "nbx.h"

namespace {
void nbx_die(const char *what)
{
    std::cerr << "nbody.x (b200 backend): " << what << ": " << nbx_last_error() << std::endl;
    std::exit(1);
}
}  // namespace

void GSimulation ::start()
{
    const int n = get_npart();

    // host storage exactly as the other backends allocate it (cuda/Compute.cu:76-87); the
    // destructor in GSimulation.cpp:216-228 free()s all ten arrays, so all ten exist
    const int alignment = 32;
    const size_t bytes = ((size_t)n * sizeof(real_type) + alignment - 1) / alignment * alignment;
    particles = (ParticleSoA *)aligned_alloc(alignment, ((sizeof(ParticleSoA) + alignment - 1) / alignment) * alignment);
    real_type **arrays[10] = {&particles->pos_x, &particles->pos_y, &particles->pos_z, &particles->vel_x, &particles->vel_y,
                              &particles->vel_z, &particles->acc_x, &particles->acc_y, &particles->acc_z, &particles->mass};
    for (real_type **a : arrays) *a = (real_type *)aligned_alloc(alignment, bytes);

    init_pos();
    init_vel();
    init_acc();
    init_mass();

    // device set-up, outside the timed region like the reference's cudaMalloc + first copies (:100-123)
    const char *env_g = std::getenv("NBODY_GPUS");
    const int G = std::max(1, env_g ? std::atoi(env_g) : 1);
    const float softeningSquared = 1.e-3f;   // cuda/Compute.cu:35
    const float Gconst = 6.67259e-11f;       // cuda/Compute.cu:36
    std::vector<nbx_ctx *> ctx((size_t)G, nullptr);
    for (int g = 0; g < G; ++g)
        if (nbx_create(&ctx[g], n, g, g, G, get_tstep(), Gconst, softeningSquared)) nbx_die("nbx_create");
    if (nbx_upload_group(ctx.data(), G, particles->pos_x, particles->pos_y, particles->pos_z, particles->vel_x,
                         particles->vel_y, particles->vel_z, particles->mass))
        nbx_die("nbx_upload_group");

    print_header();

    _totTime = 0.;
    ts0 = 0;
    ts1 = 0;
    nd = double(n);
    gflops = 1e-9 * ((11. + 18.) * nd * nd + nd * 19.);   // the reference's flop convention
    av = 0.0, dev = 0.0;
    nf = 0;

    const int sfreq = std::max(1, get_sfreq());
    std::vector<double> ke((size_t)sfreq);
    const double t0 = time.start();
    for (int done = 0; done < get_nsteps();) {
        const int chunk = std::min(sfreq, get_nsteps() - done);
        ts0 += time.start();
        if (nbx_run_group(ctx.data(), G, chunk, ke.data(), nullptr)) nbx_die("nbx_run_group");
        _kenergy = (real_type)ke[(size_t)chunk - 1];
        ts1 += time.stop();
        done += chunk;
        s = done;
        print_stats();   // prints a row (and resets ts0/ts1) when s is a multiple of sfreq: GSimulation.cpp:136-158
    }
    const double t1 = time.stop();
    _totTime = (t1 - t0);
    _totFlops = gflops * get_nsteps();

    av /= (double)(nf - 2);
    dev = sqrt(dev / (double)(nf - 2) - av * av);

    print_flops();

    // final state back into the host arrays (the reference's CUDA backend has it there every step)
    for (int g = 0; g < G; ++g)
        if (nbx_download(ctx[g], particles->pos_x, particles->pos_y, particles->pos_z, particles->vel_x,
                         particles->vel_y, particles->vel_z))
            nbx_die("nbx_download");
    for (int g = 0; g < G; ++g) nbx_destroy(ctx[g]);
}

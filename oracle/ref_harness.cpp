// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference GSimulation (compiled from where it lies under
// /root/reference by oracle/Makefile, with -fno-access-control so this file can
// read the private particle storage) and dumps the final state to a binary file.
// The reference never prints positions and prints kenergy only every 50 steps at
// 5 digits (ver0/GSimulation.cpp:176-185), so this is the only way to get
// known-answer vectors out of it without editing its sources.
//
//   ref_dump_verN <nPart> <nSteps> <outfile>
//
// File layout (little endian): char magic[4]="NBXD"; int32 n; int32 nsteps;
// float kenergy; double loop_seconds; then float32[n] x7: px py pz vx vy vz mass.
#include "GSimulation.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

int main(int argc, char **argv)
{
    if (argc != 4) { std::fprintf(stderr, "usage: %s nPart nSteps outfile\n", argv[0]); return 2; }
    const int n = std::atoi(argv[1]);
    const int steps = std::atoi(argv[2]);
    GSimulation sim;
    sim.set_number_of_particles(n);
    sim.set_number_of_steps(steps);
    sim.start();

    std::vector<float> a[7];
    for (auto &v : a) v.resize(n);
    for (int i = 0; i < n; ++i) {
#ifdef REF_AOS   /* ver0-ver2: Particle[] (ver0/Particle.hpp:26-41) */
        a[0][i] = sim.particles[i].pos[0]; a[1][i] = sim.particles[i].pos[1]; a[2][i] = sim.particles[i].pos[2];
        a[3][i] = sim.particles[i].vel[0]; a[4][i] = sim.particles[i].vel[1]; a[5][i] = sim.particles[i].vel[2];
        a[6][i] = sim.particles[i].mass;
#else            /* ver3-ver8: ParticleSoA (ver3/Particle.hpp:43-58) */
        a[0][i] = sim.particles->pos_x[i]; a[1][i] = sim.particles->pos_y[i]; a[2][i] = sim.particles->pos_z[i];
        a[3][i] = sim.particles->vel_x[i]; a[4][i] = sim.particles->vel_y[i]; a[5][i] = sim.particles->vel_z[i];
        a[6][i] = sim.particles->mass[i];
#endif
    }
    const float ke = steps > 0 ? (float)sim._kenergy : 0.0f;
    const double secs = sim._totTime;
    FILE *f = std::fopen(argv[3], "wb");
    if (!f) { std::perror("fopen"); return 1; }
    const int32_t hdr[2] = {n, steps};
    std::fwrite("NBXD", 1, 4, f);
    std::fwrite(hdr, sizeof(int32_t), 2, f);
    std::fwrite(&ke, sizeof(float), 1, f);
    std::fwrite(&secs, sizeof(double), 1, f);
    for (auto &v : a) std::fwrite(v.data(), sizeof(float), (size_t)n, f);
    std::fclose(f);
    std::fflush(stdout);
    std::_Exit(0);   // skip the reference dtor (ver0:238 mismatched delete; harmless but noisy under ASan)
}

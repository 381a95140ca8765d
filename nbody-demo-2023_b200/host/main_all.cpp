// main_all.cpp -- `./nbody_all.x [nPart [nSteps [cpu|gpu|cpu+gpu [cpu_ratio [dim0 dim1]]]]]`:
// the extended CLI of the reference's multi-backend build (ver5_all/main.cpp:23-66), so scripts
// written for a ver5_all binary also run.  Differences from verN/main.cpp kept as they are there:
// nSteps is taken whenever argc > 2, the device string is echoed before the banner, the banner is
// printed from main().  "cpu" and "cpu+gpu" cannot be honoured (no CPU path) and exit non-zero.
#include <cstdlib>
#include <iostream>
#include <string>

#include "GSimulation.hpp"

int main(int argc, char **argv)
{
    GSimulation::set_banner(false);
    GSimulation sim;
    if (argc > 1) {
        sim.set_number_of_particles(std::atoi(argv[1]));
        if (argc > 2) sim.set_number_of_steps(std::atoi(argv[2]));
        if (argc > 3) {
            const std::string a = argv[3];
            std::cout << a << std::endl;
            if (a == "cpu") sim.set_devices(1);
            if (a == "gpu") sim.set_devices(2);
            if (a == "cpu+gpu") sim.set_devices(3);
        }
        if (argc > 4) sim.set_cpu_ratio((float)std::atof(argv[4]));
        if (argc > 6) {   // the reference reads argv[6] whenever argc > 5 (out of bounds for argc == 6)
            sim.set_thread_dim0(std::atoi(argv[5]));
            sim.set_thread_dim1(std::atoi(argv[6]));
        }
    }
    std::cout << "===============================" << std::endl;
    std::cout << " Initialize Gravity Simulation" << std::endl;
    sim.start();
    return 0;
}

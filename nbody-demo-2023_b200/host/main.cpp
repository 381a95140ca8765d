// main.cpp -- `./nbody.x [nPart [nSteps]]`, argv rules of the reference CLI
// (verN/main.cpp:25-46): argv[1] overrides the particle count; the step count is
// overridden only when exactly two arguments are given (`argc == 3`), so a third extra
// argument silently leaves nSteps at its default, as in ver0-8.
#include <cstdlib>

#include "GSimulation.hpp"

int main(int argc, char **argv)
{
    GSimulation sim;
    if (argc > 1) {
        sim.set_number_of_particles(std::atoi(argv[1]));
        if (argc == 3) sim.set_number_of_steps(std::atoi(argv[2]));
    }
    sim.start();
    return 0;
}

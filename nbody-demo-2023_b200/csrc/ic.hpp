// ic.hpp -- host-side initial conditions, shared by libnbx (nbx_ic_*) and the
// GSimulation front end (init_pos / init_vel / init_mass).
//
// The uniform-cube generator makes the SAME library calls, in the same order, as the
// reference (ver0/GSimulation.cpp:44-93): three independent std::mt19937 engines
// seeded with 42, each read through std::uniform_real_distribution<float>; positions
// U(0,1) drawn x,y,z per particle, velocities U(-1,1)*1e-3f, masses n*U(0,1).
// uniform_real_distribution is implementation-defined, so calling libstdc++ (rather
// than re-deriving the bit pattern) is what keeps the build a drop-in.
#pragma once

#include <cmath>
#include <random>

namespace nbx_ic {

inline void uniform_pos(int n, float *px, float *py, float *pz)
{
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> unif(0.0f, 1.0f);
    for (int i = 0; i < n; ++i) {
        px[i] = unif(gen);
        py[i] = unif(gen);
        pz[i] = unif(gen);
    }
}

inline void uniform_vel(int n, float *vx, float *vy, float *vz)
{
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> unif(-1.0f, 1.0f);
    for (int i = 0; i < n; ++i) {
        vx[i] = unif(gen) * 1.0e-3f;
        vy[i] = unif(gen) * 1.0e-3f;
        vz[i] = unif(gen) * 1.0e-3f;
    }
}

inline void uniform_mass(int n, float *mass)
{
    const float nf = static_cast<float>(n);
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> unif(0.0f, 1.0f);
    for (int i = 0; i < n; ++i) mass[i] = nf * unif(gen);
}

// Plummer sphere (BASELINE config 3; not in the reference, which only has the cube):
// cumulative mass M(r) = r^3/(r^2+a^2)^(3/2)  =>  r = a / sqrt(u^(-2/3) - 1), a = 1,
// redrawn while r > 10a; direction isotropic (cos(theta) uniform, phi uniform).
inline void plummer_pos(int n, float *px, float *py, float *pz)
{
    std::mt19937_64 gen(20231);
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < n; ++i) {
        double r;
        do {
            double u = unif(gen);
            if (u < 1e-12) u = 1e-12;
            r = 1.0 / std::sqrt(std::pow(u, -2.0 / 3.0) - 1.0);
        } while (!(r <= 10.0));
        const double ct = 2.0 * unif(gen) - 1.0;
        const double st = std::sqrt(1.0 - ct * ct);
        const double ph = two_pi * unif(gen);
        px[i] = static_cast<float>(r * st * std::cos(ph));
        py[i] = static_cast<float>(r * st * std::sin(ph));
        pz[i] = static_cast<float>(r * ct);
    }
}

}  // namespace nbx_ic

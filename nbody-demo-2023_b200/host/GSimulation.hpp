// GSimulation.hpp -- the reference's simulation-object surface, kept so this build is a
// drop-in for any verN binary (ver3/GSimulation.hpp:36-80): GSimulation(), ~GSimulation(),
// set_number_of_particles(int), set_number_of_steps(int), start().  (`init()` is declared
// by the reference at :42 but defined in no version; it is declared here too and is a
// no-op.)  start() is the B200 backend: it talks to the GPU only through include/nbx.h.
#ifndef _GSIMULATION_HPP
#define _GSIMULATION_HPP

#include <string>
#include <vector>

typedef float real_type;   // verN/types.hpp:21

struct ParticleSoA {       // host-side SoA, as ver3/Particle.hpp:43-58 (acc_* live on the GPU only)
    std::vector<real_type> pos_x, pos_y, pos_z;
    std::vector<real_type> vel_x, vel_y, vel_z;
    std::vector<real_type> mass;
};

class GSimulation {
public:
    GSimulation();
    ~GSimulation();

    void init();
    void set_number_of_particles(int N);
    void set_number_of_steps(int N);
    void start();

    // ver5_all surface (ver5_all/GSimulation.hpp:51-58), used by nbody_all.x
    void set_devices(int N) { _devices = N; }          // 1 cpu, 2 gpu, 3 cpu+gpu
    int get_devices() const { return _devices; }
    void set_cpu_ratio(const float &r) { _cpu_ratio = r; }
    void set_thread_dim0(const int &d) { _thread_dim0 = d; }
    void set_thread_dim1(const int &d) { _thread_dim1 = d; }
    static void set_banner(bool on) { s_banner = on; } // ver5_all prints the banner from main()

    // extensions (not in the reference; defaults keep `./nbody.x N S` identical)
    void set_number_of_gpus(int G) { _ngpus = G; }
    void set_sample_frequency(int sf) { if (sf > 0) _sfreq = sf; }
    real_type kenergy() const { return _kenergy; }

private:
    ParticleSoA *particles;

    int _npart;        // number of particles
    int _nsteps;       // number of integration steps
    real_type _tstep;  // time step of the simulation
    int _sfreq;        // sample frequency
    real_type _kenergy;  // kinetic energy
    double _totTime;   // total time of the simulation
    double _totFlops;  // total number of flops
    int _ngpus;
    std::string _ic;   // "uniform" (reference) or "plummer"
    int _devices = 0, _thread_dim0 = 0, _thread_dim1 = 0;
    float _cpu_ratio = -1.0f;
    static bool s_banner;

    void init_pos();
    void init_vel();
    void init_acc();
    void init_mass();

    inline void set_npart(const int &N) { _npart = N; }
    inline int get_npart() const { return _npart; }
    inline void set_tstep(const real_type &dt) { _tstep = dt; }
    inline real_type get_tstep() const { return _tstep; }
    inline void set_nsteps(const int &n) { _nsteps = n; }
    inline int get_nsteps() const { return _nsteps; }
    inline void set_sfreq(const int &sf) { _sfreq = sf; }
    inline int get_sfreq() const { return _sfreq; }

    void print_header();
};

#endif

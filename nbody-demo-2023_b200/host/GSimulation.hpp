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
    void set_sample_frequency(int sf) { if (sf > 0) cfg.sfreq = sf; }
    real_type kenergy() const { return _kenergy; }

private:
    // State of one run.  (The reference keeps the same quantities -- verN/GSimulation.hpp:48-59 --
    // behind per-field inline accessors; only the five public calls above are its interface.)
    struct Config {
        int npart = 2000;          // ver0/GSimulation.cpp:28
        int nsteps = 500;          // :29
        real_type tstep = 0.1f;    // :30
        int sfreq = 50;            // :31, rows every sfreq steps
    } cfg;
    ParticleSoA *particles = nullptr;
    real_type _kenergy = 0;
    double _totTime = 0, _totFlops = 0;
    int _ngpus = 1;
    std::string _ic = "uniform";   // or "plummer"
    int _devices = 0, _thread_dim0 = 0, _thread_dim1 = 0;
    float _cpu_ratio = -1.0f;
    static bool s_banner;

    void init_pos();
    void init_vel();
    void init_acc();
    void init_mass();
    void print_header();
};

#endif

// cpu_time.hpp -- wall clock for the step-loop timing columns.
// Same role as the reference's CPUTime (verN/cpu_time.hpp:30-48: gettimeofday-based
// start()/stop() returning seconds since the epoch as double); here on std::chrono.
#ifndef NBX_CPU_TIME_HPP
#define NBX_CPU_TIME_HPP

#include <chrono>

class CPUTime {
public:
    CPUTime() {}
    inline double start() { return now(); }
    inline double stop() { return now(); }

private:
    static inline double now()
    {
        using namespace std::chrono;
        return duration<double>(system_clock::now().time_since_epoch()).count();
    }
};

#endif

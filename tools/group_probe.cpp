// tools/group_probe.cpp -- one-process multi-GPU bring-up probe through the C ABI: create / attach_group (multicast) /
// upload_group / run_group / destroy, printing before each call; a SIGSEGV handler prints a backtrace.
//   g++ -std=c++17 -g -O1 -rdynamic -Iinclude tools/group_probe.cpp -Lnbody-demo-2023_b200 -lnbx -Wl,-rpath,$PWD/nbody-demo-2023_b200 -o tools/group_probe
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "nbx.h"

static void on_segv(int sig)
{
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "\n*** SIGSEGV, backtrace:\n";
    write(2, msg, sizeof msg - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}
#define SAY(...) do { std::printf(__VA_ARGS__); std::printf("\n"); std::fflush(stdout); } while (0)
#define OK(call) do { SAY("%s", #call); int rc_ = (call); if (rc_) { SAY("  -> %d %s", rc_, nbx_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    signal(SIGSEGV, on_segv);
    const int G = argc > 1 ? std::atoi(argv[1]) : 2, n = argc > 2 ? std::atoi(argv[2]) : 40960, mc = argc > 3 ? std::atoi(argv[3]) : -1;
    std::vector<float> a[7];
    for (auto &v : a) v.resize(n);
    nbx_ic_uniform(n, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data());
    std::vector<nbx_ctx *> ctx(G, nullptr);
    for (int g = 0; g < G; ++g) {
        OK(nbx_create(&ctx[g], n, g, g, G, 0.1f, 6.67259e-11f, 1e-3f));
        OK(nbx_set_option(ctx[g], "multicast", mc));
    }
    OK(nbx_p2p_attach_group(ctx.data(), G));
    nbx_info info;
    nbx_get_info(ctx[0], &info);
    SAY("multicast active: %d", info.multicast);
    OK(nbx_upload_group(ctx.data(), G, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(), a[6].data()));
    std::vector<double> ke(8);
    double secs = 0;
    OK(nbx_run_group(ctx.data(), G, 5, ke.data(), &secs));
    SAY("ke[4] = %.9g  (%.3f ms)", ke[4], secs * 1e3);
    for (int g = 0; g < G; ++g) { SAY("nbx_destroy(ctx[%d])", g); nbx_destroy(ctx[g]); }
    SAY("done");
    return 0;
}

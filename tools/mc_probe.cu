// tools/mc_probe.cu -- NVSwitch multicast bring-up probe (2+ GPUs, one process): creates a multicast team
// with nbx_multicast.hpp, stores through the multicast mapping from GPU 0 with multimem.st, checks every
// GPU's copy, and releases everything step by step with a line printed before each driver call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mc_probe tools/mc_probe.cu -ldl && tools/mc_probe [ngpus]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../nbody-demo-2023_b200/csrc/nbx_multicast.hpp"

__global__ void mc_store(float4 *mc, int n, float base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = make_float4(base + i, 1.f, 2.f, 3.f);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#define SAY(...) do { std::printf(__VA_ARGS__); std::printf("\n"); std::fflush(stdout); } while (0)

int main(int argc, char **argv)
{
    int G = argc > 1 ? std::atoi(argv[1]) : 2, ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < G) { SAY("need %d GPUs, have %d", G, ndev); return 0; }
    std::vector<int> devs(G);
    for (int g = 0; g < G; ++g) devs[g] = g;
    const int n = 1 << 16;
    std::vector<nbx_mc::Buffer> bufs;
    SAY("create_team...");
    std::string why = nbx_mc::create_team(devs, (size_t)n * sizeof(float4), bufs);
    if (!why.empty()) { SAY("multicast unavailable: %s", why.c_str()); return 0; }
    SAY("team ok: size %zu, uc0 %p mc0 %p", bufs[0].size, (void *)bufs[0].uc, (void *)bufs[0].mcva);
    for (int g = 0; g < G; ++g) { cudaSetDevice(g); cudaMemset((void *)bufs[g].uc, 0, bufs[g].size); cudaDeviceSynchronize(); }
    cudaSetDevice(0);
    mc_store<<<n / 256, 256>>>((float4 *)bufs[0].mcva, n, 100.f);
    SAY("store: %s", cudaGetErrorString(cudaDeviceSynchronize()));
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(g);
        std::vector<float4> h(n);
        cudaError_t e = cudaMemcpy(h.data(), (void *)bufs[g].uc, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < n; ++i) bad += !(h[i].x == 100.f + i && h[i].w == 3.f);
        SAY("gpu %d copy: %s, %d wrong of %d", g, cudaGetErrorString(e), bad, n);
    }
    if (argc > 2) {          // release orders with TWO teams alive (the library keeps one per ping-pong replica)
        std::vector<nbx_mc::Buffer> bufs2;
        why = nbx_mc::create_team(devs, (size_t)n * sizeof(float4), bufs2);
        SAY("second team: %s", why.empty() ? "ok" : why.c_str());
        cudaSetDevice(1);
        mc_store<<<n / 256, 256>>>((float4 *)bufs2[1].mcva, n, 7.f);
        SAY("store from gpu 1 into team 2: %s", cudaGetErrorString(cudaDeviceSynchronize()));
        const std::string mode = argv[2];
        if (mode == "ctx") {           // context by context: member g of both teams, g ascending (what nbx_destroy did)
            for (int g = 0; g < G; ++g) { SAY("release(gpu %d, team 1)", g); nbx_mc::release(bufs[g]); SAY("release(gpu %d, team 2)", g); nbx_mc::release(bufs2[g]); }
        } else {                       // team by team
            for (int g = 0; g < G; ++g) { SAY("release(gpu %d, team 1)", g); nbx_mc::release(bufs[g]); }
            for (int g = 0; g < G; ++g) { SAY("release(gpu %d, team 2)", g); nbx_mc::release(bufs2[g]); }
        }
        SAY("done");
        return 0;
    }
    nbx_mc::Api &a = nbx_mc::api();
    for (int g = G - 1; g >= 0; --g) {
        nbx_mc::Buffer &b = bufs[g];
        cudaSetDevice(b.device);
        SAY("gpu %d: unmap mc", g);      SAY("  -> %d", (int)a.MemUnmap(b.mcva, b.size));
        SAY("gpu %d: free mc va", g);    SAY("  -> %d", (int)a.MemAddressFree(b.mcva, b.size));
        CUdevice dev; a.DeviceGet(&dev, b.device);
        SAY("gpu %d: unbind", g);        SAY("  -> %d", (int)a.MulticastUnbind(b.mc, dev, 0, b.size));
        SAY("gpu %d: unmap uc", g);      SAY("  -> %d", (int)a.MemUnmap(b.uc, b.size));
        SAY("gpu %d: free uc va", g);    SAY("  -> %d", (int)a.MemAddressFree(b.uc, b.size));
        SAY("gpu %d: release mem", g);   SAY("  -> %d", (int)a.MemRelease(b.mem));
    }
    SAY("release mc object");            SAY("  -> %d", (int)a.MemRelease(bufs[0].mc));
    SAY("done");
    return 0;
}

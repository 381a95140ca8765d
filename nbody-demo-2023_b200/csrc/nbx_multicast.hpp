// nbx_multicast.hpp -- NVSwitch multicast (NVLS) mapping of the position replicas.
//
// With the P2P exchange the step kernel's epilogue stores every updated record into each peer's
// replica: world-1 NVLink stores per record.  On an NVSwitch system the switch can replicate a
// single store to every GPU of a multicast team: the replica is backed by one physical allocation
// per GPU (cuMemCreate), all of them bound to one multicast object (cuMulticastCreate /
// cuMulticastBindMem), and the object is mapped a second time into each GPU's address space; a
// `multimem.st` to that mapping lands in every GPU's copy (on sm_100a ptxas lowers multimem.st to an
// ordinary STG.E.128.STRONG.SYS -- the replication is a property of the mapping, not of the opcode).  This file owns the driver-API plumbing (libcuda is dlopen'ed, so a
// single-GPU user never touches it); the kernel side is two instructions in nbx_kernels.cuh.
//
// The reference has nothing comparable: its multi-device exchange is MPI_Bcast of nine arrays per
// step through host memory (ver5_all/GSimulation.cpp:170-189).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace nbx_mc {

struct Api {
    void *lib = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
};

inline Api &api()
{
    static Api a;
    return a;
}

// Empty string on success, else why the driver API is unusable here.
inline std::string load()
{
    Api &a = api();
    if (a.lib) return "";
    void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return std::string("cannot dlopen libcuda.so.1: ") + dlerror();
    struct Sym { void **slot; const char *name; };
    const Sym syms[] = {
        {(void **)&a.GetErrorString, "cuGetErrorString"},
        {(void **)&a.DeviceGet, "cuDeviceGet"},
        {(void **)&a.DeviceGetAttribute, "cuDeviceGetAttribute"},
        {(void **)&a.MulticastCreate, "cuMulticastCreate"},
        {(void **)&a.MulticastAddDevice, "cuMulticastAddDevice"},
        {(void **)&a.MulticastBindMem, "cuMulticastBindMem"},
        {(void **)&a.MulticastUnbind, "cuMulticastUnbind"},
        {(void **)&a.MulticastGetGranularity, "cuMulticastGetGranularity"},
        {(void **)&a.MemCreate, "cuMemCreate"},
        {(void **)&a.MemRelease, "cuMemRelease"},
        {(void **)&a.MemAddressReserve, "cuMemAddressReserve"},
        {(void **)&a.MemAddressFree, "cuMemAddressFree"},
        {(void **)&a.MemMap, "cuMemMap"},
        {(void **)&a.MemUnmap, "cuMemUnmap"},
        {(void **)&a.MemSetAccess, "cuMemSetAccess"},
        {(void **)&a.MemGetAllocationGranularity, "cuMemGetAllocationGranularity"},
        {(void **)&a.MemExportToShareableHandle, "cuMemExportToShareableHandle"},
        {(void **)&a.MemImportFromShareableHandle, "cuMemImportFromShareableHandle"},
    };
    for (const Sym &s : syms) {
        *s.slot = dlsym(h, s.name);
        if (!*s.slot) return std::string("libcuda lacks ") + s.name;
    }
    a.lib = h;
    return "";
}

inline std::string err(const char *what, CUresult r)
{
    const char *s = nullptr;
    if (api().GetErrorString) api().GetErrorString(r, &s);
    return std::string(what) + " -> " + (s ? s : "unknown CUDA driver error");
}

// One GPU's share of one multicast-mapped replica.
struct Buffer {
    CUmemGenericAllocationHandle mem = 0;   // this GPU's physical memory
    CUmemGenericAllocationHandle mc = 0;    // the team's multicast object (same value on every member of a process)
    CUdeviceptr uc = 0;                     // ordinary (unicast) mapping of `mem`: what the kernels read
    CUdeviceptr mcva = 0;                   // mapping of the multicast object: multimem.st target
    size_t size = 0, gran = 0;
    int device = -1;
    bool bound = false;
    int *team_refs = nullptr;               // members still holding `mc`; the last one to leave releases the object
};

inline void release(Buffer &b)
{
    Api &a = api();
    if (!a.lib) return;
    if (b.device >= 0) cudaSetDevice(b.device);
    if (b.mcva) { a.MemUnmap(b.mcva, b.size); a.MemAddressFree(b.mcva, b.size); b.mcva = 0; }
    if (b.bound) {
        CUdevice dev;
        if (a.DeviceGet(&dev, b.device) == CUDA_SUCCESS) a.MulticastUnbind(b.mc, dev, 0, b.size);
        b.bound = false;
    }
    if (b.uc) { a.MemUnmap(b.uc, b.size); a.MemAddressFree(b.uc, b.size); b.uc = 0; }
    if (b.mem) { a.MemRelease(b.mem); b.mem = 0; }
    // the multicast object outlives every member's binding and mapping: the last member out frees it.
    // (Never hand these addresses to legacy CUDA IPC: cudaIpcGetMemHandle on a multicast-bound mapping made
    // a later cuMulticastUnbind crash inside the driver -- nbx_p2p_export skips it for such replicas.)
    if (b.mc && b.team_refs && --*b.team_refs == 0) {
        a.MemRelease(b.mc);
        delete b.team_refs;
    }
    b.mc = 0;
    b.team_refs = nullptr;
}

// Team set-up inside ONE process: `devices` = CUDA ordinals of the members, in rank order.
// On success bufs[g] holds member g's mappings of a `bytes`-sized replica (contents undefined).
// Returns "" or the reason multicast is unavailable (everything already created is released).
inline std::string create_team(const std::vector<int> &devices, size_t bytes, std::vector<Buffer> &bufs)
{
    std::string why = load();
    if (!why.empty()) return why;
    Api &a = api();
    const int G = (int)devices.size();
    bufs.assign((size_t)G, Buffer());
    std::vector<CUdevice> cudev((size_t)G);
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(devices[g]);
        cudaFree(0);                                   // make sure the primary context exists
        CUresult r = a.DeviceGet(&cudev[g], devices[g]);
        if (r != CUDA_SUCCESS) return err("cuDeviceGet", r);
        int ok = 0;
        r = a.DeviceGetAttribute(&ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev[g]);
        if (r != CUDA_SUCCESS || !ok) return "device " + std::to_string(devices[g]) + " does not support NVSwitch multicast";
    }
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)G;
    mp.handleTypes = 0;
    mp.flags = 0;
    mp.size = bytes;
    size_t gran = 0;
    CUresult r = a.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran == 0) return err("cuMulticastGetGranularity", r);
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = devices[0];
    size_t mgran = 0;
    r = a.MemGetAllocationGranularity(&mgran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || mgran == 0) return err("cuMemGetAllocationGranularity", r);
    if (mgran > gran) gran = (mgran + gran - 1) / gran * gran;
    const size_t size = (bytes + gran - 1) / gran * gran;
    mp.size = size;

    auto fail_all = [&](const std::string &w) {
        for (Buffer &b : bufs) release(b);
        return w;
    };
    CUmemGenericAllocationHandle mc = 0;
    r = a.MulticastCreate(&mc, &mp);
    if (r != CUDA_SUCCESS) return err("cuMulticastCreate", r);
    int *refs = new int(G);
    for (int g = 0; g < G; ++g) {
        bufs[g].mc = mc; bufs[g].size = size; bufs[g].device = devices[g]; bufs[g].team_refs = refs;
    }
    for (int g = 0; g < G; ++g)                        // every member joins before anyone binds
        if ((r = a.MulticastAddDevice(mc, cudev[g])) != CUDA_SUCCESS) return fail_all(err("cuMulticastAddDevice", r));
    std::vector<CUmemAccessDesc> everyone((size_t)G);
    for (int g = 0; g < G; ++g) {
        everyone[g].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        everyone[g].location.id = devices[g];
        everyone[g].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    for (int g = 0; g < G; ++g) {
        Buffer &b = bufs[g];
        cudaSetDevice(devices[g]);
        ap.location.id = devices[g];
        if ((r = a.MemCreate(&b.mem, size, &ap, 0)) != CUDA_SUCCESS) return fail_all(err("cuMemCreate", r));
        if ((r = a.MulticastBindMem(mc, 0, b.mem, 0, size, 0)) != CUDA_SUCCESS) return fail_all(err("cuMulticastBindMem", r));
        b.bound = true;
        if ((r = a.MemAddressReserve(&b.uc, size, gran, 0, 0)) != CUDA_SUCCESS) return fail_all(err("cuMemAddressReserve", r));
        if ((r = a.MemMap(b.uc, size, 0, b.mem, 0)) != CUDA_SUCCESS) return fail_all(err("cuMemMap(unicast)", r));
        // every member may also address this copy directly (the unicast P2P path stays usable)
        if ((r = a.MemSetAccess(b.uc, size, everyone.data(), (size_t)G)) != CUDA_SUCCESS) return fail_all(err("cuMemSetAccess(unicast)", r));
    }
    for (int g = 0; g < G; ++g) {
        Buffer &b = bufs[g];
        cudaSetDevice(devices[g]);
        if ((r = a.MemAddressReserve(&b.mcva, size, gran, 0, 0)) != CUDA_SUCCESS) return fail_all(err("cuMemAddressReserve(mc)", r));
        if ((r = a.MemMap(b.mcva, size, 0, mc, 0)) != CUDA_SUCCESS) return fail_all(err("cuMemMap(multicast)", r));
        if ((r = a.MemSetAccess(b.mcva, size, &everyone[g], 1)) != CUDA_SUCCESS) return fail_all(err("cuMemSetAccess(multicast)", r));
    }
    return "";
}

// ------------------------------------------------------------------------------------------------
//  The same team across PROCESSES (one process per GPU): rank 0 creates the multicast objects with an
//  exportable POSIX-fd handle type and hands the file descriptors to every peer over a Unix-domain socket
//  (SCM_RIGHTS -- a file descriptor cannot travel through NCCL or torch.distributed); each process
//  imports them, adds its own device, creates exportable physical memory, binds and maps.  The phases are
//  separate functions because the caller (nbx_p2p_attach) takes a consensus between them: either every
//  rank ends up on the multicast path or none does.
// ------------------------------------------------------------------------------------------------
inline size_t team_size(int device, int world, size_t bytes, size_t *gran_out, std::string *why)
{
    Api &a = api();
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)world;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mp.size = bytes;
    size_t gran = 0, mgran = 0;
    CUresult r = a.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran == 0) { *why = err("cuMulticastGetGranularity", r); return 0; }
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    r = a.MemGetAllocationGranularity(&mgran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || mgran == 0) { *why = err("cuMemGetAllocationGranularity", r); return 0; }
    if (mgran > gran) gran = (mgran + gran - 1) / gran * gran;
    *gran_out = gran;
    return (bytes + gran - 1) / gran * gran;
}

inline void abstract_addr(const char *name, sockaddr_un *sa, socklen_t *len)
{
    std::memset(sa, 0, sizeof *sa);
    sa->sun_family = AF_UNIX;
    const size_t n = std::min(std::strlen(name), sizeof(sa->sun_path) - 2);
    std::memcpy(sa->sun_path + 1, name, n);              // leading NUL: abstract namespace, nothing to unlink
    *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

inline bool send_fds(int sock, const int *fds, int count)
{
    char payload = 'F';
    iovec iov = {&payload, 1};
    char ctrl[CMSG_SPACE(sizeof(int) * 4)] = {};
    msghdr msg = {};
    msg.msg_iov = &iov; msg.msg_iovlen = 1;
    msg.msg_control = ctrl; msg.msg_controllen = CMSG_SPACE(sizeof(int) * count);
    cmsghdr *cm = CMSG_FIRSTHDR(&msg);
    cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int) * count);
    std::memcpy(CMSG_DATA(cm), fds, sizeof(int) * count);
    return sendmsg(sock, &msg, MSG_NOSIGNAL) == 1;      // a peer that went away is an error return, not a SIGPIPE
}

inline bool recv_fds(int sock, int *fds, int count, int timeout_ms)
{
    pollfd p = {sock, POLLIN, 0};
    if (poll(&p, 1, timeout_ms) <= 0) return false;
    char payload = 0;
    iovec iov = {&payload, 1};
    char ctrl[CMSG_SPACE(sizeof(int) * 4)] = {};
    msghdr msg = {};
    msg.msg_iov = &iov; msg.msg_iovlen = 1;
    msg.msg_control = ctrl; msg.msg_controllen = sizeof ctrl;
    if (recvmsg(sock, &msg, 0) != 1) return false;
    cmsghdr *cm = CMSG_FIRSTHDR(&msg);
    if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS || cm->cmsg_len != CMSG_LEN(sizeof(int) * count)) return false;
    std::memcpy(fds, CMSG_DATA(cm), sizeof(int) * count);
    return true;
}

// Phase 1.  Rank 0: create `nbuf` multicast objects, serve their descriptors to world-1 peers on the abstract
// socket `name`.  Other ranks: connect (retrying while rank 0 gets there), receive, import.  On success
// bufs[b].mc / .size / .device are set and bufs[b].team_refs = 1 (each process releases its own handle).
inline std::string mp_open_team(int rank, int world, int device, size_t bytes, const char *name, Buffer *bufs, int nbuf,
                                int timeout_ms = 20000)
{
    std::string why = load();
    if (!why.empty()) return why;
    Api &a = api();
    cudaSetDevice(device);
    cudaFree(0);
    CUdevice dev;
    CUresult r = a.DeviceGet(&dev, device);
    if (r != CUDA_SUCCESS) return err("cuDeviceGet", r);
    int ok = 0;
    r = a.DeviceGetAttribute(&ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    if (r != CUDA_SUCCESS || !ok) return "device does not support NVSwitch multicast";
    size_t gran = 0;
    const size_t size = team_size(device, world, bytes, &gran, &why);
    if (!size) return why;
    sockaddr_un sa; socklen_t salen;
    abstract_addr(name, &sa, &salen);
    int fds[4] = {-1, -1, -1, -1};
    if (rank == 0) {
        CUmulticastObjectProp mp = {};
        mp.numDevices = (unsigned)world;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        mp.size = size;
        auto close_fds = [&]() { for (int b = 0; b < nbuf; ++b) if (fds[b] >= 0) { close(fds[b]); fds[b] = -1; } };
        for (int b = 0; b < nbuf; ++b) {
            if ((r = a.MulticastCreate(&bufs[b].mc, &mp)) != CUDA_SUCCESS) { close_fds(); return err("cuMulticastCreate(exportable)", r); }
            bufs[b].team_refs = new int(1);
            if ((r = a.MemExportToShareableHandle(&fds[b], bufs[b].mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)) != CUDA_SUCCESS) {
                close_fds();
                return err("cuMemExportToShareableHandle", r);
            }
        }
        const int ls = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        bool good = ls >= 0 && bind(ls, (sockaddr *)&sa, salen) == 0 && listen(ls, world) == 0;
        for (int peer = 1; peer < world && good; ++peer) {
            pollfd p = {ls, POLLIN, 0};
            if (poll(&p, 1, timeout_ms) <= 0) { good = false; break; }
            const int cs = accept(ls, nullptr, nullptr);
            good = cs >= 0 && send_fds(cs, fds, nbuf);
            if (cs >= 0) close(cs);
        }
        if (ls >= 0) close(ls);
        close_fds();
        if (!good) return "could not hand the multicast descriptors to every peer (Unix socket)";
    } else {
        int cs = -1;
        for (int waited = 0; waited < timeout_ms; waited += 20) {           // rank 0 may not be listening yet
            cs = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
            if (cs >= 0 && connect(cs, (sockaddr *)&sa, salen) == 0) break;
            if (cs >= 0) close(cs);
            cs = -1;
            usleep(20000);
        }
        if (cs < 0) return "could not reach rank 0's multicast socket";
        const bool got = recv_fds(cs, fds, nbuf, timeout_ms);
        close(cs);
        if (!got) return "did not receive the multicast descriptors";
        std::string bad;
        for (int b = 0; b < nbuf; ++b) {
            r = bad.empty() ? a.MemImportFromShareableHandle(&bufs[b].mc, (void *)(uintptr_t)fds[b], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR)
                            : CUDA_SUCCESS;
            close(fds[b]);                                   // every received descriptor is closed, also after a failure
            if (!bad.empty()) continue;
            if (r != CUDA_SUCCESS) { bad = err("cuMemImportFromShareableHandle", r); continue; }
            bufs[b].team_refs = new int(1);
        }
        if (!bad.empty()) return bad;
    }
    for (int b = 0; b < nbuf; ++b) { bufs[b].size = size; bufs[b].gran = gran; bufs[b].device = device; }
    return "";
}

// Phase 2: this process's GPU joins every object (all members must have joined before anyone binds).
inline std::string mp_join(Buffer *bufs, int nbuf)
{
    Api &a = api();
    CUdevice dev;
    CUresult r = a.DeviceGet(&dev, bufs[0].device);
    if (r != CUDA_SUCCESS) return err("cuDeviceGet", r);
    for (int b = 0; b < nbuf; ++b)
        if ((r = a.MulticastAddDevice(bufs[b].mc, dev)) != CUDA_SUCCESS) return err("cuMulticastAddDevice", r);
    return "";
}

// Phase 3: exportable physical memory on this GPU, bound to the object and mapped twice.
inline std::string mp_bind_and_map(Buffer *bufs, int nbuf)
{
    Api &a = api();
    for (int b = 0; b < nbuf; ++b) {
        Buffer &B = bufs[b];
        cudaSetDevice(B.device);
        const size_t gran = B.gran;
        CUmemAllocationProp ap = {};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = B.device;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;   // imported multicast objects bind only to shareable memory
        CUmemAccessDesc me = {};
        me.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        me.location.id = B.device;
        me.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CUresult r;
        if ((r = a.MemCreate(&B.mem, B.size, &ap, 0)) != CUDA_SUCCESS) return err("cuMemCreate(exportable)", r);
        if ((r = a.MulticastBindMem(B.mc, 0, B.mem, 0, B.size, 0)) != CUDA_SUCCESS) return err("cuMulticastBindMem", r);
        B.bound = true;
        if ((r = a.MemAddressReserve(&B.uc, B.size, gran, 0, 0)) != CUDA_SUCCESS) return err("cuMemAddressReserve", r);
        if ((r = a.MemMap(B.uc, B.size, 0, B.mem, 0)) != CUDA_SUCCESS) return err("cuMemMap(unicast)", r);
        if ((r = a.MemSetAccess(B.uc, B.size, &me, 1)) != CUDA_SUCCESS) return err("cuMemSetAccess(unicast)", r);
        if ((r = a.MemAddressReserve(&B.mcva, B.size, gran, 0, 0)) != CUDA_SUCCESS) return err("cuMemAddressReserve(mc)", r);
        if ((r = a.MemMap(B.mcva, B.size, 0, B.mc, 0)) != CUDA_SUCCESS) return err("cuMemMap(multicast)", r);
        if ((r = a.MemSetAccess(B.mcva, B.size, &me, 1)) != CUDA_SUCCESS) return err("cuMemSetAccess(multicast)", r);
    }
    return "";
}

}  // namespace nbx_mc

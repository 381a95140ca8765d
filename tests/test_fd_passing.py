"""CPU suite: the descriptor hand-over the cross-process NVSwitch multicast set-up relies on (csrc/nbx_multicast.hpp:
abstract Unix socket + SCM_RIGHTS), exercised without a GPU: a parent "rank 0" listens on an abstract socket and sends
two descriptors, a forked "rank 1" connects (retrying, as in mp_open_team), receives them and reads through them."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "nbx_multicast.hpp"
#include <sys/wait.h>
#include <fcntl.h>
int main()
{
    char name[64];
    std::snprintf(name, sizeof name, "nbx-mc-test-%d", (int)getpid());
    sockaddr_un sa; socklen_t salen;
    nbx_mc::abstract_addr(name, &sa, &salen);
    int p1[2], p2[2];
    if (pipe(p1) || pipe(p2)) return 2;
    if (write(p1[1], "alpha", 5) != 5 || write(p2[1], "omega", 5) != 5) return 3;
    pid_t child = fork();
    if (child == 0) {                      // "rank 1": connect with retries, receive, read through the descriptors
        int cs = -1;
        for (int waited = 0; waited < 5000; waited += 20) {
            cs = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
            if (cs >= 0 && connect(cs, (sockaddr *)&sa, salen) == 0) break;
            if (cs >= 0) close(cs);
            cs = -1;
            usleep(20000);
        }
        if (cs < 0) _exit(10);
        int fds[2] = {-1, -1};
        if (!nbx_mc::recv_fds(cs, fds, 2, 5000)) _exit(11);
        char a[6] = {0}, b[6] = {0};
        if (read(fds[0], a, 5) != 5 || read(fds[1], b, 5) != 5) _exit(12);
        _exit(std::strcmp(a, "alpha") == 0 && std::strcmp(b, "omega") == 0 ? 0 : 13);
    }
    usleep(100000);                        // the peer is already retrying when the listener appears
    const int ls = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (ls < 0 || bind(ls, (sockaddr *)&sa, salen) || listen(ls, 4)) return 4;
    pollfd p = {ls, POLLIN, 0};
    if (poll(&p, 1, 5000) <= 0) return 5;
    const int cs = accept(ls, nullptr, nullptr);
    const int fds[2] = {p1[0], p2[0]};
    if (cs < 0 || !nbx_mc::send_fds(cs, fds, 2)) return 6;
    close(cs); close(ls);
    int st = 0;
    waitpid(child, &st, 0);
    std::printf("child exit %d\n", WIFEXITED(st) ? WEXITSTATUS(st) : -1);
    return WIFEXITED(st) && WEXITSTATUS(st) == 0 ? 0 : 7;
}
'''


def test_scm_rights_hand_over(tmp_path):
    src = tmp_path / "fdpass.cpp"
    src.write_text(SRC)
    exe = tmp_path / "fdpass"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(REPO, "nbody-demo-2023_b200", "csrc"),
                        "-I", "/usr/local/cuda/include", str(src), "-o", str(exe), "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "child exit 0" in r.stdout

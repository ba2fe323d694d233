// oracle/ref_mpirun.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
//   ref_mpirun -np N <dimacs> [-d N | -q N] [-s] [-r file]
//
// Runs the WHOLE unmodified reference program (/root/reference/NDP-5_6_7.cpp: main, process_queue's
// dispatcher on rank 0, the worker loop on ranks 1..N-1) as N forked processes over
// oracle/shim_fork/mpi.h -- what `mpirun -np N ./NDP-5_6_7 ...` (README.md:188) does on a machine with
// OpenMPI.  The program ends the way it does under mpirun: the rank that prints the report calls
// std::terminate() (NDP:1404,1447); the launcher then stops the other ranks.  Standard output of all
// ranks goes through one pipe; the BFS progress line rank 0 rewrites after every expansion (NDP:930) is
// dropped, everything else is passed on.  The launcher prints its own wall time at the end:
//   [ref_mpirun] np=N wall=S.SSS s first_exit_rank=R status=...
// Deviations from the real thing, all outside the reference source: `main` renamed; the stderr dup2
// (NDP:307-308) diverted; MPI replaced by the fork shim; -march as in oracle/Makefile.
#include <fcntl.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

static inline int ndp_harness_dup2_noop(int, int) { return 0; }
#define dup2(a, b) ndp_harness_dup2_noop(a, b)
#define main ndp_reference_main
#include "NDP-5_6_7.cpp"   // resolved through -I/root/reference ; unmodified
#undef main
#undef dup2

static double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char** argv) {
    if (argc < 4 || std::string(argv[1]) != "-np") {
        fprintf(stderr, "usage: ref_mpirun -np N <dimacs> [-d N | -q N] [-s] [-r file]\n");
        return 2;
    }
    const int np = atoi(argv[2]);
    if (np < 2) { fprintf(stderr, "ref_mpirun: the reference needs a dispatcher and at least one worker (np >= 2)\n"); return 2; }
    std::vector<char*> sub;               // argv of the reference program: argv[0] = its own name (used in file names, NDP:1566)
    static char prog[] = "NDP-5_6_7";
    sub.push_back(prog);
    for (int i = 3; i < argc; ++i) sub.push_back(argv[i]);
    sub.push_back(nullptr);

    std::vector<int> to_worker(np, -1), to_head(np, -1);
    for (int r = 1; r < np; ++r) {
        int sv[2];
        if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { perror("socketpair"); return 1; }
        int big = 4 << 20;
        setsockopt(sv[0], SOL_SOCKET, SO_SNDBUF, &big, sizeof big);
        setsockopt(sv[1], SOL_SOCKET, SO_SNDBUF, &big, sizeof big);
        to_worker[r] = sv[0];
        to_head[r] = sv[1];
    }
    int out_pipe[2];
    if (pipe(out_pipe) != 0) { perror("pipe"); return 1; }
    const double t0 = now_s();
    std::vector<pid_t> pids(np, 0);
    for (int r = 0; r < np; ++r) {
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 1; }
        if (pid == 0) {
            ::close(out_pipe[0]);
            ::dup2(out_pipe[1], STDOUT_FILENO);     // (the macro diversion above is gone here)
            ::close(out_pipe[1]);
            ndp_fork_mpi::World& w = ndp_fork_mpi::world();
            w.rank = r;
            w.size = np;
            if (r == 0) {
                w.fd = to_worker;
                w.pending.resize((size_t)np);
                for (int k = 1; k < np; ++k) ::close(to_head[k]);
            } else {
                w.fd.assign(1, to_head[r]);
                w.pending.resize(1);
                for (int k = 1; k < np; ++k) { ::close(to_worker[k]); if (k != r) ::close(to_head[k]); }
            }
            int rc = ndp_reference_main((int)sub.size() - 1, sub.data());
            fflush(stdout);
            _exit(rc);
        }
        pids[r] = pid;
    }
    for (int r = 1; r < np; ++r) { ::close(to_worker[r]); ::close(to_head[r]); }
    ::close(out_pipe[1]);
    // drain the ranks' output; drop the per-expansion BFS progress rewrites ("\r\033[K  Queue size: ...")
    {
        std::string line;
        char buf[1 << 16];
        ssize_t k;
        auto flush_line = [&](bool newline) {
            // a segment that starts with the erase-line escape is a progress rewrite
            if (line.find("\033[K  Queue size:") == std::string::npos || newline) {
                fwrite(line.data(), 1, line.size(), stdout);
                if (newline) fputc('\n', stdout);
            }
            line.clear();
        };
        while ((k = ::read(out_pipe[0], buf, sizeof buf)) > 0) {
            for (ssize_t i = 0; i < k; ++i) {
                const char c = buf[i];
                if (c == '\n') flush_line(true);
                else if (c == '\r') flush_line(false);
                else line.push_back(c);
            }
        }
        if (!line.empty()) flush_line(false);
    }
    // the pipe closes when every rank is gone; normally the first one to leave is the one that reported
    int first_rank = -1, first_status = 0;
    for (;;) {
        int status = 0;
        pid_t p = waitpid(-1, &status, 0);
        if (p <= 0) break;
        if (first_rank < 0) {
            for (int r = 0; r < np; ++r) if (pids[r] == p) first_rank = r;
            first_status = status;
            for (int r = 0; r < np; ++r) if (pids[r] != p) kill(pids[r], SIGKILL);   // what mpirun does when a rank aborts
        }
    }
    const double wall = now_s() - t0;
    printf("\n[ref_mpirun] np=%d wall=%.3f s first_exit_rank=%d status=%s%d\n", np, wall, first_rank,
           WIFSIGNALED(first_status) ? "signal " : "exit ", WIFSIGNALED(first_status) ? WTERMSIG(first_status) : WEXITSTATUS(first_status));
    return 0;
}

// oracle/shim_fork/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A multi-PROCESS stand-in for the 13 MPI entry points the reference names, so that the
// WHOLE unmodified program -- main, process_queue's dispatcher and worker loops included
// (NDP-5_6_7.cpp:1229-1454, 1540-1689) -- runs with N ranks in an image that has no OpenMPI:
// oracle/ref_mpirun.cpp forks the ranks and gives rank 0 one socketpair to every worker.  That star
// is all the reference ever uses: workers talk to rank 0 only (NDP:1408-1439), rank 0 probes
// MPI_ANY_SOURCE (NDP:1292) and answers the sender, and every collective is rooted at 0.
// Wire format of a message: { int tag; int bytes; } + payload.  Matching follows MPI: a receive takes the
// FIRST queued message of the source that carries the wanted tag (or any tag); shorter messages than the
// receive buffer are fine (the reference receives its zero-length TASK_DONE / SOLUTION_FOUND signals into
// a 1-int buffer, NDP:1295,1427,1436).  Collectives use reserved negative tags.
#ifndef NDP_ORACLE_SHIM_FORK_MPI_H
#define NDP_ORACLE_SHIM_FORK_MPI_H
#include <limits.h>   // the reference relies on <mpi.h> to bring PATH_MAX (NDP-5_6_7.cpp:334)
#include <poll.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <deque>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_INT 1
#define MPI_CHAR 2
#define MPI_DOUBLE 3
#define MPI_C_BOOL 4
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_SUCCESS 0

namespace ndp_fork_mpi {
struct Msg { int tag; std::vector<char> data; };
struct World {
    int rank = 0, size = 1;
    std::vector<int> fd;                     // rank 0: fd[r] to worker r (fd[0] unused); worker: fd[0] to rank 0
    std::vector<std::deque<Msg>> pending;    // per peer, in arrival order
};
inline World& world() { static World w; return w; }
enum { TAG_BARRIER = -100, TAG_RELEASE = -101, TAG_BCAST = -102, TAG_GATHER = -103, TAG_REDUCE = -104 };

inline size_t tsize(MPI_Datatype t) { return t == MPI_INT ? sizeof(int) : t == MPI_DOUBLE ? sizeof(double) : 1; }
inline void die(const char* what) { fprintf(stdout, "\n[ref_mpirun rank %d] %s\n", world().rank, what); fflush(stdout); _exit(70); }
inline void write_all(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n) { ssize_t k = ::write(fd, c, n); if (k <= 0) _exit(0); c += k; n -= (size_t)k; }   // peer gone: the job is over
}
inline void read_all(int fd, void* p, size_t n) {
    char* c = (char*)p;
    while (n) { ssize_t k = ::read(fd, c, n); if (k <= 0) _exit(0); c += k; n -= (size_t)k; }
}
inline int peer_fd(int peer) { return world().rank == 0 ? world().fd[peer] : world().fd[0]; }
inline void send_raw(int dest, int tag, const void* buf, size_t bytes) {
    if (world().rank != 0 && dest != 0) die("worker-to-worker message: the reference never sends one");
    int hdr[2] = {tag, (int)bytes};
    std::vector<char> frame(sizeof hdr + bytes);
    memcpy(frame.data(), hdr, sizeof hdr);
    if (bytes) memcpy(frame.data() + sizeof hdr, buf, bytes);
    write_all(peer_fd(dest), frame.data(), frame.size());   // one write per message
}
inline void pull_one(int peer) {   // read one whole message from `peer` into its queue (blocks)
    int hdr[2];
    read_all(peer_fd(peer), hdr, sizeof hdr);
    Msg m;
    m.tag = hdr[0];
    m.data.resize((size_t)hdr[1]);
    if (hdr[1]) read_all(peer_fd(peer), m.data.data(), (size_t)hdr[1]);
    world().pending[peer].push_back(std::move(m));
}
inline bool find_pending(int source, int tag, int& peer, size_t& idx) {
    World& w = world();
    const int lo = source == MPI_ANY_SOURCE ? 0 : source, hi = source == MPI_ANY_SOURCE ? (int)w.pending.size() - 1 : source;
    for (int p = lo; p <= hi; ++p)
        for (size_t i = 0; i < w.pending[p].size(); ++i) {
            const int t = w.pending[p][i].tag;
            if (tag == MPI_ANY_TAG ? t >= 0 : t == tag) { peer = p; idx = i; return true; }
        }
    return false;
}
inline bool poll_peers(int source, int timeout_ms) {   // pull every message that is readable now; true if any arrived
    World& w = world();
    std::vector<pollfd> pf;
    std::vector<int> who;
    if (w.rank == 0) {
        for (int r = 1; r < w.size; ++r)
            if (source == MPI_ANY_SOURCE || source == r) { pf.push_back({w.fd[r], POLLIN, 0}); who.push_back(r); }
    } else { pf.push_back({w.fd[0], POLLIN, 0}); who.push_back(0); }
    if (pf.empty()) return false;
    if (::poll(pf.data(), pf.size(), timeout_ms) <= 0) return false;
    bool any = false;
    for (size_t i = 0; i < pf.size(); ++i) {
        if (pf[i].revents & POLLIN) { pull_one(who[i]); any = true; }
        else if (pf[i].revents & (POLLHUP | POLLERR)) _exit(0);
    }
    return any;
}
inline void recv_raw(int source, int tag, void* buf, size_t cap, MPI_Status* st) {
    int peer; size_t idx;
    while (!find_pending(source, tag, peer, idx)) poll_peers(source, -1);
    Msg& m = world().pending[peer][idx];
    if (m.data.size() > cap) die("message longer than the receive buffer");
    if (!m.data.empty()) memcpy(buf, m.data.data(), m.data.size());
    if (st) { st->MPI_SOURCE = peer; st->MPI_TAG = m.tag; st->MPI_ERROR = 0; }
    world().pending[peer].erase(world().pending[peer].begin() + (long)idx);
}
}  // namespace ndp_fork_mpi

static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = ndp_fork_mpi::world().rank; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = ndp_fork_mpi::world().size; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { fflush(stdout); _exit(code ? code : 1); return 0; }
static inline int MPI_Get_processor_name(char* name, int* len) {
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) strcpy(name, "localhost");
    name[MPI_MAX_PROCESSOR_NAME - 1] = 0; *len = (int)strlen(name); return MPI_SUCCESS;
}
static inline int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm) {
    ndp_fork_mpi::send_raw(dest, tag, buf, (size_t)count * ndp_fork_mpi::tsize(t)); return MPI_SUCCESS;
}
static inline int MPI_Recv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Status* st) {
    ndp_fork_mpi::recv_raw(source, tag, buf, (size_t)count * ndp_fork_mpi::tsize(t), st); return MPI_SUCCESS;
}
static inline int MPI_Iprobe(int source, int tag, MPI_Comm, int* flag, MPI_Status* st) {
    using namespace ndp_fork_mpi;
    int peer; size_t idx;
    if (!find_pending(source, tag, peer, idx)) { poll_peers(source, 0); if (!find_pending(source, tag, peer, idx)) { *flag = 0; return MPI_SUCCESS; } }
    *flag = 1;
    if (st) { st->MPI_SOURCE = peer; st->MPI_TAG = world().pending[peer][idx].tag; st->MPI_ERROR = 0; }
    return MPI_SUCCESS;
}
static inline int MPI_Barrier(MPI_Comm) {
    using namespace ndp_fork_mpi;
    World& w = world();
    if (w.size == 1) return MPI_SUCCESS;
    if (w.rank == 0) {
        for (int r = 1; r < w.size; ++r) recv_raw(r, TAG_BARRIER, nullptr, 0, nullptr);
        for (int r = 1; r < w.size; ++r) send_raw(r, TAG_RELEASE, nullptr, 0);
    } else { send_raw(0, TAG_BARRIER, nullptr, 0); recv_raw(0, TAG_RELEASE, nullptr, 0, nullptr); }
    return MPI_SUCCESS;
}
static inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm) {
    using namespace ndp_fork_mpi;
    World& w = world();
    if (root != 0) die("broadcast rooted elsewhere than rank 0");
    const size_t bytes = (size_t)count * tsize(t);
    if (w.rank == 0) { for (int r = 1; r < w.size; ++r) send_raw(r, TAG_BCAST, buf, bytes); }
    else recv_raw(0, TAG_BCAST, buf, bytes, nullptr);
    return MPI_SUCCESS;
}
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int root, MPI_Comm) {
    using namespace ndp_fork_mpi;
    World& w = world();
    if (root != 0) die("gather rooted elsewhere than rank 0");
    const size_t bytes = (size_t)n * tsize(t);
    if (w.rank == 0) {
        memcpy(r, s, bytes);
        for (int k = 1; k < w.size; ++k) recv_raw(k, TAG_GATHER, (char*)r + (size_t)k * bytes, bytes, nullptr);
    } else send_raw(0, TAG_GATHER, s, bytes);
    return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm) {
    using namespace ndp_fork_mpi;
    World& w = world();
    if (root != 0) die("reduce rooted elsewhere than rank 0");
    const size_t bytes = (size_t)n * tsize(t);
    if (w.rank != 0) { send_raw(0, TAG_REDUCE, s, bytes); return MPI_SUCCESS; }
    memcpy(r, s, bytes);
    std::vector<char> tmp(bytes);
    for (int k = 1; k < w.size; ++k) {
        recv_raw(k, TAG_REDUCE, tmp.data(), bytes, nullptr);
        for (int i = 0; i < n; ++i) {
            if (t == MPI_INT) { int* a = (int*)r; const int b = ((int*)tmp.data())[i]; a[i] = op == MPI_SUM ? a[i] + b : (a[i] > b ? a[i] : b); }
            else if (t == MPI_DOUBLE) { double* a = (double*)r; const double b = ((double*)tmp.data())[i]; a[i] = op == MPI_SUM ? a[i] + b : (a[i] > b ? a[i] : b); }
            else die("reduce over an unexpected datatype");
        }
    }
    return MPI_SUCCESS;
}
#endif

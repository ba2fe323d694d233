// oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
// Single-process stand-in for the 13 MPI entry points the reference TU names
// (SURVEY.md §2.1), so that the UNMODIFIED /root/reference/NDP-5_6_7.cpp compiles in an
// image without OpenMPI.  The harness drives the reference's BFS/DFS functions directly
// and never enters process_queue(), so none of the point-to-point stubs is ever reached;
// they abort if they are.
#ifndef NDP_ORACLE_SHIM_MPI_H
#define NDP_ORACLE_SHIM_MPI_H
#include <limits.h>   // the reference relies on <mpi.h> to bring PATH_MAX (NDP-5_6_7.cpp:334)
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

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

static int ndp_shim_world_size = 1;   // harness may set this before calling reference code

static inline size_t ndp_shim_tsize(MPI_Datatype t) {
    return t == MPI_INT ? sizeof(int) : t == MPI_DOUBLE ? sizeof(double) : 1;
}
static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = ndp_shim_world_size; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { _exit(code ? code : 1); return 0; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm) {
    memcpy(r, s, (size_t)n * ndp_shim_tsize(t)); return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    memcpy(r, s, (size_t)n * ndp_shim_tsize(t)); return MPI_SUCCESS;
}
static inline int MPI_Get_processor_name(char* name, int* len) {
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) strcpy(name, "localhost");
    name[MPI_MAX_PROCESSOR_NAME - 1] = 0; *len = (int)strlen(name); return MPI_SUCCESS;
}
static inline int MPI_Iprobe(int, int, MPI_Comm, int* flag, MPI_Status*) { *flag = 0; return MPI_SUCCESS; }
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { abort(); return 0; }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { abort(); return 0; }
#endif

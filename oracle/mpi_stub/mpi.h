/* Single-rank stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * The reference includes <mpi.h> (lattice.hpp:2, ising.hpp:2) and uses exactly seven calls and
 * three constants (ising.cpp:50-53, mcrg.cpp:101-103,275-280, rgnn.cpp:28,41,125-130,252-255,355,
 * main.cpp:6,20).  No MPI exists in this image, so the unmodified reference sources are compiled
 * against this header: one rank, reductions are copies, broadcast is a no-op.
 */
#ifndef MCRG_ORACLE_MPI_STUB_H
#define MCRG_ORACLE_MPI_STUB_H
#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8 /* = sizeof(double), so count*datatype is a byte count */
#define MPI_SUM 0

static inline int MPI_Init(void *, void *) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Reduce(const void *s, void *d, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    memcpy(d, s, (size_t)n * (size_t)t);
    return 0;
}
static inline int MPI_Allreduce(const void *s, void *d, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
    memcpy(d, s, (size_t)n * (size_t)t);
    return 0;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }

#endif

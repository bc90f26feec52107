/* Single-process stand-in for <mpi.h> for programs written against the reference (main.cpp:6,20 call MPI_Init /
 * MPI_Finalize; the reference's parallelism — one Markov chain per rank plus an all-reduce, mcrg.cpp:101-103 — lives
 * inside the device library here: replicas on the GPU, NCCL between GPUs).  If a real MPI is installed, put its
 * include directory first and this file is never seen. */
#ifndef MCRG_B200_MPI_SHIM_H
#define MCRG_B200_MPI_SHIM_H
#include <string.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_SUM 0
static inline int MPI_Init(void *, void *) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Reduce(const void *s, void *d, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { memcpy(d, s, (size_t)n * (size_t)t); return 0; }
static inline int MPI_Allreduce(const void *s, void *d, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { memcpy(d, s, (size_t)n * (size_t)t); return 0; }
#endif

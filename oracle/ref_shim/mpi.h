/* Single-rank MPI stand-in so the reference's gt_mpi_gather.cc links without an MPI install.
 * Test infrastructure only (oracle/_ref build); the combine path performs no collectives. */
#pragma once
#include <string.h>
typedef int MPI_Comm; typedef int MPI_Datatype; typedef long MPI_Aint; typedef long long MPI_Count; typedef int MPI_Info;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_UNSIGNED_CHAR 1
#define MPI_INFO_NULL 0
static inline int MPI_Init(int*, char***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Abort(MPI_Comm, int) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm) { memcpy(r, s, (size_t)n*t); return 0; }
static inline int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int*, MPI_Datatype, int, MPI_Comm) { memcpy(r, s, (size_t)n*t); return 0; }

// TEST INFRASTRUCTURE ONLY.  OpenCilk is absent from this image; the reference uses
// cilk_for only as a parallel-for over independent iterations (src/kmer_set.cpp:124,179),
// so it maps onto an OpenMP parallel for (compile with -fopenmp; without it the pragma is
// ignored and the loop is serial).
#ifndef ORACLE_SHIM_CILK_H
#define ORACLE_SHIM_CILK_H
#define cilk_for _Pragma("omp parallel for schedule(dynamic)") for
#define cilk_spawn
#define cilk_sync
#endif

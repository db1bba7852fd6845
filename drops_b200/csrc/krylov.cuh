// krylov.cuh -- solver drivers shared between krylov.cu and krylov_ext.cu.
#pragma once
#include "common.cuh"

namespace db200 {

// PCG on A (ne = false) or on A A^T (ne = true: PCGNE, num/krylovsolver.h:199-252) with any
// preconditioner kind, including the procedural ChebyshevPcCL.  Contract of drops_b200_pcg_dev.
int pcg_general_device(drops_b200_matrix_s* M, bool ne, double* x, const double* b, int pc, double omega,
                       int* max_iter, double* tol, int relative, double* hist, int hist_cap);
// x = M^-1 b
int precond_apply(drops_b200_matrix_s* M, int pc, double omega, bool ne, const double* b, double* x);

}  // namespace db200

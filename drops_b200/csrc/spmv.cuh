// spmv.cuh -- launch interface of the SpMV kernels (spmv.cu).
#pragma once
#include "common.cuh"

namespace db200 {

// What happens to the row sum s_i before it is stored (all products rounded before the add, as
// in the reference's valarray expressions):
enum SpmvEpilogue {
    EPI_NONE   = 0,   // y_i = s_i                               A*x            spmat.h:1712
    EPI_JACOBI = 1,   // y_i = (omega*s_i)/d_i                   M.Apply(A,w,A*v) krylovsolver.h:377
    EPI_RESID  = 2,   // y_i = b_i - s_i                         b - A*x        krylovsolver.h:125
    EPI_AXMBY  = 3,   // y_i = s_i - beta*z_i  (beta on device)  A*q1 - b0*q0   krylovsolver.h:931
};

struct SpmvArgs {
    const double* x = nullptr;     // input vector (cols entries)
    double* y = nullptr;           // output vector (rows entries)
    int epilogue = EPI_NONE;
    const double* aux = nullptr;   // d (EPI_JACOBI), b (EPI_RESID), z (EPI_AXMBY)
    double omega = 1.0;            // EPI_JACOBI
    const double* beta = nullptr;  // EPI_AXMBY: device scalar (if null, beta_val is used)
    double beta_val = 0.0;
    // optional fused reduction  sum_i y_i * dotv_i  -> *dot_out (device scalar)
    const double* dotv = nullptr;
    double* dot_out = nullptr;
    // partitioned systems (dist.cu): a fused dot may be chained over several launches -- the result
    // (plus *dd_in, if given) is stored unrounded to *dd_out instead of rounded to *dot_out
    const dd* dd_in = nullptr;
    dd* dd_out = nullptr;
    // columns >= n_owned are read from xghost[col - n_owned] (received halo entries)
    const double* xghost = nullptr;
    long long n_owned = 0;
    // x, y, aux and dotv are vectors in the operator's row-ordered numbering (SellOp::pkp, npos entries)
    bool row_ordered = false;
    const int* done = nullptr;     // optional device flag: kernel is a no-op when *done != 0
    int slot = 0;                  // reduction workspace slot (kernels that may overlap use different slots)
    cudaStream_t stream = nullptr; // default: ctx().stream
};

// y = A x through the sliced-ELL operator (exact reference summation order, product rounded before the add).
int spmv_sell(const SellOp& A, const SpmvArgs& a);
// Same operator on its dictionary-packed stream (SellOp::pk, 4 B per entry; spmv_pk.cu); spmv_sell
// dispatches to it when the operator has one and option "spmv.pk" is not 0.
int spmv_sell_pk(const SellOp& A, const SpmvArgs& a);
// y = A x straight from CSR, LANES threads per row (FMA, shuffle tree): comparison kernel.
int spmv_csr_vector(const CsrDev& A, const SpmvArgs& a, int lanes);

// dispatch on option "spmv.kernel": 0 = sell (register-staged), 1 = csr_vector
int spmv_apply(const drops_b200_matrix_s* M, bool transposed, const SpmvArgs& a);

}  // namespace db200

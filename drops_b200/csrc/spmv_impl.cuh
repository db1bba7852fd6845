// spmv_impl.cuh -- device-side pieces shared by the SpMV kernels (spmv.cu, spmv_pk*.cu).
#pragma once
#include "spmv.cuh"

namespace db200 {

struct SellView {
    const int64_t* slice_ptr;
    const int32_t* col;
    const double* val;
    const int32_t* perm;
    long long nslices;
    const uint4* scd = nullptr;     // segment descriptors of the row-ordered twin (SellOp::scd), nullptr: none
    const uint8_t* sc = nullptr;    // per slice: 1 = columns come from scd
    int prefetch = 0;               // k_spmv_sell_scv: ask L2 for the next slice's values ("spmv.scv_prefetch")
};

struct EpiArgs {
    const double* aux;
    double omega;
    const double* beta;
    double beta_val;
    const double* dotv;
    double* dot_out;
    const int* done;
    const dd* dd_in;
    dd* dd_out;
    const double* xghost;
    int n_owned;
    dd* partials;
    unsigned* ticket;
};

// s = row sum, av = the row's aux entry (d_i, b_i or z_i)
template <int EPI>
__device__ __forceinline__ double apply_epilogue(double s, double av, const EpiArgs& e, double beta)
{
    if (EPI == EPI_JACOBI) return __ddiv_rn(__dmul_rn(e.omega, s), av);
    if (EPI == EPI_RESID)  return __dsub_rn(av, s);
    if (EPI == EPI_AXMBY)  return __dsub_rn(s, __dmul_rn(beta, av));
    return s;
}

// U = entries per thread loaded before the first use (loads in flight), MINB = resident CTAs of
// 512 threads the register allocation must allow (2: <= 64 regs, 3: <= 40, 4: <= 32).
// Stores the grid-wide fused dot product: rounded to *dot_out, or -- chained launches of a
// partitioned product -- unrounded (plus the earlier launches' *dd_in) to *dd_out.
__device__ __forceinline__ void store_dot(const EpiArgs& e, dd total)
{
    if (e.dd_in != nullptr) total = dd_add(total, *e.dd_in);
    if (e.dd_out != nullptr) *e.dd_out = total;
    else *e.dot_out = dd_round(total);
}

template <bool SPLIT>
__device__ __forceinline__ double load_x(const double* __restrict__ x, const EpiArgs& e, int c)
{
    if (SPLIT) return c < e.n_owned ? __ldg(x + c) : __ldg(e.xghost + (c - e.n_owned));
    return __ldg(x + c);
}


// host side (spmv.cu)
EpiArgs make_epi(const SpmvArgs& a);
// finishes a fused dot from `nwarps` per-warp partials in e.partials (second half of DOT == 2)
void sum_warp_partials(long long nwarps, const EpiArgs& e, cudaStream_t st);

}  // namespace db200

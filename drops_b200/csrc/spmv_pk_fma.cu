// spmv_pk_fma.cu -- the dictionary-packed SpMV with fused multiply-add ("spmv.exact" = 0): the EXACT = false half of the
// launch tables of spmv_pk_impl.cuh, in its own translation unit so that it compiles beside spmv_pk.cu.
#include "spmv_pk_impl.cuh"

namespace db200 {

void launch_pk_epi_fma(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    launch_pk_epi<false>(v, a, e, grid, threads, st);
}

}  // namespace db200

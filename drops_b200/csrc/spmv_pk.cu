// spmv_pk.cu -- y = A x on the dictionary-packed sliced-ELL stream (SellOp::pk, common.cuh).
//
// Same operator, same slices, same per-row left-to-right summation as k_spmv_sell (spmv.cu), but
// every stored entry is ONE 32-bit word  (dict index << cbits) | column : 4 bytes instead of 12 on
// the stream that bounds the kernel.  The dictionary holds the matrix' distinct fp64 bit patterns,
// so a_ij is recovered exactly and y stays bit-identical to y_Ax (num/spmat.h:1682-1709).  It
// applies to constant-coefficient systems on uniform meshes (the Poisson configs: 15 distinct
// values on the P2 Kuhn brick, 3 on P1); matrices that do not qualify have no pk and run spmv.cu.
//   * one coalesced 128 B load per entry-row of a warp, U of them in flight per thread before the
//     first gather -- at one register per staged entry (three in the 12-byte format);
//   * the dictionary sits in shared memory (<= kPkDictSmem values, copied by every CTA at start) or
//     is read through the read-only L1 path; lanes of a slice hold neighbouring rows of the same
//     stencil, so a lookup is mostly one broadcast wavefront;
//   * x through ld.global.nc; epilogues and the fused dot exactly as in spmv.cu.
#include "spmv_pk_impl.cuh"

namespace db200 {

int spmv_sell_pk(const SellOp& A, const SpmvArgs& a)
{
    Ctx& c = ctx();
    DB_ARG((A.pk != nullptr || A.vidx != nullptr) && A.dict != nullptr, "operator holds no packed stream");
    DB_ARG(!a.row_ordered || A.pkp != nullptr || (A.vidx != nullptr && A.colp != nullptr), "operator holds no row-ordered stream");
    DB_ARG(a.x != nullptr && a.y != nullptr, "null vector");
    DB_ARG(a.x != a.y, "x and y must not alias");
    DB_ARG(a.xghost == nullptr, "the packed stream has no split-x form");
    DB_ARG(a.epilogue == EPI_NONE || a.aux != nullptr, "epilogue needs aux vector");
    DB_ARG(a.dotv == nullptr || a.dot_out != nullptr || a.dd_out != nullptr, "fused dot needs an output scalar");
    DB_ARG(a.dotv != a.y, "the partner of the fused dot must not be the output vector");   // the tail pass re-reads y
    if (A.rows == 0) return 0;
    // the fused-dot product with the dictionary in the parameter block has its own geometry (measured on C2 inside
    // PCG: 15 words in flight x 384-thread CTAs x 192 slices, 3 CTAs of <= 56 registers per SM)
    const bool dot_geometry = a.dotv != nullptr && a.epilogue == EPI_NONE &&     // (the form launch_pk has the U15 kernel for)
                              A.pk != nullptr && A.ndict_host > 0 && A.ndict_host == A.ndict &&
                              A.ndict <= kPkDictSmem && opt("spmv.pkdict", 0.0) == 0.0 &&
                              opt("spmv.pk_cdict", 1.0) != 0.0 && opt("spmv.pk_cdict_dot", 1.0) != 0.0;
    // segment descriptors: the row-ordered twin of an operator that is mostly line segments ("spmv.segments")
    const bool six_ro = a.row_ordered && A.pk == nullptr && A.vidx != nullptr && A.colp != nullptr;     // twin of the six-byte form
    const bool use_sc = a.row_ordered && A.scd != nullptr && A.sc != nullptr && opt("spmv.segments", 1.0) != 0.0 &&
                        (six_ro || (A.pk != nullptr && A.ndict_host > 0 && A.ndict_host == A.ndict && A.ndict <= kPkDictSmem &&
                                    opt("spmv.pkdict", 0.0) == 0.0 && opt("spmv.pk_cdict", 1.0) != 0.0));
    int threads = (use_sc && six_ro) ? (int)opt("spmv.sci_threads", 512) : use_sc ? (int)opt("spmv.sc_threads", 512) : dot_geometry ? (int)opt("spmv.pk_dot_threads", 384) : (int)opt("spmv.pk_threads", 512);
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    threads &= ~31;
    // same decomposition as spmv_sell: CTAs own consecutive slices and are launched in matrix order
    const double per_sm = opt("spmv.ctas_per_sm", 0.0);
    long long grid;
    if (per_sm > 0.0) grid = (long long)(per_sm * c.sm_count + 0.5);
    else {
        long long spc = (use_sc && six_ro) ? (long long)opt("spmv.sci_slices_per_cta", 256.0) : use_sc ? (long long)opt("spmv.sc_slices_per_cta", 256.0) : dot_geometry ? (long long)opt("spmv.pk_dot_slices_per_cta", 192.0) : (long long)opt("spmv.pk_slices_per_cta", 128.0);
        if (spc < 1) spc = 1;
        grid = ((long long)A.nslices + spc - 1) / spc;
    }
    // small operators: shrink the CTAs until the grid fills the machine a few times over
    if (per_sm <= 0.0) {
        while (threads > 128 && grid < 8LL * c.sm_count) { threads >>= 1; grid <<= 1; }
        while (grid < 8LL * c.sm_count && grid * 2 * (threads >> 5) <= (long long)A.nslices) grid <<= 1;
        threads &= ~31;
    }
    const long long max_useful = ((long long)A.nslices * 32 + threads - 1) / threads;
    if (grid > max_useful) grid = max_useful;
    if (grid > kMaxReduceGrid) grid = kMaxReduceGrid;
    if (grid < 1) grid = 1;
    cudaStream_t st = a.stream ? a.stream : c.stream;
    PkView v{A.slice_ptr, six_ro ? A.colp : A.col, A.pk != nullptr ? nullptr : A.vidx, a.row_ordered ? A.pkp : A.pk, A.dict, a.row_ordered ? nullptr : A.perm, (long long)A.nslices,
             A.ndict, A.cbits, opt("spmv.pk_nopad", 1.0) != 0.0 ? A.pk_nopad : nullptr,
             (int)opt("spmv.sc_prefetch", 1.0), use_sc ? A.scd : nullptr, use_sc ? A.sc : nullptr,
             A.ndict_host == A.ndict ? A.ndict_host : 0, {}};
    if (v.cdict_n > 0) std::memcpy(v.cdict, A.dict_host, sizeof(double) * ((size_t)A.ndict_host + 1));
    EpiArgs e = make_epi(a);
    launch_pk_epi<true>(v, a, e, (int)grid, threads, st);      // reference arithmetic: the product is rounded before the add
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace db200

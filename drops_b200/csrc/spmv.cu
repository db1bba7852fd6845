// spmv.cu -- y = A x on sm_100a.
//
// The hot kernel of every Krylov iteration (reference: y_Ax, num/spmat.h:1682-1709).  It is
// HBM-bound (2 flops per 12 matrix bytes), so the design goal is: stream col/val exactly once,
// fully coalesced, with many loads in flight; keep x in L1/L2; write y once.  No tensor cores.
//
// k_spmv_sell: sliced ELLPACK, slice height 32 = one warp, one thread per row.
//   * inside a slice entry j of all 32 rows is contiguous -> every load instruction of a warp is
//     one 128 B (col) / 256 B (val) fully used line; rows are sigma-sorted by length at upload so
//     padding is < 1 % on the P2 brick (65/27/27/19 entries per row).
//   * each thread walks ITS row left to right: sum = ((0 + a0*x0) + a1*x1) + ... with the product
//     rounded before the add -- the reference's summation order, so the result is bit-identical
//     to y_Ax compiled without FMA contraction.  The kernel is nowhere near FP64-bound, so the unfused
//     mul/add costs nothing.
//   * matrix stream: ld.global.cs (evict-first, does not displace x); x: ld.global.nc (L1).
//   * a CTA owns a contiguous range of slices and its warps walk it round-robin, so the x lines
//     touched by neighbouring mesh lines are reused out of the SM's L1.
//   * optional epilogues (Jacobi scaling, b - Ax, Ax - beta z) and a fused dot product with a
//     second vector save whole passes over HBM in PCG / GMRES / MINRES.
// k_spmv_csr_vector: LANES threads per row on plain CSR (FMA + shuffle tree); kept as the
//   comparison point ("spmv.kernel" = 1) and for matrices where SELL padding would explode.
#include "spmv_impl.cuh"

namespace db200 {

// Sum of one row of a slice, left to right.  (16-bit column deltas, "sell.compress", were measured slower --
// the extra load per entry-row costs more than the bytes save, profiles/r01_notes.md -- and are gone.)
template <bool EXACT, int U, bool SPLIT>
__device__ __forceinline__ double slice_row_sum(const SellView& A, int64_t base, int width, int lane,
                                                const double* __restrict__ x, const EpiArgs& e)
{
    const double* vp = A.val + base + lane;
    const int32_t* cp = A.col + base + lane;
    double sum = 0.0;
    int j = 0;
    for (; j + U <= width; j += U) {
        int c[U];
        double v[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            c[u] = __ldcs(cp + (size_t)(j + u) * 32);
            v[u] = __ldcs(vp + (size_t)(j + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = c[u] >= 0 ? load_x<SPLIT>(x, e, c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (c[u] >= 0)
                sum = EXACT ? __dadd_rn(sum, __dmul_rn(v[u], xv[u])) : fma(v[u], xv[u], sum);
        }
    }
    for (; j < width; ++j) {
        const int c = __ldcs(cp + (size_t)j * 32);
        const double v = __ldcs(vp + (size_t)j * 32);
        if (c >= 0) {
            const double xv = load_x<SPLIT>(x, e, c);
            sum = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv)) : fma(v, xv, sum);
        }
    }
    return sum;
}

// DOT: 0 = none; 1 = fused dot finished in this kernel (block reduce + two-level grid reduce);
// 2 = fused dot, every WARP stores its double-double partial and k_sum_warp_partials finishes it
// (no block barrier and no atomics in the SpMV itself).
template <int EPI, bool EXACT, int DOT, int U, int MAXT, int MINB, bool SPLIT = false>
__global__ void __launch_bounds__(MAXT, MINB)
k_spmv_sell(SellView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e)
{
    constexpr int kSellUnroll = U;
    if (e.done != nullptr && *e.done != 0) return;
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long s_begin = A.nslices * (long long)blockIdx.x / gridDim.x;
    const long long s_end   = A.nslices * (long long)(blockIdx.x + 1) / gridDim.x;
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    dd acc[1] = {dd_zero()};

    for (long long s = s_begin + warp; s < s_end; s += nwarp) {
        const int64_t base = A.slice_ptr[s];
        const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
        // the row index and the row's epilogue operands are fetched up front: their latency hides
        // behind the matrix stream instead of trailing it (two dependent misses per slice otherwise)
        // row-ordered twin (perm == nullptr): position s*32+lane IS the row of y, aux and the dot partner
        const int row = A.perm != nullptr ? A.perm[s * 32 + lane] : (int)(s * 32 + lane);
        double dv = 0.0, av = 0.0;
        if (row >= 0) {
            if (DOT) dv = __ldg(e.dotv + row);
            if (EPI != EPI_NONE) av = __ldg(e.aux + row);
        }
        const double sum = slice_row_sum<EXACT, kSellUnroll, SPLIT>(A, base, width, lane, x, e);
        if (row >= 0) {
            const double out = apply_epilogue<EPI>(sum, av, e, beta);
            y[row] = out;
            if (DOT) acc[0] = dd_add_prod(acc[0], out, dv);
        }
    }
    if (DOT == 1) {
        if (grid_reduce_dd<1>(acc, red_smem, e.partials, e.ticket)) {
            if (threadIdx.x == 0) store_dot(e, acc[0]);
        }
    }
    if (DOT == 2) {
        const dd w = warp_reduce_dd(acc[0]);
        if (lane == 0) e.partials[(size_t)blockIdx.x * nwarp + warp] = w;
    }
}

// Second half of DOT == 2: deterministic sum of the per-warp partials (fixed thread assignment,
// fixed order, fixed tree).
__global__ void __launch_bounds__(kReduceThreads)
k_sum_warp_partials(long long count, const dd* __restrict__ in, EpiArgs e, dd* partials, unsigned* ticket)
{
    __shared__ dd red_smem[32];
    if (e.done != nullptr && *e.done != 0) return;
    dd acc[1] = {dd_zero()};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) acc[0] = dd_add(acc[0], in[i]);
    if (grid_reduce_dd<1>(acc, red_smem, partials, ticket)) {
        if (threadIdx.x == 0) store_dot(e, acc[0]);
    }
}

void sum_warp_partials(long long nwarps, const EpiArgs& e, cudaStream_t st)
{
    Ctx& c = ctx();
    const int g2 = vec_grid((size_t)nwarps, kReduceThreads, 2);
    k_sum_warp_partials<<<g2, kReduceThreads, 0, st>>>(nwarps, e.partials, e, c.partials + (size_t)3 * kPartialStride,
                                                     c.tickets + (size_t)3 * kTicketStride);
    count_launch();
}

template <int LANES, int EPI, bool DOT>
__global__ void __launch_bounds__(256)
k_spmv_csr_vector(long long rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                  EpiArgs e)
{
    if (e.done != nullptr && *e.done != 0) return;
    __shared__ dd red_smem[32];
    constexpr int GPW = 32 / LANES;   // row groups per warp
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES, grp = lane / LANES;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    dd acc[1] = {dd_zero()};

    for (long long rb = warp_global * GPW; rb < rows; rb += nwarps * GPW) {
        const long long row = rb + grp;
        double sum = 0.0;
        if (row < rows) {
            const int64_t b = rowptr[row], en = rowptr[row + 1];
            for (int64_t k = b + sub; k < en; k += LANES)
                sum = fma(__ldcs(val + k), __ldg(x + __ldcs(col + k)), sum);
        }
#pragma unroll
        for (int m = LANES / 2; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
        if (row < rows && sub == 0) {
            const double out = apply_epilogue<EPI>(sum, EPI != EPI_NONE ? __ldg(e.aux + row) : 0.0, e, beta);
            y[row] = out;
            if (DOT) acc[0] = dd_add_prod(acc[0], out, __ldg(e.dotv + row));
        }
    }
    if (DOT) {
        if (grid_reduce_dd<1>(acc, red_smem, e.partials, e.ticket)) {
            if (threadIdx.x == 0) store_dot(e, acc[0]);
        }
    }
}

EpiArgs make_epi(const SpmvArgs& a)
{
    Ctx& c = ctx();
    EpiArgs e;
    e.aux = a.aux;
    e.omega = a.omega;
    e.beta = a.beta;
    e.beta_val = a.beta_val;
    e.dotv = a.dotv;
    e.dot_out = a.dot_out;
    e.done = a.done;
    e.dd_in = a.dd_in;
    e.dd_out = a.dd_out;
    e.xghost = a.xghost;
    e.n_owned = (int)a.n_owned;
    e.partials = c.partials + (size_t)a.slot * kPartialStride;
    e.ticket = c.tickets + (size_t)a.slot * kTicketStride;
    return e;
}

template <int EPI, bool EXACT, int U, int MAXT, int MINB, bool SPLIT = false>
static void launch_sell_v(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads,
                          cudaStream_t st)
{
    if (threads > MAXT) threads = MAXT;
    if (a.dotv == nullptr) {
        k_spmv_sell<EPI, EXACT, 0, U, MAXT, MINB, SPLIT><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
        return;
    }
    const long long nwarps = (long long)grid * (threads >> 5);
    // small grids: finish the dot inside the kernel (one launch less per iteration, which is what
    // counts on launch-bound systems); large grids: per-warp partials + finishing kernel
    const double mode = opt("spmv.dotmode", 0.0);
    const bool two_kernels = mode == 2.0 || (mode == 0.0 && grid > (int)kReduceFlat);
    if (two_kernels && nwarps <= kPartialStride) {
        k_spmv_sell<EPI, EXACT, 2, U, MAXT, MINB, SPLIT><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
        sum_warp_partials(nwarps, e, st);
    } else {
        k_spmv_sell<EPI, EXACT, 1, U, MAXT, MINB, SPLIT><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
    }
}


// ------------------------------------------------------------------------------------------
// k_spmv_sell_scv: the row-ordered product on the 12-byte stream of an operator with segment descriptors ("spmv.segments").
// In a slice that is one line segment the columns of entry-row j are at most two runs of consecutive positions --
// a 16-byte record (SellOp::scd) instead of 32 int32 columns --, the fp64 values stay per lane: 8.5 instead of 12 bytes
// per stored entry on the stream that bounds this kernel.  Records are staged per warp in shared memory with cp.async
// (the copy for the warp's next slice travels during the epilogue).  Unflagged slices run slice_row_sum.  Same
// products in the same order as k_spmv_sell: bit-identical.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void scv_issue(const SellView& A, uint32_t sd_addr, int64_t base, int width, int lane)
{
    const uint4* src = A.scd + (base >> 5);
    for (int j = lane; j < width; j += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sd_addr + (uint32_t)j * 16u), "l"(src + j) : "memory");
}

template <bool EXACT, int UB>
__device__ __forceinline__ double scv_batch(const double* __restrict__ vp, const uint4* __restrict__ sd, int j, double sum, int lane,
                                            const double* __restrict__ x)
{
    double v[UB], xv[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) v[u] = __ldcs(vp + (size_t)(j + u) * 32);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const uint4 d = sd[j + u];
        const int c = (lane < (int)d.z ? (int)d.x : (int)d.y) + lane;
        xv[u] = __ldg(x + c);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) sum = EXACT ? __dadd_rn(sum, __dmul_rn(v[u], xv[u])) : fma(v[u], xv[u], sum);
    return sum;
}

template <int EPI, bool EXACT, int DOT, int U, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_spmv_sell_scv(SellView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e)
{
    if (e.done != nullptr && *e.done != 0) return;
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    __shared__ uint4 sdesc[(MAXT / 32) * kScMaxWidth];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int s_begin = (int)(A.nslices * (long long)blockIdx.x / gridDim.x);
    const int s_end   = (int)(A.nslices * (long long)(blockIdx.x + 1) / gridDim.x);
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    const uint4* sd = sdesc + warp * kScMaxWidth;
    const uint32_t sd_addr = (uint32_t)__cvta_generic_to_shared(sd);
    dd acc[1] = {dd_zero()};
    int64_t base = 0;
    int width = 0;
    bool seg = false;
    if (s_begin + warp < s_end) {
        const int s0 = s_begin + warp;
        base = A.slice_ptr[s0];
        width = (int)((A.slice_ptr[s0 + 1] - base) >> 5);
        seg = A.sc[s0] != 0;
        if (seg) scv_issue(A, sd_addr, base, width, lane);
    }
    for (int s = s_begin + warp; s < s_end; s += nwarp) {
        int64_t nbase = 0;
        int nwidth = 0;
        bool nseg = false;
        const int s1 = s + nwarp;
        if (s1 < s_end) {
            nbase = A.slice_ptr[s1];
            nwidth = (int)((A.slice_ptr[s1 + 1] - nbase) >> 5);
            nseg = A.sc[s1] != 0;
        }
        const int row = s * 32 + lane;                        // row-ordered twin: the position IS the row
        double dv = 0.0, av = 0.0;
        if (DOT) dv = __ldg(e.dotv + row);
        if (EPI != EPI_NONE) av = __ldg(e.aux + row);
        double sum = 0.0;
        if (seg) {
            const double* vp = A.val + base + lane;
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            int j = 0;
            for (; j + U <= width; j += U) sum = scv_batch<EXACT, U>(vp, sd, j, sum, lane, x);
            if (U > 4 && j + 4 <= width) { sum = scv_batch<EXACT, 4>(vp, sd, j, sum, lane, x); j += 4; }
            if (U > 2 && j + 2 <= width) { sum = scv_batch<EXACT, 2>(vp, sd, j, sum, lane, x); j += 2; }
            for (; j < width; ++j) sum = scv_batch<EXACT, 1>(vp, sd, j, sum, lane, x);
            __syncwarp();                                     // every lane is done with the records
        } else {
            sum = slice_row_sum<EXACT, U, false>(A, base, width, lane, x, e);
        }
        if (nseg) {
            scv_issue(A, sd_addr, nbase, nwidth, lane);
            // ... and L2 is asked for the next slice's values (two 128-byte lines per entry-row)
            if (A.prefetch != 0)
                for (int l = lane; l < 2 * nwidth; l += 32)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(A.val + nbase + (size_t)l * 16));
        }
        const double out = apply_epilogue<EPI>(sum, av, e, beta);
        y[row] = out;
        if (DOT) acc[0] = dd_add_prod(acc[0], out, dv);
        base = nbase; width = nwidth; seg = nseg;
    }
    if (DOT == 1) {
        if (grid_reduce_dd<1>(acc, red_smem, e.partials, e.ticket)) {
            if (threadIdx.x == 0) store_dot(e, acc[0]);
        }
    }
    if (DOT == 2) {
        const dd w = warp_reduce_dd(acc[0]);
        if (lane == 0) e.partials[(size_t)blockIdx.x * nwarp + warp] = w;
    }
}

template <int EPI, bool EXACT, int U, int MAXT, int MINB>
static void launch_scv(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    if (threads > MAXT) threads = MAXT;
    if (a.dotv == nullptr) {
        k_spmv_sell_scv<EPI, EXACT, 0, U, MAXT, MINB><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
        return;
    }
    const long long nwarps = (long long)grid * (threads >> 5);
    const double mode = opt("spmv.dotmode", 0.0);
    const bool two_kernels = mode == 2.0 || (mode == 0.0 && grid > (int)kReduceFlat);
    if (two_kernels && nwarps <= kPartialStride) {
        k_spmv_sell_scv<EPI, EXACT, 2, U, MAXT, MINB><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
        sum_warp_partials(nwarps, e, st);
    } else {
        k_spmv_sell_scv<EPI, EXACT, 1, U, MAXT, MINB><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
    }
}

// Measured on B200 (profiles/r01_probe_dot.json, r01_sweep1_spmv.jsonl): the plain product is fastest with U4 and <= 40
// registers (2.96 ms at C2), the product with the fused dot needs the room of U8 / <= 48 registers in 128-thread CTAs
// (3.12 ms; 3.55 ms when squeezed into 40).  The other unroll x register-cap variants of round 1 are gone.
template <int EPI, bool EXACT>
static void launch_sell(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads,
                        cudaStream_t st)
{
    if (a.xghost != nullptr) {   // boundary rows of a partition: a thin operator, one variant
        launch_sell_v<EPI, EXACT, 4, 512, 2, true>(v, a, e, grid, threads, st);
        return;
    }
    if (v.scd != nullptr && v.sc != nullptr && v.perm == nullptr) {     // row-ordered twin with segment descriptors
        // measured on C2 (profiles/r02_probe_scv.json): 8 entries in flight x 256-thread CTAs x 5 per SM 2.54 ms (k_spmv_sell: 3.14);
        // 10 x 256 x 4: 2.53, 6 x 256 x 6: 2.70; 16-64 slices per CTA and 128-thread CTAs: the same
        launch_scv<EPI, EXACT, 8, 256, 5>(v, a, e, grid, threads, st);
        return;
    }
    if (a.dotv != nullptr) launch_sell_v<EPI, EXACT, 8, 128, 10>(v, a, e, grid, threads, st);
    else launch_sell_v<EPI, EXACT, 4, 512, 3>(v, a, e, grid, threads, st);
}

template <bool EXACT>
static void launch_sell_epi(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads,
                            cudaStream_t st)
{
    switch (a.epilogue) {
        case EPI_JACOBI: launch_sell<EPI_JACOBI, EXACT>(v, a, e, grid, threads, st); break;
        case EPI_RESID:  launch_sell<EPI_RESID, EXACT>(v, a, e, grid, threads, st); break;
        case EPI_AXMBY:  launch_sell<EPI_AXMBY, EXACT>(v, a, e, grid, threads, st); break;
        default:         launch_sell<EPI_NONE, EXACT>(v, a, e, grid, threads, st); break;
    }
}

int spmv_sell(const SellOp& A, const SpmvArgs& a)
{
    Ctx& c = ctx();
    DB_ARG(a.x != nullptr && a.y != nullptr, "null vector");
    DB_ARG(a.x != a.y, "x and y must not alias");
    DB_ARG(a.epilogue == EPI_NONE || a.aux != nullptr, "epilogue needs aux vector");
    DB_ARG(a.dotv == nullptr || a.dot_out != nullptr || a.dd_out != nullptr, "fused dot needs an output scalar");
    if (A.rows == 0) return 0;
    // dictionary-packed stream (4 B per entry, spmv_pk.cu) when the operator has one
    if ((A.pk != nullptr || A.vidx != nullptr) && a.xghost == nullptr && opt("spmv.pk", 1.0) != 0.0 &&
        (!a.row_ordered || A.pkp != nullptr || (A.pk == nullptr && A.vidx != nullptr && A.colp != nullptr))) {
        return spmv_sell_pk(A, a);
    }
    DB_ARG(!a.row_ordered || A.colp != nullptr, "row-ordered vectors need the operator's row-ordered stream");
    DB_ARG(A.val != nullptr, "the 12-byte stream of this operator was released once its packed stream existed "
                             "(\"sell.release_general\" = 0 keeps it)");
    const bool use_sc = a.row_ordered && a.xghost == nullptr && A.scd != nullptr && A.sc != nullptr && opt("spmv.segments", 1.0) != 0.0;
    int threads = use_sc ? (int)opt("spmv.scv_threads", 256) : (int)opt("spmv.threads", 128);
    if (threads < 32) threads = 32;
    if (threads > 512) threads = 512;
    threads &= ~31;
    if (a.xghost != nullptr) threads = 64;       // thin boundary operator: one slice per warp (see below)
    // Work decomposition: every CTA owns `slices_per_cta` consecutive slices, CTAs are launched in
    // matrix order.  Measured on B200 (profiles/r01_spmv.md): many short ranges walked in order
    // (128 threads x 16 slices = 512 rows per CTA, 98 k CTAs at C2) reach 5.9 TB/s, a persistent
    // grid of 4 CTAs/SM only 3.3 TB/s -- all resident CTAs then stream one compact window of the
    // matrix and share one window of x in L2.  "spmv.ctas_per_sm" > 0 selects the persistent-style
    // grid instead (kept for the comparison).
    const double per_sm = opt("spmv.ctas_per_sm", 0.0);
    long long grid;
    if (per_sm > 0.0) grid = (long long)(per_sm * c.sm_count + 0.5);
    else {
        // boundary rows of a partition are a thin operator (0.4 % of the rows at C5): short CTAs, or the
        // launch does not even fill the machine once
        long long spc = (long long)(a.xghost != nullptr ? opt("spmv.bnd_slices_per_cta", 2.0)
                                    : use_sc ? opt("spmv.scv_slices_per_cta", 32.0) : opt("spmv.slices_per_cta", 16.0));
        if (spc < 1) spc = 1;
        grid = ((long long)A.nslices + spc - 1) / spc;
    }
    const long long max_useful = ((long long)A.nslices * 32 + threads - 1) / threads;
    if (grid > max_useful) grid = max_useful;
    if (grid > kMaxReduceGrid) grid = kMaxReduceGrid;
    if (grid < 1) grid = 1;
    cudaStream_t st = a.stream ? a.stream : c.stream;
    SellView v{A.slice_ptr, a.row_ordered ? A.colp : A.col, A.val, a.row_ordered ? nullptr : A.perm, (long long)A.nslices,
               use_sc ? A.scd : nullptr, use_sc ? A.sc : nullptr, (int)opt("spmv.scv_prefetch", 1.0)};
    EpiArgs e = make_epi(a);
    launch_sell_epi<true>(v, a, e, (int)grid, threads, st);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

template <int LANES, int EPI>
static void launch_csrv(const CsrDev& A, const SpmvArgs& a, const EpiArgs& e, int grid, cudaStream_t st)
{
    if (a.dotv != nullptr)
        k_spmv_csr_vector<LANES, EPI, true><<<grid, 256, 0, st>>>((long long)A.rows, A.rowptr, A.col, A.val, a.x, a.y, e);
    else
        k_spmv_csr_vector<LANES, EPI, false><<<grid, 256, 0, st>>>((long long)A.rows, A.rowptr, A.col, A.val, a.x, a.y, e);
}

template <int LANES>
static void launch_csrv_epi(const CsrDev& A, const SpmvArgs& a, const EpiArgs& e, int grid, cudaStream_t st)
{
    switch (a.epilogue) {
        case EPI_JACOBI: launch_csrv<LANES, EPI_JACOBI>(A, a, e, grid, st); break;
        case EPI_RESID:  launch_csrv<LANES, EPI_RESID>(A, a, e, grid, st); break;
        case EPI_AXMBY:  launch_csrv<LANES, EPI_AXMBY>(A, a, e, grid, st); break;
        default:         launch_csrv<LANES, EPI_NONE>(A, a, e, grid, st); break;
    }
}

int spmv_csr_vector(const CsrDev& A, const SpmvArgs& a, int lanes)
{
    Ctx& c = ctx();
    DB_ARG(a.x != nullptr && a.y != nullptr && a.x != a.y, "bad vectors");
    DB_ARG(A.rowptr != nullptr, "matrix holds no CSR copy");
    if (A.rows == 0) return 0;
    long long grid = (long long)c.sm_count * 8;
    const long long rows_per_block = 256 / lanes;
    const long long need = ((long long)A.rows + rows_per_block - 1) / rows_per_block;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaStream_t st = a.stream ? a.stream : c.stream;
    EpiArgs e = make_epi(a);
    switch (lanes) {
        case 2:  launch_csrv_epi<2>(A, a, e, (int)grid, st); break;
        case 4:  launch_csrv_epi<4>(A, a, e, (int)grid, st); break;
        case 8:  launch_csrv_epi<8>(A, a, e, (int)grid, st); break;
        case 16: launch_csrv_epi<16>(A, a, e, (int)grid, st); break;
        default: launch_csrv_epi<32>(A, a, e, (int)grid, st); break;
    }
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

int spmv_apply(const drops_b200_matrix_s* M, bool transposed, const SpmvArgs& a)
{
    ScopedTiming t(transposed ? "spmv_t" : "spmv");
    const int kind = a.row_ordered ? 0 : (int)opt("spmv.kernel", 0.0);
    if (kind == 1) {
        const int lanes = (int)opt("spmv.lanes", 8.0);
        DB_ARG(!transposed || M->csr_t.rowptr != nullptr || M->csr_t.nnz == 0,
               "the CSR form of A^T was released with its 12-byte stream (\"sell.release_general\" = 0 keeps it)");
        return spmv_csr_vector(transposed ? M->csr_t : M->csr, a, lanes);
    }
    return spmv_sell(transposed ? M->sell_t : M->sell, a);
}

}  // namespace db200

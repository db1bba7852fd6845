// spmv_tma.cu -- y = A x with the matrix stream staged through shared memory by the TMA engine.
//
// Why: the register-staged kernel (spmv.cu) is latency-bound, not byte-bound -- its loads in
// flight are limited by registers x occupancy (ncu: 85 % long-scoreboard stalls at 6.0 TB/s), and
// shrinking the stream (16-bit column deltas) did not make it faster.  Here every WARP owns a
// contiguous run of slices and streams their val / col data with `cp.async.bulk` (1-D TMA copies,
// L2 evict-first) into a private ring of STAGES x R entry-rows in shared memory, completion
// signalled on one mbarrier per stage.  The bytes in flight no longer cost registers: 4 warps x 4
// stages x 2.5 KB = 40 KB per CTA.  The consumer side is unchanged in its arithmetic: one thread
// per row, left-to-right sum, product rounded before the add (bit-identical to y_Ax), x gathered
// with ld.global.nc, the same epilogues and the same fused dot.
//
// Layout consumed: SellOp (matrix.cu).  With "sell.compress" the column stream is col16 (2 B per
// entry) + one int32 base per entry-row (ld.global, warp-uniform); slices flagged wide read their
// 32-bit columns straight from global memory.
#include "spmv_impl.cuh"
#include "tma_util.cuh"

namespace db200 {

// R = entry-rows (32 entries each) per stage, U = entry-rows whose x gathers are issued together.
template <int EPI, bool EXACT, int DOT, bool C16, int WARPS, int STAGES, int R, int U>
__global__ void __launch_bounds__(WARPS * 32)
k_spmv_sell_tma(SellView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e, int* __restrict__ fault)
{
    static_assert(R % U == 0, "a stage holds whole gather batches");
    constexpr int kColBytes = C16 ? 2 : 4;
    struct __align__(128) Stage {
        double val[R * 32];
        unsigned char col[R * 32 * kColBytes];
    };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage (*ring)[STAGES] = reinterpret_cast<Stage (*)[STAGES]>(smem_raw);
    unsigned long long (*full)[STAGES] =
        reinterpret_cast<unsigned long long (*)[STAGES]>(smem_raw + sizeof(Stage) * WARPS * STAGES);
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    if (e.done != nullptr && *e.done != 0) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long s_begin = A.nslices * (long long)blockIdx.x / gridDim.x;
    const long long s_end   = A.nslices * (long long)(blockIdx.x + 1) / gridDim.x;
    const long long cnt = s_end - s_begin;
    long long s = s_begin + cnt * warp / WARPS;               // this warp's contiguous slices
    const long long ws_end = s_begin + cnt * (warp + 1) / WARPS;
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    dd acc[1] = {dd_zero()};

    if (lane == 0)
        for (int st = 0; st < STAGES; ++st) mbar_init(&full[warp][st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    if (s < ws_end) {
        const int64_t e0 = A.slice_ptr[s];
        const long long J = (A.slice_ptr[ws_end] - e0) >> 5;  // entry-rows of this warp's stream
        const long long nchunks = (J + R - 1) / R;
        const unsigned long long policy = policy_evict_first();
        auto issue = [&](long long t) {                       // lane 0: chunk t -> stage t % STAGES
            const int st = (int)(t % STAGES);
            const long long j0 = t * R;
            const unsigned rows = (unsigned)(J - j0 < R ? J - j0 : R);
            const int64_t ent = e0 + j0 * 32;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[warp][st], rows * 32u * (8u + kColBytes));
            bulk_g2s(ring[warp][st].val, A.val + ent, rows * 256u, &full[warp][st], policy);
            if (C16) bulk_g2s(ring[warp][st].col, A.col16 + ent, rows * 64u, &full[warp][st], policy);
            else     bulk_g2s(ring[warp][st].col, A.col + ent, rows * 128u, &full[warp][st], policy);
        };
        if (lane == 0)
            for (long long t = 0; t < STAGES && t < nchunks; ++t) issue(t);

        // per-slice state, fetched one slice ahead of its use
        auto width_of = [&](long long sl) -> int { return (int)((A.slice_ptr[sl + 1] - A.slice_ptr[sl]) >> 5); };
        int rem = width_of(s);
        int row = A.perm[s * 32 + lane];
        bool wide = C16 && A.wide[s] != 0;
        double dv = 0.0, av = 0.0;
        if (row >= 0) {
            if (DOT) dv = __ldg(e.dotv + row);
            if (EPI != EPI_NONE) av = __ldg(e.aux + row);
        }
        double sum = 0.0;
        long long jglob = 0;                                  // entry-row index inside the warp's stream
        const int32_t* bp = C16 ? A.cbase + (e0 >> 5) : nullptr;

        auto finish_slice = [&]() {
            if (row >= 0) {
                const double out = apply_epilogue<EPI>(sum, av, e, beta);
                y[row] = out;
                if (DOT) acc[0] = dd_add_prod(acc[0], out, dv);
            }
            sum = 0.0;
            ++s;
            while (s < ws_end) {                              // skip slices without stored entries
                rem = width_of(s);
                row = A.perm[s * 32 + lane];
                if (rem > 0) break;
                if (row >= 0) {
                    const double out = apply_epilogue<EPI>(0.0, EPI != EPI_NONE ? __ldg(e.aux + row) : 0.0, e, beta);
                    y[row] = out;
                    if (DOT) acc[0] = dd_add_prod(acc[0], out, __ldg(e.dotv + row));
                }
                ++s;
            }
            if (s < ws_end) {
                wide = C16 && A.wide[s] != 0;
                dv = 0.0; av = 0.0;
                if (row >= 0) {
                    if (DOT) dv = __ldg(e.dotv + row);
                    if (EPI != EPI_NONE) av = __ldg(e.aux + row);
                }
            }
        };
        // leading slices without entries
        if (rem == 0) { rem = 0; finish_slice(); }

        unsigned parity = 0;
        int stg = 0;
        for (long long t = 0; t < nchunks; ++t) {
            unsigned long long* bar = &full[warp][stg];
            unsigned spins = 0;
            while (!mbar_try_wait(bar, parity)) {
                if (++spins > (1u << 26)) {                   // never hang the GPU: report and bail out
                    if (lane == 0) atomicExch(fault, 1);
                    return;
                }
            }
            const Stage& S = ring[warp][stg];
            const int rows = (int)(J - t * R < R ? J - t * R : R);
            for (int jj = 0; jj < rows; jj += U) {
                int c[U];
                double v[U], xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (jj + u < rows) {
                        v[u] = S.val[(jj + u) * 32 + lane];
                        if (C16) {
                            const int d = reinterpret_cast<const unsigned short*>(S.col)[(jj + u) * 32 + lane];
                            c[u] = d == 0xFFFF ? -1 : __ldg(bp + jglob + jj + u) + d;
                        } else {
                            c[u] = reinterpret_cast<const int*>(S.col)[(jj + u) * 32 + lane];
                        }
                    } else { v[u] = 0.0; c[u] = -1; }
                }
                if (C16 && wide) {                            // rare: columns of this slice spread > 2^16
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (jj + u < rows) c[u] = __ldcs(A.col + e0 + (jglob + jj + u) * 32 + lane);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) xv[u] = c[u] >= 0 ? __ldg(x + c[u]) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (jj + u < rows) {
                        if (c[u] >= 0) sum = EXACT ? __dadd_rn(sum, __dmul_rn(v[u], xv[u])) : fma(v[u], xv[u], sum);
                        if (--rem == 0) {
                            // a slice ends inside the batch: the columns of the following entry-rows were
                            // decoded with the OLD slice's wide flag -- redo them if the flag changes
                            const bool was_wide = wide;
                            finish_slice();
                            if (C16 && wide != was_wide) {
#pragma unroll
                                for (int w = 0; w < U; ++w)
                                    if (w > u && jj + w < rows) {
                                        if (wide) c[w] = __ldcs(A.col + e0 + (jglob + jj + w) * 32 + lane);
                                        else {
                                            const int d = reinterpret_cast<const unsigned short*>(S.col)[(jj + w) * 32 + lane];
                                            c[w] = d == 0xFFFF ? -1 : __ldg(bp + jglob + jj + w) + d;
                                        }
                                        xv[w] = c[w] >= 0 ? __ldg(x + c[w]) : 0.0;
                                    }
                            }
                        }
                    }
                }
            }
            jglob += rows;
            __syncwarp();                                     // every lane is done reading this stage
            if (lane == 0 && t + STAGES < nchunks) issue(t + STAGES);
            if (++stg == STAGES) { stg = 0; parity ^= 1u; }
        }
    }
    if (DOT == 1) {
        if (grid_reduce_dd<1>(acc, red_smem, e.partials, e.ticket)) {
            if (threadIdx.x == 0) store_dot(e, acc[0]);
        }
    }
    if (DOT == 2) {
        const dd w = warp_reduce_dd(acc[0]);
        if (lane == 0) e.partials[(size_t)blockIdx.x * WARPS + warp] = w;
    }
}

template <class K>
static int allow_smem(K kernel, size_t bytes, bool* done)
{
    if (*done) return 0;
    DB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    *done = true;
    return 0;
}

template <int EPI, bool EXACT, bool C16, int WARPS, int STAGES, int R, int U>
static int launch_tma_v(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, cudaStream_t st, int* fault)
{
    constexpr size_t kSmem = (size_t)WARPS * STAGES * (R * 32 * (8 + (C16 ? 2 : 4))) + (size_t)WARPS * STAGES * 8;
    static bool set0 = false, set2 = false;
    if (a.dotv == nullptr) {
        DB_CHECK(allow_smem(k_spmv_sell_tma<EPI, EXACT, 0, C16, WARPS, STAGES, R, U>, kSmem, &set0));
        k_spmv_sell_tma<EPI, EXACT, 0, C16, WARPS, STAGES, R, U><<<grid, WARPS * 32, kSmem, st>>>(v, a.x, a.y, e, fault);
        return 0;
    }
    const long long nwarps = (long long)grid * WARPS;
    if (nwarps > kPartialStride) { set_error("spmv_tma: too many CTAs for the partial buffer"); return DROPS_B200_E_ARG; }
    DB_CHECK(allow_smem(k_spmv_sell_tma<EPI, EXACT, 2, C16, WARPS, STAGES, R, U>, kSmem, &set2));
    k_spmv_sell_tma<EPI, EXACT, 2, C16, WARPS, STAGES, R, U><<<grid, WARPS * 32, kSmem, st>>>(v, a.x, a.y, e, fault);
    sum_warp_partials(nwarps, e, st);
    return 0;
}

template <int EPI, bool EXACT, bool C16>
static int launch_tma_geo(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, cudaStream_t st, int* fault)
{
    switch ((int)opt("spmv.tma_geometry", 0.0)) {     // warps x stages x rows per stage x gather batch
        case 1: return launch_tma_v<EPI, EXACT, C16, 4, 3, 16, 8>(v, a, e, grid, st, fault);
        case 2: return launch_tma_v<EPI, EXACT, C16, 4, 4, 8, 4>(v, a, e, grid, st, fault);
        case 3: return launch_tma_v<EPI, EXACT, C16, 8, 4, 8, 8>(v, a, e, grid, st, fault);
        case 4: return launch_tma_v<EPI, EXACT, C16, 4, 6, 8, 8>(v, a, e, grid, st, fault);
        default: return launch_tma_v<EPI, EXACT, C16, 4, 4, 8, 8>(v, a, e, grid, st, fault);
    }
}

template <int EPI, bool EXACT>
static int launch_tma_c(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, cudaStream_t st, int* fault)
{
    // the geometry variants are built for the PCG forms only (plain product with and without dot)
    if (EPI == EPI_NONE) {
        if (v.col16 != nullptr) return launch_tma_geo<EPI_NONE, EXACT, true>(v, a, e, grid, st, fault);
        return launch_tma_geo<EPI_NONE, EXACT, false>(v, a, e, grid, st, fault);
    }
    if (v.col16 != nullptr) return launch_tma_v<EPI, EXACT, true, 4, 4, 8, 8>(v, a, e, grid, st, fault);
    return launch_tma_v<EPI, EXACT, false, 4, 4, 8, 8>(v, a, e, grid, st, fault);
}

template <bool EXACT>
static int launch_tma_epi(const SellView& v, const SpmvArgs& a, const EpiArgs& e, int grid, cudaStream_t st, int* fault)
{
    switch (a.epilogue) {
        case EPI_JACOBI: return launch_tma_c<EPI_JACOBI, EXACT>(v, a, e, grid, st, fault);
        case EPI_RESID:  return launch_tma_c<EPI_RESID, EXACT>(v, a, e, grid, st, fault);
        case EPI_AXMBY:  return launch_tma_c<EPI_AXMBY, EXACT>(v, a, e, grid, st, fault);
        default:         return launch_tma_c<EPI_NONE, EXACT>(v, a, e, grid, st, fault);
    }
}

static int* g_fault = nullptr;     // device flag: a warp gave up waiting for its TMA stage

int spmv_sell_tma(const SellOp& A, const SpmvArgs& a)
{
    Ctx& c = ctx();
    DB_ARG(a.x != nullptr && a.y != nullptr && a.x != a.y, "bad vectors");
    DB_ARG(a.xghost == nullptr, "the TMA kernel does not read split x");
    DB_ARG(a.epilogue == EPI_NONE || a.aux != nullptr, "epilogue needs aux vector");
    DB_ARG(a.dotv == nullptr || a.dot_out != nullptr || a.dd_out != nullptr, "fused dot needs an output scalar");
    DB_ARG(A.col16 != nullptr || A.col != nullptr, "operator holds no columns");
    if (A.rows == 0) return 0;
    if (!g_fault) {
        DB_CHECK(dev_alloc(&g_fault, 1));
        DB_CUDA(cudaMemset(g_fault, 0, sizeof(int)));
    }
    long long spc = (long long)opt("spmv.tma_slices_per_cta", 16.0);
    if (spc < 1) spc = 1;
    long long grid = ((long long)A.nslices + spc - 1) / spc;
    if (grid > kMaxReduceGrid / 8) grid = kMaxReduceGrid / 8;
    if (grid < 1) grid = 1;
    cudaStream_t st = a.stream ? a.stream : c.stream;
    SellView v{A.slice_ptr, A.col, A.val, A.perm, (long long)A.nslices, A.col16, A.cbase, A.wide};
    EpiArgs e = make_epi(a);
    int s;
    if (opt("spmv.exact", 1.0) != 0.0) s = launch_tma_epi<true>(v, a, e, (int)grid, st, g_fault);
    else s = launch_tma_epi<false>(v, a, e, (int)grid, st, g_fault);
    DB_CHECK(s);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

// 1 if a TMA stage wait ever timed out (checked by the tests and by bench.py after the run)
int spmv_tma_fault()
{
    if (!g_fault) return 0;
    int h = 0;
    cudaMemcpy(&h, g_fault, sizeof(int), cudaMemcpyDeviceToHost);
    return h;
}

}  // namespace db200

extern "C" int drops_b200_spmv_tma_fault(void) { return db200::spmv_tma_fault() | db200::spmv_pk_tma_fault(); }

// spmv_pk_tma.cu -- the dictionary-packed SpMV with its matrix stream staged by the TMA engine.
//
// ncu of the register-staged packed kernel (spmv_pk.cu, profiles/r01_notes.md): DRAM 50 %, L2 38 %,
// L1 54 %, issue 46 % -- nothing saturated; every warp alternates between waiting for its U words
// from DRAM and waiting for the U gathers they address, so the matrix stream is in flight only
// half of the time.  Here every WARP owns a contiguous run of slices and lane 0 streams their words
// with cp.async.bulk (1-D TMA copies, L2 evict-first) into a private ring of STAGES x R entry-rows
// (R x 128 B) in shared memory, one mbarrier per stage: the stream runs STAGES chunks ahead of the
// consumer and costs no registers, and the consumer only ever waits for its gathers.  4 B per
// entry make the ring cheap: 4 warps x 3 stages x 1 KB = 12 KB per CTA (the 12-byte stream needed
// 40 KB and lost more in L1 capacity than it gained, spmv_tma.cu).
//
// Arithmetic, padding rule, epilogues and the fused dot are those of spmv_pk.cu (bit-identical y):
// one thread per row, left-to-right sum, products rounded before the add, padded products dropped
// by a select.  The fused dot is the "late" form: partner fetched after the row sum, accumulator
// parked in shared memory.  perm == nullptr selects the row-ordered twin stream.
#include "spmv_impl.cuh"
#include "tma_util.cuh"

namespace db200 {

struct PkTmaView {
    const int64_t* slice_ptr;
    const uint32_t* pk;
    const double* dict;
    const int32_t* perm;        // nullptr: row-ordered stream (row = position)
    long long nslices;
    int ndict;
    int cbits;
};

// dynamic shared memory: ring[WARPS][STAGES][R*32] words | full[WARPS][STAGES] mbarriers | dictionary
template <int EPI, bool EXACT, int DOT, int WARPS, int STAGES, int R, int U, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_spmv_pk_tma(PkTmaView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e, int* __restrict__ fault)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t (*ring)[STAGES][R * 32] = reinterpret_cast<uint32_t (*)[STAGES][R * 32]>(smem_raw);
    unsigned long long (*full)[STAGES] =
        reinterpret_cast<unsigned long long (*)[STAGES]>(smem_raw + sizeof(uint32_t) * WARPS * STAGES * R * 32);
    const double* sdict = reinterpret_cast<const double*>(smem_raw + sizeof(uint32_t) * WARPS * STAGES * R * 32 +
                                                           sizeof(unsigned long long) * WARPS * STAGES);
    __shared__ dd acc_smem[DOT != 0 ? WARPS * 32 : 1];
    if (e.done != nullptr && *e.done != 0) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long s_begin = A.nslices * (long long)blockIdx.x / gridDim.x;
    const long long s_end   = A.nslices * (long long)(blockIdx.x + 1) / gridDim.x;
    const long long cnt = s_end - s_begin;
    long long s = s_begin + cnt * warp / WARPS;               // this warp's contiguous slices
    const long long ws_end = s_begin + cnt * (warp + 1) / WARPS;
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    const int cbits = A.cbits;
    const uint32_t cmask = (1u << cbits) - 1u;
    const uint32_t wpad = (uint32_t)A.ndict << cbits;

    if (lane == 0)
        for (int st = 0; st < STAGES; ++st) mbar_init(&full[warp][st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    long long J = 0, nchunks = 0;
    int64_t e0 = 0;
    unsigned long long policy = 0;
    if (s < ws_end) {
        e0 = A.slice_ptr[s];
        J = (A.slice_ptr[ws_end] - e0) >> 5;                  // entry-rows of this warp's stream
        nchunks = (J + R - 1) / R;
        policy = policy_evict_first();
    }
    auto issue = [&](long long t) {                           // lane 0: chunk t -> stage t % STAGES
        const int st = (int)(t % STAGES);
        const long long j0 = t * R;
        const unsigned rows = (unsigned)(J - j0 < R ? J - j0 : R);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[warp][st], rows * 128u);
        bulk_g2s(ring[warp][st], A.pk + e0 + j0 * 32, rows * 128u, &full[warp][st], policy);
    };
    // the stream starts before the CTA has its dictionary: the first chunks overlap the copy below
    if (lane == 0)
        for (long long t = 0; t < STAGES && t < nchunks; ++t) issue(t);
    {
        double* sd = const_cast<double*>(sdict);
        for (int i = threadIdx.x; i <= A.ndict; i += WARPS * 32) sd[i] = A.dict[i];
        if (DOT != 0) acc_smem[threadIdx.x] = dd_zero();
        __syncthreads();
    }

    if (s < ws_end) {
        auto width_of = [&](long long sl) -> int { return (int)((A.slice_ptr[sl + 1] - A.slice_ptr[sl]) >> 5); };
        auto row_of = [&](long long sl) -> int { return A.perm != nullptr ? A.perm[sl * 32 + lane] : (int)(sl * 32 + lane); };
        auto store_row = [&](long long sl, double sum) {
            const int row = row_of(sl);
            if (row >= 0) {
                const double out = apply_epilogue<EPI>(sum, EPI != EPI_NONE ? __ldg(e.aux + row) : 0.0, e, beta);
                y[row] = out;
                if (DOT != 0) acc_smem[threadIdx.x] = dd_add_prod(acc_smem[threadIdx.x], out, __ldg(e.dotv + row));
            }
        };
        int rem = width_of(s);
        double sum = 0.0;
        auto finish_slice = [&]() {
            store_row(s, sum);
            sum = 0.0;
            ++s;
            while (s < ws_end) {                              // slices without stored entries: y = epilogue(0)
                rem = width_of(s);
                if (rem > 0) break;
                store_row(s, 0.0);
                ++s;
            }
        };
        if (rem == 0) finish_slice();                          // leading slices without entries

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
            const uint32_t* W = ring[warp][stg];
            const int rows = (int)(J - t * R < R ? J - t * R : R);
            // batches of up to U entry-rows that end at the slice end or the stage end, whichever comes
            // first: one straight-line body (loads predicated on the batch length) and one place where
            // a slice is finished
            for (int jj = 0; jj < rows;) {
                int nb = rows - jj < U ? rows - jj : U;
                nb = rem < nb ? rem : nb;
                uint32_t w[U];
                double xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) w[u] = u < nb ? W[(jj + u) * 32 + lane] : wpad;
#pragma unroll
                for (int u = 0; u < U; ++u) xv[u] = __ldg(x + (w[u] & cmask));
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double v = sdict[w[u] >> cbits];
                    const double tt = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
                    sum = w[u] < wpad ? tt : sum;
                }
                jj += nb;
                rem -= nb;
                if (rem == 0) finish_slice();
            }
            __syncwarp();                                     // every lane is done reading this stage
            if (lane == 0 && t + STAGES < nchunks) issue(t + STAGES);
            if (++stg == STAGES) { stg = 0; parity ^= 1u; }
        }
    }
    if (DOT != 0) {
        const dd w = warp_reduce_dd(acc_smem[threadIdx.x]);
        if (lane == 0) e.partials[(size_t)blockIdx.x * WARPS + warp] = w;
    }
}

template <class K>
static int allow_smem_pk(K kernel, size_t bytes)
{
    DB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

template <int EPI, bool EXACT, int WARPS, int STAGES, int R, int U, int MINB>
static int launch_pk_tma_v(const PkTmaView& v, const SpmvArgs& a, const EpiArgs& e, long long nslices, cudaStream_t st,
                           int* fault)
{
    const size_t smem = (size_t)WARPS * STAGES * R * 128 + (size_t)WARPS * STAGES * 8 + ((size_t)v.ndict + 1) * 8;
    long long spc = (long long)opt("spmv.pktma_slices_per_cta", 4.0 * WARPS);
    if (spc < 1) spc = 1;
    long long grid = (nslices + spc - 1) / spc;
    if (grid * WARPS > kPartialStride) grid = kPartialStride / WARPS;
    if (grid < 1) grid = 1;
    static size_t allowed0 = 0, allowed2 = 0;
    if (a.dotv == nullptr) {
        auto k = k_spmv_pk_tma<EPI, EXACT, 0, WARPS, STAGES, R, U, MINB>;
        if (smem > allowed0) { DB_CHECK(allow_smem_pk(k, smem)); allowed0 = smem; }
        k<<<(unsigned)grid, WARPS * 32, smem, st>>>(v, a.x, a.y, e, fault);
        return 0;
    }
    auto k = k_spmv_pk_tma<EPI, EXACT, 2, WARPS, STAGES, R, U, MINB>;
    if (smem > allowed2) { DB_CHECK(allow_smem_pk(k, smem)); allowed2 = smem; }
    k<<<(unsigned)grid, WARPS * 32, smem, st>>>(v, a.x, a.y, e, fault);
    sum_warp_partials(grid * WARPS, e, st);
    return 0;
}

// "spmv.pktma_geometry": warps x stages x entry-rows per stage x gather batch (x resident CTAs the
// register allocation must allow)
template <int EPI, bool EXACT>
static int launch_pk_tma_geo(const PkTmaView& v, const SpmvArgs& a, const EpiArgs& e, long long ns, cudaStream_t st, int* fault)
{
    if (EPI == EPI_NONE && EXACT) {
        switch ((int)opt("spmv.pktma_geometry", 0.0)) {
            case 1: return launch_pk_tma_v<EPI_NONE, true, 4, 4, 8, 8, 12>(v, a, e, ns, st, fault);
            case 2: return launch_pk_tma_v<EPI_NONE, true, 4, 2, 16, 8, 12>(v, a, e, ns, st, fault);
            case 3: return launch_pk_tma_v<EPI_NONE, true, 4, 3, 8, 8, 8>(v, a, e, ns, st, fault);
            case 4: return launch_pk_tma_v<EPI_NONE, true, 8, 3, 8, 8, 5>(v, a, e, ns, st, fault);
            case 5: return launch_pk_tma_v<EPI_NONE, true, 4, 4, 16, 16, 8>(v, a, e, ns, st, fault);
            case 6: return launch_pk_tma_v<EPI_NONE, true, 4, 3, 16, 16, 10>(v, a, e, ns, st, fault);
            case 7: return launch_pk_tma_v<EPI_NONE, true, 4, 6, 4, 4, 12>(v, a, e, ns, st, fault);
            case 8: return launch_pk_tma_v<EPI_NONE, true, 2, 3, 8, 8, 16>(v, a, e, ns, st, fault);
            default: break;
        }
    }
    return launch_pk_tma_v<EPI, EXACT, 4, 3, 8, 8, 12>(v, a, e, ns, st, fault);
}

template <bool EXACT>
static int launch_pk_tma_epi(const PkTmaView& v, const SpmvArgs& a, const EpiArgs& e, long long ns, cudaStream_t st, int* fault)
{
    switch (a.epilogue) {
        case EPI_JACOBI: return launch_pk_tma_geo<EPI_JACOBI, EXACT>(v, a, e, ns, st, fault);
        case EPI_RESID:  return launch_pk_tma_geo<EPI_RESID, EXACT>(v, a, e, ns, st, fault);
        case EPI_AXMBY:  return launch_pk_tma_geo<EPI_AXMBY, EXACT>(v, a, e, ns, st, fault);
        default:         return launch_pk_tma_geo<EPI_NONE, EXACT>(v, a, e, ns, st, fault);
    }
}

static int* g_pk_fault = nullptr;   // device flag: a warp gave up waiting for its TMA stage

int spmv_sell_pk_tma(const SellOp& A, const SpmvArgs& a)
{
    Ctx& c = ctx();
    DB_ARG(A.pk != nullptr && A.dict != nullptr, "operator holds no packed stream");
    DB_ARG(!a.row_ordered || A.pkp != nullptr, "operator holds no row-ordered stream");
    DB_ARG(a.x != nullptr && a.y != nullptr && a.x != a.y, "bad vectors");
    DB_ARG(a.xghost == nullptr, "the packed stream has no split-x form");
    DB_ARG(a.epilogue == EPI_NONE || a.aux != nullptr, "epilogue needs aux vector");
    DB_ARG(a.dotv == nullptr || a.dot_out != nullptr || a.dd_out != nullptr, "fused dot needs an output scalar");
    DB_ARG(a.dotv != a.y, "the partner of the fused dot must not be the output vector");
    if (A.rows == 0) return 0;
    if (!g_pk_fault) {
        DB_CHECK(dev_alloc(&g_pk_fault, 1));
        DB_CUDA(cudaMemset(g_pk_fault, 0, sizeof(int)));
    }
    cudaStream_t st = a.stream ? a.stream : c.stream;
    PkTmaView v{A.slice_ptr, a.row_ordered ? A.pkp : A.pk, A.dict, a.row_ordered ? nullptr : A.perm, (long long)A.nslices,
                A.ndict, A.cbits};
    EpiArgs e = make_epi(a);
    int s;
    if (opt("spmv.exact", 1.0) != 0.0) s = launch_pk_tma_epi<true>(v, a, e, (long long)A.nslices, st, g_pk_fault);
    else s = launch_pk_tma_epi<false>(v, a, e, (long long)A.nslices, st, g_pk_fault);
    DB_CHECK(s);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

int spmv_pk_tma_fault()
{
    if (!g_pk_fault) return 0;
    int h = 0;
    cudaMemcpy(&h, g_pk_fault, sizeof(int), cudaMemcpyDeviceToHost);
    return h;
}

}  // namespace db200

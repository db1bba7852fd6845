// dist.cu -- the row-partitioned hot path across the GPUs of one box.
//
// Per PCG iteration and rank (num/krylovsolver.h:149-176 on the partition):
//   pack        boundary entries of p -> send buffer                         (k_pack, solver stream)
//   halo        ncclSend/ncclRecv per neighbour, one group                   (comm stream)
//   SpMV int.   rows that reference no ghost column, fused p.q partial       (solver stream, overlaps the halo)
//   SpMV bnd.   rows with ghost columns, after the halo arrived; partial chained (dd, unrounded)
//   allgather   1 double-double per rank; k_dist_scalars sums them in rank order -> lambda
//   update      x += alpha p, r -= alpha q, partials of |r|^2 and z.r        (same kernel as on one GPU)
//   allgather   2 double-doubles per rank; k_dist_scalars -> rho, resid, iter, stop flag
//   direction   p = z + beta p
// All scalars stay on the device; every rank sums the same partials in the same order, so the
// stop decision is bit-identical on all ranks and no rank needs the host inside the loop.
// The reference needs 3 MPI synchronisation points per iteration; this path
// has 2 tiny all-gathers and 1 neighbour exchange.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process -- torch's -- or
// the system one): libdrops_b200.so itself has no link-time dependency on it.
#include "dist.cuh"
#include "spmv.cuh"
#include "vecops.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_select.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>

namespace db200 {

// ------------------------------------------------------------------------------------------
// NCCL binding
// ------------------------------------------------------------------------------------------
struct NcclApi {
    bool tried = false, ok = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
};

static NcclApi& nccl()
{
    static NcclApi api;
    if (api.tried) return api;
    api.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already loaded (torch)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
#define DB_SYM(field, name) *(void**)(&api.field) = dlsym(h, name); if (!api.field) return api;
    DB_SYM(GetUniqueId, "ncclGetUniqueId") DB_SYM(CommInitRank, "ncclCommInitRank") DB_SYM(CommDestroy, "ncclCommDestroy")
    DB_SYM(GetErrorString, "ncclGetErrorString") DB_SYM(GroupStart, "ncclGroupStart") DB_SYM(GroupEnd, "ncclGroupEnd")
    DB_SYM(Send, "ncclSend") DB_SYM(Recv, "ncclRecv") DB_SYM(AllGather, "ncclAllGather") DB_SYM(AllReduce, "ncclAllReduce")
#undef DB_SYM
    api.ok = true;
    return api;
}

static int nccl_fail(ncclResult_t r, const char* what)
{
    set_error("NCCL error %d (%s) in %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
    return DROPS_B200_E_COMM;
}
#define DB_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_fail(r__, #call); } while (0)

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
struct OwnedRanges {            // device view of the plan's lookup tables
    const int64_t* gid0; const int64_t* len; const int64_t* local0; int nranges;
    const int64_t* ghost_sorted; const int32_t* ghost_slot; long long n_ghost;
    long long n_owned;
};

__device__ __forceinline__ long long local_of(const OwnedRanges& R, int64_t gid)
{
    int lo = 0, hi = R.nranges;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (R.gid0[mid] + R.len[mid] <= gid) lo = mid + 1; else hi = mid;
    }
    if (lo < R.nranges && R.gid0[lo] <= gid) return R.local0[lo] + (gid - R.gid0[lo]);
    long long a = 0, b = R.n_ghost;
    while (a < b) {
        const long long mid = (a + b) >> 1;
        if (R.ghost_sorted[mid] < gid) a = mid + 1; else b = mid;
    }
    if (a < R.n_ghost && R.ghost_sorted[a] == gid) return R.n_owned + R.ghost_slot[a];
    return -1;
}

// One thread per row: global column ids -> local int32 columns (entry order untouched), and the
// row's class: interior (no ghost column) or boundary.
__global__ void k_remap_rows(long long rows, const int64_t* __restrict__ rowptr, const int64_t* __restrict__ colgid,
                             OwnedRanges R, int32_t* __restrict__ col, char* __restrict__ f_int,
                             char* __restrict__ f_bnd, int* __restrict__ bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    bool bnd = false;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const long long l = local_of(R, colgid[k]);
        if (l < 0) { atomicExch(bad, 1); col[k] = 0; continue; }
        col[k] = (int32_t)l;
        bnd |= (l >= R.n_owned);
    }
    f_int[i] = bnd ? 0 : 1;
    f_bnd[i] = bnd ? 1 : 0;
}

__global__ void k_flag_not_owned(long long nnz, const int64_t* __restrict__ colgid, OwnedRanges R, char* __restrict__ flag)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int64_t gid = colgid[k];
    int lo = 0, hi = R.nranges;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (R.gid0[mid] + R.len[mid] <= gid) lo = mid + 1; else hi = mid;
    }
    flag[k] = (lo < R.nranges && R.gid0[lo] <= gid) ? 0 : 1;
}

__global__ void k_iota32(long long n, int32_t* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)i;
}

__global__ void k_pack(long long n, const int32_t* __restrict__ idx, const double* __restrict__ x, double* __restrict__ out,
                       const int* done)
{
    if (done != nullptr && *done != 0) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[idx[i]];
}

// position numbering of a partition: iperm[owned row] = base + pos, ro_perm[base + pos] = owned row
__global__ void k_ro_positions(long long npos, const int32_t* __restrict__ perm, int base, int32_t* __restrict__ iperm,
                               int32_t* __restrict__ ro_perm)
{
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int32_t r = perm[pos];
    ro_perm[base + pos] = r;
    if (r >= 0) iperm[r] = (int32_t)(base + pos);
}
// boundary operator twin: owned columns -> positions, ghost columns -> npos + slot; rows -> positions
__global__ void k_ro_boundary(long long padded, const int32_t* __restrict__ col, const int32_t* __restrict__ iperm,
                              int n_owned, int npos, int32_t* __restrict__ col_ro)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    const int32_t c = col[i];
    col_ro[i] = c < 0 ? -1 : (c < n_owned ? iperm[c] : npos + (c - n_owned));
}
__global__ void k_ro_perm_shift(long long npos, const int32_t* __restrict__ perm, int base, int32_t* __restrict__ out)
{
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos < npos) out[pos] = perm[pos] >= 0 ? (int32_t)(base + pos) : -1;
}
__global__ void k_gather_i32(long long n, const int32_t* __restrict__ idx, const int32_t* __restrict__ src, int32_t* __restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_gather_flags(long long n, const int32_t* __restrict__ order, const char* __restrict__ a, const char* __restrict__ b,
                               char* __restrict__ ao, char* __restrict__ bo)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ao[i] = a[order[i]]; bo[i] = b[order[i]]; }
}

// Line segments (reorder.cu) stay whole: a segment with a boundary row goes to the boundary operator with all its 32
// rows (an interior row there only reads owned columns), so the segments behind it keep their 32-row alignment in the
// interior operator.  One warp per segment.
__global__ void k_segment_flags(long long nseg, char* __restrict__ fi, char* __restrict__ fb)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg * 32) return;                     // nseg * 32 is a multiple of the warp size: whole warps leave
    const bool any_bnd = __any_sync(0xffffffffu, fb[i] != 0);
    if (any_bnd && fi[i] != 0) { fi[i] = 0; fb[i] = 1; }
}

// ---- peer-to-peer halo ------------------------------------------------------------------------
// pack + push: entry i of my send list goes straight into the ghost buffer of the rank that needs it.
// Every block fences its stores system-wide and takes a ticket; the block that draws the last ticket
// raises my sequence number in the flag arrays of the receivers -- by then all entries are visible.
__global__ void __launch_bounds__(256)
k_pack_push(long long n_send, const int32_t* __restrict__ idx, const double* __restrict__ x,
            const size_t* __restrict__ send_off, int world, double* const* __restrict__ peer_dst,
            unsigned* const* __restrict__ peer_flag, unsigned seq, unsigned* ticket, const int* done)
{
    if (done != nullptr && *done != 0) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_send) {
        int q = 0;
        while (q + 1 < world && (long long)send_off[q + 1] <= i) ++q;
        peer_dst[q][i - (long long)send_off[q]] = x[idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(ticket, 1u) == gridDim.x - 1) {
        *ticket = 0u;
        __threadfence_system();
        for (int q = 0; q < world; ++q)
            if (send_off[q + 1] > send_off[q]) *(volatile unsigned*)peer_flag[q] = seq;
        __threadfence_system();
    }
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Consumer side: one lane per sender spins (bounded) until that sender's sequence number has arrived.
// stat (nullable): {nanoseconds this rank spent in the wait, number of waits} (drops_b200_dist_matrix_wait_stats).
__global__ void k_halo_wait(int world, const unsigned* flags, const unsigned char* __restrict__ expect, unsigned seq,
                            const int* done, int* fault, unsigned long long* stat)
{
    if (done != nullptr && *done != 0) return;
    const unsigned long long t0 = (stat != nullptr && threadIdx.x == 0) ? global_ns() : 0ull;
    const int q = threadIdx.x;
    if (q < world && expect[q]) {
        unsigned long long spins = 0;
        while ((int)(*(volatile const unsigned*)(flags + q) - seq) < 0) {
            __nanosleep(64);
            if (++spins > (1ull << 26)) { atomicExch(fault, 1); break; }       // never hang the GPU
        }
    }
    __threadfence_system();
    if (stat != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) { stat[0] += global_ns() - t0; stat[1] += 1ull; }
    }
}

// peer-to-peer all-gather of reduction partials: my nq partials go into slot `me` of the current parity half
// of EVERY rank's buffer, then my sequence number into every rank's flag array
__global__ void k_push_partials(int nq, int me, int world, const dd* __restrict__ local, dd* const* __restrict__ peer_buf,
                                unsigned* const* __restrict__ peer_flag, unsigned seq, const int* done)
{
    if (done != nullptr && *done != 0) return;
    const int t = threadIdx.x;
    if (t < world * nq) {
        const int q = t / nq, j = t - q * nq;
        dd v = local[j];
        dd* dst = peer_buf[q] + (size_t)(seq & 1u) * world * kMaxQ + (size_t)me * nq + j;
        *(volatile double*)&dst->hi = v.hi;
        *(volatile double*)&dst->lo = v.lo;
    }
    __threadfence_system();
    __syncthreads();
    if (t < world) { *(volatile unsigned*)(peer_flag[t] + me) = seq; }
    __threadfence_system();
}

enum ScalarPhase { PH_DOT = 0, PH_LAMBDA = 1, PH_SETUP = 2, PH_UPDATE = 3, PH_NORM = 4 };

// Sums the ranks' partials in rank order (the same on every rank) and does the scalar bookkeeping
// that the last block of the fused kernels does on a single GPU.
__global__ void k_dist_scalars(int phase, SolverScalars* S, double* hist, const dd* gathered, int world,
                               double* out, const unsigned* flags, unsigned seq, int* fault, unsigned long long* stat)
{
    if ((phase == PH_LAMBDA || phase == PH_UPDATE) && S->done) return;
    if (flags != nullptr) {                          // peer path: wait (bounded) for every rank's sequence number
        const unsigned long long t0 = (stat != nullptr && threadIdx.x == 0) ? global_ns() : 0ull;
        const int q = threadIdx.x;
        if (q < world) {
            unsigned long long spins = 0;
            while ((int)(*(volatile const unsigned*)(flags + q) - seq) < 0) {
                __nanosleep(32);
                if (++spins > (1ull << 26)) { atomicExch(fault, 1); break; }
            }
        }
        __threadfence_system();
        __syncthreads();
        if (stat != nullptr && threadIdx.x == 0) {   // slots: 2 = the p.Ap reduction, 4 = the (rho, residual) pair, 6 = others
            const int slot = phase == PH_LAMBDA ? 2 : (phase == PH_UPDATE ? 4 : 6);
            stat[slot] += global_ns() - t0; stat[slot + 1] += 1ull;
        }
        // a peer never showed up: its partial is stale.  Stop iterating on this rank (every later kernel of the
        // solve returns on S->done) and leave the sticky fault flag for the host, which turns it into
        // DROPS_B200_E_COMM (check_p2p_faults) instead of a plausible-looking residual.
        if (*(volatile int*)fault != 0) {
            if (threadIdx.x == 0 && S != nullptr) { S->done = 1; S->converged = 0; }
            return;
        }
    }
    if (threadIdx.x != 0) return;
    const int nq = phase == PH_SETUP ? 3 : (phase == PH_UPDATE ? 2 : 1);
    dd acc[3] = {dd_zero(), dd_zero(), dd_zero()};
    for (int r = 0; r < world; ++r)
        for (int q = 0; q < nq; ++q) {
            const volatile dd* g = gathered + r * nq + q;
            dd t; t.hi = g->hi; t.lo = g->lo;
            acc[q] = dd_add(acc[q], t);
        }
    if (phase == PH_DOT) *out = dd_round(acc[0]);
    else if (phase == PH_NORM) *out = sqrt(dd_round(acc[0]));
    else if (phase == PH_LAMBDA) S->lambda = dd_round(acc[0]);
    else if (phase == PH_SETUP) pcg_setup_finish(S, hist, dd_round(acc[0]), dd_round(acc[1]), dd_round(acc[2]));
    else pcg_update_finish(S, hist, dd_round(acc[0]), dd_round(acc[1]));
}

// dot partial of a partition whose boundary rows ran beside the interior rows: boundary + interior, the order of the
// chained form (store_dot: total = dd_add(total, *dd_in) with total = the boundary rows' sum)
__global__ void k_combine_dd(dd* out, const dd* bnd, const dd* interior, const int* done)
{
    if (done != nullptr && *done != 0) return;
    dd t = *bnd;
    if (interior != nullptr) t = dd_add(t, *interior);
    *out = t;
}

__global__ void k_zero_dd(dd* p, int n)
{
    if (threadIdx.x < n) p[threadIdx.x] = dd_zero();
}

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
static inline int blocks_for(long long n, int threads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

struct DevTables {     // device copies of the plan's lookup tables
    int64_t *gid0 = nullptr, *len = nullptr, *local0 = nullptr, *ghost_sorted = nullptr;
    int32_t* ghost_slot = nullptr;
    OwnedRanges view{};
    void release() { dev_free(gid0); dev_free(len); dev_free(local0); dev_free(ghost_sorted); dev_free(ghost_slot); }
};

static int upload_ranges(const std::vector<drops_b200_exchange_s::Range>& ranges, size_t n_owned,
                         const std::vector<int64_t>* ghost_sorted, const std::vector<int32_t>* ghost_slot, DevTables* T)
{
    cudaStream_t st = ctx().stream;
    const size_t nr = ranges.size();
    std::vector<int64_t> a(nr), b(nr), c(nr);
    for (size_t i = 0; i < nr; ++i) { a[i] = ranges[i].gid0; b[i] = ranges[i].len; c[i] = ranges[i].local0; }
    DB_CHECK(dev_alloc(&T->gid0, nr)); DB_CHECK(dev_alloc(&T->len, nr)); DB_CHECK(dev_alloc(&T->local0, nr));
    if (nr) {
        DB_CUDA(cudaMemcpyAsync(T->gid0, a.data(), nr * 8, cudaMemcpyHostToDevice, st));
        DB_CUDA(cudaMemcpyAsync(T->len, b.data(), nr * 8, cudaMemcpyHostToDevice, st));
        DB_CUDA(cudaMemcpyAsync(T->local0, c.data(), nr * 8, cudaMemcpyHostToDevice, st));
    }
    const size_t ng = ghost_sorted ? ghost_sorted->size() : 0;
    DB_CHECK(dev_alloc(&T->ghost_sorted, ng)); DB_CHECK(dev_alloc(&T->ghost_slot, ng));
    if (ng) {
        DB_CUDA(cudaMemcpyAsync(T->ghost_sorted, ghost_sorted->data(), ng * 8, cudaMemcpyHostToDevice, st));
        DB_CUDA(cudaMemcpyAsync(T->ghost_slot, ghost_slot->data(), ng * 4, cudaMemcpyHostToDevice, st));
    }
    DB_CUDA(cudaStreamSynchronize(st));    // the host vectors above go out of scope
    T->view = OwnedRanges{T->gid0, T->len, T->local0, (int)nr, T->ghost_sorted, T->ghost_slot, (long long)ng, (long long)n_owned};
    return 0;
}

// Sorted unique non-owned column ids of device-resident rows (global ids) -> host vector.
int find_ghosts_device(size_t n_owned, const int64_t* owned_host_sorted, size_t nnz, const int64_t* d_colgid,
                       std::vector<int64_t>* out)
{
    out->clear();
    if (nnz == 0) return 0;
    cudaStream_t st = ctx().stream;
    std::vector<drops_b200_exchange_s::Range> ranges;
    for (size_t i = 0; i < n_owned;) {
        size_t j = i + 1;
        while (j < n_owned && owned_host_sorted[j] == owned_host_sorted[j - 1] + 1) ++j;
        ranges.push_back({owned_host_sorted[i], (int64_t)(j - i), (int64_t)i});
        i = j;
    }
    DevTables T;
    int s = upload_ranges(ranges, n_owned, nullptr, nullptr, &T);
    char* flag = nullptr; int64_t* cand = nullptr; int64_t* cand2 = nullptr; long long* d_cnt = nullptr; void* tmp = nullptr;
    auto cleanup = [&]() { T.release(); dev_free(flag); dev_free(cand); dev_free(cand2); dev_free(d_cnt); dev_free(tmp); };
#define FG_CHECK(call) do { s = (call); if (s < 0) { cleanup(); return s; } } while (0)
#define FG_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
    if (s < 0) { cleanup(); return s; }
    FG_CHECK(dev_alloc(&flag, nnz));
    FG_CHECK(dev_alloc(&d_cnt, 1));
    k_flag_not_owned<<<blocks_for((long long)nnz, 256), 256, 0, st>>>((long long)nnz, d_colgid, T.view, flag);
    count_launch();
    // the candidates are a thin set (entries of boundary rows): count, then select exactly that many
    long long ncand = 0;
    size_t tb = 0;
    FG_CUDA(cub::DeviceReduce::Sum(nullptr, tb, (const unsigned char*)flag, d_cnt, (long long)nnz, st));
    FG_CHECK(dev_alloc((char**)&tmp, tb));
    FG_CUDA(cub::DeviceReduce::Sum(tmp, tb, (const unsigned char*)flag, d_cnt, (long long)nnz, st));
    FG_CUDA(cudaMemcpyAsync(&ncand, d_cnt, sizeof(long long), cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaStreamSynchronize(st));
    dev_free(tmp); tmp = nullptr;
    if (ncand > 0) {
        FG_CHECK(dev_alloc(&cand, (size_t)ncand));
        tb = 0;
        FG_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, d_colgid, flag, cand, d_cnt, (long long)nnz, st));
        FG_CHECK(dev_alloc((char**)&tmp, tb));
        FG_CUDA(cub::DeviceSelect::Flagged(tmp, tb, d_colgid, flag, cand, d_cnt, (long long)nnz, st));
        count_launch(2);
        FG_CUDA(cudaStreamSynchronize(st));
    }
    dev_free(flag); flag = nullptr; dev_free(tmp); tmp = nullptr;
    if (ncand > 0) {
        FG_CHECK(dev_alloc(&cand2, (size_t)ncand));
        tb = 0;
        FG_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, cand, cand2, ncand, 0, 64, st));
        FG_CHECK(dev_alloc((char**)&tmp, tb));
        FG_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, cand, cand2, ncand, 0, 64, st));
        dev_free(tmp); tmp = nullptr;
        tb = 0;
        FG_CUDA(cub::DeviceSelect::Unique(nullptr, tb, cand2, cand, d_cnt, ncand, st));
        FG_CHECK(dev_alloc((char**)&tmp, tb));
        FG_CUDA(cub::DeviceSelect::Unique(tmp, tb, cand2, cand, d_cnt, ncand, st));
        count_launch(4);
        long long nuniq = 0;
        FG_CUDA(cudaMemcpyAsync(&nuniq, d_cnt, sizeof(long long), cudaMemcpyDeviceToHost, st));
        FG_CUDA(cudaStreamSynchronize(st));
        out->resize((size_t)nuniq);
        if (nuniq) FG_CUDA(cudaMemcpyAsync(out->data(), cand, (size_t)nuniq * 8, cudaMemcpyDeviceToHost, st));
        FG_CUDA(cudaStreamSynchronize(st));
    }
    cleanup();
#undef FG_CHECK
#undef FG_CUDA
    return 0;
}

static int p2p_setup_nccl(drops_b200_dist_matrix_s* M);

static void free_dist_matrix(drops_b200_dist_matrix_s* M)
{
    if (!M) return;
    if (ctx().inited) { cudaStreamSynchronize(ctx().stream); if (M->comm && M->comm->comm_stream) cudaStreamSynchronize(M->comm->comm_stream); }
    if (M->comm && M->comm->kind == COMM_LOCAL && M->rank < (int)M->comm->parts.size() && M->comm->parts[(size_t)M->rank] == M)
        M->comm->parts[(size_t)M->rank] = nullptr;
    free_sell(M->sell_int); free_sell(M->sell_bnd); free_sell(M->sell_bnd_ro);
    dev_free(M->ro_perm); dev_free(M->diag_ro); dev_free(M->d_send_idx_ro); dev_free(M->diag_ro_idx);
    for (void* m : M->ipc_mapped) cudaIpcCloseMemHandle(m);
    dev_free(M->p2p_ticket); dev_free(M->p2p_fault); dev_free(M->wait_stats); dev_free(M->d_peer_dst); dev_free(M->d_peer_flag);
    dev_free(M->d_send_off); dev_free(M->d_expect);
    dev_free(M->diag); dev_free(M->ghost); dev_free(M->sendbuf); dev_free(M->d_send_idx);
    for (double* w : M->work) dev_free(w);
    dev_free(M->d_scalars); dev_free(M->d_hist); dev_free(M->d_small);
    delete M;
}

static double* dwork(drops_b200_dist_matrix_s* M, int i)
{
    if ((int)M->work.size() <= i) M->work.resize((size_t)i + 1, nullptr);
    if (!M->work[(size_t)i] && dev_alloc(&M->work[(size_t)i], std::max(M->n_owned, M->npos)) < 0) return nullptr;
    return M->work[(size_t)i];
}

__global__ void k_extract_diag_local(long long rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                     const double* __restrict__ val, double* __restrict__ diag, int* __restrict__ missing)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int64_t k = rowptr[i];
    const int64_t e = rowptr[i + 1];
    while (k < e && col[k] != (int32_t)i) ++k;
    if (k < e) diag[i] = val[k];
    else { diag[i] = 0.0; atomicExch(missing, 1); }
}

// Builds the partition operator from device-resident rows with global column ids: the rank-local part
// (every failure exit in here is this rank's alone -- dist_matrix_build agrees on the outcome across the
// ranks before anything collective happens).
static int dist_matrix_build_local(drops_b200_comm_s* comm, drops_b200_exchange_s* E, size_t nnz, const int64_t* d_rowptr,
                                   const int64_t* d_colgid, const double* d_val, uint64_t version, drops_b200_dist_matrix_s** out)
{
    *out = nullptr;
    DB_ARG(comm != nullptr && E != nullptr, "null comm/exchange");
    DB_ARG(E->world == comm->world, "plan and communicator disagree on the number of ranks");
    DB_ARG(comm->kind == COMM_LOCAL || E->rank == comm->rank, "plan was built for another rank");
    for (int q = 0; q < E->world; ++q)
        DB_ARG(E->send_set[(size_t)q], "halo plan is not ready: exchange_set_send is missing for a peer");
    cudaStream_t st = ctx().stream;
    const size_t n = E->owned.size(), ng = E->ghost_gid.size();
    drops_b200_dist_matrix_s* M = new drops_b200_dist_matrix_s;
    M->comm = comm; M->rank = E->rank; M->world = E->world;
    M->n_owned = n; M->n_ghost = ng; M->nnz = nnz; M->version = version;
    DevTables T;
    int32_t* col = nullptr; char* f_int = nullptr; char* f_bnd = nullptr; int* d_flag = nullptr;
    int32_t *rows_all = nullptr, *rows_int = nullptr, *rows_bnd = nullptr; long long* d_cnt = nullptr; void* tmp = nullptr;
    int s = 0;
    auto cleanup = [&]() {
        T.release(); dev_free(col); dev_free(f_int); dev_free(f_bnd); dev_free(d_flag); dev_free(rows_all);
        dev_free(rows_int); dev_free(rows_bnd); dev_free(d_cnt); dev_free(tmp);
    };
#define DM_CHECK(call) do { s = (call); if (s < 0) { cleanup(); free_dist_matrix(M); return s; } } while (0)
#define DM_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); free_dist_matrix(M); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
    DM_CHECK(upload_ranges(E->ranges, n, &E->ghost_sorted, &E->ghost_sorted_slot, &T));
    DM_CHECK(dev_alloc(&col, nnz));
    DM_CHECK(dev_alloc(&f_int, n)); DM_CHECK(dev_alloc(&f_bnd, n));
    DM_CHECK(dev_alloc(&d_flag, 2));
    DM_CUDA(cudaMemsetAsync(d_flag, 0, 2 * sizeof(int), st));
    if (n > 0) {
        k_remap_rows<<<blocks_for((long long)n, 128), 128, 0, st>>>((long long)n, d_rowptr, d_colgid, T.view, col, f_int, f_bnd, d_flag);
        count_launch();
        DM_CUDA(cudaGetLastError());
    }
    // diagonal (local column of an owned DOF == its local row)
    DM_CHECK(dev_alloc(&M->diag, n));
    if (n > 0) {
        k_extract_diag_local<<<blocks_for((long long)n, 256), 256, 0, st>>>((long long)n, d_rowptr, col, d_val, M->diag, d_flag + 1);
        count_launch();
    }
    int hflag[2] = {0, 0};
    DM_CUDA(cudaMemcpyAsync(hflag, d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    DM_CUDA(cudaStreamSynchronize(st));
    if (hflag[0]) {
        cleanup(); free_dist_matrix(M);
        set_error("dist_matrix: a column id is neither owned nor in the plan's ghost list");
        return DROPS_B200_E_ARG;
    }
    M->diag_ok = (hflag[1] == 0);
    // interior / boundary row lists
    DM_CHECK(dev_alloc(&rows_all, n)); DM_CHECK(dev_alloc(&rows_int, n)); DM_CHECK(dev_alloc(&rows_bnd, n));
    DM_CHECK(dev_alloc(&d_cnt, 2));
    long long cnt[2] = {0, 0};
    bool packs = false;
    if (n > 0) {
        // rows in blocks of the partition's own graph (reorder.cu) when it is large enough to profit
        int32_t* order = nullptr;
        size_t n_segments = 0;
        {
            CsrDev L;
            L.rows = n; L.cols = n + ng; L.nnz = nnz;
            L.rowptr = const_cast<int64_t*>(d_rowptr); L.col = col; L.val = const_cast<double*>(d_val);
            packs = false;               // graph blocks only pay together with the row-ordered twins
            DM_CHECK(will_pack_one_word(d_val, nnz, n + ng + 2 * 4096 + 64, &packs));
            if (packs && opt("dist.reorder", 1.0) != 0.0) DM_CHECK(compute_block_order(L, &order, &n_segments));
        }
        if (order) {
            char *fi = nullptr, *fb = nullptr;
            int s2 = dev_alloc(&fi, n);
            if (s2 == 0) s2 = dev_alloc(&fb, n);
            if (s2 == 0) {
                k_gather_flags<<<blocks_for((long long)n, 256), 256, 0, st>>>((long long)n, order, f_int, f_bnd, fi, fb);
                count_launch();
                if (n_segments > 0) {
                    k_segment_flags<<<blocks_for((long long)n_segments * 32, 256), 256, 0, st>>>((long long)n_segments, fi, fb);
                    count_launch();
                }
                cudaError_t e2 = cudaMemcpyAsync(rows_all, order, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
                if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(f_int, fi, n, cudaMemcpyDeviceToDevice, st);
                if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(f_bnd, fb, n, cudaMemcpyDeviceToDevice, st);
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
                if (e2 != cudaSuccess) s2 = cuda_fail(e2, "block order of a partition", __FILE__, __LINE__);
                else M->block_ordered = true;
            }
            dev_free(fi); dev_free(fb); dev_free(order);
            DM_CHECK(s2);
        } else {
            k_iota32<<<blocks_for((long long)n, 256), 256, 0, st>>>((long long)n, rows_all);
            count_launch();
            DM_CUDA(cudaGetLastError());
        }
        size_t tb = 0;
        DM_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, rows_all, f_int, rows_int, d_cnt, (long long)n, st));
        DM_CHECK(dev_alloc((char**)&tmp, tb));
        DM_CUDA(cub::DeviceSelect::Flagged(tmp, tb, rows_all, f_int, rows_int, d_cnt, (long long)n, st));
        DM_CUDA(cub::DeviceSelect::Flagged(tmp, tb, rows_all, f_bnd, rows_bnd, d_cnt + 1, (long long)n, st));
        count_launch(3);
        DM_CUDA(cudaMemcpyAsync(cnt, d_cnt, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        DM_CUDA(cudaStreamSynchronize(st));
    }
    M->n_int = (size_t)cnt[0]; M->n_bnd = (size_t)cnt[1];
    CsrDev A;   // aliases the caller's arrays; only the int32 columns are ours
    A.rows = n; A.cols = n + ng; A.nnz = nnz;
    A.rowptr = const_cast<int64_t*>(d_rowptr); A.col = col; A.val = const_cast<double*>(d_val);
    M->device_bytes = n * 8;
    {
        // head room in the column field of the packed words: the row-ordered twin stores positions, and the
        // positions of a partition run up to (windows of the interior + windows of the boundary operator)
        CsrDev Ai = A;
        Ai.cols = n + ng + 2 * 4096 + 64;
        DM_CHECK(build_sell(Ai, M->sell_int, &M->device_bytes, rows_int, M->n_int, false, packs));
        M->sell_int.cols = A.cols;
    }
    DM_CHECK(build_sell(A, M->sell_bnd, &M->device_bytes, rows_bnd, M->n_bnd));
    // halo buffers
    M->recv_off = E->recv_off;
    M->send_off.assign((size_t)E->world + 1, 0);
    std::vector<int32_t> flat;
    for (int q = 0; q < E->world; ++q) {
        flat.insert(flat.end(), E->send_idx[(size_t)q].begin(), E->send_idx[(size_t)q].end());
        M->send_off[(size_t)q + 1] = flat.size();
    }
    M->n_send = flat.size();
    DM_CHECK(dev_alloc(&M->d_send_idx, M->n_send));
    DM_CHECK(dev_alloc(&M->sendbuf, M->n_send));
    {
        // ghost entries and, behind them, the arrival flags of the peer-to-peer halo: ONE allocation of at
        // least 2 MB, so that it is a block of its own and one CUDA IPC handle maps both into the peers
        const size_t flag_off = (ng * sizeof(double) + 255) / 256 * 256;
        const size_t bytes = std::max<size_t>(flag_off + (size_t)std::max(E->world, 1) * sizeof(unsigned), (size_t)2 << 20);
        char* raw = nullptr;
        DM_CHECK(dev_alloc(&raw, bytes));
        M->ghost = reinterpret_cast<double*>(raw);
        M->halo_flags = reinterpret_cast<unsigned*>(raw + flag_off);
        M->halo_flag_offset = flag_off;
        DM_CUDA(cudaMemsetAsync(M->halo_flags, 0, (size_t)std::max(E->world, 1) * sizeof(unsigned), st));
    }
    if (M->n_send) DM_CUDA(cudaMemcpyAsync(M->d_send_idx, flat.data(), M->n_send * 4, cudaMemcpyHostToDevice, st));
    DM_CUDA(cudaMemsetAsync(M->ghost, 0, std::max<size_t>(ng, 1) * 8, st));
    DM_CUDA(cudaStreamSynchronize(st));
    M->device_bytes += (M->n_send + ng) * 8 + M->n_send * 4;
    // ---- row-ordered twins: PCG iterates in the numbering of the operators' own slices ----
    if (opt("dist.reorder", 1.0) != 0.0 && n > 0 && M->sell_int.pk != nullptr && M->n_int > 0) {
        const size_t npi = M->sell_int.nslices * 32, npb = M->sell_bnd.nslices * 32, npos = npi + npb;
        int32_t* iperm = nullptr;
        DM_CHECK(dev_alloc(&iperm, n));
        int s3 = dev_alloc(&M->ro_perm, npos);
        if (s3 == 0) {
            k_ro_positions<<<blocks_for((long long)npi, 256), 256, 0, st>>>((long long)npi, M->sell_int.perm, 0, iperm, M->ro_perm);
            k_ro_positions<<<blocks_for((long long)npb, 256), 256, 0, st>>>((long long)npb, M->sell_bnd.perm, (int)npi, iperm, M->ro_perm);
            count_launch(2);
            s3 = pack_reorder(M->sell_int, iperm, npos);
        }
        if (s3 == 0 && M->sell_int.pkp != nullptr) {
            SellOp& B = M->sell_bnd; SellOp& R = M->sell_bnd_ro;
            R.rows = B.rows; R.cols = npos + ng; R.nslices = B.nslices; R.padded = B.padded; R.full_rows = false;
            s3 = dev_alloc(&R.slice_ptr, B.nslices + 1);
            if (s3 == 0) s3 = dev_alloc(&R.col, B.padded);
            if (s3 == 0) s3 = dev_alloc(&R.val, B.padded);
            if (s3 == 0) s3 = dev_alloc(&R.perm, npb);
            if (s3 == 0) s3 = dev_alloc(&M->diag_ro, npos);
            if (s3 == 0) s3 = dev_alloc(&M->d_send_idx_ro, M->n_send);
            if (s3 == 0) {
                cudaMemcpyAsync(R.slice_ptr, B.slice_ptr, (B.nslices + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st);
                if (B.padded) {
                    cudaMemcpyAsync(R.val, B.val, B.padded * sizeof(double), cudaMemcpyDeviceToDevice, st);
                    k_ro_boundary<<<blocks_for((long long)B.padded, 256), 256, 0, st>>>((long long)B.padded, B.col, iperm, (int)n, (int)npos, R.col);
                }
                k_ro_perm_shift<<<blocks_for((long long)npb, 256), 256, 0, st>>>((long long)npb, B.perm, (int)npi, R.perm);
                if (M->n_send) k_gather_i32<<<blocks_for((long long)M->n_send, 256), 256, 0, st>>>((long long)M->n_send, M->d_send_idx, iperm, M->d_send_idx_ro);
                count_launch(3);
                s3 = vec_gather_perm(npos, M->ro_perm, M->diag, M->diag_ro, 1.0);
                // constant-coefficient partitions: the permuted diagonal as a byte index (update kernel of dist_pcg)
                if (s3 == 0 && M->diag_ok && M->sell_int.ndict_host > 0)
                    s3 = build_value_index8(M->diag_ro, npos, kDiagDict, &M->diag_ro_idx, M->diag_ro_values, &M->diag_ro_count);
            }
            if (s3 == 0) {
                cudaError_t e3 = cudaStreamSynchronize(st);
                if (e3 != cudaSuccess) s3 = cuda_fail(e3, "row-ordered twins", __FILE__, __LINE__);
            }
            if (s3 == 0) {
                M->npos = npos; M->npos_int = npi;
                M->device_bytes += M->sell_int.padded * 4 + B.padded * 12 + npos * 12 + M->n_send * 4;
            }
        }
        dev_free(iperm);
        if (s3 < 0) { cleanup(); free_dist_matrix(M); return s3; }
        if (M->npos == 0) {      // twins incomplete (bits do not fit): drop them
            dev_free(M->sell_int.pkp); M->sell_int.pkp = nullptr; M->sell_int.npos = 0;
            free_sell(M->sell_bnd_ro); dev_free(M->ro_perm); M->ro_perm = nullptr;
        }
    }
    // the interior rows never read ghost columns: once their packed stream exists its 12-byte form is dead weight
    release_general(M->sell_int, &M->device_bytes);
    cleanup();
#undef DM_CHECK
#undef DM_CUDA
    *out = M;
    return 0;
}

int dist_matrix_build(drops_b200_comm_s* comm, drops_b200_exchange_s* E, size_t nnz, const int64_t* d_rowptr,
                      const int64_t* d_colgid, const double* d_val, uint64_t version, drops_b200_dist_matrix_s** out)
{
    drops_b200_dist_matrix_s* M = nullptr;
    const int s = dist_matrix_build_local(comm, E, nnz, d_rowptr, d_colgid, d_val, version, &M);
    *out = nullptr;
    if (comm == nullptr) return s;
    if (comm->kind == COMM_LOCAL) {
        if (s < 0) return s;
        if ((int)comm->parts.size() < comm->world) comm->parts.resize((size_t)comm->world, nullptr);
        comm->parts[(size_t)M->rank] = M;
        *out = M;
        return 0;
    }
    // NCCL transport: the peer mapping below is collective (an all-gather of IPC handles).  A rank whose local build
    // failed (unknown column id, out of memory) must not leave the others blocked in it: agree on the outcome first.
    double bad = s < 0 ? 1.0 : 0.0;
    const int sa = drops_b200_comm_allreduce_max(comm, &bad);
    if (s < 0) return s;                           // this rank's own error text stands
    if (sa < 0) { free_dist_matrix(M); return sa; }
    if (bad != 0.0) {
        free_dist_matrix(M);
        set_error("dist_matrix: the build of the partition failed on another rank");
        return DROPS_B200_E_COMM;
    }
    const int ps = p2p_setup_nccl(M);              // collective; failure to map peers is not an error
    if (ps < 0) { free_dist_matrix(M); return ps; }
    *out = M;
    return 0;
}

// ------------------------------------------------------------------------------------------
// collective building blocks
// ------------------------------------------------------------------------------------------
struct PartArgs {
    const double* x = nullptr; double* y = nullptr;
    int epilogue = EPI_NONE; const double* aux = nullptr; double omega = 1.0;
    double beta_val = 0.0;              // EPI_AXMBY: y = A x - beta_val * aux
    const double* dotv = nullptr;       // fused dot partner (local partial -> gathered slot)
    const int* done = nullptr;
    bool ro = false;                    // x, y, aux, dotv are in the partition's row-ordered numbering (npos entries)
    bool p2p = false;                   // the exchange may use the peer path (solver loops only, see halo_exchange)
};

static int check_parts(int np, drops_b200_dist_matrix_s* const* P)
{
    DB_ARG(np >= 1 && P != nullptr && P[0] != nullptr, "no partitions");
    drops_b200_comm_s* comm = P[0]->comm;
    DB_ARG(comm != nullptr, "partition without communicator");
    if (comm->kind == COMM_NCCL) DB_ARG(np == 1, "an NCCL communicator drives exactly one partition per process");
    else DB_ARG(np == comm->world, "a local communicator drives all `world` partitions in one call");
    for (int p = 0; p < np; ++p) {
        DB_ARG(P[p] != nullptr && P[p]->comm == comm, "partitions of different communicators");
        DB_ARG(comm->kind == COMM_NCCL || P[p]->rank == p, "partitions must be ordered by rank");
    }
    return 0;
}

// slot of partition p's reduction partials in the rank-major gather buffer
static inline dd* gather_slot(drops_b200_dist_matrix_s* M, int nq) { return M->comm->gathered + (size_t)M->rank * nq; }

static int allgather_partials(drops_b200_comm_s* comm, int nq, const int* const* done_of = nullptr)
{
    comm->red_last_p2p = false;
    if (comm->world > 1 && comm->p2p == 1 && opt("dist.p2p_reduce", 1.0) != 0.0 && comm->world * nq <= 1024) {
        ScopedTiming t("allgather");
        ++comm->red_seq;
        const int threads = std::max(32, (comm->world * nq + 31) / 32 * 32);
        const int first = comm->kind == COMM_LOCAL ? 0 : comm->rank, last = comm->kind == COMM_LOCAL ? comm->world : comm->rank + 1;
        for (int me = first; me < last; ++me) {
            const int* done = done_of ? done_of[comm->kind == COMM_LOCAL ? me : 0] : nullptr;
            k_push_partials<<<1, threads, 0, ctx().stream>>>(nq, me, comm->world, comm->gathered + (size_t)me * nq, comm->d_peer_red,
                                                             comm->d_peer_redflag, comm->red_seq, done);
            count_launch();
        }
        DB_CUDA(cudaGetLastError());
        comm->red_last_p2p = true;
        return 0;
    }
    if (comm->kind == COMM_NCCL && comm->world > 1) {
        ScopedTiming t("allgather");
        DB_NCCL(nccl().AllGather(comm->gathered + (size_t)comm->rank * nq, comm->gathered, (size_t)nq * 2, ncclDouble,
                                 (ncclComm_t)comm->nccl, ctx().stream));
    }
    return 0;
}

// Wide reduction: nq <= kWideQ double-double partials per rank (the i+1 products of a classical Gram-Schmidt column,
// krylovsolver.h:465-494) -> every rank holds all of them (ONE all-gather per column; in-process transport: the
// partitions wrote into the shared buffer themselves) -> k_wide_sum adds them in rank order, so all ranks round the
// same sums.
static inline dd* wide_slot(drops_b200_dist_matrix_s* M, int nq) { return M->comm->wide + (size_t)M->rank * nq; }

static int wide_allgather(drops_b200_comm_s* comm, int nq)
{
    DB_ARG(nq >= 1 && nq <= kWideQ, "wide reduction: too many quantities");
    if (comm->kind == COMM_NCCL && comm->world > 1) {
        ScopedTiming t("allgather");
        DB_NCCL(nccl().AllGather(comm->wide + (size_t)comm->rank * nq, comm->wide, (size_t)nq * 2, ncclDouble,
                                 (ncclComm_t)comm->nccl, ctx().stream));
    }
    return 0;
}

__global__ void k_wide_sum(const dd* __restrict__ wide, int world, int nq, double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nq) return;
    dd a = wide[k];
    for (int r = 1; r < world; ++r) a = dd_add(a, wide[(size_t)r * nq + k]);
    out[k] = dd_round(a);
}

// Peer mappings of one partition.  NCCL transport: IPC handles of every rank's ghost buffer and flag
// array travel through one all-gather (collective: every rank builds its partition); in-process
// transport: the peers are the other partitions of this process.  Any failure leaves p2p = -1 on
// ALL ranks and the exchange stays on ncclSend/ncclRecv.
static int p2p_finish(drops_b200_dist_matrix_s* M, const std::vector<double*>& ghost_of, const std::vector<unsigned*>& flags_of,
                      const std::vector<std::vector<size_t>>& recv_off_of)
{
    const int world = M->world, me = M->rank;
    cudaStream_t st = ctx().stream;
    std::vector<double*> dst((size_t)world, nullptr);
    std::vector<unsigned*> flg((size_t)world, nullptr);
    std::vector<unsigned char> expect((size_t)world, 0);
    for (int q = 0; q < world; ++q) {
        if (M->send_off[(size_t)q + 1] > M->send_off[(size_t)q]) {
            dst[(size_t)q] = ghost_of[(size_t)q] + recv_off_of[(size_t)q][(size_t)me];
            flg[(size_t)q] = flags_of[(size_t)q] + me;
        }
        expect[(size_t)q] = M->recv_off[(size_t)q + 1] > M->recv_off[(size_t)q] ? 1 : 0;
    }
    DB_CHECK(dev_alloc(&M->d_peer_dst, (size_t)world)); DB_CHECK(dev_alloc(&M->d_peer_flag, (size_t)world));
    DB_CHECK(dev_alloc(&M->d_send_off, (size_t)world + 1)); DB_CHECK(dev_alloc(&M->d_expect, (size_t)world));
    DB_CHECK(dev_alloc(&M->p2p_ticket, 1)); DB_CHECK(dev_alloc(&M->p2p_fault, 1));
    if (opt("dist.wait_stats", 1.0) != 0.0 && !M->wait_stats) {
        DB_CHECK(dev_alloc(&M->wait_stats, 8));
        DB_CUDA(cudaMemsetAsync(M->wait_stats, 0, 8 * sizeof(unsigned long long), st));
    }
    DB_CUDA(cudaMemcpyAsync(M->d_peer_dst, dst.data(), (size_t)world * sizeof(double*), cudaMemcpyHostToDevice, st));
    DB_CUDA(cudaMemcpyAsync(M->d_peer_flag, flg.data(), (size_t)world * sizeof(unsigned*), cudaMemcpyHostToDevice, st));
    DB_CUDA(cudaMemcpyAsync(M->d_send_off, M->send_off.data(), ((size_t)world + 1) * sizeof(size_t), cudaMemcpyHostToDevice, st));
    DB_CUDA(cudaMemcpyAsync(M->d_expect, expect.data(), (size_t)world, cudaMemcpyHostToDevice, st));
    DB_CUDA(cudaMemsetAsync(M->p2p_ticket, 0, sizeof(unsigned), st));
    DB_CUDA(cudaMemsetAsync(M->p2p_fault, 0, sizeof(int), st));
    DB_CUDA(cudaStreamSynchronize(st));
    M->p2p = 1;
    return 0;
}

static int p2p_setup_nccl(drops_b200_dist_matrix_s* M)
{
    drops_b200_comm_s* comm = M->comm;
    const int world = M->world, me = M->rank;
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    M->p2p = -1;
    if (opt("dist.p2p", 1.0) == 0.0 || world < 2 || !nccl().ok || comm->nccl == nullptr) return 0;
    // blob per rank: handle of the ghost + flags allocation | offset of the flags in it | status | recv_off[world + 1]
    const size_t blob = 2 * sizeof(cudaIpcMemHandle_t) + 8 + ((size_t)world + 1) * sizeof(size_t);
    std::vector<unsigned char> mine(blob, 0), all(blob * (size_t)world, 0);
    cudaIpcMemHandle_t hg, hf;
    long long ok = 1;
    std::memset(&hg, 0, sizeof hg); std::memset(&hf, 0, sizeof hf);
    if (cudaIpcGetMemHandle(&hg, M->ghost) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        std::memset(&hg, 0, sizeof hg);
    }
    const unsigned long long foff = M->halo_flag_offset;
    std::memcpy(&hf, &foff, sizeof foff);                  // second slot carries the flag offset
    std::memcpy(mine.data(), &hg, sizeof hg);
    std::memcpy(mine.data() + sizeof hg, &hf, sizeof hf);
    std::memcpy(mine.data() + 2 * sizeof hg, &ok, 8);
    std::memcpy(mine.data() + 2 * sizeof hg + 8, M->recv_off.data(), ((size_t)world + 1) * sizeof(size_t));
    unsigned char* d_all = nullptr;
    DB_CHECK(dev_alloc(&d_all, blob * (size_t)world));
    cudaError_t e = cudaMemcpyAsync(d_all + blob * (size_t)me, mine.data(), blob, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { dev_free(d_all); return cuda_fail(e, "p2p setup", __FILE__, __LINE__); }
    ncclResult_t nr = nccl().AllGather(d_all + blob * (size_t)me, d_all, blob, ncclChar, (ncclComm_t)comm->nccl, st);
    if (nr != ncclSuccess) { dev_free(d_all); return nccl_fail(nr, "p2p handle all-gather"); }
    e = cudaMemcpyAsync(all.data(), d_all, blob * (size_t)world, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(d_all);
    if (e != cudaSuccess) return cuda_fail(e, "p2p setup", __FILE__, __LINE__);
    std::vector<double*> ghost_of((size_t)world, nullptr);
    std::vector<unsigned*> flags_of((size_t)world, nullptr);
    std::vector<std::vector<size_t>> roff((size_t)world, std::vector<size_t>((size_t)world + 1, 0));
    bool good = true;
    for (int q = 0; q < world; ++q) {
        const unsigned char* b = all.data() + blob * (size_t)q;
        long long okq = 0;
        std::memcpy(&okq, b + 2 * sizeof hg, 8);
        std::memcpy(roff[(size_t)q].data(), b + 2 * sizeof hg + 8, ((size_t)world + 1) * sizeof(size_t));
        if (!okq) good = false;
    }
    for (int q = 0; q < world && good; ++q) {
        if (q == me) { ghost_of[(size_t)q] = M->ghost; flags_of[(size_t)q] = M->halo_flags; continue; }
        if (M->send_off[(size_t)q + 1] == M->send_off[(size_t)q]) continue;      // nothing goes to q
        const unsigned char* b = all.data() + blob * (size_t)q;
        std::memcpy(&hg, b, sizeof hg); std::memcpy(&hf, b + sizeof hg, sizeof hf);
        void* pg = nullptr;
        if (cudaIpcOpenMemHandle(&pg, hg, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; break; }
        M->ipc_mapped.push_back(pg);
        unsigned long long foff_q = 0;
        std::memcpy(&foff_q, &hf, sizeof foff_q);
        ghost_of[(size_t)q] = (double*)pg; flags_of[(size_t)q] = (unsigned*)((char*)pg + foff_q);
    }
    double bad = good ? 0.0 : 1.0;                         // all ranks or none
    DB_CHECK(drops_b200_comm_allreduce_max(comm, &bad));
    if (bad != 0.0) {
        for (void* m : M->ipc_mapped) cudaIpcCloseMemHandle(m);
        M->ipc_mapped.clear();
        return 0;
    }
    return p2p_finish(M, ghost_of, flags_of, roff);
}

static int p2p_setup_local(int np, drops_b200_dist_matrix_s* const* P)
{
    std::vector<double*> ghost_of((size_t)np);
    std::vector<unsigned*> flags_of((size_t)np);
    std::vector<std::vector<size_t>> roff((size_t)np);
    for (int p = 0; p < np; ++p) { ghost_of[(size_t)p] = P[p]->ghost; flags_of[(size_t)p] = P[p]->halo_flags; roff[(size_t)p] = P[p]->recv_off; }
    for (int p = 0; p < np; ++p) {
        if (P[p]->p2p != 0) continue;
        P[p]->p2p = -1;
        if (opt("dist.p2p", 1.0) == 0.0 || np < 2) continue;
        DB_CHECK(p2p_finish(P[p], ghost_of, flags_of, roff));
    }
    return 0;
}

static int halo_exchange(int np, drops_b200_dist_matrix_s* const* P, const PartArgs* a)
{
    drops_b200_comm_s* comm = P[0]->comm;
    if (comm->world == 1) return 0;
    cudaStream_t st = ctx().stream, cs = comm->comm_stream;
    if (comm->kind == COMM_LOCAL && P[0]->p2p == 0) DB_CHECK(p2p_setup_local(np, P));
    // peer path: only inside solver loops (a[p].p2p) -- there two reductions over all ranks separate one
    // exchange from the next, so a ghost buffer is never overwritten while its owner still reads it
    bool peer = a[0].p2p && opt("dist.p2p", 1.0) != 0.0;
    for (int p = 0; p < np; ++p) peer = peer && P[p]->p2p == 1;
    if (peer) {
        // on the communication stream, beside the interior rows (as the NCCL exchange below)
        DB_CUDA(cudaEventRecord(comm->ev_pack, st));
        DB_CUDA(cudaStreamWaitEvent(cs, comm->ev_pack, 0));
        for (int p = 0; p < np; ++p) {
            drops_b200_dist_matrix_s* M = P[p];
            ++M->p2p_seq;
            if (M->n_send) {
                k_pack_push<<<blocks_for((long long)M->n_send, 256), 256, 0, cs>>>(
                    (long long)M->n_send, a[p].ro ? M->d_send_idx_ro : M->d_send_idx, a[p].x, M->d_send_off, M->world,
                    M->d_peer_dst, M->d_peer_flag, M->p2p_seq, M->p2p_ticket, a[p].done);
                count_launch();
            }
        }
        DB_CUDA(cudaGetLastError());
        DB_CUDA(cudaEventRecord(comm->ev_halo, cs));
        return 1;                                  // the caller waits with k_halo_wait before the boundary rows
    }
    {
        ScopedTiming t("halo.pack");
        for (int p = 0; p < np; ++p)
            if (P[p]->n_send) {
                k_pack<<<blocks_for((long long)P[p]->n_send, 256), 256, 0, st>>>((long long)P[p]->n_send,
                                                                                a[p].ro ? P[p]->d_send_idx_ro : P[p]->d_send_idx,
                                                                                a[p].x, P[p]->sendbuf, a[p].done);
                count_launch();
            }
    }
    DB_CUDA(cudaGetLastError());
    DB_CUDA(cudaEventRecord(comm->ev_pack, st));
    DB_CUDA(cudaStreamWaitEvent(cs, comm->ev_pack, 0));
    if (comm->kind == COMM_NCCL) {
        drops_b200_dist_matrix_s* M = P[0];
        DB_NCCL(nccl().GroupStart());
        for (int q = 0; q < comm->world; ++q) {
            const size_t ns = M->send_off[(size_t)q + 1] - M->send_off[(size_t)q];
            const size_t nr = M->recv_off[(size_t)q + 1] - M->recv_off[(size_t)q];
            if (ns) DB_NCCL(nccl().Send(M->sendbuf + M->send_off[(size_t)q], ns, ncclDouble, q, (ncclComm_t)comm->nccl, cs));
            if (nr) DB_NCCL(nccl().Recv(M->ghost + M->recv_off[(size_t)q], nr, ncclDouble, q, (ncclComm_t)comm->nccl, cs));
        }
        DB_NCCL(nccl().GroupEnd());
    } else {
        for (int r = 0; r < np; ++r)
            for (int q = 0; q < np; ++q) {
                const size_t nr = P[r]->recv_off[(size_t)q + 1] - P[r]->recv_off[(size_t)q];
                if (!nr) continue;
                const size_t ns = P[q]->send_off[(size_t)r + 1] - P[q]->send_off[(size_t)r];
                DB_ARG(ns == nr, "send and receive lists of two partitions do not match");
                DB_CUDA(cudaMemcpyAsync(P[r]->ghost + P[r]->recv_off[(size_t)q], P[q]->sendbuf + P[q]->send_off[(size_t)r],
                                        nr * 8, cudaMemcpyDeviceToDevice, cs));
            }
    }
    DB_CUDA(cudaEventRecord(comm->ev_halo, cs));
    return 0;
}

// y = epilogue(A x) on every partition; with a[p].dotv the local partial of y.dotv goes to the
// partition's gather slot (1 dd per rank) -- the caller all-gathers.
static int dist_spmv_run(int np, drops_b200_dist_matrix_s* const* P, const PartArgs* a)
{
    drops_b200_comm_s* comm = P[0]->comm;
    cudaStream_t st = ctx().stream;
    const int hx = halo_exchange(np, P, a);
    if (hx < 0) return hx;
    const bool peer = hx == 1;
    for (int p = 0; p < np; ++p) {
        drops_b200_dist_matrix_s* M = P[p];
        if (a[p].dotv && M->n_int == 0 && M->n_bnd == 0) {
            k_zero_dd<<<1, 32, 0, st>>>(gather_slot(M, 1), 1);
            count_launch();
        }
        if (M->n_int == 0) continue;
        ScopedTiming t("spmv");
        SpmvArgs s;
        s.x = a[p].x; s.y = a[p].y; s.epilogue = a[p].epilogue; s.aux = a[p].aux; s.omega = a[p].omega; s.done = a[p].done;
        s.beta_val = a[p].beta_val;
        if (a[p].dotv) { s.dotv = a[p].dotv; s.dd_out = M->n_bnd ? (dd*)M->d_scalars + 32 : gather_slot(M, 1); }
        s.row_ordered = a[p].ro;
        DB_CHECK(spmv_sell(M->sell_int, s));
    }
    // "dist.overlap_bnd" (peer path): the boundary rows run on the communication stream, behind the pack-and-push and the
    // halo wait, BESIDE the interior rows (high-priority stream: their CTAs take the next free slots) instead of after them --
    // ranks with two faces no longer finish later than ranks with one.  The two dot partials are combined in the order of
    // the chained form, so the sum keeps its bits.
    const bool overlap = peer && opt("dist.overlap_bnd", 1.0) != 0.0;
    if (overlap) {
        cudaStream_t cs = comm->comm_stream;
        {
            ScopedTiming t("halo.wait", cs);
            for (int p = 0; p < np; ++p) {
                drops_b200_dist_matrix_s* M = P[p];
                if (M->n_ghost == 0) continue;
                k_halo_wait<<<1, std::max(32, (M->world + 31) / 32 * 32), 0, cs>>>(M->world, M->halo_flags, M->d_expect, M->p2p_seq,
                                                                                 a[p].done, M->p2p_fault, M->wait_stats);
                count_launch();
            }
            DB_CUDA(cudaGetLastError());
        }
        for (int p = 0; p < np; ++p) {
            drops_b200_dist_matrix_s* M = P[p];
            if (M->n_bnd == 0) continue;
            ScopedTiming t("spmv.bnd", cs);
            SpmvArgs s;
            s.x = a[p].x; s.y = a[p].y; s.epilogue = a[p].epilogue; s.aux = a[p].aux; s.omega = a[p].omega; s.done = a[p].done;
            s.beta_val = a[p].beta_val;
            s.xghost = M->ghost; s.n_owned = a[p].ro ? (long long)M->npos : (long long)M->n_owned;
            s.stream = cs; s.slot = 2;                                   // (the interior rows reduce in slot 0 at the same time)
            if (a[p].dotv) { s.dotv = a[p].dotv; s.dd_out = (dd*)M->d_scalars + 33; }
            DB_CHECK(spmv_sell(a[p].ro ? M->sell_bnd_ro : M->sell_bnd, s));
        }
        DB_CUDA(cudaEventRecord(comm->ev_bnd, cs));
        DB_CUDA(cudaStreamWaitEvent(st, comm->ev_bnd, 0));               // also: my pushes have left
        for (int p = 0; p < np; ++p) {
            drops_b200_dist_matrix_s* M = P[p];
            if (!a[p].dotv || M->n_bnd == 0) continue;
            k_combine_dd<<<1, 1, 0, st>>>(gather_slot(M, 1), (dd*)M->d_scalars + 33, M->n_int ? (dd*)M->d_scalars + 32 : nullptr, a[p].done);
            count_launch();
        }
        DB_CUDA(cudaGetLastError());
        return 0;
    }
    if (comm->world > 1) DB_CUDA(cudaStreamWaitEvent(st, comm->ev_halo, 0));     // my sends / pushes have left
    if (peer) {
        ScopedTiming t("halo.wait");
        for (int p = 0; p < np; ++p) {
            drops_b200_dist_matrix_s* M = P[p];
            if (M->n_ghost == 0) continue;
            k_halo_wait<<<1, std::max(32, (M->world + 31) / 32 * 32), 0, st>>>(M->world, M->halo_flags, M->d_expect, M->p2p_seq,
                                                                             a[p].done, M->p2p_fault, M->wait_stats);
            count_launch();
        }
        DB_CUDA(cudaGetLastError());
    }
    for (int p = 0; p < np; ++p) {
        drops_b200_dist_matrix_s* M = P[p];
        if (M->n_bnd == 0) continue;
        ScopedTiming t("spmv.bnd");
        SpmvArgs s;
        s.x = a[p].x; s.y = a[p].y; s.epilogue = a[p].epilogue; s.aux = a[p].aux; s.omega = a[p].omega; s.done = a[p].done;
        s.beta_val = a[p].beta_val;
        s.xghost = M->ghost; s.n_owned = a[p].ro ? (long long)M->npos : (long long)M->n_owned;
        if (a[p].dotv) { s.dotv = a[p].dotv; s.dd_in = M->n_int ? (dd*)M->d_scalars + 32 : nullptr; s.dd_out = gather_slot(M, 1); }
        DB_CHECK(spmv_sell(a[p].ro ? M->sell_bnd_ro : M->sell_bnd, s));
    }
    return 0;
}

static int ensure_part_scalars(drops_b200_dist_matrix_s* M, int hist_cap)
{
    // SolverScalars followed by scratch: dd partial at +32 dd (chained interior/boundary dot)
    if (!M->d_scalars) {
        char* raw = nullptr;
        DB_CHECK(dev_alloc(&raw, 4096));
        DB_CUDA(cudaMemsetAsync(raw, 0, 4096, ctx().stream));
        M->d_scalars = raw;
    }
    if (hist_cap > M->hist_alloc) {
        dev_free(M->d_hist); M->d_hist = nullptr; M->hist_alloc = 0;
        DB_CHECK(dev_alloc(&M->d_hist, (size_t)hist_cap));
        M->hist_alloc = hist_cap;
    }
    return 0;
}
static_assert(sizeof(SolverScalars) <= 32 * sizeof(dd), "scratch offset inside the scalar block");

static int scalars_kernel(int phase, drops_b200_dist_matrix_s* M, double* hist, double* out)
{
    drops_b200_comm_s* C = M->comm;
    if (C->red_last_p2p)
        k_dist_scalars<<<1, std::max(32, (M->world + 31) / 32 * 32), 0, ctx().stream>>>(
            phase, (SolverScalars*)M->d_scalars, hist, C->red_buf + (size_t)(C->red_seq & 1u) * M->world * kMaxQ, M->world, out,
            C->red_flags, C->red_seq, C->red_fault, M->wait_stats);
    else
        k_dist_scalars<<<1, 32, 0, ctx().stream>>>(phase, (SolverScalars*)M->d_scalars, hist, C->gathered, M->world, out, nullptr, 0u, nullptr, nullptr);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

// The bounded peer-to-peer waits (k_halo_wait, k_dist_scalars) never hang the GPU: after 2^26 spins they give up and
// raise a sticky device flag.  Every solver entry point reads the flags before it returns; a set flag means ghost
// values or reduction partials of some step were stale (stalled host, crashed rank), so the solve is reported as
// failed -- on this rank; the peers run into their own time-outs.
static int check_p2p_faults(int np, drops_b200_dist_matrix_s* const* P)
{
    cudaStream_t st = ctx().stream;
    int* h = reinterpret_cast<int*>(ctx().h_scalars + 24);       // pinned
    int n = 0;
    for (int p = 0; p < np; ++p)
        if (P[p]->p2p_fault) DB_CUDA(cudaMemcpyAsync(h + n++, P[p]->p2p_fault, sizeof(int), cudaMemcpyDeviceToHost, st));
    const int n_halo = n;
    if (P[0]->comm->red_fault) DB_CUDA(cudaMemcpyAsync(h + n++, P[0]->comm->red_fault, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n == 0) return 0;
    DB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i)
        if (h[i] != 0) {
            set_error("peer-to-peer %s timed out: a rank of the partition stalled or died; the result of this solve is invalid",
                      i < n_halo ? "halo wait" : "reduction wait");
            return DROPS_B200_E_COMM;
        }
    return 0;
}

int dist_pcg(int np, drops_b200_dist_matrix_s* const* P, double* const* x, const double* const* b, int pc, double omega,
             int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(check_parts(np, P));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    DB_ARG(pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI, "unknown preconditioner");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    drops_b200_comm_s* comm = P[0]->comm;
    const int want_hist = hist ? std::max(hist_cap, 0) : 0;
    std::vector<double*> r(np), pv(np), q(np);
    std::vector<SolverScalars*> S(np);
    // row-ordered solve: every partition iterates on vectors in the numbering of its own slices; x and b are
    // gathered once and x is scattered back at the end (same products in every row, dist.cuh)
    bool ro = opt("dist.reorder", 1.0) != 0.0 && opt("spmv.pk", 1.0) != 0.0 && opt("spmv.kernel", 0.0) == 0.0;
    for (int p = 0; p < np; ++p) ro = ro && P[p]->npos > 0;
    {
        double flag = ro ? 0.0 : 1.0;                         // all ranks or none
        DB_CHECK(drops_b200_comm_allreduce_max(comm, &flag));
        ro = flag == 0.0;
    }
    std::vector<double*> x_user(x, x + np);
    std::vector<double*> xv(x, x + np);
    std::vector<const double*> bv(b, b + np), dg(np);
    std::vector<size_t> nv(np);
    std::vector<DiagIndex> dix(np);               // (idx == nullptr: the kernels read the fp64 diagonal)
    for (int p = 0; p < np; ++p) {
        dg[p] = P[p]->diag; nv[p] = P[p]->n_owned;
        if (ro && pc == DROPS_B200_PC_JACOBI && P[p]->diag_ro_idx != nullptr && P[p]->diag_ro_count > 0 &&
            opt("pcg.diag_index", 1.0) != 0.0) {
            dix[p].idx = P[p]->diag_ro_idx;
            std::memcpy(dix[p].dict.v, P[p]->diag_ro_values, sizeof(double) * (size_t)P[p]->diag_ro_count);
        }
        if (ro) {
            double* xp = dwork(P[p], 3); double* bp = dwork(P[p], 4);
            if (!xp || !bp) return DROPS_B200_E_NOMEM;
            DB_CHECK(vec_gather_perm(P[p]->npos, P[p]->ro_perm, x[p], xp, 0.0));
            DB_CHECK(vec_gather_perm(P[p]->npos, P[p]->ro_perm, b[p], bp, 0.0));
            xv[p] = xp; bv[p] = bp; dg[p] = P[p]->diag_ro; nv[p] = P[p]->npos;
        }
    }
    SolverScalars h;
    std::memset(&h, 0, sizeof h);
    h.tol = *tol; h.relative = relative; h.max_iter = *max_iter; h.pc = pc; h.omega = omega; h.hist_cap = want_hist;
    // Jacobi needs a stored diagonal on EVERY rank: agree through the gather buffer of a dot-like phase
    int diag_bad = 0;
    for (int p = 0; p < np; ++p) {
        if (pc == DROPS_B200_PC_JACOBI && !P[p]->diag_ok) diag_bad = 1;
        DB_CHECK(ensure_part_scalars(P[p], std::max(want_hist, 1)));
        S[p] = (SolverScalars*)P[p]->d_scalars;
        r[p] = dwork(P[p], 0); pv[p] = dwork(P[p], 1); q[p] = dwork(P[p], 2);
        if (!r[p] || !pv[p] || !q[p]) return DROPS_B200_E_NOMEM;
        if (ro) {
            // the boundary operator skips its padding rows (the tail of the position range): r and q must hold
            // zeros there, the vector kernels run over all npos positions
            const size_t tail0 = P[p]->npos_int + P[p]->n_bnd, tail = P[p]->npos - tail0;
            DB_CUDA(cudaMemsetAsync(r[p] + tail0, 0, tail * sizeof(double), st));
            DB_CUDA(cudaMemsetAsync(q[p] + tail0, 0, tail * sizeof(double), st));
        }
        DB_CUDA(cudaMemcpyAsync(S[p], &h, sizeof h, cudaMemcpyHostToDevice, st));
    }
    {
        double flag = diag_bad;
        DB_CHECK(drops_b200_comm_allreduce_max(comm, &flag));
        if (flag != 0.0) {
            set_error("JACPcCL: at least one row (on some rank) stores no diagonal entry");
            return DROPS_B200_E_NODIAG;
        }
    }
    double* dh = want_hist ? P[0]->d_hist : nullptr;
    std::vector<PartArgs> a(np);
    std::vector<char> defer_x(np);
    for (int p = 0; p < np; ++p)
        defer_x[p] = opt("pcg.defer_x", 1.0) != 0.0 && opt("vec.wide", 1.0) != 0.0 && nv[p] > 0 &&
                     ((reinterpret_cast<uintptr_t>(xv[p]) | reinterpret_cast<uintptr_t>(pv[p]) | reinterpret_cast<uintptr_t>(r[p]) |
                       reinterpret_cast<uintptr_t>(dg[p])) & 15u) == 0;

    // r = b - A x; setup partials; scalars
    for (int p = 0; p < np; ++p) { a[p] = PartArgs(); a[p].x = xv[p]; a[p].y = r[p]; a[p].epilogue = EPI_RESID; a[p].aux = bv[p]; a[p].ro = ro; }
    DB_CHECK(dist_spmv_run(np, P, a.data()));
    for (int p = 0; p < np; ++p)
        DB_CHECK(pcg_setup(nv[p], r[p], bv[p], dg[p], pv[p], S[p], nullptr, gather_slot(P[p], 3)));
    DB_CHECK(allgather_partials(comm, 3));
    for (int p = 0; p < np; ++p) DB_CHECK(scalars_kernel(PH_SETUP, P[p], p == 0 ? dh : nullptr, nullptr));

    int check_every = (int)opt("pcg.check_every", 16.0);
    if (check_every < 1) check_every = 1;
    struct Flags { int iter, done, converged, pad; };
    Flags* hflags = reinterpret_cast<Flags*>(c.h_scalars + 16);
    cudaEvent_t ev[2];
    DB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    DB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int status = 0, launched = 0, chunk_id = 0;
    bool finished = false;
    const int total = std::max(*max_iter, 0);
    auto enqueue_flag_read = [&](int slot) -> int {
        DB_CUDA(cudaMemcpyAsync(&hflags[slot], &S[0]->iter, sizeof(Flags), cudaMemcpyDeviceToHost, st));
        DB_CUDA(cudaEventRecord(ev[slot], st));
        return 0;
    };
    std::vector<const int*> done_ptrs(np);       // skipped iterations skip their reductions on every rank alike
    for (int p = 0; p < np; ++p) done_ptrs[p] = &S[p]->done;
    auto iteration = [&]() -> int {
        for (int p = 0; p < np; ++p) {
            a[p] = PartArgs(); a[p].x = pv[p]; a[p].y = q[p]; a[p].dotv = pv[p]; a[p].done = &S[p]->done; a[p].ro = ro; a[p].p2p = true;
        }
        DB_CHECK(dist_spmv_run(np, P, a.data()));                                     // q = A p, partial p.q
        DB_CHECK(allgather_partials(comm, 1, done_ptrs.data()));
        for (int p = 0; p < np; ++p) DB_CHECK(scalars_kernel(PH_LAMBDA, P[p], nullptr, nullptr));
        // x += alpha p rides in the direction kernel (10 instead of 11 vector passes, vecops.cuh)
        for (int p = 0; p < np; ++p) {
            if (defer_x[p]) DB_CHECK(pcg_update_r(nv[p], r[p], q[p], dg[p], S[p], nullptr, gather_slot(P[p], 2), &dix[p]));
            else DB_CHECK(pcg_update(nv[p], xv[p], r[p], pv[p], q[p], dg[p], S[p], nullptr, gather_slot(P[p], 2)));
        }
        DB_CHECK(allgather_partials(comm, 2, done_ptrs.data()));
        for (int p = 0; p < np; ++p) DB_CHECK(scalars_kernel(PH_UPDATE, P[p], p == 0 ? dh : nullptr, nullptr));
        for (int p = 0; p < np; ++p) {
            if (defer_x[p]) DB_CHECK(pcg_direction_x(nv[p], xv[p], pv[p], r[p], dg[p], S[p]));
            else DB_CHECK(pcg_direction(nv[p], pv[p], r[p], dg[p], S[p]));
        }
        return 0;
    };
    status = enqueue_flag_read(0);
    chunk_id = 1;
    while (status == 0 && !finished) {
        const int chunk = std::min(check_every, total - launched);
        for (int k = 0; k < chunk && status == 0; ++k) status = iteration();
        launched += chunk;
        if (status != 0) break;
        const int slot = chunk_id & 1;
        status = enqueue_flag_read(slot);
        if (status != 0) break;
        const int prev = slot ^ 1;
        cudaError_t e = cudaEventSynchronize(ev[prev]);
        if (e != cudaSuccess) { status = cuda_fail(e, "event sync", __FILE__, __LINE__); break; }
        if (hflags[prev].done || launched >= total) {
            e = cudaEventSynchronize(ev[slot]);
            if (e != cudaSuccess) { status = cuda_fail(e, "event sync", __FILE__, __LINE__); break; }
            if (hflags[slot].done || launched >= total) finished = true;
        }
        ++chunk_id;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (status != 0) { cudaStreamSynchronize(st); return status; }
    if (ro) for (int p = 0; p < np; ++p) DB_CHECK(vec_scatter_perm(P[p]->npos, P[p]->ro_perm, xv[p], x_user[p]));
    DB_CUDA(cudaMemcpyAsync(&h, S[0], sizeof h, cudaMemcpyDeviceToHost, st));
    DB_CUDA(cudaStreamSynchronize(st));
    if (comm->comm_stream) DB_CUDA(cudaStreamSynchronize(comm->comm_stream));
    DB_CHECK(check_p2p_faults(np, P));
    if (hist && want_hist > 0) {
        const int cnt = std::min(want_hist, h.iter + 1);
        DB_CUDA(cudaMemcpyAsync(hist, P[0]->d_hist, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
        DB_CUDA(cudaStreamSynchronize(st));
    }
    *tol = h.resid;
    if (h.converged) *max_iter = h.iter;
    return h.converged ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// GMRES(m) on the partitioned system -- the serial driver of num/krylovsolver.h:320-461 (modified
// Gram-Schmidt, Givens rotations on the host) with every reduction summed over the ranks in rank
// order: partial -> all-gather -> k_dist_scalars, result left on the device for the fused axpy of the
// next Gram-Schmidt step (no host round trip inside a column; one fetch of the column at its end).
// All ranks hold bit-identical Hessenberg columns, so they rotate, stop and restart in step.
// The reference's own distributed variants (krylovsolver.h:528-916) need one GlobalSum per
// Gram-Schmidt step as well (or one per column with classical Gram-Schmidt).
// ------------------------------------------------------------------------------------------
static void gm_gen_rot(double& dx, double& dy, double& cs, double& sn)     // :258-271
{
    if (dy == 0.0) { cs = 1.0; sn = 0.0; }
    else { const double r = std::sqrt(dx * dx + dy * dy); cs = dx / r; sn = dy / r; }
}
static void gm_app_rot(double& dx, double& dy, double cs, double sn)       // :274-279
{
    const double tmp = cs * dx + sn * dy;
    dy = -sn * dx + cs * dy;
    dx = tmp;
}

int dist_gmres(int np, drops_b200_dist_matrix_s* const* P, double* const* x, const double* const* b, int pc, double omega,
               int restart, int* max_iter, double* tol, int relative, int left)
{
    DB_CHECK(check_parts(np, P));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    DB_ARG(pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI, "unknown preconditioner");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    drops_b200_comm_s* comm = P[0]->comm;
    const int m = restart <= *max_iter ? restart : *max_iter;                  // :327
    DB_ARG(m >= 2, "GMRES needs restart >= 2 and max_iter >= 2");
    const bool jac = pc == DROPS_B200_PC_JACOBI;
    const bool cgs = opt("gmres.ortho", 0.0) == 1.0 && m <= kWideQ;          // "gmres.ortho" as in the single-GPU driver
    {
        double flag = 0.0;
        for (int p = 0; p < np; ++p) if (jac && !P[p]->diag_ok) flag = 1.0;
        DB_CHECK(drops_b200_comm_allreduce_max(comm, &flag));
        if (flag != 0.0) { set_error("JACPcCL: at least one row (on some rank) stores no diagonal entry"); return DROPS_B200_E_NODIAG; }
    }
    // per partition: r, w, tmp, z, basis v[0..m-1]; small device block h[m+1] | y[m] | vptrs[m] | scal[4]
    struct Part { double *r, *w, *tmp, *z; std::vector<double*> v; double *d_h, *d_y, *d_scal; const double** d_vptrs; size_t n; };
    std::vector<Part> W((size_t)np);
    for (int p = 0; p < np; ++p) {
        Part& w = W[(size_t)p];
        w.n = P[p]->n_owned;
        DB_CHECK(ensure_part_scalars(P[p], 1));
        w.r = dwork(P[p], 0); w.w = dwork(P[p], 1); w.tmp = dwork(P[p], 2); w.z = dwork(P[p], 3);
        if (!w.r || !w.w || !w.tmp || !w.z) return DROPS_B200_E_NOMEM;
        w.v.resize((size_t)m);
        for (int i = 0; i < m; ++i) { w.v[(size_t)i] = dwork(P[p], 4 + i); if (!w.v[(size_t)i]) return DROPS_B200_E_NOMEM; }
        const size_t need = (size_t)3 * m + 8;
        if (need > P[p]->small_alloc) {
            dev_free(P[p]->d_small); P[p]->d_small = nullptr; P[p]->small_alloc = 0;
            DB_CHECK(dev_alloc(&P[p]->d_small, need));
            P[p]->small_alloc = need;
        }
        w.d_h = P[p]->d_small; w.d_y = w.d_h + (m + 1);
        w.d_vptrs = reinterpret_cast<const double**>(w.d_y + m);
        w.d_scal = reinterpret_cast<double*>(w.d_vptrs + m);
        DB_CUDA(cudaMemcpyAsync(w.d_vptrs, w.v.data(), (size_t)m * sizeof(double*), cudaMemcpyHostToDevice, st));
    }
    std::vector<double> H((size_t)m * m, 0.0), s(m, 0.0), cs(m, 0.0), sn(m, 0.0), y(m, 0.0), hcol(m + 1, 0.0);
    auto Hat = [&](int i, int j) -> double& { return H[(size_t)j * m + i]; };
    std::vector<PartArgs> a((size_t)np);

    auto apply_pc = [&](int p, double* out, const double* in) -> int {
        if (jac) return vec_jacobi(W[(size_t)p].n, omega, P[p]->diag, in, out);
        return vec_copy(W[(size_t)p].n, out, in);
    };
    // global reduction: local partials are in the gather slots; result -> out_of(p) on every partition
    auto reduce_to = [&](int phase, double* (*slot)(Part&, int), int idx) -> int {
        DB_CHECK(allgather_partials(comm, 1));
        for (int p = 0; p < np; ++p) DB_CHECK(scalars_kernel(phase, P[p], nullptr, slot(W[(size_t)p], idx)));
        return 0;
    };
    auto h_slot = [](Part& w, int k) -> double* { return w.d_h + k; };
    auto scal_slot = [](Part& w, int k) -> double* { return w.d_scal + k; };
    auto norm_of = [&](double* Part::*vec, int scal_idx) -> int {            // scal[idx] = |vec|_2 over all ranks
        for (int p = 0; p < np; ++p)
            DB_CHECK(mgs_step(W[(size_t)p].n, W[(size_t)p].*vec, nullptr, nullptr, nullptr, nullptr, true, gather_slot(P[p], 1)));
        return reduce_to(PH_NORM, scal_slot, scal_idx);
    };
    auto residual = [&]() -> int {                                            // r = [M^-1](b - A x)
        for (int p = 0; p < np; ++p) {
            a[(size_t)p] = PartArgs(); a[(size_t)p].x = x[p]; a[(size_t)p].epilogue = EPI_RESID; a[(size_t)p].aux = b[p];
            a[(size_t)p].y = (left && jac) ? W[(size_t)p].tmp : W[(size_t)p].r;
        }
        DB_CHECK(dist_spmv_run(np, P, a.data()));
        if (left && jac) for (int p = 0; p < np; ++p) DB_CHECK(apply_pc(p, W[(size_t)p].r, W[(size_t)p].tmp));
        return 0;
    };
    auto fetch_host = [&](const double* d, double* h, size_t cnt) -> int { return fetch_small(d, h, cnt); };
    auto update_x = [&](int k) -> int {                                       // GMRES_Update :282-297
        if (k < 0) return 0;
        y = s;
        for (int i = k; i >= 0; --i) {
            y[(size_t)i] /= Hat(i, i);
            for (int j = i - 1; j >= 0; --j) y[(size_t)j] -= Hat(j, i) * y[(size_t)i];
        }
        for (int p = 0; p < np; ++p) {
            Part& w = W[(size_t)p];
            DB_CUDA(cudaMemcpyAsync(w.d_y, y.data(), (size_t)(k + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
            if (left) DB_CHECK(multi_axpy(w.n, x[p], w.d_vptrs, w.d_y, k + 1));
            else {                                                            // :416-421
                DB_CHECK(vec_fill(w.n, w.z, 0.0));
                DB_CHECK(multi_axpy(w.n, w.z, w.d_vptrs, w.d_y, k + 1));
                DB_CHECK(apply_pc(p, w.tmp, w.z));
                DB_CHECK(vec_axpy(w.n, 1.0, w.tmp, x[p]));
            }
        }
        DB_CUDA(cudaStreamSynchronize(st));
        return 0;
    };

    double beta, normb, resid;
    DB_CHECK(residual());
    DB_CHECK(norm_of(&Part::r, 0));                                            // beta = |r|
    for (int p = 0; p < np; ++p) {
        if (left) DB_CHECK(apply_pc(p, W[(size_t)p].w, b[p]));                 // normb = |M^-1 b|  :346-347
        else DB_CHECK(vec_copy(W[(size_t)p].n, W[(size_t)p].w, b[p]));         // normb = |b|       :340
    }
    DB_CHECK(norm_of(&Part::w, 1));
    DB_CHECK(fetch_host(W[0].d_scal, c.h_scalars, 2));
    beta = c.h_scalars[0]; normb = c.h_scalars[1];
    if (normb == 0.0 || !relative) normb = 1.0;
    resid = beta / normb;
    if (resid <= *tol) { *tol = resid; *max_iter = 0; return 1; }

    int j = 1;
    while (j <= *max_iter) {
        for (int p = 0; p < np; ++p) DB_CHECK(vec_scale_inv(W[(size_t)p].n, W[(size_t)p].v[0], W[(size_t)p].r, W[(size_t)p].d_scal + 0));
        std::fill(s.begin(), s.end(), 0.0);
        s[0] = beta;
        int i;
        for (i = 0; i < m - 1 && j <= *max_iter; ++i, ++j) {
            for (int p = 0; p < np; ++p) {
                Part& w = W[(size_t)p];
                a[(size_t)p] = PartArgs();
                a[(size_t)p].p2p = true;                                      // reductions over all ranks separate the exchanges
                if (left) {                                                   // w = M^-1 (A v[i])  :377
                    a[(size_t)p].x = w.v[(size_t)i]; a[(size_t)p].y = w.w;
                    if (jac) { a[(size_t)p].epilogue = EPI_JACOBI; a[(size_t)p].aux = P[p]->diag; a[(size_t)p].omega = omega; }
                } else {                                                      // w = A (M^-1 v[i])  :372-376
                    DB_CHECK(apply_pc(p, w.tmp, w.v[(size_t)i]));
                    a[(size_t)p].x = w.tmp; a[(size_t)p].y = w.w;
                }
            }
            DB_CHECK(dist_spmv_run(np, P, a.data()));
            if (cgs) {
                // classical Gram-Schmidt, the default of the reference's distributed ModGMRES (:465-494, :862-866): all
                // i+1 products against the unmodified w in ONE pass and ONE reduction over the ranks, then one update
                for (int p = 0; p < np; ++p) {
                    Part& w = W[(size_t)p];
                    DB_CHECK(cgs_dots(w.n, w.w, w.d_vptrs, i + 1, nullptr, wide_slot(P[p], i + 1)));
                }
                DB_CHECK(wide_allgather(comm, i + 1));
                for (int p = 0; p < np; ++p) {
                    Part& w = W[(size_t)p];
                    k_wide_sum<<<(i + 1 + 127) / 128, 128, 0, st>>>(comm->wide, comm->world, i + 1, w.d_h);
                    count_launch();
                    DB_CHECK(cgs_update(w.n, w.w, w.d_vptrs, w.d_h, i + 1, nullptr, gather_slot(P[p], 1)));              // + partial of |w|^2
                }
            } else {
                for (int k = 0; k <= i; ++k) {                                // modified Gram-Schmidt  :378-381
                    for (int p = 0; p < np; ++p) {
                        Part& w = W[(size_t)p];
                        DB_CHECK(mgs_step(w.n, w.w, k > 0 ? w.v[(size_t)k - 1] : nullptr, w.d_h + (k - 1), w.v[(size_t)k], nullptr, false,
                                          gather_slot(P[p], 1)));
                    }
                    DB_CHECK(reduce_to(PH_DOT, h_slot, k));
                }
                for (int p = 0; p < np; ++p) {
                    Part& w = W[(size_t)p];
                    DB_CHECK(mgs_step(w.n, w.w, w.v[(size_t)i], w.d_h + i, nullptr, nullptr, true, gather_slot(P[p], 1)));   // H(i+1,i) = |w|  :383
                }
            }
            DB_CHECK(reduce_to(PH_NORM, h_slot, i + 1));
            for (int p = 0; p < np; ++p) DB_CHECK(vec_scale_inv(W[(size_t)p].n, W[(size_t)p].v[(size_t)i + 1], W[(size_t)p].w, W[(size_t)p].d_h + i + 1));
            DB_CHECK(fetch_host(W[0].d_h, hcol.data(), (size_t)(i + 2)));
            for (int k = 0; k <= i + 1; ++k) Hat(k, i) = hcol[(size_t)k];
            for (int k = 0; k < i; ++k) gm_app_rot(Hat(k, i), Hat(k + 1, i), cs[(size_t)k], sn[(size_t)k]);
            gm_gen_rot(Hat(i, i), Hat(i + 1, i), cs[(size_t)i], sn[(size_t)i]);
            gm_app_rot(Hat(i, i), Hat(i + 1, i), cs[(size_t)i], sn[(size_t)i]);
            gm_app_rot(s[(size_t)i], s[(size_t)i + 1], cs[(size_t)i], sn[(size_t)i]);
            resid = std::abs(s[(size_t)i + 1]) / normb;                       // :393
            DB_CHECK(check_p2p_faults(np, P));                                // a stale column must not steer the rotations
            if (resid <= *tol) {                                              // :413-429
                DB_CHECK(update_x(i));
                *tol = resid; *max_iter = j;
                return 1;
            }
        }
        DB_CHECK(update_x(i - 1));                                            // :431-444
        DB_CHECK(residual());
        DB_CHECK(norm_of(&Part::r, 0));
        DB_CHECK(fetch_host(W[0].d_scal, c.h_scalars, 1));
        DB_CHECK(check_p2p_faults(np, P));
        beta = c.h_scalars[0];
        resid = beta / normb;
        if (resid <= *tol) { *tol = resid; *max_iter = j; return 1; }         // :449-456
    }
    *tol = resid;
    return 0;
}


// ------------------------------------------------------------------------------------------
// MINRES / PMINRES on the partitioned system -- num/krylovsolver.h:926-1161 (MResSolverCL :1918-1941, PMResSolverCL
// :1944-1972 with PLanczosONBCL over JACPcCL / DummyPcCL).  The serial driver of krylov.cu with every vector step on
// all partitions and every reduction as partial -> gather -> rank-ordered sum; the Lanczos scalars a1, b1 are fetched
// once per step (all ranks read identical values), the rotations run on the host of every rank, so all ranks stop in
// the same iteration.  Products carry the fused dot (a1 = (A q1 - b0 q0) . q1) over the peer-to-peer halo.
// ------------------------------------------------------------------------------------------
namespace {
struct DRing3 { int front = 0; int at(int k) const { int i = k; if (i < 0) i += 3; else if (i >= 3) i -= 3; i += front; return i < 3 ? i : i - 3; } void rotate() { front = (front + 1) % 3; } };
struct DRing2 { int front = 0; int at(int k) const { int i = k; if (i < 0) i += 2; else if (i >= 2) i -= 2; i += front; return i < 2 ? i : i - 2; } void rotate() { front = (front + 1) % 2; } };
inline void mr_gen_rot(double& dx, double& dy, double& cs, double& sn) { gm_gen_rot(dx, dy, cs, sn); }
inline void mr_app_rot(double& dx, double& dy, double cs, double sn) { gm_app_rot(dx, dy, cs, sn); }
}  // namespace

int dist_minres(int np, drops_b200_dist_matrix_s* const* P, double* const* x, const double* const* b, int precond, int pc,
                double omega, int* max_iter, double* tol, int relative)
{
    DB_CHECK(check_parts(np, P));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    DB_ARG(!precond || pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI, "unknown preconditioner");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    drops_b200_comm_s* comm = P[0]->comm;
    const bool pre = precond != 0;
    const bool jac = pre && pc == DROPS_B200_PC_JACOBI;
    {
        double flag = 0.0;
        for (int p = 0; p < np; ++p) if (jac && !P[p]->diag_ok) flag = 1.0;
        DB_CHECK(drops_b200_comm_allreduce_max(comm, &flag));
        if (flag != 0.0) { set_error("JACPcCL: at least one row (on some rank) stores no diagonal entry"); return DROPS_B200_E_NODIAG; }
    }
    // per partition: 0 tmp / r0, 1-3 q ring, 4-6 t ring, 7-9 p ring; scalars [0] a1, [1] b1 or b1^2, [2] |r|^2
    struct Part { double* tmp; double* qv[3]; double* tv[3]; double* pv[3]; double* d_scal; size_t n; };
    std::vector<Part> W((size_t)np);
    for (int p = 0; p < np; ++p) {
        Part& w = W[(size_t)p];
        w.n = P[p]->n_owned;
        DB_CHECK(ensure_part_scalars(P[p], 1));
        w.tmp = dwork(P[p], 0);
        for (int i = 0; i < 3; ++i) { w.qv[i] = dwork(P[p], 1 + i); w.tv[i] = dwork(P[p], 4 + i); w.pv[i] = dwork(P[p], 7 + i); }
        for (int i = 0; i < 3; ++i) if (!w.tmp || !w.qv[i] || !w.tv[i] || !w.pv[i]) return DROPS_B200_E_NOMEM;
        if (P[p]->small_alloc < 8) {
            dev_free(P[p]->d_small); P[p]->d_small = nullptr; P[p]->small_alloc = 0;
            DB_CHECK(dev_alloc(&P[p]->d_small, (size_t)8));
            P[p]->small_alloc = 8;
        }
        w.d_scal = P[p]->d_small;
    }
    std::vector<PartArgs> a((size_t)np);
    DRing3 q3, t3; DRing2 q2r, br;
    double bb[2] = {0.0, 0.0};
    auto Q = [&](Part& w, int k) -> double* { return pre ? w.qv[q2r.at(k)] : w.qv[q3.at(k)]; };
    auto T = [&](Part& w, int k) -> double* { return w.tv[t3.at(k)]; };
    auto B = [&](int k) -> double& { return bb[br.at(k)]; };
    auto reduce_to = [&](int phase, int idx) -> int {             // gathered partials -> d_scal[idx] on every partition
        DB_CHECK(allgather_partials(comm, 1));
        for (int p = 0; p < np; ++p) DB_CHECK(scalars_kernel(phase, P[p], nullptr, W[(size_t)p].d_scal + idx));
        return 0;
    };
    auto fetch_host = [&](int cnt) -> int {
        DB_CHECK(fetch_small(W[0].d_scal, c.h_scalars, (size_t)cnt));
        return check_p2p_faults(np, P);                           // a stale partial must not steer the rotations
    };
    auto apply_pc = [&](int p, double* out, const double* in) -> int {
        if (jac) return vec_jacobi(W[(size_t)p].n, omega, P[p]->diag, in, out);
        return vec_copy(W[(size_t)p].n, out, in);
    };
    double a0 = 0.0, norm_r0 = 0.0;
    bool nobreakdown = true;
    auto step = [&]() -> int {
        if (!pre) {                                                           // LanczosStep :926-943
            for (int p = 0; p < np; ++p) {
                Part& w = W[(size_t)p];
                a[(size_t)p] = PartArgs(); a[(size_t)p].p2p = true;
                a[(size_t)p].x = Q(w, 0); a[(size_t)p].y = Q(w, 1); a[(size_t)p].epilogue = EPI_AXMBY; a[(size_t)p].aux = Q(w, -1);
                a[(size_t)p].beta_val = B(-1); a[(size_t)p].dotv = Q(w, 0);    // q2 = A q1 - b0 q0; a1 = q2.q1
            }
            DB_CHECK(dist_spmv_run(np, P, a.data()));
            DB_CHECK(reduce_to(PH_DOT, 0));
            for (int p = 0; p < np; ++p) {
                Part& w = W[(size_t)p];
                DB_CHECK(lanczos_orth(w.n, Q(w, 1), Q(w, 0), w.d_scal + 0, nullptr, gather_slot(P[p], 1)));   // q2 -= a1 q1; |q2|^2
            }
            DB_CHECK(reduce_to(PH_NORM, 1));
            DB_CHECK(fetch_host(2));
            a0 = c.h_scalars[0];
            const double b1 = c.h_scalars[1];
            B(0) = b1;
            if (b1 < 1e-15) { nobreakdown = false; return 0; }
            for (int p = 0; p < np; ++p) DB_CHECK(vec_div_scalar(W[(size_t)p].n, Q(W[(size_t)p], 1), Q(W[(size_t)p], 1), b1, 0));
        } else {                                                              // PLanczosStep :989-1008
            for (int p = 0; p < np; ++p) {
                Part& w = W[(size_t)p];
                a[(size_t)p] = PartArgs(); a[(size_t)p].p2p = true;
                a[(size_t)p].x = Q(w, 0); a[(size_t)p].y = T(w, 1); a[(size_t)p].epilogue = EPI_AXMBY; a[(size_t)p].aux = T(w, -1);
                a[(size_t)p].beta_val = B(-1); a[(size_t)p].dotv = Q(w, 0);    // t2 = A q1 - b0 t0; a1 = t2.q1
            }
            DB_CHECK(dist_spmv_run(np, P, a.data()));
            DB_CHECK(reduce_to(PH_DOT, 0));
            for (int p = 0; p < np; ++p) {
                Part& w = W[(size_t)p];
                DB_CHECK(planczos_orth(w.n, T(w, 1), T(w, 0), w.d_scal + 0, Q(w, 1), pc, omega, P[p]->diag, nullptr, gather_slot(P[p], 1)));
            }
            DB_CHECK(reduce_to(PH_DOT, 1));
            DB_CHECK(fetch_host(2));
            a0 = c.h_scalars[0];
            const double b1 = std::sqrt(c.h_scalars[1]);
            B(0) = b1;
            if (b1 < 1e-15) { nobreakdown = false; return 0; }
            for (int p = 0; p < np; ++p) {
                Part& w = W[(size_t)p];
                DB_CHECK(vec_div_scalar(w.n, T(w, 1), T(w, 1), b1, 1));
                DB_CHECK(vec_div_scalar(w.n, Q(w, 1), Q(w, 1), b1, 1));
            }
        }
        nobreakdown = true;
        return 0;
    };
    auto true_residual_sq = [&]() -> int {                                    // tmp = b - A x; d_scal[2] = |tmp|^2 over all ranks
        for (int p = 0; p < np; ++p) {
            a[(size_t)p] = PartArgs(); a[(size_t)p].x = x[p]; a[(size_t)p].y = W[(size_t)p].tmp;
            a[(size_t)p].epilogue = EPI_RESID; a[(size_t)p].aux = b[p];
        }
        DB_CHECK(dist_spmv_run(np, P, a.data()));
        for (int p = 0; p < np; ++p) DB_CHECK(vec_dot_partial(W[(size_t)p].n, W[(size_t)p].tmp, W[(size_t)p].tmp, gather_slot(P[p], 1)));
        return reduce_to(PH_DOT, 2);
    };

    // r0 = b - A x :1158, new basis :958-967 / :1030-1040
    for (int p = 0; p < np; ++p) {
        a[(size_t)p] = PartArgs(); a[(size_t)p].x = x[p]; a[(size_t)p].y = W[(size_t)p].tmp;
        a[(size_t)p].epilogue = EPI_RESID; a[(size_t)p].aux = b[p];
    }
    DB_CHECK(dist_spmv_run(np, P, a.data()));
    if (!pre) {
        for (int p = 0; p < np; ++p) {
            Part& w = W[(size_t)p];
            DB_CHECK(vec_fill(w.n, Q(w, -1), 0.0));
            DB_CHECK(vec_dot_partial(w.n, w.tmp, w.tmp, gather_slot(P[p], 1)));
        }
        DB_CHECK(reduce_to(PH_DOT, 0));
        DB_CHECK(fetch_host(1));
        norm_r0 = std::sqrt(c.h_scalars[0]);
        for (int p = 0; p < np; ++p) {
            Part& w = W[(size_t)p];
            DB_CHECK(vec_div_scalar(w.n, Q(w, 0), w.tmp, norm_r0, 0));
            DB_CHECK(vec_fill(w.n, Q(w, 1), 0.0));
        }
    } else {
        for (int p = 0; p < np; ++p) {
            Part& w = W[(size_t)p];
            DB_CHECK(vec_fill(w.n, T(w, -1), 0.0));
            DB_CHECK(apply_pc(p, Q(w, -1), w.tmp));
            DB_CHECK(vec_dot_partial(w.n, Q(w, -1), w.tmp, gather_slot(P[p], 1)));
        }
        DB_CHECK(reduce_to(PH_DOT, 0));
        DB_CHECK(fetch_host(1));
        norm_r0 = std::sqrt(c.h_scalars[0]);
        for (int p = 0; p < np; ++p) {
            Part& w = W[(size_t)p];
            DB_CHECK(vec_div_scalar(w.n, T(w, 0), w.tmp, norm_r0, 0));
            DB_CHECK(vec_div_scalar(w.n, Q(w, 0), Q(w, -1), norm_r0, 0));
            DB_CHECK(vec_fill(w.n, T(w, 1), 0.0));
        }
    }
    B(-1) = 0.;
    DB_CHECK(step());

    // PMINRES :1069-1152
    double normb = std::fabs(norm_r0), res = norm_r0;
    bool lucky = !nobreakdown;
    double cc[3] = {0, 0, 0}, ss[3] = {0, 0, 0}, rr[3][3] = {{0}}, bq[2][2] = {{0}};
    DRing3 cr, sr, rrg, pr; DRing2 bqr;
    if (normb == 0.0 || !relative) normb = 1.0;
    if ((res = norm_r0 / normb) <= *tol) { *tol = res; *max_iter = 0; return 1; }
    for (int k = 1; k <= *max_iter; ++k) {
        double* R0 = rr[rrg.at(0)];
        double* B0 = bq[bqr.at(0)];
        const double* Bm1 = bq[bqr.at(-1)];
        int mode;
        switch (k) {
        case 1:
            mr_gen_rot(a0, B(0), cc[cr.at(0)], ss[sr.at(0)]);
            R0[0] = std::sqrt(a0 * a0 + B(0) * B(0));
            B0[0] = 1.; B0[1] = 0.;
            mr_app_rot(B0[0], B0[1], cc[cr.at(0)], ss[sr.at(0)]);
            mode = 1;
            break;
        case 2:
            R0[0] = B(-1); R0[1] = a0; R0[2] = B(0);
            mr_app_rot(R0[0], R0[1], cc[cr.at(-1)], ss[sr.at(-1)]);
            mr_gen_rot(R0[1], R0[2], cc[cr.at(0)], ss[sr.at(0)]);
            mr_app_rot(R0[1], R0[2], cc[cr.at(0)], ss[sr.at(0)]);
            B0[0] = Bm1[1]; B0[1] = 0.;
            mr_app_rot(B0[0], B0[1], cc[cr.at(0)], ss[sr.at(0)]);
            mode = 2;
            break;
        default: {
            R0[0] = 0.; R0[1] = B(-1); R0[2] = a0;
            double tmpb = B(0);
            mr_app_rot(R0[0], R0[1], cc[cr.at(-2)], ss[sr.at(-2)]);
            mr_app_rot(R0[1], R0[2], cc[cr.at(-1)], ss[sr.at(-1)]);
            mr_gen_rot(R0[2], tmpb, cc[cr.at(0)], ss[sr.at(0)]);
            mr_app_rot(R0[2], tmpb, cc[cr.at(0)], ss[sr.at(0)]);
            B0[0] = Bm1[1]; B0[1] = 0.;
            mr_app_rot(B0[0], B0[1], cc[cr.at(0)], ss[sr.at(0)]);
            mode = 3;
        } }
        for (int p = 0; p < np; ++p) {                                        // p[0] = ...; x += norm_r0 b[0][0] p[0]  :1095-1129
            Part& w = W[(size_t)p];
            DB_CHECK(minres_update(w.n, mode, Q(w, 0), w.pv[pr.at(-2)], w.pv[pr.at(-1)], w.pv[pr.at(0)], x[p],
                                   R0[0], R0[1], R0[2], norm_r0 * B0[0]));
        }
        res = std::fabs(norm_r0 * B0[1]) / normb;
        if (res <= *tol || lucky) {                                           // :1132-1141
            DB_CHECK(true_residual_sq());
            DB_CUDA(cudaMemcpyAsync(c.h_scalars, W[0].d_scal + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
            DB_CUDA(cudaStreamSynchronize(st));
            DB_CHECK(check_p2p_faults(np, P));
            const double realRes = std::sqrt(c.h_scalars[0]);
            if (!(realRes > *tol)) { *tol = res; *max_iter = k; return 1; }
        }
        if (pre) { q2r.rotate(); t3.rotate(); } else q3.rotate();            // :973-978 / :1046-1051
        br.rotate();
        DB_CHECK(step());
        if (!nobreakdown) lucky = true;
        cr.rotate(); sr.rotate(); rrg.rotate(); pr.rotate(); bqr.rotate();
    }
    *tol = res;
    return 0;
}

}  // namespace db200

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace db200;

extern "C" {

int drops_b200_comm_unique_id(void* id_bytes)
{
    DB_ARG(id_bytes != nullptr, "null id buffer");
    static_assert(sizeof(ncclUniqueId) <= DROPS_B200_UNIQUE_ID_BYTES, "unique id size");
    if (!nccl().ok) { set_error("libnccl.so.2 could not be loaded"); return DROPS_B200_E_COMM; }
    ncclUniqueId id;
    DB_NCCL(nccl().GetUniqueId(&id));
    std::memset(id_bytes, 0, DROPS_B200_UNIQUE_ID_BYTES);
    std::memcpy(id_bytes, &id, sizeof id);
    return 0;
}

static int comm_common(drops_b200_comm_s* C)
{
    int prio_lo = 0, prio_hi = 0;      // the halo must not queue behind the interior-row SpMV
    DB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    DB_CUDA(cudaStreamCreateWithPriority(&C->comm_stream, cudaStreamNonBlocking, prio_hi));
    DB_CUDA(cudaEventCreateWithFlags(&C->ev_pack, cudaEventDisableTiming));
    DB_CUDA(cudaEventCreateWithFlags(&C->ev_halo, cudaEventDisableTiming));
    DB_CUDA(cudaEventCreateWithFlags(&C->ev_bnd, cudaEventDisableTiming));
    DB_CHECK(dev_alloc(&C->gathered, (size_t)C->world * kMaxQ));
    DB_CUDA(cudaMemset(C->gathered, 0, (size_t)C->world * kMaxQ * sizeof(dd)));
    DB_CHECK(dev_alloc(&C->wide, (size_t)C->world * kWideQ));
    DB_CUDA(cudaMemset(C->wide, 0, (size_t)C->world * kWideQ * sizeof(dd)));
    // buffer of the peer-to-peer reductions: 2 parity halves of world x kMaxQ partials, then world flags; one
    // allocation of >= 2 MB so that a single CUDA IPC handle maps it into the peers
    {
        const size_t flag_off = ((size_t)2 * C->world * kMaxQ * sizeof(dd) + 255) / 256 * 256;
        const size_t bytes = std::max<size_t>(flag_off + (size_t)C->world * sizeof(unsigned), (size_t)2 << 20);
        char* raw = nullptr;
        DB_CHECK(dev_alloc(&raw, bytes));
        DB_CUDA(cudaMemset(raw, 0, bytes));
        C->red_buf = reinterpret_cast<dd*>(raw);
        C->red_flags = reinterpret_cast<unsigned*>(raw + flag_off);
        DB_CHECK(dev_alloc(&C->red_fault, 1));
        DB_CUDA(cudaMemset(C->red_fault, 0, sizeof(int)));
    }
    return 0;
}

// peer tables of the reductions: bufs[q] / flags[q] = rank q's buffer and flag array in MY address space
static int comm_p2p_finish(drops_b200_comm_s* C, const std::vector<dd*>& bufs, const std::vector<unsigned*>& flags)
{
    DB_CHECK(dev_alloc(&C->d_peer_red, (size_t)C->world));
    DB_CHECK(dev_alloc(&C->d_peer_redflag, (size_t)C->world));
    DB_CUDA(cudaMemcpy(C->d_peer_red, bufs.data(), (size_t)C->world * sizeof(dd*), cudaMemcpyHostToDevice));
    DB_CUDA(cudaMemcpy(C->d_peer_redflag, flags.data(), (size_t)C->world * sizeof(unsigned*), cudaMemcpyHostToDevice));
    C->p2p = 1;
    return 0;
}

static int comm_p2p_setup_nccl(drops_b200_comm_s* C)
{
    C->p2p = -1;
    if (opt("dist.p2p", 1.0) == 0.0 || C->world < 2 || !nccl().ok || C->nccl == nullptr) return 0;
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    const int world = C->world, me = C->rank;
    const size_t blob = sizeof(cudaIpcMemHandle_t) + 16;      // handle | flag offset | status
    std::vector<unsigned char> mine(blob, 0), all(blob * (size_t)world, 0);
    cudaIpcMemHandle_t h;
    long long ok = 1;
    if (cudaIpcGetMemHandle(&h, C->red_buf) != cudaSuccess) { cudaGetLastError(); ok = 0; std::memset(&h, 0, sizeof h); }
    const unsigned long long foff = (unsigned long long)((char*)C->red_flags - (char*)C->red_buf);
    std::memcpy(mine.data(), &h, sizeof h);
    std::memcpy(mine.data() + sizeof h, &foff, 8);
    std::memcpy(mine.data() + sizeof h + 8, &ok, 8);
    unsigned char* d_all = nullptr;
    DB_CHECK(dev_alloc(&d_all, blob * (size_t)world));
    cudaError_t e = cudaMemcpyAsync(d_all + blob * (size_t)me, mine.data(), blob, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { dev_free(d_all); return cuda_fail(e, "p2p reduce setup", __FILE__, __LINE__); }
    ncclResult_t nr = nccl().AllGather(d_all + blob * (size_t)me, d_all, blob, ncclChar, (ncclComm_t)C->nccl, st);
    if (nr != ncclSuccess) { dev_free(d_all); return nccl_fail(nr, "p2p reduce handle all-gather"); }
    e = cudaMemcpyAsync(all.data(), d_all, blob * (size_t)world, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(d_all);
    if (e != cudaSuccess) return cuda_fail(e, "p2p reduce setup", __FILE__, __LINE__);
    std::vector<dd*> bufs((size_t)world, nullptr);
    std::vector<unsigned*> flags((size_t)world, nullptr);
    bool good = true;
    for (int q = 0; q < world && good; ++q) {
        const unsigned char* b = all.data() + blob * (size_t)q;
        long long okq = 0; unsigned long long fq = 0;
        std::memcpy(&h, b, sizeof h); std::memcpy(&fq, b + sizeof h, 8); std::memcpy(&okq, b + sizeof h + 8, 8);
        if (!okq) { good = false; break; }
        if (q == me) { bufs[(size_t)q] = C->red_buf; flags[(size_t)q] = C->red_flags; continue; }
        void* pm = nullptr;
        if (cudaIpcOpenMemHandle(&pm, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; break; }
        C->ipc_mapped.push_back(pm);
        bufs[(size_t)q] = (dd*)pm; flags[(size_t)q] = (unsigned*)((char*)pm + fq);
    }
    double bad = good ? 0.0 : 1.0;                         // all ranks or none
    DB_CHECK(drops_b200_comm_allreduce_max(C, &bad));
    if (bad != 0.0) {
        for (void* m : C->ipc_mapped) cudaIpcCloseMemHandle(m);
        C->ipc_mapped.clear();
        return 0;
    }
    return comm_p2p_finish(C, bufs, flags);
}

int drops_b200_comm_create(int rank, int world, const void* id_bytes, drops_b200_comm_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr, "null out");
    *out = nullptr;
    DB_ARG(world >= 1 && rank >= 0 && rank < world, "rank/world out of range");
    drops_b200_comm_s* C = new drops_b200_comm_s;
    C->kind = COMM_NCCL; C->rank = rank; C->world = world;
    int s = comm_common(C);
    if (s == 0 && world > 1) {
        if (!id_bytes) { set_error("comm_create: unique id required for world > 1"); s = DROPS_B200_E_ARG; }
        else if (!nccl().ok) { set_error("libnccl.so.2 could not be loaded"); s = DROPS_B200_E_COMM; }
        else {
            ncclUniqueId id;
            std::memcpy(&id, id_bytes, sizeof id);
            ncclComm_t nc = nullptr;
            ncclResult_t r = nccl().CommInitRank(&nc, world, id, rank);
            if (r != ncclSuccess) s = nccl_fail(r, "ncclCommInitRank");
            C->nccl = nc;
            if (s == 0) s = comm_p2p_setup_nccl(C);          // collective; failing to map peers is not an error
        }
    }
    if (s < 0) { drops_b200_comm_destroy(C); return s; }
    *out = C;
    return 0;
}

int drops_b200_comm_create_local(int world, drops_b200_comm_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr && world >= 1, "bad argument");
    drops_b200_comm_s* C = new drops_b200_comm_s;
    C->kind = COMM_LOCAL; C->rank = 0; C->world = world;
    C->parts.assign((size_t)world, nullptr);
    int s = comm_common(C);
    if (s == 0 && world > 1 && opt("dist.p2p", 1.0) != 0.0) {     // all ranks live here: every "peer" buffer is this one
        std::vector<dd*> bufs((size_t)world, C->red_buf);
        std::vector<unsigned*> flags((size_t)world, C->red_flags);
        s = comm_p2p_finish(C, bufs, flags);
    }
    if (s < 0) { drops_b200_comm_destroy(C); *out = nullptr; return s; }
    *out = C;
    return 0;
}

int drops_b200_comm_destroy(drops_b200_comm_t C)
{
    if (!C) return 0;
    if (ctx().inited) {
        cudaStreamSynchronize(ctx().stream);
        if (C->comm_stream) cudaStreamSynchronize(C->comm_stream);
    }
    if (C->nccl && nccl().ok) nccl().CommDestroy((ncclComm_t)C->nccl);
    if (C->ev_pack) cudaEventDestroy(C->ev_pack);
    if (C->ev_halo) cudaEventDestroy(C->ev_halo);
    if (C->ev_bnd) cudaEventDestroy(C->ev_bnd);
    if (C->comm_stream) cudaStreamDestroy(C->comm_stream);
    dev_free(C->gathered); dev_free(C->wide);
    for (void* m : C->ipc_mapped) cudaIpcCloseMemHandle(m);
    dev_free(C->red_buf); dev_free(C->red_fault); dev_free(C->d_peer_red); dev_free(C->d_peer_redflag);
    delete C;
    return 0;
}

int drops_b200_comm_info(drops_b200_comm_t C, int* rank, int* world, int* is_local)
{
    DB_ARG(C != nullptr, "null comm");
    if (rank) *rank = C->rank;
    if (world) *world = C->world;
    if (is_local) *is_local = C->kind == COMM_LOCAL;
    return 0;
}

int drops_b200_comm_allreduce_max(drops_b200_comm_t C, double* value)
{
    DB_CHECK(require_ctx());
    DB_ARG(C != nullptr && value != nullptr, "null argument");
    Ctx& c = ctx();
    DB_CUDA(cudaStreamSynchronize(c.stream));
    if (C->kind == COMM_LOCAL || C->world == 1) return 0;
    double* d = c.scratch_scalars + 8;
    DB_CUDA(cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, c.stream));
    DB_NCCL(nccl().AllReduce(d, d, 1, ncclDouble, ncclMax, (ncclComm_t)C->nccl, c.stream));
    DB_CUDA(cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    DB_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int drops_b200_dist_matrix_create_device(drops_b200_comm_t comm, drops_b200_exchange_t ex, size_t nnz,
                                         const int64_t* d_rowptr, const int64_t* d_colind_global, const double* d_val,
                                         uint64_t version, drops_b200_dist_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr && d_rowptr != nullptr, "null argument");
    DB_ARG(nnz == 0 || (d_colind_global != nullptr && d_val != nullptr), "null colind/val");
    return dist_matrix_build(comm, ex, nnz, d_rowptr, d_colind_global, d_val, version, out);
}

int drops_b200_dist_matrix_create(drops_b200_comm_t comm, drops_b200_exchange_t ex, size_t nnz, const size_t* rowbeg,
                                  const size_t* colind_global, const double* val, uint64_t version,
                                  drops_b200_dist_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr && ex != nullptr && rowbeg != nullptr, "null argument");
    DB_ARG(nnz == 0 || (colind_global != nullptr && val != nullptr), "null colind/val");
    const size_t n = ex->owned.size();
    DB_ARG(rowbeg[0] == 0 && rowbeg[n] == nnz, "rowbeg[0] != 0 or rowbeg[n_owned] != nnz");
    cudaStream_t st = ctx().stream;
    int64_t *rp = nullptr, *cg = nullptr; double* v = nullptr;
    int s;
    if ((s = dev_alloc(&rp, n + 1)) < 0 || (s = dev_alloc(&cg, nnz)) < 0 || (s = dev_alloc(&v, nnz)) < 0) {
        dev_free(rp); dev_free(cg); dev_free(v);
        return s;
    }
    static_assert(sizeof(size_t) == sizeof(int64_t), "size_t must be 64 bit");
    cudaError_t e = cudaMemcpyAsync(rp, rowbeg, (n + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(cg, colind_global, nnz * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(v, val, nnz * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) s = cuda_fail(e, "partition upload", __FILE__, __LINE__);
    else s = dist_matrix_build(comm, ex, nnz, rp, cg, v, version, out);
    dev_free(rp); dev_free(cg); dev_free(v);
    return s;
}

int drops_b200_dist_matrix_destroy(drops_b200_dist_matrix_t A)
{
    free_dist_matrix(A);
    return 0;
}

int drops_b200_dist_matrix_info(drops_b200_dist_matrix_t A, size_t* n_owned, size_t* n_ghost, size_t* nnz,
                                size_t* interior_rows, size_t* boundary_rows, size_t* n_send, size_t* device_bytes)
{
    DB_ARG(A != nullptr, "null matrix");
    if (n_owned) *n_owned = A->n_owned;
    if (n_ghost) *n_ghost = A->n_ghost;
    if (nnz) *nnz = A->nnz;
    if (interior_rows) *interior_rows = A->n_int;
    if (boundary_rows) *boundary_rows = A->n_bnd;
    if (n_send) *n_send = A->n_send;
    if (device_bytes) *device_bytes = A->device_bytes;
    return 0;
}

int drops_b200_dist_matrix_segment_stream(drops_b200_dist_matrix_t A, size_t* n_slices, size_t* n_entries, size_t* padded_entries)
{
    DB_ARG(A != nullptr, "null matrix");
    const bool on = A->sell_int.scd != nullptr && A->npos > 0;
    if (n_slices) *n_slices = on ? A->sell_int.sc_slices : 0;
    if (n_entries) *n_entries = on ? A->sell_int.sc_entries : 0;
    if (padded_entries) *padded_entries = A->sell_int.padded;
    return 0;
}

int drops_b200_dist_matrix_format(drops_b200_dist_matrix_t A, int* interior_packed, int* n_distinct)
{
    DB_ARG(A != nullptr, "null matrix");
    if (interior_packed)
        *interior_packed = (A->sell_int.pk != nullptr ? 1 : 0) | (A->sell_int.vidx != nullptr ? 2 : 0) | (A->npos > 0 ? 4 : 0) |
                           (A->block_ordered ? 8 : 0) | (A->p2p == 1 ? 16 : 0);
    if (interior_packed && A->p2p == 1 && A->p2p_fault && ctx().inited) {
        int f = 0;
        cudaStreamSynchronize(ctx().stream);
        cudaMemcpy(&f, A->p2p_fault, sizeof(int), cudaMemcpyDeviceToHost);
        if (f) *interior_packed |= 32;
    }
    if (interior_packed && A->comm && A->comm->p2p == 1 && ctx().inited) {
        int f = 0;
        cudaStreamSynchronize(ctx().stream);
        cudaMemcpy(&f, A->comm->red_fault, sizeof(int), cudaMemcpyDeviceToHost);
        *interior_packed |= 64;                      // reductions go peer to peer as well
        if (f) *interior_packed |= 32;
    }
    if (interior_packed && A->diag_ro_idx != nullptr && A->diag_ro_count > 0 && opt("pcg.diag_index", 1.0) != 0.0)
        *interior_packed |= 128;                     // the update kernel reads the Jacobi diagonal as a byte index
    if (interior_packed && A->sell_int.scd != nullptr && A->npos > 0 && opt("spmv.segments", 1.0) != 0.0)
        *interior_packed |= 256;                     // interior rows run from segment descriptors (SellOp::scd)
    if (n_distinct) *n_distinct = A->sell_int.ndict;
    return 0;
}

int drops_b200_dist_matrix_wait_stats(drops_b200_dist_matrix_t A, double* out8, int reset)
{
    DB_CHECK(require_ctx());
    DB_ARG(A != nullptr && out8 != nullptr, "null argument");
    for (int k = 0; k < 8; ++k) out8[k] = 0.0;
    if (!A->wait_stats) return 0;                    // no peer path on this partition: nothing waits inside a kernel
    unsigned long long h[8];
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (A->comm && A->comm->comm_stream) DB_CUDA(cudaStreamSynchronize(A->comm->comm_stream));
    DB_CUDA(cudaMemcpy(h, A->wait_stats, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) out8[k] = (double)h[k];
    if (reset) DB_CUDA(cudaMemset(A->wait_stats, 0, sizeof(h)));
    return 0;
}

int drops_b200_dist_matrix_diag(drops_b200_dist_matrix_t A, double* diag)
{
    DB_CHECK(require_ctx());
    DB_ARG(A != nullptr && diag != nullptr, "null argument");
    if (A->n_owned) DB_CUDA(cudaMemcpyAsync(diag, A->diag, A->n_owned * 8, cudaMemcpyDeviceToHost, ctx().stream));
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

int drops_b200_dist_spmv_dev(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                             const drops_b200_vector_t* y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y, "null argument");
    DB_CHECK(check_parts(nparts, A));
    std::vector<PartArgs> a((size_t)nparts);
    for (int p = 0; p < nparts; ++p) {
        DB_ARG(x[p] && y[p] && x[p]->n == A[p]->n_owned && y[p]->n == A[p]->n_owned, "vector size != owned rows of the partition");
        DB_ARG(x[p]->d != y[p]->d, "x and y must not alias");
        DB_CHECK(ensure_part_scalars(A[p], 1));
        a[(size_t)p].x = x[p]->d; a[(size_t)p].y = y[p]->d;
    }
    return dist_spmv_run(nparts, A, a.data());
}

int drops_b200_dist_dot(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                        const drops_b200_vector_t* y, double* result)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y && result, "null argument");
    DB_CHECK(check_parts(nparts, A));
    Ctx& c = ctx();
    for (int p = 0; p < nparts; ++p) {
        DB_ARG(x[p] && y[p] && x[p]->n == A[p]->n_owned && y[p]->n == A[p]->n_owned, "vector size != owned rows of the partition");
        DB_CHECK(ensure_part_scalars(A[p], 1));
        DB_CHECK(vec_dot_partial(A[p]->n_owned, x[p]->d, y[p]->d, gather_slot(A[p], 1)));
    }
    DB_CHECK(allgather_partials(A[0]->comm, 1));
    DB_CHECK(scalars_kernel(PH_DOT, A[0], nullptr, c.scratch_scalars));
    DB_CUDA(cudaMemcpyAsync(c.h_scalars, c.scratch_scalars, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    DB_CUDA(cudaStreamSynchronize(c.stream));
    *result = c.h_scalars[0];
    return 0;
}

int drops_b200_dist_gmres_dev(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                              const drops_b200_vector_t* b, int pc, double omega, int restart, int* max_iter, double* tol,
                              int relative, int left_precond)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    DB_CHECK(check_parts(nparts, A));
    std::vector<double*> xs((size_t)nparts);
    std::vector<const double*> bs((size_t)nparts);
    for (int p = 0; p < nparts; ++p) {
        DB_ARG(x[p] && b[p] && x[p]->n == A[p]->n_owned && b[p]->n == A[p]->n_owned, "vector size != owned rows of the partition");
        xs[(size_t)p] = x[p]->d; bs[(size_t)p] = b[p]->d;
    }
    return dist_gmres(nparts, A, xs.data(), bs.data(), pc, omega, restart, max_iter, tol, relative, left_precond);
}

int drops_b200_dist_minres_dev(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                               const drops_b200_vector_t* b, int precond, int pc, double omega, int* max_iter, double* tol,
                               int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    DB_CHECK(check_parts(nparts, A));
    std::vector<double*> xs((size_t)nparts);
    std::vector<const double*> bs((size_t)nparts);
    for (int p = 0; p < nparts; ++p) {
        DB_ARG(x[p] && b[p] && x[p]->n == A[p]->n_owned && b[p]->n == A[p]->n_owned, "vector size != owned rows of the partition");
        xs[(size_t)p] = x[p]->d; bs[(size_t)p] = b[p]->d;
    }
    return dist_minres(nparts, A, xs.data(), bs.data(), precond, pc, omega, max_iter, tol, relative);
}

int drops_b200_dist_pcg_dev(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                            const drops_b200_vector_t* b, int pc, double omega, int* max_iter, double* tol, int relative,
                            double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    DB_CHECK(check_parts(nparts, A));
    std::vector<double*> xs((size_t)nparts);
    std::vector<const double*> bs((size_t)nparts);
    for (int p = 0; p < nparts; ++p) {
        DB_ARG(x[p] && b[p] && x[p]->n == A[p]->n_owned && b[p]->n == A[p]->n_owned, "vector size != owned rows of the partition");
        xs[(size_t)p] = x[p]->d; bs[(size_t)p] = b[p]->d;
    }
    return dist_pcg(nparts, A, xs.data(), bs.data(), pc, omega, max_iter, tol, relative, hist, hist_cap);
}

}  // extern "C"

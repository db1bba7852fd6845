// common.cuh -- shared device helpers and the process-wide context of libdrops_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/drops_b200.h"

namespace db200 {

// ---------------------------------------------------------------------------------------------
// Error plumbing: nothing throws across the C ABI.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define DB_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return ::db200::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define DB_CHECK(call)                                                             \
    do {                                                                           \
        int s__ = (call);                                                          \
        if (s__ < 0) return s__;                                                   \
    } while (0)

#define DB_ARG(cond, msg)                                                          \
    do {                                                                           \
        if (!(cond)) { ::db200::set_error("%s: %s", __func__, msg); return DROPS_B200_E_ARG; } \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Double-double accumulation for reductions.  The reference sums dot products with Kahan
// compensation (num/spmat.h:317-333); a double-double tree has a smaller error bound and is
// order-independent up to O(eps^2), which keeps Krylov iteration counts aligned.
// ---------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};

__host__ __device__ inline dd dd_zero() { return dd{0.0, 0.0}; }

#ifdef __CUDACC__
__device__ __forceinline__ dd two_sum(double a, double b)
{
    const double s  = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    const double e  = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    return dd{s, e};
}
__device__ __forceinline__ dd quick_two_sum(double a, double b)   // |a| >= |b|
{
    const double s = __dadd_rn(a, b);
    const double e = __dsub_rn(b, __dsub_rn(s, a));
    return dd{s, e};
}
__device__ __forceinline__ dd dd_add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    s.lo = __dadd_rn(s.lo, __dadd_rn(a.lo, b.lo));
    return quick_two_sum(s.hi, s.lo);
}
// acc += a*b, compensated (the Dot2 step of Ogita/Rump/Oishi): exact product via the FMA residual,
// error-free sum of the high parts, all error terms collected in acc.lo WITHOUT renormalising.
// The loop-carried dependency is the single add on acc.hi; the error chain runs beside it.  (The
// renormalising variant serialised ~10 fp64 operations per row at the tail of every SpMV slice,
// profiles/r01_notes.md.)  The result is as accurate as a sum in twice the working precision,
// i.e. better than the reference's Kahan sum of rounded products (num/spmat.h:317-333).
__device__ __forceinline__ dd dd_add_prod(dd acc, double a, double b)
{
    const double p  = __dmul_rn(a, b);
    const double ep = __fma_rn(a, b, -p);
    const double s  = __dadd_rn(acc.hi, p);
    const double bb = __dsub_rn(s, acc.hi);
    const double es = __dadd_rn(__dsub_rn(acc.hi, __dsub_rn(s, bb)), __dsub_rn(p, bb));
    return dd{s, __dadd_rn(acc.lo, __dadd_rn(ep, es))};
}
__device__ __forceinline__ double dd_round(dd a) { return __dadd_rn(a.hi, a.lo); }

__device__ __forceinline__ dd dd_shfl_xor(dd v, int mask)
{
    dd o;
    o.hi = __shfl_xor_sync(0xffffffffu, v.hi, mask);
    o.lo = __shfl_xor_sync(0xffffffffu, v.lo, mask);
    return o;
}
__device__ __forceinline__ dd warp_reduce_dd(dd v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = dd_add(v, dd_shfl_xor(v, m));
    return v;
}

// Block-wide sum of NQ double-double quantities (fixed tree: lanes, then warps in order).
// Result valid in thread 0.  smem must hold NQ * 32 dd.
template <int NQ>
__device__ __forceinline__ void block_reduce_dd(dd (&v)[NQ], dd* smem)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = warp_reduce_dd(v[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) smem[q * 32 + warp] = v[q];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            dd t = lane < nwarp ? smem[q * 32 + lane] : dd_zero();
            v[q] = warp_reduce_dd(t);
        }
    }
    __syncthreads();
}

// Grid-wide deterministic reduction.  Every block stores its NQ partials and takes a ticket; the
// block that draws the last ticket sums the partials in block-index order (independent of which
// block happens to be last) and returns true in ALL its threads with the totals in thread 0's
// v[].  Grids of more than kReduceFlat blocks (the SpMV launches ~10^5 short CTAs) are reduced in
// two levels -- groups of kReduceGroup consecutive blocks, then the group sums -- so that the
// serial tail executed by a single CTA stays a few loads per thread instead of gridDim/blockDim
// (measured on B200: 0.8 ms of a 3.9 ms SpMV with the flat scheme, profiles/r01_notes.md).
// `partials` holds kPartialStride dd (NQ*gridDim block partials, then NQ*ngroups group partials),
// `ticket` kTicketStride counters that are zero on entry and are reset before the kernel ends.
constexpr unsigned kReduceFlat = 2048;
constexpr unsigned kReduceGroup = 256;

template <int NQ>
__device__ __forceinline__ void sum_partials_dd(dd (&acc)[NQ], const dd* partials, unsigned stride,
                                                unsigned begin, unsigned end, dd* smem)
{
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        acc[q] = dd_zero();
        // fixed assignment of partials to threads, fixed order inside a thread
        for (unsigned b = begin + threadIdx.x; b < end; b += blockDim.x) {
            const volatile dd* p = partials + (size_t)q * stride + b;
            dd t; t.hi = p->hi; t.lo = p->lo;
            acc[q] = dd_add(acc[q], t);
        }
    }
    block_reduce_dd<NQ>(acc, smem);
}

template <int NQ>
__device__ __forceinline__ bool grid_reduce_dd(dd (&v)[NQ], dd* smem, dd* partials, unsigned* ticket)
{
    __shared__ bool is_last;
    block_reduce_dd<NQ>(v, smem);
    const unsigned nblk = gridDim.x;
    dd acc[NQ];
    if (nblk <= kReduceFlat) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) partials[(size_t)q * nblk + blockIdx.x] = v[q];
            __threadfence();
            is_last = (atomicAdd(ticket, 1u) == nblk - 1);
        }
        __syncthreads();
        if (!is_last) return false;
        __threadfence();
        sum_partials_dd<NQ>(acc, partials, nblk, 0u, nblk, smem);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) v[q] = acc[q];
            *ticket = 0u;
        }
        return true;
    }
    // ---- level 1: the last block of each group sums the group's block partials ------------------
    const unsigned ngroups = (nblk + kReduceGroup - 1) / kReduceGroup;
    const unsigned group = blockIdx.x / kReduceGroup;
    const unsigned gbegin = group * kReduceGroup;
    const unsigned gend = gbegin + kReduceGroup < nblk ? gbegin + kReduceGroup : nblk;
    dd* gpartials = partials + (size_t)NQ * nblk;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) partials[(size_t)q * nblk + blockIdx.x] = v[q];
        __threadfence();
        is_last = (atomicAdd(ticket + 1 + group, 1u) == gend - gbegin - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    sum_partials_dd<NQ>(acc, partials, nblk, gbegin, gend, smem);
    // ---- level 2: the last group finisher sums the group partials ------------------------------
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) gpartials[(size_t)q * ngroups + group] = acc[q];
        ticket[1 + group] = 0u;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == ngroups - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    sum_partials_dd<NQ>(acc, gpartials, ngroups, 0u, ngroups, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = acc[q];
        *ticket = 0u;
    }
    return true;
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// Process-wide context
// ---------------------------------------------------------------------------------------------
struct TimingSlot {
    double ms = 0.0;
    long long launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct Ctx {
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    int cc = 0;
    cudaStream_t stream = nullptr;      // all solver work
    cudaStream_t copy_stream = nullptr; // staging copies
    // reduction workspace
    dd* partials = nullptr;             // MAX_Q * max_grid
    unsigned* tickets = nullptr;        // kPartialSlots * kTicketStride counters
    double* scratch_scalars = nullptr;  // device doubles for standalone dot/norm results
    double* h_scalars = nullptr;        // pinned host mirror
    // small results the host waits for inside a solver loop (Hessenberg columns, Lanczos scalars): a kernel stores them
    // into mapped pinned memory and raises a sequence number the host polls -- 2-3 us instead of the ~10 us of a
    // copy + stream synchronisation ("fetch.poll", fetch_small)
    double* h_pub = nullptr;            // kPubDoubles payload doubles, then the sequence word
    double* d_pub = nullptr;            // the same memory, device view
    unsigned pub_seq = 0;
    int max_grid = 0;
    long long launches = 0;
    unsigned long long options_epoch = 0;   // bumped by every set_option: cached CUDA graphs are keyed on it
    std::map<std::string, double> options;
    std::map<std::string, TimingSlot> timing;
    std::vector<cudaEvent_t> event_pool;
};

Ctx& ctx();
int require_ctx();
constexpr int kPubDoubles = 600;
// hdst[0 .. count) = dsrc[0 .. count) after everything enqueued on the solver stream so far; count doubles
int fetch_small(const double* dsrc, double* hdst, size_t count);

double opt(const char* key, double dflt);

// RAII-less kernel timing: begin/end record events on ctx().stream when option "timing" != 0.
struct ScopedTiming {
    TimingSlot* slot = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t stream = nullptr;
    explicit ScopedTiming(const char* key, cudaStream_t st = nullptr);     // st: the stream the timed launches go to (default: the solver stream)
    ~ScopedTiming();
};
int timing_flush();

inline void count_launch(int n = 1) { ctx().launches += n; }

constexpr int kMaxQ = 4;            // quantities per fused reduction
constexpr int kWideQ = 512;         // quantities of a wide reduction (classical Gram-Schmidt column of the partitioned GMRES)
constexpr int kPartialStride = 1 << 20;   // dd entries of reduction workspace per slot (>= NQ * grid)
constexpr int kPartialSlots = 4;          // kernels that may overlap use different slots
constexpr int kTicketStride = 8192;       // counters per slot: 1 + groups of the two-level grid reduction
constexpr int kMaxReduceGrid = kPartialStride / 2;   // NQ = 1 kernels; fused NQ > 1 kernels use <= 8 CTAs/SM
constexpr int kReduceThreads = 256;

// Grid for streaming vector kernels: a multiple of the SM count.
int vec_grid(size_t n, int threads, int per_thread);

// device allocation helpers (status-returning)
template <typename T>
inline int dev_alloc(T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return DROPS_B200_E_NOMEM;
    }
    return 0;
}
inline void dev_free(void* p) { if (p) cudaFree(p); }

}  // namespace db200

// ---------------------------------------------------------------------------------------------
// Handle types behind the C ABI
// ---------------------------------------------------------------------------------------------
struct drops_b200_vector_s {
    size_t n = 0;
    double* d = nullptr;
    bool owns = true;
};

namespace db200 {

// Sliced-ELL (C = 32, sigma-sorted) operator + the CSR it was built from.
constexpr int kScMaxWidth = 80;      // widest slice the segment kernel stages in shared memory (P2 brick: 65)
constexpr int kDictConst = 48;       // dictionary entries (values + the padding slot) a kernel parameter block carries

struct SellOp {
    size_t rows = 0, cols = 0;
    size_t nslices = 0;
    size_t padded = 0;               // stored entries incl. padding
    int64_t* slice_ptr = nullptr;    // nslices+1 offsets into col/val (multiples of 32)
    int32_t* col = nullptr;          // -1 = padding
    double*  val = nullptr;
    int32_t* perm = nullptr;         // sorted position -> original row, -1 beyond rows
    int max_width = 0;
    bool full_rows = true;           // built over all rows of the CSR (false: a row subset, dist.cu)
    bool general_released = false;   // col / val were freed after the packed stream was built ("sell.release_general")
    // Value dictionary ("sell.valdict"): constant-coefficient systems on uniform meshes hold a handful
    // of distinct fp64 values (15 on the P2 Kuhn brick).  When the matrix has at most 65535 distinct
    // bit patterns, `dict` lists them (sorted by bit pattern) and every stored entry becomes one
    // 32-bit word  pk = (dict index << cbits) | column  (padding: index ndict = 0.0, column 0) -- lossless, so the
    // product stays bit-identical, at 4 instead of 12 bytes per entry on the stream that bounds
    // the kernel.  pk is nullptr when column bits + index bits do not fit 32 (then col/val are used).
    uint32_t* pk = nullptr;          // padded entries
    double* dict = nullptr;          // ndict distinct values + one 0.0 (the padding slot)
    int ndict = 0;
    double dict_host[kDictConst] = {};   // the same list on the host when it fits: travels as a kernel parameter
    int ndict_host = 0;                  // (constant bank: lookups leave the L1 data pipe, spmv_pk.cu "spmv.pk_cdict")
    int cbits = 0;                   // low bits of pk that hold the column
    // Row-ordered twin of pk (square operators over all rows): the unknowns renumbered by the
    // operator's own slice order -- position s*32+lane of the sorted windows -- so that row `pos`
    // IS y[pos] (no perm load, coalesced stores) and the neighbours of 32 rows of equal shape are
    // neighbours in x (a gather touches 8-9 sectors instead of ~14 on the P2 brick, where DOF classes
    // alternate along a mesh line).  Vectors of that numbering have npos = 32*nslices entries
    // (padding positions hold 0); the entries of a row keep their order, so every y_i keeps its bits.
    uint16_t* vidx = nullptr;        // instead of pk when the bits do not fit one word: 16-bit dictionary index beside col (6 B / entry)
    uint8_t* pk_nopad = nullptr;     // per slice: 1 = no padding word in the slice (the kernel then skips the selects)
    uint32_t* pkp = nullptr;         // padded entries: (dict index << cbits) | position of the column
    // Segment descriptors of the row-ordered twin ("sell.segments"): in a slice whose 32 rows are one line segment
    // (reorder.cu) entry j of all lanes holds the same dictionary index and columns that are at most two runs of
    // consecutive positions -- (d0, d1, brk, idx): lane l reads position (l < brk ? d0 : d1) + l.  Such a slice is
    // described by ONE 16-byte record per entry instead of 32 words: scd[(slice_ptr[s] >> 5) + j], sc[s] = 1.
    uint4* scd = nullptr;            // padded / 32 records (only those of flagged slices are meaningful)
    uint8_t* sc = nullptr;           // per slice: 1 = described by scd
    size_t sc_slices = 0, sc_entries = 0;    // flagged slices and the stored entries they cover
    int32_t* colp = nullptr;         // row-ordered twin of col for operators WITHOUT a packed stream (12-byte stream): position of the column, -1 = padding
    size_t npos = 0;
};

struct CsrDev {
    size_t rows = 0, cols = 0, nnz = 0;
    int64_t* rowptr = nullptr;
    int32_t* col = nullptr;
    double*  val = nullptr;
};

}  // namespace db200

struct drops_b200_matrix_s {
    db200::CsrDev csr;               // kept: value updates, download, csr kernels, transpose build
    db200::SellOp sell;              // y = A x
    bool has_t = false;
    db200::CsrDev csr_t;             // lazily built A^T (entries of a row ascending in source row)
    db200::SellOp sell_t;
    double* diag = nullptr;          // a_ii (first stored entry with col == row), rows entries
    int* diag_missing = nullptr;     // device flag
    bool diag_ok = false;
    bool square = false;
    bool block_ordered = false;      // sell was built over the block order of reorder.cu
    size_t line_segments = 0;        // complete 32-row line segments at the head of that order ("sell.line_order")
    uint64_t version = 0;
    size_t device_bytes = 0;
    // solver workspace cached with the matrix: vectors of max(rows, cols) doubles, grown on demand
    std::vector<double*> work;
    double* bridge[2] = {nullptr, nullptr};   // device copies of host x / b (host-buffer entry points)
    std::vector<double*> ro_work;    // vectors in the row-ordered numbering (sell.npos doubles)
    uint64_t ro_diag_epoch = 0;      // values epoch ro_work[0] (the permuted diagonal) was gathered at
    uint64_t ro_diag_idx_epoch = 0;  // values epoch the byte index of that diagonal was (tried to be) built at
    // the permuted diagonal as one byte per position + its distinct values, when there are at most kDiagDict
    // of them (constant-coefficient operators): the PCG vector kernels read 1 instead of 8 bytes of d ("pcg.diag_index")
    uint8_t* ro_diag_idx = nullptr;
    int ro_diag_count = 0;
    double ro_diag_values[32] = {};
    void* d_scalars = nullptr;       // db200::SolverScalars
    double* d_hist = nullptr;
    int hist_alloc = 0;
    double* d_small = nullptr;       // small device arrays (Hessenberg column, y, pointers)
    size_t small_alloc = 0;
    // CUDA graph of one chunk of PCG iterations (launch-bound systems: cdrdrops' 250 k unknowns)
    cudaGraphExec_t pcg_graph = nullptr;
    struct { const void* x; const void* b; int chunk; int cg; unsigned long long epoch; uint64_t values; } pcg_graph_key = {nullptr, nullptr, 0, 0, 0, 0};
    long long pcg_graph_launches = 0;  // kernels inside the captured chunk
    // ChebyshevPcCL [0] / ChebyshevBBTPcCL [1] (num/precond.h:1463-1567): constructor arguments,
    // lambda_max cached like SetDiag does (keyed on the values epoch instead of Version())
    double cheb_scale[2] = {1.1, 1.0};
    int cheb_degree[2] = {3, 3};
    double cheb_ratio[2] = {30.0, 30.0};
    double cheb_lambda[2] = {0.0, 0.0};
    double cheb_lambda_scale[2] = {0.0, 0.0};
    uint64_t cheb_lambda_epoch[2] = {0, 0};
    double* diag_pc = nullptr;       // DiagPcCL's vector D (num/precond.h:618-642), rows entries, set by the caller
    double* bbt_diag = nullptr;      // diag(A A^T), BBTDiag num/precond.h:56-75
    uint64_t bbt_epoch = 0;
    uint64_t values_epoch = 1;       // bumped whenever the stored values change
};

namespace db200 {
// i-th cached work vector of the matrix (allocated on first use); nullptr + error on failure.
double* work_vec(drops_b200_matrix_s* M, int i);
void free_sell(SellOp& s);
void free_csr(CsrDev& c);
int build_transpose(drops_b200_matrix_s* M);
int build_sell(const CsrDev& A, SellOp& S, size_t* bytes, const int32_t* rowmap, size_t nsub,
               bool rowmap_is_permutation = false, bool type_sort = false);
// block order of the rows of a square operator from its graph (reorder.cu); *d_order = nullptr: keep the caller's order
// n_segments (optional): complete 32-row line segments at the head of the order ("sell.line_order")
int compute_block_order(const CsrDev& A, int32_t** d_order, size_t* n_segments = nullptr);
// a vector of <= max_values distinct bit patterns as one byte per entry + the sorted value list (matrix.cu)
int build_value_index8(const double* d_val, size_t n, int max_values, uint8_t** d_idx, double* host_values, int* count);
// one pass over the values: will this operator get the one-word packed stream (and with it the row-ordered twin)?
int will_pack_one_word(const double* d_val, size_t nnz, size_t cols_and_positions, bool* yes);
// (re)builds S.dict / S.pk from S.col / S.val; leaves them null when the matrix does not qualify
int build_value_dictionary(SellOp& S, size_t* bytes);
// row-ordered twin of S.col (12-byte stream) for square operators over all their rows that have no packed twin
int build_row_ordered_columns(SellOp& S, size_t* bytes);
// row-ordered twin of S.pk under a caller-supplied numbering iperm[column] = position (< npos)
int pack_reorder(SellOp& S, const int32_t* iperm, size_t npos);
// frees col / val of an operator that holds the one-word packed stream ("sell.release_general", large operators)
void release_general(SellOp& S, size_t* bytes);
// segment descriptors of the row-ordered twin (SellOp::scd) where most slices are line segments
int build_segment_descriptors(SellOp& S, size_t* bytes);
}  // namespace db200

// spmv_pk_impl.cuh -- kernels and launch tables of the dictionary-packed SpMV and of the segment kernel (spmv_pk.cu).
// EXACT = reference arithmetic (product rounded before the add); the fused-multiply-add instantiations of round 1
// ("spmv.exact" = 0) bought nothing on kernels that are nowhere near FP64-bound and are gone.
#pragma once
#include "spmv_impl.cuh"

#include <algorithm>
#include <cstring>

namespace db200 {

constexpr int kPkDictSmem = 512;

struct PkView {
    const int64_t* slice_ptr;
    const int32_t* col;         // six-byte form: int32 column (-1 = padding) ...
    const uint16_t* vidx;       // ... + 16-bit dictionary index (slot ndict = 0.0 for padding)
    const uint32_t* pk;
    const double* dict;
    const int32_t* perm;
    long long nslices;
    int ndict;
    int cbits;
    const uint8_t* nopad;       // per slice: 1 = no padding word (nullptr: unknown, keep the selects)
    int prefetch;               // k_spmv_sell_sci: ask L2 for the next slice's index stream ("spmv.sc_prefetch")
    const uint4* scd;           // segment descriptors of the row-ordered twin (SellOp::scd), nullptr: none
    const uint8_t* sc;          // per slice: 1 = run from scd
    int cdict_n;                // > 0: cdict holds the dictionary (values + padding slot)
    double cdict[kDictConst];   // CD kernels: the dictionary inside the parameter block (constant bank)
};

// a_ij from its dictionary index: shared memory (SMEM), the kernel's parameter block (CD: an indexed
// constant load, which does not pass through the L1 data pipe the x gathers saturate), or global memory
template <bool SMEM, bool CD>
__device__ __forceinline__ double pk_value(const PkView& A, uint32_t idx)
{
    extern __shared__ double sdict[];
    if (CD) return A.cdict[idx];
    return SMEM ? sdict[idx] : __ldg(A.dict + idx);
}

// Padding is the word (ndict << cbits): column 0 and the extra dictionary slot ndict (0.0), so every
// load of the unrolled body is unconditional and straight-line; the padded product is computed and
// then dropped by a select (adding it could turn a -0.0 sum into +0.0, or poison it when x_0 is
// not finite), which keeps the row sum bit-identical.  The matrix stream is loaded ld.global.cs
// (evict-first), x through ld.global.nc.
// What was measured on C2 and NOT kept (profiles/r01_notes.md, r01_probe_*.json; the code is in the history
// of round 1): L2 prefetch of the next slice, a software-pipelined batch loop, a TMA-staged ring, a shared-memory
// window of x, L1::no_allocate stream loads, a masked tail batch, the dot accumulator in registers or in a tail
// pass -- none of them beat this form.
// One batch: UB words in flight, then their UB gathers, then the UB products added left to right.
template <bool EXACT, int UB, bool SMEM, bool NOPAD, bool CD>
__device__ __forceinline__ double pk_batch(const PkView& A, const uint32_t* __restrict__ pp, int j, double sum, uint32_t cmask,
                                           uint32_t wpad, int cbits, const double* __restrict__ x)
{
    uint32_t w[UB];
    double xv[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) w[u] = __ldcs(pp + (size_t)(j + u) * 32);
#pragma unroll
    for (int u = 0; u < UB; ++u) xv[u] = __ldg(x + (w[u] & cmask));
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const double v = pk_value<SMEM, CD>(A, w[u] >> cbits);
        const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
        sum = (NOPAD || w[u] < wpad) ? t : sum;
    }
    return sum;
}

// TAIL: what the batches of U leave over runs as batches of 8, 4, 2 (a row of 27 entries is 15 + 8 + 4: three round
// trips to memory) instead of one entry at a time.
template <bool EXACT, int U, bool SMEM, bool NOPAD, bool CD, bool TAIL>
__device__ __forceinline__ double pk_row_sum(const PkView& A, int64_t base, int width, int lane,
                                             const double* __restrict__ x)
{
    const uint32_t* pp = A.pk + base + lane;
    const int cbits = A.cbits;
    const uint32_t cmask = (1u << cbits) - 1u;
    const uint32_t wpad = (uint32_t)A.ndict << cbits;
    double sum = 0.0;
    int j = 0;
    for (; j + U <= width; j += U) sum = pk_batch<EXACT, U, SMEM, NOPAD, CD>(A, pp, j, sum, cmask, wpad, cbits, x);
    if (TAIL) {
        if (U > 8 && j + 8 <= width) { sum = pk_batch<EXACT, 8, SMEM, NOPAD, CD>(A, pp, j, sum, cmask, wpad, cbits, x); j += 8; }
        if (U > 4 && j + 4 <= width) { sum = pk_batch<EXACT, 4, SMEM, NOPAD, CD>(A, pp, j, sum, cmask, wpad, cbits, x); j += 4; }
        if (U > 2 && j + 2 <= width) { sum = pk_batch<EXACT, 2, SMEM, NOPAD, CD>(A, pp, j, sum, cmask, wpad, cbits, x); j += 2; }
    }
    for (; j < width; ++j) sum = pk_batch<EXACT, 1, SMEM, NOPAD, CD>(A, pp, j, sum, cmask, wpad, cbits, x);
    return sum;
}

// Six-byte form (V6): matrices whose column bits + index bits exceed one word keep the int32 column
// of the general stream and replace the fp64 value by a 16-bit dictionary index.
template <bool EXACT, int U, bool SMEM>
__device__ __forceinline__ double pk6_row_sum(const PkView& A, int64_t base, int width, int lane,
                                              const double* __restrict__ x)
{
    extern __shared__ double sdict[];
    const int32_t* cp = A.col + base + lane;
    const uint16_t* ip = A.vidx + base + lane;
    double sum = 0.0;
    int j = 0;
    for (; j + U <= width; j += U) {
        int c[U];
        unsigned vi[U];
        double xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { c[u] = __ldcs(cp + (size_t)(j + u) * 32); vi[u] = __ldcs(ip + (size_t)(j + u) * 32); }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = c[u] >= 0 ? __ldg(x + c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double v = SMEM ? sdict[vi[u]] : __ldg(A.dict + vi[u]);
            const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
            sum = c[u] >= 0 ? t : sum;
        }
    }
    for (; j < width; ++j) {
        const int c = __ldcs(cp + (size_t)j * 32);
        const unsigned vi = __ldcs(ip + (size_t)j * 32);
        const double xv = c >= 0 ? __ldg(x + c) : 0.0;
        const double v = SMEM ? sdict[vi] : __ldg(A.dict + vi);
        const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv)) : fma(v, xv, sum);
        sum = c >= 0 ? t : sum;
    }
    return sum;
}

// DOT as in k_spmv_sell: 0 none, 1 finished in this kernel, 2 per-warp partials + finishing kernel.
// The fused dot fetches its partner AFTER the row sum (an L1 hit: the diagonal entry just gathered it) and
// parks the accumulator in shared memory between slices, so the row sum runs with the registers of the
// plain product (same per-thread order of the same products as any other form: bit-identical results).
// LEAN (the U15 fused-dot form the PCG loop runs): 1 = batched tail; 2 = batched tail, and every warp asks L2 for
// the stream of its NEXT slice (prefetch.global.L2, one 128-byte line per lane and round) before it sums the current one;
// the extent of that slice was read one iteration earlier, so the prefetch waits for nothing.
template <int EPI, bool EXACT, int DOT, int U, int MAXT, int MINB, bool SMEM, bool V6, bool CD, int LEAN = 0>
__global__ void __launch_bounds__(MAXT, MINB)
k_spmv_sell_pk(const __grid_constant__ PkView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e)
{
    if (e.done != nullptr && *e.done != 0) return;
    extern __shared__ double sdict[];
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    __shared__ dd acc_smem[DOT != 0 ? MAXT : 1];
    if (DOT != 0) acc_smem[threadIdx.x] = dd_zero();     // the thread's own slot: no barrier needed
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int s_begin = (int)(A.nslices * (long long)blockIdx.x / gridDim.x);
    const int s_end   = (int)(A.nslices * (long long)(blockIdx.x + 1) / gridDim.x);
    if (SMEM) {
        for (int i = threadIdx.x; i <= A.ndict; i += blockDim.x) sdict[i] = A.dict[i];
        __syncthreads();
    }
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;

    // LEAN: the extents of this warp's slices run two iterations ahead of their use -- (base, width) of slice s were read
    // while s - 2 nwarp was summed, those of s + nwarp while s - nwarp was: neither the row sum nor the prefetch waits
    int64_t base = 0, nbase = 0;
    int width = 0, nwidth = 0;
    if (LEAN == 2) {
        if (s_begin + warp < s_end) {
            base = A.slice_ptr[s_begin + warp];
            width = (int)((A.slice_ptr[s_begin + warp + 1] - base) >> 5);
        }
        if (s_begin + warp + nwarp < s_end) {
            nbase = A.slice_ptr[s_begin + warp + nwarp];
            nwidth = (int)((A.slice_ptr[s_begin + warp + nwarp + 1] - nbase) >> 5);
        }
    }
    for (int s = s_begin + warp; s < s_end; s += nwarp) {
        int64_t n2base = 0;
        int n2width = 0;
        if (LEAN == 2) {
            for (int l = lane; l < nwidth; l += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(A.pk + nbase + (size_t)l * 32));
            const int s2 = s + 2 * nwarp;
            if (s2 < s_end) {
                n2base = A.slice_ptr[s2];
                n2width = (int)((A.slice_ptr[s2 + 1] - n2base) >> 5);
            }
        } else {
            base = A.slice_ptr[s];
            width = (int)((A.slice_ptr[s + 1] - base) >> 5);
        }
        // row-ordered twin (perm == nullptr): position s*32+lane IS the row of y, aux and the dot partner
        const int row = A.perm != nullptr ? A.perm[(size_t)s * 32 + lane] : s * 32 + lane;
        double av = 0.0;
        if (row >= 0 && EPI != EPI_NONE) av = __ldg(e.aux + row);
        // 99.8 % of the slices of a sorted operator hold no padding word: their row sum needs no selects
        const bool nopad = !V6 && A.nopad != nullptr && A.nopad[s] != 0;      // warp-uniform
        const double sum = V6 ? pk6_row_sum<EXACT, U, SMEM>(A, base, width, lane, x)
                         : nopad ? pk_row_sum<EXACT, U, SMEM, true, CD, (LEAN >= 1)>(A, base, width, lane, x)
                                 : pk_row_sum<EXACT, U, SMEM, false, CD, (LEAN >= 1)>(A, base, width, lane, x);
        if (row >= 0) {
            const double out = apply_epilogue<EPI>(sum, av, e, beta);
            y[row] = out;
            if (DOT) acc_smem[threadIdx.x] = dd_add_prod(acc_smem[threadIdx.x], out, __ldg(e.dotv + row));
        }
        if (LEAN == 2) { base = nbase; width = nwidth; nbase = n2base; nwidth = n2width; }
    }
    dd acc[1] = {DOT != 0 ? acc_smem[threadIdx.x] : dd_zero()};
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


// ------------------------------------------------------------------------------------------
// Segment kernel ("spmv.segments"): the row-ordered twin of an operator whose slices are mostly line segments.
// A flagged slice is described by one 16-byte record per entry-row (SellOp::scd) instead of 32 words: the warp
// copies the records of its slice into shared memory with cp.async (the copy of the NEXT slice travels while the
// epilogue of this one runs), every lane then forms its column as (lane < brk ? d0 : d1) + lane -- 32 consecutive
// positions, at most two runs -- and all lanes share the dictionary index.  Same products, same left-to-right order
// as the packed row sum: bit-identical.  Slices that are not flagged (chain tails, boundary rows) run the packed row
// sum in the same kernel.  The stream of a flagged slice is 0.5 instead of 4 bytes per stored entry.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sc_issue(const PkView& A, uint32_t sd_addr, int64_t base, int width, int lane)
{
    const uint4* src = A.scd + (base >> 5);
    for (int j = lane; j < width; j += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sd_addr + (uint32_t)j * 16u), "l"(src + j) : "memory");
}

template <bool EXACT, int UB>
__device__ __forceinline__ double sc_batch(const PkView& A, const uint4* __restrict__ sd, int j, double sum, int lane,
                                           const double* __restrict__ x)
{
    double xv[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const uint4 d = sd[j + u];                                       // one broadcast 16-byte read
        const int c = (lane < (int)d.z ? (int)d.x : (int)d.y) + lane;
        xv[u] = __ldg(x + c);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        // the index is read again: a register per entry is dearer than an LDS (packing four indices into a register was
        // measured too: more spills, slower)
        const double v = A.cdict[sd[j + u].w];
        sum = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
    }
    return sum;
}

template <bool EXACT, int U>
__device__ __forceinline__ double sc_row_sum(const PkView& A, const uint4* __restrict__ sd, int width, int lane,
                                             const double* __restrict__ x)
{
    double sum = 0.0;
    int j = 0;
    for (; j + U <= width; j += U) sum = sc_batch<EXACT, U>(A, sd, j, sum, lane, x);
    if (U > 8 && j + 8 <= width) { sum = sc_batch<EXACT, 8>(A, sd, j, sum, lane, x); j += 8; }
    if (U > 4 && j + 4 <= width) { sum = sc_batch<EXACT, 4>(A, sd, j, sum, lane, x); j += 4; }
    if (U > 2 && j + 2 <= width) { sum = sc_batch<EXACT, 2>(A, sd, j, sum, lane, x); j += 2; }
    for (; j < width; ++j) sum = sc_batch<EXACT, 1>(A, sd, j, sum, lane, x);
    return sum;
}

template <int EPI, bool EXACT, int DOT, int U, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_spmv_sell_sc(const __grid_constant__ PkView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e)
{
    if (e.done != nullptr && *e.done != 0) return;
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    __shared__ dd acc_smem[DOT != 0 ? MAXT : 1];
    __shared__ uint4 sdesc[(MAXT / 32) * kScMaxWidth];
    if (DOT != 0) acc_smem[threadIdx.x] = dd_zero();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int s_begin = (int)(A.nslices * (long long)blockIdx.x / gridDim.x);
    const int s_end   = (int)(A.nslices * (long long)(blockIdx.x + 1) / gridDim.x);
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    const uint4* sd = sdesc + warp * kScMaxWidth;
    const uint32_t sd_addr = (uint32_t)__cvta_generic_to_shared(sd);

    // extents and flag of the warp's next slice are read while the current one is summed
    int64_t base = 0;
    int width = 0;
    bool seg = false;
    if (s_begin + warp < s_end) {
        const int s0 = s_begin + warp;
        base = A.slice_ptr[s0];
        width = (int)((A.slice_ptr[s0 + 1] - base) >> 5);
        seg = A.sc[s0] != 0;
        if (seg) sc_issue(A, sd_addr, base, width, lane);
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
        double av = 0.0;
        if (EPI != EPI_NONE) av = __ldg(e.aux + row);
        double sum;
        if (seg) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            sum = sc_row_sum<EXACT, U>(A, sd, width, lane, x);
            __syncwarp();                                     // every lane is done with the records
        } else {
            sum = pk_row_sum<EXACT, U, false, false, true, true>(A, base, width, lane, x);
        }
        if (nseg) sc_issue(A, sd_addr, nbase, nwidth, lane);  // travels during the epilogue
        else if (A.prefetch != 0)                             // a slice that keeps its packed words: L2 is asked for them
            for (int l = lane; l < nwidth; l += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(A.pk + nbase + (size_t)l * 32));
        const double out = apply_epilogue<EPI>(sum, av, e, beta);
        y[row] = out;
        if (DOT) acc_smem[threadIdx.x] = dd_add_prod(acc_smem[threadIdx.x], out, __ldg(e.dotv + row));
        base = nbase; width = nwidth; seg = nseg;
    }
    dd acc[1] = {DOT != 0 ? acc_smem[threadIdx.x] : dd_zero()};
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
static void launch_sc_d(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    if (threads > MAXT) threads = MAXT;
    if (a.dotv == nullptr) {
        k_spmv_sell_sc<EPI, EXACT, 0, U, MAXT, MINB><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
        return;
    }
    const long long nwarps = (long long)grid * (threads >> 5);
    const double mode = opt("spmv.dotmode", 0.0);
    const bool two_kernels = mode == 2.0 || (mode == 0.0 && grid > (int)kReduceFlat);
    if (two_kernels && nwarps <= kPartialStride) {
        k_spmv_sell_sc<EPI, EXACT, 2, U, MAXT, MINB><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
        sum_warp_partials(nwarps, e, st);
    } else {
        k_spmv_sell_sc<EPI, EXACT, 1, U, MAXT, MINB><<<grid, threads, 0, st>>>(v, a.x, a.y, e);
    }
}


// ------------------------------------------------------------------------------------------
// k_spmv_sell_sci: segment records for the COLUMNS, a 16-bit dictionary index per stored entry for the VALUES -- 2.5 bytes
// per entry.  This is the form of a matrix DROPS assembled itself on a structured mesh: the columns are translates of
// each other, the values are equal only up to the rounding of the assembly (757 distinct bit patterns on a 40^3 P2 brick
// where exact arithmetic has 21), so lanes do not share an index, and at C2 size column bits + index bits no longer fit
// one word.  Row-ordered twin of the six-byte form: PkView::col = positions (SellOp::colp), vidx, scd / sc.  Dictionary in
// shared memory (SMEM, <= kSciDictSmem values) or through L1.  Unflagged slices read the int32 position of every entry.
// Same products in the same order as the six-byte kernel: bit-identical.
// ------------------------------------------------------------------------------------------
constexpr int kSciDictSmem = 3072;         // (24 KB of dynamic shared memory beside the 15 KB of records)

template <bool EXACT, int UB, bool SMEM>
__device__ __forceinline__ double sci_batch(const PkView& A, const uint16_t* __restrict__ ip, const uint4* __restrict__ sd, int j,
                                            double sum, int lane, const double* __restrict__ x, const double* __restrict__ sdict)
{
    unsigned vi[UB];
    double xv[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) vi[u] = __ldcs(ip + (size_t)(j + u) * 32);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const uint4 d = sd[j + u];
        const int c = (lane < (int)d.z ? (int)d.x : (int)d.y) + lane;
        xv[u] = __ldg(x + c);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
        const double v = SMEM ? sdict[vi[u]] : __ldg(A.dict + vi[u]);
        sum = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
    }
    return sum;
}

template <int EPI, bool EXACT, int DOT, int U, int MAXT, int MINB, bool SMEM>
__global__ void __launch_bounds__(MAXT, MINB)
k_spmv_sell_sci(const __grid_constant__ PkView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e)
{
    if (e.done != nullptr && *e.done != 0) return;
    extern __shared__ double sdict[];                      // SMEM: ndict + 1 values
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    __shared__ uint4 sdesc[(MAXT / 32) * kScMaxWidth];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int s_begin = (int)(A.nslices * (long long)blockIdx.x / gridDim.x);
    const int s_end   = (int)(A.nslices * (long long)(blockIdx.x + 1) / gridDim.x);
    if (SMEM) {
        for (int i = threadIdx.x; i <= A.ndict; i += blockDim.x) sdict[i] = A.dict[i];
        __syncthreads();
    }
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
        if (seg) sc_issue(A, sd_addr, base, width, lane);
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
            const uint16_t* ip = A.vidx + base + lane;
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            int j = 0;
            for (; j + U <= width; j += U) sum = sci_batch<EXACT, U, SMEM>(A, ip, sd, j, sum, lane, x, sdict);
            if (U > 8 && j + 8 <= width) { sum = sci_batch<EXACT, 8, SMEM>(A, ip, sd, j, sum, lane, x, sdict); j += 8; }
            if (U > 4 && j + 4 <= width) { sum = sci_batch<EXACT, 4, SMEM>(A, ip, sd, j, sum, lane, x, sdict); j += 4; }
            if (U > 2 && j + 2 <= width) { sum = sci_batch<EXACT, 2, SMEM>(A, ip, sd, j, sum, lane, x, sdict); j += 2; }
            for (; j < width; ++j) sum = sci_batch<EXACT, 1, SMEM>(A, ip, sd, j, sum, lane, x, sdict);
            __syncwarp();
        } else {
            sum = pk6_row_sum<EXACT, 8, SMEM>(A, base, width, lane, x);      // int32 position (-1 = padding) + index per entry
        }
        if (nseg) {
            sc_issue(A, sd_addr, nbase, nwidth, lane);
            // ... and L2 is asked for the next slice's index stream (64 bytes per entry-row): its batches then wait for L2,
            // not for HBM
            if (A.prefetch != 0)
                for (int l = lane; 2 * l < nwidth; l += 32)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(A.vidx + nbase + (size_t)l * 64));
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

template <int EPI, bool EXACT, int U, int MAXT, int MINB, bool SMEM>
static void launch_sci_d(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    if (threads > MAXT) threads = MAXT;
    const size_t sh = SMEM ? ((size_t)v.ndict + 1) * sizeof(double) : 0;
    if (a.dotv == nullptr) {
        k_spmv_sell_sci<EPI, EXACT, 0, U, MAXT, MINB, SMEM><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
        return;
    }
    const long long nwarps = (long long)grid * (threads >> 5);
    const double mode = opt("spmv.dotmode", 0.0);
    const bool two_kernels = mode == 2.0 || (mode == 0.0 && grid > (int)kReduceFlat);
    if (two_kernels && nwarps <= kPartialStride) {
        k_spmv_sell_sci<EPI, EXACT, 2, U, MAXT, MINB, SMEM><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
        sum_warp_partials(nwarps, e, st);
    } else {
        k_spmv_sell_sci<EPI, EXACT, 1, U, MAXT, MINB, SMEM><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
    }
}

template <int EPI, bool EXACT, int U, int MAXT, int MINB, bool SMEM, bool V6, bool CD, int LEAN = 0>
static void launch_pk_d(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    if (threads > MAXT) threads = MAXT;
    const size_t sh = SMEM ? ((size_t)v.ndict + 1) * sizeof(double) : 0;
    if (a.dotv == nullptr) {
        k_spmv_sell_pk<EPI, EXACT, 0, U, MAXT, MINB, SMEM, V6, CD, LEAN><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
        return;
    }
    const long long nwarps = (long long)grid * (threads >> 5);
    const double mode = opt("spmv.dotmode", 0.0);
    const bool two_kernels = mode == 2.0 || (mode == 0.0 && grid > (int)kReduceFlat);
    if (two_kernels && nwarps <= kPartialStride) {
        k_spmv_sell_pk<EPI, EXACT, 2, U, MAXT, MINB, SMEM, V6, CD, LEAN><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
        sum_warp_partials(nwarps, e, st);
    } else {
        k_spmv_sell_pk<EPI, EXACT, 1, U, MAXT, MINB, SMEM, V6, CD, LEAN><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
    }
}

// Launch table.  Measured on C2 (profiles/r01_probe_blockorder.json, r01_probe_cdict*.json):
//   * dictionary of <= 47 values (it fits the parameter block, "spmv.pk_cdict" = 1): plain product and
//     epilogue forms U8 x 512-thread CTAs (<= 40 registers, 3 CTAs/SM); product fused with the dot
//     ("spmv.pk_cdict_dot" = 1) U15 x 384-thread CTAs (<= 56 registers, 3 CTAs/SM);
//   * larger dictionaries: shared-memory copy per CTA (<= kPkDictSmem values) or through L1, U8 x 512;
//   * six-byte form: U8 x 512.
template <int EPI, bool EXACT>
static bool launch_pk(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    const bool smem_ok = v.ndict <= kPkDictSmem && opt("spmv.pkdict", 0.0) == 0.0;
    if (v.vidx != nullptr && v.scd != nullptr && v.sc != nullptr && v.perm == nullptr) {   // row-ordered twin of the six-byte form
        // measured on C2 forced into this form (profiles/r02_probe_sci2.json): 8 entries in flight x 512 threads x 3 CTAs per SM
        // 2.09 ms, 12 x 384 x 3: 2.14, the twin without records: 2.22
        if (v.ndict <= kSciDictSmem && opt("spmv.pkdict", 0.0) == 0.0) launch_sci_d<EPI, EXACT, 8, 512, 3, true>(v, a, e, grid, threads, st);
        else launch_sci_d<EPI, EXACT, 8, 512, 3, false>(v, a, e, grid, threads, st);
        return true;
    }
    if (v.vidx != nullptr) {       // six-byte form: dictionary in shared memory or through L1
        if (smem_ok) launch_pk_d<EPI, EXACT, 8, 512, 3, true, true, false>(v, a, e, grid, threads, st);
        else launch_pk_d<EPI, EXACT, 8, 512, 3, false, true, false>(v, a, e, grid, threads, st);
        return true;
    }
    if (smem_ok && v.cdict_n > 0 && opt("spmv.pk_cdict", 1.0) != 0.0) {
        if (v.scd != nullptr && v.sc != nullptr && v.perm == nullptr) {       // row-ordered twin with segment descriptors
            // measured on C2 (profiles/r02_probe_sc*.json): 12 gathers in flight x 512-thread CTAs x 3 per SM (40 registers)
            // 1.22 ms; 15 x 384 x 3: 1.26-1.33; 10 x 384 x 4: 1.26; 19 x 384 x 3: 1.34; 8 x 512 x 4: 1.48
            launch_sc_d<EPI, EXACT, 12, 512, 3>(v, a, e, grid, threads, st);
            return true;
        }
        if (EPI == EPI_NONE && EXACT && a.dotv != nullptr && opt("spmv.pk_cdict_dot", 1.0) != 0.0) {
            const int lean = (int)opt("spmv.pk_lean", 2.0);    // 1: batched tail, 2: + L2 prefetch of the warp's next slice
            if (lean == 2) launch_pk_d<EPI_NONE, true, 15, 384, 3, false, false, true, 2>(v, a, e, grid, threads, st);
            else if (lean == 1) launch_pk_d<EPI_NONE, true, 15, 384, 3, false, false, true, 1>(v, a, e, grid, threads, st);
            else launch_pk_d<EPI_NONE, true, 15, 384, 3, false, false, true>(v, a, e, grid, threads, st);
        } else
            launch_pk_d<EPI, EXACT, 8, 512, 3, false, false, true>(v, a, e, grid, threads, st);
        return true;
    }
    if (smem_ok) launch_pk_d<EPI, EXACT, 8, 512, 3, true, false, false>(v, a, e, grid, threads, st);
    else launch_pk_d<EPI, EXACT, 8, 512, 3, false, false, false>(v, a, e, grid, threads, st);
    return true;
}

template <bool EXACT>
static void launch_pk_epi(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    switch (a.epilogue) {
        case EPI_JACOBI: launch_pk<EPI_JACOBI, EXACT>(v, a, e, grid, threads, st); break;
        case EPI_RESID:  launch_pk<EPI_RESID, EXACT>(v, a, e, grid, threads, st); break;
        case EPI_AXMBY:  launch_pk<EPI_AXMBY, EXACT>(v, a, e, grid, threads, st); break;
        default:         launch_pk<EPI_NONE, EXACT>(v, a, e, grid, threads, st); break;
    }
}

}  // namespace db200

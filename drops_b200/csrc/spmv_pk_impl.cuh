// spmv_pk_impl.cuh -- kernels and launch tables of the dictionary-packed SpMV, shared by the two translation
// units that instantiate them: spmv_pk.cu (exact reference arithmetic: product rounded before the add) and
// spmv_pk_fma.cu ("spmv.exact" = 0: fused multiply-add).  Split so that the two halves compile in parallel.
#pragma once
#include "spmv_impl.cuh"

#include <algorithm>
#include <cstring>

namespace db200 {

constexpr int kPkDictSmem = 512;
constexpr size_t kXwMaxSmem = 100 * 1024;      // x window + dictionary per CTA

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
    int xw_halo;                // XW kernels: slices of x staged either side of the CTA's own rows
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

// matrix-stream load: LD = 0 ld.global.cs (evict-first in L1 and L2), 1 = L1::no_allocate + L2 evict-first
template <int LD>
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p)
{
    if (LD == 0) return __ldcs(p);
    uint32_t v;
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// x window (XW, option "spmv.pk_xwin", off): with row-ordered vectors a CTA's own rows are the positions
// [32 s_begin, 32 s_end) of x, and on a block-ordered operator (reorder.cu) most of the columns of those rows
// are positions of the same block.  The CTA copies that stretch of x (plus `halo` slices either side) into
// shared memory with coalesced loads and gathers from there; only columns outside the window go through L1.
// Measured on C2 inside PCG (profiles/r01_probe_xw.json): bit-identical and SLOWER in every geometry -- 1.68 ms
// without the window, 2.06-2.9 ms with it (the window costs L1 capacity and every gather becomes a predicated
// LDS + LDG pair).  Kept as a measured option.
struct XWin {
    uint32_t off;               // sdict[off + k] = x[base + k], k < len
    uint32_t base, len;
};

template <bool XW>
__device__ __forceinline__ double pk_gather(const double* __restrict__ x, const XWin& xw, uint32_t pos)
{
    if (XW) {
        extern __shared__ double sdict[];
        const uint32_t rel = pos - xw.base;
        return rel < xw.len ? sdict[xw.off + rel] : __ldg(x + pos);
    }
    return __ldg(x + pos);
}

// Padding is the word (ndict << cbits): column 0 and the extra dictionary slot ndict (0.0), so every
// load of the unrolled body is unconditional and straight-line; the padded product is computed and
// then dropped by a select (adding it could turn a -0.0 sum into +0.0, or poison it when x_0 is
// not finite), which keeps the row sum bit-identical.
// TM (masked tail): the width % U entries left after the full batches are ONE more batch whose loads are
// predicated on the (warp-uniform) remainder instead of a loop the compiler cuts into batches of 8, 4, 2, 1 --
// every batch costs a round trip to the matrix stream plus one to x, so a row of 27 entries at U = 14 takes
// 2 round-trip pairs instead of 4.  The masked slots hold the padding word and are dropped by the same select.
// Measured on C2 (profiles/r01_probe_cdict2.json): no gain (U14 1.500 -> 1.499 ms, U15 1.467 -> 1.581) -- the
// product is not bound by the number of round trips per row; variants 100 + n of "spmv.pk_cdict[_dot]".
template <bool EXACT, int U, bool SMEM, int LD, bool NOPAD = false, bool XW = false, bool CD = false, bool TM = false>
__device__ __forceinline__ double pk_row_sum(const PkView& A, int64_t base, int width, int lane,
                                             const double* __restrict__ x, const XWin& xw = XWin())
{
    extern __shared__ double sdict[];      // the CTA's copy of the dictionary (SMEM)
    const uint32_t* pp = A.pk + base + lane;
    const int cbits = A.cbits;
    const uint32_t cmask = (1u << cbits) - 1u;
    const uint32_t wpad = (uint32_t)A.ndict << cbits;
    double sum = 0.0;
    int j = 0;
    for (; j + U <= width; j += U) {
        uint32_t w[U];
        double xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = ld_stream<LD>(pp + (size_t)(j + u) * 32);
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = pk_gather<XW>(x, xw, w[u] & cmask);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double v = pk_value<SMEM, CD>(A, w[u] >> cbits);
            const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
            sum = (NOPAD || w[u] < wpad) ? t : sum;
        }
    }
    if constexpr (TM) {
        const int rem = width - j;                      // < U, the same for all lanes of the slice
        if (rem > 0) {
            uint32_t w[U];
            double xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) w[u] = u < rem ? ld_stream<LD>(pp + (size_t)(j + u) * 32) : wpad;
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = u < rem ? pk_gather<XW>(x, xw, w[u] & cmask) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double v = pk_value<SMEM, CD>(A, w[u] >> cbits);
                const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
                sum = w[u] < wpad ? t : sum;
            }
        }
    } else {
        for (; j < width; ++j) {
            const uint32_t w = ld_stream<LD>(pp + (size_t)j * 32);
            const double xv = pk_gather<XW>(x, xw, w & cmask);
            const double v = pk_value<SMEM, CD>(A, w >> cbits);
            const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv)) : fma(v, xv, sum);
            sum = (NOPAD || w < wpad) ? t : sum;
        }
    }
    return sum;
}

// DOT as in k_spmv_sell: 0 none, 1 finished in this kernel, 2 per-warp partials + finishing kernel.
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

// Software-pipelined form: the words of batch k+1 are requested before the gathers of batch k are
// consumed, so a warp always has U matrix loads AND U gathers in flight (the plain form alternates).
template <bool EXACT, int U, bool SMEM>
__device__ __forceinline__ double pk_row_sum_pipe(const PkView& A, int64_t base, int width, int lane,
                                                  const double* __restrict__ x)
{
    extern __shared__ double sdict[];
    const uint32_t* pp = A.pk + base + lane;
    const int cbits = A.cbits;
    const uint32_t cmask = (1u << cbits) - 1u;
    const uint32_t wpad = (uint32_t)A.ndict << cbits;
    double sum = 0.0;
    int j = 0;
    uint32_t w[U];
    if (U <= width) {
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = __ldcs(pp + (size_t)u * 32);
    }
    for (; j + U <= width; j += U) {
        uint32_t wn[U];
        double xv[U];
        const bool more = j + 2 * U <= width;
        if (more) {
#pragma unroll
            for (int u = 0; u < U; ++u) wn[u] = __ldcs(pp + (size_t)(j + U + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = __ldg(x + (w[u] & cmask));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double v = SMEM ? sdict[w[u] >> cbits] : __ldg(A.dict + (w[u] >> cbits));
            const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv[u])) : fma(v, xv[u], sum);
            sum = w[u] < wpad ? t : sum;
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < U; ++u) w[u] = wn[u];
        }
    }
    for (; j < width; ++j) {
        const uint32_t ww = __ldcs(pp + (size_t)j * 32);
        const double xv = __ldg(x + (ww & cmask));
        const double v = SMEM ? sdict[ww >> cbits] : __ldg(A.dict + (ww >> cbits));
        const double t = EXACT ? __dadd_rn(sum, __dmul_rn(v, xv)) : fma(v, xv, sum);
        sum = ww < wpad ? t : sum;
    }
    return sum;
}

// DF, the form of the fused dot (same per-thread order of the same products: bit-identical results):
//   0 = in the slice loop, partner fetched at slice start, accumulator in registers -- at <= 40
//       registers the compiler can then no longer keep the U loads of the row sum in flight
//       (2.2 instead of 1.5 ms at C2);
//   1 = a second pass over the warp's own slices (y_i re-read from L2, the partner from L1);
//   2 = in the slice loop, but the partner is fetched AFTER the row sum (an L1 hit: the diagonal
//       entry just gathered it) and the accumulator is parked in shared memory between slices, so
//       the row sum runs with the registers of the plain product.
template <int EPI, bool EXACT, int DOT, int U, int MAXT, int MINB, bool SMEM, int DF, bool V6 = false, bool XW = false, bool CD = false, bool TM = false>
__global__ void __launch_bounds__(MAXT, MINB)
k_spmv_sell_pk(const __grid_constant__ PkView A, const double* __restrict__ x, double* __restrict__ y, EpiArgs e)
{
    if (e.done != nullptr && *e.done != 0) return;
    extern __shared__ double sdict[];
    __shared__ dd red_smem[DOT == 1 ? 32 : 1];
    __shared__ dd acc_smem[(DOT != 0 && DF == 2) ? MAXT : 1];
    if (DOT != 0 && DF == 2) acc_smem[threadIdx.x] = dd_zero();     // the thread's own slot: no barrier needed
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long s_begin = A.nslices * (long long)blockIdx.x / gridDim.x;
    const long long s_end   = A.nslices * (long long)(blockIdx.x + 1) / gridDim.x;
    XWin xw{0u, 0u, 0u};
    if (SMEM) {
        for (int i = threadIdx.x; i <= A.ndict; i += blockDim.x) sdict[i] = A.dict[i];
    }
    if (XW) {
        // the window sits behind the dictionary (16-byte aligned); positions are multiples of 32 doubles
        double* swin = sdict + ((A.ndict + 2) & ~1);
        const long long lo = s_begin - A.xw_halo > 0 ? s_begin - A.xw_halo : 0;
        const long long hi = s_end + A.xw_halo < A.nslices ? s_end + A.xw_halo : A.nslices;
        xw.off = (uint32_t)((A.ndict + 2) & ~1); xw.base = (uint32_t)(lo * 32); xw.len = (uint32_t)((hi - lo) * 32);
        const double2* src = reinterpret_cast<const double2*>(x + lo * 32);
        double2* dst = reinterpret_cast<double2*>(swin);
        for (uint32_t i = threadIdx.x; i < xw.len / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (SMEM || XW) __syncthreads();
    const double beta = (EPI == EPI_AXMBY) ? (e.beta != nullptr ? *e.beta : e.beta_val) : 0.0;
    dd acc[1] = {dd_zero()};
    // PCG: x += alpha p of the previous iteration rides on this product (DRAM has the headroom, the
    // kernel is bound by the L2 -> L1 path), 3 vector passes less in the direction kernel
    const bool xupd = DOT != 0 && e.xu != nullptr && *e.xu_applied != *e.xu_iter;
    const double xalpha = xupd ? *e.xu_alpha : 0.0;

    for (long long s = s_begin + warp; s < s_end; s += nwarp) {
        const int64_t base = A.slice_ptr[s];
        const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
        // row-ordered twin (perm == nullptr): position s*32+lane IS the row of y, aux and the dot partner
        const int row = A.perm != nullptr ? A.perm[s * 32 + lane] : (int)(s * 32 + lane);
        double dv = 0.0, av = 0.0;
        if (row >= 0) {
            if (DOT && DF == 0) dv = __ldg(e.dotv + row);
            if (EPI != EPI_NONE) av = __ldg(e.aux + row);
        }
        if (e.prefetch > 0 && s + (long long)e.prefetch * nwarp < A.nslices) {
            // request the words of a slice this warp (or the next CTA) reaches soon into L2: the stream
            // is perfectly predictable, and the kernel waits on latency, not on DRAM bytes
            const long long sp = s + (long long)e.prefetch * nwarp;
            const int64_t nb = A.slice_ptr[sp];
            const int nlines = (int)((A.slice_ptr[sp + 1] - nb) >> 5);
            for (int l = lane; l < nlines; l += 32) prefetch_l2(A.pk + nb + (int64_t)l * 32);
        }
        // U < 0: software-pipelined form; U >= 100: U - 100 words in flight, stream loaded with L1::no_allocate
        constexpr int UU = U < 0 ? -U : (U >= 100 ? U - 100 : U);
        // 99.8 % of the slices of a sorted operator hold no padding word: their row sum needs no selects
        const bool nopad = !V6 && U > 0 && A.nopad != nullptr && A.nopad[s] != 0;      // warp-uniform
        const double sum = V6 ? pk6_row_sum<EXACT, UU, SMEM>(A, base, width, lane, x)
                         : U <= 0 ? pk_row_sum_pipe<EXACT, UU, SMEM>(A, base, width, lane, x)
                         : nopad ? pk_row_sum<EXACT, UU, SMEM, (U >= 100 ? 1 : 0), true, XW, CD, TM>(A, base, width, lane, x, xw)
                                 : pk_row_sum<EXACT, UU, SMEM, (U >= 100 ? 1 : 0), false, XW, CD, TM>(A, base, width, lane, x, xw);
        if (row >= 0) {
            const double out = apply_epilogue<EPI>(sum, av, e, beta);
            y[row] = out;
            if (DOT && DF == 0) acc[0] = dd_add_prod(acc[0], out, dv);
            if (DOT && DF == 2) acc_smem[threadIdx.x] = dd_add_prod(acc_smem[threadIdx.x], out, __ldg(e.dotv + row));
            if (xupd) e.xu[row] = __dadd_rn(e.xu[row], __dmul_rn(xalpha, __ldcs(e.xu_p + row)));
        }
    }
    if (DOT && DF == 2) acc[0] = acc_smem[threadIdx.x];
    if (DOT && DF == 1) {
        for (long long s = s_begin + warp; s < s_end; s += nwarp) {
            const int row = A.perm != nullptr ? A.perm[s * 32 + lane] : (int)(s * 32 + lane);
            if (row >= 0) acc[0] = dd_add_prod(acc[0], __ldcg(y + row), __ldg(e.dotv + row));
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

// shared memory of an XW launch: dictionary + the widest window a CTA of this grid can own
static size_t xw_smem_bytes(const PkView& v, int grid)
{
    const long long own = (v.nslices + grid - 1) / grid;
    return ((size_t)((v.ndict + 2) & ~1) + (size_t)(own + 2 * v.xw_halo) * 32) * sizeof(double);
}

template <typename K>
static void allow_smem(K kernel, size_t sh)
{
    if (sh > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
}

template <int EPI, bool EXACT, int U, int MAXT, int MINB, bool SMEM, int DF, bool V6 = false, bool XW = false, bool CD = false, bool TM = false>
static void launch_pk_d(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    if (threads > MAXT) threads = MAXT;
    const size_t sh = XW ? xw_smem_bytes(v, grid) : SMEM ? ((size_t)v.ndict + 1) * sizeof(double) : 0;
    if (a.dotv == nullptr) {
        if (XW) allow_smem(k_spmv_sell_pk<EPI, EXACT, 0, U, MAXT, MINB, SMEM, 0, V6, XW, CD, TM>, sh);
        k_spmv_sell_pk<EPI, EXACT, 0, U, MAXT, MINB, SMEM, 0, V6, XW, CD, TM><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
        return;
    }
    const long long nwarps = (long long)grid * (threads >> 5);
    const double mode = opt("spmv.dotmode", 0.0);
    const bool two_kernels = mode == 2.0 || (mode == 0.0 && grid > (int)kReduceFlat);
    if (two_kernels && nwarps <= kPartialStride) {
        if (XW) allow_smem(k_spmv_sell_pk<EPI, EXACT, 2, U, MAXT, MINB, SMEM, DF, V6, XW, CD, TM>, sh);
        k_spmv_sell_pk<EPI, EXACT, 2, U, MAXT, MINB, SMEM, DF, V6, XW, CD, TM><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
        sum_warp_partials(nwarps, e, st);
    } else {
        if (XW) allow_smem(k_spmv_sell_pk<EPI, EXACT, 1, U, MAXT, MINB, SMEM, DF, V6, XW, CD, TM>, sh);
        k_spmv_sell_pk<EPI, EXACT, 1, U, MAXT, MINB, SMEM, DF, V6, XW, CD, TM><<<grid, threads, sh, st>>>(v, a.x, a.y, e);
    }
}

template <int EPI, bool EXACT, int U, int MAXT, int MINB, bool BOTH_DICTS>
static void launch_pk_v(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    const bool smem = v.ndict <= kPkDictSmem && opt("spmv.pkdict", 0.0) == 0.0;
    const int df = a.dotv != nullptr ? (int)opt("spmv.pkdotform", 2.0) : 0;
    if constexpr (BOTH_DICTS) {
        if (!smem) { launch_pk_d<EPI, EXACT, U, MAXT, MINB, false, 2>(v, a, e, grid, threads, st); return; }
    }
    if (df == 1) launch_pk_d<EPI, EXACT, U, MAXT, MINB, true, 1>(v, a, e, grid, threads, st);
    else if (df == 0) launch_pk_d<EPI, EXACT, U, MAXT, MINB, true, 0>(v, a, e, grid, threads, st);
    else launch_pk_d<EPI, EXACT, U, MAXT, MINB, true, 2>(v, a, e, grid, threads, st);
}

// "spmv.pkvariant" (plain product) / "spmv.pkdotvariant" (fused dot): unroll x register cap.
//   0 = U8/<=40 regs, 1 = U8/<=48, 2 = U4/<=32, 3 = U16/<=64, 4 = U16/<=48, 5 = U8/256 threads/<=40,
//   6 = U12/<=48, 7 = U4/<=40, 8 = U8/CTAs up to 1024 threads/<=64, 9 = U8/<=512 threads/<=40,
//   10 = U16/<=512 threads/<=64; software-pipelined: 11 = U4/<=40, 12 = U8/<=48, 13 = U6/<=40, 14 = U4/<=512 threads/<=40; 15 = variant 9 with the stream loaded L1::no_allocate
constexpr int kPkCdictDotVariant = 17; // measured on C2 inside PCG (profiles/r01_probe_cdict*.json): U15, 384-thread CTAs, <= 56 registers
constexpr int kPkDefaultVariant = 9;   // measured on C2 (profiles/r01_probe_blockorder.json): 512-thread CTAs x 128 slices

template <int EPI, bool EXACT>
static bool launch_pk(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st)
{
    if (v.vidx != nullptr) {       // six-byte form: one geometry, dictionary in shared memory or through L1
        if (v.ndict <= kPkDictSmem && opt("spmv.pkdict", 0.0) == 0.0) launch_pk_d<EPI, EXACT, 8, 512, 3, true, 2, true>(v, a, e, grid, threads, st);
        else launch_pk_d<EPI, EXACT, 8, 512, 3, false, 2, true>(v, a, e, grid, threads, st);
        return true;
    }
    const bool smem_ok = v.ndict <= kPkDictSmem && opt("spmv.pkdict", 0.0) == 0.0;
    // x window in shared memory: row-ordered vectors only (the CTA's rows are then a stretch of x), at most
    // kXwMaxSmem bytes per CTA, and not with the carried x update (it would need the window of xu as well)
    if (a.row_ordered && smem_ok && a.xu == nullptr && opt("spmv.pk_xwin", 0.0) != 0.0 &&
        xw_smem_bytes(v, grid) <= kXwMaxSmem) {
        if constexpr (EPI == EPI_NONE && EXACT) {      // a measured option: the plain and the fused-dot product only
            const int xv = (int)opt("spmv.pk_xwin", 0.0);
            if (xv == 2) launch_pk_d<EPI_NONE, true, 8, 512, 2, true, 2, false, true>(v, a, e, grid, threads, st);
            else if (xv == 3) launch_pk_d<EPI_NONE, true, 8, 1024, 1, true, 2, false, true>(v, a, e, grid, threads, st);
            else if (xv == 4) launch_pk_d<EPI_NONE, true, 4, 512, 3, true, 2, false, true>(v, a, e, grid, threads, st);
            else if (xv == 5) launch_pk_d<EPI_NONE, true, 6, 512, 3, true, 2, false, true>(v, a, e, grid, threads, st);
            else launch_pk_d<EPI_NONE, true, 8, 512, 3, true, 2, false, true>(v, a, e, grid, threads, st);
            return true;
        }
    }
    // dictionary in the parameter block (CD): default geometry only
    // (measured on C2, profiles/r01_notes.md third part: fused-dot product 1.90 -> 1.73 ms, PCG +6.8 %)
    // "spmv.pk_cdict" (plain product) / "spmv.pk_cdict_dot" (fused dot): 0 = off, else unroll x CTA size x resident CTAs
    const int cv = (int)(a.dotv != nullptr ? opt("spmv.pk_cdict_dot", (double)kPkCdictDotVariant) : opt("spmv.pk_cdict", 1.0));
    if (smem_ok && v.cdict_n > 0 && cv != 0) {
        if (EPI == EPI_NONE && EXACT) {
#define DB_CD(U_, T_, B_) launch_pk_d<EPI_NONE, true, U_, T_, B_, false, 2, false, false, true>(v, a, e, grid, threads, st); return true
            switch (cv) {
                case 2: DB_CD(16, 512, 2);
                case 3: DB_CD(10, 512, 3);
                case 4: DB_CD(8, 256, 6);
                case 5: DB_CD(16, 1024, 1);
                case 6: DB_CD(12, 256, 5);
                case 7: DB_CD(16, 256, 4);
                case 8: DB_CD(14, 384, 3);
                case 9: DB_CD(12, 384, 3);
                case 10: DB_CD(20, 256, 3);
                case 11: DB_CD(32, 512, 1);
                case 12: DB_CD(24, 128, 5);
                case 13: DB_CD(14, 192, 6);
#define DB_CDT(U_, T_, B_) launch_pk_d<EPI_NONE, true, U_, T_, B_, false, 2, false, false, true, true>(v, a, e, grid, threads, st); return true
                case 101: DB_CDT(8, 512, 3);           // 100 + n: variant n with the masked tail batch
                case 102: DB_CDT(16, 512, 2);
                case 103: DB_CDT(10, 512, 3);
                case 108: DB_CDT(14, 384, 3);
                case 109: DB_CDT(12, 384, 3);
                case 110: DB_CDT(20, 256, 3);
                case 117: DB_CDT(15, 384, 3);
                case 118: DB_CDT(16, 384, 3);
                case 120: DB_CDT(7, 512, 3);
                case 121: DB_CDT(9, 512, 3);
#undef DB_CDT
#define DB_CDF(U_, T_, B_, DF_) launch_pk_d<EPI_NONE, true, U_, T_, B_, false, DF_, false, false, true>(v, a, e, grid, threads, st); return true
                case 14: DB_CDF(14, 384, 3, 0);        // dot accumulator in registers
                case 15: DB_CDF(12, 384, 3, 0);
                case 16: DB_CDF(10, 384, 3, 0);
                case 17: DB_CDF(15, 384, 3, 2);
                case 18: DB_CDF(16, 384, 3, 2);
                case 19: DB_CDF(14, 384, 3, 1);        // dot in a tail pass over the CTA's rows
#undef DB_CDF
                default: break;
            }
#undef DB_CD
        }
        launch_pk_d<EPI, EXACT, 8, 512, 3, false, 2, false, false, true>(v, a, e, grid, threads, st);
        return true;
    }
    const int variant = (int)(a.dotv != nullptr ? opt("spmv.pkdotvariant", (double)kPkDefaultVariant)
                                                : opt("spmv.pkvariant", (double)kPkDefaultVariant));
    if (EPI == EPI_NONE && EXACT && smem_ok) {
        switch (variant) {
            case 1: launch_pk_v<EPI_NONE, true, 8, 128, 10, false>(v, a, e, grid, threads, st); return true;
            case 2: launch_pk_v<EPI_NONE, true, 4, 128, 16, false>(v, a, e, grid, threads, st); return true;
            case 3: launch_pk_v<EPI_NONE, true, 16, 128, 8, false>(v, a, e, grid, threads, st); return true;
            case 4: launch_pk_v<EPI_NONE, true, 16, 128, 10, false>(v, a, e, grid, threads, st); return true;
            case 5: launch_pk_v<EPI_NONE, true, 8, 256, 6, false>(v, a, e, grid, threads, st); return true;
            case 6: launch_pk_v<EPI_NONE, true, 12, 128, 10, false>(v, a, e, grid, threads, st); return true;
            case 7: launch_pk_v<EPI_NONE, true, 4, 128, 12, false>(v, a, e, grid, threads, st); return true;
            case 8: launch_pk_v<EPI_NONE, true, 8, 1024, 1, false>(v, a, e, grid, threads, st); return true;
            case 0: launch_pk_v<EPI_NONE, true, 8, 128, 12, false>(v, a, e, grid, threads, st); return true;
            case 10: launch_pk_v<EPI_NONE, true, 16, 512, 2, false>(v, a, e, grid, threads, st); return true;
            case 11: launch_pk_v<EPI_NONE, true, -4, 128, 12, false>(v, a, e, grid, threads, st); return true;
            case 12: launch_pk_v<EPI_NONE, true, -8, 128, 10, false>(v, a, e, grid, threads, st); return true;
            case 13: launch_pk_v<EPI_NONE, true, -6, 128, 12, false>(v, a, e, grid, threads, st); return true;
            case 14: launch_pk_v<EPI_NONE, true, -4, 512, 3, false>(v, a, e, grid, threads, st); return true;
            case 15: launch_pk_v<EPI_NONE, true, 108, 512, 3, false>(v, a, e, grid, threads, st); return true;
            default: break;
        }
    }
    launch_pk_v<EPI, EXACT, 8, 512, 3, true>(v, a, e, grid, threads, st);
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

// defined in spmv_pk_fma.cu
void launch_pk_epi_fma(const PkView& v, const SpmvArgs& a, const EpiArgs& e, int grid, int threads, cudaStream_t st);

}  // namespace db200

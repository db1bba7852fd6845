// matrix.cu -- device-side SparseMatBaseCL: upload, diagonal, sliced-ELL build, transpose.
//
// Reference container: SparseMatBaseCL<double>, num/spmat.h:682-879 (CSR with size_t indices,
// 16 B per stored entry).  HBM layout here, per matrix:
//   csr  : rowptr int64[rows+1], col int32[nnz], val f64[nnz]          (12 B per entry)
//   sell : slice_ptr int64[nslices+1], col int32[padded], val f64[padded], perm int32[32*nslices]
//          slice = 32 rows (one warp), entry j of the 32 rows contiguous; rows are sorted by
//          length (descending, ties by row index) inside windows of "sell.sigma" rows so that the
//          32 rows of a slice have (almost) equal length; col = -1 marks padding.
//   diag : f64[rows]  first stored entry with col == row (what JACPcCL divides by,
//          num/precond.h:153-169; GetDiag num/spmat.h:993-1002 on ascending columns is the same).
#include "common.cuh"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <limits>

namespace db200 {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__global__ void k_extract_diag(long long rows, const int64_t* __restrict__ rowptr,
                               const int32_t* __restrict__ col, const double* __restrict__ val,
                               double* __restrict__ diag, int* __restrict__ missing)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int64_t b = rowptr[i], e = rowptr[i + 1];
    int64_t k = b;
    while (k < e && col[k] != (int32_t)i) ++k;
    if (k < e) diag[i] = val[k];
    else { diag[i] = 0.0; atomicExch(missing, 1); }
}

// One CTA sorts one window of SIGMA rows by (length descending, [stencil type,] row ascending) with a
// bitonic network in shared memory and emits perm + per-slice widths (in stored entries, i.e. width*32).
// type_val != nullptr ("sell.type_sort", operators that get the packed stream): rows of one length are
// grouped by a 20-bit hash of their value sequence.  On a constant-coefficient mesh operator that is the
// stencil class (x-, y-, z-edge ... rows), so the 32 rows of a slice hold the same stencil: their j-th
// columns are one run of the row-ordered x (fewer 128-byte lines per gather) and their j-th dictionary
// index is one address.  Any order inside a window is valid; colliding hashes only merge two groups.
// Measured on C2 (profiles/r01_probe_type.json): no gain (fused-dot product 1.474 -> 1.492 ms), off by default.
template <int SIGMA>
__global__ void __launch_bounds__(1024)
k_sell_sort(long long rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rowmap,
            int32_t* __restrict__ perm, int64_t* __restrict__ slice_entries, const double* __restrict__ type_val)
{
    // rows = number of rows of the operator; rowmap (optional) = the CSR row behind each of them
    __shared__ unsigned long long key[SIGMA];
    const long long w0 = (long long)blockIdx.x * SIGMA;
    for (int t = threadIdx.x; t < SIGMA; t += blockDim.x) {
        const long long r = w0 + t;
        if (r < rows) {
            const long long ra = rowmap ? rowmap[r] : r;
            const unsigned long long len = (unsigned long long)(rowptr[ra + 1] - rowptr[ra]);
            const unsigned long long inv = 0x7fffffffULL - (len > 0x7fffffffULL ? 0x7fffffffULL : len);
            unsigned long long type = 0;
            if (type_val != nullptr) {
                unsigned h = 2166136261u;
                for (int64_t k = rowptr[ra]; k < rowptr[ra + 1]; ++k) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(__ldg(type_val + k));
                    h = (h ^ (unsigned)bits) * 16777619u;
                    h = (h ^ (unsigned)(bits >> 32)) * 16777619u;
                }
                type = (h ^ (h >> 20)) & 0xfffffu;
            }
            key[t] = (inv << 32) | (type << 12) | (unsigned long long)(unsigned)t;      // t < SIGMA <= 4096
        } else {
            key[t] = ~0ULL;
        }
    }
    __syncthreads();
    for (int k = 2; k <= SIGMA; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < SIGMA; t += blockDim.x) {
                const int l = t ^ j;
                if (l > t) {
                    const unsigned long long a = key[t], b = key[l];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { key[t] = b; key[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < SIGMA; t += blockDim.x) {
        const unsigned long long kk = key[t];
        const bool pad = (kk == ~0ULL);
        const long long r = w0 + (long long)(kk & 0xfffULL);
        perm[w0 + t] = pad ? -1 : (rowmap ? rowmap[r] : (int32_t)r);
        if ((t & 31) == 0) {
            const long long len = pad ? 0 : (long long)(0x7fffffffULL - (kk >> 32));
            slice_entries[(w0 + t) >> 5] = len * 32;
        }
    }
}

// sigma = 0/1: identity order (no sorting), widths = max length inside each slice.
__global__ void k_sell_identity(long long rows, long long padded_rows, const int64_t* __restrict__ rowptr,
                                const int32_t* __restrict__ rowmap, int32_t* __restrict__ perm,
                                int64_t* __restrict__ slice_entries)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= padded_rows) return;
    long long len = 0;
    if (r < rows) {
        const long long ra = rowmap ? rowmap[r] : r;
        perm[r] = (int32_t)ra; len = rowptr[ra + 1] - rowptr[ra];
    }
    else perm[r] = -1;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, len, m);
        len = o > len ? o : len;
    }
    if ((threadIdx.x & 31) == 0) slice_entries[r >> 5] = len * 32;
}

// One warp fills one slice.
__global__ void k_sell_fill(long long nslices, const int64_t* __restrict__ slice_ptr,
                            const int32_t* __restrict__ perm, const int64_t* __restrict__ rowptr,
                            const int32_t* __restrict__ col, const double* __restrict__ val,
                            int32_t* __restrict__ scol, double* __restrict__ sval)
{
    const int lane = threadIdx.x & 31;
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int64_t base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) >> 5);
    const int32_t row = perm[s * 32 + lane];
    int64_t b = 0; int len = 0;
    if (row >= 0) { b = rowptr[row]; len = (int)(rowptr[row + 1] - b); }
    for (int j = 0; j < width; ++j) {
        int32_t c = -1; double v = 0.0;
        if (j < len) { c = col[b + j]; v = val[b + j]; }
        scol[base + (int64_t)j * 32 + lane] = c;
        sval[base + (int64_t)j * 32 + lane] = v;
    }
}


// ---- value dictionary -----------------------------------------------------------------------
// Distinct bit patterns of the stored values, collected in an open-addressing hash set.  Lanes of
// a warp that hold the same pattern elect one leader; a slot that already holds the key costs one
// cached read and no atomic, so the pass runs at the speed of reading `val` once.
constexpr unsigned long long kDictEmpty = ~0ULL;          // (a NaN payload; a matrix holding it is not compressed)
constexpr int kDictTableBits = 17;
constexpr int kDictMax = 65535;

__device__ __forceinline__ unsigned dict_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (unsigned)k & ((1u << kDictTableBits) - 1u);
}

__global__ void __launch_bounds__(256)
k_dict_collect(long long n, const int32_t* __restrict__ col, const double* __restrict__ val,
               unsigned long long* table, int* count)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
        int over = 0;                                                 // too many: give up early (warp-uniform)
        if ((threadIdx.x & 31) == 0) over = *(volatile int*)count > kDictMax;
        if (__shfl_sync(0xffffffffu, over, 0)) return;
        const long long i = i0 + threadIdx.x;
        const bool live = i < n && (col == nullptr || __ldcs(col + i) >= 0);   // padding holds no value
        const unsigned long long key = live ? (unsigned long long)__double_as_longlong(__ldcs(val + i)) : 0ULL;
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (!live) continue;
        const unsigned same = __match_any_sync(act, key);
        if ((threadIdx.x & 31) != __ffs(same) - 1) continue;           // not the leader of its value
        if (key == kDictEmpty) { atomicAdd(count, kDictMax + 1); continue; }
        unsigned slot = dict_hash(key);
        for (int probe = 0; probe < (1 << kDictTableBits); ++probe) {
            unsigned long long cur = *(volatile unsigned long long*)(table + slot);
            if (cur == key) break;
            if (cur == kDictEmpty) {
                cur = atomicCAS(table + slot, kDictEmpty, key);
                if (cur == kDictEmpty) { atomicAdd(count, 1); break; }
                if (cur == key) break;
            }
            slot = (slot + 1) & ((1u << kDictTableBits) - 1u);
        }
    }
}

// pk = (index of val in the sorted dictionary << cbits) | col; padding = (ndict << cbits), i.e. column 0
// and the extra dictionary slot that holds 0.0 (spmv_pk.cu drops the padded product with a select).
__global__ void __launch_bounds__(256)
k_dict_pack(long long n, const int32_t* __restrict__ col, const double* __restrict__ val,
            const unsigned long long* __restrict__ keys, int ndict, int cbits, uint32_t* __restrict__ pk)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = __ldcs(col + i);
    if (c < 0) { pk[i] = (uint32_t)ndict << cbits; return; }
    const unsigned long long key = (unsigned long long)__double_as_longlong(__ldcs(val + i));
    int lo = 0, hi = ndict;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
    }
    pk[i] = ((uint32_t)lo << cbits) | (uint32_t)c;
}

// Only the values change (same pattern): refill sval from the CSR values.
__global__ void k_sell_refill_values(long long nslices, const int64_t* __restrict__ slice_ptr,
                                     const int32_t* __restrict__ perm, const int64_t* __restrict__ rowptr,
                                     const double* __restrict__ val, double* __restrict__ sval)
{
    const int lane = threadIdx.x & 31;
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int64_t base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) >> 5);
    const int32_t row = perm[s * 32 + lane];
    int64_t b = 0; int len = 0;
    if (row >= 0) { b = rowptr[row]; len = (int)(rowptr[row + 1] - b); }
    for (int j = 0; j < width; ++j)
        sval[base + (int64_t)j * 32 + lane] = j < len ? val[b + j] : 0.0;
}

// ---- LinComb: C = a*A + b*B, row-wise merge of two ascending column lists (spmat.h:1457-1551) ----
__global__ void k_lincomb_count(long long rows, const int64_t* __restrict__ pa, const int32_t* __restrict__ ca,
                                const int64_t* __restrict__ pb, const int32_t* __restrict__ cb, int64_t* __restrict__ len)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int64_t ia = pa[r], ib = pb[r];
    const int64_t ea = pa[r + 1], eb = pb[r + 1];
    int64_t n = 0;
    while (ia < ea && ib < eb) {
        const int32_t x = ca[ia], y = cb[ib];
        ia += (x <= y); ib += (y <= x); ++n;
    }
    len[r] = n + (ea - ia) + (eb - ib);
}

__global__ void k_lincomb_fill(long long rows, double a, const int64_t* __restrict__ pa, const int32_t* __restrict__ ca,
                               const double* __restrict__ va, double b, const int64_t* __restrict__ pb,
                               const int32_t* __restrict__ cb, const double* __restrict__ vb,
                               const int64_t* __restrict__ pc, int32_t* __restrict__ cc, double* __restrict__ vc)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int64_t ia = pa[r], ib = pb[r], i = pc[r];
    const int64_t ea = pa[r + 1], eb = pb[r + 1];
    while (ia < ea && ib < eb) {
        const int32_t x = ca[ia], y = cb[ib];
        if (y < x) { vc[i] = __dmul_rn(b, vb[ib]); cc[i] = y; ++ib; }                                  // :1524-1527
        else if (x < y) { vc[i] = __dmul_rn(a, va[ia]); cc[i] = x; ++ia; }                             // :1528-1531
        else { vc[i] = __dadd_rn(__dmul_rn(a, va[ia]), __dmul_rn(b, vb[ib])); cc[i] = y; ++ia; ++ib; } // :1532-1535
        ++i;
    }
    for (; ia < ea; ++ia, ++i) { vc[i] = __dmul_rn(a, va[ia]); cc[i] = ca[ia]; }
    for (; ib < eb; ++ib, ++i) { vc[i] = __dmul_rn(b, vb[ib]); cc[i] = cb[ib]; }
}

// ---- transpose ---------------------------------------------------------------------------
__global__ void k_count_cols(long long nnz, const int32_t* __restrict__ col, unsigned long long* __restrict__ cnt)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) atomicAdd(cnt + col[k], 1ULL);
}

__global__ void k_transpose_fill(long long rows, const int64_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ col, const double* __restrict__ val,
                                 const int64_t* __restrict__ tptr, unsigned long long* __restrict__ cursor,
                                 int32_t* __restrict__ tcol, double* __restrict__ tval)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int32_t c = col[k];
        const unsigned long long pos = atomicAdd(cursor + c, 1ULL);
        tcol[tptr[c] + (int64_t)pos] = (int32_t)i;
        tval[tptr[c] + (int64_t)pos] = val[k];
    }
}

// Insertion-sort each transposed row by source row: y_ATx (num/spmat.h:1729-1746) adds the
// contributions to y[c] in ascending row order, and so must we.
__global__ void k_sort_rows(long long rows, const int64_t* __restrict__ ptr, int32_t* __restrict__ col,
                            double* __restrict__ val)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int64_t b = ptr[i], e = ptr[i + 1];
    for (int64_t k = b + 1; k < e; ++k) {
        const int32_t c = col[k]; const double v = val[k];
        int64_t j = k - 1;
        while (j >= b && col[j] > c) { col[j + 1] = col[j]; val[j + 1] = val[j]; --j; }
        col[j + 1] = c; val[j + 1] = v;
    }
}

__global__ void k_widen_i32(long long n, const int32_t* __restrict__ in, unsigned long long* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (unsigned long long)(long long)in[k];
}

__global__ void k_copy_i64(long long n, const int64_t* __restrict__ in, unsigned long long* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (unsigned long long)in[k];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int blocks_for(long long n, int threads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

static int exclusive_scan_i64(int64_t* d_inout, size_t count, cudaStream_t st)
{
    void* tmp = nullptr; size_t tmp_bytes = 0;
    DB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_inout, d_inout, (long long)count, st));
    DB_CHECK(dev_alloc((char**)&tmp, tmp_bytes));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_inout, d_inout, (long long)count, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    dev_free(tmp);
    DB_CUDA(e);
    DB_CUDA(e2);
    count_launch(2);
    return 0;
}

double* work_vec(drops_b200_matrix_s* M, int i)
{
    if ((int)M->work.size() <= i) M->work.resize(i + 1, nullptr);
    if (!M->work[i]) {
        const size_t n = std::max(std::max(M->csr.rows, M->csr.cols), M->sell.npos);   // row-ordered solves: 32 * nslices positions
        if (dev_alloc(&M->work[i], n) < 0) return nullptr;
    }
    return M->work[i];
}

void free_sell(SellOp& s)
{
    dev_free(s.slice_ptr); dev_free(s.col); dev_free(s.val); dev_free(s.perm);
    dev_free(s.colp);
    dev_free(s.pk); dev_free(s.dict); dev_free(s.pkp); dev_free(s.vidx); dev_free(s.pk_nopad);
    dev_free(s.scd); dev_free(s.sc);
    s = SellOp();
}

void free_csr(CsrDev& c)
{
    dev_free(c.rowptr); dev_free(c.col); dev_free(c.val);
    c = CsrDev();
}

// Builds the sliced-ELL operator of all rows of A (rowmap == nullptr) or of the `nsub` rows
// listed in rowmap (device array, ascending); perm then holds rows of A, so y, aux and the
// fused dot are addressed by the full row index.
int build_sell(const CsrDev& A, SellOp& S, size_t* bytes, const int32_t* rowmap, size_t nsub, bool rowmap_is_permutation,
               bool type_sort)
{
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    free_sell(S);
    const size_t op_rows = rowmap ? nsub : A.rows;
    S.rows = op_rows; S.cols = A.cols;
    S.full_rows = rowmap == nullptr || (rowmap_is_permutation && nsub == A.rows);
    int sigma = (int)opt("sell.sigma", 4096.0);
    const bool sorted = sigma >= 64;
    if (sorted) {
        if (sigma >= 4096) sigma = 4096; else if (sigma >= 2048) sigma = 2048;
        else if (sigma >= 1024) sigma = 1024; else if (sigma >= 512) sigma = 512;
        else if (sigma >= 256) sigma = 256; else if (sigma >= 128) sigma = 128; else sigma = 64;
    } else sigma = 32;
    const size_t nwin = (op_rows + sigma - 1) / sigma;
    const size_t padded_rows = std::max<size_t>(nwin * sigma, 32);
    S.nslices = padded_rows / 32;
    DB_CHECK(dev_alloc(&S.perm, padded_rows));
    DB_CHECK(dev_alloc(&S.slice_ptr, S.nslices + 1));
    DB_CUDA(cudaMemsetAsync(S.slice_ptr, 0, (S.nslices + 1) * sizeof(int64_t), st));
    if (op_rows > 0) {
        if (!sorted) {
            k_sell_identity<<<blocks_for(padded_rows, 256), 256, 0, st>>>((long long)op_rows, (long long)padded_rows, A.rowptr, rowmap, S.perm, S.slice_ptr);
        } else {
            const int g = (int)nwin;
            const double* tv = (type_sort && opt("sell.type_sort", 0.0) != 0.0) ? A.val : nullptr;
            switch (sigma) {
                case 4096: k_sell_sort<4096><<<g, 1024, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
                case 2048: k_sell_sort<2048><<<g, 1024, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
                case 1024: k_sell_sort<1024><<<g, 512, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
                case 512:  k_sell_sort<512><<<g, 256, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
                case 256:  k_sell_sort<256><<<g, 128, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
                case 128:  k_sell_sort<128><<<g, 64, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
                default:   k_sell_sort<64><<<g, 32, 0, st>>>((long long)op_rows, A.rowptr, rowmap, S.perm, S.slice_ptr, tv); break;
            }
        }
        count_launch();
        DB_CUDA(cudaGetLastError());
    } else {
        DB_CUDA(cudaMemsetAsync(S.perm, 0xff, padded_rows * sizeof(int32_t), st));
    }
    DB_CHECK(exclusive_scan_i64(S.slice_ptr, S.nslices + 1, st));
    int64_t total = 0;
    DB_CUDA(cudaMemcpyAsync(&total, S.slice_ptr + S.nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DB_CUDA(cudaStreamSynchronize(st));
    S.padded = (size_t)total;
    DB_CHECK(dev_alloc(&S.col, S.padded));
    DB_CHECK(dev_alloc(&S.val, S.padded));
    if (S.padded > 0) {
        k_sell_fill<<<blocks_for((long long)S.nslices * 32, 256), 256, 0, st>>>(
            (long long)S.nslices, S.slice_ptr, S.perm, A.rowptr, A.col, A.val, S.col, S.val);
        count_launch();
        DB_CUDA(cudaGetLastError());
    }
    DB_CUDA(cudaStreamSynchronize(st));
    const size_t col_bytes = S.padded * 4;
    if (bytes) *bytes += S.padded * 8 + col_bytes + (S.nslices + 1) * 8 + padded_rows * 4;
    DB_CHECK(build_value_dictionary(S, bytes));
    return build_row_ordered_columns(S, bytes);
}

// vidx = index of val in the sorted dictionary (16 bit); padding (col < 0) gets slot ndict = 0.0
__global__ void __launch_bounds__(256)
k_dict_index16(long long n, const int32_t* __restrict__ col, const double* __restrict__ val,
               const unsigned long long* __restrict__ keys, int ndict, uint16_t* __restrict__ vidx)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (__ldcs(col + i) < 0) { vidx[i] = (uint16_t)ndict; return; }
    const unsigned long long key = (unsigned long long)__double_as_longlong(__ldcs(val + i));
    int lo = 0, hi = ndict;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
    }
    vidx[i] = (uint16_t)lo;
}

// idx = index of val in the sorted list of distinct values (8 bit)
__global__ void __launch_bounds__(256)
k_value_index8(long long n, const double* __restrict__ val, const unsigned long long* __restrict__ keys, int nkeys,
               uint8_t* __restrict__ idx)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = (unsigned long long)__double_as_longlong(__ldcs(val + i));
    int lo = 0, hi = nkeys;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
    }
    idx[i] = (uint8_t)lo;
}

// A vector with at most max_values (<= 256) distinct bit patterns as one byte per entry + the list of values
// (host_values, sorted by bit pattern).  *count = 0 and *d_idx = nullptr when the vector holds more.
// Used for the Jacobi diagonal of constant-coefficient operators (krylov.cu): the PCG vector kernels then
// read 1 instead of 8 bytes per entry of d.
int build_value_index8(const double* d_val, size_t n, int max_values, uint8_t** d_idx, double* host_values, int* count)
{
    *d_idx = nullptr; *count = 0;
    if (n == 0 || max_values < 1) return 0;
    if (max_values > 256) max_values = 256;
    cudaStream_t st = ctx().stream;
    unsigned long long* table = nullptr; int* d_count = nullptr;
    const size_t cap = (size_t)1 << kDictTableBits;
    DB_CHECK(dev_alloc(&table, cap));
    if (dev_alloc(&d_count, 1) < 0) { dev_free(table); return DROPS_B200_E_NOMEM; }
    int status = 0, cnt = 0;
    cudaError_t e = cudaMemsetAsync(table, 0xff, cap * sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, sizeof(int), st);
    if (e == cudaSuccess) {
        const int grid = (int)std::min<long long>(((long long)n + 255) / 256, (long long)ctx().sm_count * 16);
        k_dict_collect<<<grid, 256, 0, st>>>((long long)n, nullptr, d_val, table, d_count);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && cnt >= 1 && cnt <= max_values) {
        std::vector<unsigned long long> host(cap), keys;
        e = cudaMemcpyAsync(host.data(), table, cap * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) {
            for (unsigned long long k : host) if (k != kDictEmpty) keys.push_back(k);
            std::sort(keys.begin(), keys.end());
            if ((int)keys.size() != cnt) { set_error("value index: %zu keys in the table, %d counted", keys.size(), cnt); status = DROPS_B200_E_CUDA; }
        }
        if (e == cudaSuccess && status == 0) {
            e = cudaMemcpyAsync(table, keys.data(), keys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) status = dev_alloc(d_idx, n);
            if (e == cudaSuccess && status == 0) {
                k_value_index8<<<blocks_for((long long)n, 256), 256, 0, st>>>((long long)n, d_val, table, cnt, *d_idx);
                count_launch();
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            }
            if (e == cudaSuccess && status == 0) {
                std::memcpy(host_values, keys.data(), keys.size() * sizeof(double));
                *count = cnt;
            }
        }
    }
    dev_free(table); dev_free(d_count);
    if (e != cudaSuccess) status = cuda_fail(e, "value index", __FILE__, __LINE__);
    if (status != 0) { dev_free(*d_idx); *d_idx = nullptr; *count = 0; }
    return status;
}

// one warp per slice: does the slice hold a padding word?  (rows are sorted by length inside a window, so
// only the last entry-row can differ between the lanes; slices of padding rows are all padding)
__global__ void k_slice_nopad(long long nslices, const int64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ pk,
                              uint32_t wpad, uint8_t* __restrict__ nopad)
{
    const int lane = threadIdx.x & 31;
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int64_t base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) >> 5);
    bool pad = false;
    for (int j = 0; j < width; ++j) pad |= pk[base + (int64_t)j * 32 + lane] >= wpad;
    const unsigned any = __ballot_sync(0xffffffffu, pad);
    if (lane == 0) nopad[s] = any == 0u ? 1 : 0;
}

__global__ void k_inverse_perm(long long npos, const int32_t* __restrict__ perm, int32_t* __restrict__ iperm)
{
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int32_t r = perm[pos];
    if (r >= 0) iperm[r] = (int32_t)pos;
}

// the packed stream with every column replaced by its position in the operator's row order
__global__ void __launch_bounds__(256)
k_pack_reorder(long long n, const uint32_t* __restrict__ pk, const int32_t* __restrict__ iperm, int cbits, uint32_t wpad,
               uint32_t* __restrict__ pkp)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t w = __ldcs(pk + i);
    if (w >= wpad) { pkp[i] = w; return; }
    const uint32_t cmask = (1u << cbits) - 1u;
    pkp[i] = (w & ~cmask) | (uint32_t)__ldg(iperm + (w & cmask));
}

static int bits_for(unsigned long long count)      // bits that distinguish `count` codes 0..count-1
{
    int b = 0;
    while ((1ULL << b) < count) ++b;
    return b;
}

// Will an operator with these values and `cols` columns (positions: up to cols + slack) get the one-word
// packed stream?  One pass over val; decides up front whether the rows are cut in graph blocks (reorder.cu),
// which only pays together with the row-ordered twin of that stream.
int will_pack_one_word(const double* d_val, size_t nnz, size_t cols_and_positions, bool* yes)
{
    *yes = false;
    if (nnz == 0 || opt("sell.valdict", 1.0) == 0.0) return 0;
    cudaStream_t st = ctx().stream;
    unsigned long long* table = nullptr; int* d_count = nullptr;
    const size_t cap = (size_t)1 << kDictTableBits;
    DB_CHECK(dev_alloc(&table, cap));
    if (dev_alloc(&d_count, 1) < 0) { dev_free(table); return DROPS_B200_E_NOMEM; }
    int count = 0;
    cudaError_t e = cudaMemsetAsync(table, 0xff, cap * sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, sizeof(int), st);
    if (e == cudaSuccess) {
        const int grid = (int)std::min<long long>(((long long)nnz + 255) / 256, (long long)ctx().sm_count * 16);
        k_dict_collect<<<grid, 256, 0, st>>>((long long)nnz, nullptr, d_val, table, d_count);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(table); dev_free(d_count);
    if (e != cudaSuccess) return cuda_fail(e, "value dictionary: count", __FILE__, __LINE__);
    const int cbits = std::max(1, bits_for((unsigned long long)std::max<size_t>(cols_and_positions, 1)));
    // (+1: padding may add the value 0.0; +1: the padding slot)
    *yes = count >= 1 && count + 1 <= kDictMax && cbits + bits_for((unsigned long long)count + 2) <= (int)opt("sell.pack_bits", 32.0);
    return 0;
}

int build_value_dictionary(SellOp& S, size_t* bytes)
{
    cudaStream_t st = ctx().stream;
    dev_free(S.pk); dev_free(S.dict); dev_free(S.pkp); dev_free(S.vidx); dev_free(S.pk_nopad);
    S.pk_nopad = nullptr; S.pk = nullptr; S.dict = nullptr; S.pkp = nullptr; S.vidx = nullptr; S.ndict = 0; S.ndict_host = 0; S.cbits = 0;
    dev_free(S.scd); S.scd = nullptr; dev_free(S.sc); S.sc = nullptr; S.sc_slices = S.sc_entries = 0;
    if (S.colp == nullptr) S.npos = 0;
    if (S.padded == 0 || S.col == nullptr || S.val == nullptr || opt("sell.valdict", 1.0) == 0.0) return 0;
    // columns and, for the row-ordered twin, positions (32 * nslices >= rows) share the low bits
    const int cbits = std::max(1, bits_for((unsigned long long)std::max<size_t>(std::max(S.cols, S.nslices * 32), 1)));
    if (cbits > 31) return 0;
    unsigned long long* table = nullptr; int* d_count = nullptr;
    const size_t cap = (size_t)1 << kDictTableBits;
    DB_CHECK(dev_alloc(&table, cap));
    if (dev_alloc(&d_count, 1) < 0) { dev_free(table); return DROPS_B200_E_NOMEM; }
    int status = 0, count = 0;
    std::vector<unsigned long long> keys;
    auto fail = [&](cudaError_t e, const char* what) { status = cuda_fail(e, what, __FILE__, __LINE__); };
    cudaError_t e = cudaMemsetAsync(table, 0xff, cap * sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, sizeof(int), st);
    if (e == cudaSuccess) {
        const int grid = (int)std::min<long long>(((long long)S.padded + 255) / 256, (long long)ctx().sm_count * 16);
        k_dict_collect<<<grid, 256, 0, st>>>((long long)S.padded, S.col, S.val, table, d_count);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) fail(e, "value dictionary: collect");
    // one word per entry when column bits + index bits fit ("sell.pack_bits", 32; the index field has one
    // code more than ndict: slot ndict (0.0) belongs to the padding word); otherwise the dictionary still
    // pays as a 16-bit index beside the int32 column: 6 instead of 12 bytes per entry
    const bool one_word = cbits + bits_for((unsigned long long)count + 1) <= (int)opt("sell.pack_bits", 32.0);
    if (status == 0 && count >= 1 && count <= kDictMax) {
        std::vector<unsigned long long> host(cap);
        e = cudaMemcpyAsync(host.data(), table, cap * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) fail(e, "value dictionary: download");
        else {
            for (unsigned long long k : host) if (k != kDictEmpty) keys.push_back(k);
            std::sort(keys.begin(), keys.end());
            if ((int)keys.size() != count) { set_error("value dictionary: %zu keys in the table, %d counted", keys.size(), count); status = DROPS_B200_E_CUDA; }
        }
    } else count = 0;
    dev_free(d_count);
    if (status == 0 && count > 0) {
        // the hash table's storage doubles as the sorted key list for the packing pass
        e = cudaMemcpyAsync(table, keys.data(), keys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) fail(e, "value dictionary: upload");
        std::vector<double> dict_host(keys.size() + 1, 0.0);
        std::memcpy(dict_host.data(), keys.data(), keys.size() * sizeof(double));
        if (status == 0) status = dev_alloc(&S.dict, dict_host.size());
        if (status == 0) status = one_word ? dev_alloc(&S.pk, S.padded) : dev_alloc(&S.vidx, S.padded);
        if (status == 0) {
            e = cudaMemcpyAsync(S.dict, dict_host.data(), dict_host.size() * sizeof(double), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                if (one_word)
                    k_dict_pack<<<blocks_for((long long)S.padded, 256), 256, 0, st>>>((long long)S.padded, S.col, S.val, table, count, cbits, S.pk);
                else
                    k_dict_index16<<<blocks_for((long long)S.padded, 256), 256, 0, st>>>((long long)S.padded, S.col, S.val, table, count, S.vidx);
                count_launch();
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) fail(e, "value dictionary: pack");
        }
        if (status == 0 && one_word) {
            status = dev_alloc(&S.pk_nopad, S.nslices);
            if (status == 0) {
                k_slice_nopad<<<blocks_for((long long)S.nslices * 32, 256), 256, 0, st>>>((long long)S.nslices, S.slice_ptr, S.pk,
                                                                                         (uint32_t)count << cbits, S.pk_nopad);
                count_launch();
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) fail(e, "value dictionary: padding flags");
            }
        }
        if (status == 0) {
            S.ndict_host = 0;
            if (dict_host.size() <= (size_t)kDictConst) {
                std::memcpy(S.dict_host, dict_host.data(), dict_host.size() * sizeof(double));
                S.ndict_host = count;
            }
            S.ndict = count; S.cbits = one_word ? cbits : 0;
            if (bytes) *bytes += S.padded * (one_word ? 4 : 2) + keys.size() * 8;
        }
        // row-ordered twin: square operator over all its rows (perm is a permutation of 0..rows-1)
        if (status == 0 && one_word && S.full_rows && S.rows == S.cols && opt("sell.reorder", 1.0) != 0.0) {
            int32_t* iperm = nullptr;
            const long long npos = (long long)S.nslices * 32;
            status = dev_alloc(&iperm, S.cols);
            if (status == 0) status = dev_alloc(&S.pkp, S.padded);
            if (status == 0) {
                k_inverse_perm<<<blocks_for(npos, 256), 256, 0, st>>>(npos, S.perm, iperm);
                k_pack_reorder<<<blocks_for((long long)S.padded, 256), 256, 0, st>>>((long long)S.padded, S.pk, iperm, cbits,
                                                                                   (uint32_t)count << cbits, S.pkp);
                count_launch(2);
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) fail(e, "value dictionary: reorder");
            }
            dev_free(iperm);
            if (status == 0) { S.npos = (size_t)npos; if (bytes) *bytes += S.padded * 4; }
            else { dev_free(S.pkp); S.pkp = nullptr; }
            if (status == 0) status = build_segment_descriptors(S, bytes);
        }
        if (status != 0) {
            dev_free(S.pk); dev_free(S.dict); dev_free(S.vidx); dev_free(S.pk_nopad);
            S.pk = nullptr; S.dict = nullptr; S.vidx = nullptr; S.pk_nopad = nullptr; S.ndict = 0;
        }
    }
    dev_free(table);
    return status;
}

// One warp per slice of the row-ordered twin: is every entry-row (a) one dictionary index on all lanes (packed twin; the
// twin of the 12-byte stream keeps its values per lane) and (b) at most two runs of consecutive positions?  Then the
// slice is flagged and gets its descriptors (SellOp::scd).
template <bool PACKED>
__global__ void __launch_bounds__(256)
k_sc_build(long long nslices, const int64_t* __restrict__ slice_ptr, const uint32_t* __restrict__ pkp, const int32_t* __restrict__ colp,
           const uint8_t* __restrict__ nopad, int cbits, uint4* __restrict__ scd, uint8_t* __restrict__ sc,
           unsigned long long* __restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nslices) return;
    const int64_t base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) >> 5);
    bool ok = width > 0 && width <= kScMaxWidth && (!PACKED || (nopad != nullptr && nopad[s] != 0));
    const uint32_t cmask = PACKED ? (1u << cbits) - 1u : 0u;
    for (int j = 0; j < width && ok; ++j) {
        uint32_t idx = 0;
        int pos;
        if (PACKED) {
            const uint32_t w = pkp[base + (int64_t)j * 32 + lane];
            idx = w >> cbits;
            pos = (int)(w & cmask);
        } else {
            pos = colp[base + (int64_t)j * 32 + lane];
        }
        const int d = pos - lane;
        const uint32_t idx0 = __shfl_sync(0xffffffffu, idx, 0);
        const int d0 = __shfl_sync(0xffffffffu, d, 0);
        const unsigned m0 = __ballot_sync(0xffffffffu, d == d0);
        const int brk = __popc(m0);
        bool good = __all_sync(0xffffffffu, idx == idx0 && pos >= 0) && (brk == 32 || m0 == ((1u << brk) - 1u));
        int d1 = d0;
        if (good && brk < 32) {
            d1 = __shfl_sync(0xffffffffu, d, brk);
            const unsigned m1 = __ballot_sync(0xffffffffu, d == d1);
            good = (m0 | m1) == 0xffffffffu;
        }
        // positions are >= 0 and lane < 32: d0, d1 > -32; they are stored as signed 32-bit values.
        // (A per-entry-row fall-back -- records for the regular entry-rows of a slice, the twin's own words for the others --
        // was measured: the dependent load in the middle of a batch of gathers costs more than the records save, 1.22 -> 1.63 ms
        // on C2; a slice either runs from records entirely or not at all.)
        if (good) { if (lane == 0) scd[(base >> 5) + j] = make_uint4((unsigned)d0, (unsigned)d1, (unsigned)brk, idx0); }
        else ok = false;
    }
    if (lane == 0) {
        sc[s] = ok ? 1 : 0;
        if (ok) { atomicAdd(counts, 1ull); atomicAdd(counts + 1, (unsigned long long)width * 32ull); }
    }
}

// Segment descriptors of the row-ordered twin (SellOp::scd; "sell.segments", default on): of the packed twin when the
// dictionary rides in the kernel's parameter block, of the int32 twin of the 12-byte stream otherwise (variable
// coefficients on a structured mesh: the columns are described by 0.5 bytes per entry, the fp64 values stay -- 8.5
// instead of 12 bytes per entry).  Costs one pass over the twin; 16 bytes per 32 stored entries.
int build_segment_descriptors(SellOp& S, size_t* bytes)
{
    dev_free(S.scd); S.scd = nullptr; dev_free(S.sc); S.sc = nullptr; S.sc_slices = S.sc_entries = 0;
    if (S.nslices == 0 || opt("sell.segments", 1.0) == 0.0) return 0;
    const bool packed = S.pkp != nullptr;
    if (packed && (S.pk_nopad == nullptr || S.ndict_host == 0)) return 0;
    if (!packed && S.colp == nullptr) return 0;
    cudaStream_t st = ctx().stream;
    unsigned long long* d_counts = nullptr;
    DB_CHECK(dev_alloc(&d_counts, 2));
    int status = dev_alloc(&S.scd, S.padded / 32 + 1);
    if (status == 0) status = dev_alloc(&S.sc, S.nslices);
    if (status == 0) {
        cudaError_t e = cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), st);
        const int g = blocks_for((long long)S.nslices * 32, 256);
        if (packed)
            k_sc_build<true><<<g, 256, 0, st>>>((long long)S.nslices, S.slice_ptr, S.pkp, nullptr, S.pk_nopad, S.cbits, S.scd, S.sc, d_counts);
        else
            k_sc_build<false><<<g, 256, 0, st>>>((long long)S.nslices, S.slice_ptr, nullptr, S.colp, nullptr, 0, S.scd, S.sc, d_counts);
        count_launch();
        if (e == cudaSuccess) e = cudaGetLastError();
        unsigned long long h[2] = {0, 0};
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_counts, sizeof h, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) status = cuda_fail(e, "segment descriptors", __FILE__, __LINE__);
        S.sc_slices = (size_t)h[0]; S.sc_entries = (size_t)h[1];
    }
    dev_free(d_counts);
    // not worth a second code path for a handful of slices
    if (status != 0 || S.sc_entries * 10 < S.padded) {
        dev_free(S.scd); S.scd = nullptr; dev_free(S.sc); S.sc = nullptr; S.sc_slices = S.sc_entries = 0;
        return status;
    }
    if (bytes) *bytes += (S.padded / 32 + 1) * 16 + S.nslices;
    return 0;
}

// Row-ordered twin of S.pk for a caller-supplied numbering: iperm[column] = position (dist.cu: the
// positions of a partition run over its interior and its boundary operator).
int pack_reorder(SellOp& S, const int32_t* iperm, size_t npos)
{
    dev_free(S.pkp); S.pkp = nullptr; S.npos = 0;
    dev_free(S.scd); S.scd = nullptr; dev_free(S.sc); S.sc = nullptr; S.sc_slices = S.sc_entries = 0;
    if (S.pk == nullptr || S.padded == 0 || npos > ((size_t)1 << S.cbits)) return 0;
    cudaStream_t st = ctx().stream;
    DB_CHECK(dev_alloc(&S.pkp, S.padded));
    k_pack_reorder<<<blocks_for((long long)S.padded, 256), 256, 0, st>>>((long long)S.padded, S.pk, iperm, S.cbits,
                                                                       (uint32_t)S.ndict << S.cbits, S.pkp);
    count_launch();
    DB_CUDA(cudaGetLastError());
    S.npos = npos;
    return build_segment_descriptors(S, nullptr);
}

// Row-ordered twin of S.col for a square operator over all its rows that has NO packed stream (more than 65 535
// distinct values: variable coefficients, cut cells): colp = position of the column in the operator's own slice
// order, so that PCG iterates on row-ordered vectors on the 12-byte stream as well ("pcg.reorder"; the rows are then
// cut in blocks of the matrix graph like those of a packed operator).  Depends on the pattern only.
__global__ void k_col_positions(long long n, const int32_t* __restrict__ col, const int32_t* __restrict__ iperm,
                                int32_t* __restrict__ colp)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = __ldcs(col + i);
    colp[i] = c < 0 ? -1 : __ldg(iperm + c);
}

int build_row_ordered_columns(SellOp& S, size_t* bytes)
{
    dev_free(S.colp); S.colp = nullptr;
    if (S.pkp != nullptr) return 0;                       // the packed twin exists
    S.npos = 0;
    // six-byte form (16-bit value index beside the int32 column): its twin pairs the positions with the same indices, and
    // where the rows are line segments the positions collapse into segment records (k_spmv_sell_sci: 2.5 B per entry)
    if (S.vidx != nullptr && opt("sell.reorder_six", 1.0) == 0.0) return 0;
    if (S.padded == 0 || S.col == nullptr || !S.full_rows || S.rows != S.cols || opt("sell.reorder", 1.0) == 0.0) return 0;
    cudaStream_t st = ctx().stream;
    const long long npos = (long long)S.nslices * 32;
    int32_t* iperm = nullptr;
    DB_CHECK(dev_alloc(&iperm, S.cols));
    int status = dev_alloc(&S.colp, S.padded);
    if (status == 0) {
        k_inverse_perm<<<blocks_for(npos, 256), 256, 0, st>>>(npos, S.perm, iperm);
        k_col_positions<<<blocks_for((long long)S.padded, 256), 256, 0, st>>>((long long)S.padded, S.col, iperm, S.colp);
        count_launch(2);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) status = cuda_fail(e, "row-ordered columns", __FILE__, __LINE__);
    }
    dev_free(iperm);
    if (status != 0) { dev_free(S.colp); S.colp = nullptr; return status; }
    S.npos = (size_t)npos;
    if (bytes) *bytes += S.padded * 4;
    return build_segment_descriptors(S, bytes);
}

// Footprint ("sell.release_general", operators of at least "sell.release_min_rows" rows): once the one-word packed
// stream exists, nothing on the solver path reads the int32 column + fp64 value arrays of the sliced-ELL operator any
// more (products, solvers and the transposed product all run on pk / pkp); they are 12 of the 16-20 bytes per stored
// entry the operator holds.  A value update rebuilds them from the CSR copy for the time of the rebuild.
void release_general(SellOp& S, size_t* bytes)
{
    if ((S.pk == nullptr && S.vidx == nullptr) || S.val == nullptr || opt("sell.release_general", 1.0) == 0.0) return;
    if ((double)S.rows < opt("sell.release_min_rows", 2.0e6)) return;
    size_t freed = S.padded * 8;
    if (S.pk != nullptr) { dev_free(S.col); S.col = nullptr; freed += S.padded * 4; }     // (the six-byte form reads col)
    dev_free(S.val); S.val = nullptr;
    S.general_released = true;
    if (bytes && *bytes >= freed) *bytes -= freed;
}

static int restore_general(const CsrDev& A, SellOp& S)
{
    if (!S.general_released || S.val != nullptr) return 0;
    if (S.col == nullptr) DB_CHECK(dev_alloc(&S.col, S.padded));
    DB_CHECK(dev_alloc(&S.val, S.padded));
    k_sell_fill<<<blocks_for((long long)S.nslices * 32, 256), 256, 0, ctx().stream>>>(
        (long long)S.nslices, S.slice_ptr, S.perm, A.rowptr, A.col, A.val, S.col, S.val);
    count_launch();
    DB_CUDA(cudaGetLastError());
    S.general_released = false;
    return 0;
}

int refill_sell_values(const CsrDev& A, SellOp& S)
{
    if (S.padded == 0) return 0;
    cudaStream_t st = ctx().stream;
    DB_CHECK(restore_general(A, S));
    k_sell_refill_values<<<blocks_for((long long)S.nslices * 32, 256), 256, 0, st>>>(
        (long long)S.nslices, S.slice_ptr, S.perm, A.rowptr, A.val, S.val);
    count_launch();
    DB_CUDA(cudaGetLastError());
    DB_CHECK(build_value_dictionary(S, nullptr));      // new values, new dictionary
    DB_CHECK(build_row_ordered_columns(S, nullptr));    // (the packed twin may have come or gone with them)
    release_general(S, nullptr);
    return 0;
}

int extract_diag(drops_b200_matrix_s* M)
{
    cudaStream_t st = ctx().stream;
    M->diag_ok = false;
    if (!M->square) return 0;
    if (!M->diag) DB_CHECK(dev_alloc(&M->diag, M->csr.rows));
    if (!M->diag_missing) DB_CHECK(dev_alloc(&M->diag_missing, 1));
    DB_CUDA(cudaMemsetAsync(M->diag_missing, 0, sizeof(int), st));
    if (M->csr.rows > 0) {
        k_extract_diag<<<blocks_for((long long)M->csr.rows, 256), 256, 0, st>>>(
            (long long)M->csr.rows, M->csr.rowptr, M->csr.col, M->csr.val, M->diag, M->diag_missing);
        count_launch();
        DB_CUDA(cudaGetLastError());
    }
    int missing = 0;
    DB_CUDA(cudaMemcpyAsync(&missing, M->diag_missing, sizeof(int), cudaMemcpyDeviceToHost, st));
    DB_CUDA(cudaStreamSynchronize(st));
    M->diag_ok = (missing == 0);
    return 0;
}

int build_transpose(drops_b200_matrix_s* M)
{
    if (M->has_t) return 0;
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    const CsrDev& A = M->csr;
    CsrDev& T = M->csr_t;
    free_csr(T);
    T.rows = A.cols; T.cols = A.rows; T.nnz = A.nnz;
    DB_ARG(A.rows <= (size_t)std::numeric_limits<int32_t>::max(), "too many rows for int32 column index of A^T");
    DB_CHECK(dev_alloc(&T.rowptr, T.rows + 1));
    DB_CHECK(dev_alloc(&T.col, T.nnz));
    DB_CHECK(dev_alloc(&T.val, T.nnz));
    unsigned long long* cursor = nullptr;
    DB_CHECK(dev_alloc(&cursor, T.rows + 1));
    DB_CUDA(cudaMemsetAsync(cursor, 0, (T.rows + 1) * sizeof(unsigned long long), st));
    if (A.nnz > 0) {
        k_count_cols<<<blocks_for((long long)A.nnz, 256), 256, 0, st>>>((long long)A.nnz, A.col, cursor);
        count_launch();
    }
    DB_CUDA(cudaMemcpyAsync(T.rowptr, cursor, (T.rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    DB_CHECK(exclusive_scan_i64(T.rowptr, T.rows + 1, st));
    DB_CUDA(cudaMemsetAsync(cursor, 0, (T.rows + 1) * sizeof(unsigned long long), st));
    if (A.rows > 0 && A.nnz > 0) {
        k_transpose_fill<<<blocks_for((long long)A.rows, 256), 256, 0, st>>>(
            (long long)A.rows, A.rowptr, A.col, A.val, T.rowptr, cursor, T.col, T.val);
        k_sort_rows<<<blocks_for((long long)T.rows, 256), 256, 0, st>>>((long long)T.rows, T.rowptr, T.col, T.val);
        count_launch(2);
    }
    DB_CUDA(cudaGetLastError());
    DB_CUDA(cudaStreamSynchronize(st));
    dev_free(cursor);
    size_t bytes = T.nnz * 12 + (T.rows + 1) * 8;
    DB_CHECK(build_sell(T, M->sell_t, &bytes, nullptr, 0));
    release_general(M->sell_t, &bytes);
    if (M->sell_t.general_released) {                  // the CSR form of A^T only existed to build the operator
        bytes -= std::min(bytes, T.nnz * 12 + (T.rows + 1) * 8);
        const size_t r = T.rows, cc = T.cols, z = T.nnz;
        free_csr(T);
        T.rows = r; T.cols = cc; T.nnz = z;
    }
    M->device_bytes += bytes;
    M->has_t = true;
    return 0;
}

void drop_transpose(drops_b200_matrix_s* M)
{
    if (!M->has_t) return;
    free_csr(M->csr_t);
    free_sell(M->sell_t);
    M->has_t = false;
}

// Finishes a matrix whose csr member is filled.
int finalize_matrix(drops_b200_matrix_s* M)
{
    M->square = (M->csr.rows == M->csr.cols);
    M->device_bytes = M->csr.nnz * 12 + (M->csr.rows + 1) * 8;
    DB_CHECK(extract_diag(M));
    if (M->square) M->device_bytes += M->csr.rows * 8;
    // rows in compact blocks of the matrix graph (reorder.cu) when the operator is large enough to profit
    int32_t* order = nullptr;
    bool packs = false;
    DB_CHECK(will_pack_one_word(M->csr.val, M->csr.nnz, std::max(M->csr.cols, M->csr.rows + 4096), &packs));
    // (every square operator gets a row-ordered twin, packed or not: graph blocks pay together with it)
    M->line_segments = 0;
    if (M->square && opt("sell.reorder", 1.0) != 0.0) DB_CHECK(compute_block_order(M->csr, &order, &M->line_segments));
    const int s = build_sell(M->csr, M->sell, &M->device_bytes, order, order ? M->csr.rows : 0, order != nullptr, packs);
    dev_free(order);
    M->block_ordered = s == 0 && order != nullptr;
    if (s == 0) release_general(M->sell, &M->device_bytes);
    return s;
}

}  // namespace db200

using namespace db200;

extern "C" {

int drops_b200_matrix_create(size_t rows, size_t cols, size_t nnz, const size_t* rowbeg,
                             const size_t* colind, const double* val, uint64_t version,
                             drops_b200_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr, "null out");
    DB_ARG(rowbeg != nullptr, "null rowbeg");
    DB_ARG(nnz == 0 || (colind != nullptr && val != nullptr), "null colind/val");
    *out = nullptr;
    if (rows > (size_t)std::numeric_limits<int32_t>::max() - 4096 ||
        cols > (size_t)std::numeric_limits<int32_t>::max()) {
        set_error("matrix %zux%zu exceeds the int32 row/column index of one device partition", rows, cols);
        return DROPS_B200_E_RANGE;
    }
    DB_ARG(rowbeg[rows] == nnz && rowbeg[0] == 0, "rowbeg[0] != 0 or rowbeg[rows] != nnz");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    drops_b200_matrix_s* M = new drops_b200_matrix_s;
    M->version = version;
    CsrDev& A = M->csr;
    A.rows = rows; A.cols = cols; A.nnz = nnz;
    int s;
    if ((s = dev_alloc(&A.rowptr, rows + 1)) < 0 || (s = dev_alloc(&A.col, nnz)) < 0 ||
        (s = dev_alloc(&A.val, nnz)) < 0) {
        drops_b200_matrix_destroy(M);
        return s;
    }
    static_assert(sizeof(size_t) == sizeof(int64_t), "size_t must be 64 bit");
    cudaError_t e = cudaMemcpyAsync(A.rowptr, rowbeg, (rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nnz > 0)
        e = cudaMemcpyAsync(A.val, val, nnz * sizeof(double), cudaMemcpyHostToDevice, st);
    // narrow the column indices through a pinned double buffer
    const size_t chunk = 8u << 20;
    int32_t* stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool bad_col = false;
    if (e == cudaSuccess && nnz > 0) {
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaMallocHost((void**)&stage[b], std::min(chunk, nnz) * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
        }
        int b = 0;
        for (size_t off = 0; off < nnz && e == cudaSuccess; off += chunk, b ^= 1) {
            const size_t m = std::min(chunk, nnz - off);
            e = cudaEventSynchronize(ev[b]);
            if (e != cudaSuccess) break;
            int32_t* dst = stage[b];
            const size_t* src = colind + off;
            int bad = 0;
#pragma omp parallel for reduction(| : bad)
            for (long long k = 0; k < (long long)m; ++k) {
                bad |= (src[k] >= cols);
                dst[k] = (int32_t)src[k];
            }
            bad_col |= (bad != 0);
            e = cudaMemcpyAsync(A.col + off, dst, m * sizeof(int32_t), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    for (int b = 0; b < 2; ++b) {
        if (stage[b]) cudaFreeHost(stage[b]);
        if (ev[b]) cudaEventDestroy(ev[b]);
    }
    if (e != cudaSuccess) {
        drops_b200_matrix_destroy(M);
        return cuda_fail(e, "matrix upload", __FILE__, __LINE__);
    }
    if (bad_col) {
        drops_b200_matrix_destroy(M);
        set_error("drops_b200_matrix_create: a column index is >= cols");
        return DROPS_B200_E_ARG;
    }
    s = finalize_matrix(M);
    if (s < 0) { drops_b200_matrix_destroy(M); return s; }
    *out = M;
    return 0;
}

int drops_b200_matrix_create_device(size_t rows, size_t cols, size_t nnz, const int64_t* d_rowptr,
                                    const int32_t* d_colind, const double* d_val, uint64_t version,
                                    drops_b200_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr && d_rowptr != nullptr, "null argument");
    DB_ARG(nnz == 0 || (d_colind != nullptr && d_val != nullptr), "null colind/val");
    *out = nullptr;
    if (rows > (size_t)std::numeric_limits<int32_t>::max() - 4096 ||
        cols > (size_t)std::numeric_limits<int32_t>::max()) {
        set_error("matrix %zux%zu exceeds the int32 row/column index of one device partition", rows, cols);
        return DROPS_B200_E_RANGE;
    }
    cudaStream_t st = ctx().stream;
    drops_b200_matrix_s* M = new drops_b200_matrix_s;
    M->version = version;
    CsrDev& A = M->csr;
    A.rows = rows; A.cols = cols; A.nnz = nnz;
    int s;
    if ((s = dev_alloc(&A.rowptr, rows + 1)) < 0 || (s = dev_alloc(&A.col, nnz)) < 0 ||
        (s = dev_alloc(&A.val, nnz)) < 0) {
        drops_b200_matrix_destroy(M);
        return s;
    }
    cudaError_t e = cudaMemcpyAsync(A.rowptr, d_rowptr, (rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(A.col, d_colind, nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(A.val, d_val, nnz * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        drops_b200_matrix_destroy(M);
        return cuda_fail(e, "device csr copy", __FILE__, __LINE__);
    }
    s = finalize_matrix(M);
    if (s < 0) { drops_b200_matrix_destroy(M); return s; }
    *out = M;
    return 0;
}

int drops_b200_matrix_lincomb(double a, drops_b200_matrix_t A, double b, drops_b200_matrix_t B, uint64_t version,
                              drops_b200_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(A != nullptr && B != nullptr && out != nullptr, "null argument");
    DB_ARG(A->csr.rows == B->csr.rows && A->csr.cols == B->csr.cols, "LinComb: incompatible dimensions");
    *out = nullptr;
    cudaStream_t st = ctx().stream;
    const size_t rows = A->csr.rows;
    drops_b200_matrix_s* M = new drops_b200_matrix_s;
    M->version = version;
    CsrDev& C = M->csr;
    C.rows = rows; C.cols = A->csr.cols;
    int s = dev_alloc(&C.rowptr, rows + 1);
    if (s < 0) { drops_b200_matrix_destroy(M); return s; }
    cudaError_t e = cudaMemsetAsync(C.rowptr, 0, (rows + 1) * sizeof(int64_t), st);
    if (e == cudaSuccess && rows > 0) {
        k_lincomb_count<<<blocks_for((long long)rows, 256), 256, 0, st>>>((long long)rows, A->csr.rowptr, A->csr.col,
                                                                         B->csr.rowptr, B->csr.col, C.rowptr);
        count_launch();
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { drops_b200_matrix_destroy(M); return cuda_fail(e, "lincomb count", __FILE__, __LINE__); }
    s = exclusive_scan_i64(C.rowptr, rows + 1, st);
    int64_t nnz = 0;
    if (s == 0) {
        e = cudaMemcpyAsync(&nnz, C.rowptr + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) s = cuda_fail(e, "lincomb nnz", __FILE__, __LINE__);
    }
    if (s == 0) { C.nnz = (size_t)nnz; s = dev_alloc(&C.col, C.nnz); }
    if (s == 0) s = dev_alloc(&C.val, C.nnz);
    if (s == 0 && rows > 0 && C.nnz > 0) {
        k_lincomb_fill<<<blocks_for((long long)rows, 256), 256, 0, st>>>(
            (long long)rows, a, A->csr.rowptr, A->csr.col, A->csr.val, b, B->csr.rowptr, B->csr.col, B->csr.val,
            C.rowptr, C.col, C.val);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) s = cuda_fail(e, "lincomb fill", __FILE__, __LINE__);
    }
    if (s == 0) s = finalize_matrix(M);
    if (s < 0) { drops_b200_matrix_destroy(M); return s; }
    *out = M;
    return 0;
}

int drops_b200_matrix_update_values(drops_b200_matrix_t M, const double* val, uint64_t version)
{
    DB_CHECK(require_ctx());
    DB_ARG(M != nullptr && (val != nullptr || M->csr.nnz == 0), "null argument");
    cudaStream_t st = ctx().stream;
    if (M->csr.nnz)
        DB_CUDA(cudaMemcpyAsync(M->csr.val, val, M->csr.nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    DB_CHECK(refill_sell_values(M->csr, M->sell));
    DB_CHECK(extract_diag(M));
    drop_transpose(M);
    M->version = version;
    ++M->values_epoch;
    DB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int drops_b200_matrix_destroy(drops_b200_matrix_t M)
{
    if (!M) return 0;
    if (ctx().inited) cudaStreamSynchronize(ctx().stream);
    free_csr(M->csr); free_sell(M->sell);
    free_csr(M->csr_t); free_sell(M->sell_t);
    dev_free(M->diag); dev_free(M->diag_missing); dev_free(M->bbt_diag); dev_free(M->diag_pc);
    for (double* w : M->work) dev_free(w);
    for (double* w : M->ro_work) dev_free(w);
    dev_free(M->ro_diag_idx);
    dev_free(M->bridge[0]); dev_free(M->bridge[1]);
    dev_free(M->d_scalars); dev_free(M->d_hist); dev_free(M->d_small);
    if (M->pcg_graph) cudaGraphExecDestroy(M->pcg_graph);
    delete M;
    return 0;
}

int drops_b200_matrix_info(drops_b200_matrix_t M, size_t* rows, size_t* cols, size_t* nnz,
                           uint64_t* version, size_t* device_bytes, size_t* padded_nnz)
{
    DB_ARG(M != nullptr, "null matrix");
    if (rows) *rows = M->csr.rows;
    if (cols) *cols = M->csr.cols;
    if (nnz) *nnz = M->csr.nnz;
    if (version) *version = M->version;
    if (device_bytes) *device_bytes = M->device_bytes;
    if (padded_nnz) *padded_nnz = M->sell.padded;
    return 0;
}

int drops_b200_matrix_value_dictionary(drops_b200_matrix_t M, int transposed, int* n_distinct, int* packed,
                                       int* column_bits)
{
    DB_ARG(M != nullptr, "null matrix");
    const SellOp& S = transposed ? M->sell_t : M->sell;
    if (n_distinct) *n_distinct = S.ndict;
    if (packed) *packed = S.pk != nullptr ? 1 : (S.vidx != nullptr ? 2 : 0);
    if (column_bits) *column_bits = S.cbits;
    return 0;
}

int drops_b200_matrix_layout(drops_b200_matrix_t M, int* block_ordered, int* row_ordered_stream)
{
    DB_ARG(M != nullptr, "null matrix");
    if (block_ordered) *block_ordered = M->block_ordered;
    if (row_ordered_stream) *row_ordered_stream = M->sell.pkp != nullptr ? 1 : (M->sell.colp != nullptr ? 2 : 0);
    return 0;
}

int drops_b200_matrix_line_segments(drops_b200_matrix_t M, size_t* n_segments)
{
    DB_ARG(M != nullptr && n_segments != nullptr, "null argument");
    *n_segments = M->block_ordered ? M->line_segments : 0;
    return 0;
}

int drops_b200_matrix_row_order(drops_b200_matrix_t M, int32_t* perm, size_t capacity)
{
    DB_CHECK(require_ctx());
    DB_ARG(M != nullptr && perm != nullptr, "null argument");
    const size_t npos = M->sell.nslices * 32;
    DB_ARG(capacity >= npos, "row order: buffer too small (32 * nslices entries)");
    if (npos) DB_CUDA(cudaMemcpy(perm, M->sell.perm, npos * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int drops_b200_matrix_segment_stream(drops_b200_matrix_t M, size_t* n_slices, size_t* n_entries)
{
    DB_ARG(M != nullptr, "null matrix");
    if (n_slices) *n_slices = M->sell.scd != nullptr ? M->sell.sc_slices : 0;
    if (n_entries) *n_entries = M->sell.scd != nullptr ? M->sell.sc_entries : 0;
    return 0;
}

int drops_b200_matrix_diag_index(drops_b200_matrix_t M, int* n_values)
{
    DB_ARG(M != nullptr && n_values != nullptr, "null argument");
    *n_values = (M->ro_diag_idx != nullptr && M->ro_diag_epoch == M->values_epoch && opt("pcg.diag_index", 1.0) != 0.0) ? M->ro_diag_count : 0;
    return 0;
}

int drops_b200_matrix_format(drops_b200_matrix_t M, size_t* nslices, size_t* padded_nnz)
{
    DB_ARG(M != nullptr, "null matrix");
    if (nslices) *nslices = M->sell.nslices;
    if (padded_nnz) *padded_nnz = M->sell.padded;
    return 0;
}

int drops_b200_matrix_download(drops_b200_matrix_t M, size_t* rowbeg, size_t* colind, double* val)
{
    DB_CHECK(require_ctx());
    DB_ARG(M != nullptr && rowbeg != nullptr, "null argument");
    cudaStream_t st = ctx().stream;
    const CsrDev& A = M->csr;
    DB_CUDA(cudaMemcpyAsync(rowbeg, A.rowptr, (A.rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (A.nnz) {
        DB_ARG(colind != nullptr && val != nullptr, "null colind/val");
        DB_CUDA(cudaMemcpyAsync(val, A.val, A.nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
        // widen the columns on the device in chunks, reusing a bounce buffer
        const size_t chunk = 32u << 20;
        unsigned long long* wide = nullptr;
        DB_CHECK(dev_alloc(&wide, std::min(chunk, A.nnz)));
        for (size_t off = 0; off < A.nnz; off += chunk) {
            const size_t m = std::min(chunk, A.nnz - off);
            k_widen_i32<<<blocks_for((long long)m, 256), 256, 0, st>>>((long long)m, A.col + off, wide);
            count_launch();
            cudaError_t e = cudaMemcpyAsync(colind + off, wide, m * sizeof(size_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { dev_free(wide); return cuda_fail(e, "download", __FILE__, __LINE__); }
        }
        dev_free(wide);
    }
    DB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int drops_b200_matrix_diag(drops_b200_matrix_t M, double* diag)
{
    DB_CHECK(require_ctx());
    DB_ARG(M != nullptr && diag != nullptr, "null argument");
    DB_ARG(M->square, "GetDiag needs a square matrix");
    cudaStream_t st = ctx().stream;
    DB_CUDA(cudaMemcpyAsync(diag, M->diag, M->csr.rows * sizeof(double), cudaMemcpyDeviceToHost, st));
    DB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

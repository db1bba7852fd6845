// reorder.cu -- a locality order of the rows of a square operator, computed from its graph alone.
//
// Why: with the packed 4-byte stream the SpMV is no longer bound by HBM but by the L2 -> SM path
// (ncu, profiles/r01_notes.md: 14.8 GB leave L2 per launch at C2, 9 GB of them x gathers, against
// 6.8 GB of DRAM traffic; ~85 % of the ~6300 B/clk the L2 slices deliver).  Every x entry is used by
// ~29 rows, but in the caller's numbering the rows that share it are a mesh line or a mesh plane
// apart, i.e. in other CTAs on other SMs, so L1 serves only a quarter of the gathers.  Processing
// the rows in compact 3-D blocks of the mesh lets a CTA find its neighbours' x entries in its own L1.
//
// DROPS hands over matrices, not meshes, so the blocks come from the matrix graph: three
// breadth-first distance fields from three far-apart rows (a = farthest from row 0, b = farthest
// from a, c = the row maximising d_a + d_b) are pseudo-coordinates; rows are sorted by the Morton
// code of their cell (d_a / Ta, d_b / Tb, d_c / Tc).  The sliced-ELL windows are then cut from this order (matrix.cu,
// build_sell with a row map), and the row-ordered twin stream numbers x the same way.
//
// Line segments ("sell.line_order", round 2): what bounds the packed product after the blocks is the L1 data pipe --
// one wavefront per 128-byte line a 32-lane gather touches, 4.8 lines per gather with blocks of 8^3 (runs of 8
// rows), floor 2 (profiles/r02_*).  Stencil matrices are translation invariant along the fastest direction of their
// numbering: row r + s repeats row r with every column moved up by a small amount.  Such rows are chained
// (k_line_next: same length, all column differences in [1, D], smallest s <= S), chains are ranked by pointer
// jumping and cut into segments of 32; a complete segment is ONE slice, the unit the blocks are built from.  Entry j
// of its 32 lanes is then a segment of another chain -- 32 consecutive positions of the row-ordered x, 2 to 3 lines --
// and the rows left over (chain tails, boundary rows, graphs without lines) follow in the old per-row block order.  Any permutation is valid -- the order only decides who shares an L1 -- so graphs
// without mesh structure, several components or unreachable rows cost nothing but the setup pass.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>

#include <algorithm>

namespace db200 {

// One warp per frontier row: unvisited neighbours get level+1 and join the next frontier.
__global__ void __launch_bounds__(256)
k_bfs_expand(int nf, const int32_t* __restrict__ frontier, const int64_t* __restrict__ rowptr,
             const int32_t* __restrict__ col, int32_t* dist, int level, int32_t* next, int* next_count, int n)
{
    const int lane = threadIdx.x & 31;
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (w >= nf) return;
    const int v = frontier[w];
    const int64_t b = rowptr[v], e = rowptr[v + 1];
    for (int64_t k0 = b; k0 < e; k0 += 32) {
        const int64_t k = k0 + lane;
        bool found = false;
        int u = -1;
        if (k < e) {
            u = col[k];
            if (u >= 0 && u < n && dist[u] < 0) found = atomicCAS(&dist[u], -1, level + 1) == -1;
        }
        const unsigned m = __ballot_sync(0xffffffffu, found);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(next_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (found) next[base + __popc(m & ((1u << lane) - 1u))] = u;
        }
    }
}

__global__ void k_fill_i32(long long n, int32_t* p, int32_t v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_sum_dist(long long n, const int32_t* __restrict__ a, const int32_t* __restrict__ b, int32_t* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (a[i] < 0 || b[i] < 0) ? -1 : a[i] + b[i];
}

__device__ __forceinline__ unsigned spread3(unsigned v)     // 10 bits -> every third bit
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Morton code of the cell (d_a / ta, d_b / tb, d_c / tc) of a row (30 bits); rows the searches did not reach sort last.
__device__ __forceinline__ unsigned cell_code(int a, int b, int c, int ta, int tb, int tc)
{
    if (a < 0 || b < 0 || c < 0) return 0x3fffffffu;
    const unsigned ca = min(a / ta, 1022), cb = min(b / tb, 1022), cc = min(c / tc, 1022);
    return spread3(ca) | (spread3(cb) << 1) | (spread3(cc) << 2);
}

// Sort key of a row without line segments: its cell; the stable sort keeps ascending row index inside a cell.
__global__ void k_block_keys(long long n, const int32_t* __restrict__ da, const int32_t* __restrict__ db,
                             const int32_t* __restrict__ dc, int ta, int tb, int tc,
                             unsigned long long* __restrict__ key, int32_t* __restrict__ row)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    row[i] = (int32_t)i;
    key[i] = cell_code(da[i], db[i], dc[i], ta, tb, tc);
}

// ---- line segments --------------------------------------------------------------------------------------------
// nxt[r] = the first r' in (r, r + S] whose row repeats row r translated: same length and every column difference in
// [1, D].  Two rows may claim the same successor; prv keeps the nearest one (atomicMax) and k_line_validate drops the other link.
__global__ void k_line_next(int n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int S, int D,
                            int32_t* __restrict__ nxt, int32_t* __restrict__ prv)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int64_t b = rowptr[r];
    const int len = (int)(rowptr[r + 1] - b);
    int found = -1;
    if (len > 0) {
        for (int s = 1; s <= S && r + s < n && found < 0; ++s) {
            const int64_t b2 = rowptr[r + s];
            if ((int)(rowptr[r + s + 1] - b2) != len) continue;
            // a translate moves every column by the stride of that column's own line: on a mesh with k kinds of lines the
            // differences take at most a few distinct values (1 for vertices and 2 or 7 for edges on the P2 bricks); a row of
            // ANOTHER stencil class that happens to sit next to r (DROPS numbers the edges cell by cell) has all kinds of
            // differences -- at most two distinct values are accepted
            bool ok = true;
            int da = 0, db = 0;
            for (int j = 0; j < len && ok; ++j) {
                const int d = col[b2 + j] - col[b + j];
                ok = d >= 1 && d <= D;
                if (d != da && d != db) {
                    if (da == 0) da = d;
                    else if (db == 0) db = d;
                    else ok = false;
                }
            }
            if (ok) found = r + s;
        }
    }
    nxt[r] = found;
    if (found >= 0) atomicMax(&prv[found], r);
}

// Keeps nxt[r] only where r is the predecessor its successor kept.
__global__ void k_line_validate(int n, int32_t* __restrict__ nxt, const int32_t* __restrict__ prv)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int f = nxt[r];
    if (f >= 0 && prv[f] != r) nxt[r] = -1;
}

// A line ends where the stride changes: DROPS leaves the eliminated boundary unknowns out of the numbering, so the last
// vertex of a mesh line finds the first vertex of the NEXT line three rows on and would carry the chain across the
// wrap (segments with a jump in every column).  cut[r] = the link r -> nxt[r] goes; decided from the links as they are.
__global__ void k_line_cut(int n, const int32_t* __restrict__ nxt, const int32_t* __restrict__ prv, uint8_t* __restrict__ cut)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int f = nxt[r], p = prv[r];
    cut[r] = (f >= 0 && p >= 0 && nxt[p] == r && (r - p) != (f - r)) ? 1 : 0;
}

// Applies the cuts; anc/dist start the ranking: a head is its own ancestor at distance 0, every other row points at its
// predecessor at distance 1.
__global__ void k_line_apply(int n, int32_t* __restrict__ nxt, int32_t* __restrict__ prv, const uint8_t* __restrict__ cut,
                             int32_t* __restrict__ anc, int32_t* __restrict__ dist)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int p = prv[r];
    if (p >= 0 && (nxt[p] != r || cut[p])) p = -1;        // (prv may still name a claimant that lost)
    // the links are rewritten by their owners only (nxt[r] by r, prv[r] by r); a reader that sees nxt[p] already reset
    // comes to the same conclusion as one that sees cut[p]
    if (cut[r]) nxt[r] = -1;
    prv[r] = p;
    anc[r] = p >= 0 ? p : r;
    dist[r] = p >= 0 ? 1 : 0;
}

// One round of pointer jumping (Wyllie): dist = steps to anc; *again is raised while some ancestor is not a head.
__global__ void k_line_jump(int n, const int32_t* __restrict__ anc, const int32_t* __restrict__ dist,
                            int32_t* __restrict__ anc2, int32_t* __restrict__ dist2, int* __restrict__ again)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int a = anc[r];
    const int aa = anc[a];
    anc2[r] = aa;
    dist2[r] = dist[r] + dist[a];
    if (anc[aa] != aa) *again = 1;
}

// The last row of a chain records the chain's length with its head.
__global__ void k_line_length(int n, const int32_t* __restrict__ nxt, const int32_t* __restrict__ anc,
                              const int32_t* __restrict__ dist, int32_t* __restrict__ clen)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (nxt[r] < 0) clen[anc[r]] = dist[r] + 1;
}

// Keys of the two stable sorts.  keyA puts the 32 rows of a complete segment side by side, in chain order (first row
// of the segment, rank inside it); keyB is the cell of the segment's first row.  Rows outside complete segments keep
// their own index and cell and sort after all segments (bit 30 of keyB).
__global__ void k_line_keys(int n, const int32_t* __restrict__ prv, const int32_t* __restrict__ anc,
                            const int32_t* __restrict__ dist, const int32_t* __restrict__ clen,
                            const int32_t* __restrict__ da, const int32_t* __restrict__ db, const int32_t* __restrict__ dc,
                            int ta, int tb, int tc, unsigned long long* __restrict__ keyA, unsigned* __restrict__ keyB,
                            int32_t* __restrict__ row, unsigned long long* __restrict__ nfull)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    row[r] = r;
    const int rank = dist[r];
    const bool full = ((rank >> 5) + 1) * 32 <= clen[anc[r]];
    // the tail of a chain (fewer than 32 rows) stays together in chain order as well: columns that point into it are then
    // one run of positions, although the tail shares its slices with other tails (DROPS numbers edges cell by cell, so in
    // ascending row order the tails of seven edge classes would interleave)
    int first = r;
    for (int t = rank & 31; t > 0; --t) first = prv[first];
    keyA[r] = ((unsigned long long)(unsigned)first << 5) | (unsigned long long)(rank & 31);
    keyB[r] = cell_code(da[first], db[first], dc[first], ta, tb, tc) | (full ? 0u : (1u << 30));
    if (full && (rank & 31) == 0) atomicAdd(nfull, 1ull);
}

__global__ void k_widen_u32(long long n, const unsigned* __restrict__ src, unsigned long long* __restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

__global__ void k_gather_u32(long long n, const unsigned* __restrict__ src, const int32_t* __restrict__ idx, unsigned* __restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

static inline int blocks_of(long long n, int t) { return (int)std::max<long long>(1, (n + t - 1) / t); }

// dist = breadth-first distance from src (-1 where unreachable)
static int bfs(const CsrDev& A, int src, int32_t* dist, int32_t* q0, int32_t* q1, int* d_count, cudaStream_t st)
{
    const int n = (int)A.rows;
    k_fill_i32<<<blocks_of(n, 256), 256, 0, st>>>(n, dist, -1);
    const int32_t zero = 0, s32 = src;
    DB_CUDA(cudaMemcpyAsync(dist + src, &zero, 4, cudaMemcpyHostToDevice, st));
    DB_CUDA(cudaMemcpyAsync(q0, &s32, 4, cudaMemcpyHostToDevice, st));
    int nf = 1, level = 0;
    count_launch();
    while (nf > 0) {
        DB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
        k_bfs_expand<<<blocks_of((long long)nf * 32, 256), 256, 0, st>>>(nf, q0, A.rowptr, A.col, dist, level, q1, d_count, n);
        count_launch();
        DB_CUDA(cudaMemcpyAsync(&nf, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        DB_CUDA(cudaStreamSynchronize(st));
        std::swap(q0, q1);
        ++level;
    }
    return 0;
}

static int argmax_i32(const int32_t* d, int n, int* out, void** tmp, size_t* tmp_bytes, cub::KeyValuePair<int, int32_t>* d_kv,
                      cudaStream_t st)
{
    size_t need = 0;
    DB_CUDA(cub::DeviceReduce::ArgMax(nullptr, need, d, d_kv, n, st));
    if (need > *tmp_bytes) {
        dev_free(*tmp); *tmp = nullptr; *tmp_bytes = 0;
        char* p = nullptr;
        DB_CHECK(dev_alloc(&p, need));
        *tmp = p; *tmp_bytes = need;
    }
    DB_CUDA(cub::DeviceReduce::ArgMax(*tmp, need, d, d_kv, n, st));
    cub::KeyValuePair<int, int32_t> h;
    DB_CUDA(cudaMemcpyAsync(&h, d_kv, sizeof h, cudaMemcpyDeviceToHost, st));
    DB_CUDA(cudaStreamSynchronize(st));
    count_launch();
    *out = h.key;
    return 0;
}

// *d_order (rows entries, caller frees) lists the rows in block order; nullptr when the option is off,
// the operator has more rows than columns or is too small to profit ("sell.block_min_rows").
int compute_block_order(const CsrDev& A, int32_t** d_order, size_t* n_segments)
{
    *d_order = nullptr;
    // columns >= rows (the ghost columns of a partition) are no graph vertices: k_bfs_expand skips them
    if (opt("sell.block_order", 1.0) == 0.0 || A.rows > A.cols || A.rows == 0 || A.nnz == 0) return 0;
    if ((double)A.rows < opt("sell.block_min_rows", 2.0e6)) return 0;
    int tile = (int)opt("sell.block_tile", 8.0);
    if (tile < 1) tile = 1;
    // per-field tiles ("sell.block_tile_a/b/c", 0 = block_tile): cells need not be cubes
    const int ta = std::max(1, (int)opt("sell.block_tile_a", 0.0) > 0 ? (int)opt("sell.block_tile_a", 0.0) : tile);
    const int tb = std::max(1, (int)opt("sell.block_tile_b", 0.0) > 0 ? (int)opt("sell.block_tile_b", 0.0) : tile);
    const int tc = std::max(1, (int)opt("sell.block_tile_c", 0.0) > 0 ? (int)opt("sell.block_tile_c", 0.0) : tile);
    cudaStream_t st = ctx().stream;
    const int n = (int)A.rows;
    int32_t *da = nullptr, *db = nullptr, *dc = nullptr, *q0 = nullptr, *q1 = nullptr, *rows_in = nullptr, *rows_out = nullptr;
    unsigned long long *key_in = nullptr, *key_out = nullptr;
    const bool lines = opt("sell.line_order", 1.0) != 0.0;
    if (n_segments) *n_segments = 0;
    int* d_count = nullptr;
    cub::KeyValuePair<int, int32_t>* d_kv = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int status = 0;
    auto cleanup = [&]() {
        dev_free(da); dev_free(db); dev_free(dc); dev_free(q0); dev_free(q1); dev_free(rows_in); dev_free(key_in);
        dev_free(key_out); dev_free(d_count); dev_free(d_kv); dev_free(tmp);
    };
#define RO_CHECK(call) do { status = (call); if (status < 0) { cleanup(); dev_free(rows_out); return status; } } while (0)
#define RO_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); dev_free(rows_out); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
    RO_CHECK(dev_alloc(&da, (size_t)n)); RO_CHECK(dev_alloc(&db, (size_t)n)); RO_CHECK(dev_alloc(&dc, (size_t)n));
    RO_CHECK(dev_alloc(&q0, (size_t)n)); RO_CHECK(dev_alloc(&q1, (size_t)n));
    RO_CHECK(dev_alloc(&d_count, 1)); RO_CHECK(dev_alloc(&d_kv, 1));
    int a = 0, b = 0, c = 0;
    RO_CHECK(bfs(A, 0, dc, q0, q1, d_count, st));                       // scratch field from row 0
    RO_CHECK(argmax_i32(dc, n, &a, &tmp, &tmp_bytes, d_kv, st));
    RO_CHECK(bfs(A, a, da, q0, q1, d_count, st));
    RO_CHECK(argmax_i32(da, n, &b, &tmp, &tmp_bytes, d_kv, st));
    RO_CHECK(bfs(A, b, db, q0, q1, d_count, st));
    k_sum_dist<<<blocks_of(n, 256), 256, 0, st>>>(n, da, db, dc);
    count_launch();
    RO_CHECK(argmax_i32(dc, n, &c, &tmp, &tmp_bytes, d_kv, st));
    RO_CHECK(bfs(A, c, dc, q0, q1, d_count, st));
    RO_CHECK(dev_alloc(&key_in, (size_t)n)); RO_CHECK(dev_alloc(&key_out, (size_t)n));
    RO_CHECK(dev_alloc(&rows_in, (size_t)n)); RO_CHECK(dev_alloc(&rows_out, (size_t)n));
    auto sort_pairs = [&](int bits) -> int {          // key_in/rows_in -> key_out/rows_out, stable
        size_t need = 0;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, key_in, key_out, rows_in, rows_out, n, 0, bits, st);
        if (e != cudaSuccess) return cuda_fail(e, "block order: sort", __FILE__, __LINE__);
        if (need > tmp_bytes) {
            dev_free(tmp); tmp = nullptr; tmp_bytes = 0;
            char* p = nullptr;
            const int r = dev_alloc(&p, need);
            if (r < 0) return r;
            tmp = p; tmp_bytes = need;
        }
        e = cub::DeviceRadixSort::SortPairs(tmp, need, key_in, key_out, rows_in, rows_out, n, 0, bits, st);
        if (e != cudaSuccess) return cuda_fail(e, "block order: sort", __FILE__, __LINE__);
        count_launch();
        return 0;
    };
    if (lines) {
        // q0 = nxt, q1 = prv, then the ranking ping-pongs between (anc, dist) pairs
        int32_t *anc[2] = {nullptr, nullptr}, *dst[2] = {nullptr, nullptr}, *clen = nullptr;
        unsigned *kb = nullptr, *kb2 = nullptr;
        unsigned long long* d_nfull = nullptr;
        auto lfree = [&]() { dev_free(anc[0]); dev_free(anc[1]); dev_free(dst[0]); dev_free(dst[1]); dev_free(clen); dev_free(kb);
                             dev_free(kb2); dev_free(d_nfull); };
#define LN_CHECK(call) do { status = (call); if (status < 0) { lfree(); cleanup(); dev_free(rows_out); return status; } } while (0)
#define LN_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { lfree(); cleanup(); dev_free(rows_out); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
        for (int i = 0; i < 2; ++i) { LN_CHECK(dev_alloc(&anc[i], (size_t)n)); LN_CHECK(dev_alloc(&dst[i], (size_t)n)); }
        LN_CHECK(dev_alloc(&clen, (size_t)n)); LN_CHECK(dev_alloc(&kb, (size_t)n)); LN_CHECK(dev_alloc(&kb2, (size_t)n));
        LN_CHECK(dev_alloc(&d_nfull, 1));
        const int S = std::max(1, (int)opt("sell.line_stride", 8.0)), D = std::max(1, (int)opt("sell.line_maxdiff", 64.0));
        const int g = blocks_of(n, 256);
        k_fill_i32<<<g, 256, 0, st>>>(n, q1, -1);
        k_line_next<<<g, 256, 0, st>>>(n, A.rowptr, A.col, S, D, q0, q1);
        k_line_validate<<<g, 256, 0, st>>>(n, q0, q1);
        // kb2 doubles as the cut flags (n bytes of its 4 n)
        k_line_cut<<<g, 256, 0, st>>>(n, q0, q1, reinterpret_cast<uint8_t*>(kb2));
        k_line_apply<<<g, 256, 0, st>>>(n, q0, q1, reinterpret_cast<const uint8_t*>(kb2), anc[0], dst[0]);
        count_launch(5);
        int cur = 0;
        for (int round = 0; round < 32; ++round) {
            int again = 0;
            LN_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
            k_line_jump<<<g, 256, 0, st>>>(n, anc[cur], dst[cur], anc[cur ^ 1], dst[cur ^ 1], d_count);
            count_launch();
            LN_CUDA(cudaMemcpyAsync(&again, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
            LN_CUDA(cudaStreamSynchronize(st));
            cur ^= 1;
            if (!again) break;
        }
        k_fill_i32<<<g, 256, 0, st>>>(n, clen, 0);
        k_line_length<<<g, 256, 0, st>>>(n, q0, anc[cur], dst[cur], clen);
        LN_CUDA(cudaMemsetAsync(d_nfull, 0, sizeof(unsigned long long), st));
        k_line_keys<<<g, 256, 0, st>>>(n, q1, anc[cur], dst[cur], clen, da, db, dc, ta, tb, tc, key_in, kb, rows_in, d_nfull);
        count_launch(); count_launch(); count_launch();
        LN_CHECK(sort_pairs(37));                                            // keyA: (first row of the segment, rank in it)
        k_gather_u32<<<g, 256, 0, st>>>(n, kb, rows_out, kb2);               // keyB in that order, widened for the sorter
        k_widen_u32<<<g, 256, 0, st>>>(n, kb2, key_in);
        count_launch(); count_launch();
        std::swap(rows_in, rows_out);
        LN_CHECK(sort_pairs(31));                                            // keyB: segments before single rows, by cell
        unsigned long long nfull = 0;
        LN_CUDA(cudaMemcpyAsync(&nfull, d_nfull, sizeof nfull, cudaMemcpyDeviceToHost, st));
        LN_CUDA(cudaStreamSynchronize(st));
        if (n_segments) *n_segments = (size_t)nfull;
        lfree();
#undef LN_CHECK
#undef LN_CUDA
    } else {
        k_block_keys<<<blocks_of(n, 256), 256, 0, st>>>(n, da, db, dc, ta, tb, tc, key_in, rows_in);
        count_launch();
        RO_CHECK(sort_pairs(30));                                            // stable: equal keys keep ascending row index
    }
    RO_CUDA(cudaStreamSynchronize(st));
#undef RO_CHECK
#undef RO_CUDA
    cleanup();
    *d_order = rows_out;
    return 0;
}

}  // namespace db200

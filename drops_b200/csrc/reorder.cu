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
// code of their cell (d_a / Ta, d_b / Tb, d_c / Tc) and, inside a cell, along the lines of the graph (k_block_keys).  The sliced-ELL windows are then cut
// from this order (matrix.cu, build_sell with a row map), and the row-ordered twin stream numbers x
// the same way.  Any permutation is valid -- the order only decides who shares an L1 -- so graphs
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

// Sort key of a row: the Morton code of its cell (d_a / ta, d_b / tb, d_c / tc) in the high bits and, below it,
// what orders the rows INSIDE a cell ("sell.block_inner"):
//   0 = nothing (the stable sort keeps ascending row index: runs follow the caller's numbering),
//   1 = (d_a, d_b, d_c) lexicographic, d_c fastest.  Neighbouring rows of the graph differ by at most 1 in every
//       field, and on a mesh-like graph each field is piecewise an affine coordinate, so rows that agree in two
//       fields and are consecutive in the third are a LINE of the mesh.  With a long cell in that direction
//       ("sell.block_tile_c" = 64) the sigma windows, which then sort by row length, cut slices that are 32
//       consecutive same-shape rows of one line: entry j of the 32 lanes gathers 32 consecutive positions of the
//       row-ordered x (2-3 cache lines instead of ~5: the L1 wavefronts per gather are what bounds the packed
//       product, profiles/r01_notes.md), and neighbouring lines reuse them from L1.
__global__ void k_block_keys(long long n, const int32_t* __restrict__ da, const int32_t* __restrict__ db,
                             const int32_t* __restrict__ dc, int ta, int tb, int tc, int inner,
                             unsigned long long* __restrict__ key, int32_t* __restrict__ row)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    row[i] = (int32_t)i;
    const int a = da[i], b = db[i], c = dc[i];
    if (a < 0 || b < 0 || c < 0) { key[i] = 0x3fffffffull << 30; return; }      // not reachable: after everything else
    const unsigned ca = min(a / ta, 1022), cb = min(b / tb, 1022), cc = min(c / tc, 1022);
    unsigned long long k = (unsigned long long)(spread3(ca) | (spread3(cb) << 1) | (spread3(cc) << 2)) << 30;
    if (inner == 1)
        k |= ((unsigned long long)min(a - (int)ca * ta, 1023) << 20) | ((unsigned long long)min(b - (int)cb * tb, 1023) << 10) |
             (unsigned long long)min(c - (int)cc * tc, 1023);
    key[i] = k;
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
int compute_block_order(const CsrDev& A, int32_t** d_order)
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
    const int inner = (int)opt("sell.block_inner", 1.0);
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
    dev_free(q0); q0 = nullptr; dev_free(q1); q1 = nullptr;
    RO_CHECK(dev_alloc(&key_in, (size_t)n)); RO_CHECK(dev_alloc(&key_out, (size_t)n));
    RO_CHECK(dev_alloc(&rows_in, (size_t)n)); RO_CHECK(dev_alloc(&rows_out, (size_t)n));
    k_block_keys<<<blocks_of(n, 256), 256, 0, st>>>(n, da, db, dc, ta, tb, tc, inner, key_in, rows_in);
    count_launch();
    size_t need = 0;
    RO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, key_in, key_out, rows_in, rows_out, n, 0, 60, st));
    if (need > tmp_bytes) {
        dev_free(tmp); tmp = nullptr; tmp_bytes = 0;
        char* p = nullptr;
        RO_CHECK(dev_alloc(&p, need));
        tmp = p; tmp_bytes = need;
    }
    RO_CUDA(cub::DeviceRadixSort::SortPairs(tmp, need, key_in, key_out, rows_in, rows_out, n, 0, 60, st));   // stable: equal keys keep ascending row index
    count_launch();
    RO_CUDA(cudaStreamSynchronize(st));
#undef RO_CHECK
#undef RO_CUDA
    cleanup();
    *d_order = rows_out;
    return 0;
}

}  // namespace db200

// brick.cu -- synthetic Poisson systems on the uniform Kuhn brick.
//
// Not part of the solver path: this is the deterministic workload generator that feeds the SAME
// CSR arrays to the oracle and to the CUDA path (the DROPS executables and their FE assembly
// cannot be built here: ParamCL needs Boost).  What it restates from the reference:
//   mesh     : BrickBuilderCL, geom/builder.cpp:151-178 -- every cell is cut into the 6 Kuhn
//              tetrahedra that share the main diagonal (vertex paths x-y-z, x-z-y, y-x-z, y-z-x,
//              z-x-y, z-y-x).
//   P1       : coup[i][j] = alpha * G_i.G_j / 6 * |det|          poisson/poissonaccus.tpp:327
//   P2       : coup[i][j] = alpha * QuadGrad(Grad,i,j) * |det| with the 5-point rule
//              (sum of the 4 vertex values)/120 + 2/15 * barycentre   poisson/poisson.tpp:828-838, :898-905
//   pattern  : every pair of DOFs sharing a tetrahedron is stored, zero or not      poisson.tpp:912-917
//   Dirichlet: boundary DOFs carry no unknown                                   misc/problem.h:1022-1042
//   numbering: 0 = lexicographic over the DOF lattice; 1 = vertices first, then edges
//              (misc/problem.cpp:707-719)
// On a uniform brick all interior rows of one lattice-parity class share one stencil, so the
// matrix is generated from a small table (8 classes x <= 65 offsets) -- on the host for oracle
// sized systems and by a kernel, directly into device CSR, for the 50 M / 400 M row systems.
// Host and device use the same table and the same index functions, so they agree bit for bit.
#include "dist.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

extern "C" int drops_b200_matrix_create_device(size_t rows, size_t cols, size_t nnz, const int64_t* d_rowptr,
                                               const int32_t* d_colind, const double* d_val,
                                               uint64_t version, drops_b200_matrix_t* out);

namespace db200 {

constexpr int kMaxStencil = 65;

struct StencilTable {
    int fe;                                   // 1 or 2
    int count[8];                             // entries per parity class (P1: class 0 only)
    signed char off[2][8][kMaxStencil][3];    // [numbering][class][entry] -> (dx,dy,dz), column-ascending
    double val[2][8][kMaxStencil];            // value for h = 1
};

// ---- table assembly (host) ------------------------------------------------------------------
static const int kPerm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
static const int kEdgeVerts[6][2] = {{0, 1}, {0, 2}, {1, 2}, {0, 3}, {1, 3}, {2, 3}};

static void tet_local_matrix(int fe, const double G[4][3], double coup[10][10])
{
    if (fe == 1) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                coup[i][j] = (G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2]) / 6.0;
        return;
    }
    // gradients of the 10 P2 shape functions at the 4 vertices and the barycentre
    double grad[10][5][3];
    for (int qp = 0; qp < 5; ++qp) {
        double lam[4];
        for (int k = 0; k < 4; ++k) lam[k] = qp < 4 ? (k == qp ? 1.0 : 0.0) : 0.25;
        for (int i = 0; i < 4; ++i)
            for (int d = 0; d < 3; ++d) grad[i][qp][d] = (4.0 * lam[i] - 1.0) * G[i][d];
        for (int e = 0; e < 6; ++e) {
            const int a = kEdgeVerts[e][0], b = kEdgeVerts[e][1];
            for (int d = 0; d < 3; ++d) grad[4 + e][qp][d] = 4.0 * (lam[a] * G[b][d] + lam[b] * G[a][d]);
        }
    }
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j) {
            double t[5];
            for (int qp = 0; qp < 5; ++qp)
                t[qp] = grad[i][qp][0] * grad[j][qp][0] + grad[i][qp][1] * grad[j][qp][1] + grad[i][qp][2] * grad[j][qp][2];
            coup[i][j] = (t[0] + t[1] + t[2] + t[3]) / 120. + 2. / 15. * t[4];
        }
}

static const StencilTable& stencil(int fe)
{
    static StencilTable tabs[2];
    static bool built[2] = {false, false};
    StencilTable& T = tabs[fe - 1];
    if (built[fe - 1]) return T;
    std::memset(&T, 0, sizeof T);
    T.fe = fe;
    const int nclass = fe == 2 ? 8 : 1;
    const int unit = fe == 2 ? 2 : 1;        // lattice steps per cell edge
    const int ndof = fe == 2 ? 10 : 4;
    for (int cl = 0; cl < nclass; ++cl) {
        const int P[3] = {unit + (cl & 1), unit + ((cl >> 1) & 1), unit + ((cl >> 2) & 1)};
        std::map<std::tuple<int, int, int>, double> row;   // keyed (dz,dy,dx): lexicographic column order
        for (int cz = 0; cz < 3; ++cz) for (int cy = 0; cy < 3; ++cy) for (int cx = 0; cx < 3; ++cx) {
            const int cell[3] = {cx, cy, cz};
            for (int t = 0; t < 6; ++t) {
                int v[4][3];
                for (int d = 0; d < 3; ++d) v[0][d] = cell[d];
                for (int s = 0; s < 3; ++s) {
                    for (int d = 0; d < 3; ++d) v[s + 1][d] = v[s][d];
                    v[s + 1][kPerm[t][s]] += 1;
                }
                // barycentric gradients for unit cells: -e_a, e_a - e_b, e_b - e_c, e_c
                double G[4][3] = {{0}};
                G[0][kPerm[t][0]] = -1.0;
                G[1][kPerm[t][0]] = 1.0; G[1][kPerm[t][1]] = -1.0;
                G[2][kPerm[t][1]] = 1.0; G[2][kPerm[t][2]] = -1.0;
                G[3][kPerm[t][2]] = 1.0;
                int pos[10][3];
                for (int k = 0; k < 4; ++k) for (int d = 0; d < 3; ++d) pos[k][d] = unit * v[k][d];
                if (fe == 2)
                    for (int e = 0; e < 6; ++e)
                        for (int d = 0; d < 3; ++d) pos[4 + e][d] = v[kEdgeVerts[e][0]][d] + v[kEdgeVerts[e][1]][d];
                int me = -1;
                for (int k = 0; k < ndof; ++k)
                    if (pos[k][0] == P[0] && pos[k][1] == P[1] && pos[k][2] == P[2]) me = k;
                if (me < 0) continue;
                double coup[10][10];
                tet_local_matrix(fe, G, coup);
                for (int k = 0; k < ndof; ++k)
                    row[std::make_tuple(pos[k][2] - P[2], pos[k][1] - P[1], pos[k][0] - P[0])] += coup[me][k];
            }
        }
        T.count[cl] = (int)row.size();
        // numbering 0: ascending (dz,dy,dx).  numbering 1: vertex columns first, then edge columns.
        int n0 = 0;
        for (auto& kv : row) {
            T.off[0][cl][n0][0] = (signed char)std::get<2>(kv.first);
            T.off[0][cl][n0][1] = (signed char)std::get<1>(kv.first);
            T.off[0][cl][n0][2] = (signed char)std::get<0>(kv.first);
            T.val[0][cl][n0] = kv.second;
            ++n0;
        }
        int n1 = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (int e = 0; e < n0; ++e) {
                bool vert = true;
                if (fe == 2)
                    for (int d = 0; d < 3; ++d) vert = vert && (((P[d] + T.off[0][cl][e][d]) & 1) == 0);
                if ((pass == 0) == vert) {
                    for (int d = 0; d < 3; ++d) T.off[1][cl][n1][d] = T.off[0][cl][e][d];
                    T.val[1][cl][n1] = T.val[0][cl][e];
                    ++n1;
                }
            }
    }
    built[fe - 1] = true;
    return T;
}

// ---- index functions shared by host and device ------------------------------------------------
struct BrickGeom {
    int fe, N, numbering;
    long long M;      // interior lattice points per direction
    long long nv1;    // N-1 (vertices per direction, P2)
    double h;         // cell size 1/N
    double upper_scale;
};

__host__ __device__ inline long long nat_index(const BrickGeom& g, long long i, long long j, long long k)
{
    return ((k - 1) * g.M + (j - 1)) * g.M + (i - 1);
}

__host__ __device__ inline long long dof_index(const BrickGeom& g, long long i, long long j, long long k)
{
    const long long nat = nat_index(g, i, j, k);
    if (g.fe == 1 || g.numbering == 0) return nat;
    const bool vert = ((i | j | k) & 1) == 0;
    if (vert) return ((k / 2 - 1) * g.nv1 + (j / 2 - 1)) * g.nv1 + (i / 2 - 1);
    long long before = ((k - 1) / 2) * g.nv1 * g.nv1;        // vertices in complete planes below
    if ((k & 1) == 0) {
        before += ((j - 1) / 2) * g.nv1;
        if ((j & 1) == 0) before += (i - 1) / 2;
    }
    return g.nv1 * g.nv1 * g.nv1 + nat - before;
}

__host__ __device__ inline int dof_class(const BrickGeom& g, long long i, long long j, long long k)
{
    return g.fe == 1 ? 0 : (int)((i & 1) | ((j & 1) << 1) | ((k & 1) << 2));
}

// Emits row (i,j,k): returns the number of stored entries; writes them when col/val != null.
template <typename IdxT>
__host__ __device__ inline int emit_row(const BrickGeom& g, const StencilTable& T, long long i, long long j,
                                        long long k, long long row, IdxT* col, double* val)
{
    const int cl = dof_class(g, i, j, k);
    const int nb = (g.fe == 2 && g.numbering == 1) ? 1 : 0;
    int cnt = 0;
    for (int e = 0; e < T.count[cl]; ++e) {
        const long long ii = i + T.off[nb][cl][e][0], jj = j + T.off[nb][cl][e][1], kk = k + T.off[nb][cl][e][2];
        if (ii < 1 || jj < 1 || kk < 1 || ii > g.M || jj > g.M || kk > g.M) continue;
        if (col) {
            const long long c = dof_index(g, ii, jj, kk);
            double v = T.val[nb][cl][e] * g.h;
            if (c > row) v = v * g.upper_scale;
            col[cnt] = (IdxT)c;
            val[cnt] = v;
        }
        ++cnt;
    }
    return cnt;
}

__host__ __device__ inline double xstar_at(const BrickGeom& g, long long i, long long j, long long k)
{
    const double hh = g.fe == 2 ? g.h * 0.5 : g.h;
    const double x = (double)i * hh, y = (double)j * hh, z = (double)k * hh;
#ifdef __CUDA_ARCH__
    double v = __dmul_rn(64.0, x);
    v = __dmul_rn(v, y); v = __dmul_rn(v, z);
    v = __dmul_rn(v, __dsub_rn(1.0, x)); v = __dmul_rn(v, __dsub_rn(1.0, y)); v = __dmul_rn(v, __dsub_rn(1.0, z));
    return v;
#else
    volatile double v = 64.0 * x;
    v = v * y; v = v * z;
    v = v * (1.0 - x); v = v * (1.0 - y); v = v * (1.0 - z);
    return v;
#endif
}

static int make_geom(int fe, int N, int numbering, double upper_scale, BrickGeom* g)
{
    DB_ARG(fe == 1 || fe == 2, "fe must be 1 (P1) or 2 (P2)");
    DB_ARG(N >= 2, "N must be >= 2");
    DB_ARG(numbering == 0 || numbering == 1, "numbering must be 0 or 1");
    g->fe = fe; g->N = N; g->numbering = numbering;
    g->M = fe == 2 ? 2LL * N - 1 : (long long)N - 1;
    g->nv1 = (long long)N - 1;
    g->h = 1.0 / (double)N;
    g->upper_scale = upper_scale;
    return 0;
}

// ---- device generation --------------------------------------------------------------------------
__global__ void k_brick_count(BrickGeom g, const StencilTable* __restrict__ T, long long n, int64_t* __restrict__ len)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long i = r % g.M + 1, j = (r / g.M) % g.M + 1, k = r / (g.M * g.M) + 1;
    const long long row = dof_index(g, i, j, k);
    len[row] = emit_row<int32_t>(g, *T, i, j, k, row, nullptr, nullptr);
}

__global__ void k_brick_fill(BrickGeom g, const StencilTable* __restrict__ T, long long n,
                             const int64_t* __restrict__ rowptr, int32_t* __restrict__ col, double* __restrict__ val)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long i = r % g.M + 1, j = (r / g.M) % g.M + 1, k = r / (g.M * g.M) + 1;
    const long long row = dof_index(g, i, j, k);
    const int64_t b = rowptr[row];
    emit_row<int32_t>(g, *T, i, j, k, row, col + b, val + b);
}

__global__ void k_brick_xstar(BrickGeom g, long long n, double* __restrict__ x)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long i = r % g.M + 1, j = (r / g.M) % g.M + 1, k = r / (g.M * g.M) + 1;
    x[dof_index(g, i, j, k)] = xstar_at(g, i, j, k);
}


// ---- z-slab partitions (multi-GPU) ---------------------------------------------------------------
// Rank r of `world` owns the lattice planes [k0, k1).  A cut face costs a rank time the rows do not show (its boundary rows
// run from the 12-byte stream beside the interior rows, halo pack and push, two more kernels in the chain), and the end
// slabs have one face where the others have two: with an even split they idle at every reduction (8 GPUs, measured
// inside the waiting kernels: 170 of 1955 us per iteration).  "brick.face_weight" = w counts a face as w planes: every
// rank gets the same planes + w * faces, i.e. the end slabs a few more planes (4 GPUs, w = 0 -> 6: waits of the end
// slabs 166 / 77 -> 20 / 32 us, 2088 -> 2134 C2-equivalent it/s).  w = 0: the even split.
static inline long long slab_begin(long long M, int world, int r, int fe)
{
    const long long w1000 = world > 2 ? (long long)(opt("brick.face_weight", 6.0) * 1000.0 + 0.5) : 0;
    if (r <= 0) return 1;
    if (w1000 <= 0) return 1 + M * r / world;
    const long long units = 1000 * M + w1000 * (2LL * world - 2);          // milli-planes incl. the faces
    if (units / world - 2 * w1000 < 2000) return 1 + M * r / world;        // slabs of a few planes: nothing to balance
    const long long before = (long long)r * units / world - w1000 * (2LL * r - 1);
    long long b = 1 + before / 1000;
    // P2: rows of even lattice planes reach two planes, rows of odd planes one, so a slab with an odd number of planes has
    // 2 boundary + 4 ghost planes or 4 + 2, alternating from rank to rank, and the ranks with 4 boundary planes are 130 us
    // per iteration behind (8 GPUs, 91-plane slabs, r02_notes.md section 12).  Cut above even planes only -- between
    // element layers, where DROPS' element-wise partition cuts -- and every inner slab has 3 + 3.
    if (fe == 2 && (b & 1) == 0) b += 1;
    return b;
}

static inline int slab_owner(long long M, int world, long long k, int fe)
{
    int r = (int)(((k - 1) * world) / M);          // first guess, then walk to the slab that holds k
    while (r > 0 && k < slab_begin(M, world, r, fe)) --r;
    while (r + 1 < world && k >= slab_begin(M, world, r + 1, fe)) ++r;
    return r;
}

// lattice plane k (1-based) of a global DOF id -- the inverse of dof_index as far as k goes
__host__ __device__ inline long long plane_of_gid(const BrickGeom& g, long long gid)
{
    if (g.fe == 1 || g.numbering == 0) return gid / (g.M * g.M) + 1;
    const long long nv3 = g.nv1 * g.nv1 * g.nv1;
    if (gid < nv3) return 2 * (gid / (g.nv1 * g.nv1) + 1);
    const long long e = gid - nv3;
    const long long pair = 2 * g.M * g.M - g.nv1 * g.nv1;   // edge DOFs in an (odd, even) plane pair
    const long long p = e / pair, rem = e % pair;
    return rem < g.M * g.M ? 2 * p + 1 : 2 * p + 2;
}

__global__ void k_slab_gids(BrickGeom g, long long r0, long long n, int64_t* __restrict__ gid, int64_t* __restrict__ lat)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long r = r0 + t;
    const long long i = r % g.M + 1, j = (r / g.M) % g.M + 1, k = r / (g.M * g.M) + 1;
    gid[t] = dof_index(g, i, j, k);
    lat[t] = r;
}

__global__ void k_slab_count(BrickGeom g, const StencilTable* __restrict__ T, long long n, const int64_t* __restrict__ lat,
                             int64_t* __restrict__ len)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long r = lat[t];
    const long long i = r % g.M + 1, j = (r / g.M) % g.M + 1, k = r / (g.M * g.M) + 1;
    len[t] = emit_row<int64_t>(g, *T, i, j, k, 0, nullptr, nullptr);
}

__global__ void k_slab_fill(BrickGeom g, const StencilTable* __restrict__ T, long long n, const int64_t* __restrict__ lat,
                            const int64_t* __restrict__ gid, const int64_t* __restrict__ rowptr,
                            int64_t* __restrict__ col, double* __restrict__ val)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long r = lat[t];
    const long long i = r % g.M + 1, j = (r / g.M) % g.M + 1, k = r / (g.M * g.M) + 1;
    const int64_t b = rowptr[t];
    emit_row<int64_t>(g, *T, i, j, k, gid[t], col + b, val + b);
}

__global__ void k_slab_xstar(BrickGeom g, long long n, const int64_t* __restrict__ lat, double* __restrict__ x)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long r = lat[t];
    x[t] = xstar_at(g, r % g.M + 1, (r / g.M) % g.M + 1, r / (g.M * g.M) + 1);
}

}  // namespace db200

struct drops_b200_brick_slab_s {
    db200::BrickGeom g;
    int world = 1, rank = 0;
    long long k0 = 0, k1 = 0;
    size_t n_owned = 0, nnz = 0;
    int64_t* d_gid = nullptr;      // owned global ids, ascending
    int64_t* d_lat = nullptr;      // lattice index of each owned row
    int64_t* d_rowptr = nullptr;
    int64_t* d_col = nullptr;      // global column ids
    double* d_val = nullptr;
    std::vector<int64_t> owned, ghost_gid;
    std::vector<int32_t> ghost_owner;
};

namespace db200 {

}  // namespace db200

using namespace db200;

extern "C" {

int drops_b200_brick_sizes(int fe, int N, size_t* n_out, size_t* nnz_out)
{
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, 0, 1.0, &g));
    const StencilTable& T = stencil(fe);
    const long long M = g.M;
    if (n_out) *n_out = (size_t)(M * M * M);
    if (nnz_out) {
        // separable count: per class and offset, the number of rows for which the neighbour is interior
        unsigned long long nnz = 0;
        const int nclass = fe == 2 ? 8 : 1;
        for (int cl = 0; cl < nclass; ++cl)
            for (int e = 0; e < T.count[cl]; ++e) {
                unsigned long long prod = 1;
                for (int d = 0; d < 3; ++d) {
                    const int o = T.off[0][cl][e][d];
                    long long cnt = 0;
                    for (long long c = 1; c <= M; ++c) {
                        if (fe == 2 && ((c & 1) != ((cl >> d) & 1))) continue;
                        if (c + o >= 1 && c + o <= M) ++cnt;
                    }
                    prod *= (unsigned long long)cnt;
                }
                nnz += prod;
            }
        *nnz_out = (size_t)nnz;
    }
    return 0;
}

int drops_b200_brick_csr_host(int fe, int N, int numbering, double upper_scale, size_t* rowbeg, size_t* colind,
                              double* val)
{
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, numbering, upper_scale, &g));
    DB_ARG(rowbeg && colind && val, "null output");
    const StencilTable& T = stencil(fe);
    const long long M = g.M, n = M * M * M;
    std::vector<int> len((size_t)n);
#pragma omp parallel for
    for (long long r = 0; r < n; ++r) {
        const long long i = r % M + 1, j = (r / M) % M + 1, k = r / (M * M) + 1;
        const long long row = dof_index(g, i, j, k);
        len[(size_t)row] = emit_row<size_t>(g, T, i, j, k, row, nullptr, nullptr);
    }
    rowbeg[0] = 0;
    for (long long r = 0; r < n; ++r) rowbeg[r + 1] = rowbeg[r] + (size_t)len[(size_t)r];
#pragma omp parallel for
    for (long long r = 0; r < n; ++r) {
        const long long i = r % M + 1, j = (r / M) % M + 1, k = r / (M * M) + 1;
        const long long row = dof_index(g, i, j, k);
        emit_row<size_t>(g, T, i, j, k, row, colind + rowbeg[row], val + rowbeg[row]);
    }
    return 0;
}

int drops_b200_brick_solution_host(int fe, int N, int numbering, double* xstar)
{
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, numbering, 1.0, &g));
    DB_ARG(xstar != nullptr, "null output");
    const long long M = g.M, n = M * M * M;
#pragma omp parallel for
    for (long long r = 0; r < n; ++r) {
        const long long i = r % M + 1, j = (r / M) % M + 1, k = r / (M * M) + 1;
        xstar[dof_index(g, i, j, k)] = xstar_at(g, i, j, k);
    }
    return 0;
}

int drops_b200_brick_matrix_device(int fe, int N, int numbering, double upper_scale, drops_b200_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr, "null out");
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, numbering, upper_scale, &g));
    const StencilTable& T = stencil(fe);
    const long long M = g.M, n = M * M * M;
    DB_ARG(n < 2147480000LL, "brick too large for one device partition (int32 row index)");
    cudaStream_t st = ctx().stream;
    StencilTable* dT = nullptr;
    int64_t* rowptr = nullptr; int32_t* col = nullptr; double* val = nullptr;
    void* tmp = nullptr;
    int s = 0;
    auto cleanup = [&]() { dev_free(dT); dev_free(rowptr); dev_free(col); dev_free(val); dev_free(tmp); };
#define BR_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
#define BR_CHECK(call) do { s = (call); if (s < 0) { cleanup(); return s; } } while (0)
    BR_CHECK(dev_alloc(&dT, 1));
    BR_CUDA(cudaMemcpyAsync(dT, &T, sizeof T, cudaMemcpyHostToDevice, st));
    BR_CHECK(dev_alloc(&rowptr, (size_t)n + 1));
    BR_CUDA(cudaMemsetAsync(rowptr, 0, ((size_t)n + 1) * sizeof(int64_t), st));
    const int threads = 256;
    const int grid = (int)((n + threads - 1) / threads);
    k_brick_count<<<grid, threads, 0, st>>>(g, dT, n, rowptr);
    count_launch();
    BR_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    BR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, rowptr, rowptr, n + 1, st));
    BR_CHECK(dev_alloc((char**)&tmp, tmp_bytes));
    BR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rowptr, rowptr, n + 1, st));
    count_launch(2);
    int64_t nnz = 0;
    BR_CUDA(cudaMemcpyAsync(&nnz, rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BR_CUDA(cudaStreamSynchronize(st));
    BR_CHECK(dev_alloc(&col, (size_t)nnz));
    BR_CHECK(dev_alloc(&val, (size_t)nnz));
    k_brick_fill<<<grid, threads, 0, st>>>(g, dT, n, rowptr, col, val);
    count_launch();
    BR_CUDA(cudaGetLastError());
    BR_CUDA(cudaStreamSynchronize(st));
    s = drops_b200_matrix_create_device((size_t)n, (size_t)n, (size_t)nnz, rowptr, col, val, 1, out);
    cleanup();
#undef BR_CUDA
#undef BR_CHECK
    return s;
}

int drops_b200_brick_solution_device(int fe, int N, int numbering, drops_b200_vector_t xstar)
{
    DB_CHECK(require_ctx());
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, numbering, 1.0, &g));
    const long long n = g.M * g.M * g.M;
    DB_ARG(xstar != nullptr && (long long)xstar->n == n, "vector size does not match the brick");
    k_brick_xstar<<<(int)((n + 255) / 256), 256, 0, ctx().stream>>>(g, n, xstar->d);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

int drops_b200_brick_slab_destroy(drops_b200_brick_slab_t S)
{
    if (!S) return 0;
    if (ctx().inited) cudaStreamSynchronize(ctx().stream);
    dev_free(S->d_gid); dev_free(S->d_lat); dev_free(S->d_rowptr); dev_free(S->d_col); dev_free(S->d_val);
    delete S;
    return 0;
}

int drops_b200_brick_slab_generate(int fe, int N, int numbering, double upper_scale, int world, int rank,
                                   drops_b200_brick_slab_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr, "null out");
    *out = nullptr;
    DB_ARG(world >= 1 && rank >= 0 && rank < world, "rank/world out of range");
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, numbering, upper_scale, &g));
    DB_ARG(g.M >= world, "fewer lattice planes than ranks");
    const StencilTable& T = stencil(fe);
    cudaStream_t st = ctx().stream;
    drops_b200_brick_slab_s* S = new drops_b200_brick_slab_s;
    S->g = g; S->world = world; S->rank = rank;
    S->k0 = slab_begin(g.M, world, rank, g.fe);
    S->k1 = rank + 1 == world ? g.M + 1 : slab_begin(g.M, world, rank + 1, g.fe);
    const long long n = (S->k1 - S->k0) * g.M * g.M;
    S->n_owned = (size_t)n;
    StencilTable* dT = nullptr; void* tmp = nullptr; int64_t* gid2 = nullptr; int64_t* lat2 = nullptr;
    int s = 0;
    auto cleanup = [&]() { dev_free(dT); dev_free(tmp); dev_free(gid2); dev_free(lat2); };
#define SL_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); drops_b200_brick_slab_destroy(S); return cuda_fail(e_, #call, __FILE__, __LINE__); } } while (0)
#define SL_CHECK(call) do { s = (call); if (s < 0) { cleanup(); drops_b200_brick_slab_destroy(S); return s; } } while (0)
    SL_CHECK(dev_alloc(&dT, 1));
    SL_CUDA(cudaMemcpyAsync(dT, &T, sizeof T, cudaMemcpyHostToDevice, st));
    SL_CHECK(dev_alloc(&S->d_gid, (size_t)n)); SL_CHECK(dev_alloc(&S->d_lat, (size_t)n));
    const int threads = 256;
    const int grid = (int)std::max<long long>(1, (n + threads - 1) / threads);
    const long long r0 = (S->k0 - 1) * g.M * g.M;
    k_slab_gids<<<grid, threads, 0, st>>>(g, r0, n, S->d_gid, S->d_lat);
    count_launch();
    if (fe == 2 && numbering == 1) {     // lattice order is not ascending in the DROPS numbering: sort
        SL_CHECK(dev_alloc(&gid2, (size_t)n)); SL_CHECK(dev_alloc(&lat2, (size_t)n));
        size_t tb = 0;
        SL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, S->d_gid, gid2, S->d_lat, lat2, n, 0, 40, st));
        SL_CHECK(dev_alloc((char**)&tmp, tb));
        SL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, S->d_gid, gid2, S->d_lat, lat2, n, 0, 40, st));
        count_launch(4);
        SL_CUDA(cudaStreamSynchronize(st));
        std::swap(S->d_gid, gid2); std::swap(S->d_lat, lat2);
        dev_free(gid2); gid2 = nullptr; dev_free(lat2); lat2 = nullptr; dev_free(tmp); tmp = nullptr;
    }
    SL_CHECK(dev_alloc(&S->d_rowptr, (size_t)n + 1));
    SL_CUDA(cudaMemsetAsync(S->d_rowptr, 0, ((size_t)n + 1) * sizeof(int64_t), st));
    k_slab_count<<<grid, threads, 0, st>>>(g, dT, n, S->d_lat, S->d_rowptr);
    count_launch();
    size_t tb = 0;
    SL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, S->d_rowptr, S->d_rowptr, n + 1, st));
    SL_CHECK(dev_alloc((char**)&tmp, tb));
    SL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, S->d_rowptr, S->d_rowptr, n + 1, st));
    count_launch(2);
    int64_t nnz = 0;
    SL_CUDA(cudaMemcpyAsync(&nnz, S->d_rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SL_CUDA(cudaStreamSynchronize(st));
    S->nnz = (size_t)nnz;
    SL_CHECK(dev_alloc(&S->d_col, S->nnz)); SL_CHECK(dev_alloc(&S->d_val, S->nnz));
    k_slab_fill<<<grid, threads, 0, st>>>(g, dT, n, S->d_lat, S->d_gid, S->d_rowptr, S->d_col, S->d_val);
    count_launch();
    SL_CUDA(cudaGetLastError());
    S->owned.resize((size_t)n);
    if (n) SL_CUDA(cudaMemcpyAsync(S->owned.data(), S->d_gid, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    SL_CUDA(cudaStreamSynchronize(st));
    SL_CHECK(find_ghosts_device(S->n_owned, S->owned.data(), S->nnz, S->d_col, &S->ghost_gid));
    S->ghost_owner.resize(S->ghost_gid.size());
    for (size_t i = 0; i < S->ghost_gid.size(); ++i)
        S->ghost_owner[i] = slab_owner(g.M, world, plane_of_gid(g, S->ghost_gid[i]), g.fe);
    cleanup();
#undef SL_CUDA
#undef SL_CHECK
    *out = S;
    return 0;
}

int drops_b200_brick_slab_info(drops_b200_brick_slab_t S, size_t* n_owned, size_t* nnz, size_t* n_ghost, long long* k0,
                               long long* k1)
{
    DB_ARG(S != nullptr, "null slab");
    if (n_owned) *n_owned = S->n_owned;
    if (nnz) *nnz = S->nnz;
    if (n_ghost) *n_ghost = S->ghost_gid.size();
    if (k0) *k0 = S->k0;
    if (k1) *k1 = S->k1;
    return 0;
}

int drops_b200_brick_slab_owned(drops_b200_brick_slab_t S, int64_t* owned_gid)
{
    DB_ARG(S != nullptr && owned_gid != nullptr, "null argument");
    std::copy(S->owned.begin(), S->owned.end(), owned_gid);
    return 0;
}

int drops_b200_brick_slab_ghosts(drops_b200_brick_slab_t S, int64_t* ghost_gid, int32_t* ghost_owner)
{
    DB_ARG(S != nullptr, "null slab");
    if (ghost_gid) std::copy(S->ghost_gid.begin(), S->ghost_gid.end(), ghost_gid);
    if (ghost_owner) std::copy(S->ghost_owner.begin(), S->ghost_owner.end(), ghost_owner);
    return 0;
}

int drops_b200_brick_slab_matrix(drops_b200_comm_t comm, drops_b200_exchange_t ex, drops_b200_brick_slab_t S,
                                 drops_b200_dist_matrix_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(comm && ex && S && out, "null argument");
    DB_ARG(S->d_col != nullptr, "the slab's rows were already consumed");
    DB_ARG(ex->owned.size() == S->n_owned, "plan and slab disagree on the owned rows");
    const int s = dist_matrix_build(comm, ex, S->nnz, S->d_rowptr, S->d_col, S->d_val, 1, out);
    // the operator holds its own copy: release the 20 B/entry generator output
    dev_free(S->d_rowptr); dev_free(S->d_col); dev_free(S->d_val);
    S->d_rowptr = nullptr; S->d_col = nullptr; S->d_val = nullptr;
    return s;
}

int drops_b200_brick_slab_solution(drops_b200_brick_slab_t S, drops_b200_vector_t xstar)
{
    DB_CHECK(require_ctx());
    DB_ARG(S != nullptr && xstar != nullptr && xstar->n == S->n_owned, "vector size does not match the slab");
    const long long n = (long long)S->n_owned;
    if (n == 0) return 0;
    k_slab_xstar<<<(int)((n + 255) / 256), 256, 0, ctx().stream>>>(S->g, n, S->d_lat, xstar->d);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

int drops_b200_brick_owner(int fe, int N, int numbering, int world, size_t count, const int64_t* gids, int32_t* owner)
{
    BrickGeom g;
    DB_CHECK(make_geom(fe, N, numbering, 1.0, &g));
    DB_ARG(world >= 1 && (count == 0 || (gids != nullptr && owner != nullptr)), "bad argument");
    for (size_t i = 0; i < count; ++i) {
        DB_ARG(gids[i] >= 0 && gids[i] < g.M * g.M * g.M, "global id outside the brick");
        owner[i] = slab_owner(g.M, world, plane_of_gid(g, gids[i]), g.fe);
    }
    return 0;
}

}  // extern "C"

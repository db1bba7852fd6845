// vecops.cu -- streaming vector kernels of the Krylov solvers, fused so that every iteration
// touches each vector the minimum number of times (13 vector passes per
// PCG-Jacobi iteration instead of the reference's ~21), with deterministic double-double
// reductions finished on the device by the last block (no host round trip per iteration).
//
// Element arithmetic follows the reference's valarray expressions exactly: products are rounded
// before the add (__dmul_rn/__dadd_rn, no FMA contraction), divisions are IEEE divisions.
#include "vecops.cuh"

#include <algorithm>

namespace db200 {

// ------------------------------------------------------------------------------------------
// generic streaming kernel: Op provides
//   bool begin()                         -- load scalars; false = whole kernel is a no-op
//   void elem(size_t i, dd* acc)         -- per element work, accumulates NQ reductions
//   void finish(const dd* acc)           -- run by ONE thread after the grid-wide reduction
// ------------------------------------------------------------------------------------------
template <int NQ, class Op>
__global__ void __launch_bounds__(kReduceThreads)
k_stream(size_t n, Op op, dd* partials, unsigned* ticket)
{
    constexpr int NA = NQ > 0 ? NQ : 1;
    __shared__ dd red_smem[NA * 32];
    if (!op.begin()) return;
    dd acc[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) acc[q] = dd_zero();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) op.elem(i, acc);
    if constexpr (NQ > 0) {
        if (grid_reduce_dd<NA>(acc, red_smem, partials, ticket)) {
            if (threadIdx.x == 0) op.finish(acc);
        }
    }
}

template <int NQ, class Op>
static int launch_stream(size_t n, const Op& op, const char* timing_key, int slot = 0)
{
    Ctx& c = ctx();
    ScopedTiming t(timing_key);
    const int grid = vec_grid(n, kReduceThreads, 4);
    k_stream<NQ, Op><<<grid, kReduceThreads, 0, c.stream>>>(
        n, op, c.partials + (size_t)slot * kPartialStride, c.tickets + (size_t)slot * kTicketStride);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

// z = M^-1 r for the two preconditioners on the path
__device__ __forceinline__ double precond(int pc, double omega, double r, const double* d, size_t i)
{
    return pc == DROPS_B200_PC_JACOBI ? dvd(mul(omega, r), d[i]) : r;
}

// ------------------------------------------------------------------------------------------
// standalone ops
// ------------------------------------------------------------------------------------------
struct OpDot {
    const double* x; const double* y; double* out;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc) { acc[0] = dd_add_prod(acc[0], x[i], y[i]); }
    __device__ void finish(const dd* acc) { *out = dd_round(acc[0]); }
};
int vec_dot(size_t n, const double* x, const double* y, double* d_out)
{
    return launch_stream<1>(n, OpDot{x, y, d_out}, "vec");
}
struct OpDotUnless {
    const double* x; const double* y; double* out; const int* done;
    __device__ bool begin() { return *done == 0; }
    __device__ void elem(size_t i, dd* acc) { acc[0] = dd_add_prod(acc[0], x[i], y[i]); }
    __device__ void finish(const dd* acc) { *out = dd_round(acc[0]); }
};
int vec_dot_unless(size_t n, const double* x, const double* y, double* d_out, const int* done)
{
    return launch_stream<1>(n, OpDotUnless{x, y, d_out, done}, "vec");
}
struct OpDotPartial {
    const double* x; const double* y; dd* out;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc) { acc[0] = dd_add_prod(acc[0], x[i], y[i]); }
    __device__ void finish(const dd* acc) { *out = acc[0]; }
};
int vec_dot_partial(size_t n, const double* x, const double* y, dd* part_out)
{
    return launch_stream<1>(n, OpDotPartial{x, y, part_out}, "vec");
}

__global__ void k_supnorm(size_t n, const double* __restrict__ x, double* partial)
{
    __shared__ double sm[32];
    double m = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmax(m, fabs(x[i]));
    for (int k = 16; k >= 1; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        for (int k = 16; k >= 1; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
        if (threadIdx.x == 0) partial[blockIdx.x] = m;
    }
}
__global__ void k_supnorm_final(int nb, const double* partial, double* out)
{
    double m = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) m = fmax(m, partial[i]);
    for (int k = 16; k >= 1; k >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, k));
    if (threadIdx.x == 0) *out = m;
}
int vec_supnorm(size_t n, const double* x, double* d_out)
{
    Ctx& c = ctx();
    const int grid = vec_grid(n, kReduceThreads, 4);
    double* partial = reinterpret_cast<double*>(c.partials);
    k_supnorm<<<grid, kReduceThreads, 0, c.stream>>>(n, x, partial);
    k_supnorm_final<<<1, 32, 0, c.stream>>>(grid, partial, d_out);
    count_launch(2);
    DB_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_gather_perm(size_t npos, const int32_t* __restrict__ perm, const double* __restrict__ src,
                              double* __restrict__ dst, double fill)
{
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int32_t r = perm[pos];
    dst[pos] = r >= 0 ? src[r] : fill;
}
__global__ void k_scatter_perm(size_t npos, const int32_t* __restrict__ perm, const double* __restrict__ src,
                               double* __restrict__ dst)
{
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int32_t r = perm[pos];
    if (r >= 0) dst[r] = src[pos];
}
int vec_gather_perm(size_t npos, const int32_t* perm, const double* src, double* dst, double fill)
{
    if (npos == 0) return 0;
    k_gather_perm<<<(unsigned)((npos + 255) / 256), 256, 0, ctx().stream>>>(npos, perm, src, dst, fill);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}
int vec_scatter_perm(size_t npos, const int32_t* perm, const double* src, double* dst)
{
    if (npos == 0) return 0;
    k_scatter_perm<<<(unsigned)((npos + 255) / 256), 256, 0, ctx().stream>>>(npos, perm, src, dst);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

struct OpAxpy {
    double a; const double* x; double* y;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { y[i] = add(y[i], mul(a, x[i])); }
    __device__ void finish(const dd*) {}
};
int vec_axpy(size_t n, double a, const double* x, double* y) { return launch_stream<0>(n, OpAxpy{a, x, y}, "vec"); }

struct OpZxpay {
    double* z; const double* x; double a; const double* y;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { z[i] = add(x[i], mul(a, y[i])); }
    __device__ void finish(const dd*) {}
};
int vec_z_xpay(size_t n, double* z, const double* x, double a, const double* y)
{
    return launch_stream<0>(n, OpZxpay{z, x, a, y}, "vec");
}

struct OpZxpaypby2 {
    double* z; const double* x; double a; const double* y; double b; const double* y2;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { z[i] = add(add(x[i], mul(a, y[i])), mul(b, y2[i])); }
    __device__ void finish(const dd*) {}
};
int vec_z_xpaypby2(size_t n, double* z, const double* x, double a, const double* y, double b, const double* y2)
{
    return launch_stream<0>(n, OpZxpaypby2{z, x, a, y, b, y2}, "vec");
}

struct OpJacobi {
    double omega; const double* d; const double* b; double* x;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { x[i] = dvd(mul(omega, b[i]), d[i]); }
    __device__ void finish(const dd*) {}
};
int vec_jacobi(size_t n, double omega, const double* d, const double* b, double* x)
{
    return launch_stream<0>(n, OpJacobi{omega, d, b, x}, "vec");
}

struct OpFill {
    double* x; double v;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { x[i] = v; }
    __device__ void finish(const dd*) {}
};
int vec_fill(size_t n, double* x, double v) { return launch_stream<0>(n, OpFill{x, v}, "vec"); }

struct OpFillHash {
    double* x;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*)
    {
        const unsigned long long h = ((unsigned long long)i * 2654435761ULL) & 0xffffffffULL;
        x[i] = (double)h / 4294967296.0 - 0.5;
    }
    __device__ void finish(const dd*) {}
};
int vec_fill_hash(size_t n, double* x) { return launch_stream<0>(n, OpFillHash{x}, "vec"); }

int vec_copy(size_t n, double* dst, const double* src)
{
    if (n == 0 || dst == src) return 0;
    DB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
    return 0;
}

struct OpSub {
    double* out; const double* a; const double* b;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { out[i] = sub(a[i], b[i]); }
    __device__ void finish(const dd*) {}
};
int vec_sub(size_t n, double* out, const double* a, const double* b) { return launch_stream<0>(n, OpSub{out, a, b}, "vec"); }

// ------------------------------------------------------------------------------------------
// PCG
// ------------------------------------------------------------------------------------------
struct OpPcgSetup {
    const double* r; const double* b; const double* d; double* p; SolverScalars* S; double* hist;
    int pc; double omega; dd* part_out;
    __device__ bool begin() { pc = S->pc; omega = S->omega; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double ri = r[i], bi = b[i];
        const double z = precond(pc, omega, ri, d, i);
        p[i] = z;
        acc[0] = dd_add_prod(acc[0], ri, ri);
        acc[1] = dd_add_prod(acc[1], bi, bi);
        acc[2] = dd_add_prod(acc[2], z, ri);
    }
    __device__ void finish(const dd* acc)
    {
        if (part_out) { part_out[0] = acc[0]; part_out[1] = acc[1]; part_out[2] = acc[2]; return; }
        pcg_setup_finish(S, hist, dd_round(acc[0]), dd_round(acc[1]), dd_round(acc[2]));
    }
};
int pcg_setup(size_t n, const double* r, const double* b, const double* d, double* p, SolverScalars* S, double* hist,
              dd* part_out)
{
    return launch_stream<3>(n, OpPcgSetup{r, b, d, p, S, hist, 0, 1.0, part_out}, "pcg.setup");
}

struct OpPcgUpdate {
    double* x; double* r; const double* p; const double* q; const double* d; SolverScalars* S; double* hist;
    double alpha; int pc; double omega; dd* part_out;
    __device__ bool begin()
    {
        if (S->done) return false;
        alpha = dvd(S->rho, S->lambda);                         // :152
        pc = S->pc; omega = S->omega;
        return true;
    }
    __device__ void elem(size_t i, dd* acc)
    {
        x[i] = add(x[i], mul(alpha, p[i]));                     // :154
        const double ri = add(r[i], mul(-alpha, q[i]));         // :155 (axpy(-alpha, q, r))
        r[i] = ri;
        const double z = precond(pc, omega, ri, d, i);          // :170
        acc[0] = dd_add_prod(acc[0], ri, ri);                   // :158
        acc[1] = dd_add_prod(acc[1], z, ri);                    // :172
    }
    __device__ void finish(const dd* acc)
    {
        if (part_out) { part_out[0] = acc[0]; part_out[1] = acc[1]; return; }
        pcg_update_finish(S, hist, dd_round(acc[0]), dd_round(acc[1]));
    }
};
// 16-byte vectorised form of the two per-iteration PCG kernels: every thread has 5 (update) or 3
// (direction) independent 16 B loads in flight before the first use, which is what lifts these
// latency-bound streams to HBM speed (ncu, profiles/r01: scalar form 4.1 TB/s at 58 % occupancy).
// Same element arithmetic as OpPcgUpdate / OpPcgDirection.
static inline bool aligned16(const void* a, const void* b, const void* c, const void* d = nullptr,
                             const void* e = nullptr)
{
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
             reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e)) & 15u) == 0;
}

__global__ void __launch_bounds__(kReduceThreads)
k_pcg_update2(size_t n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
              const double* __restrict__ q, const double* __restrict__ d, SolverScalars* S, double* hist,
              dd* part_out, dd* partials, unsigned* ticket)
{
    __shared__ dd red_smem[2 * 32];
    OpPcgUpdate op{x, r, p, q, d, S, hist, 0.0, 0, 1.0, part_out};
    if (!op.begin()) return;
    const double alpha = op.alpha, omega = op.omega;
    const bool jac = op.pc == DROPS_B200_PC_JACOBI;
    dd acc[2] = {dd_zero(), dd_zero()};
    const size_t n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* q2 = reinterpret_cast<const double2*>(q);
    const double2* d2 = reinterpret_cast<const double2*>(d);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 xv = x2[i], rv = r2[i];
        const double2 pv = p2[i], qv = __ldcs(q2 + i);          // q is dead after this kernel
        double2 dv = make_double2(1.0, 1.0);
        if (jac) dv = d2[i];
        xv.x = add(xv.x, mul(alpha, pv.x)); xv.y = add(xv.y, mul(alpha, pv.y));
        rv.x = add(rv.x, mul(-alpha, qv.x)); rv.y = add(rv.y, mul(-alpha, qv.y));
        x2[i] = xv; r2[i] = rv;
        const double zx = jac ? dvd(mul(omega, rv.x), dv.x) : rv.x;
        const double zy = jac ? dvd(mul(omega, rv.y), dv.y) : rv.y;
        acc[0] = dd_add_prod(acc[0], rv.x, rv.x); acc[0] = dd_add_prod(acc[0], rv.y, rv.y);
        acc[1] = dd_add_prod(acc[1], zx, rv.x);   acc[1] = dd_add_prod(acc[1], zy, rv.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.elem(n - 1, acc);
    if (grid_reduce_dd<2>(acc, red_smem, partials, ticket)) {
        if (threadIdx.x == 0) op.finish(acc);
    }
}

int pcg_update(size_t n, double* x, double* r, const double* p, const double* q, const double* d,
               SolverScalars* S, double* hist, dd* part_out)
{
    if (opt("vec.wide", 1.0) != 0.0 && aligned16(x, r, p, q, d)) {
        Ctx& c = ctx();
        ScopedTiming t("pcg.update");
        const int grid = vec_grid(n, kReduceThreads, 4);
        k_pcg_update2<<<grid, kReduceThreads, 0, c.stream>>>(n, x, r, p, q, d, S, hist, part_out, c.partials, c.tickets);
        count_launch();
        DB_CUDA(cudaGetLastError());
        return 0;
    }
    return launch_stream<2>(n, OpPcgUpdate{x, r, p, q, d, S, hist, 0.0, 0, 1.0, part_out}, "pcg.update");
}

struct OpPcgDirection {
    double* p; const double* r; const double* d; SolverScalars* S;
    double beta; int pc; double omega;
    __device__ bool begin()
    {
        if (S->done) return false;
        beta = dvd(S->rho, S->rho_1);                           // :173
        pc = S->pc; omega = S->omega;
        return true;
    }
    __device__ void elem(size_t i, dd*)
    {
        const double z = precond(pc, omega, r[i], d, i);
        p[i] = add(z, mul(beta, p[i]));                         // z_xpay(p, z, beta, p)
    }
    __device__ void finish(const dd*) {}
};
__global__ void __launch_bounds__(kReduceThreads)
k_pcg_direction2(size_t n, double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ d,
                 SolverScalars* S)
{
    OpPcgDirection op{p, r, d, S, 0.0, 0, 1.0};
    if (!op.begin()) return;
    const double beta = op.beta, omega = op.omega;
    const bool jac = op.pc == DROPS_B200_PC_JACOBI;
    const size_t n2 = n >> 1;
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    const double2* d2 = reinterpret_cast<const double2*>(d);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 pv = p2[i];
        const double2 rv = r2[i];
        double2 dv = make_double2(1.0, 1.0);
        if (jac) dv = d2[i];
        const double zx = jac ? dvd(mul(omega, rv.x), dv.x) : rv.x;
        const double zy = jac ? dvd(mul(omega, rv.y), dv.y) : rv.y;
        pv.x = add(zx, mul(beta, pv.x)); pv.y = add(zy, mul(beta, pv.y));
        p2[i] = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.elem(n - 1, nullptr);
}

int pcg_direction(size_t n, double* p, const double* r, const double* d, SolverScalars* S)
{
    if (opt("vec.wide", 1.0) != 0.0 && aligned16(p, r, d)) {
        Ctx& c = ctx();
        ScopedTiming t("pcg.direction");
        const int grid = vec_grid(n, kReduceThreads, 4);
        k_pcg_direction2<<<grid, kReduceThreads, 0, c.stream>>>(n, p, r, d, S);
        count_launch();
        DB_CUDA(cudaGetLastError());
        return 0;
    }
    return launch_stream<0>(n, OpPcgDirection{p, r, d, S, 0.0, 0, 1.0}, "pcg.direction");
}

// ---- deferred x update: 4 + 6 vector passes instead of 7 + 4 ---------------------------------------
struct OpPcgUpdateR {
    double* r; const double* q; const double* d; SolverScalars* S; double* hist; dd* part_out;
    double alpha; int pc; double omega;
    __device__ bool begin()
    {
        if (S->done) return false;
        alpha = dvd(S->rho, S->lambda);                         // :152
        pc = S->pc; omega = S->omega;
        return true;
    }
    __device__ void elem(size_t i, dd* acc)
    {
        const double ri = add(r[i], mul(-alpha, q[i]));         // :155
        r[i] = ri;
        const double z = precond(pc, omega, ri, d, i);          // :170
        acc[0] = dd_add_prod(acc[0], ri, ri);                   // :158
        acc[1] = dd_add_prod(acc[1], z, ri);                    // :172
    }
    __device__ void finish(const dd* acc)
    {
        S->alpha_prev = alpha;
        if (part_out) { part_out[0] = acc[0]; part_out[1] = acc[1]; return; }      // partitioned: k_dist_scalars finishes
        pcg_update_finish(S, hist, dd_round(acc[0]), dd_round(acc[1]));
    }
};

// d_i of the element pair i: from the vector, or (DIDX) from the byte index + the dictionary in the parameter block
// Two steps so that every global load of a trip is issued before the first dictionary lookup: the lookup (an
// indexed constant load) waits for its byte pair, and placed between the loads it would split them in two batches.
struct DiagRaw { double2 v; uchar2 k; };

template <bool DIDX>
__device__ __forceinline__ DiagRaw load_diag_raw(const double2* __restrict__ d2, const uint8_t* __restrict__ didx, size_t i)
{
    DiagRaw w;
    w.v = make_double2(1.0, 1.0); w.k = make_uchar2(0, 0);
    if (DIDX) w.k = __ldg(reinterpret_cast<const uchar2*>(didx) + i);
    else w.v = d2[i];
    return w;
}

// (the value list is copied from the parameter block to shared memory: neighbouring positions hold rows of
// different stencil types, and a constant load whose index diverges inside the warp is replayed per address --
// measured on C2, direction kernel: 0.364 ms with the fp64 vector, 0.443 ms with indexed constant loads)
template <bool DIDX>
__device__ __forceinline__ double2 resolve_diag(const DiagRaw& w, const double* sdict)
{
    return DIDX ? make_double2(sdict[w.k.x], sdict[w.k.y]) : w.v;
}

template <int STEP, bool DIDX>
__global__ void __launch_bounds__(kReduceThreads)
k_pcg_update_r(size_t n, double* __restrict__ r, const double* __restrict__ q, const double* __restrict__ d,
               SolverScalars* S, double* hist, dd* part_out, dd* partials, unsigned* ticket,
               const uint8_t* __restrict__ didx, const __grid_constant__ DiagDict dict)
{
    __shared__ dd red_smem[2 * 32];
    __shared__ double sdict[DIDX ? kDiagDict : 1];
    OpPcgUpdateR op{r, q, d, S, hist, part_out, 0.0, 0, 1.0};
    if (!op.begin()) return;
    if (DIDX) {
        if (threadIdx.x < kDiagDict) sdict[threadIdx.x] = dict.v[threadIdx.x];
        __syncthreads();
    }
    const double alpha = op.alpha, omega = op.omega;
    const bool jac = op.pc == DROPS_B200_PC_JACOBI;
    dd acc[2] = {dd_zero(), dd_zero()};
    const size_t n2 = n >> 1;
    double2* r2 = reinterpret_cast<double2*>(r);
    const double2* q2 = reinterpret_cast<const double2*>(q);
    const double2* d2 = reinterpret_cast<const double2*>(d);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // two elements of the grid-stride sequence per trip: 6 independent 16 B loads in flight per thread; the
    // per-thread order of the accumulated products is unchanged (i, i + stride, i + 2 stride, ...)
    auto body = [&](size_t i, double2 rv, const double2 qv, const DiagRaw& dw) {
        const double2 dv = jac ? resolve_diag<DIDX>(dw, sdict) : make_double2(1.0, 1.0);
        rv.x = add(rv.x, mul(-alpha, qv.x)); rv.y = add(rv.y, mul(-alpha, qv.y));
        r2[i] = rv;
        const double zx = jac ? dvd(mul(omega, rv.x), dv.x) : rv.x;
        const double zy = jac ? dvd(mul(omega, rv.y), dv.y) : rv.y;
        acc[0] = dd_add_prod(acc[0], rv.x, rv.x); acc[0] = dd_add_prod(acc[0], rv.y, rv.y);
        acc[1] = dd_add_prod(acc[1], zx, rv.x);   acc[1] = dd_add_prod(acc[1], zy, rv.y);
    };
    const double2 one2 = make_double2(1.0, 1.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += STEP * stride) {
        const size_t j = i + stride;
        const bool two = STEP == 2 && j < n2;
        const double2 rv0 = r2[i], qv0 = __ldcs(q2 + i);            // q is dead after this kernel
        DiagRaw dw0, dw1;
        dw0.v = one2; dw0.k = make_uchar2(0, 0); dw1 = dw0;
        if (jac) dw0 = load_diag_raw<DIDX>(d2, didx, i);
        double2 rv1 = one2, qv1 = one2;
        if (two) { rv1 = r2[j]; qv1 = __ldcs(q2 + j); if (jac) dw1 = load_diag_raw<DIDX>(d2, didx, j); }
        body(i, rv0, qv0, dw0);
        if (two) body(j, rv1, qv1, dw1);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) op.elem(n - 1, acc);
    if (grid_reduce_dd<2>(acc, red_smem, partials, ticket)) {
        if (threadIdx.x == 0) op.finish(acc);
    }
}

int pcg_update_r(size_t n, double* r, const double* q, const double* d, SolverScalars* S, double* hist, dd* part_out,
                 const DiagIndex* di)
{
    if (di != nullptr && di->idx != nullptr && (n & 1) == 0 && opt("vec.wide", 1.0) != 0.0 && aligned16(r, q, d)) {
        Ctx& c = ctx();
        ScopedTiming t("pcg.update");
        const int grid = vec_grid(n, kReduceThreads, 4);
        if (opt("vec.unroll2_update_idx", 1.0) != 0.0)       // with the byte index two elements per trip pay: 0.263 -> 0.252 ms
            k_pcg_update_r<2, true><<<grid, kReduceThreads, 0, c.stream>>>(n, r, q, d, S, hist, part_out, c.partials, c.tickets, di->idx, di->dict);
        else
            k_pcg_update_r<1, true><<<grid, kReduceThreads, 0, c.stream>>>(n, r, q, d, S, hist, part_out, c.partials, c.tickets, di->idx, di->dict);
        count_launch();
        DB_CUDA(cudaGetLastError());
        return 0;
    }
    if (opt("vec.wide", 1.0) != 0.0 && aligned16(r, q, d)) {
        Ctx& c = ctx();
        ScopedTiming t("pcg.update");
        const int grid = vec_grid(n, kReduceThreads, 4);
        static const DiagDict no_dict = {};
        // measured on C2: two elements per trip help the direction kernel (0.375 -> 0.362 ms) and hurt this
        // one (0.282 -> 0.306 ms: 74 registers), so the default differs ("vec.unroll2_update")
        if (opt("vec.unroll2_update", 0.0) != 0.0)
            k_pcg_update_r<2, false><<<grid, kReduceThreads, 0, c.stream>>>(n, r, q, d, S, hist, part_out, c.partials, c.tickets, nullptr, no_dict);
        else
            k_pcg_update_r<1, false><<<grid, kReduceThreads, 0, c.stream>>>(n, r, q, d, S, hist, part_out, c.partials, c.tickets, nullptr, no_dict);
        count_launch();
        DB_CUDA(cudaGetLastError());
        return 0;
    }
    return launch_stream<2>(n, OpPcgUpdateR{r, q, d, S, hist, part_out, 0.0, 0, 1.0}, "pcg.update");
}

// x += alpha p for the iteration pcg_update_r just completed (S->iter counts it, S->x_iter does not
// yet); then, unless that iteration ended the solve, p = z + beta p.  After the rotation in
// pcg_update_finish the iteration's rho sits in rho_1, so alpha = rho_1 / lambda is the quotient the
// update kernel used (:152).  The block that finishes last marks the x update as applied; kernels of
// iterations that are skipped after the stop see x_iter == iter and do nothing.
template <int STEP, bool DIDX>
__global__ void __launch_bounds__(kReduceThreads)
k_pcg_direction_x(size_t n, double* __restrict__ x, double* __restrict__ p, const double* __restrict__ r,
                  const double* __restrict__ d, SolverScalars* S, unsigned* ticket,
                  const uint8_t* __restrict__ didx, const __grid_constant__ DiagDict dict)
{
    __shared__ double sdict[DIDX ? kDiagDict : 1];
    const int it = S->iter;
    if (S->x_iter == it) return;
    if (DIDX) {
        if (threadIdx.x < kDiagDict) sdict[threadIdx.x] = dict.v[threadIdx.x];
        __syncthreads();
    }
    const bool more = S->done == 0;
    const double alpha = dvd(S->rho_1, S->lambda);
    const double beta = more ? dvd(S->rho, S->rho_1) : 0.0;     // :173
    const double omega = S->omega;
    const bool jac = S->pc == DROPS_B200_PC_JACOBI;
    const size_t n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* p2 = reinterpret_cast<double2*>(p);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    const double2* d2 = reinterpret_cast<const double2*>(d);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    auto body = [&](size_t i, double2 xv, double2 pv, const double2 rv, const DiagRaw& dw) {
        xv.x = add(xv.x, mul(alpha, pv.x)); xv.y = add(xv.y, mul(alpha, pv.y));   // :154
        x2[i] = xv;
        if (more) {
            const double2 dv = jac ? resolve_diag<DIDX>(dw, sdict) : make_double2(1.0, 1.0);
            const double zx = jac ? dvd(mul(omega, rv.x), dv.x) : rv.x;
            const double zy = jac ? dvd(mul(omega, rv.y), dv.y) : rv.y;
            pv.x = add(zx, mul(beta, pv.x)); pv.y = add(zy, mul(beta, pv.y));     // :174
            p2[i] = pv;
        }
    };
    const double2 one2 = make_double2(1.0, 1.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += STEP * stride) {    // STEP elements per trip
        const size_t j = i + stride;
        const bool two = STEP == 2 && j < n2;
        const double2 xv0 = x2[i], pv0 = p2[i];
        const double2 rv0 = more ? r2[i] : one2;
        DiagRaw dw0, dw1;
        dw0.v = one2; dw0.k = make_uchar2(0, 0); dw1 = dw0;
        if (more && jac) dw0 = load_diag_raw<DIDX>(d2, didx, i);
        double2 xv1 = one2, pv1 = one2, rv1 = one2;
        if (two) { xv1 = x2[j]; pv1 = p2[j]; if (more) { rv1 = r2[j]; if (jac) dw1 = load_diag_raw<DIDX>(d2, didx, j); } }
        body(i, xv0, pv0, rv0, dw0);
        if (two) body(j, xv1, pv1, rv1, dw1);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const size_t i = n - 1;
        const double pi = p[i];
        x[i] = add(x[i], mul(alpha, pi));
        if (more) p[i] = add(precond(S->pc, omega, r[i], d, i), mul(beta, pi));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { S->x_iter = it; *ticket = 0u; }
    }
}

// CTAs of `kernel` that are resident at once on the whole device: a grid-stride kernel launched with 8 CTAs per SM
// but 3 resident (74 registers) runs 2.67 waves, the last one a third empty
template <typename K>
static int resident_ctas(K kernel, int threads)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 8;
    }
    return per_sm * ctx().sm_count;
}

int pcg_direction_x(size_t n, double* x, double* p, const double* r, const double* d, SolverScalars* S, const DiagIndex* di)
{
    Ctx& c = ctx();
    DB_ARG(aligned16(x, p, r, d), "pcg_direction_x needs 16-byte aligned vectors");
    ScopedTiming t("pcg.direction");
    int grid = vec_grid(n, kReduceThreads, 4);
    unsigned* ticket = c.tickets + (size_t)2 * kTicketStride;
    static const DiagDict no_dict = {};
    const bool two = opt("vec.unroll2", 1.0) != 0.0;
    const bool idx = di != nullptr && di->idx != nullptr && (n & 1) == 0 && opt("pcg.diag_index_direction", 0.0) != 0.0;
    // Measured on C2 (profiles/r01_notes.md, third part): both off.  One resident wave: 0.363 -> 0.385 ms (the block
    // scheduler backfills an oversubscribed grid better than the stride loop balances one wave); the byte index in
    // THIS kernel: 0.364 -> 0.42-0.45 ms although 0.35 GB less are read -- the kernel is bound by loads in flight
    // at 3 CTAs/SM, and a 2-byte load in flight buys less bandwidth than the 16-byte load it replaces.
    if (opt("vec.one_wave", 0.0) != 0.0) {           // (no reduction in this kernel: the grid does not touch any result)
        static int res[4] = {0, 0, 0, 0};
        const int which = (idx ? 2 : 0) + (two ? 1 : 0);
        if (res[which] == 0)
            res[which] = which == 3 ? resident_ctas(k_pcg_direction_x<2, true>, kReduceThreads)
                       : which == 2 ? resident_ctas(k_pcg_direction_x<1, true>, kReduceThreads)
                       : which == 1 ? resident_ctas(k_pcg_direction_x<2, false>, kReduceThreads)
                                    : resident_ctas(k_pcg_direction_x<1, false>, kReduceThreads);
        grid = std::min(grid, res[which]);
    }
    if (idx) {
        if (two) k_pcg_direction_x<2, true><<<grid, kReduceThreads, 0, c.stream>>>(n, x, p, r, d, S, ticket, di->idx, di->dict);
        else k_pcg_direction_x<1, true><<<grid, kReduceThreads, 0, c.stream>>>(n, x, p, r, d, S, ticket, di->idx, di->dict);
    } else if (two)
        k_pcg_direction_x<2, false><<<grid, kReduceThreads, 0, c.stream>>>(n, x, p, r, d, S, ticket, nullptr, no_dict);
    else
        k_pcg_direction_x<1, false><<<grid, kReduceThreads, 0, c.stream>>>(n, x, p, r, d, S, ticket, nullptr, no_dict);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// PCG with an external preconditioner, Chebyshev, IDR(s) helpers
// ------------------------------------------------------------------------------------------
struct OpPcgSetRho {
    const double* z; const double* r; SolverScalars* S;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc) { acc[0] = dd_add_prod(acc[0], z[i], r[i]); }
    __device__ void finish(const dd* acc) { S->rho = dd_round(acc[0]); }
};
int pcg_set_rho(size_t n, const double* z, const double* r, SolverScalars* S)
{
    return launch_stream<1>(n, OpPcgSetRho{z, r, S}, "pcg.rho");
}

struct OpPcgDirectionZ {
    double* p; const double* z; SolverScalars* S; double beta;
    __device__ bool begin()
    {
        if (S->done) return false;
        beta = dvd(S->rho, S->rho_1);
        return true;
    }
    __device__ void elem(size_t i, dd*) { p[i] = add(z[i], mul(beta, p[i])); }
    __device__ void finish(const dd*) {}
};
int pcg_direction_z(size_t n, double* p, const double* z, SolverScalars* S)
{
    return launch_stream<0>(n, OpPcgDirectionZ{p, z, S, 0.0}, "pcg.direction");
}

struct OpChebFirst {
    double* W; double* x; const double* res; const double* d; double oot;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*)
    {
        const double w = mul(mul(dvd(1.0, d[i]), res[i]), oot);
        W[i] = w; x[i] = w;
    }
    __device__ void finish(const dd*) {}
};
int cheb_first(size_t n, double* W, double* x, const double* res, const double* d, double one_over_theta)
{
    return launch_stream<0>(n, OpChebFirst{W, x, res, d, one_over_theta}, "cheb");
}

struct OpChebStep {
    double* W; double* x; const double* res; const double* d; double dt1, dt2;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*)
    {
        const double w = add(mul(W[i], dt1), mul(mul(dt2, dvd(1.0, d[i])), res[i]));
        W[i] = w;
        x[i] = add(x[i], w);
    }
    __device__ void finish(const dd*) {}
};
int cheb_step(size_t n, double* W, double* x, const double* res, const double* d, double dt1, double dt2)
{
    return launch_stream<0>(n, OpChebStep{W, x, res, d, dt1, dt2}, "cheb");
}

struct OpMul {
    double* z; const double* x; const double* y;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { z[i] = mul(x[i], y[i]); }
    __device__ void finish(const dd*) {}
};
int vec_mul(size_t n, double* z, const double* x, const double* y)
{
    return launch_stream<0>(n, OpMul{z, x, y}, "vec");
}

struct OpAxpby {
    double a; double* y; double b; const double* x;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*) { y[i] = add(mul(a, y[i]), mul(b, x[i])); }
    __device__ void finish(const dd*) {}
};
int vec_axpby(size_t n, double a, double* y, double b, const double* x)
{
    return launch_stream<0>(n, OpAxpby{a, y, b, x}, "vec");
}

template <int K>
struct OpMultiDot {
    const double* x; const double* v[K]; double* out;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double xi = x[i];
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = dd_add_prod(acc[j], v[j][i], xi);
    }
    __device__ void finish(const dd* acc)
    {
#pragma unroll
        for (int j = 0; j < K; ++j) out[j] = dd_round(acc[j]);
    }
};
int multi_dot(size_t n, const double* x, const double* const* h_vptrs, int k, double* d_out)
{
    for (int j0 = 0; j0 < k; j0 += 4) {
        const int m = k - j0 < 4 ? k - j0 : 4;
        const double* const* v = h_vptrs + j0;
        int s = 0;
        switch (m) {
            case 1: s = launch_stream<1>(n, OpMultiDot<1>{x, {v[0]}, d_out + j0}, "vec"); break;
            case 2: s = launch_stream<2>(n, OpMultiDot<2>{x, {v[0], v[1]}, d_out + j0}, "vec"); break;
            case 3: s = launch_stream<3>(n, OpMultiDot<3>{x, {v[0], v[1], v[2]}, d_out + j0}, "vec"); break;
            default: s = launch_stream<4>(n, OpMultiDot<4>{x, {v[0], v[1], v[2], v[3]}, d_out + j0}, "vec"); break;
        }
        if (s < 0) return s;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// CG (sign convention of the reference: r = A x - b, d = -r)
// ------------------------------------------------------------------------------------------
struct OpCgSetup {
    const double* r; const double* b; double* dv; SolverScalars* S; double* hist;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double ri = r[i], bi = b[i];
        dv[i] = -ri;                                            // :65
        acc[0] = dd_add_prod(acc[0], ri, ri);
        acc[1] = dd_add_prod(acc[1], bi, bi);
    }
    __device__ void finish(const dd* acc)
    {
        double normb = sqrt(dd_round(acc[1]));                  // :62
        if (normb == 0.0 || !S->relative) normb = 1.0;          // :67
        const double rr = dd_round(acc[0]);                     // resid = Norm_sq :64
        const double res = sqrt(rr) / normb;
        S->normb = normb; S->rr = rr; S->resid = res; S->iter = 0;
        if (hist && S->hist_cap > 0) hist[0] = sqrt(rr);
        if (res <= S->tol) { S->done = 1; S->converged = 1; }   // :69-74
        else if (S->max_iter <= 0) S->done = 1;
    }
};
int cg_setup(size_t n, const double* r, const double* b, double* dvec, SolverScalars* S, double* hist)
{
    return launch_stream<2>(n, OpCgSetup{r, b, dvec, S, hist}, "cg.setup");
}

struct OpCgUpdate {
    double* x; double* r; const double* dv; const double* Ad; SolverScalars* S; double* hist;
    double alpha;
    __device__ bool begin()
    {
        if (S->done) return false;
        alpha = dvd(S->rr, S->lambda);                          // alpha = resid/delta :80
        return true;
    }
    __device__ void elem(size_t i, dd* acc)
    {
        x[i] = add(x[i], mul(alpha, dv[i]));                    // :83
        const double ri = add(r[i], mul(alpha, Ad[i]));         // :84
        r[i] = ri;
        acc[0] = dd_add_prod(acc[0], ri, ri);                   // :86
    }
    __device__ void finish(const dd* acc)
    {
        const double rr = dd_round(acc[0]);
        const int it = S->iter + 1;
        S->rho_1 = S->rr;                                       // beta = resid_new / resid_old :98
        S->rr = rr; S->iter = it;
        const double res = sqrt(rr) / S->normb;
        S->resid = res;
        if (hist && it < S->hist_cap) hist[it] = sqrt(rr);      // :88-90
        if (res <= S->tol) { S->done = 1; S->converged = 1; }
        else if (it >= S->max_iter) S->done = 1;
    }
};
int cg_update(size_t n, double* x, double* r, const double* dvec, const double* Ad, SolverScalars* S, double* hist)
{
    return launch_stream<1>(n, OpCgUpdate{x, r, dvec, Ad, S, hist, 0.0}, "cg.update");
}

struct OpCgDirection {
    double* dv; const double* r; SolverScalars* S; double beta;
    __device__ bool begin()
    {
        if (S->done) return false;
        beta = dvd(S->rr, S->rho_1);
        return true;
    }
    __device__ void elem(size_t i, dd*) { dv[i] = sub(mul(beta, dv[i]), r[i]); }   // :99
    __device__ void finish(const dd*) {}
};
int cg_direction(size_t n, double* dvec, const double* r, SolverScalars* S)
{
    return launch_stream<0>(n, OpCgDirection{dvec, r, S, 0.0}, "cg.direction");
}

// ------------------------------------------------------------------------------------------
// GMRES pieces
// ------------------------------------------------------------------------------------------
struct OpScaleInv {
    double* out; const double* in; const double* scale; double inv; const int* done;
    __device__ bool begin() { if (done != nullptr && *done != 0) return false; inv = dvd(1.0, *scale); return true; }
    __device__ void elem(size_t i, dd*) { out[i] = mul(in[i], inv); }
    __device__ void finish(const dd*) {}
};
int vec_scale_inv(size_t n, double* out, const double* in, const double* d_scale, const int* done)
{
    return launch_stream<0>(n, OpScaleInv{out, in, d_scale, 0.0, done}, "vec");
}

template <bool PREV, bool NORM>
struct OpMgs {
    double* w; const double* vprev; const double* hprev; const double* vk; double* hout; double h; dd* part_out;
    __device__ bool begin() { h = PREV ? *hprev : 0.0; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        double wi = w[i];
        if (PREV) { wi = sub(wi, mul(h, vprev[i])); w[i] = wi; }   // w -= H(k,i)*v[k]  :380
        acc[0] = NORM ? dd_add_prod(acc[0], wi, wi) : dd_add_prod(acc[0], wi, vk[i]);
    }
    __device__ void finish(const dd* acc)
    {
        if (part_out) { *part_out = acc[0]; return; }      // partitioned: the sum over ranks (and the sqrt) come later
        *hout = NORM ? sqrt(dd_round(acc[0])) : dd_round(acc[0]);
    }
};
int mgs_step(size_t n, double* w, const double* vprev, const double* d_hprev, const double* vk,
             double* d_hout, bool want_norm, dd* part_out)
{
    if (vprev) {
        if (want_norm) return launch_stream<1>(n, OpMgs<true, true>{w, vprev, d_hprev, vk, d_hout, 0.0, part_out}, "gmres.mgs");
        return launch_stream<1>(n, OpMgs<true, false>{w, vprev, d_hprev, vk, d_hout, 0.0, part_out}, "gmres.mgs");
    }
    if (want_norm) return launch_stream<1>(n, OpMgs<false, true>{w, vprev, d_hprev, vk, d_hout, 0.0, part_out}, "gmres.mgs");
    return launch_stream<1>(n, OpMgs<false, false>{w, vprev, d_hprev, vk, d_hout, 0.0, part_out}, "gmres.mgs");
}

struct OpMultiAxpy {
    double* x; const double* const* v; const double* y; int k;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*)
    {
        double xi = x[i];
        for (int j = 0; j < k; ++j) xi = add(xi, mul(y[j], v[j][i]));   // x += y[i]*v[i]  :295-296
        x[i] = xi;
    }
    __device__ void finish(const dd*) {}
};
int multi_axpy(size_t n, double* x, const double* const* d_vptrs, const double* d_y, int k)
{
    return launch_stream<0>(n, OpMultiAxpy{x, d_vptrs, d_y, k}, "gmres.update");
}

// ------------------------------------------------------------------------------------------
// Classical Gram-Schmidt: StandardGramSchmidt, num/krylovsolver.h:465-494 -- what the reference's distributed
// ModGMRES runs by default (:862-866): all i+1 products h_k = w.v_k against the UNMODIFIED w, one synchronisation
// point, then w -= sum_k h_k v_k.  Two kernels per Hessenberg column instead of i+1 dependent ones, and
// 2(i+1) + 2 instead of 4(i+1) vector passes.
//
// k_cgs_dots: CTA c owns a contiguous range of rows and walks it in tiles of kCgsTile rows staged in shared
// memory; its warps take the basis vectors round-robin (no two warps share an accumulator), every (vector, tile)
// partial is reduced over the warp's lanes and added to the CTA's accumulator of that vector in tile order; the
// CTA that draws the last ticket sums the per-CTA partials of each vector in CTA order.  w and every v_k are
// read once; the result does not depend on scheduling.
// ------------------------------------------------------------------------------------------
constexpr int kCgsTile = 1024;
constexpr int kCgsMaxVec = 512;

__global__ void __launch_bounds__(256)
k_cgs_dots(size_t n, const double* __restrict__ w, const double* const* __restrict__ v, int nv, dd* partials,
           unsigned* ticket, double* h_out, dd* part_out, const int* done)
{
    if (done != nullptr && *done != 0) return;
    __shared__ double wt[kCgsTile];
    __shared__ dd acc[kCgsMaxVec];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const size_t ntiles = (n + kCgsTile - 1) / kCgsTile;
    const size_t t_begin = ntiles * blockIdx.x / gridDim.x, t_end = ntiles * (blockIdx.x + 1) / gridDim.x;
    for (int k = threadIdx.x; k < nv; k += blockDim.x) acc[k] = dd_zero();
    for (size_t t = t_begin; t < t_end; ++t) {
        const size_t r0 = t * kCgsTile;
        __syncthreads();                                   // the previous tile is consumed (and acc is initialised)
        for (int j = threadIdx.x; j < kCgsTile; j += blockDim.x) wt[j] = r0 + j < n ? w[r0 + j] : 0.0;
        __syncthreads();
        const int len = (int)(n - r0 < (size_t)kCgsTile ? n - r0 : (size_t)kCgsTile);
        for (int k = warp; k < nv; k += nwarp) {
            const double* __restrict__ vk = v[k] + r0;
            // four independent accumulators per lane: the double-double update is a long dependent chain
            dd a = dd_zero(), a1 = dd_zero(), a2 = dd_zero(), a3 = dd_zero();
            int j = lane;
            for (; j + 96 < len; j += 128) {
                const double v0 = __ldcs(vk + j), v1 = __ldcs(vk + j + 32), v2 = __ldcs(vk + j + 64), v3 = __ldcs(vk + j + 96);
                a = dd_add_prod(a, wt[j], v0); a1 = dd_add_prod(a1, wt[j + 32], v1);
                a2 = dd_add_prod(a2, wt[j + 64], v2); a3 = dd_add_prod(a3, wt[j + 96], v3);
            }
            for (; j < len; j += 32) a = dd_add_prod(a, wt[j], __ldcs(vk + j));
            a = dd_add(dd_add(a, a1), dd_add(a2, a3));
            a = warp_reduce_dd(a);
            if (lane == 0) acc[k] = dd_add(acc[k], a);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nv; k += blockDim.x) partials[(size_t)k * gridDim.x + blockIdx.x] = acc[k];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int k = warp; k < nv; k += nwarp) {
        dd a = dd_zero();
        for (unsigned c = lane; c < gridDim.x; c += 32) {
            const volatile dd* p = partials + (size_t)k * gridDim.x + c;
            dd t; t.hi = p->hi; t.lo = p->lo;
            a = dd_add(a, t);
        }
        a = warp_reduce_dd(a);
        if (lane == 0) {
            if (part_out) part_out[k] = a;                 // partitioned system: summed over the ranks later
            else h_out[k] = dd_round(a);
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

int cgs_dots(size_t n, const double* w, const double* const* d_vptrs, int nv, double* d_h, dd* part_out, const int* done)
{
    DB_ARG(nv >= 1 && nv <= kCgsMaxVec, "classical Gram-Schmidt: more basis vectors than the kernel holds accumulators for");
    Ctx& c = ctx();
    ScopedTiming t("gmres.cgs");
    const size_t ntiles = (n + kCgsTile - 1) / kCgsTile;
    // small systems (level-set, 10^5 rows): 2 CTAs per SM keep the finishing pass short; large ones need the whole
    // machine in flight (8 CTAs of 256 threads per SM)
    long long grid = std::min<long long>((long long)std::max<size_t>(ntiles, 1), (n >= ((size_t)1 << 22) ? 8LL : 2LL) * c.sm_count);
    grid = std::min<long long>(grid, kPartialStride / nv);
    k_cgs_dots<<<(int)grid, 256, 0, c.stream>>>(n, w, d_vptrs, nv, c.partials + (size_t)1 * kPartialStride,
                                                c.tickets + (size_t)1 * kTicketStride, d_h, part_out, done);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

// w -= h_k v_k for k = 0 .. nv-1 in that order (the loop of :492-493), then *d_norm = |w|_2 (:383 / :868)
struct OpCgsUpdate {
    double* w; const double* const* v; const double* h; int nv; double* norm_out; dd* part_out; const int* done;
    __device__ bool begin() { return !(done != nullptr && *done != 0); }
    __device__ void elem(size_t i, dd* acc)
    {
        double wi = w[i];
        int k = 0;
        for (; k + 4 <= nv; k += 4) {                  // four loads in flight; the subtractions stay in order
            const double v0 = __ldcs(v[k] + i), v1 = __ldcs(v[k + 1] + i), v2 = __ldcs(v[k + 2] + i), v3 = __ldcs(v[k + 3] + i);
            wi = sub(wi, mul(__ldg(h + k), v0)); wi = sub(wi, mul(__ldg(h + k + 1), v1));
            wi = sub(wi, mul(__ldg(h + k + 2), v2)); wi = sub(wi, mul(__ldg(h + k + 3), v3));
        }
        for (; k < nv; ++k) wi = sub(wi, mul(__ldg(h + k), __ldcs(v[k] + i)));
        w[i] = wi;
        acc[0] = dd_add_prod(acc[0], wi, wi);
    }
    __device__ void finish(const dd* acc)
    {
        if (part_out) { *part_out = acc[0]; return; }
        *norm_out = sqrt(dd_round(acc[0]));
    }
};
int cgs_update(size_t n, double* w, const double* const* d_vptrs, const double* d_h, int nv, double* d_norm, dd* part_out,
               const int* done)
{
    return launch_stream<1>(n, OpCgsUpdate{w, d_vptrs, d_h, nv, d_norm, part_out, done}, "gmres.cgs");
}

// Givens rotations of one GMRES column on the device (krylovsolver.h:258-279, :385-393): the new column h[0 .. i+1] gets the
// i rotations of the earlier columns, a new rotation zeroes h[i+1], the right-hand side s follows; the rotated column is
// kept in H (column-major, ld = m) for the back substitution on the host, the residual estimate |s[i+1]| / normb decides
// whether the solve stops.  The same operations in the same order as the host loop, rounded one by one (no contraction):
// the same bits.  Thread 0 of block 0 runs the dependent chain out of shared memory.
// It rides on the kernel that scales the new basis vector, v_{i+1} = w / h_{i+1,i}: block 0 rotates the column while the
// grid scales (one launch less per step: on 10^5 rows every launch is a sixth of the step).  A block that sees the stop
// flag raised by block 0 of this very launch skips a vector nobody reads any more.
__global__ void __launch_bounds__(256)
k_gmres_scale_givens(size_t n, double* __restrict__ out, const double* __restrict__ in, int i, int m, const double* __restrict__ h,
                     double* __restrict__ cs, double* __restrict__ sn, double* __restrict__ s, double* __restrict__ H, GmresCtrl* ctrl)
{
    if (ctrl->done != 0) return;
    const double inv = dvd(1.0, h[i + 1]);
    if (blockIdx.x == 0) {
        extern __shared__ double sm[];
        double* col = sm;
        double* cc = sm + (i + 2);
        double* ss = cc + i;
        for (int k = threadIdx.x; k <= i + 1; k += blockDim.x) col[k] = h[k];
        for (int k = threadIdx.x; k < i; k += blockDim.x) { cc[k] = cs[k]; ss[k] = sn[k]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 0; k < i; ++k) {
                const double dx = col[k], dy = col[k + 1];
                col[k] = __dadd_rn(__dmul_rn(cc[k], dx), __dmul_rn(ss[k], dy));
                col[k + 1] = __dadd_rn(__dmul_rn(-ss[k], dx), __dmul_rn(cc[k], dy));
            }
            double c1, s1;
            const double dx = col[i], dy = col[i + 1];
            if (dy == 0.0) { c1 = 1.0; s1 = 0.0; }
            else {
                const double r = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                c1 = __ddiv_rn(dx, r); s1 = __ddiv_rn(dy, r);
            }
            col[i] = __dadd_rn(__dmul_rn(c1, dx), __dmul_rn(s1, dy));
            col[i + 1] = __dadd_rn(__dmul_rn(-s1, dx), __dmul_rn(c1, dy));
            cs[i] = c1; sn[i] = s1;
            const double sx = s[i], sy = s[i + 1];
            s[i] = __dadd_rn(__dmul_rn(c1, sx), __dmul_rn(s1, sy));
            const double s_next = __dadd_rn(__dmul_rn(-s1, sx), __dmul_rn(c1, sy));
            s[i + 1] = s_next;
            const double resid = __ddiv_rn(fabs(s_next), ctrl->normb);
            ctrl->resid = resid;
            ctrl->steps = i + 1;
            if (resid <= ctrl->tol) ctrl->done = 1;
        }
        __syncthreads();
        for (int k = threadIdx.x; k <= i + 1 && k < m; k += blockDim.x) H[(size_t)i * m + k] = col[k];
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) out[e] = mul(in[e], inv);
}

int gmres_scale_givens(size_t n, double* out, const double* in, int i, int m, const double* d_h, double* d_cs, double* d_sn,
                       double* d_s, double* d_H, GmresCtrl* ctrl)
{
    Ctx& c = ctx();
    ScopedTiming t("vec");
    const int grid = vec_grid(n, 256, 4);
    k_gmres_scale_givens<<<grid, 256, (size_t)(3 * i + 4) * sizeof(double), c.stream>>>(n, out, in, i, m, d_h, d_cs, d_sn, d_s, d_H, ctrl);
    count_launch();
    DB_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// MINRES pieces
// ------------------------------------------------------------------------------------------
struct OpLanczosOrth {
    double* q2; const double* q1; const double* a1p; double* b1; double a1; dd* part_out;
    __device__ bool begin() { a1 = *a1p; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double v = sub(q2[i], mul(a1, q1[i]));            // q2 -= a1*q1  :933
        q2[i] = v;
        acc[0] = dd_add_prod(acc[0], v, v);
    }
    __device__ void finish(const dd* acc)
    {
        if (part_out) { *part_out = acc[0]; return; }           // partitioned system: |q2|^2 summed over the ranks later
        *b1 = sqrt(dd_round(acc[0]));                           // :934
    }
};
int lanczos_orth(size_t n, double* q2, const double* q1, const double* d_a1, double* d_b1, dd* part_out)
{
    return launch_stream<1>(n, OpLanczosOrth{q2, q1, d_a1, d_b1, 0.0, part_out}, "minres.orth");
}

struct OpPLanczosOrth {
    double* t2; const double* t1; const double* a1p; double* q2; int pc; double omega; const double* d;
    double* b1sq; double a1; dd* part_out;
    __device__ bool begin() { a1 = *a1p; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double t = sub(t2[i], mul(a1, t1[i]));            // t2 -= a1*t1     :1000
        t2[i] = t;
        const double q = precond(pc, omega, t, d, i);           // M.Apply(A,q2,t2) :1001
        q2[i] = q;
        acc[0] = dd_add_prod(acc[0], q, t);                     // dot(q2,t2)      :1002
    }
    __device__ void finish(const dd* acc)
    {
        if (part_out) { *part_out = acc[0]; return; }
        *b1sq = dd_round(acc[0]);
    }
};
int planczos_orth(size_t n, double* t2, const double* t1, const double* d_a1, double* q2, int pc,
                  double omega, const double* d, double* d_b1sq, dd* part_out)
{
    return launch_stream<1>(n, OpPLanczosOrth{t2, t1, d_a1, q2, pc, omega, d, d_b1sq, 0.0, part_out}, "minres.orth");
}

struct OpDivScalar {
    double* out; const double* in; double s; int mode;
    __device__ bool begin() { if (mode) s = dvd(1., s); return true; }
    __device__ void elem(size_t i, dd*) { out[i] = mode ? mul(in[i], s) : dvd(in[i], s); }
    __device__ void finish(const dd*) {}
};
int vec_div_scalar(size_t n, double* out, const double* in, double s, int mode)
{
    return launch_stream<0>(n, OpDivScalar{out, in, s, mode}, "vec");
}

struct OpMinresUpdate {
    int mode; const double* q; const double* pm2; const double* pm1; double* p0; double* x;
    double r0, r1, r2, f, inv;
    __device__ bool begin() { inv = dvd(1, r2); return true; }
    __device__ void elem(size_t i, dd*)
    {
        double p;
        if (mode == 1) p = dvd(q[i], r0);                                        // :1097
        else if (mode == 2) p = dvd(sub(q[i], mul(r0, pm1[i])), r1);             // :1110
        else p = mul(sub(sub(q[i], mul(r0, pm2[i])), mul(r1, pm1[i])), inv);     // :1124
        p0[i] = p;
        x[i] = add(x[i], mul(f, p));                                             // :1128-1129
    }
    __device__ void finish(const dd*) {}
};
int minres_update(size_t n, int mode, const double* q, const double* pm2, const double* pm1, double* p0,
                  double* x, double r0, double r1, double r2, double f)
{
    return launch_stream<0>(n, OpMinresUpdate{mode, q, pm2, pm1, p0, x, r0, r1, r2, f, 0.0}, "minres.update");
}

// ------------------------------------------------------------------------------------------
// BiCGStab pieces
// ------------------------------------------------------------------------------------------
struct OpBicgP {
    bool first; double* p; double* phat; const double* r; const double* v; double beta, omega; int pc; double opc; const double* d;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd*)
    {
        const double pi = first ? r[i] : add(r[i], mul(beta, sub(p[i], mul(omega, v[i]))));   // :1216-1220
        p[i] = pi;
        phat[i] = precond(pc, opc, pi, d, i);                                                  // :1222
    }
    __device__ void finish(const dd*) {}
};
int bicg_p(size_t n, bool first, double* p, double* phat, const double* r, const double* v, double beta, double omega,
           int pc, double omega_pc, const double* d)
{
    return launch_stream<0>(n, OpBicgP{first, p, phat, r, v, beta, omega, pc, omega_pc, d}, "bicgstab.vec");
}

struct OpBicgS {
    double* s; double* shat; const double* r; const double* v; double alpha; int pc; double opc; const double* d; double* out;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double si = sub(r[i], mul(alpha, v[i]));                                         // :1225
        s[i] = si;
        shat[i] = precond(pc, opc, si, d, i);                                                  // :1232
        acc[0] = dd_add_prod(acc[0], si, si);
    }
    __device__ void finish(const dd* acc) { out[0] = dd_round(acc[0]); }
};
int bicg_s(size_t n, double* s, double* shat, const double* r, const double* v, double alpha, int pc, double omega_pc,
           const double* d, double* d_out)
{
    return launch_stream<1>(n, OpBicgS{s, shat, r, v, alpha, pc, omega_pc, d, d_out}, "bicgstab.vec");
}

struct OpBicgTs {
    const double* t; const double* s; double* out;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        acc[0] = dd_add_prod(acc[0], t[i], s[i]);
        acc[1] = dd_add_prod(acc[1], t[i], t[i]);
    }
    __device__ void finish(const dd* acc) { out[0] = dd_round(acc[0]); out[1] = dd_round(acc[1]); }
};
int bicg_ts(size_t n, const double* t, const double* s, double* d_out)
{
    return launch_stream<2>(n, OpBicgTs{t, s, d_out}, "bicgstab.vec");
}

struct OpBicgX {
    double* x; double* r; const double* phat; const double* shat; const double* s; const double* t; const double* rtilde;
    double alpha, omega; double* out;
    __device__ bool begin() { return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        x[i] = add(x[i], add(mul(alpha, phat[i]), mul(omega, shat[i])));                       // :1235
        const double ri = sub(s[i], mul(omega, t[i]));                                         // :1236
        r[i] = ri;
        acc[0] = dd_add_prod(acc[0], ri, ri);
        acc[1] = dd_add_prod(acc[1], rtilde[i], ri);
    }
    __device__ void finish(const dd* acc) { out[0] = dd_round(acc[0]); out[1] = dd_round(acc[1]); }
};
int bicg_x(size_t n, double* x, double* r, const double* phat, const double* shat, const double* s, const double* t,
           const double* rtilde, double alpha, double omega, double* d_out)
{
    return launch_stream<2>(n, OpBicgX{x, r, phat, shat, s, t, rtilde, alpha, omega, d_out}, "bicgstab.vec");
}

// ------------------------------------------------------------------------------------------
// GCR pieces
// ------------------------------------------------------------------------------------------
struct OpGcrOrth {
    double* vn; double* sn; double* vnacc; const double* vprev; const double* sprev; const double* vaccprev;
    const double* aprev; const double* vacc_cur; double* out; double a;
    __device__ bool begin() { a = vprev ? *aprev : 0.0; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        double vi = vn[i], wi = vnacc ? vnacc[i] : 0.0;
        if (vprev) {                                                                           // :1445-1448
            vi = sub(vi, mul(a, vprev[i])); vn[i] = vi;
            if (vnacc) { wi = sub(wi, mul(a, vaccprev[i])); vnacc[i] = wi; }
            sn[i] = sub(sn[i], mul(a, sprev[i]));
        }
        const double acc_v = vnacc ? wi : vi;
        acc[0] = vacc_cur ? dd_add_prod(acc[0], acc_v, vacc_cur[i]) : dd_add_prod(acc[0], acc_v, acc_v);
    }
    __device__ void finish(const dd* acc) { *out = vacc_cur ? dd_round(acc[0]) : sqrt(dd_round(acc[0])); }
};
int gcr_orth(size_t n, double* vn, double* sn, double* vnacc, const double* vprev, const double* sprev,
             const double* vaccprev, const double* d_aprev, const double* vacc_cur, double* d_out)
{
    return launch_stream<1>(n, OpGcrOrth{vn, sn, vnacc, vprev, sprev, vaccprev, d_aprev, vacc_cur, d_out, 0.0}, "gcr.vec");
}

struct OpGcrScale {
    double* vn; double* sn; double* vnacc; const double* racc; const double* betap; double* out; double beta;
    __device__ bool begin() { beta = *betap; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        const double vi = dvd(vn[i], beta);                                                    // :1452-1454
        vn[i] = vi;
        double wi = vi;
        if (vnacc) { wi = dvd(vnacc[i], beta); vnacc[i] = wi; }
        sn[i] = dvd(sn[i], beta);
        acc[0] = dd_add_prod(acc[0], racc[i], wi);                                             // :1455
    }
    __device__ void finish(const dd* acc) { *out = dd_round(acc[0]); }
};
int gcr_scale(size_t n, double* vn, double* sn, double* vnacc, const double* racc, const double* d_beta, double* d_gamma)
{
    return launch_stream<1>(n, OpGcrScale{vn, sn, vnacc, racc, d_beta, d_gamma, 0.0}, "gcr.vec");
}

struct OpGcrUpdate {
    double* x; double* r; double* racc; const double* sn; const double* vn; const double* vnacc; const double* gp;
    double* out; double g;
    __device__ bool begin() { g = *gp; return true; }
    __device__ void elem(size_t i, dd* acc)
    {
        x[i] = add(x[i], mul(g, sn[i]));                                                       // :1459
        double ri = sub(r[i], mul(g, vn[i]));                                                  // :1460
        r[i] = ri;
        if (racc) { ri = sub(racc[i], mul(g, vnacc[i])); racc[i] = ri; }                       // :1461
        acc[0] = dd_add_prod(acc[0], ri, ri);
    }
    __device__ void finish(const dd* acc) { *out = dd_round(acc[0]); }
};
int gcr_update(size_t n, double* x, double* r, double* racc, const double* sn, const double* vn, const double* vnacc,
               const double* d_gamma, double* d_rr)
{
    return launch_stream<1>(n, OpGcrUpdate{x, r, racc, sn, vn, vnacc, d_gamma, d_rr, 0.0}, "gcr.vec");
}

}  // namespace db200

// ------------------------------------------------------------------------------------------
// C ABI: vectors
// ------------------------------------------------------------------------------------------
using namespace db200;

static int fetch_scalar(const double* d_s, double* out)
{
    Ctx& c = ctx();
    DB_CUDA(cudaMemcpyAsync(c.h_scalars, d_s, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    DB_CUDA(cudaStreamSynchronize(c.stream));
    *out = c.h_scalars[0];
    return 0;
}

extern "C" {

int drops_b200_vector_create(size_t n, drops_b200_vector_t* out)
{
    DB_CHECK(require_ctx());
    DB_ARG(out != nullptr, "null out");
    drops_b200_vector_s* v = new drops_b200_vector_s;
    v->n = n;
    int s = dev_alloc(&v->d, n);
    if (s < 0) { delete v; *out = nullptr; return s; }
    cudaError_t e = cudaMemsetAsync(v->d, 0, (n ? n : 1) * sizeof(double), ctx().stream);
    if (e != cudaSuccess) { dev_free(v->d); delete v; *out = nullptr; return cuda_fail(e, "memset", __FILE__, __LINE__); }
    *out = v;
    return 0;
}

int drops_b200_vector_destroy(drops_b200_vector_t v)
{
    if (!v) return 0;
    if (ctx().inited) cudaStreamSynchronize(ctx().stream);
    if (v->owns) dev_free(v->d);
    delete v;
    return 0;
}

int drops_b200_vector_size(drops_b200_vector_t v, size_t* n)
{
    DB_ARG(v != nullptr && n != nullptr, "null argument");
    *n = v->n;
    return 0;
}

int drops_b200_vector_upload(drops_b200_vector_t v, const double* host)
{
    DB_CHECK(require_ctx());
    DB_ARG(v != nullptr && (host != nullptr || v->n == 0), "null argument");
    if (v->n) DB_CUDA(cudaMemcpyAsync(v->d, host, v->n * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

int drops_b200_vector_download(drops_b200_vector_t v, double* host)
{
    DB_CHECK(require_ctx());
    DB_ARG(v != nullptr && (host != nullptr || v->n == 0), "null argument");
    if (v->n) DB_CUDA(cudaMemcpyAsync(host, v->d, v->n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    DB_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

int drops_b200_vector_fill(drops_b200_vector_t v, double value)
{
    DB_CHECK(require_ctx());
    DB_ARG(v != nullptr, "null vector");
    return vec_fill(v->n, v->d, value);
}

int drops_b200_vector_fill_hash(drops_b200_vector_t v)
{
    DB_CHECK(require_ctx());
    DB_ARG(v != nullptr, "null vector");
    return vec_fill_hash(v->n, v->d);
}

int drops_b200_vector_copy(drops_b200_vector_t dst, drops_b200_vector_t src)
{
    DB_CHECK(require_ctx());
    DB_ARG(dst != nullptr && src != nullptr && dst->n == src->n, "size mismatch");
    return vec_copy(dst->n, dst->d, src->d);
}

int drops_b200_vector_ptr(drops_b200_vector_t v, double** dptr)
{
    DB_ARG(v != nullptr && dptr != nullptr, "null argument");
    *dptr = v->d;
    return 0;
}

int drops_b200_dot(drops_b200_vector_t x, drops_b200_vector_t y, double* result)
{
    DB_CHECK(require_ctx());
    DB_ARG(x && y && result && x->n == y->n, "dot: incompatible dimensions");
    DB_CHECK(vec_dot(x->n, x->d, y->d, ctx().scratch_scalars));
    return fetch_scalar(ctx().scratch_scalars, result);
}

int drops_b200_norm_sq(drops_b200_vector_t x, double* result) { return drops_b200_dot(x, x, result); }

int drops_b200_norm(drops_b200_vector_t x, double* result)
{
    double s = 0.0;
    DB_CHECK(drops_b200_dot(x, x, &s));
    *result = sqrt(s);
    return 0;
}

int drops_b200_supnorm(drops_b200_vector_t x, double* result)
{
    DB_CHECK(require_ctx());
    DB_ARG(x && result, "null argument");
    DB_CHECK(vec_supnorm(x->n, x->d, ctx().scratch_scalars));
    return fetch_scalar(ctx().scratch_scalars, result);
}

int drops_b200_axpy(double a, drops_b200_vector_t x, drops_b200_vector_t y)
{
    DB_CHECK(require_ctx());
    DB_ARG(x && y && x->n == y->n, "axpy: incompatible dimensions");
    return vec_axpy(x->n, a, x->d, y->d);
}

int drops_b200_z_xpay(drops_b200_vector_t z, drops_b200_vector_t x, double a, drops_b200_vector_t y)
{
    DB_CHECK(require_ctx());
    DB_ARG(z && x && y && z->n == x->n && z->n == y->n, "z_xpay: incompatible dimensions");
    return vec_z_xpay(z->n, z->d, x->d, a, y->d);
}

int drops_b200_z_xpaypby2(drops_b200_vector_t z, drops_b200_vector_t x, double a, drops_b200_vector_t y,
                          double b, drops_b200_vector_t y2)
{
    DB_CHECK(require_ctx());
    DB_ARG(z && x && y && y2 && z->n == x->n && z->n == y->n && z->n == y2->n, "z_xpaypby2: incompatible dimensions");
    return vec_z_xpaypby2(z->n, z->d, x->d, a, y->d, b, y2->d);
}

}  // extern "C"

// krylov_ext.cu -- the callers one step outside the PCG/GMRES/MINRES core:
//   * BlockMatrixBaseCL / CompositeMatrixBaseCL products   num/spblockmat.h:225-275, :376-404
//   * BBTDiag                                             num/precond.h:56-75
//   * ChebyshevPcCL / ChebyshevBBTPcCL                     num/precond.h:1080-1096, :1359-1408, :1463-1567
//   * PCG with a preconditioner that is a procedure (Chebyshev) and PCGNE  num/krylovsolver.h:107-179, :199-252
//   * IDR(s)                                              num/krylovsolver.h:1579-1698
// Everything here is y = A x / y = A^T x through the same SELL kernels plus fused streaming kernels.
#include "krylov.cuh"
#include "spmv.cuh"
#include "vecops.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

namespace db200 {

namespace {

int fetch_sync(const void* dsrc, void* hdst, size_t bytes)
{
    if (bytes % sizeof(double) == 0)
        return fetch_small(static_cast<const double*>(dsrc), static_cast<double*>(hdst), bytes / sizeof(double));
    Ctx& c = ctx();
    DB_CUDA(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, c.stream));
    DB_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int ensure_small_ext(drops_b200_matrix_s* M, size_t doubles)
{
    if (doubles > M->small_alloc) {
        dev_free(M->d_small); M->d_small = nullptr; M->small_alloc = 0;
        DB_CHECK(dev_alloc(&M->d_small, doubles));
        M->small_alloc = doubles;
    }
    return 0;
}

// ret_i = sum_l B_il * B_il, left to right from 0 (precond.h:61-75)
__global__ void k_bbt_diag(long long rows, const int64_t* __restrict__ rowptr, const double* __restrict__ val,
                           double* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0.0;
    for (int64_t l = rowptr[i]; l < rowptr[i + 1]; ++l) {
        const double b = val[l];
        s = __dadd_rn(s, __dmul_rn(b, b));
    }
    out[i] = s;
}

}  // namespace

// diag(B B^T), cached on the handle until the values change
int bbt_diag(drops_b200_matrix_s* M, const double** out)
{
    if (!M->bbt_diag || M->bbt_epoch != M->values_epoch) {
        if (!M->bbt_diag) DB_CHECK(dev_alloc(&M->bbt_diag, M->csr.rows));
        const long long rows = (long long)M->csr.rows;
        if (rows > 0) {
            k_bbt_diag<<<(unsigned)((rows + 255) / 256), 256, 0, ctx().stream>>>(rows, M->csr.rowptr, M->csr.val,
                                                                                  M->bbt_diag);
            count_launch();
            DB_CUDA(cudaGetLastError());
        }
        M->bbt_epoch = M->values_epoch;
    }
    *out = M->bbt_diag;
    return 0;
}

// y = op(A) x with an optional epilogue; op = 0: A, 1: A^T
static int apply_op(drops_b200_matrix_s* M, int transposed, SpmvArgs a)
{
    if (transposed) DB_CHECK(build_transpose(M));
    return spmv_apply(M, transposed != 0, a);
}

// The operator the generic solvers iterate on: A (square) or the composite A A^T of PCGNE
// (CompositeMatrixBaseCL(&A, TRANSP_MUL, ex, &A, MUL, ex), precond.h:1549, :1562).
struct Oper {
    drops_b200_matrix_s* M;
    bool ne;
    size_t n() const { return M->csr.rows; }
    // y = b - Op x (resid) or y = Op x; tmp: cols entries (only used when ne)
    int apply(const double* x, double* y, const double* b_for_resid, double* tmp) const
    {
        SpmvArgs a;
        if (!ne) {
            a.x = x; a.y = y;
            if (b_for_resid) { a.epilogue = EPI_RESID; a.aux = b_for_resid; }
            return spmv_apply(M, false, a);
        }
        SpmvArgs t; t.x = x; t.y = tmp;
        DB_CHECK(apply_op(M, 1, t));
        a.x = tmp; a.y = y;
        if (b_for_resid) { a.epilogue = EPI_RESID; a.aux = b_for_resid; }
        return spmv_apply(M, false, a);
    }
};

// ------------------------------------------------------------------------------------------
// MaxEigenvalueScaledMatrix (precond.h:1080-1096): 10 steps of the power method on D^-1 Op,
// started from the normalised vector of ones.  Returns scale * lambda, cached per operator.
// ------------------------------------------------------------------------------------------
static int cheb_diag(drops_b200_matrix_s* M, bool ne, const double** d)
{
    if (ne) return bbt_diag(M, d);
    if (!M->diag_ok) {
        set_error("ChebyshevPcCL: at least one row stores no diagonal entry");
        return DROPS_B200_E_NODIAG;
    }
    *d = M->diag;
    return 0;
}

int cheb_lambda(drops_b200_matrix_s* M, bool ne, double* out)
{
    const int k = ne ? 1 : 0;
    if (M->cheb_lambda_epoch[k] == M->values_epoch && M->cheb_lambda_scale[k] == M->cheb_scale[k]) {
        *out = M->cheb_lambda[k];
        return 0;
    }
    Ctx& c = ctx();
    const size_t n = M->csr.rows;
    const double* d = nullptr;
    DB_CHECK(cheb_diag(M, ne, &d));
    // work vectors 12.. are reserved for the preconditioner (the solvers use 0..11)
    double* y = work_vec(M, 12); double* t = work_vec(M, 13); double* tmp = work_vec(M, 14);
    if (!y || !t || !tmp) return DROPS_B200_E_NOMEM;
    double* d_s = c.scratch_scalars;
    double* hs = c.h_scalars;
    // y = D; y /= D; y /= Norm(y)
    DB_CHECK(vec_jacobi(n, 1.0, d, d, y));                          // (1*d)/d
    DB_CHECK(vec_dot(n, y, y, d_s));
    DB_CHECK(fetch_sync(d_s, hs, sizeof(double)));
    DB_CHECK(vec_div_scalar(n, y, y, std::sqrt(hs[0]), 0));
    double lambda = 0.0;
    for (int i = 0; i < 10; ++i) {
        // y = (Op*y)/D
        if (!ne) {
            SpmvArgs a; a.x = y; a.y = t; a.epilogue = EPI_JACOBI; a.aux = d; a.omega = 1.0;
            DB_CHECK(spmv_apply(M, false, a));
        } else {
            SpmvArgs a0; a0.x = y; a0.y = tmp;
            DB_CHECK(apply_op(M, 1, a0));
            SpmvArgs a; a.x = tmp; a.y = t; a.epilogue = EPI_JACOBI; a.aux = d; a.omega = 1.0;
            DB_CHECK(spmv_apply(M, false, a));
        }
        DB_CHECK(vec_dot(n, t, t, d_s));
        DB_CHECK(fetch_sync(d_s, hs, sizeof(double)));
        lambda = std::sqrt(hs[0]);
        DB_CHECK(vec_div_scalar(n, y, t, lambda, 0));               // y /= lambda
    }
    M->cheb_lambda[k] = M->cheb_scale[k] * lambda;
    M->cheb_lambda_scale[k] = M->cheb_scale[k];
    M->cheb_lambda_epoch[k] = M->values_epoch;
    *out = M->cheb_lambda[k];
    return 0;
}

// ------------------------------------------------------------------------------------------
// Chebyshev<false> (precond.h:1359-1408): x = p_degree(D^-1 Op) D^-1 b, start vector 0.
// ------------------------------------------------------------------------------------------
int cheb_apply(drops_b200_matrix_s* M, bool ne, double* x, const double* b)
{
    const int k = ne ? 1 : 0;
    const int degree = M->cheb_degree[k];
    if (degree == 0) return 0;                                       // :1362-1363: x untouched
    const size_t n = M->csr.rows;
    const double* d = nullptr;
    DB_CHECK(cheb_diag(M, ne, &d));
    double lmax = 0.0;
    DB_CHECK(cheb_lambda(M, ne, &lmax));
    double* W = work_vec(M, 12); double* res = work_vec(M, 13); double* tmp = work_vec(M, 14);
    if (!W || !res || !tmp) return DROPS_B200_E_NOMEM;
    const double alpha = lmax / M->cheb_ratio[k];
    const double beta = lmax;
    const double delta = 2.0 / (beta - alpha);
    const double theta = 0.5 * (beta + alpha);
    const double s1 = theta * delta;
    const double one_over_theta = 1.0 / theta;
    DB_CHECK(cheb_first(n, W, x, b, d, one_over_theta));             // res = b; W = invDiag*res/theta; x = W
    double rhok = 1.0 / s1;
    Oper op{M, ne};
    for (int it = 0; it < degree - 1; ++it) {
        DB_CHECK(op.apply(x, res, b, tmp));                          // res = b - A*x
        const double rhokp1 = 1.0 / (2.0 * s1 - rhok);
        const double dt1 = rhokp1 * rhok;
        const double dt2 = 2.0 * rhokp1 * delta;
        rhok = rhokp1;
        DB_CHECK(cheb_step(n, W, x, res, d, dt1, dt2));
    }
    return 0;
}

// x = M^-1 b for any preconditioner kind (x must not alias b for Chebyshev)
int precond_apply(drops_b200_matrix_s* M, int pc, double omega, bool ne, const double* b, double* x)
{
    const size_t n = M->csr.rows;
    if (pc == DROPS_B200_PC_DUMMY) return vec_copy(n, x, b);
    if (pc == DROPS_B200_PC_JACOBI) {
        const double* d = nullptr;
        DB_CHECK(cheb_diag(M, ne, &d));
        return vec_jacobi(n, omega, d, b, x);
    }
    if (pc == DROPS_B200_PC_CHEBYSHEV) return cheb_apply(M, ne, x, b);
    if (pc == DROPS_B200_PC_DIAG) {                       // DiagPcCL::Apply: x = D_ * b   precond.h:626-631
        if (M->diag_pc == nullptr) { set_error("DiagPcCL: no vector set (drops_b200_matrix_set_diag_pc)"); return DROPS_B200_E_ARG; }
        return vec_mul(n, x, M->diag_pc, b);
    }
    set_error("unknown preconditioner %d", pc);
    return DROPS_B200_E_ARG;
}

// ------------------------------------------------------------------------------------------
// PCG on Op = A with a procedural preconditioner, and PCGNE (Op = A A^T, krylovsolver.h:199-252).
// Same device-side scalar block as the fast PCG path: the host looks at the stop flag every
// "pcg.check_every" iterations only.
//   PCGNE per iteration: pt = A^T q; lambda = |pt|^2 (:233-234); u += alpha q; r -= alpha (A pt);
//   z = M^-1 r; resid; rho = z.r; q = z + (rho/rho_1) q  -- PCG on A A^T with q.(A A^T q) taken
//   as |A^T q|^2.
// ------------------------------------------------------------------------------------------
int pcg_general_device(drops_b200_matrix_s* M, bool ne, double* x, const double* b, int pc, double omega,
                       int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    DB_ARG(M != nullptr, "null matrix");
    DB_ARG(ne || M->square, "PCG needs a square matrix");
    DB_ARG(max_iter && tol, "null max_iter/tol");
    DB_ARG(pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI || pc == DROPS_B200_PC_CHEBYSHEV || pc == DROPS_B200_PC_DIAG,
           "unknown preconditioner");
    DB_ARG(pc != DROPS_B200_PC_DIAG || M->diag_pc != nullptr, "DiagPcCL: no vector set (drops_b200_matrix_set_diag_pc)");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    const size_t n = M->csr.rows;
    // preconditioners that are a procedure (ChebyshevPcCL) or carry their own vector (DiagPcCL): z = M^-1 r is a separate
    // step after the fused update, which then runs as if unpreconditioned
    const bool external = pc == DROPS_B200_PC_CHEBYSHEV || pc == DROPS_B200_PC_DIAG;
    const double* d = M->diag;
    if (pc == DROPS_B200_PC_JACOBI || pc == DROPS_B200_PC_CHEBYSHEV) DB_CHECK(cheb_diag(M, ne, &d));
    if (pc == DROPS_B200_PC_CHEBYSHEV) { double l; DB_CHECK(cheb_lambda(M, ne, &l)); }    // M.SetDiag(...)        :122, :204
    const int want_hist = hist ? std::max(hist_cap, 0) : 0;
    if (!M->d_scalars) DB_CHECK(dev_alloc((SolverScalars**)&M->d_scalars, 1));
    if (std::max(want_hist, 1) > M->hist_alloc) {
        dev_free(M->d_hist); M->d_hist = nullptr; M->hist_alloc = 0;
        DB_CHECK(dev_alloc(&M->d_hist, (size_t)std::max(want_hist, 1)));
        M->hist_alloc = std::max(want_hist, 1);
    }
    SolverScalars* S = (SolverScalars*)M->d_scalars;
    double* r = work_vec(M, 0); double* p = work_vec(M, 1); double* q = work_vec(M, 2);
    double* z = work_vec(M, 3); double* pt = work_vec(M, 4);
    if (!r || !p || !q || !z || !pt) return DROPS_B200_E_NOMEM;

    SolverScalars h;
    std::memset(&h, 0, sizeof h);
    h.tol = *tol; h.relative = relative; h.max_iter = *max_iter; h.omega = omega;
    h.pc = external ? DROPS_B200_PC_DUMMY : pc;      // the fused kernels then compute z = r, rho = r.r
    h.hist_cap = want_hist;
    DB_CUDA(cudaMemcpyAsync(S, &h, sizeof h, cudaMemcpyHostToDevice, st));
    double* dh = want_hist ? M->d_hist : nullptr;

    Oper op{M, ne};
    DB_CHECK(op.apply(x, r, b, pt));                                 // r = b - Op x          :125, :205-207
    DB_CHECK(pcg_setup(n, r, b, d, p, S, dh));                       // normb, resid, z, rho, p = z
    if (external) {
        DB_CHECK(precond_apply(M, pc, omega, ne, r, z));
        DB_CHECK(pcg_set_rho(n, z, r, S));
        DB_CHECK(vec_copy(n, p, z));
    }
    int check_every = (int)opt("pcg.check_every", 16.0);
    if (check_every < 1) check_every = 1;
    struct Flags { int iter, done, converged, pad; };
    Flags fl;
    DB_CHECK(fetch_sync(&S->iter, &fl, sizeof fl));
    const int total = std::max(*max_iter, 0);
    int launched = 0;
    while (!fl.done && launched < total) {
        const int chunk = std::min(check_every, total - launched);
        for (int k = 0; k < chunk; ++k) {
            if (!ne) {
                SpmvArgs s; s.x = p; s.y = q; s.dotv = p; s.dot_out = &S->lambda; s.done = &S->done;
                DB_CHECK(spmv_apply(M, false, s));                   // q = A p, lambda = q.p :149-151
            } else {
                SpmvArgs t; t.x = p; t.y = pt; t.done = &S->done;
                DB_CHECK(apply_op(M, 1, t));                         // pt = A^T q            :233
                DB_CHECK(vec_dot(M->csr.cols, pt, pt, &S->lambda));             // |pt|^2                :234
                SpmvArgs s; s.x = pt; s.y = q; s.done = &S->done;
                DB_CHECK(spmv_apply(M, false, s));                   // A*ptacc               :236
            }
            DB_CHECK(pcg_update(n, x, r, p, q, d, S, dh));
            if (external) {
                DB_CHECK(precond_apply(M, pc, omega, ne, r, z));
                DB_CHECK(pcg_set_rho(n, z, r, S));
                DB_CHECK(pcg_direction_z(n, p, z, S));
            } else {
                DB_CHECK(pcg_direction(n, p, r, d, S));
            }
        }
        launched += chunk;
        DB_CHECK(fetch_sync(&S->iter, &fl, sizeof fl));
    }
    DB_CHECK(fetch_sync(S, &h, sizeof h));
    if (hist && want_hist > 0) {
        const int cnt = std::min(want_hist, h.iter + 1);
        DB_CHECK(fetch_sync(M->d_hist, hist, (size_t)cnt * sizeof(double)));
    }
    *tol = h.resid;
    if (h.converged) *max_iter = h.iter;
    return h.converged ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// IDR(s) -- krylovsolver.h:1579-1698.  Small dense work (M, f, c) on the host as in the reference,
// every vector operation a device kernel.  P: the s shadow vectors BEFORE orthonormalisation
// (host, s*n doubles) or nullptr: drawn like the reference does (srand(time(0)); 2*rand()/RAND_MAX-1).
// Return value and outputs follow the reference literally, including its final test (:1689-1696):
// 1 unless tol > |r|/normb at exit; *tol = |r| (absolute), *max_iter = iterations.
// ------------------------------------------------------------------------------------------
int idrs_device(drops_b200_matrix_s* M, double* x, const double* b, int pc, double omega_pc, int s, double omega_bound,
                const double* P_host, int* max_iter, double* tol, int relative)
{
    DB_ARG(M != nullptr && M->square, "IDR(s) needs a square matrix");
    DB_ARG(max_iter && tol, "null max_iter/tol");
    DB_ARG(s >= 1 && s <= 64, "IDR(s): 1 <= s <= 64");
    DB_ARG(pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI, "IDR(s): preconditioner must be DummyPcCL or JACPcCL");
    if (pc == DROPS_B200_PC_JACOBI && !M->diag_ok) {
        set_error("JACPcCL: at least one row stores no diagonal entry");
        return DROPS_B200_E_NODIAG;
    }
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    const size_t n = M->csr.rows;
    // work vectors: 0 v, 1 t, 2 resid, 3.. P (s), G (s), U (s)
    std::vector<double*> W(3 + 3 * (size_t)s);
    for (size_t i = 0; i < W.size(); ++i) { W[i] = work_vec(M, (int)(16 + i)); if (!W[i]) return DROPS_B200_E_NOMEM; }
    double *v = W[0], *t = W[1], *resid = W[2];
    double** P = &W[3]; double** G = &W[3 + s]; double** U = &W[3 + 2 * s];
    DB_CHECK(ensure_small_ext(M, (size_t)s + 8));
    double* d_s = M->d_small;
    std::vector<double> hsv((size_t)s + 8);
    double* hs = hsv.data();
    auto dotn = [&](const double* a, const double* bb, double* out) -> int {
        DB_CHECK(vec_dot(n, a, bb, d_s));
        return fetch_sync(d_s, out, sizeof(double));
    };

    double omega = 1.0, tmp = 0.0;
    DB_CHECK(dotn(b, b, &tmp));
    double normb = std::sqrt(tmp);                                                  // :1591
    if (normb == 0.0 || !relative) normb = 1.0;
    SpmvArgs a0; a0.x = x; a0.y = resid; a0.epilogue = EPI_RESID; a0.aux = b;       // resid = rhs - A*x   :1595
    DB_CHECK(spmv_apply(M, false, a0));
    DB_CHECK(dotn(resid, resid, &tmp));
    double normres = std::sqrt(tmp);
    if (normres / normb <= *tol) { *tol = normres; *max_iter = 0; return 1; }       // :1597-1601

    // shadow space                                                                  :1603-1617
    {
        std::vector<double> gen;
        const double* src = P_host;
        if (!src) {
            gen.resize((size_t)s * n);
            std::srand((unsigned)std::time(nullptr));
            for (size_t i = 0; i < gen.size(); ++i) gen[i] = (double)(2.0 * std::rand()) / RAND_MAX - 1.0;
            src = gen.data();
        }
        for (int i = 0; i < s; ++i)
            if (n) DB_CUDA(cudaMemcpyAsync(P[i], src + (size_t)i * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
        DB_CUDA(cudaStreamSynchronize(st));
    }
    for (int k = 0; k < s; ++k) {                                                   // mod. Gram-Schmidt
        for (int j = 0; j < k; ++j) {
            double sm; DB_CHECK(dotn(P[j], P[k], &sm));
            DB_CHECK(vec_axpy(n, -sm, P[j], P[k]));                                 // P[k] -= sm*P[j]
        }
        double nn; DB_CHECK(dotn(P[k], P[k], &nn));
        DB_CHECK(vec_div_scalar(n, P[k], P[k], std::sqrt(nn), 0));
    }
    for (int i = 0; i < s; ++i) { DB_CHECK(vec_fill(n, G[i], 0.0)); DB_CHECK(vec_fill(n, U[i], 0.0)); }
    std::vector<double> Mm((size_t)s * s, 0.0), f((size_t)s, 0.0), cc((size_t)s, 0.0);
    auto Mat = [&](int i, int j) -> double& { return Mm[(size_t)i * s + j]; };
    for (int i = 0; i < s; ++i) Mat(i, i) = 1.0;
    const double* d = M->diag;
    int it = 0;
    const double tol_in = *tol;
    while (normres / normb > tol_in && it < *max_iter) {
        DB_CHECK(multi_dot(n, resid, P, s, d_s));                                   // f = P^T resid       :1628
        DB_CHECK(fetch_sync(d_s, f.data(), (size_t)s * sizeof(double)));
        for (int k = 0; k < s; ++k) {
            for (int j = 0; j < s - k; ++j) {                                       // M[k:s,k:s] c = f[k:s]
                double cs = f[k + j];
                for (int l = 0; l < j; ++l) cs -= Mat(k + j, k + l) * cc[l];
                cc[j] = cs / Mat(k + j, k + j);
            }
            DB_CHECK(vec_copy(n, v, resid));                                        // v = resid           :1639
            for (int j = 0; j < s - k; ++j) DB_CHECK(vec_axpy(n, -cc[j], G[k + j], v));
            if (pc == DROPS_B200_PC_JACOBI) DB_CHECK(vec_jacobi(n, omega_pc, d, v, v));   // pc.Apply(A, v, v)
            DB_CHECK(vec_axpby(n, cc[0], U[k], omega, v));                          // U[k] = c0 U[k] + omega v
            for (int j = 1; j < s - k; ++j) DB_CHECK(vec_axpy(n, cc[j], U[k + j], U[k]));
            SpmvArgs ag; ag.x = U[k]; ag.y = G[k];                                  // G[k] = A U[k]       :1646
            DB_CHECK(spmv_apply(M, false, ag));
            for (int i = 0; i < k; ++i) {                                           // bi-orthogonalise    :1648-1652
                double pg; DB_CHECK(dotn(P[i], G[k], &pg));
                const double al = pg / Mat(i, i);
                DB_CHECK(vec_axpy(n, -al, G[i], G[k]));
                DB_CHECK(vec_axpy(n, -al, U[i], U[k]));
            }
            DB_CHECK(multi_dot(n, G[k], P + k, s - k, d_s));                        // M(k+j,k) = P[k+j].G[k]
            DB_CHECK(fetch_sync(d_s, hs, (size_t)(s - k) * sizeof(double)));
            for (int j = 0; j < s - k; ++j) Mat(k + j, k) = hs[j];
            if (Mat(k, k) == 0.0) {
                set_error("IDR(s): M(k,k) ==0");
                return DROPS_B200_E_ARG;
            }
            const double beta = f[k] / Mat(k, k);
            DB_CHECK(vec_axpy(n, -beta, G[k], resid));                              // resid -= beta G[k]
            DB_CHECK(vec_axpy(n, beta, U[k], x));                                   // x += beta U[k]
            DB_CHECK(dotn(resid, resid, &tmp));
            normres = std::sqrt(tmp);
            ++it;
            if (normres / normb <= tol_in) break;
            if (k + 1 < s)
                for (int j = 1; j < s - k; ++j) f[k + j] -= beta * Mat(k + j, k);
        }
        if (normres / normb <= tol_in) break;
        // entering G+                                                                :1672-1685
        if (pc == DROPS_B200_PC_JACOBI) DB_CHECK(vec_jacobi(n, omega_pc, d, resid, v));
        else DB_CHECK(vec_copy(n, v, resid));
        SpmvArgs at; at.x = v; at.y = t;
        DB_CHECK(spmv_apply(M, false, at));
        DB_CHECK(bicg_ts(n, t, resid, d_s));                                        // t.resid, t.t
        DB_CHECK(fetch_sync(d_s, hs, 2 * sizeof(double)));
        const double tn = std::sqrt(hs[1]), tr = hs[0];
        omega = tr / (tn * tn);
        const double rho = std::fabs(tr) / (tn * normres);
        if (rho < omega_bound) omega *= omega_bound / rho;
        if (omega == 0.0) {
            set_error("IDR(s): omega ==0");
            return DROPS_B200_E_ARG;
        }
        DB_CHECK(vec_axpy(n, -omega, t, resid));
        DB_CHECK(vec_axpy(n, omega, v, x));
        DB_CHECK(dotn(resid, resid, &tmp));
        normres = std::sqrt(tmp);
        ++it;
    }
    *max_iter = it;
    const bool quirk_false = tol_in > normres / normb;                              // :1689-1693
    *tol = normres;
    return quirk_false ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// BlockMatrixBaseCL operator* (spblockmat.h:225-267) / CompositeMatrixBaseCL operator* (:376-397)
// ------------------------------------------------------------------------------------------
static size_t op_rows(const drops_b200_matrix_s* A, int op) { return op == 0 ? A->csr.rows : A->csr.cols; }
static size_t op_cols(const drops_b200_matrix_s* A, int op) { return op == 0 ? A->csr.cols : A->csr.rows; }

int block_dims(const drops_b200_matrix_t* blk, const int* ops, size_t rows[2], size_t cols[2])
{
    // num_rows(block_row) / num_cols(block_col), spblockmat.h:112-190: the first non-null block decides
    for (int br = 0; br < 2; ++br) {
        const int b0 = 2 * br, b1 = 2 * br + 1;
        if (blk[b0]) rows[br] = op_rows(blk[b0], ops[b0]);
        else if (blk[b1]) rows[br] = op_rows(blk[b1], ops[b1]);
        else { set_error("BlockMatrixBaseCL::num_rows: All pointers are 0."); return DROPS_B200_E_ARG; }
    }
    for (int bc = 0; bc < 2; ++bc) {
        const int b0 = bc, b1 = bc + 2;
        if (blk[b0]) cols[bc] = op_cols(blk[b0], ops[b0]);
        else if (blk[b1]) cols[bc] = op_cols(blk[b1], ops[b1]);
        else { set_error("BlockMatrixBaseCL::num_cols: All pointers are 0."); return DROPS_B200_E_ARG; }
    }
    for (int b = 0; b < 4; ++b) {
        if (!blk[b]) continue;
        if (op_rows(blk[b], ops[b]) != rows[b / 2] || op_cols(blk[b], ops[b]) != cols[b % 2]) {
            set_error("BlockMatrixBaseCL: block %d has incompatible dimensions", b);
            return DROPS_B200_E_ARG;
        }
    }
    return 0;
}

int block_apply(const drops_b200_matrix_t* blk, const int* ops, const double* x, double* y)
{
    size_t rows[2], cols[2];
    DB_CHECK(block_dims(blk, ops, rows, cols));
    const double* xs[2] = {x, x + cols[0]};
    double* ys[2] = {y, y + rows[0]};
    for (int br = 0; br < 2; ++br) {
        const int b0 = 2 * br, b1 = 2 * br + 1;
        if (blk[b0]) {                                                // r = op(block) * x0
            SpmvArgs a; a.x = xs[0]; a.y = ys[br];
            DB_CHECK(apply_op(blk[b0], ops[b0], a));
        } else {
            DB_CHECK(vec_fill(rows[br], ys[br], 0.0));                // Vec r(n): zeros
        }
        if (blk[b1]) {                                                // r += op(block) * x1
            SpmvArgs a; a.x = xs[1]; a.y = ys[br];
            a.epilogue = EPI_AXMBY; a.aux = ys[br]; a.beta_val = -1.0;   // s - (-1)*r = r + s
            DB_CHECK(apply_op(blk[b1], ops[b1], a));
        }
    }
    return 0;
}

}  // namespace db200

using namespace db200;

namespace {
struct Bridge {
    double* dx = nullptr; double* dy = nullptr;
    ~Bridge() { dev_free(dx); dev_free(dy); }
    int up(size_t nx, const double* x, size_t ny, const double* y_or_null)
    {
        cudaStream_t st = ctx().stream;
        DB_CHECK(dev_alloc(&dx, nx));
        DB_CHECK(dev_alloc(&dy, ny));
        if (nx) DB_CUDA(cudaMemcpyAsync(dx, x, nx * sizeof(double), cudaMemcpyHostToDevice, st));
        if (ny && y_or_null) DB_CUDA(cudaMemcpyAsync(dy, y_or_null, ny * sizeof(double), cudaMemcpyHostToDevice, st));
        return 0;
    }
    int down(size_t n, const double* dsrc, double* h)
    {
        cudaStream_t st = ctx().stream;
        if (n) DB_CUDA(cudaMemcpyAsync(h, dsrc, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        DB_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
};

int check_pc_system(drops_b200_matrix_t A, int pc, bool ne)
{
    DB_ARG(A != nullptr, "null matrix");
    DB_ARG(pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI || pc == DROPS_B200_PC_CHEBYSHEV || pc == DROPS_B200_PC_DIAG,
           "unknown preconditioner");
    if (pc == DROPS_B200_PC_DIAG) {
        DB_ARG(!ne && A->diag_pc != nullptr, "DiagPcCL: no vector set (drops_b200_matrix_set_diag_pc)");
        return 0;
    }
    if (!ne && pc != DROPS_B200_PC_DUMMY && !A->diag_ok) {
        set_error("preconditioner needs a stored diagonal entry in every row");
        return DROPS_B200_E_NODIAG;
    }
    return 0;
}
}  // namespace

extern "C" {

int drops_b200_matrix_set_chebyshev(drops_b200_matrix_t A, int normal_eq, double lambda_scale, int degree, double eig_ratio)
{
    DB_ARG(A != nullptr, "null matrix");
    DB_ARG(degree >= 0 && eig_ratio > 0.0, "ChebyshevPcCL: degree >= 0, ratio > 0");
    const int k = normal_eq ? 1 : 0;
    A->cheb_scale[k] = lambda_scale; A->cheb_degree[k] = degree; A->cheb_ratio[k] = eig_ratio;
    return 0;
}

int drops_b200_matrix_set_diag_pc(drops_b200_matrix_t A, const double* D)
{
    DB_CHECK(require_ctx());
    DB_ARG(A != nullptr, "null matrix");
    if (D == nullptr) { dev_free(A->diag_pc); A->diag_pc = nullptr; return 0; }
    if (!A->diag_pc) DB_CHECK(dev_alloc(&A->diag_pc, A->csr.rows));
    cudaStream_t st = ctx().stream;
    if (A->csr.rows) DB_CUDA(cudaMemcpyAsync(A->diag_pc, D, A->csr.rows * sizeof(double), cudaMemcpyHostToDevice, st));
    DB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int drops_b200_diagpc_apply(drops_b200_vector_t D, drops_b200_vector_t b, drops_b200_vector_t x)
{
    DB_CHECK(require_ctx());
    DB_ARG(D && b && x && D->n == b->n && x->n == b->n, "DiagPcCL: incompatible dimensions");     // precond.h:629
    return vec_mul(b->n, x->d, D->d, b->d);
}

int drops_b200_chebyshev_lambda_max(drops_b200_matrix_t A, int normal_eq, double* lambda_max)
{
    DB_CHECK(require_ctx());
    DB_ARG(A != nullptr && lambda_max != nullptr, "null argument");
    DB_ARG(normal_eq || A->square, "ChebyshevPcCL needs a square matrix");
    return cheb_lambda(A, normal_eq != 0, lambda_max);
}

int drops_b200_matrix_bbt_diag(drops_b200_matrix_t A, double* diag)
{
    DB_CHECK(require_ctx());
    DB_ARG(A != nullptr && diag != nullptr, "null argument");
    const double* d = nullptr;
    DB_CHECK(bbt_diag(A, &d));
    Bridge br;
    return br.down(A->csr.rows, d, diag);
}

int drops_b200_precond_apply_dev(drops_b200_matrix_t A, int pc, double omega, int normal_eq, drops_b200_vector_t b,
                                 drops_b200_vector_t x)
{
    DB_CHECK(require_ctx());
    DB_CHECK(check_pc_system(A, pc, normal_eq != 0));
    DB_ARG(normal_eq || A->square, "preconditioner needs a square matrix");
    DB_ARG(b && x && b->n == A->csr.rows && x->n == A->csr.rows, "Apply: incompatible dimensions");
    DB_ARG(b->d != x->d || pc != DROPS_B200_PC_CHEBYSHEV, "ChebyshevPcCL::Apply: x must not alias b");
    return precond_apply(A, pc, omega, normal_eq != 0, b->d, x->d);
}

int drops_b200_pcg_general_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc, double omega,
                               int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_CHECK(check_pc_system(A, pc, false));
    DB_ARG(A->square, "PCG needs a square matrix");
    DB_ARG(x && b && x->n == A->csr.rows && b->n == A->csr.rows, "PCG: incompatible dimensions");
    return pcg_general_device(A, false, x->d, b->d, pc, omega, max_iter, tol, relative, hist, hist_cap);
}

int drops_b200_pcgne_dev(drops_b200_matrix_t A, drops_b200_vector_t u, drops_b200_vector_t b, int pc, double omega,
                         int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_CHECK(check_pc_system(A, pc, true));
    DB_ARG(u && b && u->n == A->csr.rows && b->n == A->csr.rows, "PCGNE: incompatible dimensions");
    return pcg_general_device(A, true, u->d, b->d, pc, omega, max_iter, tol, relative, hist, hist_cap);
}

int drops_b200_pcgne(drops_b200_matrix_t A, double* u, const double* b, int pc, double omega, int* max_iter,
                     double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_CHECK(check_pc_system(A, pc, true));
    DB_ARG(u && b, "null argument");
    Bridge br;
    const size_t n = A->csr.rows;
    DB_CHECK(br.up(n, u, n, b));
    const int r = pcg_general_device(A, true, br.dx, br.dy, pc, omega, max_iter, tol, relative, hist, hist_cap);
    if (r < 0) return r;
    DB_CHECK(br.down(n, br.dx, u));
    return r;
}

int drops_b200_idrs_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc, double omega,
                        int s, double omega_bound, const double* shadow, int* max_iter, double* tol, int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "IDRS: incompatible dimensions");
    return idrs_device(A, x->d, b->d, pc, omega, s, omega_bound, shadow, max_iter, tol, relative);
}

int drops_b200_idrs(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega, int s, double omega_bound,
                    const double* shadow, int* max_iter, double* tol, int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    Bridge br;
    const size_t n = A->csr.rows;
    DB_CHECK(br.up(n, x, n, b));
    const int r = idrs_device(A, br.dx, br.dy, pc, omega, s, omega_bound, shadow, max_iter, tol, relative);
    if (r < 0) return r;
    DB_CHECK(br.down(n, br.dx, x));
    return r;
}

int drops_b200_block_dims(const drops_b200_matrix_t blocks[4], const int ops[4], size_t* rows, size_t* cols)
{
    DB_ARG(blocks && ops, "null argument");
    size_t r[2], c[2];
    DB_CHECK(block_dims(blocks, ops, r, c));
    if (rows) *rows = r[0] + r[1];
    if (cols) *cols = c[0] + c[1];
    return 0;
}

int drops_b200_block_spmv_dev(const drops_b200_matrix_t blocks[4], const int ops[4], drops_b200_vector_t x,
                              drops_b200_vector_t y)
{
    DB_CHECK(require_ctx());
    DB_ARG(blocks && ops && x && y, "null argument");
    size_t r[2], c[2];
    DB_CHECK(block_dims(blocks, ops, r, c));
    DB_ARG(x->n == c[0] + c[1] && y->n == r[0] + r[1], "BlockMatrixBaseCL * Vec: incompatible dimensions");
    DB_ARG(x->d != y->d, "BlockMatrixBaseCL * Vec: x must not alias y");
    return block_apply(blocks, ops, x->d, y->d);
}

int drops_b200_block_spmv(const drops_b200_matrix_t blocks[4], const int ops[4], const double* x, double* y)
{
    DB_CHECK(require_ctx());
    DB_ARG(blocks && ops && x && y, "null argument");
    size_t r[2], c[2];
    DB_CHECK(block_dims(blocks, ops, r, c));
    Bridge br;
    DB_CHECK(br.up(c[0] + c[1], x, r[0] + r[1], nullptr));
    DB_CHECK(block_apply(blocks, ops, br.dx, br.dy));
    return br.down(r[0] + r[1], br.dy, y);
}

int drops_b200_composite_spmv_dev(drops_b200_matrix_t A0, int op0, drops_b200_matrix_t A1, int op1,
                                  drops_b200_vector_t x, drops_b200_vector_t y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A0 && A1 && x && y, "null argument");
    const size_t mid = op0 == 0 ? A0->csr.rows : A0->csr.cols;
    DB_ARG(x->n == (op0 == 0 ? A0->csr.cols : A0->csr.rows) && mid == (op1 == 0 ? A1->csr.cols : A1->csr.rows) &&
           y->n == (op1 == 0 ? A1->csr.rows : A1->csr.cols), "CompositeMatrixBaseCL * Vec: incompatible dimensions");
    double* tmp = work_vec(A0, 15);
    if (!tmp) return DROPS_B200_E_NOMEM;
    SpmvArgs a; a.x = x->d; a.y = tmp;
    DB_CHECK(apply_op(A0, op0, a));
    SpmvArgs b; b.x = tmp; b.y = y->d;
    return apply_op(A1, op1, b);
}

int drops_b200_composite_spmv(drops_b200_matrix_t A0, int op0, drops_b200_matrix_t A1, int op1, const double* x,
                              double* y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A0 && A1 && x && y, "null argument");
    const size_t nx = op0 == 0 ? A0->csr.cols : A0->csr.rows;
    const size_t mid = op0 == 0 ? A0->csr.rows : A0->csr.cols;
    const size_t ny = op1 == 0 ? A1->csr.rows : A1->csr.cols;
    DB_ARG(mid == (op1 == 0 ? A1->csr.cols : A1->csr.rows), "CompositeMatrixBaseCL: incompatible dimensions");
    Bridge br;
    DB_CHECK(br.up(nx, x, ny, nullptr));
    double* tmp = work_vec(A0, 15);
    if (!tmp) return DROPS_B200_E_NOMEM;
    SpmvArgs a; a.x = br.dx; a.y = tmp;
    DB_CHECK(apply_op(A0, op0, a));
    SpmvArgs b; b.x = tmp; b.y = br.dy;
    DB_CHECK(apply_op(A1, op1, b));
    return br.down(ny, br.dy, y);
}

}  // extern "C"

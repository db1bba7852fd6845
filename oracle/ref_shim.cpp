// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" veneer over the UNMODIFIED DROPS reference headers, compiled where
// they lie under /root/reference/src (see oracle/Makefile). The resulting
// oracle/_ref/libdrops_ref.so is the checker for the CUDA path and the CPU baseline
// that bench.py times ("cpu_baseline.kind": "reference").  Nothing below restates
// reference arithmetic: every numeric call goes straight into the reference's own
// templates:
//   A*x            -> DROPS::operator*            src/num/spmat.h:1712-1724  (y_Ax :1682-1709)
//   transp_mul     -> DROPS::transp_mul           src/num/spmat.h:1748-1760  (y_ATx :1729-1746)
//   dot/norm/axpy  -> src/num/spmat.h:164-237, KahanInnerProd :317-333
//   JACPcCL        -> src/num/precond.h:874 (PreGSCL<P_JAC0>, SolveGSstep<PB_JAC0> :153-169)
//   CG / PCG       -> src/num/krylovsolver.h:49-103 / :107-179, drivers :1782-1832
//   GMRES          -> src/num/krylovsolver.h:320-461, driver :1924-1973
//   MINRES/PMINRES -> src/num/krylovsolver.h:926-1161, drivers :1868-1921
//   DummyExchangeCL-> src/parallel/exchange.h:45-85
// IDRS (krylovsolver.h:1610-1613) draws its shadow space with srand(time(NULL)) / rand().  To make a
// reference run repeatable the two calls are redirected -- by macro, after every standard header that
// could mention them has been included, so the reference sources stay untouched -- to a queue the test
// fills (ref_idrs); with an empty queue rand() is the C library's.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <valarray>
#include <vector>
namespace shimrand {
static std::vector<int> queue;
static size_t pos = 0;
inline int next() { return pos < queue.size() ? queue[pos++] : std::rand(); }
inline void seed(unsigned) {}
}
#define rand() shimrand::next()
#define srand(x) shimrand::seed(x)

#include "num/krylovsolver.h"
#include "num/precond.h"
#include "num/spblockmat.h"
#include "parallel/exchange.h"

#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct RefMatrix {
    DROPS::MatrixCL A;
};

// The reference solvers chat on std::cout (GMRES/PMINRES); keep test logs clean and
// let callers fetch the text if they want to look at it.
struct CoutCapture {
    std::ostringstream sink;
    std::streambuf* old;
    CoutCapture() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutCapture() { std::cout.rdbuf(old); }
};

std::string g_last_log;

// Parse "PCG: <i> resid <r>" / "CG: <i> resid <r>" lines written by the reference
// into hist[i] (hist[0] is left to the caller).
int parse_history(const std::string& s, const char* tag, double* hist, int cap)
{
    std::istringstream in(s);
    std::string line;
    int last = 0;
    const size_t tl = std::strlen(tag);
    while (std::getline(in, line)) {
        if (line.compare(0, tl, tag) != 0) continue;
        std::istringstream ls(line.substr(tl));
        int it; std::string word; double r;
        if (ls >> it >> word >> r) {
            if (hist && it >= 0 && it < cap) hist[it] = r;
            last = it;
        }
    }
    return last;
}

DROPS::VectorCL to_vec(const double* p, size_t n) { return n ? DROPS::VectorCL(p, n) : DROPS::VectorCL(); }
void from_vec(const DROPS::VectorCL& v, double* p) { if (v.size()) std::memcpy(p, &v[0], v.size()*sizeof(double)); }

} // namespace

template <typename PC>
static int run_bicgstab(const DROPS::MatrixCL& A, PC& pc, double* x, const double* b, int* max_iter, double* tol,
                        int relative)
{
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    bool ok;
    {
        CoutCapture cap;
        ok = DROPS::BICGSTAB(A, xv, bv, DROPS::DummyExchangeCL(), pc, *max_iter, *tol, relative != 0);
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    return ok ? 1 : 0;
}

template <typename PC>
static int run_gcr(const DROPS::MatrixCL& A, PC& pc, double* x, const double* b, int m, int* max_iter, double* tol,
                   int relative)
{
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    bool ok;
    {
        CoutCapture cap;
        ok = DROPS::GCR(A, xv, bv, DROPS::DummyExchangeCL(), pc, m, *max_iter, *tol, relative != 0);
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    return ok ? 1 : 0;
}

extern "C" {

int ref_num_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_set_num_threads(int t)
{
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

const char* ref_last_log() { return g_last_log.c_str(); }

// Builds a DROPS::MatrixCL through SparseMatBaseCL(rows, cols, nnz) + raw_* (spmat.h:717,789-794).
void* ref_matrix_create(size_t rows, size_t cols, size_t nnz,
                        const size_t* rowbeg, const size_t* colind, const double* val)
{
    RefMatrix* m = new RefMatrix{DROPS::MatrixCL(rows, cols, nnz)};
    std::memcpy(m->A.raw_row(), rowbeg, (rows + 1)*sizeof(size_t));
    if (nnz) {
        std::memcpy(m->A.raw_col(), colind, nnz*sizeof(size_t));
        std::memcpy(m->A.raw_val(), val, nnz*sizeof(double));
    }
    return m;
}

// Builds through the hash-map SparseMatBuilderCL the reference tests use (spmat.h:420-676):
// entries given as (i, j, v) triplets, duplicates accumulate.
void* ref_matrix_build_triplets(size_t rows, size_t cols, size_t ntrip,
                                const size_t* ti, const size_t* tj, const double* tv)
{
    RefMatrix* m = new RefMatrix;
    DROPS::MatrixBuilderCL B(&m->A, rows, cols);
    for (size_t k = 0; k < ntrip; ++k) B(ti[k], tj[k]) += tv[k];
    B.Build();
    return m;
}

void ref_matrix_destroy(void* h) { delete static_cast<RefMatrix*>(h); }
size_t ref_matrix_rows(void* h) { return static_cast<RefMatrix*>(h)->A.num_rows(); }
size_t ref_matrix_cols(void* h) { return static_cast<RefMatrix*>(h)->A.num_cols(); }
size_t ref_matrix_nnz (void* h) { return static_cast<RefMatrix*>(h)->A.num_nonzeros(); }
size_t ref_matrix_version(void* h) { return static_cast<RefMatrix*>(h)->A.Version(); }
void ref_matrix_get_csr(void* h, size_t* rowbeg, size_t* colind, double* val)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    std::memcpy(rowbeg, A.raw_row(), (A.num_rows() + 1)*sizeof(size_t));
    if (A.num_nonzeros()) {
        std::memcpy(colind, A.raw_col(), A.num_nonzeros()*sizeof(size_t));
        std::memcpy(val, A.raw_val(), A.num_nonzeros()*sizeof(double));
    }
}

void ref_spmv(void* h, const double* x, double* y)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    DROPS::VectorCL xv(to_vec(x, A.num_cols()));
    DROPS::VectorCL yv(A*xv);
    from_vec(yv, y);
}

// In-place y_Ax on caller memory: what bench.py times (no valarray temporaries).
void ref_y_Ax_raw(void* h, const double* x, double* y)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    DROPS::y_Ax(y, A.num_rows(), A.raw_val(), A.raw_row(), A.raw_col(), x);
}

void ref_spmv_transpose(void* h, const double* x, double* y)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    DROPS::VectorCL xv(to_vec(x, A.num_rows()));
    DROPS::VectorCL yv(DROPS::transp_mul(A, xv));
    from_vec(yv, y);
}

void ref_get_diag(void* h, double* d)
{
    from_vec(static_cast<RefMatrix*>(h)->A.GetDiag(), d);
}

double ref_dot(size_t n, const double* x, const double* y)
{
    return DROPS::dot(to_vec(x, n), to_vec(y, n));
}
double ref_norm_sq(size_t n, const double* x) { return DROPS::norm_sq(to_vec(x, n)); }
double ref_norm(size_t n, const double* x) { return DROPS::norm(to_vec(x, n)); }
double ref_supnorm(size_t n, const double* x) { return DROPS::supnorm(to_vec(x, n)); }
double ref_kahan_dot(size_t n, const double* x, const double* y)
{
    return DROPS::DummyExchangeCL().ParDot(to_vec(x, n), false, to_vec(y, n), false);
}
double ref_kahan_norm(size_t n, const double* x)
{
    return DROPS::DummyExchangeCL().Norm(to_vec(x, n), false);
}
void ref_axpy(size_t n, double a, const double* x, double* y)
{
    DROPS::VectorCL yv(to_vec(y, n));
    DROPS::axpy(a, to_vec(x, n), yv);
    from_vec(yv, y);
}
void ref_z_xpay(size_t n, double* z, const double* x, double a, const double* y)
{
    DROPS::VectorCL zv(n);
    DROPS::z_xpay(zv, to_vec(x, n), a, to_vec(y, n));
    from_vec(zv, z);
}
void ref_z_xpaypby2(size_t n, double* z, const double* x, double a, const double* y, double b, const double* y2)
{
    DROPS::VectorCL zv(n);
    DROPS::z_xpaypby2(zv, to_vec(x, n), a, to_vec(y, n), b, to_vec(y2, n));
    from_vec(zv, z);
}

void ref_jacobi_apply(void* h, double omega, const double* b, double* x)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    DROPS::JACPcCL pc(omega);
    DROPS::VectorCL xv(A.num_rows()), bv(to_vec(b, A.num_rows()));
    pc.Apply(A, xv, bv, DROPS::DummyExchangeCL());
    from_vec(xv, x);
}

// PCGSolverCL<JACPcCL>::Solve  (krylovsolver.h:1805-1832).  hist (may be null) receives the
// per-iteration residuals the reference prints through SolverBaseCL::output_.
int ref_pcg_jacobi(void* h, double* x, const double* b, double omega,
                   int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::JACPcCL pc(omega);
    std::ostringstream os; os.precision(17);
    DROPS::PCGSolverCL<DROPS::JACPcCL> solver(pc, *max_iter, *tol, relative != 0, hist ? &os : 0);
    solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
    from_vec(xv, x);
    if (hist) parse_history(os.str(), "PCG:", hist, hist_cap);
    const bool conv = solver.GetResid() <= *tol;
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return conv ? 1 : 0;
}

// PCG with DummyPcCL (identity), same driver.
int ref_pcg_dummy(void* h, double* x, const double* b,
                  int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::DummyPcCL pc;
    std::ostringstream os; os.precision(17);
    DROPS::PCGSolverCL<DROPS::DummyPcCL> solver(pc, *max_iter, *tol, relative != 0, hist ? &os : 0);
    solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
    from_vec(xv, x);
    if (hist) parse_history(os.str(), "PCG:", hist, hist_cap);
    const bool conv = solver.GetResid() <= *tol;
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return conv ? 1 : 0;
}

// Free function PCG (krylovsolver.h:107-179) timed without any output stream: bench baseline.
int ref_pcg_jacobi_timed(void* h, double* x, const double* b, double omega,
                         int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::JACPcCL pc(omega);
    const bool conv = DROPS::PCG(A, xv, bv, DROPS::DummyExchangeCL(), pc, *max_iter, *tol, relative != 0);
    from_vec(xv, x);
    return conv ? 1 : 0;
}

// PCGSolverCL<DiagPcCL>::Solve with DiagPcCL(D) (precond.h:618-642): x = D o b.
int ref_pcg_diagpc(void* h, double* x, const double* b, const double* D, int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n)), Dv(to_vec(D, n));
    DROPS::DiagPcCL pc(Dv);
    DROPS::PCGSolverCL<DROPS::DiagPcCL> solver(pc, *max_iter, *tol, relative != 0);
    solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
    from_vec(xv, x);
    const bool conv = solver.GetResid() <= *tol;
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return conv ? 1 : 0;
}

// DiagPcCL::Apply (precond.h:626-631).
void ref_diagpc_apply(size_t n, const double* D, const double* b, double* x)
{
    DROPS::VectorCL Dv(to_vec(D, n)), bv(to_vec(b, n)), xv(n);
    DROPS::DiagPcCL pc(Dv);
    DROPS::MatrixCL dummy;
    pc.Apply(dummy, xv, bv, DROPS::DummyExchangeCL());
    from_vec(xv, x);
}

// CGSolverCL::Solve (krylovsolver.h:1782-1802).
int ref_cg(void* h, double* x, const double* b, int* max_iter, double* tol, int relative,
           double* hist, int hist_cap)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    std::ostringstream os; os.precision(17);
    DROPS::CGSolverCL solver(*max_iter, *tol, relative != 0, hist ? &os : 0);
    solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
    from_vec(xv, x);
    if (hist) parse_history(os.str(), "CG:", hist, hist_cap);
    const bool conv = solver.GetResid() <= *tol;
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return conv ? 1 : 0;
}

// GMResSolverCL<JACPcCL>::Solve (krylovsolver.h:1924-1973 -> serial GMRES :320-461).
// left_precond: 1 = LeftPreconditioning (the driver's default), 0 = RightPreconditioning.
int ref_gmres_jacobi(void* h, double* x, const double* b, double omega, int restart,
                     int* max_iter, double* tol, int relative, int left_precond)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::JACPcCL pc(omega);
    DROPS::GMResSolverCL<DROPS::JACPcCL> solver(pc, restart, *max_iter, *tol, relative != 0, false,
        left_precond ? DROPS::LeftPreconditioning : DROPS::RightPreconditioning);
    const double tol_in = *tol;
    {
        CoutCapture cap;
        solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return solver.GetResid() <= tol_in ? 1 : 0;
}

// BiCGStabSolverCL<PC>::Solve (krylovsolver.h:1978-2013 -> serial BICGSTAB :1189-1258) and
// GCRSolverCL<PC>::Solve (:2017-2053 -> GCR :1410-1483).  pc_kind 0 = DummyPcCL, 1 = JACPcCL.
// Return: 1 converged, 0 not (a breakdown also returns 0; the message is in ref_last_log).

int ref_bicgstab(void* h, double* x, const double* b, int pc_kind, double omega, int* max_iter, double* tol,
                 int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    if (pc_kind == 1) { DROPS::JACPcCL pc(omega); return run_bicgstab(A, pc, x, b, max_iter, tol, relative); }
    DROPS::DummyPcCL pc;
    return run_bicgstab(A, pc, x, b, max_iter, tol, relative);
}


int ref_gcr(void* h, double* x, const double* b, int pc_kind, double omega, int m, int* max_iter, double* tol,
            int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    if (pc_kind == 1) { DROPS::JACPcCL pc(omega); return run_gcr(A, pc, x, b, m, max_iter, tol, relative); }
    DROPS::DummyPcCL pc;
    return run_gcr(A, pc, x, b, m, max_iter, tol, relative);
}

int ref_gmres_dummy(void* h, double* x, const double* b, int restart,
                    int* max_iter, double* tol, int relative, int left_precond)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::DummyPcCL pc;
    DROPS::GMResSolverCL<DROPS::DummyPcCL> solver(pc, restart, *max_iter, *tol, relative != 0, false,
        left_precond ? DROPS::LeftPreconditioning : DROPS::RightPreconditioning);
    const double tol_in = *tol;
    {
        CoutCapture cap;
        solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return solver.GetResid() <= tol_in ? 1 : 0;
}

// MResSolverCL::Solve (krylovsolver.h:1868-1889 -> MINRES :1155-1161 -> PMINRES :1069-1152).
int ref_minres(void* h, double* x, const double* b, int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::MResSolverCL solver(*max_iter, *tol, relative != 0);
    bool conv;
    {
        CoutCapture cap;
        int it = *max_iter; double res = *tol;
        conv = DROPS::MINRES(A, xv, bv, DROPS::DummyExchangeCL(), it, res, relative != 0);
        *max_iter = it; *tol = res;
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    return conv ? 1 : 0;
}

// PMResSolverCL<PLanczosONBCL<VectorCL, JACPcCL>> (krylovsolver.h:989-1052, :1892-1921).
int ref_pminres_jacobi(void* h, double* x, const double* b, double omega,
                       int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::JACPcCL pc(omega);
    typedef DROPS::PLanczosONBCL<DROPS::VectorCL, DROPS::JACPcCL> LanczosT;
    LanczosT q(pc);
    DROPS::PMResSolverCL<LanczosT> solver(q, *max_iter, *tol, relative != 0);
    const double tol_in = *tol;
    {
        CoutCapture cap;
        solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return solver.GetResid() <= tol_in ? 1 : 0;
}

// PMResSolverCL<PLanczosONBCL<VectorCL, DummyPcCL>>: TEST5 of src/tests/minres.cpp.
int ref_pminres_dummy(void* h, double* x, const double* b, int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::DummyPcCL pc;
    typedef DROPS::PLanczosONBCL<DROPS::VectorCL, DROPS::DummyPcCL> LanczosT;
    LanczosT q(pc);
    DROPS::PMResSolverCL<LanczosT> solver(q, *max_iter, *tol, relative != 0);
    const double tol_in = *tol;
    {
        CoutCapture cap;
        solver.Solve(A, xv, bv, DROPS::DummyExchangeCL());
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    *max_iter = solver.GetIter();
    *tol = solver.GetResid();
    return solver.GetResid() <= tol_in ? 1 : 0;
}

// Lanczos basis of src/tests/minres.cpp TEST1: returns the number of basis vectors generated
// before the lucky breakdown; Q receives them row-major (max_vecs x n), a/b the coefficients.
int ref_lanczos_basis(void* h, const double* r0, int max_vecs, double* Q, double* a, double* bcoef)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::LanczosONBCL<DROPS::VectorCL> onb;
    onb.new_basis(A, to_vec(r0, n), DROPS::DummyExchangeCL());
    int k = 0;
    do {
        if (k < max_vecs) { from_vec(onb.q[0], Q + size_t(k)*n); a[k] = onb.a0; bcoef[k] = onb.b[0]; }
        ++k;
    } while (k < max_vecs && onb.next(A, DROPS::DummyExchangeCL()));
    if (k < max_vecs) { from_vec(onb.q[0], Q + size_t(k)*n); a[k] = onb.a0; bcoef[k] = onb.b[0]; ++k; }
    return k;
}

// BlockMatrixCL / CompositeMatrixCL products of src/tests/blockmat.cpp (spblockmat.h):
//   S = [A  B^T; B  0],  returns S*b and transp_mul(S, b).
void ref_blockmat_saddle(void* hA, void* hB, const double* b, double* Sb, double* STb)
{
    DROPS::MatrixCL& A = static_cast<RefMatrix*>(hA)->A;
    DROPS::MatrixCL& B = static_cast<RefMatrix*>(hB)->A;
    DROPS::BlockMatrixCL S(&A, DROPS::MUL, &B, DROPS::TRANSP_MUL, &B, DROPS::MUL, 0, DROPS::MUL);
    DROPS::VectorCL bv(to_vec(b, S.num_cols()));
    from_vec(DROPS::VectorCL(S*bv), Sb);
    from_vec(DROPS::VectorCL(transp_mul(S, to_vec(b, S.num_rows()))), STb);
}

// Matrix text I/O of the reference itself (src/num/spmat.h:1213-1255): operator<< (MatrixMarket),
// out() / in() (CSR dump).  Used to check drops_b200/matrix_io.py and the drop-in headers' I/O.
int ref_matrix_write(void* h, const char* path, int format, int precision)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    std::ofstream f(path);
    if (!f) return -1;
    if (format == 0) f << A;
    else { f.precision(precision); DROPS::out(f, A); }
    return f ? 0 : -1;
}

void* ref_matrix_read_text(const char* path)
{
    std::ifstream f(path);
    if (!f) return 0;
    RefMatrix* M = new RefMatrix;
    DROPS::in(f, M->A);
    return M;
}

// SparseMatBaseCL::LinComb (src/num/spmat.h:1457-1551): returns a new matrix handle a*A + b*B.
void* ref_matrix_lincomb(double a, void* hA, double b, void* hB)
{
    RefMatrix* M = new RefMatrix;
    M->A.LinComb(a, static_cast<RefMatrix*>(hA)->A, b, static_cast<RefMatrix*>(hB)->A);
    return M;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// PCGNE, Chebyshev, IDR(s), block / composite products.
// ---------------------------------------------------------------------------------------------
namespace {
// PCGNE needs M.SetDiag(A, ex, ex) and M.Apply(A, x, b, ex, ex) (krylovsolver.h:204, :218).  Of the
// serial preconditioners only NEGSPcCL and ChebyshevBBTPcCL have both; these two adapters give the
// reference PCGNE an identity and a Jacobi (diag(A A^T) = BBTDiag, the _PAR ParJacNEG0CL of
// precond.h:1190-1229, applied with the reference's own Jacobi0 :1293-1311).
struct ShimIdentityNEPc {
    bool RetAcc() const { return false; }
    bool NeedDiag() const { return false; }
    template <typename Mat, typename ExT> void SetDiag(const Mat&, const ExT&, const ExT&) {}
    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat&, Vec& x, const Vec& b, const ExT&, const ExT&) const { x = b; }
};
struct ShimJacobiNEPc {
    DROPS::VectorCL diag_;
    double omega_;
    explicit ShimJacobiNEPc(double omega) : omega_(omega) {}
    bool RetAcc() const { return false; }
    bool NeedDiag() const { return true; }
    template <typename Mat, typename ExT> void SetDiag(const Mat& A, const ExT& rowex, const ExT& colex)
    {
        diag_.resize(A.num_rows());
        diag_ = DROPS::BBTDiag(A, rowex, colex);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat& A, Vec& x, const Vec& b, const ExT&, const ExT&) const
    {
        DROPS::Jacobi0(A, diag_, x, b, omega_, std::fabs(omega_ - 1.) > DROPS::DoubleEpsC);
    }
};

template <typename PC>
int run_pcgne(const DROPS::MatrixCL& A, PC& pc, double* u, const double* b, int* max_iter, double* tol, int relative)
{
    const size_t n = A.num_rows();
    DROPS::VectorCL uv(to_vec(u, n)), bv(to_vec(b, n));
    DROPS::DummyExchangeCL ex;
    bool ok;
    {
        CoutCapture cap;
        ok = DROPS::PCGNE(A, uv, bv, ex, ex, pc, *max_iter, *tol, relative != 0);
        g_last_log = cap.sink.str();
    }
    from_vec(uv, u);
    return ok ? 1 : 0;
}
} // namespace

extern "C" {

void ref_bbt_diag(void* h, double* d)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    from_vec(DROPS::BBTDiag(A, DROPS::DummyExchangeCL(), DROPS::DummyExchangeCL()), d);
}

// MaxEigenvalueScaledMatrix (precond.h:1080-1096) on A (ne = 0) or on the composite A A^T (ne = 1)
double ref_max_eig_scaled(void* h, int ne, const double* D, int iter)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    DROPS::DummyExchangeCL ex;
    DROPS::VectorCL d(to_vec(D, A.num_rows()));
    if (!ne) return DROPS::MaxEigenvalueScaledMatrix(A, d, ex, iter);
    DROPS::CompositeMatrixBaseCL<DROPS::MatrixCL, DROPS::MatrixCL, DROPS::DummyExchangeCL>
        mat(&A, DROPS::TRANSP_MUL, ex, &A, DROPS::MUL, ex);
    return DROPS::MaxEigenvalueScaledMatrix(mat, d, ex, iter);
}

// ChebyshevPcCL (ne = 0, precond.h:1463-1529) / ChebyshevBBTPcCL (ne = 1, :1531-1567): SetDiag + Apply.
// Returns lambda_max as printed by SetDiag.
double ref_chebyshev_apply(void* h, int ne, double scale, int degree, double ratio, const double* b, double* x)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::DummyExchangeCL ex;
    DROPS::VectorCL bv(to_vec(b, n)), xv(n);
    CoutCapture cap;
    if (!ne) {
        DROPS::ChebyshevPcCL pc(scale, degree, ratio);
        pc.SetDiag(A, ex);
        pc.Apply(A, xv, bv, ex);
    } else {
        DROPS::ChebyshevBBTPcCL pc(scale, degree, ratio);
        pc.SetDiag(A, ex, ex);
        pc.Apply(A, xv, bv, ex, ex);
    }
    g_last_log = cap.sink.str();
    from_vec(xv, x);
    const size_t p = g_last_log.rfind("lambda_max = ");
    return p == std::string::npos ? 0.0 : std::atof(g_last_log.c_str() + p + 13);
}

// PCGSolverCL<ChebyshevPcCL>
int ref_pcg_chebyshev(void* h, double* x, const double* b, double scale, int degree, double ratio,
                      int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    DROPS::ChebyshevPcCL pc(scale, degree, ratio);
    bool ok;
    {
        CoutCapture cap;
        ok = DROPS::PCG(A, xv, bv, DROPS::DummyExchangeCL(), pc, *max_iter, *tol, relative != 0);
        g_last_log = cap.sink.str();
    }
    from_vec(xv, x);
    return ok ? 1 : 0;
}

// PCGNE (krylovsolver.h:199-252).  pc_kind 0: identity adapter, 1: Jacobi-on-BBTDiag adapter,
// 2: the reference's ChebyshevBBTPcCL(scale, degree, ratio).
int ref_pcgne(void* h, double* u, const double* b, int pc_kind, double omega, double scale, int degree, double ratio,
              int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    if (pc_kind == 2) { DROPS::ChebyshevBBTPcCL pc(scale, degree, ratio); return run_pcgne(A, pc, u, b, max_iter, tol, relative); }
    if (pc_kind == 1) { ShimJacobiNEPc pc(omega); return run_pcgne(A, pc, u, b, max_iter, tol, relative); }
    ShimIdentityNEPc pc;
    return run_pcgne(A, pc, u, b, max_iter, tol, relative);
}

// IDRS (krylovsolver.h:1579-1698).  draws: the s*n values rand() returns at :1613, in call order.
// Returns the reference's bool, -1 when it throws (M(k,k) == 0 / omega == 0).
int ref_idrs(void* h, double* x, const double* b, int pc_kind, double omega_pc, int s, double omega_bound,
             const int* draws, size_t ndraws, int* max_iter, double* tol, int relative)
{
    const DROPS::MatrixCL& A = static_cast<RefMatrix*>(h)->A;
    const size_t n = A.num_rows();
    DROPS::VectorCL xv(to_vec(x, n)), bv(to_vec(b, n));
    shimrand::queue.assign(draws, draws + ndraws);
    shimrand::pos = 0;
    // The reference reads uninitialised memory here: DMatrixCL M(s,s) is a bare new T[] (misc/container.h:51)
    // and the first cycle solves with the strictly lower triangle of M before any of it was written
    // (krylovsolver.h:1620-1636).  The values do not matter as long as they are finite (G and U are still
    // zero), but recycled heap chunks can hold NaN bit patterns.  Park a few zeroed chunks of exactly that
    // size in the allocator's free list so that M starts from finite memory and the run is repeatable.
    {
        double* park[6];
        for (int i = 0; i < 6; ++i) { park[i] = new double[(size_t)s * s]; std::memset(park[i], 0, sizeof(double) * s * s); }
        for (int i = 0; i < 6; ++i) delete[] park[i];
    }
    int ret;
    try {
        CoutCapture cap;
        bool ok;
        if (pc_kind == 1) { DROPS::JACPcCL pc(omega_pc); ok = DROPS::IDRS(A, xv, bv, DROPS::DummyExchangeCL(), pc, *max_iter, *tol, relative != 0, s, omega_bound); }
        else { DROPS::DummyPcCL pc; ok = DROPS::IDRS(A, xv, bv, DROPS::DummyExchangeCL(), pc, *max_iter, *tol, relative != 0, s, omega_bound); }
        ret = ok ? 1 : 0;
    } catch (DROPS::DROPSErrCL&) {
        ret = -1;
    }
    shimrand::queue.clear();
    shimrand::pos = 0;
    from_vec(xv, x);
    return ret;
}

// BlockMatrixCL(A, opA, B, opB, C, opC, D, opD) * x and transp_mul (spblockmat.h:225-273); null handles
// are absent blocks.  rows/cols: num_rows()/num_cols() of the block matrix.
void ref_block_dims(void* const* hb, const int* ops, size_t* rows, size_t* cols)
{
    const DROPS::MatrixCL* m[4];
    for (int i = 0; i < 4; ++i) m[i] = hb[i] ? &static_cast<RefMatrix*>(hb[i])->A : 0;
    DROPS::BlockMatrixCL S(m[0], (DROPS::BlockMatrixOperationT)ops[0], m[1], (DROPS::BlockMatrixOperationT)ops[1],
                           m[2], (DROPS::BlockMatrixOperationT)ops[2], m[3], (DROPS::BlockMatrixOperationT)ops[3]);
    *rows = S.num_rows(); *cols = S.num_cols();
}

void ref_block_mul(void* const* hb, const int* ops, int transposed, const double* x, double* y)
{
    const DROPS::MatrixCL* m[4];
    for (int i = 0; i < 4; ++i) m[i] = hb[i] ? &static_cast<RefMatrix*>(hb[i])->A : 0;
    DROPS::BlockMatrixCL S(m[0], (DROPS::BlockMatrixOperationT)ops[0], m[1], (DROPS::BlockMatrixOperationT)ops[1],
                           m[2], (DROPS::BlockMatrixOperationT)ops[2], m[3], (DROPS::BlockMatrixOperationT)ops[3]);
    if (!transposed) from_vec(DROPS::VectorCL(S * to_vec(x, S.num_cols())), y);
    else from_vec(DROPS::VectorCL(transp_mul(S, to_vec(x, S.num_rows()))), y);
}

// CompositeMatrixCL(A0, op0, ex, A1, op1, ex) * x (spblockmat.h:376-397)
void ref_composite_mul(void* h0, int op0, void* h1, int op1, int transposed, const double* x, double* y)
{
    const DROPS::MatrixCL& A0 = static_cast<RefMatrix*>(h0)->A;
    const DROPS::MatrixCL& A1 = static_cast<RefMatrix*>(h1)->A;
    DROPS::DummyExchangeCL ex;
    DROPS::CompositeMatrixCL Cm(&A0, (DROPS::BlockMatrixOperationT)op0, ex, &A1, (DROPS::BlockMatrixOperationT)op1, ex);
    if (!transposed) from_vec(DROPS::VectorCL(Cm * to_vec(x, Cm.num_cols())), y);
    else from_vec(DROPS::VectorCL(transp_mul(Cm, to_vec(x, Cm.num_rows()))), y);
}

} // extern "C"

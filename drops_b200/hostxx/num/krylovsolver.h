// drops_b200/hostxx/num/krylovsolver.h
//
// B200 drop-in for the part of DROPS' src/num/krylovsolver.h named by the hot path: the free
// functions CG (:49-103), PCG (:107-179), GMRES (:320-461), MINRES (:1155-1161), PMINRES
// (:1069-1152), the orthonormal-basis classes LanczosONBCL (:945-979) / PLanczosONBCL
// (:1010-1052) and the drivers CGSolverCL (:1782-1802), PCGSolverCL (:1805-1832), MResSolverCL
// (:1868-1889), PMResSolverCL (:1892-1921), GMResSolverCL (:1924-1973); next to them PCGNE (:199-252,
// PCGNESolverCL :1836-1865), BICGSTAB, GCR and IDRS (:1579-1698, IDRsSolverCL :2095-2135).
//
// Signatures, in/out conventions and return values are the reference's: x is the start vector
// on entry and the solution on exit, max_iter IN limit / OUT iterations used, tol IN tolerance /
// OUT residual reached, the bool result says "converged"; the drivers record the outcome in
// SolverBaseCL::iter_/res_.  The bodies do not iterate on the host: they hand the system to the
// device solvers behind include/drops_b200.h (vectors are copied in and out once per Solve, the
// matrix copy is cached on SparseMatBaseCL::Version()).  Only DummyExchangeCL-style serial
// exchange objects are meaningful here; multi-GPU runs use the row-partitioned device solver.
#ifndef DROPS_SOLVER_H
#define DROPS_SOLVER_H

#include <iostream>
#include <vector>

#include "misc/container.h"
#include "num/precond.h"
#include "num/solverbase.h"
#include "num/spmat.h"

namespace DROPS
{

namespace b200
{
/// Prints the residual history the way the reference's loops do ("PCG: 3 resid 1.2e-05").
inline void print_history(std::ostream* out, const char* tag, const std::vector<double>& hist, int iters)
{
    if (!out) return;
    for (int i = 1; i <= iters && i < static_cast<int>(hist.size()); ++i)
        (*out) << tag << ": " << i << " resid " << hist[i] << std::endl;
}

template <typename Mat, typename Vec>
inline void check_dims(const Mat& A, const Vec& x, const Vec& b, const char* who)
{
    if (A.num_rows() != A.num_cols() || x.size() != A.num_rows() || b.size() != A.num_rows())
        throw DROPSErrCL(std::string(who) + ": incompatible dimensions");
}
}  // namespace b200

//=============================================================================
//  Free functions
//=============================================================================

/// Conjugate gradients without preconditioner.
template <typename Mat, typename Vec, typename ExCL>
bool CG(const Mat& A, Vec& x, const Vec& b, const ExCL&, int& max_iter, double& tol, bool measure_relative_tol,
        std::ostream* output = 0)
{
    const MatrixCL& M = b200::finest(A);
    b200::check_dims(M, x, b, "CG");
    std::vector<double> hist(output ? max_iter + 1 : 0);
    int it = max_iter;
    const int r = drops_b200_cg(M.device(), &x[0], Addr(b), &it, &tol, measure_relative_tol ? 1 : 0,
                                output ? hist.data() : 0, static_cast<int>(hist.size()));
    b200::check(r);
    if (r == 1) max_iter = it;
    b200::print_history(output, "CG", hist, r == 1 ? it : max_iter);
    return r == 1;
}

/// Preconditioned conjugate gradients; M is JACPcCL or DummyPcCL.
template <typename Mat, typename Vec, typename PreCon, typename ExCL>
bool PCG(const Mat& A, Vec& x, const Vec& b, const ExCL& ex, PreCon& M, int& max_iter, double& tol,
         bool measure_relative_tol = false, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "PCG");
    std::vector<double> hist(output ? max_iter + 1 : 0);
    int it = max_iter;
    if (M.NeedDiag()) M.SetDiag(A, ex);                 // krylovsolver.h:123-124 (Chebyshev: lambda_max, printed)
    b200::PcTraits<PreCon>::prepare(M, Mat0.device(), false);
    const int r = drops_b200_pcg(Mat0.device(), &x[0], Addr(b), b200::PcTraits<PreCon>::kind(),
                                 b200::PcTraits<PreCon>::omega(M), &it, &tol, measure_relative_tol ? 1 : 0,
                                 output ? hist.data() : 0, static_cast<int>(hist.size()));
    b200::check(r);
    if (r == 1) max_iter = it;
    b200::print_history(output, "PCG", hist, r == 1 ? it : max_iter);
    return r == 1;
}

/// PCG for the normal equations A A^T u = b of a rectangular A (src/num/krylovsolver.h:199-252); M is
/// ChebyshevBBTPcCL or DummyPcCL, or JACPcCL (then Jacobi on diag(A A^T), what the MPI build's
/// ParJacNEG0CL does).  The residual log follows the reference: iteration 0 and every tenth.
template <typename Mat, typename Vec, typename PreCon, typename ExACL, typename ExATranspCL>
bool PCGNE(const Mat& A, Vec& u, const Vec& b, const ExACL& ExAX, const ExATranspCL& ExATranspX, PreCon& M, int& max_iter, double& tol,
           bool measure_relative_tol = false, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    if (u.size() != Mat0.num_rows() || b.size() != Mat0.num_rows()) throw DROPSErrCL("PCGNE: incompatible dimensions");
    std::vector<double> hist(output ? max_iter + 1 : 0);
    int it = max_iter;
    M.SetDiag(A, ExATranspX, ExAX);                      // krylovsolver.h:204
    b200::PcTraits<PreCon>::prepare(M, Mat0.device(), true);
    const int r = drops_b200_pcgne(Mat0.device(), &u[0], Addr(b), b200::PcTraits<PreCon>::kind(),
                                   b200::PcTraits<PreCon>::omega(M), &it, &tol, measure_relative_tol ? 1 : 0,
                                   output ? hist.data() : 0, static_cast<int>(hist.size()));
    b200::check(r);
    if (r == 1) max_iter = it;
    if (output) {
        const int last = r == 1 ? it : max_iter;
        for (int i = 0; i <= last && i < static_cast<int>(hist.size()); ++i)
            if (i == 0 || i % 10 == 0) (*output) << "PCGNE: iter: " << i << " resid: " << hist[i] << '\n';
    }
    return r == 1;
}

enum PreMethGMRES { RightPreconditioning, LeftPreconditioning };

/// Restarted GMRES(m) with modified Gram-Schmidt; the residual is the preconditioned one for
/// LeftPreconditioning.  calculate2norm (a debugging aid of the reference) is accepted and ignored.
template <typename Mat, typename Vec, typename PreCon, typename ExT>
bool GMRES(const Mat& A, Vec& x, const Vec& b, const ExT&, PreCon& M, int m, int& max_iter, double& tol,
           bool measure_relative_tol = true, bool /*calculate2norm*/ = false,
           PreMethGMRES method = RightPreconditioning, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "GMRES");
    int it = max_iter;
    const int r = drops_b200_gmres(Mat0.device(), &x[0], Addr(b), b200::PcTraits<PreCon>::kind(),
                                   b200::PcTraits<PreCon>::omega(M), m, &it, &tol, measure_relative_tol ? 1 : 0,
                                   method == LeftPreconditioning ? 1 : 0);
    b200::check(r);
    if (r == 1) max_iter = it;
    if (output)
        (*output) << (r == 1 ? "GMRES converged in " : "GMRES DIDN'T converge after ") << it
                  << " iterations with residual: " << tol << std::endl;
    return r == 1;
}

/// BiCGStab (src/num/krylovsolver.h:1189-1258).  On a breakdown the reference prints a message,
/// sets max_iter to the iteration and returns false; so does this.
template <class Mat, class Vec, class Preconditioner, class ExT>
bool BICGSTAB(const Mat& A, Vec& x, const Vec& b, const ExT&, Preconditioner& M, int& max_iter, double& tol,
              bool measure_relative_tol = true)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "BICGSTAB");
    int it = max_iter, breakdown = 0;
    const int r = drops_b200_bicgstab(Mat0.device(), &x[0], Addr(b), b200::PcTraits<Preconditioner>::kind(),
                                      b200::PcTraits<Preconditioner>::omega(M), &it, &tol,
                                      measure_relative_tol ? 1 : 0, &breakdown);
    b200::check(r);
    if (breakdown == 1) std::cout << "BiCGSTAB: Breakdown with rho_1 = 0.\n";
    if (breakdown == 2) std::cout << "BiCGSTAB: Breakdown with omega_ = 0.\n";
    if (r == 1 || breakdown) max_iter = it;
    return r == 1;
}

/// IDR(s) (src/num/krylovsolver.h:1579-1698).  The shadow space is drawn like the reference draws it
/// (srand(time(0)), rand()); outputs follow the reference literally: tol returns the ABSOLUTE residual
/// norm, max_iter the iterations done, and the bool is false when the tolerance was beaten inside the
/// loop (:1689-1693) -- callers in DROPS ignore it and read GetResid().
template <class Mat, class Vec, class PC, class ExT>
bool IDRS(const Mat& A, Vec& x, const Vec& rhs, const ExT&, PC& pc, int& max_iter, double& tol,
          bool measure_relative_tol = false, const int s = 4, typename Vec::value_type omega_bound = 0.7)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, rhs, "IDRS");
    const int r = drops_b200_idrs(Mat0.device(), &x[0], Addr(rhs), b200::PcTraits<PC>::kind(), b200::PcTraits<PC>::omega(pc),
                                  s, omega_bound, 0, &max_iter, &tol, measure_relative_tol ? 1 : 0);
    if (r < 0) throw DROPSErrCL(std::string(drops_b200_last_error()));     // "IDR(s): M(k,k) ==0" / "omega ==0"
    return r == 1;
}

/// GCR(m) with min-alpha truncation (src/num/krylovsolver.h:1410-1483).
template <typename Mat, typename Vec, typename PreCon, typename ExCL>
bool GCR(const Mat& A, Vec& x, const Vec& b, const ExCL&, PreCon& M, int m, int& max_iter, double& tol,
         bool measure_relative_tol = true, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "GCR");
    int it = max_iter;
    const int r = drops_b200_gcr(Mat0.device(), &x[0], Addr(b), b200::PcTraits<PreCon>::kind(),
                                 b200::PcTraits<PreCon>::omega(M), m, &it, &tol, measure_relative_tol ? 1 : 0);
    b200::check(r);
    if (r == 1) max_iter = it;
    if (output) (*output) << "GCR: k: " << it << "\tresidual: " << tol << std::endl;
    return r == 1;
}

//-----------------------------------------------------------------------------
// Lanczos bases.  They are usable on their own (new_basis / next walk the three-term recurrence
// with the product on the GPU), and they select the MINRES flavour when handed to PMResSolverCL.
//-----------------------------------------------------------------------------
template <typename Mat, typename Vec>
bool LanczosStep(const Mat& A, const Vec& q0, const Vec& q1, Vec& q2, double& a1, const double b0, double& b1)
{
    q2 = A * q1 - b0 * q0;
    a1 = dot(q2, q1);
    q2 -= a1 * q1;
    b1 = norm(q2);
    if (b1 < 1e-15) return false;   // lucky breakdown: the Krylov space is A-invariant
    q2 /= b1;
    return true;
}

template <typename Vec>
class LanczosONBCL
{
  private:
    bool nobreakdown_;
    double norm_r0_;

  public:
    SBufferCL<Vec, 3> q;
    double a0;
    SBufferCL<double, 2> b;

    LanczosONBCL() : nobreakdown_(true), norm_r0_(0.), a0(0.) {}

    template <typename Mat, typename ExT>
    void new_basis(const Mat& A, const Vec& r0, const ExT&)
    {
        q[-1].resize(r0.size(), 0.);
        norm_r0_ = norm(r0);
        q[0].resize(r0.size(), 0.); q[0] = r0 / norm_r0_;
        q[1].resize(r0.size(), 0.);
        b[-1] = 0.;
        nobreakdown_ = LanczosStep(A, q[-1], q[0], q[1], a0, b[-1], b[0]);
    }
    double norm_r0() const { return norm_r0_; }
    bool breakdown() const { return !nobreakdown_; }

    template <typename Mat, typename ExT>
    bool next(const Mat& A, const ExT&)
    {
        q.rotate(); b.rotate();
        return (nobreakdown_ = LanczosStep(A, q[-1], q[0], q[1], a0, b[-1], b[0]));
    }
};

template <typename Vec, typename PreCon>
class PLanczosONBCL
{
  public:
    const PreCon& M;
    PLanczosONBCL(const PreCon& M_) : M(M_) {}
};

namespace b200
{
/// How a Lanczos type maps onto drops_b200_minres.
template <typename L>
struct LanczosTraits;
template <typename Vec>
struct LanczosTraits<LanczosONBCL<Vec> > {
    static int precond() { return 0; }
    static int kind(const LanczosONBCL<Vec>&) { return DROPS_B200_PC_DUMMY; }
    static double omega(const LanczosONBCL<Vec>&) { return 1.0; }
};
template <typename Vec, typename PC>
struct LanczosTraits<PLanczosONBCL<Vec, PC> > {
    static int precond() { return 1; }
    static int kind(const PLanczosONBCL<Vec, PC>&) { return PcTraits<PC>::kind(); }
    static double omega(const PLanczosONBCL<Vec, PC>& q) { return PcTraits<PC>::omega(q.M); }
};

template <typename Mat, typename Vec>
inline bool run_minres(const Mat& A, Vec& x, const Vec& rhs, int precond, int kind, double omega, int& max_iter,
                       double& tol, bool rel)
{
    const MatrixCL& Mat0 = finest(A);
    check_dims(Mat0, x, rhs, "MINRES");
    int it = max_iter;
    const int r = drops_b200_minres(Mat0.device(), &x[0], Addr(rhs), precond, kind, omega, &it, &tol, rel ? 1 : 0);
    check(r);
    if (r == 1) max_iter = it;
    return r == 1;
}
}  // namespace b200

/// (Preconditioned) MINRES driven by a Lanczos basis type; the basis object only selects the
/// flavour, the recurrence itself runs on the device starting from rhs - A*x.
template <typename Mat, typename Vec, typename Lanczos, typename ExT>
bool PMINRES(const Mat& A, Vec& x, const Vec& rhs, const ExT&, Lanczos& q, int& max_iter, double& tol,
             bool measure_relative_tol = false)
{
    typedef b200::LanczosTraits<Lanczos> LT;
    return b200::run_minres(A, x, rhs, LT::precond(), LT::kind(q), LT::omega(q), max_iter, tol, measure_relative_tol);
}

template <typename Mat, typename Vec, typename ExT>
bool MINRES(const Mat& A, Vec& x, const Vec& rhs, const ExT&, int& max_iter, double& tol,
            bool measure_relative_tol = false)
{
    return b200::run_minres(A, x, rhs, 0, DROPS_B200_PC_DUMMY, 1.0, max_iter, tol, measure_relative_tol);
}

//=============================================================================
//  Drivers
//=============================================================================

class CGSolverCL : public SolverBaseCL
{
  public:
    CGSolverCL(int maxiter, double tol, bool rel = false, std::ostream* output = 0)
        : SolverBaseCL(maxiter, tol, rel, output) {}

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        CG(A, x, b, ex, iter_, res_, rel_, output_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        CG(A, x, b, ex, numIter, resid, rel_, output_);
    }
};

template <typename PC>
class PCGSolverCL : public SolverBaseCL
{
  private:
    PC& _pc;

  public:
    PCGSolverCL(PC& pc, int maxiter, double tol, bool rel = false, std::ostream* output = 0)
        : SolverBaseCL(maxiter, tol, rel, output), _pc(pc) {}

    PC&       GetPc()       { return _pc; }
    const PC& GetPc() const { return _pc; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        PCG(A, x, b, ex, _pc, iter_, res_, rel_, output_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        PCG(A, x, b, ex, _pc, numIter, resid, rel_, output_);
    }
};

/// PCG for A A^T; note that A, not A A^T, is handed to Solve (src/num/krylovsolver.h:1836-1865).
template <typename PC>
class PCGNESolverCL : public SolverBaseCL
{
  private:
    PC& pc_;

  public:
    PCGNESolverCL(PC& pc, int maxiter, double tol, bool rel = false, std::ostream* output = 0)
        : SolverBaseCL(maxiter, tol, rel, output), pc_(pc) {}

    PC&       GetPc()       { return pc_; }
    const PC& GetPc() const { return pc_; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, const ExT& ex_transp)
    {
        res_ = tol_; iter_ = maxiter_;
        PCGNE(A, x, b, ex, ex_transp, pc_, iter_, res_, rel_, output_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, const ExT& ex_transp, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        PCGNE(A, x, b, ex, ex_transp, pc_, numIter, resid, rel_, output_);
    }
};

class MResSolverCL : public SolverBaseCL
{
  public:
    MResSolverCL(int maxiter, double tol, bool rel = false) : SolverBaseCL(maxiter, tol, rel) {}

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        MINRES(A, x, b, ex, iter_, res_, rel_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        MINRES(A, x, b, ex, numIter, resid, rel_);
    }
};

template <typename Lanczos>
class PMResSolverCL : public SolverBaseCL
{
  private:
    Lanczos& q_;

  public:
    PMResSolverCL(Lanczos& q, int maxiter, double tol, bool rel = false) : SolverBaseCL(maxiter, tol, rel), q_(q) {}

    Lanczos&       GetONB()       { return q_; }
    const Lanczos& GetONB() const { return q_; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        PMINRES(A, x, b, ex, q_, iter_, res_, rel_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        PMINRES(A, x, b, ex, q_, numIter, resid, rel_);
    }
};

template <typename PC>
class GMResSolverCL : public SolverBaseCL
{
  private:
    PC&          pc_;
    int          restart_;
    bool         calculate2norm_;
    PreMethGMRES method_;

  public:
    /// mod / useModGS select Gram-Schmidt variants of the MPI build of DROPS; the serial path
    /// (and this one) always runs modified Gram-Schmidt.
    GMResSolverCL(PC& pc, int restart, int maxiter, double tol, bool relative = true, bool calculate2norm = false,
                  PreMethGMRES method = LeftPreconditioning, bool /*mod*/ = true, bool /*useModGS*/ = false,
                  std::ostream* output = 0)
        : SolverBaseCL(maxiter, tol, relative, output), pc_(pc), restart_(restart),
          calculate2norm_(calculate2norm), method_(method) {}

    PC&       GetPc()       { return pc_; }
    const PC& GetPc() const { return pc_; }
    int       GetRestart() const { return restart_; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        GMRES(A, x, b, ex, pc_, restart_, iter_, res_, rel_, calculate2norm_, method_, output_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        GMRES(A, x, b, ex, pc_, restart_, numIter, resid, rel_, calculate2norm_, method_, output_);
    }
};

/// BiCGStab driver (src/num/krylovsolver.h:1978-2013); `mod` selects ModBICGSTAB in MPI builds only.
template <typename PC>
class BiCGStabSolverCL : public SolverBaseCL
{
  private:
    PC&  pc_;
    bool mod_;

  public:
    BiCGStabSolverCL(PC& pc, int maxiter, double tol, bool relative = true, bool mod = true)
        : SolverBaseCL(maxiter, tol, relative), pc_(pc), mod_(mod) {}

    PC&       GetPc()       { return pc_; }
    const PC& GetPc() const { return pc_; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        BICGSTAB(A, x, b, ex, pc_, iter_, res_, rel_);
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        BICGSTAB(A, x, b, ex, pc_, numIter, resid, rel_);
    }
};

/// IDR(s) driver (src/num/krylovsolver.h:2095-2135).
template <typename PC>
class IDRsSolverCL : public SolverBaseCL
{
  private:
    PC&          pc_;
    const int    s_;
    const double omega_bound_;

  public:
    IDRsSolverCL(PC& pc, int maxiter, double tol, bool relative = true, const int s = 4, const double omega_bound = 0.7,
                 std::ostream* output = 0)
        : SolverBaseCL(maxiter, tol, relative, output), pc_(pc), s_(s), omega_bound_(omega_bound) {}

    PC&       GetPc()       { return pc_; }
    const PC& GetPc() const { return pc_; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        IDRS(A, x, b, ex, pc_, iter_, res_, rel_, s_, omega_bound_);
        if (output_ != 0)
            *output_ << "IDRsSolverCL: iterations: " << GetIter() << "\tresidual: " << GetResid() << std::endl;
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        IDRS(A, x, b, ex, pc_, numIter, resid, rel_, s_, omega_bound_);
    }
};

/// GCR driver (src/num/krylovsolver.h:2017-2053).
template <typename PC>
class GCRSolverCL : public SolverBaseCL
{
  private:
    PC& pc_;
    int truncate_;

  public:
    GCRSolverCL(PC& pc, int truncate, int maxiter, double tol, bool relative = true, std::ostream* output = 0)
        : SolverBaseCL(maxiter, tol, relative, output), pc_(pc), truncate_(truncate) {}

    PC&       GetPc()       { return pc_; }
    const PC& GetPc() const { return pc_; }
    int       GetTruncate() const { return truncate_; }

    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex)
    {
        res_ = tol_; iter_ = maxiter_;
        GCR(A, x, b, ex, pc_, truncate_, iter_, res_, rel_, output_);
        if (output_ != 0)
            *output_ << "GCRSolverCL: iterations: " << GetIter() << "\tresidual: " << GetResid() << std::endl;
    }
    template <typename Mat, typename Vec, typename ExT>
    void Solve(const Mat& A, Vec& x, const Vec& b, const ExT& ex, int& numIter, double& resid) const
    {
        resid = tol_; numIter = maxiter_;
        GCR(A, x, b, ex, pc_, truncate_, numIter, resid, rel_, output_);
    }
};

}  // namespace DROPS

#endif

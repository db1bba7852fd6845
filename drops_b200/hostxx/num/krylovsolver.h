// drops_b200/hostxx/num/krylovsolver.h -- OVERLAY on DROPS' own src/num/krylovsolver.h.
//
// Pulls in the reference header unchanged (#include_next) -- all driver classes (CGSolverCL, PCGSolverCL<PC>,
// GMResSolverCL<PC>, MResSolverCL, PMResSolverCL<Lanczos>, BiCGStabSolverCL, GCRSolverCL, IDRsSolverCL,
// PCGNESolverCL, :1782-2135) and every other solver stay the reference's own code -- and adds, in namespace
// DROPS, MORE SPECIALISED overloads of the free function templates those drivers call:
//   CG :49-103, PCG :107-179, PCGNE :199-252, GMRES :320-461, MINRES :1155-1161, PMINRES :1069-1152,
//   BICGSTAB :1189-1258, GCR :1407-1483, IDRS :1587-1698
// for Mat in {MatrixCL, MLMatrixCL (finest level, spmat.h:1890-1894)}, Vec = VectorCL, the serial
// DummyExchangeCL and the preconditioners that exist on the device (JACPcCL, DummyPcCL, ChebyshevPcCL /
// ChebyshevBBTPcCL).  Partial ordering of function templates picks them over the reference's generic
// templates, so PCGSolverCL<JACPcCL>::Solve, GMResSolverCL<JACPcCL>::Solve, MResSolverCL::Solve ... -- and
// with them PoissonSolverFactoryCL's codes 202 / 302 / 207 (num/poissonsolverfactory.h:128-171, :234-236) --
// run on the GPU through the C ABI (include/drops_b200.h), while e.g. PCGSolverCL<SSORPcCL> (203) keeps
// running the reference's template (its products go through the GPU operator* of the spmat.h overlay).
//
// Contract of every overload = the reference's: x start vector in / solution out, max_iter IN limit / OUT
// iterations used, tol IN tolerance / OUT residual reached, bool "converged"; the drivers record iter_ / res_.
#ifndef DROPS_B200_OVERLAY_NUM_KRYLOVSOLVER_H
#define DROPS_B200_OVERLAY_NUM_KRYLOVSOLVER_H

#include "num/spmat.h"
#include "num/precond.h"            // the preconditioner types the overloads below are written for
#include_next "num/krylovsolver.h"

#include <cstring>
#include <type_traits>
#include <vector>

namespace DROPS
{
namespace b200
{

template <typename Mat> struct IsHostMatrix : std::false_type {};
template <> struct IsHostMatrix<MatrixCL> : std::true_type {};
template <> struct IsHostMatrix<MLMatrixCL> : std::true_type {};

/// The single-level matrix a solver works on.
inline const MatrixCL& finest(const MatrixCL& A) { return A; }
inline const MatrixCL& finest(const MLMatrixCL& A) { return A.GetFinest(); }

/// How a preconditioner type of num/precond.h maps onto the device solvers.  on_device = false: the
/// reference's template runs (host vector loops, products on the GPU).
template <typename PC> struct PcTraits { static const bool on_device = false; };

#ifndef _PAR
/// serial JACPcCL = PreGSCL<P_JAC0> (precond.h:403-427, :874): x_i = omega b_i / a_ii.  Its only data member is
/// the private `double _omega` with no accessor; it is read through the object representation.
template <> struct PcTraits<JACPcCL> {
    static const bool on_device = true;
    static int kind() { return DROPS_B200_PC_JACOBI; }
    static double omega(const JACPcCL& pc)
    {
        static_assert(sizeof(JACPcCL) == sizeof(double), "PreGSCL<P_JAC0,false> is expected to hold exactly _omega");
        double om;
        std::memcpy(&om, &pc, sizeof om);
        return om;
    }
    static void prepare(const JACPcCL&, drops_b200_matrix_t, bool) {}
};
#endif
template <> struct PcTraits<DummyPcCL> {       // precond.h:1032-1065
    static const bool on_device = true;
    static int kind() { return DROPS_B200_PC_DUMMY; }
    static double omega(const DummyPcCL&) { return 1.0; }
    static void prepare(const DummyPcCL&, drops_b200_matrix_t, bool) {}
};
/// ChebyshevPcCL (precond.h:1463-1529) keeps scale_, degree_, EigRatio_ protected: read them through
/// pointers to members formed in a derived class.
struct ChebyshevPeekCL : public ChebyshevPcCL {
    static double scale(const ChebyshevPcCL& c) { return c.*(&ChebyshevPeekCL::scale_); }
    static int degree(const ChebyshevPcCL& c) { return c.*(&ChebyshevPeekCL::degree_); }
    static double ratio(const ChebyshevPcCL& c) { return c.*(&ChebyshevPeekCL::EigRatio_); }
};
template <> struct PcTraits<ChebyshevPcCL> {
    static const bool on_device = true;
    static int kind() { return DROPS_B200_PC_CHEBYSHEV; }
    static double omega(const ChebyshevPcCL&) { return 1.0; }
    static void prepare(const ChebyshevPcCL& pc, drops_b200_matrix_t h, bool normal_eq)
    {
        check(drops_b200_matrix_set_chebyshev(h, normal_eq ? 1 : 0, ChebyshevPeekCL::scale(pc), ChebyshevPeekCL::degree(pc),
                                              ChebyshevPeekCL::ratio(pc)));
    }
};
template <> struct PcTraits<ChebyshevBBTPcCL> : public PcTraits<ChebyshevPcCL> {};   // precond.h:1531-1567
/// DiagPcCL (precond.h:618-642): x = D o b with the caller's vector D.  Its only data member is the private
/// reference `const VectorCL& D_` (stored as a pointer), read through the object representation; the vector is
/// sent to the device with every Solve (the reference re-reads it in every Apply, so it may change between solves).
template <> struct PcTraits<DiagPcCL> {
    static const bool on_device = true;
    static int kind() { return DROPS_B200_PC_DIAG; }
    static double omega(const DiagPcCL&) { return 1.0; }
    static void prepare(const DiagPcCL& pc, drops_b200_matrix_t h, bool normal_eq)
    {
        static_assert(sizeof(DiagPcCL) == sizeof(const VectorCL*), "DiagPcCL is expected to hold exactly the reference D_");
        if (normal_eq) throw DROPSErrCL("DiagPcCL: not a preconditioner of the normal equations");
        const VectorCL* D;
        std::memcpy(&D, &pc, sizeof D);
        size_t rows = 0;
        check(drops_b200_matrix_info(h, &rows, 0, 0, 0, 0, 0));
        if (D->size() != rows) throw DROPSErrCL("DiagPcCL: incompatible dimensions");           // precond.h:629
        check(drops_b200_matrix_set_diag_pc(h, Addr(*D)));
    }
};
/// Preconditioners the device runs inside PCG only: with them GMRES, BiCGStab, GCR ... keep the reference's templates.
template <typename PC> struct PcgOnly
    : std::integral_constant<bool, std::is_base_of<ChebyshevPcCL, PC>::value || std::is_same<DiagPcCL, PC>::value> {};

template <typename Mat, typename PC>
struct OnDevice : std::integral_constant<bool, IsHostMatrix<Mat>::value && PcTraits<PC>::on_device> {};

inline void check_dims(const MatrixCL& A, const VectorCL& x, const VectorCL& b, const char* who)
{
    if (A.num_rows() != A.num_cols() || x.size() != A.num_rows() || b.size() != A.num_rows())
        throw DROPSErrCL(std::string(who) + ": incompatible dimensions");
}
/// The residual log of the reference's loops ("PCG: 3 resid 1.2e-05", krylovsolver.h:88-89, :162-163).
inline void print_history(std::ostream* out, const char* tag, const std::vector<double>& hist, int iters)
{
    if (!out) return;
    for (int i = 1; i <= iters && i < static_cast<int>(hist.size()); ++i)
        (*out) << tag << ": " << i << " resid " << hist[i] << std::endl;
}
inline bool run_minres(const MatrixCL& A, VectorCL& x, const VectorCL& rhs, int precond, int kind, double omega,
                       int& max_iter, double& tol, bool rel)
{
    check_dims(A, x, rhs, "MINRES");
    int it = max_iter;
    const int r = drops_b200_minres(device(A), &x[0], Addr(rhs), precond, kind, omega, &it, &tol, rel ? 1 : 0);
    check(r);
    if (r == 1) max_iter = it;
    return r == 1;
}

}  // namespace b200

#define DROPS_B200_IF(...) typename std::enable_if<__VA_ARGS__, bool>::type

/// CG :49-103 on the device.
template <typename Mat>
DROPS_B200_IF(b200::IsHostMatrix<Mat>::value)
CG(const Mat& A, VectorCL& x, const VectorCL& b, const DummyExchangeCL&, int& max_iter, double& tol,
   bool measure_relative_tol, std::ostream* output = 0)
{
    const MatrixCL& M = b200::finest(A);
    b200::check_dims(M, x, b, "CG");
    std::vector<double> hist(output ? max_iter + 1 : 0);
    int it = max_iter;
    const int r = drops_b200_cg(b200::device(M), &x[0], Addr(b), &it, &tol, measure_relative_tol ? 1 : 0,
                                output ? hist.data() : 0, static_cast<int>(hist.size()));
    b200::check(r);
    if (r == 1) max_iter = it;
    b200::print_history(output, "CG", hist, r == 1 ? it : max_iter);
    return r == 1;
}

/// PCG :107-179 on the device (PCGSolverCL<JACPcCL>: factory code 202).
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value)
PCG(const Mat& A, VectorCL& x, const VectorCL& b, const DummyExchangeCL& ex, PreCon& M, int& max_iter, double& tol,
    bool measure_relative_tol = false, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "PCG");
    std::vector<double> hist(output ? max_iter + 1 : 0);
    int it = max_iter;
    if (M.NeedDiag()) M.SetDiag(A, ex);                 // :123-124 (ChebyshevPcCL: lambda_max, printed by the reference's code)
    const drops_b200_matrix_t h = b200::device(Mat0);
    b200::PcTraits<PreCon>::prepare(M, h, false);
    b200::TraceCL trace("PCG solve (vectors up, iterations, solution down)");
    const int r = drops_b200_pcg(h, &x[0], Addr(b), b200::PcTraits<PreCon>::kind(), b200::PcTraits<PreCon>::omega(M), &it, &tol,
                                 measure_relative_tol ? 1 : 0, output ? hist.data() : 0, static_cast<int>(hist.size()));
    b200::check(r);
    if (r == 1) max_iter = it;
    b200::print_history(output, "PCG", hist, r == 1 ? it : max_iter);
    return r == 1;
}

/// PCGNE :199-252 on the device: CG on A A^T u = b for a rectangular A.
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value)
PCGNE(const Mat& A, VectorCL& u, const VectorCL& b, const DummyExchangeCL& ExAX, const DummyExchangeCL& ExATranspX, PreCon& M,
      int& max_iter, double& tol, bool measure_relative_tol = false, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    if (u.size() != Mat0.num_rows() || b.size() != Mat0.num_rows()) throw DROPSErrCL("PCGNE: incompatible dimensions");
    std::vector<double> hist(output ? max_iter + 1 : 0);
    int it = max_iter;
    M.SetDiag(A, ExATranspX, ExAX);                     // :204
    const drops_b200_matrix_t h = b200::device(Mat0);
    b200::PcTraits<PreCon>::prepare(M, h, true);
    const int r = drops_b200_pcgne(h, &u[0], Addr(b), b200::PcTraits<PreCon>::kind(), b200::PcTraits<PreCon>::omega(M), &it, &tol,
                                   measure_relative_tol ? 1 : 0, output ? hist.data() : 0, static_cast<int>(hist.size()));
    b200::check(r);
    if (r == 1) max_iter = it;
    if (output) {                                       // :213-214, :240
        const int last = r == 1 ? it : max_iter;
        for (int i = 0; i <= last && i < static_cast<int>(hist.size()); ++i)
            if (i == 0 || i % 10 == 0) (*output) << "PCGNE: iter: " << i << " resid: " << hist[i] << '\n';
    }
    return r == 1;
}

#ifndef _PAR
/// serial GMRES(m) :320-461 on the device, modified Gram-Schmidt (GMResSolverCL<JACPcCL>: factory code 302).
/// calculate2norm (a debugging aid of the reference) is accepted and ignored.
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value && !b200::PcgOnly<PreCon>::value)
GMRES(const Mat& A, VectorCL& x, const VectorCL& b, const DummyExchangeCL&, PreCon& M, int m, int& max_iter, double& tol,
      bool measure_relative_tol = true, bool /*calculate2norm*/ = false, PreMethGMRES method = RightPreconditioning,
      std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "GMRES");
    int it = max_iter;
    const drops_b200_matrix_t h = b200::device(Mat0);
    b200::TraceCL trace("GMRES solve (vectors up, iterations, solution down)");
    const int r = drops_b200_gmres(h, &x[0], Addr(b), b200::PcTraits<PreCon>::kind(), b200::PcTraits<PreCon>::omega(M),
                                   m, &it, &tol, measure_relative_tol ? 1 : 0, method == LeftPreconditioning ? 1 : 0);
    b200::check(r);
    if (r == 1) max_iter = it;
    if (output)
        (*output) << (r == 1 ? "GMRES converged in " : "GMRES DIDN'T converge after ") << it
                  << " iterations with residual: " << tol << std::endl;
    return r == 1;
}
#endif

/// MINRES :1155-1161 on the device (MResSolverCL).
template <typename Mat>
DROPS_B200_IF(b200::IsHostMatrix<Mat>::value)
MINRES(const Mat& A, VectorCL& x, const VectorCL& rhs, const DummyExchangeCL&, int& max_iter, double& tol,
       bool measure_relative_tol = false)
{
    return b200::run_minres(b200::finest(A), x, rhs, 0, DROPS_B200_PC_DUMMY, 1.0, max_iter, tol, measure_relative_tol);
}

/// PMINRES :1069-1152 on the device.  The Lanczos object selects the flavour (and hands over its preconditioner);
/// the recurrence itself restarts on the device from rhs - A x, which is what q was built from
/// (PMResSolverCL::Solve :1906-1912, MINRES :1158-1159).
template <typename Mat>
DROPS_B200_IF(b200::IsHostMatrix<Mat>::value)
PMINRES(const Mat& A, VectorCL& x, const VectorCL& rhs, const DummyExchangeCL&, LanczosONBCL<VectorCL>&, int& max_iter,
        double& tol, bool measure_relative_tol = false)
{
    return b200::run_minres(b200::finest(A), x, rhs, 0, DROPS_B200_PC_DUMMY, 1.0, max_iter, tol, measure_relative_tol);
}
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value && !b200::PcgOnly<PreCon>::value)
PMINRES(const Mat& A, VectorCL& x, const VectorCL& rhs, const DummyExchangeCL&, PLanczosONBCL<VectorCL, PreCon>& q,
        int& max_iter, double& tol, bool measure_relative_tol = false)
{
    return b200::run_minres(b200::finest(A), x, rhs, 1, b200::PcTraits<PreCon>::kind(), b200::PcTraits<PreCon>::omega(q.M),
                            max_iter, tol, measure_relative_tol);
}

/// BICGSTAB :1189-1258 on the device.  On a breakdown the reference prints a message, sets max_iter to the
/// iteration and returns false; so does this.
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value && !b200::PcgOnly<PreCon>::value)
BICGSTAB(const Mat& A, VectorCL& x, const VectorCL& b, const DummyExchangeCL&, PreCon& M, int& max_iter, double& tol,
         bool measure_relative_tol = true)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "BICGSTAB");
    int it = max_iter, breakdown = 0;
    const int r = drops_b200_bicgstab(b200::device(Mat0), &x[0], Addr(b), b200::PcTraits<PreCon>::kind(),
                                      b200::PcTraits<PreCon>::omega(M), &it, &tol, measure_relative_tol ? 1 : 0, &breakdown);
    b200::check(r);
    if (breakdown == 1) std::cout << "BiCGSTAB: Breakdown with rho_1 = 0.\n";
    if (breakdown == 2) std::cout << "BiCGSTAB: Breakdown with omega_ = 0.\n";
    if (r == 1 || breakdown) max_iter = it;
    return r == 1;
}

/// GCR(m) :1407-1483 on the device (min-alpha truncation).
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value && !b200::PcgOnly<PreCon>::value)
GCR(const Mat& A, VectorCL& x, const VectorCL& b, const DummyExchangeCL&, PreCon& M, int m, int& max_iter, double& tol,
    bool measure_relative_tol = true, std::ostream* output = 0)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, b, "GCR");
    int it = max_iter;
    const int r = drops_b200_gcr(b200::device(Mat0), &x[0], Addr(b), b200::PcTraits<PreCon>::kind(), b200::PcTraits<PreCon>::omega(M),
                                 m, &it, &tol, measure_relative_tol ? 1 : 0);
    b200::check(r);
    if (r == 1) max_iter = it;
    if (output) (*output) << "GCR: k: " << it << "\tresidual: " << tol << std::endl;
    return r == 1;
}

/// IDR(s) :1587-1698 on the device.  The shadow space is drawn like the reference draws it (srand(time(0)),
/// rand()); outputs follow the reference literally: tol returns the ABSOLUTE residual norm, max_iter the
/// iterations done, and the bool is false when the tolerance was beaten inside the loop (:1689-1693).
template <typename Mat, typename PreCon>
DROPS_B200_IF(b200::OnDevice<Mat, PreCon>::value && !b200::PcgOnly<PreCon>::value)
IDRS(const Mat& A, VectorCL& x, const VectorCL& rhs, const DummyExchangeCL&, PreCon& pc, int& max_iter, double& tol,
     bool measure_relative_tol = false, const int s = 4, double omega_bound = 0.7)
{
    const MatrixCL& Mat0 = b200::finest(A);
    b200::check_dims(Mat0, x, rhs, "IDRS");
    const int r = drops_b200_idrs(b200::device(Mat0), &x[0], Addr(rhs), b200::PcTraits<PreCon>::kind(), b200::PcTraits<PreCon>::omega(pc),
                                  s, omega_bound, 0, &max_iter, &tol, measure_relative_tol ? 1 : 0);
    if (r < 0) throw DROPSErrCL(std::string(drops_b200_last_error()));     // "IDR(s): M(k,k) ==0" / "omega ==0"
    return r == 1;
}

#undef DROPS_B200_IF

}  // namespace DROPS

#endif

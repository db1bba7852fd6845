// drops_b200/hostxx/num/precond.h
//
// Preconditioners of the B200 drop-in.  The sparse-solver hot path knows two:
//   JACPcCL   x_i = omega * b_i / a_ii   -- serial DROPS: PreGSCL<P_JAC0> (src/num/precond.h:874,
//             :403-427, SolveGSstep<PB_JAC0> :153-169); MPI DROPS: the PreBaseCL subclass with a
//             cached diagonal (:1155-1188, Jacobi0 :1293-1311).  Same arithmetic in both.
//   DummyPcCL x = b                      -- src/num/precond.h:1032-1065
// Both keep the preconditioner concept the Krylov templates rely on: Apply(A, x, b, ex),
// RetAcc(), NeedDiag(), SetDiag(...).  Inside PCGSolverCL / GMResSolverCL / PMResSolverCL they
// are not called element by element: the solver reads b200::PcTraits<PC> and the device kernels
// apply the scaling fused with their other vector passes.  Apply() itself (used by DROPS code
// outside a solve, e.g. as a smoother) runs on the GPU through drops_b200_jacobi_apply.
// Next to them: the Chebyshev polynomial preconditioners ChebyshevPcCL
// (src/num/precond.h:1463-1529) and ChebyshevBBTPcCL (:1531-1567, for PCGNE) -- products with A / A^T
// only, so they run on the same SpMV kernels -- and BBTDiag (:56-75).
// Gauss-Seidel/SSOR are sequential sweeps and not part of this path.
#ifndef DROPS_PRECOND_H
#define DROPS_PRECOND_H

#include <iostream>

#include "num/spmat.h"

namespace DROPS
{

/// Jacobi with start vector 0 (one step): x = omega * D^-1 b.
class JACPcCL
{
  private:
    double omega_;

  public:
    JACPcCL(double om = 1.0) : omega_(om) {}

    double GetOmega() const { return omega_; }
    void   SetOmega(double om) { omega_ = om; }

    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat& A, Vec& x, const Vec& b, const ExT&) const
    {
        const MatrixCL& M = b200::finest(A);
        if (M.num_rows() != b.size()) throw DROPSErrCL("JACPcCL::Apply: incompatible dimensions");
        if (x.size() != b.size()) x.resize(b.size());
        if (b.size() == 0) return;
        drops_b200_vector_t db = 0, dx = 0;
        b200::ensure_init();
        b200::check(drops_b200_vector_create(b.size(), &db));
        int s = drops_b200_vector_create(b.size(), &dx);
        if (s >= 0) s = drops_b200_vector_upload(db, Addr(b));
        if (s >= 0) s = drops_b200_jacobi_apply(M.device(), omega_, db, dx);
        if (s >= 0) s = drops_b200_vector_download(dx, &x[0]);
        drops_b200_vector_destroy(db);
        drops_b200_vector_destroy(dx);
        b200::check(s);
    }
    template <typename Mat, typename Vec, typename ExT>
    void transp_Apply(const Mat& A, Vec& x, const Vec& b, const ExT& ex) const { Apply(A, x, b, ex); }

    bool RetAcc() const { return false; }    ///< Apply returns a distributed (not accumulated) vector
    bool NeedDiag() const { return false; }  ///< the diagonal lives with the device matrix
    void SetDiag(const VectorCL&) {}
    template <typename Mat, typename ExT>
    void SetDiag(const Mat&, const ExT&) {}
    template <typename Mat, typename ExT>
    void SetDiag(const Mat&, const ExT&, const ExT&) {}   ///< PCGNE: Jacobi on diag(A A^T), kept on the device
};

/// Identity preconditioner.
class DummyPcCL
{
  public:
    bool RetAcc() const { return false; }
    bool NeedDiag() const { return false; }
    void SetDiag(const VectorCL&) {}
    template <typename Mat, typename ExT>
    void SetDiag(const Mat&, const ExT&) {}
    template <typename Mat, typename ExT>
    void SetDiag(const Mat&, const ExT&, const ExT&) {}

    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat&, Vec& x, const Vec& b, const ExT&) const { x = b; }
    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat&, Vec& x, const Vec& b, const ExT&, const ExT&) const { x = b; }
    template <typename Mat, typename Vec, typename ExT>
    void transp_Apply(const Mat&, Vec& x, const Vec& b, const ExT&) const { x = b; }
};

/// diag(B B^T) (src/num/precond.h:56-75), computed on the device copy of B.
template <typename T, typename ExT>
VectorBaseCL<T> BBTDiag(const SparseMatBaseCL<T>& B, const ExT&, const ExT&)
{
    VectorBaseCL<T> ret(B.num_rows());
    if (B.num_rows()) b200::check(drops_b200_matrix_bbt_diag(B.device(), &ret[0]));
    return ret;
}

/// Chebyshev polynomial preconditioner x = p(D^-1 A) D^-1 b with start vector 0; lambda_max comes from
/// ten steps of the power method on D^-1 A (MaxEigenvalueScaledMatrix), scaled by lambda_scale.
class ChebyshevPcCL
{
  protected:
    double scale_;
    int    degree_;
    double EigRatio_;
    mutable double lambda_max_;

    template <typename Mat, typename Vec>
    void apply_on_device(const Mat& A, Vec& x, const Vec& b, bool normal_eq) const
    {
        const MatrixCL& M = b200::finest(A);
        if (M.num_rows() != b.size()) throw DROPSErrCL("ChebyshevPcCL::Apply: incompatible dimensions");
        if (x.size() != b.size()) x.resize(b.size());
        if (b.size() == 0 || degree_ == 0) return;
        configure(M.device(), normal_eq);
        drops_b200_vector_t db = 0, dx = 0;
        b200::check(drops_b200_vector_create(b.size(), &db));
        int s = drops_b200_vector_create(b.size(), &dx);
        if (s >= 0) s = drops_b200_vector_upload(db, Addr(b));
        if (s >= 0) s = drops_b200_precond_apply_dev(M.device(), DROPS_B200_PC_CHEBYSHEV, 1.0, normal_eq ? 1 : 0, db, dx);
        if (s >= 0) s = drops_b200_vector_download(dx, &x[0]);
        drops_b200_vector_destroy(db);
        drops_b200_vector_destroy(dx);
        b200::check(s);
    }
    template <typename Mat>
    void set_diag_on_device(const Mat& A, bool normal_eq, const char* who) const
    {
        const MatrixCL& M = b200::finest(A);
        configure(M.device(), normal_eq);
        const double before = lambda_max_;
        b200::check(drops_b200_chebyshev_lambda_max(M.device(), normal_eq ? 1 : 0, &lambda_max_));
        if (lambda_max_ != before) std::cout << who << ": lambda_max = " << lambda_max_ << std::endl;
    }

  public:
    ChebyshevPcCL(double lambda_scale = 1.1, int degree = 3, double ratio = 30.0)
        : scale_(lambda_scale), degree_(degree), EigRatio_(ratio), lambda_max_(0.0) {}

    /// Hands the constructor arguments to the device matrix (what the solvers call before a solve).
    void configure(drops_b200_matrix_t h, bool normal_eq) const
    {
        b200::check(drops_b200_matrix_set_chebyshev(h, normal_eq ? 1 : 0, scale_, degree_, EigRatio_));
    }
    double GetLambdaMax() const { return lambda_max_; }

    bool RetAcc() const { return true; }
    bool NeedDiag() const { return true; }
    void SetDiag(const VectorCL&) {}
    template <typename Mat, typename ExT>
    void SetDiag(const Mat& A, const ExT&) { set_diag_on_device(A, false, "ChebyshevPcCL"); }

    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat& A, Vec& x, const Vec& b, const ExT&) const { apply_on_device(A, x, b, false); }
};

/// The same polynomial on the normal equations B B^T (the PCGNE preconditioner), D = BBTDiag(B).
class ChebyshevBBTPcCL : public ChebyshevPcCL
{
  public:
    ChebyshevBBTPcCL(double lambda_scale = 1, int degree = 3, double ratio = 30.0)
        : ChebyshevPcCL(lambda_scale, degree, ratio) {}

    template <typename Mat, typename ExT>
    void SetDiag(const Mat& A, const ExT&, const ExT&) { set_diag_on_device(A, true, "ChebyshevBBTPcCL"); }

    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat& A, Vec& x, const Vec& b, const ExT&, const ExT&) const { apply_on_device(A, x, b, true); }
};

namespace b200
{
/// What the device solvers need to know about a preconditioner type.  Only the types above are on
/// the B200 path; any other PC makes the solver templates fail to compile instead of silently
/// running somewhere else.  prepare() runs before a solve (Chebyshev: constructor arguments to the
/// device matrix); normal_eq says whether the solver iterates on A A^T (PCGNE).
template <typename PC>
struct PcTraits;

template <>
struct PcTraits<JACPcCL> {
    static int kind() { return DROPS_B200_PC_JACOBI; }
    static double omega(const JACPcCL& pc) { return pc.GetOmega(); }
    static void prepare(const JACPcCL&, drops_b200_matrix_t, bool) {}
};
template <>
struct PcTraits<DummyPcCL> {
    static int kind() { return DROPS_B200_PC_DUMMY; }
    static double omega(const DummyPcCL&) { return 1.0; }
    static void prepare(const DummyPcCL&, drops_b200_matrix_t, bool) {}
};
template <>
struct PcTraits<ChebyshevPcCL> {
    static int kind() { return DROPS_B200_PC_CHEBYSHEV; }
    static double omega(const ChebyshevPcCL&) { return 1.0; }
    static void prepare(const ChebyshevPcCL& pc, drops_b200_matrix_t h, bool normal_eq) { pc.configure(h, normal_eq); }
};
template <>
struct PcTraits<ChebyshevBBTPcCL> {
    static int kind() { return DROPS_B200_PC_CHEBYSHEV; }
    static double omega(const ChebyshevBBTPcCL&) { return 1.0; }
    static void prepare(const ChebyshevBBTPcCL& pc, drops_b200_matrix_t h, bool normal_eq) { pc.configure(h, normal_eq); }
};
}  // namespace b200

}  // namespace DROPS

#endif

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
// Gauss-Seidel/SSOR/Chebyshev are sequential sweeps and not part of this path.
#ifndef DROPS_PRECOND_H
#define DROPS_PRECOND_H

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

    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat&, Vec& x, const Vec& b, const ExT&) const { x = b; }
    template <typename Mat, typename Vec, typename ExT>
    void Apply(const Mat&, Vec& x, const Vec& b, const ExT&, const ExT&) const { x = b; }
    template <typename Mat, typename Vec, typename ExT>
    void transp_Apply(const Mat&, Vec& x, const Vec& b, const ExT&) const { x = b; }
};

namespace b200
{
/// What the device solvers need to know about a preconditioner type.  Only the two types above
/// are on the B200 path; any other PC makes the solver templates fail to compile instead of
/// silently running somewhere else.
template <typename PC>
struct PcTraits;

template <>
struct PcTraits<JACPcCL> {
    static int kind() { return DROPS_B200_PC_JACOBI; }
    static double omega(const JACPcCL& pc) { return pc.GetOmega(); }
};
template <>
struct PcTraits<DummyPcCL> {
    static int kind() { return DROPS_B200_PC_DUMMY; }
    static double omega(const DummyPcCL&) { return 1.0; }
};
}  // namespace b200

}  // namespace DROPS

#endif

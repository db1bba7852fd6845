// drops_b200/hostxx/num/solverbase.h
//
// Settings/outcome record shared by the iterative solvers of the B200 drop-in.  It offers the
// interface DROPS callers use on src/num/solverbase.h:33-60 -- SetTol/SetMaxIter/SetRelError/
// SetOutput and GetTol/GetMaxIter/GetResid/GetIter/GetRelError, all virtual as there -- with the
// same meaning: after Solve(), GetIter() is the number of iterations used (the unchanged limit
// if the solver did not converge) and GetResid() the final residual, relative iff GetRelError().
#ifndef DROPS_SOLVERBASE_H
#define DROPS_SOLVERBASE_H

#include <iosfwd>

namespace DROPS
{

class SolverBaseCL
{
  public:
    // ---- settings ------------------------------------------------------------------------
    virtual void SetMaxIter(int limit) { maxiter_ = limit; }
    virtual void SetTol(double t) { tol_ = t; }
    virtual void SetRelError(bool relative) { rel_ = relative; }
    virtual void SetOutput(std::ostream* sink) { output_ = sink; }

    virtual int    GetMaxIter() const { return maxiter_; }
    virtual double GetTol() const { return tol_; }
    virtual bool   GetRelError() const { return rel_; }

    // ---- outcome of the last Solve ---------------------------------------------------------
    virtual int    GetIter() const { return iter_; }
    virtual double GetResid() const { return res_; }

  protected:
    SolverBaseCL(int maxiter, double tol, bool rel = false, std::ostream* output = 0)
        : maxiter_(maxiter), iter_(-1), tol_(tol), res_(-1.), rel_(rel), output_(output)
    {}
    virtual ~SolverBaseCL() {}

    int maxiter_;                    ///< iteration limit
    mutable int iter_;               ///< iterations used by the last Solve (-1: none yet)
    double tol_;                     ///< stopping tolerance
    mutable double res_;             ///< residual reached by the last Solve (-1: none yet)
    bool rel_;                       ///< measure the residual relative to the right-hand side
    mutable std::ostream* output_;   ///< per-iteration residual log, may be null
};

}  // namespace DROPS

#endif

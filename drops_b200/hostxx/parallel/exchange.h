// drops_b200/hostxx/parallel/exchange.h
//
// Serial exchange object of the B200 drop-in: the role DummyExchangeCL plays in
// src/parallel/exchange.h:45-85.  In a serial run every vector is already "accumulated", so
// Accumulate is a no-op, GetAccumulate a copy, and the inner products are compensated (Kahan)
// sums that optionally hand back copies of their arguments -- which is exactly how PCG obtains
// q_acc / z_acc (krylovsolver.h:144,151).  These members run on the host: they exist for DROPS
// code that calls them on host vectors outside a solve; inside Solve() the device kernels
// (double-double reductions, at least as accurate) take their place.
#ifndef DROPS_EXCHANGE_H
#define DROPS_EXCHANGE_H

#include <cmath>
#include <vector>

#include "num/spmat.h"

namespace DROPS
{

class DummyExchangeCL
{
  public:
    void Accumulate(VectorCL&, const Ulint = 0) const {}
    void Accumulate(std::vector<VectorCL>&) const {}
    VectorCL GetAccumulate(const VectorCL& u) const { return u; }
    std::vector<VectorCL> GetAccumulate(const std::vector<VectorCL>& u) const { return u; }

    /// (x, y) without a reduction over processes; copies x/y into *x_acc/*y_acc when asked to.
    double LocalDot(const VectorCL& x, bool, const VectorCL& y, bool, VectorCL* x_acc = 0, VectorCL* y_acc = 0) const
    {
        if (x_acc) { x_acc->resize(x.size()); *x_acc = x; }
        if (y_acc) { y_acc->resize(y.size()); *y_acc = y; }
        return KahanInnerProd(x, y, double());
    }
    double ParDot(const VectorCL& x, bool xa, const VectorCL& y, bool ya, VectorCL* x_acc = 0, VectorCL* y_acc = 0) const
    {
        return LocalDot(x, xa, y, ya, x_acc, y_acc);
    }
    double LocalNorm_sq(const VectorCL& x, bool a, VectorCL* x_acc = 0) const { return LocalDot(x, a, x, a, x_acc); }
    double Norm_sq(const VectorCL& x, bool a, VectorCL* x_acc = 0) const { return LocalNorm_sq(x, a, x_acc); }
    double Norm(const VectorCL& x, bool a, VectorCL* x_acc = 0) const { return std::sqrt(LocalNorm_sq(x, a, x_acc)); }
};

}  // namespace DROPS

#endif

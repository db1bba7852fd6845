// drops_b200/hostxx/num/spmat.h -- B200 drop-in for DROPS' src/num/spmat.h.
//
// Same names, argument meaning and error behaviour as the reference header for everything the
// sparse-solver hot path and its callers use:
//   VectorBaseCL<T> / VectorCL            src/num/spmat.h:82-145, :1909
//   dot, norm_sq, norm, supnorm           :164-210
//   axpy, z_xpay, z_xpaypby2              :212-237
//   KahanInnerProd                        :317-333
//   SparseMatBuilderCL (operator(), Build) :420-676
//   SparseMatBaseCL<T> / MatrixCL         :682-879 (raw_*, num_*, row_beg/col_ind/val, Version(),
//                                          operator(), *=, /=, resize, GetDiag)
//   operator*(A, x) = y_Ax                :1682-1724
//   transp_mul(A, x) = y_ATx              :1729-1760
//   MLSparseMatBaseCL / MLMatrixCL        :1870-1913 (finest level forwards)
// Storage stays host-side and reference-compatible (std::valarray, size_t indices) so the FE
// assembly of a DROPS executable keeps filling it; products and solves go to the GPU through the
// C ABI (include/drops_b200.h).  The device copy of a matrix is cached in the object and rebuilt
// when Version() changes (the reference's own change-detection idiom, :807-808).
#ifndef DROPS_SPMAT_H
#define DROPS_SPMAT_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <numeric>
#include <string>
#include <type_traits>
#include <valarray>
#include <vector>

#include "misc/utils.h"
#include "drops_b200.h"

namespace DROPS
{

namespace b200
{
/// C-ABI status -> DROPSErrCL (the reference reports errors by exception, misc/utils.h:369).
inline void check(int status)
{
    if (status < 0) throw DROPSErrCL(std::string("drops_b200: ") + drops_b200_last_error());
}
/// Bind the process to its GPU on first use (DROPS_B200_DEVICE, default LOCAL_RANK, default 0).
inline void ensure_init()
{
    static const bool once = [] {
        const char* d = std::getenv("DROPS_B200_DEVICE");
        if (!d) d = std::getenv("LOCAL_RANK");
        check(drops_b200_init(d ? std::atoi(d) : 0));
        return true;
    }();
    (void)once;
}
}  // namespace b200

//*****************************************************************************
//  VectorBaseCL: numerical vector on top of std::valarray
//*****************************************************************************
template <typename T>
class VectorBaseCL : public std::valarray<T>
{
  public:
    typedef T value_type;
    typedef std::valarray<T> base_type;

    VectorBaseCL() : base_type() {}
    VectorBaseCL(size_t s) : base_type(s) {}
    VectorBaseCL(T c, size_t s) : base_type(c, s) {}
    VectorBaseCL(const T* tp, size_t s) : base_type(tp, s) {}
    VectorBaseCL(const VectorBaseCL&) = default;
    VectorBaseCL(VectorBaseCL&&) = default;
    /// valarray expression templates, slices, plain valarrays
    template <class X, class = typename std::enable_if<!std::is_arithmetic<X>::value &&
                                                       std::is_constructible<base_type, const X&>::value>::type>
    explicit VectorBaseCL(const X& x) : base_type(x) {}

    VectorBaseCL& operator=(const VectorBaseCL&) = default;
    VectorBaseCL& operator=(VectorBaseCL&&) = default;

#define DROPS_B200_VEC_OP(op)                                                             \
    VectorBaseCL& operator op(const T s) { *static_cast<base_type*>(this) op s; return *this; } \
    template <class VT, class = typename std::enable_if<!std::is_arithmetic<VT>::value>::type> \
    VectorBaseCL& operator op(const VT& v) { *static_cast<base_type*>(this) op v; return *this; }
    DROPS_B200_VEC_OP(=)
    DROPS_B200_VEC_OP(+=)
    DROPS_B200_VEC_OP(-=)
    DROPS_B200_VEC_OP(*=)
    DROPS_B200_VEC_OP(/=)
#undef DROPS_B200_VEC_OP

    VectorBaseCL append(const VectorBaseCL& v) const
    {
        VectorBaseCL res(this->size() + v.size());
        for (size_t i = 0; i < this->size(); ++i) res[i] = (*this)[i];
        for (size_t i = 0; i < v.size(); ++i) res[this->size() + i] = v[i];
        return res;
    }
    void swap(VectorBaseCL& v) { std::swap(*static_cast<base_type*>(this), *static_cast<base_type*>(&v)); }

    const T* raw() const { return Addr(*this); }
    T*       raw()       { return &(*this)[0]; }
};

template <class T>
inline T dot(const VectorBaseCL<T>& v, const VectorBaseCL<T>& w)
{
    Assert(v.size() == w.size(), "dot: incompatible dimensions", DebugNumericC);
    return std::inner_product(Addr(v), Addr(v) + v.size(), Addr(w), T());
}

template <class VT>
inline typename VT::value_type norm_sq(const VT& v)
{
    typename VT::value_type ret = typename VT::value_type();
    for (size_t i = 0, s = v.size(); i < s; ++i) ret += v[i] * v[i];
    return ret;
}

template <class VT>
inline typename VT::value_type norm(const VT& v) { return std::sqrt(norm_sq(v)); }

template <class VT>
inline typename VT::value_type supnorm(const VT& v)
{
    typedef typename VT::value_type valueT;
    const size_t s = v.size();
    valueT ret = s == 0 ? valueT() : std::abs(v[0]);
    for (size_t i = 1; i < s; ++i) ret = std::max(ret, valueT(std::abs(v[i])));
    return ret;
}

template <typename T>
inline void axpy(T a, const VectorBaseCL<T>& x, VectorBaseCL<T>& y)
{
    Assert(x.size() == y.size(), "axpy: incompatible dimensions", DebugNumericC);
    y += a * x;
}

template <typename T>
inline void z_xpay(VectorBaseCL<T>& z, const VectorBaseCL<T>& x, T a, const VectorBaseCL<T>& y)
{
    Assert(z.size() == x.size() && z.size() == y.size(), "z_xpay: incompatible dimensions", DebugNumericC);
    z = x + a * y;
}

template <typename T>
inline void z_xpaypby2(VectorBaseCL<T>& z, const VectorBaseCL<T>& x, T a, const VectorBaseCL<T>& y, T b,
                       const VectorBaseCL<T>& y2)
{
    Assert(z.size() == x.size() && z.size() == y.size() && z.size() == y2.size(),
           "z_xpaypby2: incompatible dimensions", DebugNumericC);
    z = x + a * y + b * y2;
}

/// Compensated inner product: what DummyExchangeCL::ParDot / Norm are made of.
template <class ValueT, template <class> class VecT>
inline ValueT
#if defined(__GNUC__) && !defined(__clang__)
__attribute__((optimize("no-associative-math")))
#endif
KahanInnerProd(const VecT<ValueT>& a, const VecT<ValueT>& b, const ValueT init = ValueT(0))
{
    Assert(a.size() == b.size(), "KahanInnerProd: sizes don't match", DebugNumericC);
    ValueT sum = init, c = ValueT(0);
    for (size_t i = 0; i < a.size(); ++i) {
        const ValueT y = a[i] * b[i] - c;
        const ValueT t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
    return sum;
}

//*****************************************************************************
//  SparseMatBaseCL: CSR matrix, host storage + cached device operator
//*****************************************************************************
template <typename T>
class SparseMatBaseCL;

template <typename T = double>
class SparseMatBuilderCL
{
  private:
    SparseMatBaseCL<T>* mat_;
    size_t rows_, cols_;
    bool reuse_;
    std::vector<std::map<size_t, T> > coupl_;

  public:
    typedef T value_type;
    /// \param reuse keep the sparsity pattern of *mat and only reset its values (spmat.h:537-560)
    SparseMatBuilderCL(SparseMatBaseCL<T>* mat, size_t rows, size_t cols, bool reuse = false)
        : mat_(mat), rows_(rows), cols_(cols), reuse_(reuse && mat->num_rows() == rows && mat->num_cols() == cols)
    {
        mat_->IncrementVersion();
        if (reuse_) mat_->zero_values();
        else coupl_.resize(rows);
    }

    T& operator()(size_t i, size_t j)
    {
        Assert(i < rows_ && j < cols_, "SparseMatBuilderCL (): index out of bounds", DebugNumericC);
        if (reuse_) return mat_->entry_ref(i, j);
        return coupl_[i][j];
    }

    void Build()
    {
        if (reuse_) return;
        size_t nz = 0;
        for (size_t i = 0; i < rows_; ++i) nz += coupl_[i].size();
        mat_->resize(rows_, cols_, nz);
        size_t* rb = mat_->raw_row();
        size_t* ci = mat_->raw_col();
        T* va = mat_->raw_val();
        nz = 0;
        for (size_t i = 0; i < rows_; ++i) {
            rb[i] = nz;
            for (typename std::map<size_t, T>::const_iterator it = coupl_[i].begin(); it != coupl_[i].end(); ++it) {
                ci[nz] = it->first;
                va[nz] = it->second;
                ++nz;
            }
        }
        rb[rows_] = nz;
        coupl_.clear();
    }
};

template <typename T>
class SparseMatBaseCL
{
  private:
    size_t _rows, _cols, nnz_;
    size_t version_;
    std::valarray<size_t> _rowbeg;
    std::valarray<size_t> _colind;
    std::valarray<T>      _val;

    // device copy, keyed on version_
    mutable drops_b200_matrix_t dev_;
    mutable size_t dev_version_;
    mutable size_t dev_pattern_version_;
    size_t pattern_version_;

    void drop_device() const
    {
        if (dev_) { drops_b200_matrix_destroy(dev_); dev_ = 0; }
        dev_version_ = 0;
    }

  public:
    typedef T value_type;

    SparseMatBaseCL() : _rows(0), _cols(0), nnz_(0), version_(1), _rowbeg(size_t(0), 1), dev_(0), dev_version_(0),
                        dev_pattern_version_(0), pattern_version_(1) {}
    SparseMatBaseCL(size_t rows, size_t cols, size_t nnz = 0)   ///< fields allocated, not initialised
        : _rows(rows), _cols(cols), nnz_(nnz), version_(1), _rowbeg(rows + 1), _colind(nnz), _val(nnz), dev_(0),
          dev_version_(0), dev_pattern_version_(0), pattern_version_(1) {}
    SparseMatBaseCL(const std::valarray<T>& d)                   ///< square diagonal matrix
        : _rows(d.size()), _cols(d.size()), nnz_(d.size()), version_(1), _rowbeg(d.size() + 1), _colind(d.size()),
          _val(d), dev_(0), dev_version_(0), dev_pattern_version_(0), pattern_version_(1)
    {
        for (size_t i = 0; i < _rows; ++i) _rowbeg[i] = _colind[i] = i;
        _rowbeg[_rows] = _rows;
    }
    SparseMatBaseCL(const SparseMatBaseCL& m)
        : _rows(m._rows), _cols(m._cols), nnz_(m.nnz_), version_(1), _rowbeg(m._rowbeg), _colind(m._colind),
          _val(m._val), dev_(0), dev_version_(0), dev_pattern_version_(0), pattern_version_(1) {}
    ~SparseMatBaseCL() { drop_device(); }

    SparseMatBaseCL& operator=(const SparseMatBaseCL& m)
    {
        if (&m == this) return *this;
        IncrementVersion();
        ++pattern_version_;
        _rows = m._rows; _cols = m._cols; nnz_ = m.nnz_;
        _rowbeg.resize(m._rowbeg.size()); _rowbeg = m._rowbeg;
        _colind.resize(m._colind.size()); _colind = m._colind;
        _val.resize(m._val.size());       _val = m._val;
        return *this;
    }

    const T*      raw_val() const { return Addr(_val); }
    T*            raw_val()       { return Addr(_val); }
    const size_t* raw_row() const { return Addr(_rowbeg); }
    size_t*       raw_row()       { return Addr(_rowbeg); }
    const size_t* raw_col() const { return Addr(_colind); }
    size_t*       raw_col()       { return Addr(_colind); }

    size_t num_rows() const { return _rows; }
    size_t num_cols() const { return _cols; }
    size_t num_nonzeros() const { return nnz_; }

    size_t row_beg(size_t i) const { return _rowbeg[i]; }
    size_t col_ind(size_t i) const { return _colind[i]; }
    T      val(size_t i) const { return _val[i]; }

    void   IncrementVersion() { ++version_; }
    size_t Version() const { return version_; }

    const size_t* GetFirstCol(size_t i) const { return Addr(_colind) + _rowbeg[i]; }
    size_t*       GetFirstCol(size_t i)       { return Addr(_colind) + _rowbeg[i]; }
    const T*      GetFirstVal(size_t i) const { return Addr(_val) + _rowbeg[i]; }
    T*            GetFirstVal(size_t i)       { return Addr(_val) + _rowbeg[i]; }

    /// A(i,j) by binary search over the ascending columns of row i; T() if not stored.
    T operator()(size_t i, size_t j) const
    {
        Assert(i < num_rows() && j < num_cols(), "SparseMatBaseCL (): index out of bounds", DebugNumericC);
        const size_t* e = GetFirstCol(i + 1);
        const size_t* pos = std::lower_bound(GetFirstCol(i), e, j);
        return (pos != e && *pos == j) ? _val[pos - GetFirstCol(0)] : T();
    }

    SparseMatBaseCL& operator*=(T c) { IncrementVersion(); _val *= c; return *this; }
    SparseMatBaseCL& operator/=(T c) { IncrementVersion(); _val /= c; return *this; }

    /// New dimensions; the old content is lost.
    void resize(size_t rows, size_t cols, size_t nnz)
    {
        IncrementVersion();
        ++pattern_version_;
        _rows = rows; _cols = cols; nnz_ = nnz;
        _rowbeg.resize(rows + 1); _colind.resize(nnz); _val.resize(nnz);
    }
    void clear() { resize(0, 0, 0); }

    /// *this = coeffA*A + coeffB*B, pattern = row-wise merge of the operands' ascending column lists
    /// (src/num/spmat.h:1457-1551).  DROPS builds its time-step matrix this way right before every
    /// solve (levelset/coupling.tpp:96, poisson/cdrdrops.cpp:112).  The host copy is merged as in the
    /// reference; when both operands are resident on the GPU the device copy is merged THERE
    /// (drops_b200_matrix_lincomb), so the next Solve needs no upload.
    SparseMatBaseCL& LinComb(double coeffA, const SparseMatBaseCL& A, double coeffB, const SparseMatBaseCL& B)
    {
        if (A.num_rows() != B.num_rows() || A.num_cols() != B.num_cols())
            throw DROPSErrCL("LinComb: incompatible dimensions");
        const size_t rows = A.num_rows();
        std::valarray<size_t> rb(rows + 1);
        rb[0] = 0;
        for (size_t row = 0; row < rows; ++row) {
            const size_t *rA = A.GetFirstCol(row), *rB = B.GetFirstCol(row);
            const size_t* const rAend = A.GetFirstCol(row + 1);
            const size_t* const rBend = B.GetFirstCol(row + 1);
            size_t i = 0;
            for (; rA != rAend && rB != rBend; ++i)
                if (*rB < *rA) ++rB;
                else if (*rA < *rB) ++rA;
                else { ++rA; ++rB; }
            rb[row + 1] = rb[row] + i + (rAend - rA) + (rBend - rB);
        }
        const size_t nnz = rb[rows];
        std::valarray<size_t> ci(nnz);
        std::valarray<T> va(nnz);
#pragma omp parallel for
        for (long row = 0; row < (long)rows; ++row) {
            size_t i = rb[row], iA = A.row_beg(row), iB = B.row_beg(row);
            const size_t rAend = A.row_beg(row + 1), rBend = B.row_beg(row + 1);
            for (; iA != rAend && iB != rBend; ++i)
                if (B.col_ind(iB) < A.col_ind(iA)) { va[i] = coeffB * B._val[iB]; ci[i] = B._colind[iB++]; }
                else if (A.col_ind(iA) < B.col_ind(iB)) { va[i] = coeffA * A._val[iA]; ci[i] = A._colind[iA++]; }
                else { va[i] = coeffA * A._val[iA++] + coeffB * B._val[iB]; ci[i] = B._colind[iB++]; }
            for (; iA < rAend; ++iA, ++i) { va[i] = coeffA * A._val[iA]; ci[i] = A._colind[iA]; }
            for (; iB < rBend; ++iB, ++i) { va[i] = coeffB * B._val[iB]; ci[i] = B._colind[iB]; }
        }
        const bool on_device = &A != this && &B != this && A.dev_ && A.dev_version_ == A.version_ && B.dev_ &&
                               B.dev_version_ == B.version_;
        drops_b200_matrix_t merged = 0;
        IncrementVersion();
        ++pattern_version_;
        if (on_device) b200::check(drops_b200_matrix_lincomb(coeffA, A.dev_, coeffB, B.dev_, version_, &merged));
        _rows = rows; _cols = A.num_cols(); nnz_ = nnz;
        _rowbeg.resize(rows + 1); _rowbeg = rb;
        _colind.resize(nnz);      _colind = ci;
        _val.resize(nnz);         _val = va;
        drop_device();
        if (merged) { dev_ = merged; dev_version_ = version_; dev_pattern_version_ = pattern_version_; }
        return *this;
    }
    /// true when the device copy is current (no upload pending)
    bool DeviceIsCurrent() const { return dev_ && dev_version_ == version_; }

    /// Writes the matrix in MatrixMarket coordinate format (src/num/spmat.h:1085-1089).
    SparseMatBaseCL& exportMTX(const std::string& path);
    VectorBaseCL<T> GetDiag() const
    {
        Assert(num_rows() == num_cols(), "SparseMatBaseCL::GetDiag: no square Matrix", DebugParallelC);
        VectorBaseCL<T> d(num_rows());
        for (size_t i = 0; i < num_rows(); ++i) d[i] = (*this)(i, i);
        return d;
    }

    // --- used by SparseMatBuilderCL in pattern-reuse mode -------------------------------------
    void zero_values() { _val = T(); }
    T& entry_ref(size_t i, size_t j)
    {
        size_t* e = GetFirstCol(i + 1);
        size_t* pos = std::lower_bound(GetFirstCol(i), e, j);
        if (pos == e || *pos != j) throw DROPSErrCL("SparseMatBuilderCL: entry not in the reused pattern");
        return _val[pos - GetFirstCol(0)];
    }

    // --- the device side --------------------------------------------------------------------------
    /// Handle of the device copy; uploaded on first use, values refreshed when only they changed,
    /// rebuilt when the pattern changed.
    drops_b200_matrix_t device() const
    {
        static_assert(std::is_same<T, double>::value, "the B200 path computes in double");
        b200::ensure_init();
        if (dev_ && dev_version_ == version_) return dev_;
        if (dev_ && dev_pattern_version_ == pattern_version_) {
            b200::check(drops_b200_matrix_update_values(dev_, raw_val(), version_));
        } else {
            drop_device();
            b200::check(drops_b200_matrix_create(_rows, _cols, nnz_, raw_row(), raw_col(), raw_val(), version_, &dev_));
            dev_pattern_version_ = pattern_version_;
        }
        dev_version_ = version_;
        return dev_;
    }
    /// Call after writing through raw_col()/raw_row() on an existing matrix (new pattern).
    void PatternChanged() { ++pattern_version_; IncrementVersion(); }
};

/// y = A*x on the GPU (reference: y_Ax, per-row left-to-right sums; bit-identical).
template <typename _VecEntry>
VectorBaseCL<_VecEntry> operator*(const SparseMatBaseCL<double>& A, const VectorBaseCL<_VecEntry>& x)
{
    if (A.num_cols() != x.size()) throw DROPSErrCL("SparseMatBaseCL * VectorBaseCL: incompatible dimensions");
    VectorBaseCL<_VecEntry> ret(A.num_rows());
    if (A.num_rows() == 0) return ret;
    b200::check(drops_b200_spmv(A.device(), Addr(x), &ret[0]));
    return ret;
}

/// y = A^T*x on the GPU (reference: y_ATx scatter in ascending row order; bit-identical).
template <typename _VecEntry>
VectorBaseCL<_VecEntry> transp_mul(const SparseMatBaseCL<double>& A, const VectorBaseCL<_VecEntry>& x)
{
    if (A.num_rows() != x.size()) throw DROPSErrCL("transp_mul: incompatible dimensions");
    VectorBaseCL<_VecEntry> ret(_VecEntry(), A.num_cols());
    if (A.num_cols() == 0 || A.num_rows() == 0) return ret;
    b200::check(drops_b200_spmv_transpose(A.device(), Addr(x), &ret[0]));
    return ret;
}

//*****************************************************************************
//  MLSparseMatBaseCL: one matrix per multigrid level; products use the finest
//*****************************************************************************
template <typename T>
class MLSparseMatBaseCL : public std::list<SparseMatBaseCL<T> >
{
  public:
    typedef std::list<SparseMatBaseCL<T> > base_;
    explicit MLSparseMatBaseCL(size_t lvl = 1) : base_(lvl) {}
    SparseMatBaseCL<T>&       GetFinest()       { return this->back(); }
    const SparseMatBaseCL<T>& GetFinest() const { return this->back(); }
    SparseMatBaseCL<T>*       GetFinestPtr()       { return &this->back(); }
    const SparseMatBaseCL<T>* GetFinestPtr() const { return &this->back(); }
    size_t num_rows() const { return GetFinest().num_rows(); }
    size_t num_cols() const { return GetFinest().num_cols(); }
    size_t Version() const { return GetFinest().Version(); }
    VectorBaseCL<T> GetDiag() const { return GetFinest().GetDiag(); }
};

template <typename _VecEntry>
VectorBaseCL<_VecEntry> operator*(const MLSparseMatBaseCL<double>& A, const VectorBaseCL<_VecEntry>& x)
{
    return A.GetFinest() * x;
}
template <typename _VecEntry>
VectorBaseCL<_VecEntry> transp_mul(const MLSparseMatBaseCL<double>& A, const VectorBaseCL<_VecEntry>& x)
{
    return transp_mul(A.GetFinest(), x);
}

//*****************************************************************************
//  Matrix / vector text I/O (src/num/spmat.h:1196-1288): same formats as the reference, so matrices
//  written by a DROPS run (exportMTX :1085-1089, out) replay through this path and vice versa.
//*****************************************************************************
/// MatrixMarket coordinate output, 1-based, 15 digits (:1213-1224).
template <typename T>
std::ostream& operator<<(std::ostream& os, const SparseMatBaseCL<T>& A)
{
    const size_t M = A.num_rows();
    os << "%%MatrixMarket matrix coordinate real general\n" << A.num_rows() << ' ' << A.num_cols() << ' '
       << A.num_nonzeros() << '\n';
    os.precision(15);
    for (size_t row = 0; row < M; ++row)
        for (size_t col = A.row_beg(row), rowend = A.row_beg(row + 1); col < rowend; ++col)
            os << row + 1 << ' ' << A.col_ind(col) + 1 << ' ' << std::scientific << A.val(col) << '\n';
    return os << std::flush;
}

template <typename T>
SparseMatBaseCL<T>& SparseMatBaseCL<T>::exportMTX(const std::string& path)
{
    std::ofstream out(path);
    out << *this;
    return *this;
}

/// Plain CSR dump, readable by in() (:1227-1238).  Uses the stream's current precision.
template <typename T>
void out(std::ostream& os, const SparseMatBaseCL<T>& A)
{
    os << A.num_rows() << ' ' << A.num_cols() << ' ' << A.num_nonzeros() << '\n';
    for (size_t row = 0; row <= A.num_rows(); ++row) os << A.row_beg(row) << ' ';
    os << '\n';
    for (size_t nz = 0; nz < A.num_nonzeros(); ++nz) os << A.col_ind(nz) << ' ';
    os << '\n';
    for (size_t nz = 0; nz < A.num_nonzeros(); ++nz) os << A.val(nz) << ' ';
    os << std::endl;
}

/// Reads the output of out() (:1241-1255).
template <typename T>
void in(std::istream& is, SparseMatBaseCL<T>& A)
{
    size_t numrows, numcols, numnz;
    is >> numrows >> numcols >> numnz;
    A.resize(numrows, numcols, numnz);
    T* val = A.raw_val();
    size_t* row_beg = A.raw_row();
    size_t* col_ind = A.raw_col();
    for (size_t row = 0; row <= numrows; ++row) is >> row_beg[row];
    for (size_t nz = 0; nz < numnz; ++nz) is >> col_ind[nz];
    for (size_t nz = 0; nz < numnz; ++nz) is >> val[nz];
}

/// Reads "% rows x cols nonzeros" + 1-based triplets (:1258-1288; the older DROPS text format --
/// like the reference, it does NOT read the MatrixMarket header that operator<< writes today).
template <typename T>
std::istream& operator>>(std::istream& in, SparseMatBaseCL<T>& A)
{
    size_t numrows, numcols, numnz;
    while (in && in.get() != '%') ;
    in >> numrows;
    while (in && in.get() != 'x') ;
    in >> numcols >> numnz;
    while (in && in.get() != '\n') ;
    if (!in) throw DROPSErrCL("SparseMatBaseCL operator>>: Missing \"% rows cols nz\" comment.\n");
    SparseMatBuilderCL<T> B(&A, numrows, numcols);
    size_t r, c;
    T v;
    for (size_t nz = 0; nz < numnz; ++nz) {
        in >> r >> c >> v;
        if (!in) throw DROPSErrCL("SparseMatBaseCL operator>>: Stream corrupt.\n");
        if (r > numrows || c > numcols) {
            std::cout << "r: " << r << "\tc: " << c << "\tv: " << v << "\tnz: " << nz << '\n';
            throw DROPSErrCL("SparseMatBaseCL operator>>: Inconsistent data.\n");
        }
        B(r - 1, c - 1) = v;
    }
    B.Build();
    return in;
}

/// Vector I/O (:1196-1210).
template <typename T>
void out(std::ostream& os, const VectorBaseCL<T>& v)
{
    os << v.size() << '\n';
    for (size_t i = 0; i < v.size(); ++i) os << v[i] << ' ';
    os << std::endl;
}
template <typename T>
void in(std::istream& is, VectorBaseCL<T>& v)
{
    size_t s;
    is >> s;
    v.resize(s);
    for (size_t i = 0; i < s; ++i) is >> v[i];
}

typedef VectorBaseCL<double>       VectorCL;
typedef SparseMatBaseCL<double>    MatrixCL;
typedef SparseMatBuilderCL<double> MatrixBuilderCL;
typedef MLSparseMatBaseCL<double>  MLMatrixCL;

namespace b200
{
/// The single-level matrix a solver works on (MLMatrixCL -> finest level, spmat.h:1890-1894).
inline const MatrixCL& finest(const MatrixCL& A) { return A; }
inline const MatrixCL& finest(const MLMatrixCL& A) { return A.GetFinest(); }
}  // namespace b200

}  // namespace DROPS

#endif

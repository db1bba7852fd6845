// drops_b200/hostxx/num/spblockmat.h
//
// B200 drop-in for DROPS' src/num/spblockmat.h: the 2x2 block operator ( A B ; C D ) and the product of
// two operators, as the Stokes / Oseen / Schur-complement code uses them (BlockMatrixBaseCL :42-79,
// operator* :225-267, transp_mul :269-273, BuildMatrix :282-313; CompositeMatrixBaseCL :322-374,
// operator* :376-397, transp_mul :399-404).  The objects hold pointers and MUL / TRANSP_MUL flags only;
// a product hands the device handles of the blocks to drops_b200_block_spmv / drops_b200_composite_spmv,
// so the whole saddle-point product runs on the GPU (x up, y down once) with the reference's arithmetic:
// every block product is the bit-identical y_Ax / y_ATx, the second block of a row is added with one
// rounded add per entry.
#ifndef DROPS_SPBLOCKMAT_H
#define DROPS_SPBLOCKMAT_H

#include "num/spmat.h"
#include "parallel/exchange.h"

namespace DROPS
{

enum BlockMatrixOperationT { MUL, TRANSP_MUL };

namespace b200
{
inline int c_op(BlockMatrixOperationT op) { return op == MUL ? DROPS_B200_OP_MUL : DROPS_B200_OP_TRANSP_MUL; }
inline BlockMatrixOperationT flip(BlockMatrixOperationT op) { return op == MUL ? TRANSP_MUL : MUL; }
template <typename MatT>
inline size_t applied_rows(const MatT& m, BlockMatrixOperationT op) { return op == MUL ? m.num_rows() : m.num_cols(); }
template <typename MatT>
inline size_t applied_cols(const MatT& m, BlockMatrixOperationT op) { return op == MUL ? m.num_cols() : m.num_rows(); }
}  // namespace b200

template <class MatT>
class BlockMatrixBaseCL
{
  public:
    typedef BlockMatrixOperationT OperationT;

  private:
    const MatT* block_[4];      // row-major: A B / C D, null = zero block
    OperationT  operation_[4];

    // size of a block row / column: taken from the first block of it that exists
    size_t extent(size_t first, size_t second, bool rows) const
    {
        const size_t cand[2] = {first, second};
        for (int k = 0; k < 2; ++k) {
            const size_t b = cand[k];
            if (block_[b])
                return rows ? b200::applied_rows(*block_[b], operation_[b]) : b200::applied_cols(*block_[b], operation_[b]);
        }
        throw DROPSErrCL(rows ? "BlockMatrixBaseCL::num_rows: All pointers are 0.\n"
                              : "BlockMatrixBaseCL::num_cols: All pointers are 0.\n");
    }

  public:
    BlockMatrixBaseCL(const MatT* A, OperationT Aop, const MatT* B, OperationT Bop, const MatT* C, OperationT Cop,
                      const MatT* D = 0, OperationT Dop = MUL)
    {
        const MatT* m[4] = {A, B, C, D};
        const OperationT o[4] = {Aop, Bop, Cop, Dop};
        for (int i = 0; i < 4; ++i) { block_[i] = m[i]; operation_[i] = o[i]; }
    }

    size_t num_rows(size_t block_row) const { return block_row < 2 ? extent(2 * block_row, 2 * block_row + 1, true) : 0; }
    size_t num_cols(size_t block_col) const { return block_col < 2 ? extent(block_col, block_col + 2, false) : 0; }
    size_t num_rows() const { return num_rows(0) + num_rows(1); }
    size_t num_cols() const { return num_cols(0) + num_cols(1); }

    const MatT* GetBlock(size_t b) const { return block_[b]; }
    void        SetBlock(size_t b, const MatT* m) { block_[b] = m; }
    OperationT  GetOperation(size_t b) const { return operation_[b]; }
    OperationT  GetTransposeOperation(size_t b) const { return b200::flip(operation_[b]); }

    /// ( A B ; C D )^T = ( A^T C^T ; B^T D^T )
    BlockMatrixBaseCL GetTranspose() const
    {
        return BlockMatrixBaseCL(block_[0], GetTransposeOperation(0), block_[2], GetTransposeOperation(2),
                                 block_[1], GetTransposeOperation(1), block_[3], GetTransposeOperation(3));
    }

    /// Diagonal of the (square) block matrix; a missing diagonal block counts as the identity.
    VectorBaseCL<typename MatT::value_type> GetDiag() const
    {
        VectorBaseCL<typename MatT::value_type> diag(num_rows());
        size_t off = 0;
        for (size_t k = 0; k < 2; ++k) {
            const size_t len = num_rows(k);
            if (block_[3 * k]) diag[std::slice(off, len, 1)] = block_[3 * k]->GetDiag();
            else diag[std::slice(off, len, 1)] = 1.;
            off += len;
        }
        return diag;
    }
};

namespace b200
{
template <typename Mat>
inline void block_handles(const BlockMatrixBaseCL<Mat>& S, drops_b200_matrix_t h[4], int op[4])
{
    for (size_t b = 0; b < 4; ++b) {
        h[b] = S.GetBlock(b) ? finest(*S.GetBlock(b)).device() : 0;
        op[b] = c_op(S.GetOperation(b));
    }
}
}  // namespace b200

/// ( A B ; C D ) * x on the GPU.
template <typename Mat, typename Vec>
Vec operator*(const BlockMatrixBaseCL<Mat>& S, const Vec& x)
{
    if (x.size() != S.num_cols()) throw DROPSErrCL("BlockMatrixBaseCL * Vec: incompatible dimensions");
    Vec ret(S.num_rows());
    if (ret.size() == 0) return ret;
    drops_b200_matrix_t h[4];
    int op[4];
    b200::block_handles(S, h, op);
    b200::check(drops_b200_block_spmv(h, op, Addr(x), &ret[0]));
    return ret;
}

template <typename Mat, typename Vec>
Vec transp_mul(const BlockMatrixBaseCL<Mat>& S, const Vec& x)
{
    return S.GetTranspose() * x;
}

/// Assembles the block matrix into one CSR matrix (host side; setup code, not on the hot path).
template <typename Mat>
Mat BuildMatrix(BlockMatrixBaseCL<Mat>& S)
{
    Mat ret;
    SparseMatBuilderCL<typename Mat::value_type> bld(&ret, S.num_rows(), S.num_cols());
    for (int b = 0; b < 4; ++b) {
        const Mat* m = S.GetBlock(b);
        if (!m) continue;
        const size_t r0 = b / 2 ? S.num_rows(0) : 0, c0 = b % 2 ? S.num_cols(0) : 0;
        const bool plain = S.GetOperation(b) == MUL;
        for (size_t i = 0; i < m->num_rows(); ++i)
            for (size_t nz = m->row_beg(i); nz < m->row_beg(i + 1); ++nz) {
                const size_t j = m->col_ind(nz);
                if (plain) bld(r0 + i, c0 + j) = m->val(nz);
                else bld(r0 + j, c0 + i) = m->val(nz);
            }
    }
    bld.Build();
    return ret;
}

/// Product of two operators, applied as block1 * (block0 * v).
template <class MatT0, class MatT1, class ExT>
class CompositeMatrixBaseCL
{
  public:
    typedef BlockMatrixOperationT OperationT;

  private:
    const ExT& ex0_;
    const ExT& ex1_;
    const MatT0* block0_;
    const MatT1* block1_;
    OperationT operation_[2];

  public:
    CompositeMatrixBaseCL(const MatT0* A, OperationT Aop, const ExT& ex0, const MatT1* B, OperationT Bop, const ExT& ex1)
        : ex0_(ex0), ex1_(ex1), block0_(A), block1_(B)
    {
        operation_[0] = Aop;
        operation_[1] = Bop;
    }

    size_t num_rows() const { return b200::applied_rows(*block1_, operation_[1]); }
    size_t num_cols() const { return b200::applied_cols(*block0_, operation_[0]); }
    size_t intermediate_dim() const { return b200::applied_rows(*block0_, operation_[0]); }

    const ExT& GetEx0() const { return ex0_; }
    const ExT& GetEx1() const { return ex1_; }
    VectorCL GetDiag() const { throw DROPSErrCL("CompositeMatrixBaseCL:GetDiag(): Not implemented"); }

    const MatT0* GetBlock0() const { return block0_; }
    const MatT1* GetBlock1() const { return block1_; }
    void SetBlock0(const MatT0* p) { block0_ = p; }
    void SetBlock1(const MatT1* p) { block1_ = p; }
    OperationT GetOperation(size_t b) const { return operation_[b]; }
    OperationT GetTransposeOperation(size_t b) const { return b200::flip(operation_[b]); }

    CompositeMatrixBaseCL<MatT1, MatT0, ExT> GetTranspose() const
    {
        return CompositeMatrixBaseCL<MatT1, MatT0, ExT>(block1_, GetTransposeOperation(1), ex1_,
                                                        block0_, GetTransposeOperation(0), ex0_);
    }
    size_t Version() const { return block0_->Version(); }
};

/// block1 * (block0 * x): both products on the GPU, the intermediate vector never leaves it.
template <typename _MatT0, typename _MatT1, class ExT, typename _VecEntry>
VectorBaseCL<_VecEntry> operator*(const CompositeMatrixBaseCL<_MatT0, _MatT1, ExT>& C, const VectorBaseCL<_VecEntry>& x)
{
    if (x.size() != C.num_cols()) throw DROPSErrCL("CompositeMatrixBaseCL * Vec: incompatible dimensions");
    VectorBaseCL<_VecEntry> ret(C.num_rows());
    if (ret.size() == 0) return ret;
    b200::check(drops_b200_composite_spmv(b200::finest(*C.GetBlock0()).device(), b200::c_op(C.GetOperation(0)),
                                          b200::finest(*C.GetBlock1()).device(), b200::c_op(C.GetOperation(1)),
                                          Addr(x), &ret[0]));
    return ret;
}

template <typename _MatT0, typename _MatT1, class ExT, typename _VecEntry>
VectorBaseCL<_VecEntry> transp_mul(const CompositeMatrixBaseCL<_MatT0, _MatT1, ExT>& C, const VectorBaseCL<_VecEntry>& x)
{
    return C.GetTranspose() * x;
}

typedef BlockMatrixBaseCL<MatrixCL>   BlockMatrixCL;
typedef BlockMatrixBaseCL<MLMatrixCL> MLBlockMatrixCL;
typedef CompositeMatrixBaseCL<MatrixCL, MatrixCL, DummyExchangeCL> CompositeMatrixCL;

}  // namespace DROPS

#endif

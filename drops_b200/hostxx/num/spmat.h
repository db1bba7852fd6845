// drops_b200/hostxx/num/spmat.h -- OVERLAY on DROPS' own src/num/spmat.h.
//
// Put drops_b200/hostxx BEFORE $DROPS/src on the include path.  This file pulls in the reference header
// unchanged (#include_next) and adds, in namespace DROPS, exactly what moves the sparse products of the hot
// path to the GPU:
//   operator*(const MatrixCL&, const VectorCL&)      replaces y_Ax   / operator*   src/num/spmat.h:1682-1724
//   transp_mul(const MatrixCL&, const VectorCL&)     replaces y_ATx  / transp_mul  src/num/spmat.h:1729-1760
// Both are NON-template overloads: overload resolution prefers them to the reference's function templates
// wherever the arguments are MatrixCL / VectorCL -- also inside the reference's own templates (MLMatrixCL
// products :1878-1894, BlockMatrixBaseCL / CompositeMatrixBaseCL num/spblockmat.h:225-273, :376-420, every
// Krylov template of num/krylovsolver.h), whose unqualified calls are resolved at instantiation by
// argument-dependent lookup.  Everything else of the header (containers, assembly, LinComb, I/O, float
// matrices, ...) is the reference's code, untouched: every DROPS translation unit keeps compiling.
//
// The device copy of a matrix lives in a side table keyed on the object's address and validated by
// SparseMatBaseCL::Version() (the reference's change-detection idiom, spmat.h:807-808), the array
// addresses, the sizes and a sampled checksum of pattern and values -- SparseMatBaseCL gets no new member.
// The checksum exists because every new matrix starts at version 1 (an object re-created at the same
// address would otherwise alias a stale copy) and because a few reference routines write through raw_val()
// without IncrementVersion() (num/renumber.tpp:284-317): a sample of 4096 entries catches wholesale
// changes; DROPS_B200_FULL_HASH=1 checks every entry (debugging).  After a partial in-place edit through
// the raw pointers call A.IncrementVersion(), as the reference's own mutators do.
#ifndef DROPS_B200_OVERLAY_NUM_SPMAT_H
#define DROPS_B200_OVERLAY_NUM_SPMAT_H

#include_next "num/spmat.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>

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

/// DROPS_B200_TRACE=1: one line on stderr per one-time cost (CUDA context, matrix upload) and per solve, so that
/// the time a DROPS driver prints around solver.Solve (e.g. poisson/cdrdrops.cpp:117-119) can be taken apart.
inline bool trace_on()
{
    static const bool on = [] { const char* t = std::getenv("DROPS_B200_TRACE"); return t && std::atoi(t) != 0; }();
    return on;
}
class TraceCL
{
  private:
    const char* what_;
    std::chrono::steady_clock::time_point t0_;
  public:
    explicit TraceCL(const char* what) : what_(what), t0_(std::chrono::steady_clock::now()) {}
    ~TraceCL()
    {
        if (!trace_on()) return;
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
        std::fprintf(stderr, "drops_b200 trace: %s %.6f s\n", what_, s);
    }
};

/// Bind the process to its GPU on first use (DROPS_B200_DEVICE, else LOCAL_RANK, else 0).  There is no CPU
/// path behind this boundary: without an sm_100 device the first product or solve throws.
inline void ensure_init()
{
    static const bool once = [] {
        TraceCL t("init (CUDA context, streams, workspace)");
        const char* d = std::getenv("DROPS_B200_DEVICE");
        if (!d) d = std::getenv("LOCAL_RANK");
        check(drops_b200_init(d ? std::atoi(d) : 0));
        // DROPS_B200_OPTIONS="key=value,key=value": library options (include/drops_b200.h, drops_b200_set_option) for an
        // executable that has no other way to pass them
        if (const char* o = std::getenv("DROPS_B200_OPTIONS")) {
            std::string all(o);
            size_t pos = 0;
            while (pos < all.size()) {
                size_t end = all.find(',', pos);
                if (end == std::string::npos) end = all.size();
                const std::string kv = all.substr(pos, end - pos);
                const size_t eq = kv.find('=');
                if (eq != std::string::npos) check(drops_b200_set_option(kv.substr(0, eq).c_str(), std::atof(kv.c_str() + eq + 1)));
                pos = end + 1;
            }
        }
        return true;
    }();
    (void)once;
}

/// Device copies of host matrices.
class DeviceCacheCL
{
  private:
    struct EntryT {
        drops_b200_matrix_t handle = nullptr;
        size_t version = 0, rows = 0, cols = 0, nnz = 0;
        const void *val = nullptr, *col = nullptr, *row = nullptr;
        std::uint64_t pattern_sum = 0, value_sum = 0;
        std::uint64_t last_use = 0;
    };
    std::unordered_map<const void*, EntryT> table_;
    std::uint64_t clock_ = 0;
    size_t max_entries_;
    bool full_hash_;

    static std::uint64_t mix(std::uint64_t h, std::uint64_t v)
    {
        h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        return h;
    }
    template <typename U>
    std::uint64_t sample(const U* p, size_t n) const
    {
        static_assert(sizeof(U) == 8, "64-bit entries");
        std::uint64_t h = 1469598103934665603ull, w;
        const size_t step = (full_hash_ || n <= 4096) ? 1 : n / 4096;
        for (size_t i = 0; i < n; i += step) { std::memcpy(&w, p + i, 8); h = mix(h, w); }
        if (n) { std::memcpy(&w, p + (n - 1), 8); h = mix(h, w); }
        return h;
    }
    void evict_lru(const void* keep)
    {
        while (table_.size() > max_entries_) {
            auto victim = table_.end();
            for (auto it = table_.begin(); it != table_.end(); ++it)
                if (it->first != keep && (victim == table_.end() || it->second.last_use < victim->second.last_use)) victim = it;
            if (victim == table_.end()) return;
            drops_b200_matrix_destroy(victim->second.handle);
            table_.erase(victim);
        }
    }

  public:
    DeviceCacheCL()
    {
        const char* m = std::getenv("DROPS_B200_CACHE_ENTRIES");
        max_entries_ = m ? std::max(1, std::atoi(m)) : 32;
        const char* f = std::getenv("DROPS_B200_FULL_HASH");
        full_hash_ = f && std::atoi(f) != 0;
    }
    ~DeviceCacheCL() { clear(); }
    DeviceCacheCL(const DeviceCacheCL&) = delete;
    DeviceCacheCL& operator=(const DeviceCacheCL&) = delete;

    static DeviceCacheCL& instance()
    {
        static DeviceCacheCL cache;      // one per process (inline function: shared by all translation units)
        return cache;
    }
    size_t size() const { return table_.size(); }
    void clear()
    {
        for (auto& kv : table_) drops_b200_matrix_destroy(kv.second.handle);
        table_.clear();
    }
    /// Drop the copy of A (e.g. before the host object goes away, to give the HBM back early).
    void forget(const SparseMatBaseCL<double>& A)
    {
        auto it = table_.find(&A);
        if (it == table_.end()) return;
        drops_b200_matrix_destroy(it->second.handle);
        table_.erase(it);
    }
    /// Handle of the device copy of A: uploaded on first use, values refreshed when only they changed
    /// (drops_b200_matrix_update_values), rebuilt when the pattern changed.
    drops_b200_matrix_t get(const SparseMatBaseCL<double>& A)
    {
        ensure_init();
        EntryT& e = table_[&A];
        e.last_use = ++clock_;
        const size_t nnz = A.num_nonzeros();
        const std::uint64_t psum = mix(sample(A.raw_row(), A.num_rows() + 1), sample(A.raw_col(), nnz));
        const std::uint64_t vsum = sample(A.raw_val(), nnz);
        const bool same_pattern = e.handle && e.rows == A.num_rows() && e.cols == A.num_cols() && e.nnz == nnz &&
                                  e.pattern_sum == psum;
        if (same_pattern && e.version == A.Version() && e.val == A.raw_val() && e.col == A.raw_col() &&
            e.row == A.raw_row() && e.value_sum == vsum)
            return e.handle;
        TraceCL t(same_pattern ? "matrix values refreshed on the device" : "matrix uploaded (CSR -> device operator)");
        if (same_pattern) {
            check(drops_b200_matrix_update_values(e.handle, A.raw_val(), A.Version()));
        } else {
            if (e.handle) { drops_b200_matrix_destroy(e.handle); e.handle = nullptr; }
            int s = drops_b200_matrix_create(A.num_rows(), A.num_cols(), nnz, A.raw_row(), A.raw_col(), A.raw_val(),
                                             A.Version(), &e.handle);
            if (s == DROPS_B200_E_NOMEM && table_.size() > 1) {        // give the other copies back and retry once
                const size_t keep = max_entries_;
                max_entries_ = 1; evict_lru(&A); max_entries_ = keep;
                s = drops_b200_matrix_create(A.num_rows(), A.num_cols(), nnz, A.raw_row(), A.raw_col(), A.raw_val(),
                                             A.Version(), &e.handle);
            }
            if (s < 0) { table_.erase(&A); check(s); }
        }
        if (trace_on()) {                    // how the operator is stored: one stderr line per upload
            int bo = 0, ro = 0, nd = 0, pk = 0, cb = 0;
            size_t ls = 0, ss = 0, se = 0;
            drops_b200_matrix_layout(e.handle, &bo, &ro);
            drops_b200_matrix_line_segments(e.handle, &ls);
            drops_b200_matrix_segment_stream(e.handle, &ss, &se);
            drops_b200_matrix_value_dictionary(e.handle, 0, &nd, &pk, &cb);
            std::fprintf(stderr, "drops_b200 layout: rows %zu nnz %zu dictionary %d packed %d block_ordered %d row_ordered %d "
                                 "line_segments %zu segment_slices %zu segment_entries %zu\n",
                         (size_t)A.num_rows(), nnz, nd, pk, bo, ro, ls, ss, se);
        }
        e.version = A.Version(); e.rows = A.num_rows(); e.cols = A.num_cols(); e.nnz = nnz;
        e.val = A.raw_val(); e.col = A.raw_col(); e.row = A.raw_row();
        e.pattern_sum = psum; e.value_sum = vsum;
        evict_lru(&A);
        return e.handle;
    }
};

inline drops_b200_matrix_t device(const SparseMatBaseCL<double>& A) { return DeviceCacheCL::instance().get(A); }

}  // namespace b200

/// y = A*x on the GPU (per-row left-to-right sums of rounded products: bit-identical to y_Ax built without
/// FMA contraction).  x is uploaded, y downloaded inside the call.
inline VectorBaseCL<double> operator*(const SparseMatBaseCL<double>& A, const VectorBaseCL<double>& x)
{
    if (A.num_cols() != x.size()) throw DROPSErrCL("SparseMatBaseCL * VectorBaseCL: incompatible dimensions");
    VectorBaseCL<double> ret(A.num_rows());
    if (A.num_rows() == 0) return ret;
    b200::check(drops_b200_spmv(b200::device(A), Addr(x), &ret[0]));
    return ret;
}

/// y = A^T*x on the GPU (each y_c accumulated in ascending row order like the reference's scatter loop:
/// bit-identical).
inline VectorBaseCL<double> transp_mul(const SparseMatBaseCL<double>& A, const VectorBaseCL<double>& x)
{
    if (A.num_rows() != x.size()) throw DROPSErrCL("transp_mul: incompatible dimensions");
    VectorBaseCL<double> ret(0., A.num_cols());
    if (A.num_cols() == 0 || A.num_rows() == 0) return ret;
    b200::check(drops_b200_spmv_transpose(b200::device(A), Addr(x), &ret[0]));
    return ret;
}

}  // namespace DROPS

#endif

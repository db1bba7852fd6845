/* drops_b200.h -- C ABI of the B200-native DROPS sparse-solver hot path.
 *
 * One process drives one GPU.  Plain pointers and sizes only; no C++ / torch types cross this
 * boundary and no exception ever does: every entry point returns an int status
 * (>= 0 success, < 0 one of DROPS_B200_E_*; drops_b200_last_error() has the text).
 * There is no CPU fallback: without a usable sm_100 device every compute entry point fails with
 * DROPS_B200_E_NODEVICE.
 *
 * The DROPS reference (56th/drops) has no FFI; its boundary for this path is the header-template
 * interface of src/num/spmat.h, src/num/krylovsolver.h, src/num/precond.h, src/num/solverbase.h
 * and src/parallel/exchange.h.  drops_b200/hostxx/ re-creates those headers on top of this ABI
 * (see INTEGRATION.md); each entry point below names the reference interface it replaces
 * (paths relative to the reference's src/).
 *
 * Layout conventions (same as the reference):
 *   matrices : CSR, rowbeg[rows+1], colind[nnz] ascending inside a row, val[nnz]; the host entry
 *              points take the reference's own types (size_t indices, double values;
 *              SparseMatBaseCL::raw_row/raw_col/raw_val, num/spmat.h:789-794);
 *   vectors  : contiguous double (VectorBaseCL::raw(), num/spmat.h:131-132).
 */
#ifndef DROPS_B200_H
#define DROPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define DROPS_B200_API __attribute__((visibility("default")))
#else
#  define DROPS_B200_API
#endif

/* ---- status codes ------------------------------------------------------------------------ */
#define DROPS_B200_OK            0
#define DROPS_B200_E_ARG        -1   /* bad argument (null handle, size mismatch, ...)           */
#define DROPS_B200_E_CUDA       -2   /* CUDA runtime error (text in drops_b200_last_error)      */
#define DROPS_B200_E_NODEVICE   -3   /* no CUDA device / not sm_100: the product has no CPU path*/
#define DROPS_B200_E_NOTINIT    -4   /* drops_b200_init has not been called                     */
#define DROPS_B200_E_NODIAG     -5   /* Jacobi needs a stored diagonal entry in every row       */
#define DROPS_B200_E_RANGE      -6   /* index does not fit the device index type                */
#define DROPS_B200_E_COMM       -7   /* NCCL error                                              */
#define DROPS_B200_E_NOMEM      -8   /* device or host allocation failed                        */

/* preconditioner selector: the PC template argument of the reference solvers */
#define DROPS_B200_PC_DUMMY      0   /* DummyPcCL, num/precond.h:1032-1065: x = b               */
#define DROPS_B200_PC_JACOBI     1   /* JACPcCL,   num/precond.h:874 / :1155-1188: x = omega*b/a_ii */
#define DROPS_B200_PC_CHEBYSHEV  2   /* ChebyshevPcCL num/precond.h:1463-1529 (PCG); ChebyshevBBTPcCL :1531-1567
                                      * (PCGNE).  Parameters: drops_b200_matrix_set_chebyshev.              */
#define DROPS_B200_PC_DIAG       3   /* DiagPcCL, num/precond.h:618-642: x = D o b with a caller-supplied vector D
                                      * (drops_b200_matrix_set_diag_pc); PCG only.                           */
/* BlockMatrixOperationT of num/spblockmat.h:40 */
#define DROPS_B200_OP_MUL        0
#define DROPS_B200_OP_TRANSP_MUL 1

typedef struct drops_b200_matrix_s* drops_b200_matrix_t;
typedef struct drops_b200_vector_s* drops_b200_vector_t;

/* ---- process-wide context ---------------------------------------------------------------- */
/* Bind this process to CUDA device `device` (cudaSetDevice), create the solver stream and the
 * reduction workspace.  Idempotent for the same device. */
DROPS_B200_API int drops_b200_init(int device);
DROPS_B200_API int drops_b200_shutdown(void);
DROPS_B200_API const char* drops_b200_last_error(void);
/* Library/ABI version and the sm architecture the kernels were compiled for (100). */
DROPS_B200_API int drops_b200_abi_version(void);
DROPS_B200_API int drops_b200_compiled_arch(void);
/* Device facts: sm count, compute capability major*10+minor, free/total bytes. */
DROPS_B200_API int drops_b200_device_info(int* sm_count, int* cc, size_t* free_bytes, size_t* total_bytes);
/* Tunables: "spmv.kernel", "spmv.threads", "spmv.slices_per_cta", "spmv.exact",
 * "spmv.dotmode", "sell.sigma", "vec.wide", "pcg.check_every", "gmres.ortho", "timing", ... */
DROPS_B200_API int drops_b200_set_option(const char* key, double value);
DROPS_B200_API int drops_b200_get_option(const char* key, double* value);
/* Accumulated CUDA-event time (ms) and launch count of a named kernel class since the last
 * reset; only collected while option "timing" is 1.  Keys: "spmv", "pcg.update", "pcg.direction",
 * "vec", ... */
DROPS_B200_API int drops_b200_timing_get(const char* key, double* ms, long long* launches);
DROPS_B200_API int drops_b200_timing_reset(void);
/* Device stopwatch: CUDA events recorded on the solver stream.  timer_stop synchronises and
 * returns the milliseconds since the matching timer_start (what bench.py reports). */
DROPS_B200_API int drops_b200_timer_start(void);
DROPS_B200_API int drops_b200_timer_stop(double* ms);
/* Number of kernels this library launched since init (all streams). */
DROPS_B200_API long long drops_b200_launch_count(void);
DROPS_B200_API int drops_b200_synchronize(void);

/* ---- matrices: SparseMatBaseCL<double>, num/spmat.h:682-879 -------------------------------- */
/* Copies a host CSR matrix in the reference layout to the device: narrows colind to int32,
 * extracts the diagonal (GetDiag, num/spmat.h:993-1002), builds the sliced-ELL operator.
 * `version` is SparseMatBaseCL::Version() (num/spmat.h:807-808); the hostxx layer keys its
 * handle cache on (address, version). */
DROPS_B200_API int drops_b200_matrix_create(size_t rows, size_t cols, size_t nnz,
                                            const size_t* rowbeg, const size_t* colind,
                                            const double* val, uint64_t version,
                                            drops_b200_matrix_t* out);
/* Same, from CSR arrays already resident on this device (int64 row pointers, int32 columns).
 * The arrays are consumed (copied or adopted) before the call returns. */
DROPS_B200_API int drops_b200_matrix_create_device(size_t rows, size_t cols, size_t nnz,
                                                   const int64_t* d_rowptr, const int32_t* d_colind,
                                                   const double* d_val, uint64_t version,
                                                   drops_b200_matrix_t* out);
/* New values on the unchanged pattern (operator*=, LinComb results with equal pattern). */
DROPS_B200_API int drops_b200_matrix_update_values(drops_b200_matrix_t A, const double* val, uint64_t version);
/* C = a*A + b*B on the device: SparseMatBaseCL::LinComb, num/spmat.h:1457-1551 (row-wise merge of the
 * two ascending column lists; an entry present in both is a*A_ij + b*B_ij with both products rounded
 * first).  DROPS calls it every time step right before the solve (levelset/coupling.tpp:96,
 * poisson/cdrdrops.cpp:112); with the operands already resident nothing crosses PCIe. */
DROPS_B200_API int drops_b200_matrix_lincomb(double a, drops_b200_matrix_t A, double b, drops_b200_matrix_t B,
                                             uint64_t version, drops_b200_matrix_t* out);
DROPS_B200_API int drops_b200_matrix_destroy(drops_b200_matrix_t A);
DROPS_B200_API int drops_b200_matrix_info(drops_b200_matrix_t A, size_t* rows, size_t* cols, size_t* nnz,
                                          uint64_t* version, size_t* device_bytes, size_t* padded_nnz);
/* Storage facts of the sliced-ELL operator: slices of 32 rows, stored entries incl. padding. */
DROPS_B200_API int drops_b200_matrix_format(drops_b200_matrix_t A, size_t* nslices, size_t* padded_nnz);
/* Value dictionary of the sliced-ELL operator ("sell.valdict", default on): a matrix with at most
 * 65535 distinct fp64 bit patterns whose column and dictionary-index bits fit 32 is ALSO stored as one
 * 32-bit word per entry (lossless; products stay bit-identical).  n_distinct = 0 and packed = 0 when
 * the matrix does not qualify; column_bits = low bits of a word that hold the column.  packed = 2: the
 * dictionary is used as a 16-bit index beside the int32 column (6 bytes per entry) because column bits +
 * index bits exceed one word ("sell.pack_bits", 32).  `transposed`
 * != 0 asks about the A^T operator (0/0 until the first transp_mul built it). */
DROPS_B200_API int drops_b200_matrix_value_dictionary(drops_b200_matrix_t A, int transposed, int* n_distinct,
                                                      int* packed, int* column_bits);
/* Row order of the sliced-ELL operator: block_ordered != 0 when the slices were cut from the block
 * order of the matrix graph ("sell.block_order", operators of at least "sell.block_min_rows" rows:
 * three breadth-first distance fields as pseudo-coordinates, rows sorted by the Morton code of their
 * cell); row_ordered_stream != 0 when the operator has its row-ordered twin, which is what PCG iterates on
 * ("pcg.reorder"): 1 = twin of the packed stream, 2 = twin of the int32 columns of the 12-byte stream (operators
 * without a value dictionary).  Both only change who shares a cache, never a result bit of A*x. */
DROPS_B200_API int drops_b200_matrix_layout(drops_b200_matrix_t A, int* block_ordered, int* row_ordered_stream);
/* "sell.line_order" (default on, part of the block order): rows that repeat each other translated along the fastest
 * direction of the caller's numbering (same length, every column moved up by 1..."sell.line_maxdiff", at most
 * "sell.line_stride" rows apart) are chained and cut into segments of 32; a complete segment is one slice, so entry j of
 * its lanes gathers 32 consecutive positions of the row-ordered x.  *n_segments = complete segments of A's operator
 * (0: no lines found, or A is not block ordered).  Never changes a result bit of A*x. */
DROPS_B200_API int drops_b200_matrix_line_segments(drops_b200_matrix_t A, size_t* n_segments);
/* The row order of the sliced-ELL operator: perm[32 * s + l] = row that lane l of slice s sums (-1: padding position);
 * capacity >= 32 * nslices (drops_b200_matrix_format).  Position p of this order is entry p of the row-ordered vectors
 * the solvers iterate on.  Introspection for tests and tools. */
DROPS_B200_API int drops_b200_matrix_row_order(drops_b200_matrix_t A, int32_t* perm, size_t capacity);
/* "sell.segments" (default on): slices of the row-ordered twin whose 32 rows are one line segment hold, entry by entry,
 * ONE dictionary index and columns that are at most two runs of consecutive positions; they are described by one 16-byte
 * record per entry-row instead of 32 packed words (0.5 instead of 4 bytes per stored entry) and run by the segment kernel
 * ("spmv.segments").  *n_slices / *n_entries = such slices and the stored entries they cover (0: none, or too few to
 * pay).  Same products in the same order: never changes a result bit. */
DROPS_B200_API int drops_b200_matrix_segment_stream(drops_b200_matrix_t A, size_t* n_slices, size_t* n_entries);
/* Jacobi diagonal of the row-ordered PCG as one byte per row ("pcg.diag_index"; JACPcCL's diag_, num/precond.h:1162-1177,
 * holds a handful of distinct values on constant-coefficient operators): *n_values = length of the value list the
 * last PCG-Jacobi solve on A built, 0 when the diagonal is read as an fp64 vector.  Never changes a result bit. */
DROPS_B200_API int drops_b200_matrix_diag_index(drops_b200_matrix_t A, int* n_values);
/* Copies the device CSR back in the reference layout (for tests and the CPU baseline). */
DROPS_B200_API int drops_b200_matrix_download(drops_b200_matrix_t A, size_t* rowbeg, size_t* colind, double* val);

/* y = A*x      : operator*(SparseMatBaseCL, VectorBaseCL), num/spmat.h:1712-1724 (y_Ax :1682-1709).
 * y = A^T*x    : transp_mul, num/spmat.h:1748-1760 (y_ATx :1729-1746).
 * Host buffers; the copies are part of the call.  x has cols (rows) entries, y rows (cols). */
DROPS_B200_API int drops_b200_spmv(drops_b200_matrix_t A, const double* x, double* y);
DROPS_B200_API int drops_b200_spmv_transpose(drops_b200_matrix_t A, const double* x, double* y);
/* A.GetDiag(), num/spmat.h:993-1002. */
DROPS_B200_API int drops_b200_matrix_diag(drops_b200_matrix_t A, double* diag);

/* ---- device-resident vectors: VectorBaseCL<double>, num/spmat.h:82-145 ---------------------- */
DROPS_B200_API int drops_b200_vector_create(size_t n, drops_b200_vector_t* out);
DROPS_B200_API int drops_b200_vector_destroy(drops_b200_vector_t v);
DROPS_B200_API int drops_b200_vector_size(drops_b200_vector_t v, size_t* n);
DROPS_B200_API int drops_b200_vector_upload(drops_b200_vector_t v, const double* host);
DROPS_B200_API int drops_b200_vector_download(drops_b200_vector_t v, double* host);
DROPS_B200_API int drops_b200_vector_fill(drops_b200_vector_t v, double value);
DROPS_B200_API int drops_b200_vector_copy(drops_b200_vector_t dst, drops_b200_vector_t src);
/* x_i = ((i*2654435761) mod 2^32)/2^32 - 0.5: the seedless SpMV input of SURVEY.md section 8(d). */
DROPS_B200_API int drops_b200_vector_fill_hash(drops_b200_vector_t v);
/* Raw device pointer (for callers that own CUDA code themselves). */
DROPS_B200_API int drops_b200_vector_ptr(drops_b200_vector_t v, double** dptr);

DROPS_B200_API int drops_b200_spmv_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t y);
DROPS_B200_API int drops_b200_spmv_transpose_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t y);
/* y = A*x and *result = dot(y, w) in one kernel: the q = A*p; lambda = ParDot(q, p) pair of PCG
 * (num/krylovsolver.h:149-151) as the solver runs it.  w may be x. */
DROPS_B200_API int drops_b200_spmv_dot_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t y,
                                           drops_b200_vector_t w, double* result);

/* dot (num/spmat.h:164-170), norm_sq/norm (:180-196), supnorm (:198-210): one deterministic
 * double-double tree reduction; DummyExchangeCL::ParDot/Norm (parallel/exchange.h:58-84, Kahan)
 * map to the same kernels (error <= the reference's). */
DROPS_B200_API int drops_b200_dot(drops_b200_vector_t x, drops_b200_vector_t y, double* result);
DROPS_B200_API int drops_b200_norm_sq(drops_b200_vector_t x, double* result);
DROPS_B200_API int drops_b200_norm(drops_b200_vector_t x, double* result);
DROPS_B200_API int drops_b200_supnorm(drops_b200_vector_t x, double* result);
/* axpy: y += a*x (num/spmat.h:212-218); z_xpay: z = x + a*y (:220-227);
 * z_xpaypby2: z = x + a*y + b*y2 (:229-237).  Products are rounded before the add (no FMA),
 * like the reference's valarray expressions. */
DROPS_B200_API int drops_b200_axpy(double a, drops_b200_vector_t x, drops_b200_vector_t y);
DROPS_B200_API int drops_b200_z_xpay(drops_b200_vector_t z, drops_b200_vector_t x, double a, drops_b200_vector_t y);
DROPS_B200_API int drops_b200_z_xpaypby2(drops_b200_vector_t z, drops_b200_vector_t x, double a,
                                         drops_b200_vector_t y, double b, drops_b200_vector_t y2);
/* JACPcCL::Apply: x_i = omega*b_i/a_ii (num/precond.h:153-169, :1293-1311). */
DROPS_B200_API int drops_b200_jacobi_apply(drops_b200_matrix_t A, double omega,
                                           drops_b200_vector_t b, drops_b200_vector_t x);

/* ---- Krylov solvers ------------------------------------------------------------------------
 * Common contract (num/krylovsolver.h:50-59): x is in/out (initial guess -> solution), b is
 * const; *max_iter IN: iteration limit, OUT: iterations used (unchanged when not converged);
 * *tol IN: tolerance, OUT: final residual (relative iff `relative`).  Return value: 1 converged,
 * 0 not converged within *max_iter, < 0 error.  hist (may be NULL) receives the residual the
 * reference would print through SolverBaseCL::output_ after each iteration, hist[0] = initial.
 *
 * pcg    : PCG, num/krylovsolver.h:107-179 (PCGSolverCL<PC>::Solve :1818-1824) with
 *          DummyExchangeCL; pc = DROPS_B200_PC_JACOBI gives PCGSolverCL<JACPcCL>.
 * cg     : CG, :49-103 (CGSolverCL :1782-1802).
 * gmres  : serial GMRES(m), :320-461 (GMResSolverCL<PC>::Solve :1946-1958), modified
 *          Gram-Schmidt; left_precond = 1 is the driver's default (:1937).
 * minres : MINRES :1155-1161 / PMINRES :1069-1152 (MResSolverCL :1868-1889); precond = 1 runs
 *          the preconditioned Lanczos basis (PLanczosONBCL :1010-1052) with `pc`.
 * Host-buffer forms copy x,b in and x out inside the call; *_dev forms work on device vectors. */
DROPS_B200_API int drops_b200_pcg(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega,
                                  int* max_iter, double* tol, int relative, double* hist, int hist_cap);
DROPS_B200_API int drops_b200_pcg_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                      int pc, double omega, int* max_iter, double* tol, int relative,
                                      double* hist, int hist_cap);
DROPS_B200_API int drops_b200_cg(drops_b200_matrix_t A, double* x, const double* b,
                                 int* max_iter, double* tol, int relative, double* hist, int hist_cap);
DROPS_B200_API int drops_b200_cg_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                     int* max_iter, double* tol, int relative, double* hist, int hist_cap);
DROPS_B200_API int drops_b200_gmres(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega,
                                    int restart, int* max_iter, double* tol, int relative, int left_precond);
DROPS_B200_API int drops_b200_gmres_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                        int pc, double omega, int restart, int* max_iter, double* tol,
                                        int relative, int left_precond);
DROPS_B200_API int drops_b200_minres(drops_b200_matrix_t A, double* x, const double* b, int precond, int pc,
                                     double omega, int* max_iter, double* tol, int relative);
DROPS_B200_API int drops_b200_minres_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                         int precond, int pc, double omega, int* max_iter, double* tol,
                                         int relative);

/* bicgstab : BICGSTAB, num/krylovsolver.h:1189-1258 (BiCGStabSolverCL<PC>::Solve :1978-2013, serial
 *            branch).  Returns 0 for "not converged" and for a breakdown, as the reference does;
 *            *breakdown (may be NULL) is then 1 (rho_1 == 0) or 2 (omega == 0).  Note the
 *            reference's strict "< tol" inside the loop (:1226, :1240).
 * gcr      : GCR(m), num/krylovsolver.h:1410-1483 (GCRSolverCL<PC>::Solve :2017-2053), min-alpha
 *            truncation with `truncate` stored directions, including the reference's behaviour of not
 *            refreshing vacc[min_idx] after a truncation. */
DROPS_B200_API int drops_b200_bicgstab(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega,
                                       int* max_iter, double* tol, int relative, int* breakdown);
DROPS_B200_API int drops_b200_bicgstab_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                           int pc, double omega, int* max_iter, double* tol, int relative,
                                           int* breakdown);
DROPS_B200_API int drops_b200_gcr(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega,
                                  int truncate, int* max_iter, double* tol, int relative);
DROPS_B200_API int drops_b200_gcr_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                      int pc, double omega, int truncate, int* max_iter, double* tol, int relative);

/* ---- the callers next to the core path --------------------------------
 * Chebyshev polynomial preconditioner: ChebyshevPcCL(lambda_scale = 1.1, degree = 3, ratio = 30) for
 * a square A (num/precond.h:1474) and ChebyshevBBTPcCL(lambda_scale = 1, degree = 3, ratio = 30) for
 * the normal equations A A^T (:1533); normal_eq selects which parameter set is changed.  The
 * defaults are the reference's constructor defaults. */
DROPS_B200_API int drops_b200_matrix_set_chebyshev(drops_b200_matrix_t A, int normal_eq, double lambda_scale,
                                                   int degree, double eig_ratio);
/* What SetDiag computes (num/precond.h:1481-1491, :1537-1553): lambda_scale * MaxEigenvalueScaledMatrix
 * (:1080-1096, 10 power iterations on D^-1 A or D^-1 A A^T with D = BBTDiag).  Cached on the handle. */
DROPS_B200_API int drops_b200_chebyshev_lambda_max(drops_b200_matrix_t A, int normal_eq, double* lambda_max);
/* DiagPcCL(D), num/precond.h:618-642 (the Schur-complement preconditioner of num/oseenprecond.h:911-919): the vector
 * the preconditioner multiplies with, rows entries; copied to the device and kept with the handle.  NULL forgets it.
 * DiagPcCL::Apply itself is drops_b200_diagpc_apply: x_i = D_i * b_i (product rounded once, as the valarray expression). */
DROPS_B200_API int drops_b200_matrix_set_diag_pc(drops_b200_matrix_t A, const double* D);
DROPS_B200_API int drops_b200_diagpc_apply(drops_b200_vector_t D, drops_b200_vector_t b, drops_b200_vector_t x);
/* BBTDiag, num/precond.h:56-75: diag(A A^T). */
DROPS_B200_API int drops_b200_matrix_bbt_diag(drops_b200_matrix_t A, double* diag);
/* x = M^-1 b: PC::Apply(A, x, b, ex) (normal_eq = 0) or PC::Apply(A, x, b, rowex, colex) on A A^T
 * (normal_eq = 1; Jacobi then divides by BBTDiag).  x must not alias b for Chebyshev. */
DROPS_B200_API int drops_b200_precond_apply_dev(drops_b200_matrix_t A, int pc, double omega, int normal_eq,
                                                drops_b200_vector_t b, drops_b200_vector_t x);
/* PCG through the generic-preconditioner driver (what drops_b200_pcg uses for PC_CHEBYSHEV; with
 * PC_DUMMY / PC_JACOBI it gives the results of drops_b200_pcg_dev without CUDA graphs). */
DROPS_B200_API int drops_b200_pcg_general_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b,
                                              int pc, double omega, int* max_iter, double* tol, int relative,
                                              double* hist, int hist_cap);
/* pcgne : PCGNE, num/krylovsolver.h:199-252 (PCGNESolverCL<PC>::Solve :1851-1864): CG on A A^T u = b
 *         for a rectangular A (rows x cols), u and b have rows entries.  Contract as drops_b200_pcg. */
DROPS_B200_API int drops_b200_pcgne(drops_b200_matrix_t A, double* u, const double* b, int pc, double omega,
                                    int* max_iter, double* tol, int relative, double* hist, int hist_cap);
DROPS_B200_API int drops_b200_pcgne_dev(drops_b200_matrix_t A, drops_b200_vector_t u, drops_b200_vector_t b,
                                        int pc, double omega, int* max_iter, double* tol, int relative,
                                        double* hist, int hist_cap);
/* idrs  : IDRS, num/krylovsolver.h:1579-1698 (IDRsSolverCL<PC>::Solve :2112-2126).  shadow: the s
 *         vectors P BEFORE orthonormalisation (host, s*n doubles, vector i at shadow + i*n), or NULL to
 *         draw them as the reference does (srand(time(0)), 2*rand()/RAND_MAX - 1).  Outputs follow the
 *         reference literally: *tol OUT is the ABSOLUTE residual norm, *max_iter OUT the iteration
 *         count in every case, and the return value is 0 when tol_in > |r|/normb at exit (:1689-1693,
 *         i.e. on convergence inside the loop) and 1 otherwise.  A zero pivot / omega (DROPSErrCL in
 *         the reference) returns DROPS_B200_E_ARG. */
DROPS_B200_API int drops_b200_idrs(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega, int s,
                                   double omega_bound, const double* shadow, int* max_iter, double* tol,
                                   int relative);
DROPS_B200_API int drops_b200_idrs_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc,
                                       double omega, int s, double omega_bound, const double* shadow,
                                       int* max_iter, double* tol, int relative);
/* BlockMatrixBaseCL operator* (num/spblockmat.h:225-267): blocks[4] = {A, B, C, D} of ( A B ; C D ),
 * any of them may be NULL, ops[i] = DROPS_B200_OP_MUL / _TRANSP_MUL says how block i is applied.
 * r0 = op(A) x0 [+ op(B) x1], r1 = op(C) x0 [+ op(D) x1], each "+=" a rounded elementwise add as in
 * the reference.  transp_mul(BlockMatrix, x) (:269-273) is the same call on {A, C, B, D} with the
 * operations flipped.  block_dims: num_rows() / num_cols() (:112-190). */
DROPS_B200_API int drops_b200_block_dims(const drops_b200_matrix_t blocks[4], const int ops[4], size_t* rows,
                                         size_t* cols);
DROPS_B200_API int drops_b200_block_spmv(const drops_b200_matrix_t blocks[4], const int ops[4], const double* x,
                                         double* y);
DROPS_B200_API int drops_b200_block_spmv_dev(const drops_b200_matrix_t blocks[4], const int ops[4],
                                             drops_b200_vector_t x, drops_b200_vector_t y);
/* CompositeMatrixBaseCL operator* (num/spblockmat.h:376-397): y = op1(A1) * (op0(A0) * x). */
DROPS_B200_API int drops_b200_composite_spmv(drops_b200_matrix_t A0, int op0, drops_b200_matrix_t A1, int op1,
                                             const double* x, double* y);
DROPS_B200_API int drops_b200_composite_spmv_dev(drops_b200_matrix_t A0, int op0, drops_b200_matrix_t A1, int op1,
                                                 drops_b200_vector_t x, drops_b200_vector_t y);

/* ---- synthetic brick systems ---------------------------------------
 * Uniform Kuhn brick (BrickBuilderCL, geom/builder.cpp:99-180: 6 tetrahedra per cell sharing the
 * main diagonal) on the unit cube with N cells per direction, homogeneous Dirichlet on all faces,
 * alpha = 1.  fe = 1: P1 stiffness (poisson/poissonaccus.tpp:327), n = (N-1)^3, 15 stored
 * couplings per interior row.  fe = 2: P2 stiffness (poisson/poisson.tpp:828-838, :898-917),
 * n = (2N-1)^3, 65/27/27/19 stored couplings per interior row.  All couplings of DOFs sharing a
 * tetrahedron are stored, numerically zero or not (poisson.tpp:912-917).
 * numbering = 0: lexicographic over the DOF lattice; 1: vertices first, then edges
 * (misc/problem.cpp:707-719).  upper_scale != 1 multiplies the strict upper triangle (the
 * nonsymmetric GMRES stand-in).  z-slab [z0, z1) of lattice planes restricts the generated rows
 * (multi-GPU); z0 = z1 = 0 means all rows. */
DROPS_B200_API int drops_b200_brick_sizes(int fe, int N, size_t* n, size_t* nnz);
DROPS_B200_API int drops_b200_brick_csr_host(int fe, int N, int numbering, double upper_scale,
                                             size_t* rowbeg, size_t* colind, double* val);
DROPS_B200_API int drops_b200_brick_matrix_device(int fe, int N, int numbering, double upper_scale,
                                                  drops_b200_matrix_t* out);
/* x*(p) = 64 xyz(1-x)(1-y)(1-z) sampled at the DOF positions (poissonsolverfactory.h:25-26). */
DROPS_B200_API int drops_b200_brick_solution_host(int fe, int N, int numbering, double* xstar);
DROPS_B200_API int drops_b200_brick_solution_device(int fe, int N, int numbering, drops_b200_vector_t xstar);

/* ---- row-partitioned systems across the GPUs of one box --------------------------------------
 * Stand-in for the reference's distributed layer: ExchangeCL (parallel/exchange.h:194-243,
 * Accumulate parallel/exchange.cpp:333-362, ParDot :449-462, Norm :487-495) over ProcCL/MPI
 * (parallel/parallel.h:256-263), with ParMETIS deciding who owns what
 * (parallel/partitioner.cpp:251-257).  DROPS partitions tetrahedra and duplicates interface DOFs;
 * here the matrix is partitioned by ROWS (owner computes), so a vector has one form only, the
 * product y = A x needs the neighbours' x entries of the referenced-but-not-owned columns
 * ("ghosts": ncclSend/ncclRecv, overlapped with the rows that reference no ghost), and a dot
 * product is a local double-double partial + one ncclAllGather summed in rank order on every rank
 * (bit-identical everywhere, so all ranks take the same branches).  One process per GPU.
 *
 * Local numbering of a rank: [0, n_owned) owned DOFs by ascending global id, then the ghosts
 * grouped by owner rank.  A distributed vector is a drops_b200_vector_t of n_owned entries. */
typedef struct drops_b200_exchange_s*    drops_b200_exchange_t;
typedef struct drops_b200_comm_s*        drops_b200_comm_t;
typedef struct drops_b200_dist_matrix_s* drops_b200_dist_matrix_t;

/* -- halo plan (host only; no device needed).  Built in two steps because the send side needs the
 *    neighbours' requests: create -> request(q) for every q -> [caller transports the lists] ->
 *    set_send(q, what q requested from us).  ghost_gid must be unique and not owned. */
DROPS_B200_API int drops_b200_exchange_create(int rank, int world, size_t n_owned, const int64_t* owned_gid,
                                              size_t n_ghost, const int64_t* ghost_gid, const int32_t* ghost_owner,
                                              drops_b200_exchange_t* out);
DROPS_B200_API int drops_b200_exchange_destroy(drops_b200_exchange_t ex);
DROPS_B200_API int drops_b200_exchange_sizes(drops_b200_exchange_t ex, size_t* n_owned, size_t* n_ghost,
                                             size_t* n_send, int* n_neighbours);
/* Global ids this rank needs from `peer`, in the order they will be received (gids may be NULL). */
DROPS_B200_API int drops_b200_exchange_request(drops_b200_exchange_t ex, int peer, size_t* count, int64_t* gids);
/* The ids `peer` requested from this rank; all must be owned here. */
DROPS_B200_API int drops_b200_exchange_set_send(drops_b200_exchange_t ex, int peer, size_t count, const int64_t* gids);
DROPS_B200_API int drops_b200_exchange_send_list(drops_b200_exchange_t ex, int peer, size_t* count, int64_t* local_idx);
/* local index of each global id (owned: position; ghost: n_owned + slot; else -1) */
DROPS_B200_API int drops_b200_exchange_local_index(drops_b200_exchange_t ex, size_t count, const int64_t* gids,
                                                   int64_t* local);
/* 1 when every peer's send list has been set */
DROPS_B200_API int drops_b200_exchange_ready(drops_b200_exchange_t ex);
/* Sorted unique column ids of the local rows that are not in owned_gid.  Writes at most `cap` ids,
 * *n_ghost receives the full count (call with cap = 0 first).  colind: reference layout (size_t). */
DROPS_B200_API int drops_b200_find_ghosts(size_t n_owned, const int64_t* owned_gid, size_t nnz, const size_t* colind,
                                          int64_t* ghost_out, size_t cap, size_t* n_ghost);
/* Row partitioner on the matrix graph (breadth-first level structure cut into chunks of equal
 * stored-entry count): part[row] in [0, nparts).  Host only. */
DROPS_B200_API int drops_b200_partition_rows(size_t rows, const size_t* rowbeg, const size_t* colind, int nparts,
                                             int32_t* part);

/* -- communicator.  unique_id: rank 0 fills DROPS_B200_UNIQUE_ID_BYTES bytes (ncclGetUniqueId), the
 *    caller broadcasts them (torch.distributed / MPI / files), every rank calls comm_create.
 *    comm_create_local: all `world` partitions are held by THIS process on this GPU (device-to-
 *    device copies instead of NCCL) -- the transport of the single-GPU parity tests. */
#define DROPS_B200_UNIQUE_ID_BYTES 128
DROPS_B200_API int drops_b200_comm_unique_id(void* id_bytes);
DROPS_B200_API int drops_b200_comm_create(int rank, int world, const void* id_bytes, drops_b200_comm_t* out);
DROPS_B200_API int drops_b200_comm_create_local(int world, drops_b200_comm_t* out);
DROPS_B200_API int drops_b200_comm_destroy(drops_b200_comm_t comm);
DROPS_B200_API int drops_b200_comm_info(drops_b200_comm_t comm, int* rank, int* world, int* is_local);
/* Device-side barrier + max over ranks of a host double (timing plumbing for bench.py). */
DROPS_B200_API int drops_b200_comm_allreduce_max(drops_b200_comm_t comm, double* value);

/* -- the rows of one rank.  Rows in the order of the plan's owned ids, GLOBAL column ids, entries of
 *    a row in the reference's order (ascending global column): the per-row summation order, hence
 *    every y_i, is bit-identical to the single-GPU product.  The plan must be ready. */
DROPS_B200_API int drops_b200_dist_matrix_create(drops_b200_comm_t comm, drops_b200_exchange_t ex, size_t nnz,
                                                 const size_t* rowbeg, const size_t* colind_global, const double* val,
                                                 uint64_t version, drops_b200_dist_matrix_t* out);
DROPS_B200_API int drops_b200_dist_matrix_create_device(drops_b200_comm_t comm, drops_b200_exchange_t ex, size_t nnz,
                                                        const int64_t* d_rowptr, const int64_t* d_colind_global,
                                                        const double* d_val, uint64_t version,
                                                        drops_b200_dist_matrix_t* out);
DROPS_B200_API int drops_b200_dist_matrix_destroy(drops_b200_dist_matrix_t A);
DROPS_B200_API int drops_b200_dist_matrix_info(drops_b200_dist_matrix_t A, size_t* n_owned, size_t* n_ghost, size_t* nnz,
                                               size_t* interior_rows, size_t* boundary_rows, size_t* n_send,
                                               size_t* device_bytes);
/* Storage facts of a partition: interior_packed is a bit set -- 1: the interior-row operator has the
 * dictionary-packed 4-byte stream, 2: the six-byte form, 4: the partition has its row-ordered twins (PCG
 * iterates in the numbering of the operators' own slices, "dist.reorder"), 8: rows were cut in blocks of the
 * partition's graph, 16: the peer-to-peer halo is set up (the pack kernel stores into the neighbours' ghost
 * buffers over NVLink, "dist.p2p"; otherwise ncclSend/ncclRecv), 32: a peer-to-peer wait timed out, 64: the
 * reduction partials travel peer to peer as well ("dist.p2p_reduce"; otherwise ncclAllGather), 128: the update kernel of the
 * partitioned PCG reads the Jacobi diagonal as a byte index, 256: the interior rows run from segment descriptors
 * (drops_b200_matrix_segment_stream); n_distinct = size of the interior operator's dictionary. */
DROPS_B200_API int drops_b200_dist_matrix_format(drops_b200_dist_matrix_t A, int* interior_packed, int* n_distinct);
/* Segment descriptors of the interior-row operator (see drops_b200_matrix_segment_stream); padded_entries = stored
 * entries of that operator incl. padding. */
DROPS_B200_API int drops_b200_dist_matrix_segment_stream(drops_b200_dist_matrix_t A, size_t* n_slices, size_t* n_entries,
                                                         size_t* padded_entries);
/* Time this partition spent waiting for its peers inside the iteration (peer-to-peer path only; zeros otherwise):
 * out8 = 4 x {nanoseconds, number of waits} for the halo wait before the boundary rows, the p.Ap reduction, the
 * (rho, residual) reduction of PCG and all other reductions.  A wait is the time between the start of the waiting
 * kernel and the arrival of the last peer's sequence number; reset != 0 clears the counters.  "dist.wait_stats" (1). */
DROPS_B200_API int drops_b200_dist_matrix_wait_stats(drops_b200_dist_matrix_t A, double* out8, int reset);
DROPS_B200_API int drops_b200_dist_matrix_diag(drops_b200_dist_matrix_t A, double* diag);

/* Collective operations.  `nparts` partitions driven by this process: 1 with an NCCL communicator
 * (every rank makes the same call), `world` with a local communicator (arrays ordered by rank). */
DROPS_B200_API int drops_b200_dist_spmv_dev(int nparts, const drops_b200_dist_matrix_t* A,
                                            const drops_b200_vector_t* x, const drops_b200_vector_t* y);
DROPS_B200_API int drops_b200_dist_dot(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                                       const drops_b200_vector_t* y, double* result);
/* PCG (num/krylovsolver.h:107-179) on the partitioned system; contract as drops_b200_pcg_dev.
 * Every rank returns the same iteration count and residual. */
DROPS_B200_API int drops_b200_dist_pcg_dev(int nparts, const drops_b200_dist_matrix_t* A,
                                           const drops_b200_vector_t* x, const drops_b200_vector_t* b, int pc,
                                           double omega, int* max_iter, double* tol, int relative, double* hist,
                                           int hist_cap);
/* GMRES(m) on the partitioned system: the serial driver of num/krylovsolver.h:320-461 (modified Gram-Schmidt,
 * left or right preconditioning with JACPcCL / DummyPcCL), every dot product summed over the ranks in rank
 * order (one small all-gather per Gram-Schmidt step -- what the reference's distributed variants,
 * krylovsolver.h:528-916, do with GlobalSum; "gmres.ortho" = 1: classical Gram-Schmidt as ModGMRES :465-494, :862-866,
 * all i+1 products of a column in one pass and ONE reduction over the ranks).  All ranks compute identical Hessenberg columns and stop in the
 * same iteration.  Arguments and return value as drops_b200_gmres_dev. */
DROPS_B200_API int drops_b200_dist_gmres_dev(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                                             const drops_b200_vector_t* b, int pc, double omega, int restart,
                                             int* max_iter, double* tol, int relative, int left_precond);

/* MINRES / PMINRES on the partitioned system (num/krylovsolver.h:926-1161; MResSolverCL :1918-1941, PMResSolverCL
 * :1944-1972 -- the solvers the reference's parallel Stokes drivers instantiate with ExchangeCL): precond = 0 plain
 * Lanczos, 1 = preconditioned Lanczos with pc in {DUMMY, JACOBI}.  Two rank-ordered reductions per Lanczos step (the
 * reference: ParDot + Norm, exchange.h), rotations on the host of every rank from identical scalars.  Arguments and
 * return value as drops_b200_minres_dev. */
DROPS_B200_API int drops_b200_dist_minres_dev(int nparts, const drops_b200_dist_matrix_t* A, const drops_b200_vector_t* x,
                                              const drops_b200_vector_t* b, int precond, int pc, double omega,
                                              int* max_iter, double* tol, int relative);

/* -- synthetic brick, z-slab partition: rank r owns the lattice planes [k0, k1) (1-based, M planes
 *    split evenly), rows in ascending global id.  Generated on the device, global ids as above. */
typedef struct drops_b200_brick_slab_s* drops_b200_brick_slab_t;
DROPS_B200_API int drops_b200_brick_slab_generate(int fe, int N, int numbering, double upper_scale, int world, int rank,
                                                  drops_b200_brick_slab_t* out);
DROPS_B200_API int drops_b200_brick_slab_destroy(drops_b200_brick_slab_t slab);
DROPS_B200_API int drops_b200_brick_slab_info(drops_b200_brick_slab_t slab, size_t* n_owned, size_t* nnz, size_t* n_ghost,
                                              long long* k0, long long* k1);
DROPS_B200_API int drops_b200_brick_slab_owned(drops_b200_brick_slab_t slab, int64_t* owned_gid);
DROPS_B200_API int drops_b200_brick_slab_ghosts(drops_b200_brick_slab_t slab, int64_t* ghost_gid, int32_t* ghost_owner);
DROPS_B200_API int drops_b200_brick_slab_matrix(drops_b200_comm_t comm, drops_b200_exchange_t ex,
                                                drops_b200_brick_slab_t slab, drops_b200_dist_matrix_t* out);
DROPS_B200_API int drops_b200_brick_slab_solution(drops_b200_brick_slab_t slab, drops_b200_vector_t xstar);
/* Host helper: owner rank (z-slab) of global ids of the brick. */
DROPS_B200_API int drops_b200_brick_owner(int fe, int N, int numbering, int world, size_t count, const int64_t* gids,
                                          int32_t* owner);

#ifdef __cplusplus
}
#endif
#endif /* DROPS_B200_H */

// vecops.cuh -- streaming vector kernels with fused deterministic reductions (vecops.cu).
#pragma once
#include "common.cuh"

namespace db200 {

// Device-side scalar block shared by the kernels of one solver run.  Everything a later kernel
// needs (alpha, beta, stop flag) is produced on the device by the last block of an earlier
// kernel, so an iteration needs no host round trip.
struct SolverScalars {
    double rho, rho_1, lambda, rr, resid, normb, tol, bb;
    double a1, b1, b0;           // Lanczos coefficients
    double h[2];                 // scratch
    double omega;
    double alpha_prev;           // alpha of the last completed PCG iteration (x update carried by the next SpMV)
    int iter, done, converged, relative, max_iter, pc, hist_cap;
    int x_iter;                  // PCG with the deferred x update: iterations whose alpha*p is already in x
};

#ifdef __CUDACC__
// Scalar bookkeeping after the reductions of the PCG setup / update pass; run by one thread
// (the last block of the fused kernel on one GPU, k_dist_scalars on a partitioned system).
__device__ __forceinline__ void pcg_setup_finish(SolverScalars* S, double* hist, double rr, double bb, double rho)
{
    double normb = sqrt(bb);                                    // krylovsolver.h:130-133
    if (normb == 0.0 || !S->relative) normb = 1.0;
    const double resid = sqrt(rr) / normb;                      // :129, :134
    S->normb = normb; S->resid = resid; S->rr = rr;
    S->rho = rho; S->rho_1 = 0.0;                               // :144
    S->iter = 0;
    if (hist && S->hist_cap > 0) hist[0] = resid;
    if (resid <= S->tol) { S->done = 1; S->converged = 1; }     // :137-141
    else if (S->max_iter <= 0) S->done = 1;
}
__device__ __forceinline__ void pcg_update_finish(SolverScalars* S, double* hist, double rr, double rho)
{
    const double resid = sqrt(rr) / S->normb;
    const int it = S->iter + 1;
    S->iter = it; S->rr = rr; S->resid = resid;
    S->rho_1 = S->rho; S->rho = rho;
    if (hist && it < S->hist_cap) hist[it] = resid;
    if (resid <= S->tol) { S->done = 1; S->converged = 1; }     // :164-168
    else if (it >= S->max_iter) S->done = 1;
}
#endif

// ---- standalone reductions / updates (C ABI: dot, norm, axpy, ...) ---------------------------
int vec_dot(size_t n, const double* x, const double* y, double* d_out);          // d_out: device scalar
int vec_supnorm(size_t n, const double* x, double* d_out);
int vec_axpy(size_t n, double a, const double* x, double* y);                    // y += a*x
int vec_z_xpay(size_t n, double* z, const double* x, double a, const double* y); // z = x + a*y
int vec_z_xpaypby2(size_t n, double* z, const double* x, double a, const double* y, double b, const double* y2);
int vec_jacobi(size_t n, double omega, const double* d, const double* b, double* x);  // x = (omega*b)/d
int vec_mul(size_t n, double* z, const double* x, const double* y);               // z = x o y   (DiagPcCL::Apply, precond.h:626-631)
int vec_fill(size_t n, double* x, double v);
int vec_fill_hash(size_t n, double* x);
int vec_copy(size_t n, double* dst, const double* src);
// dst[pos] = perm[pos] >= 0 ? src[perm[pos]] : fill   /   dst[perm[pos]] = src[pos]   (row-ordered numbering)
int vec_gather_perm(size_t npos, const int32_t* perm, const double* src, double* dst, double fill);
int vec_scatter_perm(size_t npos, const int32_t* perm, const double* src, double* dst);

// ---- PCG (num/krylovsolver.h:107-179) ------------------------------------------------------------
// setup: given r = b - A x: normb, resid, hist[0], z = M^-1 r, rho = z.r, p = z; may set done.
// part_out != nullptr (partitioned system): the reductions are stored unrounded to part_out[0..2]
// (|r|^2, |b|^2, z.r) and the scalar bookkeeping is left to the caller.
int pcg_setup(size_t n, const double* r, const double* b, const double* d, double* p,
              SolverScalars* S, double* hist, dd* part_out = nullptr);
// update: alpha = rho/lambda; x += alpha p; r -= alpha q; resid = |r|/normb; rho_1 = rho;
//         rho = (M^-1 r).r; ++iter; hist[iter] = resid; done if resid <= tol or iter == max_iter.
// part_out != nullptr: partials (|r|^2, z.r) to part_out[0..1], no bookkeeping.
int pcg_update(size_t n, double* x, double* r, const double* p, const double* q, const double* d,
               SolverScalars* S, double* hist, dd* part_out = nullptr);
// local double-double partial of x.y -> *part_out (no rounding)
int vec_dot_partial(size_t n, const double* x, const double* y, dd* part_out);
// direction: p = M^-1 r + (rho/rho_1) p
int pcg_direction(size_t n, double* p, const double* r, const double* d, SolverScalars* S);
// The same iteration with the x update moved from the update into the direction kernel (10 instead of
// 11 vector passes: p is read once): pcg_update_r does r -= alpha q, z, the two reductions and the
// scalar bookkeeping; pcg_direction_x does x += alpha p for the iteration that just ran (also when it
// was the last one) and then p = z + beta p unless the solve is done.  Element arithmetic unchanged.
// di (optional): d as one byte per entry + its distinct values (the values travel in the kernels' parameter
// block); d_i = di->dict.v[di->idx[i]] has the bits of d[i], and 1 instead of 8 bytes per entry are read.
constexpr int kDiagDict = 32;
struct DiagDict { double v[kDiagDict]; };
struct DiagIndex { const uint8_t* idx = nullptr; DiagDict dict; };
int pcg_update_r(size_t n, double* r, const double* q, const double* d, SolverScalars* S, double* hist,
                 dd* part_out = nullptr, const DiagIndex* di = nullptr);
int pcg_direction_x(size_t n, double* x, double* p, const double* r, const double* d, SolverScalars* S,
                    const DiagIndex* di = nullptr);
// dot product into a device scalar, skipped while *done != 0
int vec_dot_unless(size_t n, const double* x, const double* y, double* d_out, const int* done);

// ---- CG (num/krylovsolver.h:49-103) --------------------------------------------------------------
int cg_setup(size_t n, const double* r, const double* b, double* dvec, SolverScalars* S, double* hist);
int cg_update(size_t n, double* x, double* r, const double* dvec, const double* Ad, SolverScalars* S, double* hist);
int cg_direction(size_t n, double* dvec, const double* r, SolverScalars* S);

// ---- GMRES building blocks (num/krylovsolver.h:320-461) -------------------------------------------
// out = in * (1/(*d_scale))                                                      :366, :384
int vec_scale_inv(size_t n, double* out, const double* in, const double* d_scale, const int* done = nullptr);
// if vprev: w -= (*d_hprev) * vprev.  Then *d_hout = dot(w, vk) (vk != null) or |w|_2 (vk == null).
// part_out != nullptr (partitioned system): the local double-double partial of the reduction goes to
// *part_out unrounded (for the norm: of w.w, without the sqrt) and d_hout is not written.
int mgs_step(size_t n, double* w, const double* vprev, const double* d_hprev, const double* vk,
             double* d_hout, bool want_norm, dd* part_out = nullptr);
// x_e = (((x_e + y0 v0_e) + y1 v1_e) + ...) for k vectors; y on device.         :282-297
int multi_axpy(size_t n, double* x, const double* const* d_vptrs, const double* d_y, int k);

// Classical Gram-Schmidt (StandardGramSchmidt :465-494, the reference's ModGMRES default :862-866):
// d_h[k] = dot(w, v_k), k < nv, all against the unmodified w in ONE pass (part_out != nullptr: unrounded local
// partials part_out[k] instead, for a partitioned system) ...
int cgs_dots(size_t n, const double* w, const double* const* d_vptrs, int nv, double* d_h, dd* part_out = nullptr,
             const int* done = nullptr);
// ... then w -= d_h[k] v_k for k = 0..nv-1 in that order and *d_norm = |w|_2 (part_out: local partial of w.w)
int cgs_update(size_t n, double* w, const double* const* d_vptrs, const double* d_h, int nv, double* d_norm,
               dd* part_out = nullptr, const int* done = nullptr);
// Device-side Givens step of GMRES column i ("gmres.device_givens"): see k_gmres_scale_givens.  done stays the first member:
// the kernels of later steps take &ctrl->done as their stop flag.
struct GmresCtrl { int done; int steps; double resid; double normb; double tol; };
// v_{i+1} = w / h_{i+1,i} and the Givens step of column i in one launch
int gmres_scale_givens(size_t n, double* out, const double* in, int i, int m, const double* d_h, double* d_cs, double* d_sn,
                       double* d_s, double* d_H, GmresCtrl* ctrl);

// ---- MINRES building blocks (num/krylovsolver.h:926-1152) ----------------------------------------
// q2 -= (*d_a1) * q1;  *d_b1 = |q2|_2                                           :933-934
int lanczos_orth(size_t n, double* q2, const double* q1, const double* d_a1, double* d_b1, dd* part_out = nullptr);
// PLanczos: t2 -= a1*t1 (t1 may differ from the dot partner), then q2 = M^-1 t2, b1sq = q2.t2
int planczos_orth(size_t n, double* t2, const double* t1, const double* d_a1, double* q2,
                  int pc, double omega, const double* d, double* d_b1sq, dd* part_out = nullptr);
// out = in / s (mode 0, :941, :965)  or  out = in * (1/s) (mode 1, :1005-1006); out may alias in
int vec_div_scalar(size_t n, double* out, const double* in, double s, int mode);
// p0 = (q - r0*pm2 - r1*pm1) {/ r2 | * (1/r2)};  x += f*p0                        :1090-1131
// mode 1: p0 = q/r0c (k = 1);  mode 2: p0 = (q - r0*pm1)/r1 (k = 2);  mode 3: general.
int minres_update(size_t n, int mode, const double* q, const double* pm2, const double* pm1, double* p0,
                  double* x, double r0, double r1, double r2, double f);
// ---- BiCGStab building blocks (num/krylovsolver.h:1189-1258) ---------------------------------------
// p = first ? r : r + beta*(p - omega*v);  phat = M^-1 p                                  :1216-1222
int bicg_p(size_t n, bool first, double* p, double* phat, const double* r, const double* v, double beta, double omega,
           int pc, double omega_pc, const double* d);
// s = r - alpha*v;  shat = M^-1 s;  *d_out = s.s                                          :1225-1232
int bicg_s(size_t n, double* s, double* shat, const double* r, const double* v, double alpha, int pc, double omega_pc,
           const double* d, double* d_out);
// d_out[0] = t.s, d_out[1] = t.t                                                          :1234
int bicg_ts(size_t n, const double* t, const double* s, double* d_out);
// x += alpha*phat + omega*shat;  r = s - omega*t;  d_out[0] = r.r, d_out[1] = rtilde.r     :1235-1239, :1209
int bicg_x(size_t n, double* x, double* r, const double* phat, const double* shat, const double* s, const double* t,
           const double* rtilde, double alpha, double omega, double* d_out);
// x += a*y with the reference's rounding (x = x + a*y)
// ---- GCR building blocks (num/krylovsolver.h:1410-1483) --------------------------------------------
// One orthogonalisation step: if vprev: vn -= a*vprev, sn -= a*sprev, (vnacc -= a*vaccprev) with a = *d_aprev;
// then *d_out = dot(vnacc, vacc_cur)  (or |vnacc|_2 when vacc_cur == nullptr).  vnacc == nullptr: the
// accumulated twin is vn itself.
int gcr_orth(size_t n, double* vn, double* sn, double* vnacc, const double* vprev, const double* sprev,
             const double* vaccprev, const double* d_aprev, const double* vacc_cur, double* d_out);
// vn /= beta, sn /= beta, (vnacc /= beta) with beta = *d_beta;  *d_gamma = dot(racc, vnacc)
int gcr_scale(size_t n, double* vn, double* sn, double* vnacc, const double* racc, const double* d_beta, double* d_gamma);
// x += g*sn, r -= g*vn, (racc -= g*vnacc) with g = *d_gamma;  *d_rr = racc.racc
int gcr_update(size_t n, double* x, double* r, double* racc, const double* sn, const double* vn, const double* vnacc,
               const double* d_gamma, double* d_rr);

// ---- generic-preconditioner PCG (PCG / PCGNE with a preconditioner that is not an elementwise
//      formula: ChebyshevPcCL) ----------------------------------------------------------------------
// S->rho = z.r   (run after pcg_setup / pcg_update executed with pc = DUMMY, which left rho = r.r)
int pcg_set_rho(size_t n, const double* z, const double* r, SolverScalars* S);
// p = z + (rho/rho_1) p                                               krylovsolver.h:173-174, :246
int pcg_direction_z(size_t n, double* p, const double* z, SolverScalars* S);

// ---- Chebyshev polynomial preconditioner (num/precond.h:1359-1408) --------------------------------
// W = ((1/d)*res)*one_over_theta;  x = W                                       precond.h:1383-1389
int cheb_first(size_t n, double* W, double* x, const double* res, const double* d, double one_over_theta);
// W = W*dt1 + ((dt2*(1/d))*res);  x += W                                       precond.h:1402-1405
int cheb_step(size_t n, double* W, double* x, const double* res, const double* d, double dt1, double dt2);
// y = (a*y) + (b*x)                                          IDR(s): U[k] = c[0]*U[k] + omega*v  :1630
int vec_axpby(size_t n, double a, double* y, double b, const double* x);
// d_out[j] = dot(v_j, x), j < k <= 4: one pass over x for k reductions
int multi_dot(size_t n, const double* x, const double* const* h_vptrs, int k, double* d_out);
// x *= (1/s) style helpers are vec_div_scalar; d_out[0] = |t|^2, d_out[1] = t.r       IDR(s) :1676
// (bicg_ts(n, t, r, d_out) gives d_out[0] = t.r, d_out[1] = t.t)

// r = b - t  style helpers
int vec_sub(size_t n, double* out, const double* a, const double* b);            // out = a - b

}  // namespace db200

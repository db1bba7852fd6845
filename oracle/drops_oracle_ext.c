/* oracle/drops_oracle_ext.c -- TEST INFRASTRUCTURE ONLY (same rules as drops_oracle.c: only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it, as the checker).
 *
 * Plain-C restatement of the callers next to the core path; paths
 * relative to /root/reference/src:
 *   BBTDiag                        num/precond.h:56-75
 *   MaxEigenvalueScaledMatrix      num/precond.h:1080-1096
 *   Chebyshev<false>               num/precond.h:1359-1408
 *   PCG with ChebyshevPcCL         num/krylovsolver.h:107-179, num/precond.h:1463-1529
 *   PCGNE (+ ChebyshevBBTPcCL)     num/krylovsolver.h:199-252, num/precond.h:1531-1567
 *   IDRS                           num/krylovsolver.h:1579-1698
 *   BlockMatrixBaseCL operator*    num/spblockmat.h:225-273
 *   CompositeMatrixBaseCL operator* num/spblockmat.h:376-404
 * Operation order as in the reference (left-to-right sums, products rounded before the add).
 * Pinned by oracle/_ref (the compiled reference) in tests/test_oracle.py; the reference's own
 * tests hold a block-matrix KAT (tests/blockmat.cpp) but none for PCGNE / Chebyshev / IDR(s):
 * for those "parity unpinned by reference tests", pinned by the compiled reference only.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* from drops_oracle.c */
void orc_y_Ax(double* y, size_t num_rows, const double* Aval, const size_t* Arow, const size_t* Acol, const double* x);
void orc_y_ATx(double* y, size_t num_rows, size_t num_cols, const double* Aval, const size_t* Arow,
               const size_t* Acol, const double* x);
void orc_get_diag(double* diag, size_t n, const double* Aval, const size_t* Arow, const size_t* Acol);
double orc_dot(size_t n, const double* v, const double* w);
double orc_norm(size_t n, const double* v);
double orc_kahan_dot(size_t n, const double* a, const double* b);
double orc_kahan_norm(size_t n, const double* a);
void orc_jacobi0(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol, double omega, const double* b,
                 double* x);

static double* vnew(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

typedef struct {
    size_t rows, cols;
    const double* val; const size_t* row; const size_t* col;
    int ne;                 /* 0: operator A (square), 1: A A^T */
    double* tmp;            /* cols entries (ne) */
} ext_op;

/* y = Op x: A*x, or CompositeMatrixBaseCL(&A, TRANSP_MUL, ex, &A, MUL, ex)*x = A*(A^T x)  spblockmat.h:376-397 */
static void op_mul(const ext_op* O, double* y, const double* x)
{
    if (!O->ne) { orc_y_Ax(y, O->rows, O->val, O->row, O->col, x); return; }
    orc_y_ATx(O->tmp, O->rows, O->cols, O->val, O->row, O->col, x);
    orc_y_Ax(y, O->rows, O->val, O->row, O->col, O->tmp);
}

/* precond.h:61-75 */
ORC_API void orc_bbt_diag(size_t rows, const double* Aval, const size_t* Arow, double* ret)
{
    size_t i, l;
    for (i = 0; i < rows; ++i) {
        ret[i] = 0.0;
        for (l = Arow[i]; l < Arow[i + 1]; ++l) { const double Bik = Aval[l]; ret[i] += Bik * Bik; }
    }
}

/* precond.h:1080-1096 with DummyExchangeCL (Norm = sqrt of the Kahan sum, exchange.h:76-84) */
static double max_eig_scaled(const ext_op* O, const double* D, int iter)
{
    const size_t n = O->rows;
    double *y = vnew(n), *t = vnew(n), lambda = 0.0, nrm;
    size_t k; int i;
    for (k = 0; k < n; ++k) y[k] = D[k];
    for (k = 0; k < n; ++k) y[k] /= D[k];
    nrm = orc_kahan_norm(n, y);
    for (k = 0; k < n; ++k) y[k] /= nrm;
    for (i = 0; i < iter; ++i) {
        op_mul(O, t, y);
        for (k = 0; k < n; ++k) y[k] = t[k] / D[k];
        lambda = orc_kahan_norm(n, y);
        for (k = 0; k < n; ++k) y[k] /= lambda;
    }
    free(y); free(t);
    return lambda;
}

ORC_API double orc_max_eig_scaled(size_t rows, size_t cols, const double* Aval, const size_t* Arow, const size_t* Acol,
                                  int ne, const double* D, int iter)
{
    ext_op O = { rows, cols, Aval, Arow, Acol, ne, NULL };
    double l;
    O.tmp = vnew(cols);
    l = max_eig_scaled(&O, D, iter);
    free(O.tmp);
    return l;
}

/* precond.h:1359-1408, InitialGuess == false */
static void chebyshev0(const ext_op* O, double* x, const double* b, const double* diag, int degree, double lmax,
                       double ratio)
{
    const size_t n = O->rows;
    double alpha, beta, delta, theta, s1, oot, rhok, rhokp1, dt1, dt2;
    double *inv, *W, *res, *t;
    size_t k; int it;
    if (degree == 0) return;
    inv = vnew(n); W = vnew(n); res = vnew(n); t = vnew(n);
    for (k = 0; k < n; ++k) inv[k] = 1.0 / diag[k];
    alpha = lmax / ratio; beta = lmax;
    delta = 2.0 / (beta - alpha);
    theta = 0.5 * (beta + alpha);
    s1 = theta * delta;
    oot = 1.0 / theta;
    for (k = 0; k < n; ++k) res[k] = b[k];
    for (k = 0; k < n; ++k) W[k] = inv[k] * res[k] * oot;
    for (k = 0; k < n; ++k) x[k] = W[k];
    rhok = 1.0 / s1;
    for (it = 0; it < degree - 1; ++it) {
        op_mul(O, t, x);
        for (k = 0; k < n; ++k) res[k] = b[k] - t[k];
        rhokp1 = 1.0 / (2.0 * s1 - rhok);
        dt1 = rhokp1 * rhok;
        dt2 = 2.0 * rhokp1 * delta;
        rhok = rhokp1;
        for (k = 0; k < n; ++k) W[k] *= dt1;
        for (k = 0; k < n; ++k) W[k] += dt2 * inv[k] * res[k];
        for (k = 0; k < n; ++k) x[k] += W[k];
    }
    free(inv); free(W); free(res); free(t);
}

ORC_API void orc_chebyshev(size_t rows, size_t cols, const double* Aval, const size_t* Arow, const size_t* Acol, int ne,
                           double* x, const double* b, const double* diag, int degree, double lmax, double ratio)
{
    ext_op O = { rows, cols, Aval, Arow, Acol, ne, NULL };
    O.tmp = vnew(cols);
    chebyshev0(&O, x, b, diag, degree, lmax, ratio);
    free(O.tmp);
}

/* preconditioner state for the generic PCG below
 *   kind 0: x = b (DummyPcCL)
 *   kind 1: x = omega*b/diag   (JACPcCL on A; for the normal equations the diagonal is BBTDiag --
 *           the _PAR JACPcCL specialisation for CompositeMatrixBaseCL, precond.h:1190-1229)
 *   kind 2: ChebyshevPcCL / ChebyshevBBTPcCL
 *   kind 3: DiagPcCL, x = D o b with the caller's vector (precond.h:618-642); diag holds D */
typedef struct {
    int kind; double omega; double* diag; int degree; double lmax, ratio;
} ext_pc;

static void pc_setup(ext_pc* P, const ext_op* O, int kind, double omega, double scale, int degree, double ratio)
{
    P->kind = kind; P->omega = omega; P->degree = degree; P->ratio = ratio; P->lmax = 0.0;
    P->diag = vnew(O->rows);
    if (O->ne) orc_bbt_diag(O->rows, O->val, O->row, P->diag);
    else orc_get_diag(P->diag, O->rows, O->val, O->row, O->col);
    if (kind == 2) P->lmax = scale * max_eig_scaled(O, P->diag, 10);      /* SetDiag  precond.h:1481-1491, :1537-1553 */
}

static void pc_run(const ext_pc* P, const ext_op* O, double* x, const double* b)
{
    size_t k;
    if (P->kind == 0) memcpy(x, b, O->rows * sizeof(double));
    else if (P->kind == 1) for (k = 0; k < O->rows; ++k) x[k] = P->omega * b[k] / P->diag[k];
    else if (P->kind == 3) for (k = 0; k < O->rows; ++k) x[k] = P->diag[k] * b[k];               /* x= D_*b  :630 */
    else chebyshev0(O, x, b, P->diag, P->degree, P->lmax, P->ratio);
}

/* PCG (krylovsolver.h:107-179) on a square A, and PCGNE (:199-252) on A A^T.  Same loop shape;
 * the differences are kept literal: PCG takes lambda = ParDot(q, p) with q = A*p (:149-151),
 * PCGNE alpha = rho / Norm_sq(A^T q) (:233-234) and r -= alpha*(A*pt) (:236). */
static int pcg_ext_core(size_t rows, size_t cols, const double* Aval, const size_t* Arow, const size_t* Acol, int ne,
                        int pc_kind, double omega, double scale, int degree, double ratio, const double* user_D,
                        double* x, const double* b, int* max_iter, double* tol, int relative,
                        double* hist, int hist_cap, double* lambda_out)
{
    ext_op O = { rows, cols, Aval, Arow, Acol, ne, NULL };
    ext_pc P;
    const size_t n = rows;
    double *r = vnew(n), *z = vnew(n), *p = vnew(n), *q = vnew(n), *pt = vnew(cols);
    double rho, rho_1, resid, normb;
    size_t k; int i, conv = 0;
    O.tmp = vnew(cols);
    pc_setup(&P, &O, pc_kind, omega, scale, degree, ratio);
    if (pc_kind == 3) memcpy(P.diag, user_D, n * sizeof(double));
    if (lambda_out) *lambda_out = P.lmax;

    op_mul(&O, r, x);
    for (k = 0; k < n; ++k) r[k] = b[k] - r[k];
    if (!ne) {
        resid = orc_kahan_norm(n, r);                      /* :129 */
        normb = orc_kahan_norm(n, b);
        if (normb == 0.0 || !relative) normb = 1.0;
        resid = resid / normb;
    } else {
        normb = orc_kahan_norm(n, b);                      /* :209 */
        if (normb == 0.0 || !relative) normb = 1.0;
        resid = orc_kahan_norm(n, r) / normb;              /* :212 */
    }
    if (hist && hist_cap > 0) hist[0] = resid;
    if (resid <= *tol) { *tol = resid; *max_iter = 0; conv = 1; goto done; }
    pc_run(&P, &O, z, r);
    rho = orc_kahan_dot(n, z, r);
    memcpy(p, z, n * sizeof(double));
    for (i = 1; i <= *max_iter; ++i) {
        double alpha;
        if (!ne) {
            op_mul(&O, q, p);
            alpha = rho / orc_kahan_dot(n, q, p);
        } else {
            orc_y_ATx(pt, rows, cols, Aval, Arow, Acol, p);
            alpha = rho / orc_kahan_dot(cols, pt, pt);
            orc_y_Ax(q, rows, Aval, Arow, Acol, pt);
        }
        for (k = 0; k < n; ++k) x[k] = x[k] + alpha * p[k];
        if (!ne) for (k = 0; k < n; ++k) r[k] = r[k] + (-alpha) * q[k];     /* axpy(-alpha, q, r) :155 */
        else     for (k = 0; k < n; ++k) r[k] = r[k] - alpha * q[k];        /* r -= alpha*(A*ptacc) :236 */
        if (!ne) {
            resid = orc_kahan_norm(n, r) / normb;                            /* :158 (before Apply) */
            if (hist && i < hist_cap) hist[i] = resid;
            if (resid <= *tol) { *tol = resid; *max_iter = i; conv = 1; goto done; }
            pc_run(&P, &O, z, r);
        } else {
            pc_run(&P, &O, z, r);                                            /* :237 */
            resid = orc_kahan_norm(n, r) / normb;
            if (hist && i < hist_cap) hist[i] = resid;
            if (resid <= *tol) { *tol = resid; *max_iter = i; conv = 1; goto done; }
        }
        rho_1 = rho;
        rho = orc_kahan_dot(n, z, r);
        for (k = 0; k < n; ++k) p[k] = z[k] + (rho / rho_1) * p[k];
    }
    *tol = resid;
done:
    free(r); free(z); free(p); free(q); free(pt); free(O.tmp); free(P.diag);
    return conv;
}

ORC_API int orc_pcg_ext(size_t rows, size_t cols, const double* Aval, const size_t* Arow, const size_t* Acol, int ne,
                        int pc_kind, double omega, double scale, int degree, double ratio,
                        double* x, const double* b, int* max_iter, double* tol, int relative,
                        double* hist, int hist_cap, double* lambda_out)
{
    if (pc_kind < 0 || pc_kind > 2) return -1;
    return pcg_ext_core(rows, cols, Aval, Arow, Acol, ne, pc_kind, omega, scale, degree, ratio, NULL, x, b, max_iter, tol,
                        relative, hist, hist_cap, lambda_out);
}

/* PCGSolverCL<DiagPcCL> (krylovsolver.h:107-179 with precond.h:618-642): D has n entries. */
ORC_API int orc_pcg_diagpc(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol, const double* D,
                           double* x, const double* b, int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    return pcg_ext_core(n, n, Aval, Arow, Acol, 0, 3, 1.0, 1.0, 0, 1.0, D, x, b, max_iter, tol, relative, hist, hist_cap, NULL);
}

/* IDRS, krylovsolver.h:1579-1698.  Praw: the s shadow vectors as drawn at :1603-1608 (s*n, vector i
 * at Praw + i*n).  dot/norm are the plain sequential ones (spmat.h:164-196).  Returns the
 * reference's return value (:1689-1696), -1 for the two DROPSErrCL exits. */
ORC_API int orc_idrs(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol, int pc_kind, double omega_pc,
                     int s, double omega_bound, const double* Praw, double* x, const double* rhs,
                     int* max_iter, double* tol, int relative)
{
    double omega = 1.0, normb, normres;
    double *v = vnew(n), *t = vnew(n), *f = vnew((size_t)s), *resid = vnew(n), *c = vnew((size_t)s);
    double *P = vnew((size_t)s * n), *G = vnew((size_t)s * n), *U = vnew((size_t)s * n), *M = vnew((size_t)s * s);
    size_t e; int i, j, k, l, it, ret = -1;
#define PV(i) (P + (size_t)(i) * n)
#define GV(i) (G + (size_t)(i) * n)
#define UV(i) (U + (size_t)(i) * n)
#define MM(i, j) M[(size_t)(i) * s + (j)]
    normb = orc_norm(n, rhs);
    if (normb == 0.0 || !relative) normb = 1.0;
    orc_y_Ax(resid, n, Aval, Arow, Acol, x);
    for (e = 0; e < n; ++e) resid[e] = rhs[e] - resid[e];
    normres = orc_norm(n, resid);
    if (normres / normb <= *tol) { *tol = normres; *max_iter = 0; ret = 1; goto done; }
    memcpy(P, Praw, (size_t)s * n * sizeof(double));
    for (k = 0; k < s; ++k) {
        double nn;
        for (j = 0; j < k; ++j) {
            const double sm = orc_dot(n, PV(j), PV(k));
            for (e = 0; e < n; ++e) PV(k)[e] -= sm * PV(j)[e];
        }
        nn = orc_norm(n, PV(k));
        for (e = 0; e < n; ++e) PV(k)[e] /= nn;
    }
    for (i = 0; i < s; ++i) MM(i, i) = 1.0;
    it = 0;
    while (normres / normb > *tol && it < *max_iter) {
        for (i = 0; i < s; ++i) f[i] = orc_dot(n, PV(i), resid);
        for (k = 0; k < s; ++k) {
            double beta;
            for (j = 0; j < s - k; ++j) {
                double cs = f[k + j];
                for (l = 0; l < j; ++l) cs -= MM(k + j, k + l) * c[l];
                c[j] = cs / MM(k + j, k + j);
            }
            memcpy(v, resid, n * sizeof(double));
            for (j = 0; j < s - k; ++j) for (e = 0; e < n; ++e) v[e] -= c[j] * GV(k + j)[e];
            if (pc_kind == 1) orc_jacobi0(n, Aval, Arow, Acol, omega_pc, v, v);
            for (e = 0; e < n; ++e) UV(k)[e] = c[0] * UV(k)[e] + omega * v[e];
            for (j = 1; j < s - k; ++j) for (e = 0; e < n; ++e) UV(k)[e] += c[j] * UV(k + j)[e];
            orc_y_Ax(GV(k), n, Aval, Arow, Acol, UV(k));
            for (i = 0; i < k; ++i) {
                const double alpha = orc_dot(n, PV(i), GV(k)) / MM(i, i);
                for (e = 0; e < n; ++e) GV(k)[e] -= alpha * GV(i)[e];
                for (e = 0; e < n; ++e) UV(k)[e] -= alpha * UV(i)[e];
            }
            for (j = 0; j < s - k; ++j) MM(k + j, k) = orc_dot(n, PV(k + j), GV(k));
            if (MM(k, k) == 0.0) goto done;
            beta = f[k] / MM(k, k);
            for (e = 0; e < n; ++e) resid[e] -= beta * GV(k)[e];
            for (e = 0; e < n; ++e) x[e] += beta * UV(k)[e];
            normres = orc_norm(n, resid);
            it++;
            if (normres / normb <= *tol) break;
            if (k + 1 < s) for (j = 1; j < s - k; ++j) f[k + j] -= beta * MM(k + j, k);
        }
        if (normres / normb <= *tol) break;
        {
            double tn, tr, rho;
            if (pc_kind == 1) orc_jacobi0(n, Aval, Arow, Acol, omega_pc, resid, v);
            else memcpy(v, resid, n * sizeof(double));
            orc_y_Ax(t, n, Aval, Arow, Acol, v);
            tn = orc_norm(n, t); tr = orc_dot(n, t, resid);
            omega = tr / (tn * tn);
            rho = fabs(tr) / (tn * normres);
            if (rho < omega_bound) omega *= omega_bound / rho;
            if (omega == 0.0) goto done;
            for (e = 0; e < n; ++e) resid[e] -= omega * t[e];
            for (e = 0; e < n; ++e) x[e] += omega * v[e];
            normres = orc_norm(n, resid);
            it++;
        }
    }
    ret = (*tol > normres / normb) ? 0 : 1;
    *max_iter = it;
    *tol = normres;
done:
    free(v); free(t); free(f); free(resid); free(c); free(P); free(G); free(U); free(M);
    return ret;
#undef PV
#undef GV
#undef UV
#undef MM
}

/* BlockMatrixBaseCL operator*, spblockmat.h:225-267.  Block b (0..3 = A, B, C, D) is absent when
 * val[b] == NULL; op[b]: 0 MUL, 1 TRANSP_MUL; brows/bcols: stored dimensions of block b.
 * nr[2] / nc[2]: num_rows(block_row) / num_cols(block_col) as the caller derived them. */
ORC_API void orc_block_mul(const double* const* val, const size_t* const* row, const size_t* const* col,
                           const size_t* brows, const size_t* bcols, const int* op, const size_t* nr, const size_t* nc,
                           const double* x, double* ret)
{
    const double* xs[2]; double* rs[2];
    double* t = vnew(nr[0] > nr[1] ? nr[0] : nr[1]);
    int br; size_t k;
    xs[0] = x; xs[1] = x + nc[0];
    rs[0] = ret; rs[1] = ret + nr[0];
    for (br = 0; br < 2; ++br) {
        const int b0 = 2 * br, b1 = 2 * br + 1;
        for (k = 0; k < nr[br]; ++k) rs[br][k] = 0.0;
        if (val[b0]) {
            if (op[b0] == 0) orc_y_Ax(rs[br], brows[b0], val[b0], row[b0], col[b0], xs[0]);
            else orc_y_ATx(rs[br], brows[b0], bcols[b0], val[b0], row[b0], col[b0], xs[0]);
        }
        if (val[b1]) {
            if (op[b1] == 0) orc_y_Ax(t, brows[b1], val[b1], row[b1], col[b1], xs[1]);
            else orc_y_ATx(t, brows[b1], bcols[b1], val[b1], row[b1], col[b1], xs[1]);
            for (k = 0; k < nr[br]; ++k) rs[br][k] += t[k];
        }
    }
    free(t);
}

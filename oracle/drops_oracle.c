/* oracle/drops_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * called from the product (drops_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Plain-C restatement of the DROPS sparse-solver hot path (reference: 56th/drops, paths below
 * relative to /root/reference/src).  Every function states the reference lines it follows and
 * keeps the reference's operation ORDER (left-to-right sums, product rounded before the add,
 * no FMA, no re-association), so that on the same CSR arrays it is bit-identical to the
 * reference compiled without --fast-math.  Build: see oracle/Makefile
 * (gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math).
 *
 * Parity pinning: tests/test_oracle.py checks this file against
 *   (1) the reference's own known-answer tests (src/tests/minres.cpp + TestMinres.ref,
 *       src/tests/blockmat.cpp + Testblockmat.ref), and
 *   (2) oracle/_ref/libdrops_ref.so = the unmodified reference headers compiled here,
 *       bit for bit, on seeded systems, and (3) the fixtures under tests/golden/ that
 *       oracle/make_golden.py generated from (2).
 * PCG+JACPcCL and serial GMRES have no test in the reference; for those
 * the pin is (2)/(3) only.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------
 * Sparse products
 * ------------------------------------------------------------------------------------- */

/* y = A x.   num/spmat.h:1682-1709 (y_Ax): per row, sum starts at 0.0 and runs over the
 * stored entries left to right; rows are distributed with "omp parallel for". */
ORC_API void orc_y_Ax(double* y, size_t num_rows, const double* Aval, const size_t* Arow,
                      const size_t* Acol, const double* x)
{
    long i;
#pragma omp parallel for
    for (i = 0; i < (long)num_rows; ++i) {
        double sum = 0.0;
        size_t nz;
        const size_t rowend = Arow[i + 1];
        for (nz = Arow[i]; nz < rowend; ++nz)
            sum += Aval[nz] * x[Acol[nz]];
        y[i] = sum;
    }
}

/* y = A^T x. num/spmat.h:1729-1760 (y_ATx via transp_mul): the result is zero-initialised and
 * every stored entry is scattered y[col] += val * x[row], rows ascending, serial. */
ORC_API void orc_y_ATx(double* y, size_t num_rows, size_t num_cols, const double* Aval,
                       const size_t* Arow, const size_t* Acol, const double* x)
{
    size_t i, nz = 0;
    for (i = 0; i < num_cols; ++i) y[i] = 0.0;
    for (i = 0; i < num_rows; ++i) {
        const double xrow = x[i];
        const size_t rowend = Arow[i + 1];
        for (; nz < rowend; ++nz)
            y[Acol[nz]] += Aval[nz] * xrow;
    }
}

/* A(i,j): num/spmat.h:956-963 -- lower_bound over the (ascending) columns of row i. */
static double orc_entry(size_t i, size_t j, const double* Aval, const size_t* Arow, const size_t* Acol)
{
    size_t lo = Arow[i], hi = Arow[i + 1];
    const size_t end = hi;
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (Acol[mid] < j) lo = mid + 1; else hi = mid;
    }
    return (lo != end && Acol[lo] == j) ? Aval[lo] : 0.0;
}

/* diag = A.GetDiag(): num/spmat.h:993-1002. */
ORC_API void orc_get_diag(double* diag, size_t n, const double* Aval, const size_t* Arow, const size_t* Acol)
{
    size_t i;
    for (i = 0; i < n; ++i) diag[i] = orc_entry(i, i, Aval, Arow, Acol);
}

/* ---------------------------------------------------------------------------------------
 * Vector kernels
 * ------------------------------------------------------------------------------------- */

/* num/spmat.h:164-170: std::inner_product, sequential from T(). */
ORC_API double orc_dot(size_t n, const double* v, const double* w)
{
    double s = 0.0; size_t i;
    for (i = 0; i < n; ++i) s = s + v[i] * w[i];
    return s;
}

/* num/spmat.h:180-189. */
ORC_API double orc_norm_sq(size_t n, const double* v)
{
    double s = 0.0; size_t i;
    for (i = 0; i < n; ++i) s += v[i] * v[i];
    return s;
}

/* num/spmat.h:191-196. */
ORC_API double orc_norm(size_t n, const double* v) { return sqrt(orc_norm_sq(n, v)); }

/* num/spmat.h:198-210. */
ORC_API double orc_supnorm(size_t n, const double* v)
{
    double r = n == 0 ? 0.0 : fabs(v[0]); size_t i;
    for (i = 1; i < n; ++i) { const double t = fabs(v[i]); if (t > r) r = t; }
    return r;
}

/* num/spmat.h:317-333: Kahan-compensated inner product (what DummyExchangeCL::ParDot /
 * Norm use, parallel/exchange.h:58-84). */
ORC_API double orc_kahan_dot(size_t n, const double* a, const double* b)
{
    double sum = 0.0, c = 0.0, t, y; size_t i;
    for (i = 0; i < n; ++i) {
        y = a[i] * b[i] - c;
        t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
    return sum;
}

ORC_API double orc_kahan_norm(size_t n, const double* a) { return sqrt(orc_kahan_dot(n, a, a)); }

/* num/spmat.h:212-218: y += a*x. */
ORC_API void orc_axpy(size_t n, double a, const double* x, double* y)
{
    size_t i; for (i = 0; i < n; ++i) y[i] = y[i] + a * x[i];
}

/* num/spmat.h:220-227: z = x + a*y. */
ORC_API void orc_z_xpay(size_t n, double* z, const double* x, double a, const double* y)
{
    size_t i; for (i = 0; i < n; ++i) z[i] = x[i] + a * y[i];
}

/* num/spmat.h:229-237: z = x + a*y + b*y2. */
ORC_API void orc_z_xpaypby2(size_t n, double* z, const double* x, double a, const double* y,
                            double b, const double* y2)
{
    size_t i; for (i = 0; i < n; ++i) z[i] = (x[i] + a * y[i]) + b * y2[i];
}

/* ---------------------------------------------------------------------------------------
 * Preconditioners.  pc_kind: 0 = DummyPcCL (precond.h:1032-1065, x = b),
 *                            1 = serial JACPcCL = PreGSCL<P_JAC0> (precond.h:874, :403-427).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    size_t n;
    const double* val; const size_t* row; const size_t* col;
    int pc_kind; double omega;
} orc_sys;

/* SolveGSstep<PB_JAC0> precond.h:153-169: linear search for the first entry whose column is
 * the row index, then x_i = omega*b_i / a_ii evaluated left to right ((omega*b_i)/a_ii).
 * P_JAC0 has HasOmega == true (precond.h:112). */
ORC_API void orc_jacobi0(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                         double omega, const double* b, double* x)
{
    long i;
#pragma omp parallel for
    for (i = 0; i < (long)n; ++i) {
        size_t nz = Arow[i];
        const size_t end = Arow[i + 1];
        while (Acol[nz] != (size_t)i && nz < end) ++nz;
        x[i] = omega * b[i] / Aval[nz];
    }
}

static void pc_apply(const orc_sys* S, double* x, const double* b)
{
    if (S->pc_kind == 1) orc_jacobi0(S->n, S->val, S->row, S->col, S->omega, b, x);
    else memcpy(x, b, S->n * sizeof(double));
}

static double* vec_new(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

/* ---------------------------------------------------------------------------------------
 * CG  -- num/krylovsolver.h:49-103 with DummyExchangeCL (Norm/Norm_sq/ParDot are Kahan sums
 * and copy into the *_acc argument, parallel/exchange.h:58-84).
 * hist (may be NULL, capacity hist_cap) receives sqrt(resid) per iteration as printed at :86.
 * ------------------------------------------------------------------------------------- */
ORC_API int orc_cg(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                   double* x, const double* b, int* max_iter, double* tol, int relative,
                   double* hist, int hist_cap)
{
    double *r = vec_new(n), *racc = vec_new(n), *d = vec_new(n), *Ad = vec_new(n);
    double normb, res = 0.0, resid;
    size_t k; int i, conv = 0;

    orc_y_Ax(r, n, Aval, Arow, Acol, x);
    for (k = 0; k < n; ++k) r[k] = r[k] - b[k];            /* r = A*x - b       :61 */
    normb = orc_kahan_norm(n, b);                          /* ExX.Norm(b)       :62 */
    resid = orc_kahan_dot(n, r, r);                        /* Norm_sq(r,&r_acc) :64 */
    memcpy(racc, r, n * sizeof(double));
    for (k = 0; k < n; ++k) d[k] = -racc[k];               /* d = -r_acc        :65 */
    if (normb == 0.0 || !relative) normb = 1.0;

    if ((res = sqrt(resid) / normb) <= *tol) {
        *tol = res; *max_iter = 0; conv = 1; goto done;
    }
    for (i = 1; i <= *max_iter; ++i) {
        double delta, alpha, beta;
        orc_y_Ax(Ad, n, Aval, Arow, Acol, d);
        delta = orc_kahan_dot(n, Ad, d);
        alpha = resid / delta;
        beta = resid;
        orc_axpy(n, alpha, d, x);
        orc_axpy(n, alpha, Ad, r);
        resid = orc_kahan_dot(n, r, r);
        memcpy(racc, r, n * sizeof(double));
        if (hist && i < hist_cap) hist[i] = sqrt(resid);
        if ((res = sqrt(resid) / normb) <= *tol) {
            *tol = res; *max_iter = i; conv = 1; goto done;
        }
        beta = resid / beta;
        for (k = 0; k < n; ++k) d[k] = beta * d[k] - racc[k];   /* :99 */
    }
    *tol = res;
done:
    free(r); free(racc); free(d); free(Ad);
    return conv;
}

/* ---------------------------------------------------------------------------------------
 * PCG -- num/krylovsolver.h:107-179 with DummyExchangeCL.
 * ------------------------------------------------------------------------------------- */
ORC_API int orc_pcg(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                    int pc_kind, double omega, double* x, const double* b,
                    int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    orc_sys S = { n, Aval, Arow, Acol, pc_kind, omega };
    double *p = vec_new(n), *z = vec_new(n), *zacc = vec_new(n), *q = vec_new(n),
           *qacc = vec_new(n), *r = vec_new(n), *racc = vec_new(n);
    double rho, rho_1 = 0.0, resid, normb;
    size_t k; int i, conv = 0;
    (void)rho_1;

    orc_y_Ax(r, n, Aval, Arow, Acol, x);
    for (k = 0; k < n; ++k) r[k] = b[k] - r[k];            /* r = b - A*x        :125 */
    resid = orc_kahan_norm(n, r);                          /* Norm(r,&r_acc)     :129 */
    memcpy(racc, r, n * sizeof(double));
    normb = orc_kahan_norm(n, b);                          /*                    :130 */
    if (normb == 0.0 || !relative) normb = 1.0;
    resid = resid / normb;
    if (hist && hist_cap > 0) hist[0] = resid;
    if (resid <= *tol) { *tol = resid; *max_iter = 0; conv = 1; goto done; }

    pc_apply(&S, z, r);                                    /* :143 */
    rho = orc_kahan_dot(n, z, racc);                       /* :144, copies z -> z_acc */
    memcpy(zacc, z, n * sizeof(double));
    memcpy(p, zacc, n * sizeof(double));

    for (i = 1; i <= *max_iter; ++i) {
        double lambda, alpha;
        orc_y_Ax(q, n, Aval, Arow, Acol, p);               /* :149 */
        lambda = orc_kahan_dot(n, q, p);                   /* :151, copies q -> q_acc */
        memcpy(qacc, q, n * sizeof(double));
        alpha = rho / lambda;
        orc_axpy(n, alpha, p, x);
        orc_axpy(n, -alpha, q, r);
        orc_axpy(n, -alpha, qacc, racc);
        resid = orc_kahan_norm(n, racc) / normb;           /* :158 */
        if (hist && i < hist_cap) hist[i] = resid;
        if (resid <= *tol) { *tol = resid; *max_iter = i; conv = 1; goto done; }
        pc_apply(&S, z, r);                                /* :170 */
        rho_1 = rho;
        rho = orc_kahan_dot(n, z, racc);
        memcpy(zacc, z, n * sizeof(double));
        orc_z_xpay(n, p, zacc, rho / rho_1, p);            /* :173 */
    }
    *tol = resid;
done:
    free(p); free(z); free(zacc); free(q); free(qacc); free(r); free(racc);
    return conv;
}

/* ---------------------------------------------------------------------------------------
 * LinComb -- num/spmat.h:1457-1551: C = coeffA*A + coeffB*B, pattern = row-wise merge of the two
 * ascending column lists.  Two calls: with Ccol == NULL only Crow (row_beg of the result) is filled.
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_lincomb(size_t rows, double coeffA, const double* Aval, const size_t* Arow, const size_t* Acol,
                         double coeffB, const double* Bval, const size_t* Brow, const size_t* Bcol,
                         size_t* Crow, size_t* Ccol, double* Cval)
{
    size_t row;
    if (Ccol == NULL) {
        Crow[0] = 0;
        for (row = 0; row < rows; ++row) {                               /* :1487-1503 */
            size_t i = 0, iA = Arow[row], iB = Brow[row];
            const size_t eA = Arow[row + 1], eB = Brow[row + 1];
            for (; iA != eA && iB != eB; ++i)
                if (Bcol[iB] < Acol[iA]) ++iB;
                else if (Acol[iA] < Bcol[iB]) ++iA;
                else { ++iA; ++iB; }
            Crow[row + 1] = Crow[row] + i + (eA - iA) + (eB - iB);
        }
        return;
    }
    for (row = 0; row < rows; ++row) {                                   /* :1515-1545 */
        size_t i = Crow[row], iA = Arow[row], iB = Brow[row];
        const size_t eA = Arow[row + 1], eB = Brow[row + 1];
        for (; iA != eA && iB != eB; ++i)
            if (Bcol[iB] < Acol[iA]) { Cval[i] = coeffB * Bval[iB]; Ccol[i] = Bcol[iB++]; }
            else if (Acol[iA] < Bcol[iB]) { Cval[i] = coeffA * Aval[iA]; Ccol[i] = Acol[iA++]; }
            else { Cval[i] = coeffA * Aval[iA++] + coeffB * Bval[iB]; Ccol[i] = Bcol[iB++]; }
        for (; iA < eA; ++iA, ++i) { Cval[i] = coeffA * Aval[iA]; Ccol[i] = Acol[iA]; }
        for (; iB < eB; ++iB, ++i) { Cval[i] = coeffB * Bval[iB]; Ccol[i] = Bcol[iB]; }
    }
}

/* ---------------------------------------------------------------------------------------
 * BiCGStab -- num/krylovsolver.h:1189-1258 (BICGSTAB, serial).  Plain dot / norm
 * (std::inner_product, spmat.h:164-196), not the Kahan sums of the exchange class.
 * Returns 1 converged, 0 not converged, -1 breakdown (the reference prints a message and
 * returns false with max_iter = i).
 * ------------------------------------------------------------------------------------- */
ORC_API int orc_bicgstab(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                         int pc_kind, double omega_pc, double* x, const double* b,
                         int* max_iter, double* tol, int relative)
{
    orc_sys S = { n, Aval, Arow, Acol, pc_kind, omega_pc };
    double rho_1 = 0.0, rho_2 = 0.0, alpha = 0.0, beta = 0.0, omega = 0.0, normb, resid;
    double *p = vec_new(n), *phat = vec_new(n), *s = vec_new(n), *shat = vec_new(n), *t = vec_new(n),
           *v = vec_new(n), *r = vec_new(n), *rtilde = vec_new(n);
    size_t k; int i, ret = 0;

    normb = orc_norm(n, b);                                            /* :1196 */
    orc_y_Ax(r, n, Aval, Arow, Acol, x);
    for (k = 0; k < n; ++k) r[k] = b[k] - r[k];                        /* :1197 */
    memcpy(rtilde, r, n * sizeof(double));
    if (normb == 0.0 || !relative) normb = 1.0;
    resid = orc_norm(n, r) / normb;
    if (resid <= *tol) { *tol = resid; *max_iter = 0; ret = 1; goto done; }
    for (i = 1; i <= *max_iter; ++i) {
        rho_1 = orc_dot(n, rtilde, r);                                 /* :1209 */
        if (rho_1 == 0.0) { *tol = orc_norm(n, r) / normb; *max_iter = i; ret = -1; goto done; }
        if (i == 1) memcpy(p, r, n * sizeof(double));
        else {
            beta = (rho_1 / rho_2) * (alpha / omega);                  /* :1219 */
            for (k = 0; k < n; ++k) p[k] = r[k] + beta * (p[k] - omega * v[k]);
        }
        pc_apply(&S, phat, p);
        orc_y_Ax(v, n, Aval, Arow, Acol, phat);
        alpha = rho_1 / orc_dot(n, rtilde, v);                         /* :1224 */
        for (k = 0; k < n; ++k) s[k] = r[k] - alpha * v[k];
        if ((resid = orc_norm(n, s) / normb) < *tol) {                 /* :1226 */
            for (k = 0; k < n; ++k) x[k] += alpha * phat[k];
            *tol = resid; *max_iter = i; ret = 1; goto done;
        }
        pc_apply(&S, shat, s);
        orc_y_Ax(t, n, Aval, Arow, Acol, shat);
        omega = orc_dot(n, t, s) / orc_dot(n, t, t);                   /* :1234 */
        for (k = 0; k < n; ++k) x[k] += alpha * phat[k] + omega * shat[k];
        for (k = 0; k < n; ++k) r[k] = s[k] - omega * t[k];
        rho_2 = rho_1;
        if ((resid = orc_norm(n, r) / normb) < *tol) { *tol = resid; *max_iter = i; ret = 1; goto done; }
        if (omega == 0.0) { *tol = orc_norm(n, r) / normb; *max_iter = i; ret = -1; goto done; }
    }
    *tol = resid;
done:
    free(p); free(phat); free(s); free(shat); free(t); free(v); free(r); free(rtilde);
    return ret;
}

/* ---------------------------------------------------------------------------------------
 * GCR(m) -- num/krylovsolver.h:1410-1483 with DummyExchangeCL: Norm / ParDot are Kahan sums,
 * GetAccumulate is a copy.  The reference carries "accumulated" twins (racc, vnacc, vacc[i])
 * beside r, vn, v[i]; serially they hold the same values UNTIL the truncation branch
 * (:1466-1477) replaces s[min_idx] and v[min_idx] but not vacc[min_idx].  From then on the
 * twins differ, and the coefficients (computed from the acc twins) no longer match the vectors
 * they are applied to.  This restatement keeps the twins, i.e. reproduces that behaviour.
 * ------------------------------------------------------------------------------------- */
ORC_API int orc_gcr(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                    int pc_kind, double omega_pc, double* x, const double* b, int m,
                    int* max_iter, double* tol, int relative)
{
    orc_sys S = { n, Aval, Arow, Acol, pc_kind, omega_pc };
    double *r = vec_new(n), *racc = vec_new(n), *sn = vec_new(n), *vn = vec_new(n), *vnacc = vec_new(n);
    double **s, **v, **vacc, *a;
    double normb, resid;
    size_t q; int k, i, stored = 0, ret = 0;

    m = (m <= *max_iter) ? m : *max_iter;                              /* :1416 */
    if (m < 1) m = 1;
    s = (double**)calloc((size_t)m, sizeof(double*));
    v = (double**)calloc((size_t)m, sizeof(double*));
    vacc = (double**)calloc((size_t)m, sizeof(double*));
    a = (double*)calloc((size_t)m, sizeof(double));
    orc_y_Ax(r, n, Aval, Arow, Acol, x);
    for (q = 0; q < n; ++q) r[q] = b[q] - r[q];                        /* :1418 */
    normb = orc_kahan_norm(n, b);
    if (normb == 0.0 || !relative) normb = 1.0;
    resid = orc_kahan_norm(n, r) / normb;                              /* Norm(r, false, &racc) */
    memcpy(racc, r, n * sizeof(double));
    for (k = 0; k < *max_iter; ++k) {
        double beta, gamma;
        if (resid < *tol) { *tol = resid; *max_iter = k; ret = 1; goto done; }
        pc_apply(&S, sn, r);
        orc_y_Ax(vn, n, Aval, Arow, Acol, sn);                         /* vn = A*GetAccumulate(sn) */
        memcpy(vnacc, vn, n * sizeof(double));
        for (i = 0; i < k && i < m; ++i) {                             /* :1442-1449 */
            const double alpha = orc_kahan_dot(n, vnacc, vacc[i]);
            a[i] = alpha;
            for (q = 0; q < n; ++q) vn[q] -= alpha * v[i][q];
            for (q = 0; q < n; ++q) vnacc[q] -= alpha * vacc[i][q];
            for (q = 0; q < n; ++q) sn[q] -= alpha * s[i][q];
        }
        beta = orc_kahan_norm(n, vnacc);
        for (q = 0; q < n; ++q) vn[q] /= beta;
        for (q = 0; q < n; ++q) vnacc[q] /= beta;
        for (q = 0; q < n; ++q) sn[q] /= beta;
        gamma = orc_kahan_dot(n, racc, vnacc);
        for (q = 0; q < n; ++q) x[q] += gamma * sn[q];
        for (q = 0; q < n; ++q) r[q] -= gamma * vn[q];
        for (q = 0; q < n; ++q) racc[q] -= gamma * vnacc[q];
        resid = orc_kahan_norm(n, racc) / normb;
        if (k < m) {
            s[stored] = vec_new(n); v[stored] = vec_new(n); vacc[stored] = vec_new(n);
            memcpy(s[stored], sn, n * sizeof(double)); memcpy(v[stored], vn, n * sizeof(double));
            memcpy(vacc[stored], vnacc, n * sizeof(double));
            ++stored;
        } else {
            int min_idx = 0;
            double a_min = fabs(a[0]);
            for (i = 1; i < k && i < m; ++i)
                if (fabs(a[i]) < a_min) { min_idx = i; a_min = fabs(a[i]); }
            memcpy(s[min_idx], sn, n * sizeof(double)); memcpy(v[min_idx], vn, n * sizeof(double));
            /* vacc[min_idx] keeps its old contents, as in the reference */
        }
    }
    *tol = resid;
done:
    for (i = 0; i < stored; ++i) { free(s[i]); free(v[i]); free(vacc[i]); }
    free(s); free(v); free(vacc); free(a); free(r); free(racc); free(sn); free(vn); free(vnacc);
    return ret;
}

/* ---------------------------------------------------------------------------------------
 * GMRES (serial) -- num/krylovsolver.h:258-297 (rotations, update), :320-461 (driver).
 * H is stored column-major m x m like DMatrixCL (misc/container.h:44-74).
 * ------------------------------------------------------------------------------------- */
static void gen_rot(double* dx, double* dy, double* cs, double* sn)   /* :258-271 */
{
    if (*dy == 0.0) { *cs = 1.0; *sn = 0.0; }
    else {
        const double r = sqrt(*dx * *dx + *dy * *dy);
        *cs = *dx / r; *sn = *dy / r;
    }
}
static void app_rot(double* dx, double* dy, const double* cs, const double* sn)   /* :274-279 */
{
    const double tmp = *cs * *dx + *sn * *dy;
    *dy = -*sn * *dx + *cs * *dy;
    *dx = tmp;
}
#define HH(i, j) H[(size_t)(j) * (size_t)m + (size_t)(i)]

static void gmres_update(size_t n, double* x, int k, int m, const double* H, const double* s,
                         double** v, double* y)               /* :282-297 */
{
    int i, j; size_t t;
    memcpy(y, s, (size_t)m * sizeof(double));
    for (i = k; i >= 0; --i) {
        y[i] /= HH(i, i);
        for (j = i - 1; j >= 0; --j) y[j] -= HH(j, i) * y[i];
    }
    for (i = 0; i <= k; ++i)
        for (t = 0; t < n; ++t) x[t] = x[t] + y[i] * v[i][t];
}

ORC_API int orc_gmres(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                      int pc_kind, double omega, double* x, const double* b, int m,
                      int* max_iter, double* tol, int relative, int left_precond)
{
    orc_sys S = { n, Aval, Arow, Acol, pc_kind, omega };
    double *H, *s, *cs, *sn, *yk, *w, *r, *z, *t, *tmp, **v;
    double beta, normb, resid = 0.0;
    int i, j, k, conv = 0; size_t e;

    if (m > *max_iter) m = *max_iter;                      /* :327 */
    if (m < 2) return -1;   /* the reference's inner loop never advances for m < 2 */
    H = vec_new((size_t)m * m); s = vec_new(m); cs = vec_new(m); sn = vec_new(m); yk = vec_new(m);
    w = vec_new(n); r = vec_new(n); z = vec_new(n); t = vec_new(n); tmp = vec_new(n);
    v = (double**)malloc((size_t)m * sizeof(double*));
    for (i = 0; i < m; ++i) v[i] = vec_new(n);

    orc_y_Ax(tmp, n, Aval, Arow, Acol, x);
    for (e = 0; e < n; ++e) tmp[e] = b[e] - tmp[e];
    if (!left_precond) {                                   /* :336-341 */
        memcpy(r, tmp, n * sizeof(double));
        beta = orc_norm(n, r);
        normb = orc_norm(n, b);
    } else {                                               /* :343-348 */
        pc_apply(&S, r, tmp);
        beta = orc_norm(n, r);
        pc_apply(&S, w, b);
        normb = orc_norm(n, w);
    }
    if (normb == 0.0 || !relative) normb = 1.0;
    resid = beta / normb;
    if (resid <= *tol) { *tol = resid; *max_iter = 0; conv = 1; goto done; }

    j = 1;
    while (j <= *max_iter) {
        const double ib = 1.0 / beta;
        for (e = 0; e < n; ++e) v[0][e] = r[e] * ib;      /* :366 */
        for (i = 0; i < m; ++i) s[i] = 0.0;
        s[0] = beta;
        for (i = 0; i < m - 1 && j <= *max_iter; ++i, ++j) {
            if (!left_precond) {                           /* :372-376 */
                pc_apply(&S, tmp, v[i]);
                orc_y_Ax(w, n, Aval, Arow, Acol, tmp);
            } else {                                       /* :377 */
                orc_y_Ax(tmp, n, Aval, Arow, Acol, v[i]);
                pc_apply(&S, w, tmp);
            }
            for (k = 0; k <= i; ++k) {                     /* modified Gram-Schmidt :378-381 */
                const double h = orc_dot(n, w, v[k]);
                HH(k, i) = h;
                for (e = 0; e < n; ++e) w[e] = w[e] - h * v[k][e];
            }
            HH(i + 1, i) = orc_norm(n, w);
            {
                const double ih = 1.0 / HH(i + 1, i);
                for (e = 0; e < n; ++e) v[i + 1][e] = w[e] * ih;
            }
            for (k = 0; k < i; ++k) app_rot(&HH(k, i), &HH(k + 1, i), &cs[k], &sn[k]);
            gen_rot(&HH(i, i), &HH(i + 1, i), &cs[i], &sn[i]);
            app_rot(&HH(i, i), &HH(i + 1, i), &cs[i], &sn[i]);
            app_rot(&s[i], &s[i + 1], &cs[i], &sn[i]);
            resid = fabs(s[i + 1]) / normb;                /* :393 */
            if (resid <= *tol) {                           /* :413-429 */
                if (!left_precond) {
                    memset(z, 0, n * sizeof(double));
                    gmres_update(n, z, i, m, H, s, v, yk);
                    pc_apply(&S, t, z);
                    for (e = 0; e < n; ++e) x[e] = x[e] + t[e];
                } else gmres_update(n, x, i, m, H, s, v, yk);
                *tol = resid; *max_iter = j; conv = 1; goto done;
            }
        }
        if (!left_precond) {                               /* :432-439 */
            memset(z, 0, n * sizeof(double));
            gmres_update(n, z, i - 1, m, H, s, v, yk);
            pc_apply(&S, t, z);
            for (e = 0; e < n; ++e) x[e] = x[e] + t[e];
            orc_y_Ax(tmp, n, Aval, Arow, Acol, x);
            for (e = 0; e < n; ++e) r[e] = b[e] - tmp[e];
        } else {                                           /* :440-444 */
            gmres_update(n, x, i - 1, m, H, s, v, yk);
            orc_y_Ax(tmp, n, Aval, Arow, Acol, x);
            for (e = 0; e < n; ++e) tmp[e] = b[e] - tmp[e];
            pc_apply(&S, r, tmp);
        }
        beta = orc_norm(n, r);
        resid = beta / normb;
        if (resid <= *tol) { *tol = resid; *max_iter = j; conv = 1; goto done; }
    }
    *tol = resid;
done:
    for (i = 0; i < m; ++i) free(v[i]);
    free(v); free(H); free(s); free(cs); free(sn); free(yk);
    free(w); free(r); free(z); free(t); free(tmp);
    return conv;
}

/* ---------------------------------------------------------------------------------------
 * Lanczos / MINRES -- num/krylovsolver.h:926-1161.
 * The reference keeps q, t, b, c, s, r, p in ring buffers (SBufferCL, misc/container.h:201-251):
 * index k of a buffer of size S is slot (Front + (k mod S)) mod S and rotate() advances Front.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    size_t n; const orc_sys* S;
    int precond;            /* 0: LanczosONBCL :945-979, 1: PLanczosONBCL :1010-1052 */
    double* q[3]; int qf;   /* q ring: size 3 (plain) or 2 (preconditioned) */
    double* t[3]; int tf;   /* t ring (preconditioned only), size 3 */
    double bb[2]; int bf;   /* b ring, size 2 */
    double a0, norm_r0; int nobreakdown;
    double* tmp;
} orc_lanczos;

static int ring(int front, int k, int size) { int i = k; if (i < 0) i += size; else if (i >= size) i -= size; i += front; return i < size ? i : i - size; }
#define LQ(L, k) ((L)->q[ring((L)->qf, (k), (L)->precond ? 2 : 3)])
#define LT(L, k) ((L)->t[ring((L)->tf, (k), 3)])
#define LB(L, k) ((L)->bb[ring((L)->bf, (k), 2)])

/* LanczosStep :926-943. */
static int lanczos_step(orc_lanczos* L)
{
    const size_t n = L->n; size_t e;
    const double *q0 = LQ(L, -1), *q1 = LQ(L, 0); double* q2 = LQ(L, 1);
    const double b0 = LB(L, -1);
    double a1, b1;
    orc_y_Ax(L->tmp, n, L->S->val, L->S->row, L->S->col, q1);
    for (e = 0; e < n; ++e) q2[e] = L->tmp[e] - b0 * q0[e];
    a1 = orc_dot(n, q2, q1);
    for (e = 0; e < n; ++e) q2[e] = q2[e] - a1 * q1[e];
    b1 = orc_norm(n, q2);
    L->a0 = a1; LB(L, 0) = b1;
    if (b1 < 1e-15) return 0;
    for (e = 0; e < n; ++e) q2[e] = q2[e] / b1;
    return 1;
}

/* PLanczosStep :989-1008. */
static int planczos_step(orc_lanczos* L)
{
    const size_t n = L->n; size_t e;
    const double *q1 = LQ(L, 0); double* q2 = LQ(L, 1);
    const double *t0 = LT(L, -1), *t1 = LT(L, 0); double* t2 = LT(L, 1);
    const double b0 = LB(L, -1);
    double a1, b1, ib;
    orc_y_Ax(L->tmp, n, L->S->val, L->S->row, L->S->col, q1);
    for (e = 0; e < n; ++e) t2[e] = L->tmp[e] - b0 * t0[e];
    a1 = orc_dot(n, t2, q1);
    for (e = 0; e < n; ++e) t2[e] = t2[e] - a1 * t1[e];
    pc_apply(L->S, q2, t2);
    b1 = sqrt(orc_dot(n, q2, t2));
    L->a0 = a1; LB(L, 0) = b1;
    if (b1 < 1e-15) return 0;
    ib = 1. / b1;
    for (e = 0; e < n; ++e) t2[e] = t2[e] * ib;
    for (e = 0; e < n; ++e) q2[e] = q2[e] * ib;
    return 1;
}

static void lanczos_new_basis(orc_lanczos* L, const double* r0)
{
    const size_t n = L->n; size_t e;
    L->qf = L->tf = L->bf = 0;
    if (!L->precond) {                                     /* :958-967 */
        memset(LQ(L, -1), 0, n * sizeof(double));
        L->norm_r0 = orc_norm(n, r0);
        for (e = 0; e < n; ++e) LQ(L, 0)[e] = r0[e] / L->norm_r0;
        memset(LQ(L, 1), 0, n * sizeof(double));
        LB(L, -1) = 0.;
        L->nobreakdown = lanczos_step(L);
    } else {                                               /* :1030-1040 */
        memset(LT(L, -1), 0, n * sizeof(double));
        pc_apply(L->S, LQ(L, -1), r0);
        L->norm_r0 = sqrt(orc_dot(n, LQ(L, -1), r0));
        for (e = 0; e < n; ++e) LT(L, 0)[e] = r0[e] / L->norm_r0;
        {   /* q[0] = q[-1]/norm; in a ring of size 2, q[1] aliases q[-1] afterwards */
            double* qm1 = LQ(L, -1); double* q0 = LQ(L, 0);
            for (e = 0; e < n; ++e) q0[e] = qm1[e] / L->norm_r0;
        }
        memset(LT(L, 1), 0, n * sizeof(double));
        LB(L, -1) = 0.;
        L->nobreakdown = planczos_step(L);
    }
}

static int lanczos_next(orc_lanczos* L)                    /* :973-978, :1046-1051 */
{
    const int qs = L->precond ? 2 : 3;
    L->qf = (L->qf + 1) % qs;
    if (L->precond) L->tf = (L->tf + 1) % 3;
    L->bf = (L->bf + 1) % 2;
    return (L->nobreakdown = L->precond ? planczos_step(L) : lanczos_step(L));
}

/* PMINRES :1069-1152.  Returns converged; lucky-breakdown handling and the true-residual
 * acceptance test (:1134-1141) as in the reference. */
static int pminres(const orc_sys* S, orc_lanczos* L, double* x, const double* rhs,
                   int* max_iter, double* tol, int relative)
{
    const size_t n = S->n; size_t e;
    double* dx = vec_new(n); double* tmp = vec_new(n);
    double* p[3] = { vec_new(n), vec_new(n), vec_new(n) };
    double c[3] = {0, 0, 0}, s[3] = {0, 0, 0}, r[3][3] = {{0}}, b[2][2] = {{0}};
    int cf = 0, sf = 0, rf = 0, pf = 0, bf = 0;
    const double norm_r0 = L->norm_r0;
    double normb = fabs(norm_r0), res = norm_r0;
    int lucky = !L->nobreakdown, k, conv = 0;
#define C_(k) c[ring(cf, (k), 3)]
#define S_(k) s[ring(sf, (k), 3)]
#define R_(k) r[ring(rf, (k), 3)]
#define P_(k) p[ring(pf, (k), 3)]
#define B_(k) b[ring(bf, (k), 2)]
    if (normb == 0.0 || !relative) normb = 1.0;
    if ((res = norm_r0 / normb) <= *tol) { *tol = res; *max_iter = 0; conv = 1; goto done; }
    for (k = 1; k <= *max_iter; ++k) {
        const double* qq = LQ(L, 0);
        switch (k) {
        case 1:
            gen_rot(&L->a0, &LB(L, 0), &C_(0), &S_(0));
            R_(0)[0] = sqrt(L->a0 * L->a0 + LB(L, 0) * LB(L, 0));
            for (e = 0; e < n; ++e) P_(0)[e] = qq[e] / R_(0)[0];
            B_(0)[0] = 1.; B_(0)[1] = 0.;
            app_rot(&B_(0)[0], &B_(0)[1], &C_(0), &S_(0));
            break;
        case 2:
            R_(0)[0] = LB(L, -1); R_(0)[1] = L->a0; R_(0)[2] = LB(L, 0);
            app_rot(&R_(0)[0], &R_(0)[1], &C_(-1), &S_(-1));
            gen_rot(&R_(0)[1], &R_(0)[2], &C_(0), &S_(0));
            app_rot(&R_(0)[1], &R_(0)[2], &C_(0), &S_(0));
            for (e = 0; e < n; ++e) P_(0)[e] = (qq[e] - R_(0)[0] * P_(-1)[e]) / R_(0)[1];
            B_(0)[0] = B_(-1)[1]; B_(0)[1] = 0.;
            app_rot(&B_(0)[0], &B_(0)[1], &C_(0), &S_(0));
            break;
        default: {
            double tmpb = LB(L, 0), inv;
            R_(0)[0] = 0.; R_(0)[1] = LB(L, -1); R_(0)[2] = L->a0;
            app_rot(&R_(0)[0], &R_(0)[1], &C_(-2), &S_(-2));
            app_rot(&R_(0)[1], &R_(0)[2], &C_(-1), &S_(-1));
            gen_rot(&R_(0)[2], &tmpb, &C_(0), &S_(0));
            app_rot(&R_(0)[2], &tmpb, &C_(0), &S_(0));
            inv = 1 / R_(0)[2];
            for (e = 0; e < n; ++e)
                P_(0)[e] = ((qq[e] - R_(0)[0] * P_(-2)[e]) - R_(0)[1] * P_(-1)[e]) * inv;
            B_(0)[0] = B_(-1)[1]; B_(0)[1] = 0.;
            app_rot(&B_(0)[0], &B_(0)[1], &C_(0), &S_(0));
        } }
        {
            const double f = norm_r0 * B_(0)[0];
            for (e = 0; e < n; ++e) dx[e] = f * P_(0)[e];
            for (e = 0; e < n; ++e) x[e] = x[e] + dx[e];
        }
        res = fabs(norm_r0 * B_(0)[1]) / normb;
        if (res <= *tol || lucky) {
            double realRes;
            orc_y_Ax(tmp, n, S->val, S->row, S->col, x);
            for (e = 0; e < n; ++e) tmp[e] = rhs[e] - tmp[e];
            realRes = orc_norm(n, tmp);
            if (!(realRes > *tol)) { *tol = res; *max_iter = k; conv = 1; goto done; }
        }
        lanczos_next(L);
        if (!L->nobreakdown) lucky = 1;
        cf = (cf + 1) % 3; sf = (sf + 1) % 3; rf = (rf + 1) % 3; pf = (pf + 1) % 3; bf = (bf + 1) % 2;
    }
    *tol = res;
done:
    free(dx); free(tmp); free(p[0]); free(p[1]); free(p[2]);
    return conv;
#undef C_
#undef S_
#undef R_
#undef P_
#undef B_
}

/* MINRES :1155-1161 (precond = 0) and PMResSolverCL<PLanczosONBCL<Vec,PC>>::Solve :1903-1909
 * (precond = 1; pc_kind selects DummyPcCL or JACPcCL). */
ORC_API int orc_minres(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                       int precond, int pc_kind, double omega, double* x, const double* b,
                       int* max_iter, double* tol, int relative)
{
    orc_sys S = { n, Aval, Arow, Acol, pc_kind, omega };
    orc_lanczos L; double* r0 = vec_new(n); int i, conv; size_t e;
    memset(&L, 0, sizeof L);
    L.n = n; L.S = &S; L.precond = precond; L.tmp = vec_new(n);
    for (i = 0; i < 3; ++i) { L.q[i] = vec_new(n); L.t[i] = vec_new(n); }
    orc_y_Ax(r0, n, Aval, Arow, Acol, x);
    for (e = 0; e < n; ++e) r0[e] = b[e] - r0[e];
    lanczos_new_basis(&L, r0);
    conv = pminres(&S, &L, x, b, max_iter, tol, relative);
    for (i = 0; i < 3; ++i) { free(L.q[i]); free(L.t[i]); }
    free(L.tmp); free(r0);
    return conv;
}

/* Lanczos ONB as driven by src/tests/minres.cpp:40-70 (TEST1). */
ORC_API int orc_lanczos_basis(size_t n, const double* Aval, const size_t* Arow, const size_t* Acol,
                              const double* r0, int max_vecs, double* Q, double* a, double* bc)
{
    orc_sys S = { n, Aval, Arow, Acol, 0, 1.0 };
    orc_lanczos L; int i, k = 0, go;
    memset(&L, 0, sizeof L);
    L.n = n; L.S = &S; L.precond = 0; L.tmp = vec_new(n);
    for (i = 0; i < 3; ++i) { L.q[i] = vec_new(n); L.t[i] = vec_new(n); }
    lanczos_new_basis(&L, r0);
    do {
        if (k < max_vecs) { memcpy(Q + (size_t)k * n, LQ(&L, 0), n * sizeof(double)); a[k] = L.a0; bc[k] = LB(&L, 0); }
        ++k;
        go = k < max_vecs && lanczos_next(&L);
    } while (go);
    if (k < max_vecs) { memcpy(Q + (size_t)k * n, LQ(&L, 0), n * sizeof(double)); a[k] = L.a0; bc[k] = LB(&L, 0); ++k; }
    for (i = 0; i < 3; ++i) { free(L.q[i]); free(L.t[i]); }
    free(L.tmp);
    return k;
}

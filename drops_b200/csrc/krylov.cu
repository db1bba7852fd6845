// krylov.cu -- the Krylov drivers of num/krylovsolver.h on the device: PCG, CG, GMRES(m), MINRES.
//
// PCG / CG run without any host round trip per iteration: three kernels per iteration
// (SpMV + fused dot, fused update + 2 reductions, direction update), scalars (alpha, beta,
// residual, stop flag) live in a device-side SolverScalars block, and the host only polls the
// stop flag every "pcg.check_every" iterations, one chunk behind the launch front.
// GMRES and MINRES need their small dense problems (Givens rotations) on the host, so they
// fetch 2..m scalars per iteration; all vector work is fused device kernels.
#include "krylov.cuh"
#include "spmv.cuh"
#include "vecops.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace db200 {

static int ensure_scalars(drops_b200_matrix_s* M, int hist_cap)
{
    if (!M->d_scalars) DB_CHECK(dev_alloc((SolverScalars**)&M->d_scalars, 1));
    if (hist_cap > M->hist_alloc) {
        dev_free(M->d_hist); M->d_hist = nullptr; M->hist_alloc = 0;
        DB_CHECK(dev_alloc(&M->d_hist, (size_t)hist_cap));
        M->hist_alloc = hist_cap;
    }
    return 0;
}

static int ensure_small(drops_b200_matrix_s* M, size_t doubles)
{
    if (doubles > M->small_alloc) {
        dev_free(M->d_small); M->d_small = nullptr; M->small_alloc = 0;
        DB_CHECK(dev_alloc(&M->d_small, doubles));
        M->small_alloc = doubles;
    }
    return 0;
}

static int check_system(drops_b200_matrix_s* M, int pc)
{
    DB_ARG(M != nullptr, "null matrix");
    DB_ARG(M->square, "Krylov solvers need a square matrix");
    DB_ARG(pc == DROPS_B200_PC_DUMMY || pc == DROPS_B200_PC_JACOBI, "unknown preconditioner");
    if (pc == DROPS_B200_PC_JACOBI && !M->diag_ok) {
        set_error("JACPcCL: at least one row stores no diagonal entry (the reference would read past the row, "
                  "num/precond.h:163-164)");
        return DROPS_B200_E_NODIAG;
    }
    return 0;
}

static int fetch(const void* dsrc, void* hdst_pinned_or_not, size_t bytes)
{
    if (bytes % sizeof(double) == 0)         // scalars and Hessenberg columns: published to mapped memory, the host polls
        return fetch_small(static_cast<const double*>(dsrc), static_cast<double*>(hdst_pinned_or_not), bytes / sizeof(double));
    Ctx& c = ctx();
    DB_CUDA(cudaMemcpyAsync(hdst_pinned_or_not, dsrc, bytes, cudaMemcpyDeviceToHost, c.stream));
    DB_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// Row-ordered solves ("pcg.reorder" for PCG, "krylov.reorder" for GMRES / MINRES; operators with a row-ordered
// stream): the whole iteration runs on vectors renumbered by the operator's slice order -- x, b and the diagonal are
// gathered once, x is scattered back at the end.  Every row sum keeps its bits (entry order is untouched); the
// reductions visit the same products in another order, which the double-double accumulation absorbs.  Padding
// positions hold 0 (diagonal 1) and stay 0: a padding row of the twin stream sums to 0.
// ------------------------------------------------------------------------------------------
namespace {
struct RoScope {
    drops_b200_matrix_s* M = nullptr;
    bool on = false;
    size_t nv = 0;                 // length of the vectors the iteration works on
    double* x = nullptr;           // where the iteration reads / writes the solution
    const double* b = nullptr;
    const double* diag = nullptr;
    double* x_user = nullptr;
};

bool ro_available(const drops_b200_matrix_s* M)
{
    return opt("spmv.kernel", 0.0) == 0.0 &&
           ((M->sell.pkp != nullptr && opt("spmv.pk", 1.0) != 0.0) || M->sell.colp != nullptr);
}

int ro_enter(drops_b200_matrix_s* M, bool want, double* x, const double* b, RoScope* R)
{
    R->M = M; R->on = want && ro_available(M);
    R->nv = M->csr.rows; R->x = x; R->b = b; R->diag = M->diag; R->x_user = x;
    if (!R->on) return 0;
    const size_t npos = M->sell.npos;
    // 0 diagonal, 1 x, 2 b (3 r, 4 p, 5 q of PCG)
    if (M->ro_work.size() < 6) M->ro_work.resize(6, nullptr);
    for (size_t k = 0; k < 3; ++k)
        if (!M->ro_work[k]) DB_CHECK(dev_alloc(&M->ro_work[k], npos));
    double* dp = M->ro_work[0]; double* xp = M->ro_work[1]; double* bp = M->ro_work[2];
    if (M->ro_diag_epoch != M->values_epoch) {
        DB_CHECK(vec_gather_perm(npos, M->sell.perm, M->diag, dp, 1.0));
        dev_free(M->ro_diag_idx); M->ro_diag_idx = nullptr; M->ro_diag_count = 0;
        M->ro_diag_idx_epoch = 0;
        M->ro_diag_epoch = M->values_epoch;
    }
    DB_CHECK(vec_gather_perm(npos, M->sell.perm, x, xp, 0.0));
    DB_CHECK(vec_gather_perm(npos, M->sell.perm, b, bp, 0.0));
    R->nv = npos; R->x = xp; R->b = bp; R->diag = dp;
    return 0;
}

int ro_leave(const RoScope& R)
{
    if (!R.on) return 0;
    return vec_scatter_perm(R.M->sell.npos, R.M->sell.perm, R.x, R.x_user);
}
}  // namespace

// ------------------------------------------------------------------------------------------
// PCG -- num/krylovsolver.h:107-179
// ------------------------------------------------------------------------------------------
int pcg_device(drops_b200_matrix_s* M, double* x, const double* b, int pc, double omega, int* max_iter,
               double* tol, int relative, double* hist, int hist_cap, bool unpreconditioned_cg)
{
    DB_CHECK(check_system(M, unpreconditioned_cg ? DROPS_B200_PC_DUMMY : pc));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    const size_t n = M->csr.rows;
    const int want_hist = hist ? std::max(hist_cap, 0) : 0;
    DB_CHECK(ensure_scalars(M, std::max(want_hist, 1)));
    SolverScalars* S = (SolverScalars*)M->d_scalars;
    RoScope R;
    DB_CHECK(ro_enter(M, !unpreconditioned_cg && opt("pcg.reorder", 1.0) != 0.0, x, b, &R));
    const bool ro = R.on;
    if (ro) {
        const size_t npos = M->sell.npos;
        for (size_t k = 3; k < 6; ++k)
            if (!M->ro_work[k]) DB_CHECK(dev_alloc(&M->ro_work[k], npos));
        double* dp = M->ro_work[0];
        // constant-coefficient operators: a handful of distinct diagonal values -> one byte per position.  Built on the
        // first JACOBI solve on these values (a DummyPcCL solve before it must not mark the index as done).
        if (pc == DROPS_B200_PC_JACOBI && M->ro_diag_idx_epoch != M->values_epoch) {
            if (M->sell.ndict_host > 0)
                DB_CHECK(build_value_index8(dp, npos, kDiagDict, &M->ro_diag_idx, M->ro_diag_values, &M->ro_diag_count));
            M->ro_diag_idx_epoch = M->values_epoch;
        }
    }
    x = R.x; b = R.b;
    const double* diag = R.diag;
    const size_t nv = R.nv;                       // length of the vectors the iteration works on
    DiagIndex dix;
    const bool use_dix = ro && M->ro_diag_idx != nullptr && M->ro_diag_count > 0 && opt("pcg.diag_index", 1.0) != 0.0;
    if (use_dix) {
        dix.idx = M->ro_diag_idx;
        std::memcpy(dix.dict.v, M->ro_diag_values, sizeof(double) * (size_t)M->ro_diag_count);
    }
    const DiagIndex* di = use_dix ? &dix : nullptr;
    double* r = ro ? M->ro_work[3] : work_vec(M, 0);
    double* p = ro ? M->ro_work[4] : work_vec(M, 1);
    double* q = ro ? M->ro_work[5] : work_vec(M, 2);
    if (!r || !p || !q) return DROPS_B200_E_NOMEM;

    SolverScalars h;
    std::memset(&h, 0, sizeof h);
    h.tol = *tol; h.relative = relative; h.max_iter = *max_iter; h.pc = pc; h.omega = omega;
    h.hist_cap = want_hist;
    DB_CUDA(cudaMemcpyAsync(S, &h, sizeof h, cudaMemcpyHostToDevice, st));
    double* dh = want_hist ? M->d_hist : nullptr;

    SpmvArgs a;
    if (!unpreconditioned_cg) {
        // r = b - A x                                                           :125
        a.x = x; a.y = r; a.epilogue = EPI_RESID; a.aux = b; a.row_ordered = ro;
        DB_CHECK(spmv_apply(M, false, a));
        DB_CHECK(pcg_setup(nv, r, b, diag, p, S, dh));
    } else {
        // r = A x - b  (CG :61): q = A x, r = q - b
        a.x = x; a.y = q;
        DB_CHECK(spmv_apply(M, false, a));
        DB_CHECK(vec_sub(n, r, q, b));
        DB_CHECK(cg_setup(n, r, b, p, S, dh));
    }

    int check_every = (int)opt("pcg.check_every", 16.0);
    if (check_every < 1) check_every = 1;
    struct Flags { int iter, done, converged, pad; };
    Flags* hflags = reinterpret_cast<Flags*>(c.h_scalars + 16);   // two pinned slots
    cudaEvent_t ev[2];
    DB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    DB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));

    int status = 0, launched = 0, chunk_id = 0;
    bool finished = false;
    const int total = std::max(*max_iter, 0);
    auto enqueue_flag_read = [&](int slot) -> int {
        DB_CUDA(cudaMemcpyAsync(&hflags[slot], &S->iter, sizeof(Flags), cudaMemcpyDeviceToHost, st));
        DB_CUDA(cudaEventRecord(ev[slot], st));
        return 0;
    };
    // "pcg.fused_dot" = 0: q = A p and p.q as two kernels (the dot re-reads p and q from HBM);
    // "pcg.defer_x" = 1: x += alpha p moves into the direction kernel (10 instead of 11 vector passes)
    const bool fused_dot = opt("pcg.fused_dot", 1.0) != 0.0;
    const bool defer_x = !unpreconditioned_cg && opt("pcg.defer_x", 1.0) != 0.0 && opt("vec.wide", 1.0) != 0.0 &&
                         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(r) |
                           reinterpret_cast<uintptr_t>(diag)) & 15u) == 0;
    int iter_no = 0;                                                                // iterations enqueued so far
    auto one_iteration = [&]() -> int {
        SpmvArgs s;
        ++iter_no;
        s.x = p; s.y = q; s.done = &S->done; s.row_ordered = ro;                    // q = A p, lambda = q.p  :149-151
        if (fused_dot) { s.dotv = p; s.dot_out = &S->lambda; }
        DB_CHECK(spmv_apply(M, false, s));
        if (!fused_dot) DB_CHECK(vec_dot_unless(nv, q, p, &S->lambda, &S->done));
        if (defer_x) {
            DB_CHECK(pcg_update_r(nv, r, q, diag, S, dh, nullptr, di));
            DB_CHECK(pcg_direction_x(nv, x, p, r, diag, S, di));
            return 0;
        }
        DB_CHECK(unpreconditioned_cg ? cg_update(n, x, r, p, q, S, dh) : pcg_update(nv, x, r, p, q, diag, S, dh));
        DB_CHECK(unpreconditioned_cg ? cg_direction(n, p, r, S) : pcg_direction(nv, p, r, diag, S));
        return 0;
    };
    // A chunk of iterations as ONE CUDA graph launch: on small systems (cdrdrops' example has 250 k
    // unknowns: ~10 us of HBM time per iteration) the 4 kernel launches per iteration are the cost.
    // All kernel arguments are pointers that stay fixed over the solve, so the graph is captured
    // once per (matrix, x, b, options) and replayed.  Not used while per-kernel timing is on.
    const bool use_graph = opt("pcg.graph", 1.0) != 0.0 && opt("timing", 0.0) == 0.0 && total >= check_every &&
                           dh == nullptr;
    if (use_graph) {
        auto& key = M->pcg_graph_key;
        if (!M->pcg_graph || key.x != x || key.b != b || key.chunk != check_every ||
            key.cg != (int)unpreconditioned_cg + 2 * pc || key.epoch != c.options_epoch || key.values != M->values_epoch) {
            // (a value update rebuilds the packed streams and the dictionary: new device addresses)
            if (M->pcg_graph) { cudaGraphExecDestroy(M->pcg_graph); M->pcg_graph = nullptr; }
            const long long l0 = c.launches;
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (ce != cudaSuccess) status = cuda_fail(ce, "begin capture", __FILE__, __LINE__);
            for (int k = 0; k < check_every && status == 0; ++k) status = one_iteration();
            iter_no = 0;                                       // captured, not run
            ce = cudaStreamEndCapture(st, &graph);
            if (status == 0 && ce != cudaSuccess) status = cuda_fail(ce, "end capture", __FILE__, __LINE__);
            if (status == 0) {
                ce = cudaGraphInstantiate(&M->pcg_graph, graph, 0);
                if (ce != cudaSuccess) status = cuda_fail(ce, "graph instantiate", __FILE__, __LINE__);
            }
            if (graph) cudaGraphDestroy(graph);
            M->pcg_graph_launches = c.launches - l0;
            c.launches = l0;                                   // captured, not launched
            key.x = x; key.b = b; key.chunk = check_every; key.cg = (int)unpreconditioned_cg + 2 * pc;
            key.epoch = c.options_epoch; key.values = M->values_epoch;
            if (status != 0 && M->pcg_graph) { cudaGraphExecDestroy(M->pcg_graph); M->pcg_graph = nullptr; }
        }
    }
    // the setup kernels may already have finished the solve (resid <= tol at iteration 0)
    if (status == 0) status = enqueue_flag_read(0);
    chunk_id = 1;
    while (status == 0 && !finished) {
        const int chunk = std::min(check_every, total - launched);
        if (use_graph && chunk == check_every && M->pcg_graph) {
            cudaError_t ce = cudaGraphLaunch(M->pcg_graph, st);
            if (ce != cudaSuccess) status = cuda_fail(ce, "graph launch", __FILE__, __LINE__);
            c.launches += M->pcg_graph_launches;
            iter_no += check_every;
        } else {
            for (int k = 0; k < chunk && status == 0; ++k) status = one_iteration();
        }
        launched += chunk;
        if (status != 0) break;
        const int slot = chunk_id & 1;
        status = enqueue_flag_read(slot);
        if (status != 0) break;
        // look at the flags of the PREVIOUS chunk while this one runs
        const int prev = slot ^ 1;
        cudaError_t e = cudaEventSynchronize(ev[prev]);
        if (e != cudaSuccess) { status = cuda_fail(e, "event sync", __FILE__, __LINE__); break; }
        if (hflags[prev].done || launched >= total) {
            e = cudaEventSynchronize(ev[slot]);
            if (e != cudaSuccess) { status = cuda_fail(e, "event sync", __FILE__, __LINE__); break; }
            if (hflags[slot].done || launched >= total) finished = true;
        }
        ++chunk_id;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (status != 0) { cudaStreamSynchronize(st); return status; }

    DB_CHECK(ro_leave(R));
    DB_CHECK(fetch(S, &h, sizeof h));
    if (hist && want_hist > 0) {
        const int cnt = std::min(want_hist, h.iter + 1);
        DB_CHECK(fetch(M->d_hist, hist, (size_t)cnt * sizeof(double)));
    }
    *tol = h.resid;
    if (h.converged) *max_iter = h.iter;
    return h.converged ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// GMRES(m) -- num/krylovsolver.h:258-297, :320-461 (serial variant, modified Gram-Schmidt)
// ------------------------------------------------------------------------------------------
static void gen_rot(double& dx, double& dy, double& cs, double& sn)     // :258-271
{
    if (dy == 0.0) { cs = 1.0; sn = 0.0; }
    else { const double r = std::sqrt(dx * dx + dy * dy); cs = dx / r; sn = dy / r; }
}
static void app_rot(double& dx, double& dy, double cs, double sn)       // :274-279
{
    const double tmp = cs * dx + sn * dy;
    dy = -sn * dx + cs * dy;
    dx = tmp;
}

int gmres_device(drops_b200_matrix_s* M, double* x, const double* b, int pc, double omega, int restart,
                 int* max_iter, double* tol, int relative, int left)
{
    DB_CHECK(check_system(M, pc));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    Ctx& c = ctx();
    cudaStream_t st = c.stream;
    // "krylov.reorder" (1): the iteration runs on row-ordered vectors (RoScope above)
    RoScope R;
    DB_CHECK(ro_enter(M, opt("krylov.reorder", 1.0) != 0.0, x, b, &R));
    x = R.x; b = R.b;
    const double* diag = R.diag;
    const bool ro = R.on;
    const size_t n = R.nv;
    int m = restart <= *max_iter ? restart : *max_iter;                 // :327
    DB_ARG(m >= 2, "GMRES needs restart >= 2 and max_iter >= 2");
    const bool jac = (pc == DROPS_B200_PC_JACOBI);
    // "gmres.ortho": 0 = modified Gram-Schmidt, the serial GMRES of the reference (:378-381) and what parity is judged
    // against; 1 = classical Gram-Schmidt, the default of the reference's distributed ModGMRES (:862-866): one fused
    // multi-dot + one fused update per column -- 2 kernels instead of i+1 dependent ones, which is what counts on the
    // small, launch-bound level-set systems (twophasedrops, restart 200)
    const bool cgs = opt("gmres.ortho", 0.0) == 1.0 && m <= 512;

    // work vectors: 0 r, 1 w, 2 tmp, 3 z (right precond), 4.. basis v[0..m-1]
    double* r = work_vec(M, 0); double* w = work_vec(M, 1); double* tmp = work_vec(M, 2); double* z = work_vec(M, 3);
    if (!r || !w || !tmp || !z) return DROPS_B200_E_NOMEM;
    std::vector<double*> v(m);
    for (int i = 0; i < m; ++i) { v[i] = work_vec(M, 4 + i); if (!v[i]) return DROPS_B200_E_NOMEM; }
    // "gmres.device_givens" (classical Gram-Schmidt only): the rotations of a column run in block 0 of the kernel that
    // scales the new basis vector and the stop flag lives on the device, so the host enqueues "gmres.check_every" steps
    // of a cycle without waiting for each Hessenberg column and a step is 4 launches instead of 5 (10^5 rows, GMRES(200):
    // 57.7 -> 54.0 us per step; the same bits, test_gmres_device_side_givens_keeps_every_bit)
    // (restart lengths up to 1024: the column and the rotations sit in shared memory, the rotated columns in an m x m block)
    const bool dev_givens = cgs && m <= 1024 && opt("gmres.device_givens", 1.0) != 0.0;
    // small device block: h[m+1] | y[m] | vptrs[m] | scal[4] | cs[m] | sn[m] | s[m+1] | ctrl (8 doubles) | H[m*m]
    DB_CHECK(ensure_small(M, (size_t)3 * m + 8 + (dev_givens ? (size_t)3 * m + 16 + (size_t)m * m : 0)));
    double* d_h = M->d_small;
    double* d_y = d_h + (m + 1);
    const double** d_vptrs = reinterpret_cast<const double**>(d_y + m);
    double* d_scal = reinterpret_cast<double*>(d_vptrs + m);            // [0] beta, [1] normb
    double* d_cs = d_scal + 4;
    double* d_sn = d_cs + m;
    double* d_s = d_sn + m;
    GmresCtrl* d_ctrl = reinterpret_cast<GmresCtrl*>(d_s + (m + 1) + 1);
    double* d_H = reinterpret_cast<double*>(d_ctrl) + 8;
    DB_CUDA(cudaMemcpyAsync(d_vptrs, v.data(), (size_t)m * sizeof(double*), cudaMemcpyHostToDevice, st));

    std::vector<double> H((size_t)m * m, 0.0), s(m, 0.0), cs(m, 0.0), sn(m, 0.0), y(m, 0.0), hcol(m + 1, 0.0);
    auto Hat = [&](int i, int j) -> double& { return H[(size_t)j * m + i]; };

    auto apply_pc = [&](double* out, const double* in) -> int {         // M.Apply(A, out, in)
        if (jac) return vec_jacobi(n, omega, diag, in, out);
        return vec_copy(n, out, in);
    };
    auto residual = [&](double* out) -> int {                           // out = [M^-1](b - A x)
        SpmvArgs a; a.x = x; a.epilogue = EPI_RESID; a.aux = b; a.row_ordered = ro;
        if (left && jac) { a.y = tmp; DB_CHECK(spmv_apply(M, false, a)); return apply_pc(out, tmp); }
        a.y = out;
        return spmv_apply(M, false, a);
    };
    auto update_x = [&](int k) -> int {                                 // GMRES_Update(x, k, H, s, v) :282-297
        if (k < 0) return 0;
        y = s;
        for (int i = k; i >= 0; --i) {
            y[i] /= Hat(i, i);
            for (int j = i - 1; j >= 0; --j) y[j] -= Hat(j, i) * y[i];
        }
        DB_CUDA(cudaMemcpyAsync(d_y, y.data(), (size_t)(k + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
        if (left) {
            DB_CHECK(multi_axpy(n, x, d_vptrs, d_y, k + 1));
        } else {                                                        // :416-421
            DB_CHECK(vec_fill(n, z, 0.0));
            DB_CHECK(multi_axpy(n, z, d_vptrs, d_y, k + 1));
            DB_CHECK(apply_pc(tmp, z));
            DB_CHECK(vec_axpy(n, 1.0, tmp, x));
        }
        DB_CUDA(cudaStreamSynchronize(st));   // y is reused by the host
        return 0;
    };

    double beta, normb, resid;
    DB_CHECK(residual(r));
    DB_CHECK(mgs_step(n, r, nullptr, nullptr, nullptr, d_scal + 0, true));          // beta = norm(r)
    if (left) {
        DB_CHECK(apply_pc(w, b));
        DB_CHECK(mgs_step(n, w, nullptr, nullptr, nullptr, d_scal + 1, true));      // normb = norm(M^-1 b) :346-347
    } else {
        DB_CHECK(vec_copy(n, w, b));
        DB_CHECK(mgs_step(n, w, nullptr, nullptr, nullptr, d_scal + 1, true));      // normb = norm(b)      :340
    }
    DB_CHECK(fetch(d_scal, c.h_scalars, 2 * sizeof(double)));
    beta = c.h_scalars[0]; normb = c.h_scalars[1];
    if (normb == 0.0 || !relative) normb = 1.0;
    resid = beta / normb;
    if (resid <= *tol) { *tol = resid; *max_iter = 0; return 1; }                   // :352-364 (x untouched)

    int j = 1;
    while (j <= *max_iter) {
        DB_CHECK(vec_scale_inv(n, v[0], r, d_scal + 0));                            // v[0] = r*(1/beta) :366
        std::fill(s.begin(), s.end(), 0.0);
        s[0] = beta;
        int i;
        if (dev_givens) {
            const int steps = std::min(m - 1, *max_iter - j + 1);          // what the host loop below would run
            GmresCtrl hc;
            hc.done = 0; hc.steps = 0; hc.resid = resid; hc.normb = normb; hc.tol = *tol;
            DB_CUDA(cudaMemcpyAsync(d_ctrl, &hc, sizeof hc, cudaMemcpyHostToDevice, st));
            DB_CUDA(cudaMemcpyAsync(d_s, s.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
            const int* d_done = &d_ctrl->done;
            int chunk = (int)opt("gmres.check_every", 8.0);
            if (chunk < 1) chunk = 1;
            double hctl[4] = {0, 0, 0, 0};
            int enq = 0;
            bool stop = false;
            while (enq < steps && !stop) {
                const int upto = std::min(steps, enq + chunk);
                for (int ii = enq; ii < upto; ++ii) {
                    SpmvArgs a;
                    a.row_ordered = ro; a.done = d_done;
                    if (left) {
                        a.x = v[ii]; a.y = w;
                        if (jac) { a.epilogue = EPI_JACOBI; a.aux = diag; a.omega = omega; }
                        DB_CHECK(spmv_apply(M, false, a));
                    } else {
                        if (jac) DB_CHECK(vec_jacobi(n, omega, diag, v[ii], tmp)); else DB_CHECK(vec_copy(n, tmp, v[ii]));
                        a.x = tmp; a.y = w;
                        DB_CHECK(spmv_apply(M, false, a));
                    }
                    DB_CHECK(cgs_dots(n, w, d_vptrs, ii + 1, d_h, nullptr, d_done));
                    DB_CHECK(cgs_update(n, w, d_vptrs, d_h, ii + 1, d_h + ii + 1, nullptr, d_done));
                    DB_CHECK(gmres_scale_givens(n, v[ii + 1], w, ii, m, d_h, d_cs, d_sn, d_s, d_H, d_ctrl));
                }
                enq = upto;
                // (done, steps) as two ints + resid: 3 doubles' worth of the control block
                DB_CHECK(fetch(d_ctrl, hctl, 3 * sizeof(double)));
                GmresCtrl now;
                std::memcpy(&now, hctl, sizeof(int) * 2 + sizeof(double));
                if (now.done) stop = true;
            }
            GmresCtrl fin;
            DB_CHECK(fetch(d_ctrl, hctl, 3 * sizeof(double)));
            std::memcpy(&fin, hctl, sizeof(int) * 2 + sizeof(double));
            const int ran = fin.steps;                                     // columns built in this cycle
            // the rotated columns and the right-hand side come back once per cycle
            if (ran > 0) {
                std::vector<double> Hd((size_t)ran * m);
                DB_CUDA(cudaMemcpyAsync(Hd.data(), d_H, Hd.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
                DB_CUDA(cudaMemcpyAsync(s.data(), d_s, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
                DB_CUDA(cudaStreamSynchronize(st));
                for (int col = 0; col < ran; ++col)
                    for (int k = 0; k <= col + 1 && k < m; ++k) Hat(k, col) = Hd[(size_t)col * m + k];
            }
            resid = fin.resid;
            j += ran;
            if (fin.done) {                                                    // :413-429
                DB_CHECK(update_x(ran - 1));
                DB_CHECK(ro_leave(R));
                *tol = resid; *max_iter = j - 1;
                return 1;
            }
            i = ran;
        } else
        for (i = 0; i < m - 1 && j <= *max_iter; ++i, ++j) {
            SpmvArgs a;
            a.row_ordered = ro;
            if (left) {                                                             // w = M^-1 (A v[i])  :377
                a.x = v[i]; a.y = w;
                if (jac) { a.epilogue = EPI_JACOBI; a.aux = diag; a.omega = omega; }
                DB_CHECK(spmv_apply(M, false, a));
            } else {                                                                // w = A (M^-1 v[i])  :372-376
                DB_CHECK(apply_pc(tmp, v[i]));
                a.x = tmp; a.y = w;
                DB_CHECK(spmv_apply(M, false, a));
            }
            if (cgs) {
                // classical Gram-Schmidt: all products against the unmodified w, then one update  :465-494, :862-866
                DB_CHECK(cgs_dots(n, w, d_vptrs, i + 1, d_h));
                DB_CHECK(cgs_update(n, w, d_vptrs, d_h, i + 1, d_h + i + 1));
            } else {
                // modified Gram-Schmidt: axpy of step k-1 fused with the dot of step k      :378-381
                for (int k = 0; k <= i; ++k)
                    DB_CHECK(mgs_step(n, w, k > 0 ? v[k - 1] : nullptr, d_h + (k - 1), v[k], d_h + k, false));
                DB_CHECK(mgs_step(n, w, v[i], d_h + i, nullptr, d_h + i + 1, true));    // H(i+1,i) = norm(w) :383
            }
            DB_CHECK(vec_scale_inv(n, v[i + 1], w, d_h + i + 1));                   // v[i+1] = w*(1/H(i+1,i))
            DB_CHECK(fetch(d_h, hcol.data(), (size_t)(i + 2) * sizeof(double)));
            for (int k = 0; k <= i + 1; ++k) Hat(k, i) = hcol[k];
            for (int k = 0; k < i; ++k) app_rot(Hat(k, i), Hat(k + 1, i), cs[k], sn[k]);
            gen_rot(Hat(i, i), Hat(i + 1, i), cs[i], sn[i]);
            app_rot(Hat(i, i), Hat(i + 1, i), cs[i], sn[i]);
            app_rot(s[i], s[i + 1], cs[i], sn[i]);
            resid = std::abs(s[i + 1]) / normb;                                     // :393
            if (resid <= *tol) {                                                    // :413-429
                DB_CHECK(update_x(i));
                DB_CHECK(ro_leave(R));
                *tol = resid; *max_iter = j;
                return 1;
            }
        }
        DB_CHECK(update_x(i - 1));                                                  // :431-444
        DB_CHECK(residual(r));
        DB_CHECK(mgs_step(n, r, nullptr, nullptr, nullptr, d_scal + 0, true));
        DB_CHECK(fetch(d_scal, c.h_scalars, sizeof(double)));
        beta = c.h_scalars[0];
        resid = beta / normb;
        if (resid <= *tol) { DB_CHECK(ro_leave(R)); *tol = resid; *max_iter = j; return 1; }   // :449-456
    }
    DB_CHECK(ro_leave(R));
    *tol = resid;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Lanczos / MINRES -- num/krylovsolver.h:926-1161
// ------------------------------------------------------------------------------------------
namespace {

struct Ring3 {   // SBufferCL<T,3> index arithmetic, misc/container.h:226-247
    int front = 0;
    int at(int k) const { int i = k; if (i < 0) i += 3; else if (i >= 3) i -= 3; i += front; return i < 3 ? i : i - 3; }
    void rotate() { front = (front + 1) % 3; }
};
struct Ring2 {
    int front = 0;
    int at(int k) const { int i = k; if (i < 0) i += 2; else if (i >= 2) i -= 2; i += front; return i < 2 ? i : i - 2; }
    void rotate() { front = (front + 1) % 2; }
};

struct LanczosDev {
    drops_b200_matrix_s* M;
    size_t n;
    bool precond; int pc; double omega;
    bool ro = false; const double* diag = nullptr;   // row-ordered solve (RoScope): every product on the twin stream, its diagonal
    double* qv[3]; Ring3 q3; Ring2 q2r;      // q ring: 3 slots (plain) / 2 slots (preconditioned)
    double* tv[3]; Ring3 t3;
    double bb[2]; Ring2 br;
    double a0 = 0.0, norm_r0 = 0.0;
    bool nobreakdown = true;
    double* d_scal;                          // device: [0] a1, [1] b1 / b1sq

    double* Q(int k) { return precond ? qv[q2r.at(k)] : qv[q3.at(k)]; }
    double* T(int k) { return tv[t3.at(k)]; }
    double& B(int k) { return bb[br.at(k)]; }

    int step()
    {
        Ctx& c = ctx();
        if (!precond) {                                                     // LanczosStep :926-943
            SpmvArgs a;
            a.row_ordered = ro;
            a.x = Q(0); a.y = Q(1); a.epilogue = EPI_AXMBY; a.aux = Q(-1); a.beta_val = B(-1);
            a.dotv = Q(0); a.dot_out = d_scal + 0;                          // q2 = A q1 - b0 q0; a1 = q2.q1
            DB_CHECK(spmv_apply(M, false, a));
            DB_CHECK(lanczos_orth(n, Q(1), Q(0), d_scal + 0, d_scal + 1));  // q2 -= a1 q1; b1 = |q2|
            DB_CHECK(fetch(d_scal, c.h_scalars, 2 * sizeof(double)));
            a0 = c.h_scalars[0];
            const double b1 = c.h_scalars[1];
            B(0) = b1;
            if (b1 < 1e-15) { nobreakdown = false; return 0; }
            DB_CHECK(vec_div_scalar(n, Q(1), Q(1), b1, 0));                 // q2 /= b1
        } else {                                                            // PLanczosStep :989-1008
            SpmvArgs a;
            a.row_ordered = ro;
            a.x = Q(0); a.y = T(1); a.epilogue = EPI_AXMBY; a.aux = T(-1); a.beta_val = B(-1);
            a.dotv = Q(0); a.dot_out = d_scal + 0;                          // t2 = A q1 - b0 t0; a1 = t2.q1
            DB_CHECK(spmv_apply(M, false, a));
            DB_CHECK(planczos_orth(n, T(1), T(0), d_scal + 0, Q(1), pc, omega, diag, d_scal + 1));
            DB_CHECK(fetch(d_scal, c.h_scalars, 2 * sizeof(double)));
            a0 = c.h_scalars[0];
            const double b1 = std::sqrt(c.h_scalars[1]);
            B(0) = b1;
            if (b1 < 1e-15) { nobreakdown = false; return 0; }
            DB_CHECK(vec_div_scalar(n, T(1), T(1), b1, 1));                 // t2 *= 1/b1
            DB_CHECK(vec_div_scalar(n, Q(1), Q(1), b1, 1));                 // q2 *= 1/b1
        }
        nobreakdown = true;
        return 0;
    }

    int new_basis(const double* r0)
    {
        Ctx& c = ctx();
        q3 = Ring3(); q2r = Ring2(); t3 = Ring3(); br = Ring2();
        if (!precond) {                                                     // :958-967
            DB_CHECK(vec_fill(n, Q(-1), 0.0));
            DB_CHECK(vec_dot(n, r0, r0, d_scal + 0));
            DB_CHECK(fetch(d_scal, c.h_scalars, sizeof(double)));
            norm_r0 = std::sqrt(c.h_scalars[0]);
            DB_CHECK(vec_div_scalar(n, Q(0), r0, norm_r0, 0));
            DB_CHECK(vec_fill(n, Q(1), 0.0));
        } else {                                                            // :1030-1040
            DB_CHECK(vec_fill(n, T(-1), 0.0));
            if (pc == DROPS_B200_PC_JACOBI) DB_CHECK(vec_jacobi(n, omega, diag, r0, Q(-1)));
            else DB_CHECK(vec_copy(n, Q(-1), r0));
            DB_CHECK(vec_dot(n, Q(-1), r0, d_scal + 0));
            DB_CHECK(fetch(d_scal, c.h_scalars, sizeof(double)));
            norm_r0 = std::sqrt(c.h_scalars[0]);
            DB_CHECK(vec_div_scalar(n, T(0), r0, norm_r0, 0));
            DB_CHECK(vec_div_scalar(n, Q(0), Q(-1), norm_r0, 0));
            DB_CHECK(vec_fill(n, T(1), 0.0));
        }
        B(-1) = 0.;
        return step();
    }

    int next()                                                              // :973-978 / :1046-1051
    {
        if (precond) { q2r.rotate(); t3.rotate(); } else q3.rotate();
        br.rotate();
        return step();
    }
};

}  // namespace

int minres_device(drops_b200_matrix_s* M, double* x, const double* b, int precond, int pc, double omega,
                  int* max_iter, double* tol, int relative)
{
    DB_CHECK(check_system(M, precond ? pc : DROPS_B200_PC_DUMMY));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    Ctx& c = ctx();
    RoScope R;                                                              // "krylov.reorder"
    DB_CHECK(ro_enter(M, opt("krylov.reorder", 1.0) != 0.0, x, b, &R));
    x = R.x; b = R.b;
    const size_t n = R.nv;
    // work: 0 tmp/r0, 1-3 q, 4-6 t, 7-9 p
    double* tmp = work_vec(M, 0);
    LanczosDev L;
    L.M = M; L.n = n; L.precond = precond != 0; L.pc = pc; L.omega = omega; L.ro = R.on; L.diag = R.diag;
    for (int i = 0; i < 3; ++i) {
        L.qv[i] = work_vec(M, 1 + i);
        L.tv[i] = work_vec(M, 4 + i);
        if (!L.qv[i] || !L.tv[i]) return DROPS_B200_E_NOMEM;
    }
    double* pv[3];
    for (int i = 0; i < 3; ++i) { pv[i] = work_vec(M, 7 + i); if (!pv[i]) return DROPS_B200_E_NOMEM; }
    if (!tmp) return DROPS_B200_E_NOMEM;
    DB_CHECK(ensure_small(M, 8));
    L.d_scal = M->d_small;

    {   // r0 = b - A x                                                     :1158
        SpmvArgs a; a.x = x; a.y = tmp; a.epilogue = EPI_RESID; a.aux = b; a.row_ordered = R.on;
        DB_CHECK(spmv_apply(M, false, a));
    }
    DB_CHECK(L.new_basis(tmp));

    // PMINRES :1069-1152
    const double norm_r0 = L.norm_r0;
    double normb = std::fabs(norm_r0), res = norm_r0;
    bool lucky = !L.nobreakdown;
    double cc[3] = {0, 0, 0}, ss[3] = {0, 0, 0}, rr[3][3] = {{0}}, bq[2][2] = {{0}};
    Ring3 cr, sr, rrg, pr; Ring2 bqr;
    if (normb == 0.0 || !relative) normb = 1.0;
    if ((res = norm_r0 / normb) <= *tol) { *tol = res; *max_iter = 0; return 1; }
    for (int k = 1; k <= *max_iter; ++k) {
        double* R0 = rr[rrg.at(0)];
        double* B0 = bq[bqr.at(0)];
        const double* Bm1 = bq[bqr.at(-1)];
        int mode;
        switch (k) {
        case 1:
            gen_rot(L.a0, L.B(0), cc[cr.at(0)], ss[sr.at(0)]);
            R0[0] = std::sqrt(L.a0 * L.a0 + L.B(0) * L.B(0));
            B0[0] = 1.; B0[1] = 0.;
            app_rot(B0[0], B0[1], cc[cr.at(0)], ss[sr.at(0)]);
            mode = 1;
            break;
        case 2:
            R0[0] = L.B(-1); R0[1] = L.a0; R0[2] = L.B(0);
            app_rot(R0[0], R0[1], cc[cr.at(-1)], ss[sr.at(-1)]);
            gen_rot(R0[1], R0[2], cc[cr.at(0)], ss[sr.at(0)]);
            app_rot(R0[1], R0[2], cc[cr.at(0)], ss[sr.at(0)]);
            B0[0] = Bm1[1]; B0[1] = 0.;
            app_rot(B0[0], B0[1], cc[cr.at(0)], ss[sr.at(0)]);
            mode = 2;
            break;
        default: {
            R0[0] = 0.; R0[1] = L.B(-1); R0[2] = L.a0;
            double tmpb = L.B(0);
            app_rot(R0[0], R0[1], cc[cr.at(-2)], ss[sr.at(-2)]);
            app_rot(R0[1], R0[2], cc[cr.at(-1)], ss[sr.at(-1)]);
            gen_rot(R0[2], tmpb, cc[cr.at(0)], ss[sr.at(0)]);
            app_rot(R0[2], tmpb, cc[cr.at(0)], ss[sr.at(0)]);
            B0[0] = Bm1[1]; B0[1] = 0.;
            app_rot(B0[0], B0[1], cc[cr.at(0)], ss[sr.at(0)]);
            mode = 3;
        } }
        // p[0] = ...; dx = norm_r0*b[0][0]*p[0]; x += dx                    :1095-1129
        DB_CHECK(minres_update(n, mode, L.Q(0), pv[pr.at(-2)], pv[pr.at(-1)], pv[pr.at(0)], x,
                               R0[0], R0[1], R0[2], norm_r0 * B0[0]));
        res = std::fabs(norm_r0 * B0[1]) / normb;
        if (res <= *tol || lucky) {                                         // :1132-1141
            SpmvArgs a; a.x = x; a.y = tmp; a.epilogue = EPI_RESID; a.aux = b; a.row_ordered = R.on;
            DB_CHECK(spmv_apply(M, false, a));
            DB_CHECK(vec_dot(n, tmp, tmp, L.d_scal + 2));
            DB_CHECK(fetch(L.d_scal + 2, c.h_scalars, sizeof(double)));
            const double realRes = std::sqrt(c.h_scalars[0]);
            if (!(realRes > *tol)) { DB_CHECK(ro_leave(R)); *tol = res; *max_iter = k; return 1; }
        }
        DB_CHECK(L.next());
        if (!L.nobreakdown) lucky = true;
        cr.rotate(); sr.rotate(); rrg.rotate(); pr.rotate(); bqr.rotate();
    }
    DB_CHECK(ro_leave(R));
    *tol = res;
    return 0;
}

// ------------------------------------------------------------------------------------------
// BiCGStab -- num/krylovsolver.h:1189-1258 (BiCGStabSolverCL :1978-2013, serial branch)
// Returns 1 converged, 0 not converged or breakdown (the reference returns false for both and
// leaves a message on std::cout); *breakdown tells them apart.
// ------------------------------------------------------------------------------------------
int bicgstab_device(drops_b200_matrix_s* M, double* x, const double* b, int pc, double omega_pc, int* max_iter,
                    double* tol, int relative, int* breakdown)
{
    DB_CHECK(check_system(M, pc));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    Ctx& c = ctx();
    const size_t n = M->csr.rows;
    if (breakdown) *breakdown = 0;
    double* w[8];
    for (int i = 0; i < 8; ++i) { w[i] = work_vec(M, i); if (!w[i]) return DROPS_B200_E_NOMEM; }
    double *r = w[0], *rtilde = w[1], *p = w[2], *phat = w[3], *sv = w[4], *shat = w[5], *t = w[6], *v = w[7];
    DB_CHECK(ensure_small(M, 8));
    double* d_s = M->d_small;
    double* hs = c.h_scalars;
    SpmvArgs a;
    a.x = x; a.y = r; a.epilogue = EPI_RESID; a.aux = b;                       // r = b - A x          :1197
    DB_CHECK(spmv_apply(M, false, a));
    DB_CHECK(vec_copy(n, rtilde, r));                                          // rtilde = r           :1198
    DB_CHECK(vec_dot(n, b, b, d_s + 0));
    DB_CHECK(vec_dot(n, r, r, d_s + 1));
    DB_CHECK(vec_dot(n, rtilde, r, d_s + 2));
    DB_CHECK(fetch(d_s, hs, 3 * sizeof(double)));
    double normb = std::sqrt(hs[0]);
    if (normb == 0.0 || !relative) normb = 1.0;                                //                      :1200
    double resid = std::sqrt(hs[1]) / normb, rr = hs[1];
    if (resid <= *tol) { *tol = resid; *max_iter = 0; return 1; }
    double rho_1 = hs[2], rho_2 = 0.0, alpha = 0.0, beta = 0.0, omega = 0.0;
    for (int i = 1; i <= *max_iter; ++i) {
        if (rho_1 == 0.0) {                                                    //                      :1210-1215
            *tol = std::sqrt(rr) / normb; *max_iter = i;
            if (breakdown) *breakdown = 1;
            return 0;
        }
        if (i > 1) beta = (rho_1 / rho_2) * (alpha / omega);                   //                      :1219
        DB_CHECK(bicg_p(n, i == 1, p, phat, r, v, beta, omega, pc, omega_pc, M->diag));
        SpmvArgs av; av.x = phat; av.y = v; av.dotv = rtilde; av.dot_out = d_s + 0;   // v = A phat; rtilde.v  :1223-1224
        DB_CHECK(spmv_apply(M, false, av));
        DB_CHECK(fetch(d_s, hs, sizeof(double)));
        alpha = rho_1 / hs[0];
        DB_CHECK(bicg_s(n, sv, shat, r, v, alpha, pc, omega_pc, M->diag, d_s + 0));
        DB_CHECK(fetch(d_s, hs, sizeof(double)));
        if ((resid = std::sqrt(hs[0]) / normb) < *tol) {                       //                      :1226-1231
            DB_CHECK(vec_axpy(n, alpha, phat, x));
            *tol = resid; *max_iter = i;
            return 1;
        }
        SpmvArgs at; at.x = shat; at.y = t;                                    // t = A shat           :1233
        DB_CHECK(spmv_apply(M, false, at));
        DB_CHECK(bicg_ts(n, t, sv, d_s + 0));
        DB_CHECK(fetch(d_s, hs, 2 * sizeof(double)));
        omega = hs[0] / hs[1];                                                 //                      :1234
        DB_CHECK(bicg_x(n, x, r, phat, shat, sv, t, rtilde, alpha, omega, d_s + 0));
        DB_CHECK(fetch(d_s, hs, 2 * sizeof(double)));
        rr = hs[0];
        rho_2 = rho_1;
        rho_1 = hs[1];                                                         // next iteration's dot(rtilde, r)
        if ((resid = std::sqrt(rr) / normb) < *tol) { *tol = resid; *max_iter = i; return 1; }
        if (omega == 0.0) {                                                    //                      :1245-1250
            *tol = std::sqrt(rr) / normb; *max_iter = i;
            if (breakdown) *breakdown = 2;
            return 0;
        }
    }
    *tol = resid;
    return 0;
}

// ------------------------------------------------------------------------------------------
// GCR(m) -- num/krylovsolver.h:1410-1483 (GCRSolverCL :2017-2053) with DummyExchangeCL.
// The reference keeps "accumulated" twins racc / vnacc / vacc[i] of r / vn / v[i]; serially they
// are copies until the truncation branch (:1466-1477) replaces s[min] and v[min] but NOT vacc[min].
// To give the reference's results the twins are reproduced -- lazily: they alias the originals
// until the first truncation.
// ------------------------------------------------------------------------------------------
int gcr_device(drops_b200_matrix_s* M, double* x, const double* b, int pc, double omega_pc, int m, int* max_iter,
               double* tol, int relative)
{
    DB_CHECK(check_system(M, pc));
    DB_ARG(max_iter && tol, "null max_iter/tol");
    DB_ARG(m >= 1, "GCR: truncation parameter m must be >= 1");
    Ctx& c = ctx();
    const size_t n = M->csr.rows;
    m = (m <= *max_iter) ? m : *max_iter;                                      //                      :1416
    if (m < 1) m = 1;
    // work: 0 r, 1 sn, 2 vn, 3 racc, 4 vnacc, 5.. s[i], v[i], vacc_own[i] interleaved
    double* r = work_vec(M, 0); double* sn = work_vec(M, 1); double* vn = work_vec(M, 2);
    if (!r || !sn || !vn) return DROPS_B200_E_NOMEM;
    double *racc = nullptr, *vnacc = nullptr;                                  // twins: nullptr = alias
    std::vector<double*> s((size_t)m, nullptr), v((size_t)m, nullptr), vacc_own((size_t)m, nullptr);
    std::vector<double> a((size_t)m, 0.0);
    DB_CHECK(ensure_small(M, (size_t)m + 8));
    double* d_a = M->d_small;                                                  // a[m] | beta | gamma | rr | bb
    double* d_beta = d_a + m; double* d_gamma = d_beta + 1; double* d_rr = d_gamma + 1; double* d_bb = d_rr + 1;
    std::vector<double> h((size_t)m + 4);
    SpmvArgs ar;
    ar.x = x; ar.y = r; ar.epilogue = EPI_RESID; ar.aux = b;                   // r = b - A x          :1418
    DB_CHECK(spmv_apply(M, false, ar));
    DB_CHECK(vec_dot(n, b, b, d_bb));
    DB_CHECK(vec_dot(n, r, r, d_rr));
    DB_CHECK(fetch(d_rr, h.data(), 2 * sizeof(double)));
    double normb = std::sqrt(h[1]);
    if (normb == 0.0 || !relative) normb = 1.0;
    double resid = std::sqrt(h[0]) / normb;
    int stored = 0;
    for (int k = 0; k < *max_iter; ++k) {
        if (resid < *tol) { *tol = resid; *max_iter = k; return 1; }           //                      :1433-1437
        if (pc == DROPS_B200_PC_JACOBI) DB_CHECK(vec_jacobi(n, omega_pc, M->diag, r, sn));
        else DB_CHECK(vec_copy(n, sn, r));                                     // M.Apply(A, sn, r)    :1438
        SpmvArgs av; av.x = sn; av.y = vn;                                     // vn = A sn            :1440
        DB_CHECK(spmv_apply(M, false, av));
        if (vnacc) DB_CHECK(vec_copy(n, vnacc, vn));                           // vnacc = GetAccumulate(vn)
        const int nst = std::min(k, m);
        auto vacc = [&](int i) -> const double* { return vacc_own[(size_t)i] ? vacc_own[(size_t)i] : v[(size_t)i]; };
        for (int i = 0; i <= nst; ++i) {                                       // orthogonalise        :1442-1451
            const bool prev = i > 0;
            DB_CHECK(gcr_orth(n, vn, sn, vnacc, prev ? v[(size_t)i - 1] : nullptr, prev ? s[(size_t)i - 1] : nullptr,
                              prev ? vacc(i - 1) : nullptr, prev ? d_a + (i - 1) : nullptr,
                              i < nst ? vacc(i) : nullptr, i < nst ? d_a + i : d_beta));
        }
        DB_CHECK(gcr_scale(n, vn, sn, vnacc, racc ? racc : r, d_beta, d_gamma));      //               :1452-1455
        DB_CHECK(gcr_update(n, x, r, racc, sn, vn, vnacc, d_gamma, d_rr));            //               :1456-1462
        DB_CHECK(fetch(d_a, h.data(), ((size_t)m + 3) * sizeof(double)));
        for (int i = 0; i < nst; ++i) a[(size_t)i] = h[(size_t)i];
        resid = std::sqrt(h[(size_t)m + 2]) / normb;
        if (k < m) {                                                           // keep the direction   :1463-1467
            s[(size_t)stored] = work_vec(M, 5 + 3 * stored);
            v[(size_t)stored] = work_vec(M, 6 + 3 * stored);
            if (!s[(size_t)stored] || !v[(size_t)stored]) return DROPS_B200_E_NOMEM;
            DB_CHECK(vec_copy(n, s[(size_t)stored], sn));
            DB_CHECK(vec_copy(n, v[(size_t)stored], vn));
            ++stored;
        } else {                                                               // min-alpha truncation :1468-1480
            int min_idx = 0;
            double a_min = std::fabs(a[0]);
            for (int i = 1; i < k && i < m; ++i)
                if (std::fabs(a[(size_t)i]) < a_min) { min_idx = i; a_min = std::fabs(a[(size_t)i]); }
            if (!racc) {                                                       // the twins part ways here
                racc = work_vec(M, 3); vnacc = work_vec(M, 4);
                if (!racc || !vnacc) return DROPS_B200_E_NOMEM;
                DB_CHECK(vec_copy(n, racc, r));
                DB_CHECK(vec_copy(n, vnacc, vn));
            }
            if (!vacc_own[(size_t)min_idx]) {                                  // vacc[min_idx] keeps the OLD vector
                vacc_own[(size_t)min_idx] = work_vec(M, 7 + 3 * min_idx);
                if (!vacc_own[(size_t)min_idx]) return DROPS_B200_E_NOMEM;
                DB_CHECK(vec_copy(n, vacc_own[(size_t)min_idx], v[(size_t)min_idx]));
            }
            DB_CHECK(vec_copy(n, s[(size_t)min_idx], sn));
            DB_CHECK(vec_copy(n, v[(size_t)min_idx], vn));
        }
    }
    *tol = resid;
    return 0;
}

}  // namespace db200

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace db200;

namespace {

// Temporary device copies of host vectors for the host-buffer entry points.
struct HostBridge {
    // x/b device copies cached with the matrix handle: no cudaMalloc/cudaFree per Solve
    drops_b200_matrix_s* M;
    double* dx = nullptr; double* db = nullptr;
    explicit HostBridge(drops_b200_matrix_s* m) : M(m) {}
    int up(size_t n, const double* x, const double* b)
    {
        cudaStream_t st = ctx().stream;
        for (int i = 0; i < 2; ++i)
            if (!M->bridge[i]) DB_CHECK(dev_alloc(&M->bridge[i], std::max(M->csr.rows, M->csr.cols)));
        dx = M->bridge[0]; db = M->bridge[1];
        if (n && x) DB_CUDA(cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, st));
        if (n && b) DB_CUDA(cudaMemcpyAsync(db, b, n * sizeof(double), cudaMemcpyHostToDevice, st));
        return 0;
    }
    int down(size_t n, double* x)
    {
        cudaStream_t st = ctx().stream;
        if (n) DB_CUDA(cudaMemcpyAsync(x, dx, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        DB_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
};

}  // namespace

extern "C" {

int drops_b200_spmv_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y, "null argument");
    DB_ARG(x->n == A->csr.cols && y->n == A->csr.rows, "SparseMatBaseCL * VectorBaseCL: incompatible dimensions");
    SpmvArgs a; a.x = x->d; a.y = y->d;
    return spmv_apply(A, false, a);
}

int drops_b200_spmv_dot_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t y,
                            drops_b200_vector_t w, double* result)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y && w && result, "null argument");
    DB_ARG(x->n == A->csr.cols && y->n == A->csr.rows && w->n == A->csr.rows,
           "SparseMatBaseCL * VectorBaseCL: incompatible dimensions");
    Ctx& c = ctx();
    SpmvArgs a; a.x = x->d; a.y = y->d; a.dotv = w->d; a.dot_out = c.scratch_scalars;
    DB_CHECK(spmv_apply(A, false, a));
    return fetch(c.scratch_scalars, result, sizeof(double));
}

int drops_b200_spmv_transpose_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y, "null argument");
    DB_ARG(x->n == A->csr.rows && y->n == A->csr.cols, "transp_mul: incompatible dimensions");
    DB_CHECK(build_transpose(A));
    SpmvArgs a; a.x = x->d; a.y = y->d;
    return spmv_apply(A, true, a);
}

int drops_b200_spmv(drops_b200_matrix_t A, const double* x, double* y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y, "null argument");
    cudaStream_t st = ctx().stream;
    double *dx = nullptr, *dy = nullptr;
    DB_CHECK(dev_alloc(&dx, A->csr.cols));
    int s = dev_alloc(&dy, A->csr.rows);
    if (s == 0 && A->csr.cols) {
        cudaError_t e = cudaMemcpyAsync(dx, x, A->csr.cols * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) s = cuda_fail(e, "H2D x", __FILE__, __LINE__);
    }
    if (s == 0) { SpmvArgs a; a.x = dx; a.y = dy; s = spmv_apply(A, false, a); }
    if (s == 0 && A->csr.rows) {
        cudaError_t e = cudaMemcpyAsync(y, dy, A->csr.rows * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) s = cuda_fail(e, "D2H y", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(st);
    dev_free(dx); dev_free(dy);
    return s;
}

int drops_b200_spmv_transpose(drops_b200_matrix_t A, const double* x, double* y)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && y, "null argument");
    DB_CHECK(build_transpose(A));
    cudaStream_t st = ctx().stream;
    double *dx = nullptr, *dy = nullptr;
    DB_CHECK(dev_alloc(&dx, A->csr.rows));
    int s = dev_alloc(&dy, A->csr.cols);
    if (s == 0 && A->csr.rows) {
        cudaError_t e = cudaMemcpyAsync(dx, x, A->csr.rows * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) s = cuda_fail(e, "H2D x", __FILE__, __LINE__);
    }
    if (s == 0) { SpmvArgs a; a.x = dx; a.y = dy; s = spmv_apply(A, true, a); }
    if (s == 0 && A->csr.cols) {
        cudaError_t e = cudaMemcpyAsync(y, dy, A->csr.cols * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) s = cuda_fail(e, "D2H y", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(st);
    dev_free(dx); dev_free(dy);
    return s;
}

int drops_b200_jacobi_apply(drops_b200_matrix_t A, double omega, drops_b200_vector_t b, drops_b200_vector_t x)
{
    DB_CHECK(require_ctx());
    DB_CHECK(check_system(A, DROPS_B200_PC_JACOBI));
    DB_ARG(b && x && b->n == A->csr.rows && x->n == A->csr.rows, "JACPcCL::Apply: incompatible dimensions");
    return vec_jacobi(A->csr.rows, omega, A->diag, b->d, x->d);
}

int drops_b200_pcg_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc, double omega,
                       int* max_iter, double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "PCG: incompatible dimensions");
    if (pc == DROPS_B200_PC_CHEBYSHEV || pc == DROPS_B200_PC_DIAG)      // ChebyshevPcCL is a procedure, DiagPcCL carries its own vector
        return pcg_general_device(A, false, x->d, b->d, pc, omega, max_iter, tol, relative, hist, hist_cap);
    return pcg_device(A, x->d, b->d, pc, omega, max_iter, tol, relative, hist, hist_cap, false);
}

int drops_b200_pcg(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega, int* max_iter,
                   double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    HostBridge hb(A);
    DB_CHECK(hb.up(A->csr.rows, x, b));
    const int r = (pc == DROPS_B200_PC_CHEBYSHEV || pc == DROPS_B200_PC_DIAG)
        ? pcg_general_device(A, false, hb.dx, hb.db, pc, omega, max_iter, tol, relative, hist, hist_cap)
        : pcg_device(A, hb.dx, hb.db, pc, omega, max_iter, tol, relative, hist, hist_cap, false);
    if (r < 0) return r;
    DB_CHECK(hb.down(A->csr.rows, x));
    return r;
}

int drops_b200_cg_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int* max_iter,
                      double* tol, int relative, double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "CG: incompatible dimensions");
    return pcg_device(A, x->d, b->d, DROPS_B200_PC_DUMMY, 1.0, max_iter, tol, relative, hist, hist_cap, true);
}

int drops_b200_cg(drops_b200_matrix_t A, double* x, const double* b, int* max_iter, double* tol, int relative,
                  double* hist, int hist_cap)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    HostBridge hb(A);
    DB_CHECK(hb.up(A->csr.rows, x, b));
    const int r = pcg_device(A, hb.dx, hb.db, DROPS_B200_PC_DUMMY, 1.0, max_iter, tol, relative, hist, hist_cap, true);
    if (r < 0) return r;
    DB_CHECK(hb.down(A->csr.rows, x));
    return r;
}

int drops_b200_gmres_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc, double omega,
                         int restart, int* max_iter, double* tol, int relative, int left_precond)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "GMRES: incompatible dimensions");
    return gmres_device(A, x->d, b->d, pc, omega, restart, max_iter, tol, relative, left_precond);
}

int drops_b200_gmres(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega, int restart,
                     int* max_iter, double* tol, int relative, int left_precond)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    HostBridge hb(A);
    DB_CHECK(hb.up(A->csr.rows, x, b));
    const int r = gmres_device(A, hb.dx, hb.db, pc, omega, restart, max_iter, tol, relative, left_precond);
    if (r < 0) return r;
    DB_CHECK(hb.down(A->csr.rows, x));
    return r;
}

int drops_b200_minres_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int precond, int pc,
                          double omega, int* max_iter, double* tol, int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "MINRES: incompatible dimensions");
    return minres_device(A, x->d, b->d, precond, pc, omega, max_iter, tol, relative);
}

int drops_b200_minres(drops_b200_matrix_t A, double* x, const double* b, int precond, int pc, double omega,
                      int* max_iter, double* tol, int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    HostBridge hb(A);
    DB_CHECK(hb.up(A->csr.rows, x, b));
    const int r = minres_device(A, hb.dx, hb.db, precond, pc, omega, max_iter, tol, relative);
    if (r < 0) return r;
    DB_CHECK(hb.down(A->csr.rows, x));
    return r;
}

int drops_b200_bicgstab_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc, double omega,
                            int* max_iter, double* tol, int relative, int* breakdown)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "BICGSTAB: incompatible dimensions");
    return bicgstab_device(A, x->d, b->d, pc, omega, max_iter, tol, relative, breakdown);
}

int drops_b200_bicgstab(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega, int* max_iter,
                        double* tol, int relative, int* breakdown)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    HostBridge hb(A);
    DB_CHECK(hb.up(A->csr.rows, x, b));
    const int r = bicgstab_device(A, hb.dx, hb.db, pc, omega, max_iter, tol, relative, breakdown);
    if (r < 0) return r;
    DB_CHECK(hb.down(A->csr.rows, x));
    return r;
}

int drops_b200_gcr_dev(drops_b200_matrix_t A, drops_b200_vector_t x, drops_b200_vector_t b, int pc, double omega,
                       int truncate, int* max_iter, double* tol, int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b && x->n == A->csr.rows && b->n == A->csr.rows, "GCR: incompatible dimensions");
    return gcr_device(A, x->d, b->d, pc, omega, truncate, max_iter, tol, relative);
}

int drops_b200_gcr(drops_b200_matrix_t A, double* x, const double* b, int pc, double omega, int truncate,
                   int* max_iter, double* tol, int relative)
{
    DB_CHECK(require_ctx());
    DB_ARG(A && x && b, "null argument");
    HostBridge hb(A);
    DB_CHECK(hb.up(A->csr.rows, x, b));
    const int r = gcr_device(A, hb.dx, hb.db, pc, omega, truncate, max_iter, tol, relative);
    if (r < 0) return r;
    DB_CHECK(hb.down(A->csr.rows, x));
    return r;
}

}  // extern "C"

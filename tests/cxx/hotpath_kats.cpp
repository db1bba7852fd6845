// tests/cxx/hotpath_kats.cpp -- known-answer program for the sparse-solver hot path, written against
// the DROPS header interface only (num/spmat.h, num/krylovsolver.h, num/precond.h,
// parallel/exchange.h).  The SAME source is compiled twice:
//   * against the unmodified reference headers (/root/reference/src)  -> expected output,
//     committed as tests/golden/hotpath_kats.expected.txt by tests/cxx/make_expected.sh;
//   * against the B200 drop-in headers (drops_b200/hostxx) + libdrops_b200.so -> run on the GPU.
// Style follows the reference's own test mains (src/tests/minres.cpp, blockmat.cpp, vectest.cpp):
// one line "NAME: value ..." per checked quantity, scraped by tests/test_hostxx.py.  The small
// systems are the reference's own (minres.cpp:94-103, blockmat.cpp:34-42) plus a 3D 7-point
// Laplacian for PCG/GMRES with JACPcCL, which the reference never tests.
#include "num/krylovsolver.h"
#include "num/precond.h"
#include "num/spblockmat.h"
#include "num/spmat.h"
#include "parallel/exchange.h"

#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <iostream>

using namespace DROPS;

static void print_vec(const char* name, const VectorCL& v)
{
    std::printf("%s:", name);
    for (size_t i = 0; i < v.size(); ++i) std::printf(" %.15e", v[i]);
    std::printf("\n");
}

static void fill_4x4(MatrixCL& A, VectorCL& b)
{
    static const double m[4][4] = {{-249., -453., -397., -52.}, {-453., -731., -601., -78.},
                                   {-397., -601., -648., -91.}, {-52., -78., -91., -13.}};
    MatrixBuilderCL AB(&A, 4, 4);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) AB(i, j) = m[i][j];
    AB.Build();
    b.resize(4);
    b[0] = -1277. / 2.; b[1] = -2015. / 2.; b[2] = -3907. / 4.; b[3] = -533. / 4.;
}

// 7-point Laplacian on an n^3 grid plus a diagonal shift, assembled entry by entry like an FE code would.
static void fill_laplace3d(MatrixCL& A, int n, double shift, double upper_scale)
{
    const size_t N = size_t(n) * n * n;
    MatrixBuilderCL AB(&A, N, N);
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                const size_t r = (size_t(k) * n + j) * n + i;
                AB(r, r) += 6. + shift;
                if (i > 0) AB(r, r - 1) += -1.;
                if (i < n - 1) AB(r, r + 1) += -1. * upper_scale;
                if (j > 0) AB(r, r - n) += -1.;
                if (j < n - 1) AB(r, r + n) += -1. * upper_scale;
                if (k > 0) AB(r, r - size_t(n) * n) += -1.;
                if (k < n - 1) AB(r, r + size_t(n) * n) += -1. * upper_scale;
            }
    AB.Build();
}

int main()
{
  try {
    DummyExchangeCL ex;

    // --- Lanczos lucky breakdown, MINRES 2x2 (minres.cpp TEST1/TEST2) -------------------------
    {
        MatrixCL A;
        MatrixBuilderCL AB(&A, 2, 2);
        AB(0, 1) = 1.; AB(1, 0) = 1.;
        AB.Build();
        VectorCL r0(2);
        r0[0] = 1.; r0[1] = 2.;
        LanczosONBCL<VectorCL> onb;
        onb.new_basis(A, r0, ex);
        int vecs = 1;
        while (onb.next(A, ex)) ++vecs;
        ++vecs;
        std::printf("LANCZOS vectors: %d\n", vecs);

        VectorCL b(2), x(0., 2);
        b[0] = 1.; b[1] = 2.;
        int mi = 10; double tol = 1e-10;
        const bool ok = MINRES(A, x, b, ex, mi, tol);
        std::printf("MINRES2 converged: %d iterations: %d\n", int(ok), mi);
        print_vec("MINRES2 x", x);
        print_vec("MINRES2 Ax-b", VectorCL(A * x - b));
    }
    // --- MINRES / PMINRES 4x4 (minres.cpp TEST3-TEST5) ------------------------------------------
    {
        MatrixCL A; VectorCL b;
        fill_4x4(A, b);
        {
            VectorCL x(0., 4);
            int mi = 10; double tol = 1e-10;
            MINRES(A, x, b, ex, mi, tol);
            std::printf("MINRES4 iterations: %d residual: %.6e\n", mi, tol);
            print_vec("MINRES4 x", x);
        }
        {
            VectorCL x(0., 4);
            LanczosONBCL<VectorCL> q;
            PMResSolverCL<LanczosONBCL<VectorCL> > pmr(q, 10, 1e-10);
            pmr.Solve(A, x, b, ex);
            std::printf("PMINRES4 iterations: %d residual: %.6e\n", pmr.GetIter(), pmr.GetResid());
            print_vec("PMINRES4 x", x);
        }
        {
            VectorCL x(0., 4);
            DummyPcCL pc;
            PLanczosONBCL<VectorCL, DummyPcCL> q(pc);
            PMResSolverCL<PLanczosONBCL<VectorCL, DummyPcCL> > pmr(q, 10, 1e-10);
            pmr.Solve(A, x, b, ex);
            std::printf("PMINRES4D iterations: %d residual: %.6e\n", pmr.GetIter(), pmr.GetResid());
            print_vec("PMINRES4D x", x);
        }
    }
    // --- BiCGStab / GCR 2x2 and 4x4 (bicgstab.cpp + TestBiCGStab.ref, gcr.cpp + TestGCR.ref) ---------
    {
        MatrixCL A2;
        MatrixBuilderCL AB(&A2, 2, 2);
        AB(0, 1) = 1.; AB(1, 0) = 1.;
        AB.Build();
        VectorCL b2(2);
        b2[0] = 1.; b2[1] = 2.;
        MatrixCL A4; VectorCL b4;
        fill_4x4(A4, b4);
        DummyPcCL pc;
        {
            VectorCL x(0., 2);
            int mi = 10; double tol = 1e-10;
            BICGSTAB(A2, x, b2, ex, pc, mi, tol);
            std::printf("BICGSTAB2 iterations: %d residual: %.6e\n", mi, tol);
            print_vec("BICGSTAB2 x", x);
        }
        {
            VectorCL x(0., 4);
            int mi = 10; double tol = 1e-10;
            BICGSTAB(A4, x, b4, ex, pc, mi, tol);
            std::printf("BICGSTAB4 iterations: %d residual: %.6e\n", mi, tol);
            print_vec("BICGSTAB4 x", x);
        }
        {
            VectorCL x(3., 2);
            int mi = 100; double tol = 1e-10;
            GCR(A2, x, b2, ex, pc, 10, mi, tol, false);
            std::printf("GCR2 iterations: %d residual: %.6e\n", mi, tol);
            print_vec("GCR2 x", x);
        }
        {
            VectorCL x(0., 4);
            int mi = 100; double tol = 1e-10;
            GCR(A4, x, b4, ex, pc, 10, mi, tol, false);
            std::printf("GCR4 iterations: %d residual: %.6e\n", mi, tol);
            print_vec("GCR4 x", x);
        }
    }
    // --- transp_mul (blockmat.cpp) -----------------------------------------------------------------
    {
        MatrixCL B;
        MatrixBuilderCL BB(&B, 1, 2);
        BB(0, 0) = 1.; BB(0, 1) = 1.;
        BB.Build();
        VectorCL b2(1);
        b2[0] = 3.;
        print_vec("TRANSP_MUL", transp_mul(B, b2));
        VectorCL c(2);
        c[0] = 1.; c[1] = 2.;
        print_vec("B*c", VectorCL(B * c));
    }
    // --- BlockMatrixCL / CompositeMatrixCL (blockmat.cpp + Testblockmat.ref) ------------------------------
    {
        MatrixCL A, B;
        MatrixBuilderCL AB(&A, 2, 2);
        AB(0, 1) = 1.; AB(1, 0) = 1.;
        AB.Build();
        MatrixBuilderCL BB(&B, 1, 2);
        BB(0, 0) = 1.; BB(0, 1) = 1.;
        BB.Build();
        BlockMatrixCL S(&A, MUL, &B, TRANSP_MUL, &B, MUL);
        VectorCL b(3);
        b[0] = 1.; b[1] = 2.; b[2] = 3.;
        std::printf("BLOCK dims: %zu %zu %zu %zu %zu %zu\n", S.num_rows(0), S.num_rows(1), S.num_cols(0), S.num_cols(1),
                    S.num_rows(), S.num_cols());
        print_vec("BLOCK S*b", VectorCL(S * b));
        print_vec("BLOCK transp_mul", VectorCL(transp_mul(S, b)));
        BlockMatrixCL S2(0, MUL, &B, TRANSP_MUL, &B, MUL, 0, MUL);            // no diagonal blocks
        print_vec("BLOCK nodiag", VectorCL(S2 * b));
        MatrixCL Sfull(BuildMatrix(S));
        print_vec("BLOCK built", VectorCL(Sfull * b));
        DummyExchangeCL dex;
        CompositeMatrixCL BBt(&B, TRANSP_MUL, dex, &B, MUL, dex);             // B * B^T  (1x1)
        VectorCL one(1);
        one[0] = 2.;
        std::printf("COMPOSITE dims: %zu %zu %zu\n", BBt.num_rows(), BBt.num_cols(), BBt.intermediate_dim());
        print_vec("COMPOSITE BBt", VectorCL(BBt * one));
        CompositeMatrixCL AtB(&B, TRANSP_MUL, dex, &A, MUL, dex);             // A * B^T  (2x1)
        print_vec("COMPOSITE ABt", VectorCL(AtB * one));
        VectorCL two(2);
        two[0] = 1.; two[1] = -3.;
        print_vec("COMPOSITE ABt_T", VectorCL(transp_mul(AtB, two)));
    }
    // --- text I/O: out/in round trip, MatrixMarket header of operator<< (spmat.h:1213-1255) -------------
    {
        MatrixCL A; VectorCL b;
        fill_4x4(A, b);
        std::stringstream ss;
        ss.precision(17);
        out(ss, A);
        MatrixCL B;
        in(ss, B);
        double maxdiff = 0.;
        for (size_t k = 0; k < A.num_nonzeros(); ++k) maxdiff = std::max(maxdiff, std::abs(A.val(k) - B.val(k)));
        std::printf("IO out/in rows: %zu nnz: %zu maxdiff: %.3e\n", B.num_rows(), B.num_nonzeros(), maxdiff);
        std::stringstream mm;
        mm << A;
        std::string l1, l2, l3;
        std::getline(mm, l1); std::getline(mm, l2); std::getline(mm, l3);
        std::printf("IO mtx header ok: %d dims: %s first: %s\n", int(l1 == "%%MatrixMarket matrix coordinate real general"),
                    l2.c_str(), l3.c_str());
        std::stringstream vs;
        vs.precision(17);
        out(vs, b);
        VectorCL c;
        in(vs, c);
        std::printf("IO vec size: %zu diff: %.3e\n", c.size(), norm(VectorCL(b - c)));
    }
    // --- vector expressions (vectest.cpp) ------------------------------------------------------------
    {
        const size_t n = 1000;
        VectorCL v1(n), v2(n), v3(n);
        for (size_t i = 0; i < n; ++i) { v1[i] = 1. / (1. + i); v2[i] = 0.5 * i; v3[i] = 3. - 1e-3 * i; }
        const double a = 0.25;
        VectorCL z(n);
        z = v1 + 3. * v2 + a * v3;
        double maxdiff = 0.;
        for (size_t i = 0; i < n; ++i) maxdiff = std::max(maxdiff, std::abs(z[i] - (v1[i] + 3. * v2[i] + a * v3[i])));
        std::printf("VECEXPR maxdiff: %.3e\n", maxdiff);
        VectorCL y(v2);
        axpy(a, v1, y);
        z_xpay(z, v1, a, v2);
        std::printf("VECOPS dot: %.15e norm: %.15e supnorm: %.15e axpy: %.15e z_xpay: %.15e\n", dot(v1, v2), norm(v3),
                    supnorm(v2), y[7], z[7]);
        std::printf("EXCHANGE pardot: %.15e norm: %.15e\n", ex.ParDot(v1, false, v3, false), ex.Norm(v3, false));
    }
    // --- PCG + JACPcCL, CG, GMRES + JACPcCL on a 3D Laplacian ----------------------------------------
    {
        const int n = 12;
        MatrixCL A;
        fill_laplace3d(A, n, 0.05, 1.0);
        const size_t N = A.num_rows();
        VectorCL xs(N), x(0., N);
        for (size_t i = 0; i < N; ++i) xs[i] = 1. + 0.01 * double(i % 17);
        VectorCL b(A * xs);
        std::printf("LAPLACE rows: %zu nnz: %zu version>0: %d\n", A.num_rows(), A.num_nonzeros(), int(A.Version() > 0));
        {
            JACPcCL pc(1.0);
            PCGSolverCL<JACPcCL> solver(pc, 500, 1e-10, true);
            solver.Solve(A, x, b, ex);
            std::printf("PCGJAC iterations: %d residual: %.3e error: %.3e\n", solver.GetIter(), solver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
        }
        {
            x = 0.;
            JACPcCL pc(0.8);
            int it = 500; double tol = 1e-9;
            const bool ok = PCG(A, x, b, ex, pc, it, tol, false);
            std::printf("PCGJAC08 converged: %d iterations: %d error: %.3e\n", int(ok), it, norm(VectorCL(x - xs)) / norm(xs));
        }
        {
            x = 0.;
            CGSolverCL solver(500, 1e-10, true);
            solver.Solve(A, x, b, ex);
            std::printf("CG iterations: %d residual: %.3e error: %.3e\n", solver.GetIter(), solver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
        }
        {
            x = 0.;
            JACPcCL pc(1.0);
            PCGSolverCL<JACPcCL> solver(pc, 5, 1e-10, true);   // too few iterations: not converged
            solver.Solve(A, x, b, ex);
            std::printf("PCGJAC5 iterations: %d converged: %d\n", solver.GetIter(), int(solver.GetResid() <= 1e-10));
        }
        {
            x = 0.;
            JACPcCL pc(1.0);
            VectorCL z(N);
            pc.Apply(A, z, b, ex);
            std::printf("JACAPPLY z7: %.15e z100: %.15e\n", z[7], z[100]);
        }
        {
            // PCG with DiagPcCL (precond.h:618-642): D = 1 / (a_ii + a perturbation that is not the Jacobi diagonal)
            x = 0.;
            VectorCL D(A.GetDiag());
            for (size_t i = 0; i < N; ++i) D[i] = 1. / (D[i] * (1. + 0.05 * double(i % 5)));
            DiagPcCL dpc(D);
            PCGSolverCL<DiagPcCL> solver(dpc, 500, 1e-10, true);
            solver.Solve(A, x, b, ex);
            std::printf("PCGDIAG iterations: %d residual: %.3e error: %.3e\n", solver.GetIter(), solver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
            VectorCL z(N);
            dpc.Apply(A, z, b, ex);
            std::printf("DIAGAPPLY z7: %.15e z100: %.15e\n", z[7], z[100]);
        }
        {
            // PCG with the Chebyshev polynomial preconditioner (precond.h:1463-1529); SetDiag prints lambda_max
            x = 0.;
            ChebyshevPcCL cpc(1.1, 3, 30.0);
            PCGSolverCL<ChebyshevPcCL> solver(cpc, 500, 1e-10, true);
            std::ostringstream sink;
            std::streambuf* oldbuf = std::cout.rdbuf(sink.rdbuf());
            solver.Solve(A, x, b, ex);
            std::cout.rdbuf(oldbuf);
            std::printf("PCGCHEB iterations: %d residual: %.3e error: %.3e\n", solver.GetIter(), solver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
            const std::string log = sink.str();
            const size_t pos = log.find("lambda_max = ");
            std::printf("PCGCHEB lambda_max: %.5g\n", pos == std::string::npos ? -1. : std::atof(log.c_str() + pos + 13));
            // PCGNE with ChebyshevBBTPcCL on B = every second row of A (krylovsolver.h:199-252, precond.h:1531-1567)
            MatrixCL Bh;
            {
                MatrixBuilderCL BB(&Bh, (N + 1) / 2, N);
                for (size_t i = 0; i < N; i += 2)
                    for (size_t nz = A.row_beg(i); nz < A.row_beg(i + 1); ++nz) BB(i / 2, A.col_ind(nz)) = A.val(nz);
                BB.Build();
            }
            VectorCL us(Bh.num_rows());
            for (size_t i = 0; i < us.size(); ++i) us[i] = 1. + 0.01 * double(i % 17);
            const VectorCL bu(Bh * VectorCL(transp_mul(Bh, us)));
            VectorCL u(Bh.num_rows());
            ChebyshevBBTPcCL bpc(1., 3, 30.0);
            PCGNESolverCL<ChebyshevBBTPcCL> nesolver(bpc, 500, 1e-10, true);
            oldbuf = std::cout.rdbuf(sink.rdbuf());
            nesolver.Solve(Bh, u, bu, ex, ex);
            std::cout.rdbuf(oldbuf);
            std::printf("PCGNE iterations: %d residual: %.3e error: %.3e\n", nesolver.GetIter(), nesolver.GetResid(),
                        norm(VectorCL(u - us)) / norm(us));
            const VectorCL bbt(BBTDiag(Bh, ex, ex));
            std::printf("BBTDIAG d0: %.15e d9: %.15e\n", bbt[0], bbt[9]);
            VectorCL zc(N);
            cpc.Apply(A, zc, b, ex);
            std::printf("CHEBAPPLY z7: %.12e z100: %.12e\n", zc[7], zc[100]);
        }
        {
            MatrixCL G;
            fill_laplace3d(G, n, 0.5, 1.1);         // nonsymmetric
            VectorCL bg(G * xs);
            x = 0.;
            JACPcCL pc(1.0);
            GMResSolverCL<JACPcCL> solver(pc, 20, 500, 1e-10);
            solver.Solve(G, x, bg, ex);
            std::printf("GMRESJAC iterations: %d residual: %.3e error: %.3e\n", solver.GetIter(), solver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
            x = 0.;
            GMResSolverCL<JACPcCL> rsolver(pc, 20, 500, 1e-10, true, false, RightPreconditioning);
            rsolver.Solve(G, x, bg, ex);
            std::printf("GMRESJACR iterations: %d residual: %.3e error: %.3e\n", rsolver.GetIter(), rsolver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
            x = 0.;
            BiCGStabSolverCL<JACPcCL> bsolver(pc, 500, 1e-10);
            bsolver.Solve(G, x, bg, ex);
            std::printf("BICGSTABJAC iterations: %d residual: %.3e error: %.3e\n", bsolver.GetIter(), bsolver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
            x = 0.;
            GCRSolverCL<JACPcCL> gsolver(pc, 100, 500, 1e-10);
            gsolver.Solve(G, x, bg, ex);
            std::printf("GCRJAC iterations: %d residual: %.3e error: %.3e\n", gsolver.GetIter(), gsolver.GetResid(),
                        norm(VectorCL(x - xs)) / norm(xs));
            x = 0.;
            IDRsSolverCL<JACPcCL> isolver(pc, 500, 1e-10, true, 4, 0.7);      // random shadow space: only the
            isolver.Solve(G, x, bg, ex);                                       // outcome is comparable
            std::printf("IDRS solved: %d error ok: %d\n", int(isolver.GetResid() / norm(bg) <= 1e-10),
                        int(norm(VectorCL(x - xs)) / norm(xs) <= 1e-8));
            // time-step matrix L = c1*A + c2*G (LinComb, spmat.h:1457-1551): merged on the device when the
            // operands are resident -- no upload before the solve
            {
                MatrixCL Lm;
                const size_t v0 = Lm.Version();
                Lm.LinComb(2.5, A, 1., G);
                VectorCL bl(Lm * xs);
                double md = 0.;
                const VectorCL chk(2.5 * VectorCL(A * xs) + VectorCL(G * xs));
                for (size_t i = 0; i < N; ++i) md = std::max(md, std::abs(bl[i] - chk[i]) / (1. + std::abs(chk[i])));
                std::printf("LINCOMB nnz: %zu version bumped: %d maxreldiff: %.1e L(3,3): %.15e\n", Lm.num_nonzeros(),
                            int(Lm.Version() > v0), md > 1e-13 ? md : 0., Lm(3, 3));
            }
            // matrix modified in place: Version() changes, the product must follow
            G *= 2.;
            VectorCL b2(G * xs);
            std::printf("VERSIONED ratio: %.15e\n", b2[5] / bg[5]);
        }
    }
    std::printf("hotpath_kats finished regularly\n");
    return 0;
  }
  catch (DROPSErrCL err) { err.handle(); }
  return 1;
}

"""The reference's own Poisson executable, cdrdrops (poisson/cdrdrops.cpp), as the drop-in test of the boundary:
tests/cxx/drops_exe/Makefile compiles the UNMODIFIED sources under /root/reference twice -- against the reference
headers ("ref") and with drops_b200/hostxx first on the include path + libdrops_b200.so ("b200") -- and the two
binaries must tell the same story on the same parameter files: mesh from BrickBuilderCL, DROPS' own FE assembly,
PoissonSolverFactoryCL codes 202 (PCGSolverCL<JACPcCL>), 302 (GMResSolverCL<JACPcCL>), 207, 203.

tests/golden/cdrdrops_reference.json holds what the ref binary printed here (tests/cxx/drops_exe/make_golden.py).
CPU: both variants build, reference TUs outside the path compile against the overlay, the golden file is
consistent with the reference's own TestCDRDrops.ref-style expectations.  GPU: the b200 binary reproduces the
reference's iteration counts (+-1), residual bound, unknown count and discretisation error.
/root/reference does not exist on the GPU box: there the prebuilt binaries under tests/cxx/build/ are used.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cxx", "drops_exe"))
import cases  # noqa: E402

HAVE_REF = os.path.isdir("/root/reference/src/num")


def golden():
    return json.load(open(cases.GOLDEN))


def ensure_built(variant):
    if HAVE_REF:
        subprocess.check_call(["make", "-C", cases.HERE, variant], stdout=subprocess.DEVNULL)
    if not os.path.exists(cases.exe(variant)):
        pytest.skip(f"{cases.exe(variant)} not built and /root/reference is absent on this box")
    return cases.exe(variant)


@pytest.mark.skipif(not HAVE_REF, reason="needs the DROPS source tree")
def test_cdrdrops_builds_against_reference_and_overlay(lib_built):
    ref, b200 = ensure_built("ref"), ensure_built("b200")
    assert "libdrops_b200.so" in subprocess.run(["ldd", b200], capture_output=True, text=True).stdout
    assert "libdrops_b200.so" not in subprocess.run(["ldd", ref], capture_output=True, text=True).stdout
    # the overlay's solvers were chosen for the Jacobi-preconditioned drivers of the factory (202 / 302): the b200
    # binary holds no instantiation of the reference's PCG / GMRES templates for JACPcCL = PreGSCL<P_JAC0> ...
    sym = subprocess.run(["nm", "-C", b200], capture_output=True, text=True).stdout
    jac = "DROPS::PreGSCL<(DROPS::PreMethGS)12, false>"
    ref_templates = [l for l in sym.splitlines() if (" bool DROPS::PCG<" in l or " bool DROPS::GMRES<" in l) and jac in l]
    assert not ref_templates, ref_templates[:3]
    # ... while the reference binary does, and SSOR (203 / 303) still runs the reference's templates in both
    sym_ref = subprocess.run(["nm", "-C", ref], capture_output=True, text=True).stdout
    assert any((" bool DROPS::PCG<" in l or " bool DROPS::GMRES<" in l) and jac in l for l in sym_ref.splitlines())
    assert any(" bool DROPS::PCG<" in l and "(DROPS::PreMethGS)7" in l for l in sym.splitlines())


@pytest.mark.skipif(not HAVE_REF, reason="needs the DROPS source tree")
def test_reference_units_outside_the_path_compile_against_the_overlay(lib_built):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cxx"), "syntax"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for unit in ("num/MGsolver.cpp", "geom/simplex.cpp", "num/fe.cpp", "misc/problem.cpp"):
        assert f"ok {unit}" in r.stdout


def test_golden_holds_the_reference_runs():
    g = golden()
    assert set(g) == set(cases.CASES)
    c1 = g["c1_p1_pcg_jac"]                       # BASELINE configs[0]: 4 refinements of the 4x4x4 brick, one Neumann face
    assert c1["unknowns"] == 63 * 63 * 64 and c1["factory_code"] == 202
    for name, rec in g.items():
        assert rec["finished"] and 0 < rec["iterations"] < 1000, name
        assert rec["residuum"] <= 1e-10, name
    assert abs(g["c1_p1_gmres_jac"]["l2_error"] - c1["l2_error"]) <= 1e-6 * c1["l2_error"]    # same discrete solution


@pytest.mark.skipif(not HAVE_REF, reason="needs the DROPS source tree")
def test_reference_binary_reproduces_the_golden_file(tmp_path):
    ensure_built("ref")
    _, rec = cases.run("ref", "small_p1_pcg_jac", str(tmp_path))
    g = golden()["small_p1_pcg_jac"]
    assert rec["finished"] and rec["unknowns"] == g["unknowns"] and rec["iterations"] == g["iterations"]
    assert abs(rec["l2_error"] - g["l2_error"]) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(cases.CASES))
def test_overlay_binary_matches_the_reference(gpu, case, tmp_path):
    ensure_built("b200")
    out, rec = cases.run("b200", case, str(tmp_path))
    g = golden()[case]
    assert rec["finished"], out[-2000:]
    assert rec["unknowns"] == g["unknowns"]
    assert abs(rec["iterations"] - g["iterations"]) <= 1, (rec["iterations"], g["iterations"])
    assert rec["residuum"] <= 1e-10                                  # Poisson.Solver.Tol, absolute (useRelTol 0)
    # "real residuum" is recomputed by cdrdrops.cpp:122 with A*x -- the GPU product of the overlay
    assert rec["real_residuum"] <= 1.5 * max(g["real_residuum"], 1e-10)
    # the discretisation error against the known solution is a property of the solution vector: 6 printed digits
    assert abs(rec["l2_error"] - g["l2_error"]) <= 2e-6 * g["l2_error"], (rec["l2_error"], g["l2_error"])


@pytest.mark.gpu
def test_line_segments_on_a_matrix_drops_assembled_itself(gpu, tmp_path):
    """The segment path is not a property of the synthetic generator: DROPS' own mesh (BrickBuilderCL, 40 cells per
    direction, unrefined), numbering (misc/problem.cpp) and P2 assembly hand over a matrix in which 94 % of the rows have a
    translate at most 8 rows away; with the size threshold lifted the overlay cuts it into line segments, runs most stored
    entries from segment descriptors, and cdrdrops prints the reference's iteration count and discretisation error.
    (The hierarchical numbering after uniform refinement -- case p2_pcg_jac -- has no such rows: the per-row graph blocks
    and the packed kernel run there.)"""
    ensure_built("b200")
    env = dict(os.environ, DROPS_B200_TRACE="1", DROPS_B200_OPTIONS="sell.block_min_rows=0")
    out, rec = cases.run("b200", "p2_brick40_pcg_jac", str(tmp_path), env=env)
    g = golden()["p2_brick40_pcg_jac"]
    assert rec["finished"] and rec["unknowns"] == g["unknowns"] and abs(rec["iterations"] - g["iterations"]) <= 1
    assert abs(rec["l2_error"] - g["l2_error"]) <= 2e-6 * g["l2_error"]
    lay = max(rec["layouts"], key=lambda l: l["rows"])
    assert lay["rows"] == g["unknowns"] and lay["block_ordered"] == 1 and lay["row_ordered"] >= 1, lay
    assert lay["line_segments"] * 32 >= 0.75 * lay["rows"], lay            # 82 % of the rows in complete segments at N = 40
    # DROPS' assembly rounds: hundreds of distinct values where exact arithmetic has 21 -- no shared index per entry-row, so
    # the one-word form keeps its packed words; the six-byte form (what C2 size needs: 26 column bits + 10 index bits)
    # takes the segment records
    assert lay["dictionary"] > 100, lay
    env6 = dict(env, DROPS_B200_OPTIONS="sell.block_min_rows=0,sell.pack_bits=0")
    out6, rec6 = cases.run("b200", "p2_brick40_pcg_jac", str(tmp_path / "six"), env=env6)
    lay6 = max(rec6["layouts"], key=lambda l: l["rows"])
    # (lines of 39 rows: every complete segment starts next to the boundary rows of its line, and an entry-row that meets
    #  them is three runs, not two -- a fifth of the slices qualify here, most of them on a mesh with longer lines)
    assert lay6["packed"] == 2 and lay6["row_ordered"] == 2 and lay6["segment_entries"] >= 0.1 * lay6["nnz"], lay6
    assert rec6["finished"] and abs(rec6["iterations"] - g["iterations"]) <= 1
    assert abs(rec6["l2_error"] - g["l2_error"]) <= 2e-6 * g["l2_error"]
    # the refined-mesh numbering: no lines, graph blocks per row
    out2, rec2 = cases.run("b200", "p2_pcg_jac", str(tmp_path / "refined"), env=env)
    lay2 = max(rec2["layouts"], key=lambda l: l["rows"])
    assert lay2["block_ordered"] == 1 and lay2["line_segments"] * 32 <= 0.05 * lay2["rows"], lay2
    assert abs(rec2["iterations"] - golden()["p2_pcg_jac"]["iterations"]) <= 1


@pytest.mark.gpu
def test_overlay_binary_against_live_reference_binary(gpu, tmp_path):
    """Both binaries run side by side on this box (the ref one travels prebuilt)."""
    ensure_built("b200")
    if not os.path.exists(cases.exe("ref")):
        pytest.skip("reference binary not built")
    _, a = cases.run("ref", "small_p1_pcg_jac", str(tmp_path / "ref"))
    _, b = cases.run("b200", "small_p1_pcg_jac", str(tmp_path / "b200"))
    assert a["unknowns"] == b["unknowns"] and abs(a["iterations"] - b["iterations"]) <= 1
    assert abs(a["l2_error"] - b["l2_error"]) <= 2e-6 * a["l2_error"]

"""The C++ host side above the C ABI: drops_b200/hostxx, overlay headers that #include_next DROPS' own
num/spmat.h and num/krylovsolver.h and add the overloads that send the hot path to the GPU.

tests/cxx/hotpath_kats.cpp is compiled against the UNMODIFIED reference headers (expected output,
committed as tests/golden/hotpath_kats.expected.txt) and with the overlay first on the include path +
libdrops_b200.so.  CPU: the overlay build compiles and links (where the DROPS tree exists), and the committed
expectation holds the reference's .ref values.  GPU: the overlay binary reproduces the reference's output line
by line (prebuilt binary where /root/reference is absent).
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX_DIR = os.path.join(ROOT, "tests", "cxx")
EXPECTED = os.path.join(ROOT, "tests", "golden", "hotpath_kats.expected.txt")

KEYS = ("LANCZOS", "MINRES2", "MINRES4", "PMINRES4", "PMINRES4D", "TRANSP_MUL", "B*c", "VECEXPR", "VECOPS",
        "EXCHANGE", "LAPLACE", "PCGJAC", "PCGJAC08", "CG", "PCGJAC5", "JACAPPLY", "GMRESJAC", "GMRESJACR",
        "VERSIONED", "hotpath_kats", "BICGSTAB2", "BICGSTAB4", "GCR2", "GCR4", "BICGSTABJAC", "GCRJAC", "IO out", "IO mtx", "IO vec", "LINCOMB", "BLOCK", "COMPOSITE", "PCGCHEB", "PCGNE",
        "BBTDIAG", "CHEBAPPLY", "IDRS", "PCGDIAG", "DIAGAPPLY")


def parse(text):
    """{first token(s) before ':' -> list of numbers on the line} for the program's own lines."""
    out = {}
    for line in text.splitlines():
        m = re.match(r"^([A-Za-z0-9_*]+(?: [A-Za-z\-]+)?):? ?(.*)$", line)
        if not m or not line.startswith(KEYS):
            continue
        key = line.split(":")[0] if ":" in line else line
        nums = [float(t) for t in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line[len(key):])]
        out[key] = nums
    return out


HAVE_REF = os.path.isdir("/root/reference/src/num")


def build_b200(lib_built):
    exe = os.path.join(CXX_DIR, "build", "kats_b200")
    if not HAVE_REF and not os.path.exists(exe):
        pytest.skip("tests/cxx/build/kats_b200 not prebuilt and /root/reference is absent on this box")
    subprocess.check_call(["make", "-C", CXX_DIR, "b200"], stdout=subprocess.DEVNULL)
    return exe


@pytest.mark.skipif(not HAVE_REF, reason="needs the DROPS source tree")
def test_overlay_headers_compile_and_link(lib_built):
    exe = build_b200(lib_built)
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libdrops_b200.so" in out
    # the overlay adds no header of its own for anything but the two hot-path headers
    hostxx = os.path.join(ROOT, "drops_b200", "hostxx")
    files = sorted(os.path.relpath(os.path.join(d, f), hostxx) for d, _, fs in os.walk(hostxx) for f in fs)
    assert files == ["num/krylovsolver.h", "num/spmat.h"], files
    for f in files:
        assert "#include_next" in open(os.path.join(hostxx, f)).read()


def test_overlay_binary_refuses_to_run_without_a_gpu(lib_built):
    """No CPU fallback: on a box without a device the first product throws (DROPSErrCL from the C ABI status)."""
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    exe = build_b200(lib_built)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 or "finished regularly" not in r.stdout
    assert "no CPU fallback" in (r.stdout + r.stderr)


def test_expected_output_holds_the_reference_kats():
    exp = parse(open(EXPECTED).read())
    assert exp["LANCZOS vectors"] == [2]                               # "lucky breakdown after 2 vectors"
    assert abs(exp["MINRES4 iterations"][1] - 2.87588e-11) < 1e-10      # TestMinres.ref Residual3
    assert abs(exp["PMINRES4D iterations"][1] - 2.87596e-11) < 1e-10    # TestMinres.ref Residual5
    assert exp["TRANSP_MUL"] == [3.0, 3.0]                              # Testblockmat.ref
    assert exp["VECEXPR maxdiff"][0] <= 1e-16                           # Testvectest.ref
    # bicgstab.cpp / TestBiCGStab.ref and gcr.cpp / TestGCR.ref: residuals 0 +- 1e-10
    for key in ("BICGSTAB2 iterations", "BICGSTAB4 iterations", "GCR2 iterations", "GCR4 iterations"):
        assert abs(exp[key][1]) < 1e-10, key
    assert exp["BLOCK S*b"] == [5.0, 4.0, 3.0] and exp["BLOCK transp_mul"] == [5.0, 4.0, 3.0]   # Testblockmat.ref
    assert exp["IDRS solved"] == [1.0, 1.0]
    assert "hotpath_kats finished regularly" in open(EXPECTED).read()


@pytest.mark.gpu
def test_dropin_binary_matches_reference_output(lib_built):
    exe = build_b200(lib_built)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got, exp = parse(r.stdout), parse(open(EXPECTED).read())
    assert "hotpath_kats finished regularly" in r.stdout
    missing = sorted(set(exp) - set(got))
    assert not missing, missing
    # exact lines: host-side vector code and bit-identical device products
    for key in ("LANCZOS vectors", "TRANSP_MUL", "B*c", "VECEXPR maxdiff", "VECOPS dot", "EXCHANGE pardot",
                "LAPLACE rows", "VERSIONED ratio", "JACAPPLY z7", "MINRES2 converged", "LINCOMB nnz", "IO out/in rows",
                "IO mtx header ok", "IO vec size"):
        assert got[key] == exp[key], (key, got[key], exp[key])
    # solutions of the tiny systems
    for key in ("MINRES2 x", "MINRES4 x", "PMINRES4 x", "PMINRES4D x"):
        assert all(abs(a - b) <= 1e-9 * max(1.0, abs(b)) for a, b in zip(got[key], exp[key])), key
    # iterations (+-1) and residuals (the .ref files use an absolute 1e-10)
    for key in ("MINRES4 iterations", "PMINRES4 iterations", "PMINRES4D iterations"):
        assert abs(got[key][0] - exp[key][0]) <= 1 and abs(got[key][1] - exp[key][1]) < 1e-10, key
    for key in ("PCGJAC iterations", "CG iterations", "GMRESJAC iterations", "GMRESJACR iterations"):
        assert abs(got[key][0] - exp[key][0]) <= 1, (key, got[key], exp[key])
        assert got[key][1] <= 1e-10 and got[key][2] <= 1e-8, (key, got[key])
    # BiCGStab / GCR known-answer systems (TestBiCGStab.ref, TestGCR.ref: residual 0 +- 1e-10)
    for key in ("BICGSTAB2", "BICGSTAB4", "GCR2", "GCR4"):
        assert abs(got[key + " iterations"][0] - exp[key + " iterations"][0]) <= 1, key
        assert abs(got[key + " iterations"][1]) < 1e-10, key
        tolx = 1e-6 if key == "BICGSTAB4" else 1e-9      # the reference itself stops at 6e-12 on the 4x4 system
        assert all(abs(a - b) <= tolx * max(1.0, abs(b)) for a, b in zip(got[key + " x"], exp[key + " x"])), key
    for key in ("BICGSTABJAC iterations", "GCRJAC iterations"):
        assert abs(got[key][0] - exp[key][0]) <= 1, (key, got[key], exp[key])
        assert got[key][1] <= 1e-10 and got[key][2] <= 1e-8, (key, got[key])
    # section 8(f) callers: block / composite products exact, Chebyshev-preconditioned solvers +-1 iteration
    for key in ("BLOCK dims", "BLOCK S*b", "BLOCK transp_mul", "BLOCK nodiag", "BLOCK built", "COMPOSITE dims",
                "COMPOSITE BBt", "COMPOSITE ABt", "COMPOSITE ABt_T", "BBTDIAG d0", "IDRS solved"):
        assert got[key] == exp[key], (key, got[key], exp[key])
    assert got["DIAGAPPLY z7"] == exp["DIAGAPPLY z7"]                          # DiagPcCL::Apply, precond.h:626-631
    for key in ("PCGCHEB iterations", "PCGNE iterations", "PCGDIAG iterations"):
        assert abs(got[key][0] - exp[key][0]) <= 1, (key, got[key], exp[key])
        assert got[key][1] <= 1e-10 and got[key][2] <= 1e-8, (key, got[key])
    assert got["PCGCHEB lambda_max"] == exp["PCGCHEB lambda_max"]              # as printed by SetDiag (6 digits)
    assert all(abs(a - b) <= 1e-11 * abs(b) for a, b in zip(got["CHEBAPPLY z7"], exp["CHEBAPPLY z7"]))
    assert abs(got["PCGJAC08 converged"][1] - exp["PCGJAC08 converged"][1]) <= 1 and got["PCGJAC08 converged"][0] == 1
    assert got["PCGJAC5 iterations"] == exp["PCGJAC5 iterations"] == [5.0, 0.0]

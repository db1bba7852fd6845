"""Parameter sets for the reference's own Poisson executable `cdrdrops` (poisson/cdrdrops.cpp) and the scraper of
its stdout -- shared by tests/test_drops_exe.py and make_golden.py (TEST INFRASTRUCTURE).

Base = the reference's example param/poisson/cdrdrops/statpoissonEx.json (restated here as a dict: unit cube,
BrickBuilder 4x4x4, one Neumann face, known solution Example1), with the error estimator and the VTK output
switched off.  BASELINE.json configs[0] is case "c1_p1_pcg_jac": 4 uniform refinements (64 cells per direction,
254 016 unknowns), Poisson.Solver.Solver = 202 = PCGSolverCL<JACPcCL> (num/poissonsolverfactory.h:234-236)."""
import copy
import json
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
EXE = {"b200": os.path.join(ROOT, "tests", "cxx", "build", "drops_exe", "b200", "cdrdrops"),      # overlay + libdrops_b200
       "ref": os.path.join(ROOT, "oracle", "_ref", "drops_exe", "ref", "cdrdrops")}                # the compiled reference
LIBDIR = os.path.join(ROOT, "oracle", "_ref", "drops_exe", "lib")     # the reference's function containers (dlopen'ed)
GOLDEN = os.path.join(ROOT, "tests", "golden", "cdrdrops_reference.json")

BASE = {
    "General": {"DynamicLibs": ["poisson/libpoisson-poissonCoeff", "misc/libmisc-scalarFunctions",
                                "misc/libmisc-vectorFunctions"],
                "DynamicLibsPrefix": LIBDIR + "/", "ProgressBar": 0},
    "Poisson": {
        "InitialValue": "Example1_InitialValue",
        "Solver": {"MG": {"NumLvl": -1, "SmoothingSteps": 1}, "Solver": 202, "Iter": 1000, "Tol": 1e-10,
                   "useRelTol": 0, "Restart": 100, "Relax": 1, "PcIter": 1, "PcTol": 0.1},
        "Coeff": {"Diffusion": 1, "Reaction": "Example1_Reaction", "withConvection": 0,
                  "Flowfield": "Example1_Flowfield", "Source": "Example1_Source"},
        "BoundaryData": {"Default": ["DirBC", "One"], "1": ["NatBC", "Example1_Neumann"]},
        "SolutionIsKnown": 1, "Solution": "Example1_Solution", "P1": 1},
    "Mesh": {"Type": "BrickBuilder", "Origin": [0, 0, 0], "E1": [1, 0, 0], "E2": [0, 1, 0], "E3": [0, 0, 1],
             "N1": 4, "N2": 4, "N3": 4, "AdaptRef": {"FinestLevel": 2}},
    "Restart": {"InputData": "none"},
    "Time": {"NumSteps": 0, "Scheme": 1, "FinalTime": 0, "Theta": 1},
    "Error": {"Meas": 0.875, "NumRef": 0, "RelReduction": 0.1, "DoMark": 0, "MinRatio": 0, "Threshold": 0.6,
              "DoErrorEstimate": 0},
    "VTK": {"VTKDir": "vtk", "Binary": 0, "Freq": 0, "VTKName": "poisson_ex_", "ReUseTimeFile": "0",
            "UseDeformation": "0", "UseOnlyP1": "0"},
    "Stabilization": {"SUPG": "0", "Magnitude": "1", "Grids": "1"},
    "ALE": {"wavy": "0", "P1": "1", "OnlyBndCurved": "1", "Interface": "Zero"},
}

# name -> (P1 elements?, refinement levels, factory code, where it runs with the overlay)
CASES = {
    "small_p1_pcg_jac":  (1, 2, 202, "device"),     # quick
    "c1_p1_pcg_jac":     (1, 4, 202, "device"),     # BASELINE configs[0]
    "c1_p1_gmres_jac":   (1, 4, 302, "device"),     # GMResSolverCL<JACPcCL>, restart 100
    "p2_pcg_jac":        (0, 3, 202, "device"),     # P2 elements, 254 016 unknowns as well
    "p1_pcg_chebyshev":  (1, 3, 207, "device"),     # PCGSolverCL<ChebyshevPcCL>
    "small_p1_pcg_ssor": (1, 2, 203, "reference template, products on the GPU"),   # out of scope: must keep working
    # BrickBuilder with 40 cells per direction and NO refinement (the way a 185^3 mesh -- BASELINE configs[1] -- has to be
    # built: 185 is not 4 * 2^k): DROPS numbers vertices lexicographically and edges cell by cell, so rows that are
    # translates of each other are a few rows apart -- the matrix the line segments / segment descriptors are for
    "p2_brick40_pcg_jac": (0, 0, 202, "device", 40),
}


def params(case):
    p1, level, code = CASES[case][:3]
    P = copy.deepcopy(BASE)
    if len(CASES[case]) > 4:
        P["Mesh"]["N1"] = P["Mesh"]["N2"] = P["Mesh"]["N3"] = CASES[case][4]
    P["Poisson"]["P1"] = p1
    P["Poisson"]["Solver"]["Solver"] = code
    P["Mesh"]["AdaptRef"]["FinestLevel"] = level
    return P


def exe(variant):
    return EXE[variant]


def run(variant, case, workdir, timeout=1200, env=None):
    """Runs cdrdrops of `variant` ("ref" / "b200") on `case` in workdir; returns (stdout, parsed dict)."""
    os.makedirs(workdir, exist_ok=True)
    path = os.path.join(workdir, case + ".json")
    json.dump(params(case), open(path, "w"), indent=1)
    json.dump({}, open(os.path.join(workdir, "default.json"), "w"))        # everything is in the case file
    r = subprocess.run([exe(variant), path], cwd=workdir, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"cdrdrops ({variant}, {case}) failed with {r.returncode}:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    rec = parse(r.stdout)
    # DROPS_B200_TRACE=1 (overlay binary): one stderr line per one-time cost and per solve
    rec["trace"] = {m.group(1).strip(): float(m.group(2))
                    for m in re.finditer(r"drops_b200 trace: (.*?) ([0-9.eE+-]+) s\n", r.stderr)}
    # ... and one "drops_b200 layout:" line per uploaded operator
    rec["layouts"] = [dict((k, int(v)) for k, v in re.findall(r"(\w+) (\d+)", m.group(1)))
                      for m in re.finditer(r"drops_b200 layout: (.*)\n", r.stderr)]
    return r.stdout, rec


def parse(text):
    def num(pat):
        m = re.search(pat, text)
        return float(m.group(1)) if m else None
    return {"unknowns": int(num(r"o number of unknowns\s+(\d+)") or -1),
            "iterations": int(num(r"- iterations\s+(-?\d+)") if num(r"- iterations\s+(-?\d+)") is not None else -1),
            "residuum": num(r"- residuum\s+(\S+)"),
            "real_residuum": num(r"- real residuum\s+(\S+)"),
            "solve_time_s": num(r"- time\s+(\S+) s"),
            "l2_error": num(r"L2-Norm= (\S+)"),
            "finished": "cdrdrops finished regularly" in text}

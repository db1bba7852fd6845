"""Runs the UNMODIFIED reference executable (oracle/_ref/drops_exe/ref/cdrdrops, built by this directory's
Makefile from /root/reference) on every case of cases.py and writes what it printed -- unknowns, iterations,
residuals, L2 error against the known solution -- to tests/golden/cdrdrops_reference.json.
    python tests/cxx/drops_exe/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases

subprocess.check_call(["make", "-C", cases.HERE, "ref"], stdout=subprocess.DEVNULL)
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name in cases.CASES:
        _, rec = cases.run("ref", name, tmp)
        assert rec["finished"], name
        rec["factory_code"] = cases.CASES[name][2]
        out[name] = rec
        print(name, rec, flush=True)
json.dump(out, open(cases.GOLDEN, "w"), indent=1, sort_keys=True)
print("wrote", cases.GOLDEN)

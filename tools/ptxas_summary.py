"""Registers / spills per kernel of one .cu (nvcc -Xptxas -v), demangled template arguments only."""
import re, subprocess, sys
src, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-I", "include",
       "-Xptxas", "-v", "-c", src, "-o", "/tmp/_ptxas.o"]
log = subprocess.run(cmd, capture_output=True, text=True).stderr
cur = None
for line in log.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("db200::", "").replace("void ", "")
        spill = ""
    elif "spill" in line:
        spill = re.sub(r"\s+", " ", line.strip())
    elif "Used" in line and cur and pat in cur:
        regs = re.search(r"Used (\d+) registers", line).group(1)
        sm = re.search(r"(\d+) bytes smem", line)
        print(f"{cur:70s} regs={regs:3s} smem={sm.group(1) if sm else 0} {'' if spill.startswith('0 bytes stack') else spill}")

"""Register / stack / spill summary per kernel from `nvcc -Xptxas -v` output on stdin (demangled with c++filt)."""
import re, subprocess, sys
txt = sys.stdin.read()
cur = None; rows = {}
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", ln)
    if m: cur = m.group(1); rows[cur] = {}; continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and cur: rows[cur].update(stack=int(m.group(1)), spill=int(m.group(2)))
    m = re.search(r"Used (\d+) registers", ln)
    if m and cur: rows[cur]["regs"] = int(m.group(1))
names = list(rows)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for n, d in zip(names, dem):
    if flt in d:
        r = rows[n]; print(f"{r.get('regs'):>4} regs {r.get('stack', 0):>4} stack {r.get('spill', 0):>4} spill  {d[:110]}")

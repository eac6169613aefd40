"""Hot spots of one kernel from an .ncu-rep source page: python tools/ncu_src_hot.py rep kernel_regex [top]"""
import csv, collections, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[0]; b = starts[1] if len(starts) > 1 else len(rows)
print(rows[a][1])
hdr = rows[a + 1]; data = rows[a + 2:b]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[ix["# Samples"]] or 0)
tot = sum(S(r) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("samples", tot, "instructions", len(data))
agg = collections.Counter()
for r in data:
    for h in stalls: agg[h] += int(r[ix[h]] or 0)
print(agg.most_common(10))
order = sorted(range(len(data)), key=lambda i: -S(data[i]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{i:5d} {S(r):5d} {100*S(r)/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {st}")

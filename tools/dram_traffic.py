"""Sum the DRAM bytes of all library kernels in an `ncu --csv` log and (optionally) record them per bench config.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv \
      --log-file gpurun_out/dram_c3.csv python tools/prof_case.py 1024 108 0 1
  python tools/dram_traffic.py gpurun_out/dram_c3.csv 108 [--config c3 --units-are fwd+adjoint --json profiles/r02_dram_traffic.json]

prof_case.py runs `reps` forward calls (complex64 -> |U|^2) and `reps` adjoint calls of B samples each, so the totals divided
by B*reps are the DRAM bytes of one fwd+adjoint unit (for the forward-only config c2 use --fwd-only to count forward kernels'
share: the tool then expects a log made with `prof_case.py N B pad reps 1`).  bench.py reads the JSON for `roofline.traffic`.
"""
import argparse
import collections
import csv
import json
import os
import subprocess
import time

ap = argparse.ArgumentParser()
ap.add_argument("log")
ap.add_argument("samples", type=int, help="B x reps of the profiled run")
ap.add_argument("--config", default="")
ap.add_argument("--json", default="")
ap.add_argument("--note", default="")
a = ap.parse_args()
rows = [r for r in csv.reader(open(a.log)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot = collections.Counter()
per_kernel = collections.defaultdict(collections.Counter)
for r in rows[1:]:
    name = r[ki]
    if "asmb::" in name or name.startswith("k32") or name.startswith("k_"):
        v = float(r[vi].replace(",", ""))
        tot[r[mi]] += v
        per_kernel[name.split("(")[0]][r[mi]] += v
rd, wr = tot["dram__bytes_read.sum"], tot["dram__bytes_write.sum"]
print(f"dram read {rd / 1e6:.1f} MB  write {wr / 1e6:.1f} MB  -> per unit read {rd / a.samples / 1e6:.2f} MB "
      f"write {wr / a.samples / 1e6:.2f} MB, total {(rd + wr) / a.samples / 1e6:.2f} MB ({a.samples} samples)")
for k, c in sorted(per_kernel.items()):
    print(f"  {k[:70]:70s} read {c['dram__bytes_read.sum'] / a.samples / 1e6:8.2f} MB/unit  write {c['dram__bytes_write.sum'] / a.samples / 1e6:8.2f} MB/unit")
if a.json and a.config:
    doc = {"configs": {}}
    if os.path.isfile(a.json):
        doc = json.load(open(a.json))
    try:
        rev = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        rev = ""
    doc["configs"][a.config] = {
        "dram_bytes_per_unit": (rd + wr) / a.samples, "read_bytes_per_unit": rd / a.samples, "write_bytes_per_unit": wr / a.samples,
        "samples": a.samples, "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum over all library kernels, --cache-control none "
                                     "--replay-mode application, tools/prof_case.py", "log": os.path.basename(a.log),
        "git": rev, "when": time.strftime("%Y-%m-%d"), "note": a.note,
        "per_kernel_MB_per_unit": {k: {"read": c["dram__bytes_read.sum"] / a.samples / 1e6, "write": c["dram__bytes_write.sum"] / a.samples / 1e6}
                                   for k, c in per_kernel.items()}}
    json.dump(doc, open(a.json, "w"), indent=1)
    print("wrote", a.json)

"""Sum dram bytes over all library kernels of one forward call (ncu --csv log on stdin path).  python tools/dram_traffic.py log.csv B"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value')
B = int(sys.argv[2])
tot = collections.Counter()
for r in rows[1:]:
    if 'asmb::' in r[ki] or r[ki].startswith('k32') or r[ki].startswith('k_'):
        tot[r[mi]] += float(r[vi].replace(',', ''))
rd, wr = tot['dram__bytes_read.sum'], tot['dram__bytes_write.sum']
print(f"dram read {rd/1e6:.1f} MB  write {wr/1e6:.1f} MB  -> per fwd+adjoint unit read {rd/B/1e6:.2f} MB write {wr/B/1e6:.2f} MB, total {(rd+wr)/B/1e6:.1f} MB ({B} samples per call)")

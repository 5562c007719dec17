"""Stall breakdown of an instruction range of an .ncu-rep source page: ncu_region.py rep.csv first last  (csv from `ncu -i rep --page source --csv`)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
a, b = int(sys.argv[2]), int(sys.argv[3])
verbose = len(sys.argv) > 4
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = {hdr[i]: 0 for i in cols}
ns = 0
for r in data[a:b + 1]:
    for i in cols: tot[hdr[i]] += int(r[i] or 0)
    ns += int(r[4])
print("samples", ns, {k: round(v / max(ns, 1), 3) for k, v in sorted(tot.items(), key=lambda x: -x[1]) if v})
if verbose:
    for k, r in enumerate(data[a:b + 1]):
        st = {hdr[i][6:]: int(r[i] or 0) for i in cols if int(r[i] or 0) > 0.15 * int(r[4] or 1)}
        print(a + k, r[1][:70].ljust(70), r[4].rjust(7), r[5].rjust(10), r[17].rjust(10), st)

"""Top stall-sample SASS lines of an `ncu --page source --csv` export (first kernel in the file)."""
import csv
import gzip
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
hdr = None
data = []
kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
        print(r[1][:150])
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr) - 2:
        data.append(r)
i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
i_ex = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[i_s] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda k: -int(data[k][i_s] or 0))[:top]
for k in sorted(order):
    r = data[k]
    s = int(r[i_s] or 0)
    why = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print("%5d %5.1f%% exec=%-8s %-70s %s" % (k, 100.0 * s / max(tot, 1), r[i_ex], r[i_src].strip()[:70],
                                             " ".join("%s:%d" % (n, c) for c, n in why if c)))

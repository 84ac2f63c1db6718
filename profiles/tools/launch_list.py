"""Prints the per-kernel times of the LAST 1/k-th of an ncu --metrics gpu__time_duration.sum CSV (one frame / step)."""
import csv, sys
path, k = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
names = [(r[ki][:70], float(r[vi].replace(",", ""))) for r in rows[1:]]
m = len(names) // k
tot = 0.0
for nme, v in names[-m:]:
    print(f"{v/1000:9.1f} us  {nme}"); tot += v
print("total us", round(tot / 1000, 1), "kernels", m)

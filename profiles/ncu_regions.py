"""Instruction / sample shares of the k_knn search by code region and by source line, summed over the captured launches:
   python profiles/ncu_regions.py report.ncu-rep [kernel_name] [queries_per_launch]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_knn"
nq = float(sys.argv[3]) if len(sys.argv) > 3 else 128381.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout.splitlines()
fname = ""
hdr = None
agg = {}
lines = {}
launches = 0
for line in out:
    if line.startswith('"Kernel Name"'):
        launches += 1
    if line.startswith('"File Path"'):
        fname = next(csv.reader([line]))[1].split("/")[-1]
        continue
    if line.startswith('"Line No"'):
        hdr = next(csv.reader([line]))
        continue
    if hdr is None or not line.startswith('"'):
        continue
    r = next(csv.reader([line]))
    if len(r) < len(hdr) or r[2] != "-":
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    try:
        ln = int(r[0]); smp = int(r[ix["# Samples"]]); ins = int(r[ix["Instructions Executed"]])
        th = int(r[ix["Thread Instructions Executed"]])
    except ValueError:
        continue
    src = r[1].strip()
    reg = fname
    a = agg.setdefault(reg, [0, 0, 0]); a[0] += smp; a[1] += ins; a[2] += th
    l = lines.setdefault((fname, ln), [0, 0, 0, src]); l[0] += smp; l[1] += ins; l[2] += th
launches = max(launches, 1)
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"{kern}: {launches} launches, {ti / launches / 1e6:.1f} M warp instructions per launch")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} smp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% lanes {a[2]/max(1,a[1]):5.1f}  thread-inst/query {a[2]/launches/nq:8.1f}")
print()
for (f, ln), l in sorted(lines.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{f}:{ln:4d} smp {100*l[0]/ts:4.1f}% inst {100*l[1]/ti:4.1f}% lanes {l[2]/max(1,l[1]):4.1f} {l[3][:100]}")

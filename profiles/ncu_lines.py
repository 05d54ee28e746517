"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo and --import-source on):
   python profiles/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout.splitlines()
rows = []
fname = ""
hdr = None
first_kernel = None
for line in out:
    if line.startswith('"File Path"'):
        fname = next(csv.reader([line]))[1].split("/")[-1]
        continue
    if line.startswith('"Function Name"'):
        fn = next(csv.reader([line]))[1]
        if first_kernel is None:
            first_kernel = fn
        continue
    if line.startswith('"Line No"'):
        hdr = next(csv.reader([line]))
        continue
    if hdr is None or not line.startswith('"'):
        continue
    r = next(csv.reader([line]))
    if len(r) < len(hdr) or r[2] != "-":      # source-line rows have "-" in the Address column
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    try:
        rows.append((fname, int(r[0]), r[1].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]),
                     int(r[ix["Thread Instructions Executed"]])))
    except ValueError:
        pass
agg = {}
for r in rows:   # the same line of every captured launch (and every inlined copy) is one entry
    a = agg.setdefault((r[0], r[1]), [r[0], r[1], r[2], 0, 0, 0])
    a[3] += r[3]
    a[4] += r[4]
    a[5] += r[5]
rows = [tuple(a) for a in agg.values()]
ts = sum(r[3] for r in rows) or 1
ti = sum(r[4] for r in rows) or 1
tt = sum(r[5] for r in rows)
print(f"{first_kernel}\n  samples {ts}, warp instructions {ti}, lanes/instr {tt / ti:.1f}")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{r[0]:12s}:{r[1]:4d} smp {100 * r[3] / ts:5.1f}%  inst {100 * r[4] / ti:5.1f}%  lanes {r[5] / max(1, r[4]):5.1f}  {r[2][:90]}")

"""Aggregate an .ncu-rep source page per CUDA source line: instructions executed and stall samples."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fpath = ""
rows = []
hdr = None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] in ("Function Name",):
        continue
    if r[0] == "":
        continue  # SASS rows (already summed into the CUDA line row)
    d = dict(zip(hdr, r))
    try:
        rows.append((fpath, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"]), d))
    except ValueError:
        pass
tot_i = sum(x[3] for x in rows) or 1
tot_s = sum(x[4] for x in rows) or 1
print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
stall_keys = [k for k in (hdr or []) if k.startswith("stall_") and "Not Issued" not in k]
for key, title in ((3, "by instructions executed"), (4, "by stall samples")):
    print("\n== top lines", title)
    for x in sorted(rows, key=lambda x: -x[key])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        d = x[5]
        st = sorted(((int(d[k] or 0), k[6:]) for k in stall_keys), reverse=True)[:3]
        print(f"{x[0]:14s}:{x[1]:4d} inst {100 * x[3] / tot_i:5.1f}%  smp {100 * x[4] / tot_s:5.1f}%  "
              f"{' '.join(f'{n}={v}' for v, n in st if v)} | {x[2]}")

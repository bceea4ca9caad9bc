"""Top stalled SASS instructions of one kernel from an .ncu-rep source page.
Usage: python tools/ncu_hot.py rep.ncu-rep <launch-skip> [top]"""
import csv
import io
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:120])
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[hdr.index("# Samples")].isdigit():
        if body:
            break  # a second view (or kernel) starts here
        continue
    body.append(r)
si = hdr.index("# Samples")
src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):7d} {100.0 * int(r[si]) / tot:5.1f}%  {r[src].strip()[:70]:70s} {st}")

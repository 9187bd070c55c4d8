"""Top stall locations (SASS) of one kernel from an ncu report:  python tools/ncu_hot.py REPORT KERNEL_REGEX [N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, ci = h.index("# Samples"), h.index("Source")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[hi + 1:]:
    if len(r) <= si or r[0] == "Address" or r[0] == "Kernel Name":
        break
    try:
        v = float(r[si])
    except ValueError:
        continue
    st = sorted(((float(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    data.append((v, len(data), r[ci].strip()[:90], st))
tot = sum(d[0] for d in data) or 1.0
print(f"{kern}: {int(tot)} samples, {len(data)} instructions")
for v, idx, src, st in sorted(data, reverse=True)[:top]:
    print(f"{100 * v / tot:5.1f}%  #{idx:4d}  {src:90s} {st[0][1]}={int(st[0][0])} {st[1][1]}={int(st[1][0])}")

"""Per-source-line instruction counts and stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv
import sys

cur = None
out = []
for r in csv.reader(open(sys.argv[1])):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or not r[0].isdigit() or r[2] != "-":
        continue
    samples = int(r[4]) if r[4].isdigit() else 0
    inst = int(r[7]) if r[7].isdigit() else 0
    out.append((inst, samples, cur, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out) or 1
ts = sum(o[1] for o in out) or 1
print("total warp instructions", tot, "stall samples", ts)
key = 1 if len(sys.argv) > 2 and sys.argv[2] == "samples" else 0
for o in sorted(out, key=lambda o: -o[key])[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print("%9d %5.1f%%  smp %5.1f%%  %s:%d  %s" % (o[0], 100 * o[0] / tot, 100 * o[1] / ts, o[2], o[3], o[4]))

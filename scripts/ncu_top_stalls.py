"""Top stall-sample instructions of an ncu report (SASS source page): usage python scripts/ncu_top_stalls.py rep [N]"""
import csv, io, subprocess, sys
lines = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
hi = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[hi:]))))
K = "Warp Stall Sampling (All Samples)"
tot = sum(float(r[K] or 0) for r in rows)
idx = sorted(range(len(rows)), key=lambda i: -float(rows[i][K] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 10]
for i in idx:
    print("----", i, "%.1f%%" % (100 * float(rows[i][K]) / tot))
    for j in range(max(0, i - 3), i + 1):
        print("   ", rows[j]["Address"][-5:], rows[j]["Source"].strip(), rows[j][K])

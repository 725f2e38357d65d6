"""Executed-instruction mix of one kernel from an ncu report's SASS source page.

usage: python scripts/ncu_sass_mix.py report.ncu-rep [path_days] [--blocks] [--kernel=REGEX]
Prints warp-level executed instructions by opcode class (and per basic block with --blocks),
divided by `path_days / 32` when given, i.e. instructions per path-day per lane.
"""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
pd = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
ksel = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--kernel=")]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + ksel[0]] if ksel else []),
                     capture_output=True, text=True).stdout
lines = txt.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[hdr_i:]))))
scale = 1.0 / (pd / 32.0) if pd else 1.0
ops = collections.Counter(); stall = collections.Counter()
blocks = []
cur = None
for r in rows:
    src = r["Source"].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0].rstrip(";") if src else "?"
    n = float(r["Instructions Executed"] or 0)
    s = float(r["Warp Stall Sampling (All Samples)"] or 0)
    ops[op.split(".")[0]] += n; stall[op.split(".")[0]] += s
    if cur is None or abs(cur["n"] - n) > 0.02 * max(cur["n"], n, 1):
        cur = {"n": n, "cnt": 0, "tot": 0.0, "stall": 0.0, "first": r["Address"], "ops": collections.Counter()}
        blocks.append(cur)
    cur["cnt"] += 1; cur["tot"] += n; cur["stall"] += s; cur["ops"][op.split(".")[0]] += 1
tot = sum(ops.values()); stot = sum(stall.values())
print(f"total warp instr {tot:.4g}  per path-day {tot*scale:.1f}  samples {stot:.0f}")
for op, n in ops.most_common(40):
    print(f"  {op:12s} {n*scale:10.2f}  {100*n/tot:5.1f}%  stall {100*stall[op]/max(stot,1):5.1f}%")
if "--blocks" in sys.argv:
    print("blocks (static count, executions each, share of instr, share of stall samples):")
    for b in blocks:
        if b["tot"] / tot < 0.004: continue
        top = " ".join(f"{k}:{v}" for k, v in b["ops"].most_common(8))
        print(f"  {b['first'][-6:]} n={b['cnt']:4d} x {b['n']*scale:9.3f}  {100*b['tot']/tot:5.1f}%  stall {100*b['stall']/max(stot,1):5.1f}%  {top}")

"""Copy one measurement pass (gpurun_out/<dir>) into profiles/ under a version tag and print a summary table.

usage: python scripts/collect_profiles.py gpurun_out/final2 v14 [round prefix, default r1] [workload=trials ...]
(round 2: python scripts/collect_profiles.py gpurun_out/r2a v1 r2 c3=113664; the bench default is c3 from round 2 on)
Writes profiles/<round>_<workload>_<tag>_bench.json, ..._ncu_full_summary.csv, ..._sass_mix.txt, the launch list and the
reference-arm line, and refreshes profiles/ncu_constants.json from the captures it finds."""
import json, os, shutil, subprocess, sys

src, tag = sys.argv[1], sys.argv[2]
RND = sys.argv[3] if len(sys.argv) > 3 else "r1"
DEFAULT_W = "c2" if RND == "r1" else "c3"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
TRIALS = {"c2": 100000, "c3": 75776, "c4": 50560, "boot": 50560, "sweep19": 37888, "series19": 100000}
DAYS = {"series19": 2520, "series": 2520}        # simulated days per trial (default 7560)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    TRIALS[k] = int(v)


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


rows = []
for f in sorted(os.listdir(src)):
    if f.startswith("bench_") and f.endswith(".json"):
        w = f[6:-5]
        name = {"default": f"{RND}_{DEFAULT_W}_{tag}_bench.json", "reference": f"{RND}_{DEFAULT_W}_{tag}_reference_arm.json"}.get(w, f"{RND}_{w}_{tag}_bench.json")
        shutil.copy(os.path.join(src, f), os.path.join(prof, name))
        try:
            d = last_json(os.path.join(src, f))
            r = d.get("roofline") or {}
            rows.append((w, d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("achieved"), r.get("unit")))
        except Exception as e:
            rows.append((w, "ERR", str(e)))
for lw in ("c2", "c3"):
    if os.path.exists(os.path.join(src, f"launches_{lw}.csv")):
        shutil.copy(os.path.join(src, f"launches_{lw}.csv"), os.path.join(prof, f"{RND}_{lw}_{tag}_launches.csv"))
consts = {"_comment": "constants read from the committed ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum per "
                      "launch / trials of that launch; smsp__inst_executed.sum x 32 / path-days of that launch); bench.py scales them "
                      "to its launch and copies them into the JSON line"}
KERNEL_RE = {0: "letf_jpair_kernel|letf_path_kernel", 1: "letf_path_kernel"}
for w, n in TRIALS.items():
    rep = os.path.join(src, f"prof_{w}.ncu-rep")
    if not os.path.exists(rep):
        continue
    # a report of the split strategy pipeline holds two kernels: the pair producer and the rule kernel
    parts = []
    for ki in (0, 1):
        out = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_metrics.py"), rep, str(ki)], capture_output=True, text=True).stdout
        if out.strip():
            parts.append(out)
    split = len(parts) == 2
    dram = inst = ms = 0.0
    for ki, out in enumerate(parts):
        first = "_records" if w.startswith("series") else "_producer"
        sfx = (first if ki == 0 else "_rules") if split else ""
        days = DAYS.get(w, 7560)
        with open(os.path.join(prof, f"{RND}_{w}{sfx}_{tag}_ncu_full_summary.csv"), "w") as fh:
            fh.write("metric,unit,value\n" + out)
        kre = (("letf_series_records_kernel" if w.startswith("series") else "letf_jpair_kernel") if ki == 0 else "letf_path_kernel") if split else None
        mix = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_sass_mix.py"), rep, str(n * days), "--blocks"] +
                             ([f"--kernel={kre}"] if kre else []), capture_output=True, text=True).stdout
        with open(os.path.join(prof, f"{RND}_{w}{sfx}_{tag}_sass_mix.txt"), "w") as fh:
            fh.write(mix)
        m = {}
        for line in out.splitlines():
            k, u, v = line.split(",")[:3]
            if k == "kernel":
                continue
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            m[k] = float(v) * scale
        dram += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        inst += m.get("smsp__inst_executed.sum", 0.0)
        ms += m.get("gpu__time_duration.sum", 0)
        print(w, sfx, "ms %.2f" % (m.get("gpu__time_duration.sum", 0)), "instr/path-day %.1f" % (m.get("smsp__inst_executed.sum", 0) * 32 / (n * days)),
              "issue %.1f%%" % m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
              "DRAM GB %.2f" % ((m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / 1e9))
    consts[w] = {"dram_bytes_per_trial": dram / n, "executed_warp_instr_per_path_day": inst * 32 / (n * DAYS.get(w, 7560)),
                 "source": f"profiles/{RND}_{w}" + ("_{producer|records,rules}" if split else "") + f"_{tag}_ncu_full_summary.csv ({n} trials)"}
cpath = os.path.join(prof, "ncu_constants.json")
merged = json.load(open(cpath)) if os.path.exists(cpath) else {}
merged.update(consts)             # workloads not in this pass keep their constants
json.dump(merged, open(cpath, "w"), indent=1)
for r in rows:
    print(r)

"""profiles/r02_kernels_summary.md from an `ncu --set full` report of tools/kernel_zoo.py: per kernel the duration, DRAM bytes and
throughput, L2 hit rate, and the achieved GB/s against the measured HBM peak (MEASURED_PEAKS.json)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, log, out = sys.argv[1], sys.argv[2], sys.argv[3]
alg = {}
for line in open(log, errors="ignore"):
    if line.startswith("ZOO_ALGORITHMIC "):
        alg = json.loads(line[len("ZOO_ALGORITHMIC "):])
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, body = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[col[k]].replace(",", ""))
    except Exception: return float("nan")
agg = {}
for r in body:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("nncf::", "")
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "dram_pct": 0.0, "l2hit": 0.0, "regs": 0, "grid": 0})
    a["n"] += 1
    a["us"] += f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(rows[1][col["gpu__time_duration.sum"]], 1.0)
    def by(k):
        unit = rows[1][col[k]]
        return f(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    a["rd"] += by("dram__bytes_read.sum"); a["wr"] += by("dram__bytes_write.sum")
    a["dram_pct"] = max(a["dram_pct"], f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
    a["l2hit"] += f(r, "lts__t_sector_hit_rate.pct")
    a["regs"] = int(f(r, "launch__registers_per_thread")); a["grid"] = int(f(r, "launch__grid_size"))
lines = ["| kernel | launches | avg us | DRAM MB / launch (read + write) | DRAM GB/s | of HBM peak (%.0f GB/s) | ncu dram %% of peak (max) | L2 hit %% | regs | grid |" % peak,
         "|---|---|---|---|---|---|---|---|---|---|"]
for name, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    us = a["us"] / a["n"]; mb = (a["rd"] + a["wr"]) / a["n"] / 1e6
    gbs = mb / 1e3 / (us * 1e-6) if us > 0 else 0.0
    lines.append("| `%s` | %d | %.1f | %.2f | %.0f | %.2f | %.1f | %.0f | %d | %d |" % (name, a["n"], us, mb, gbs, gbs / peak, a["dram_pct"], a["l2hit"] / a["n"], a["regs"], a["grid"]))
lines += ["", "Algorithmic bytes (tools/kernel_zoo.py):", ""]
for k, v in alg.items():
    lines.append("* `%s`: %s%s" % (k, ("%.1f MB, " % (v["bytes"] / 1e6)) if v.get("bytes") else ("%.1f MB per pass, " % (v["bytes_per_pass"] / 1e6) if v.get("bytes_per_pass") else ""), v["what"]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

"""Turns the ncu reports in gpurun_out/ into the small text/CSV summaries committed under profiles/."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main(tag):
    lines = []
    for name in ("score", "gather", "finalize", "eval"):
        rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, name))
        if not os.path.exists(rep):
            continue
        hdr, units, rows = raw(rep)
        for r in rows:
            kn = r[hdr.index("Kernel Name")]
            lines.append("## %s  (%s)" % (kn, os.path.basename(rep)))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    lines.append("%-70s %s %s" % (k, r[i], units[i]))
            lines.append("")
    open(os.path.join(OUT, "%s_ncu_full_summary.txt" % tag), "w").write("\n".join(lines))
    for f in ("launches_train", "launches_eval"):
        src = os.path.join(ROOT, "gpurun_out", "%s_%s.csv" % (tag, f))
        if not os.path.exists(src):
            continue
        rows = [r for r in csv.reader(open(src)) if len(r) > 5]
        hdr = rows[0]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = {}
        for r in rows[1:]:
            name = r[ik].split("(")[0]
            agg.setdefault(name, []).append(float(r[iv].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(OUT, "%s_%s_summary.csv" % (tag, f)), "w") as fp:
            fp.write("kernel,launches,avg_us,total_us,share\n")
            for name, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
                fp.write("%s,%d,%.3f,%.3f,%.4f\n" % (name, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, sum(v) / tot))
    print(open(os.path.join(OUT, "%s_ncu_full_summary.txt" % tag)).read()[:6000])
    for f in ("launches_train", "launches_eval"):
        p = os.path.join(OUT, "%s_%s_summary.csv" % (tag, f))
        if os.path.exists(p):
            print(open(p).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")

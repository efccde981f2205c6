"""Summarise an .ncu-rep (raw page) into a small table: python tests/perf/ncu_summary.py file.ncu-rep [out.md]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("l1tex__t_bytes.sum", "l1_bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
lines = ["| kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|" + "---|" * len(cols)]
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0][-60:]
    vals = []
    for m, _ in cols:
        if m in ix:
            v = r[ix[m]]
            try:
                f = float(v.replace(",", ""))
                v = ("%.3g" % f) + (" " + units[ix[m]] if units[ix[m]] not in ("", "%") else "")
            except ValueError:
                pass
            vals.append(v)
        else:
            vals.append("-")
    lines.append("| %s | %s |" % (name, " | ".join(vals)))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")

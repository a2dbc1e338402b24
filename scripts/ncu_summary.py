"""Per-kernel summary of an ncu report (--set full): duration, FMA pipe, issue rate, registers, DRAM traffic."""
import csv, subprocess, sys, re
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "dur"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%act"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed", "fma%el"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.per_cycle_active", "warps/SM"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"), ("launch__occupancy_limit_registers", "occR"),
        ("launch__shared_mem_per_block_dynamic", "smem")]
print("%-62s" % "kernel" + " ".join("%9s" % c[1] for c in cols))
units = rows[1]
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("nvf::fast::", "").replace("<unnamed>::", "")[:60]
    vals = []
    for c, _ in cols:
        v = r[ix[c]] if c in ix else ""
        u = units[ix[c]] if c in ix else ""
        try:
            f = float(v.replace(",", ""))
            vals.append(("%9.1f" % f) + "")
            if u: vals[-1] = ("%6.1f%s" % (f, u[:3])).rjust(9)
        except ValueError:
            vals.append("%9s" % v[:9])
    print("%-62s" % name + " ".join(vals))

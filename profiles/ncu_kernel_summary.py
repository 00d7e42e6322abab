"""Per-launch summary of an `ncu --set full` report: duration, DRAM bytes / utilisation, pipe utilisation, issue rate,
cache hit rates and the warp-stall sampling breakdown.   usage: python profiles/ncu_kernel_summary.py <file.ncu-rep> [out.txt]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
M = [("gpu__time_duration.sum", "dur"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
     ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram_busy%"),
     ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy%"),
     ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma%"),
     ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu%"),
     ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
     ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
     ("smsp__inst_executed.sum", "warp_inst"), ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefront%")]
out = []
for r in data:
    g = lambda name: r[hdr.index(name)] if name in hdr else "n/a"
    u = lambda name: units[hdr.index(name)] if name in hdr else ""
    out.append("%s  (id %s)" % (g("Kernel Name")[:60], g("ID")))
    out.append("   " + "  ".join("%s %s%s" % (short, g(name), (" " + u(name)) if u(name) not in ("%", "") else "") for name, short in M))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try: st.append((float(r[i]), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1.0
    out.append("   stall samples: " + ", ".join("%s %.1f%%" % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:7]))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("ncu --set full --clock-control none, report %s\n" % rep.split("/")[-1] + txt + "\n")

"""Builds profiles/r2_ncu_constants.json (read by bench.py for roofline.traffic) from per-launch ncu metric lists of the
accumulate-phase kernels, captured by profiles/run_1gpu_capture.sh with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
usage: python profiles/make_ncu_constants.py <tag>     (reads gpurun_out/<tag>_acc_{g1_n21,g2_n18}.csv)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = {"commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
       "captured_with": "ncu --clock-control none, one MSM after two warm-up MSMs (tests/gpu_one_dev.py), all launches of the accumulate phase"}
for wl in ("g1_n21", "g2_n18"):
    fn = os.path.join(ROOT, "gpurun_out", "%s_acc_%s.csv" % (tag, wl))
    if not os.path.exists(fn):
        continue
    rows = [l for l in open(fn) if l.startswith('"')]
    launches = {}
    for r in csv.DictReader(rows):
        launches.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"].split("(")[0]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = sorted(launches)
    # the LAST MSM of the run: from the last launch of round 0 (ba_round_kernel<.., 1>) or accumulate_kernel to the end
    first = max(i for i in ids if "accumulate_kernel" in launches[i]["kernel"] or ", 1>" in launches[i]["kernel"] or "true" in launches[i]["kernel"])
    sel = [launches[i] for i in ids if i >= first]
    dur = sum(l["gpu__time_duration.sum"] for l in sel)
    pipe = sum(l["gpu__time_duration.sum"] * l["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"] for l in sel) / dur
    out[wl] = {
        "kernels": sorted(set(l["kernel"] for l in sel)), "launches": len(sel), "duration_ms_under_ncu": dur / 1e6,
        "dram_bytes_accumulate_phase": int(sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in sel)),
        "dram_bytes_read": int(sum(l["dram__bytes_read.sum"] for l in sel)), "dram_bytes_write": int(sum(l["dram__bytes_write.sum"] for l in sel)),
        "fmaheavy_pct_accumulate_phase": pipe,
        "per_launch": [{"kernel": l["kernel"], "ms": l["gpu__time_duration.sum"] / 1e6, "dram_read": int(l["dram__bytes_read.sum"]),
                        "dram_write": int(l["dram__bytes_write.sum"]), "fmaheavy_pct": l["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]} for l in sel],
    }
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json"), "w"), indent=1)
print(json.dumps({k: ({kk: vv for kk, vv in v.items() if kk != "per_launch"} if isinstance(v, dict) else v) for k, v in out.items()}, indent=1))

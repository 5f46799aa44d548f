"""Pick the judged metrics out of `ncu -i X.ncu-rep --page raw --csv` (one `--set full` capture), one block per kernel
(the longest capture of each: the polled Krylov path also launches iterations that exit at once on the done flag).
    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/raw.csv; python scripts/ncu_full_summarise.py /tmp/raw.csv "<cmd>" out.txt traffic.json"""
import csv, json, re, sys
path, cmd, out_txt, out_json = sys.argv[1:5]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__inst_executed.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]
with open(path, newline="") as fh:
    rd = list(csv.reader(fh))
hdr, units, rows = rd[0], rd[1], rd[2:]
col = {}
for i, h in enumerate(hdr):
    for w in WANT:
        if h == w or h.endswith("." + w):
            col.setdefault(w, i)
kn = hdr.index("Kernel Name")
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return float("nan")
best = {}
for r in rows:
    name = re.sub(r"^void |\(.*$", "", r[kn].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("vch::", "")
    d = num(r[col["gpu__time_duration.sum"]])
    if name not in best or d > best[name][0]:
        best[name] = (d, r)
traffic = {}
with open(out_txt, "w") as fo:
    fo.write(f"# {cmd}\n# longest capture of each kernel; graphs off so that ncu sees plain launches\n")
    for name, (d, r) in sorted(best.items()):
        fo.write(f"\n== {name}\n")
        for w in WANT:
            if w in col:
                fo.write(f"  {w:76s} {r[col[w]]:>16s} {units[col[w]]}\n")
        rdb, wrb = num(r[col["dram__bytes_read.sum"]]), num(r[col["dram__bytes_write.sum"]])
        scale = lambda v, u: v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic[name] = scale(rdb, units[col["dram__bytes_read.sum"]]) + scale(wrb, units[col["dram__bytes_write.sum"]])
json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (cold cache), 2D 1024^2", "kernels_raw": traffic},
          open(out_json, "w"), indent=1)
print(open(out_txt).read()[:3000])

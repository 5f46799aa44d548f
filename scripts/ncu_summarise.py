"""Fold an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
    python scripts/ncu_summarise.py gpurun_out/launches.csv "<command line that was profiled>" > profiles/rNN_ncu_launch_list.txt
Launches of the polled (graph-less) Krylov path that return at once on the device `done` flag are listed separately."""
import csv, re, sys
from collections import defaultdict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    name = r["Kernel Name"]
    name = re.sub(r"^void |\(.*$", "", name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    name = name.replace("vch::", "")
    if not re.match(r"(dct_|rows16|cols16|bicg_|residual|dmu_|schur|trial|step_setup|adj_|clip_mass|mass_shift|solve_w|set_solve|cost_|grad_prox|mu_init|xbar|halo|lap_|jac_|kkt|energy|publish_scalars|copy_kernel)", name):
        name = "torch (setup: targets, zeros)"
    elif us < 4.5 and re.match(r"(dct_fft_kernel|bicg_x_kernel|rows16_kernel|cols16)", name):
        name += " [exits on done flag]"
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, u in rows:
    agg[n][0] += 1; agg[n][1] += u
print(f"# {cmd}")
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES with the event-timed table of the bench line.")
print(f"{'kernel':52s} {'launches':>9s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}")
for n, (k, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:52s} {k:9d} {u:12.1f} {u / k:9.2f} {u / tot:7.3f}")

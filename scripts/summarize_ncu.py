"""Turns the raw ncu outputs of scripts/collect_evidence.sh into the small summaries kept under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv > profiles/r01_ncu_launch_summary.csv
  python scripts/summarize_ncu.py full gpurun_out/prof_batch_add_r01.ncu-rep > profiles/r01_ncu_full_k_batch_add.csv
  python scripts/summarize_ncu.py traffic profiles/r01_ncu_full_k_batch_add.csv "round 2 of 5" > profiles/ncu_traffic.json
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = OrderedDict()
    nmsm = 0
    rounds = []
    for r in rd:
        name = r["Kernel Name"]
        m = re.search(r"(k_[a-z_]+)", name)
        if not m:
            continue      # torch fill / copy kernels of the harness
        k = m.group(1)
        if k == "k_batch_add":
            k += "<round0>" if re.search(r"false,\s*true>|0, *1>\(", name) or "(bool)1>" in name.split("k_batch_add")[1][:200] and name.rstrip().endswith("1>") else ""
        us = float(r["Metric Value"]) / 1e3
        if k == "k_final":
            nmsm += 1
        d = per.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += us
        if k.startswith("k_batch_add"):
            rounds.append(us)
    harness = {k: per.pop(k) for k in list(per) if k in ("k_random_points", "k_imad", "k_mulbench", "k_set_points")}
    tot = sum(v[1] for v in per.values())
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none) of `bench.py --steps 2 --warmup 1`: %d MSM steps of BLS12-377 2^20, c=16;" % nmsm)
    print("# per-step averages; ncu times are cold-cache and serialised -- compare SHARES with bench.py's phases, not absolutes")
    print("kernel,launches_per_step,us_per_step,share")
    for k, (cnt, us) in per.items():
        print("%s,%.1f,%.1f,%.3f" % (k, cnt / nmsm, us / nmsm, us / tot))
    print("total,,%.1f,1.000" % (tot / nmsm))
    for k, (cnt, us) in harness.items():
        print("# outside the step (set-up / live peak measurement): %s x%d, %.0f us in all" % (k, cnt, us))
    per_msm = len(rounds) // nmsm
    print("# k_batch_add rounds of the last step (us): " + " ".join("%.0f" % u for u in rounds[-per_msm:]))


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none, one launch: %s" % vals[col["Kernel Name"]][:160].replace(",", ";"))
    print("# grid %s block %s" % (vals[col["Grid Size"]], vals[col["Block Size"]]))
    print("metric,value,unit")
    for w in WANT:
        if w in col:
            print("%s,%s,%s" % (w, vals[col[w]].replace(",", ""), units[col[w]]))
    for h in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
        if m and float(vals[col[h]] or 0) >= 0.05:
            print("stall_%s,%s,warps per issue" % (m.group(1), vals[col[h]]))


def traffic(path, which=""):
    """profiles/*_ncu_full_*.csv -> the small JSON bench.py reads for roofline.traffic (DRAM bytes of that one launch)."""
    import json
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
    vals, kernel = {}, ""
    for ln in open(path):
        if ln.startswith("# ncu") and "one launch:" in ln:
            kernel = ln.split("one launch:")[1].strip()
        f = ln.strip().split(",")
        if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            vals[f[0]] = float(f[1]) * scale.get(f[2], 1)
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    print(json.dumps({
        "kernel": "k_batch_add", "launch": which, "source": path, "dram_read_bytes": rd, "dram_write_bytes": wr,
        "bytes_per_launch": rd + wr, "launch_us_under_ncu": vals.get("gpu__time_duration.sum"),
        "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE k_batch_add launch (%s) from the committed ncu --set full capture %s; "
                "roofline.achieved is over the whole MSM (all launches)" % (which, path),
        "kernel_name": kernel[:120]}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

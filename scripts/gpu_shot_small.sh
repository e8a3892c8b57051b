#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "16 bls12-377" "18 pallas" "18 ed-on-bls12-377"; do
  set -- $cfg
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/small_$1_$2.csv python scripts/profile_msm.py $1 3 $2 > /dev/null 2>&1
  echo "== $cfg (last MSM of 3)"
  python - gpurun_out/small_$1_$2.csv <<'PY'
import csv, io, re, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = list(csv.DictReader(io.StringIO("".join(rows))))
names = [(re.search(r"(k_[a-z_0-9]+)", r["Kernel Name"]).group(1) if re.search(r"(k_[a-z_0-9]+)", r["Kernel Name"]) else "?", float(r["Metric Value"]) / 1e3) for r in rd]
# last MSM = from the last k_digits group to the end
idx = max(i for i, (n, _) in enumerate(names) if n == "k_digits")
while idx > 0 and names[idx - 1][0] == "k_digits": idx -= 1
tot = 0
for n, us in names[idx:]:
    print("  %-22s %8.1f us" % (n, us)); tot += us
print("  total %.1f us" % tot)
PY
done

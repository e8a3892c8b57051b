#!/bin/bash
# round 2, third GPU call: pipe utilisation of the field multiplication against resident warps (one or two products in
# flight per warp), default tile sizes after the npairs fix, the twisted-Edwards Horner regression
set -u
mkdir -p gpurun_out
python - > gpurun_out/r2c_microbench.txt 2>&1 <<'PY'
import ctypes, json
from montgomery_b200 import _native
L = _native.lib()
for mode in (4, 6, 13):
    for bps, thr in ((1, 128), (1, 256), (1, 384), (1, 512), (2, 256), (3, 256), (4, 256)):
        ops = ctypes.c_double(); ms = ctypes.c_float()
        rc = L.mgb_microbench(0, mode, bps, thr, 3000, ctypes.byref(ops), ctypes.byref(ms))
        print(json.dumps({"mode": mode, "blocks_per_sm": bps, "threads": thr, "warps_per_smsp": bps * thr / 128, "rc": rc,
                          "G_mults_per_s": round(ops.value / 1e9, 2), "ms": round(ms.value, 3)}))
PY
cat gpurun_out/r2c_microbench.txt
{
for i in 1 2 3; do timeout 60 python scripts/quick_time.py 20; done
MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | tail -6
for cfg in "18 ed-on-bls12-377" "16 bls12-377" "18 pallas"; do timeout 60 python scripts/quick_time.py $cfg; done
MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_winv.so timeout 60 python scripts/quick_time.py 18 ed-on-bls12-377
} > gpurun_out/r2c_times.txt 2>&1
cat gpurun_out/r2c_times.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches_ed.csv \
    python scripts/profile_msm.py 18 2 ed-on-bls12-377 > gpurun_out/r2c_ncu_ed.log 2>&1
grep -E "k_final|k_normalize|k_window|k_tree|k_pair|k_group|k_bucket" gpurun_out/r2c_launches_ed.csv | awk -F'","' '{print $5, $NF}' | tail -24

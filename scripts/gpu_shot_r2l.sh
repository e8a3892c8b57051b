#!/bin/bash
# round 2, GPU call 12: affine bucket reduction (parity + A/B), new window table / CH rule, 2^24 window check
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.txt 2>&1; tail -15 gpurun_out/r2l_pytest.txt
python - > gpurun_out/r2l_affine.txt 2>&1 <<'PY'
import torch, montgomery_b200 as m
from montgomery_b200 import inputs
cv = m.curves.BLS12_377
for logn in (16, 18, 20):
    n = 1 << logn
    eng = m.MsmEngine(cv, 0, n); eng.random_points(n, 1)
    sc = inputs.random_scalars(cv.q, n, 2); d = torch.from_numpy(sc).cuda(); torch.cuda.synchronize()
    for aff in (False, True):
        best = None
        for i in range(6):
            res, tm = eng.msm(None, n=n, device_ptr=d.data_ptr(), affine_reduction=aff)
            if i >= 2 and (best is None or tm["total"] < best["total"]): best = tm
        print("2^%d affine_reduction=%s" % (logn, aff), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()}, hex(res["x"])[:18])
    eng.close()
PY
cat gpurun_out/r2l_affine.txt
{ for cfg in "20 bls12-377" "16 bls12-377" "18 bls12-377" "18 pallas" "21 bls12-377" "23 bls12-377" "24 bls12-377 18" "24 bls12-377 19" "24 bls12-377 20"; do timeout 120 python scripts/quick_time.py $cfg; done; } > gpurun_out/r2l_times.txt 2>&1
cat gpurun_out/r2l_times.txt

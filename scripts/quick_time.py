"""Quick phase timing of one curve/size: python scripts/quick_time.py [logn] [label] [c]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montgomery_b200 as m
from montgomery_b200 import inputs
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
label = sys.argv[2] if len(sys.argv) > 2 else "bls12-377"
c = int(sys.argv[3]) if len(sys.argv) > 3 else None
cv = m.curves.BY_LABEL[label]
n = 1 << logn
eng = m.MsmEngine(cv, 0, n)
eng.random_points(n, 1)
sc = inputs.random_scalars(cv.q, n, 2)
import torch
d = torch.from_numpy(sc).cuda()
best = None
for i in range(6):
    res, tm = eng.msm(None, n=n, c=c, device_ptr=d.data_ptr())
    if i >= 2 and (best is None or tm["total"] < best["total"]):
        best = tm
print(label, logn, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()})

"""Small driver for ncu: a few MSM calls at 2^20 (BLS12-377).  Run under ncu via gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montgomery_b200 as m
from montgomery_b200 import inputs
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
label = sys.argv[3] if len(sys.argv) > 3 else "bls12-377"
cv = m.curves.BY_LABEL[label]
n = 1 << logn
eng = m.MsmEngine(cv, 0, n)
eng.random_points(n, 1)
sc = inputs.random_scalars(cv.q, n, 2)
for i in range(reps):
    res, tm = eng.msm(sc, n=n)
print(tm)

"""Closed-form check + timing at a large size on one GPU: python scripts/check_large.py [logn] [label]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import montgomery_b200 as m
from montgomery_b200 import inputs
from tests.helpers import OracleCurve
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
label = sys.argv[2] if len(sys.argv) > 2 else "bls12-377"
cv = m.curves.BY_LABEL[label]; O = OracleCurve(label)
n = 1 << logn
eng = m.MsmEngine(cv, 0, n)
t0 = time.time(); eng.random_points(n, 4242); print("points generated in %.2f s" % (time.time() - t0), flush=True)
sc = inputs.random_scalars(cv.q, n, 4243)
for i in range(3):
    res, tm = eng.msm(sc, n=n)
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()})
a = inputs.known_dlogs(4242, n)
t0 = time.time()
# sum s_i * a_i with Python ints in chunks (object arrays)
s_ints = np.array(inputs.scalars_to_ints(sc), dtype=object)
k = int(np.dot(s_ints, a.astype(object))) % O.q
print("closed form scalar in %.1f s" % (time.time() - t0))
exp = O.result_of(O.scale(k, O.G))
print("MATCH" if res == exp else "MISMATCH", hex(res["x"])[:24])

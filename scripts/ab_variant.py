"""A/B of an experiment build against the product library on one GPU, in one process per library:
    python scripts/ab_variant.py <lib.so> [logn ...]
Loads the given library (MGB_LIB), checks the warp-cooperative product and two MSMs against the closed form
[(sum s_i a_i) mod q] G of known-dlog points, then prints best-of-6 phase times per size."""
import json, os, sys
lib = sys.argv[1]
os.environ["MGB_LIB"] = os.path.abspath(lib)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, random
import numpy as np
import torch
import montgomery_b200 as m
from montgomery_b200 import _native, inputs
from oracle.params import BLS12_377

out = {"lib": os.path.basename(lib)}
L = _native.lib()
# --- field parity of op 8 (warp-cooperative product)
p, nl = BLS12_377.p, 12
rnd = random.Random(1)
a = [rnd.randrange(p) for _ in range(2049)] + [0, 1, p - 1]
b = [rnd.randrange(p) for _ in range(2049)] + [p - 1, p - 1, p - 1]
A = np.frombuffer(b"".join(v.to_bytes(48, "little") for v in a), dtype=np.uint8).copy()
B = np.frombuffer(b"".join(v.to_bytes(48, "little") for v in b), dtype=np.uint8).copy()
O = np.zeros_like(A)
vp = ctypes.c_void_p
rc = L.mgb_field_op(0, 0, 8, A.ctypes.data_as(vp), B.ctypes.data_as(vp), O.ctypes.data_as(vp), len(a))
got = [int.from_bytes(O[i * 48:(i + 1) * 48].tobytes(), "little") for i in range(len(a))]
out["warp_mul_ok"] = rc == 0 and got == [x * y % p for x, y in zip(a, b)]
for mode in (10, 11):
    ops = ctypes.c_double(); ms = ctypes.c_float()
    L.mgb_microbench(0, mode, 1, 32, 2000, ctypes.byref(ops), ctypes.byref(ms))
    out["ns_per_product_mode%d" % mode] = round(ms.value * 1e6 / 4000, 1)
# --- lane-parallel inverse (op 9; experiment): parity + latency against the one-lane routine (modes 9 / 12, one warp per SM)
O2 = np.zeros_like(A)
rc = L.mgb_field_op(0, 0, 9, A.ctypes.data_as(vp), B.ctypes.data_as(vp), O2.ctypes.data_as(vp), len(a))
got = [int.from_bytes(O2[i * 48:(i + 1) * 48].tobytes(), "little") for i in range(len(a))]
out["warp_inv_ok"] = rc == 0 and got == [pow(x, -1, p) if x else 0 for x in a]
for mode in (9, 12):
    ops = ctypes.c_double(); ms = ctypes.c_float()
    L.mgb_microbench(0, mode, 1, 32, 200, ctypes.byref(ops), ctypes.byref(ms))
    out["us_per_inversion_mode%d" % mode] = round(ms.value * 1e3 / 200, 2)
# --- MSM parity (closed form) + timing
from tests.helpers import OracleCurve


def closed_form(label, n, seed, eng=None, sc=None, res=None):
    """msm over the known-dlog points a_i G == [(sum s_i a_i) mod q] G"""
    cv = m.curves.BY_LABEL[label]
    O = OracleCurve(label)
    if eng is None:
        eng = m.MsmEngine(cv, 0, n)
        eng.random_points(n, seed)
        sc = inputs.random_scalars(cv.q, n, seed + 1)
        res, _ = eng.msm(sc, n=n)
        eng.close()
    a = inputs.known_dlogs(seed, n)
    k = int(np.dot(np.array(inputs.scalars_to_ints(sc), dtype=object), a.astype(object))) % O.q
    return res == O.result_of(O.scale(k, O.G))


for label, logn in (("bls12-377", 12), ("ed-on-bls12-377", 12), ("pallas", 12), ("bls12-381", 12), ("bls12-377", 16)):
    out["closed_form_%s_2^%d" % (label, logn)] = bool(closed_form(label, 1 << logn, 5))
# msmProjective (WeierstrassBasicPolicy: its own k_final / k_window_assemble instances) == batched-affine msm
cv = m.curves.BY_LABEL["bls12-377"]
eng = m.MsmEngine(cv, 0, 1 << 12)
eng.random_points(1 << 12, 9)
sc = inputs.random_scalars(cv.q, 1 << 12, 10)
out["projective_equals_affine"] = eng.msm(sc, n=1 << 12)[0] == eng.msm(sc, n=1 << 12, projective=True)[0]
eng.close()
for logn in [int(x) for x in sys.argv[2:]] or [20]:
    cv = m.curves.BY_LABEL["bls12-377"]
    n = 1 << logn
    eng = m.MsmEngine(cv, 0, n)
    eng.random_points(n, 1)
    sc = inputs.random_scalars(cv.q, n, 2)
    d = torch.from_numpy(sc).cuda()
    best = None
    for i in range(8):
        res, tm = eng.msm(None, n=n, device_ptr=d.data_ptr())
        if i >= 2 and (best is None or tm["total"] < best["total"]):
            best = tm
    out["2^%d" % logn] = {k: round(v, 4) for k, v in best.items() if isinstance(v, float)}
    out["2^%d_closed_form" % logn] = bool(closed_form("bls12-377", n, 1, eng, sc, res))
    eng.close()
print(json.dumps(out))

"""A torch-free, one-minute GPU check of the last changes of round 2 (run by hand: `python tests/gpu_final_check.py`;
pytest does not collect it -- the same checks live in tests/test_gpu_round2.py / test_gpu_api.py, which import torch and
take longer on a fresh box).  Through the C ABI (ctypes), checked against the closed form of known-dlog points:

  * smoke() of __graft_entry__ (oracle comparison on two curves);
  * byte ingestion in chunks (mgb_set_points above 2^18 points, and with a forced small chunk);
  * the sharded entry point on one rank (normalisation fused into k_final) and mgb_combine_partials (k_normalize);
  * the headline MSM (BLS12-377, 2^20): parity and device time.

Prints one JSON line per check.
"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry
import montgomery_b200 as m
from montgomery_b200 import inputs
from tests.helpers import OracleCurve

t00 = time.time()
SMALL = bool(os.environ.get("MGB_FINAL_CHECK_SMALL"))       # dry run of this script on the emulated host build (MGB_LIB=...)


def closed_form(label, seed, sc):
    O = OracleCurve(label)
    return O.result_of(O.scale(inputs.dot_known_dlogs(sc, inputs.known_dlogs(seed, sc.shape[0])) % O.q, O.G))


def report(name, ok, **kw):
    print(json.dumps({"check": name, "ok": bool(ok), "t_s": round(time.time() - t00, 1), **kw}), flush=True)
    return ok


ok = True
entry.smoke()
ok &= report("smoke", True)

for label in ("bls12-377", "ed-on-bls12-377"):
    cv = m.curves.BY_LABEL[label]
    n = 150 if SMALL else (1 << 18) + 4097
    a, b = m.MsmEngine(cv, 0, n), m.MsmEngine(cv, 0, n)
    a.random_points(n, seed=70)
    xy, z = a.get_points(0, n)
    sc = inputs.random_scalars(cv.q, n, 71)
    exp = closed_form(label, 70, sc)
    good = a.msm(sc)[0] == exp
    for chunk in (None, "13" if SMALL else "10007"):
        if chunk:
            os.environ["MGB_DEBUG_INGEST_CHUNK"] = chunk
        t0 = time.time()
        b.set_points(xy.reshape(-1), z if cv.kind == "weierstrass" else None)
        dt = time.time() - t0
        back, bz = b.get_points(0, n)
        good &= bool(np.array_equal(back, xy) and np.array_equal(bz, z)) and b.msm(sc)[0] == exp
        b.random_points(16, seed=1)
        ok &= report("chunked_ingestion", good, curve=label, chunk=chunk or "default", n=n, set_points_ms=round(dt * 1e3, 2))
    os.environ.pop("MGB_DEBUG_INGEST_CHUNK", None)
    # sharded entry point on one rank, its empty-shard case (k_normalize on one partial), and two partial sums combined by
    # k_normalize (device buffer from cudaMalloc through ctypes: no torch in this script)
    res = a.msm_sharded(sc.ctypes.data, False, n)[0]
    neutral = a.msm_sharded(None, False, 0)[0]
    ok &= report("sharded_one_rank", res == exp and neutral["isZero"], curve=label)
    pb = a.partial_bytes
    dbuf = ctypes.c_void_p()
    if SMALL:
        keep = np.zeros(2 * pb + 64, np.uint8)
        dbuf.value = (keep.ctypes.data + 63) // 64 * 64
    else:
        rt = ctypes.CDLL("libcudart.so.12")
        assert rt.cudaMalloc(ctypes.byref(dbuf), ctypes.c_size_t(2 * pb)) == 0
    h = n // 2
    a.msm_partial(sc.ctypes.data, False, h, dbuf.value)
    b.set_points(xy[h:].reshape(-1), z[h:] if cv.kind == "weierstrass" else None)
    b.msm_partial(sc[h:].ctypes.data, False, n - h, dbuf.value + pb)
    ok &= report("partials_combined", a.combine_partials(dbuf.value, 2) == exp, curve=label)
    if not SMALL:
        rt.cudaFree(dbuf)
    a.close()
    b.close()

cv = m.curves.BLS12_377
n = 200 if SMALL else 1 << 20
eng = m.MsmEngine(cv, 0, n)
eng.random_points(n, seed=3)
sc = inputs.random_scalars(cv.q, n, 4)
best = None
for i in range(1 if SMALL else 5):
    res, tm = eng.msm(sc)
    best = tm if best is None or tm["total"] < best["total"] else best
ok &= report("headline_2p20", res == closed_form("bls12-377", 3, sc), device_ms=round(best["total"], 3),
             phases={k: round(best[k], 3) for k in ("decompose_slice", "sort", "accumulate", "reduce", "final_sum")})
eng.close()
print("FINAL", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)

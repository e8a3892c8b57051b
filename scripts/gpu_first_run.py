"""First GPU run: microbenchmarks + quick MSM timings -> gpurun_out/first_run.json"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import montgomery_b200 as m
from montgomery_b200 import _native, inputs
lib = _native.lib()
os.makedirs("gpurun_out", exist_ok=True)
out = {"microbench": [], "msm": []}
names = {0: "mad.lo.u32", 1: "mad.hi.u32", 2: "mad.wide.u32", 3: "wide carry chain", 4: "Fp377 mul (call)", 5: "Fr377 mul (call)", 6: "Fp377 mul (inline)", 7: "Fr377 mul (inline)"}
for mode in range(8):
    for bps, thr in ((1, 256), (2, 256), (4, 256), (2, 512), (2, 1024), (8, 256)):
        iters = 2000 if mode < 4 else 400
        ops = ctypes.c_double(); ms = ctypes.c_float()
        rc = lib.mgb_microbench(0, mode, bps, thr, iters, ctypes.byref(ops), ctypes.byref(ms))
        rec = {"mode": mode, "name": names[mode], "blocks_per_sm": bps, "threads": thr, "rc": rc, "ops_per_s": ops.value, "ms": ms.value}
        out["microbench"].append(rec); print(rec, flush=True)
for label, cv, logn in (("bls12-377", m.curves.BLS12_377, 16), ("bls12-377", m.curves.BLS12_377, 20), ("pallas", m.curves.PALLAS, 18), ("ed-on-bls12-377", m.curves.ED_ON_BLS12_377, 18)):
    n = 1 << logn
    eng = m.MsmEngine(cv, 0, n)
    t0 = time.time(); eng.random_points(n, 1); t1 = time.time()
    sc = inputs.random_scalars(cv.q, n, 2)
    for c in (None,):
        for it in range(3):
            res, tm = eng.msm(sc, n=n, c=c)
        rec = {"curve": label, "logn": logn, "gen_points_s": t1 - t0, "timing": tm, "x": hex(res["x"])}
        out["msm"].append(rec); print(rec, flush=True)
    eng.close()
json.dump(out, open("gpurun_out/first_run.json", "w"), indent=1)

"""Integer-pipe and field-multiplication microbenchmarks -> JSON lines (run on the GPU box)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montgomery_b200 import _native
lib = _native.lib()
names = {0: "mad.lo.u32 (IMAD)", 1: "mad.hi.u32 (IMAD.HI)", 2: "mad.wide.u32 64-bit accumulate", 3: "IMAD.WIDE.U32.X carry chain",
         4: "Fp377 mul (call)", 5: "Fr377 mul (call)", 6: "Fp377 mul (inline)", 7: "Fr377 mul (inline)"}
for mode in range(8):
    ops = ctypes.c_double(); ms = ctypes.c_float()
    rc = lib.mgb_microbench(0, mode, 2, 1024, 2000 if mode < 4 else 400, ctypes.byref(ops), ctypes.byref(ms))
    print(json.dumps({"mode": mode, "name": names[mode], "rc": rc, "ops_per_s": ops.value, "ms": ms.value}))
# latency of ONE dependent chain of Fp377 products on a warp that has its scheduler to itself (one 32-thread block per SM)
lat = {10: "Fp377 product latency, lane 0 alone (Field::mul)", 11: "Fp377 product latency, warp-cooperative (WarpField::mul)"}
for mode, name in lat.items():
    ops = ctypes.c_double(); ms = ctypes.c_float()
    iters = 2000
    rc = lib.mgb_microbench(0, mode, 1, 32, iters, ctypes.byref(ops), ctypes.byref(ms))
    print(json.dumps({"mode": mode, "name": name, "rc": rc, "ns_per_product": ms.value * 1e6 / (2 * iters), "ms": ms.value}))

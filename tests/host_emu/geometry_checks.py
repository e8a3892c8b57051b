"""Manual tool (run by scripts/asan_emulated.sh): the PRODUCTION geometries on the emulated host build, meant for the
AddressSanitizer build -- small inputs, but the bucket counts, reduction geometry and tile sizes of the large configs,
so that every buffer size in msm.cu is checked against the kernels' accesses at the shapes the GPU runs:

  * window sizes 16 (2^18 .. 2^21 points: three 5-bit digits, 256 partial sums per group, 2.6e5 buckets) and, with
    --c18, 18 (2^22 points and more: four digits, 2048 partial sums per group, 1e6 buckets; 26 minutes under ASan);
  * tile sizes of the accumulation as the 2^20 plan picks them (E = 56 / 28 / 14 pairs per lane) and the maximum (64).

    python tests/host_emu/geometry_checks.py path/to/libmgb_emu.so [--c18]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tests.test_host_emu_pipeline as T  # noqa: E402
from montgomery_b200 import inputs  # noqa: E402

host = T.EmuHost(sys.argv[1])
ok = True
for c in [16] + ([18] if "--c18" in sys.argv else []):
    ctx = host.create("pallas", 64)
    pts = ctx.random_points(48, seed=5)
    sc = inputs.random_scalars(ctx.cv.q, 48, 6)
    t = time.time()
    res, tm = ctx.msm(sc, c=c)
    good = res == T.oracle_msm("pallas", sc, pts)
    ok &= good
    print("window", c, "ok" if good else "MISMATCH", "%.0f s" % (time.time() - t), {k: tm[k] for k in ("K", "rounds", "n_launches")}, flush=True)
    ctx.close()
ctx = host.create("bls12-377", 320)
pts = ctx.random_points(300, seed=15)
sc = inputs.random_scalars(ctx.cv.q, 300, 16)
exp = T.oracle_msm("bls12-377", sc, pts)
os.environ["MGB_DEBUG_NROUNDS"] = "3"
for E in ("56,28,14", "64,64,64", "7,5,4"):
    os.environ["MGB_DEBUG_E"] = E
    res, tm = ctx.msm(sc, c=6)
    ok &= res == exp
    print("tile sizes", E, "ok" if res == exp else "MISMATCH", flush=True)
ctx.close()
sys.exit(0 if ok else 1)

"""GPU tests added in round 2 (all through the C ABI, bit-exact against the oracle):

  * the plain-C host of bindings/c/example_msm.c RUN on the GPU and its printed point checked (a second host language
    on the boundary, not only ctypes);
  * scalars outside the range a path accepts: the reference traps (src/wasm/glv.ts:131,158) -- here the GLV paths
    accept any 256-bit value and return (s mod q) P, the paths without decomposition reject with MGB_E_INVALID;
  * every PointSet handle owns its point table (src/parallel.ts:97-116: every pointPtr is its own memory);
  * n is validated against the scalar buffer;
  * the sharded entry point on one rank, empty shards, and -- when the box has >= 2 GPUs -- the real thing:
    one process per GPU under torchrun (NCCL communicator owned by the context) and one process driving all GPUs;
  * closed form at 2^22 (and 2^24 with MGB_TEST_2P24=1): sizes the reference cannot run (src/field-msm.ts:55-56).
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import montgomery_b200 as m
from montgomery_b200 import _native, inputs
from montgomery_b200.api import MsmError
from tests.helpers import OracleCurve, scalars_to_bytes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CURVES = {"bls12-377": m.curves.BLS12_377, "pallas": m.curves.PALLAS, "ed-on-bls12-377": m.curves.ED_ON_BLS12_377,
          "bls12-381": m.curves.BLS12_381}


def closed_form(label, seeds_and_scalars):
    """[(sum over shards of sum_i s_i a_i) mod q] G for known-dlog shards [(seed, (n, 32) uint8 scalars), ...]"""
    O = OracleCurve(label)
    k = 0
    for seed, sc in seeds_and_scalars:
        k += inputs.dot_known_dlogs(sc, inputs.known_dlogs(seed, sc.shape[0]))
    return O.result_of(O.scale(k % O.q, O.G))


def test_c99_host_runs_on_gpu(tmp_path):
    """bindings/c/example_msm.c compiled with gcc -std=c99 and run here: random points a_i G (seed 0x6d6f6e74), scalars from
    the program's own LCG; the point it prints must be the closed form."""
    lib_dir = os.path.dirname(_native.LIB_PATH)
    exe = str(tmp_path / "example_msm")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "bindings", "c", "example_msm.c"), "-L", lib_dir, "-lmontgomery_b200",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    logn = 12
    res = subprocess.run([exe, str(logn)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    x = int(re.search(r"x = 0x([0-9a-f]+)", res.stdout).group(1), 16)
    y = int(re.search(r"y = 0x([0-9a-f]+)", res.stdout).group(1), 16)
    assert "isZero = 0" in res.stdout
    n = 1 << logn
    sc = np.zeros(32 * n, dtype=np.uint8)          # fill_scalars() of the C program
    seed = 1
    for i in range(32 * n):
        seed = (seed * 6364136223846793005 + 1442695040888963407) & (2**64 - 1)
        sc[i] = seed >> 56
        if (i & 31) == 31:
            sc[i] &= 0x0F
    assert closed_form("bls12-377", [(0x6D6F6E74, sc.reshape(n, 32))]) == {"x": x, "y": y, "isZero": False}


def test_scalars_out_of_range():
    n = 64
    # GLV paths: any 32-byte value is a valid scalar, the result is (s mod q) P
    for label in ("bls12-377", "pallas", "bls12-381"):
        cv = CURVES[label]
        O = OracleCurve(label)
        eng = m.MsmEngine(cv, 0, n)
        try:
            eng.random_points(n, seed=3)
            a = inputs.known_dlogs(3, n)
            big = [cv.q, cv.q + 5, 2 * cv.q + 1, (1 << 256) - 1, (1 << 255) + 12345, cv.q - 1, 0, 1] * (n // 8)
            res, _ = eng.msm(scalars_to_bytes(big), n=n)
            k = sum((s % cv.q) * int(x) for s, x in zip(big, a)) % cv.q
            assert res == O.result_of(O.scale(k, O.G)), label
            # msmProjective takes the raw scalar: fine below 2^bits(q), rejected above
            ok = [cv.q + 5 if (cv.q + 5).bit_length() <= cv.q.bit_length() else cv.q - 1] * n
            res, _ = eng.msm(scalars_to_bytes(ok), n=n, projective=True)
            k = sum((s % cv.q) * int(x) for s, x in zip(ok, a)) % cv.q
            assert res == O.result_of(O.scale(k, O.G)), label
            bad = [1] * n
            bad[17] = 1 << cv.q.bit_length()
            with pytest.raises(MsmError) as ei:
                eng.msm(scalars_to_bytes(bad), n=n, projective=True)
            assert ei.value.code == _native.E_INVALID and "out of range" in str(ei.value)
            assert eng.msm(scalars_to_bytes([1] * n), n=n)[0] == O.result_of(O.scale(sum(int(x) for x in a), O.G))   # still usable
        finally:
            eng.close()
    # twisted Edwards (no decomposition): scalars must be below 2^251
    cv = CURVES["ed-on-bls12-377"]
    O = OracleCurve("ed-on-bls12-377")
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=4)
        a = inputs.known_dlogs(4, n)
        for s_bad in (1 << 251, (1 << 256) - 1, 1 << 252):
            bad = [2] * n
            bad[n - 1] = s_bad
            with pytest.raises(MsmError) as ei:
                eng.msm(scalars_to_bytes(bad), n=n)
            assert ei.value.code == _native.E_INVALID and "out of range" in str(ei.value)
        top = [(1 << 251) - 1, cv.q - 1, cv.q, cv.q + 1] * (n // 4)          # in range, some >= q
        res, _ = eng.msm(scalars_to_bytes(top), n=n)
        k = sum(s * int(x) for s, x in zip(top, a)) % cv.q
        assert res == O.result_of(O.scale(k, O.G))
    finally:
        eng.close()


def test_point_sets_own_their_tables():
    """Two live PointSets of one curve module never alias (the advisor's round-1 finding): creating the second set
    must not change what msm over the first one returns."""
    cv = CURVES["bls12-377"]
    mod = m.Weierstrass.create(cv)
    P = mod.Parallel
    pts1 = P.randomPointsFast(64, seed=11)
    sc = P.randomScalars(64, seed=12)
    r1 = P.msm(sc, pts1, 64)["result"]
    pts2 = P.randomPointsFast(200, seed=13)                  # larger than the first engine's capacity
    assert pts2.engine is not pts1.engine
    assert P.msm(sc, pts1, 64)["result"] == r1 == closed_form("bls12-377", [(11, sc)])
    sc2 = P.randomScalars(200, seed=14)
    assert P.msm(sc2, pts2, 200)["result"] == closed_form("bls12-377", [(13, sc2)])
    with pytest.raises(ValueError):
        P.msm(sc2, pts1, 200)                                # more scalars than points in the set
    eng2 = pts2.engine
    pts2.close()
    with pytest.raises(MsmError):
        P.msm(sc2, pts2, 200)                                # closed handle
    pts3 = P.randomPointsFast(100, seed=15)                  # the released engine is reused
    assert pts3.engine is eng2
    sc3 = P.randomScalars(100, seed=16)
    assert P.msm(sc3, pts3, 100)["result"] == closed_form("bls12-377", [(15, sc3)])
    assert P.msm(sc, pts1, 64)["result"] == r1


def test_n_is_validated_against_the_scalar_buffer():
    cv = CURVES["bls12-377"]
    eng = m.MsmEngine(cv, 0, 256)
    try:
        eng.random_points(256, seed=1)
        sc = inputs.random_scalars(cv.q, 100, 2)
        with pytest.raises(ValueError):
            eng.msm(sc, n=256)                               # would read 32 * 156 bytes past the host array
        d = torch.from_numpy(sc).cuda()
        torch.cuda.synchronize()
        with pytest.raises(ValueError):
            eng.msm(None, device_ptr=d.data_ptr())           # n unknown
        with pytest.raises(ValueError):
            eng.msm(None, n=10, device_ptr=d.data_ptr() + 4)  # misaligned
        assert eng.msm(None, n=100, device_ptr=d.data_ptr())[0] == eng.msm(sc)[0]
    finally:
        eng.close()


@pytest.mark.parametrize("label", list(CURVES))
def test_sharded_entry_point_on_one_rank_and_empty_shards(label):
    cv = CURVES[label]
    n = 1000
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=21)
        sc = inputs.random_scalars(cv.q, n, 22)
        whole, _ = eng.msm(sc, n=n)
        assert eng.comm_info()["world"] == 1
        got, tm = eng.msm_sharded(sc.ctypes.data, False, n)                      # no communicator: the plain MSM
        assert got == whole and tm["total"] > 0
        neutral, _ = eng.msm_sharded(0, False, 0)                                # an empty shard is the neutral element
        assert neutral["isZero"]
        pb = eng.partial_bytes
        parts = torch.zeros(3 * pb, dtype=torch.uint8, device="cuda")
        d = torch.from_numpy(sc).cuda()
        torch.cuda.synchronize()
        eng.msm_partial(0, True, 0, parts.data_ptr())                            # shard 0: empty
        eng.msm_partial(d.data_ptr(), True, n, parts.data_ptr() + pb)            # shard 1: everything
        eng.msm_partial(0, False, 0, parts.data_ptr() + 2 * pb)                  # shard 2: empty
        assert eng.combine_partials(parts.data_ptr(), 3) == whole
    finally:
        eng.close()


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_one_process_multi_device_api(label):
    """mgb_multi_*: one host process, one context per device, contiguous shards.  With one visible GPU this runs the
    single-device degenerate case (no NCCL); with two or more, the communicators of ncclCommInitAll."""
    cv = CURVES[label]
    O = OracleCurve(label)
    ndev = min(torch.cuda.device_count(), 4)
    for devs in ([0], list(range(ndev))) if ndev > 1 else ([0],):
        G = len(devs)
        n = 5000 + 3                                                             # not a multiple of the device count
        mg = m.MultiGpuMsm(cv, devs, 8192)
        try:
            mg.random_points(n, seed=40)
            per = -(-n // G)
            sc = inputs.random_scalars(cv.q, n, 41)
            shards = [(40 + g, sc[g * per:min(n, (g + 1) * per)]) for g in range(G)]
            res, tm = mg.msm(sc)
            assert res == closed_form(label, shards), (label, devs)
            # prefix of the pairs: later shards are partly / completely empty
            k = per + 7 if G > 1 else 77
            part = [(40 + g, sc[g * per:min(k, (g + 1) * per)]) for g in range(G) if g * per < k]
            assert mg.msm(sc, n=k)[0] == closed_form(label, part)
            # read-back across the shard boundary, and byte ingestion of the same points gives the same sum
            xy, z = mg.get_points(0, n)
            a0 = inputs.known_dlogs(40, 4)
            cb = cv.coord_bytes
            assert (int.from_bytes(xy[1, :cb].tobytes(), "little"), int.from_bytes(xy[1, cb:].tobytes(), "little")) == O.scale(int(a0[1]), O.G)
            mg.set_points(xy.reshape(-1), z if cv.kind == "weierstrass" else None)
            assert mg.msm(sc)[0] == res
        finally:
            mg.close()


def test_failed_shard_is_reported_not_hung():
    """mgb_msm_sharded / mgb_multi_msm: a device whose shard holds an out-of-range scalar joins the all-gather flagged; the
    call returns that device's error (not a hang, not a sum without the shard) and the handle stays usable.  Needs two
    GPUs for the collective; the CPU suite runs the same scenario on the emulated host (tests/test_host_emu_pipeline.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    label = "ed-on-bls12-377"
    cv = CURVES[label]
    n = 4096
    mg = m.MultiGpuMsm(cv, [0, 1], n)
    try:
        mg.random_points(n, seed=60)
        sc = inputs.random_scalars(cv.q, n, 61)
        bad = sc.copy()
        bad[n - 5, 31] = 0xFF                                                    # in device 1's shard
        with pytest.raises(MsmError) as ei:
            mg.msm(bad)
        assert ei.value.code == _native.E_INVALID and "device 1" in str(ei.value) and "out of range" in str(ei.value)
        per = n // 2
        assert mg.msm(sc)[0] == closed_form(label, [(60, sc[:per]), (61, sc[per:])])
    finally:
        mg.close()


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_chunked_point_ingestion(label, monkeypatch):
    """mgb_set_points above 2^18 points: chunks through two staging halves, the copy of the next chunk overlapping the
    conversion of the current one.  Points made on the device are read back as bytes, fed to a second
    context, and must give the same table (byte-identical read-back) and the same, closed-form-checked, MSM -- with the
    default chunk (two chunks, the last one partial) and with 27 small ones."""
    cv = CURVES[label]
    n = (1 << 18) + 4097
    a = m.MsmEngine(cv, 0, n)
    b = m.MsmEngine(cv, 0, n)
    try:
        a.random_points(n, seed=70)
        xy, z = a.get_points(0, n)
        sc = inputs.random_scalars(cv.q, n, 71)
        exp = closed_form(label, [(70, sc)])
        assert a.msm(sc)[0] == exp
        for chunk in (None, "10007"):
            if chunk:
                monkeypatch.setenv("MGB_DEBUG_INGEST_CHUNK", chunk)
            assert b.set_points(xy.reshape(-1), z if cv.kind == "weierstrass" else None) == n
            back, bz = b.get_points(0, n)
            assert np.array_equal(back, xy) and np.array_equal(bz, z)
            assert b.msm(sc)[0] == exp
            b.random_points(16, seed=1)                                          # scribble over the head of the table between the passes
    finally:
        a.close()
        b.close()


def test_multi_gpu_one_process_per_gpu():
    """The bench's multi-GPU path: torchrun, one process per GPU, communicator owned by the context (mgb_comm_init),
    mgb_msm_sharded; closed form over all ranks' shards on three curves, plus ranks with empty shards."""
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ndev < 4 else 4
    port = 29700 + os.getpid() % 200
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "check_multi_gpu.py"), "14"],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("multi-gpu ok") == 3 and "empty shards ok" in res.stdout


@pytest.mark.parametrize("logn", [22] + ([24] if os.environ.get("MGB_TEST_2P24") else []))
def test_closed_form_beyond_the_reference_limit(logn):
    """BASELINE config 5 is 2^24 pairs; the reference cannot hold more than ~2^21 points in its 4 GiB wasm memory
    (src/field-msm.ts:55-56), so nothing pins these sizes but the closed form.  2^24 runs with MGB_TEST_2P24=1
    (scripts/gpu_strong_2p24.sh records it)."""
    cv = CURVES["bls12-377"]
    n = 1 << logn
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=4242)
        sc = inputs.random_scalars(cv.q, n, 4243)
        res, tm = eng.msm(sc, n=n)
        assert res == closed_form("bls12-377", [(4242, sc)])
        assert tm["n_pairs"] > 10 * n
    finally:
        eng.close()


@pytest.mark.parametrize("label", ["bls12-377", "pallas", "bls12-381"])
def test_affine_bucket_reduction_equals_default(label):
    """mgb_opts.affine_reduction (the reference's reduceBucketsAffine idea): group sums by batched-affine
    trees instead of XYZZ additions -- same canonical point as the default reduction and as the oracle, for several
    sizes, window sizes (incl. a sparse, sub-bucket-spread top window) and degenerate inputs."""
    cv = CURVES[label]
    O = OracleCurve(label)
    n = 1 << 13
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=60)
        sc = inputs.random_scalars(cv.q, n, 61)
        for k, c in [(n, None), (n, 7), (n, 11), (n, 14), (3000, 9), (257, 5), (64, None), (1, None)]:
            ref, _ = eng.msm(sc[:k], n=k, c=c)
            got, tm = eng.msm(sc[:k], n=k, c=c, affine_reduction=True)
            assert got == ref, (label, k, c)
        assert got == O.result_of(O.scale(int.from_bytes(sc[0].tobytes(), "little") * int(inputs.known_dlogs(60, 1)[0]), O.G))
        assert closed_form(label, [(60, sc)]) == eng.msm(sc, n=n, affine_reduction=True)[0]
        # all scalars equal (one huge bucket per window, everything else empty), zeros, and P + (-P) inside a group
        same = np.repeat(sc[:1], n, axis=0)
        assert eng.msm(same, n=n, affine_reduction=True)[0] == eng.msm(same, n=n)[0]
        zeros = np.zeros((n, 32), dtype=np.uint8)
        assert eng.msm(zeros, n=n, affine_reduction=True)[0]["isZero"]
        q1 = scalars_to_bytes([1, cv.q - 1] * (n // 2))       # pairs s, -s over different points: no cancellation expected, just signs
        assert eng.msm(q1, n=n, affine_reduction=True)[0] == eng.msm(q1, n=n)[0]
    finally:
        eng.close()
    # the twisted-Edwards engine has no batched-affine path: the flag is accepted and ignored
    cv = CURVES["ed-on-bls12-377"]
    eng = m.MsmEngine(cv, 0, 512)
    try:
        eng.random_points(512, seed=62)
        sc = inputs.random_scalars(cv.q, 512, 63)
        assert eng.msm(sc, affine_reduction=True)[0] == eng.msm(sc)[0]
    finally:
        eng.close()


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_identical_points_and_scalars(label):
    """Every pair is (s, P) with the same s and the same P: every addition of the bucket trees, of the leftover sums
    (affine + affine with equal operands: the doubling branch of mmadd) and of the reduction is a doubling or meets
    equal operands; the result is [n s] P."""
    cv = CURVES[label]
    O = OracleCurve(label)
    n = 5000
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(1, seed=70)
        xy, z = eng.get_points(0, 1)
        eng.set_points(np.tile(xy.reshape(-1), n), np.zeros(n, dtype=np.uint8) if cv.kind == "weierstrass" else None)
        s = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCD % cv.q
        sc = scalars_to_bytes([s] * n)
        P = O.scale(int(inputs.known_dlogs(70, 1)[0]), O.G)
        exp = O.result_of(O.scale(n * s, P))
        for c in (None, 8, 12):
            assert eng.msm(sc, n=n, c=c)[0] == exp, (label, c)
    finally:
        eng.close()


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_prefetched_scalars(label):
    """mgb_msm_prefetch: scalar sets uploaded ahead of their MSM give the same point as the plain call; a set that was
    prefetched but is not the one passed is simply not used; at most two sets wait; the sharded entry point consumes
    them too (include/montgomery_b200.h)."""
    cv = CURVES[label]
    n = 3000
    eng = m.MsmEngine(cv, 0, 4096)
    try:
        eng.random_points(4096, seed=21)
        sets = [torch.from_numpy(inputs.random_scalars(cv.q, n, 30 + i)).pin_memory() for i in range(4)]
        plain = [eng.msm(s.numpy())[0] for s in sets]
        assert plain[0] == closed_form(label, [(21, sets[0].numpy())])
        # the pipeline of bench.py: the next set travels while the current MSM runs
        eng.prefetch(sets[0].numpy())
        got = []
        for i in range(4):
            if i + 1 < 4:
                eng.prefetch(sets[i + 1].numpy())
            got.append(eng.msm(sets[i].numpy())[0])
        assert got == plain
        # a prefetched set that is not the one passed (other pointer, or other n) stays put; the call uploads its own
        eng.prefetch(sets[0].numpy())
        assert eng.msm(sets[1].numpy())[0] == plain[1]
        assert eng.msm(sets[0].numpy(), n=n - 1)[0] == eng.msm(sets[0].numpy()[: n - 1].copy())[0]
        assert eng.msm(sets[0].numpy())[0] == plain[0]                     # ... and is still there for its own call
        # the same buffer again with new contents: uploaded again
        buf = sets[3].numpy()
        eng.prefetch(buf)
        buf[:] = sets[2].numpy()
        eng.prefetch(buf)
        assert eng.msm(buf)[0] == plain[2]
        # two sets may wait, a third is refused
        eng.prefetch(sets[0].numpy())
        eng.prefetch(sets[1].numpy())
        with pytest.raises(MsmError):
            eng.prefetch(sets[2].numpy())
        assert eng.msm(sets[1].numpy())[0] == plain[1]
        assert eng.msm_sharded(sets[0].data_ptr(), False, n)[0] == plain[0]
        with pytest.raises(MsmError):
            eng.prefetch(sets[0].data_ptr(), n=4097)                       # more than the context holds (raw address: the library checks)
    finally:
        eng.close()

"""CPU test of the WHOLE path behind the C ABI: montgomery_b200/csrc/msm.cu -- context and buffer management, window /
reduction-geometry choice, round planning, every kernel launch in its shipped order, error paths -- compiled with g++
against a stand-in CUDA runtime (tests/host_emu/cuda_rt_emu.h: "device" memory is host memory, a launch runs the grid
block after block as lockstep host threads, tests/host_emu/cuda_emu.h) and driven exactly as a C host drives the
product: mgb_create -> mgb_set_points / mgb_random_points -> mgb_msm, results bit-exact against the oracle.

The only edit to the shipped source is mechanical and done here at build time (tests/host_emu/make_emu_host.py): each
`kernel<<<grid, block, 0, stream>>>(args)` becomes `emu_launch(grid, block, [&] { kernel(args); })`.  The emulated device
has 2 "SMs", so persistent grids are 8 blocks; inputs are a few hundred points (one MSM takes seconds).  This library is
test infrastructure: it is built into a temporary directory, bound here with ctypes, and never loaded by the package --
the product has no CPU path (montgomery_b200/_native.py fails loudly without the CUDA library).  The GPU versions of
these checks are tests/test_gpu_*.py."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from montgomery_b200 import _native, curves, inputs
from tests.helpers import OracleCurve, points_to_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(900)
E_INVALID, E_STATE = -1, -4
# The default run keeps the CPU suite at a few minutes; MGB_TEST_FULL=1 (set by scripts/asan_emulated.sh) adds the cases
# that repeat a path with other parameters (more window sizes, the tuning mechanisms on a skewed input).
full_only = pytest.mark.skipif(not os.environ.get("MGB_TEST_FULL"), reason="repeats covered paths with other parameters; MGB_TEST_FULL=1 runs it")


class EmuHost:
    """The calls of include/montgomery_b200.h that the tests use, bound on the emulated build."""

    def __init__(self, path):
        lib = ctypes.CDLL(path)
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        po, pt = ctypes.POINTER(_native.MgbOpts), ctypes.POINTER(_native.MgbTiming)
        lib.mgb_create.argtypes = [ctypes.POINTER(vp), ci, ci, sz]
        lib.mgb_set_points.argtypes = [vp, vp, vp, sz]
        lib.mgb_random_points.argtypes = [vp, ctypes.c_uint64, sz]
        lib.mgb_get_points.argtypes = [vp, sz, sz, vp, vp]
        lib.mgb_msm.argtypes = [vp, vp, sz, po, vp, ctypes.POINTER(ci), pt]
        lib.mgb_msm_prefetch.argtypes = [vp, vp, sz]
        lib.mgb_partial_bytes.argtypes = [vp]
        lib.mgb_partial_bytes.restype = sz
        lib.mgb_msm_partial.argtypes = [vp, vp, ci, sz, po, vp, pt]
        lib.mgb_combine_partials.argtypes = [vp, vp, ci, vp, ctypes.POINTER(ci)]
        lib.mgb_msm_sharded.argtypes = [vp, vp, ci, sz, po, vp, ctypes.POINTER(ci), pt]
        lib.mgb_last_error.argtypes = [vp]
        lib.mgb_last_error.restype = ctypes.c_char_p
        lib.mgb_destroy.argtypes = [vp]
        lib.mgb_destroy.restype = None
        pi = ctypes.POINTER(ci)
        lib.mgb_comm_unique_id.argtypes = [vp]
        lib.mgb_comm_init.argtypes = [vp, vp, ci, ci]
        lib.mgb_comm_info.argtypes = [vp, pi, pi, pi]
        lib.mgb_multi_create.argtypes = [ctypes.POINTER(vp), ci, pi, ci, sz]
        lib.mgb_multi_set_points.argtypes = [vp, vp, vp, sz]
        lib.mgb_multi_random_points.argtypes = [vp, ctypes.c_uint64, sz]
        lib.mgb_multi_get_points.argtypes = [vp, sz, sz, vp, vp]
        lib.mgb_multi_msm.argtypes = [vp, vp, sz, po, vp, pi, pt]
        lib.mgb_multi_last_error.argtypes = [vp]
        lib.mgb_multi_last_error.restype = ctypes.c_char_p
        lib.mgb_multi_destroy.argtypes = [vp]
        lib.mgb_multi_destroy.restype = None
        self.lib = lib

    def create(self, label, max_points):
        cv = curves.BY_LABEL[label]
        h = ctypes.c_void_p()
        rc = self.lib.mgb_create(ctypes.byref(h), cv.curve_id, 0, max_points)
        assert rc == 0, self.lib.mgb_last_error(None)
        return Ctx(self.lib, h, cv, label)


class Ctx:
    def __init__(self, lib, h, cv, label):
        self.lib, self.h, self.cv, self.label = lib, h, cv, label

    def error(self):
        return (self.lib.mgb_last_error(self.h) or b"").decode()

    def random_points(self, n, seed):
        assert self.lib.mgb_random_points(self.h, seed, n) == 0, self.error()
        return self.get_points(n)

    def set_points(self, pts):
        xy, z = points_to_bytes(pts, self.cv.coord_bytes)
        assert self.lib.mgb_set_points(self.h, xy.ctypes.data, z.ctypes.data, len(pts)) == 0, self.error()

    def get_points(self, n):
        cb = self.cv.coord_bytes
        xy = np.zeros(n * 2 * cb, np.uint8)
        z = np.zeros(n, np.uint8)
        assert self.lib.mgb_get_points(self.h, 0, n, xy.ctypes.data, z.ctypes.data) == 0, self.error()
        rows = xy.reshape(n, 2 * cb)
        return [None if f else (int.from_bytes(r[:cb].tobytes(), "little"), int.from_bytes(r[cb:].tobytes(), "little")) for r, f in zip(rows, z)]

    def _point(self, out, flag):
        cb = self.cv.coord_bytes
        return {"x": int.from_bytes(out[:cb].tobytes(), "little"), "y": int.from_bytes(out[cb:].tobytes(), "little"), "isZero": bool(flag.value)}

    def msm(self, sc, n=None, expect_rc=0, **o):
        n = sc.shape[0] if n is None else n
        out = np.zeros(2 * self.cv.coord_bytes, np.uint8)
        flag = ctypes.c_int(0)
        tm = _native.MgbTiming()
        opts = _native.MgbOpts(o.get("c", 0), 0, 0, o.get("projective", 0), o.get("affine_reduction", 0))
        rc = self.lib.mgb_msm(self.h, sc.ctypes.data, n, ctypes.byref(opts), out.ctypes.data, ctypes.byref(flag), ctypes.byref(tm))
        assert rc == expect_rc, (rc, self.error())
        return (self._point(out, flag), tm.as_dict()) if rc == 0 else (None, None)

    def close(self):
        self.lib.mgb_destroy(self.h)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu_host")
    src, so = str(d / "msm_emu.cpp"), str(d / "libmgb_emu.so")
    emu = os.path.join(ROOT, "tests", "host_emu")
    subprocess.check_call([sys.executable, os.path.join(emu, "make_emu_host.py"), os.path.join(ROOT, "montgomery_b200", "csrc", "msm.cu"), src])
    # MGB_EMU_CXXFLAGS: e.g. "-O1 -g -fsanitize=address -fno-omit-frame-pointer" (run pytest under LD_PRELOAD=libasan.so with
    # ASAN_OPTIONS=detect_leaks=0): "device" buffers are heap blocks in the emulation, so AddressSanitizer checks every
    # kernel's global-memory accesses against the sizes msm.cu allocated -- a memcheck of the whole path without a GPU
    flags = os.environ.get("MGB_EMU_CXXFLAGS", "-O2").split()
    subprocess.check_call(["g++", "-std=c++17", *flags, "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-DMGB_HOST_EMU", "-I", emu,
                           "-I", os.path.join(ROOT, "montgomery_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           "-include", "cuda_rt_emu.h", src, "-o", so, "-ldl"])
    # the collective of the multi-GPU entry points: msm.cu binds NCCL with dlopen; the emulated build gets an in-process
    # stand-in (ranks = host threads) through the same MGB_NCCL_LIB override a host with an unusual NCCL path would use
    nccl = str(d / "libfake_nccl.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", os.path.join(emu, "fake_nccl.cpp"), "-o", nccl])
    os.environ["MGB_NCCL_LIB"] = nccl
    yield EmuHost(so)
    os.environ.pop("MGB_NCCL_LIB", None)


def oracle_msm(label, sc, pts):
    return OracleCurve(label).msm(inputs.scalars_to_ints(sc), pts)


@pytest.mark.parametrize("label", ["bls12-377", "pallas", "ed-on-bls12-377", pytest.param("bls12-381", marks=full_only)])
def test_whole_msm_default_plan(host, label):
    """random points made on the (emulated) device, read back, one MSM with the engine's own window and round choice"""
    ctx = host.create(label, 128)
    try:
        pts = ctx.random_points(96, seed=11)
        sc = inputs.random_scalars(ctx.cv.q, 96, 12)
        res, tm = ctx.msm(sc)
        assert res == oracle_msm(label, sc, pts), tm
        assert tm["n_launches"] >= 10 and tm["K"] * tm["c"] >= 126
        # fewer scalars than points held: the first n pairs; n = 0: the neutral element
        assert ctx.msm(sc, n=17)[0] == oracle_msm(label, sc[:17], pts[:17])
        assert ctx.msm(sc, n=0)[0] == OracleCurve(label).result_of(None if label != "ed-on-bls12-377" else (0, 1))
    finally:
        ctx.close()


def test_deep_trees_and_both_reductions(host, monkeypatch):
    """three tree rounds of batched-affine additions (pair lists handed from round to round), bucket finish, then the default
    (XYZZ, digit-decomposed) and the affine bucket reduction; points uploaded as bytes, with a point at infinity, a repeated
    point and a pair P, -P among them"""
    label = "bls12-377"
    ctx = host.create(label, 320)
    try:
        O = OracleCurve(label)
        pts = [O.scale(3 + 5 * i, O.G) for i in range(300)]
        pts[7] = None
        pts[20] = pts[21] = pts[22]
        pts[31] = (pts[30][0], (-pts[30][1]) % O.prm.p)
        ctx.set_points(pts)
        assert ctx.get_points(300) == pts
        sc = inputs.random_scalars(ctx.cv.q, 300, 5)
        sc[20] = sc[21] = sc[22]                              # same point, same scalar: doublings inside a bucket
        sc[31] = sc[30]
        exp = oracle_msm(label, sc, pts)
        monkeypatch.setenv("MGB_DEBUG_NROUNDS", "3")
        res, tm = ctx.msm(sc, c=6)
        assert res == exp and tm["rounds"] == 3, tm
        res, tm = ctx.msm(sc, c=6, affine_reduction=1)         # trees run to completion, group trees through k_batch_add
        assert res == exp, tm
        monkeypatch.delenv("MGB_DEBUG_NROUNDS")
        assert ctx.msm(sc, c=9)[0] == exp                      # another window size: two digits of the bucket index
        assert ctx.msm(sc, projective=1)[0] == exp             # msmProjective: no GLV, no batched-affine rounds
    finally:
        ctx.close()


@pytest.mark.parametrize("label", ["bls12-381", "ed-on-bls12-377"])
def test_chunked_ingestion(host, label, monkeypatch):
    """mgb_set_points in several chunks through the two staging halves (the path of point sets above 2^18, forced here with
    7 points per chunk): every point and every infinity flag lands at its own index; chunk borders, a last partial chunk,
    flags absent; then an MSM over the table"""
    ctx = host.create(label, 64)
    try:
        O = OracleCurve(label)
        pts = [O.scale(2 + 3 * i, O.G) for i in range(52)]
        if O.kind == "weierstrass":
            pts[6] = pts[7] = pts[51] = None                   # the last point of a chunk, the first of the next, the very last
        monkeypatch.setenv("MGB_DEBUG_INGEST_CHUNK", "7")
        ctx.set_points(pts)
        assert ctx.get_points(52) == pts
        xy, _ = points_to_bytes(pts[8:50], ctx.cv.coord_bytes)
        assert ctx.lib.mgb_set_points(ctx.h, xy.ctypes.data, None, 42) == 0, ctx.error()      # no flags: 6 whole chunks
        assert ctx.get_points(42) == pts[8:50]
        monkeypatch.delenv("MGB_DEBUG_INGEST_CHUNK")
        sc = inputs.random_scalars(ctx.cv.q, 42, 6)
        assert ctx.msm(sc)[0] == oracle_msm(label, sc, pts[8:50])
    finally:
        ctx.close()


@full_only
def test_skewed_buckets_and_tuning_paths(host, monkeypatch):
    """one scalar value for most of the input: a single bucket per window holds nearly every point, so the trees get their
    forced depth (at most 16 leftovers per bucket) and every round hands a long pair list on; then the same input through
    the mechanisms that stay in the library for tuning -- window groups on separate streams, the reduction with and
    without the bucket-finish kernel, another chunking of the group sums -- all the same canonical point"""
    label = "pallas"
    ctx = host.create(label, 160)
    try:
        pts = ctx.random_points(150, seed=31)
        sc = inputs.random_scalars(ctx.cv.q, 150, 32)
        sc[10:140] = sc[9]                                     # 131 points share every digit
        exp = oracle_msm(label, sc, pts)
        res, tm = ctx.msm(sc, c=7)
        assert res == exp and tm["max_bucket"] >= 131 and tm["rounds"] >= 4, tm
        for knob, val in (("MGB_DEBUG_GROUPS", "3"), ("MGB_DEBUG_FINISH", "0"), ("MGB_DEBUG_CH", "4")):
            monkeypatch.setenv(knob, val)
            res, tm = ctx.msm(sc, c=7)
            assert res == exp, (knob, val, tm)
            monkeypatch.delenv(knob)
    finally:
        ctx.close()


def test_error_paths_and_state(host):
    ctx = host.create("ed-on-bls12-377", 64)
    try:
        sc = inputs.random_scalars(ctx.cv.q, 64, 3)
        ctx.msm(sc, expect_rc=E_STATE)                         # no points yet
        pts = ctx.random_points(64, seed=2)
        ctx.msm(sc, c=30, expect_rc=E_INVALID)                 # window out of range
        assert "window" in ctx.error()
        big = sc.copy()
        big[5, 31] = 0xFF                                      # >= 2^251 on the path without decomposition
        ctx.msm(big, expect_rc=E_INVALID)
        assert "out of range" in ctx.error()
        assert ctx.msm(sc)[0] == oracle_msm("ed-on-bls12-377", sc, pts)    # the context is still usable
        h2 = ctypes.c_void_p()
        assert host.lib.mgb_create(ctypes.byref(h2), 17, 0, 8) == E_INVALID   # unknown curve
    finally:
        ctx.close()


def test_partials_combine_and_sharded_entry(host):
    """the multi-GPU building blocks on one emulated device: two shards -> partial accumulators -> combine = the whole MSM;
    the sharded entry point without a communicator is a one-rank job"""
    label = "pallas"
    a, b = host.create(label, 64), host.create(label, 64)
    try:
        O = OracleCurve(label)
        pts = [O.scale(7 + 3 * i, O.G) for i in range(100)]
        a.set_points(pts[:50])
        b.set_points(pts[50:])
        sc = inputs.random_scalars(a.cv.q, 100, 9)
        pb = host.lib.mgb_partial_bytes(a.h)
        parts = np.zeros(2 * pb, np.uint8)
        opts = _native.MgbOpts(0, 0, 0, 0, 0)
        for k, (cx, lo) in enumerate(((a, 0), (b, 50))):
            rc = host.lib.mgb_msm_partial(cx.h, sc[lo:lo + 50].ctypes.data, 0, 50, ctypes.byref(opts), parts.ctypes.data + k * pb, None)
            assert rc == 0, cx.error()
        out = np.zeros(2 * a.cv.coord_bytes, np.uint8)
        flag = ctypes.c_int(0)
        assert host.lib.mgb_combine_partials(a.h, parts.ctypes.data, 2, out.ctypes.data, ctypes.byref(flag)) == 0, a.error()
        assert a._point(out, flag) == oracle_msm(label, sc, pts)
        flag = ctypes.c_int(0)
        rc = host.lib.mgb_msm_sharded(a.h, sc[:50].ctypes.data, 0, 50, ctypes.byref(opts), out.ctypes.data, ctypes.byref(flag), None)
        assert rc == 0 and a._point(out, flag) == oracle_msm(label, sc[:50], pts[:50])
    finally:
        a.close()
        b.close()


def test_prefetch_bookkeeping(host):
    """mgb_msm_prefetch: registered sets are uploaded behind the next MSM's first round and consumed by the call that passes
    the same pointer and n; anything else falls back to a plain upload; a third waiting set is refused"""
    label = "pallas"
    ctx = host.create(label, 64)
    try:
        pts = ctx.random_points(64, seed=4)
        sets = [inputs.random_scalars(ctx.cv.q, 64, 40 + i) for i in range(3)]
        exp = [oracle_msm(label, s, pts) for s in sets]
        pf = lambda s, n=64: host.lib.mgb_msm_prefetch(ctx.h, s.ctypes.data, n)
        assert pf(sets[0]) == 0 and pf(sets[1]) == 0
        assert pf(sets[2]) == E_INVALID and "waiting" in ctx.error()
        assert ctx.msm(sets[0])[0] == exp[0]                   # never uploaded ahead (no MSM ran since): plain upload; starts set 1
        sets[2][:] = sets[0]                                   # a buffer that is NOT registered may change freely
        assert ctx.msm(sets[1])[0] == exp[1]                   # consumed from the prefetched copy
        assert pf(sets[1]) == 0
        assert ctx.msm(sets[2])[0] == exp[0]                   # other pointer: own upload; set 1 is uploaded meanwhile
        assert ctx.msm(sets[1], n=63)[0] == oracle_msm(label, sets[1][:63], pts[:63])   # other n: not the registered set
        assert ctx.msm(sets[1])[0] == exp[1]
        assert pf(sets[0], 65) == E_INVALID                    # more than the context holds
    finally:
        ctx.close()


@pytest.mark.parametrize("label,c", [("pallas", 2), pytest.param("pallas", 4, marks=full_only), pytest.param("pallas", 8, marks=full_only), ("pallas", 11),
                                     pytest.param("ed-on-bls12-377", 3, marks=full_only), ("ed-on-bls12-377", 9)])
def test_window_sizes(host, label, c):
    """one, two and three reduction digits, a clipped top window (c = 11: 128 = 11 * 11 + 7 bits), the smallest windows"""
    ctx = host.create(label, 64)
    try:
        pts = ctx.random_points(48, seed=100 + c)
        sc = inputs.random_scalars(ctx.cv.q, 48, 200 + c)
        sc[3] = 0                                              # a zero scalar contributes to no bucket
        sc[4] = np.frombuffer(int(ctx.cv.q - 1).to_bytes(32, "little"), dtype=np.uint8)
        res, tm = ctx.msm(sc, c=c)
        assert res == oracle_msm(label, sc, pts) and tm["c"] == c, tm
    finally:
        ctx.close()


def _read_point(cv, out, flag):
    cb = cv.coord_bytes
    return {"x": int.from_bytes(out[:cb].tobytes(), "little"), "y": int.from_bytes(out[cb:].tobytes(), "little"), "isZero": bool(flag.value)}


def test_one_process_several_devices(host):
    """mgb_multi_*: three (emulated) devices behind one handle, one host thread per device, the all-gather of the partial
    accumulators through the stand-in NCCL: contiguous shards, reads across shard borders, n below the point count (a
    partly used and an empty shard), n = 0, device-made points"""
    label = "pallas"
    cv = curves.BY_LABEL[label]
    lib = host.lib
    devs = (ctypes.c_int * 3)(0, 1, 2)
    m = ctypes.c_void_p()
    assert lib.mgb_multi_create(ctypes.byref(m), cv.curve_id, devs, 3, 40) == 0, lib.mgb_multi_last_error(None)
    try:
        O = OracleCurve(label)
        pts = [O.scale(11 + 7 * i, O.G) for i in range(100)]
        pts[35] = None                                         # a point at infinity in the second shard
        xy, z = points_to_bytes(pts, cv.coord_bytes)
        assert lib.mgb_multi_set_points(m, xy.ctypes.data, z.ctypes.data, 100) == 0, lib.mgb_multi_last_error(m)

        def get(first, n):
            cb = cv.coord_bytes
            o, f = np.zeros(n * 2 * cb, np.uint8), np.zeros(n, np.uint8)
            assert lib.mgb_multi_get_points(m, first, n, o.ctypes.data, f.ctypes.data) == 0, lib.mgb_multi_last_error(m)
            rows = o.reshape(n, 2 * cb)
            return [None if fl else (int.from_bytes(r[:cb].tobytes(), "little"), int.from_bytes(r[cb:].tobytes(), "little")) for r, fl in zip(rows, f)]

        assert get(0, 100) == pts and get(30, 45) == pts[30:75]   # shards of 34 / 34 / 32 points

        def msm(sc, n, expect_rc=0):
            out, flag, tm = np.zeros(2 * cv.coord_bytes, np.uint8), ctypes.c_int(0), _native.MgbTiming()
            rc = lib.mgb_multi_msm(m, sc.ctypes.data, n, None, out.ctypes.data, ctypes.byref(flag), ctypes.byref(tm))
            assert rc == expect_rc, (rc, lib.mgb_multi_last_error(m))
            return _read_point(cv, out, flag) if rc == 0 else None

        sc = inputs.random_scalars(cv.q, 100, 21)
        assert msm(sc, 100) == oracle_msm(label, sc, pts)
        assert msm(sc, 40) == oracle_msm(label, sc[:40], pts[:40])     # device 1 uses 6 of its points, device 2 none
        assert msm(sc, 0) == O.result_of(None)
        msm(sc, 101, expect_rc=E_INVALID)
        big_xy, big_z = np.zeros(121 * 2 * cv.coord_bytes, np.uint8), np.ones(121, np.uint8)
        assert lib.mgb_multi_set_points(m, big_xy.ctypes.data, big_z.ctypes.data, 121) == E_INVALID   # 41 per device > 40
        assert b"device" in lib.mgb_multi_last_error(m)
        # points made on the devices: shard g is the known-dlog set of seed + g
        assert lib.mgb_multi_random_points(m, 77, 60) == 0, lib.mgb_multi_last_error(m)
        rp = get(0, 60)
        assert all(O.A.is_on_curve(p) for p in rp) and len(set(rp)) == 60
        assert msm(sc, 60) == oracle_msm(label, sc[:60], rp)
    finally:
        lib.mgb_multi_destroy(m)


def test_one_context_per_rank_with_communicator(host, monkeypatch):
    """mgb_comm_unique_id / mgb_comm_init / mgb_msm_sharded as an SPMD host uses them (here: one thread per rank): every
    rank ends with the same canonical sum; an empty shard; state errors; an asynchronous NCCL failure is reported"""
    import threading
    label = "pallas"
    lib = host.lib
    ranks = [host.create(label, 64) for _ in range(2)]
    try:
        O = OracleCurve(label)
        pts = [O.scale(5 + 9 * i, O.G) for i in range(80)]
        ranks[0].set_points(pts[:48])
        ranks[1].set_points(pts[48:])
        sc = inputs.random_scalars(ranks[0].cv.q, 80, 33)
        opts = _native.MgbOpts(0, 0, 0, 0, 0)
        out = np.zeros(2 * ranks[0].cv.coord_bytes, np.uint8)
        flag = ctypes.c_int(0)
        # before mgb_comm_init the context is a one-rank job
        rk, wd, ver = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(-1)
        assert lib.mgb_comm_info(ranks[0].h, ctypes.byref(rk), ctypes.byref(wd), ctypes.byref(ver)) == 0
        assert (rk.value, wd.value, ver.value) == (0, 1, 99999)      # the version the stand-in reports: it is the library bound
        uid = np.zeros(_native.COMM_ID_BYTES, np.uint8)
        assert lib.mgb_comm_unique_id(uid.ctypes.data) == 0
        assert lib.mgb_comm_init(ranks[0].h, uid.ctypes.data, 2, 2) == E_INVALID   # rank out of range

        def on_ranks(fn):
            res = [None, None]
            th = [threading.Thread(target=lambda r=r: res.__setitem__(r, fn(r))) for r in range(2)]
            [t.start() for t in th]
            [t.join(600) for t in th]
            assert not any(t.is_alive() for t in th), "a rank hangs in the collective"
            return res

        assert on_ranks(lambda r: lib.mgb_comm_init(ranks[r].h, uid.ctypes.data, r, 2)) == [0, 0]
        assert lib.mgb_comm_init(ranks[0].h, uid.ctypes.data, 0, 2) == E_STATE     # already has a communicator
        assert lib.mgb_comm_info(ranks[1].h, ctypes.byref(rk), ctypes.byref(wd), None) == 0 and (rk.value, wd.value) == (1, 2)

        def sharded(r, lo, hi):
            o, f = np.zeros_like(out), ctypes.c_int(0)
            rc = lib.mgb_msm_sharded(ranks[r].h, sc[lo:hi].ctypes.data if hi > lo else None, 0, hi - lo, ctypes.byref(opts), o.ctypes.data, ctypes.byref(f), None)
            return rc, (ranks[r]._point(o, f) if rc == 0 else ranks[r].error())

        exp = oracle_msm(label, sc, pts)
        assert on_ranks(lambda r: sharded(r, *((0, 48), (48, 80))[r])) == [(0, exp), (0, exp)]
        exp0 = oracle_msm(label, sc[:48], pts[:48])
        assert on_ranks(lambda r: sharded(r, *((0, 48), (48, 48))[r])) == [(0, exp0), (0, exp0)]    # rank 1's shard is empty
        monkeypatch.setenv("MGB_FAKE_NCCL_ASYNC_ERROR", "1")
        res = on_ranks(lambda r: sharded(r, *((0, 48), (48, 80))[r]))
        assert [rc for rc, _ in res] == [_native.E_COMM] * 2 and "asynchronous" in res[0][1]
    finally:
        for cx in ranks:
            cx.close()


def test_failed_shard_does_not_hang_the_collective(host):
    """a rank that cannot compute its shard (here: a scalar out of range on the path without decomposition) still joins the
    all-gather, flagged: it returns its own error, its peers return MGB_E_COMM instead of waiting forever or of returning a
    sum that misses a shard; the contexts stay usable.  Same through mgb_multi_msm, which reports the root cause."""
    import threading
    label = "ed-on-bls12-377"
    cv = curves.BY_LABEL[label]
    lib = host.lib
    ranks = [host.create(label, 32) for _ in range(2)]
    m = ctypes.c_void_p()
    try:
        pts = ranks[0].random_points(24, seed=8) + ranks[1].random_points(24, seed=9)
        sc = inputs.random_scalars(cv.q, 48, 44)
        bad = sc.copy()
        bad[30, 31] = 0xFF                                     # in rank 1's shard
        uid = np.zeros(_native.COMM_ID_BYTES, np.uint8)
        assert lib.mgb_comm_unique_id(uid.ctypes.data) == 0
        opts = _native.MgbOpts(0, 0, 0, 0, 0)

        def on_ranks(fn):
            res = [None, None]
            th = [threading.Thread(target=lambda r=r: res.__setitem__(r, fn(r)), daemon=True) for r in range(2)]
            [t.start() for t in th]
            [t.join(300) for t in th]
            assert not any(t.is_alive() for t in th), "a rank hangs in the collective"
            return res

        def sharded(r, s, n=24):
            o, f = np.zeros(2 * cv.coord_bytes, np.uint8), ctypes.c_int(0)
            rc = lib.mgb_msm_sharded(ranks[r].h, s[24 * r:24 * r + 24].ctypes.data, 0, n, ctypes.byref(opts), o.ctypes.data, ctypes.byref(f), None)
            return rc, (ranks[r]._point(o, f) if rc == 0 else ranks[r].error())

        assert on_ranks(lambda r: lib.mgb_comm_init(ranks[r].h, uid.ctypes.data, r, 2)) == [0, 0]
        res = on_ranks(lambda r: sharded(r, bad))
        assert res[0][0] == _native.E_COMM and "another rank" in res[0][1]
        assert res[1][0] == E_INVALID and "out of range" in res[1][1]
        # an argument error only one rank makes (more pairs than it holds points): rejected before any work, and the rank
        # still joins, flagged
        res = on_ranks(lambda r: sharded(r, sc, n=(24, 25)[r]))
        assert res[0][0] == _native.E_COMM and res[1][0] == E_INVALID and "exceeds" in res[1][1]
        exp = oracle_msm(label, sc, pts)
        assert on_ranks(lambda r: sharded(r, sc)) == [(0, exp), (0, exp)]

        devs = (ctypes.c_int * 2)(0, 1)
        assert lib.mgb_multi_create(ctypes.byref(m), cv.curve_id, devs, 2, 32) == 0, lib.mgb_multi_last_error(None)
        xy, z = points_to_bytes(pts, cv.coord_bytes)
        assert lib.mgb_multi_set_points(m, xy.ctypes.data, z.ctypes.data, 48) == 0
        out, flag = np.zeros(2 * cv.coord_bytes, np.uint8), ctypes.c_int(0)
        assert lib.mgb_multi_msm(m, bad.ctypes.data, 48, None, out.ctypes.data, ctypes.byref(flag), None) == E_INVALID
        err = lib.mgb_multi_last_error(m).decode()
        assert "device 1" in err and "out of range" in err
        assert lib.mgb_multi_msm(m, sc.ctypes.data, 48, None, out.ctypes.data, ctypes.byref(flag), None) == 0
        assert _read_point(cv, out, flag) == exp
    finally:
        for cx in ranks:
            cx.close()
        lib.mgb_multi_destroy(m)

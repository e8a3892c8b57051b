"""ctypes wrapper of oracle/msm_cpu.cpp (CPU restatement of the reference's MSM; oracle = test and
baseline infrastructure).  Builds oracle/libmsm_cpu.so on demand with the Makefile beside it."""
import ctypes
import os
import subprocess

import numpy as np

from .glv import GlvScalar
from .params import BLS12_377, BLS12_381, ED_ON_BLS12_377, PALLAS

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmsm_cpu.so")


class _Glv(ctypes.Structure):
    _fields_ = [("m0", ctypes.c_uint64 * 3), ("m1", ctypes.c_uint64 * 3), ("v", (ctypes.c_uint64 * 2) * 4),
                ("sm0", ctypes.c_int32), ("sm1", ctypes.c_int32), ("sv", ctypes.c_int32 * 4),
                ("m_bits", ctypes.c_int32), ("k_bits", ctypes.c_int32), ("max_bits", ctypes.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "msm_cpu.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmsm_cpu.so"])
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.ref_msm.restype = ctypes.c_int
    return _lib


def _limbs64(x, n):
    return (ctypes.c_uint64 * n)(*[(x >> (64 * i)) & (2**64 - 1) for i in range(n)])


def _glv_struct(prm):
    g = GlvScalar(prm.q, prm.lam)
    s = _Glv()
    s.m0 = _limbs64(abs(g.m0), 3)
    s.m1 = _limbs64(abs(g.m1), 3)
    for i, v in enumerate(g.V):
        s.v[i] = _limbs64(abs(v), 2)
        s.sv[i] = 1 if v >= 0 else -1
    s.sm0 = 1 if g.m0 >= 0 else -1
    s.sm1 = 1 if g.m1 >= 0 else -1
    s.m_bits, s.k_bits, s.max_bits = g.m, g.k, g.max_bits
    return s


_CURVES = {"bls12-377": (0, BLS12_377, 6, 48), "pallas": (1, PALLAS, 4, 32), "ed-on-bls12-377": (2, ED_ON_BLS12_377, 4, 32),
           "bls12-381": (3, BLS12_381, 6, 48)}


def msm(label, scalars_bytes: np.ndarray, points_bytes: np.ndarray, n: int, threads: int = 0, c: int = 0):
    """Returns ({x, y, isZero}, milliseconds of the MSM proper)."""
    lib = _load()
    cid, prm, nl, cb = _CURVES[label]
    threads = threads or os.cpu_count() or 1
    mod = _limbs64(prm.p, nl)
    if cid == 2:
        aux = _limbs64(prm.d, nl)
        glv = None
    else:
        aux = _limbs64(prm.beta, nl)
        glv = ctypes.byref(_glv_struct(prm))
    sc = np.ascontiguousarray(scalars_bytes.reshape(-1))
    pt = np.ascontiguousarray(points_bytes.reshape(-1))
    assert sc.size >= 32 * n and pt.size >= 2 * cb * n
    out = np.zeros(2 * cb, dtype=np.uint8)
    is_zero = ctypes.c_int(0)
    ms = ctypes.c_double(0)
    vp = ctypes.c_void_p
    rc = lib.ref_msm(cid, mod, aux, glv, sc.ctypes.data_as(vp), pt.ctypes.data_as(vp), ctypes.c_size_t(n), threads, c,
                     out.ctypes.data_as(vp), ctypes.byref(is_zero), ctypes.byref(ms))
    assert rc == 0
    res = {"x": int.from_bytes(out[:cb].tobytes(), "little"), "y": int.from_bytes(out[cb:].tobytes(), "little"),
           "isZero": bool(is_zero.value)}
    return res, ms.value


def known_dlog_points(label, seed: int, n: int, threads: int = 0) -> np.ndarray:
    """(n, 2*coord_bytes) uint8 array of the points a_i*G, a_i = splitmix64(seed, i) -- the same set
    `mgb_random_points(seed, n)` generates on the GPU."""
    lib = _load()
    cid, prm, nl, cb = _CURVES[label]
    threads = threads or os.cpu_count() or 1
    out = np.zeros((n, 2 * cb), dtype=np.uint8)
    aux = _limbs64(prm.d, nl) if cid == 2 else None
    rc = lib.ref_known_dlog_points(cid, _limbs64(prm.p, nl), aux, _limbs64(prm.G[0], nl), _limbs64(prm.G[1], nl),
                                   ctypes.c_uint64(seed), ctypes.c_size_t(n), threads, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out

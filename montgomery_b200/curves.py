"""Curve descriptors used by the host layer (same parameter sets as the reference's
src/concrete/bls12-377.params.ts, pasta.params.ts, ed-on-bls12-377.params.ts, bls12-381.params.ts)."""
from dataclasses import dataclass

from . import _native


@dataclass(frozen=True)
class CurveInfo:
    label: str
    curve_id: int
    kind: str           # "weierstrass" | "twisted-edwards"
    p: int              # base field modulus
    q: int              # scalar field modulus (group order)
    coord_bytes: int    # bytes per coordinate in the reference byte format

    @property
    def point_bytes(self):
        return 2 * self.coord_bytes


BLS12_377 = CurveInfo(
    "bls12-377", _native.BLS12_377_G1, "weierstrass",
    0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001, 48)
PALLAS = CurveInfo(
    "pallas", _native.PALLAS, "weierstrass",
    0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
    0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001, 32)
ED_ON_BLS12_377 = CurveInfo(
    "ed-on-bls12-377", _native.ED_ON_BLS12_377, "twisted-edwards",
    0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    0x4AAD957A68B2955982D1347970DEC005293A3AFC43C8AFEB95AEE9AC33FD9FF, 32)

BLS12_381 = CurveInfo(
    "bls12-381", _native.BLS12_381_G1, "weierstrass",
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001, 48)

BY_LABEL = {c.label: c for c in (BLS12_377, PALLAS, ED_ON_BLS12_377, BLS12_381)}

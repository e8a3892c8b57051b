"""Curve parameters, restated from the reference (oracle = test infrastructure).

  BLS12-377        /root/reference/src/concrete/bls12-377.params.ts:11-46
  ed-on-BLS12-377  /root/reference/src/concrete/ed-on-bls12-377.params.ts:5-31
  Pallas           /root/reference/src/concrete/pasta.params.ts:10-46
"""
from dataclasses import dataclass
from typing import Optional, Tuple


@dataclass(frozen=True)
class WeierstrassParams:
    label: str
    p: int          # base field modulus
    q: int          # group order (scalar field)
    h: int          # cofactor
    a: int
    b: int
    G: Tuple[int, int]
    lam: int        # endomorphism scalar (cube root of 1 in Fq)
    beta: int       # endomorphism base (cube root of 1 in Fp)
    kind: str = "weierstrass"


@dataclass(frozen=True)
class TwistedEdwardsParams:
    label: str
    p: int
    q: int
    h: int
    d: int
    G: Tuple[int, int]
    kind: str = "twisted-edwards"


# bls12-377.params.ts:11-34
BLS12_377 = WeierstrassParams(
    label="bls12-377",
    p=0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    q=0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    h=0x170B5D44300000000000000000000000,
    a=0,
    b=1,
    G=(
        0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF,
        0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6,
    ),
    lam=0x12AB655E9A2CA55660B44D1E5C37B00114885F32400000000000000000000000,
    beta=0x1AE3A4617C510EABC8756BA8F8C524EB8882A75CC9BC8E359064EE822FB5BFFD1E945779FFFFFFFFFFFFFFFFFFFFFFF,
)

# ed-on-bls12-377.params.ts:5-31  (-x^2 + y^2 = 1 + d x^2 y^2)
ED_ON_BLS12_377 = TwistedEdwardsParams(
    label="ed-on-bls12-377",
    p=0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    q=0x4AAD957A68B2955982D1347970DEC005293A3AFC43C8AFEB95AEE9AC33FD9FF,
    h=4,
    d=3021,
    G=(
        0x9F1B5A5BAF6ACF06FED91C9AE9EBFA06068DD2835790980894E2328F3EBCA05,
        0x9A20DF36571AC3CD906B256080BA8454453C177AAF3131BB50A67BF1A806781,
    ),
)


def _pallas() -> WeierstrassParams:
    # pasta.params.ts:10-46 -- lambda and beta are *computed* there, same here
    p = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
    q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
    lam = pow(5, (q - 1) // 3, q)
    assert pow(lam, 3, q) == 1 and lam != 1
    beta2 = pow(5, (p - 1) // 3, p)
    beta = beta2 * beta2 % p
    assert beta2 * beta % p == 1
    return WeierstrassParams(
        label="pallas", p=p, q=q, h=1, a=0, b=5,
        G=(1, 0x1B74B5A30A12937C53DFA9F06378EE548F655BD4333D477119CF7A23CAED2ABB),
        lam=lam, beta=beta,
    )


PALLAS = _pallas()

CURVES = {"bls12-377": BLS12_377, "pallas": PALLAS, "ed-on-bls12-377": ED_ON_BLS12_377}

# known-answer fixtures
# scripts/zprize23/submission-test-bls377.ts:6-10
KAT_BLS12_377_POINT = (
    111871295567327857271108656266735188604298176728428155068227918632083036401841336689521497731900230387779623820740,
    76860045326390600098227152997486448974650822224305058012700629806287380625419427989664237630603922765089083164740,
)
# scripts/zprize23/submission-test.ts:5-10  (x, y, t with z = 1)
KAT_ED377_POINT = (
    2796670805570508460920584878396618987767121022598342527208237783066948667246,
    8134280397689638111748378379571739274369602049665521098046934931245960532166,
    3446088593515175914550487355059397868296219355049460558182099906777968652023,
)

"""CPU oracle for the MSM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, with Python arbitrary-precision integers, the plain-bigint layer of the
reference (`/root/reference/src/bigint/*.ts`, `src/glv/glv.ts`, `src/wasm/glv.ts`,
`src/concrete/*.params.ts`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package.
The product (`montgomery_b200/`) never does.

Parity status: the reference cannot be executed in this image (no Node / Wasm
runtime), and its MSM tests hold no golden output vectors.
The oracle is pinned against everything the reference does hold for this path:
  * the two known-answer fixtures of `scripts/zprize23/submission-test*.ts`,
  * the generators / on-curve / subgroup facts of `src/bigint/curves.test.ts`,
  * the algebraic MSM identities of `src/bigint/msm.test.ts`,
  * the GLV identity `s0 + s1*lambda = s (mod q)` of `src/scalar-glv.ts:92-103`,
and cross-checked by an independent C restatement (`oracle/msm_cpu.cpp`).
For random inputs the reference holds no outputs at all (its tests compare two of its own
implementations at run time), so there the repository pins itself: oracle-generated golden
vectors (`tests/golden/`, Pippenger == defining sum) and the closed form
[(sum s_i a_i) mod q] G for known-dlog points at every size.
"""

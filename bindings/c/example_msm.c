/* Plain-C host of the engine: the smallest program a maintainer would write against
 * include/montgomery_b200.h.  It plays scripts/run-msm-377.ts (random points, random scalars, one MSM,
 * print the affine result and the phase log):
 *
 *     gcc -std=c99 -I include bindings/c/example_msm.c -L montgomery_b200 -lmontgomery_b200 \
 *         -Wl,-rpath,$PWD/montgomery_b200 -o example_msm && ./example_msm 16
 *
 * Exit code 0 = MSM done, 3 = no CUDA device (mgb_create reports MGB_E_CUDA; there is no CPU path).
 * tests/test_abi.py builds this file to check that the header is valid C99 and that every call links. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "montgomery_b200.h"

/* uniform scalar below 2^252 (< q of BLS12-377): 32 little-endian bytes from a 64-bit LCG */
static void fill_scalars(uint8_t* s, size_t n, uint64_t seed) {
  size_t i;
  for (i = 0; i < 32 * n; i++) {
    seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    s[i] = (uint8_t)(seed >> 56);
    if ((i & 31) == 31) s[i] &= 0x0f;
  }
}

int main(int argc, char** argv) {
  int logn = argc > 1 ? atoi(argv[1]) : 16;
  size_t n = (size_t)1 << logn, i;
  mgb_ctx* ctx = NULL;
  mgb_opts opts;
  mgb_timing tm;
  uint8_t out[96];
  uint8_t* scalars;
  int is_zero = 0, rc;

  rc = mgb_create(&ctx, MGB_BLS12_377_G1, 0, n);
  if (rc != MGB_OK) {
    fprintf(stderr, "mgb_create: %d (%s)\n", rc, mgb_last_error(NULL));
    return rc == MGB_E_CUDA ? 3 : 1;
  }
  rc = mgb_random_points(ctx, 0x6d6f6e74u, n);                 /* randomPointsFast(n) */
  if (rc != MGB_OK) { fprintf(stderr, "mgb_random_points: %s\n", mgb_last_error(ctx)); return 1; }
  scalars = (uint8_t*)malloc(32 * n);
  fill_scalars(scalars, n, 1);
  memset(&opts, 0, sizeof opts);                               /* c = 0: engine default window */
  rc = mgb_msm(ctx, scalars, n, &opts, out, &is_zero, &tm);    /* Parallel.msmUnsafe + toAffine + toBigint */
  if (rc != MGB_OK) { fprintf(stderr, "mgb_msm: %s\n", mgb_last_error(ctx)); return 1; }
  printf("msm of 2^%d points: %.3f ms on the device (c = %d, %d windows, %u kernels)\n  x = 0x", logn, tm.total, tm.c, tm.K, tm.n_launches);
  for (i = 48; i-- > 0;) printf("%02x", out[i]);
  printf("\n  y = 0x");
  for (i = 96; i-- > 48;) printf("%02x", out[i]);
  printf("\n  isZero = %d\n", is_zero);
  free(scalars);
  mgb_destroy(ctx);
  return 0;
}

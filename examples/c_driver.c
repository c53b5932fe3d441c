/*
 * examples/c_driver.c -- the reference's fenton.py driver loop (fenton.py:155-187) written against
 * the C ABI only (include/fib_b200.h): no Python, no torch.  Shows that the drop-in boundary is a
 * plain C shared library.
 *
 *   gcc -O2 -Iinclude examples/c_driver.c -Lfib_tf_b200 -lfibb200 -Wl,-rpath,$PWD/fib_tf_b200 -o c_driver
 *   ./c_driver [N=256] [iterations=60]      prints a checksum of U and the probe at [20, N/2]
 *
 * Exit code 0 on success, 2 when the library reports an error (e.g. no CUDA device: there is no
 * CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fib_b200.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) < 0) {                                                      \
      fprintf(stderr, "%s failed: %s\n", #call, fib_last_error());         \
      return 2;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 256;
  const int iters = argc > 2 ? atoi(argv[2]) : 60;
  fib_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = sizeof cfg;
  cfg.model = FIB_FENTON4V;
  cfg.height = cfg.width = n;
  cfg.dt = 0.1;
  cfg.diff = 1.5;
  fib_ctx *ctx = NULL;
  CHECK(fib_create(&cfg, &ctx));

  /* initial state of fenton.py:116-123: U = 0 (column 1 = 1: the S1 stimulus), V = W = 1, S = 0 */
  const size_t cells = (size_t)n * n;
  float *plane = (float *)malloc(cells * sizeof(float));
  for (int v = 0; v < fib_num_vars(ctx); ++v) {
    const char *name = fib_var_name(ctx, v);
    const float fill = (!strcmp(name, "V") || !strcmp(name, "W")) ? 1.0f : 0.0f;
    for (size_t i = 0; i < cells; ++i) plane[i] = fill;
    if (!strcmp(name, "U"))
      for (int r = 0; r < n; ++r) plane[(size_t)r * n + 1] = 1.0f;
    CHECK(fib_set_state(ctx, v, plane, cells));
  }
  const int U = fib_var_index(ctx, "U");
  for (int i = 0; i < iters; ++i) {
    CHECK(fib_step(ctx, FIB_OP_ODE, 1));                 /* one run() iteration = 10 time steps */
    if (i == iters / 2)                                  /* S2: 'luq' = rows 1:H/2, cols 1:W/2 */
      CHECK(fib_stimulate(ctx, U, 1, n / 2, 1, n / 2, 1.0f, 0.0f));
  }
  float probe = 0.f;
  CHECK(fib_probe(ctx, U, 20 < n ? 20 : n - 1, n / 2, &probe));
  CHECK(fib_get_state(ctx, U, plane, cells));
  double sum = 0.0;
  for (size_t i = 0; i < cells; ++i) sum += plane[i];
  unsigned long long launches = 0;
  CHECK(fib_launch_count(ctx, (uint64_t *)&launches));
  printf("n=%d iterations=%d time_steps=%d kernels=%llu sum(U)=%.6f probe=%.6f\n", n, iters,
         iters * fib_dt_per_step(ctx), launches, sum, probe);
  free(plane);
  CHECK(fib_destroy(ctx));
  return 0;
}

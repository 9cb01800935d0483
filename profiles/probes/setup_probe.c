/* setup_probe.c -- cs_b200_create on an nr x nc 8-neighbour raster Laplacian built in C (no Python):
 * CS_B200_VERBOSE=2 ./setup_probe 3163 3163   prints the setup phase timers of the library.
 * gcc -O2 -I include -o setup_probe setup_probe.c -L circuitscape_b200/lib -lcsb200 -lm -Wl,-rpath,... */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "cs_b200.h"
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }
int main(int argc, char** argv) {
  const int nr = argc > 1 ? atoi(argv[1]) : 1000, nc = argc > 2 ? atoi(argv[2]) : 1000, reps = argc > 3 ? atoi(argv[3]) : 2;
  const int64_t n = (int64_t)nr * nc;
  int32_t* rp = malloc((n + 1) * sizeof(int32_t));
  int32_t* ci = malloc(9 * n * sizeof(int32_t));
  double* va = malloc(9 * n * sizeof(double));
  double* g = malloc(n * sizeof(double));
  uint64_t sd = 88172645463325252ULL;
  for (int64_t i = 0; i < n; ++i) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; g[i] = 1.0 / (1.0 + 9.0 * (double)(sd >> 11) / 9007199254740992.0); }
  int64_t nnz = 0;
  for (int c = 0; c < nc; ++c)
    for (int r = 0; r < nr; ++r) {
      const int64_t i = (int64_t)c * nr + r;
      rp[i] = (int32_t)nnz;
      double deg = 0; int64_t dpos = -1;
      for (int dc = -1; dc <= 1; ++dc)
        for (int dr = -1; dr <= 1; ++dr) {
          const int rr = r + dr, cc = c + dc;
          if (rr < 0 || rr >= nr || cc < 0 || cc >= nc) continue;
          const int64_t j = (int64_t)cc * nr + rr;
          if (j == i) { dpos = nnz; ci[nnz++] = (int32_t)i; continue; }
          const double w = (dr && dc) ? (g[i] + g[j]) / (2.0 * sqrt(2.0)) : (g[i] + g[j]) / 2.0;
          ci[nnz] = (int32_t)j; va[nnz++] = -w; deg += w;
        }
      va[dpos] = deg;
    }
  rp[n] = (int32_t)nnz;
  printf("n=%lld nnz=%lld\n", (long long)n, (long long)nnz);
  for (int rep = 0; rep < reps; ++rep) {
    cs_b200_opts o = {0};
    o.precond = CS_B200_PRECOND_AMG;
    cs_b200_handle* h = NULL;
    const double t0 = now();
    int rc = cs_b200_create(n, nnz, rp, ci, va, 32, 0, CS_B200_F64, 0, &o, &h);
    const double t1 = now();
    if (rc) { fprintf(stderr, "create failed %d: %s\n", rc, cs_b200_last_error(NULL)); return 1; }
    cs_b200_stats st; cs_b200_get_stats(h, &st);
    int64_t src[8] = {1, 2, 3, 4, 5, 6, 7, 8}, dst[8]; for (int k = 0; k < 8; ++k) dst[k] = n - 1 - 17 * k;
    double R[8]; int64_t it[8]; double rr[8];
    const double t2 = now();
    rc = cs_b200_solve_pairs(h, 8, src, dst, NULL, 1e-6, 100000, R, NULL, NULL, 1, it, rr);
    const double t3 = now();
    printf("rep %d: create %.1f ms wall (device %.1f ms), 8 pairs %.1f ms, iters %lld R0 %.6f rc %d\n", rep, t1 - t0, st.setup_ms, t3 - t2, (long long)it[0], R[0], rc);
    const double t4 = now();
    cs_b200_destroy(h);
    printf("        destroy %.1f ms\n", now() - t4);
  }
  return 0;
}

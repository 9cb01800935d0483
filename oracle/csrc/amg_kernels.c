/* amg_kernels.c -- sequential kernels of the CPU oracle's "CG+AMG" solver.
 * TEST/BASELINE INFRASTRUCTURE, not product code (see oracle/amg.py).
 * Restates the published algorithms the reference reaches through
 * AlgebraicMultigrid.jl 1.2 (src/core.jl:164-167): standard (Vanek) aggregation
 * and Gauss-Seidel sweeps, plus a CSR mat-vec and the PCG loop glue is in Python. */
#include <stdint.h>
#include <stdlib.h>

/* y = A x */
void csr_matvec(int64_t n, const int32_t* ip, const int32_t* ix, const double* a,
                const double* x, double* y) {
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int32_t j = ip[i]; j < ip[i + 1]; ++j) s += a[j] * x[ix[j]];
    y[i] = s;
  }
}

/* one Gauss-Seidel sweep on A x = b; dir = +1 forward, -1 backward */
void gauss_seidel(int64_t n, const int32_t* ip, const int32_t* ix, const double* a,
                  double* x, const double* b, int dir) {
  int64_t i = dir > 0 ? 0 : n - 1;
  for (int64_t c = 0; c < n; ++c, i += dir) {
    double s = b[i], d = 0.0;
    for (int32_t j = ip[i]; j < ip[i + 1]; ++j) {
      if (ix[j] == i) d += a[j]; else s -= a[j] * x[ix[j]];
    }
    if (d != 0.0) x[i] = s / d;
  }
}

/* Standard aggregation on the strength graph S (CSR pattern, no diagonal needed).
 * agg[i] = aggregate id or -1 (isolated).  Returns number of aggregates. */
int64_t standard_aggregation(int64_t n, const int32_t* ip, const int32_t* ix, int32_t* agg) {
  int64_t nagg = 0;
  for (int64_t i = 0; i < n; ++i) agg[i] = -1;
  /* pass 1: roots whose whole strong neighbourhood is free */
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    int has_nbr = 0, free_nbrs = 1;
    for (int32_t j = ip[i]; j < ip[i + 1]; ++j) {
      if (ix[j] == i) continue;
      has_nbr = 1;
      if (agg[ix[j]] != -1) { free_nbrs = 0; break; }
    }
    if (!has_nbr) { agg[i] = -2; continue; } /* isolated */
    if (!free_nbrs) continue;
    agg[i] = (int32_t)nagg;
    for (int32_t j = ip[i]; j < ip[i + 1]; ++j) agg[ix[j]] = (int32_t)nagg;
    ++nagg;
  }
  /* pass 2: attach leftovers to a neighbouring pass-1 aggregate */
  int32_t* tmp = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  for (int64_t i = 0; i < n; ++i) tmp[i] = agg[i];
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    for (int32_t j = ip[i]; j < ip[i + 1]; ++j) {
      if (tmp[ix[j]] >= 0) { agg[i] = tmp[ix[j]]; break; }
    }
  }
  free(tmp);
  /* pass 3: remaining nodes form aggregates with their free neighbours */
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    agg[i] = (int32_t)nagg;
    for (int32_t j = ip[i]; j < ip[i + 1]; ++j)
      if (agg[ix[j]] == -1) agg[ix[j]] = (int32_t)nagg;
    ++nagg;
  }
  for (int64_t i = 0; i < n; ++i) if (agg[i] == -2) agg[i] = -1;
  return nagg;
}

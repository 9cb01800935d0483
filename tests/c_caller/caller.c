/* A non-Python caller of libcsb200.so standing in for the Julia `ccall` glue of INTEGRATION.md:
 * construct_cholesky_factor -> cs_b200_create with 1-based Int64 colptr / rowval (what a
 * SparseMatrixCSC{Float64,Int64} holds, ext/CircuitscapePardisoExt.jl:31-45 is the shape matched),
 * solve_linear_system -> cs_b200_solve_rhs on a column-major n x k rhs, then destroy.
 * Builds the Laplacian of an nr x nc 4-neighbour grid with unit conductances, injects -1 / +1 at two
 * corners in column 0 and a second pair in column 1, prints the effective resistances and checks
 * them against the series/parallel value known for the 1 x m path (when nr == 1).
 * usage: caller nr nc      exit code 0 = all checks passed */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "cs_b200.h"

int main(int argc, char** argv) {
  const int nr = argc > 1 ? atoi(argv[1]) : 1, nc = argc > 2 ? atoi(argv[2]) : 50;
  const int64_t n = (int64_t)nr * nc;
  int64_t* colptr = (int64_t*)malloc((n + 1) * sizeof(int64_t));
  int64_t* rowval = (int64_t*)malloc(5 * n * sizeof(int64_t));
  double* nzval = (double*)malloc(5 * n * sizeof(double));
  int64_t nnz = 0;
  for (int c = 0; c < nc; ++c)
    for (int r = 0; r < nr; ++r) {                     /* column-major node numbering, 1-based */
      const int64_t i = (int64_t)c * nr + r;
      colptr[i] = nnz + 1;
      int deg = (r > 0) + (r + 1 < nr) + (c > 0) + (c + 1 < nc);
      if (c > 0) { rowval[nnz] = i - nr + 1; nzval[nnz++] = -1.0; }
      if (r > 0) { rowval[nnz] = i - 1 + 1; nzval[nnz++] = -1.0; }
      rowval[nnz] = i + 1; nzval[nnz++] = (double)deg;
      if (r + 1 < nr) { rowval[nnz] = i + 1 + 1; nzval[nnz++] = -1.0; }
      if (c + 1 < nc) { rowval[nnz] = i + nr + 1; nzval[nnz++] = -1.0; }
    }
  colptr[n] = nnz + 1;
  cs_b200_opts opts = {0};
  opts.precond = CS_B200_PRECOND_AMG;
  cs_b200_handle* h = NULL;
  int rc = cs_b200_create(n, nnz, colptr, rowval, nzval, 64, 1, CS_B200_F64, 0, &opts, &h);
  if (rc) { fprintf(stderr, "create failed: %d %s\n", rc, cs_b200_last_error(NULL)); return 2; }
  const int k = 2;
  double* rhs = (double*)calloc((size_t)n * k, sizeof(double));
  double* lhs = (double*)calloc((size_t)n * k, sizeof(double));
  const int64_t s0 = 0, d0 = n - 1, s1 = nr - 1, d1 = n - nr;     /* opposite corners */
  rhs[s0] = -1.0; rhs[d0] = 1.0;
  rhs[n + s1] += -1.0; rhs[n + d1] += 1.0;
  int64_t iters[2];
  double relres[2];
  rc = cs_b200_solve_rhs(h, k, rhs, lhs, 1e-10, 100000, iters, relres);
  if (rc) { fprintf(stderr, "solve failed: %d %s\n", rc, cs_b200_last_error(h)); return 3; }
  const double R0 = lhs[d0] - lhs[s0], R1 = lhs[n + d1] - lhs[n + s1];
  printf("n=%lld nnz=%lld R0=%.12f R1=%.12f iters=%lld,%lld relres=%.2e,%.2e\n", (long long)n, (long long)nnz, R0, R1,
         (long long)iters[0], (long long)iters[1], relres[0], relres[1]);
  int bad = !(relres[0] < 1e-4 && relres[1] < 1e-4);
  if (nr == 1) bad |= fabs(R0 - (double)(nc - 1)) > 1e-6 * (nc - 1);   /* nc - 1 unit resistors in series */
  bad |= fabs(R0 - R1) > 1e-6 * fabs(R0);                               /* the grid is symmetric */
  cs_b200_stats st;
  cs_b200_get_stats(h, &st);
  printf("setup_ms=%.2f solve_ms=%.2f launches=%lld\n", st.setup_ms, st.solve_ms, (long long)st.kernel_launches);
  cs_b200_destroy(h);
  free(colptr); free(rowval); free(nzval); free(rhs); free(lhs);
  return bad ? 1 : 0;
}

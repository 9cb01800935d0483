// amg_host.hpp -- host-side setup of the smoothed-aggregation hierarchy used as the
// CG preconditioner on the device (role of `smoothed_aggregation(matrix; ...)` at
// src/core.jl:164-167 and src/raster/advanced.jl:308 of the reference).
//
// The hierarchy is built ONCE per connected component on the host from the CSR the
// caller hands to cs_b200_create, then uploaded; every V-cycle runs on the GPU
// (DESIGN.md §5).  Algorithm (Vanek/Mandel/Brezina smoothed aggregation):
//   strength      every off-diagonal nonzero (symmetric strength, theta = 0)
//   aggregation   greedy root + neighbourhood, leftovers join a neighbour
//   tentative T   piecewise constant, columns normalised (candidate = ones)
//   prolongator   P = (I - (4/3)/rho * D^-1 A) T,  rho ~ lambda_max(D^-1 A) (power iteration),
//                 capped by ||D^-1 A||_inf
//   coarse op     A_c = P^T A P  (Galerkin)
//   coarsest      (<= 200 nodes) dense symmetric pseudo-inverse, cyclic Jacobi eigen-solver
// The smoother on the device is damped Jacobi with omega = (4/3)/rho_l per level, so
// the V(1,1) cycle is a symmetric positive (semi-)definite operator as CG requires.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace csb_amg {

// OpenMP team size for a loop over `work` items: serial below 2e5 (the golden-test sized
// components would only pay fork/join latency), at most 16 threads above (setup shares the
// host with the other ranks of a multi-GPU job)
inline int team(int64_t work) {
#ifdef _OPENMP
  return work < 200000 ? 1 : std::min(16, omp_get_max_threads());
#else
  (void)work; return 1;
#endif
}


struct Csr {
  int64_t nrows = 0, ncols = 0;
  std::vector<int> ptr, idx;
  std::vector<double> val;
  int64_t nnz() const { return (int64_t)idx.size(); }
};

struct HostLevel {
  Csr A;                       // operator of this level
  Csr P;                       // prolongator to this level from the next (nrows = n_l, ncols = n_{l+1})
  Csr R;                       // P^T
  std::vector<double> dinv;    // 1/diag(A)
  double omega = 2.0 / 3.0;    // Jacobi damping of this level
};

struct Hierarchy {
  std::vector<HostLevel> levels;     // levels.back() has no P/R
  std::vector<double> coarse_pinv;   // dense n_c x n_c (row-major) pseudo-inverse of levels.back().A
  double operator_complexity() const {
    double s = 0;
    for (auto& l : levels) s += (double)l.A.nnz();
    return s / (double)levels[0].A.nnz();
  }
};

inline Csr transpose(const Csr& a) {
  Csr t;
  t.nrows = a.ncols; t.ncols = a.nrows;
  t.ptr.assign(t.nrows + 1, 0);
  for (int c : a.idx) t.ptr[c + 1]++;
  for (int64_t i = 0; i < t.nrows; ++i) t.ptr[i + 1] += t.ptr[i];
  t.idx.resize(a.idx.size());
  t.val.resize(a.val.size());
  std::vector<int> cur(t.ptr.begin(), t.ptr.end() - 1);
  for (int64_t r = 0; r < a.nrows; ++r)
    for (int j = a.ptr[r]; j < a.ptr[r + 1]; ++j) {
      const int d = cur[a.idx[j]]++;
      t.idx[d] = (int)r;
      t.val[d] = a.val[j];
    }
  return t;
}

// C = A * B (Gustavson; columns sorted on output).  Two passes over contiguous row chunks
// dealt to OpenMP threads: pass 1 counts the distinct columns of every output row, a prefix
// sum sizes the result, pass 2 accumulates and writes straight into it (no per-thread
// buffers to merge; the result does not depend on the thread count).  `max_nnz` > 0 is a
// budget: if the product would have more entries the multiplication is abandoned after the
// counting pass and `*overflow` set (the caller is only probing whether coarsening pays).
inline Csr spgemm(const Csr& a, const Csr& b, int64_t max_nnz = 0, bool* overflow = nullptr) {
  Csr c;
  c.nrows = a.nrows; c.ncols = b.ncols;
  c.ptr.assign(c.nrows + 1, 0);
  if (overflow) *overflow = false;
  const int64_t n = a.nrows;
  const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(512, n / 1024));
  const int nt = team(a.nnz());
  std::vector<int64_t> chunk_nnz(nchunk, 0);
#pragma omp parallel num_threads(nt)
  {
    std::vector<int> marker(b.ncols, -1);
#pragma omp for schedule(dynamic, 1)
    for (int ch = 0; ch < nchunk; ++ch) {
      int64_t tot = 0;
      for (int64_t i = n * ch / nchunk; i < n * (ch + 1) / nchunk; ++i) {
        int cnt = 0;
        for (int ja = a.ptr[i]; ja < a.ptr[i + 1]; ++ja) {
          const int k = a.idx[ja];
          for (int jb = b.ptr[k]; jb < b.ptr[k + 1]; ++jb) {
            const int col = b.idx[jb];
            if (marker[col] != (int)i) { marker[col] = (int)i; ++cnt; }
          }
        }
        c.ptr[i + 1] = cnt;
        tot += cnt;
      }
      chunk_nnz[ch] = tot;
    }
  }
  int64_t total = 0;
  for (int ch = 0; ch < nchunk; ++ch) total += chunk_nnz[ch];
  if ((max_nnz > 0 && total > max_nnz) || total > (int64_t)std::numeric_limits<int>::max()) {
    if (overflow) *overflow = true;
    c.ptr.assign(c.nrows + 1, 0);
    return c;
  }
  for (int64_t i = 0; i < n; ++i) c.ptr[i + 1] += c.ptr[i];
  c.idx.resize((size_t)total);
  c.val.resize((size_t)total);
#pragma omp parallel num_threads(nt)
  {
    std::vector<int> marker(b.ncols, -1);
    std::vector<double> acc(b.ncols, 0.0);
#pragma omp for schedule(dynamic, 1)
    for (int ch = 0; ch < nchunk; ++ch) {
      for (int64_t i = n * ch / nchunk; i < n * (ch + 1) / nchunk; ++i) {
        int* oc = c.idx.data() + c.ptr[i];
        int cnt = 0;
        for (int ja = a.ptr[i]; ja < a.ptr[i + 1]; ++ja) {
          const int k = a.idx[ja];
          const double av = a.val[ja];
          for (int jb = b.ptr[k]; jb < b.ptr[k + 1]; ++jb) {
            const int col = b.idx[jb];
            if (marker[col] != (int)i) { marker[col] = (int)i; acc[col] = 0.0; oc[cnt++] = col; }
            acc[col] += av * b.val[jb];
          }
        }
        std::sort(oc, oc + cnt);
        double* ov = c.val.data() + c.ptr[i];
        for (int q = 0; q < cnt; ++q) ov[q] = acc[oc[q]];
      }
    }
  }
  return c;
}

// Greedy aggregation over the off-diagonal pattern.  agg[i] in [0, nagg) or -1 (isolated).
inline int aggregate(const Csr& a, std::vector<int>& agg) {
  const int64_t n = a.nrows;
  agg.assign(n, -1);
  std::vector<char> isolated(n, 0);
  int nagg = 0;
  // phase 1: a node whose whole neighbourhood is still free seeds an aggregate
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    bool any = false, all_free = true;
    for (int j = a.ptr[i]; j < a.ptr[i + 1] && all_free; ++j) {
      const int c = a.idx[j];
      if (c == i || a.val[j] == 0.0) continue;
      any = true;
      if (agg[c] >= 0) all_free = false;
    }
    if (!any) { isolated[i] = 1; continue; }
    if (!all_free) continue;
    agg[i] = nagg;
    for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j)
      if (a.idx[j] != i && a.val[j] != 0.0) agg[a.idx[j]] = nagg;
    ++nagg;
  }
  // phase 2: leftovers join the aggregate of their strongest already-aggregated neighbour
  const std::vector<int> seeded(agg);
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] >= 0 || isolated[i]) continue;
    double best = 0.0;
    int pick = -1;
    for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j) {
      const int c = a.idx[j];
      if (c == i || seeded[c] < 0) continue;
      const double w = std::fabs(a.val[j]);
      if (w > best) { best = w; pick = seeded[c]; }
    }
    if (pick >= 0) agg[i] = pick;
  }
  // phase 3: whatever is left groups with its free neighbours
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] >= 0 || isolated[i]) continue;
    agg[i] = nagg;
    for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j) {
      const int c = a.idx[j];
      if (c != i && agg[c] < 0 && !isolated[c] && a.val[j] != 0.0) agg[c] = nagg;
    }
    ++nagg;
  }
  return nagg;
}


// ---- MIS(2) aggregation: the PARALLEL scheme the device setup runs (setup_device.cu), restated
// sequentially with the same synchronous rounds, the same hash and the same tie rules, so that
// host and device produce identical aggregates (tests compare them).
//   roots      = a maximal independent set of the distance-2 strength graph (Bell/Dalton/Olson):
//                every undecided node carries the key (state, hash(i), i); after two rounds of
//                neighbourhood max-propagation a node whose own key came back is a root, a node
//                that saw a root's key is out
//   aggregates = root + its neighbours (strongest root wins) + distance-2 nodes (strongest
//                already-aggregated neighbour wins); numbered by root id order
inline uint32_t mis_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
inline uint64_t mis_key(uint32_t state, uint32_t i) {
  return ((uint64_t)state << 62) | ((uint64_t)(mis_hash(i) & 0x3fffffffU) << 32) | (uint64_t)i;
}
inline int aggregate_mis2(const Csr& a, std::vector<int>& agg) {
  const int64_t n = a.nrows;
  enum : uint32_t { OUT = 0, UND = 1, IN = 2 };
  std::vector<uint64_t> key(n), t1(n), t2(n);
  std::vector<char> isolated(n, 0);
  int64_t undecided = 0;
  for (int64_t i = 0; i < n; ++i) {
    bool any = false;
    for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j)
      if (a.idx[j] != i && a.val[j] != 0.0) { any = true; break; }
    isolated[i] = !any;
    key[i] = mis_key(any ? UND : OUT, (uint32_t)i);
    undecided += any;
  }
  auto nbmax = [&](const std::vector<uint64_t>& in, std::vector<uint64_t>& out) {
    for (int64_t i = 0; i < n; ++i) {
      uint64_t m = in[i];
      for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j)
        if (a.idx[j] != i && a.val[j] != 0.0) m = std::max(m, in[a.idx[j]]);
      out[i] = m;
    }
  };
  while (undecided > 0) {
    nbmax(key, t1);
    nbmax(t1, t2);
    for (int64_t i = 0; i < n; ++i) {
      if ((key[i] >> 62) != UND) continue;
      if (t2[i] == key[i]) { key[i] = mis_key(IN, (uint32_t)i); --undecided; }
      else if ((t2[i] >> 62) == IN) { key[i] = mis_key(OUT, (uint32_t)i); --undecided; }
    }
  }
  agg.assign(n, -1);
  int nagg = 0;
  for (int64_t i = 0; i < n; ++i) if ((key[i] >> 62) == IN) agg[i] = nagg++;
  for (int pass = 0; pass < 2; ++pass) {
    const std::vector<int> snap(agg);
    for (int64_t i = 0; i < n; ++i) {
      if (snap[i] >= 0 || isolated[i]) continue;
      double best = -1.0;
      int pick = -1;
      for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j) {
        const int c = a.idx[j];
        if (c == i || a.val[j] == 0.0 || snap[c] < 0) continue;
        if (pass == 0 && (key[c] >> 62) != IN) continue;       // pass 0: roots only
        const double w = std::fabs(a.val[j]);
        if (w > best) { best = w; pick = snap[c]; }
      }
      if (pick >= 0) agg[i] = pick;
    }
  }
  return nagg;
}

// dense symmetric pseudo-inverse (n <= a few hundred) from the eigen-decomposition A = V diag(d) V^T:
// Householder reduction to tridiagonal form, then implicit-shift QL with accumulated transformations
// (the classic EISPACK tred2 / tql2 pair as restated in the public-domain JAMA package) -- ~4 n^3 flops
// instead of the ~25 n^3 of the cyclic Jacobi sweeps used in round 1 (110 ms -> ~15 ms at n = 196).
// Eigenvalues below 1e-10 n lambda_max are treated as the null space (singular Neumann operators).
inline void sym_tridiag(int n, std::vector<double>& V, std::vector<double>& d, std::vector<double>& e) {
  auto v = [&](int i, int j) -> double& { return V[(size_t)i * n + j]; };
  for (int j = 0; j < n; ++j) d[j] = v(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = v(i - 1, j); v(i, j) = 0.0; v(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        v(j, i) = f;
        g = e[j] + v(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += v(k, j) * d[k]; e[k] += v(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; ++k) v(k, j) -= (f * e[k] + g * d[k]);
        d[j] = v(i - 1, j);
        v(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {           // accumulate the transformations
    v(n - 1, i) = v(i, i);
    v(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = v(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += v(k, i + 1) * v(k, j);
        for (int k = 0; k <= i; ++k) v(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) v(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = v(n - 1, j); v(n - 1, j) = 0.0; }
  v(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
}

inline void tridiag_ql(int n, std::vector<double>& V, std::vector<double>& d, std::vector<double>& e) {
  auto v = [&](int i, int j) -> double& { return V[(size_t)i * n + j]; };
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::numeric_limits<double>::epsilon();
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n - 1 && std::fabs(e[m]) > eps * tst1) ++m;
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = v(k, i + 1);
            v(k, i + 1) = s * v(k, i) + c * h;
            v(k, i) = c * v(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] += f;
    e[l] = 0.0;
  }
}

inline std::vector<double> dense_pinv(const Csr& a) {
  const int n = (int)a.nrows;
  std::vector<double> V((size_t)n * n, 0.0), d(n, 0.0), e(n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j) V[(size_t)i * n + a.idx[j]] += a.val[j];
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double s = 0.5 * (V[(size_t)i * n + j] + V[(size_t)j * n + i]);
      V[(size_t)i * n + j] = V[(size_t)j * n + i] = s;
    }
  std::vector<double> out((size_t)n * n, 0.0);
  if (n == 0) return out;
  if (n == 1) { out[0] = V[0] != 0.0 ? 1.0 / V[0] : 0.0; return out; }
  sym_tridiag(n, V, d, e);
  tridiag_ql(n, V, d, e);
  double lmax = 0.0;
  for (int i = 0; i < n; ++i) lmax = std::max(lmax, std::fabs(d[i]));
  const double cut = lmax * 1e-10 * std::max(1, n);
  // out = sum over kept eigenpairs  v v^T / lambda ; built as W W^T with W = V diag(1/sqrt|lambda|) sign-aware
  for (int ev = 0; ev < n; ++ev) {
    const double lam = d[ev];
    if (std::fabs(lam) <= cut) continue;
    const double inv = 1.0 / lam;
    for (int i = 0; i < n; ++i) {
      const double vi = V[(size_t)i * n + ev] * inv;
      if (vi == 0.0) continue;
      double* o = out.data() + (size_t)i * n;
      for (int j = 0; j < n; ++j) o[j] += vi * V[(size_t)j * n + ev];
    }
  }
  return out;
}

// dinv = 1/diag(A);  rho = estimate of lambda_max(D^-1 A): 8 power iterations (Rayleigh
// quotient in the D inner product, fixed start vector => deterministic), kept inside
// [0.7, 1] x the rigorous bound ||D^-1 A||_inf.  For raster stencils the bound is 2
// while lambda_max ~ 1.6; the sharper value gives a larger Jacobi / prolongator-smoothing
// weight (omega = (4/3)/rho) and ~12 % fewer CG iterations.
inline void diag_and_rho(const Csr& a, std::vector<double>& dinv, double& rho) {
  const int64_t n = a.nrows;
  dinv.assign(n, 0.0);
  double rho_inf = 0.0;
#pragma omp parallel for reduction(max : rho_inf) schedule(static) num_threads(team(a.nnz()))
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0, s = 0.0;
    for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j) {
      if (a.idx[j] == i) d += a.val[j];
      s += std::fabs(a.val[j]);
    }
    if (d != 0.0) {
      dinv[i] = 1.0 / d;
      rho_inf = std::max(rho_inf, s / std::fabs(d));
    }
  }
  if (!(rho_inf > 0.0)) { rho = 1.0; return; }
  std::vector<double> x(n), y(n);
  for (int64_t i = 0; i < n; ++i) x[i] = dinv[i] != 0.0 ? 1.0 + (double)((i * 2654435761ULL) % 1024) / 1024.0 * ((i & 1) ? 1.0 : -1.0) : 0.0;
  double lam = 0.0;
  const int nchunk = 64;   // fixed partition => the sums do not depend on the thread count
  const int nt = team(a.nnz());
  for (int it = 0; it < 8; ++it) {
    double pn[nchunk], pd[nchunk], pm[nchunk];
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int ch = 0; ch < nchunk; ++ch) {
      double sn = 0.0, sd = 0.0, sm = 0.0;
      for (int64_t i = n * ch / nchunk; i < n * (ch + 1) / nchunk; ++i) {
        double acc = 0.0;
        for (int j = a.ptr[i]; j < a.ptr[i + 1]; ++j) acc += a.val[j] * x[a.idx[j]];
        sn += x[i] * acc;                                   // x' A x
        sd += dinv[i] != 0.0 ? x[i] * x[i] / dinv[i] : 0.0; // x' D x
        y[i] = dinv[i] * acc;
        sm = std::max(sm, std::fabs(y[i]));
      }
      pn[ch] = sn; pd[ch] = sd; pm[ch] = sm;
    }
    double num = 0.0, den = 0.0, nrm = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) { num += pn[ch]; den += pd[ch]; nrm = std::max(nrm, pm[ch]); }
    if (!(den > 0.0)) break;
    lam = num / den;
    if (!(nrm > 0.0)) break;
    const double inv = 1.0 / nrm;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n; ++i) x[i] = y[i] * inv;
  }
  // lam is a Rayleigh quotient, i.e. a LOWER bound of lambda_max <= rho_inf.  The floor
  // 0.7 rho_inf keeps omega * lambda_max <= (4/3)/0.7 < 2 whatever the estimate did, so the
  // Jacobi smoother stays convergent and the V-cycle positive definite.
  rho = lam > 0.0 ? std::min(rho_inf, std::max(lam, 0.7 * rho_inf)) : rho_inf;
}

inline Hierarchy build_hierarchy(Csr a0, int max_levels = 12, int max_coarse = 200, bool mis2 = false) {
  Hierarchy h;
  h.levels.emplace_back();
  h.levels.back().A = std::move(a0);
  for (;;) {
    HostLevel& lv = h.levels.back();
    double rho;
    diag_and_rho(lv.A, lv.dinv, rho);
    lv.omega = (4.0 / 3.0) / rho;
    const int64_t n = lv.A.nrows;
    if ((int)h.levels.size() >= max_levels || n <= max_coarse) break;
    std::vector<int> agg;
    const int nagg = mis2 ? aggregate_mis2(lv.A, agg) : aggregate(lv.A, agg);
    if (nagg <= 0 || nagg >= n) break;
    std::vector<double> cnt(nagg, 0.0);
    for (int64_t i = 0; i < n; ++i) if (agg[i] >= 0) cnt[agg[i]] += 1.0;
    // S = I - omega D^-1 A   applied to T on the fly:  P = T - omega D^-1 (A T)
    Csr T;
    T.nrows = n; T.ncols = nagg;
    T.ptr.assign(n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
      if (agg[i] >= 0) { T.idx.push_back(agg[i]); T.val.push_back(1.0 / std::sqrt(cnt[agg[i]])); }
      T.ptr[i + 1] = (int)T.idx.size();
    }
    Csr AT = spgemm(lv.A, T);
    Csr P;
    P.nrows = n; P.ncols = nagg;
    P.ptr.assign(n + 1, 0);
    P.idx.reserve(AT.idx.size());
    P.val.reserve(AT.idx.size());
    for (int64_t i = 0; i < n; ++i) {
      const double sc = lv.omega * lv.dinv[i];
      const int mine = agg[i];
      const double tv = mine >= 0 ? 1.0 / std::sqrt(cnt[mine]) : 0.0;
      bool placed = mine < 0;
      for (int j = AT.ptr[i]; j < AT.ptr[i + 1]; ++j) {
        const int c = AT.idx[j];
        double v = -sc * AT.val[j];
        if (!placed && c > mine) { P.idx.push_back(mine); P.val.push_back(tv); placed = true; }
        if (c == mine) { v += tv; placed = true; }
        P.idx.push_back(c);
        P.val.push_back(v);
      }
      if (!placed) { P.idx.push_back(mine); P.val.push_back(tv); }
      P.ptr[i + 1] = (int)P.idx.size();
    }
    // expander-like graphs (power-law networks) densify under smoothed aggregation: a coarse
    // operator with more entries than the one it came from buys nothing on a bandwidth-bound
    // machine -- stop here (the cycle ends at this level; with no level at all the solver is
    // plain Jacobi-PCG, which such graphs like: 27 iterations on a 2e5-node BA graph).  The
    // products carry an nnz budget so that finding this out costs O(nnz), not a dense fill.
    bool over = false;
    Csr AP = spgemm(lv.A, P, 4 * lv.A.nnz(), &over);
    if (over) break;
    Csr R = transpose(P);
    Csr Ac = spgemm(R, AP, lv.A.nnz(), &over);
    if (over || Ac.nnz() > lv.A.nnz()) break;
    lv.P = std::move(P);
    lv.R = std::move(R);
    h.levels.emplace_back();
    h.levels.back().A = std::move(Ac);
  }
  // exact coarse solve only while the dense eigen-solve stays cheap; otherwise (coarsening
  // stalled on an irregular graph) the device falls back to a few Jacobi sweeps there
  if (h.levels.back().A.nrows <= 320) h.coarse_pinv = dense_pinv(h.levels.back().A);
  return h;
}

}  // namespace csb_amg

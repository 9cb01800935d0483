// Test harness (tests only): exposes circuitscape_b200/csrc/amg_host.hpp to Python so
// the host-built hierarchy can be checked on a box without a GPU.
#include "../circuitscape_b200/csrc/amg_host.hpp"
#include <cstring>

using namespace csb_amg;
extern "C" {
void* amgh_build(long n, long nnz, const int* ptr, const int* idx, const double* val) {
  Csr a; a.nrows = a.ncols = n;
  a.ptr.assign(ptr, ptr + n + 1); a.idx.assign(idx, idx + nnz); a.val.assign(val, val + nnz);
  return new Hierarchy(build_hierarchy(std::move(a)));
}
void* amgh_build2(long n, long nnz, const int* ptr, const int* idx, const double* val, int mis2) {
  Csr a; a.nrows = a.ncols = n;
  a.ptr.assign(ptr, ptr + n + 1); a.idx.assign(idx, idx + nnz); a.val.assign(val, val + nnz);
  return new Hierarchy(build_hierarchy(std::move(a), 12, 200, mis2 != 0));
}
int amgh_nlevels(void* h) { return (int)((Hierarchy*)h)->levels.size(); }
static const Csr& pick(void* h, int l, int which) {
  HostLevel& L = ((Hierarchy*)h)->levels[l];
  return which == 0 ? L.A : which == 1 ? L.P : L.R;
}
void amgh_dims(void* h, int l, int which, long* nrows, long* ncols, long* nnz, double* omega) {
  const Csr& m = pick(h, l, which);
  *nrows = m.nrows; *ncols = m.ncols; *nnz = m.nnz(); *omega = ((Hierarchy*)h)->levels[l].omega;
}
void amgh_copy(void* h, int l, int which, int* ptr, int* idx, double* val) {
  const Csr& m = pick(h, l, which);
  std::memcpy(ptr, m.ptr.data(), m.ptr.size() * sizeof(int));
  std::memcpy(idx, m.idx.data(), m.idx.size() * sizeof(int));
  std::memcpy(val, m.val.data(), m.val.size() * sizeof(double));
}
void amgh_pinv(void* h, double* out) {
  auto& p = ((Hierarchy*)h)->coarse_pinv;
  std::memcpy(out, p.data(), p.size() * sizeof(double));
}
void amgh_free(void* h) { delete (Hierarchy*)h; }
}

// ---- windowed row-block form (win_host.hpp): host emulation of what k_spmm_win does ----
#include "../circuitscape_b200/csrc/win_host.hpp"
extern "C" long winh_spmv_rect(long n, long ncols, long nnz, const int* ptr, const int* idx, const double* val,
                          const double* x, double* y, long* nblocks, long* max_wrows);
extern "C" long winh_spmv(long n, long nnz, const int* ptr, const int* idx, const double* val,
                          const double* x, double* y, long* nblocks, long* max_wrows) {
  return winh_spmv_rect(n, n, nnz, ptr, idx, val, x, y, nblocks, max_wrows);
}
extern "C" long winh_spmv_rect(long n, long ncols, long nnz, const int* ptr, const int* idx, const double* val,
                          const double* x, double* y, long* nblocks, long* max_wrows) {
  const long n_pad = (ncols + 3) / 4 * 4;
  std::vector<double> xp(n_pad, 0.0);
  std::memcpy(xp.data(), x, ncols * sizeof(double));
  const int wcap = (double)nnz / (double)n >= 20.0 ? csb_win::WCAP_WIDE : csb_win::WCAP;   // as cs_b200.cu
  csb_win::Windowed w = csb_win::build(ptr, idx, n, n_pad, 8, wcap);
  *nblocks = (long)w.meta.size();
  *max_wrows = 0;
  std::vector<double> win(csb_win::WCAP_WIDE);
  for (const auto& m : w.meta) {
    if (m.nseg == 0) {
      for (int row = m.row0; row < m.row0 + m.nrows; ++row) {
        double s = 0;
        for (int j = ptr[row]; j < ptr[row + 1]; ++j) s += val[j] * xp[idx[j]];
        y[row] = s;
      }
      continue;
    }
    if (m.wrows > *max_wrows) *max_wrows = m.wrows;
    int slot = 0;
    for (int k = 0; k < m.nseg; ++k) {
      if (m.seg_lo[k] % csb_win::ALN || m.seg_len[k] % csb_win::ALN || m.seg_lo[k] + m.seg_len[k] > n_pad) return -1;
      for (int i = 0; i < m.seg_len[k]; ++i) win[slot + i] = xp[m.seg_lo[k] + i];
      slot += m.seg_len[k];
    }
    if (slot != m.wrows || m.ent_off % 8 || w.roff_off[&m - w.meta.data()] % 8 || m.blob_off16 < 0) return -2;
    for (int rl = 0; rl < m.nrows; ++rl) {
      const int ro = w.roff_off[&m - w.meta.data()];
      const int a = w.roff[ro + rl], b = w.roff[ro + rl + 1];
      double s = 0;
      for (int j = a; j < b; ++j) {
        const int p = w.perm_off[m.ent_off + j];
        if (p < 0) return -3;
        s += val[p] * win[w.lcol[m.ent_off + j]];
      }
      y[m.row0 + rl] = s;
      if (n == ncols && m.self_slot >= 0 && win[m.self_slot + rl] != xp[m.row0 + rl]) return -4;
    }
  }
  return w.windowed_blocks;
}

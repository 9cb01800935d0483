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

// win_host.hpp -- host-side construction of the "windowed row-block" form of a CSR
// operator, the layout the TMA-staged SpMM kernel (k_spmm_win, kernels.cuh) consumes.
//
// For every row block (<= 128 consecutive rows, <= 1152 nnz) the set of columns it
// touches is covered by a few contiguous *segments* of the X panel (for the raster
// stencil with column-major numbering: three strips of ~130 rows).  The kernel bulk-
// copies (cp.async.bulk + mbarrier) those segments, the block's slice of values, a
// 16-bit *window-local* column index per entry and the block's row offsets into
// shared memory, double-buffered across blocks, and then works out of shared memory
// only.  Blocks whose columns do not fit (hub rows, scattered graphs) are flagged
// nseg = 0 and take the direct-gather path on the plain CSR.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace csb_win {

constexpr int RB = 128;        // rows per block
constexpr int NNZ_CAP = 1152;  // entries per block (128 rows x 9)
constexpr int WCAP = 512;      // X rows staged per block (all segments together)
constexpr int WCAP_WIDE = 1024; // ... for wide-row operators (restrictions)
constexpr int MAXSEG = 8;
constexpr int ALN = 4;         // segment start/length granularity in X rows (16 B at KT=1, fp32)
constexpr int MERGE_GAP = 16;  // runs closer than this are merged into one segment

struct BlockMeta {             // 96 bytes, one per row block (device + host identical)
  int row0, nrows;
  int nnz;                     // entries of the block
  int ent_off;                 // first entry in vals_p / lcol_p (multiple of 8)
  int blob_off16;              // byte offset / 16 of the block's record in the packed blob:
                               //   [ values nnzp*sizeof(T) | (1/diag rowsp*sizeof(T)) | local cols nnzp*2 | row offsets roffp*2 ]
  int nseg;                    // 0 => direct-gather path
  int self_slot;               // window slot of X[row0] if rows row0.. are contiguous in it, else -1
  int wrows;                   // total staged X rows
  int seg_lo[MAXSEG];
  int seg_len[MAXSEG];
};

struct Windowed {
  std::vector<BlockMeta> meta;
  std::vector<uint16_t> lcol;   // packed, per block padded to 8
  std::vector<int> perm_off;    // per packed entry: index into the CSR value array or -1 (padding)
  std::vector<uint16_t> roff;   // packed row offsets (nrows+1 per block, padded to 8)
  std::vector<int> roff_off;    // per block: first entry in roff
  int64_t windowed_blocks = 0;
  int64_t blob_bytes = 0;       // size of the device blob for value size `vsize`
};

// greedy row blocks of the windowed form: <= RB rows and <= NNZ_CAP entries; a longer row
// stands alone (and takes the direct-gather path)
inline std::vector<int> row_blocks(const int* rowptr, int64_t n) {
  std::vector<int> bstart{0};
  int64_t r = 0;
  while (r < n) {
    int64_t r1 = r + 1;
    const int64_t base = rowptr[r];
    while (r1 < n && (r1 - r) < RB && (int64_t)rowptr[r1 + 1] - base <= NNZ_CAP) ++r1;
    bstart.push_back((int)r1);
    r = r1;
  }
  return bstart;
}

inline Windowed build(const int* rowptr, const int* colidx, int64_t nrows,
                      int64_t ncols_pad /* X rows available (n_pad of the input panel) */,
                      int vsize = 8 /* sizeof(T) of the device values */, int wcap = WCAP,
                      bool with_dinv = false /* record carries 1/diag of the block's rows */) {
  Windowed w;
  const std::vector<int> bstart = row_blocks(rowptr, nrows);
  const int nb = (int)bstart.size() - 1;
  w.meta.resize(nb);
  // pass 1: sizes
  std::vector<int64_t> ent_off(nb + 1, 0), roff_off(nb + 1, 0);
  for (int b = 0; b < nb; ++b) {
    const int r0 = bstart[b], r1 = bstart[b + 1];
    const int cnt = rowptr[r1] - rowptr[r0];
    const bool fits = cnt <= NNZ_CAP && (r1 - r0) <= RB;
    ent_off[b + 1] = ent_off[b] + (fits ? (cnt + 7) / 8 * 8 : 0);
    roff_off[b + 1] = roff_off[b] + (fits ? (r1 - r0 + 1 + 7) / 8 * 8 : 0);
  }
  w.roff_off.resize(nb);
  {
    int64_t bo = 0;   // blob offsets (bytes, multiples of 16)
    for (int b = 0; b < nb; ++b) {
      const int r0 = bstart[b], r1 = bstart[b + 1];
      const int cnt = rowptr[r1] - rowptr[r0];
      const bool fits = cnt <= NNZ_CAP && (r1 - r0) <= RB;
      w.meta[b].blob_off16 = (int)(bo / 16);
      w.roff_off[b] = (int)roff_off[b];
      if (fits) bo += (int64_t)((cnt + 7) / 8 * 8) * (vsize + 2) + (int64_t)((r1 - r0 + 1 + 7) / 8 * 8) * 2 +
                      (with_dinv ? (int64_t)((r1 - r0 + 7) / 8 * 8) * vsize : 0);
    }
    w.blob_bytes = bo + 64;
  }
  w.lcol.assign((size_t)ent_off[nb] + 8, 0);
  w.perm_off.assign((size_t)ent_off[nb] + 8, -1);
  w.roff.assign((size_t)roff_off[nb] + 8, 0);
  int64_t nwin = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nwin) num_threads(nb < 512 ? 1 : 16)
  for (int b = 0; b < nb; ++b) {
    BlockMeta& m = w.meta[b];
    const int r0 = bstart[b], r1 = bstart[b + 1];
    const int s = rowptr[r0], e = rowptr[r1];
    const int keep_blob = m.blob_off16;
    m = BlockMeta{};
    m.blob_off16 = keep_blob;
    m.row0 = r0; m.nrows = r1 - r0; m.nnz = e - s;
    m.ent_off = (int)ent_off[b];
    const int my_roff = (int)roff_off[b];
    m.nseg = 0; m.self_slot = -1; m.wrows = 0;
    if (m.nnz > NNZ_CAP || m.nrows > RB || m.nnz == 0) continue;
    // unique sorted columns of the block
    std::vector<int> cols(colidx + s, colidx + e);
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    // runs -> segments (aligned to ALN rows, merged when close)
    int nseg = 0, total = 0;
    int lo[MAXSEG], len[MAXSEG];
    bool ok = true;
    size_t i = 0;
    while (i < cols.size()) {
      int a = cols[i] / ALN * ALN;
      int hi = cols[i] + 1;
      size_t j = i + 1;
      while (j < cols.size() && cols[j] - hi < MERGE_GAP) { hi = cols[j] + 1; ++j; }
      int l = (hi - a + ALN - 1) / ALN * ALN;
      if ((int64_t)a + l > ncols_pad) l = (int)(ncols_pad - a);
      if (nseg > 0 && a < lo[nseg - 1] + len[nseg - 1]) {   // alignment made it touch the previous one
        const int nl = a + l - lo[nseg - 1];
        total += nl - len[nseg - 1];
        len[nseg - 1] = nl;
      } else {
        if (nseg == MAXSEG) { ok = false; break; }
        lo[nseg] = a; len[nseg] = l; total += l; ++nseg;
      }
      i = j;
    }
    if (!ok || total > wcap) continue;
    m.nseg = nseg; m.wrows = total;
    int off[MAXSEG];
    int acc = 0;
    for (int k = 0; k < nseg; ++k) { m.seg_lo[k] = lo[k]; m.seg_len[k] = len[k]; off[k] = acc; acc += len[k]; }
    // window-local column of every entry, packed values permutation, row offsets
    for (int j = s; j < e; ++j) {
      const int c = colidx[j];
      int k = 0;
      while (k + 1 < nseg && c >= lo[k + 1]) ++k;
      w.lcol[(size_t)m.ent_off + (j - s)] = (uint16_t)(off[k] + (c - lo[k]));
      w.perm_off[(size_t)m.ent_off + (j - s)] = j;
    }
    for (int r = r0; r <= r1; ++r) w.roff[(size_t)my_roff + (r - r0)] = (uint16_t)(rowptr[r] - s);
    // are the block's own rows one contiguous stretch of the window?
    for (int k = 0; k < nseg; ++k)
      if (r0 >= lo[k] && r1 <= lo[k] + len[k]) { m.self_slot = off[k] + (r0 - lo[k]); break; }
    nwin += 1;
  }
  w.windowed_blocks = nwin;
  return w;
}

}  // namespace csb_win

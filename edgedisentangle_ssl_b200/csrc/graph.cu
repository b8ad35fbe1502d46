// Host-side graph builder + device graph handle (CSR / CSC / chunked work schedules).
// Integer work: must be bit-exact against the reference's `load_data` pipeline
// (/root/reference/data_load.py:39-77, utils.py:163-170) -- see include/edis.h.
#include <algorithm>
#include <functional>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "edis_common.cuh"

namespace edis {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int launch_grid(const void* kernel, int block, size_t smem, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  return per_sm * sm_count;
}

// Stable counting sort of `len` records by key[] in [0, n): fills order[] with source indices.
static void counting_order(int64_t n, int64_t len, const int64_t* key, const int64_t* src_order,
                           int64_t* out_order, std::vector<int64_t>& cnt) {
  cnt.assign(static_cast<size_t>(n) + 1, 0);
  for (int64_t k = 0; k < len; ++k) cnt[key[src_order ? src_order[k] : k] + 1]++;
  for (int64_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
  for (int64_t k = 0; k < len; ++k) {
    const int64_t s = src_order ? src_order[k] : k;
    out_order[cnt[key[s]]++] = s;
  }
}

static Schedule build_schedule_host(int64_t n, const int64_t* ptr, int max_chunk,
                                    std::vector<Item>& items, std::vector<SplitRow>& split) {
  Schedule s;
  items.clear();
  split.clear();
  items.reserve(static_cast<size_t>(n) + 1024);
  int64_t slots = 0;
  for (int64_t r = 0; r < n; ++r) {
    const int64_t b = ptr[r], e = ptr[r + 1], deg = e - b;
    if (deg <= max_chunk) {
      items.push_back({static_cast<int32_t>(r), static_cast<int32_t>(b), static_cast<int32_t>(e), -1});
    } else {
      const int64_t nch = (deg + max_chunk - 1) / max_chunk;
      const int64_t sz = (deg + nch - 1) / nch;
      split.push_back({static_cast<int32_t>(r), static_cast<int32_t>(slots), static_cast<int32_t>(nch), 0});
      for (int64_t c = 0; c < nch; ++c) {
        const int64_t cb = b + c * sz, ce = std::min(e, cb + sz);
        items.push_back({static_cast<int32_t>(r), static_cast<int32_t>(cb), static_cast<int32_t>(ce),
                         static_cast<int32_t>(slots + c)});
      }
      slots += nch;
    }
  }
  s.n_items = static_cast<int64_t>(items.size());
  s.n_slots = slots;
  s.n_split = static_cast<int64_t>(split.size());
  return s;
}

template <class T>
static int upload(T** dev, const T* host, size_t count) {
  *dev = nullptr;
  if (count == 0) count = 1;
  EDIS_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), count * sizeof(T)));
  if (host) EDIS_CUDA(cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
  return EDIS_OK;
}

}  // namespace edis

using namespace edis;

extern "C" const char* edis_last_error(void) { return g_err.c_str(); }
extern "C" const char* edis_version(void) { return "edis 0.1 sm_100a"; }

// ---------------------------------------------------------------------------------------
extern "C" int64_t edis_build_adjacency_host(int64_t n, int64_t m, const int64_t* rows,
                                             const int64_t* cols, const double* vals,
                                             int64_t* out_row, int64_t* out_col, float* out_val) {
  if (n <= 0 || m < 0 || (m > 0 && (!rows || !cols)) || !out_row || !out_col || !out_val) {
    set_error("edis_build_adjacency_host: bad arguments");
    return EDIS_ERR_ARG;
  }
  // Row buckets instead of a global sort: pass 1 validates and counts the directed records per row
  // (every off-diagonal entry (r, c) contributes (r, c) with flag 1 and its mirror (c, r) with flag
  // 2; np.fill_diagonal overwrites the diagonal, data_load.py:69), pass 2 scatters (col, value,
  // flag) into the row's bucket, pass 3 sorts each (short) bucket by column and merges duplicates.
  // Sequential reads, one scattered write per record, no permutation gathers.
  std::vector<int64_t> ptr(n + 1, 0);
  for (int64_t k = 0; k < m; ++k) {
    const int64_t r = rows[k], c = cols[k];
    if (r < 0 || r >= n || c < 0 || c >= n) {
      set_error("edis_build_adjacency_host: entry %lld = (%lld, %lld) out of range n=%lld",
                (long long)k, (long long)r, (long long)c, (long long)n);
      return EDIS_ERR_ARG;
    }
    if (r == c) continue;
    ++ptr[r + 1];
    ++ptr[c + 1];
  }
  for (int64_t r = 0; r < n; ++r) ptr[r + 1] += ptr[r];
  const int64_t len = ptr[n];
  struct Rec {
    int64_t col;
    double v;
    uint8_t fl;
  };
  std::vector<Rec> rec(static_cast<size_t>(len));
  {
    std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
    for (int64_t k = 0; k < m; ++k) {
      const int64_t r = rows[k], c = cols[k];
      if (r == c) continue;
      const double v = vals ? vals[k] : 1.0;
      rec[pos[r]++] = {c, v, 1};
      rec[pos[c]++] = {r, v, 2};
    }
  }
  // dedup: value = max(A_ij, A_ji) with absent = 0 (data_load.py:71), then drop zeros; the diagonal
  // (value 1) is merged at its row-major position; rows are normalised by their sum in emission order
  int64_t E = 0;
  for (int64_t r = 0; r < n; ++r) {
    Rec* b = rec.data() + ptr[r];
    Rec* e = rec.data() + ptr[r + 1];
    std::sort(b, e, [](const Rec& x, const Rec& y) { return x.col < y.col; });
    double deg = 0.0;
    bool diag_done = false;
    // unique columns + merged values, compacted in place (the write index never overtakes the read index)
    Rec* w = b;
    for (Rec* q = b; q < e;) {
      const int64_t c = q->col;
      double best = q->v;
      uint8_t fl = q->fl;
      for (++q; q < e && q->col == c; ++q) {
        best = std::max(best, q->v);
        fl |= q->fl;
      }
      if (fl != 3) best = std::max(best, 0.0);   // the other orientation is an implicit 0
      if (best != 0.0) *w++ = {c, best, fl};
    }
    // row sum in row-major order with the diagonal in place
    for (Rec* q = b; q < w; ++q) {
      if (!diag_done && q->col > r) {
        deg += 1.0;
        diag_done = true;
      }
      deg += q->v;
    }
    if (!diag_done) deg += 1.0;
    // normalize_adj (data_load.py:12-20): r_inv = rowsum**-1 (inf -> 0), values r_inv * a (float64)
    double rinv = std::pow(deg, -1.0);
    if (std::isinf(rinv)) rinv = 0.0;
    diag_done = false;
    for (Rec* q = b; q < w; ++q) {
      if (!diag_done && q->col > r) {
        out_row[E] = r; out_col[E] = r; out_val[E] = static_cast<float>(rinv * 1.0); ++E;
        diag_done = true;
      }
      out_row[E] = r; out_col[E] = q->col; out_val[E] = static_cast<float>(rinv * q->v); ++E;
    }
    if (!diag_done) {
      out_row[E] = r; out_col[E] = r; out_val[E] = static_cast<float>(rinv * 1.0); ++E;
    }
  }
  return E;
}

// ---------------------------------------------------------------------------------------
extern "C" int edis_graph_create(int64_t n, int64_t e_in, const int64_t* row, const int64_t* col,
                                 int max_chunk, int device, edis_graph** out) {
  return edis_graph_create_rect(n, n, e_in, row, col, max_chunk, device, out);
}

extern "C" int edis_graph_create_rect(int64_t n, int64_t n_cols, int64_t e_in, const int64_t* row,
                                      const int64_t* col, int max_chunk, int device, edis_graph** out) {
  EDIS_CHECK_ARG(out && n > 0 && n_cols >= n && e_in >= 0 && (e_in == 0 || (row && col)),
                 "edis_graph_create: bad arguments (need n_cols >= n_rows > 0)");
  EDIS_CHECK_ARG(e_in < (int64_t(1) << 31) - 64, "edis_graph_create: e must fit int32");
  // bits 30-31 of a device-side neighbour id carry its L2 heat level and every kernel masks them off
  EDIS_CHECK_ARG(n_cols < (int64_t(1) << kHeatShift),
                 "edis_graph_create: node ids must stay below 2^%d (got n_cols=%lld)", kHeatShift, (long long)n_cols);
  if (max_chunk <= 0) max_chunk = 256;
  bool sorted = true;
  for (int64_t k = 0; k < e_in; ++k) {
    if (row[k] < 0 || row[k] >= n || col[k] < 0 || col[k] >= n_cols) {
      set_error("edis_graph_create: entry %lld out of range", (long long)k);
      return EDIS_ERR_ARG;
    }
    if (k > 0 && (row[k] < row[k - 1] || (row[k] == row[k - 1] && col[k] <= col[k - 1]))) sorted = false;
  }
  edis_graph* g = new edis_graph();
  g->n = n;
  g->n_cols = n_cols;
  g->e_in = e_in;
  g->device = device;
  g->was_sorted = sorted;
  g->h_perm = new int64_t[std::max<int64_t>(e_in, 1)];
  std::vector<int64_t> srow, scol;
  if (sorted) {
    for (int64_t k = 0; k < e_in; ++k) g->h_perm[k] = k;
    srow.assign(row, row + e_in);
    scol.assign(col, col + e_in);
  } else {
    // stable (col, then row) counting sorts + coalesce duplicates (adj.coalesce(), layers.py:344)
    std::vector<int64_t> o1(e_in), o2(e_in), cnt;
    counting_order(n_cols, e_in, col, nullptr, o1.data(), cnt);
    counting_order(n, e_in, row, o1.data(), o2.data(), cnt);
    srow.reserve(e_in);
    scol.reserve(e_in);
    for (int64_t k = 0; k < e_in; ++k) {
      const int64_t s = o2[k];
      if (srow.empty() || srow.back() != row[s] || scol.back() != col[s]) {
        srow.push_back(row[s]);
        scol.push_back(col[s]);
      }
      g->h_perm[s] = static_cast<int64_t>(srow.size()) - 1;
    }
  }
  const int64_t e = static_cast<int64_t>(srow.size());
  g->e = e;
  g->h_rowptr = new int64_t[n + 1]();
  g->h_col = new int32_t[std::max<int64_t>(e, 1)];
  g->h_cscptr = new int64_t[n_cols + 1]();
  g->h_cscrow = new int32_t[std::max<int64_t>(e, 1)];
  g->h_csceid = new int32_t[std::max<int64_t>(e, 1)];
  for (int64_t k = 0; k < e; ++k) {
    g->h_rowptr[srow[k] + 1]++;
    g->h_cscptr[scol[k] + 1]++;
    g->h_col[k] = static_cast<int32_t>(scol[k]);
  }
  for (int64_t i = 0; i < n; ++i) {
    g->max_in = std::max(g->max_in, g->h_rowptr[i + 1]);
    g->h_rowptr[i + 1] += g->h_rowptr[i];
  }
  for (int64_t i = 0; i < n_cols; ++i) {
    g->max_out = std::max(g->max_out, g->h_cscptr[i + 1]);
    g->h_cscptr[i + 1] += g->h_cscptr[i];
  }
  {
    std::vector<int64_t> cur(g->h_cscptr, g->h_cscptr + n_cols);
    for (int64_t k = 0; k < e; ++k) {
      const int64_t pos = cur[scol[k]]++;
      g->h_cscrow[pos] = static_cast<int32_t>(srow[k]);
      g->h_csceid[pos] = static_cast<int32_t>(k);
    }
  }
  std::vector<Item> items;
  std::vector<SplitRow> split;
  g->max_chunk = max_chunk;
  auto keep = [](const std::vector<Item>& it, const std::vector<SplitRow>& sp, Item** hi, SplitRow** hs) {
    *hi = new Item[std::max<size_t>(it.size(), 1)];
    *hs = new SplitRow[std::max<size_t>(sp.size(), 1)];
    if (!it.empty()) memcpy(*hi, it.data(), it.size() * sizeof(Item));
    if (!sp.empty()) memcpy(*hs, sp.data(), sp.size() * sizeof(SplitRow));
  };
  if (device < 0) {
    // structure-only handle (device == -1): host mirrors and schedules, nothing uploaded.  For
    // edis_graph_info / edis_graph_export (host-side checks of the builder); every op rejects it.
    g->dst = build_schedule_host(n, g->h_rowptr, max_chunk, items, split);
    keep(items, split, &g->h_dst_items, &g->h_dst_split);
    g->src = build_schedule_host(n_cols, g->h_cscptr, max_chunk, items, split);
    keep(items, split, &g->h_src_items, &g->h_src_split);
    *out = g;
    return EDIS_OK;
  }
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  int rc = EDIS_OK;
  auto fail = [&](int code) {
    cudaSetDevice(prev_dev);
    edis_graph_destroy(g);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error("edis_graph_create: cudaSetDevice(%d) failed (no CUDA device? there is no CPU fallback)", device);
    return fail(EDIS_ERR_CUDA);
  }
  cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
  g->dst = build_schedule_host(n, g->h_rowptr, max_chunk, items, split);
  keep(items, split, &g->h_dst_items, &g->h_dst_split);
  if ((rc = upload(&g->dst.items, items.data(), items.size())) != EDIS_OK) return fail(rc);
  if ((rc = upload(&g->dst.split, split.data(), split.size())) != EDIS_OK) return fail(rc);
  g->src = build_schedule_host(n_cols, g->h_cscptr, max_chunk, items, split);
  keep(items, split, &g->h_src_items, &g->h_src_split);
  if ((rc = upload(&g->src.items, items.data(), items.size())) != EDIS_OK) return fail(rc);
  if ((rc = upload(&g->src.split, split.data(), split.size())) != EDIS_OK) return fail(rc);
  if ((rc = upload(&g->rowptr, g->h_rowptr, n + 1)) != EDIS_OK) return fail(rc);
  // device copies of the neighbour arrays carry the neighbour's heat level in bits 30-31
  // (see edis_common.cuh); the host mirrors stay plain
  std::vector<int32_t> dcol(g->h_col, g->h_col + e), dcscrow(g->h_cscrow, g->h_cscrow + e);
  if (e > 0) {
    auto heat_of = [](const int64_t* ptr, int64_t count) {
      std::vector<int64_t> deg(count);
      for (int64_t i = 0; i < count; ++i) deg[i] = ptr[i + 1] - ptr[i];
      std::vector<int64_t> srt(deg);
      std::sort(srt.begin(), srt.end(), std::greater<int64_t>());
      auto thr = [&](int64_t k) { return k < count ? std::max<int64_t>(srt[k], 2) : int64_t(2); };
      const int64_t t3 = thr(4096), t2 = thr(16384), t1 = thr(65536);
      std::vector<uint8_t> h(count);
      for (int64_t i = 0; i < count; ++i) h[i] = deg[i] > t3 ? 3 : deg[i] > t2 ? 2 : deg[i] > t1 ? 1 : 0;
      return h;
    };
    const std::vector<uint8_t> hsrc = heat_of(g->h_cscptr, n_cols);   // how often a source row is gathered
    const std::vector<uint8_t> hdst = heat_of(g->h_rowptr, n);        // ... a destination row (src pass)
    for (int64_t k = 0; k < e; ++k) {
      dcol[k] |= static_cast<int32_t>(static_cast<uint32_t>(hsrc[g->h_col[k]]) << kHeatShift);
      dcscrow[k] |= static_cast<int32_t>(static_cast<uint32_t>(hdst[g->h_cscrow[k]]) << kHeatShift);
    }
  }
  if ((rc = upload(&g->col, dcol.data(), e)) != EDIS_OK) return fail(rc);
  if ((rc = upload(&g->cscptr, g->h_cscptr, n_cols + 1)) != EDIS_OK) return fail(rc);
  if ((rc = upload(&g->cscrow, dcscrow.data(), e)) != EDIS_OK) return fail(rc);
  if ((rc = upload(&g->csceid, g->h_csceid, e)) != EDIS_OK) return fail(rc);
  cudaSetDevice(prev_dev);
  *out = g;
  return EDIS_OK;
}

extern "C" void edis_graph_destroy(edis_graph* g) {
  if (!g) return;
  if (g->device >= 0) {
  cudaFree(g->rowptr); cudaFree(g->col); cudaFree(g->cscptr); cudaFree(g->cscrow); cudaFree(g->csceid);
  cudaFree(g->dst.items); cudaFree(g->dst.split); cudaFree(g->src.items); cudaFree(g->src.split);
  }
  delete[] g->h_rowptr; delete[] g->h_col; delete[] g->h_perm;
  delete[] g->h_cscptr; delete[] g->h_cscrow; delete[] g->h_csceid;
  delete[] g->h_dst_items; delete[] g->h_dst_split; delete[] g->h_src_items; delete[] g->h_src_split;
  delete g;
}

extern "C" int edis_graph_info(const edis_graph* g, int64_t info[10]) {
  EDIS_CHECK_ARG(g && info, "edis_graph_info: null argument");
  info[0] = g->n; info[1] = g->e;
  info[2] = g->dst.n_items; info[3] = g->dst.n_slots;
  info[4] = g->src.n_items; info[5] = g->src.n_slots;
  info[6] = g->max_in; info[7] = g->max_out; info[8] = g->was_sorted ? 1 : 0;
  info[9] = g->n_cols;
  return EDIS_OK;
}

extern "C" int64_t edis_graph_input_entries(const edis_graph* g) { return g ? g->e_in : EDIS_ERR_ARG; }

extern "C" int edis_graph_export(const edis_graph* g, int64_t* rowptr, int32_t* col, int64_t* perm,
                                 int64_t* cscptr, int32_t* cscrow, int32_t* csceid) {
  EDIS_CHECK_ARG(g, "edis_graph_export: null graph");
  if (rowptr) memcpy(rowptr, g->h_rowptr, (g->n + 1) * sizeof(int64_t));
  if (col) memcpy(col, g->h_col, g->e * sizeof(int32_t));
  if (perm) memcpy(perm, g->h_perm, g->e_in * sizeof(int64_t));  // sized by the INPUT entry count
  if (cscptr) memcpy(cscptr, g->h_cscptr, (g->n_cols + 1) * sizeof(int64_t));
  if (cscrow) memcpy(cscrow, g->h_cscrow, g->e * sizeof(int32_t));
  if (csceid) memcpy(csceid, g->h_csceid, g->e * sizeof(int32_t));
  return EDIS_OK;
}

// ---------------------------------------------------------------------------------------
// On-disk graph cache: everything edis_graph_create derives from an edge list, in one
// flat file that edis_graph_load memory-maps and uploads.  Replaces the reference's dense detour on
// every start (data_load.py:39-77) and, being keyed by the caller's content hash, its stale
// `./resource/<ds>/DisEdges.pt` hazard (pretrainer.py:390-398): a file written for other input, another
// chunk size or another format version is rejected, never silently used.
namespace {
constexpr uint64_t kCacheMagic = 0x3147534944450a0dull;   // "\r\nEDISG1"
constexpr uint32_t kCacheVersion = 2;
struct CacheHeader {
  uint64_t magic;
  uint32_t version, max_chunk;
  uint64_t key;
  int64_t n, n_cols, e, e_in, max_in, max_out, was_sorted;
  int64_t dst_items, dst_slots, dst_split, src_items, src_slots, src_split;
  uint64_t payload_bytes, payload_hash;
};
inline uint64_t mix64(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  return h ^ (h >> 32);
}
uint64_t hash_bytes(const void* p, size_t bytes, uint64_t seed) {
  // 4 independent lanes over 8-byte words (memory-bound on one core: ~5 GB/s)
  const uint64_t* w = static_cast<const uint64_t*>(p);
  const size_t nw = bytes / 8;
  uint64_t h[4] = {seed, seed ^ 0xa5a5a5a5a5a5a5a5ull, seed + 0x1234567ull, ~seed};
  size_t i = 0;
  for (; i + 4 <= nw; i += 4)
    for (int k = 0; k < 4; ++k) h[k] = mix64(h[k], w[i + k]);
  for (; i < nw; ++i) h[0] = mix64(h[0], w[i]);
  uint64_t tail = 0;
  memcpy(&tail, static_cast<const uint8_t*>(p) + nw * 8, bytes - nw * 8);
  uint64_t r = mix64(h[0], tail);
  for (int k = 1; k < 4; ++k) r = mix64(r, h[k]);
  return mix64(r, bytes);
}
struct Section {
  const void* p;
  size_t bytes;
};
size_t pad8(size_t b) { return (b + 7) & ~size_t(7); }
}  // namespace

// Content key of an edge list: what `edis_graph_save` / `edis_graph_load` files are keyed by.
extern "C" uint64_t edis_edge_list_key(int64_t n, int64_t n_cols, int64_t e_in, const int64_t* row,
                                       const int64_t* col, int max_chunk) {
  uint64_t h = mix64(mix64(mix64(0x45444953ull, n), n_cols), e_in);
  h = mix64(h, max_chunk <= 0 ? 256 : max_chunk);
  if (e_in > 0 && row && col) {
    h = hash_bytes(row, static_cast<size_t>(e_in) * 8, h);
    h = hash_bytes(col, static_cast<size_t>(e_in) * 8, h);
  }
  return h;
}

extern "C" int edis_graph_save(const edis_graph* g, const char* path, uint64_t key) {
  EDIS_CHECK_ARG(g && path, "edis_graph_save: null argument");
  EDIS_CHECK_ARG(g->h_dst_items && g->h_src_items, "edis_graph_save: handle carries no host schedules");
  // neighbour arrays are stored WITH their heat bits when the handle is device-resident; a
  // structure-only handle stores plain ids (heat 0) and the loader recomputes nothing
  std::vector<int32_t> dcol(std::max<int64_t>(g->e, 1)), dcscrow(std::max<int64_t>(g->e, 1));
  if (g->device >= 0 && g->e > 0) {
    if (cudaMemcpy(dcol.data(), g->col, g->e * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(dcscrow.data(), g->cscrow, g->e * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
      set_error("edis_graph_save: reading the device arrays failed");
      return EDIS_ERR_CUDA;
    }
  } else if (g->e > 0) {
    memcpy(dcol.data(), g->h_col, g->e * 4);
    memcpy(dcscrow.data(), g->h_cscrow, g->e * 4);
  }
  const Section sec[] = {
      {g->h_rowptr, static_cast<size_t>(g->n + 1) * 8},        {dcol.data(), static_cast<size_t>(g->e) * 4},
      {g->h_perm, static_cast<size_t>(g->e_in) * 8},            {g->h_cscptr, static_cast<size_t>(g->n_cols + 1) * 8},
      {dcscrow.data(), static_cast<size_t>(g->e) * 4},          {g->h_csceid, static_cast<size_t>(g->e) * 4},
      {g->h_dst_items, static_cast<size_t>(g->dst.n_items) * sizeof(Item)},
      {g->h_dst_split, static_cast<size_t>(g->dst.n_split) * sizeof(SplitRow)},
      {g->h_src_items, static_cast<size_t>(g->src.n_items) * sizeof(Item)},
      {g->h_src_split, static_cast<size_t>(g->src.n_split) * sizeof(SplitRow)}};
  CacheHeader h = {};
  h.magic = kCacheMagic; h.version = kCacheVersion; h.max_chunk = static_cast<uint32_t>(g->max_chunk);
  h.key = key;
  h.n = g->n; h.n_cols = g->n_cols; h.e = g->e; h.e_in = g->e_in;
  h.max_in = g->max_in; h.max_out = g->max_out; h.was_sorted = g->was_sorted ? 1 : 0;
  h.dst_items = g->dst.n_items; h.dst_slots = g->dst.n_slots; h.dst_split = g->dst.n_split;
  h.src_items = g->src.n_items; h.src_slots = g->src.n_slots; h.src_split = g->src.n_split;
  uint64_t ph = 0x5eedull;
  for (const Section& s : sec) {
    h.payload_bytes += pad8(s.bytes);
    ph = mix64(ph, hash_bytes(s.p, s.bytes, s.bytes));
  }
  h.payload_hash = ph;
  const std::string tmp = std::string(path) + ".tmp." + std::to_string(static_cast<long long>(getpid()));
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) {
    set_error("edis_graph_save: cannot open %s for writing", tmp.c_str());
    return EDIS_ERR_ARG;
  }
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  const uint64_t zero = 0;
  for (const Section& s : sec) {
    if (s.bytes) ok = ok && fwrite(s.p, 1, s.bytes, f) == s.bytes;
    if (pad8(s.bytes) != s.bytes) ok = ok && fwrite(&zero, 1, pad8(s.bytes) - s.bytes, f) == pad8(s.bytes) - s.bytes;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {       // atomic publish: readers never see a partial file
    remove(tmp.c_str());
    set_error("edis_graph_save: writing %s failed", path);
    return EDIS_ERR_ARG;
  }
  return EDIS_OK;
}

// Returns EDIS_OK and a handle, or EDIS_ERR_STALE (no handle) when the file is missing, truncated,
// corrupt, of another format version, or was written for another key / chunk size.
extern "C" int edis_graph_load(const char* path, uint64_t key, int max_chunk, int device, int verify,
                               edis_graph** out) {
  EDIS_CHECK_ARG(path && out, "edis_graph_load: null argument");
  *out = nullptr;
  if (max_chunk <= 0) max_chunk = 256;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    set_error("edis_graph_load: %s: no such cache file", path);
    return EDIS_ERR_STALE;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || static_cast<size_t>(st.st_size) < sizeof(CacheHeader)) {
    close(fd);
    set_error("edis_graph_load: %s: truncated", path);
    return EDIS_ERR_STALE;
  }
  const size_t fbytes = static_cast<size_t>(st.st_size);
  void* map = mmap(nullptr, fbytes, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) {
    set_error("edis_graph_load: mmap of %s failed", path);
    return EDIS_ERR_STALE;
  }
  madvise(map, fbytes, MADV_SEQUENTIAL);
  const CacheHeader h = *static_cast<const CacheHeader*>(map);
  auto stale = [&](const char* why) {
    munmap(map, fbytes);
    set_error("edis_graph_load: %s: %s", path, why);
    return EDIS_ERR_STALE;
  };
  if (h.magic != kCacheMagic || h.version != kCacheVersion) return stale("not an edis graph cache of this version");
  if (h.key != key) return stale("stale: written for a different input (content key mismatch)");
  if (static_cast<int>(h.max_chunk) != max_chunk) return stale("stale: written for a different max_chunk");
  if (h.n <= 0 || h.n_cols < h.n || h.e < 0 || h.e_in < 0) return stale("corrupt header");
  const size_t sizes[] = {static_cast<size_t>(h.n + 1) * 8, static_cast<size_t>(h.e) * 4,
                          static_cast<size_t>(h.e_in) * 8,   static_cast<size_t>(h.n_cols + 1) * 8,
                          static_cast<size_t>(h.e) * 4,      static_cast<size_t>(h.e) * 4,
                          static_cast<size_t>(h.dst_items) * sizeof(Item), static_cast<size_t>(h.dst_split) * sizeof(SplitRow),
                          static_cast<size_t>(h.src_items) * sizeof(Item), static_cast<size_t>(h.src_split) * sizeof(SplitRow)};
  size_t total = 0;
  for (size_t b : sizes) total += pad8(b);
  if (total != h.payload_bytes || sizeof(CacheHeader) + total != fbytes) return stale("truncated or corrupt (size mismatch)");
  const uint8_t* base = static_cast<const uint8_t*>(map) + sizeof(CacheHeader);
  const uint8_t* secp[10];
  {
    const uint8_t* q = base;
    for (int k = 0; k < 10; ++k) {
      secp[k] = q;
      q += pad8(sizes[k]);
    }
  }
  if (verify) {
    uint64_t ph = 0x5eedull;
    for (int k = 0; k < 10; ++k) ph = mix64(ph, hash_bytes(secp[k], sizes[k], sizes[k]));
    if (ph != h.payload_hash) return stale("corrupt (payload checksum mismatch)");
  }
  edis_graph* g = new edis_graph();
  g->n = h.n; g->n_cols = h.n_cols; g->e = h.e; g->e_in = h.e_in;
  g->max_in = h.max_in; g->max_out = h.max_out; g->was_sorted = h.was_sorted != 0;
  g->device = device;
  g->max_chunk = max_chunk;
  g->dst.n_items = h.dst_items; g->dst.n_slots = h.dst_slots; g->dst.n_split = h.dst_split;
  g->src.n_items = h.src_items; g->src.n_slots = h.src_slots; g->src.n_split = h.src_split;
  const int64_t e1 = std::max<int64_t>(h.e, 1);
  g->h_rowptr = new int64_t[h.n + 1];
  g->h_col = new int32_t[e1];
  g->h_perm = new int64_t[std::max<int64_t>(h.e_in, 1)];
  g->h_cscptr = new int64_t[h.n_cols + 1];
  g->h_cscrow = new int32_t[e1];
  g->h_csceid = new int32_t[e1];
  g->h_dst_items = new Item[std::max<int64_t>(h.dst_items, 1)];
  g->h_dst_split = new SplitRow[std::max<int64_t>(h.dst_split, 1)];
  g->h_src_items = new Item[std::max<int64_t>(h.src_items, 1)];
  g->h_src_split = new SplitRow[std::max<int64_t>(h.src_split, 1)];
  memcpy(g->h_rowptr, secp[0], sizes[0]);
  memcpy(g->h_perm, secp[2], sizes[2]);
  memcpy(g->h_cscptr, secp[3], sizes[3]);
  memcpy(g->h_csceid, secp[5], sizes[5]);
  memcpy(g->h_dst_items, secp[6], sizes[6]);
  memcpy(g->h_dst_split, secp[7], sizes[7]);
  memcpy(g->h_src_items, secp[8], sizes[8]);
  memcpy(g->h_src_split, secp[9], sizes[9]);
  const int32_t* dcol = reinterpret_cast<const int32_t*>(secp[1]);
  const int32_t* dcscrow = reinterpret_cast<const int32_t*>(secp[4]);
  for (int64_t k = 0; k < h.e; ++k) {
    g->h_col[k] = dcol[k] & kIdMask;
    g->h_cscrow[k] = dcscrow[k] & kIdMask;
  }
  int rc = EDIS_OK;
  if (device >= 0) {
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    if (cudaSetDevice(device) != cudaSuccess) {
      set_error("edis_graph_load: cudaSetDevice(%d) failed (no CUDA device? there is no CPU fallback)", device);
      rc = EDIS_ERR_CUDA;
    } else {
      cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
      // straight from the mapped pages: no intermediate host buffer
      if (rc == EDIS_OK) rc = upload(&g->rowptr, reinterpret_cast<const int64_t*>(secp[0]), h.n + 1);
      if (rc == EDIS_OK) rc = upload(&g->col, dcol, h.e);
      if (rc == EDIS_OK) rc = upload(&g->cscptr, reinterpret_cast<const int64_t*>(secp[3]), h.n_cols + 1);
      if (rc == EDIS_OK) rc = upload(&g->cscrow, dcscrow, h.e);
      if (rc == EDIS_OK) rc = upload(&g->csceid, reinterpret_cast<const int32_t*>(secp[5]), h.e);
      if (rc == EDIS_OK) rc = upload(&g->dst.items, reinterpret_cast<const Item*>(secp[6]), h.dst_items);
      if (rc == EDIS_OK) rc = upload(&g->dst.split, reinterpret_cast<const SplitRow*>(secp[7]), h.dst_split);
      if (rc == EDIS_OK) rc = upload(&g->src.items, reinterpret_cast<const Item*>(secp[8]), h.src_items);
      if (rc == EDIS_OK) rc = upload(&g->src.split, reinterpret_cast<const SplitRow*>(secp[9]), h.src_split);
      cudaSetDevice(prev_dev);
    }
  }
  munmap(map, fbytes);
  if (rc != EDIS_OK) {
    edis_graph_destroy(g);
    return rc;
  }
  *out = g;
  return EDIS_OK;
}

extern "C" int64_t edis_graph_workspace_bytes(const edis_graph* g, int64_t width) {
  if (!g || width <= 0) return EDIS_ERR_ARG;
  const int64_t slots = std::max(g->dst.n_slots, g->src.n_slots);
  return (slots * width + 64) * static_cast<int64_t>(sizeof(float));
}

// ---------------------------------------------------------------------------------------
extern "C" int64_t edis_merge_pairs_host(int64_t n_hit, const int64_t* hit_key, int64_t n_forced,
                                         const int64_t* forced_key, int64_t n_pos,
                                         const int64_t* pos_key, int64_t* out_key, float* out_label) {
  if (n_hit < 0 || n_forced < 0 || n_pos < 0 || !out_key || !out_label) {
    set_error("edis_merge_pairs_host: bad arguments");
    return EDIS_ERR_ARG;
  }
  std::vector<int64_t> f(forced_key, forced_key + n_forced);
  std::sort(f.begin(), f.end());
  const bool hit_sorted = std::is_sorted(hit_key, hit_key + n_hit);
  std::vector<int64_t> hs;
  const int64_t* h = hit_key;
  if (!hit_sorted) {
    hs.assign(hit_key, hit_key + n_hit);
    std::sort(hs.begin(), hs.end());
    h = hs.data();
  }
  // merge + dedup (mask.nonzero() is row-major sorted and unique, pretrainer.py:703)
  int64_t a = 0, b = 0, m = 0, p = 0;
  while (a < n_hit || b < n_forced) {
    int64_t key;
    if (b >= n_forced || (a < n_hit && h[a] <= f[b])) key = h[a++]; else key = f[b++];
    if (m > 0 && out_key[m - 1] == key) continue;
    while (p < n_pos && pos_key[p] < key) ++p;
    out_label[m] = (p < n_pos && pos_key[p] == key) ? 1.0f : 0.0f;  // label[indices] (704)
    out_key[m++] = key;
  }
  return m;
}

// ---------------------------------------------------------------------------------------
// Bit-exact replay of `(torch.rand(n_draws) < thr).nonzero()` on torch's CPU generator without
// materialising the uniforms (pretrainer.py:692 draws N x N of them per sample_train call).
//   state   the 5056-byte blob of torch.get_rng_state(): CPUGeneratorImplState =
//           { u64 seed; i32 left; i32 seeded; u64 next; u64 mt[624]; normal-cache fields ... }
//           (aten/src/ATen/CPUGeneratorImpl.cpp); advanced IN PLACE by n_draws 32-bit outputs, so
//           torch.set_rng_state(state) afterwards leaves the generator exactly where the
//           reference's torch.rand would have left it.
//   thr24   the float32 threshold as an integer: a float32 uniform is (y & 0xFFFFFF) * 2^-24
//           (ATen/core/TransformationHelper.h uniform_real<float>), so  u < thr  <=>
//           (y & 0xFFFFFF) < ceil(float32(thr) * 2^24)
//   out     linear draw indices of the hits (row-major: k = i * N + j), ascending; capacity `cap`
// Returns the number of hits, or -(hits needed) if cap is too small (state untouched then).
// The engine is ATen's mt19937 (ATen/core/MT19937RNGEngine.h): standard MT19937, `left` counts
// the outputs remaining in the current 624-word block, regenerated when --left == 0.
namespace {
inline uint32_t temper24(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y & 0xFFFFFFu;
}
// hits among the raw MT words w[0, cnt): draw index base + t for every t with temper24(w[t]) < thr24
void scan_words(const uint32_t* w, int64_t cnt, int64_t base, uint32_t thr24, std::vector<int64_t>* hits) {
  constexpr int B = 1024;
  uint32_t tmp[B];
  for (int64_t o = 0; o < cnt; o += B) {
    const int len = static_cast<int>(std::min<int64_t>(B, cnt - o));
    for (int t = 0; t < len; ++t) tmp[t] = temper24(w[o + t]);      // vectorises
    for (int t = 0; t < len; ++t)
      if (tmp[t] < thr24) hits->push_back(base + o + t);            // rare
  }
}
}  // namespace

extern "C" int64_t edis_rand_hits_host(uint8_t* state, int64_t state_bytes, int64_t n_draws,
                                       uint32_t thr24, int64_t* out, int64_t cap) {
  if (!state || state_bytes < 24 + 624 * 8 || n_draws < 0 || (!out && cap > 0)) {
    set_error("edis_rand_hits_host: bad arguments (state must be torch's 5056-byte CPU generator state)");
    return EDIS_ERR_ARG;
  }
  constexpr int N = 624, M = 397;
  int32_t left;
  uint64_t next64;
  uint32_t mt[N];
  std::memcpy(&left, state + 8, 4);
  std::memcpy(&next64, state + 16, 8);
  for (int i = 0; i < N; ++i) {
    uint64_t v;
    std::memcpy(&v, state + 24 + 8 * i, 8);
    mt[i] = static_cast<uint32_t>(v);
  }
  uint32_t next = static_cast<uint32_t>(next64);
  if (left < 1 || left > N || next > static_cast<uint32_t>(N)) {
    set_error("edis_rand_hits_host: implausible generator state (left=%d next=%u)", left, next);
    return EDIS_ERR_ARG;
  }
  auto twist = [](uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((0u - (v & 1u)) & 0x9908b0dfu);
  };
  // next block `nw` from the previous one `old` (MT19937RNGEngine.h next_state(), written out of
  // place: no load/store hazards, so the three loops vectorise)
  auto next_block = [&](const uint32_t* __restrict__ old, uint32_t* __restrict__ nw) {
    for (int i = 0; i < N - M; ++i) nw[i] = old[i + M] ^ twist(old[i], old[i + 1]);
    // nw[i] needs nw[i - 227]: independent within runs of 227, done as two such runs
    for (int i = N - M; i < 2 * (N - M); ++i) nw[i] = nw[i + M - N] ^ twist(old[i], old[i + 1]);
    for (int i = 2 * (N - M); i < N - 1; ++i) nw[i] = nw[i + M - N] ^ twist(old[i], old[i + 1]);
    nw[N - 1] = nw[M - 1] ^ twist(old[N - 1], nw[0]);
  };
  // ATen's engine: a draw does `if (--left == 0) regenerate (left = 624, next = 0)` and then takes
  // mt[next++].  So left - 1 words of the current block are still unread, starting at mt[next].
  std::vector<int64_t> hits;
  int64_t k = 0;
  {
    const int64_t a0 = std::min<int64_t>(n_draws, left - 1);
    scan_words(mt + next, a0, 0, thr24, &hits);
    next += static_cast<uint32_t>(a0);
    left -= static_cast<int32_t>(a0);
    k = a0;
  }
  // Whole blocks: the recurrence is sequential (this thread), tempering + threshold test is not:
  // blocks are generated into one half of a double buffer while helper threads scan the other half.
  constexpr int64_t kChunkBlocks = 2048;                           // 2048 * 624 words = 5 MB per half
  const int helpers = static_cast<int>(std::max(1u, std::min(4u, std::thread::hardware_concurrency() > 1
                                                                      ? std::thread::hardware_concurrency() - 1 : 1u)));
  std::vector<uint32_t> buf[2];
  std::vector<std::thread> workers;
  std::vector<std::vector<int64_t>> part(helpers);
  auto drain = [&]() {
    for (auto& t : workers) t.join();
    workers.clear();
    for (auto& p : part) {
      hits.insert(hits.end(), p.begin(), p.end());
      p.clear();
    }
  };
  int half = 0;
  while (k < n_draws) {
    const int64_t want = std::min<int64_t>(n_draws - k, kChunkBlocks * N);
    const int64_t blocks = (want + N - 1) / N;
    std::vector<uint32_t>& b = buf[half];
    b.resize(static_cast<size_t>(blocks) * N);
    for (int64_t q = 0; q < blocks; ++q) next_block(q ? b.data() + (q - 1) * N : mt, b.data() + q * N);
    std::memcpy(mt, b.data() + (blocks - 1) * N, sizeof(mt));
    const int64_t last = want - (blocks - 1) * N;                  // words consumed from the last block
    left = N + 1 - static_cast<int32_t>(last);
    next = static_cast<uint32_t>(last);
    drain();                                                       // the other half is free again
    const int64_t per = (want + helpers - 1) / helpers;
    for (int h = 0; h < helpers; ++h) {
      const int64_t lo = std::min<int64_t>(want, h * per), hi = std::min<int64_t>(want, lo + per);
      if (hi > lo) workers.emplace_back(scan_words, b.data() + lo, hi - lo, k + lo, thr24, &part[h]);
    }
    k += want;
    half ^= 1;
  }
  drain();
  const int64_t nh = static_cast<int64_t>(hits.size());
  if (nh > cap) return -nh;
  if (nh) std::memcpy(out, hits.data(), static_cast<size_t>(nh) * sizeof(int64_t));
  std::memcpy(state + 8, &left, 4);
  next64 = next;
  std::memcpy(state + 16, &next64, 8);
  for (int i = 0; i < N; ++i) {
    const uint64_t v = mt[i];
    std::memcpy(state + 24 + 8 * i, &v, 8);
  }
  return nh;
}

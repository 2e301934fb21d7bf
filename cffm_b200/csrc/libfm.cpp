// Native libfm loader: restates LoadData (reference LoadData.py:25-112).
//   - vocabulary: the whole "idx:val" token string is the key; ids are handed out in order of
//     first appearance while scanning train, test, validation (LoadData.py:33-55);
//   - rows: label = float(items[0]) (and 1/0 by >0 for log_loss), features = ids of items[1:]
//     (LoadData.py:81-103); a line is split exactly like Python's line.strip().split(' ');
//   - rows are re-ordered by ascending row length (LoadData.py:105-112), stably;
//   - output is CSR (row_ptr, ids) + labels in page-locked host memory, ready for
//     cudaMemcpyAsync; when no CUDA device is present the buffers are plain host memory.
#include <cuda_runtime_api.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/cffm.h"

namespace {

thread_local std::string g_io_err;

struct Mapped {
  const char* p = nullptr; size_t n = 0; int fd = -1;
  bool open(const char* path, std::string* err) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { *err = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { *err = std::string("fstat ") + path; return false; }
    n = (size_t)st.st_size;
    if (n == 0) { p = ""; return true; }
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { *err = std::string("mmap ") + path + ": " + strerror(errno); return false; }
    madvise(m, n, MADV_SEQUENTIAL);
    p = (const char*)m;
    return true;
  }
  ~Mapped() { if (p && n) munmap((void*)p, n); if (fd >= 0) ::close(fd); }
};

inline bool py_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// Open-addressing hash map from token bytes to id; tokens live in one arena.
struct Vocab {
  std::vector<char> arena;
  std::vector<uint64_t> tok_off;   // start of token id in arena
  std::vector<uint32_t> tok_len;
  std::vector<int32_t> slots;      // -1 empty, else id
  std::vector<uint64_t> hashes;    // hash per id
  size_t mask = 0;
  Vocab() { slots.assign(1 << 16, -1); mask = slots.size() - 1; }
  static uint64_t hash(const char* s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32;
    return h * 0x9E3779B97F4A7C15ull;
  }
  void grow() {
    std::vector<int32_t> ns(slots.size() * 2, -1);
    const size_t nm = ns.size() - 1;
    for (size_t id = 0; id < hashes.size(); ++id) {
      size_t s = (hashes[id] >> 7) & nm;
      while (ns[s] >= 0) s = (s + 1) & nm;
      ns[s] = (int32_t)id;
    }
    slots.swap(ns); mask = nm;
  }
  // returns id or -1
  int32_t find(const char* s, size_t n, uint64_t h) const {
    size_t sl = (h >> 7) & mask;
    while (true) {
      const int32_t id = slots[sl];
      if (id < 0) return -1;
      if (hashes[id] == h && tok_len[id] == n && memcmp(arena.data() + tok_off[id], s, n) == 0) return id;
      sl = (sl + 1) & mask;
    }
  }
  int32_t find_or_add(const char* s, size_t n) {
    const uint64_t h = hash(s, n);
    int32_t id = find(s, n, h);
    if (id >= 0) return id;
    if ((hashes.size() + 1) * 2 > slots.size()) grow();
    id = (int32_t)hashes.size();
    tok_off.push_back(arena.size()); tok_len.push_back((uint32_t)n); hashes.push_back(h);
    arena.insert(arena.end(), s, s + n);
    size_t sl = (h >> 7) & mask;
    while (slots[sl] >= 0) sl = (sl + 1) & mask;
    slots[sl] = id;
    return id;
  }
};

// Calls fn(token_begin, token_len, index_in_line) for every item of line.strip().split(' ').
template <class Fn>
inline void split_line(const char* b, const char* e, Fn&& fn) {
  while (b < e && py_space(*b)) ++b;
  while (e > b && py_space(e[-1])) --e;
  int idx = 0;
  const char* t = b;
  for (const char* c = b;; ++c) {
    if (c == e || *c == ' ') {
      fn(t, (size_t)(c - t), idx++);
      if (c == e) break;
      t = c + 1;
    }
  }
}

template <class Fn>
inline void for_each_line(const Mapped& f, Fn&& fn) {
  const char* p = f.p; const char* end = f.p + f.n;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    fn(p, le);
    p = nl ? nl + 1 : end;
  }
}

struct Split {
  int64_t n_rows = 0, nnz = 0;
  int64_t* row_ptr = nullptr; int32_t* ids = nullptr; float* y_raw = nullptr; float* y_log = nullptr;
};

void* host_alloc(size_t bytes, bool* pinned) {
  if (bytes == 0) bytes = 8;
  void* p = nullptr;
  if (*pinned) {
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) return p;
    cudaGetLastError();
    *pinned = false;  // no device / no driver: plain host memory
  }
  if (posix_memalign(&p, 256, bytes) != 0) return nullptr;
  return p;
}

}  // namespace

struct cffm_libfm {
  Vocab vocab;
  Split split[3];
  bool pinned = true;
  std::vector<void*> blocks;
  std::vector<bool> block_pinned;
  void* alloc(size_t bytes) {
    bool pin = pinned;
    void* p = host_alloc(bytes, &pin);
    pinned = pin;
    blocks.push_back(p); block_pinned.push_back(pin);
    return p;
  }
};

static bool scan_vocab(const char* path, Vocab* v, std::string* err) {
  Mapped f;
  if (!f.open(path, err)) return false;
  for_each_line(f, [&](const char* b, const char* e) {
    split_line(b, e, [&](const char* t, size_t n, int idx) { if (idx > 0) v->find_or_add(t, n); });
  });
  return true;
}

static bool read_split(const char* path, cffm_libfm* d, Split* out, std::string* err) {
  Mapped f;
  if (!f.open(path, err)) return false;
  std::vector<int64_t> rp; rp.push_back(0);
  std::vector<int32_t> ids;
  std::vector<double> y;
  bool ok = true;
  int64_t lineno = 0;
  for_each_line(f, [&](const char* b, const char* e) {
    ++lineno;
    if (!ok) return;
    split_line(b, e, [&](const char* t, size_t n, int idx) {
      if (!ok) return;
      if (idx == 0) {
        char buf[64];
        if (n == 0 || n >= sizeof(buf)) { ok = false; *err = std::string(path) + ":" + std::to_string(lineno) + ": cannot parse label"; return; }
        memcpy(buf, t, n); buf[n] = 0;
        char* endp = nullptr;
        const double v = strtod(buf, &endp);
        if (endp != buf + n) { ok = false; *err = std::string(path) + ":" + std::to_string(lineno) + ": cannot parse label '" + buf + "'"; return; }
        y.push_back(v);
      } else {
        const int32_t id = d->vocab.find(t, n, Vocab::hash(t, n));
        if (id < 0) { ok = false; *err = std::string(path) + ":" + std::to_string(lineno) + ": token not in vocabulary"; return; }
        ids.push_back(id);
      }
    });
    rp.push_back((int64_t)ids.size());
  });
  if (!ok) return false;
  const int64_t n = (int64_t)y.size();
  // LoadData.py:105-112: rows ordered by row length (stable here; see oracle/libfm_ref.py)
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  bool ragged = false;
  for (int64_t i = 1; i < n && !ragged; ++i) ragged = (rp[i + 1] - rp[i]) != (rp[1] - rp[0]);
  if (ragged) std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return rp[a + 1] - rp[a] < rp[b + 1] - rp[b]; });
  out->n_rows = n; out->nnz = (int64_t)ids.size();
  out->row_ptr = (int64_t*)d->alloc(sizeof(int64_t) * (n + 1));
  out->ids = (int32_t*)d->alloc(sizeof(int32_t) * ids.size());
  out->y_raw = (float*)d->alloc(sizeof(float) * n);
  out->y_log = (float*)d->alloc(sizeof(float) * n);
  if (!out->row_ptr || !out->ids || !out->y_raw || !out->y_log) { *err = "out of host memory"; return false; }
  int64_t w = 0;
  out->row_ptr[0] = 0;
  for (int64_t r = 0; r < n; ++r) {
    const int64_t src = order[r];
    const int64_t len = rp[src + 1] - rp[src];
    if (len) memcpy(out->ids + w, ids.data() + rp[src], sizeof(int32_t) * len);
    w += len;
    out->row_ptr[r + 1] = w;
    out->y_raw[r] = (float)y[src];
    out->y_log[r] = y[src] > 0 ? 1.f : 0.f;
  }
  return true;
}

extern "C" int cffm_libfm_load(const char* train_path, const char* test_path, const char* validation_path, cffm_libfm** out) {
  if (!train_path || !test_path || !validation_path || !out) return CFFM_ERR_INVALID;
  *out = nullptr;
  cffm_libfm* d = nullptr;
  try {
    d = new cffm_libfm();
    std::string err;
    // vocabulary pass order: train, test, validation (LoadData.py:35-39)
    if (!scan_vocab(train_path, &d->vocab, &err) || !scan_vocab(test_path, &d->vocab, &err) ||
        !scan_vocab(validation_path, &d->vocab, &err) ||
        // data pass order: train, validation, test (LoadData.py:58-72); split index 0,1,2
        !read_split(train_path, d, &d->split[0], &err) || !read_split(validation_path, d, &d->split[1], &err) ||
        !read_split(test_path, d, &d->split[2], &err)) {
      g_io_err = err;
      cffm_libfm_free(d);
      return CFFM_ERR_IO;
    }
  } catch (...) {
    g_io_err = "out of host memory";
    if (d) cffm_libfm_free(d);
    return CFFM_ERR_NOMEM;
  }
  *out = d;
  return CFFM_OK;
}

extern "C" const char* cffm_libfm_last_error(void) { return g_io_err.c_str(); }

extern "C" int64_t cffm_libfm_features_M(const cffm_libfm* d) { return d ? (int64_t)d->vocab.hashes.size() : -1; }

extern "C" int cffm_libfm_split(const cffm_libfm* d, int split, int64_t* n_rows, const int64_t** row_ptr, const int32_t** ids,
                                const float** labels_raw, const float** labels_log) {
  if (!d || split < 0 || split > 2) return CFFM_ERR_INVALID;
  const Split& s = d->split[split];
  if (n_rows) *n_rows = s.n_rows;
  if (row_ptr) *row_ptr = s.row_ptr;
  if (ids) *ids = s.ids;
  if (labels_raw) *labels_raw = s.y_raw;
  if (labels_log) *labels_log = s.y_log;
  return CFFM_OK;
}

extern "C" int cffm_libfm_token(const cffm_libfm* d, int64_t id, char* buf, int cap) {
  if (!d || id < 0 || id >= (int64_t)d->vocab.hashes.size() || !buf || cap < 1) return CFFM_ERR_INVALID;
  const int n = (int)d->vocab.tok_len[id];
  const int k = n < cap - 1 ? n : cap - 1;
  memcpy(buf, d->vocab.arena.data() + d->vocab.tok_off[id], k);
  buf[k] = 0;
  return n;
}

extern "C" int cffm_libfm_free(cffm_libfm* d) {
  if (!d) return CFFM_OK;
  for (size_t i = 0; i < d->blocks.size(); ++i) {
    if (!d->blocks[i]) continue;
    if (d->block_pinned[i]) cudaFreeHost(d->blocks[i]); else free(d->blocks[i]);
  }
  delete d;
  return CFFM_OK;
}

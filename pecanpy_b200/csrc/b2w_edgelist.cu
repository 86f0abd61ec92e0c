// b2w_edgelist.cu -- host-side parser of `.edg` edge lists (no device code; lives in libb2w.so for the C ABI).
//
// Replaces the per-line Python of AdjlstGraph.read / _read_edge_line / add_edge / add_node
// (reference graph.py:160-180, 217-236, 258-305): ~9 us per line even when vectorised with NumPy object arrays,
// i.e. minutes at 10^7 edges, after which the CSR build on the device (b2w_csr_build.cu) takes 86 ms.  Same
// conventions, line by line:
//   terms = line.strip().split(delimiter); id1, id2 = terms[0].strip(), terms[1].strip()      (graph.py:166-167)
//   weighted: exactly three terms, weight = float(terms[-1]); unweighted: weight 1, extra columns ignored
//   weight <= 0: the line is ignored and its ids are NOT registered                            (graph.py:182-193, 283)
//   node index = order of first appearance, id1 before id2                                     (graph.py:217-236)
// Output: endpoints and weights in FILE ORDER (duplicates are resolved later: "a later line wins" is a property of
// the stable sort in b2w_csr_from_edges / of the host lexsort), plus the id strings as one NUL-separated blob.
// Whitespace is the ASCII set of Python's str.strip(); files with non-ASCII bytes are refused
// (B2W_ERR_UNSUPPORTED) so that the caller can fall back to Python's Unicode-aware strip.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "b2w_common.cuh"

namespace {

inline bool is_py_space(unsigned char c) {
  return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f);
}

struct IdTable {
  // open addressing over (offset, length) views into `blob`
  std::vector<uint32_t> slots;     // node index + 1, 0 = empty
  std::vector<uint64_t> off;       // per node: offset into blob
  std::vector<uint32_t> len;       // per node: byte length
  std::string blob;                // ids, NUL separated, in node order
  uint64_t mask = 0;

  static uint64_t hash(const char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;                               // FNV-1a, finalised with a multiply-xorshift
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32; h *= 0x9e3779b97f4a7c15ull; h ^= h >> 29;
    return h;
  }
  void init(size_t cap_pow2) { slots.assign(cap_pow2, 0u); mask = cap_pow2 - 1; }
  void grow() {
    std::vector<uint32_t> fresh(slots.size() * 2, 0u);
    const uint64_t m2 = fresh.size() - 1;
    for (uint32_t node = 0; node < off.size(); ++node) {
      uint64_t h = hash(blob.data() + off[node], len[node]) & m2;
      while (fresh[h]) h = (h + 1) & m2;
      fresh[h] = node + 1;
    }
    slots.swap(fresh);
    mask = m2;
  }
  uint32_t intern(const char* p, size_t n) {
    uint64_t h = hash(p, n) & mask;
    while (slots[h]) {
      const uint32_t node = slots[h] - 1;
      if (len[node] == n && memcmp(blob.data() + off[node], p, n) == 0) return node;
      h = (h + 1) & mask;
    }
    const uint32_t node = (uint32_t)off.size();
    slots[h] = node + 1;
    off.push_back(blob.size());
    len.push_back((uint32_t)n);
    blob.append(p, n);
    blob.push_back('\0');
    if ((uint64_t)off.size() * 2 > slots.size()) grow();
    return node;
  }
};

}  // namespace

struct b2w_edgelist {
  std::vector<uint32_t> src, dst;
  std::vector<double> w;
  IdTable ids;
  uint64_t dropped = 0;            // lines ignored for weight <= 0
  std::vector<uint64_t> dropped_lines;   // first few of them (1-based line numbers), for the caller's warnings
};

extern "C" int b2w_edgelist_parse(const char* path, int weighted, const char* delimiter, b2w_edgelist** out,
                                  uint64_t* num_edges, uint32_t* num_nodes, uint64_t* names_bytes, uint64_t* num_dropped) {
  if (!path || !delimiter || !out || !delimiter[0]) { b2w_set_error("b2w_edgelist_parse: null / empty argument"); return B2W_ERR_INVALID; }
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { b2w_set_error("b2w_edgelist_parse: cannot open %s: %s", path, strerror(errno)); return B2W_ERR_INVALID; }
  std::string buf;
  try {
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.append(chunk, got);
  } catch (const std::bad_alloc&) {
    fclose(f);
    b2w_set_error("b2w_edgelist_parse: out of memory");
    return B2W_ERR_NOMEM;
  }
  fclose(f);
  // Inputs on which this byte-level parser and the reference's text-mode read could disagree are left to the Python
  // parser (B2W_ERR_UNSUPPORTED): non-ASCII bytes (str.strip() is Unicode aware), NUL (the id blob is NUL separated),
  // a bare '\r' (universal newlines make it a line end).
  for (size_t i = 0; i < buf.size(); ++i) {
    const unsigned char c = (unsigned char)buf[i];
    if (c >= 0x80 || c == 0 || (c == '\r' && (i + 1 == buf.size() || buf[i + 1] != '\n'))) {
      b2w_set_error("b2w_edgelist_parse: non-ASCII / NUL / bare CR in the input (use the Python parser)");
      return B2W_ERR_UNSUPPORTED;
    }
  }
  b2w_edgelist* E = new (std::nothrow) b2w_edgelist();
  if (!E) { b2w_set_error("b2w_edgelist_parse: out of memory"); return B2W_ERR_NOMEM; }
  E->ids.init(1u << 16);
  const size_t dl = strlen(delimiter);
  const char* p = buf.data();
  const char* const end = p + buf.size();
  uint64_t line_no = 0;
  try {
    while (p < end) {
      const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
      const char* next = eol ? eol + 1 : end;
      const char* a = p;
      const char* b = eol ? eol : end;
      ++line_no;
      while (a < b && is_py_space((unsigned char)*a)) ++a;              // line.strip()
      while (b > a && is_py_space((unsigned char)b[-1])) --b;
      p = next;
      if (a == b) continue;                                             // blank line
      // split by the delimiter string
      const char* t0 = a; const char* t0e = nullptr;
      const char* t1 = nullptr; const char* t1e = nullptr;
      const char* tl = nullptr; const char* tle = nullptr;              // last term
      int nterms = 0;
      const char* s = a;
      for (;;) {
        const char* hit = nullptr;
        for (const char* q = s; q + dl <= b; ++q)
          if (*q == delimiter[0] && memcmp(q, delimiter, dl) == 0) { hit = q; break; }
        const char* te = hit ? hit : b;
        if (nterms == 0) t0e = te;
        else if (nterms == 1) { t1 = s; t1e = te; }
        tl = s; tle = te;
        ++nterms;
        if (!hit) break;
        s = hit + dl;
      }
      if (nterms < 2) {
        b2w_set_error("b2w_edgelist_parse: line %llu has fewer than two columns", (unsigned long long)line_no);
        delete E; return B2W_ERR_GRAPH;
      }
      double weight = 1.0;
      if (weighted) {
        if (nterms != 3) {
          b2w_set_error("Expecting three columns in the edge list file for a weighted graph, got %d instead (line %llu)",
                        nterms, (unsigned long long)line_no);
          delete E; return B2W_ERR_GRAPH;
        }
        while (tl < tle && is_py_space((unsigned char)*tl)) ++tl;       // float() strips whitespace
        while (tle > tl && is_py_space((unsigned char)tle[-1])) --tle;
        std::string tok(tl, tle);
        char* stop = nullptr;
        if (tok.find_first_of("_(") != std::string::npos) {             // float() accepts 1_0 and rejects nan(1): strtod differs
          delete E;
          b2w_set_error("b2w_edgelist_parse: weight token '%s' needs Python's float() (line %llu)", tok.c_str(), (unsigned long long)line_no);
          return B2W_ERR_UNSUPPORTED;
        }
        bool bad = tok.empty() || tok.find_first_of("xXpP") != std::string::npos;   // float() has no hex floats
        if (!bad) { weight = strtod(tok.c_str(), &stop); bad = stop != tok.c_str() + tok.size(); }
        if (bad) {
          b2w_set_error("could not convert string to float: '%s' (line %llu)", tok.c_str(), (unsigned long long)line_no);
          delete E; return B2W_ERR_GRAPH;
        }
      }
      if (weight <= 0) {                                                // ignored; ids not registered (graph.py:283)
        if (E->dropped_lines.size() < 20) E->dropped_lines.push_back(line_no);
        ++E->dropped;
        continue;
      }
      while (t0 < t0e && is_py_space((unsigned char)*t0)) ++t0;         // terms[0].strip(), terms[1].strip()
      while (t0e > t0 && is_py_space((unsigned char)t0e[-1])) --t0e;
      while (t1 < t1e && is_py_space((unsigned char)*t1)) ++t1;
      while (t1e > t1 && is_py_space((unsigned char)t1e[-1])) --t1e;
      const uint32_t ia = E->ids.intern(t0, (size_t)(t0e - t0));
      const uint32_t ib = E->ids.intern(t1, (size_t)(t1e - t1));
      E->src.push_back(ia);
      E->dst.push_back(ib);
      E->w.push_back(weight);
    }
  } catch (const std::bad_alloc&) {
    delete E;
    b2w_set_error("b2w_edgelist_parse: out of memory");
    return B2W_ERR_NOMEM;
  }
  *out = E;
  if (num_edges) *num_edges = E->src.size();
  if (num_nodes) *num_nodes = (uint32_t)E->ids.off.size();
  if (names_bytes) *names_bytes = E->ids.blob.size();
  if (num_dropped) *num_dropped = E->dropped;
  return B2W_OK;
}

extern "C" int b2w_edgelist_fetch(const b2w_edgelist* E, uint32_t* h_src, uint32_t* h_dst, double* h_weight,
                                  char* h_names, uint64_t* h_dropped_lines /* up to 20 */, uint32_t* n_dropped_lines) {
  if (!E) { b2w_set_error("b2w_edgelist_fetch: null handle"); return B2W_ERR_INVALID; }
  const size_t m = E->src.size();
  if (h_src && m) memcpy(h_src, E->src.data(), m * sizeof(uint32_t));
  if (h_dst && m) memcpy(h_dst, E->dst.data(), m * sizeof(uint32_t));
  if (h_weight && m) memcpy(h_weight, E->w.data(), m * sizeof(double));
  if (h_names && !E->ids.blob.empty()) memcpy(h_names, E->ids.blob.data(), E->ids.blob.size());
  if (h_dropped_lines) memcpy(h_dropped_lines, E->dropped_lines.data(), E->dropped_lines.size() * sizeof(uint64_t));
  if (n_dropped_lines) *n_dropped_lines = (uint32_t)E->dropped_lines.size();
  return B2W_OK;
}

extern "C" void b2w_edgelist_free(b2w_edgelist* E) { delete E; }

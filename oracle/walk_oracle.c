/*
 * walk_oracle.c -- CPU restatement of PecanPy's biased random-walk path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA engine in
 * pecanpy_b200/csrc.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load it; the product path never does.
 *
 * It restates, in plain C, the algorithm of the reference's Numba kernels.  All
 * file:line citations are relative to /root/reference/src/pecanpy/ :
 *
 *   orc_walk_*            <- pecanpy.py:164-210   Base._random_walks (walker loop)
 *   sparse_probs()        <- rw/sparse_rw.py:51-91   get_normalized_probs
 *                            rw/sparse_rw.py:93-130  get_extended_normalized_probs
 *   isnotin()             <- rw/sparse_rw.py:142-230
 *   isnotin_extended()    <- rw/sparse_rw.py:233-295
 *   otf_draw()            <- pecanpy.py:556-559 (cumsum + searchsorted + indices[...])
 *   alias_setup()         <- pecanpy.py:617-665
 *   alias_draw()          <- pecanpy.py:668-677
 *   precomp build         <- pecanpy.py:442-507
 *   precomp step          <- pecanpy.py:409-438
 *   dense_probs()         <- rw/dense_rw.py:34-72, 74-118
 *   dense step            <- pecanpy.py:596-612
 *   first-order modes     <- pecanpy.py:293-361 (FirstOrderUnweighted, PreCompFirstOrder)
 *
 * Numeric semantics: weights are copied as f32; `w /= q` is
 * f32(f64(w)/f64(q)); arr.sum() and np.cumsum are sequential left-to-right f32 (f64
 * for the dense graph); np.searchsorted(side='left') against a 53-bit uniform;
 * indices[indptr[cur]+choice] is read without a bounds check (choice may equal deg).
 *
 * Random streams (rng_mode):
 *   ORC_RNG_WORDS  : one global stream of raw 32-bit MT19937 outputs consumed by walkers
 *                    in row order -- replays the unmodified reference at 1 thread
 *                    (numba/cpython/randomimpl.py:134-147 random(), :454-533 randint()).
 *   ORC_RNG_FEED   : U[row, step-1] doubles supplied by the caller (OTF modes only).
 *   ORC_RNG_PHILOX : Philox4x32-10, key=(seed lo, seed hi), counter=(row lo,row hi,step,block).
 *
 * Parity pinning: tests/golden/*.npz were produced by importing the reference itself
 * (oracle/gen_golden.py); tests/test_oracle_golden.py checks this file against them,
 * including the reference's own known-answer vectors (test/test_walk.py:21-82).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_RNG_WORDS 0
#define ORC_RNG_FEED 1
#define ORC_RNG_PHILOX 2

#define ORC_MODE_SPARSE_OTF 0
#define ORC_MODE_PRECOMP 1
#define ORC_MODE_DENSE_OTF 2
#define ORC_MODE_FIRST_ORDER_UNWEIGHTED 3
#define ORC_MODE_PRECOMP_FIRST_ORDER 4

/* ------------------------------------------------------------------ Philox4x32-10 */
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------ RNG front end */
typedef struct {
  int mode;
  /* WORDS */
  const uint32_t* words; uint64_t pos; uint64_t nwords; int exhausted;
  /* FEED */
  const double* feed; uint32_t feed_ld;
  /* PHILOX */
  uint32_t key[2]; uint32_t buf[4]; uint32_t widx; uint32_t block;
  /* walker position */
  uint64_t row; uint32_t step;
} rng_t;

static inline void rng_begin_step(rng_t* r, uint64_t row, uint32_t step) {
  r->row = row; r->step = step; r->widx = 4; r->block = 0;
}

static inline uint32_t rng_word(rng_t* r) {
  if (r->mode == ORC_RNG_WORDS) {
    if (r->pos >= r->nwords) { r->exhausted = 1; return 0; }
    return r->words[r->pos++];
  }
  /* PHILOX: words of (row, step) = blocks b=0,1,... concatenated */
  if (r->widx == 4) {
    uint32_t ctr[4] = {(uint32_t)r->row, (uint32_t)(r->row >> 32), r->step, r->block};
    philox4x32_10(ctr, r->key, r->buf);
    r->block++; r->widx = 0;
  }
  return r->buf[r->widx++];
}

/* numba/cpython/randomimpl.py:134-147: (a>>5, b>>6) -> 53-bit double */
static inline double rng_uniform(rng_t* r) {
  if (r->mode == ORC_RNG_FEED) return r->feed[r->row * (uint64_t)r->feed_ld + (r->step - 1)];
  uint32_t a = rng_word(r) >> 5, b = rng_word(r) >> 6;
  return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

/* numba/cpython/randomimpl.py:454-533 (state 'np'): masked rejection on the low
 * bitlen(n-1) bits of successive 32-bit words; no word consumed when n == 1. */
static inline uint32_t rng_randint(rng_t* r, uint32_t n) {
  if (n == 1) return 0;
  uint32_t nm1 = n - 1;
  int nbits = 32 - __builtin_clz(nm1);
  uint32_t mask = 0xFFFFFFFFu >> (32 - nbits);
  uint32_t v;
  do { v = rng_word(r) & mask; } while (v >= n && !r->exhausted);
  return v;
}

/* ------------------------------------------------------------------ graph views */
typedef struct {
  uint32_t n;
  const uint32_t* indptr;   /* [n+1] */
  const uint32_t* indices;  /* [nnz+1]: one pad element (see `choice == deg`) */
  const float* data;        /* [nnz] */
} csr_t;

typedef struct {
  uint32_t n;
  const double* data;       /* [n,n] */
  const uint8_t* nonzero;   /* [n,n] */
} dense_t;

/* rw/sparse_rw.py:142-230 */
static void isnotin(const uint32_t* a1, uint32_t n1, const uint32_t* a2, uint32_t n2, uint8_t* ind) {
  for (uint32_t i = 0; i < n1; ++i) ind[i] = 1;
  uint32_t idx2 = 0;
  for (uint32_t idx1 = 0; idx1 < n1; ++idx1) {
    if (idx2 == n2) break;
    uint32_t p1 = a1[idx1], p2 = a2[idx2];
    if (p1 < p2) continue;
    if (p1 == p2) { ind[idx1] = 0; idx2++; continue; }
    for (uint32_t j = idx2; j < n2; ++j) {
      p2 = a2[j];
      if (p2 == p1) { ind[idx1] = 0; idx2 = j + 1; break; }
      if (p2 > p1) { idx2 = j; break; }
    }
  }
}

/* rw/sparse_rw.py:233-295 */
static void isnotin_extended(const uint32_t* a1, uint32_t n1, const uint32_t* a2, const float* w2,
                             uint32_t n2, const float* thr, uint8_t* ind, float* t) {
  for (uint32_t i = 0; i < n1; ++i) { ind[i] = 1; t[i] = 0.0f; }
  uint32_t idx2 = 0;
  for (uint32_t idx1 = 0; idx1 < n1; ++idx1) {
    if (idx2 >= n2) break;
    uint32_t p1 = a1[idx1], p2 = a2[idx2];
    if (p1 < p2) continue;
    if (p1 == p2) {
      if (w2[idx2] >= thr[p2]) ind[idx1] = 0; else t[idx1] = w2[idx2] / thr[p2];
      idx2++;
      continue;
    }
    for (uint32_t j = idx2 + 1; j < n2; ++j) {
      p2 = a2[j];
      if (p2 == p1) {
        if (w2[j] >= thr[p2]) ind[idx1] = 0; else t[idx1] = w2[j] / thr[p2];
        idx2 = j + 1;
        break;
      }
      if (p2 > p1) { idx2 = j; break; }
    }
  }
}

typedef struct { float* w; float* t; uint8_t* ind; uint32_t* stack; double* wd; uint32_t* nz; size_t cap; } scratch_t;

static void scratch_init(scratch_t* s, size_t cap) {
  s->cap = cap ? cap : 1;
  s->w = (float*)malloc(sizeof(float) * s->cap);
  s->t = (float*)malloc(sizeof(float) * s->cap);
  s->ind = (uint8_t*)malloc(s->cap);
  s->stack = (uint32_t*)malloc(sizeof(uint32_t) * 2 * s->cap);
  s->wd = NULL; s->nz = NULL;
}
static void scratch_init_dense(scratch_t* s, size_t n) {
  scratch_init(s, 1);
  s->wd = (double*)malloc(sizeof(double) * (n ? n : 1));
  s->nz = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
}
static void scratch_free(scratch_t* s) {
  free(s->w); free(s->t); free(s->ind); free(s->stack); free(s->wd); free(s->nz);
}

/* rw/sparse_rw.py:51-91 (extend=0) and :93-130 (extend=1).  Writes the NORMALISED
 * probabilities of cur's neighbours into s->w and returns the degree. */
static uint32_t sparse_probs(const csr_t* g, double p, double q, uint32_t cur, int has_prev,
                             uint32_t prev, int extend, const float* thr, scratch_t* s) {
  uint32_t start = g->indptr[cur], deg = g->indptr[cur + 1] - start;
  const uint32_t* nbrs = g->indices + start;
  float* w = s->w;
  for (uint32_t k = 0; k < deg; ++k) w[k] = g->data[start + k];      /* get_nbrs: f32 copy (:138) */
  if (has_prev) {
    uint32_t ps = g->indptr[prev], pdeg = g->indptr[prev + 1] - ps;
    const uint32_t* pn = g->indices + ps;
    if (!extend) {
      isnotin(nbrs, deg, pn, pdeg, s->ind);                           /* :83 */
      for (uint32_t k = 0; k < deg; ++k) if (nbrs[k] == prev) s->ind[k] = 0;   /* :84 */
      for (uint32_t k = 0; k < deg; ++k)
        if (s->ind[k]) w[k] = (float)((double)w[k] / q);              /* :86 */
    } else {
      isnotin_extended(nbrs, deg, pn, g->data + ps, pdeg, thr, s->ind, s->t);   /* :110-115 */
      for (uint32_t k = 0; k < deg; ++k) if (nbrs[k] == prev) s->ind[k] = 0;   /* :116 */
      double invq = 1.0 / q;
      double supp = invq < 1.0 ? invq : 1.0;                          /* np.minimum(1, 1/q) :124 */
      float thr_cur = thr[cur];
      for (uint32_t k = 0; k < deg; ++k) {
        if (!s->ind[k]) continue;
        double alpha = invq + (1.0 - invq) * (double)s->t[k];         /* :119 */
        if (w[k] < thr_cur) alpha = supp;                             /* :122-124 */
        w[k] = (float)((double)w[k] * alpha);                         /* :125 */
      }
    }
    for (uint32_t k = 0; k < deg; ++k)
      if (nbrs[k] == prev) w[k] = (float)((double)w[k] / p);          /* :87 / :126 */
  }
  float sum = 0.0f;
  for (uint32_t k = 0; k < deg; ++k) sum = sum + w[k];                /* sequential f32 sum */
  for (uint32_t k = 0; k < deg; ++k) w[k] = w[k] / sum;               /* :89 */
  return deg;
}

/* pecanpy.py:556-557: cdf = cumsum(probs) (sequential f32); searchsorted(cdf, u) left */
static inline uint32_t cumsum_search_f32(const float* pr, uint32_t deg, double u) {
  float c = 0.0f;
  for (uint32_t i = 0; i < deg; ++i) {
    c = c + pr[i];
    if (!((double)c < u)) return i;
  }
  return deg;
}

/* pecanpy.py:617-665 */
static void alias_setup(const float* probs, uint32_t k, uint32_t* j, float* q, uint32_t* smaller,
                        uint32_t* larger) {
  uint32_t sp = 0, lp = 0;
  for (uint32_t kk = 0; kk < k; ++kk) { q[kk] = 0.0f; j[kk] = 0; }
  for (uint32_t kk = 0; kk < k; ++kk) {
    q[kk] = (float)((double)k * (double)probs[kk]);                   /* int64 * f32 -> f64 -> f32 */
    if (q[kk] < 1.0f) smaller[sp++] = kk; else larger[lp++] = kk;
  }
  while (sp > 0 && lp > 0) {
    uint32_t small = smaller[--sp];
    uint32_t large = larger[--lp];
    j[small] = large;
    q[large] = (float)((double)(float)(q[large] + q[small]) - 1.0);   /* f32 add, then f64 - 1.0 */
    if (q[large] < 1.0f) smaller[sp++] = large; else larger[lp++] = large;
  }
}

/* pecanpy.py:668-677 */
static inline uint32_t alias_draw(const uint32_t* j, const float* q, uint32_t k, rng_t* r) {
  uint32_t kk = rng_randint(r, k);
  double u = rng_uniform(r);
  return (u < (double)q[kk]) ? kk : j[kk];
}

/* ------------------------------------------------------------------ exported: probabilities */
/* Normalised transition probabilities of one (cur, prev) pair; prev < 0 -> first order. */
uint32_t orc_sparse_probs(uint32_t n, const uint32_t* indptr, const uint32_t* indices, const float* data,
                          double p, double q, int extend, const float* thr, uint32_t cur, int64_t prev,
                          float* out) {
  csr_t g = {n, indptr, indices, data};
  scratch_t s; scratch_init(&s, indptr[cur + 1] - indptr[cur]);
  uint32_t deg = sparse_probs(&g, p, q, cur, prev >= 0, (uint32_t)(prev >= 0 ? prev : 0), extend, thr, &s);
  memcpy(out, s.w, sizeof(float) * deg);
  scratch_free(&s);
  return deg;
}

/* pecanpy.py:442-507: all (node, neighbour-slot) alias tables.  alias_indptr u64[n+1]
 * must hold cumsum(deg^2) (pecanpy.py:474-476). */
void orc_alias_build(uint32_t n, const uint32_t* indptr, const uint32_t* indices, const float* data,
                     double p, double q, int extend, const float* thr, const uint64_t* alias_indptr,
                     uint32_t* alias_j, float* alias_q) {
  csr_t g = {n, indptr, indices, data};
  uint32_t maxdeg = 0;
  for (uint32_t i = 0; i < n; ++i) { uint32_t d = indptr[i + 1] - indptr[i]; if (d > maxdeg) maxdeg = d; }
#pragma omp parallel
  {
    scratch_t s; scratch_init(&s, maxdeg);
#pragma omp for schedule(dynamic, 64)
    for (int64_t idx = 0; idx < (int64_t)n; ++idx) {
      uint32_t start = indptr[idx], deg = indptr[idx + 1] - start;
      for (uint32_t nb = 0; nb < deg; ++nb) {
        uint32_t prev = indices[start + nb];
        sparse_probs(&g, p, q, (uint32_t)idx, 1, prev, extend, thr, &s);
        uint64_t off = alias_indptr[idx] + (uint64_t)deg * nb;
        alias_setup(s.w, deg, alias_j + off, alias_q + off, s.stack, s.stack + maxdeg);
      }
    }
    scratch_free(&s);
  }
}

/* pecanpy.py:336-361: one first-order alias table per node (PreCompFirstOrder). */
void orc_alias_build_first_order(uint32_t n, const uint32_t* indptr, const uint32_t* indices,
                                 const float* data, uint32_t* alias_j, float* alias_q) {
  csr_t g = {n, indptr, indices, data};
  uint32_t maxdeg = 0;
  for (uint32_t i = 0; i < n; ++i) { uint32_t d = indptr[i + 1] - indptr[i]; if (d > maxdeg) maxdeg = d; }
  scratch_t s; scratch_init(&s, maxdeg);
  for (uint32_t idx = 0; idx < n; ++idx) {
    uint32_t start = indptr[idx], deg = indptr[idx + 1] - start;
    if (!deg) continue;
    sparse_probs(&g, 1.0, 1.0, idx, 0, 0, 0, NULL, &s);   /* rw/sparse_rw.py:37-49 */
    alias_setup(s.w, deg, alias_j + start, alias_q + start, s.stack, s.stack + maxdeg);
  }
  scratch_free(&s);
}

/* ------------------------------------------------------------------ dense probabilities */
/* rw/dense_rw.py:34-72 (extend=0), :74-118 (extend=1).  Writes the normalised
 * probabilities of the compressed row into s->wd, neighbour columns into s->nz. */
static uint32_t dense_probs(const dense_t* g, double p, double q, uint32_t cur, int has_prev,
                            uint32_t prev, int extend, const float* thr, scratch_t* s) {
  uint32_t n = g->n;
  const double* row = g->data + (size_t)cur * n;
  const uint8_t* nzc = g->nonzero + (size_t)cur * n;
  uint32_t d = 0;
  if (!has_prev) {
    for (uint32_t k = 0; k < n; ++k) if (nzc[k]) { s->wd[d] = row[k]; s->nz[d] = k; d++; }
  } else if (!extend) {
    const uint8_t* nzp = g->nonzero + (size_t)prev * n;
    for (uint32_t k = 0; k < n; ++k) {
      double w = row[k];
      int out = nzc[k] && !nzp[k] && (k != prev);                      /* :63-64 */
      if (out) w = w / q;                                             /* :66 */
      if (k == prev) w = w / p;                                       /* :67 (unconditional) */
      if (nzc[k]) { s->wd[d] = w; s->nz[d] = k; d++; }                /* :69 */
    }
  } else {
    const double* prow = g->data + (size_t)prev * n;
    double invq = 1.0 / q;
    double supp = invq < 1.0 ? invq : 1.0;
    double thr_cur = (double)thr[cur];
    for (uint32_t k = 0; k < n; ++k) {
      double w = row[k];
      int out = nzc[k] && (prow[k] < (double)thr[k]) && (k != prev);   /* :94-95 */
      if (out) {
        double t = prow[k] / (double)thr[k];                          /* :101 */
        double alpha = invq + (1.0 - invq) * t;                       /* :106 */
        if (w < thr_cur) alpha = supp;                                /* :109-111 */
        w = w * alpha;                                                /* :112 */
      }
      if (k == prev) w = w / p;                                       /* :113 */
      if (nzc[k]) { s->wd[d] = w; s->nz[d] = k; d++; }                /* :115 */
    }
  }
  double sum = 0.0;
  for (uint32_t k = 0; k < d; ++k) sum = sum + s->wd[k];
  for (uint32_t k = 0; k < d; ++k) s->wd[k] = s->wd[k] / sum;
  return d;
}

uint32_t orc_dense_probs(uint32_t n, const double* data, const uint8_t* nonzero, double p, double q,
                         int extend, const float* thr, uint32_t cur, int64_t prev, double* out,
                         uint32_t* out_cols) {
  dense_t g = {n, data, nonzero};
  scratch_t s; scratch_init_dense(&s, n);
  uint32_t d = dense_probs(&g, p, q, cur, prev >= 0, (uint32_t)(prev >= 0 ? prev : 0), extend, thr, &s);
  memcpy(out, s.wd, sizeof(double) * d);
  if (out_cols) memcpy(out_cols, s.nz, sizeof(uint32_t) * d);
  scratch_free(&s);
  return d;
}

/* ------------------------------------------------------------------ the walker loop */
typedef struct {
  int mode; int extend; double p, q;
  csr_t csr; dense_t dense;
  const float* thr;
  const uint64_t* alias_indptr; const uint32_t* alias_j; const float* alias_q;
} walk_cfg_t;

static inline int has_nbrs(const walk_cfg_t* c, uint32_t v) {
  if (c->mode == ORC_MODE_DENSE_OTF) {                                /* rw/dense_rw.py:25-30 */
    const uint8_t* nz = c->dense.nonzero + (size_t)v * c->dense.n;
    for (uint32_t k = 0; k < c->dense.n; ++k) if (nz[k]) return 1;
    return 0;
  }
  return c->csr.indptr[v] != c->csr.indptr[v + 1];                    /* rw/sparse_rw.py:17-18 */
}

static uint32_t move_forward(const walk_cfg_t* c, uint32_t cur, int has_prev, uint32_t prev, rng_t* r,
                             scratch_t* s) {
  switch (c->mode) {
    case ORC_MODE_SPARSE_OTF: {                                       /* pecanpy.py:543-559 */
      uint32_t deg = sparse_probs(&c->csr, c->p, c->q, cur, has_prev, prev, c->extend, c->thr, s);
      uint32_t choice = cumsum_search_f32(s->w, deg, rng_uniform(r));
      return c->csr.indices[c->csr.indptr[cur] + choice];
    }
    case ORC_MODE_PRECOMP: {                                          /* pecanpy.py:409-438 */
      uint32_t start = c->csr.indptr[cur], deg = c->csr.indptr[cur + 1] - start;
      uint32_t choice;
      if (!has_prev) {
        /* first step always uses the NON-extended probs (pecanpy.py:402,413) */
        sparse_probs(&c->csr, c->p, c->q, cur, 0, 0, 0, NULL, s);
        choice = cumsum_search_f32(s->w, deg, rng_uniform(r));
      } else {
        /* np.searchsorted(indices[start:end], prev) (left) */
        uint32_t lo = 0, hi = deg;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (c->csr.indices[start + mid] < prev) lo = mid + 1; else hi = mid; }
        uint64_t off = c->alias_indptr[cur] + (uint64_t)deg * lo;
        choice = alias_draw(c->alias_j + off, c->alias_q + off, deg, r);
      }
      return c->csr.indices[start + choice];
    }
    case ORC_MODE_DENSE_OTF: {                                        /* pecanpy.py:596-612 */
      uint32_t d = dense_probs(&c->dense, c->p, c->q, cur, has_prev, prev, c->extend, c->thr, s);
      double u = rng_uniform(r);
      double cs = 0.0; uint32_t choice = d;
      for (uint32_t i = 0; i < d; ++i) { cs = cs + s->wd[i]; if (!(cs < u)) { choice = i; break; } }
      /* nbrs[choice] with choice == d is out of bounds in the reference (UB); the engine
       * and this oracle both clamp to the last neighbour. */
      if (choice >= d) choice = d - 1;
      return s->nz[choice];
    }
    case ORC_MODE_FIRST_ORDER_UNWEIGHTED: {                           /* pecanpy.py:304-307 */
      uint32_t start = c->csr.indptr[cur], deg = c->csr.indptr[cur + 1] - start;
      return c->csr.indices[start + rng_randint(r, deg)];
    }
    case ORC_MODE_PRECOMP_FIRST_ORDER: {                              /* pecanpy.py:327-332 */
      uint32_t start = c->csr.indptr[cur], deg = c->csr.indptr[cur + 1] - start;
      uint32_t choice = alias_draw(c->alias_j + start, c->alias_q + start, deg, r);
      return c->csr.indices[start + choice];
    }
  }
  return 0;
}

/* pecanpy.py:189-206 for one row */
static void walk_row(const walk_cfg_t* c, uint64_t row, uint32_t start_node, uint32_t L, rng_t* r,
                     scratch_t* s, uint32_t* out /* [L+2] */) {
  memset(out, 0, sizeof(uint32_t) * (L + 2));
  out[0] = start_node;
  out[L + 1] = L + 1;
  if (!has_nbrs(c, start_node)) { out[L + 1] = 1; return; }
  rng_begin_step(r, row, 1);
  out[1] = move_forward(c, start_node, 0, 0, r, s);
  for (uint32_t j = 2; j <= L; ++j) {
    uint32_t cur = out[j - 1];
    if (!has_nbrs(c, cur)) { out[L + 1] = j; break; }
    rng_begin_step(r, row, j);
    out[j] = move_forward(c, cur, 1, out[j - 2], r, s);
  }
}

static uint32_t csr_maxdeg(const csr_t* g) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < g->n; ++i) { uint32_t d = g->indptr[i + 1] - g->indptr[i]; if (d > m) m = d; }
  return m;
}

/* Walks rows [row0, row0+n_rows) of the (already shuffled) start array.  `start` and `out`
 * are indexed from row0 (start[0] is row row0).  Returns 0, or -1 when the word stream
 * ran dry (ORC_RNG_WORDS). `nthreads` <= 0 -> all available (ignored for ORC_RNG_WORDS). */
static int run_walks(const walk_cfg_t* c, const uint32_t* start, uint64_t row0, uint64_t n_rows, uint32_t L,
                     int rng_mode, uint64_t seed, const uint32_t* words, uint64_t nwords,
                     const double* feed, uint32_t* out, int nthreads, uint64_t* words_used) {
  uint32_t maxdeg = (c->mode == ORC_MODE_DENSE_OTF) ? 0 : csr_maxdeg(&c->csr);
  int rc = 0;
  if (rng_mode == ORC_RNG_WORDS) {
    rng_t r; memset(&r, 0, sizeof r);
    r.mode = rng_mode; r.words = words; r.nwords = nwords;
    scratch_t s;
    if (c->mode == ORC_MODE_DENSE_OTF) scratch_init_dense(&s, c->dense.n); else scratch_init(&s, maxdeg);
    for (uint64_t i = 0; i < n_rows; ++i)
      walk_row(c, row0 + i, start[i], L, &r, &s, out + i * (uint64_t)(L + 2));
    scratch_free(&s);
    if (words_used) *words_used = r.pos;
    return r.exhausted ? -1 : 0;
  }
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    rng_t r; memset(&r, 0, sizeof r);
    r.mode = rng_mode; r.feed = feed; r.feed_ld = L;
    r.key[0] = (uint32_t)seed; r.key[1] = (uint32_t)(seed >> 32);
    scratch_t s;
    if (c->mode == ORC_MODE_DENSE_OTF) scratch_init_dense(&s, c->dense.n); else scratch_init(&s, maxdeg);
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)n_rows; ++i) {
      /* FEED is addressed by the row index relative to row0 (the caller passes U for these rows) */
      uint64_t row = (rng_mode == ORC_RNG_FEED) ? (uint64_t)i : row0 + (uint64_t)i;
      walk_row(c, row, start[i], L, &r, &s, out + (uint64_t)i * (L + 2));
    }
    scratch_free(&s);
  }
  return rc;
}

int orc_walk_csr(int mode, uint32_t n, const uint32_t* indptr, const uint32_t* indices, const float* data,
                 double p, double q, int extend, const float* thr,
                 const uint64_t* alias_indptr, const uint32_t* alias_j, const float* alias_q,
                 const uint32_t* start, uint64_t row0, uint64_t n_rows, uint32_t L,
                 int rng_mode, uint64_t seed, const uint32_t* words, uint64_t nwords, const double* feed,
                 uint32_t* out, int nthreads, uint64_t* words_used) {
  walk_cfg_t c; memset(&c, 0, sizeof c);
  c.mode = mode; c.extend = extend; c.p = p; c.q = q;
  c.csr.n = n; c.csr.indptr = indptr; c.csr.indices = indices; c.csr.data = data;
  c.thr = thr; c.alias_indptr = alias_indptr; c.alias_j = alias_j; c.alias_q = alias_q;
  return run_walks(&c, start, row0, n_rows, L, rng_mode, seed, words, nwords, feed, out, nthreads, words_used);
}

int orc_walk_dense(uint32_t n, const double* data, const uint8_t* nonzero, double p, double q, int extend,
                   const float* thr, const uint32_t* start, uint64_t row0, uint64_t n_rows, uint32_t L,
                   int rng_mode, uint64_t seed, const uint32_t* words, uint64_t nwords, const double* feed,
                   uint32_t* out, int nthreads, uint64_t* words_used) {
  walk_cfg_t c; memset(&c, 0, sizeof c);
  c.mode = ORC_MODE_DENSE_OTF; c.extend = extend; c.p = p; c.q = q;
  c.dense.n = n; c.dense.data = data; c.dense.nonzero = nonzero; c.thr = thr;
  return run_walks(&c, start, row0, n_rows, L, rng_mode, seed, words, nwords, feed, out, nthreads, words_used);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// b2w_start.cu -- the start array of Base.simulate_walks (reference pecanpy.py:135-141), on the host, natively.
//
//     nodes = np.array(range(num_nodes), dtype=np.uint32)
//     start = np.concatenate([nodes] * num_walks)
//     np.random.seed(random_state); np.random.shuffle(start)          # "for balanced work load"
//
// The shuffle fixes the ROW ORDER of the walk matrix, so it is part of the result and is kept bit for bit.  It is
// NumPy's LEGACY generator (third party, not under /root/reference: numpy/random/mtrand.pyx RandomState.shuffle ->
// _shuffle_raw, and random_interval in numpy/random/src/distributions/distributions.c):
//     for i = n-1 .. 1:   j = random_interval(i);   swap(x[i], x[j])
//     random_interval(max): mask = smallest 2^k - 1 >= max;  draw 32-bit MT19937 words (64-bit = two words, high
//                           first, when max > 0xffffffff) until (word & mask) <= max
// At 10^7 walkers NumPy needs 0.3-0.7 s for it -- five to ten times the GPU's whole walk + copy (64 ms) -- mostly
// in the rejection loop, whose test fails unpredictably for up to half of the words: the positions j depend on the
// generator only, never on the data, so here they are drawn a chunk ahead of the swaps with a branch-free loop (and
// prefetched).  The caller seeds NumPy exactly as the reference does, hands over the state of the
// global generator (np.random.get_state(): key[624], pos) and stores the returned state back (np.random.set_state):
// array AND generator state after the call are the reference's (tests/test_start_array.py, against NumPy itself).
#include <sys/mman.h>

#include <cstdint>
#include <cstring>

#include "b2w_common.cuh"

namespace {

struct Mt19937 {
  uint32_t* key;   // [624], caller's storage
  int pos;

  void twist() {                                                      // mt19937_gen (numpy/random/src/mt19937/mt19937.c)
    constexpr int N = 624, M = 397;
    constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    uint32_t y;
    int kk = 0;
    for (; kk < N - M; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1u) & MATRIX_A);
    }
    for (; kk < N - 1; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1u) & MATRIX_A);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1u) & MATRIX_A);
    pos = 0;
  }
  inline uint32_t next32() {
    if (pos == 624) twist();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline uint64_t interval(uint64_t max) {                            // random_interval (distributions.c)
    if (max == 0) return 0;
    uint64_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t v;
    if (max <= 0xffffffffull) {
      while ((v = (next32() & mask)) > max) {}
    } else {
      for (;;) {
        const uint64_t hi = next32();
        v = ((hi << 32) | next32()) & mask;                           // mt19937_next64: high word first
        if (v <= max) break;
      }
    }
    return v;
  }
};

}  // namespace

extern "C" int b2w_shuffled_start(uint32_t num_nodes, uint32_t num_walks, uint32_t* mt_key, int32_t* mt_pos,
                                  uint32_t* h_start) {
  if (!mt_key || !mt_pos || (!h_start && num_nodes && num_walks)) { b2w_set_error("b2w_shuffled_start: null pointer"); return B2W_ERR_INVALID; }
  if (*mt_pos < 0 || *mt_pos > 624) { b2w_set_error("b2w_shuffled_start: MT19937 position %d out of range", (int)*mt_pos); return B2W_ERR_INVALID; }
  const uint64_t n = (uint64_t)num_nodes * num_walks;
#ifdef MADV_HUGEPAGE
  // the swaps touch the array at random: with 4 KB pages every one of them is a TLB miss as well.  The caller's
  // buffer is normally fresh and untouched (np.empty): ask for huge pages before the first touch (a hint; ignored
  // where transparent huge pages are off)
  if (n * 4 >= (8u << 20)) {
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(h_start) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(h_start) + n * 4) & ~(uintptr_t)((2u << 20) - 1);
    if (hi > lo) (void)madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
  }
#endif
  for (uint32_t w = 0; w < num_walks; ++w) {
    uint32_t* dst = h_start + (uint64_t)w * num_nodes;
    for (uint32_t v = 0; v < num_nodes; ++v) dst[v] = v;
  }
  if (n < 2) return B2W_OK;                                           // (np.random.shuffle draws nothing)
  Mt19937 g{mt_key, (int)*mt_pos};
  uint64_t i = n - 1;
  // positions that need 64-bit draws (arrays of 2^32 entries and more): the plain loop
  for (; i > 0xffffffffull; --i) {
    const uint64_t j = g.interval(i);
    const uint32_t t = h_start[j]; h_start[j] = h_start[i]; h_start[i] = t;
  }
  // 32-bit draws.  The rejection test `(word & mask) <= i` fails for up to half of the words, unpredictably: as a
  // branch it is the whole cost of the shuffle.  Here the partners of the next CHUNK positions are drawn first,
  // branch-free (every candidate is stored, the cursor advances only when it was accepted), then the swaps run.
  constexpr uint32_t CHUNK = 2048;
  uint32_t partner[CHUNK + 1];
  while (i >= 1) {
    const uint32_t want = (uint32_t)(i < CHUNK ? i : CHUNK);
    uint32_t cnt = 0, cur = (uint32_t)i;
    while (cnt < want) {
      const uint32_t v = g.next32() & (0xffffffffu >> __builtin_clz(cur));   // mask: smallest 2^k - 1 >= cur
      const uint32_t ok = v <= cur;
      partner[cnt] = v;
      cnt += ok;
      cur -= ok;
    }
    for (uint32_t k = 0; k < want; ++k) __builtin_prefetch(h_start + partner[k], 1);
    for (uint32_t k = 0; k < want; ++k) {
      const uint64_t a = i - k, j = partner[k];
      const uint32_t t = h_start[j]; h_start[j] = h_start[a]; h_start[a] = t;
    }
    i -= want;
  }
  *mt_pos = g.pos;
  return B2W_OK;
}

// Host-side exact emulation of CPython's `random.sample` stream (MT19937 + getrandbits + _randbelow +
// the pool / set selection of Lib/random.py).  The reference draws its contrastive negatives with Python's
// global `random` inside forward() (model/DCNet_model.py:87 and :413); to reproduce the same indices without
// ~B*B*N0 interpreter-level calls per step the stream is advanced here in C and handed back to Python
// (random.setstate) so later users of `random` see the state the reference would have left.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/dcnet_b200.h"
#include "host_error.h"

namespace {

// MT19937 on a private copy of the state (no aliasing with the caller's buffer), CPython's genrand_uint32 tempering.
struct MT {
  uint32_t mt[624];
  uint32_t pos;
  explicit MT(const uint32_t* st) { std::memcpy(mt, st, sizeof(mt)); pos = st[624]; }
  void store(uint32_t* st) const { std::memcpy(st, mt, sizeof(mt)); st[624] = pos; }
  void regen() {
    const uint32_t N = 624, M = 397;
    uint32_t kk;
    for (kk = 0; kk < N - M; kk++) {
      const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < N - 1; kk++) {
      const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk - (N - M)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    const uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    pos = 0;
  }
  inline uint32_t next() {
    if (__builtin_expect(pos >= 624, 0)) regen();
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // random.Random._randbelow_with_getrandbits(n), n >= 1; shift = 32 - n.bit_length()
  inline uint32_t randbelow(uint32_t n, int shift) {
    uint32_t r = next() >> shift;
    while (r >= n) r = next() >> shift;
    return r;
  }
};

inline int bit_shift(uint32_t n) { return 32 - (32 - __builtin_clz(n)); }

inline int set_size_threshold(int k) {
  int setsize = 21;
  if (k > 5) setsize += (int)std::pow(4.0, std::ceil(std::log((double)k * 3.0) / std::log(4.0)));
  return setsize;
}

// random.sample(range(n), k) -> positions.  `pool` is scratch of >= n ints.
inline void sample_positions(MT& g, int n, int k, int setsize, int* pool, int* out) {
  if (n <= setsize) {
    for (int i = 0; i < n; i++) pool[i] = i;
    for (int i = 0; i < k; i++) {
      const uint32_t m = (uint32_t)(n - i);
      const uint32_t j = g.randbelow(m, bit_shift(m));
      out[i] = pool[j];
      pool[j] = pool[n - i - 1];
    }
  } else {
    const int sh = bit_shift((uint32_t)n);
    for (int i = 0; i < k; i++) {
      uint32_t j;
      bool dup;
      do {
        j = g.randbelow((uint32_t)n, sh);
        dup = false;
        for (int t = 0; t < i; t++) dup |= (out[t] == (int)j);
      } while (dup);
      out[i] = (int)j;
    }
  }
}

}  // namespace

extern "C" int dcnet_pyrandom_interframe(uint32_t* st, int P, int top_k, int N0, int neg_n, int* negpos) {
  if (!st || !negpos || P < 0 || top_k < 0 || neg_n < 0 || N0 - 1 < neg_n) return dcnet_set_error(-1, "pyrandom_interframe: bad arguments");
  MT g(st);
  const int setsize = set_size_threshold(neg_n);
  std::vector<int> pool(N0);
  for (int p = 0; p < P; p++)
    for (int r = 0; r < top_k; r++)
      sample_positions(g, N0 - 1, neg_n, setsize, pool.data(), negpos + ((size_t)p * top_k + r) * neg_n);
  g.store(st);
  return 0;
}

extern "C" int dcnet_pyrandom_crossmodal(uint32_t* st, int B, int N0, int neg_n, long long* negidx) {
  if (!st || !negidx || B < 1 || neg_n < 0 || N0 - 1 < neg_n) return dcnet_set_error(-1, "pyrandom_crossmodal: bad arguments");
  MT g(st);
  const int setsize = set_size_threshold(neg_n);
  std::vector<int> pool(N0);
  std::vector<int> tmp(neg_n > 0 ? neg_n : 1);
  for (int ii = 0; ii < B; ii++)
    for (int jj = 0; jj < N0; jj++) {
      for (int index = 0; index < B; index++) {
        const int n = (index == ii) ? N0 - 1 : N0;
        sample_positions(g, n, neg_n, setsize, pool.data(), tmp.data());
        if (index == B - 1) {
          long long* o = negidx + ((size_t)ii * N0 + jj) * neg_n;
          // population = range(N0) without jj when index == ii: position -> pixel
          for (int t = 0; t < neg_n; t++) o[t] = (index == ii && tmp[t] >= jj) ? tmp[t] + 1 : tmp[t];
        }
      }
    }
  g.store(st);
  return 0;
}

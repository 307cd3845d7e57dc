// Host-side exact emulation of CPython's `random.sample` stream (MT19937 + getrandbits + _randbelow +
// the pool / set selection of Lib/random.py).  The reference draws its contrastive negatives with Python's
// global `random` inside forward() (model/DCNet_model.py:87 and :413); to reproduce the same indices without
// ~B*B*N0 interpreter-level calls per step the stream is advanced here in C and handed back to Python
// (random.setstate) so later users of `random` see the state the reference would have left.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dcnet_b200.h"
#include "host_error.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define DCNET_RNG_X86 1
#endif

namespace {

// MT19937 on a private copy of the state (no aliasing with the caller's buffer), CPython's genrand_uint32 tempering.
struct MT {
  uint32_t mt[624];
  uint32_t tv[624];          // the tempered outputs of the current block (filled by regen: both loops vectorise)
  uint32_t pos;
  bool lazy_tv = false;      // the block sampler tempers inside its own pass over mt[] and does not need tv[]
  explicit MT(const uint32_t* st) {
    std::memcpy(mt, st, sizeof(mt));
    pos = st[624];
    temper_block();
  }
  void store(uint32_t* st) const { std::memcpy(st, mt, sizeof(mt)); st[624] = pos; }
  void temper_block() {
    for (int i = 0; i < 624; i++) {
      uint32_t y = mt[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      tv[i] = y;
    }
  }
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
    if (!lazy_tv) temper_block();
  }
  static inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline uint32_t next() {
    if (__builtin_expect(pos >= 624, 0)) regen();
    return tv[pos++];
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
  if (k <= 0) return;
  if (n <= setsize) {
    for (int i = 0; i < n; i++) pool[i] = i;
    for (int i = 0; i < k; i++) {
      const uint32_t m = (uint32_t)(n - i);
      const uint32_t j = g.randbelow(m, bit_shift(m));
      out[i] = pool[j];
      pool[j] = pool[n - i - 1];
    }
  } else if (n <= 64) {
    // set path, population of at most 64: the selected set is a 64-bit mask and one raw draw is one branch-free iteration
    // (rejected by _randbelow, rejected as a duplicate, or accepted -- the stream position only depends on the count)
    const int sh = bit_shift((uint32_t)n);
    uint64_t mask = 0;
    int i = 0;
    do {
      const uint32_t r = g.next() >> sh;
      const uint64_t bit = 1ull << (r & 63u);
      const bool ok = (r < (uint32_t)n) & ((mask & bit) == 0);
      out[i] = (int)r;
      mask |= ok ? bit : 0ull;
      i += ok;
    } while (i < k);
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


// ---- block sampler for the cross-modal draw ---------------------------------------------------------------------------------
// The reference draws B*N0*B samples per step and keeps one in B of them (model/DCNet_model.py:81-96); the others only advance
// the stream, by a count that depends on every rejection and duplicate, so all of them have to be replayed.  Per 624-word MT
// block the raw words are reduced once to the list of values _randbelow(N0) would accept (branch-free compaction); a sample
// over range(N0) then takes the next K accepted values, which are pairwise distinct 85-95 % of the time (checked without
// branches), and only otherwise falls back to the one-by-one loop.  The one sample in B over range(N0-1) (own image, pixel jj
// removed) may use another bit length, so it reads the raw words directly from the exact stream position and the accepted-list
// cursor is re-derived from the per-word prefix counts afterwards.
struct CompactLut {
  alignas(16) uint8_t shuf[256][16];     // pshufb control that moves the 16-bit lanes selected by the mask to the front
  uint8_t pop[256];
  uint8_t nth[256][8];                   // nth[m][j] = bit position of the (j+1)-th set bit of m
  CompactLut() {
    for (int m = 0; m < 256; m++) {
      int c = 0;
      for (int i = 0; i < 16; i++) shuf[m][i] = 0x80;
      for (int b = 0; b < 8; b++) {
        nth[m][b] = 0;
        if (m & (1 << b)) {
          shuf[m][2 * c] = (uint8_t)(2 * b);
          shuf[m][2 * c + 1] = (uint8_t)(2 * b + 1);
          nth[m][c] = (uint8_t)b;
          c++;
        }
      }
      pop[m] = (uint8_t)c;
    }
  }
};
const CompactLut g_lut;

constexpr int NGRP = 624 / 8;

// One pass over the untempered block: temper, >> sh, keep the values below n (in order) in acc[]; per group of 8 words the
// acceptance mask m8[] and the number of values accepted before the group c8[].  Returns the number accepted.
int compact_scalar(const uint32_t* mt, int sh, uint32_t n, uint16_t* acc, uint8_t* m8, uint16_t* c8) {
  int c = 0;
  for (int g = 0; g < NGRP; g++) {
    int m = 0;
    c8[g] = (uint16_t)c;
    for (int b = 0; b < 8; b++) {
      const uint32_t r = MT::temper(mt[8 * g + b]) >> sh;
      acc[c] = (uint16_t)r;
      const int ok = r < n;
      c += ok;
      m |= ok << b;
    }
    m8[g] = (uint8_t)m;
  }
  c8[NGRP] = (uint16_t)c;
  return c;
}

#ifdef DCNET_RNG_X86
__attribute__((target("ssse3")))
int compact_ssse3(const uint32_t* mt, int sh, uint32_t n, uint16_t* acc, uint8_t* m8, uint16_t* c8) {
  const __m128i shc = _mm_cvtsi32_si128(sh);
  const __m128i nv = _mm_set1_epi16((short)n);                 // n <= 16384, values < 32768: signed 16-bit compares are exact
  const __m128i k1 = _mm_set1_epi32((int)0x9d2c5680u), k2 = _mm_set1_epi32((int)0xefc60000u);
  int c = 0;
  for (int g = 0; g < NGRP; g++) {
    __m128i y0 = _mm_loadu_si128((const __m128i*)(mt + 8 * g));
    __m128i y1 = _mm_loadu_si128((const __m128i*)(mt + 8 * g + 4));
    y0 = _mm_xor_si128(y0, _mm_srli_epi32(y0, 11));
    y1 = _mm_xor_si128(y1, _mm_srli_epi32(y1, 11));
    y0 = _mm_xor_si128(y0, _mm_and_si128(_mm_slli_epi32(y0, 7), k1));
    y1 = _mm_xor_si128(y1, _mm_and_si128(_mm_slli_epi32(y1, 7), k1));
    y0 = _mm_xor_si128(y0, _mm_and_si128(_mm_slli_epi32(y0, 15), k2));
    y1 = _mm_xor_si128(y1, _mm_and_si128(_mm_slli_epi32(y1, 15), k2));
    y0 = _mm_xor_si128(y0, _mm_srli_epi32(y0, 18));
    y1 = _mm_xor_si128(y1, _mm_srli_epi32(y1, 18));
    const __m128i r = _mm_packs_epi32(_mm_srl_epi32(y0, shc), _mm_srl_epi32(y1, shc));
    const __m128i lt = _mm_cmpgt_epi16(nv, r);
    const int m = _mm_movemask_epi8(_mm_packs_epi16(lt, lt)) & 0xff;
    _mm_storeu_si128((__m128i*)(acc + c), _mm_shuffle_epi8(r, _mm_load_si128((const __m128i*)g_lut.shuf[m])));
    m8[g] = (uint8_t)m;
    c8[g] = (uint16_t)c;
    c += g_lut.pop[m];
  }
  c8[NGRP] = (uint16_t)c;
  return c;
}
#endif

typedef int (*compact_fn)(const uint32_t*, int, uint32_t, uint16_t*, uint8_t*, uint16_t*);
compact_fn pick_compact() {
#ifdef DCNET_RNG_X86
  if (__builtin_cpu_supports("ssse3")) return compact_ssse3;
#endif
  return compact_scalar;
}

struct BlockSampler {
  MT& g;
  const uint32_t n_main;
  const int sh_main;
  const compact_fn compact_block;
  uint16_t acc[624 + 8];
  uint8_t m8[NGRP];
  uint16_t c8[NGRP + 1];
  int cnt = 0, a = 0;
  int p = 0;              // exact raw position, valid when p_exact
  bool p_exact = true;

  BlockSampler(MT& g_, int n, compact_fn f) : g(g_), n_main((uint32_t)n), sh_main(bit_shift((uint32_t)n)), compact_block(f) {
    g.lazy_tv = true;
    compact();
    p = (int)g.pos;
    a = accepted_before(p);
  }
  void compact() { cnt = compact_block(g.mt, sh_main, n_main, acc, m8, c8); }
  void next_block() { g.regen(); compact(); }
  // number of accepted values among the raw words [0, raw)
  int accepted_before(int raw) const {
    if (raw >= 624) return cnt;
    return c8[raw >> 3] + g_lut.pop[m8[raw >> 3] & ((1u << (raw & 7)) - 1u)];
  }
  // raw position of the first word not consumed yet
  int raw_pos() const {
    if (p_exact) return p;
    // a >= 1 values of this block are consumed, the last one an accepted word: the position after accepted value #a
    int lo = 0, hi = NGRP - 1;                 // largest group with c8[group] < a
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (c8[mid] < a) lo = mid; else hi = mid - 1;
    }
    return 8 * lo + g_lut.nth[m8[lo]][a - c8[lo] - 1] + 1;
  }
  template <int K>
  inline void take_main(bool emit, int* out) {
    if (a + K <= cnt) {
      const uint16_t* v = acc + a;
      bool distinct = true;
#pragma GCC unroll 8
      for (int i = 1; i < K; i++)
#pragma GCC unroll 8
        for (int j = 0; j < i; j++) distinct &= (v[i] != v[j]);
      if (__builtin_expect(distinct, 1)) {
        if (emit)
          for (int i = 0; i < K; i++) out[i] = v[i];
        a += K;
        p_exact = false;
        return;
      }
    }
    int sel[K];
    int i = 0;
    while (i < K) {
      if (a == cnt) { next_block(); a = 0; }
      const int v = acc[a++];
      bool ok = true;
      for (int t = 0; t < i; t++) ok &= (sel[t] != v);
      if (ok) sel[i++] = v;
    }
    if (emit)
      for (int t = 0; t < K; t++) out[t] = sel[t];
    p_exact = false;
  }
  // random.sample(range(n), K) through the set path, read word by word from the exact position
  template <int K>
  inline void take_direct(uint32_t n, int* out) {
    int q = raw_pos();
    const int sh = bit_shift(n);
    int i = 0;
    while (i < K) {
      if (q == 624) { next_block(); q = 0; }
      const int v = (int)(MT::temper(g.mt[q++]) >> sh);
      bool ok = (uint32_t)v < n;
      for (int t = 0; t < i; t++) ok &= (out[t] != v);
      if (ok) out[i++] = v;
    }
    p = q;
    p_exact = true;
    a = accepted_before(q);
  }
  void finish() {
    g.pos = (uint32_t)raw_pos();
    g.lazy_tv = false;
  }
};

template <int K>
void crossmodal_blocks(MT& g, int B, int N0, long long* negidx, compact_fn f) {
  BlockSampler bs(g, N0, f);
  int tmp[K];
  for (int ii = 0; ii < B; ii++)
    for (int jj = 0; jj < N0; jj++) {
      for (int index = 0; index < B; index++) {
        const bool emit = index == B - 1;
        if (index == ii) bs.template take_direct<K>((uint32_t)(N0 - 1), tmp);
        else bs.template take_main<K>(emit, tmp);
        if (emit) {
          long long* o = negidx + ((size_t)ii * N0 + jj) * K;
          for (int t = 0; t < K; t++) o[t] = (index == ii && tmp[t] >= jj) ? tmp[t] + 1 : tmp[t];
        }
      }
    }
  bs.finish();
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
  if (neg_n == 5 && N0 - 1 > setsize && N0 <= 16384) {      // both populations on the set path: block sampler
    static const compact_fn f = pick_compact();
    crossmodal_blocks<5>(g, B, N0, negidx, getenv("DCNET_RNG_SCALAR") ? compact_scalar : f);
    g.store(st);
    return 0;
  }
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

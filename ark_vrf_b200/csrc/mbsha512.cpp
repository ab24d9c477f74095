// Multi-buffer SHA-512 (see mbsha512.h).  Host-only translation unit, compiled by the host compiler; the AVX-512
// code is confined to one function with a target attribute and selected at run time.
#include "mbsha512.h"

#include <immintrin.h>
#include <string.h>

#include <algorithm>

namespace avrf {

static const uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL,
    0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL,
    0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL,
    0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL,
    0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL,
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL,
    0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL,
    0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL,
    0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL,
    0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL,
    0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL,
    0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL,
    0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

static const uint64_t IV512[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static inline uint64_t load_be64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return __builtin_bswap64(v);
}

void sha512_blocks_x1(uint64_t h[8], const uint8_t* p, size_t nblk) {
  for (size_t blk = 0; blk < nblk; blk++, p += 128) {
    uint64_t w[16], s[8];
    for (int i = 0; i < 16; i++) w[i] = load_be64(p + 8 * i);
    for (int i = 0; i < 8; i++) s[i] = h[i];
    for (int r = 0; r < 80; r++) {
      if (r >= 16) {
        uint64_t w15 = w[(r + 1) & 15], w2 = w[(r + 14) & 15];
        w[r & 15] += (rotr(w15, 1) ^ rotr(w15, 8) ^ (w15 >> 7)) + w[(r + 9) & 15] + (rotr(w2, 19) ^ rotr(w2, 61) ^ (w2 >> 6));
      }
      uint64_t a = s[0], b = s[1], c = s[2], e = s[4], f = s[5], g = s[6];
      uint64_t t1 = s[7] + (rotr(e, 14) ^ rotr(e, 18) ^ rotr(e, 41)) + ((e & f) ^ (~e & g)) + K512[r] + w[r & 15];
      uint64_t t2 = (rotr(a, 28) ^ rotr(a, 34) ^ rotr(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
      s[7] = g; s[6] = f; s[5] = e; s[4] = s[3] + t1; s[3] = c; s[2] = b; s[1] = a; s[0] = t1 + t2;
    }
    for (int i = 0; i < 8; i++) h[i] += s[i];
  }
}

__attribute__((target("avx512f,avx512bw"))) static void blocks_x8_avx512(uint64_t h[8][8], const uint8_t* const ptr[8],
                                                                          size_t nblk, unsigned mask) {
  const __m512i bsw = _mm512_set_epi8(56, 57, 58, 59, 60, 61, 62, 63, 48, 49, 50, 51, 52, 53, 54, 55, 40, 41, 42, 43, 44, 45,
                                      46, 47, 32, 33, 34, 35, 36, 37, 38, 39, 24, 25, 26, 27, 28, 29, 30, 31, 16, 17, 18, 19,
                                      20, 21, 22, 23, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7);
  alignas(64) uint64_t tr[8][8];                       // tr[word][lane]
  for (int i = 0; i < 8; i++)
    for (int l = 0; l < 8; l++) tr[i][l] = h[l][i];
  __m512i st[8];
  for (int i = 0; i < 8; i++) st[i] = _mm512_load_si512((const void*)tr[i]);
  __m512i addr = _mm512_loadu_si512((const void*)ptr);  // eight lane pointers
  const __m512i step = _mm512_set1_epi64(128);
  for (size_t blk = 0; blk < nblk; blk++) {
    __m512i w[16];
    for (int t = 0; t < 16; t++) {
      __m512i v = _mm512_i64gather_epi64(_mm512_add_epi64(addr, _mm512_set1_epi64(8 * t)), nullptr, 1);
      w[t] = _mm512_shuffle_epi8(v, bsw);
    }
    addr = _mm512_add_epi64(addr, step);
    __m512i a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], hh = st[7];
#define AVRF_RND(a, b, c, d, e, f, g, h, kw)                                                                             \
  {                                                                                                                     \
    __m512i s1 = _mm512_ternarylogic_epi64(_mm512_ror_epi64(e, 14), _mm512_ror_epi64(e, 18), _mm512_ror_epi64(e, 41), 0x96); \
    __m512i ch = _mm512_ternarylogic_epi64(e, f, g, 0xCA);                                                              \
    __m512i t1 = _mm512_add_epi64(_mm512_add_epi64(h, s1), _mm512_add_epi64(ch, kw));                                   \
    __m512i s0 = _mm512_ternarylogic_epi64(_mm512_ror_epi64(a, 28), _mm512_ror_epi64(a, 34), _mm512_ror_epi64(a, 39), 0x96); \
    __m512i mj = _mm512_ternarylogic_epi64(a, b, c, 0xE8);                                                              \
    d = _mm512_add_epi64(d, t1);                                                                                        \
    h = _mm512_add_epi64(t1, _mm512_add_epi64(s0, mj));                                                                 \
  }
#define AVRF_KW(i) _mm512_add_epi64(w[i], _mm512_set1_epi64((long long)K512[r + i]))
    for (int r = 0; r < 80; r += 16) {
      if (r) {
        for (int t = 0; t < 16; t++) {
          __m512i w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
          __m512i s0 = _mm512_ternarylogic_epi64(_mm512_ror_epi64(w15, 1), _mm512_ror_epi64(w15, 8), _mm512_srli_epi64(w15, 7), 0x96);
          __m512i s1 = _mm512_ternarylogic_epi64(_mm512_ror_epi64(w2, 19), _mm512_ror_epi64(w2, 61), _mm512_srli_epi64(w2, 6), 0x96);
          w[t] = _mm512_add_epi64(_mm512_add_epi64(w[t], s0), _mm512_add_epi64(w[(t + 9) & 15], s1));
        }
      }
      AVRF_RND(a, b, c, d, e, f, g, hh, AVRF_KW(0)) AVRF_RND(hh, a, b, c, d, e, f, g, AVRF_KW(1))
      AVRF_RND(g, hh, a, b, c, d, e, f, AVRF_KW(2)) AVRF_RND(f, g, hh, a, b, c, d, e, AVRF_KW(3))
      AVRF_RND(e, f, g, hh, a, b, c, d, AVRF_KW(4)) AVRF_RND(d, e, f, g, hh, a, b, c, AVRF_KW(5))
      AVRF_RND(c, d, e, f, g, hh, a, b, AVRF_KW(6)) AVRF_RND(b, c, d, e, f, g, hh, a, AVRF_KW(7))
      AVRF_RND(a, b, c, d, e, f, g, hh, AVRF_KW(8)) AVRF_RND(hh, a, b, c, d, e, f, g, AVRF_KW(9))
      AVRF_RND(g, hh, a, b, c, d, e, f, AVRF_KW(10)) AVRF_RND(f, g, hh, a, b, c, d, e, AVRF_KW(11))
      AVRF_RND(e, f, g, hh, a, b, c, d, AVRF_KW(12)) AVRF_RND(d, e, f, g, hh, a, b, c, AVRF_KW(13))
      AVRF_RND(c, d, e, f, g, hh, a, b, AVRF_KW(14)) AVRF_RND(b, c, d, e, f, g, hh, a, AVRF_KW(15))
    }
#undef AVRF_RND
#undef AVRF_KW
    st[0] = _mm512_add_epi64(st[0], a); st[1] = _mm512_add_epi64(st[1], b); st[2] = _mm512_add_epi64(st[2], c);
    st[3] = _mm512_add_epi64(st[3], d); st[4] = _mm512_add_epi64(st[4], e); st[5] = _mm512_add_epi64(st[5], f);
    st[6] = _mm512_add_epi64(st[6], g); st[7] = _mm512_add_epi64(st[7], hh);
  }
  for (int i = 0; i < 8; i++) _mm512_store_si512((void*)tr[i], st[i]);
  for (int l = 0; l < 8; l++)
    if (mask & (1u << l))
      for (int i = 0; i < 8; i++) h[l][i] = tr[i][l];
}

bool MbSha512::simd_available() {
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
  return ok;
}

void sha512_blocks_x8(uint64_t h[8][8], const uint8_t* const ptr[8], size_t nblk, unsigned mask) {
  if (MbSha512::simd_available()) {
    blocks_x8_avx512(h, ptr, nblk, mask);
    return;
  }
  for (int l = 0; l < 8; l++)
    if (mask & (1u << l)) sha512_blocks_x1(h[l], ptr[l], nblk);
}

MbSha512::MbSha512() {
  for (auto& l : lanes_) memcpy(l.h, IV512, sizeof IV512);
  th_ = std::thread([this] { run(); });
}

MbSha512::~MbSha512() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  if (th_.joinable()) th_.join();
}

int MbSha512::acquire() {
  std::lock_guard<std::mutex> lk(mu_);
  for (int i = 0; i < LANES; i++)
    if (!lanes_[i].used) {
      Lane& l = lanes_[i];
      l.used = true;
      memcpy(l.h, IV512, sizeof IV512);
      l.buflen = 0;
      l.total = 0;
      return i;
    }
  return -1;
}

void MbSha512::release(int lane) {
  sync(lane);
  std::lock_guard<std::mutex> lk(mu_);
  lanes_[lane].used = false;
}

void MbSha512::reset(int lane) {
  sync(lane);
  std::lock_guard<std::mutex> lk(mu_);
  Lane& l = lanes_[lane];
  memcpy(l.h, IV512, sizeof IV512);
  l.buflen = 0;
  l.total = 0;
}

void MbSha512::update(int lane, const uint8_t* p, size_t n) {
  if (!n) return;
  {
    std::lock_guard<std::mutex> lk(mu_);
    lanes_[lane].q.push_back(Seg{p, n});
    lanes_[lane].total += n;
  }
  cv_work_.notify_one();
}

void MbSha512::sync(int lane) {
  std::unique_lock<std::mutex> lk(mu_);
  cv_idle_.wait(lk, [&] { return lanes_[lane].q.empty() && !lanes_[lane].busy; });
}

void MbSha512::digest(int lane, uint8_t out[64]) {
  sync(lane);
  uint64_t h[8];
  uint8_t tail[256];
  size_t bl;
  uint64_t total;
  {
    std::lock_guard<std::mutex> lk(mu_);
    Lane& l = lanes_[lane];
    memcpy(h, l.h, sizeof h);
    bl = l.buflen;
    memcpy(tail, l.buf, bl);
    total = l.total;
  }
  size_t padded = (bl + 1 + 16 <= 128) ? 128 : 256;
  memset(tail + bl, 0, padded - bl);
  tail[bl] = 0x80;
  uint64_t bits = total * 8;                            // < 2^64 bits: the high length word stays 0
  for (int i = 0; i < 8; i++) tail[padded - 1 - i] = (uint8_t)(bits >> (8 * i));
  sha512_blocks_x1(h, tail, padded / 128);
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(h[i] >> (56 - 8 * k));
}

// scalar absorb of a short piece (buffer top-up and tails); hashing thread only, lane not shared meanwhile
void MbSha512::absorb(Lane& l, const uint8_t* p, size_t n) {
  while (n) {
    size_t take = std::min(n, (size_t)128 - l.buflen);
    memcpy(l.buf + l.buflen, p, take);
    l.buflen += take;
    p += take;
    n -= take;
    if (l.buflen == 128) {
      sha512_blocks_x1(l.h, l.buf, 1);
      l.buflen = 0;
    }
  }
}

void MbSha512::run() {
  static const uint8_t dummy[128] = {0};
  (void)dummy;
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_work_.wait(lk, [&] {
      if (stop_) return true;
      for (auto& l : lanes_) if (!l.q.empty()) return true;
      return false;
    });
    bool any = false;
    for (auto& l : lanes_) any |= !l.q.empty();
    if (!any) {
      if (stop_) return;
      continue;
    }
    // Take the head segment of every lane that has one.  Bring each to a block boundary with the scalar path,
    // then run the common number of whole blocks in lockstep; what is left of a segment goes back to the front.
    Seg seg[LANES];
    unsigned mask = 0;
    for (int i = 0; i < LANES; i++)
      if (!lanes_[i].q.empty()) {
        seg[i] = lanes_[i].q.front();
        lanes_[i].q.pop_front();
        lanes_[i].busy = true;
        mask |= 1u << i;
      }
    lk.unlock();
    size_t common = SIZE_MAX;
    for (int i = 0; i < LANES; i++)
      if (mask & (1u << i)) {
        Lane& l = lanes_[i];
        if (l.buflen) {                                   // top the partial block up first
          size_t take = std::min(seg[i].n, (size_t)128 - l.buflen);
          absorb(l, seg[i].p, take);
          seg[i].p += take;
          seg[i].n -= take;
        }
        if (seg[i].n >= 128) common = std::min(common, seg[i].n / 128);
      }
    unsigned vmask = 0;
    const uint8_t* ptr[LANES];
    for (int i = 0; i < LANES; i++) {
      ptr[i] = dummy;
      if ((mask & (1u << i)) && seg[i].n >= 128) {
        vmask |= 1u << i;
        ptr[i] = seg[i].p;
      }
    }
    if (vmask) {
      common = std::min(common, (size_t)1 << 13);        // <= 1 MiB per lane per pass: lanes that arrive late join soon
      if (__builtin_popcount(vmask) == 1 || !simd_available()) {
        for (int i = 0; i < LANES; i++)
          if (vmask & (1u << i)) sha512_blocks_x1(lanes_[i].h, ptr[i], common);
      } else {
        uint64_t hs[8][8];
        int first = __builtin_ctz(vmask);
        for (int i = 0; i < LANES; i++) {
          memcpy(hs[i], lanes_[i].h, sizeof hs[i]);
          if (!(vmask & (1u << i))) ptr[i] = ptr[first];   // idle lanes re-read an active lane's data; result dropped
        }
        sha512_blocks_x8(hs, ptr, common, vmask);
        for (int i = 0; i < LANES; i++)
          if (vmask & (1u << i)) memcpy(lanes_[i].h, hs[i], sizeof hs[i]);
      }
      for (int i = 0; i < LANES; i++)
        if (vmask & (1u << i)) {
          seg[i].p += common * 128;
          seg[i].n -= common * 128;
        }
    }
    for (int i = 0; i < LANES; i++)
      if ((mask & (1u << i)) && seg[i].n < 128 && seg[i].n) {   // tail shorter than a block: buffer it
        absorb(lanes_[i], seg[i].p, seg[i].n);
        seg[i].n = 0;
      }
    lk.lock();
    for (int i = 0; i < LANES; i++)
      if (mask & (1u << i)) {
        if (seg[i].n) lanes_[i].q.push_front(seg[i]);
        lanes_[i].busy = false;
      }
    cv_idle_.notify_all();
  }
}

}  // namespace avrf

// SHA-512 for the Fiat-Shamir transcripts (replaces the `sha2` 0.10 crate behind
// `HashTranscript<Sha512>`, src/utils/transcript.rs:209-289).
//
// A transcript is a byte stream absorbed into a SHA-512 state; squeezing finalises the
// hash into a 64-byte seed and reads SHA512(seed || LE64(counter)) blocks
// (transcript.rs:230-273).  One thread owns one transcript: the 80 rounds of a compression
// are a serial dependency chain, so spreading one hash over a warp would idle 31 lanes.
#pragma once
#include <stdint.h>
#include "fp.cuh"

namespace avrf {

#define AVRF_SHA512_K_LIST \
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL, \
    0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, \
    0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, \
    0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL, \
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL, \
    0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, \
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, \
    0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL, \
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL, \
    0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, \
    0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, \
    0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL, \
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL, \
    0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, \
    0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, \
    0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL
static const uint64_t SHA512_K_HOST[80] = {AVRF_SHA512_K_LIST};
#ifdef __CUDACC__
static __constant__ uint64_t SHA512_K_DEV[80] = {AVRF_SHA512_K_LIST};
#endif

#ifdef __CUDA_ARCH__
#define AVRF_SHA_K(i) SHA512_K_DEV[i]
#else
#define AVRF_SHA_K(i) SHA512_K_HOST[i]
#endif

// 64-bit rotate.  On the device: two funnel shifts (SHF.R.W) on the 32-bit halves; the plain C expression compiles to two
// shift pairs joined by ORs (8 more instructions per SHA-512 round, a quarter of the compression).
AVRF_HD uint64_t rotr64(uint64_t x, int n) {
#ifdef __CUDA_ARCH__
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32), a, b;
  if (n < 32) {
    a = __funnelshift_r(lo, hi, n);
    b = __funnelshift_r(hi, lo, n);
  } else {
    a = __funnelshift_r(hi, lo, n - 32);
    b = __funnelshift_r(lo, hi, n - 32);
  }
  return ((uint64_t)b << 32) | a;
#else
  return (x >> n) | (x << (64 - n));
#endif
}

AVRF_HD uint64_t bswap64(uint64_t x) {
  x = ((x & 0x00ff00ff00ff00ffULL) << 8) | ((x >> 8) & 0x00ff00ff00ff00ffULL);
  x = ((x & 0x0000ffff0000ffffULL) << 16) | ((x >> 16) & 0x0000ffff0000ffffULL);
  return (x << 32) | (x >> 32);
}

struct Sha512 {
  uint64_t h[8];
  uint64_t w[16];   // current block, big-endian words
  uint32_t len;     // total bytes absorbed
};

AVRF_HD void sha512_init(Sha512& c) {
  c.h[0] = 0x6a09e667f3bcc908ULL; c.h[1] = 0xbb67ae8584caa73bULL; c.h[2] = 0x3c6ef372fe94f82bULL;
  c.h[3] = 0xa54ff53a5f1d36f1ULL; c.h[4] = 0x510e527fade682d1ULL; c.h[5] = 0x9b05688c2b3e6c1fULL;
  c.h[6] = 0x1f83d9abfb41bd6bULL; c.h[7] = 0x5be0cd19137e2179ULL;
#pragma unroll
  for (int i = 0; i < 16; i++) c.w[i] = 0;
  c.len = 0;
}

#define AVRF_SHA_ROUND(i, kw)                                                    \
  {                                                                              \
    uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);                 \
    uint64_t ch = (e & f) ^ (~e & g);                                            \
    uint64_t t1 = hh + S1 + ch + (kw);                                           \
    uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);                 \
    uint64_t mj = (a & b) ^ (a & c) ^ (b & c);                                   \
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;     \
  }

// One compression of block w[0..15] (destroyed) into h.
AVRF_HD void sha512_compress_inl(uint64_t* h, uint64_t* w) {
  uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 16; i++) AVRF_SHA_ROUND(i, AVRF_SHA_K(i) + w[i]);
#pragma unroll 1
  for (int r = 16; r < 80; r += 16) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      uint64_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      uint64_t s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
      uint64_t s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
      w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
      AVRF_SHA_ROUND(i, AVRF_SHA_K(r + i) + w[i]);
    }
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// Compress the context's current block and clear it (single pointer argument, see fp.cuh).
AVRF_HD_CALL void sha512_block(Sha512& c) {
  sha512_compress_inl(c.h, c.w);
#pragma unroll
  for (int i = 0; i < 16; i++) c.w[i] = 0;
}

AVRF_HD void sha512_put_byte(Sha512& c, uint32_t byte) {
  uint32_t pos = c.len & 127;
  c.w[pos >> 3] |= (uint64_t)byte << (56 - 8 * (pos & 7));
  c.len++;
  if ((c.len & 127) == 0) sha512_block(c);
}

AVRF_HD void sha512_update(Sha512& c, const uint8_t* p, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) sha512_put_byte(c, p[i]);
}

// Absorb a little-endian 64-bit value (8 bytes, least significant first).
AVRF_HD void sha512_put_le64(Sha512& c, uint64_t v) {
  uint32_t pos = c.len & 127;
  if ((pos & 7) == 0) {                 // aligned: one word store
    c.w[pos >> 3] = bswap64(v);
    c.len += 8;
    if ((c.len & 127) == 0) sha512_block(c);
  } else {                              // straddles two block words: two shifted stores instead of eight byte RMWs
    uint64_t x = bswap64(v);
    uint32_t sh = 8 * (pos & 7), idx = pos >> 3;
    c.w[idx] |= x >> sh;
    c.len += 8;
    if (idx == 15) sha512_block(c);     // the block filled up with the first part
    c.w[(idx + 1) & 15] = x << (64 - sh);
  }
}

// Absorb eight little-endian 32-bit words (a 32-byte encoded point or scalar).
AVRF_HD void sha512_put_words(Sha512& c, const uint32_t* w8) {
#pragma unroll 1
  for (int i = 0; i < 8; i += 2) sha512_put_le64(c, (uint64_t)w8[i] | ((uint64_t)w8[i + 1] << 32));
}

struct Digest { uint64_t w[8]; };   // 8 big-endian words (w[0] holds digest bytes 0..7)

// Finalise; digest written as 8 big-endian words (out[0] holds digest bytes 0..7).
AVRF_HD void sha512_final(Sha512& c, uint64_t* out) {
  uint64_t bits = (uint64_t)c.len * 8;
  sha512_put_byte(c, 0x80);
  if ((c.len & 127) > 112) sha512_block(c);
  c.w[15] = bits;
  sha512_block(c);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = c.h[i];
}

// Counter-mode squeeze block: SHA512(seed || LE64(ctr))  (transcript.rs:255-273).
AVRF_HD_CALL Digest sha512_xof_block_v(Digest seed, uint64_t ctr) {
  uint64_t h[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
  uint64_t w[16];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = seed.w[i];
  w[8] = bswap64(ctr);
  w[9] = 0x8000000000000000ULL;
#pragma unroll
  for (int i = 10; i < 15; i++) w[i] = 0;
  w[15] = 72 * 8;
  sha512_compress_inl(h, w);
  Digest out;
#pragma unroll
  for (int i = 0; i < 8; i++) out.w[i] = h[i];
  return out;
}
AVRF_HD void sha512_xof_block(uint64_t* out, const uint64_t* seed, uint64_t ctr) {
  Digest s;
#pragma unroll
  for (int i = 0; i < 8; i++) s.w[i] = seed[i];
  Digest o = sha512_xof_block_v(s, ctr);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = o.w[i];
}

// 16 bytes at byte offset `off` (multiple of 16) of a digest -> little-endian 128-bit
// integer as 4 words (challenge_scalar, common.rs:72-76: from_le_bytes_mod_order of 16 B).
AVRF_HD void digest_le128(uint32_t* out4, const uint64_t* digest, uint32_t off) {
  uint64_t lo = bswap64(digest[off >> 3]);        // bytes off..off+7, LE
  uint64_t hi = bswap64(digest[(off >> 3) + 1]);
  out4[0] = (uint32_t)lo; out4[1] = (uint32_t)(lo >> 32);
  out4[2] = (uint32_t)hi; out4[3] = (uint32_t)(hi >> 32);
}

}  // namespace avrf

// Carry-free Montgomery multiplication experiment: 9 limbs of 29 bits, R = 2^261, product scanning
// into 64-bit column accumulators with plain IMAD.WIDE.U32 (no carry predicates: 2.33 instead of 4.06
// issue cycles per wide MAC on sm_100a).
//
// Status: measured by avrf_microbench(kind 5) and checked by tests/hostemu; not yet used by the MSM.
#pragma once
#include "fp.cuh"

namespace avrf {

struct Fe29 { uint32_t v[9]; };
constexpr uint32_t M29 = (1u << 29) - 1;

// 8x32 canonical-form limbs -> 9x29 limbs (pure re-slicing of the same integer)
AVRF_HD void to29(Fe29& r, const Fe& a) {
#pragma unroll
  for (int k = 0; k < 9; k++) {
    int bit = 29 * k, w = bit >> 5, sh = bit & 31;
    uint64_t lo = a.v[w];
    uint64_t hi = (w + 1 < 8) ? a.v[w + 1] : 0;
    r.v[k] = (uint32_t)(((lo | (hi << 32)) >> sh) & M29);
  }
}

// 9x29 limbs (each < 2^29, value < 2^256) -> 8x32
AVRF_HD void from29(Fe& r, const Fe29& a) {
  uint32_t out[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 9; k++) {
    int bit = 29 * k, w = bit >> 5, sh = bit & 31;
    uint64_t v = (uint64_t)a.v[k] << sh;
    out[w] |= (uint32_t)v;
    if (w + 1 < 9) out[w + 1] |= (uint32_t)(v >> 32);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = out[i];
}

struct Field29Consts {
  uint32_t p[9];     // modulus in 29-bit limbs
  uint32_t n0;       // -p^{-1} mod 2^29
};

// c += a * b as ONE IMAD.WIDE.U32 with its 64-bit addend (kept as an accumulate chain: letting the
// compiler re-associate the column sums into trees turns half of them into IADD3 pairs on the ALU pipe)
AVRF_HD void mac29(uint64_t& c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
#else
  c += (uint64_t)a * b;
#endif
}

// r = a * b * 2^-261 mod p, limbs of a and b < 2^30, result limbs < 2^29 and value < p + 2^252 (lazy).
template <bool P0_IS_ONE>
AVRF_HD void mont_mul29(Fe29& r, const Fe29& a, const Fe29& b, const Field29Consts& F) {
  uint64_t c[18];
#pragma unroll
  for (int k = 0; k < 18; k++) c[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = 0; j < 9; j++) mac29(c[i + j], a.v[i], b.v[j]);
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t m;
    if (P0_IS_ONE) {
      m = (0u - (uint32_t)c[i]) & M29;                 // n0 = 2^29 - 1
      c[i] += m;                                       // m * p[0]
    } else {
      m = ((uint32_t)c[i] * F.n0) & M29;
      mac29(c[i], m, F.p[0]);
    }
#pragma unroll
    for (int j = 1; j < 9; j++) mac29(c[i + j], m, F.p[j]);
    c[i + 1] += c[i] >> 29;                            // low 29 bits of c[i] are zero now
  }
#pragma unroll
  for (int k = 9; k < 17; k++) {
    r.v[k - 9] = (uint32_t)c[k] & M29;
    c[k + 1] += c[k] >> 29;
  }
  r.v[8] = (uint32_t)c[17];
}

}  // namespace avrf

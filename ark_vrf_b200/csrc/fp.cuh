// 256-bit prime-field arithmetic in 8x32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces (for the GPU path) the arkworks `ark-ff` 0.6 Montgomery backend that the
// reference reaches through `BaseField<S>` / `ScalarField<S>` (src/lib.rs:113-128).  The
// in-memory limbs are identical to arkworks' 4x64-bit little-endian Montgomery limbs, so
// field elements cross the C ABI without conversion.
//
// All moduli in scope are < 2^255 (one spare bit), which the lazy carry handling in
// mont_mul relies on.  The multiplier is written for the sm_100a IMAD pipe: every
// 32x32->64 product is a mad.lo/mad.hi pair that ptxas fuses into one IMAD.WIDE.U32 with a
// carry predicate; partial products of even and odd limbs live in two accumulators so
// that each row is one unbroken carry chain (no per-limb carry fix-ups).
//
// Every function is __host__ __device__: the host variants emulate the PTX carry flag in
// plain C so that tests/ can exercise the exact limb algorithms without a GPU
// (tests/hostemu).  The product (libavrf_gpu.so entry points) only ever runs the device
// variants; there is no CPU fallback.
#pragma once
#include <stdint.h>
#include "constants_gen.h"

#ifdef __CUDACC__
#define AVRF_HD __host__ __device__ __forceinline__
#define AVRF_D __device__ __forceinline__
// Out-of-line variants: everything outside the bucket-accumulation hot loop calls these, which
// keeps code size (and ptxas time) bounded; the call overhead is irrelevant there.
#define AVRF_HD_CALL __host__ __device__ __noinline__
#else
#define AVRF_HD inline
#define AVRF_D inline
#define AVRF_HD_CALL inline
#endif

namespace avrf {

// Field ids (index into the constant table).
enum : int { FQ_BAND = 0, FQ_ED = 1, FQ_BJJ = 2, FR_BAND = 3, FR_ED = 4, FR_BJJ = 5, N_FIELDS = 6 };

struct FieldConsts {
  uint32_t p[8];      // modulus
  uint32_t r1[8];     // R mod p         (Montgomery one)
  uint32_t r2[8];     // R^2 mod p
  uint32_t pm2[8];    // p - 2           (inversion exponent)
  uint32_t phalf[8];  // (p - 1) / 2     (sign rule x > p - x  <=>  x > (p-1)/2 ; Legendre exponent)
  uint32_t n0;        // -p^{-1} mod 2^32
  uint32_t pad[7];
};

// One copy per translation unit (the device library is a single TU, so no -rdc needed).
static const FieldConsts FC_HOST[N_FIELDS] = AVRF_FIELD_CONSTS_INIT;
#ifdef __CUDACC__
static __constant__ FieldConsts FC_DEV[N_FIELDS] = AVRF_FIELD_CONSTS_INIT;
#endif

#ifdef __CUDA_ARCH__
#define AVRF_FC(F) FC_DEV[F]
#else
#define AVRF_FC(F) FC_HOST[F]
#endif

struct alignas(16) Fe {
  uint32_t v[8];
};

// ---------------------------------------------------------------------------------------
// Carry-flag primitives: PTX on the device, emulated flag on the host.
// ---------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define AVRF_ASM2(name, ins)                                                        \
  AVRF_D uint32_t name(uint32_t a, uint32_t b) {                                    \
    uint32_t r;                                                                     \
    asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                    \
    return r;                                                                       \
  }
#define AVRF_ASM3(name, ins)                                                        \
  AVRF_D uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                        \
    uint32_t r;                                                                     \
    asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));        \
    return r;                                                                       \
  }
AVRF_ASM2(add_cc, "add.cc.u32")
AVRF_ASM2(addc_cc, "addc.cc.u32")
AVRF_ASM2(addc, "addc.u32")
AVRF_ASM2(sub_cc, "sub.cc.u32")
AVRF_ASM2(subc_cc, "subc.cc.u32")
AVRF_ASM2(subc, "subc.u32")
AVRF_ASM3(mad_lo_cc, "mad.lo.cc.u32")
AVRF_ASM3(madc_lo_cc, "madc.lo.cc.u32")
AVRF_ASM3(madc_hi_cc, "madc.hi.cc.u32")
AVRF_ASM3(madc_hi, "madc.hi.u32")
AVRF_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
AVRF_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
static thread_local uint32_t g_cf = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cf; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cf; g_cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(a * b) + c + g_cf; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + g_cf; g_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + g_cf; }
#endif

// ---------------------------------------------------------------------------------------
// Montgomery multiplication
// ---------------------------------------------------------------------------------------

// acc[0..7] = a[0,2,4,6] * b   (four independent 64-bit products)
AVRF_HD void mul_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    uint64_t t = (uint64_t)a[j] * b;      // one IMAD.WIDE.U32
    acc[j] = (uint32_t)t;
    acc[j + 1] = (uint32_t)(t >> 32);
  }
}

// acc[0..7] += a[0,2,4,6] * b  as one carry chain; the carry out is left in CC.
AVRF_HD void mad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
  acc[0] = mad_lo_cc(a[0], b, acc[0]);
  acc[1] = madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    acc[j] = madc_lo_cc(a[j], b, acc[j]);
    acc[j + 1] = madc_hi_cc(a[j], b, acc[j + 1]);
  }
}

// Same with carry-in from CC, reading the accumulator two limbs further up
// (acc[j] = a*b + acc[j+2]): the two-limb right shift that follows two reduction steps
// costs no instruction.  The top pair is formed from the product and the carry alone.
AVRF_HD void mad_row_rshift(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    acc[j] = madc_lo_cc(a[j], b, acc[j + 2]);
    acc[j + 1] = madc_hi_cc(a[j], b, acc[j + 3]);
  }
  acc[6] = madc_lo_cc(a[6], b, 0);
  acc[7] = madc_hi(a[6], b, 0);
}

// One operand-scanning step: (even, odd) += a * bi, then one Montgomery reduction limb.
// Value convention: V = sum even[k] B^k + sum odd[k] B^(k+1).  On entry (unless FIRST)
// the caller has swapped the roles of the two accumulators, which is the division by B.
template <int F, bool FIRST>
AVRF_HD void mad_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
  if (FIRST) {
    mul_row(odd, a + 1, bi);
    mul_row(even, a, bi);
  } else {
    even[0] = add_cc(even[0], odd[1]);
    mad_row_rshift(odd, a + 1, bi);
    mad_row(even, a, bi);
    odd[7] = addc(odd[7], 0);
  }
  if (F == FQ_BAND) {
    // BLS12-381 Fr: p[0] = 1, p[1] = 2^32 - 1, so n0 = 2^32 - 1 and three of the seventeen
    // multiplier issues of a reduction step turn into adds on the ALU pipe:
    //   mi = -even[0];   mi * p[0] = mi;   mi * p[1] = (mi << 32) - mi = (mi - [mi != 0]) : (-mi)
    // mi = -e0, written as (e0 ^ n0) + 1 with n0 = 2^32-1 read from the constant bank: a literal
    // negation would be folded into the multiplies as an operand modifier, which IMAD.WIDE does
    // not have, and ptxas then emits unfused IMAD.X / IMAD.HI.X pairs (twice the issues).
    uint32_t e0 = even[0];
    uint32_t mi = (e0 ^ AVRF_FC(F).n0) + 1u;
    uint32_t nz = e0 != 0u ? 1u : 0u;
    uint32_t hi = mi - nz;               // mi * (2^32 - 1) = hi : e0
    odd[0] = add_cc(odd[0], e0);
    odd[1] = addc_cc(odd[1], hi);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      odd[j] = madc_lo_cc(AVRF_FC(F).p[j + 1], mi, odd[j]);
      odd[j + 1] = madc_hi_cc(AVRF_FC(F).p[j + 1], mi, odd[j + 1]);
    }
    even[0] = 0;                         // e0 + mi = 0 mod 2^32, carry = (e0 != 0)
    even[1] = add_cc(even[1], nz);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      even[j] = madc_lo_cc(AVRF_FC(F).p[j], mi, even[j]);
      even[j + 1] = madc_hi_cc(AVRF_FC(F).p[j], mi, even[j + 1]);
    }
    odd[7] = addc(odd[7], 0);
  } else if (F == FQ_ED) {
    // p = 2^255 - 19:  V += m*p  is  V += (m << 255) - 19*m  with  m = t0 / 19 mod 2^32 (n0 = 19^-1).
    // One plain IMAD.WIDE (19*m) and ALU adds/borrows replace the sixteen carry-chained wide MACs of a
    // generic reduction step.  Value limbs: even[k] at k, odd[k] at k+1; odd[7] (limb 8) is the top limb
    // and absorbs the carry of limb 7 and the borrow of the subtraction (the total stays in [0, B^9)).
    uint32_t t0 = even[0];
    uint32_t mi = mul_lo(t0, AVRF_FC(F).n0);
    uint64_t q = (uint64_t)mi * 19u;                   // low word equals t0 by construction
    uint32_t qhi = (uint32_t)(q >> 32);
    even[7] = add_cc(even[7], mi << 31);               // + m * 2^255 : limbs 7 and 8
    odd[7] = addc(odd[7], mi >> 1);
    even[0] = 0;                                       // t0 - lo(19 m) = 0, no borrow
    even[1] = sub_cc(even[1], qhi);
#pragma unroll
    for (int j = 2; j < 8; j++) even[j] = subc_cc(even[j], 0);
    odd[7] = subc(odd[7], 0);
  } else {
    uint32_t mi = mul_lo(even[0], AVRF_FC(F).n0);
    mad_row(odd, AVRF_FC(F).p + 1, mi);  // cannot carry out: odd*B <= V < B^9
    mad_row(even, AVRF_FC(F).p, mi);
    odd[7] = addc(odd[7], 0);
  }
}

// t = a >= p ? a - p : a      (a < 2p)
template <int F>
AVRF_HD void cond_sub_p(uint32_t* r, const uint32_t* a) {
  uint32_t t[8];
  t[0] = sub_cc(a[0], AVRF_FC(F).p[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = subc_cc(a[i], AVRF_FC(F).p[i]);
  uint32_t borrow = subc(0, 0);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = borrow ? a[i] : t[i];
}

// r = a * b * R^-1 mod p, fully reduced to [0, p).  a, b in [0, p).
// b may be any 256-bit value (it is only scanned limb by limb); a must be below p: the running value is bounded by a + p.
template <int F>
AVRF_HD void mont_mul(Fe& r, const Fe& a, const Fe& b) {
  uint32_t even[8], odd[8];
  mad_redc<F, true>(even, odd, a.v, b.v[0]);
  mad_redc<F, false>(odd, even, a.v, b.v[1]);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    mad_redc<F, false>(even, odd, a.v, b.v[i]);
    mad_redc<F, false>(odd, even, a.v, b.v[i + 1]);
  }
  // final division by B: result limb k = even[k] + odd[k+1]
  even[0] = add_cc(even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) even[i] = addc_cc(even[i], odd[i + 1]);
  even[7] = addc(even[7], 0);
  cond_sub_p<F>(r.v, even);
}

template <int F>
AVRF_HD void mont_sqr(Fe& r, const Fe& a) { mont_mul<F>(r, a, a); }

// Out-of-line multiply / square.  Out-of-line helpers take and return VALUES: nvcc 12.9 was
// seen to miscompile a by-reference noinline helper whose output aliased an input (the
// caller's stack slots of two live locals were merged), so no helper takes pointers to
// caller locals.
template <int F>
AVRF_HD_CALL Fe mont_mul_v(Fe a, Fe b) {
  Fe r;
  mont_mul<F>(r, a, b);
  return r;
}
template <int F>
AVRF_HD void mont_mul_c(Fe& r, const Fe& a, const Fe& b) { r = mont_mul_v<F>(a, b); }
template <int F>
AVRF_HD void mont_sqr_c(Fe& r, const Fe& a) { mont_mul_c<F>(r, a, a); }

// ---------------------------------------------------------------------------------------
// Lazy reduction: wide (unreduced) products combined as plain 512-bit integers, ONE Montgomery reduction per sum.
// A sum of products  sum_i a_i b_i  costs one reduction (48 of the 112 multiplier issues of a multiplication on
// BLS12-381 Fr) instead of one per product; the mixed addition of the bucket accumulation uses it for the pair
// (X1 y2 + Y1 x2,  Y1 y2 - a X1 x2).
// ---------------------------------------------------------------------------------------

// t[0..15] = a * b for any 256-bit a, b.  Operand scanning with the same even/odd accumulators as mont_mul: the
// products of row i land in limbs i .. i+8; the chain of the odd multiplicand limbs creates the two new top limbs
// and takes the carry of the other chain (the total a * (b mod B^(i+1)) < B^(i+9), so that limb cannot overflow).
AVRF_HD void mul_wide(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  uint32_t even[16], odd[16];
  mul_row(even, a, b[0]);
  mul_row(odd, a + 1, b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    uint32_t* P = (i & 1) ? odd + (i - 1) : even + i;       // a[0,2,4,6] * b[i]: limbs i .. i+7, all present
    uint32_t* Q = (i & 1) ? even + (i + 1) : odd + i;       // a[1,3,5,7] * b[i]: limbs i+1 .. i+8, the top pair new
    Q[0] = mad_lo_cc(a[1], b[i], Q[0]);
    Q[1] = madc_hi_cc(a[1], b[i], Q[1]);
    Q[2] = madc_lo_cc(a[3], b[i], Q[2]);
    Q[3] = madc_hi_cc(a[3], b[i], Q[3]);
    Q[4] = madc_lo_cc(a[5], b[i], Q[4]);
    Q[5] = madc_hi_cc(a[5], b[i], Q[5]);
    Q[6] = madc_lo_cc(a[7], b[i], 0);
    Q[7] = madc_hi(a[7], b[i], 0);
    mad_row(P, a, b[i]);
    Q[7] = addc(Q[7], 0);
  }
  // limb k = even[k] + odd[k-1]  (odd reaches index 13)
  t[0] = even[0];
  t[1] = add_cc(even[1], odd[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) t[k] = addc_cc(even[k], odd[k - 1]);
  t[15] = addc(even[15], 0);
}

// 512-bit integer helpers (no reduction)
AVRF_HD void add16(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 15; i++) r[i] = addc_cc(a[i], b[i]);
  r[15] = addc(a[15], b[15]);
}
AVRF_HD void sub16(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 15; i++) r[i] = subc_cc(a[i], b[i]);
  r[15] = subc(a[15], b[15]);
}
AVRF_HD void shl2_16(uint32_t* r, const uint32_t* a) {
#pragma unroll
  for (int i = 15; i > 0; i--) r[i] = (a[i] << 2) | (a[i - 1] >> 30);
  r[0] = a[0] << 2;
}
// 256-bit sum without reduction (the caller knows it fits: both operands < p < 2^255)
AVRF_HD void add8_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r[i] = addc_cc(a[i], b[i]);
  r[7] = addc(a[7], b[7]);
}
// t -= p * 2^256 when the top half of t is >= p  (t < 2p * 2^256 on entry, < p * 2^256 on exit)
template <int F>
AVRF_HD void cond_sub_p_top(uint32_t* t) { cond_sub_p<F>(t + 8, t + 8); }

// One limb of the Montgomery reduction of a wide value held in the (even, odd) accumulators of mont_mul, without a
// product row.  `ev` enters as the odd accumulator of the previous step and leaves as the even one of this step; `od`
// enters as the previous even accumulator (its limb 0 cancelled) and leaves as the odd one, two limbs down - the
// shift costs nothing because every limb passes through a multiplier issue or an add anyway.  `tnext` is the next
// limb of the wide value, taken in at the top.
template <int F, bool FIRST>
AVRF_HD void redc_step(uint32_t* ev, uint32_t* od, uint32_t tnext) {
  const uint32_t* p = AVRF_FC(F).p;
  if (F == FQ_BAND) {                    // p[0] = 1, p[1] = 2^32 - 1  (see mad_redc)
    uint32_t e0 = FIRST ? ev[0] : add_cc(ev[0], od[1]);
    uint32_t mi = (e0 ^ AVRF_FC(F).n0) + 1u;
    uint32_t nz = e0 != 0u ? 1u : 0u;
    uint32_t hi = mi - nz;
    if (FIRST) {
      od[0] = e0;
      od[1] = hi;
#pragma unroll
      for (int j = 2; j < 8; j += 2) {
        uint64_t t = (uint64_t)p[j + 1] * mi;
        od[j] = (uint32_t)t;
        od[j + 1] = (uint32_t)(t >> 32);
      }
    } else {
      od[0] = addc_cc(od[2], e0);
      od[1] = addc_cc(od[3], hi);
      od[2] = madc_lo_cc(p[3], mi, od[4]);
      od[3] = madc_hi_cc(p[3], mi, od[5]);
      od[4] = madc_lo_cc(p[5], mi, od[6]);
      od[5] = madc_hi_cc(p[5], mi, od[7]);
      od[6] = madc_lo_cc(p[7], mi, tnext);
      od[7] = madc_hi(p[7], mi, 0);
    }
    ev[0] = 0;
    ev[1] = add_cc(ev[1], nz);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      ev[j] = madc_lo_cc(p[j], mi, ev[j]);
      ev[j + 1] = madc_hi_cc(p[j], mi, ev[j + 1]);
    }
    od[7] = addc(od[7], 0);
  } else {
    uint32_t e0 = FIRST ? ev[0] : add_cc(ev[0], od[1]);
    uint32_t mi = mul_lo(e0, AVRF_FC(F).n0);
    if (FIRST) {
      mul_row(od, p + 1, mi);
    } else {
      od[0] = madc_lo_cc(p[1], mi, od[2]);
      od[1] = madc_hi_cc(p[1], mi, od[3]);
      od[2] = madc_lo_cc(p[3], mi, od[4]);
      od[3] = madc_hi_cc(p[3], mi, od[5]);
      od[4] = madc_lo_cc(p[5], mi, od[6]);
      od[5] = madc_hi_cc(p[5], mi, od[7]);
      od[6] = madc_lo_cc(p[7], mi, tnext);
      od[7] = madc_hi(p[7], mi, 0);
    }
    ev[0] = mad_lo_cc(p[0], mi, e0);
    ev[1] = madc_hi_cc(p[0], mi, ev[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      ev[j] = madc_lo_cc(p[j], mi, ev[j]);
      ev[j + 1] = madc_hi_cc(p[j], mi, ev[j + 1]);
    }
    od[7] = addc(od[7], 0);
  }
}

// r = t * R^-1 mod p, fully reduced, for a 16-limb t < p * 2^256.  The high limbs of t ride in on the addends of the
// top multiplier issue of every step, so the reduction of a wide value costs exactly the reduction half of mont_mul.
template <int F, bool NO_FINAL_SUB = false>
AVRF_HD void redc_wide(Fe& r, const uint32_t* t) {
  uint32_t x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = t[i];
  redc_step<F, true>(x, y, 0);
  redc_step<F, false>(y, x, t[8]);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    redc_step<F, false>(x, y, t[7 + i]);
    redc_step<F, false>(y, x, t[8 + i]);
  }
  // after the last step y is the even accumulator (limb 0 cancelled), x the odd one: limb k = y[k+1] + x[k]
  x[0] = add_cc(x[0], y[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) x[i] = addc_cc(x[i], y[i + 1]);
  x[7] = addc(x[7], t[15]);
  if (NO_FINAL_SUB) {                  // the caller takes a value in [0, 2p)
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = x[i];
  } else {
    cond_sub_p<F>(r.v, x);
  }
}

// ---------------------------------------------------------------------------------------
// Additive ops (inputs and outputs in [0, p))
// ---------------------------------------------------------------------------------------
template <int F>
AVRF_HD void fe_add(Fe& r, const Fe& a, const Fe& b) {
  uint32_t s[8];
  s[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) s[i] = addc_cc(a.v[i], b.v[i]);
  s[7] = addc(a.v[7], b.v[7]);  // < 2^256 because p < 2^255
  cond_sub_p<F>(r.v, s);
}

template <int F>
AVRF_HD void fe_sub(Fe& r, const Fe& a, const Fe& b) {
  uint32_t s[8], t[8];
  s[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0, 0);
  t[0] = add_cc(s[0], AVRF_FC(F).p[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) t[i] = addc_cc(s[i], AVRF_FC(F).p[i]);
  t[7] = addc(s[7], AVRF_FC(F).p[7]);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = borrow ? t[i] : s[i];
}

AVRF_HD bool fe_is_zero(const Fe& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}

AVRF_HD bool fe_eq(const Fe& a, const Fe& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

template <int F>
AVRF_HD void fe_neg(Fe& r, const Fe& a) {
  uint32_t t[8];
  t[0] = sub_cc(AVRF_FC(F).p[0], a.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) t[i] = subc_cc(AVRF_FC(F).p[i], a.v[i]);
  t[7] = subc(AVRF_FC(F).p[7], a.v[7]);
  bool z = fe_is_zero(a);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : t[i];
}

// r = neg ? -a : a
template <int F>
AVRF_HD void fe_cneg(Fe& r, const Fe& a, bool neg) {
  Fe t;
  fe_neg<F>(t, a);
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = neg ? t.v[i] : a.v[i];
}

template <int F>
AVRF_HD void fe_dbl(Fe& r, const Fe& a) { fe_add<F>(r, a, a); }

AVRF_HD void fe_set(Fe& r, const uint32_t* w) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = w[i];
}

template <int F>
AVRF_HD void fe_one(Fe& r) { fe_set(r, AVRF_FC(F).r1); }

AVRF_HD void fe_zero(Fe& r) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
}

// a > b as 256-bit integers
AVRF_HD bool limbs_gt(const uint32_t* a, const uint32_t* b) {
  // b - a borrows  <=>  a > b
  sub_cc(b[0], a[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) subc_cc(b[i], a[i]);
  return subc(0, 0) != 0;
}

// canonical -> Montgomery, Montgomery -> canonical
template <int F>
AVRF_HD void to_mont(Fe& r, const Fe& a) {
  Fe r2;
  fe_set(r2, AVRF_FC(F).r2);
  mont_mul_c<F>(r, a, r2);
}

template <int F>
AVRF_HD void from_mont(Fe& r, const Fe& a) {
  Fe one;
  fe_zero(one);
  one.v[0] = 1;
  mont_mul_c<F>(r, a, one);
}

// Reduce an arbitrary 256-bit integer into [0, p) (at most a few subtractions since
// every modulus in scope exceeds 2^250).
template <int F>
AVRF_HD void reduce_once(Fe& r, const Fe& a) {
  Fe t = a;
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
    if (limbs_gt(AVRF_FC(F).p, t.v)) break;  // p > t
    uint32_t s[8];
    s[0] = sub_cc(t.v[0], AVRF_FC(F).p[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = subc_cc(t.v[i], AVRF_FC(F).p[i]);
    fe_set(t, s);
  }
  r = t;
}

// r = a^e, e given as 8 limbs in constant memory (public exponent).  Fixed 4-bit windows, most significant first:
// 14 multiplications for the table, then 4 squarings and at most one multiplication per window (about 335
// multiplications for a 255-bit exponent of any weight, against 255 + popcount for square-and-multiply).
template <int F>
AVRF_HD_CALL Fe fe_pow_v(Fe a, const uint32_t* e) {
  Fe tbl[16];
  fe_one<F>(tbl[0]);
  tbl[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tbl[i] = mont_mul_v<F>(tbl[i - 1], a);
  Fe acc;
  fe_one<F>(acc);
  bool started = false;
#pragma unroll 1
  for (int w = 63; w >= 0; w--) {
    uint32_t dgt = (e[w >> 3] >> (4 * (w & 7))) & 15u;
    if (started) {
      acc = mont_mul_v<F>(acc, acc);
      acc = mont_mul_v<F>(acc, acc);
      acc = mont_mul_v<F>(acc, acc);
      acc = mont_mul_v<F>(acc, acc);
      if (dgt) acc = mont_mul_v<F>(acc, tbl[dgt]);
    } else if (dgt) {
      acc = tbl[dgt];
      started = true;
    }
  }
  return acc;
}
template <int F>
AVRF_HD void fe_pow(Fe& r, const Fe& a, const uint32_t* e) { r = fe_pow_v<F>(a, e); }

template <int F>
AVRF_HD void fe_inv(Fe& r, const Fe& a) { fe_pow<F>(r, a, AVRF_FC(F).pm2); }

// Jacobi symbol (a / p) of a 256-bit value 0 <= a < p (p an odd prime: the Legendre symbol): +1, -1, or 0 for a = 0.
// Binary algorithm - trailing-zero shifts, compares, swaps and subtractions on eight limbs, nothing on the multiplier:
// ~300 iterations of ~45 instructions against the ~75 000 instructions of the exponentiation a^((p-1)/2).
//   (2 / n) = -1 iff n = 3, 5 (mod 8);   (a / n)(n / a) = -1 iff a = n = 3 (mod 4);   (a / n) = ((a - n) / n)
// The Montgomery representative has the same symbol as the value it stands for (R = 2^256 is a square).
template <int F>
AVRF_HD_CALL int fe_jacobi_v(Fe x) {
  uint32_t a[8], n[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = x.v[i]; n[i] = AVRF_FC(F).p[i]; }
  uint32_t flip = 0;
  if (fe_is_zero(x)) return 0;
#pragma unroll 1
  for (;;) {
    // a != 0: strip its trailing zeros (whole limbs first: 32 is even, no sign change)
#pragma unroll 1
    while (a[0] == 0) {
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = a[i + 1];
      a[7] = 0;
    }
    uint32_t low = a[0];
#ifdef __CUDA_ARCH__
    int tz = __ffs((int)low) - 1;
#else
    int tz = 0;
    while (((low >> tz) & 1u) == 0) tz++;
#endif
    if (tz) {
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = (a[i] >> tz) | (a[i + 1] << (32 - tz));
      a[7] >>= tz;
      uint32_t r8 = n[0] & 7u;
      if ((tz & 1) && (r8 == 3u || r8 == 5u)) flip ^= 1u;
    }
    // both odd now; keep a >= n (quadratic reciprocity when they trade places)
    if (limbs_gt(n, a)) {
      if ((a[0] & n[0] & 2u) != 0) flip ^= 1u;
#pragma unroll
      for (int i = 0; i < 8; i++) { uint32_t t = a[i]; a[i] = n[i]; n[i] = t; }
    }
    a[0] = sub_cc(a[0], n[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) a[i] = subc_cc(a[i], n[i]);
    a[7] = subc(a[7], n[7]);
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= a[i];
    if (nz == 0) break;
  }
  uint32_t rest = n[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < 8; i++) rest |= n[i];
  if (rest != 0) return 0;                             // gcd(a, p) != 1: cannot happen for a prime modulus and 0 < a < p
  return flip ? -1 : 1;
}

// Legendre symbol: returns true iff a is a non-zero square.
template <int F>
AVRF_HD bool fe_is_nonzero_square(const Fe& a) { return fe_jacobi_v<F>(a) == 1; }

// The same through Euler's criterion a^((p-1)/2) (kept for the tests: the two must agree).
template <int F>
AVRF_HD bool fe_is_nonzero_square_pow(const Fe& a) {
  Fe t, one;
  fe_pow<F>(t, a, AVRF_FC(F).phalf);
  fe_one<F>(one);
  return fe_eq(t, one);
}

}  // namespace avrf

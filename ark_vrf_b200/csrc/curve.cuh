// Twisted-Edwards group arithmetic (a*x^2 + y^2 = 1 + d*x^2*y^2) for the three suites in
// scope: Bandersnatch (a=-5), Ed25519 (a=-1), Baby-JubJub (a=1).
//
// Replaces the `ark-ec` 0.6 twisted_edwards group law used by the reference through
// `AffinePoint<S>` (src/lib.rs:113-128) - same unified extended-coordinate addition
// (add-2008-hwcd), so the same inputs give the same group elements.  Identity is (0,1)
// (src/lib.rs:460-462).
#pragma once
#include "fp.cuh"

namespace avrf {

enum : int { SUITE_BAND = 0, SUITE_ED = 1, SUITE_BJJ = 2, N_SUITES = 3 };

struct CurveConsts {
  uint32_t d[8], gx[8], gy[8], gk[8];    // Montgomery: d, generator, d*gx*gy
  uint32_t ts_exp[8];                    // (q-1)/2 with p-1 = 2^s q   (plain integer)
  uint32_t ts_root[8];                   // z^q, z a non-residue       (Montgomery)
  uint32_t d2[8];                        // 2d                         (Montgomery)
  uint32_t jk[8], k2inv[8], kk[8], zz[8];  // Elligator2: J/K, 1/K^2, K, Z (Montgomery)
  uint32_t zcw[8];                       // Z^((q-1)/2), p-1 = 2^s q (Montgomery): sqrt(Z*a) from the chain of sqrt(a)
  uint32_t g_enc[8];                     // compressed generator (A.2 encoding, LE words)
  uint32_t bx[8], by[8], bk[8];          // Pedersen blinding base B and d*bx*by (Montgomery)
  uint32_t mt_alpha[8];                  // cofactor-4 curves: a root of u^2 + A u + 1 on the Montgomery model (Montgomery form)
  uint32_t ts_s, cof_log2, sid_len, p_bits, r_bits, pad[3];
  uint8_t suite_id[32];
};

static const CurveConsts CC_HOST[N_SUITES] = AVRF_CURVE_CONSTS_INIT;
#ifdef __CUDACC__
static __constant__ CurveConsts CC_DEV[N_SUITES] = AVRF_CURVE_CONSTS_INIT;
#endif
#ifdef __CUDA_ARCH__
#define AVRF_CC(S) CC_DEV[S]
#else
#define AVRF_CC(S) CC_HOST[S]
#endif

template <int S> struct SuiteT {
  static constexpr int FQ = S;        // base-field id
  static constexpr int FR = S + 3;    // scalar-field id
};

struct Affine { Fe x, y; };           // 64 B, arkworks `Affine<P>` memory image (Montgomery)
// 96 B MSM base with the addition's per-base work hoisted:
//   Bandersnatch, Baby-JubJub: (x, y, k = d*x*y)            -> unified 8-multiplication mixed addition
//   Ed25519 (a = -1)         : (y - x, y + x, k = 2*d*x*y)  -> the 7-multiplication form (madd-2008-hwcd-3)
struct AffineK { Fe x, y, k; };
// The same as it lies in HBM: one 128-byte line per base.  k_accumulate gathers one record per addition and prefetches
// it into L2; 96-byte records straddle two 128-byte lines half of the time (measured: 180 B of DRAM traffic per
// addition instead of 96), a padded record costs exactly one line.
struct alignas(128) BaseRec { Fe x, y, k, pad; };
struct Ext { Fe x, y, z, t; };        // extended coordinates, T = XY/Z

// r = -a * v  (v Montgomery) for the curve coefficient a in {-5,-1,1}:  B - a*A below.
template <int S>
AVRF_HD void sub_a_times(Fe& r, const Fe& B, const Fe& A) {
  constexpr int FQ = SuiteT<S>::FQ;
  if (S == SUITE_BAND) {            // a = -5 : B + 5A
    Fe t;
    fe_dbl<FQ>(t, A);
    fe_dbl<FQ>(t, t);
    fe_add<FQ>(t, t, A);
    fe_add<FQ>(r, B, t);
  } else if (S == SUITE_ED) {       // a = -1 : B + A
    fe_add<FQ>(r, B, A);
  } else {                          // a = 1  : B - A
    fe_sub<FQ>(r, B, A);
  }
}

// r = a * v
template <int S>
AVRF_HD void a_times(Fe& r, const Fe& A) {
  constexpr int FQ = SuiteT<S>::FQ;
  if (S == SUITE_BAND) {
    Fe t;
    fe_dbl<FQ>(t, A);
    fe_dbl<FQ>(t, t);
    fe_add<FQ>(t, t, A);
    fe_neg<FQ>(r, t);
  } else if (S == SUITE_ED) {
    fe_neg<FQ>(r, A);
  } else {
    r = A;
  }
}

template <int S>
AVRF_HD void ext_identity(Ext& p) {
  fe_zero(p.x);
  fe_one<SuiteT<S>::FQ>(p.y);
  fe_one<SuiteT<S>::FQ>(p.z);
  fe_zero(p.t);
}

template <int S>
AVRF_HD bool ext_is_identity(const Ext& p) {
  return fe_is_zero(p.x) && fe_eq(p.y, p.z);
}

template <int S>
AVRF_HD bool affine_is_identity(const Affine& p) {
  Fe one;
  fe_one<SuiteT<S>::FQ>(one);
  return fe_is_zero(p.x) && fe_eq(p.y, one);
}

template <int S>
AVRF_HD void affine_to_ext(Ext& r, const Affine& p) {
  r.x = p.x;
  r.y = p.y;
  fe_one<SuiteT<S>::FQ>(r.z);
  mont_mul_c<SuiteT<S>::FQ>(r.t, p.x, p.y);
}

template <int S>
AVRF_HD void affine_to_k(AffineK& r, const Affine& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe d, t;
  mont_mul_c<FQ>(t, p.x, p.y);
  if (S == SUITE_ED) {
    fe_set(d, AVRF_CC(S).d2);
    mont_mul_c<FQ>(r.k, t, d);
    Fe ym, yp;
    fe_sub<FQ>(ym, p.y, p.x);
    fe_add<FQ>(yp, p.y, p.x);
    r.x = ym;
    r.y = yp;
  } else {
    fe_set(d, AVRF_CC(S).d);
    mont_mul_c<FQ>(r.k, t, d);
    r.x = p.x;
    r.y = p.y;
  }
}

// base <- neg ? -base : base   (the sign of a signed-digit entry)
template <int S>
AVRF_HD void base_cneg(AffineK& q, bool neg) {
  constexpr int FQ = SuiteT<S>::FQ;
  if (S == SUITE_ED) {                   // -(x, y): (y - x, y + x) swap, k changes sign
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t a = q.x.v[i], b = q.y.v[i];
      q.x.v[i] = neg ? b : a;
      q.y.v[i] = neg ? a : b;
    }
  } else {
    fe_cneg<FQ>(q.x, q.x, neg);
  }
  fe_cneg<FQ>(q.k, q.k, neg);
}

#ifndef AVRF_LAZY_MADD
#define AVRF_LAZY_MADD 1
#endif
// Unified mixed addition  acc += base  (base as produced by affine_to_k)
// generic: add-2008-hwcd with Z2 = 1 and d*T2 precomputed per base [8 field multiplications]
template <int S>
AVRF_HD void ext_madd(Ext& acc, const Fe& x2, const Fe& y2, const Fe& k2) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe A, B, C, E, F, G, H, t0, t1;
  if (S == SUITE_ED) {
    // a = -1, base given as (ym, yp, k) = (y2 - x2, y2 + x2, 2 d x2 y2)   [7 field multiplications]
    fe_sub<FQ>(t0, acc.y, acc.x);
    fe_add<FQ>(t1, acc.y, acc.x);
    mont_mul<FQ>(A, t0, x2);
    mont_mul<FQ>(B, t1, y2);
    mont_mul<FQ>(C, acc.t, k2);
    fe_dbl<FQ>(t0, acc.z);                 // D = 2 Z1
    fe_sub<FQ>(E, B, A);
    fe_sub<FQ>(F, t0, C);
    fe_add<FQ>(G, t0, C);
    fe_add<FQ>(H, B, A);
  } else if (S == SUITE_BAND && AVRF_LAZY_MADD) {
    // a = -5.  E = X1 y2 + Y1 x2 and H = Y1 y2 + 5 X1 x2 from three WIDE products and two reductions (lazy
    // reduction, fp.cuh): with A = X1 x2, B = Y1 y2, M = (X1 + Y1)(x2 + y2) as plain integers (the sums unreduced,
    // < 2p < 2^256):  E = M - A - B < 2 p^2  and  H = (A + B) + 4A < 6 p^2, brought under p 2^256 by two conditional
    // subtractions on the top half.  One reduction (48 multiplier issues of 904) and ~120 ALU instructions fewer
    // than three multiplications followed by modular additions.
    mont_mul<FQ>(C, acc.t, k2);
    uint32_t wa[16], wb[16], wm[16];
    add8_raw(t0.v, acc.x.v, acc.y.v);
    add8_raw(t1.v, x2.v, y2.v);
    mul_wide(wm, t0.v, t1.v);
    mul_wide(wa, acc.x.v, x2.v);
    sub16(wm, wm, wa);
    mul_wide(wb, acc.y.v, y2.v);
    sub16(wm, wm, wb);                     // E (wide)
    add16(wb, wb, wa);                     // A + B            top half < 0.9 p
    shl2_16(wa, wa);                       // 4A               top half < 1.8 p
    cond_sub_p_top<FQ>(wa);                //                  top half < p
    add16(wb, wb, wa);                     // H (wide)         top half < 1.9 p
    cond_sub_p_top<FQ>(wb);
    redc_wide<FQ, true>(E, wm);            // E in [0, 2p): it only meets F, H < p as the scanned operand below
    redc_wide<FQ>(H, wb);
    fe_sub<FQ>(F, acc.z, C);
    add8_raw(G.v, acc.z.v, C.v);           // G in [0, 2p) likewise
  } else if (S == SUITE_BJJ && AVRF_LAZY_MADD) {
    // a = 1 (Baby-JubJub in its Edwards form): the same lazy reduction with H = B - A.  The difference of the two wide
    // products is kept non-negative by adding p 2^256 when it borrows (H + p 2^256 < p 2^256 then).
    mont_mul<FQ>(C, acc.t, k2);
    uint32_t wa[16], wb[16], wm[16];
    add8_raw(t0.v, acc.x.v, acc.y.v);      // < 2p < 2^255
    add8_raw(t1.v, x2.v, y2.v);
    mul_wide(wm, t0.v, t1.v);
    mul_wide(wa, acc.x.v, x2.v);
    sub16(wm, wm, wa);
    mul_wide(wb, acc.y.v, y2.v);
    sub16(wm, wm, wb);                     // E (wide) < 2 p^2
    wb[0] = sub_cc(wb[0], wa[0]);          // H (wide) = B - A ...
#pragma unroll
    for (int i = 1; i < 16; i++) wb[i] = subc_cc(wb[i], wa[i]);
    uint32_t borrow = subc(0, 0);          // all ones when B < A
    wb[8] = add_cc(wb[8], AVRF_FC(FQ).p[0] & borrow);   // ... + p 2^256 in that case
#pragma unroll
    for (int i = 1; i < 7; i++) wb[8 + i] = addc_cc(wb[8 + i], AVRF_FC(FQ).p[i] & borrow);
    wb[15] = addc(wb[15], AVRF_FC(FQ).p[7] & borrow);
    redc_wide<FQ, true>(E, wm);            // E in [0, 2p)
    redc_wide<FQ>(H, wb);
    fe_sub<FQ>(F, acc.z, C);
    add8_raw(G.v, acc.z.v, C.v);           // G in [0, 2p)
  } else {
    mont_mul<FQ>(A, acc.x, x2);
    mont_mul<FQ>(B, acc.y, y2);
    mont_mul<FQ>(C, acc.t, k2);
    fe_add<FQ>(t0, acc.x, acc.y);
    fe_add<FQ>(t1, x2, y2);
    mont_mul<FQ>(E, t0, t1);
    fe_sub<FQ>(E, E, A);
    fe_sub<FQ>(E, E, B);
    fe_sub<FQ>(F, acc.z, C);
    fe_add<FQ>(G, acc.z, C);
    sub_a_times<S>(H, B, A);
  }
  // E and G may be in [0, 2p) (lazy form): they go in as the SCANNED operand b, for which mont_mul only needs 256 bits;
  // the multiplicand a must stay below p (its running value is bounded by a + p < 2^256).
  mont_mul<FQ>(acc.x, F, E);
  mont_mul<FQ>(acc.y, H, G);
  mont_mul<FQ>(acc.t, H, E);
  mont_mul<FQ>(acc.z, F, G);
}

// Unified full addition  r = p + q   [10 field multiplications]
template <int S>
AVRF_HD void ext_add(Ext& r, const Ext& p, const Ext& q) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe A, B, C, D, E, F, G, H, t0, t1, d;
  fe_set(d, AVRF_CC(S).d);
  mont_mul<FQ>(A, p.x, q.x);
  mont_mul<FQ>(B, p.y, q.y);
  mont_mul<FQ>(C, p.t, q.t);
  mont_mul<FQ>(C, C, d);
  mont_mul<FQ>(D, p.z, q.z);
  fe_add<FQ>(t0, p.x, p.y);
  fe_add<FQ>(t1, q.x, q.y);
  mont_mul<FQ>(E, t0, t1);
  fe_sub<FQ>(E, E, A);
  fe_sub<FQ>(E, E, B);
  fe_sub<FQ>(F, D, C);
  fe_add<FQ>(G, D, C);
  sub_a_times<S>(H, B, A);
  mont_mul<FQ>(r.x, E, F);
  mont_mul<FQ>(r.y, G, H);
  mont_mul<FQ>(r.t, E, H);
  mont_mul<FQ>(r.z, F, G);
}

// Doubling (dbl-2008-hwcd)   [4 squarings + 4 multiplications]
template <int S>
AVRF_HD void ext_dbl(Ext& r, const Ext& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe A, B, C, D, E, F, G, H, t0;
  mont_sqr<FQ>(A, p.x);
  mont_sqr<FQ>(B, p.y);
  mont_sqr<FQ>(C, p.z);
  fe_dbl<FQ>(C, C);
  a_times<S>(D, A);
  fe_add<FQ>(t0, p.x, p.y);
  mont_sqr<FQ>(E, t0);
  fe_sub<FQ>(E, E, A);
  fe_sub<FQ>(E, E, B);
  fe_add<FQ>(G, D, B);
  fe_sub<FQ>(F, G, C);
  fe_sub<FQ>(H, D, B);
  mont_mul<FQ>(r.x, E, F);
  mont_mul<FQ>(r.y, G, H);
  mont_mul<FQ>(r.t, E, H);
  mont_mul<FQ>(r.z, F, G);
}

// Doubling whose result is only doubled again: T3 = E*H is not needed by dbl-2008-hwcd, one multiplication less.
// The T coordinate of the result is NOT valid.
template <int S>
AVRF_HD void ext_dbl_not(Ext& r, const Ext& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe A, B, C, D, E, F, G, H, t0;
  mont_sqr<FQ>(A, p.x);
  mont_sqr<FQ>(B, p.y);
  mont_sqr<FQ>(C, p.z);
  fe_dbl<FQ>(C, C);
  a_times<S>(D, A);
  fe_add<FQ>(t0, p.x, p.y);
  mont_sqr<FQ>(E, t0);
  fe_sub<FQ>(E, E, A);
  fe_sub<FQ>(E, E, B);
  fe_add<FQ>(G, D, B);
  fe_sub<FQ>(F, G, C);
  fe_sub<FQ>(H, D, B);
  mont_mul<FQ>(r.x, E, F);
  mont_mul<FQ>(r.y, G, H);
  mont_mul<FQ>(r.z, F, G);
  fe_zero(r.t);
}

// out-of-line variants for everything outside the accumulation hot loop
template <int S>
AVRF_HD_CALL Ext ext_add_v(Ext p, Ext q) {
  Ext r;
  ext_add<S>(r, p, q);
  return r;
}
template <int S>
AVRF_HD_CALL Ext ext_dbl_v(Ext p) {
  Ext r;
  ext_dbl<S>(r, p);
  return r;
}
template <int S>
AVRF_HD_CALL Ext ext_dbl_not_v(Ext p) {
  Ext r;
  ext_dbl_not<S>(r, p);
  return r;
}
template <int S>
AVRF_HD void ext_add_c(Ext& r, const Ext& p, const Ext& q) { r = ext_add_v<S>(p, q); }
template <int S>
AVRF_HD void ext_dbl_c(Ext& r, const Ext& p) { r = ext_dbl_v<S>(p); }

template <int S>
AVRF_HD void ext_neg(Ext& r, const Ext& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  fe_neg<FQ>(r.x, p.x);
  r.y = p.y;
  r.z = p.z;
  fe_neg<FQ>(r.t, p.t);
}

template <int S>
AVRF_HD void ext_to_affine(Affine& r, const Ext& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe zi;
  fe_inv<FQ>(zi, p.z);
  mont_mul_c<FQ>(r.x, p.x, zi);
  mont_mul_c<FQ>(r.y, p.y, zi);
}

// r = k * p, k a plain (non-Montgomery) integer of `bits` bits given as 8 limbs.
// Left-to-right with a 2-bit fixed window (small table: register/local-memory pressure).
template <int S>
AVRF_HD_CALL Ext ext_scalar_mul_v(Ext p, Fe k, int bits) {
  Ext p2 = ext_dbl_v<S>(p);
  Ext p3 = ext_add_v<S>(p2, p);
  Ext acc;
  ext_identity<S>(acc);
  int top = (bits + 1) & ~1;
#pragma unroll 1
  for (int i = top - 2; i >= 0; i -= 2) {
    acc = ext_dbl_v<S>(acc);
    acc = ext_dbl_v<S>(acc);
    uint32_t dgt = (k.v[i >> 5] >> (i & 31)) & 3;
    if (dgt == 1) acc = ext_add_v<S>(acc, p);
    else if (dgt == 2) acc = ext_add_v<S>(acc, p2);
    else if (dgt == 3) acc = ext_add_v<S>(acc, p3);
  }
  return acc;
}
// Long scalars: radix-16 Booth digits d_i = -8 k_{4i+3} + 4 k_{4i+2} + 2 k_{4i+1} + k_{4i} + k_{4i-1} in [-8, 8] (no carry
// chain: every digit is read off five adjacent bits), an 8-entry table P .. 8P and a conditional negation.  Every
// window adds table[|d|] with the unified addition - d = 0 adds the identity - so the lanes of a warp never diverge
// on their scalars.  The table is 1 KiB per thread of local memory (a 16-entry table of unsigned digits made the
// kernels spill ~3 KB of DRAM traffic per scalar multiplication); three of every four doublings skip T.
// 256 bits: 256 doublings + 65 additions + 7 for the table.
template <int S>
AVRF_HD_CALL Ext ext_scalar_mul_w4_v(Ext p, Fe k, int bits) {
  constexpr int FQ = SuiteT<S>::FQ;
  Ext tbl[8];
  tbl[0] = p;
#pragma unroll 1
  for (int i = 1; i < 8; i++) tbl[i] = (i & 1) ? ext_dbl_v<S>(tbl[i >> 1]) : ext_add_v<S>(tbl[i - 1], p);
  Ext acc;
  ext_identity<S>(acc);
  const int top = (bits + 3) >> 2;                    // digits top .. 0 (digit `top` is the Booth carry: 0 or 1)
#pragma unroll 1
  for (int i = top; i >= 0; i--) {
    // (inlining the five point operations of a window was measured: no gain for the scalar-multiplication kernel, which is
    // bound by its 3-4 resident warps per scheduler, and 20 % slower k_prove / k_verify_each, which call this four times)
    if (i != top) {
      acc = ext_dbl_not_v<S>(acc);                    // T is only needed by the addition after the fourth doubling
      acc = ext_dbl_not_v<S>(acc);
      acc = ext_dbl_not_v<S>(acc);
      acc = ext_dbl_v<S>(acc);
    }
    // five bits 4i-1 .. 4i+3 of k (bit -1 and bits >= 256 are zero)
    int lo = 4 * i - 1;
    uint32_t f;
    if (lo < 0) {
      f = (k.v[0] << 1) & 0x1fu;
    } else {
      uint32_t w0 = lo < 256 ? k.v[lo >> 5] : 0u, w1 = (lo >> 5) + 1 < 8 ? k.v[(lo >> 5) + 1] : 0u;
      uint64_t ww = ((uint64_t)w1 << 32) | w0;
      f = (uint32_t)(ww >> (lo & 31)) & 0x1fu;
    }
    int d = (int)((f >> 1) & 7u) + (int)(f & 1u) - 8 * (int)(f >> 4);
    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    Ext q;
    if (mag == 0) {
      ext_identity<S>(q);
    } else {
      q = tbl[mag - 1];
      if (d < 0) { fe_neg<FQ>(q.x, q.x); fe_neg<FQ>(q.t, q.t); }
    }
    acc = ext_add_v<S>(acc, q);
  }
  return acc;
}
// ---------------------------------------------------------------------------------------
// GLV scalar multiplication on Bandersnatch.  The curve has the degree-2 endomorphism
//     psi(x, y) = ( x (y^2 + E0) / (C1 y) ,  (y^2 + BN) / (BD y^2 - 1) ),     psi^2 = [-2],
// which acts on the prime-order subgroup as multiplication by lambda = sqrt(-2) mod r.  A scalar k splits into
// k1 + k2 lambda (mod r) with |k1|, |k2| < 2^128, so k P = k1 P + k2 psi(P) needs 128 doublings instead of 256.
// ONLY valid for P in the prime-order subgroup (psi(P) = lambda P does not hold on the 2-torsion): used for the
// library's own hash-to-curve outputs and the generator, never for caller-supplied points.
// Constants: tools/gen_constants.py (fitted to (P, lambda P) samples and self-checked there).
// ---------------------------------------------------------------------------------------
struct GlvConsts {
  uint32_t bd[8], bn[8], c1[8], e0[8];     // Montgomery
  uint32_t g1[9], g2[9];                   // rounding multipliers, c_i = (k g_i) >> 384
  uint32_t A1[10], A2[10], B1[10], B2[10]; // 320-bit two's complement
  uint32_t lambda[8];
};
static const GlvConsts GLV_HOST = AVRF_GLV_CONSTS_INIT;
#ifdef __CUDACC__
static __constant__ GlvConsts GLV_DEV = AVRF_GLV_CONSTS_INIT;
#endif
#ifdef __CUDA_ARCH__
#define AVRF_GLV GLV_DEV
#else
#define AVRF_GLV GLV_HOST
#endif

struct GlvSplit { Fe k1, k2; bool neg1, neg2; };

// acc[0..9] += m[0..4] * A[0..9]   (mod 2^320; A in two's complement, m unsigned)
AVRF_HD void glv_mac320(uint32_t* acc, const uint32_t* m, const uint32_t* A) {
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    uint64_t carry = 0;
#pragma unroll 1
    for (int j = 0; i + j < 10; j++) {
      uint64_t t = (uint64_t)m[i] * A[j] + acc[i + j] + carry;
      acc[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
  }
}

// m[0..4] = (k[0..7] * g[0..8]) >> 384
AVRF_HD void glv_round_mul(uint32_t* m, const uint32_t* k, const uint32_t* g) {
  uint32_t t[17];
#pragma unroll 1
  for (int i = 0; i < 17; i++) t[i] = 0;
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    uint64_t carry = 0;
#pragma unroll 1
    for (int j = 0; j < 9; j++) {
      uint64_t v = (uint64_t)k[i] * g[j] + t[i + j] + carry;
      t[i + j] = (uint32_t)v;
      carry = v >> 32;
    }
    t[i + 9] = (uint32_t)carry;
  }
#pragma unroll 1
  for (int i = 0; i < 5; i++) m[i] = t[12 + i];
}

AVRF_HD_CALL GlvSplit glv_split_v(Fe k) {
  GlvSplit out;
  uint32_t m1[5], m2[5], a[10], b[10];
  glv_round_mul(m1, k.v, AVRF_GLV.g1);
  glv_round_mul(m2, k.v, AVRF_GLV.g2);
  // a = m1 A1 + m2 A2 ;  b = m1 B1 + m2 B2   (mod 2^320)
#pragma unroll 1
  for (int i = 0; i < 10; i++) a[i] = b[i] = 0;
  glv_mac320(a, m1, AVRF_GLV.A1);
  glv_mac320(a, m2, AVRF_GLV.A2);
  glv_mac320(b, m1, AVRF_GLV.B1);
  glv_mac320(b, m2, AVRF_GLV.B2);
  // k1 = k - a ;  k2 = -b
  uint32_t k1[10], k2[10];
  uint64_t br = 0;
#pragma unroll 1
  for (int i = 0; i < 10; i++) {
    uint64_t ki = i < 8 ? k.v[i] : 0u;
    uint64_t d = ki - a[i] - br;
    k1[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  br = 0;
#pragma unroll 1
  for (int i = 0; i < 10; i++) {
    uint64_t d = (uint64_t)0 - b[i] - br;
    k2[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
  // sign and magnitude (both fit 128 bits; the upper limbs are the sign extension)
  out.neg1 = (k1[9] >> 31) != 0;
  out.neg2 = (k2[9] >> 31) != 0;
  uint64_t c1 = out.neg1 ? 1 : 0, c2 = out.neg2 ? 1 : 0;
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    uint64_t v1 = (uint64_t)(out.neg1 ? ~k1[i] : k1[i]) + c1;
    uint64_t v2 = (uint64_t)(out.neg2 ? ~k2[i] : k2[i]) + c2;
    out.k1.v[i] = (uint32_t)v1; c1 = v1 >> 32;
    out.k2.v[i] = (uint32_t)v2; c2 = v2 >> 32;
  }
  return out;
}

// psi of an affine point (Montgomery coordinates), result in extended coordinates.  8 multiplications.
template <int S>
AVRF_HD_CALL Ext glv_psi_v(Affine P) {
  constexpr int FQ = SuiteT<S>::FQ;
  Ext r;
  Fe y2, n1, d1, n2, d2, t, one, c;
  fe_one<FQ>(one);
  mont_sqr_c<FQ>(y2, P.y);
  fe_set(c, AVRF_GLV.e0);
  fe_add<FQ>(t, y2, c);
  mont_mul_c<FQ>(n1, P.x, t);               // x (y^2 + E0)
  fe_set(c, AVRF_GLV.c1);
  mont_mul_c<FQ>(d1, c, P.y);               // C1 y
  fe_set(c, AVRF_GLV.bn);
  fe_add<FQ>(n2, y2, c);                    // y^2 + BN
  fe_set(c, AVRF_GLV.bd);
  mont_mul_c<FQ>(d2, c, y2);
  fe_sub<FQ>(d2, d2, one);                  // BD y^2 - 1
  mont_mul_c<FQ>(r.x, n1, d2);
  mont_mul_c<FQ>(r.y, n2, d1);
  mont_mul_c<FQ>(r.z, d1, d2);
  mont_mul_c<FQ>(r.t, n1, n2);
  return r;
}

// k * P for P (affine, Montgomery coordinates) in the prime-order subgroup of Bandersnatch.  Joint radix-16 Booth
// digits of k1 and k2 over two 8-entry tables: 132 doublings + 2 x 34 additions + 14 for the tables.
template <int S>
AVRF_HD_CALL Ext ext_scalar_mul_glv_v(Affine P, Fe k) {
  constexpr int FQ = SuiteT<S>::FQ;
  GlvSplit sp = glv_split_v(k);
  Ext p1, p2;
  affine_to_ext<S>(p1, P);
  p2 = glv_psi_v<S>(P);
  if (sp.neg1) ext_neg<S>(p1, p1);
  if (sp.neg2) ext_neg<S>(p2, p2);
  Ext tbl[2][8];
  tbl[0][0] = p1;
  tbl[1][0] = p2;
#pragma unroll 1
  for (int i = 1; i < 8; i++) {
    tbl[0][i] = (i & 1) ? ext_dbl_v<S>(tbl[0][i >> 1]) : ext_add_v<S>(tbl[0][i - 1], p1);
    tbl[1][i] = (i & 1) ? ext_dbl_v<S>(tbl[1][i >> 1]) : ext_add_v<S>(tbl[1][i - 1], p2);
  }
  Ext acc;
  ext_identity<S>(acc);
  const int top = 33;                                 // 4 * 33 = 132 bits cover |k_i| < 2^128 and the Booth carry
#pragma unroll 1
  for (int i = top; i >= 0; i--) {
    if (i != top) {
      acc = ext_dbl_not_v<S>(acc);
      acc = ext_dbl_not_v<S>(acc);
      acc = ext_dbl_not_v<S>(acc);
      acc = ext_dbl_v<S>(acc);
    }
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const Fe& kk = h ? sp.k2 : sp.k1;
      int lo = 4 * i - 1;
      uint32_t f;
      if (lo < 0) {
        f = (kk.v[0] << 1) & 0x1fu;
      } else {
        uint32_t w0 = (lo >> 5) < 8 ? kk.v[lo >> 5] : 0u, w1 = (lo >> 5) + 1 < 8 ? kk.v[(lo >> 5) + 1] : 0u;
        uint64_t ww = ((uint64_t)w1 << 32) | w0;
        f = (uint32_t)(ww >> (lo & 31)) & 0x1fu;
      }
      int d = (int)((f >> 1) & 7u) + (int)(f & 1u) - 8 * (int)(f >> 4);
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      Ext q;
      if (mag == 0) {
        ext_identity<S>(q);
      } else {
        q = tbl[h][mag - 1];
        if (d < 0) { fe_neg<FQ>(q.x, q.x); fe_neg<FQ>(q.t, q.t); }
      }
      acc = ext_add_v<S>(acc, q);
    }
  }
  return acc;
}

// ---------------------------------------------------------------------------------------
// Shared-scalar GLV multiplication: ONE secret for a whole batch of points (Secret::output of one key over many
// inputs, src/lib.rs:391-393 in a loop).  The GLV split and the width-5 non-adjacent forms of the two halves are
// computed once (on the host); every thread then follows the same sparse addition chain - uniform control flow with
// additions only at the non-zero digits: ~129 doublings (T-less unless an addition follows) + ~43 additions + 16 point
// operations for the two tables of odd multiples, against 132 + 68 + 14 for the per-thread radix-16 Booth form.
// Same precondition as ext_scalar_mul_glv_v: P in the prime-order subgroup.
// ---------------------------------------------------------------------------------------
#ifndef AVRF_NAF_W
#define AVRF_NAF_W 5            // window width of the shared-scalar plan: digits odd in (-2^(W-1), 2^(W-1)); W = 4 halves the
                                // table (1.1 instead of 3.2 GB of DRAM reads per 2^20 points) at the same speed: 35.8 vs 35.5 ms
#endif
constexpr int NAF_TBL = 1 << (AVRF_NAF_W - 2);   // odd multiples 1, 3, .. per base
struct NafPlan {
  int8_t d1[132], d2[132];      // digit i of k1 / k2: 0 or odd in [-15, 15]
  int32_t top;                  // highest index with a non-zero digit in either form (-1: k = 0)
  uint32_t neg1, neg2;          // signs of k1, k2
};

AVRF_HD void naf5(int8_t* d, int& top, const Fe& k) {
  uint32_t v[9];
#pragma unroll 1
  for (int i = 0; i < 8; i++) v[i] = k.v[i];
  v[8] = 0;
  top = -1;
#pragma unroll 1
  for (int i = 0; i < 132; i++) {
    int dg = 0;
    if (v[0] & 1u) {
      dg = (int)(v[0] & ((1u << AVRF_NAF_W) - 1u));
      if (dg >= (1 << (AVRF_NAF_W - 1))) dg -= 1 << AVRF_NAF_W;
      // v -= dg
      if (dg > 0) {
        uint64_t br = (uint64_t)dg;
        for (int l = 0; l < 9 && br; l++) { uint64_t t = (uint64_t)v[l] - br; v[l] = (uint32_t)t; br = (t >> 32) & 1; }
      } else {
        uint64_t c = (uint64_t)(-dg);
        for (int l = 0; l < 9 && c; l++) { uint64_t t = (uint64_t)v[l] + c; v[l] = (uint32_t)t; c = t >> 32; }
      }
      top = i;
    }
    d[i] = (int8_t)dg;
    for (int l = 0; l < 8; l++) v[l] = (v[l] >> 1) | (v[l + 1] << 31);
    v[8] >>= 1;
  }
}

AVRF_HD void naf_plan(NafPlan& pl, const Fe& k_canonical) {
  GlvSplit sp = glv_split_v(k_canonical);
  int t1, t2;
  naf5(pl.d1, t1, sp.k1);
  naf5(pl.d2, t2, sp.k2);
  pl.top = t1 > t2 ? t1 : t2;
  pl.neg1 = sp.neg1 ? 1u : 0u;
  pl.neg2 = sp.neg2 ? 1u : 0u;
}

template <int S>
AVRF_HD_CALL Ext ext_scalar_mul_glv_plan_v(Affine P, const NafPlan& pl) {
  constexpr int FQ = SuiteT<S>::FQ;
  Ext p1, p2, acc;
  ext_identity<S>(acc);
  if (pl.top < 0) return acc;
  affine_to_ext<S>(p1, P);
  p2 = glv_psi_v<S>(P);
  if (pl.neg1) ext_neg<S>(p1, p1);
  if (pl.neg2) ext_neg<S>(p2, p2);
  Ext tbl[2][NAF_TBL];                                 // (2i + 1) p_h
  Ext dd = ext_dbl_v<S>(p1);
  tbl[0][0] = p1;
#pragma unroll 1
  for (int i = 1; i < NAF_TBL; i++) tbl[0][i] = ext_add_v<S>(tbl[0][i - 1], dd);
  dd = ext_dbl_v<S>(p2);
  tbl[1][0] = p2;
#pragma unroll 1
  for (int i = 1; i < NAF_TBL; i++) tbl[1][i] = ext_add_v<S>(tbl[1][i - 1], dd);
#pragma unroll 1
  for (int i = pl.top; i >= 0; i--) {
    const int e1 = pl.d1[i], e2 = pl.d2[i];
    if (i != pl.top) acc = ((e1 | e2) || i == 0) ? ext_dbl_v<S>(acc) : ext_dbl_not_v<S>(acc);   // T only when an addition follows (or at the end)
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const int e = h ? e2 : e1;
      if (e == 0) continue;
      Ext q = tbl[h][((e < 0 ? -e : e) - 1) >> 1];
      if (e < 0) { fe_neg<FQ>(q.x, q.x); fe_neg<FQ>(q.t, q.t); }
      acc = ext_add_v<S>(acc, q);
    }
  }
  return acc;
}

template <int S>
AVRF_HD void ext_scalar_mul(Ext& r, const Ext& p, const uint32_t* k, int bits) {
  Fe kk;
  fe_set(kk, k);
  r = bits > 32 ? ext_scalar_mul_w4_v<S>(p, kk, bits) : ext_scalar_mul_v<S>(p, kk, bits);
}

// Compressed encoding (ark-serialize 0.6): 32-byte LE canonical y, bit 7
// of byte 31 set iff x > p - x.  `out` receives 8 little-endian words.
template <int S>
AVRF_HD void affine_compress(uint32_t* out, const Affine& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe xc, yc;
  from_mont<FQ>(xc, p.x);
  from_mont<FQ>(yc, p.y);
  bool neg = limbs_gt(xc.v, AVRF_FC(FQ).phalf);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = yc.v[i];
  if (neg) out[7] |= 0x80000000u;
}

}  // namespace avrf

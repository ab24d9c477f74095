// Hash-to-curve on the device.
//
// Elligator2 + expand_message_xmd (Bandersnatch): replaces
//   utils::hash_to_curve_ell2_xmd (src/utils/hash_to_curve.rs:66-100) and, behind it,
//   ark-ec 0.6 MapToCurveBasedHasher / Elligator2Map and ark-ff 0.6 DefaultFieldHasher
//   (behaviour as pinned by the `h` golden vectors; SURVEY.md Appendix A.6 - note Z_pad is
//   48 bytes, not the 128-byte SHA-512 block).
// Try-and-increment (Ed25519, Baby-JubJub): replaces utils::hash_to_curve_tai
//   (src/utils/hash_to_curve.rs:34-57) incl. ark-ec `Affine::from_random_bytes`.
#pragma once
#include "curve.cuh"
#include "sha512.cuh"
#include "ts_tables_gen.h"

namespace avrf {

// r = sqrt(a) (Tonelli-Shanks over p-1 = 2^s q); returns false when a is a non-residue.
// Either root may be returned - every caller normalises the sign afterwards.
struct SqrtRes { Fe r; bool ok; };

// Windowed Tonelli-Shanks tables (tools/gen_constants.py, ts_tables): index 0 = BLS12-381 Fr, 1 = BN254 Fr.
static const uint32_t TS_POW_HOST[2][4][256][8] = AVRF_TS_POW_INIT;
static const uint16_t TS_LOOK_HOST[2][1024] = AVRF_TS_LOOK_INIT;
#ifdef __CUDACC__
static __device__ const uint32_t TS_POW_DEV[2][4][256][8] = AVRF_TS_POW_INIT;
static __device__ const uint16_t TS_LOOK_DEV[2][1024] = AVRF_TS_LOOK_INIT;
#endif
#ifdef __CUDA_ARCH__
#define AVRF_TS_POW(f, i, j) TS_POW_DEV[f][i][j]
#define AVRF_TS_LOOK(f, k) TS_LOOK_DEV[f][k]
#else
#define AVRF_TS_POW(f, i, j) TS_POW_HOST[f][i][j]
#define AVRF_TS_LOOK(f, k) TS_LOOK_HOST[f][k]
#endif

template <int S> struct TsTab { static constexpr int IDX = S == SUITE_BAND ? 0 : (S == SUITE_BJJ ? 1 : -1); };

// Tonelli-Shanks with the 2-power discrete logarithm read off in four windows (p - 1 = 2^s q, s = 32 or 28): given
// x^2 = a * b with b in the subgroup of order 2^s, find e with b = g^e (g = z^q) from tables and return
// x * g^(-e/2); ok = false when e is odd (a is a non-residue).  6 w squarings (w = s/4) and 7 multiplications instead
// of the ~s^2/4 squarings of the textbook loop - the square roots of Elligator2 and of point decompression are the
// bulk of those kernels.
template <int S>
AVRF_HD_CALL SqrtRes ts_loop_v(Fe x, Fe b) {
  constexpr int FQ = SuiteT<S>::FQ;
  constexpr int TI = TsTab<S>::IDX;
  SqrtRes res;
  res.ok = true;
  if (TI >= 0) {
    constexpr int ti = TI >= 0 ? TI : 0;
    const uint32_t w = AVRF_CC(S).ts_s >> 2, mask = (1u << w) - 1;
    uint32_t e = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < 4; i++) {
      Fe t = b;
#pragma unroll 1
      for (uint32_t k = 0; k < w * (3 - i); k++) t = mont_mul_v<FQ>(t, t);
      // t = (g^(2^(3w)))^(e_i): look e_i up
      uint32_t slot = t.v[0] & 1023u, dgt = 0xffffu;
#pragma unroll 1
      for (uint32_t probe = 0; probe < 1024; probe++) {
        uint32_t j = AVRF_TS_LOOK(ti, slot);
        if (j == 0xffffu) break;
        Fe cand;
        fe_set(cand, AVRF_TS_POW(ti, 3, ((1u << w) - j) & mask));      // g^(j 2^(3w)) = inverse of g^(-j 2^(3w))
        if (fe_eq(cand, t)) { dgt = j; break; }
        slot = (slot + 1) & 1023u;
      }
      if (dgt == 0xffffu) { res.ok = false; res.r = x; return res; }     // not in the subgroup: cannot happen for b = a^q
      if (i == 0 && (dgt & 1)) { res.ok = false; res.r = x; return res; }  // e odd: non-residue
      e |= dgt << (w * i);
      if (dgt) {
        Fe m;
        fe_set(m, AVRF_TS_POW(ti, i, dgt));
        b = mont_mul_v<FQ>(b, m);
      }
    }
    uint32_t f = e >> 1;
#pragma unroll 1
    for (uint32_t i = 0; i < 4; i++) {
      uint32_t dgt = (f >> (w * i)) & mask;
      if (dgt) {
        Fe m;
        fe_set(m, AVRF_TS_POW(ti, i, dgt));
        x = mont_mul_v<FQ>(x, m);
      }
    }
    res.r = x;
    return res;
  }
  // textbook loop (2-adicity 2 for 2^255 - 19: at most one round)
  Fe one, z;
  fe_one<FQ>(one);
  fe_set(z, AVRF_CC(S).ts_root);
  uint32_t v = AVRF_CC(S).ts_s;
#pragma unroll 1
  while (!fe_eq(b, one)) {
    uint32_t k = 0;
    Fe t = b;
#pragma unroll 1
    while (!fe_eq(t, one)) {
      mont_sqr_c<FQ>(t, t);
      k++;
      if (k == v) { res.ok = false; res.r = x; return res; }   // b has order 2^v: a is a non-residue
    }
    Fe g = z;
#pragma unroll 1
    for (uint32_t i = 0; i + k + 1 < v; i++) mont_sqr_c<FQ>(g, g);
    mont_sqr_c<FQ>(z, g);
    mont_mul_c<FQ>(b, b, z);
    mont_mul_c<FQ>(x, x, g);
    v = k;
  }
  res.r = x;
  return res;
}

template <int S>
AVRF_HD_CALL SqrtRes fe_sqrt_v(Fe a) {
  constexpr int FQ = SuiteT<S>::FQ;
  SqrtRes res;
  fe_zero(res.r);
  res.ok = true;
  if (fe_is_zero(a)) return res;
  Fe w, x, b;
  fe_pow<FQ>(w, a, AVRF_CC(S).ts_exp);   // a^((q-1)/2)
  mont_mul_c<FQ>(x, a, w);                 // a^((q+1)/2)
  mont_mul_c<FQ>(b, x, w);                 // a^q
  return ts_loop_v<S>(x, b);
}

// sqrt(a) when a is a residue (ok = true), otherwise sqrt(Z * a) (ok = false; Z the Elligator2 non-residue, so
// Z * a is a residue).  One exponentiation serves both: (Z a)^((q-1)/2) = Z^((q-1)/2) * a^((q-1)/2).  a != 0.
template <int S>
AVRF_HD_CALL SqrtRes fe_sqrt_or_z_v(Fe a) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe w, x, b, za, cw;
  fe_pow<FQ>(w, a, AVRF_CC(S).ts_exp);
  mont_mul_c<FQ>(x, a, w);
  mont_mul_c<FQ>(b, x, w);
  SqrtRes r = ts_loop_v<S>(x, b);
  if (r.ok) return r;
  fe_set(za, AVRF_CC(S).zz);
  fe_set(cw, AVRF_CC(S).zcw);
  mont_mul_c<FQ>(za, za, a);               // Z a
  mont_mul_c<FQ>(w, w, cw);                // (Z a)^((q-1)/2)
  mont_mul_c<FQ>(x, za, w);
  mont_mul_c<FQ>(b, x, w);
  r = ts_loop_v<S>(x, b);
  r.ok = false;
  return r;
}
template <int S>
AVRF_HD bool fe_sqrt(Fe& r, const Fe& a) {
  SqrtRes q = fe_sqrt_v<S>(a);
  r = q.r;
  return q.ok;
}

// 48 big-endian bytes (six BE 64-bit words, most significant first) reduced into field F,
// result in Montgomery form.  value = hi*2^256 + lo, hi < 2^128.
template <int F>
AVRF_HD void fe_from_be48_mont(Fe& r, const uint64_t* be6) {
  Fe lo, hi, t;
  fe_zero(hi);
  // be6[0] is the most significant 8 bytes
  hi.v[3] = (uint32_t)(be6[0] >> 32); hi.v[2] = (uint32_t)be6[0];
  hi.v[1] = (uint32_t)(be6[1] >> 32); hi.v[0] = (uint32_t)be6[1];
  lo.v[7] = (uint32_t)(be6[2] >> 32); lo.v[6] = (uint32_t)be6[2];
  lo.v[5] = (uint32_t)(be6[3] >> 32); lo.v[4] = (uint32_t)be6[3];
  lo.v[3] = (uint32_t)(be6[4] >> 32); lo.v[2] = (uint32_t)be6[4];
  lo.v[1] = (uint32_t)(be6[5] >> 32); lo.v[0] = (uint32_t)be6[5];
  reduce_once<F>(lo, lo);
  to_mont<F>(lo, lo);                    // lo * R
  to_mont<F>(t, hi);                     // hi * R     (= canonical value of hi*2^256)
  to_mont<F>(t, t);                      // hi * R * R (= Montgomery form of hi*2^256)
  fe_add<F>(r, lo, t);
}

// 48 little-endian bytes (six LE 64-bit words, least significant first) reduced into
// field F, canonical (non-Montgomery) result: nonce_scalar (src/utils/common.rs:66-70).
template <int F>
AVRF_HD void fe_from_le48(Fe& r, const uint64_t* le6) {
  Fe lo, hi, t;
  fe_zero(hi);
#pragma unroll
  for (int i = 0; i < 4; i++) { lo.v[2 * i] = (uint32_t)le6[i]; lo.v[2 * i + 1] = (uint32_t)(le6[i] >> 32); }
  hi.v[0] = (uint32_t)le6[4]; hi.v[1] = (uint32_t)(le6[4] >> 32);
  hi.v[2] = (uint32_t)le6[5]; hi.v[3] = (uint32_t)(le6[5] >> 32);
  reduce_once<F>(lo, lo);
  to_mont<F>(t, hi);                     // canonical value of hi * 2^256 mod p
  fe_add<F>(r, lo, t);
}

// Elligator2 map of one field element u (Montgomery form) to the twisted-Edwards model, in extended coordinates.
// Same point as ark-ec 0.6 Elligator2Map::map_to_curve (behaviour pinned by the `h` golden vectors),
// arranged for few long operations: the caller supplies 1/den (den = 1 + Z u^2, or 1 when that is zero; the two maps of
// one hash share a single inversion), the second candidate reuses the first one's exponentiation through
// g(x2) = Z u^2 g(x1), and the Montgomery -> Edwards change of model is done projectively (no inversion).
template <int S>
AVRF_HD_CALL Ext ell2_map_ext_v(Fe u, Fe inv_den, bool den_was_zero) {
  constexpr int FQ = SuiteT<S>::FQ;
  Ext out;
  Fe one, jk, k2inv, K, x1, x, gx, y, t, a;
  fe_one<FQ>(one);
  fe_set(jk, AVRF_CC(S).jk);
  fe_set(k2inv, AVRF_CC(S).k2inv);
  fe_set(K, AVRF_CC(S).kk);
  mont_mul_c<FQ>(x1, jk, inv_den);
  fe_neg<FQ>(x1, x1);                    // x1 = -(J/K) / den
  // g(x) = x^3 + (J/K) x^2 + x / K^2 = x * (x * (x + J/K) + 1/K^2)
  fe_add<FQ>(a, x1, jk);
  mont_mul_c<FQ>(a, a, x1);
  fe_add<FQ>(a, a, k2inv);
  mont_mul_c<FQ>(gx, a, x1);
  bool sgn;
  if (fe_is_zero(gx)) {                  // g(x1) = 0 counts as "not a square" (as upstream): x2, y = 0
    fe_add<FQ>(x, x1, jk);
    fe_neg<FQ>(x, x);
    fe_zero(y);
    sgn = false;
  } else {
    SqrtRes r = fe_sqrt_or_z_v<S>(gx);
    if (r.ok) {
      x = x1;
      y = r.r;
      sgn = true;
    } else {
      fe_add<FQ>(x, x1, jk);
      fe_neg<FQ>(x, x);                  // x2 = -x1 - J/K
      // g(x2) = Z u^2 g(x1)  =>  sqrt(g(x2)) = u * sqrt(Z g(x1));  in the exceptional case den = 0 upstream takes
      // den = 1, which makes x2 = 0 and g(x2) = 0
      if (den_was_zero) fe_zero(y);
      else mont_mul_c<FQ>(y, r.r, u);
      sgn = false;
    }
  }
  Fe yc;
  from_mont<FQ>(yc, y);
  if (((yc.v[0] & 1) != 0) != sgn) fe_neg<FQ>(y, y);
  // Montgomery (s, t) = (x K, y K)  ->  Edwards (s / t, (s - 1) / (s + 1)); (0, 1) when (s + 1) t = 0
  Fe sM, tM, yn, yd;
  mont_mul_c<FQ>(sM, x, K);
  mont_mul_c<FQ>(tM, y, K);
  fe_sub<FQ>(yn, sM, one);
  fe_add<FQ>(yd, sM, one);
  mont_mul_c<FQ>(out.z, tM, yd);
  if (fe_is_zero(out.z)) {
    ext_identity<S>(out);
    return out;
  }
  mont_mul_c<FQ>(out.x, sM, yd);
  mont_mul_c<FQ>(out.y, yn, tM);
  mont_mul_c<FQ>(out.t, sM, yn);
  (void)t;
  return out;
}

// expand_message_xmd(SHA-512) as ark-ff 0.6 does it, 96 output bytes -> (u0, u1).
template <int S>
AVRF_HD void ell2_hash_to_field(Fe& u0, Fe& u1, const uint8_t* msg, uint32_t len) {
  constexpr int FQ = SuiteT<S>::FQ;
  const uint32_t sid_len = AVRF_CC(S).sid_len;
  const uint32_t dst_len = sid_len + 1;
  Sha512 c;
  uint64_t b0[8], b1[8], b2[8];
  auto put_dst_prime = [&](Sha512& h) {
    for (uint32_t i = 0; i < sid_len; i++) sha512_put_byte(h, AVRF_CC(S).suite_id[i]);
    sha512_put_byte(h, 0x60);
    sha512_put_byte(h, dst_len);
  };
  sha512_init(c);
  for (int i = 0; i < 48; i++) sha512_put_byte(c, 0);          // Z_pad = L bytes (arkworks)
  sha512_update(c, msg, len);
  sha512_put_byte(c, 0); sha512_put_byte(c, 96);               // I2OSP(96, 2)
  sha512_put_byte(c, 0);
  put_dst_prime(c);
  sha512_final(c, b0);
  sha512_init(c);
  for (int i = 0; i < 8; i++) sha512_put_le64(c, bswap64(b0[i]));
  sha512_put_byte(c, 1);
  put_dst_prime(c);
  sha512_final(c, b1);
  sha512_init(c);
  for (int i = 0; i < 8; i++) sha512_put_le64(c, bswap64(b0[i] ^ b1[i]));
  sha512_put_byte(c, 2);
  put_dst_prime(c);
  sha512_final(c, b2);
  uint64_t e1[6] = {b1[6], b1[7], b2[0], b2[1], b2[2], b2[3]};
  fe_from_be48_mont<FQ>(u0, b1);
  fe_from_be48_mont<FQ>(u1, e1);
}

template <int S>
AVRF_HD void hash_to_curve_ell2(Affine& out, const uint8_t* msg, uint32_t len) {
  constexpr int FQ = SuiteT<S>::FQ;
  Fe u0, u1, one, Z, d0, d1, t, inv;
  ell2_hash_to_field<S>(u0, u1, msg, len);
  fe_one<FQ>(one);
  fe_set(Z, AVRF_CC(S).zz);
  mont_sqr_c<FQ>(t, u0);
  mont_mul_c<FQ>(t, t, Z);
  fe_add<FQ>(d0, one, t);                // 1 + Z u0^2
  mont_sqr_c<FQ>(t, u1);
  mont_mul_c<FQ>(t, t, Z);
  fe_add<FQ>(d1, one, t);
  bool z0 = fe_is_zero(d0), z1 = fe_is_zero(d1);
  if (z0) d0 = one;
  if (z1) d1 = one;
  mont_mul_c<FQ>(t, d0, d1);
  fe_inv<FQ>(inv, t);                    // one inversion for both maps
  Fe i0, i1;
  mont_mul_c<FQ>(i0, inv, d1);
  mont_mul_c<FQ>(i1, inv, d0);
  Ext e0 = ell2_map_ext_v<S>(u0, i0, z0);
  Ext e1 = ell2_map_ext_v<S>(u1, i1, z1);
  ext_add_c<S>(e0, e0, e1);
  for (uint32_t i = 0; i < AVRF_CC(S).cof_log2; i++) ext_dbl_c<S>(e0, e0);
  ext_to_affine<S>(out, e0);
}

// ark-ec `Affine::get_point_from_y_unchecked`: x from y, larger root iff `greatest`.
struct PointRes { Affine p; bool ok; };

template <int S>
AVRF_HD_CALL PointRes point_from_y_v(Fe y, bool greatest) {
  constexpr int FQ = SuiteT<S>::FQ;
  PointRes res;
  res.ok = false;
  fe_zero(res.p.x);
  res.p.y = y;
  Fe one, d, y2, num, den, inv, x2, x, xc;
  fe_one<FQ>(one);
  fe_set(d, AVRF_CC(S).d);
  mont_sqr_c<FQ>(y2, y);
  fe_sub<FQ>(num, one, y2);              // 1 - y^2
  mont_mul_c<FQ>(den, d, y2);
  Fe a1;
  a_times<S>(a1, one);                   // a
  fe_sub<FQ>(den, a1, den);              // a - d y^2
  if (fe_is_zero(den)) return res;
  fe_inv<FQ>(inv, den);
  mont_mul_c<FQ>(x2, num, inv);
  if (!fe_sqrt<S>(x, x2)) return res;
  from_mont<FQ>(xc, x);
  bool is_big = limbs_gt(xc.v, AVRF_FC(FQ).phalf);
  if (is_big != greatest) fe_neg<FQ>(x, x);
  res.p.x = x;
  res.ok = true;
  return res;
}
template <int S>
AVRF_HD bool point_from_y(Affine& out, const Fe& y, bool greatest) {
  PointRes q = point_from_y_v<S>(y, greatest);
  out = q.p;
  return q.ok;
}

// Is P in the prime-order subgroup?  Cofactor-4 curves with full rational 2-torsion (Bandersnatch): E(Fq)/2E(Fq) has
// order 4 and 2E(Fq) IS the prime-order subgroup, so membership is a 2-descent - two quadratic characters on the
// Montgomery model u = (1 + y)/(1 - y):  chi(u) and chi(u - alpha), alpha a root of u^2 + A u + 1.  For this curve the
// subgroup is the class where both are -1 (the curve is the quadratic twist side; checked against [r]P for all four
// classes by tests/test_hostemu_cpu.py).  Two Legendre exponentiations (~620 multiplications) instead of a 253-bit
// scalar multiplication (~2900).  Other curves: [r]P == O.
template <int S>
AVRF_HD_CALL bool in_prime_subgroup_v(Affine P) {
  constexpr int FQ = SuiteT<S>::FQ;
  constexpr int FR = SuiteT<S>::FR;
  if (S == SUITE_BAND) {
    Fe one, y1, y2, t, al;
    fe_one<FQ>(one);
    if (fe_is_zero(P.x)) return fe_eq(P.y, one);       // (0, 1) is in, (0, -1) has order 2
    fe_sub<FQ>(y1, one, P.y);                          // 1 - y
    fe_add<FQ>(y2, one, P.y);                          // 1 + y
    mont_mul_c<FQ>(t, y1, y2);                         // 1 - y^2 ~ u up to squares
    if (fe_is_zero(t) || fe_is_nonzero_square<FQ>(t)) return false;
    fe_set(al, AVRF_CC(S).mt_alpha);
    mont_mul_c<FQ>(t, al, y1);
    fe_sub<FQ>(t, y2, t);                              // (1 + y) - alpha (1 - y)
    mont_mul_c<FQ>(t, t, y1);                          // ~ u - alpha up to squares
    if (fe_is_zero(t) || fe_is_nonzero_square<FQ>(t)) return false;
    return true;
  }
  Ext e, r;
  affine_to_ext<S>(e, P);
  ext_scalar_mul<S>(r, e, AVRF_FC(FR).p, 256);         // [r]P
  return ext_is_identity<S>(r);
}

// Try-and-increment (hash_to_curve.rs:34-57).  Returns false if all 256 counters fail.
template <int S>
AVRF_HD bool hash_to_curve_tai(Affine& out, const uint8_t* msg, uint32_t len) {
  constexpr int FQ = SuiteT<S>::FQ;
  Sha512 prefix;
  sha512_init(prefix);
  for (uint32_t i = 0; i < AVRF_CC(S).sid_len; i++) sha512_put_byte(prefix, AVRF_CC(S).suite_id[i]);
  sha512_put_byte(prefix, 0x60);
  sha512_put_le64(prefix, len);
  sha512_update(prefix, msg, len);
#pragma unroll 1
  for (uint32_t ctr = 0; ctr < 256; ctr++) {
    Sha512 t = prefix;
    sha512_put_byte(t, ctr);
    uint64_t seed[8], blk[8];
    sha512_final(t, seed);
    sha512_xof_block(blk, seed, 0);
    Fe y;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint64_t le = bswap64(blk[i]);
      y.v[2 * i] = (uint32_t)le;
      y.v[2 * i + 1] = (uint32_t)(le >> 32);
    }
    bool flag = (y.v[7] >> 31) & 1;
    y.v[7] &= 0xFFFFFFFFu >> (256 - AVRF_CC(S).p_bits);
    if (!limbs_gt(AVRF_FC(FQ).p, y.v)) continue;   // y >= p
    to_mont<FQ>(y, y);
    Affine P;
    if (!point_from_y<S>(P, y, flag)) continue;
    Ext e;
    affine_to_ext<S>(e, P);
    for (uint32_t i = 0; i < AVRF_CC(S).cof_log2; i++) ext_dbl_c<S>(e, e);
    if (ext_is_identity<S>(e)) continue;
    ext_to_affine<S>(out, e);
    return true;
  }
  return false;
}

template <int S>
AVRF_HD bool data_to_point(Affine& out, const uint8_t* msg, uint32_t len) {
  if (S == SUITE_BAND) { hash_to_curve_ell2<S>(out, msg, len); return true; }
  return hash_to_curve_tai<S>(out, msg, len);
}

}  // namespace avrf

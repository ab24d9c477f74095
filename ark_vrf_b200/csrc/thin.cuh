// Thin-VRF protocol hashing on the device: one thread owns one proof's transcript.
//
// Replaces, for the batch path, BatchVerifier::prepare (src/thin.rs:209-226) together with
// vrf_transcript_base / chain_ios / absorb_ios (src/utils/common.rs:159-173,231-240,
// 377-383), DelinearizeScalars (common.rs:335-369), challenge (common.rs:270-280) and,
// for the prover, nonce (common.rs:313-328).  Byte layout: SURVEY.md Appendix A.4.
#pragma once
#include "curve.cuh"
#include "sha512.cuh"
#include "h2c.cuh"

namespace avrf {

enum : uint32_t {
  DOM_THIN = 0x01, DOM_NONCE_EXPAND = 0x10, DOM_NONCE = 0x11, DOM_POINT_TO_HASH = 0x20,
  DOM_DELINEARIZE = 0x30, DOM_CHALLENGE = 0x40, DOM_BATCH = 0x50, DOM_H2C = 0x60
};

// T = SUITE_ID || 0x01 || LE64(M+1) || enc(G) || enc(pk)   (the I/O pairs, then ad, follow)
template <int S>
AVRF_HD void thin_transcript_begin(Sha512& t, uint32_t n_ios, const uint32_t* pk_enc) {
  sha512_init(t);
  for (uint32_t i = 0; i < AVRF_CC(S).sid_len; i++) sha512_put_byte(t, AVRF_CC(S).suite_id[i]);
  sha512_put_byte(t, DOM_THIN);
  sha512_put_le64(t, (uint64_t)n_ios + 1);
  sha512_put_words(t, AVRF_CC(S).g_enc);
  sha512_put_words(t, pk_enc);
}

AVRF_HD void thin_transcript_ad(Sha512& t, const uint8_t* ad, uint32_t ad_len) {
  sha512_put_le64(t, ad_len);
  sha512_update(t, ad, ad_len);
}

// Delinearisation scalars z_1..z_M (16-byte LE each) from a fork of the transcript.
// `put(i, z4)` receives z_{i+1} as four 32-bit words.
template <typename Put>
AVRF_HD void thin_delinearize(const Sha512& t, uint32_t n_ios, Put put) {
  if (n_ios == 0) return;
  Sha512 tz = t;
  sha512_put_byte(tz, DOM_DELINEARIZE);
  uint64_t seed[8], blk[8];
  sha512_final(tz, seed);
  for (uint32_t i = 0; i < n_ios; i++) {
    if ((i & 3) == 0) sha512_xof_block(blk, seed, i >> 2);
    uint32_t z4[4];
    digest_le128(z4, blk, 16 * (i & 3));
    put(i, z4);
  }
}

// c = first 16 bytes of stream(T || 0x40 || enc(R)); consumes the transcript.
AVRF_HD void thin_challenge(Sha512& t, const uint32_t* r_enc, uint32_t* c4) {
  sha512_put_byte(t, DOM_CHALLENGE);
  sha512_put_words(t, r_enc);
  uint64_t seed[8], blk[8];
  sha512_final(t, seed);
  sha512_xof_block(blk, seed, 0);
  digest_le128(c4, blk, 0);
}

// Deterministic nonce (common.rs:313-328): canonical scalar k; `sk` canonical.
template <int S>
AVRF_HD void thin_nonce(Fe& k, const Sha512& t, const Fe& sk) {
  constexpr int FR = SuiteT<S>::FR;
  Sha512 te = t;
  sha512_put_byte(te, DOM_NONCE_EXPAND);
  sha512_put_words(te, sk.v);
  uint64_t seed[8], skh[8], blk[8];
  sha512_final(te, seed);
  sha512_xof_block(skh, seed, 0);          // 64-byte expanded key
  Sha512 tn = t;
  sha512_put_byte(tn, DOM_NONCE);
  for (int i = 0; i < 8; i++) sha512_put_le64(tn, bswap64(skh[i]));
  sha512_final(tn, seed);
  sha512_xof_block(blk, seed, 0);
  uint64_t le6[6];
  for (int i = 0; i < 6; i++) le6[i] = bswap64(blk[i]);   // first 48 bytes, LE
  fe_from_le48<FR>(k, le6);
}

// point_to_hash (common.rs:290-305): first 32 bytes of stream(SUITE_ID || 0x20 || enc(P)).
template <int S>
AVRF_HD void point_to_hash32(uint32_t* out8, const uint32_t* p_enc) {
  Sha512 t;
  sha512_init(t);
  for (uint32_t i = 0; i < AVRF_CC(S).sid_len; i++) sha512_put_byte(t, AVRF_CC(S).suite_id[i]);
  sha512_put_byte(t, DOM_POINT_TO_HASH);
  sha512_put_words(t, p_enc);
  uint64_t seed[8], blk[8];
  sha512_final(t, seed);
  sha512_xof_block(blk, seed, 0);
  for (int i = 0; i < 4; i++) {
    uint64_t le = bswap64(blk[i]);
    out8[2 * i] = (uint32_t)le;
    out8[2 * i + 1] = (uint32_t)(le >> 32);
  }
}

// Signed 16-bit window recoding of a canonical scalar (< 2^253): digits d_k in
// [-32768, 32767], sum d_k 2^(16k) = scalar.  Zero digits contribute nothing.
AVRF_HD void recode_signed16(int32_t* dg, const Fe& k) {
  uint32_t carry = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    uint32_t v = ((k.v[i >> 1] >> (16 * (i & 1))) & 0xffffu) + carry;
    if (v >= 32768u) { dg[i] = (int32_t)v - 65536; carry = 1; }
    else { dg[i] = (int32_t)v; carry = 0; }
  }
}

// Full thin prove for one proof (thin.rs:111-129).  sk canonical; pk Montgomery affine; get_io(k) returns point k of
// the pair list I_0, O_0, I_1, O_1, ... as a Montgomery affine point (any number of pairs: nothing is staged).
// Outputs R (Montgomery affine) and s (canonical).
template <int S, typename GetIo>
AVRF_HD void thin_prove_one_g(Affine& R, Fe& s_out, const Fe& sk, const Affine& pk, GetIo get_io,
                              uint32_t n_ios, const uint8_t* ad, uint32_t ad_len) {
  constexpr int FR = SuiteT<S>::FR;
  Sha512 t;
  uint32_t enc[8];
  affine_compress<S>(enc, pk);
  thin_transcript_begin<S>(t, n_ios, enc);
  for (uint32_t i = 0; i < 2 * n_ios; i++) {
    Affine P = get_io(i);
    affine_compress<S>(enc, P);
    sha512_put_words(t, enc);
  }
  thin_transcript_ad(t, ad, ad_len);
  // merged input I_m = G + sum z_i I_i   (common.rs:389-419 with z_0 = 1)
  Ext im;
  {
    Affine g;
    fe_set(g.x, AVRF_CC(S).gx);
    fe_set(g.y, AVRF_CC(S).gy);
    affine_to_ext<S>(im, g);
  }
  thin_delinearize(t, n_ios, [&](uint32_t i, const uint32_t* z4) {
    uint32_t z8[8] = {z4[0], z4[1], z4[2], z4[3], 0, 0, 0, 0};
    Ext e, m;
    Affine P = get_io(2 * i);
    affine_to_ext<S>(e, P);
    ext_scalar_mul<S>(m, e, z8, 128);
    ext_add_c<S>(im, im, m);
  });
  Fe k;
  thin_nonce<S>(k, t, sk);
  Ext rr;
  ext_scalar_mul<S>(rr, im, k.v, 256);
  ext_to_affine<S>(R, rr);
  affine_compress<S>(enc, R);
  uint32_t c4[4];
  thin_challenge(t, enc, c4);
  // s = k + c * sk mod r
  Fe c8, skm, cs;
  fe_zero(c8);
  for (int i = 0; i < 4; i++) c8.v[i] = c4[i];
  to_mont<FR>(skm, sk);
  mont_mul_c<FR>(cs, c8, skm);             // c * sk  (canonical)
  fe_add<FR>(s_out, k, cs);
}

template <int S>
AVRF_HD void thin_prove_one(Affine& R, Fe& s_out, const Fe& sk, const Affine& pk, const Affine* ios /*I,O pairs*/,
                            uint32_t n_ios, const uint8_t* ad, uint32_t ad_len) {
  thin_prove_one_g<S>(R, s_out, sk, pk, [ios](uint32_t k) { return ios[k]; }, n_ios, ad, ad_len);
}

}  // namespace avrf

// Test-only host build of the device arithmetic headers (fp.cuh / curve.cuh) with the
// PTX carry flag emulated in C.  Lets pytest check the exact limb algorithms against
// Python big integers without a GPU.  NOT part of the product: libavrf_gpu.so never links it.
#include "../../ark_vrf_b200/csrc/curve.cuh"
#include <string.h>
using namespace avrf;

template <int F> static void fop(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fe x, y, r; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
  switch (op) {
    case 0: mont_mul<F>(r, x, y); break;
    case 1: fe_add<F>(r, x, y); break;
    case 2: fe_sub<F>(r, x, y); break;
    case 3: fe_neg<F>(r, x); break;
    case 4: to_mont<F>(r, x); break;
    case 5: from_mont<F>(r, x); break;
    case 6: fe_inv<F>(r, x); break;
    case 7: reduce_once<F>(r, x); break;
    case 8: fe_zero(r); r.v[0] = fe_is_nonzero_square<F>(x); break;
    case 9: fe_zero(r); r.v[0] = fe_is_nonzero_square_pow<F>(x); break;
    case 10: fe_zero(r); r.v[0] = (uint32_t)(fe_jacobi_v<F>(x) + 1); break;
  }
  memcpy(out, r.v, 32);
}

template <int S> static void pop(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  // a: Ext (32 words), b: AffineK (24 words) or Ext or scalar; out: Ext
  Ext p; memcpy(&p, a, 128);
  Ext r;
  switch (op) {
    case 0: { AffineK q; memcpy(&q, b, 96); r = p; ext_madd<S>(r, q.x, q.y, q.k); break; }
    case 1: { Ext q; memcpy(&q, b, 128); ext_add<S>(r, p, q); break; }
    case 2: ext_dbl<S>(r, p); break;
    case 3: ext_scalar_mul<S>(r, p, b, 256); break;
    case 5: { Affine q; memcpy(&q, b, 64); AffineK k; affine_to_k<S>(k, q); r = p; ext_madd<S>(r, k.x, k.y, k.k); break; }
    case 6: { Affine q; memcpy(&q, b, 64); AffineK k; affine_to_k<S>(k, q); base_cneg<S>(k, true); r = p; ext_madd<S>(r, k.x, k.y, k.k); break; }
    case 4: { Affine q; ext_to_affine<S>(q, p); uint32_t c[8]; affine_compress<S>(c, q); memset(&r, 0, 128); memcpy(&r, c, 32); break; }
  }
  memcpy(out, &r, 128);
}

extern "C" {
void emu_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: fop<0>(op, a, b, out); break; case 1: fop<1>(op, a, b, out); break;
    case 2: fop<2>(op, a, b, out); break; case 3: fop<3>(op, a, b, out); break;
    case 4: fop<4>(op, a, b, out); break; case 5: fop<5>(op, a, b, out); break;
  }
}
void emu_point_op(int suite, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (suite) {
    case 0: pop<0>(op, a, b, out); break; case 1: pop<1>(op, a, b, out); break;
    case 2: pop<2>(op, a, b, out); break;
  }
}
}

#include "../../ark_vrf_b200/csrc/sha512.cuh"
extern "C" void emu_sha512(const uint8_t* msg, uint32_t n, uint8_t* out) {
  Sha512 c; sha512_init(c); sha512_update(c, msg, n);
  uint64_t d[8]; sha512_final(c, d);
  for (int i = 0; i < 8; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(d[i] >> (56 - 8 * k));
}

#include "../../ark_vrf_b200/csrc/thin.cuh"
template <int S> static int h2c_t(const uint8_t* msg, uint32_t n, uint32_t* out16) {
  Affine p; bool ok = data_to_point<S>(p, msg, n); memcpy(out16, &p, 64); return ok;
}
template <int S> static void prove_t(const uint32_t* sk, const uint32_t* pk, const uint32_t* ios, uint32_t n_ios,
                                     const uint8_t* ad, uint32_t ad_len, uint32_t* r16, uint32_t* s8) {
  Fe k; memcpy(k.v, sk, 32); Affine P; memcpy(&P, pk, 64); Affine R; Fe s;
  thin_prove_one<S>(R, s, k, P, (const Affine*)ios, n_ios, ad, ad_len);
  memcpy(r16, &R, 64); memcpy(s8, s.v, 32);
}
template <int S> static int sqrt_t(const uint32_t* a8, uint32_t* out8) {
  Fe a, r; memcpy(a.v, a8, 32); bool ok = fe_sqrt<S>(r, a); memcpy(out8, r.v, 32); return ok;
}
template <int S> static int sqrt_or_z_t(const uint32_t* a8, uint32_t* out8) {
  Fe a; memcpy(a.v, a8, 32); SqrtRes q = fe_sqrt_or_z_v<S>(a); memcpy(out8, q.r.v, 32); return q.ok;
}
template <int S> static void compress_t(const uint32_t* p16, uint32_t* out8) {
  Affine P; memcpy(&P, p16, 64); affine_compress<S>(out8, P);
}
extern "C" {
int emu_h2c(int suite, const uint8_t* msg, uint32_t n, uint32_t* out16) {
  return suite == 0 ? h2c_t<0>(msg, n, out16) : suite == 1 ? h2c_t<1>(msg, n, out16) : h2c_t<2>(msg, n, out16);
}
void emu_prove(int suite, const uint32_t* sk, const uint32_t* pk, const uint32_t* ios, uint32_t n_ios,
               const uint8_t* ad, uint32_t ad_len, uint32_t* r16, uint32_t* s8) {
  if (suite == 0) prove_t<0>(sk, pk, ios, n_ios, ad, ad_len, r16, s8);
  else if (suite == 1) prove_t<1>(sk, pk, ios, n_ios, ad, ad_len, r16, s8);
  else prove_t<2>(sk, pk, ios, n_ios, ad, ad_len, r16, s8);
}
int emu_sqrt(int suite, const uint32_t* a8, uint32_t* out8) {      // Montgomery in / out; returns 1 when a is a square
  return suite == 0 ? sqrt_t<0>(a8, out8) : suite == 1 ? sqrt_t<1>(a8, out8) : sqrt_t<2>(a8, out8);
}
int emu_in_subgroup(int suite, const uint32_t* p16) {                 // Montgomery affine point
  Affine P; memcpy(&P, p16, 64);
  return suite == 0 ? in_prime_subgroup_v<0>(P) : suite == 1 ? in_prime_subgroup_v<1>(P) : in_prime_subgroup_v<2>(P);
}
void emu_glv_split(const uint32_t* k8, uint32_t* k1, uint32_t* k2, int* neg) {
  Fe k; memcpy(k.v, k8, 32); GlvSplit sp = glv_split_v(k);
  memcpy(k1, sp.k1.v, 32); memcpy(k2, sp.k2.v, 32); neg[0] = sp.neg1; neg[1] = sp.neg2;
}
void emu_glv_psi(const uint32_t* p16, uint32_t* out32) {             // Montgomery affine in, extended out
  Affine P; memcpy(&P, p16, 64); Ext r = glv_psi_v<0>(P); memcpy(out32, &r, 128);
}
void emu_glv_mul(const uint32_t* p16, const uint32_t* k8, uint32_t* out32) {
  Affine P; memcpy(&P, p16, 64); Fe k; memcpy(k.v, k8, 32); Ext r = ext_scalar_mul_glv_v<0>(P, k); memcpy(out32, &r, 128);
}
int emu_sqrt_or_z(const uint32_t* a8, uint32_t* out8) { return sqrt_or_z_t<0>(a8, out8); }
void emu_compress(int suite, const uint32_t* p16, uint32_t* out8) {
  if (suite == 0) compress_t<0>(p16, out8); else if (suite == 1) compress_t<1>(p16, out8); else compress_t<2>(p16, out8);
}
}

// lazy reduction pieces (fp.cuh): the wide product and the Montgomery reduction of a wide value
extern "C" void emu_mul_wide(const uint32_t* a8, const uint32_t* b8, uint32_t* out16) { mul_wide(out16, a8, b8); }
extern "C" void emu_redc_wide(int field, const uint32_t* t16, uint32_t* out8) {
  Fe r;
  switch (field) {
    case 0: redc_wide<0>(r, t16); break; case 1: redc_wide<1>(r, t16); break; case 2: redc_wide<2>(r, t16); break;
    case 3: redc_wide<3>(r, t16); break; case 4: redc_wide<4>(r, t16); break; case 5: redc_wide<5>(r, t16); break;
  }
  memcpy(out8, r.v, 32);
}

extern "C" void emu_glv_mul_plan(const uint32_t* p16, const uint32_t* k8, uint32_t* out32, int32_t* top) {   // Montgomery affine in, extended out
  Affine P; memcpy(&P, p16, 64); Fe k; memcpy(k.v, k8, 32);
  NafPlan pl; naf_plan(pl, k);
  *top = pl.top;
  Ext r = ext_scalar_mul_glv_plan_v<0>(P, pl);
  memcpy(out32, &r, 128);
}

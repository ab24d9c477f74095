// Kernels either side of the batch-verification path:
// hash-to-curve, VRF outputs / public keys, bulk proving, point compression and output hashing,
// wire-format ingest, per-proof verdicts.  One thread per item.
#pragma once
#include "prepare.cuh"

namespace avrf {
__global__ void k_rebase(uint32_t* off, uint64_t count, uint32_t base) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) off[i] += base;
}

template <int S>
__global__ void __launch_bounds__(128) k_h2c(const uint8_t* msgs, const uint32_t* off, uint32_t n, Affine* out_aff,
                                             uint32_t* out_enc, uint8_t* ok, int canonical) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Affine P;
  bool good = data_to_point<S>(P, msgs + off[j], off[j + 1] - off[j]);
  if (!good) {
    fe_zero(P.x);
    fe_one<SuiteT<S>::FQ>(P.y);
  }
  if (ok) ok[j] = good ? 1 : 0;
  if (out_enc) {
    uint32_t enc[8];
    affine_compress<S>(enc, P);
    for (int i = 0; i < 8; i++) out_enc[8 * (size_t)j + i] = enc[i];
  }
  if (out_aff) store_affine_fmt<S>(out_aff + j, P, canonical);
}

// out_j = sk_j * in_j  (in == nullptr: the generator)
template <int S>
__global__ void __launch_bounds__(128) k_scalar_mul(const Fe* sk, uint32_t sk_stride_words, const Affine* in, uint32_t n,
                                                    Affine* out, int canonical) {
  constexpr int FR = SuiteT<S>::FR;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Fe k;
  load_fe(k, reinterpret_cast<const Fe*>(reinterpret_cast<const uint32_t*>(sk) + (size_t)j * sk_stride_words));
  if (!canonical) from_mont<FR>(k, k);
  Affine P;
  if (in) load_affine_fmt<S>(P, in + j, canonical);
  else { fe_set(P.x, AVRF_CC(S).gx); fe_set(P.y, AVRF_CC(S).gy); }
  Ext e, r;
  affine_to_ext<S>(e, P);
  ext_scalar_mul<S>(r, e, k.v, 256);
  ext_to_affine<S>(P, r);
  store_affine_fmt<S>(out + j, P, canonical);
}

struct ProveArgs {
  const Fe* sk;
  const Affine* pk;
  const Affine* ios;
  const uint32_t* io_off;
  const uint32_t* ad_off;
  const uint8_t* ad;
  Affine* r;
  Fe* s;
  uint32_t n;
  int canonical;
};

template <int S>
__global__ void __launch_bounds__(128) k_prove(ProveArgs a) {
  constexpr int FR = SuiteT<S>::FR;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  uint32_t io0 = a.io_off[j], m = a.io_off[j + 1] - io0;
  Affine pk, R;
  load_affine_fmt<S>(pk, a.pk + j, a.canonical);
  Fe sk, s;
  load_fe(sk, a.sk + j);
  if (!a.canonical) from_mont<FR>(sk, sk);
  uint32_t ad0 = a.ad_off[j];
  const Affine* base = a.ios + 2 * (size_t)io0;
  int canonical = a.canonical;
  // any number of I/O pairs: the points are read from device memory where they are needed (transcript, merge)
  thin_prove_one_g<S>(R, s, sk, pk, [base, canonical](uint32_t k) {
    Affine P;
    load_affine_fmt<S>(P, base + k, canonical);
    return P;
  }, m, a.ad + ad0, a.ad_off[j + 1] - ad0);
  store_affine_fmt<S>(a.r + j, R, a.canonical);
  if (!a.canonical) to_mont<FR>(s, s);
  store_fe(a.s + j, s);
}

template <int S>
__global__ void __launch_bounds__(128) k_compress(const Affine* in, uint64_t n, uint32_t* out, int canonical, int hash) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Affine P;
  load_affine_fmt<S>(P, in + j, canonical);
  uint32_t enc[8], h[8];
  affine_compress<S>(enc, P);
  if (hash) {
    point_to_hash32<S>(h, enc);
    for (int i = 0; i < 8; i++) out[8 * j + i] = h[i];
  } else {
    for (int i = 0; i < 8; i++) out[8 * j + i] = enc[i];
  }
}

// CanonicalDeserialize with Validate::Yes of compressed points (ark-serialize 0.6; reference
// src/lib.rs:410-433 Public, :471-494 Input, :552-575 Output, src/thin.rs:42 Proof.r):
// y < p, x = sqrt((1-y^2)/(a-d y^2)) picked by the sign flag, prime-subgroup check [r]P = O, and for
// kind = 1 (Public / Input / Output) the identity is rejected as well.
template <int S>
__global__ void __launch_bounds__(128) k_deserialize(const uint32_t* in, uint64_t n, int kind, Affine* out, uint8_t* ok,
                                                     int canonical) {
  constexpr int FQ = SuiteT<S>::FQ;
  constexpr int FR = SuiteT<S>::FR;
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Fe y;
#pragma unroll
  for (int i = 0; i < 8; i++) y.v[i] = in[8 * j + i];
  bool flag = (y.v[7] >> 31) & 1;
  y.v[7] &= 0x7fffffffu;
  Affine P;
  fe_zero(P.x);
  fe_one<FQ>(P.y);
  bool good = limbs_gt(AVRF_FC(FQ).p, y.v);          // y < p
  if (good) {
    to_mont<FQ>(y, y);
    good = point_from_y<S>(P, y, flag);
  }
  if (good && kind == 1 && affine_is_identity<S>(P)) good = false;
  if (good) {
    Ext e, r;
    affine_to_ext<S>(e, P);
    ext_scalar_mul<S>(r, e, AVRF_FC(FR).p, 256);     // [r]P
    good = ext_is_identity<S>(r);
  }
  if (!good) { fe_zero(P.x); fe_one<FQ>(P.y); }
  ok[j] = good ? 1 : 0;
  store_affine_fmt<S>(out + j, P, canonical);
}

// thin::Verifier::verify for every proof of a prepared batch (src/thin.rs:131-165), one thread per
// proof, reusing c_j and z_ij of k_prepare: status 0 Ok / 1 VerificationFailure / 2 InvalidData.
struct EachArgs {
  const Affine* pk;         // the pushed inputs (not the MSM bases, whose layout is suite specific)
  const Affine* r;
  const Affine* ios;
  int canonical;
  const uint32_t* cs;
  const uint32_t* z;
  const uint32_t* io_off;
  int32_t* status;
  uint32_t n;
};

template <int S>
__global__ void __launch_bounds__(128) k_verify_each(EachArgs a) {
  constexpr int FQ = SuiteT<S>::FQ;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  uint32_t io0 = a.io_off[j], m = a.io_off[j + 1] - io0;
  Affine R, pk, t;
  load_affine_fmt<S>(R, a.r + j, a.canonical);
  load_affine_fmt<S>(pk, a.pk + j, a.canonical);
  bool bad = affine_is_identity<S>(pk);
  Ext im, om, e, q;
  {
    Affine g;
    fe_set(g.x, AVRF_CC(S).gx);
    fe_set(g.y, AVRF_CC(S).gy);
    affine_to_ext<S>(im, g);              // I_m = G + sum z_i I_i
  }
  affine_to_ext<S>(om, pk);               // O_m = pk + sum z_i O_i
  for (uint32_t i = 0; i < m; i++) {
    uint32_t z8[8] = {a.z[4 * (size_t)(io0 + i)], a.z[4 * (size_t)(io0 + i) + 1], a.z[4 * (size_t)(io0 + i) + 2],
                      a.z[4 * (size_t)(io0 + i) + 3], 0, 0, 0, 0};
    load_affine_fmt<S>(t, a.ios + 2 * (size_t)(io0 + i) + 1, a.canonical);   // O_i
    bad |= affine_is_identity<S>(t);
    affine_to_ext<S>(e, t);
    ext_scalar_mul<S>(q, e, z8, 128);
    ext_add_c<S>(om, om, q);
    load_affine_fmt<S>(t, a.ios + 2 * (size_t)(io0 + i), a.canonical);       // I_i
    bad |= affine_is_identity<S>(t);
    affine_to_ext<S>(e, t);
    ext_scalar_mul<S>(q, e, z8, 128);
    ext_add_c<S>(im, im, q);
  }
  const uint32_t* csj = a.cs + 16 * (size_t)j;
  uint32_t c8[8] = {csj[0], csj[1], csj[2], csj[3], 0, 0, 0, 0};
  uint32_t s8[8];
  for (int i = 0; i < 8; i++) s8[i] = csj[8 + i];
  ext_scalar_mul<S>(q, im, s8, 256);      // s * I_m
  ext_scalar_mul<S>(e, om, c8, 128);      // c * O_m
  ext_neg<S>(e, e);
  ext_add_c<S>(q, q, e);
  affine_to_ext<S>(e, R);
  ext_neg<S>(e, e);
  ext_add_c<S>(q, q, e);                  // s I_m - c O_m - R
  (void)FQ;
  a.status[j] = bad ? AVRF_INVALID_DATA : (ext_is_identity<S>(q) ? AVRF_OK : AVRF_VERIFICATION_FAILURE);
}

}  // namespace avrf

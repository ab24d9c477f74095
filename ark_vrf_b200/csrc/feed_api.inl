// Handle-less entry points of libavrf_gpu.so (include/avrf.h, "Feeder operations" and "Measurement helpers"):
// textually included inside the extern "C" block of avrf_gpu.cu.
// ---- feeder operations ---------------------------------------------------------------------
// The field inversions of these operations are batched (feeders.cuh, "Pipelined feeder kernels").
static int batch_inv(uint32_t suite, Fe* v, Fe* scratch, uint64_t n, cudaStream_t st) {
  uint64_t threads = std::max<uint64_t>(128, (n + 31) / 32);       // 32 elements per thread
  uint32_t blocks = cdiv(threads, 128);
  DISPATCH(suite, (k_batch_inv<SuiteT<S>::FQ><<<blocks, 128, 0, st>>>(v, scratch, n)));
  LAUNCHED("k_batch_inv");
  return 0;
}

// Device-side Input::new for n messages (already on the device): Montgomery affine points in d_aff (n x 64 B, device),
// optional compressed encodings / caller-format copies / ok flags.  Scratch buffers are sized by the caller.
static int h2c_device(uint32_t suite, const uint8_t* d_msgs, const uint32_t* d_off, uint64_t n, Affine* d_aff, Affine* d_fmt,
                      uint32_t* d_enc, uint8_t* d_ok, int canonical, H2cScratch& w, cudaStream_t st) {
  int rc;
  if (suite != AVRF_SUITE_BANDERSNATCH_SHA512_ELL2) {
    // try-and-increment: data-dependent retries, kept as one kernel
    DISPATCH(suite, (k_h2c<S><<<cdiv(n, 128), 128, 0, st>>>(d_msgs, d_off, (uint32_t)n, d_aff, d_enc, d_ok, 0)));
    LAUNCHED("k_h2c");
    if (d_fmt) {
      DISPATCH(suite, (k_refmt<S><<<cdiv(n, 128), 128, 0, st>>>(d_aff, n, d_fmt, canonical)));
      LAUNCHED("k_refmt");
    }
    return 0;
  }
  if ((rc = w.u01.reserve(64 * n, 0, st)) || (rc = w.den.reserve(32 * n, 0, st)) || (rc = w.scr.reserve(32 * n, 0, st))) return rc;
  DISPATCH(suite, (k_h2f<S><<<cdiv(n, 128), 128, 0, st>>>(d_msgs, d_off, (uint32_t)n, w.u01.as<Fe>(), w.den.as<Fe>())));
  LAUNCHED("k_h2f");
  if ((rc = batch_inv(suite, w.den.as<Fe>(), w.scr.as<Fe>(), n, st))) return rc;
  DISPATCH(suite, (k_ell2_maps<S><<<cdiv(n, 128), 128, 0, st>>>(w.u01.as<Fe>(), w.den.as<Fe>(), (uint32_t)n, d_aff)));
  LAUNCHED("k_ell2_maps");
  if ((rc = batch_inv(suite, w.den.as<Fe>(), w.scr.as<Fe>(), n, st))) return rc;
  DISPATCH(suite, (k_affine_finish<S><<<cdiv(n, 128), 128, 0, st>>>(d_aff, w.den.as<Fe>(), n, d_aff, d_fmt, d_enc, d_ok, canonical)));
  LAUNCHED("k_affine_finish");
  return 0;
}

// Device-side scalar multiplication out_j = sk_j * in_j (in == nullptr: generator).
// h_sk_shared: host copy of the ONE scalar (caller format) when sk_stride == 0, else nullptr.
static int smul_device(uint32_t suite, const Fe* d_sk, uint32_t sk_stride_words, const Affine* d_in, int in_is_dev, int in_subgroup,
                       uint64_t n, Affine* d_out_dev, Affine* d_out_fmt, uint32_t* d_enc, int canonical, H2cScratch& w, cudaStream_t st,
                       const uint8_t* h_sk_shared = nullptr) {
  int rc;
  if ((rc = w.u01.reserve(64 * n, 0, st)) || (rc = w.den.reserve(32 * n, 0, st)) || (rc = w.scr.reserve(32 * n, 0, st))) return rc;
  Affine* xy = w.u01.as<Affine>();
  bool planned = false;
  if (suite == AVRF_SUITE_BANDERSNATCH_SHA512_ELL2 && in_subgroup && sk_stride_words == 0 && h_sk_shared) {
    // one secret over many subgroup points: plan the addition chain once, on the host
    Fe k;
    memcpy(k.v, h_sk_shared, 32);
    if (limbs_gt(AVRF_FC(FR_BAND).p, k.v)) {             // a field element (otherwise the generic kernel deals with it)
      if (!canonical) from_mont<FR_BAND>(k, k);
      NafPlan pl;
      naf_plan(pl, k);
      k_scalar_mul_plan<SUITE_BAND><<<cdiv(n, 128), 128, 0, st>>>(pl, d_in, (uint32_t)n, xy, w.den.as<Fe>(), canonical, in_is_dev);
      LAUNCHED("k_scalar_mul_plan");
      planned = true;
    }
  }
  if (!planned) {
    DISPATCH(suite, (k_scalar_mul_proj<S><<<cdiv(n, 128), 128, 0, st>>>(d_sk, sk_stride_words, d_in, (uint32_t)n, xy, w.den.as<Fe>(),
                                                                         canonical, in_is_dev, in_subgroup)));
    LAUNCHED("k_scalar_mul_proj");
  }
  if ((rc = batch_inv(suite, w.den.as<Fe>(), w.scr.as<Fe>(), n, st))) return rc;
  DISPATCH(suite, (k_affine_finish<S><<<cdiv(n, 128), 128, 0, st>>>(xy, w.den.as<Fe>(), n, d_out_dev, d_out_fmt, d_enc, nullptr, canonical)));
  LAUNCHED("k_affine_finish");
  return 0;
}

int avrf_hash_to_curve(uint32_t suite, uint32_t fmt, const uint8_t* msgs, const uint32_t* offsets, uint64_t n,
                       uint8_t* out_affine, uint8_t* out_compressed, uint8_t* ok) {
  if (suite > 2 || fmt > 1 || !offsets || (n && offsets[n] && !msgs)) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  if (n >= (1ull << 32)) return fail(AVRF_ERR_ARG, "too many messages for one call");
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& dm = fp.b[0]; DevBuf& doff = fp.b[1]; DevBuf& daff = fp.b[2]; DevBuf& dfmt = fp.b[3]; DevBuf& denc = fp.b[4]; DevBuf& dok = fp.b[5];
  H2cScratch& w = fp.w;
  int rc;
  if ((rc = dm.reserve(offsets[n] + 16)) || (rc = doff.reserve(4 * (n + 1))) || (rc = daff.reserve(64 * n)) ||
      (rc = dfmt.reserve(64 * n)) || (rc = denc.reserve(32 * n)) || (rc = dok.reserve(n)))
    return rc;
  if (offsets[n]) CK(cudaMemcpyAsync(dm.p, msgs, offsets[n], cudaMemcpyHostToDevice, gs()));
  CK(cudaMemcpyAsync(doff.p, offsets, 4 * (n + 1), cudaMemcpyHostToDevice, gs()));
  if ((rc = h2c_device(suite, dm.as<uint8_t>(), doff.as<uint32_t>(), n, daff.as<Affine>(), out_affine ? dfmt.as<Affine>() : nullptr,
                       out_compressed ? denc.as<uint32_t>() : nullptr, dok.as<uint8_t>(), fmt == AVRF_FMT_CANONICAL, w, gs())))
    return rc;
  if (out_affine) CK(cudaMemcpyAsync(out_affine, dfmt.p, 64 * n, cudaMemcpyDeviceToHost, gs()));
  if (out_compressed) CK(cudaMemcpyAsync(out_compressed, denc.p, 32 * n, cudaMemcpyDeviceToHost, gs()));
  if (ok) CK(cudaMemcpyAsync(ok, dok.p, n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaStreamSynchronize(gs()));
  return 0;
}

static int scalar_mul_impl(uint32_t suite, uint32_t fmt, const uint8_t* sk, uint32_t sk_stride, const uint8_t* inputs,
                           uint64_t n, uint8_t* outputs) {
  if (suite > 2 || fmt > 1 || !sk || !outputs || (sk_stride != 0 && sk_stride != 32)) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  if (n >= (1ull << 32)) return fail(AVRF_ERR_ARG, "too many items for one call");
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& dsk = fp.b[0]; DevBuf& din = fp.b[1]; DevBuf& dout = fp.b[2];
  H2cScratch& w = fp.w;
  int rc;
  size_t skb = sk_stride ? 32 * n : 32;
  if ((rc = dsk.reserve(skb)) || (rc = dout.reserve(64 * n))) return rc;
  if (inputs && (rc = din.reserve(64 * n))) return rc;
  CK(cudaMemcpyAsync(dsk.p, sk, skb, cudaMemcpyHostToDevice, gs()));
  if (inputs) CK(cudaMemcpyAsync(din.p, inputs, 64 * n, cudaMemcpyHostToDevice, gs()));
  // the generator is in the prime-order subgroup; caller-supplied inputs are taken as they are (src/lib.rs:391-393
  // multiplies whatever point it is given)
  if ((rc = smul_device(suite, dsk.as<Fe>(), sk_stride / 4, inputs ? din.as<Affine>() : nullptr, 0, inputs ? 0 : 1, n, nullptr,
                        dout.as<Affine>(), nullptr, fmt == AVRF_FMT_CANONICAL, w, gs(), sk_stride ? nullptr : sk)))
    return rc;
  CK(cudaMemcpyAsync(outputs, dout.p, 64 * n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaStreamSynchronize(gs()));
  return 0;
}

int avrf_vrf_output(uint32_t suite, uint32_t fmt, const uint8_t* sk, uint32_t sk_stride, const uint8_t* inputs,
                    uint64_t n, uint8_t* outputs) {
  if (!inputs) return fail(AVRF_ERR_ARG, "null inputs");
  return scalar_mul_impl(suite, fmt, sk, sk_stride, inputs, n, outputs);
}

int avrf_public_keys(uint32_t suite, uint32_t fmt, const uint8_t* sk, uint64_t n, uint8_t* pk) {
  return scalar_mul_impl(suite, fmt, sk, 32, nullptr, n, pk);
}

// Input::new + Secret::output (+ Output::hash) for n messages in one call: the points never leave the device
// between the two steps (BASELINE.json configs[4]).  Any of the three outputs may be NULL.
int avrf_vrf_io_many(uint32_t suite, uint32_t fmt, const uint8_t* msgs, const uint32_t* offsets, uint64_t n, const uint8_t* sk,
                     uint32_t sk_stride, uint8_t* out_inputs, uint8_t* out_outputs, uint8_t* out_hashes, uint8_t* ok) {
  if (suite > 2 || fmt > 1 || !offsets || !sk || (n && offsets[n] && !msgs) || (sk_stride != 0 && sk_stride != 32))
    return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  if (n >= (1ull << 32)) return fail(AVRF_ERR_ARG, "too many messages for one call");
  const int canonical = fmt == AVRF_FMT_CANONICAL;
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& dm = fp.b[0]; DevBuf& doff = fp.b[1]; DevBuf& dsk = fp.b[2]; DevBuf& din = fp.b[3]; DevBuf& dfmt = fp.b[4]; DevBuf& dout = fp.b[5]; DevBuf& dofmt = fp.b[6]; DevBuf& denc = fp.b[7]; DevBuf& dok = fp.b[8];
  H2cScratch& w = fp.w;
  int rc;
  size_t skb = sk_stride ? 32 * n : 32;
  if ((rc = dm.reserve(offsets[n] + 16)) || (rc = doff.reserve(4 * (n + 1))) || (rc = dsk.reserve(skb)) || (rc = din.reserve(64 * n)) ||
      (rc = dfmt.reserve(64 * n)) || (rc = dout.reserve(64 * n)) || (rc = dofmt.reserve(64 * n)) || (rc = denc.reserve(32 * n)) ||
      (rc = dok.reserve(n)))
    return rc;
  if (offsets[n]) CK(cudaMemcpyAsync(dm.p, msgs, offsets[n], cudaMemcpyHostToDevice, gs()));
  CK(cudaMemcpyAsync(doff.p, offsets, 4 * (n + 1), cudaMemcpyHostToDevice, gs()));
  CK(cudaMemcpyAsync(dsk.p, sk, skb, cudaMemcpyHostToDevice, gs()));
  if ((rc = h2c_device(suite, dm.as<uint8_t>(), doff.as<uint32_t>(), n, din.as<Affine>(), out_inputs ? dfmt.as<Affine>() : nullptr,
                       nullptr, dok.as<uint8_t>(), canonical, w, gs())))
    return rc;
  if (out_inputs) CK(cudaMemcpyAsync(out_inputs, dfmt.p, 64 * n, cudaMemcpyDeviceToHost, gs()));
  if ((rc = smul_device(suite, dsk.as<Fe>(), sk_stride / 4, din.as<Affine>(), 1, 1, n, out_hashes ? dout.as<Affine>() : nullptr,
                        out_outputs ? dofmt.as<Affine>() : nullptr, nullptr, canonical, w, gs(), sk_stride ? nullptr : sk)))
    return rc;
  if (out_outputs) CK(cudaMemcpyAsync(out_outputs, dofmt.p, 64 * n, cudaMemcpyDeviceToHost, gs()));
  if (out_hashes) {
    DISPATCH(suite, (k_compress<S><<<cdiv(n, 128), 128, 0, gs()>>>(dout.as<Affine>(), n, denc.as<uint32_t>(), 0, 1)));
    LAUNCHED("k_compress");
    CK(cudaMemcpyAsync(out_hashes, denc.p, 32 * n, cudaMemcpyDeviceToHost, gs()));
  }
  if (ok) CK(cudaMemcpyAsync(ok, dok.p, n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaStreamSynchronize(gs()));
  return 0;
}

int avrf_thin_prove_many(uint32_t suite, uint32_t fmt, uint64_t n, const uint8_t* sk, const uint8_t* pk,
                         const uint8_t* ios, const uint32_t* io_offsets, const uint8_t* ad_blob,
                         const uint32_t* ad_offsets, uint8_t* r, uint8_t* s) {
  if (suite > 2 || fmt > 1 || !sk || !pk || !io_offsets || !ad_offsets || !r || !s) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  size_t nio = io_offsets[n], nad = ad_offsets[n];
  if ((nio && !ios) || (nad && !ad_blob)) return fail(AVRF_ERR_ARG, "null argument");
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& dsk = fp.b[0]; DevBuf& dpk = fp.b[1]; DevBuf& dios = fp.b[2]; DevBuf& dio = fp.b[3]; DevBuf& dao = fp.b[4]; DevBuf& dad = fp.b[5]; DevBuf& dr = fp.b[6]; DevBuf& ds = fp.b[7];
  int rc;
  if ((rc = dsk.reserve(32 * n)) || (rc = dpk.reserve(64 * n)) || (rc = dios.reserve(128 * nio + 128)) ||
      (rc = dio.reserve(4 * (n + 1))) || (rc = dao.reserve(4 * (n + 1))) || (rc = dad.reserve(nad + 16)) ||
      (rc = dr.reserve(64 * n)) || (rc = ds.reserve(32 * n)))
    return rc;
  CK(cudaMemcpyAsync(dsk.p, sk, 32 * n, cudaMemcpyHostToDevice, gs()));
  CK(cudaMemcpyAsync(dpk.p, pk, 64 * n, cudaMemcpyHostToDevice, gs()));
  if (nio) CK(cudaMemcpyAsync(dios.p, ios, 128 * nio, cudaMemcpyHostToDevice, gs()));
  CK(cudaMemcpyAsync(dio.p, io_offsets, 4 * (n + 1), cudaMemcpyHostToDevice, gs()));
  CK(cudaMemcpyAsync(dao.p, ad_offsets, 4 * (n + 1), cudaMemcpyHostToDevice, gs()));
  if (nad) CK(cudaMemcpyAsync(dad.p, ad_blob, nad, cudaMemcpyHostToDevice, gs()));
  ProveArgs a;
  a.sk = dsk.as<Fe>(); a.pk = dpk.as<Affine>(); a.ios = dios.as<Affine>(); a.io_off = dio.as<uint32_t>();
  a.ad_off = dao.as<uint32_t>(); a.ad = dad.as<uint8_t>(); a.r = dr.as<Affine>(); a.s = ds.as<Fe>();
  a.n = (uint32_t)n; a.canonical = fmt == AVRF_FMT_CANONICAL;
  DISPATCH(suite, (k_prove<S><<<cdiv(n, 128), 128, 0, gs()>>>(a)));
  LAUNCHED("k_prove");
  CK(cudaMemcpyAsync(r, dr.p, 64 * n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaMemcpyAsync(s, ds.p, 32 * n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaStreamSynchronize(gs()));
  return 0;
}

static int compress_impl(uint32_t suite, uint32_t fmt, const uint8_t* points, uint64_t n, uint8_t* out32, int hash) {
  if (suite > 2 || fmt > 1 || !points || !out32) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& din = fp.b[0]; DevBuf& dout = fp.b[1];
  int rc;
  if ((rc = din.reserve(64 * n)) || (rc = dout.reserve(32 * n))) return rc;
  CK(cudaMemcpyAsync(din.p, points, 64 * n, cudaMemcpyHostToDevice, gs()));
  DISPATCH(suite, (k_compress<S><<<cdiv(n, 128), 128, 0, gs()>>>(din.as<Affine>(), n, dout.as<uint32_t>(),
                                                                     fmt == AVRF_FMT_CANONICAL, hash)));
  LAUNCHED("k_compress");
  CK(cudaMemcpyAsync(out32, dout.p, 32 * n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaStreamSynchronize(gs()));
  return 0;
}

int avrf_points_deserialize(uint32_t suite, uint32_t fmt, uint32_t kind, const uint8_t* in32, uint64_t n, uint8_t* out64,
                            uint8_t* ok) {
  if (suite > 2 || fmt > 1 || kind > 1 || !in32 || !out64 || !ok) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& din = fp.b[0]; DevBuf& dyn = fp.b[1]; DevBuf& dden = fp.b[2]; DevBuf& dscr = fp.b[3]; DevBuf& dfl = fp.b[4]; DevBuf& dout = fp.b[5]; DevBuf& dok = fp.b[6];
  int rc;
  if ((rc = din.reserve(32 * n)) || (rc = dyn.reserve(64 * n)) || (rc = dden.reserve(32 * n)) || (rc = dscr.reserve(32 * n)) ||
      (rc = dfl.reserve(n)) || (rc = dout.reserve(64 * n)) || (rc = dok.reserve(n)))
    return rc;
  CK(cudaMemcpyAsync(din.p, in32, 32 * n, cudaMemcpyHostToDevice, gs()));
  // y -> (1 - y^2, a - d y^2); one batched inversion; square root, sign, identity rule, subgroup test
  DISPATCH(suite, (k_dec_prep<S><<<cdiv(n, 128), 128, 0, gs()>>>(din.as<uint32_t>(), n, dyn.as<Fe>(), dden.as<Fe>(), dfl.as<uint8_t>())));
  LAUNCHED("k_dec_prep");
  if ((rc = batch_inv(suite, dden.as<Fe>(), dscr.as<Fe>(), n, gs()))) return rc;
  DISPATCH(suite, (k_dec_finish<S><<<cdiv(n, 128), 128, 0, gs()>>>(dyn.as<Fe>(), dden.as<Fe>(), dfl.as<uint8_t>(), n, (int)kind,
                                                                        dout.as<Affine>(), dok.as<uint8_t>(), fmt == AVRF_FMT_CANONICAL)));
  LAUNCHED("k_dec_finish");
  CK(cudaMemcpyAsync(out64, dout.p, 64 * n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaMemcpyAsync(ok, dok.p, n, cudaMemcpyDeviceToHost, gs()));
  CK(cudaStreamSynchronize(gs()));
  return 0;
}

// thin::BatchVerifier::push for proofs still in WIRE format (ark-serialize compressed, what `Proof::deserialize_compressed`,
// `Public` / `Input` / `Output` deserialisation take: src/thin.rs:42, src/lib.rs:410-433,471-494,552-575): the points are
// decoded and validated on the device (k_dec_prep -> k_batch_inv -> k_dec_finish) and handed to the push pipeline
// without leaving it.  All or nothing, like the reference where a proof that does not deserialize never reaches push.
int avrf_thin_batch_push_compressed(avrf_batch* b, uint64_t n, const uint8_t* pk32, const uint8_t* ios32,
                                    const uint32_t* io_offsets, const uint8_t* ad_blob, const uint32_t* ad_offsets,
                                    const uint8_t* r32, const uint8_t* s, uint8_t* ok, uint64_t* n_bad) {
  if (!b) return fail(AVRF_ERR_ARG, "null batch");
  if (b->scheme != 0) return fail(AVRF_ERR_ARG, "not a Thin-VRF batch");
  if (n_bad) *n_bad = 0;
  if (n == 0) return 0;
  if (!pk32 || !io_offsets || !ad_offsets || !r32 || !s) return fail(AVRF_ERR_ARG, "null argument");
  if (io_offsets[0] != 0 || ad_offsets[0] != 0) return fail(AVRF_ERR_ARG, "offsets must start at 0");
  const uint64_t nio = io_offsets[n];
  if ((nio && !ios32) || (ad_offsets[n] && !ad_blob)) return fail(AVRF_ERR_ARG, "null argument");
  if (n >= (1ull << 30) || nio >= (1ull << 30)) return fail(AVRF_ERR_ARG, "batch too large");
  ENTER(b);
  int rc = flush_pending(b);
  if (rc) return rc;
  const uint32_t suite = b->suite;
  const int canonical = b->fmt == AVRF_FMT_CANONICAL;
  // One PREP_CHUNK of proofs at a time: decode (this stream) -> push pipeline (H2D-stream copy, k_prepare, D2H of (c,s), the
  // hasher thread), so the batch-seed hash of chunk k runs under the decoding of chunk k+1.  The proofs are pushed
  // optimistically; a proof that does not decode rolls the whole call back at the end.
  size_t np_max = 0;
  for (uint64_t c0 = 0; c0 < n; c0 += PREP_CHUNK) {
    uint64_t c1 = std::min<uint64_t>(n, c0 + PREP_CHUNK);
    np_max = std::max<size_t>(np_max, 2 * (c1 - c0) + 2 * (size_t)(io_offsets[c1] - io_offsets[c0]));
  }
  const uint64_t n0 = b->n, i0 = b->n_ios, a0 = b->ad_bytes;
  FeedPool& fp = g_feed[b->device];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  cudaStream_t st = b->st_h2d;                           // the handle's own copy stream: ordered before its push pipeline
  DevBuf& din = fp.b[0]; DevBuf& dyn = fp.b[1]; DevBuf& dden = fp.b[2]; DevBuf& dscr = fp.b[3]; DevBuf& dfl = fp.b[4];
  DevBuf& dout = fp.b[5]; DevBuf& dok = fp.b[6]; DevBuf& dsc = fp.b[7]; DevBuf& dpo = fp.b[8]; DevBuf& doff = fp.b[9];
  const size_t bad_at = (n + 15) & ~(size_t)7;
  const size_t cmax = std::min<uint64_t>(n, PREP_CHUNK);
  if ((rc = din.reserve(32 * np_max)) || (rc = dyn.reserve(64 * np_max)) || (rc = dden.reserve(32 * np_max)) ||
      (rc = dscr.reserve(32 * np_max)) || (rc = dfl.reserve(np_max)) || (rc = dout.reserve(64 * np_max)) || (rc = dok.reserve(np_max)) ||
      (rc = dsc.reserve(32 * cmax)) || (rc = dpo.reserve(bad_at + 8)) || (rc = doff.reserve(4 * (cmax + 1))))
    return rc;
  if ((rc = b->h_small.reserve(4096))) return rc;
  unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(dpo.as<uint8_t>() + bad_at);
  CK(cudaMemsetAsync(d_bad, 0, 8, st));
  cudaEvent_t chunk_ev;
  CK(cudaEventCreateWithFlags(&chunk_ev, cudaEventDisableTiming));
  struct EvGuard { cudaEvent_t e; ~EvGuard() { cudaEventDestroy(e); } } ev_guard{chunk_ev};
  for (uint64_t c0 = 0; c0 < n; c0 += PREP_CHUNK) {
    const uint64_t c1 = std::min<uint64_t>(n, c0 + PREP_CHUNK), cnt = c1 - c0;
    const uint32_t q0 = io_offsets[c0], q1 = io_offsets[c1];
    const uint64_t npc = 2 * cnt + 2 * (uint64_t)(q1 - q0);       // order on the device: R (cnt) | pk (cnt) | I/O pairs
    CK(cudaMemcpyAsync(din.as<uint8_t>(), r32 + 32 * c0, 32 * cnt, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(din.as<uint8_t>() + 32 * cnt, pk32 + 32 * c0, 32 * cnt, cudaMemcpyHostToDevice, st));
    if (q1 > q0) CK(cudaMemcpyAsync(din.as<uint8_t>() + 64 * cnt, ios32 + 64 * (size_t)q0, 64 * (size_t)(q1 - q0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(doff.p, io_offsets + c0, 4 * (cnt + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dsc.p, s + 32 * c0, 32 * cnt, cudaMemcpyHostToDevice, st));
    DISPATCH(suite, (k_dec_prep<S><<<cdiv(npc, 128), 128, 0, st>>>(din.as<uint32_t>(), npc, dyn.as<Fe>(), dden.as<Fe>(), dfl.as<uint8_t>())));
    LAUNCHED("k_dec_prep");
    if ((rc = batch_inv(suite, dden.as<Fe>(), dscr.as<Fe>(), npc, st))) return rc;
    // R: bare AffinePoint (identity allowed, src/thin.rs:42); pk, I, O: Public / Input / Output (identity rejected)
    DISPATCH(suite, (k_dec_finish<S><<<cdiv(cnt, 128), 128, 0, st>>>(dyn.as<Fe>(), dden.as<Fe>(), dfl.as<uint8_t>(), cnt, 0,
                                                                      dout.as<Affine>(), dok.as<uint8_t>(), canonical)));
    LAUNCHED("k_dec_finish");
    DISPATCH(suite, (k_dec_finish<S><<<cdiv(npc - cnt, 128), 128, 0, st>>>(dyn.as<Fe>() + 2 * cnt, dden.as<Fe>() + cnt, dfl.as<uint8_t>() + cnt,
                                                                            npc - cnt, 1, dout.as<Affine>() + cnt, dok.as<uint8_t>() + cnt, canonical)));
    LAUNCHED("k_dec_finish");
    k_proof_ok<<<cdiv(cnt, 256), 256, 0, st>>>(dok.as<uint8_t>(), dok.as<uint8_t>() + cnt, dok.as<uint8_t>() + 2 * cnt, doff.as<uint32_t>(),
                                               q0, cnt, dpo.as<uint8_t>() + c0, d_bad);
    LAUNCHED("k_proof_ok");
    if (!canonical) {
      DISPATCH(suite, (k_scalars_to_mont<S><<<cdiv(cnt, 128), 128, 0, st>>>(dsc.as<Fe>(), cnt)));
      LAUNCHED("k_scalars_to_mont");
    }
    CK(cudaEventRecord(chunk_ev, st));                   // a handle outside its eager pipeline copies on the compute stream
    CK(cudaStreamWaitEvent(b->st, chunk_ev, 0));
    rc = push_many_impl(b, cnt, dout.as<uint8_t>() + 64 * cnt, dout.as<uint8_t>() + 128 * cnt, io_offsets + c0,
                        ad_blob ? ad_blob + ad_offsets[c0] : nullptr, ad_offsets + c0, dout.as<uint8_t>(), dsc.as<uint8_t>());
    if (rc) {                                            // a system error half way: leave the batch as it was before the call
      if (b->hasher) b->hasher->drain();
      quiesce(b);
      b->n = n0; b->n_ios = i0; b->ad_bytes = a0;
      b->prepared = b->have_seed = false;
      b->hashed = 0;
      b->push_launches = 0;
      b->prep_ev_chunks = 0;
      return rc;
    }
    // the pool's buffers are rewritten by the next chunk: only after this chunk's copies out of them (issued on the copy
    // stream by the pipeline, on the compute stream otherwise) are done
    CK(cudaEventRecord(chunk_ev, b->st));
    CK(cudaStreamWaitEvent(st, chunk_ev, 0));
  }
  unsigned long long* h_bad = reinterpret_cast<unsigned long long*>((uint8_t*)b->h_small.p + 2048);
  CK(cudaMemcpyAsync(h_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
  if (ok) CK(cudaMemcpyAsync(ok, dpo.p, n, cudaMemcpyDeviceToHost, st));
  CK(hsync(b, st));
  CK(hsync(b, b->prepared ? b->st_prep : b->st));        // the device has consumed the pool's buffers
  if (n_bad) *n_bad = *h_bad;
  if (*h_bad) {
    // roll back: nothing of this call stays in the batch (`ok` names the undecodable proofs).  What was already hashed
    // is dropped with it: the next verify re-derives the seed of the remaining proofs from the device stream.
    if (b->hasher) b->hasher->drain();
    if ((rc = quiesce(b))) return rc;
    b->n = n0;
    b->n_ios = i0;
    b->ad_bytes = a0;
    b->prepared = b->have_seed = false;
    b->hashed = 0;
    b->push_launches = 0;
    b->prep_ev_chunks = 0;
  }
  return 0;
}

int avrf_point_compress(uint32_t suite, uint32_t fmt, const uint8_t* points, uint64_t n, uint8_t* out32) {
  return compress_impl(suite, fmt, points, n, out32, 0);
}

int avrf_point_to_hash(uint32_t suite, uint32_t fmt, const uint8_t* points, uint64_t n, uint8_t* out32) {
  return compress_impl(suite, fmt, points, n, out32, 1);
}

int avrf_microbench(uint32_t kind, uint32_t iters, double* per_second, float* ms_out) {
  if (!per_second || iters == 0) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, g_device.load()));
  int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  FeedPool& fp = g_feed[g_device.load()];
  std::lock_guard<std::mutex> pool_lock(fp.mu);
  DevBuf& out = fp.b[0]; DevBuf& pts = fp.b[1];
  int rc;
  double work = 0;
  float ms = 0;
  if (kind == 0) {
    int blocks = sms * 8, threads = 256;
    if ((rc = out.reserve(8ull * blocks * threads))) return rc;
    k_mb_imad<<<blocks, threads, 0, gs()>>>(out.as<uint64_t>(), 16, 1);   // warm-up
    CK(cudaEventRecord(e0, gs()));
    k_mb_imad<<<blocks, threads, 0, gs()>>>(out.as<uint64_t>(), iters, 2);
    CK(cudaEventRecord(e1, gs()));
    work = (double)blocks * threads * iters * 32.0;
  } else if (kind == 1) {
    int blocks = sms * 16, threads = 128;
    if ((rc = out.reserve(32ull * blocks * threads))) return rc;
    k_mb_mul<<<blocks, threads, 0, gs()>>>(out.as<Fe>(), 4);
    CK(cudaEventRecord(e0, gs()));
    k_mb_mul<<<blocks, threads, 0, gs()>>>(out.as<Fe>(), iters);
    CK(cudaEventRecord(e1, gs()));
    work = (double)blocks * threads * iters * 2.0;
  } else if (kind == 2) {
    int blocks = sms * 16, threads = 128;
    uint32_t npts = 1u << 20;  // (bases are arbitrary field elements: the formulas do not care)
    if ((rc = out.reserve(128ull * blocks * threads)) || (rc = pts.reserve(128ull * npts))) return rc;
    CK(cudaMemsetAsync(pts.p, 0x11, 128ull * npts, gs()));
    k_mb_madd<<<blocks, threads, 0, gs()>>>(out.as<Ext>(), pts.as<BaseRec>(), npts, 2);
    CK(cudaEventRecord(e0, gs()));
    k_mb_madd<<<blocks, threads, 0, gs()>>>(out.as<Ext>(), pts.as<BaseRec>(), npts, iters);
    CK(cudaEventRecord(e1, gs()));
    work = (double)blocks * threads * iters;
  } else if (kind == 6 || kind == 7) {
    int blocks = sms * 8, threads = 256;
    if ((rc = out.reserve(8ull * blocks * threads))) return rc;
    if (kind == 6) {
      k_mb_imadc<0><<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, gs()));
      k_mb_imadc<0><<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), iters, 2);
    } else {
      k_mb_imadc<1><<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, gs()));
      k_mb_imadc<1><<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), iters, 2);
    }
    CK(cudaEventRecord(e1, gs()));
    work = (double)blocks * threads * iters * 32.0;
  } else if (kind == 3 || kind == 4) {
    int blocks = sms * 8, threads = 256;
    if ((rc = out.reserve(8ull * blocks * threads))) return rc;
    if (kind == 3) {
      k_mb_imadx<<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, gs()));
      k_mb_imadx<<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), iters, 2);
    } else {
      k_mb_imad32<<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, gs()));
      k_mb_imad32<<<blocks, threads, 0, gs()>>>(out.as<uint32_t>(), iters, 2);
    }
    CK(cudaEventRecord(e1, gs()));
    work = (double)blocks * threads * iters * 32.0;
  } else {
    return fail(AVRF_ERR_ARG, "unknown microbench kind");
  }
  LAUNCHED("microbench");
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  *per_second = work / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}


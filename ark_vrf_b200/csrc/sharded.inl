// One batch over several GPUs of ONE process (include/avrf.h, "Multi-GPU batches"): textually included at the end
// of avrf_gpu.cu.  The multi-GPU form of thin::BatchVerifier (reference src/thin.rs:198-325):
//   push    every call is cut into contiguous parts, one per device; each device runs H2D -> k_prepare -> D2H of
//           its (c_j, s_j) chunks, and ONE hasher thread absorbs the chunks of all devices in global proof order
//           (thin.rs:273-279: the serial SHA-512 runs once, on one core, while the devices work);
//   verify  finalise the seed, every device reduces its proofs to one partial point with weights addressed by
//           global proof index (thin.rs:289), k_fold stores the partial and the identity-gate flag straight into
//           a peer-mapped slot on device 0 (NVLink), device 0 adds the slots and writes the verdict.
// No collective library is involved: the only exchange is ndev x 144 bytes.
struct avrf_sharded {
  uint32_t suite = 0, fmt = 0;
  int ndev = 0;
  std::vector<avrf_batch*> sub;           // one handle per device, in g_dev_list order
  std::vector<bool> peer;                 // device d can store into device 0's memory
  std::vector<DevBuf*> local_slot;        // staging slot on devices without peer access
  std::vector<cudaEvent_t> done;
  cudaEvent_t init_ev = nullptr;
  std::unique_ptr<Hasher> hasher;
  uint64_t n = 0;
  DevBuf slots, out, flags;               // on device 0
  PinBuf h_small;
  cudaStream_t st0 = nullptr;
  uint8_t seed[64] = {};
  bool have_seed = false;
  avrf_sharded_timings tm = {};
};

__global__ void k_init_slots(ShardSlot* slots, uint32_t n, int suite) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ext e;
  if (suite == 0) ext_identity<0>(e);
  else if (suite == 1) ext_identity<1>(e);
  else ext_identity<2>(e);
  store_ext(&slots[i].partial, e);
  slots[i].gate = 0;
  slots[i].is_identity = 1;
}

extern "C" {

void avrf_thin_sharded_free(avrf_sharded* sh) {
  if (!sh) return;
  for (avrf_batch* b : sh->sub) if (b) { bind_device(b->device); quiesce(b); }
  sh->hasher.reset();
  for (avrf_batch* b : sh->sub) avrf_thin_batch_free(b);
  for (size_t d = 0; d < sh->local_slot.size(); d++)
    if (sh->local_slot[d]) { bind_device(g_dev_list[d]); delete sh->local_slot[d]; }
  bind_device(g_dev_list[0]);
  for (cudaEvent_t e : sh->done) if (e) cudaEventDestroy(e);
  if (sh->init_ev) cudaEventDestroy(sh->init_ev);
  if (sh->st0) { cudaStreamSynchronize(sh->st0); cudaStreamDestroy(sh->st0); }
  sh->slots.release(); sh->out.release(); sh->flags.release();
  delete sh;
}

avrf_sharded* avrf_thin_sharded_new(uint32_t suite, uint32_t fmt) {
  if (suite > 2 || fmt > 1) { fail(AVRF_ERR_ARG, "bad suite/fmt"); return nullptr; }
  if (ensure_init()) return nullptr;
  int nd = g_ndev.load();
  avrf_sharded* sh = new (std::nothrow) avrf_sharded();
  if (!sh) { fail(AVRF_ERR_NOMEM, "host allocation"); return nullptr; }
  sh->suite = suite;
  sh->fmt = fmt;
  sh->ndev = nd;
  int dev0 = g_dev_list[0];
  bool ok = bind_device(dev0) == 0 && sh->slots.reserve(sizeof(ShardSlot) * nd) == 0 && sh->out.reserve(sizeof(Ext)) == 0 &&
            sh->flags.reserve(64) == 0 && sh->h_small.reserve(256) == 0 &&
            cudaStreamCreateWithFlags(&sh->st0, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&sh->init_ev, cudaEventDisableTiming) == cudaSuccess;
  sh->hasher.reset(new (std::nothrow) Hasher(dev0));
  ok = ok && sh->hasher;
  sh->sub.assign(nd, nullptr);
  sh->peer.assign(nd, false);
  sh->local_slot.assign(nd, nullptr);
  sh->done.assign(nd, nullptr);
  for (int d = 0; ok && d < nd; d++) {
    int dev = g_dev_list[d];
    avrf_batch* b = avrf_thin_batch_new_on(dev, suite, fmt);
    if (!b) { ok = false; break; }
    sh->sub[d] = b;
    b->ext_hasher = sh->hasher.get();
    int can = dev == dev0;
    if (!can) cudaDeviceCanAccessPeer(&can, dev, dev0);
    sh->peer[d] = can != 0;
    if (can) {
      b->remote_slot = sh->slots.as<ShardSlot>() + d;
    } else {
      sh->local_slot[d] = new (std::nothrow) DevBuf();
      ok = sh->local_slot[d] && sh->local_slot[d]->reserve(sizeof(ShardSlot), 0, b->st) == 0;
      if (ok) b->remote_slot = sh->local_slot[d]->as<ShardSlot>();
    }
    ok = ok && cudaEventCreateWithFlags(&sh->done[d], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    if (g_err.empty()) fail(AVRF_ERR_CUDA, "avrf_thin_sharded_new", cudaGetErrorString(cudaGetLastError()));
    avrf_thin_sharded_free(sh);
    return nullptr;
  }
  return sh;
}

int avrf_thin_sharded_devices(const avrf_sharded* sh) { return sh ? sh->ndev : -1; }
int64_t avrf_thin_sharded_len(const avrf_sharded* sh) { return sh ? (int64_t)sh->n : -1; }
avrf_batch* avrf_thin_sharded_shard(avrf_sharded* sh, int index) {
  if (!sh || index < 0 || index >= sh->ndev) { fail(AVRF_ERR_ARG, "bad shard index"); return nullptr; }
  return sh->sub[index];
}

int avrf_thin_sharded_clear(avrf_sharded* sh) {
  if (!sh) return fail(AVRF_ERR_ARG, "null handle");
  sh->hasher->drain();
  for (avrf_batch* b : sh->sub) { int rc = avrf_thin_batch_clear(b); if (rc) return rc; }
  sh->n = 0;
  sh->have_seed = false;
  return 0;
}

int avrf_thin_sharded_push_many(avrf_sharded* sh, uint64_t n, const uint8_t* pk, const uint8_t* ios,
                                const uint32_t* io_offsets, const uint8_t* ad_blob, const uint32_t* ad_offsets,
                                const uint8_t* r, const uint8_t* s) {
  if (!sh) return fail(AVRF_ERR_ARG, "null handle");
  if (n == 0) return 0;
  if (!pk || !io_offsets || !ad_offsets || !r || !s) return fail(AVRF_ERR_ARG, "null argument");
  if (io_offsets[0] != 0 || ad_offsets[0] != 0) return fail(AVRF_ERR_ARG, "offsets must start at 0");
  if ((io_offsets[n] && !ios) || (ad_offsets[n] && !ad_blob)) return fail(AVRF_ERR_ARG, "null argument");
  int rc;
  if (sh->n == 0) {
    size_t sl;
    const unsigned char* sid = suite_id_of(sh->suite, &sl);
    unsigned char prefix[40];
    memcpy(prefix, sid, sl);
    prefix[sl] = DOM_BATCH;
    if ((rc = sh->hasher->begin(prefix, sl + 1))) return rc;
  }
  // contiguous parts, multiples of 32 proofs (a weight block is 4 proofs, a tree leaf 32)
  uint64_t per = ((n + sh->ndev - 1) / sh->ndev + 31) / 32 * 32;
  if ((rc = sh->hasher->reserve_stage(64 * std::min<uint64_t>(n, 16 * PREP_CHUNK) + 64 * sh->ndev * (n / PREP_CHUNK + 2)))) return rc;
  for (int d = 0; d < sh->ndev; d++) {
    uint64_t c0 = std::min<uint64_t>(n, d * per), c1 = std::min<uint64_t>(n, c0 + per);
    if (c1 == c0) continue;
    avrf_batch* b = sh->sub[d];
    if ((rc = enter(b))) return rc;
    // a further run of consecutive global indices on this device
    if (b->segs.size() >= 2 && b->segs[b->segs.size() - 1] + (b->n - b->segs[b->segs.size() - 2]) == sh->n + c0) {
      // contiguous with the previous run (single-device case): nothing to add
    } else {
      b->segs.push_back(b->n);
      b->segs.push_back(sh->n + c0);
      b->segs_dirty = true;
    }
    rc = push_many_impl(b, c1 - c0, pk + 64 * c0, ios ? ios + 128 * (size_t)io_offsets[c0] : nullptr, io_offsets + c0,
                        ad_blob ? ad_blob + ad_offsets[c0] : nullptr, ad_offsets + c0, r + 64 * c0, s + 32 * c0);
    if (rc) return rc;
  }
  sh->n += n;
  sh->have_seed = false;
  // the caller's arrays are borrowed for the duration of the call: wait until every device has consumed its part
  for (avrf_batch* b : sh->sub) {
    if ((rc = bind_device(b->device))) return rc;
    CK(hsync(b, b->st_h2d));
    CK(hsync(b, b->st_prep));
  }
  return 0;
}

int avrf_thin_sharded_verify(avrf_sharded* sh, int32_t* status) {
  if (!sh || !status) return fail(AVRF_ERR_ARG, "null argument");
  if (sh->n == 0) { *status = AVRF_OK; return 0; }                 // thin.rs:262-264
  auto t0 = std::chrono::steady_clock::now();
  int rc;
  if ((rc = sh->hasher->digest(sh->seed))) return rc;                // waits for the one hashing thread
  sh->have_seed = true;
  auto t1 = std::chrono::steady_clock::now();
  int dev0 = g_dev_list[0];
  if ((rc = bind_device(dev0))) return rc;
  k_init_slots<<<1, 32, 0, sh->st0>>>(sh->slots.as<ShardSlot>(), (uint32_t)sh->ndev, (int)sh->suite);
  LAUNCHED("k_init_slots");
  CK(cudaEventRecord(sh->init_ev, sh->st0));
  for (int d = 0; d < sh->ndev; d++) {
    avrf_batch* b = sh->sub[d];
    if (b->n == 0) continue;
    if ((rc = avrf_thin_batch_prepare(b, nullptr))) return rc;       // binds the device, orders st after the push pipeline
    CK(cudaStreamWaitEvent(b->st, sh->init_ev, 0));
    memcpy(b->seed, sh->seed, 64);
    b->have_seed = true;
    b->tm = avrf_timings{};
    if ((rc = run_msm(b, sh->seed, b->segs.size() >= 2 ? b->segs[1] : 0))) return rc;
    if (!sh->peer[d])
      CK(cudaMemcpyPeerAsync(sh->slots.as<ShardSlot>() + d, dev0, sh->local_slot[d]->p, b->device, sizeof(ShardSlot), b->st));
    CK(cudaEventRecord(sh->done[d], b->st));
  }
  if ((rc = bind_device(dev0))) return rc;
  for (int d = 0; d < sh->ndev; d++)
    if (sh->sub[d]->n) CK(cudaStreamWaitEvent(sh->st0, sh->done[d], 0));
  DISPATCH(sh->suite, (k_combine_shards<S><<<1, 32, 0, sh->st0>>>(sh->slots.as<ShardSlot>(), (uint32_t)sh->ndev, sh->out.as<Ext>(),
                                                                  sh->flags.as<int>())));
  LAUNCHED("k_combine_shards");
  CK(cudaMemcpyAsync(sh->h_small.p, sh->flags.p, 8, cudaMemcpyDeviceToHost, sh->st0));
  auto t2 = std::chrono::steady_clock::now();
  CK(cudaStreamSynchronize(sh->st0));
  auto t3 = std::chrono::steady_clock::now();
  const int* fl = sh->h_small.as<int>();
  if (fl[0] & 2) return fail(AVRF_ERR_ARG, "an input coordinate or scalar is not below its modulus (not a field element)");
  if (fl[0] & 1) *status = AVRF_INVALID_DATA;                        // thin.rs:266-271
  else *status = fl[1] ? AVRF_OK : AVRF_VERIFICATION_FAILURE;        // thin.rs:320-324
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<float, std::milli>(b - a).count();
  };
  sh->tm.hash_wait_ms = ms(t0, t1);
  sh->tm.host_hash_ms = sh->hasher->hash_ms();
  sh->tm.issue_ms = ms(t1, t2);
  sh->tm.device_wait_ms = ms(t2, t3);
  sh->tm.total_ms = ms(t0, t3);
  float worst = 0;
  for (int d = 0; d < sh->ndev; d++) {
    avrf_batch* b = sh->sub[d];
    if (!b->n) continue;
    bind_device(b->device);
    collect_timings(b, false);
    worst = std::max(worst, b->tm.scalars_ms + b->tm.sort_ms + b->tm.accumulate_ms + b->tm.reduce_ms);
  }
  sh->tm.shard_msm_ms_max = worst;
  return 0;
}

int avrf_thin_sharded_seed(const avrf_sharded* sh, uint8_t seed[64]) {
  if (!sh || !seed) return fail(AVRF_ERR_ARG, "null argument");
  if (!sh->have_seed) return fail(AVRF_ERR_STATE, "no seed yet: call verify first");
  memcpy(seed, sh->seed, 64);
  return 0;
}

int avrf_thin_sharded_timings(const avrf_sharded* sh, avrf_sharded_timings* out) {
  if (!sh || !out) return fail(AVRF_ERR_ARG, "null argument");
  *out = sh->tm;
  return 0;
}

}  // extern "C"

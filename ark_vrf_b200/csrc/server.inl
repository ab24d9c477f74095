// Batch server of libavrf_gpu.so (include/avrf.h, "Batch server"): textually included at the end of avrf_gpu.cu - it
// uses that file's handle type, error helpers and entry points and adds nothing to the device code.
// =========================================================================================
// Batch server: a pool of worker threads, one batch handle (own CUDA streams) each.
// The serial SHA-512 of every batch (thin.rs:273-279) then runs on its own core while the
// kernels of the workers share the GPU - the throughput mode of the verifier.
// =========================================================================================
struct ServerJob {
  int64_t ticket;
  uint64_t n;
  const uint8_t *pk, *ios, *ad_blob, *r, *s;
  const uint32_t *io_offsets, *ad_offsets;
};
struct ServerResult {
  int rc = 0;
  int32_t status = -1;
  std::string err;
};
struct avrf_server {
  uint32_t suite = 0, fmt = 0;
  std::vector<std::unique_ptr<MbSha512>> hashers;   // empty: every worker hashes its own batch (one core each)
  uint32_t n_own = 0;                               // workers 0 .. n_own-1 hash on their own thread even when hashers exist
  uint32_t idle_lane_workers = 0;                   // shared-lane workers waiting for a job (guarded by mu)
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::deque<ServerJob> queue;
  std::map<int64_t, ServerResult> done;
  int64_t next_ticket = 0;
  bool stop = false;
};

static void server_worker(avrf_server* sv, uint32_t index) {
  avrf_batch* h = nullptr;
  // Mixed pools: a job goes to an idle shared-lane worker first and to an own-thread worker only when none is idle.  In a
  // burst the pushes are staggered by their host-to-device copies: the early arrivals can afford the slower lanes, the
  // late ones - whose hash ends the burst - get a core of their own.
  const bool lane = !sv->hashers.empty() && index >= sv->n_own;
  for (;;) {
    ServerJob job;
    {
      std::unique_lock<std::mutex> lk(sv->mu);
      if (lane) sv->idle_lane_workers++;
      sv->cv_job.wait(lk, [&] { return sv->stop || (!sv->queue.empty() && (lane || sv->idle_lane_workers == 0)); });
      if (lane) sv->idle_lane_workers--;
      if (sv->queue.empty()) break;          // stop requested and nothing left to do
      job = sv->queue.front();
      sv->queue.pop_front();
      if (!sv->queue.empty()) sv->cv_job.notify_all();      // the own-thread workers re-evaluate
    }
    ServerResult res;
    if (!h) {
      h = avrf_thin_batch_new(sv->suite, sv->fmt);
      if (h) h->blocking = true;                         // many workers per core: sleep in waits, do not spin
      // shared multi-buffer hashing: the handle's hasher forwards its chunks to a lane of hasher i % n
      if (h && !sv->hashers.empty() && index >= sv->n_own) h->mb = sv->hashers[(index - sv->n_own) % sv->hashers.size()].get();
    }
    if (!h) {
      res.rc = AVRF_ERR_CUDA;
    } else {
      res.rc = avrf_thin_batch_clear(h);
      if (!res.rc) res.rc = avrf_thin_batch_push_many(h, job.n, job.pk, job.ios, job.io_offsets, job.ad_blob, job.ad_offsets, job.r, job.s);
      if (!res.rc) res.rc = avrf_thin_batch_verify(h, &res.status);
    }
    if (res.rc) res.err = g_err;             // this worker's thread-local message
    {
      std::lock_guard<std::mutex> lk(sv->mu);
      sv->done.emplace(job.ticket, std::move(res));
    }
    sv->cv_done.notify_all();
  }
  if (h) avrf_thin_batch_free(h);
}

extern "C" {

avrf_server* avrf_server_new(uint32_t suite, uint32_t fmt, uint32_t n_workers) {
  return avrf_server_new_ex(suite, fmt, n_workers, 0);
}

avrf_server* avrf_server_new_ex(uint32_t suite, uint32_t fmt, uint32_t n_workers, uint32_t n_hashers) {
  return avrf_server_new_mixed(suite, fmt, n_workers, n_hashers, 0);
}

avrf_server* avrf_server_new_mixed(uint32_t suite, uint32_t fmt, uint32_t n_workers, uint32_t n_hashers, uint32_t n_own) {
  if (suite > 2 || fmt > 1 || n_workers == 0 || n_workers > 256 || n_hashers > 64) { fail(AVRF_ERR_ARG, "bad suite/fmt/worker count"); return nullptr; }
  if (ensure_init()) return nullptr;
  avrf_server* sv = new (std::nothrow) avrf_server();
  if (!sv) { fail(AVRF_ERR_NOMEM, "host allocation"); return nullptr; }
  sv->suite = suite;
  sv->fmt = fmt;
  sv->n_own = n_own;
  try {
    for (uint32_t i = 0; i < n_hashers; i++) sv->hashers.emplace_back(new MbSha512());
    for (uint32_t i = 0; i < n_workers; i++) sv->workers.emplace_back(server_worker, sv, i);
  } catch (...) {
    fail(AVRF_ERR_NOMEM, "cannot start worker threads");
    avrf_server_free(sv);
    return nullptr;
  }
  return sv;
}

void avrf_server_free(avrf_server* sv) {
  if (!sv) return;
  {
    std::lock_guard<std::mutex> lk(sv->mu);
    sv->stop = true;
  }
  sv->cv_job.notify_all();
  for (auto& t : sv->workers) if (t.joinable()) t.join();   // queued batches are still verified
  delete sv;
}

int64_t avrf_server_submit(avrf_server* sv, uint64_t n, const uint8_t* pk, const uint8_t* ios, const uint32_t* io_offsets,
                           const uint8_t* ad_blob, const uint32_t* ad_offsets, const uint8_t* r, const uint8_t* s) {
  if (!sv) return fail(AVRF_ERR_ARG, "null server");
  if (n && (!pk || !io_offsets || !ad_offsets || !r || !s)) return fail(AVRF_ERR_ARG, "null argument");
  int64_t t;
  {
    std::lock_guard<std::mutex> lk(sv->mu);
    if (sv->stop) return fail(AVRF_ERR_STATE, "server is shutting down");
    t = sv->next_ticket++;
    sv->queue.push_back(ServerJob{t, n, pk, ios, ad_blob, r, s, io_offsets, ad_offsets});
  }
  sv->cv_job.notify_all();
  return t;
}

int avrf_server_wait(avrf_server* sv, int64_t ticket, int32_t* status) {
  if (!sv || !status) return fail(AVRF_ERR_ARG, "null argument");
  ServerResult res;
  {
    std::unique_lock<std::mutex> lk(sv->mu);
    if (ticket < 0 || ticket >= sv->next_ticket) return fail(AVRF_ERR_ARG, "unknown ticket");
    sv->cv_done.wait(lk, [&] { return sv->done.count(ticket) != 0; });
    res = std::move(sv->done[ticket]);
    sv->done.erase(ticket);
  }
  if (res.rc) return fail(res.rc, "batch server worker", res.err.c_str());
  *status = res.status;
  return 0;
}

// Host-only: SHA-512 of n independent streams through the multi-buffer hasher of the batch server, fed in
// interleaved `chunk`-byte updates (tests, diagnostics; needs no GPU).  *simd = 1 when the AVX-512 path ran.
int avrf_mb_sha512(uint32_t n_streams, const uint8_t* const* data, const uint64_t* lens, uint64_t chunk, uint8_t* digests,
                   int32_t* simd) {
  if ((n_streams && (!data || !lens || !digests)) || chunk == 0) return fail(AVRF_ERR_ARG, "bad argument");
  if (simd) *simd = MbSha512::simd_available() ? 1 : 0;
  MbSha512 mb;
  for (uint32_t g0 = 0; g0 < n_streams; g0 += MbSha512::LANES) {
    uint32_t g1 = std::min(n_streams, g0 + (uint32_t)MbSha512::LANES);
    int lane[MbSha512::LANES];
    for (uint32_t i = g0; i < g1; i++) {
      lane[i - g0] = mb.acquire();
      if (lane[i - g0] < 0) return fail(AVRF_ERR_STATE, "no free hash lane");
    }
    bool more = true;
    for (uint64_t off = 0; more; off += chunk) {
      more = false;
      for (uint32_t i = g0; i < g1; i++)
        if (off < lens[i]) {
          mb.update(lane[i - g0], data[i] + off, (size_t)std::min<uint64_t>(chunk, lens[i] - off));
          more = true;
        }
    }
    for (uint32_t i = g0; i < g1; i++) {
      mb.digest(lane[i - g0], digests + 64 * (size_t)i);
      mb.release(lane[i - g0]);
    }
  }
  return 0;
}

}  // extern "C"

// libavrf_gpu.so - host pipeline and C ABI (include/avrf.h) of the B200-native Thin-VRF
// batch verifier.  Single translation unit: all kernels are instantiated here for sm_100a
// (msm.cuh: Pippenger; prepare.cuh: per-proof transcripts; feeders.cuh: hash-to-curve, outputs,
// proving, ingest, per-proof verdicts; microbench.cuh: roofline probes).
//
// Path implemented (reference file:line):
//   push / prepare     src/thin.rs:209-243      -> k_prepare   (transcripts, z_i, c, point prep)
//   verify             src/thin.rs:257-325      -> host seed (serial SHA-512) + k_scalars + MSM kernels
//   Verifier::verify   src/thin.rs:131-165      -> batch of one
//   Input::new         src/lib.rs:500-502       -> k_h2c
//   Secret::output     src/lib.rs:391-393       -> k_output
//   Prover::prove      src/thin.rs:111-129      -> k_prove      (synthetic-input generator)
// There is no CPU fallback: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>
#include <openssl/evp.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/avrf.h"
#include "feeders.cuh"      // -> prepare.cuh -> msm.cuh -> thin.cuh, curve.cuh, fp.cuh, sha512.cuh
#include "microbench.cuh"
#include "mbsha512.h"

using namespace avrf;

// =========================================================================================
// Small host utilities
// =========================================================================================
static thread_local std::string g_err;
static std::atomic<int> g_device{-1};
// Stream of the handle-less entry points (hash-to-curve, outputs, proving, ingest, combine, microbenchmarks).
// Every batch handle owns its own four streams (struct avrf_batch), so handles driven from different host
// threads run concurrently on the device and never wait on each other's work.
static cudaStream_t g_stream = nullptr;
static cudaStream_t g_copy = nullptr;     // overlapped D2H inside avrf_thin_seed_dev
static int g_prio_hi = 0;
static std::mutex g_pin_mu;               // guards g_pin_stream (avrf_thin_seed_dev)

static int fail(int code, const char* what, const char* detail = "") {
  g_err = std::string(what) + (detail[0] ? ": " : "") + detail;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      char buf__[256];                                                                   \
      snprintf(buf__, sizeof buf__, "%s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return fail(e__ == cudaErrorMemoryAllocation ? AVRF_ERR_NOMEM : AVRF_ERR_CUDA, #call, buf__); \
    }                                                                                    \
  } while (0)

#define NEED_DEVICE()                                                                    \
  do {                                                                                   \
    int rc__ = ensure_init();                                                            \
    if (rc__) return rc__;                                                               \
  } while (0)

static int ensure_init() {
  if (g_device < 0) return avrf_init(0);
  static thread_local int bound = -1;      // a new host thread starts on device 0: bind it to the library's device
  int dev = g_device.load();
  if (bound != dev) {
    CK(cudaSetDevice(dev));
    bound = dev;
  }
  return 0;
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }       // temporaries in the entry points free their memory on every return path
  // Grow to at least `bytes`; keep the first `keep` bytes.
  int reserve(size_t bytes, size_t keep = 0, cudaStream_t st = nullptr) {
    if (bytes <= cap) return 0;
    if (!st) st = g_stream;
    size_t ncap = cap ? cap : 256;
    while (ncap < bytes) ncap += ncap / 2 + 256;
    void* q = nullptr;
    CK(cudaMalloc(&q, ncap));
    if (keep && p) CK(cudaMemcpyAsync(q, p, keep, cudaMemcpyDeviceToDevice, st));
    if (p) {
      CK(cudaStreamSynchronize(st));
      cudaFree(p);
    }
    p = q;
    cap = ncap;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  ~PinBuf() { release(); }
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

static inline uint32_t cdiv(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

// =========================================================================================
// Batch handle
// =========================================================================================
// proofs per k_prepare launch: one full wave of the 126-register kernel (148 SMs x 4 blocks x 128 threads; a 65536-proof
// chunk filled only 86 % of the block slots); 4.6 MiB of (c,s) stream per chunk
constexpr size_t PREP_CHUNK = 148 * 4 * 128;

struct avrf_batch {
  uint32_t suite = 0, fmt = 0, weights_mode = AVRF_WEIGHTS_REFERENCE;
  uint32_t scheme = 0;                  // 0 = Thin VRF, 1 = Pedersen VRF (src/pedersen.rs)
  uint64_t n = 0, n_ios = 0, ad_bytes = 0;
  bool prepared = false;
  bool have_seed = false;
  bool want_taps = false;
  uint8_t seed[64];
  uint64_t first_index = 0;             // global index of this handle's first proof in the last MSM run
  // inputs on the device
  DevBuf pk, r, s, ios, io_off, ad_off, ad;
  DevBuf ok, sb;                        // Pedersen only (pk holds the key commitments)
  // single-push staging on the host
  std::vector<uint8_t> h_pk, h_r, h_s, h_ios, h_ad;
  std::vector<uint32_t> h_io_off{0}, h_ad_off{0};
  // derived
  DevBuf pts, cs, z, renc, digits, hist, cursor, offs, toff, btot, totals, entries, tasks, task_out, chunk_out, wsum,
      partial, gpart, flags, w_tap, scalars_tap;
  PinBuf h_cs, h_small;
  std::vector<cudaEvent_t> prep_ev;     // one per PREP_CHUNK proofs: cs chunk i is ready
  // eager path: push = H2D + prepare + D2H + incremental SHA-512 of the batch transcript, pipelined
  bool eager = true;
  EVP_MD_CTX* hctx = nullptr;           // SHA-512 state after SUITE_ID || 0x50 || (c,s) of proofs [0, hashed)
  // batch-server handles: the same running hash kept by a shared multi-buffer hasher (mbsha512.h) instead of
  // hctx, and host waits that sleep instead of spinning (many worker threads per core)
  MbSha512* mb = nullptr;
  int mb_lane = -1;
  uint8_t mb_prefix[40] = {};
  bool blocking = false;
  cudaEvent_t sync_ev = nullptr;
  uint64_t hashed = 0;
  float push_hash_ms = 0, push_total_ms = 0;
  // the handle's own streams: compute + ordered copies; overlapped D2H of the (c,s) stream; chunked H2D of
  // pushed proofs; k_prepare of the eager push pipeline (high priority: it feeds the host hash)
  cudaStream_t st = nullptr, st_copy = nullptr, st_h2d = nullptr, st_prep = nullptr;
  // asynchronous verify: MSM enqueued on st, verdict read back at wait
  bool inflight = false;
  bool inflight_did_prepare = false;
  int32_t early_status = -1;            // >= 0: verdict known without waiting (empty batch)
  cudaEvent_t done_ev = nullptr;
  std::chrono::steady_clock::time_point t_verify0;
  cudaEvent_t ev[10] = {};
  avrf_timings tm = {};
};

// Wait for a stream of the handle: spinning (lowest latency) by default, sleeping for batch-server handles.
static cudaError_t hsync(avrf_batch* b, cudaStream_t st) {
  if (!b->blocking) return cudaStreamSynchronize(st);
  cudaError_t e;
  if (!b->sync_ev && (e = cudaEventCreateWithFlags(&b->sync_ev, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(b->sync_ev, st)) != cudaSuccess) return e;
  return cudaEventSynchronize(b->sync_ev);
}
static unsigned ev_flags(const avrf_batch* b) { return cudaEventDisableTiming | (b->blocking ? cudaEventBlockingSync : 0); }

static size_t npoints_of(const avrf_batch* b) { return b->scheme ? 5 * b->n + 2 : 2 * b->n + 2 * b->n_ios + 1; }
static size_t cs_stride(const avrf_batch* b) { return b->scheme ? 96 : 64; }

#define DISPATCH(suite, STMT)                      \
  switch (suite) {                                 \
    case 0: { constexpr int S = 0; STMT; } break;  \
    case 1: { constexpr int S = 1; STMT; } break;  \
    case 2: { constexpr int S = 2; STMT; } break;  \
    default: return fail(AVRF_ERR_ARG, "unknown suite"); \
  }

static int launch_check(const char* name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(AVRF_ERR_CUDA, name, cudaGetErrorString(e));
  return 0;
}
#define LAUNCHED(name) do { int rc__ = launch_check(name); if (rc__) return rc__; } while (0)

// =========================================================================================
// C ABI
// =========================================================================================
static const unsigned char* suite_id_of(uint32_t suite, size_t* len) {
  *len = CC_HOST[suite].sid_len;
  return CC_HOST[suite].suite_id;
}

extern "C" {

static int finish_inflight(avrf_batch* b);

const char* avrf_last_error(void) { return g_err.c_str(); }
const char* avrf_version(void) { return "ark-vrf_b200 0.1 (sm_100a)"; }

int avrf_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(AVRF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(AVRF_ERR_ARG, "device index out of range");
  static std::mutex init_mu;
  std::lock_guard<std::mutex> lock(init_mu);
  if (g_device == device && g_stream) { CK(cudaSetDevice(device)); return 0; }   // binds the calling thread too
  if (g_device >= 0 && g_device != device) return fail(AVRF_ERR_STATE, "already initialised on another device");
  CK(cudaSetDevice(device));
  if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  if (!g_copy) CK(cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking));
  int lo_prio = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &g_prio_hi));
  g_device = device;
  return 0;
}

int avrf_shutdown(void) {
  if (g_stream) cudaStreamDestroy(g_stream);
  if (g_copy) cudaStreamDestroy(g_copy);
  g_stream = g_copy = nullptr;
  g_device = -1;
  return 0;
}

avrf_batch* avrf_thin_batch_new(uint32_t suite, uint32_t fmt) {
  if (suite > 2 || fmt > 1) { fail(AVRF_ERR_ARG, "bad suite/fmt"); return nullptr; }
  if (ensure_init()) return nullptr;
  avrf_batch* b = new (std::nothrow) avrf_batch();
  if (!b) { fail(AVRF_ERR_NOMEM, "host allocation"); return nullptr; }
  b->suite = suite;
  b->fmt = fmt;
  for (auto& e : b->ev) cudaEventCreate(&e);
  if (cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&b->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&b->st_h2d, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&b->st_prep, cudaStreamNonBlocking, g_prio_hi) != cudaSuccess) {
    fail(AVRF_ERR_CUDA, "cudaStreamCreate", cudaGetErrorString(cudaGetLastError()));
    avrf_thin_batch_free(b);
    return nullptr;
  }
  return b;
}

void avrf_thin_batch_free(avrf_batch* b) {
  if (!b) return;
  if (b->inflight) finish_inflight(b);
  if (b->done_ev) cudaEventDestroy(b->done_ev);
  for (cudaStream_t q : {b->st, b->st_copy, b->st_h2d, b->st_prep}) if (q) cudaStreamSynchronize(q);
  DevBuf* bufs[] = {&b->ok, &b->sb, &b->pk, &b->r, &b->s, &b->ios, &b->io_off, &b->ad_off, &b->ad, &b->pts, &b->cs, &b->z, &b->renc,
                    &b->digits, &b->hist, &b->cursor, &b->offs, &b->toff, &b->btot, &b->totals, &b->entries, &b->tasks,
                    &b->task_out, &b->chunk_out, &b->wsum, &b->partial, &b->gpart, &b->flags, &b->w_tap, &b->scalars_tap};
  for (DevBuf* d : bufs) d->release();
  b->h_cs.release();
  b->h_small.release();
  for (auto& e : b->ev) if (e) cudaEventDestroy(e);
  for (auto& e : b->prep_ev) cudaEventDestroy(e);
  if (b->hctx) EVP_MD_CTX_free(b->hctx);
  if (b->mb && b->mb_lane >= 0) b->mb->release(b->mb_lane);
  if (b->sync_ev) cudaEventDestroy(b->sync_ev);
  for (cudaStream_t q : {b->st, b->st_copy, b->st_h2d, b->st_prep}) if (q) cudaStreamDestroy(q);
  delete b;
}

int avrf_thin_batch_clear(avrf_batch* b) {
  if (!b) return fail(AVRF_ERR_ARG, "null batch");
  if (b->inflight) { int rc = finish_inflight(b); if (rc) return rc; }
  b->n = b->n_ios = b->ad_bytes = 0;
  b->prepared = b->have_seed = false;
  b->hashed = 0;
  b->push_hash_ms = b->push_total_ms = 0;
  b->h_pk.clear(); b->h_r.clear(); b->h_s.clear(); b->h_ios.clear(); b->h_ad.clear();
  b->h_io_off.assign(1, 0);
  b->h_ad_off.assign(1, 0);
  return 0;
}

int avrf_thin_batch_invalidate(avrf_batch* b) {
  if (!b) return fail(AVRF_ERR_ARG, "null batch");
  b->prepared = b->have_seed = false;
  b->hashed = 0;
  return 0;
}

int avrf_thin_batch_set_eager(avrf_batch* b, int eager) {
  if (!b) return fail(AVRF_ERR_ARG, "null batch");
  b->eager = eager != 0;
  return 0;
}

void* avrf_stream(void) { return (void*)g_stream; }
void* avrf_thin_batch_stream(avrf_batch* b) { return b ? (void*)b->st : nullptr; }

int64_t avrf_thin_batch_len(const avrf_batch* b) { return b ? (int64_t)(b->n + b->h_io_off.size() - 1) : -1; }

int avrf_thin_batch_set_weights_mode(avrf_batch* b, uint32_t mode) {
  if (!b || mode > AVRF_WEIGHTS_TREE) return fail(AVRF_ERR_ARG, "bad weights mode");
  b->weights_mode = mode;
  b->have_seed = false;
  return 0;
}

// Leaf digests (64 B each) of this handle's (c,s) stream; first_index = global index of its first proof
// (must be a multiple of 32 unless it is 0).
int avrf_thin_batch_tree_leaves(avrf_batch* b, uint64_t first_index, uint8_t* out, uint64_t* n_leaves) {
  if (!b || !out || !n_leaves || (first_index % TREE_LEAF)) return fail(AVRF_ERR_ARG, "bad argument");
  int rc = avrf_thin_batch_prepare(b, nullptr);
  if (rc) return rc;
  uint32_t nl = (uint32_t)((b->n + TREE_LEAF - 1) / TREE_LEAF);
  *n_leaves = nl;
  if (!nl) return 0;
  if ((rc = b->gpart.reserve(64 * (size_t)nl + 64))) return rc;
  k_tree_leaves<<<cdiv(nl, 64), 64, 0, b->st>>>(b->cs.as<uint32_t>(), (uint32_t)b->n, first_index / TREE_LEAF,
                                                   b->gpart.as<uint64_t>());
  LAUNCHED("k_tree_leaves");
  CK(cudaMemcpyAsync(out, b->gpart.p, 64 * (size_t)nl, cudaMemcpyDeviceToHost, b->st));
  CK(hsync(b, b->st));
  return 0;
}

// seed = SHA512(SUITE_ID || 0x50 || 0x01 || LE64(n_total) || leaf_0 || leaf_1 || ...)
int avrf_thin_seed_tree(uint32_t suite, uint64_t n_total, const uint8_t* leaves, uint64_t n_leaves, uint8_t seed[64]) {
  if (suite > 2 || !seed || (n_leaves && !leaves)) return fail(AVRF_ERR_ARG, "bad argument");
  size_t sl;
  const unsigned char* sid = suite_id_of(suite, &sl);
  EVP_MD_CTX* ctx = EVP_MD_CTX_new();
  if (!ctx) return fail(AVRF_ERR_NOMEM, "EVP_MD_CTX_new");
  unsigned char tag[2] = {DOM_BATCH, 0x01}, le[8];
  for (int i = 0; i < 8; i++) le[i] = (unsigned char)(n_total >> (8 * i));
  unsigned int outl = 64;
  EVP_DigestInit_ex(ctx, EVP_sha512(), nullptr);
  EVP_DigestUpdate(ctx, sid, sl);
  EVP_DigestUpdate(ctx, tag, 2);
  EVP_DigestUpdate(ctx, le, 8);
  if (n_leaves) EVP_DigestUpdate(ctx, leaves, 64 * n_leaves);
  EVP_DigestFinal_ex(ctx, seed, &outl);
  EVP_MD_CTX_free(ctx);
  return 0;
}

int avrf_thin_batch_push(avrf_batch* b, const uint8_t pk[64], const uint8_t* ios, uint32_t n_ios, const uint8_t* ad,
                         uint32_t ad_len, const uint8_t r[64], const uint8_t s[32]) {
  if (!b || !pk || !r || !s || (n_ios && !ios) || (ad_len && !ad)) return fail(AVRF_ERR_ARG, "null argument");
  if (b->scheme != 0) return fail(AVRF_ERR_ARG, "not a Thin-VRF batch");
  b->h_pk.insert(b->h_pk.end(), pk, pk + 64);
  b->h_r.insert(b->h_r.end(), r, r + 64);
  b->h_s.insert(b->h_s.end(), s, s + 32);
  if (n_ios) b->h_ios.insert(b->h_ios.end(), ios, ios + 128 * (size_t)n_ios);
  if (ad_len) b->h_ad.insert(b->h_ad.end(), ad, ad + ad_len);
  b->h_io_off.push_back(b->h_io_off.back() + n_ios);
  b->h_ad_off.push_back(b->h_ad_off.back() + ad_len);
  return 0;
}

static int push_many_impl(avrf_batch* b, uint64_t n, const uint8_t* pk, const uint8_t* ios, const uint32_t* io_offsets,
                          const uint8_t* ad_blob, const uint32_t* ad_offsets, const uint8_t* r, const uint8_t* s,
                          const uint8_t* ok = nullptr, const uint8_t* sb = nullptr) {
  // thin: pk = public keys.  Pedersen (b->scheme == 1): pk = key commitments, plus ok (64 B) and sb (32 B) per proof.
  const bool ped = b->scheme == 1;
  const size_t stride = cs_stride(b);
  if (n == 0) return 0;
  uint64_t add_ios = io_offsets[n], add_ad = ad_offsets[n];
  uint64_t n0 = b->n, i0 = b->n_ios, a0 = b->ad_bytes;
  if (n0 + n >= (1ull << 30) || i0 + add_ios >= (1ull << 30) || a0 + add_ad >= (1ull << 32))
    return fail(AVRF_ERR_ARG, "batch too large");
  int rc;
  if ((rc = finish_inflight(b))) return rc;
  if ((rc = b->pk.reserve(64 * (n0 + n), 64 * n0, b->st))) return rc;
  if ((rc = b->r.reserve(64 * (n0 + n), 64 * n0, b->st))) return rc;
  if ((rc = b->s.reserve(32 * (n0 + n), 32 * n0, b->st))) return rc;
  if (ped && ((rc = b->ok.reserve(64 * (n0 + n), 64 * n0, b->st)) || (rc = b->sb.reserve(32 * (n0 + n), 32 * n0, b->st)))) return rc;
  if ((rc = b->ios.reserve(128 * (i0 + add_ios) + 128, 128 * i0, b->st))) return rc;
  if ((rc = b->ad.reserve(a0 + add_ad + 16, a0, b->st))) return rc;
  if ((rc = b->io_off.reserve(4 * (n0 + n + 1), 4 * (n0 + 1), b->st))) return rc;
  if ((rc = b->ad_off.reserve(4 * (n0 + n + 1), 4 * (n0 + 1), b->st))) return rc;
  auto tpush = std::chrono::steady_clock::now();
  // offsets first (small), rebased on the device
  bool pipeline = b->eager && (b->prepared || n0 == 0) && b->hashed == n0;
  cudaStream_t ost = pipeline ? b->st_prep : b->st;
  CK(cudaMemcpyAsync(b->io_off.as<uint32_t>() + n0, io_offsets, 4 * (n + 1), cudaMemcpyHostToDevice, ost));
  CK(cudaMemcpyAsync(b->ad_off.as<uint32_t>() + n0, ad_offsets, 4 * (n + 1), cudaMemcpyHostToDevice, ost));
  if (i0) { k_rebase<<<cdiv(n + 1, 256), 256, 0, ost>>>(b->io_off.as<uint32_t>() + n0, n + 1, (uint32_t)i0); LAUNCHED("k_rebase"); }
  if (a0) { k_rebase<<<cdiv(n + 1, 256), 256, 0, ost>>>(b->ad_off.as<uint32_t>() + n0, n + 1, (uint32_t)a0); LAUNCHED("k_rebase"); }
  if (!pipeline) {
    CK(cudaMemcpyAsync(b->pk.as<uint8_t>() + 64 * n0, pk, 64 * n, cudaMemcpyHostToDevice, b->st));
    CK(cudaMemcpyAsync(b->r.as<uint8_t>() + 64 * n0, r, 64 * n, cudaMemcpyHostToDevice, b->st));
    CK(cudaMemcpyAsync(b->s.as<uint8_t>() + 32 * n0, s, 32 * n, cudaMemcpyHostToDevice, b->st));
    if (ped) {
      CK(cudaMemcpyAsync(b->ok.as<uint8_t>() + 64 * n0, ok, 64 * n, cudaMemcpyHostToDevice, b->st));
      CK(cudaMemcpyAsync(b->sb.as<uint8_t>() + 32 * n0, sb, 32 * n, cudaMemcpyHostToDevice, b->st));
    }
    if (add_ios) CK(cudaMemcpyAsync(b->ios.as<uint8_t>() + 128 * i0, ios, 128 * add_ios, cudaMemcpyHostToDevice, b->st));
    if (add_ad) CK(cudaMemcpyAsync(b->ad.as<uint8_t>() + a0, ad_blob, add_ad, cudaMemcpyHostToDevice, b->st));
    b->n += n;
    b->n_ios += add_ios;
    b->ad_bytes += add_ad;
    b->prepared = b->have_seed = false;
    b->hashed = 0;
    return 0;
  }
  // ---- eager pipeline: per chunk  H2D (b->st_h2d) -> k_prepare (b->st_prep) -> D2H of (c,s) (b->st_copy) -> host SHA-512 ----
  size_t np_new = ped ? 5 * (n0 + n) + 2 : 2 * (n0 + n) + 2 * (i0 + add_ios) + 1;
  size_t np_old = n0 ? (ped ? 5 * n0 : 2 * n0 + 2 * i0) : 0;
  if ((rc = b->flags.reserve(64))) return rc;
  if ((rc = b->h_small.reserve(4096))) return rc;
  if ((rc = b->pts.reserve(sizeof(AffineK) * np_new, sizeof(AffineK) * np_old, b->st))) return rc;
  if ((rc = b->cs.reserve(stride * (n0 + n) + 64, stride * n0, b->st))) return rc;
  if ((rc = b->z.reserve(16 * (i0 + add_ios) + 16, 16 * i0, b->st))) return rc;
  if ((rc = b->renc.reserve(32 * (n0 + n) + 32, 32 * n0, b->st))) return rc;
  if ((rc = b->h_cs.reserve(stride * n + 64))) return rc;
  size_t sl;
  const unsigned char* sid = suite_id_of(b->suite, &sl);
  if (n0 == 0) {
    CK(cudaMemsetAsync(b->flags.p, 0, 64, b->st_prep));
    unsigned char tag = DOM_BATCH;
    if (b->mb) {
      b->mb->reset(b->mb_lane);
      memcpy(b->mb_prefix, sid, sl);
      b->mb_prefix[sl] = tag;
      b->mb->update(b->mb_lane, b->mb_prefix, sl + 1);
    } else {
      if (!b->hctx) b->hctx = EVP_MD_CTX_new();
      EVP_DigestInit_ex(b->hctx, EVP_sha512(), nullptr);
      EVP_DigestUpdate(b->hctx, sid, sl);
      EVP_DigestUpdate(b->hctx, &tag, 1);
    }
  }
  size_t nch = (n + PREP_CHUNK - 1) / PREP_CHUNK;
  while (b->prep_ev.size() < nch) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    b->prep_ev.push_back(e);
  }
  std::vector<cudaEvent_t> h2d_ev(nch), d2h_ev(nch);
  cudaEvent_t off_ev;
  CK(cudaEventCreateWithFlags(&off_ev, cudaEventDisableTiming));
  CK(cudaEventRecord(off_ev, b->st_prep));
  CK(cudaStreamWaitEvent(b->st_h2d, off_ev, 0));         // also orders after any device-side realloc copies
  PrepArgs a;
  PedPrepArgs pa;
  if (ped) {
    pa.pkcom = b->pk.as<Affine>(); pa.r = b->r.as<Affine>(); pa.ok = b->ok.as<Affine>(); pa.s = b->s.as<Fe>();
    pa.sb = b->sb.as<Fe>(); pa.ios = b->ios.as<Affine>(); pa.io_off = b->io_off.as<uint32_t>();
    pa.ad_off = b->ad_off.as<uint32_t>(); pa.ad = b->ad.as<uint8_t>(); pa.pts = b->pts.as<AffineK>();
    pa.cs = b->cs.as<uint32_t>(); pa.flags = b->flags.as<int>();
    pa.canonical = b->fmt == AVRF_FMT_CANONICAL;
  } else {
    a.pk = b->pk.as<Affine>(); a.r = b->r.as<Affine>(); a.s = b->s.as<Fe>(); a.ios = b->ios.as<Affine>();
    a.io_off = b->io_off.as<uint32_t>(); a.ad_off = b->ad_off.as<uint32_t>(); a.ad = b->ad.as<uint8_t>();
    a.pts = b->pts.as<AffineK>(); a.cs = b->cs.as<uint32_t>(); a.z = b->z.as<uint32_t>();
    a.renc = b->renc.as<uint32_t>(); a.flags = b->flags.as<int>();
    a.canonical = b->fmt == AVRF_FMT_CANONICAL;
  }
  for (size_t c = 0; c < nch; c++) {
    size_t c0 = c * PREP_CHUNK, c1 = std::min((size_t)n, c0 + PREP_CHUNK), cnt = c1 - c0;
    size_t q0 = io_offsets[c0], q1 = io_offsets[c1], d0 = ad_offsets[c0], d1 = ad_offsets[c1];
    CK(cudaEventCreateWithFlags(&h2d_ev[c], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d2h_ev[c], ev_flags(b)));
    CK(cudaMemcpyAsync(b->pk.as<uint8_t>() + 64 * (n0 + c0), pk + 64 * c0, 64 * cnt, cudaMemcpyHostToDevice, b->st_h2d));
    CK(cudaMemcpyAsync(b->r.as<uint8_t>() + 64 * (n0 + c0), r + 64 * c0, 64 * cnt, cudaMemcpyHostToDevice, b->st_h2d));
    CK(cudaMemcpyAsync(b->s.as<uint8_t>() + 32 * (n0 + c0), s + 32 * c0, 32 * cnt, cudaMemcpyHostToDevice, b->st_h2d));
    if (ped) {
      CK(cudaMemcpyAsync(b->ok.as<uint8_t>() + 64 * (n0 + c0), ok + 64 * c0, 64 * cnt, cudaMemcpyHostToDevice, b->st_h2d));
      CK(cudaMemcpyAsync(b->sb.as<uint8_t>() + 32 * (n0 + c0), sb + 32 * c0, 32 * cnt, cudaMemcpyHostToDevice, b->st_h2d));
    }
    if (q1 > q0) CK(cudaMemcpyAsync(b->ios.as<uint8_t>() + 128 * (i0 + q0), ios + 128 * q0, 128 * (q1 - q0), cudaMemcpyHostToDevice, b->st_h2d));
    if (d1 > d0) CK(cudaMemcpyAsync(b->ad.as<uint8_t>() + a0 + d0, ad_blob + d0, d1 - d0, cudaMemcpyHostToDevice, b->st_h2d));
    CK(cudaEventRecord(h2d_ev[c], b->st_h2d));
    CK(cudaStreamWaitEvent(b->st_prep, h2d_ev[c], 0));
    if (ped) {
      pa.first = (uint32_t)(n0 + c0);
      pa.n = (uint32_t)(n0 + c1);
      DISPATCH(b->suite, (k_prepare_ped<S><<<cdiv(cnt, 128), 128, 0, b->st_prep>>>(pa)));
    } else {
      a.first = (uint32_t)(n0 + c0);
      a.n = (uint32_t)(n0 + c1);
      DISPATCH(b->suite, (k_prepare<S><<<cdiv(cnt, 128), 128, 0, b->st_prep>>>(a)));
    }
    LAUNCHED("k_prepare");
    CK(cudaEventRecord(b->prep_ev[c], b->st_prep));
    CK(cudaStreamWaitEvent(b->st_copy, b->prep_ev[c], 0));
    CK(cudaMemcpyAsync((uint8_t*)b->h_cs.p + stride * c0, b->cs.as<uint8_t>() + stride * (n0 + c0), stride * cnt, cudaMemcpyDeviceToHost, b->st_copy));
    CK(cudaEventRecord(d2h_ev[c], b->st_copy));
  }
  auto th = std::chrono::steady_clock::now();
  for (size_t c = 0; c < nch; c++) {
    size_t c0 = c * PREP_CHUNK, c1 = std::min((size_t)n, c0 + PREP_CHUNK);
    CK(cudaEventSynchronize(d2h_ev[c]));
    if (b->mb) b->mb->update(b->mb_lane, (uint8_t*)b->h_cs.p + stride * c0, stride * (c1 - c0));
    else EVP_DigestUpdate(b->hctx, (uint8_t*)b->h_cs.p + stride * c0, stride * (c1 - c0));
    cudaEventDestroy(d2h_ev[c]);
    cudaEventDestroy(h2d_ev[c]);
  }
  cudaEventDestroy(off_ev);
  if (b->mb) b->mb->sync(b->mb_lane);      // the pinned (c,s) staging buffer is reused by the next push
  auto tend = std::chrono::steady_clock::now();
  b->push_hash_ms += std::chrono::duration<float, std::milli>(tend - th).count();
  b->push_total_ms += std::chrono::duration<float, std::milli>(tend - tpush).count();
  b->n += n;
  b->n_ios += add_ios;
  b->ad_bytes += add_ad;
  b->hashed = b->n;
  b->prepared = true;
  b->have_seed = false;
  b->tm.kernel_launches = nch;
  return 0;
}

static int flush_pending(avrf_batch* b) {
  uint64_t pend = b->h_io_off.size() - 1;
  if (!pend) return 0;
  int rc = push_many_impl(b, pend, b->h_pk.data(), b->h_ios.data(), b->h_io_off.data(), b->h_ad.data(),
                          b->h_ad_off.data(), b->h_r.data(), b->h_s.data());
  if (rc) return rc;
  CK(hsync(b, b->st));   // host vectors are about to be cleared
  b->h_pk.clear(); b->h_r.clear(); b->h_s.clear(); b->h_ios.clear(); b->h_ad.clear();
  b->h_io_off.assign(1, 0);
  b->h_ad_off.assign(1, 0);
  return 0;
}

int avrf_thin_batch_push_many(avrf_batch* b, uint64_t n, const uint8_t* pk, const uint8_t* ios,
                              const uint32_t* io_offsets, const uint8_t* ad_blob, const uint32_t* ad_offsets,
                              const uint8_t* r, const uint8_t* s) {
  if (!b) return fail(AVRF_ERR_ARG, "null batch");
  if (b->scheme != 0) return fail(AVRF_ERR_ARG, "not a Thin-VRF batch");
  if (n == 0) return 0;
  if (!pk || !io_offsets || !ad_offsets || !r || !s) return fail(AVRF_ERR_ARG, "null argument");
  if (io_offsets[0] != 0 || ad_offsets[0] != 0) return fail(AVRF_ERR_ARG, "offsets must start at 0");
  if ((io_offsets[n] && !ios) || (ad_offsets[n] && !ad_blob)) return fail(AVRF_ERR_ARG, "null argument");
  NEED_DEVICE();
  int rc = flush_pending(b);
  if (rc) return rc;
  rc = push_many_impl(b, n, pk, ios, io_offsets, ad_blob, ad_offsets, r, s);
  if (rc) return rc;
  // the caller's buffers are only borrowed for the duration of the call (thin.rs:218-225)
  CK(hsync(b, b->st));
  return 0;
}

int avrf_thin_batch_prepare(avrf_batch* b, int32_t* invalid) {
  if (!b) return fail(AVRF_ERR_ARG, "null batch");
  NEED_DEVICE();
  int rc = flush_pending(b);
  if (rc) return rc;
  if ((rc = b->flags.reserve(64))) return rc;
  if ((rc = b->h_small.reserve(4096))) return rc;
  if (!b->prepared) {
    size_t np = npoints_of(b);
    if ((rc = b->pts.reserve(sizeof(AffineK) * np))) return rc;
    if ((rc = b->cs.reserve(cs_stride(b) * b->n + 64))) return rc;
    if ((rc = b->z.reserve(16 * b->n_ios + 16))) return rc;
    if ((rc = b->renc.reserve(32 * b->n + 32))) return rc;
    CK(cudaMemsetAsync(b->flags.p, 0, 64, b->st));
    size_t nch = (b->n + PREP_CHUNK - 1) / PREP_CHUNK;
    while (b->prep_ev.size() < nch) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      b->prep_ev.push_back(e);
    }
    PedPrepArgs pa;
    PrepArgs ta;
    if (b->scheme == 1) {
      pa.pkcom = b->pk.as<Affine>(); pa.r = b->r.as<Affine>(); pa.ok = b->ok.as<Affine>(); pa.s = b->s.as<Fe>();
      pa.sb = b->sb.as<Fe>(); pa.ios = b->ios.as<Affine>(); pa.io_off = b->io_off.as<uint32_t>();
      pa.ad_off = b->ad_off.as<uint32_t>(); pa.ad = b->ad.as<uint8_t>(); pa.pts = b->pts.as<AffineK>();
      pa.cs = b->cs.as<uint32_t>(); pa.flags = b->flags.as<int>(); pa.n = (uint32_t)b->n;
      pa.canonical = b->fmt == AVRF_FMT_CANONICAL;
    } else {
      ta.pk = b->pk.as<Affine>(); ta.r = b->r.as<Affine>(); ta.s = b->s.as<Fe>(); ta.ios = b->ios.as<Affine>();
      ta.io_off = b->io_off.as<uint32_t>(); ta.ad_off = b->ad_off.as<uint32_t>(); ta.ad = b->ad.as<uint8_t>();
      ta.pts = b->pts.as<AffineK>(); ta.cs = b->cs.as<uint32_t>(); ta.z = b->z.as<uint32_t>();
      ta.renc = b->renc.as<uint32_t>(); ta.flags = b->flags.as<int>(); ta.n = (uint32_t)b->n;
      ta.canonical = b->fmt == AVRF_FMT_CANONICAL;
    }
    // one launch per chunk, an event after each: the D2H + host hash of chunk i start while chunk i+1 runs
    if (nch) cudaEventRecord(b->ev[0], b->st);
    for (size_t c = 0; c < nch; c++) {
      size_t cnt = std::min((size_t)PREP_CHUNK, (size_t)b->n - c * PREP_CHUNK);
      if (b->scheme == 1) {
        pa.first = (uint32_t)(c * PREP_CHUNK);
        DISPATCH(b->suite, (k_prepare_ped<S><<<cdiv(cnt, 128), 128, 0, b->st>>>(pa)));
      } else {
        ta.first = (uint32_t)(c * PREP_CHUNK);
        DISPATCH(b->suite, (k_prepare<S><<<cdiv(cnt, 128), 128, 0, b->st>>>(ta)));
      }
      LAUNCHED(b->scheme == 1 ? "k_prepare_ped" : "k_prepare");
      CK(cudaEventRecord(b->prep_ev[c], b->st));
    }
    if (nch) {
      cudaEventRecord(b->ev[1], b->st);
      b->tm.kernel_launches = nch;
    }
    b->prepared = true;
    b->have_seed = false;
  }
  if (invalid) {
    CK(cudaMemcpyAsync(b->h_small.p, b->flags.p, 8, cudaMemcpyDeviceToHost, b->st));
    CK(hsync(b, b->st));
    int fl = reinterpret_cast<int*>(b->h_small.p)[0];
    if (fl & 2) return fail(AVRF_ERR_ARG, "an input coordinate or scalar is not below its modulus (not a field element)");
    *invalid = fl & 1;
  }
  return 0;
}

int avrf_thin_batch_cs_stream(avrf_batch* b, uint8_t* out) {
  if (!b || !out) return fail(AVRF_ERR_ARG, "null argument");
  int rc = avrf_thin_batch_prepare(b, nullptr);
  if (rc) return rc;
  if (b->n) CK(cudaMemcpyAsync(out, b->cs.p, cs_stride(b) * b->n, cudaMemcpyDeviceToHost, b->st));
  CK(hsync(b, b->st));
  return 0;
}

void* avrf_thin_batch_cs_dev(avrf_batch* b) {
  if (!b) { fail(AVRF_ERR_ARG, "null batch"); return nullptr; }
  if (avrf_thin_batch_prepare(b, nullptr)) return nullptr;
  if (hsync(b, b->st) != cudaSuccess) return nullptr;
  return b->cs.p;
}

int avrf_thin_seed(uint32_t suite, const uint8_t* cs_stream, uint64_t n_items, uint8_t seed[64]) {
  if (suite > 2 || !seed || (n_items && !cs_stream)) return fail(AVRF_ERR_ARG, "bad argument");
  size_t sl;
  const unsigned char* sid = suite_id_of(suite, &sl);
  EVP_MD_CTX* ctx = EVP_MD_CTX_new();
  if (!ctx) return fail(AVRF_ERR_NOMEM, "EVP_MD_CTX_new");
  unsigned char tag = DOM_BATCH;
  unsigned int outl = 64;
  EVP_DigestInit_ex(ctx, EVP_sha512(), nullptr);
  EVP_DigestUpdate(ctx, sid, sl);
  EVP_DigestUpdate(ctx, &tag, 1);
  if (n_items) EVP_DigestUpdate(ctx, cs_stream, 64 * n_items);
  EVP_DigestFinal_ex(ctx, seed, &outl);
  EVP_MD_CTX_free(ctx);
  return 0;
}

// Device->host copy of a (c,s) stream in chunks on the copy stream, each chunk hashed on the host
// as soon as it lands: the serial SHA-512 of thin.rs:273-279.
static int seed_of_device_stream(cudaStream_t st, cudaStream_t st_copy, uint32_t suite, const uint8_t* cs_dev,
                                 size_t total, PinBuf& pin, uint8_t seed[64], float* hash_ms,
                                 const std::vector<cudaEvent_t>* chunk_ready = nullptr, size_t stride = 64) {
  int rc;
  if ((rc = pin.reserve(total + 64))) return rc;
  const size_t CH = stride * PREP_CHUNK;  // one k_prepare chunk: 4.6 MiB (thin) / 6.9 MiB (pedersen)
  size_t nch = (total + CH - 1) / CH;
  std::vector<cudaEvent_t> evs(nch);
  cudaEvent_t ready;
  CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  if (!chunk_ready) {
    CK(cudaEventRecord(ready, st));
    CK(cudaStreamWaitEvent(st_copy, ready, 0));
  }
  for (size_t i = 0; i < nch; i++) {
    size_t off = i * CH, len = std::min(CH, total - off);
    if (chunk_ready) CK(cudaStreamWaitEvent(st_copy, (*chunk_ready)[i], 0));
    CK(cudaMemcpyAsync((uint8_t*)pin.p + off, cs_dev + off, len, cudaMemcpyDeviceToHost, st_copy));
    CK(cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming));
    CK(cudaEventRecord(evs[i], st_copy));
  }
  auto t0 = std::chrono::steady_clock::now();
  size_t sl;
  const unsigned char* sid = suite_id_of(suite, &sl);
  EVP_MD_CTX* ctx = EVP_MD_CTX_new();
  unsigned char tag = DOM_BATCH;
  unsigned int outl = 64;
  EVP_DigestInit_ex(ctx, EVP_sha512(), nullptr);
  EVP_DigestUpdate(ctx, sid, sl);
  EVP_DigestUpdate(ctx, &tag, 1);
  for (size_t i = 0; i < nch; i++) {
    size_t off = i * CH, len = std::min(CH, total - off);
    cudaEventSynchronize(evs[i]);
    EVP_DigestUpdate(ctx, (uint8_t*)pin.p + off, len);
    cudaEventDestroy(evs[i]);
  }
  EVP_DigestFinal_ex(ctx, seed, &outl);
  EVP_MD_CTX_free(ctx);
  cudaEventDestroy(ready);
  if (hash_ms) *hash_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

static int seed_from_device(avrf_batch* b) {
  int rc = seed_of_device_stream(b->st, b->st_copy, b->suite, b->cs.as<uint8_t>(), cs_stride(b) * b->n, b->h_cs, b->seed,
                                 &b->tm.host_hash_ms, &b->prep_ev, cs_stride(b));
  if (rc) return rc;
  b->have_seed = true;
  return 0;
}

static PinBuf g_pin_stream;

static void seed_to_words(Seed64& sd, const uint8_t seed[64]) {
  for (int i = 0; i < 8; i++) {
    uint64_t w = 0;
    for (int k = 0; k < 8; k++) w = (w << 8) | seed[8 * i + k];
    sd.w[i] = w;
  }
}

// Everything after the seed: scalars, sort, accumulate, reduce.  Leaves the partial point in
// b->partial and flags[1].
static int run_msm(avrf_batch* b, const uint8_t seed[64], uint64_t first_index) {
  int rc;
  size_t np = npoints_of(b);
  if (np >= (1ull << 28)) return fail(AVRF_ERR_ARG, "batch too large for one handle (2^28 MSM terms): shard it");
  size_t max_entries = np * MSM_NWIN;
  uint32_t nblk = cdiv(b->n, 128);
  // segment length: ~450k segments (6 waves of 148 SMs x 512 threads), between 8 and 128 entries
  uint32_t lshift = 3;
  while (lshift < 7 && (np * 14) >> (lshift + 1) >= 450000) lshift++;
  size_t max_segs = (max_entries >> lshift) + 1;
  size_t max_slots = max_segs + MSM_NBINS + 1;
  if ((rc = b->digits.reserve(32 * np))) return rc;
  if ((rc = b->hist.reserve(4 * MSM_NBINS))) return rc;
  if ((rc = b->cursor.reserve(64 * np))) return rc;                  // ranks: 16 x u32 per point
  if ((rc = b->offs.reserve(4 * (MSM_NBINS + 1)))) return rc;
  if ((rc = b->toff.reserve(4 * (MSM_NBINS + 1)))) return rc;       // nzr: rank among non-empty bins
  if ((rc = b->btot.reserve(4 * 1024))) return rc;
  if ((rc = b->totals.reserve(64))) return rc;
  if ((rc = b->entries.reserve(4 * max_entries))) return rc;
  if ((rc = b->tasks.reserve(4 * (size_t)MSM_NBINS))) return rc;    // list of bins with many partial sums
  if ((rc = b->task_out.reserve(sizeof(Ext) * max_slots))) return rc;
  if ((rc = b->chunk_out.reserve(sizeof(Ext) * MSM_NWIN * MSM_NCHUNK))) return rc;
  if ((rc = b->wsum.reserve(sizeof(Ext) * MSM_NWIN))) return rc;
  if ((rc = b->partial.reserve(sizeof(Ext)))) return rc;
  if ((rc = b->gpart.reserve(80 * (size_t)nblk + 80))) return rc;
  if (b->want_taps) {
    if ((rc = b->w_tap.reserve(32 * b->n + 32))) return rc;
    if ((rc = b->scalars_tap.reserve(32 * np))) return rc;
  }
  cudaStream_t st = b->st;
  CK(cudaMemsetAsync(b->hist.p, 0, 4 * MSM_NBINS, st));

  ScalArgs a;
  a.cs = b->cs.as<uint32_t>(); a.z = b->z.as<uint32_t>(); a.io_off = b->io_off.as<uint32_t>();
  a.digits = b->digits.as<uint4>(); a.hist = b->hist.as<uint32_t>(); a.gpart = b->gpart.as<uint32_t>();
  a.ranks = b->cursor.as<uint4>();
  a.w_tap = b->want_taps ? b->w_tap.as<uint32_t>() : nullptr;
  a.scalars_tap = b->want_taps ? b->scalars_tap.as<Fe>() : nullptr;
  seed_to_words(a.seed, seed);
  a.first_index = first_index;
  b->first_index = first_index;
  a.n = (uint32_t)b->n;
  uint32_t* hist = b->hist.as<uint32_t>();
  uint32_t* offs = b->offs.as<uint32_t>();
  uint32_t* nzr = b->toff.as<uint32_t>();
  uint32_t* totals = b->totals.as<uint32_t>();
  Ext* slots = b->task_out.as<Ext>();
  cudaEventRecord(b->ev[2], st);
  if (b->scheme == 1) {
    PedScalArgs pa;
    pa.cs = a.cs; pa.digits = a.digits; pa.ranks = a.ranks; pa.hist = a.hist; pa.gpart = a.gpart;
    pa.w_tap = b->want_taps ? b->w_tap.as<uint32_t>() : nullptr;
    pa.scalars_tap = a.scalars_tap; pa.seed = a.seed; pa.first_index = first_index; pa.n = a.n;
    DISPATCH(b->suite, (k_scalars_ped<S><<<nblk, 128, 0, st>>>(pa)));
    LAUNCHED("k_scalars_ped");
    DISPATCH(b->suite, (k_gscalar_ped<S><<<1, 32, 0, st>>>(a.gpart, nblk, a.digits, a.ranks, a.hist, a.scalars_tap,
                                                            b->pts.as<AffineK>(), np - 2)));
    LAUNCHED("k_gscalar_ped");
  } else {
    DISPATCH(b->suite, (k_scalars<S><<<nblk, 128, 0, st>>>(a)));
    LAUNCHED("k_scalars");
    DISPATCH(b->suite, (k_gscalar<S><<<1, 256, 0, st>>>(a.gpart, nblk, a.digits, a.ranks, a.hist, a.scalars_tap,
                                                         b->pts.as<AffineK>(), np - 1)));
    LAUNCHED("k_gscalar");
  }
  cudaEventRecord(b->ev[3], st);
  k_scan_local<<<MSM_NBINS / 1024, 1024, 0, st>>>(hist, offs, nzr, b->btot.as<uint32_t>());
  LAUNCHED("k_scan_local");
  k_scan_totals<<<1, 512, 0, st>>>(b->btot.as<uint32_t>(), totals, offs);
  LAUNCHED("k_scan_totals");
  k_scan_add<<<MSM_NBINS / 1024, 1024, 0, st>>>(offs, nzr, b->btot.as<uint32_t>());
  LAUNCHED("k_scan_add");
  k_scatter<<<cdiv(np, 256), 256, 0, st>>>(b->digits.as<uint4>(), b->cursor.as<uint4>(), offs,
                                           b->entries.as<uint32_t>(), np);
  LAUNCHED("k_scatter");
  cudaEventRecord(b->ev[4], st);
  AccArgs ac;
  ac.entries = b->entries.as<uint32_t>(); ac.offs = offs; ac.hist = hist; ac.nzr = nzr; ac.totals = totals;
  ac.pts = b->pts.as<AffineK>(); ac.slots = slots; ac.lshift = lshift;
  DISPATCH(b->suite, (k_accumulate<S, 5><<<cdiv(max_segs, 128), 128, 0, st>>>(ac)));
  LAUNCHED("k_accumulate");
  cudaEventRecord(b->ev[5], st);
  DISPATCH(b->suite, (k_combine<S><<<MSM_NBINS / 128, 128, 0, st>>>(hist, offs, nzr, slots, lshift, totals,
                                                                    b->tasks.as<uint32_t>())));
  LAUNCHED("k_combine");
  DISPATCH(b->suite, (k_combine_big<S><<<64, 256, 0, st>>>(hist, offs, nzr, slots, lshift, totals,
                                                           b->tasks.as<uint32_t>())));
  LAUNCHED("k_combine_big");
  DISPATCH(b->suite, (k_bucket_reduce<S><<<MSM_NWIN * MSM_NCHUNK / 128, 128, 0, st>>>(hist, offs, nzr, lshift, slots,
                                                                                      b->chunk_out.as<Ext>())));
  LAUNCHED("k_bucket_reduce");
  DISPATCH(b->suite, (k_window_sum<S><<<MSM_NWIN, 256, 0, st>>>(b->chunk_out.as<Ext>(), b->wsum.as<Ext>())));
  LAUNCHED("k_window_sum");
  DISPATCH(b->suite, (k_fold<S><<<1, 32, 0, st>>>(b->wsum.as<Ext>(), b->partial.as<Ext>(), b->flags.as<int>())));
  LAUNCHED("k_fold");
  cudaEventRecord(b->ev[6], st);
  b->tm.kernel_launches += 12;
  b->tm.n_points = np;
  return 0;
}

static void collect_timings(avrf_batch* b, bool with_prepare) {
  float ms = 0;
  if (with_prepare && cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]) == cudaSuccess) b->tm.prepare_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev[2], b->ev[3]) == cudaSuccess) b->tm.scalars_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev[3], b->ev[4]) == cudaSuccess) b->tm.sort_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev[4], b->ev[5]) == cudaSuccess) b->tm.accumulate_ms = ms;
  if (cudaEventElapsedTime(&ms, b->ev[5], b->ev[6]) == cudaSuccess) b->tm.reduce_ms = ms;
  (void)cudaGetLastError();   // an event pair that was never recorded (prepare done at push time) is not an error
}

int avrf_thin_seed_dev(uint32_t suite, const void* cs_stream_dev, uint64_t n_items, uint8_t seed[64]) {
  if (suite > 2 || !seed || (n_items && !cs_stream_dev)) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  std::lock_guard<std::mutex> lock(g_pin_mu);
  return seed_of_device_stream(g_stream, g_copy, suite, (const uint8_t*)cs_stream_dev, 64 * n_items, g_pin_stream, seed, nullptr);
}

int avrf_thin_batch_partial(avrf_batch* b, const uint8_t seed[64], uint64_t first_index, uint8_t partial[128]) {
  if (!b || !seed || !partial) return fail(AVRF_ERR_ARG, "null argument");
  int rc = avrf_thin_batch_prepare(b, nullptr);
  if (rc) return rc;
  memcpy(b->seed, seed, 64);
  b->have_seed = true;
  if ((rc = run_msm(b, seed, first_index))) return rc;
  CK(cudaMemcpyAsync(b->h_small.p, b->partial.p, 128, cudaMemcpyDeviceToHost, b->st));
  CK(cudaMemcpyAsync((uint8_t*)b->h_small.p + 128, b->totals.p, 8, cudaMemcpyDeviceToHost, b->st));
  CK(hsync(b, b->st));
  memcpy(partial, b->h_small.p, 128);
  b->tm.n_entries = reinterpret_cast<uint32_t*>((uint8_t*)b->h_small.p + 128)[0];
  b->tm.n_tasks = reinterpret_cast<uint32_t*>((uint8_t*)b->h_small.p + 128)[1];
  collect_timings(b, true);
  return 0;
}

int avrf_thin_combine_partials(uint32_t suite, const uint8_t* partials, uint32_t n, int32_t* status) {
  if (suite > 2 || !partials || !status || n == 0) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  DevBuf in, out, flags;
  int rc;
  if ((rc = in.reserve(128 * (size_t)n)) || (rc = out.reserve(128)) || (rc = flags.reserve(64))) return rc;
  CK(cudaMemcpyAsync(in.p, partials, 128 * (size_t)n, cudaMemcpyHostToDevice, g_stream));
  DISPATCH(suite, (k_combine_partials<S><<<1, 32, 0, g_stream>>>(in.as<Ext>(), n, out.as<Ext>(), flags.as<int>())));
  LAUNCHED("k_combine_partials");
  int h[2] = {0, 0};
  CK(cudaMemcpyAsync(h, flags.p, 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  *status = h[1] ? AVRF_OK : AVRF_VERIFICATION_FAILURE;
  in.release(); out.release(); flags.release();
  return 0;
}

// verify = verify_async + verify_wait.  verify_async enqueues everything up to the device->host copy of the
// verdict words and returns; with the eager push pipeline it does not block on the GPU at all, so the
// next batch's push (host SHA-512 + its own prepare on g_prep) overlaps this batch's MSM on g_stream.
int avrf_thin_batch_verify_async(avrf_batch* b) {
  if (!b) return fail(AVRF_ERR_ARG, "null argument");
  NEED_DEVICE();
  int rc = finish_inflight(b);
  if (rc) return rc;
  b->t_verify0 = std::chrono::steady_clock::now();
  if ((rc = flush_pending(b))) return rc;
  b->early_status = -1;
  if (b->n == 0) { b->early_status = AVRF_OK; b->inflight = true; return 0; }   // thin.rs:262-264
  bool did_prepare = !b->prepared;
  uint64_t push_launches = b->tm.kernel_launches;
  b->tm = avrf_timings{};
  if (!did_prepare) b->tm.kernel_launches = push_launches;
  if ((rc = avrf_thin_batch_prepare(b, nullptr))) return rc;
  if (b->weights_mode == AVRF_WEIGHTS_TREE) {
    uint64_t nl = 0;
    if ((rc = b->h_cs.reserve(64 * ((b->n + TREE_LEAF - 1) / TREE_LEAF) + 64))) return rc;
    auto th = std::chrono::steady_clock::now();
    if ((rc = avrf_thin_batch_tree_leaves(b, 0, (uint8_t*)b->h_cs.p, &nl))) return rc;
    if ((rc = avrf_thin_seed_tree(b->suite, b->n, (const uint8_t*)b->h_cs.p, nl, b->seed))) return rc;
    b->tm.host_hash_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - th).count();
    b->have_seed = true;
  } else if (b->eager && (b->hctx || b->mb) && b->hashed == b->n && !did_prepare) {
    // every (c_j, s_j) was absorbed at push time: finalise a copy of the running SHA-512 state
    if (b->mb) {
      b->mb->digest(b->mb_lane, b->seed);
    } else {
      EVP_MD_CTX* fin = EVP_MD_CTX_new();
      unsigned int outl = 64;
      EVP_MD_CTX_copy_ex(fin, b->hctx);
      EVP_DigestFinal_ex(fin, b->seed, &outl);
      EVP_MD_CTX_free(fin);
    }
    b->tm.host_hash_ms = 0;
    b->have_seed = true;
  } else if ((rc = seed_from_device(b))) return rc;                // also orders after k_prepare
  // The identity gate (thin.rs:266-271) is decided at wait time from flags[0]; it takes precedence over
  // the MSM verdict, so running the MSM regardless does not change any result.
  if ((rc = run_msm(b, b->seed, 0))) return rc;
  CK(cudaMemcpyAsync(b->h_small.p, b->flags.p, 8, cudaMemcpyDeviceToHost, b->st));
  CK(cudaMemcpyAsync((uint8_t*)b->h_small.p + 128, b->totals.p, 8, cudaMemcpyDeviceToHost, b->st));
  if (!b->done_ev) CK(cudaEventCreateWithFlags(&b->done_ev, ev_flags(b)));
  CK(cudaEventRecord(b->done_ev, b->st));
  b->inflight = true;
  b->inflight_did_prepare = did_prepare;
  return 0;
}

int avrf_thin_batch_verify_wait(avrf_batch* b, int32_t* status) {
  if (!b || !status) return fail(AVRF_ERR_ARG, "null argument");
  if (!b->inflight) return fail(AVRF_ERR_STATE, "no verify in flight");
  b->inflight = false;
  if (b->early_status >= 0) { *status = b->early_status; return 0; }
  CK(cudaEventSynchronize(b->done_ev));
  const int* fl = reinterpret_cast<const int*>(b->h_small.p);
  if (fl[0] & 2) return fail(AVRF_ERR_ARG, "an input coordinate or scalar is not below its modulus (not a field element)");
  if (fl[0] & 1) *status = AVRF_INVALID_DATA;                                  // thin.rs:266-271
  else *status = fl[1] ? AVRF_OK : AVRF_VERIFICATION_FAILURE;                  // thin.rs:320-324
  b->tm.n_entries = reinterpret_cast<uint32_t*>((uint8_t*)b->h_small.p + 128)[0];
  b->tm.n_tasks = reinterpret_cast<uint32_t*>((uint8_t*)b->h_small.p + 128)[1];
  collect_timings(b, b->inflight_did_prepare);
  b->tm.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - b->t_verify0).count();
  return 0;
}

static int finish_inflight(avrf_batch* b) {
  if (!b->inflight) return 0;
  int32_t st;
  return avrf_thin_batch_verify_wait(b, &st);
}

int avrf_thin_batch_verify(avrf_batch* b, int32_t* status) {
  if (!b || !status) return fail(AVRF_ERR_ARG, "null argument");
  int rc = avrf_thin_batch_verify_async(b);
  if (rc) return rc;
  return avrf_thin_batch_verify_wait(b, status);
}

// ---- Pedersen VRF batch verifier on the same engine (reference src/pedersen.rs:255-427) -------
avrf_batch* avrf_pedersen_batch_new(uint32_t suite, uint32_t fmt) {
  avrf_batch* b = avrf_thin_batch_new(suite, fmt);
  if (b) b->scheme = 1;      // eager seeding like the thin verifier: push pipelines H2D, k_prepare_ped, D2H and the host SHA-512
  return b;
}

int avrf_pedersen_batch_push_many(avrf_batch* b, uint64_t n, const uint8_t* ios, const uint32_t* io_offsets,
                                  const uint8_t* ad_blob, const uint32_t* ad_offsets, const uint8_t* pk_com,
                                  const uint8_t* r, const uint8_t* ok, const uint8_t* s, const uint8_t* sb) {
  if (!b || b->scheme != 1) return fail(AVRF_ERR_ARG, "not a Pedersen batch");
  if (n == 0) return 0;
  if (!io_offsets || !ad_offsets || !pk_com || !r || !ok || !s || !sb) return fail(AVRF_ERR_ARG, "null argument");
  if (io_offsets[0] != 0 || ad_offsets[0] != 0) return fail(AVRF_ERR_ARG, "offsets must start at 0");
  uint64_t add_ios = io_offsets[n], add_ad = ad_offsets[n];
  if ((add_ios && !ios) || (add_ad && !ad_blob)) return fail(AVRF_ERR_ARG, "null argument");
  NEED_DEVICE();
  // same pipeline as the thin verifier: chunked H2D -> k_prepare_ped -> D2H of (c, s, sb) -> incremental SHA-512
  int rc = push_many_impl(b, n, pk_com, ios, io_offsets, ad_blob, ad_offsets, r, s, ok, sb);
  if (rc) return rc;
  CK(hsync(b, b->st));     // the caller's buffers are only borrowed for the duration of the call
  return 0;
}

int avrf_pedersen_batch_verify(avrf_batch* b, int32_t* status) {
  if (!b || b->scheme != 1) return fail(AVRF_ERR_ARG, "not a Pedersen batch");
  return avrf_thin_batch_verify(b, status);
}

int avrf_thin_verify_one(uint32_t suite, uint32_t fmt, const uint8_t pk[64], const uint8_t* ios, uint32_t n_ios,
                         const uint8_t* ad, uint32_t ad_len, const uint8_t r[64], const uint8_t s[32], int32_t* status) {
  avrf_batch* b = avrf_thin_batch_new(suite, fmt);
  if (!b) return AVRF_ERR_ARG;
  int rc = avrf_thin_batch_push(b, pk, ios, n_ios, ad, ad_len, r, s);
  if (!rc) rc = avrf_thin_batch_verify(b, status);
  avrf_thin_batch_free(b);
  return rc;
}

int avrf_thin_batch_timings(const avrf_batch* b, avrf_timings* out) {
  if (!b || !out) return fail(AVRF_ERR_ARG, "null argument");
  *out = b->tm;
  return 0;
}

int avrf_thin_batch_tap(avrf_batch* b, uint32_t what, void* out, size_t out_bytes) {
  if (!b || !out) return fail(AVRF_ERR_ARG, "null argument");
  NEED_DEVICE();
  int rc;
  size_t np = npoints_of(b);
  auto d2h = [&](const void* src, size_t bytes) -> int {
    if (out_bytes < bytes) return fail(AVRF_ERR_ARG, "tap buffer too small");
    if (bytes) CK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, b->st));
    CK(hsync(b, b->st));
    return 0;
  };
  switch (what) {
    case AVRF_TAP_C: {
      if ((rc = avrf_thin_batch_prepare(b, nullptr))) return rc;
      if (out_bytes < 16 * b->n) return fail(AVRF_ERR_ARG, "tap buffer too small");
      if (b->n) CK(cudaMemcpy2DAsync(out, 16, b->cs.p, cs_stride(b), 16, b->n, cudaMemcpyDeviceToHost, b->st));
      CK(hsync(b, b->st));
      return 0;
    }
    case AVRF_TAP_Z:
      if ((rc = avrf_thin_batch_prepare(b, nullptr))) return rc;
      return d2h(b->z.p, 16 * b->n_ios);
    case AVRF_TAP_R_COMPRESSED:
      if ((rc = avrf_thin_batch_prepare(b, nullptr))) return rc;
      return d2h(b->renc.p, 32 * b->n);
    case AVRF_TAP_SEED:
      if (!b->have_seed) return fail(AVRF_ERR_STATE, "no seed yet: call verify or partial first");
      if (out_bytes < 64) return fail(AVRF_ERR_ARG, "tap buffer too small");
      memcpy(out, b->seed, 64);
      return 0;
    case AVRF_TAP_W:
    case AVRF_TAP_SCALARS: {
      if (!b->have_seed) return fail(AVRF_ERR_STATE, "no seed yet: call verify or partial first");
      if (b->n == 0) return 0;
      b->want_taps = true;
      rc = run_msm(b, b->seed, b->first_index);
      b->want_taps = false;
      if (rc) return rc;
      return what == AVRF_TAP_W ? d2h(b->w_tap.p, (b->scheme ? 32 : 16) * b->n) : d2h(b->scalars_tap.p, 32 * np);
    }
    case AVRF_TAP_PARTIAL:
      if (!b->have_seed || !b->partial.p) return fail(AVRF_ERR_STATE, "no partial yet");
      return d2h(b->partial.p, 128);
    default:
      return fail(AVRF_ERR_ARG, "unknown tap");
  }
}

// ---- feeder operations ---------------------------------------------------------------------
int avrf_hash_to_curve(uint32_t suite, uint32_t fmt, const uint8_t* msgs, const uint32_t* offsets, uint64_t n,
                       uint8_t* out_affine, uint8_t* out_compressed, uint8_t* ok) {
  if (suite > 2 || fmt > 1 || !offsets || (n && offsets[n] && !msgs)) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  DevBuf dm, doff, daff, denc, dok;
  int rc;
  if ((rc = dm.reserve(offsets[n] + 16)) || (rc = doff.reserve(4 * (n + 1))) || (rc = daff.reserve(64 * n)) ||
      (rc = denc.reserve(32 * n)) || (rc = dok.reserve(n)))
    return rc;
  if (offsets[n]) CK(cudaMemcpyAsync(dm.p, msgs, offsets[n], cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(doff.p, offsets, 4 * (n + 1), cudaMemcpyHostToDevice, g_stream));
  DISPATCH(suite, (k_h2c<S><<<cdiv(n, 128), 128, 0, g_stream>>>(dm.as<uint8_t>(), doff.as<uint32_t>(), (uint32_t)n,
                                                                daff.as<Affine>(), denc.as<uint32_t>(),
                                                                dok.as<uint8_t>(), fmt == AVRF_FMT_CANONICAL)));
  LAUNCHED("k_h2c");
  if (out_affine) CK(cudaMemcpyAsync(out_affine, daff.p, 64 * n, cudaMemcpyDeviceToHost, g_stream));
  if (out_compressed) CK(cudaMemcpyAsync(out_compressed, denc.p, 32 * n, cudaMemcpyDeviceToHost, g_stream));
  if (ok) CK(cudaMemcpyAsync(ok, dok.p, n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  dm.release(); doff.release(); daff.release(); denc.release(); dok.release();
  return 0;
}

static int scalar_mul_impl(uint32_t suite, uint32_t fmt, const uint8_t* sk, uint32_t sk_stride, const uint8_t* inputs,
                           uint64_t n, uint8_t* outputs) {
  if (suite > 2 || fmt > 1 || !sk || !outputs || (sk_stride != 0 && sk_stride != 32)) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  DevBuf dsk, din, dout;
  int rc;
  size_t skb = sk_stride ? 32 * n : 32;
  if ((rc = dsk.reserve(skb)) || (rc = dout.reserve(64 * n))) return rc;
  if (inputs && (rc = din.reserve(64 * n))) return rc;
  CK(cudaMemcpyAsync(dsk.p, sk, skb, cudaMemcpyHostToDevice, g_stream));
  if (inputs) CK(cudaMemcpyAsync(din.p, inputs, 64 * n, cudaMemcpyHostToDevice, g_stream));
  DISPATCH(suite, (k_scalar_mul<S><<<cdiv(n, 128), 128, 0, g_stream>>>(dsk.as<Fe>(), sk_stride / 4,
                                                                       inputs ? din.as<Affine>() : nullptr, (uint32_t)n,
                                                                       dout.as<Affine>(), fmt == AVRF_FMT_CANONICAL)));
  LAUNCHED("k_scalar_mul");
  CK(cudaMemcpyAsync(outputs, dout.p, 64 * n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  dsk.release(); din.release(); dout.release();
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

int avrf_thin_prove_many(uint32_t suite, uint32_t fmt, uint64_t n, const uint8_t* sk, const uint8_t* pk,
                         const uint8_t* ios, const uint32_t* io_offsets, const uint8_t* ad_blob,
                         const uint32_t* ad_offsets, uint8_t* r, uint8_t* s) {
  if (suite > 2 || fmt > 1 || !sk || !pk || !io_offsets || !ad_offsets || !r || !s) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  size_t nio = io_offsets[n], nad = ad_offsets[n];
  if ((nio && !ios) || (nad && !ad_blob)) return fail(AVRF_ERR_ARG, "null argument");
  DevBuf dsk, dpk, dios, dio, dao, dad, dr, ds, derr;
  int rc;
  if ((rc = dsk.reserve(32 * n)) || (rc = dpk.reserve(64 * n)) || (rc = dios.reserve(128 * nio + 128)) ||
      (rc = dio.reserve(4 * (n + 1))) || (rc = dao.reserve(4 * (n + 1))) || (rc = dad.reserve(nad + 16)) ||
      (rc = dr.reserve(64 * n)) || (rc = ds.reserve(32 * n)) || (rc = derr.reserve(64)))
    return rc;
  CK(cudaMemcpyAsync(dsk.p, sk, 32 * n, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dpk.p, pk, 64 * n, cudaMemcpyHostToDevice, g_stream));
  if (nio) CK(cudaMemcpyAsync(dios.p, ios, 128 * nio, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dio.p, io_offsets, 4 * (n + 1), cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dao.p, ad_offsets, 4 * (n + 1), cudaMemcpyHostToDevice, g_stream));
  if (nad) CK(cudaMemcpyAsync(dad.p, ad_blob, nad, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemsetAsync(derr.p, 0, 64, g_stream));
  ProveArgs a;
  a.sk = dsk.as<Fe>(); a.pk = dpk.as<Affine>(); a.ios = dios.as<Affine>(); a.io_off = dio.as<uint32_t>();
  a.ad_off = dao.as<uint32_t>(); a.ad = dad.as<uint8_t>(); a.r = dr.as<Affine>(); a.s = ds.as<Fe>();
  a.n = (uint32_t)n; a.canonical = fmt == AVRF_FMT_CANONICAL;
  DISPATCH(suite, (k_prove<S><<<cdiv(n, 128), 128, 0, g_stream>>>(a, derr.as<int>())));
  LAUNCHED("k_prove");
  int herr = 0;
  CK(cudaMemcpyAsync(r, dr.p, 64 * n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(s, ds.p, 32 * n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(&herr, derr.p, 4, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  DevBuf* bufs[] = {&dsk, &dpk, &dios, &dio, &dao, &dad, &dr, &ds, &derr};
  for (DevBuf* d : bufs) d->release();
  if (herr) return fail(AVRF_ERR_ARG, "avrf_thin_prove_many supports at most 8 I/O pairs per proof");
  return 0;
}

static int compress_impl(uint32_t suite, uint32_t fmt, const uint8_t* points, uint64_t n, uint8_t* out32, int hash) {
  if (suite > 2 || fmt > 1 || !points || !out32) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  DevBuf din, dout;
  int rc;
  if ((rc = din.reserve(64 * n)) || (rc = dout.reserve(32 * n))) return rc;
  CK(cudaMemcpyAsync(din.p, points, 64 * n, cudaMemcpyHostToDevice, g_stream));
  DISPATCH(suite, (k_compress<S><<<cdiv(n, 128), 128, 0, g_stream>>>(din.as<Affine>(), n, dout.as<uint32_t>(),
                                                                     fmt == AVRF_FMT_CANONICAL, hash)));
  LAUNCHED("k_compress");
  CK(cudaMemcpyAsync(out32, dout.p, 32 * n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  din.release(); dout.release();
  return 0;
}

int avrf_points_deserialize(uint32_t suite, uint32_t fmt, uint32_t kind, const uint8_t* in32, uint64_t n, uint8_t* out64,
                            uint8_t* ok) {
  if (suite > 2 || fmt > 1 || kind > 1 || !in32 || !out64 || !ok) return fail(AVRF_ERR_ARG, "bad argument");
  NEED_DEVICE();
  if (n == 0) return 0;
  DevBuf din, dout, dok;
  int rc;
  if ((rc = din.reserve(32 * n)) || (rc = dout.reserve(64 * n)) || (rc = dok.reserve(n))) return rc;
  CK(cudaMemcpyAsync(din.p, in32, 32 * n, cudaMemcpyHostToDevice, g_stream));
  DISPATCH(suite, (k_deserialize<S><<<cdiv(n, 128), 128, 0, g_stream>>>(din.as<uint32_t>(), n, (int)kind, dout.as<Affine>(),
                                                                         dok.as<uint8_t>(), fmt == AVRF_FMT_CANONICAL)));
  LAUNCHED("k_deserialize");
  CK(cudaMemcpyAsync(out64, dout.p, 64 * n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(ok, dok.p, n, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  din.release(); dout.release(); dok.release();
  return 0;
}

int avrf_thin_batch_verify_each(avrf_batch* b, int32_t* statuses) {
  if (!b || !statuses) return fail(AVRF_ERR_ARG, "null argument");
  if (b->scheme != 0) return fail(AVRF_ERR_ARG, "not a Thin-VRF batch");
  int rc = avrf_thin_batch_prepare(b, nullptr);
  if (rc) return rc;
  if (b->n == 0) return 0;
  DevBuf dst;
  if ((rc = dst.reserve(4 * b->n))) return rc;
  EachArgs a;
  a.pk = b->pk.as<Affine>(); a.r = b->r.as<Affine>(); a.ios = b->ios.as<Affine>();
  a.canonical = b->fmt == AVRF_FMT_CANONICAL;
  a.cs = b->cs.as<uint32_t>(); a.z = b->z.as<uint32_t>();
  a.io_off = b->io_off.as<uint32_t>(); a.status = dst.as<int32_t>(); a.n = (uint32_t)b->n;
  DISPATCH(b->suite, (k_verify_each<S><<<cdiv(b->n, 128), 128, 0, b->st>>>(a)));
  LAUNCHED("k_verify_each");
  CK(cudaMemcpyAsync(statuses, dst.p, 4 * b->n, cudaMemcpyDeviceToHost, b->st));
  CK(hsync(b, b->st));
  dst.release();
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
  CK(cudaGetDeviceProperties(&prop, g_device));
  int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  DevBuf out, pts;
  int rc;
  double work = 0;
  float ms = 0;
  if (kind == 0) {
    int blocks = sms * 8, threads = 256;
    if ((rc = out.reserve(8ull * blocks * threads))) return rc;
    k_mb_imad<<<blocks, threads, 0, g_stream>>>(out.as<uint64_t>(), 16, 1);   // warm-up
    CK(cudaEventRecord(e0, g_stream));
    k_mb_imad<<<blocks, threads, 0, g_stream>>>(out.as<uint64_t>(), iters, 2);
    CK(cudaEventRecord(e1, g_stream));
    work = (double)blocks * threads * iters * 32.0;
  } else if (kind == 1) {
    int blocks = sms * 16, threads = 128;
    if ((rc = out.reserve(32ull * blocks * threads))) return rc;
    k_mb_mul<<<blocks, threads, 0, g_stream>>>(out.as<Fe>(), 4);
    CK(cudaEventRecord(e0, g_stream));
    k_mb_mul<<<blocks, threads, 0, g_stream>>>(out.as<Fe>(), iters);
    CK(cudaEventRecord(e1, g_stream));
    work = (double)blocks * threads * iters * 2.0;
  } else if (kind == 2) {
    int blocks = sms * 16, threads = 128;
    uint32_t npts = 1u << 20;  // (bases are arbitrary field elements: the formulas do not care)
    if ((rc = out.reserve(128ull * blocks * threads)) || (rc = pts.reserve(96ull * npts))) return rc;
    CK(cudaMemsetAsync(pts.p, 0x11, 96ull * npts, g_stream));
    k_mb_madd<<<blocks, threads, 0, g_stream>>>(out.as<Ext>(), pts.as<AffineK>(), npts, 2);
    CK(cudaEventRecord(e0, g_stream));
    k_mb_madd<<<blocks, threads, 0, g_stream>>>(out.as<Ext>(), pts.as<AffineK>(), npts, iters);
    CK(cudaEventRecord(e1, g_stream));
    work = (double)blocks * threads * iters;
  } else if (kind == 5) {
    int blocks = sms * 16, threads = 128;
    if ((rc = out.reserve(36ull * blocks * threads))) return rc;
    k_mb_mul29<<<blocks, threads, 0, g_stream>>>(out.as<Fe29>(), 4);
    CK(cudaEventRecord(e0, g_stream));
    k_mb_mul29<<<blocks, threads, 0, g_stream>>>(out.as<Fe29>(), iters);
    CK(cudaEventRecord(e1, g_stream));
    work = (double)blocks * threads * iters * 2.0;
  } else if (kind == 6 || kind == 7) {
    int blocks = sms * 8, threads = 256;
    if ((rc = out.reserve(8ull * blocks * threads))) return rc;
    if (kind == 6) {
      k_mb_imadc<0><<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, g_stream));
      k_mb_imadc<0><<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), iters, 2);
    } else {
      k_mb_imadc<1><<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, g_stream));
      k_mb_imadc<1><<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), iters, 2);
    }
    CK(cudaEventRecord(e1, g_stream));
    work = (double)blocks * threads * iters * 32.0;
  } else if (kind == 3 || kind == 4) {
    int blocks = sms * 8, threads = 256;
    if ((rc = out.reserve(8ull * blocks * threads))) return rc;
    if (kind == 3) {
      k_mb_imadx<<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, g_stream));
      k_mb_imadx<<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), iters, 2);
    } else {
      k_mb_imad32<<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), 16, 1);
      CK(cudaEventRecord(e0, g_stream));
      k_mb_imad32<<<blocks, threads, 0, g_stream>>>(out.as<uint32_t>(), iters, 2);
    }
    CK(cudaEventRecord(e1, g_stream));
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
  out.release();
  pts.release();
  return 0;
}

}  // extern "C"

#include "server.inl"   // avrf_server_*: worker pool + shared multi-buffer SHA-512 threads, avrf_mb_sha512

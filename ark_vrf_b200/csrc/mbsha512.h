// Multi-buffer SHA-512 on the host: up to 8 independent byte streams hashed in lockstep, one stream per
// 64-bit lane of an AVX-512 register.
//
// Why it exists: the reference seeds the batch weights with ONE serial SHA-512 over all (c_j, s_j) of a batch
// (reference src/thin.rs:273-279) - 64 MiB for 2^20 proofs, ~82 ms on one core, the Amdahl term of the whole
// path.  A hash chain cannot be parallelised, but DIFFERENT batches are independent: in the throughput mode
// (avrf_server_*) one hashing thread advances eight batches' chains at once at 4-5x the aggregate rate of eight
// scalar hashes on eight cores' worth of time, so the host stops being the limit when several GPUs share a box.
// Digests are bit-identical to any SHA-512 (tests/test_mbsha512_cpu.py compares with hashlib).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace avrf {

class MbSha512 {
 public:
  static constexpr int LANES = 8;
  MbSha512();
  ~MbSha512();
  MbSha512(const MbSha512&) = delete;
  MbSha512& operator=(const MbSha512&) = delete;

  static bool simd_available();           // AVX-512F + BW at run time; otherwise lanes are hashed one after the other

  int acquire();                          // a free lane (reset to the SHA-512 IV) or -1
  void release(int lane);
  void reset(int lane);                   // waits for queued data, then back to the IV
  // Queue n bytes for the lane and return at once.  The bytes must stay valid and unchanged until sync() or
  // digest() on the lane has returned.
  void update(int lane, const uint8_t* p, size_t n);
  void sync(int lane);                    // all queued bytes of the lane have been absorbed
  // SHA-512 of everything absorbed so far (sync first); the lane keeps its state and can absorb more.
  void digest(int lane, uint8_t out[64]);

 private:
  struct Seg { const uint8_t* p; size_t n; };
  struct Lane {
    uint64_t h[8];
    uint8_t buf[128];
    size_t buflen = 0;
    uint64_t total = 0;                   // bytes absorbed or queued
    std::deque<Seg> q;                    // queued, not yet absorbed (owned by the hashing thread once popped)
    bool busy = false;                    // the hashing thread is working on a segment of this lane
    bool used = false;
  };
  void run();
  void absorb(Lane& l, const uint8_t* p, size_t n);   // scalar path for odd bytes (hashing thread only)
  Lane lanes_[LANES];
  std::mutex mu_;
  std::condition_variable cv_work_, cv_idle_;
  bool stop_ = false;
  std::thread th_;
};

// SHA-512 compression of `nblk` consecutive 128-byte blocks per lane, eight lanes in lockstep.
// h[lane][0..7] is updated for the lanes whose bit is set in `mask`; ptr[lane] must be readable for every lane.
void sha512_blocks_x8(uint64_t h[8][8], const uint8_t* const ptr[8], size_t nblk, unsigned mask);
// One lane, scalar (fallback and odd blocks).
void sha512_blocks_x1(uint64_t h[8], const uint8_t* p, size_t nblk);

}  // namespace avrf

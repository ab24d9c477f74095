// GPU test driver, no Python in the process: one batch over every GPU of the box through the C++ mirror
// (include/avrf.hpp: init_multi + thin::ShardedBatchVerifier over avrf_thin_sharded_*).  Reads the proof blob of
// tests/test_gpu_cpp.py (see thin_driver.cpp for the layout), repeats it to a few thousand proofs so that every
// device gets a share, and checks accept / reject / InvalidData against a single-device BatchVerifier.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "avrf.hpp"

using namespace ark_vrf;
using Suite = BandersnatchSha512Ell2;

struct Item { AffinePoint pk; std::vector<VrfIo> ios; std::vector<uint8_t> ad; thin::Proof proof; };

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  size_t p = 0;
  auto u32 = [&]() { uint32_t v; memcpy(&v, &buf[p], 4); p += 4; return v; };
  auto bytes = [&](void* dst, size_t n) { memcpy(dst, &buf[p], n); p += n; };
  uint32_t n = u32();
  std::vector<Item> items(n);
  for (auto& it : items) {
    bytes(it.pk.data(), 64);
    it.ios.resize(u32());
    for (auto& io : it.ios) { bytes(io.input.data(), 64); bytes(io.output.data(), 64); }
    it.ad.resize(u32());
    if (!it.ad.empty()) bytes(it.ad.data(), it.ad.size());
    bytes(it.proof.r.data(), 64);
    bytes(it.proof.s.data(), 32);
  }
  int ndev = thin::init_multi(0);
  std::printf("devices: %d\n", ndev);
  const size_t reps = 700;                       // 6 proofs x 700 = 4200 proofs
  auto build = [&](long bad_at, long ident_at) {
    thin::Batch b;
    for (size_t r = 0; r < reps; r++)
      for (size_t i = 0; i < items.size(); i++) {
        long idx = (long)(r * items.size() + i);
        auto pf = items[i].proof;
        auto pk = items[i].pk;
        if (idx == bad_at) pf.s[0] ^= 1;
        if (idx == ident_at) { pk.fill(0); pk[32] = 1; }
        b.push(pk, items[i].ios, items[i].ad, pf);
      }
    return b;
  };
  int fails = 0;
  auto expect = [&](const char* what, int got, int want) {
    std::printf("%-44s status %d (want %d)\n", what, got, want);
    if (got != want) fails++;
  };
  long total = (long)(reps * items.size());
  thin::ShardedBatchVerifier<Suite, AVRF_FMT_CANONICAL> sh;
  if (sh.devices() != ndev) fails++;
  {
    thin::Batch b = build(-1, -1);
    sh.push_many(b);
    expect("sharded: valid batch", sh.verify().status, AVRF_OK);
    expect("sharded: verify is repeatable", sh.verify().status, AVRF_OK);
    uint8_t seed_sh[64], seed_one[64];
    check(avrf_thin_sharded_seed(sh.handle(), seed_sh));
    thin::BatchVerifier<Suite, AVRF_FMT_CANONICAL> one;
    check(avrf_thin_batch_push_many(one.handle(), b.len(), b.pk[0].data(), b.ios[0].input.data(), b.io_offsets.data(),
                                    b.ad.data(), b.ad_offsets.data(), b.r[0].data(), b.s[0].data()));
    expect("single device: valid batch", one.verify().status, AVRF_OK);
    check(avrf_thin_batch_tap(one.handle(), AVRF_TAP_SEED, seed_one, 64));
    expect("seed equals the single-device seed", memcmp(seed_sh, seed_one, 64) == 0, 1);
  }
  for (long bad : {0L, total / 2, total - 1}) {
    thin::Batch b = build(bad, -1);
    sh.clear();
    sh.push_many(b);
    expect("sharded: one bad response", sh.verify().status, AVRF_VERIFICATION_FAILURE);
  }
  {
    thin::Batch b = build(3, total - 2);
    sh.clear();
    sh.push_many(b);
    expect("sharded: identity pk + bad response", sh.verify().status, AVRF_INVALID_DATA);
  }
  {
    sh.clear();
    expect("sharded: empty batch", sh.verify().status, AVRF_OK);
  }
  std::printf("%s\n", fails ? "FAIL" : "ALL OK");
  return fails ? 1 : 0;
}

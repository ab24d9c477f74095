// GPU test driver for the C++ mirror of the reference API (include/avrf.hpp): reads proofs from a
// binary blob written by tests/test_gpu_cpp.py and replays the reference's batch scenarios
// (src/thin.rs:346-384, 418-433) through ark_vrf::thin::BatchVerifier / Public::verify.
//
// blob: u32 n, then per proof: pk[64] n_ios(u32) ios[128*n_ios] ad_len(u32) ad[ad_len] r[64] s[32]  (canonical)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "avrf.hpp"

using namespace ark_vrf;
using Suite = BandersnatchSha512Ell2;
using BV = thin::BatchVerifier<Suite, AVRF_FMT_CANONICAL>;

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
  int fails = 0;
  auto expect = [&](const char* what, Result r, int want) {
    std::printf("%-40s status %d (want %d)\n", what, r.status, want);
    if (r.status != want) fails++;
  };
  {  // all proofs, pushed one by one; single verify of each
    BV bv;
    for (auto& it : items) bv.push(it.pk, it.ios, it.ad, it.proof);
    expect("batch of valid proofs", bv.verify(), AVRF_OK);
    expect("verify is repeatable", bv.verify(), AVRF_OK);
    for (auto& it : items) {
      Public<Suite, AVRF_FMT_CANONICAL> pub{it.pk};
      if (!pub.verify(it.ios, it.ad, it.proof).is_ok()) fails++;
    }
  }
  {  // prepare + push_prepared, one wrong ad
    BV bv;
    for (size_t i = 0; i < items.size(); i++) {
      auto ad = items[i].ad;
      if (i == 1) ad.push_back('!');
      bv.push_prepared(BV::prepare(items[i].pk, items[i].ios, ad, items[i].proof));
    }
    Result r = bv.verify();
    expect("one wrong ad", r, AVRF_VERIFICATION_FAILURE);
    if (r.is_ok() || r.unwrap_err() != Error::VerificationFailure) fails++;
  }
  {  // identity public key -> InvalidData, also with a bad response elsewhere
    BV bv;
    AffinePoint ident{};
    ident[32] = 1;  // (0, 1) canonical
    auto pf = items[0].proof;
    pf.s[0] ^= 1;
    bv.push(items[0].pk, items[0].ios, items[0].ad, pf);
    bv.push(ident, items[1].ios, items[1].ad, items[1].proof);
    expect("identity pk + bad s", bv.verify(), AVRF_INVALID_DATA);
  }
  {
    BV bv;
    expect("empty batch", bv.verify(), AVRF_OK);
  }
  {  // worker pool: whole batches submitted from one thread, verdicts by ticket
    thin::BatchServer<Suite, AVRF_FMT_CANONICAL> srv(2);
    thin::Batch good, bad, none;
    for (auto& it : items) good.push(it.pk, it.ios, it.ad, it.proof);
    for (size_t i = 0; i < items.size(); i++) {
      auto pf = items[i].proof;
      if (i + 1 == items.size()) pf.s[3] ^= 8;
      bad.push(items[i].pk, items[i].ios, items[i].ad, pf);
    }
    int64_t t[5] = {srv.submit(good), srv.submit(bad), srv.submit(none), srv.submit(good), srv.submit(bad)};
    int want[5] = {AVRF_OK, AVRF_VERIFICATION_FAILURE, AVRF_OK, AVRF_OK, AVRF_VERIFICATION_FAILURE};
    for (int i = 4; i >= 0; i--) expect("server ticket", srv.wait(t[i]), want[i]);
  }
  {  // wire format: the same proofs as serialize_compressed bytes, decoded and validated on the device
    size_t nio = 0;
    for (auto& it : items) nio += it.ios.size();
    std::vector<uint8_t> pk32(32 * n), r32(32 * n), s32(32 * n), ios32(64 * nio + 64), ad;
    std::vector<uint32_t> io_off{0}, ad_off{0};
    size_t q = 0;
    for (size_t i = 0; i < n; i++) {
      if (avrf_point_compress(Suite::ID, AVRF_FMT_CANONICAL, items[i].pk.data(), 1, &pk32[32 * i])) fails++;
      if (avrf_point_compress(Suite::ID, AVRF_FMT_CANONICAL, items[i].proof.r.data(), 1, &r32[32 * i])) fails++;
      memcpy(&s32[32 * i], items[i].proof.s.data(), 32);
      for (auto& io : items[i].ios) {
        if (avrf_point_compress(Suite::ID, AVRF_FMT_CANONICAL, io.input.data(), 1, &ios32[64 * q])) fails++;
        if (avrf_point_compress(Suite::ID, AVRF_FMT_CANONICAL, io.output.data(), 1, &ios32[64 * q + 32])) fails++;
        q++;
      }
      ad.insert(ad.end(), items[i].ad.begin(), items[i].ad.end());
      io_off.push_back((uint32_t)q);
      ad_off.push_back((uint32_t)ad.size());
    }
    ad.resize(ad.size() + 16);
    BV bv;
    std::vector<uint8_t> ok;
    uint64_t bad = bv.push_compressed(n, pk32.data(), ios32.data(), io_off.data(), ad.data(), ad_off.data(), r32.data(), s32.data(), &ok);
    std::printf("%-40s bad %llu len %lld\n", "wire-format push", (unsigned long long)bad, (long long)bv.len());
    if (bad != 0 || bv.len() != (int64_t)n) fails++;
    expect("wire-format batch", bv.verify(), AVRF_OK);
    memset(&pk32[32], 0xff, 32);                                  // y >= p: does not deserialize
    bad = bv.push_compressed(n, pk32.data(), ios32.data(), io_off.data(), ad.data(), ad_off.data(), r32.data(), s32.data(), &ok);
    std::printf("%-40s bad %llu len %lld ok[1] %d\n", "wire-format push, one bad encoding", (unsigned long long)bad, (long long)bv.len(), ok[1]);
    if (bad != 1 || ok[1] != 0 || ok[0] != 1 || bv.len() != (int64_t)n) fails++;
    expect("batch unchanged by the refused push", bv.verify(), AVRF_OK);
  }
  std::printf("%s\n", fails ? "FAIL" : "ALL OK");
  return fails ? 1 : 0;
}

"""CPU: include/avrf.hpp (the C++ mirror of the reference's Rust API) compiles against the C ABI
and links against libavrf_gpu.so."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "avrf.hpp"
#include <cstdio>
using namespace ark_vrf;
int main() {
  // no GPU on the build box: constructing a verifier must fail loudly, not fall back
  try {
    thin::BatchVerifier<BandersnatchSha512Ell2> bv;
    thin::Proof p{};
    bv.push(AffinePoint{}, {}, {}, p);
    Result r = bv.verify();
    std::printf("status %d\n", r.status);
    thin::BatchServer<BandersnatchSha512Ell2> srv(2);
    thin::Batch b;
    b.push(AffinePoint{}, {}, {}, p);
    std::printf("status %d\n", srv.wait(srv.submit(b)).status);
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
  }
  return 0;
}
'''


def test_cpp_mirror_compiles_and_links():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(d, "t")
        lib = os.path.join(ROOT, "ark_vrf_b200")
        subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L", lib, "-l:libavrf_gpu.so", "-Wl,-rpath," + lib])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0
        assert "status" in out.stdout or "error" in out.stdout

"""GPU: the C++ mirror of the reference's Rust API (include/avrf.hpp) exercised by a compiled driver
against oracle-generated proofs - the compiled-language host path above the C ABI."""
import os
import struct
import subprocess
import tempfile

import pytest

from oracle import pyref as o
from helpers import pt_bytes, sc_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_driver_on_gpu():
    S = o.BANDERSNATCH
    pr = o.synth_proofs(S, 4, 1, signers=2)
    extra = o.synth_proofs(S, 2, 3, signers=2)
    for f in ("pk", "ios", "ad", "r", "s"):
        getattr(pr, f).extend(getattr(extra, f))
    blob = struct.pack("<I", len(pr.pk))
    for j in range(len(pr.pk)):
        blob += pt_bytes(pr.pk[j]) + struct.pack("<I", len(pr.ios[j]))
        for a, b in pr.ios[j]:
            blob += pt_bytes(a) + pt_bytes(b)
        blob += struct.pack("<I", len(pr.ad[j])) + pr.ad[j] + pt_bytes(pr.r[j]) + sc_bytes(pr.s[j])
    with tempfile.TemporaryDirectory() as d:
        exe, data = os.path.join(d, "thin_driver"), os.path.join(d, "proofs.bin")
        open(data, "wb").write(blob)
        lib = os.path.join(ROOT, "ark_vrf_b200")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", "thin_driver.cpp"), "-o", exe,
                               "-L", lib, "-l:libavrf_gpu.so", "-Wl,-rpath," + lib])
        out = subprocess.run([exe, data], capture_output=True, text=True, timeout=300)
        print(out.stdout, out.stderr)
        assert out.returncode == 0 and "ALL OK" in out.stdout

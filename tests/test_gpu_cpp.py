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


def _blob():
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
    return blob


def _run_driver(name):
    with tempfile.TemporaryDirectory() as d:
        exe, data = os.path.join(d, name), os.path.join(d, "proofs.bin")
        open(data, "wb").write(_blob())
        lib = os.path.join(ROOT, "ark_vrf_b200")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
                               "-L", lib, "-l:libavrf_gpu.so", "-Wl,-rpath," + lib])
        out = subprocess.run([exe, data], capture_output=True, text=True, timeout=300)
        print(out.stdout, out.stderr)
        assert out.returncode == 0 and "ALL OK" in out.stdout
        return out.stdout


def test_cpp_driver_on_gpu():
    _run_driver("thin_driver")


def test_cpp_sharded_driver_on_all_gpus():
    """One batch over every GPU of the box from a C++ host (no Python, no torch.distributed in that process):
    avrf_init_multi + avrf_thin_sharded_* through include/avrf.hpp.  Under `gpurun --gpus N` this drives N devices."""
    import torch
    out = _run_driver("sharded_driver")
    assert "devices: %d" % torch.cuda.device_count() in out

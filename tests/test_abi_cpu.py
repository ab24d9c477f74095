"""CPU: the C-ABI library loads, exports every symbol include/avrf.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "avrf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avrf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import ark_vrf_b200 as av
    lib = av.load()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libavrf_gpu.so does not export {s}"
    # and the Python binding declares a prototype for each of them
    from ark_vrf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ark_vrf_b200 as av
    with pytest.raises(av.AvrfError) as e:
        av.BatchVerifier(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    lib = av.load()
    st = ctypes.c_int32(-7)
    rc = lib.avrf_thin_verify_one(0, 1, bytes(64), None, 0, None, 0, bytes(64), bytes(32), ctypes.byref(st))
    assert rc < 0 and st.value == -7          # a system error, never a verdict
    from ark_vrf_b200 import ops
    with pytest.raises(av.AvrfError):
        ops.hash_to_curve(0, [b"x"])


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "ark_vrf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".rs", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("the oracle's", "").replace("oracle-backed", "") or f == "dist.py", (dirpath, f)


def test_host_side_helpers():
    from ark_vrf_b200 import synth
    from ark_vrf_b200.dist import shard_bounds
    from oracle import pyref as o
    for sid, S in o.SUITES.items():
        seed = (7).to_bytes(8, "little") + bytes(24)
        assert synth.secret_from_seed(sid, seed) == o.secret_from_seed(S, seed)
    for n, w in [(1 << 20, 8), (10, 4), (7, 2), (0, 2), (5, 8)]:
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(lo % 32 == 0 or lo == n for lo, _ in b)

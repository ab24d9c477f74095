"""CPU: the batch server's multi-buffer SHA-512 (ark_vrf_b200/csrc/mbsha512.cpp) against hashlib - ragged stream
lengths (empty, sub-block, block-aligned, multi-MiB), more streams than lanes, odd chunkings.  The digest is the
batch seed of the reference (src/thin.rs:273-279), so it has to be bit-exact SHA-512."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from ark_vrf_b200 import _lib


def _hash_streams(streams, chunk):
    lib = _lib.load()
    n = len(streams)
    bufs = [np.frombuffer(s, dtype=np.uint8).copy() if len(s) else np.zeros(1, dtype=np.uint8) for s in streams]
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
    lens = (C.c_uint64 * max(n, 1))(*[len(s) for s in streams])
    out = np.zeros(64 * max(n, 1), dtype=np.uint8)
    simd = C.c_int32(-1)
    _lib.check(lib.avrf_mb_sha512(n, ptrs, lens, chunk, out.ctypes.data, C.byref(simd)))
    return [bytes(out[64 * i:64 * i + 64]) for i in range(n)], simd.value


@pytest.mark.parametrize("chunk", [1, 127, 128, 4096, 1 << 20])
def test_ragged_streams(chunk):
    rng = np.random.default_rng(chunk)
    lens = [0, 1, 27, 111, 112, 127, 128, 129, 255, 256, 1000, 4096, 65536 + 28, 300001]
    if chunk < 128:
        lens = [x for x in lens if x <= 4096]
    streams = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in lens]
    got, simd = _hash_streams(streams, chunk)
    assert simd in (0, 1)
    assert got == [hashlib.sha512(s).digest() for s in streams]


def test_batch_seed_shape():
    """Eight 'batch transcripts' of the reference's shape: SUITE_ID || 0x50 || 64 bytes per proof."""
    rng = np.random.default_rng(5)
    sid = b"Bandersnatch-SHA512-ELL2-v1"
    streams = [sid + b"\x50" + rng.integers(0, 256, size=64 * n, dtype=np.uint8).tobytes() for n in (1, 2, 3, 1000, 4097, 70000, 5, 64)]
    got, _ = _hash_streams(streams, 75776 * 64)
    assert got == [hashlib.sha512(s).digest() for s in streams]

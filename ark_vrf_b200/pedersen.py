"""Pedersen VRF batch verification on the same GPU engine.

Mirror of `ark_vrf::pedersen::BatchVerifier` (reference src/pedersen.rs:322-427):
`new()`, `push(ios, ad, proof)`, `push_prepared(item)`, `verify()`.  A proof is
`(pk_com, r, ok, s, sb)` (pedersen.rs:43-50); the MSM has 5N+2 points."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Union

import numpy as np

from . import _lib
from .thin import Format, Suite, Tap, _bytes_of, _ios_bytes, _raise_for_status, ptr


@dataclass
class Proof:                              # pedersen::Proof (src/pedersen.rs:43-50)
    pk_com: bytes
    r: bytes
    ok: bytes
    s: bytes
    sb: bytes


@dataclass
class BatchItem:                          # pedersen::BatchItem; hashing deferred to the GPU
    ios: bytes
    n_ios: int
    ad: bytes
    proof: Proof


class BatchVerifier:
    def __init__(self, suite: Union[Suite, int], fmt: Union[Format, int] = Format.CANONICAL):
        self._lib = _lib.load()
        self.suite, self.fmt = Suite(suite), Format(fmt)
        self._h = self._lib.avrf_pedersen_batch_new(int(self.suite), int(self.fmt))
        if not self._h:
            msg = self._lib.avrf_last_error()
            raise _lib.AvrfError(msg.decode() if msg else "avrf_pedersen_batch_new failed")
        self._n = 0

    @staticmethod
    def prepare(ios, ad, proof: Proof) -> BatchItem:
        iob, k = _ios_bytes(ios)
        return BatchItem(iob, k, bytes(ad), proof)

    def push_prepared(self, e: BatchItem) -> None:
        io_off = np.array([0, e.n_ios], dtype=np.uint32)
        ad_off = np.array([0, len(e.ad)], dtype=np.uint32)
        f = lambda b, n: np.frombuffer(_bytes_of(b, n), dtype=np.uint8).copy()
        self.push_many(np.frombuffer(e.ios + bytes(128), dtype=np.uint8).copy(), io_off,
                       np.frombuffer(e.ad + bytes(16), dtype=np.uint8).copy(), ad_off,
                       f(e.proof.pk_com, 64), f(e.proof.r, 64), f(e.proof.ok, 64), f(e.proof.s, 32), f(e.proof.sb, 32))

    def push(self, ios, ad, proof: Proof) -> None:
        self.push_prepared(self.prepare(ios, ad, proof))

    def push_many(self, ios, io_offsets, ad_blob, ad_offsets, pk_com, r, ok, s, sb) -> None:
        n = len(io_offsets) - 1
        _lib.check(self._lib.avrf_pedersen_batch_push_many(self._h, n, ptr(ios), ptr(io_offsets), ptr(ad_blob),
                                                           ptr(ad_offsets), ptr(pk_com), ptr(r), ptr(ok), ptr(s), ptr(sb)))
        self._n += n

    def clear(self) -> None:
        """Forget all pushed proofs, keep allocations."""
        _lib.check(self._lib.avrf_thin_batch_clear(self._h))
        self._n = 0

    def verify_status(self) -> int:
        st = C.c_int32(-1)
        _lib.check(self._lib.avrf_pedersen_batch_verify(self._h, C.byref(st)))
        return st.value

    def verify(self) -> None:
        _raise_for_status(self.verify_status())

    def tap(self, what: Tap) -> np.ndarray:
        n = self._n
        size = {Tap.C: 16 * n, Tap.W: 32 * n, Tap.SEED: 64, Tap.SCALARS: 32 * (5 * n + 2)}[Tap(what)]
        buf = np.zeros(max(size, 1), dtype=np.uint8)
        _lib.check(self._lib.avrf_thin_batch_tap(self._h, int(what), ptr(buf), buf.nbytes))
        return buf

    def __len__(self) -> int:
        return self._n

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.avrf_thin_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

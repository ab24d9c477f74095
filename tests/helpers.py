"""Shared helpers for the parity tests: oracle <-> byte-array conversions."""
import numpy as np

from oracle import pyref as o

GOLDEN_SEEDS = [1, 2, 3, 4, 5, 5, 6]           # reference src/testing.rs:291-299


def pt_bytes(P) -> bytes:
    return P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little")


def pt_from_bytes(b) -> tuple:
    b = bytes(b)
    return (int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little"))


def sc_bytes(k: int) -> bytes:
    return int(k).to_bytes(32, "little")


def to_mont_bytes(b: bytes, mod: int) -> bytes:
    return ((int.from_bytes(b, "little") << 256) % mod).to_bytes(32, "little")


def arrays_from_proofs(pr: "o.Proofs", montgomery: bool = False):
    """Oracle Proofs -> the arrays avrf_thin_batch_push_many takes (canonical or Montgomery)."""
    S = pr.suite
    n = len(pr.pk)

    def fp(x):
        return ((x << 256) % S.p if montgomery else x).to_bytes(32, "little")

    def fr(x):
        return ((x << 256) % S.r if montgomery else x).to_bytes(32, "little")

    def pb(P):
        return fp(P[0]) + fp(P[1])

    pk = np.frombuffer(b"".join(pb(P) for P in pr.pk), dtype=np.uint8).reshape(n, 64).copy()
    r = np.frombuffer(b"".join(pb(P) for P in pr.r), dtype=np.uint8).reshape(n, 64).copy()
    s = np.frombuffer(b"".join(fr(x) for x in pr.s), dtype=np.uint8).reshape(n, 32).copy()
    iob = b"".join(pb(i) + pb(oo) for ios in pr.ios for (i, oo) in ios)
    ios = np.frombuffer(iob + bytes(128), dtype=np.uint8).copy()
    io_off = np.zeros(n + 1, dtype=np.uint32)
    io_off[1:] = np.cumsum([len(x) for x in pr.ios])
    ad_off = np.zeros(n + 1, dtype=np.uint32)
    ad_off[1:] = np.cumsum([len(a) for a in pr.ad])
    ad = np.frombuffer(b"".join(pr.ad) + bytes(16), dtype=np.uint8).copy()
    return pk, ios, io_off, ad, ad_off, r, s


def oracle_items(pr: "o.Proofs"):
    return [o.batch_prepare(pr.suite, pr.pk[j], pr.ios[j], pr.ad[j], pr.r[j], pr.s[j]) for j in range(len(pr.pk))]


def golden_proofs(S, vectors) -> "o.Proofs":
    pr = o.Proofs(S)
    for v in vectors:
        pk = o.dec_point(S, bytes.fromhex(v["pk"]))
        h = o.dec_point(S, bytes.fromhex(v["h"]))
        gamma = o.dec_point(S, bytes.fromhex(v["gamma"]))
        pr.pk.append(pk)
        pr.ios.append([(h, gamma)])
        pr.ad.append(bytes.fromhex(v["ad"]))
        pr.r.append(o.dec_point(S, bytes.fromhex(v["proof_r"])))
        pr.s.append(int.from_bytes(bytes.fromhex(v["proof_s"]), "little"))
    return pr

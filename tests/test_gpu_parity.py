"""GPU parity tests: the CUDA engine (through the C ABI) against the oracle and the
reference's golden vectors.  Bar: bit-exact bytes for every intermediate that has a byte
representation (h, pk, gamma, beta, proof_r, proof_s, c, z, w, seed, MSM scalars) and the
same Ok / VerificationFailure / InvalidData verdict as the reference semantics."""
import os

import numpy as np
import pytest

from oracle import pyref as o
from helpers import (GOLDEN_SEEDS, arrays_from_proofs, golden_proofs, oracle_items, pt_bytes, pt_from_bytes, sc_bytes)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def av():
    import ark_vrf_b200 as a
    a.load().avrf_init(0)
    return a


def _push_all(av, S_id, pr, montgomery=False):
    bv = av.BatchVerifier(S_id, av.Format.MONTGOMERY if montgomery else av.Format.CANONICAL)
    bv.push_many(*arrays_from_proofs(pr, montgomery))
    return bv


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_golden_feeders(av, sid, golden):
    """h, pk, gamma, beta, proof_r, proof_s of every golden vector from the GPU kernels."""
    from ark_vrf_b200 import ops
    S = o.SUITES[sid]
    vs = golden[sid]
    sks = np.frombuffer(b"".join(bytes.fromhex(v["sk"]) for v in vs), dtype=np.uint8).reshape(-1, 32).copy()
    pk = ops.public_keys(sid, sks)
    enc = ops.point_compress(sid, pk)
    assert [bytes(e).hex() for e in enc] == [v["pk"] for v in vs]
    h, henc, ok = ops.hash_to_curve(sid, [bytes.fromhex(v["alpha"]) for v in vs], want_compressed=True)
    assert ok.all()
    assert [bytes(e).hex() for e in henc] == [v["h"] for v in vs]
    gamma = ops.vrf_output(sid, sks, h)
    assert [bytes(e).hex() for e in ops.point_compress(sid, gamma)] == [v["gamma"] for v in vs]
    assert [bytes(e).hex() for e in ops.point_to_hash(sid, gamma)] == [v["beta"] for v in vs]
    ios = np.ascontiguousarray(np.concatenate([h, gamma], axis=1))
    io_off = np.arange(len(vs) + 1, dtype=np.uint32)
    ads = [bytes.fromhex(v["ad"]) for v in vs]
    ad_off = np.zeros(len(vs) + 1, dtype=np.uint32)
    ad_off[1:] = np.cumsum([len(a) for a in ads])
    ad = np.frombuffer(b"".join(ads) + bytes(16), dtype=np.uint8).copy()
    r, s = ops.thin_prove_many(sid, sks, pk, ios, io_off, ad, ad_off)
    assert [bytes(e).hex() for e in ops.point_compress(sid, r)] == [v["proof_r"] for v in vs]
    assert [bytes(e).hex() for e in s] == [v["proof_s"] for v in vs]
    # same through the Montgomery (arkworks memory image) format
    skm = np.frombuffer(b"".join(((int.from_bytes(bytes.fromhex(v["sk"]), "little") << 256) % S.r).to_bytes(32, "little")
                                  for v in vs), dtype=np.uint8).reshape(-1, 32).copy()
    pkm = ops.public_keys(sid, skm, av.Format.MONTGOMERY)
    assert [bytes(e).hex() for e in ops.point_compress(sid, pkm, av.Format.MONTGOMERY)] == [v["pk"] for v in vs]


@pytest.mark.parametrize("sid", [0, 1, 2])
@pytest.mark.parametrize("montgomery", [False, True])
def test_golden_batch_and_taps(av, sid, montgomery, golden):
    S = o.SUITES[sid]
    pr = golden_proofs(S, golden[sid])
    items = oracle_items(pr)
    bv = _push_all(av, sid, pr, montgomery)
    assert len(bv) == 7
    assert bv.verify_status() == 0
    bv.verify()                                   # repeatable (benches/thin.rs:84-88)
    c = bv.tap(av.Tap.C).reshape(-1, 16)
    z = bv.tap(av.Tap.Z).reshape(-1, 16)
    assert [bytes(x) for x in c] == [e.c.to_bytes(16, "little") for e in items]
    assert [bytes(x) for x in z] == [e.zs[1].to_bytes(16, "little") for e in items]
    seed = o.batch_seed(S, items)
    assert bytes(bv.tap(av.Tap.SEED)) == seed
    w = bv.tap(av.Tap.W).reshape(-1, 16)
    assert [bytes(x) for x in w] == [x.to_bytes(16, "little") for x in o.batch_weights(S, seed, 7)]
    renc = bv.tap(av.Tap.R_COMPRESSED).reshape(-1, 32)
    assert [bytes(x).hex() for x in renc] == [v["proof_r"] for v in golden[sid]]
    _, scalars = o.batch_msm_terms(S, items)
    sc = bv.tap(av.Tap.SCALARS).reshape(-1, 32)
    assert [bytes(x) for x in sc] == [sc_bytes(k) for k in scalars]
    if sid == 0:
        assert bytes(w[0]).hex() == "b2a5366d770f3656a54dd068c18ccc9d"
        assert bytes(c[0]).hex() == "08be086526bd2ca18d27746c16fc8a55"
    # partial point equals the oracle's MSM value
    part = bytes(bv.tap(av.Tap.PARTIAL))
    Rm = 1 << 256
    X, Y, Z, T = [int.from_bytes(part[32 * i:32 * i + 32], "little") * pow(Rm, -1, S.p) % S.p for i in range(4)]
    assert X == 0 and Y == Z                       # identity (valid batch)


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_golden_single_verify(av, sid, golden):
    S = o.SUITES[sid]
    for v in golden[sid]:
        pk = o.dec_point(S, bytes.fromhex(v["pk"]))
        h = o.dec_point(S, bytes.fromhex(v["h"]))
        g = o.dec_point(S, bytes.fromhex(v["gamma"]))
        R = o.dec_point(S, bytes.fromhex(v["proof_r"]))
        proof = av.Proof(pt_bytes(R), bytes.fromhex(v["proof_s"]))
        pub = av.Public(sid, pt_bytes(pk))
        pub.verify((pt_bytes(h), pt_bytes(g)), bytes.fromhex(v["ad"]), proof)
        with pytest.raises(av.VerificationFailure):
            pub.verify((pt_bytes(h), pt_bytes(g)), bytes.fromhex(v["ad"]) + b"x", proof)


@pytest.mark.parametrize("sid,m", [(0, 1), (0, 3), (0, 0), (1, 1), (2, 4)])
def test_random_batch_vs_oracle(av, sid, m):
    """Seeded synthetic batch (oracle-generated): taps and verdicts, then injected faults."""
    S = o.SUITES[sid]
    n = 24 if m <= 1 else 10
    pr = o.synth_proofs(S, n, m, signers=5)
    items = oracle_items(pr)
    bv = _push_all(av, sid, pr)
    assert bv.verify_status() == o.batch_verify(S, items) == 0
    c = bv.tap(av.Tap.C).reshape(-1, 16)
    assert [bytes(x) for x in c] == [e.c.to_bytes(16, "little") for e in items]
    if m:
        z = bv.tap(av.Tap.Z).reshape(-1, 16)
        assert [bytes(x) for x in z] == [zz.to_bytes(16, "little") for e in items for zz in e.zs[1:]]
    _, scalars = o.batch_msm_terms(S, items)
    sc = bv.tap(av.Tap.SCALARS).reshape(-1, 32)
    assert [bytes(x) for x in sc] == [sc_bytes(k) for k in scalars]

    def status_of(mut):
        import copy
        q = copy.deepcopy(pr)
        mut(q)
        b2 = _push_all(av, sid, q)
        st = b2.verify_status()
        assert st == o.batch_verify(S, oracle_items(q))
        return st

    def msm_value_matches(mut):
        """The MSM sum itself (thin.rs:319) - a non-trivial group element for a bad batch - against the oracle."""
        import copy
        q = copy.deepcopy(pr)
        mut(q)
        its = oracle_items(q)
        b2 = _push_all(av, sid, q)
        assert b2.verify_status() == 1
        part = bytes(b2.tap(av.Tap.PARTIAL))
        Rinv = pow(1 << 256, -1, S.p)
        X, Y, Z, T = [int.from_bytes(part[32 * i:32 * i + 32], "little") * Rinv % S.p for i in range(4)]
        zi = pow(Z, -1, S.p)
        bases, scalars = o.batch_msm_terms(S, its)
        want = o.IDENTITY
        for P, k in zip(bases, scalars):
            want = o.pt_add(S, want, o.pt_mul_raw(S, P, k))
        assert (X * zi % S.p, Y * zi % S.p) == tuple(want) and want != o.IDENTITY
        assert T * Z % S.p == X * Y % S.p

    def bad_s(q): q.s[3] = (q.s[3] + 1) % S.r
    msm_value_matches(bad_s)
    def bad_ad(q): q.ad[n - 1] = q.ad[n - 1] + b"!"
    def bad_r(q): q.r[0] = o.pt_add(S, q.r[0], S.G)
    def id_pk(q): q.pk[2] = o.IDENTITY
    def id_pk_and_bad_s(q): q.pk[2] = o.IDENTITY; q.s[5] = (q.s[5] + 1) % S.r
    assert status_of(bad_s) == 1
    assert status_of(bad_ad) == 1
    assert status_of(bad_r) == 1
    assert status_of(id_pk) == 2
    assert status_of(id_pk_and_bad_s) == 2        # InvalidData takes precedence (thin.rs:266-271)
    if m >= 1:
        def swap_o(q): q.ios[1][0], q.ios[2][0] = (q.ios[1][0][0], q.ios[2][0][1]), (q.ios[2][0][0], q.ios[1][0][1])
        assert status_of(swap_o) == 1
    if m >= 2:
        def id_in(q): q.ios[4][m - 1] = (o.IDENTITY, q.ios[4][m - 1][1])
        assert status_of(id_in) == 2


def test_reference_batch_scenarios(av):
    """reference src/thin.rs:346-384, 418-471 re-enacted through push / prepare+push_prepared."""
    sid, S = 0, o.BANDERSNATCH
    sk = o.secret_from_seed(S, bytes(32))
    pk = o.public_key(S, sk)
    inp = o.data_to_point(S, b"foo-input")
    io = (inp, o.pt_mul(S, inp, sk))
    iob = (pt_bytes(io[0]), pt_bytes(io[1]))
    p1 = o.thin_prove(S, sk, [io], b"foo")
    p2 = o.thin_prove(S, sk, [io], b"bar")
    P1 = av.Proof(pt_bytes(p1[0]), sc_bytes(p1[1]))
    P2 = av.Proof(pt_bytes(p2[0]), sc_bytes(p2[1]))
    bv = av.BatchVerifier(sid)
    bv.push(pt_bytes(pk), iob, b"foo", P1)
    bv.push(pt_bytes(pk), iob, b"bar", P2)
    bv.verify()
    bv = av.BatchVerifier(sid)
    bv.push_prepared(av.BatchVerifier.prepare(pt_bytes(pk), iob, b"foo", P1))
    bv.push_prepared(av.BatchVerifier.prepare(pt_bytes(pk), [iob], b"bar", P2))
    bv.verify()
    av.BatchVerifier(sid).verify()                # empty batch is Ok (thin.rs:262-264)
    bv = av.BatchVerifier(sid)
    bv.push(pt_bytes(pk), iob, b"foo", P1)
    bv.push(pt_bytes(pk), iob, b"wrong", P2)
    with pytest.raises(av.VerificationFailure):
        bv.verify()
    # identity public key forgery (thin.rs:418-433)
    sf = 0x5EED
    forged = av.Proof(pt_bytes(o.pt_mul(S, S.G, sf)), sc_bytes(sf))
    with pytest.raises(av.InvalidData):
        av.Public(sid, pt_bytes(o.IDENTITY)).verify([], b"forgery", forged)
    bv = av.BatchVerifier(sid)
    bv.push(pt_bytes(o.IDENTITY), [], b"forgery", forged)
    with pytest.raises(av.InvalidData):
        bv.verify()
    # R may be the identity without being InvalidData (thin.rs:90-94): it is just a failing proof
    bv = av.BatchVerifier(sid)
    bv.push(pt_bytes(pk), iob, b"foo", av.Proof(pt_bytes(o.IDENTITY), sc_bytes(p1[1])))
    with pytest.raises(av.VerificationFailure):
        bv.verify()
    # incremental pushes after a verify keep earlier items (verify takes &self)
    bv = av.BatchVerifier(sid)
    bv.push(pt_bytes(pk), iob, b"foo", P1)
    bv.verify()
    bv.push(pt_bytes(pk), iob, b"bar", P2)
    assert len(bv) == 2
    bv.verify()


@pytest.mark.parametrize("sid,m,n", [(0, 1, 20000), (1, 1, 6000), (2, 4, 3000)])
def test_generated_batch_properties(av, sid, m, n):
    """GPU-generated workload (ark_vrf_b200.synth): all-valid accepts; each fault kind rejects;
    a sample of the generated proofs is re-verified by the oracle."""
    from ark_vrf_b200 import synth
    S = o.SUITES[sid]
    b = synth.make_batch(sid, n, m, fmt=av.Format.CANONICAL)
    for j in [0, 1, n // 2, n - 1]:
        ios = [(pt_from_bytes(b.ios[j * m + i][:64]), pt_from_bytes(b.ios[j * m + i][64:])) for i in range(m)]
        ad = bytes(b.ad_blob[b.ad_offsets[j]:b.ad_offsets[j + 1]])
        assert ad == b"ad-%d" % j
        sk = o.secret_from_seed(S, o.synth_seed(j % 4096))
        assert pt_from_bytes(b.pk[j]) == o.public_key(S, sk)
        assert ios[0][0] == o.data_to_point(S, o.synth_msg(j, 0))
        R, s = o.thin_prove(S, sk, ios, ad)
        assert pt_from_bytes(b.r[j]) == R and int.from_bytes(bytes(b.s[j]), "little") == s
        assert o.thin_verify(S, pt_from_bytes(b.pk[j]), ios, ad, R, s) == 0
    bv = av.BatchVerifier(sid, av.Format.CANONICAL)
    args = lambda: (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    bv.push_many(*args())
    assert bv.verify_status() == 0
    # weights tap at full size against the reference stream definition
    c = bv.tap(av.Tap.C).reshape(-1, 16)
    stream = np.concatenate([c, np.zeros((n, 16), np.uint8), b.s], axis=1)
    import hashlib
    seed = hashlib.sha512(S.suite_id + b"\x50" + stream.tobytes()).digest()
    assert bytes(bv.tap(av.Tap.SEED)) == seed
    w = bv.tap(av.Tap.W).reshape(-1, 16)
    for j in [0, 1, 2, 3, 4, n - 1]:
        blk = hashlib.sha512(seed + (j // 4).to_bytes(8, "little")).digest()
        assert bytes(w[j]) == blk[16 * (j % 4):16 * (j % 4) + 16]
    # faults at the first and last item and at splitmix64 positions
    pos = sorted({0, n - 1, synth.splitmix64(0xBAD5EED) % n, synth.splitmix64(0xBAD5EED + 1) % n})
    for p in pos:
        s2 = b.s.copy()
        s2[p, 0] ^= 1
        bv.clear()
        bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
        assert bv.verify_status() == 1
    ad2 = b.ad_blob.copy()
    ad2[b.ad_offsets[pos[1]]] ^= 0x20
    bv.clear()
    bv.push_many(b.pk, b.ios, b.io_offsets, ad2, b.ad_offsets, b.r, b.s)
    assert bv.verify_status() == 1
    pk2 = b.pk.copy()
    pk2[pos[-1]] = np.frombuffer(pt_bytes(o.IDENTITY), dtype=np.uint8)
    bv.clear()
    bv.push_many(pk2, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    assert bv.verify_status() == 2
    if m > 1:
        io2 = b.ios.copy()
        io2[pos[1] * m + m - 1, :64] = np.frombuffer(pt_bytes(o.IDENTITY), dtype=np.uint8)
        bv.clear()
        bv.push_many(b.pk, io2, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
        assert bv.verify_status() == 2
    # the remaining reject kinds of SURVEY 8(d)
    def status(pk=b.pk, ios=b.ios, r=b.r, s_=b.s):
        bv.clear()
        bv.push_many(pk, ios, b.io_offsets, b.ad_blob, b.ad_offsets, r, s_)
        return bv.verify_status()
    if m >= 1:                                               # O swapped with a neighbour's
        io2 = b.ios.copy()
        q = pos[1] if pos[1] + 1 < n else pos[1] - 1
        io2[q * m, 64:], io2[(q + 1) * m, 64:] = b.ios[(q + 1) * m, 64:].copy(), b.ios[q * m, 64:].copy()
        assert status(ios=io2) == 1
    r2 = b.r.copy()                                          # R <- R + G
    r2[pos[-1]] = np.frombuffer(pt_bytes(o.pt_add(S, pt_from_bytes(b.r[pos[-1]]), S.G)), dtype=np.uint8)
    assert status(r=r2) == 1
    rng = np.random.default_rng(0xBAD5EED)                   # 1 % random faults
    s3 = b.s.copy()
    s3[rng.choice(n, size=max(1, n // 100), replace=False), 1] ^= 0x40
    assert status(s_=s3) == 1
    s4 = b.s.copy()                                          # identity pk AND a bad s: InvalidData wins
    s4[0, 0] ^= 1
    assert status(pk=pk2, s_=s4) == 2
    assert status() == 0
    # sharding property: partial sums of two shards, weights from the global seed.
    # Every valid proof contributes the identity, so each shard's partial is the identity; a
    # bad proof in shard 1 makes that partial (and the combination) non-trivial.
    bv.clear()
    bv.push_many(*args())
    assert bv.prepare_device() is False
    seed2 = av.seed_of_stream(sid, bv.cs_stream())
    assert seed2 == seed
    half = (n // 8) * 4

    def shard_partials(svals, seed_):
        parts = b""
        for lo, hi in [(0, half), (half, n)]:
            sh = av.BatchVerifier(sid, av.Format.CANONICAL)
            sh.push_many(b.pk[lo:hi], b.ios[lo * m:hi * m], (b.io_offsets[lo:hi + 1] - b.io_offsets[lo]).astype(np.uint32),
                         b.ad_blob[b.ad_offsets[lo]:], (b.ad_offsets[lo:hi + 1] - b.ad_offsets[lo]).astype(np.uint32),
                         b.r[lo:hi], np.ascontiguousarray(svals[lo:hi]))
            sh.prepare_device()
            parts += sh.partial(seed_, lo)
        return parts
    parts = shard_partials(b.s, seed)
    assert av.combine_partials(sid, parts) == 0
    assert av.combine_partials(sid, parts[:128]) == 0 and av.combine_partials(sid, parts[128:]) == 0
    s2 = b.s.copy()
    s2[n - 2, 0] ^= 1
    stream2 = np.concatenate([c, np.zeros((n, 16), np.uint8), s2], axis=1)
    seed_bad = hashlib.sha512(S.suite_id + b"\x50" + stream2.tobytes()).digest()
    parts = shard_partials(s2, seed_bad)
    assert av.combine_partials(sid, parts[:128]) == 0
    assert av.combine_partials(sid, parts[128:]) == 1
    assert av.combine_partials(sid, parts) == 1


FULL = os.environ.get("AVRF_FULL_CONFIGS", "1") == "1"      # BASELINE.json sizes by default; AVRF_FULL_CONFIGS=0: quick run


@pytest.mark.parametrize("sid,m,log2n", [(0, 1, 20 if FULL else 16), (1, 1, 20 if FULL else 16), (2, 4, 22 if FULL else 15)])
def test_baseline_configs(av, sid, m, log2n):
    """BASELINE.json configs[1..3] (full sizes with AVRF_FULL_CONFIGS=1, else scaled down): size-independent
    properties - an all-valid batch accepts, one tampered response anywhere rejects, an identity input
    is InvalidData, and the batch seed equals SHA-512 of the reference's stream definition."""
    import hashlib
    import time
    from ark_vrf_b200 import synth
    S = o.SUITES[sid]
    n = 1 << log2n
    t0 = time.time()
    b = synth.make_batch(sid, n, m, fmt=av.Format.MONTGOMERY)
    bv = av.BatchVerifier(sid, av.Format.MONTGOMERY)
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    t1 = time.time()
    assert bv.verify_status() == 0
    t2 = time.time()
    cs = bv.cs_stream()
    assert bytes(bv.tap(av.Tap.SEED)) == hashlib.sha512(S.suite_id + b"\x50" + cs.tobytes()).digest()
    # canonical s of the stream equals the Montgomery input converted back
    j = n - 1
    s_can = int.from_bytes(bytes(cs[j][32:]), "little")
    assert (s_can << 256) % S.r == int.from_bytes(bytes(b.s[j]), "little")
    pos = synth.splitmix64(0xBAD5EED) % n
    s2 = b.s.copy()
    s2[pos] = np.frombuffer((((s_can if pos == j else int.from_bytes(bytes(cs[pos][32:]), "little")) + 1) % S.r << 256).__mod__(S.r).to_bytes(32, "little"), dtype=np.uint8)
    bv.clear()
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
    assert bv.verify_status() == 1
    io2 = b.ios.copy()
    ident = np.zeros(64, dtype=np.uint8)
    ident[32:] = np.frombuffer(((1 << 256) % S.p).to_bytes(32, "little"), dtype=np.uint8)     # (0, 1) in Montgomery form
    io2[(n - 1) * m + (m - 1), :64] = ident
    bv.clear()
    bv.push_many(b.pk, io2, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    assert bv.verify_status() == 2
    print(f"\nconfig suite={sid} M={m} N=2^{log2n}: generate+push {t1 - t0:.2f}s, verify {t2 - t1:.3f}s, {bv.timings()}")


@pytest.mark.parametrize("log2n", [24 if FULL else 14])
def test_bulk_h2c_and_output(av, log2n):
    """BASELINE.json configs[4]: bulk Elligator2 hash-to-curve + VRF output on Bandersnatch; a sample is
    checked against the oracle, the rest through the compressed encodings being valid points."""
    from ark_vrf_b200 import ops, synth
    S = o.BANDERSNATCH
    n = 1 << log2n
    chunk = min(n, 1 << 22)
    sk = synth.secret_from_seed(0, bytes(32))
    skb = np.frombuffer(sk.to_bytes(32, "little"), dtype=np.uint8).copy()
    for first in range(0, n, chunk):
        j = np.arange(first, first + chunk, dtype=np.uint64)
        blob = np.concatenate([j.view(np.uint8), np.zeros(16, np.uint8)])
        off = (np.arange(chunk + 1, dtype=np.uint64) * 8).astype(np.uint32)
        pts, enc, ok = ops.hash_to_curve(0, blob, off, want_compressed=True)
        assert ok.all()
        out = ops.vrf_output(0, skb, pts)
        for q in [0, chunk // 3, chunk - 1]:
            h = o.hash_to_curve_ell2(S, int(first + q).to_bytes(8, "little"))
            assert pt_from_bytes(pts[q]) == h and bytes(enc[q]) == o.enc_point(S, h)
            assert pt_from_bytes(out[q]) == o.pt_mul(S, h, sk)


def test_sharded_inputs_outputs_single_process(av):
    """configs[4] helper without a process group: whole range on this GPU; inputs, outputs and the checksum
    against the oracle."""
    from ark_vrf_b200 import dist as avdist
    S = o.BANDERSNATCH
    sk = o.secret_from_seed(S, bytes(32))
    n = 96
    lo, hi, inp, outp, dg = avdist.sharded_inputs_outputs(0, n, np.frombuffer(sk.to_bytes(32, "little"), dtype=np.uint8),
                                                          fmt=int(av.Format.CANONICAL))
    assert (lo, hi) == (0, n)
    x = 0
    for j in range(n):
        P = o.data_to_point(S, j.to_bytes(8, "little"))
        Q = o.pt_mul(S, P, sk)
        assert pt_from_bytes(inp[j]) == P and pt_from_bytes(outp[j]) == Q
        e = o.enc_point(S, Q)
        for k in range(4):
            x ^= int.from_bytes(e[8 * k:8 * k + 8], "little")
    assert dg == x


@pytest.mark.parametrize("sid,m,n", [(0, 1, 100), (2, 2, 37)])
def test_tree_weights_mode(av, sid, m, n):
    """Opt-in AVRF_WEIGHTS_TREE: the seed follows its documented definition (oracle batch_seed_tree),
    verdicts are unchanged, and the default mode still yields the reference's seed."""
    S = o.SUITES[sid]
    pr = o.synth_proofs(S, n, m, signers=3)
    items = oracle_items(pr)
    bv = _push_all(av, sid, pr)
    bv.set_weights_mode(1)
    assert bv.verify_status() == 0
    assert bytes(bv.tap(av.Tap.SEED)) == o.batch_seed_tree(S, items)
    leaves = bv.tree_leaves(0)
    assert av.thin.seed_of_tree(sid, n, leaves) == o.batch_seed_tree(S, items)
    bv.set_weights_mode(0)
    assert bv.verify_status() == 0
    assert bytes(bv.tap(av.Tap.SEED)) == o.batch_seed(S, items)
    pr.s[n // 2] = (pr.s[n // 2] + 1) % S.r
    b2 = _push_all(av, sid, pr)
    b2.set_weights_mode(1)
    assert b2.verify_status() == 1
    pr.pk[0] = o.IDENTITY
    b3 = _push_all(av, sid, pr)
    b3.set_weights_mode(1)
    assert b3.verify_status() == 2


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_wire_format_ingest(av, sid, golden):
    """bulk CanonicalDeserialize with validation vs the oracle: golden encodings decode
    to the oracle's points; y >= p, non-residues, low-order and cofactor-tainted points, and (for
    Public/Input/Output) the identity are rejected exactly where the oracle rejects them."""
    import random
    from ark_vrf_b200 import ops
    S = o.SUITES[sid]
    rnd = random.Random(7 + sid)
    encs = []
    for v in golden[sid]:
        encs += [bytes.fromhex(v[k]) for k in ("pk", "h", "gamma", "proof_r")]
    encs.append(o.enc_point(S, o.IDENTITY))
    encs.append(((S.p - 1)).to_bytes(32, "little"))                     # (0, -1): order 2
    encs.append((S.p).to_bytes(32, "little"))                            # y = p: not canonical
    encs.append((S.p + 5).to_bytes(32, "little") if S.p + 5 < (1 << 255) else (S.p).to_bytes(32, "little"))
    for _ in range(60):                                                  # random y: mostly off-subgroup or no root
        y = rnd.randrange(S.p)
        b = bytearray(y.to_bytes(32, "little"))
        if rnd.random() < 0.5:
            b[31] |= 0x80
        encs.append(bytes(b))
    for _ in range(10):                                                  # genuine subgroup points
        encs.append(o.enc_point(S, o.pt_mul(S, S.G, rnd.randrange(1, S.r))))
    arr = np.frombuffer(b"".join(encs), dtype=np.uint8).reshape(-1, 32)
    for kind in (0, 1):
        pts, ok = ops.points_deserialize(sid, arr, kind=kind)
        for j, e in enumerate(encs):
            want = o.deserialize_point(S, e, reject_identity=(kind == 1))
            assert bool(ok[j]) == (want is not None), (sid, kind, j, e.hex())
            if want is not None:
                assert pt_from_bytes(pts[j]) == want
    assert ok[:28].all()                                                  # every golden encoding is valid


@pytest.mark.parametrize("sid,m", [(0, 1), (2, 3), (1, 0)])
def test_verify_each(av, sid, m):
    """per-proof verdicts equal thin::Verifier::verify of the oracle on every item."""
    S = o.SUITES[sid]
    n = 12
    pr = o.synth_proofs(S, n, m, signers=4)
    pr.s[3] = (pr.s[3] + 1) % S.r
    pr.ad[7] += b"x"
    pr.pk[9] = o.IDENTITY
    pr.r[10] = o.pt_add(S, pr.r[10], S.G)
    if m:
        pr.ios[5][m - 1] = (pr.ios[5][m - 1][0], o.pt_add(S, pr.ios[5][m - 1][1], S.G))
    bv = _push_all(av, sid, pr)
    want = [o.thin_verify(S, pr.pk[j], pr.ios[j], pr.ad[j], pr.r[j], pr.s[j]) for j in range(n)]
    got = list(bv.verify_each())
    assert got == want
    assert want[3] == 1 and want[7] == 1 and want[9] == 2 and want[10] == 1 and want[0] == 0
    assert bv.verify_status() == 2                                        # batch: identity pk dominates


def test_find_invalid(av):
    from ark_vrf_b200 import synth
    n = 5000
    b = synth.make_batch(0, n, 1, fmt=av.Format.MONTGOMERY)
    s2 = b.s.copy()
    bad = [0, 17, 2500, n - 1]
    for j in bad:
        s2[j, 1] ^= 2
    bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
    assert bv.verify_status() == 1
    assert list(bv.find_invalid()) == bad
    bv.clear()
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    assert bv.verify_status() == 0 and len(bv.find_invalid()) == 0


def test_incremental_pushes_and_reuse(av):
    """The eager push pipeline across several pushes (bulk, single, mixed M), clear / invalidate / re-verify:
    seed and verdict must always equal the oracle's for the proofs currently in the batch."""
    sid, S = 0, o.BANDERSNATCH
    pr1 = o.synth_proofs(S, 9, 1, signers=2)
    pr2 = o.synth_proofs(S, 5, 2, signers=2)
    pr3 = o.synth_proofs(S, 3, 0, signers=2)
    allp = o.Proofs(S)
    bv = av.BatchVerifier(sid)

    def extend(pr):
        for f in ("pk", "ios", "ad", "r", "s"):
            getattr(allp, f).extend(getattr(pr, f))

    def check(expect=0):
        items = oracle_items(allp)
        assert bv.verify_status() == expect == o.batch_verify(S, items)
        if expect != 2:
            assert bytes(bv.tap(av.Tap.SEED)) == o.batch_seed(S, items)
            _, scalars = o.batch_msm_terms(S, items)
            sc = bv.tap(av.Tap.SCALARS).reshape(-1, 32)
            assert [bytes(x) for x in sc] == [sc_bytes(k) for k in scalars]
    bv.push_many(*arrays_from_proofs(pr1)); extend(pr1); check()
    bv.push_many(*arrays_from_proofs(pr2)); extend(pr2); check()          # second bulk push, different M
    for j in range(3):                                                    # single pushes on top
        bv.push(pt_bytes(pr3.pk[j]), [], pr3.ad[j], av.Proof(pt_bytes(pr3.r[j]), sc_bytes(pr3.s[j])))
    extend(pr3); check()
    check()                                                               # verify is repeatable
    bv.invalidate(); check()                                              # full re-prepare on resident inputs
    bad = o.synth_proofs(S, 2, 1, signers=2)
    bad.s[1] = (bad.s[1] + 1) % S.r
    bv.push_many(*arrays_from_proofs(bad)); extend(bad); check(1)         # a bad proof pushed after a verify
    bv.clear()
    allp = o.Proofs(S)
    assert bv.verify_status() == 0 and len(bv) == 0                       # empty again
    bv.push_many(*arrays_from_proofs(pr2)); extend(pr2); check()          # reuse after clear
    # lazy (non-eager) handle gives the same results
    lz = av.BatchVerifier(sid, eager_seed=False)
    lz.push_many(*arrays_from_proofs(pr2))
    assert lz.verify_status() == 0
    assert bytes(lz.tap(av.Tap.SEED)) == bytes(bv.tap(av.Tap.SEED))


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_pedersen_batch(av, sid):
    """pedersen::BatchVerifier on the same engine.  Golden vectors
    (tests/golden/*_pedersen.json) accept; c, the (t_i, u_i) weights, the seed and all 5N+2 MSM
    scalars are bit-exact against the oracle; faults and identities give the reference's verdicts."""
    import json
    from ark_vrf_b200 import pedersen as ped
    S = o.SUITES[sid]
    vs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"{S.name}_pedersen.json")))
    cases = []
    for v in vs:
        h = o.dec_point(S, bytes.fromhex(v["h"]))
        g = o.dec_point(S, bytes.fromhex(v["gamma"]))
        pf = o.PedersenProof(o.dec_point(S, bytes.fromhex(v["proof_pk_com"])), o.dec_point(S, bytes.fromhex(v["proof_r"])),
                             o.dec_point(S, bytes.fromhex(v["proof_ok"])), int.from_bytes(bytes.fromhex(v["proof_s"]), "little"),
                             int.from_bytes(bytes.fromhex(v["proof_sb"]), "little"))
        cases.append(([(h, g)], bytes.fromhex(v["ad"]), pf))
    sk = o.secret_from_seed(S, bytes(32))
    for m in (0, 3):                                     # multi-pair (merged on the GPU) and zero-pair proofs
        ios = []
        for i in range(m):
            inp = o.data_to_point(S, bytes([i + 1]))
            ios.append((inp, o.pt_mul(S, inp, sk)))
        pf, _ = o.pedersen_prove(S, sk, ios, b"bar")
        cases.append((ios, b"bar", pf))

    def run(cs):
        bv = ped.BatchVerifier(sid)
        for ios, ad, pf in cs:
            bv.push([(pt_bytes(a), pt_bytes(b)) for a, b in ios], ad,
                    ped.Proof(pt_bytes(pf.pk_com), pt_bytes(pf.r), pt_bytes(pf.ok), sc_bytes(pf.s), sc_bytes(pf.sb)))
        items = [o.pedersen_batch_prepare(S, ios, ad, pf) for ios, ad, pf in cs]
        st = bv.verify_status()
        assert st == o.pedersen_batch_verify(S, items)
        return bv, items, st
    bv, items, st = run(cases)
    assert st == 0
    c = bv.tap(av.Tap.C).reshape(-1, 16)
    assert [bytes(x) for x in c] == [e.c.to_bytes(16, "little") for e in items]
    assert bytes(bv.tap(av.Tap.SEED)) == o.pedersen_batch_seed(S, items)
    _, scalars = o.pedersen_batch_terms(S, items)
    sc = bv.tap(av.Tap.SCALARS).reshape(-1, 32)
    assert [bytes(x) for x in sc] == [sc_bytes(k) for k in scalars]
    import copy
    bad = copy.deepcopy(cases); bad[1][2].sb = (bad[1][2].sb + 1) % S.r
    assert run(bad)[2] == 1
    bad = copy.deepcopy(cases); bad[2][2].s = (bad[2][2].s + 1) % S.r
    assert run(bad)[2] == 1
    bad = copy.deepcopy(cases); bad[0] = (bad[0][0], bad[0][1] + b"x", bad[0][2])
    assert run(bad)[2] == 1
    bad = copy.deepcopy(cases); bad[3][2].pk_com = o.IDENTITY
    assert run(bad)[2] == 2
    bad = copy.deepcopy(cases); bad[-1][0][1] = (o.IDENTITY, bad[-1][0][1][1])
    assert run(bad)[2] == 2
    assert ped.BatchVerifier(sid).verify_status() == 0


@pytest.mark.parametrize("sid,m,n", [(0, 1, 4096), (0, 3, 1024), (1, 1, 2048), (2, 4, 1024)])
def test_mid_size_taps_vs_c_oracle(av, sid, m, n):
    """Every tap (c, z, seed, w, all MSM scalars) of a few-thousand-proof batch is bit-exact against
    the C oracle (sizes beyond the Python oracle's reach)."""
    from oracle import corc
    arrs = corc.synth_batch(sid, n, m, signers=4096, nthreads=8)
    st, _, taps = corc.thin_batch_verify(sid, *arrs, nthreads=8, taps=True)
    assert st == 0
    bv = av.BatchVerifier(sid, av.Format.CANONICAL)
    bv.push_many(*[np.ascontiguousarray(a) for a in arrs])
    assert bv.verify_status() == 0
    assert (bv.tap(av.Tap.C).reshape(-1, 16) == taps["c"]).all()
    assert (bv.tap(av.Tap.Z).reshape(-1, 16) == taps["z"]).all()
    assert bytes(bv.tap(av.Tap.SEED)) == taps["seed"].tobytes()
    assert (bv.tap(av.Tap.W).reshape(-1, 16) == taps["w"]).all()
    assert (bv.tap(av.Tap.SCALARS).reshape(-1, 32) == taps["scalars"]).all()
    # and the GPU-generated workload is byte-identical to the oracle-generated one
    from ark_vrf_b200 import synth
    b = synth.make_batch(sid, min(n, 512), m, fmt=av.Format.CANONICAL)
    k = min(n, 512)
    assert (b.pk == arrs[0][:k]).all() and (b.ios == arrs[1][:k * m]).all()
    assert (b.r == arrs[5][:k]).all() and (b.s == arrs[6][:k]).all()


def test_async_verify_two_handles(av):
    """verify_async / verify_wait with two handles in flight (the serving pattern): verdicts are those of
    the synchronous call, whatever is pushed on the other handle meanwhile."""
    from ark_vrf_b200 import synth
    n = 70000
    good = synth.make_batch(0, n, 1, fmt=av.Format.MONTGOMERY)
    bad_s = good.s.copy()
    bad_s[n - 3, 0] ^= 1
    A = av.BatchVerifier(0, av.Format.MONTGOMERY)
    B = av.BatchVerifier(0, av.Format.MONTGOMERY)
    args = lambda s_: (good.pk, good.ios, good.io_offsets, good.ad_blob, good.ad_offsets, good.r, s_)
    expect = []
    got = []
    A.push_many(*args(good.s)); A.verify_async(); expect.append(0)
    for it in range(4):
        cur, other = (B, A) if it % 2 == 0 else (A, B)
        s_ = bad_s if it in (1, 2) else good.s
        cur.clear()
        cur.push_many(*args(s_))              # overlaps the other handle's MSM
        got.append(other.verify_wait())
        cur.verify_async()
        expect.append(1 if it in (1, 2) else 0)
    got.append((A if 3 % 2 == 1 else B).verify_wait())
    assert got == expect
    # a call on a handle with a verify in flight completes it first
    A.verify_async()
    assert A.verify_status() in (0, 1)
    assert len(A) == n
    empty = av.BatchVerifier(0)
    empty.verify_async()
    assert empty.verify_wait() == 0


def test_concurrent_handles_from_threads(av):
    """Different handles driven from different host threads at the same time (BatchVerifier is plain owned
    data, Send + Sync: thin.rs:188-198).  Every thread owns one handle (own CUDA streams) and runs several
    whole verifications; verdicts, seeds and per-proof challenges must be those of the same batches
    verified one at a time."""
    import threading
    from ark_vrf_b200 import synth
    sizes = [30011, 70000, 4097, 65536, 12345]
    suites = [0, 0, 2, 1, 0]
    jobs = []
    for k, (n, sid) in enumerate(zip(sizes, suites)):
        b = synth.make_batch(sid, n, 1, fmt=av.Format.MONTGOMERY, first=1000 * k)
        s_bad = b.s.copy()
        s_bad[(7 * n) // 11, 0] ^= 1
        jobs.append((sid, b, s_bad))

    def run(sid, b, s_, taps):
        h = av.BatchVerifier(sid, av.Format.MONTGOMERY)
        out = []
        for rep in range(3):
            bad = rep == 1
            h.clear()
            h.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s_ if bad else b.s)
            out.append(h.verify_status())
        taps.append((out, bytes(h.tap(av.Tap.SEED)), h.tap(av.Tap.C).copy()))
        h.close()

    serial = []
    for sid, b, s_bad in jobs:
        run(sid, b, s_bad, serial)
    conc = [[] for _ in jobs]
    errs = []

    def guarded(i):
        try:
            run(*jobs[i], conc[i])
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    ts = [threading.Thread(target=guarded, args=(i,)) for i in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(len(jobs)):
        assert conc[i][0][0] == serial[i][0] == [0, 1, 0]
        assert conc[i][0][1] == serial[i][1]
        assert (conc[i][0][2] == serial[i][2]).all()


@pytest.mark.parametrize("workers,hashers", [(3, 0), (3, 1), (10, 1)])     # (10, 1): two workers find no free hash lane
def test_batch_server(av, workers, hashers):
    """The native worker pool (avrf_server_*): tickets come back with the verdicts of the same batches
    verified one at a time, in any wait order, and queued batches survive `close`."""
    from ark_vrf_b200 import synth
    n = 50000
    b = synth.make_batch(0, n, 1, fmt=av.Format.MONTGOMERY)
    bad = b.s.copy()
    bad[n // 3, 2] ^= 4
    ident = b.pk.copy()
    ident[5, :32] = 0
    ident[5, 32:] = np.frombuffer(((1 << 256) % o.SUITES[0].p).to_bytes(32, "little"), dtype=np.uint8)
    variants = [(b.pk, b.s, 0), (b.pk, bad, 1), (ident, b.s, 2), (ident, bad, 2)]
    srv = av.BatchServer(0, av.Format.MONTGOMERY, workers=workers, hashers=hashers)   # hashers: shared multi-buffer SHA-512 threads
    tickets = []
    for k in range(4 * workers):
        pk, s_, want = variants[k % 4]
        if k % 3 == 2:                                   # ragged: a shorter batch, other verdict pattern
            h = n // 2 + k
            tickets.append((srv.submit(pk[:h], b.ios[:h], b.io_offsets[:h + 1], b.ad_blob, b.ad_offsets[:h + 1], b.r[:h], s_[:h]),
                            {0: 0, 1: 1, 2: 2}[want] if want != 1 else (1 if n // 3 < h else 0)))
        else:
            tickets.append((srv.submit(pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s_), want))
    empty = np.zeros(1, dtype=np.uint32)
    t_empty = srv.submit(b.pk[:0], b.ios[:0], empty, b.ad_blob[:0], empty, b.r[:0], b.s[:0])
    for t, want in reversed(tickets):
        assert srv.wait(t) == want
    assert srv.wait(t_empty) == 0                       # empty batch => Ok (thin.rs:262-264)
    with pytest.raises(av.AvrfError):
        srv.wait(10 ** 6)                               # unknown ticket
    t_last = srv.submit(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    with pytest.raises(av.VerificationFailure):
        srv.verify(srv.submit(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, bad))
    assert srv.wait(t_last) == 0
    srv.submit(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    srv.close()                                         # finishes the queued batch, joins the workers


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_out_of_range_inputs_are_argument_errors(av, sid):
    """A coordinate >= p or a response >= r is not a field element (the reference's types cannot hold one): the
    library reports a system error instead of mis-reducing it into some verdict.  p - 1 and r - 1 are accepted."""
    from ark_vrf_b200 import synth
    S = o.SUITES[sid]
    b = synth.make_batch(sid, 300, 1, fmt=av.Format.CANONICAL)

    def status(pk=b.pk, ios=b.ios, r=b.r, s_=b.s):
        bv = av.BatchVerifier(sid, av.Format.CANONICAL)
        bv.push_many(pk, ios, b.io_offsets, b.ad_blob, b.ad_offsets, r, s_)
        return bv.verify_status()
    assert status() == 0
    le = lambda x: np.frombuffer(x.to_bytes(32, "little"), dtype=np.uint8)
    pk2 = b.pk.copy(); pk2[7, :32] = le(S.p)
    io2 = b.ios.copy(); io2[299, 96:] = le(S.p + 5)
    r2 = b.r.copy(); r2[0, 32:] = le((1 << 256) - 1)
    s2 = b.s.copy(); s2[150] = le(S.r)
    for kw in (dict(pk=pk2), dict(ios=io2), dict(r=r2), dict(s_=s2)):
        with pytest.raises(av.AvrfError):
            status(**kw)
    pk3 = b.pk.copy(); pk3[7, :32] = le(S.p - 1)          # in range (not on the curve: unchecked, just a bad proof)
    s3 = b.s.copy(); s3[150] = le(S.r - 1)
    assert status(pk=pk3) == 1 and status(s_=s3) == 1


@pytest.mark.parametrize("sid,montgomery", [(0, False), (0, True), (2, False)])
def test_ragged_batch(av, sid, montgomery):
    """Ragged inputs in ONE push: M_j in {0..5} varies per proof (src/thin.rs:282 allows it), ad lengths from
    0 to 300 bytes (transcripts of 1 to 5 SHA-512 blocks).  All taps bit-exact against the oracles; the
    C oracle and the Python oracle agree on the same ragged batch."""
    import random
    from oracle import corc
    S = o.SUITES[sid]
    rnd = random.Random(11 + sid)
    sks = [o.secret_from_seed(S, o.synth_seed(k)) for k in range(3)]
    pr = o.Proofs(S)
    ms = [0, 1, 2, 5, 3, 1, 0, 4, 1, 2, 1, 1, 5, 0, 2, 3]
    ads = [0, 1, 7, 8, 9, 100, 127, 128, 129, 300, 19, 20, 21, 63, 64, 65]
    for j, (m, al) in enumerate(zip(ms, ads)):
        sk = sks[j % 3]
        ios = []
        for i in range(m):
            inp = o.data_to_point(S, o.synth_msg(1000 + j, i))
            ios.append((inp, o.pt_mul(S, inp, sk)))
        ad = bytes(rnd.randrange(256) for _ in range(al))
        R, s = o.thin_prove(S, sk, ios, ad)
        pr.pk.append(o.public_key(S, sk)); pr.ios.append(ios); pr.ad.append(ad); pr.r.append(R); pr.s.append(s)
    items = oracle_items(pr)
    bv = _push_all(av, sid, pr, montgomery)
    assert bv.verify_status() == 0 == o.batch_verify(S, items)
    assert [bytes(x) for x in bv.tap(av.Tap.C).reshape(-1, 16)] == [e.c.to_bytes(16, "little") for e in items]
    assert [bytes(x) for x in bv.tap(av.Tap.Z).reshape(-1, 16)] == [z.to_bytes(16, "little") for e in items for z in e.zs[1:]]
    _, scalars = o.batch_msm_terms(S, items)
    assert [bytes(x) for x in bv.tap(av.Tap.SCALARS).reshape(-1, 32)] == [sc_bytes(k) for k in scalars]
    st, _, taps = corc.thin_batch_verify(sid, *arrays_from_proofs(pr), nthreads=3, taps=True)
    assert st == 0 and taps["seed"].tobytes() == bytes(bv.tap(av.Tap.SEED))
    assert list(bv.verify_each()) == [0] * len(ms)
    # one bad proof among the ragged ones
    pr.s[3] = (pr.s[3] + 1) % S.r
    b2 = _push_all(av, sid, pr, montgomery)
    assert b2.verify_status() == 1
    assert list(b2.verify_each()) == [1 if j == 3 else 0 for j in range(len(ms))]

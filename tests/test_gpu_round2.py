"""GPU parity tests added in round 2: the reference's security-regression scenarios, the exact single-proof
verifier, staged single pushes, the in-library multi-GPU path, proving with many I/O pairs, handle-state
robustness.  Same bar as tests/test_gpu_parity.py: bit-exact bytes, identical verdicts."""
import ctypes
import hashlib

import numpy as np
import pytest

from oracle import pyref as o
from helpers import arrays_from_proofs, golden_proofs, oracle_items, pt_bytes, sc_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def av():
    import ark_vrf_b200 as a
    a.load().avrf_init(0)
    return a


def _proofs(S, pk, ios, ad, R, s):
    pr = o.Proofs(S)
    pr.pk.append(pk); pr.ios.append(list(ios)); pr.ad.append(ad); pr.r.append(R); pr.s.append(s)
    return pr


def _batch_status(av, sid, pr):
    bv = av.BatchVerifier(sid, av.Format.CANONICAL)
    bv.push_many(*arrays_from_proofs(pr))
    return bv.verify_status()


def _single_status(av, sid, pk, ios, ad, R, s):
    lib = av.load()
    st = ctypes.c_int32(-1)
    iob = b"".join(pt_bytes(i) + pt_bytes(oo) for (i, oo) in ios)
    rc = lib.avrf_thin_verify_one(sid, 1, pt_bytes(pk), iob if ios else None, len(ios), ad if ad else None, len(ad),
                                  pt_bytes(R), sc_bytes(s), ctypes.byref(st))
    assert rc == 0, lib.avrf_last_error()
    return st.value


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_known_dlog_input_forgery(av, sid):
    """reference src/thin.rs:656-733: with I = d*G known to the prover, a proof for an ARBITRARY output verifies -
    the verifier must not (and does not) add checks of its own: same accept as the reference, single and batch."""
    S = o.SUITES[sid]
    sk, d, tt, k = 42, 7, 1234, 9999
    pk = o.pt_mul(S, S.G, sk)
    inp = o.pt_mul(S, S.G, d)
    honest = o.pt_mul(S, inp, sk)
    fake = o.pt_mul(S, S.G, tt)
    assert fake != honest
    ad = b"attack"
    ios = [(inp, fake)]
    t, zs = o.thin_transcript(S, pk, ios, ad)
    z0, z1 = zs
    im = o.ext_to_affine(S, o.ext_add(S, o.ext_mul(S, o.to_ext(S.G), z0), o.ext_mul(S, o.to_ext(inp), z1)))
    x = (z0 * sk + z1 * tt) * pow(z0 + z1 * d, -1, S.r) % S.r
    R = o.pt_mul(S, im, k)
    c = o.challenge(S, [R], t)
    s = (k + c * x) % S.r
    assert o.thin_verify(S, pk, ios, ad, R, s) == 0                       # the reference accepts (thin.rs:728-732)
    assert _single_status(av, sid, pk, ios, ad, R, s) == 0
    pr = _proofs(S, pk, ios, ad, R, s)
    assert _batch_status(av, sid, pr) == o.batch_verify(S, oracle_items(pr)) == 0
    # the honest output with the same forged response does not verify
    assert _single_status(av, sid, pk, [(inp, honest)], ad, R, s) == o.thin_verify(S, pk, [(inp, honest)], ad, R, s) == 1


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_low_order_output_malleability(av, sid):
    """Thin-VRF analogue of reference src/lib.rs:770-870: O' = O + L with L of order 2, ad ground until z_1 is even, so
    that c*z_1*L vanishes in the EXACT single-proof equation.  The single verifier must accept exactly when the
    reference's does; the batch verifier reduces w*c*z_1 mod r first, and must agree with the reference's batch
    arithmetic on these non-subgroup points too.  R shifted by L is always rejected by the single verifier,
    whatever the parity of the batch weight (a batch of one could accept it)."""
    S = o.SUITES[sid]
    L = (0, S.p - 1)
    assert o.on_curve(S, L) and o.pt_add(S, L, L) == o.IDENTITY
    sk = o.secret_from_seed(S, bytes([9]) + bytes(31))
    pk = o.public_key(S, sk)
    inp = o.data_to_point(S, b"uniqueness attack")
    out = o.pt_mul(S, inp, sk)
    bad_out = o.pt_add(S, out, L)
    assert bad_out != out
    found = None
    for ctr in range(200):
        ad = b"ad-%d" % ctr
        ios = [(inp, bad_out)]
        t, zs = o.thin_transcript(S, pk, ios, ad)
        if zs[1] % 2 == 0:
            found = (ad, ios, t, zs)
            break
    assert found
    ad, ios, t, zs = found
    im, _ = o.merged_io(S, pk, ios, zs)
    k = 12345
    R = o.ext_to_affine(S, o.ext_mul(S, im, k))
    c = o.challenge(S, [R], t)
    s = (k + c * sk) % S.r
    want = o.thin_verify(S, pk, ios, ad, R, s)
    assert want == 0                                                      # c*z_1 is even: the forged output passes
    assert _single_status(av, sid, pk, ios, ad, R, s) == want
    pr = _proofs(S, pk, ios, ad, R, s)
    assert _batch_status(av, sid, pr) == o.batch_verify(S, oracle_items(pr))
    # odd z_1 and odd c: the reference rejects, so must we
    for ctr in range(200):
        ad2 = b"odd-%d" % ctr
        t2, zs2 = o.thin_transcript(S, pk, ios, ad2)
        im2, _ = o.merged_io(S, pk, ios, zs2)
        R2 = o.ext_to_affine(S, o.ext_mul(S, im2, k + ctr))
        c2 = o.challenge(S, [R2], t2)
        if zs2[1] % 2 == 1 and c2 % 2 == 1:
            s2 = (k + ctr + c2 * sk) % S.r
            assert o.thin_verify(S, pk, ios, ad2, R2, s2) == 1
            assert _single_status(av, sid, pk, ios, ad2, R2, s2) == 1
            break
    else:
        pytest.fail("no odd (z_1, c) found")
    # a valid proof whose R is shifted by L afterwards: the transcript changes, rejected
    good_ios = [(inp, out)]
    Rg, sg = o.thin_prove(S, sk, good_ios, b"x")
    assert _single_status(av, sid, pk, good_ios, b"x", Rg, sg) == 0
    RL = o.pt_add(S, Rg, L)
    assert _single_status(av, sid, pk, good_ios, b"x", RL, sg) == o.thin_verify(S, pk, good_ios, b"x", RL, sg) == 1
    # R' = R0 + L committed INSIDE the transcript (the attack of ADVICE r1): s answers c' = H(.., R'), so that
    # s*I_m - c'*O_m = R0 = R' - L.  The exact equation rejects it for every (c', s); a weighted batch of one
    # would accept whenever its weight is even.
    tg, zsg = o.thin_transcript(S, pk, good_ios, b"y")
    img, _ = o.merged_io(S, pk, good_ios, zsg)
    for kk in range(50, 60):
        R0 = o.ext_to_affine(S, o.ext_mul(S, img, kk))
        Rp = o.pt_add(S, R0, L)
        cp = o.challenge(S, [Rp], tg.clone())
        sp = (kk + cp * sk) % S.r
        assert o.thin_verify(S, pk, good_ios, b"y", Rp, sp) == 1
        assert _single_status(av, sid, pk, good_ios, b"y", Rp, sp) == 1


@pytest.mark.parametrize("sid,m", [(0, 0), (0, 1), (0, 5), (1, 2), (2, 9)])
def test_single_verify_many_pairs(av, sid, m):
    """thin::Verifier::verify (thin.rs:131-165) for 0 .. 9 pairs (more than one round of eight quads): accept, and
    reject on a tampered output / input / ad / response, against the oracle."""
    S = o.SUITES[sid]
    pr = o.synth_proofs(S, 2, m, signers=2)
    for j in range(2):
        args = (pr.pk[j], pr.ios[j], pr.ad[j], pr.r[j], pr.s[j])
        assert _single_status(av, sid, *args) == o.thin_verify(S, *args) == 0
        assert _single_status(av, sid, pr.pk[j], pr.ios[j], pr.ad[j] + b"!", pr.r[j], pr.s[j]) == 1
        assert _single_status(av, sid, pr.pk[j], pr.ios[j], pr.ad[j], pr.r[j], (pr.s[j] + 1) % S.r) == 1
        if m:
            ios = list(pr.ios[j])
            ios[m - 1] = (ios[m - 1][0], o.pt_add(S, ios[m - 1][1], S.G))
            assert _single_status(av, sid, pr.pk[j], ios, pr.ad[j], pr.r[j], pr.s[j]) == 1
            ios = list(pr.ios[j])
            ios[0] = (o.IDENTITY, ios[0][1])
            assert _single_status(av, sid, pr.pk[j], ios, pr.ad[j], pr.r[j], pr.s[j]) == 2
    assert _single_status(av, sid, o.IDENTITY, pr.ios[0], pr.ad[0], pr.r[0], pr.s[0]) == 2


def test_single_pushes_are_pipelined(av):
    """avrf_thin_batch_push x N (pinned staging, one chunk shipped at a time, hash on the handle's thread) gives the
    same seed, weights and verdict as one push_many; ragged M and ad; a tampered last proof rejects."""
    from ark_vrf_b200 import synth
    lib = av.load()
    n = 160000                                   # more than two 75776-proof chunks
    b = synth.make_batch(0, n, 1, signers=64, fmt=av.Format.MONTGOMERY)
    ref = av.BatchVerifier(0, av.Format.MONTGOMERY)
    ref.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    assert ref.verify_status() == 0
    seed = bytes(ref.tap(av.Tap.SEED))
    bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
    pk, ios, r, s, ad = (np.ascontiguousarray(x) for x in (b.pk, b.ios, b.r, b.s, b.ad_blob))
    push = lib.avrf_thin_batch_push
    for rounds in range(2):
        for j in range(n):
            a0, a1 = int(b.ad_offsets[j]), int(b.ad_offsets[j + 1])
            rc = push(bv._h, pk[j].ctypes.data, ios[j].ctypes.data, 1, ad[a0:].ctypes.data, a1 - a0, r[j].ctypes.data,
                      s[j].ctypes.data)
            assert rc == 0
        bv._n_ios = n
        assert len(bv) == n
        assert bv.verify_status() == 0
        assert bytes(bv.tap(av.Tap.SEED)) == seed
        assert bv.verify_status() == 0           # repeatable
        bv.clear()
    # tampered last proof, pushed alone after a bulk push
    bv.push_many(b.pk[:n - 1], b.ios[:n - 1], b.io_offsets[:n], b.ad_blob, b.ad_offsets[:n], b.r[:n - 1], b.s[:n - 1])
    s_bad = s[n - 1].copy()
    s_bad[0] ^= 1
    a0, a1 = int(b.ad_offsets[n - 1]), int(b.ad_offsets[n])
    assert push(bv._h, pk[n - 1].ctypes.data, ios[n - 1].ctypes.data, 1, ad[a0:].ctypes.data, a1 - a0, r[n - 1].ctypes.data,
                s_bad.ctypes.data) == 0
    assert bv.verify_status() == 1
    bv.clear()
    bv.reserve(1000, 4000, 100000)
    pr = o.synth_proofs(o.BANDERSNATCH, 6, 3, signers=2)
    bc = av.BatchVerifier(0, av.Format.CANONICAL)
    for j in range(6):
        bc.push(pt_bytes(pr.pk[j]), [(pt_bytes(a), pt_bytes(c)) for a, c in pr.ios[j]], pr.ad[j],
                av.Proof(pt_bytes(pr.r[j]), sc_bytes(pr.s[j])))
    assert bc.verify_status() == 0
    items = oracle_items(pr)
    assert bytes(bc.tap(av.Tap.SEED)) == o.batch_seed(o.BANDERSNATCH, items)


@pytest.mark.parametrize("sid,m,n", [(0, 1, 5000), (2, 3, 700), (1, 0, 300)])
def test_sharded_in_library(av, sid, m, n):
    """avrf_thin_sharded_*: one batch over every GPU of the process (1 on a single-GPU box, all of them under
    `gpurun --gpus N`): seed, verdicts and per-shard weights bit-identical to a single-device handle; rejects with
    the fault on the first / last proof of every shard; InvalidData precedence; several pushes (several runs of
    global indices per device)."""
    import torch
    from ark_vrf_b200 import synth
    S = o.SUITES[sid]
    nd = av.init_multi(torch.cuda.device_count())
    assert nd == torch.cuda.device_count()
    b = synth.make_batch(sid, n, m, signers=16, fmt=av.Format.CANONICAL)
    args = (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    one = av.BatchVerifier(sid, av.Format.CANONICAL)
    one.push_many(*args)
    assert one.verify_status() == 0
    seed = bytes(one.tap(av.Tap.SEED))
    w_all = one.tap(av.Tap.W).reshape(n, 16)
    sh = av.ShardedBatchVerifier(sid, av.Format.CANONICAL)
    assert sh.devices == nd
    sh.push_many(*args)
    assert len(sh) == n
    assert sh.verify_status() == 0
    assert sh.seed() == seed == hashlib.sha512(S.suite_id + b"\x50" + one.cs_stream().tobytes()).digest()
    per = ((n + nd - 1) // nd + 31) // 32 * 32
    for d in range(nd):
        lo, hi = min(n, d * per), min(n, d * per + per)
        if hi > lo:
            v = sh.shard(d, (hi - lo) * m)
            assert len(v) == hi - lo
            assert (v.tap(av.Tap.W).reshape(-1, 16) == w_all[lo:hi]).all()
    assert sh.verify_status() == 0                # repeatable
    t = sh.timings()
    assert t["total_ms"] > 0
    # faults at the first and last proof of every shard
    edges = sorted({x for d in range(nd) for x in (min(n - 1, d * per), min(n, d * per + per) - 1)})
    for pos in edges:
        s2 = b.s.copy()
        s2[pos, 0] ^= 1
        sh.clear()
        sh.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
        assert sh.verify_status() == 1, pos
    ident = np.frombuffer(pt_bytes(o.IDENTITY), dtype=np.uint8)
    pk2 = b.pk.copy()
    pk2[n - 1] = ident
    s2 = b.s.copy()
    s2[0, 0] ^= 1
    sh.clear()
    sh.push_many(pk2, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
    assert sh.verify_status() == 2                # InvalidData takes precedence (thin.rs:266-271)
    # the same batch in three pushes: global order = push order
    sh.clear()
    cuts = [0, n // 3, n // 3 + 17, n]
    for a, c in zip(cuts[:-1], cuts[1:]):
        io0, io1 = int(b.io_offsets[a]), int(b.io_offsets[c])
        ad0, ad1 = int(b.ad_offsets[a]), int(b.ad_offsets[c])
        sh.push_many(b.pk[a:c], b.ios[io0:io1 + 1], (b.io_offsets[a:c + 1] - io0).astype(np.uint32),
                     np.concatenate([b.ad_blob[ad0:ad1], np.zeros(16, np.uint8)]),
                     (b.ad_offsets[a:c + 1] - ad0).astype(np.uint32), b.r[a:c], b.s[a:c])
    assert sh.verify_status() == 0
    assert sh.seed() == seed
    sh.close()


def test_prove_more_than_eight_pairs(av):
    """avrf_thin_prove_many with 12 I/O pairs per proof (round 1 stopped at 8): R, s equal the oracle's prover."""
    from ark_vrf_b200 import ops
    S = o.BANDERSNATCH
    m = 12
    pr = o.synth_proofs(S, 3, m, signers=3)
    sks = [o.secret_from_seed(S, o.synth_seed(j % 3)) for j in range(3)]
    pk, ios, io_off, ad, ad_off, _, _ = arrays_from_proofs(pr)
    sk_arr = np.frombuffer(b"".join(sc_bytes(x) for x in sks), dtype=np.uint8).reshape(3, 32).copy()
    r, s = ops.thin_prove_many(0, sk_arr, pk, ios, io_off, ad, ad_off, fmt=av.Format.CANONICAL)
    for j in range(3):
        assert bytes(r[j]) == pt_bytes(pr.r[j]) and bytes(s[j]) == sc_bytes(pr.s[j])


def test_handle_state_robustness(av):
    """calls that arrive while a verify is in flight complete it first; several eager pushes followed by
    set_eager(0) + invalidate do not read stale chunk events; TREE weights are refused for Pedersen handles."""
    from ark_vrf_b200 import synth, pedersen
    b = synth.make_batch(0, 3000, 1, signers=8, fmt=av.Format.CANONICAL)
    args = (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    s2 = b.s.copy()
    s2[1234, 0] ^= 1
    bv = av.BatchVerifier(0, av.Format.CANONICAL)
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
    bv.verify_async()
    w = bv.tap(av.Tap.W)                          # must not clobber the verdict words of the verify in flight
    assert bv.verify_status() == 1
    bv.verify_async()
    assert (bv.verify_each() != 0).sum() == 1
    assert bv.verify_status() == 1
    assert (bv.tap(av.Tap.W) == w).all()
    # two eager pushes, then the non-eager seed path
    bv.clear()
    h = b.io_offsets
    bv.push_many(b.pk[:1500], b.ios[:1501], h[:1501], b.ad_blob, b.ad_offsets[:1501], b.r[:1500], b.s[:1500])
    a0 = int(b.ad_offsets[1500])
    bv.push_many(b.pk[1500:], b.ios[1500:], (h[1500:] - h[1500]).astype(np.uint32),
                 b.ad_blob[a0:], (b.ad_offsets[1500:] - a0).astype(np.uint32), b.r[1500:], b.s[1500:])
    assert bv.verify_status() == 0
    seed = bytes(bv.tap(av.Tap.SEED))
    bv.invalidate()
    assert bv.verify_status() == 0
    assert bytes(bv.tap(av.Tap.SEED)) == seed
    # blocking waits: same results
    bv.set_blocking(True)
    bv.invalidate()
    assert bv.verify_status() == 0 and bytes(bv.tap(av.Tap.SEED)) == seed
    bv.set_blocking(False)
    lib = av.load()
    pb = lib.avrf_pedersen_batch_new(0, 1)
    assert pb
    assert lib.avrf_thin_batch_set_weights_mode(pb, 1) < 0
    out = np.zeros(64, np.uint8)
    nl = ctypes.c_uint64(0)
    assert lib.avrf_thin_batch_tree_leaves(pb, 0, out.ctypes.data, ctypes.byref(nl)) < 0
    lib.avrf_thin_batch_free(pb)
    st = ctypes.c_int32(-9)
    assert lib.avrf_thin_verify_one(7, 1, bytes(64), None, 0, None, 0, bytes(64), bytes(32), ctypes.byref(st)) == -2


@pytest.mark.parametrize("sid,n", [(0, 5000), (1, 600), (2, 600)])
def test_vrf_io_many(av, sid, n):
    """avrf_vrf_io_many (Input::new + Secret::output + Output::hash fused, every inversion batched) against the
    separate entry points and the oracle; per-item and shared secrets, both formats."""
    from ark_vrf_b200 import ops, synth
    S = o.SUITES[sid]
    msgs = [b"msg-%d" % j + bytes(j % 7) for j in range(n)]
    sks = [synth.secret_from_seed(sid, bytes([j % 5]) + bytes(31)) for j in range(n)]
    sk_arr = np.frombuffer(b"".join(sc_bytes(k) for k in sks), dtype=np.uint8).reshape(n, 32).copy()
    res = ops.vrf_io_many(sid, msgs, None, sk_arr, want_hashes=True)
    assert res["ok"].all()
    pts, ok = ops.hash_to_curve(sid, msgs)
    assert ok.all() and (pts == res["inputs"]).all()
    outs = ops.vrf_output(sid, sk_arr, pts)
    assert (outs == res["outputs"]).all()
    assert (ops.point_to_hash(sid, outs) == res["hashes"]).all()
    for j in (0, 1, n // 2, n - 1):
        h = o.data_to_point(S, msgs[j])
        g = o.pt_mul(S, h, sks[j])
        assert bytes(res["inputs"][j]) == pt_bytes(h) and bytes(res["outputs"][j]) == pt_bytes(g)
        assert bytes(res["hashes"][j]) == o.point_to_hash(S, g)
    # shared secret, Montgomery format, outputs only
    sk0 = sk_arr[0].copy()
    skm = np.frombuffer(((sks[0] << 256) % S.r).to_bytes(32, "little"), dtype=np.uint8).copy()
    r2 = ops.vrf_io_many(sid, msgs[:100], None, skm, fmt=av.Format.MONTGOMERY, want_inputs=False)
    assert "inputs" not in r2
    can = ops.vrf_output(sid, sk0, pts[:100])
    for j in (0, 57, 99):
        x = int.from_bytes(bytes(r2["outputs"][j][:32]), "little") * pow(1 << 256, -1, S.p) % S.p
        assert x == int.from_bytes(bytes(can[j][:32]), "little")


def test_ingest_bulk_matches_scalar_mul_check(av):
    """Bulk ingest at size: 2^16 random Bandersnatch encodings (every coset of the prime-order subgroup, non-residues,
    both sign flags): the 2-descent subgroup test accepts exactly the points whose [r]P is the identity, decided here by
    an independent route - the accepted points times the cofactor-free order through avrf_vrf_output."""
    from ark_vrf_b200 import ops
    S = o.BANDERSNATCH
    rng = np.random.default_rng(5)
    n = 1 << 16
    enc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    enc[:, 31] &= 0xF3                            # y < 2^255 * ..., keeps most candidates below p; flag bit random
    good = [o.enc_point(S, o.pt_mul(S, S.G, int(k))) for k in rng.integers(1, 1 << 62, size=64)]
    enc[:64] = np.frombuffer(b"".join(good), dtype=np.uint8).reshape(64, 32)
    pts, ok = ops.points_deserialize(0, enc, kind=0)
    assert ok[:64].all()
    frac = ok[64:].mean()
    assert 0.08 < frac < 0.17                     # ~1/2 have a root, 1/4 of those are in the subgroup
    # accepted points: r * P must be the identity (scalar multiplication by the group order, canonical scalar r)
    acc = pts[ok.astype(bool)]
    rb = np.frombuffer(S.r.to_bytes(32, "little"), dtype=np.uint8).copy()
    back = ops.vrf_output(0, rb, acc)
    ident = np.frombuffer(pt_bytes(o.IDENTITY), dtype=np.uint8)
    assert (back == ident).all()
    # a sample of the rejected ones against the oracle's decision
    rej = np.nonzero(~ok.astype(bool))[0][:40]
    for j in rej:
        assert o.deserialize_point(S, bytes(enc[j]), reject_identity=False) is None


def test_hash_pool_handles(av):
    """avrf_hash_pool_*: handles whose batch seeds are hashed in the lanes of shared multi-buffer threads produce the
    same seed (SHA-512 of the stream definition, thin.rs:273-279), weights and verdicts as a handle hashing on its own
    thread - pushed eagerly, re-verified after invalidate, and after leaving the pool again."""
    import threading
    from ark_vrf_b200 import synth
    n = 9000
    b = synth.make_batch(0, n, 1, signers=16, fmt=av.Format.CANONICAL)
    args = (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    ref = av.BatchVerifier(0, av.Format.CANONICAL)
    ref.push_many(*args)
    assert ref.verify_status() == 0
    seed, w = bytes(ref.tap(av.Tap.SEED)), ref.tap(av.Tap.W).copy()
    pool = av.HashPool(2)
    hs = [av.BatchVerifier(0, av.Format.CANONICAL) for _ in range(11)]      # more handles than one thread has lanes
    s_bad = b.s.copy()
    s_bad[n - 1, 0] ^= 1
    res = {}

    def work(i, h):
        h.set_hash_pool(pool)
        h.set_blocking(True)
        h.push_many(*(args[:6] + ((s_bad if i == 3 else b.s),)))
        st = [h.verify_status()]
        sd = bytes(h.tap(av.Tap.SEED))
        h.invalidate()
        st.append(h.verify_status())
        res[i] = (st, sd, bytes(h.tap(av.Tap.SEED)), h.tap(av.Tap.W).copy())
    ths = [threading.Thread(target=work, args=(i, h)) for i, h in enumerate(hs)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(len(hs)):
        st, sd0, sd1, wi = res[i]
        if i == 3:
            assert st == [1, 1] and sd0 == sd1 != seed
        else:
            assert st == [0, 0] and sd0 == sd1 == seed and (wi == w).all()
    hs[0].set_hash_pool(None)
    hs[0].invalidate()
    assert hs[0].verify_status() == 0 and bytes(hs[0].tap(av.Tap.SEED)) == seed
    for h in hs:
        h.close()
    pool.close()


def _wire_arrays(S, pr):
    """Oracle Proofs -> the wire-format arrays of avrf_thin_batch_push_compressed."""
    n = len(pr.pk)
    pk32 = np.frombuffer(b"".join(o.enc_point(S, P) for P in pr.pk), dtype=np.uint8).reshape(n, 32).copy()
    r32 = np.frombuffer(b"".join(o.enc_point(S, P) for P in pr.r), dtype=np.uint8).reshape(n, 32).copy()
    s = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in pr.s), dtype=np.uint8).reshape(n, 32).copy()
    iob = b"".join(o.enc_point(S, i) + o.enc_point(S, oo) for ios in pr.ios for (i, oo) in ios)
    ios32 = np.frombuffer(iob + bytes(64), dtype=np.uint8).copy()
    io_off = np.zeros(n + 1, dtype=np.uint32)
    io_off[1:] = np.cumsum([len(x) for x in pr.ios])
    ad_off = np.zeros(n + 1, dtype=np.uint32)
    ad_off[1:] = np.cumsum([len(a) for a in pr.ad])
    ad = np.frombuffer(b"".join(pr.ad) + bytes(16), dtype=np.uint8).copy()
    return pk32, ios32, io_off, ad, ad_off, r32, s


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_push_compressed_golden(av, sid, golden):
    """avrf_thin_batch_push_compressed on the reference's golden vectors, taken as the hex strings they are published
    as (compressed pk, h, gamma, proof_r; canonical proof_s): accepted, and challenges / seed / weights bit-identical
    to the same proofs pushed as decoded points - with handles of either in-memory format."""
    S = o.SUITES[sid]
    vs = golden[sid]
    n = len(vs)
    pr = golden_proofs(S, vs)
    pk32 = np.frombuffer(b"".join(bytes.fromhex(v["pk"]) for v in vs), dtype=np.uint8).reshape(n, 32).copy()
    r32 = np.frombuffer(b"".join(bytes.fromhex(v["proof_r"]) for v in vs), dtype=np.uint8).reshape(n, 32).copy()
    s = np.frombuffer(b"".join(bytes.fromhex(v["proof_s"]) for v in vs), dtype=np.uint8).reshape(n, 32).copy()
    ios32 = np.frombuffer(b"".join(bytes.fromhex(v["h"]) + bytes.fromhex(v["gamma"]) for v in vs), dtype=np.uint8).copy()
    io_off = np.arange(n + 1, dtype=np.uint32)
    ad_off = np.zeros(n + 1, dtype=np.uint32)
    ad_off[1:] = np.cumsum([len(a) for a in pr.ad])
    ad = np.frombuffer(b"".join(pr.ad) + bytes(16), dtype=np.uint8).copy()
    ref = av.BatchVerifier(sid, av.Format.CANONICAL)
    ref.push_many(*arrays_from_proofs(pr))
    assert ref.verify_status() == 0
    for fmt in (av.Format.CANONICAL, av.Format.MONTGOMERY):
        bv = av.BatchVerifier(sid, fmt)
        ok = bv.push_compressed(pk32, ios32, io_off, ad, ad_off, r32, s)
        assert ok.all() and len(bv) == n
        assert bv.verify_status() == 0
        for tap in (av.Tap.C, av.Tap.SEED, av.Tap.W, av.Tap.R_COMPRESSED, av.Tap.Z):
            assert (bv.tap(tap) == ref.tap(tap)).all(), (sid, fmt, tap)
        assert [bytes(x).hex() for x in bv.tap(av.Tap.R_COMPRESSED).reshape(n, 32)] == [v["proof_r"] for v in vs]


@pytest.mark.parametrize("sid,m,n", [(0, 1, 200000), (2, 3, 500), (1, 0, 300), (0, 4, 700)])
def test_push_compressed_bulk_and_rejects(av, sid, m, n):
    """Wire-format push at size and with ragged I/O counts: same seed and verdict as the decoded push; a second
    compressed push appends; encodings the oracle's deserialiser refuses (y >= p, no root, off-subgroup, identity where a
    Public / Input / Output is expected) are named in ok[] and leave the handle untouched; an identity R decodes (it is
    a bare AffinePoint, thin.rs:42) and the batch then fails verification."""
    import random
    S = o.SUITES[sid]
    rnd = random.Random(90 + sid + m)
    pr = None
    if n <= 700:                                       # ragged: M_j cycles through 0..m, proofs made by the oracle
        sks = [o.secret_from_seed(S, o.synth_seed(k)) for k in range(3)]
        pr = o.Proofs(S)
        for j in range(24):
            sk = sks[j % 3]
            ios = []
            for i in range(j % (m + 1)):
                inp = o.data_to_point(S, o.synth_msg(500 + j, i))
                ios.append((inp, o.pt_mul(S, inp, sk)))
            ad = b"wire-%d" % j
            R, sc = o.thin_prove(S, sk, ios, ad)
            pr.pk.append(o.public_key(S, sk)); pr.ios.append(ios); pr.ad.append(ad); pr.r.append(R); pr.s.append(sc)
    if pr is None:
        from ark_vrf_b200 import synth, ops
        b = synth.make_batch(sid, n, m, signers=64, fmt=av.Format.CANONICAL)
        pk32 = ops.point_compress(sid, b.pk, fmt=av.Format.CANONICAL)
        r32 = ops.point_compress(sid, b.r, fmt=av.Format.CANONICAL)
        nio = int(b.io_offsets[n])
        ios32 = ops.point_compress(sid, b.ios[:128 * nio].reshape(-1, 64), fmt=av.Format.CANONICAL).reshape(-1)
        wire = (pk32, np.concatenate([ios32, np.zeros(64, np.uint8)]), b.io_offsets, b.ad_blob, b.ad_offsets, r32, b.s)
        dec = (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    else:
        n = len(pr.pk)
        wire = _wire_arrays(S, pr)
        dec = arrays_from_proofs(pr)
    ref = av.BatchVerifier(sid, av.Format.CANONICAL)
    ref.push_many(*dec)
    assert ref.verify_status() == 0
    bv = av.BatchVerifier(sid, av.Format.MONTGOMERY)
    ok = bv.push_compressed(*wire)
    assert ok.all() and len(bv) == n
    assert bv.verify_status() == 0
    assert bytes(bv.tap(av.Tap.SEED)) == bytes(ref.tap(av.Tap.SEED))
    assert (bv.tap(av.Tap.W) == ref.tap(av.Tap.W)).all()
    # appending: the same proofs again
    ok = bv.push_compressed(*wire)
    assert ok.all() and len(bv) == 2 * n and bv.verify_status() == 0
    # ---- encodings that do not deserialize -------------------------------------------------------------------------
    pk32, ios32, io_off, ad, ad_off, r32, s = [np.array(x, copy=True) for x in wire]
    bad_encs = [(S.p).to_bytes(32, "little"), o.enc_point(S, o.IDENTITY), (S.p - 1).to_bytes(32, "little")]
    while len(bad_encs) < 6:                            # random y: no root or off the prime-order subgroup
        e = bytearray(rnd.randrange(S.p).to_bytes(32, "little"))
        if o.deserialize_point(S, bytes(e), reject_identity=True) is None:
            bad_encs.append(bytes(e))
    victims = rnd.sample(range(n), 5)
    pk32[victims[0]] = np.frombuffer(bad_encs[0], np.uint8)
    pk32[victims[1]] = np.frombuffer(bad_encs[1], np.uint8)           # identity public key: refused (lib.rs:410-433)
    r32[victims[2]] = np.frombuffer(bad_encs[3], np.uint8)
    want_bad = {victims[0], victims[1], victims[2]}
    if m:
        owners = [j for j in range(n) if io_off[j + 1] > io_off[j]]
        j3, j4 = rnd.sample(owners, 2)
        ios32[64 * int(io_off[j3]):64 * int(io_off[j3]) + 32] = np.frombuffer(bad_encs[4], np.uint8)          # an input
        q = int(io_off[j4 + 1]) - 1
        ios32[64 * q + 32:64 * q + 64] = np.frombuffer(bad_encs[2], np.uint8)                                 # an output: (0, -1)
        want_bad |= {j3, j4}
    fresh = av.BatchVerifier(sid, av.Format.CANONICAL)
    ok = fresh.push_compressed(pk32, ios32, io_off, ad, ad_off, r32, s)
    assert set(np.nonzero(ok == 0)[0].tolist()) == want_bad
    assert len(fresh) == 0 and fresh.verify_status() == 0             # nothing was pushed: an empty batch verifies
    # the flags agree with the oracle's deserialiser point by point
    for j in sorted(want_bad | set(rnd.sample(range(n), min(n, 60)))):
        pts_ok = o.deserialize_point(S, bytes(pk32[j]), True) is not None and o.deserialize_point(S, bytes(r32[j]), False) is not None
        for i in range(2 * int(io_off[j]), 2 * int(io_off[j + 1])):
            pts_ok = pts_ok and o.deserialize_point(S, bytes(ios32[32 * i:32 * i + 32]), True) is not None
        assert bool(ok[j]) == pts_ok, j
    # the same refusal on a handle that already holds proofs: it keeps exactly those (rolled back), same seed as before
    seed1 = bytes(ref.tap(av.Tap.SEED))
    held = av.BatchVerifier(sid, av.Format.MONTGOMERY)
    assert held.push_compressed(*wire).all() and held.verify_status() == 0
    ok = held.push_compressed(pk32, ios32, io_off, ad, ad_off, r32, s)
    assert set(np.nonzero(ok == 0)[0].tolist()) == want_bad
    assert len(held) == n and held.verify_status() == 0 and bytes(held.tap(av.Tap.SEED)) == seed1
    assert held.push_compressed(*wire).all() and len(held) == 2 * n and held.verify_status() == 0
    # identity R: decodes, fails the equation
    pk32, ios32, io_off, ad, ad_off, r32, s = [np.array(x, copy=True) for x in wire]
    r32[n // 2] = np.frombuffer(o.enc_point(S, o.IDENTITY), np.uint8)
    ok = fresh.push_compressed(pk32, ios32, io_off, ad, ad_off, r32, s)
    assert ok.all() and len(fresh) == n and fresh.verify_status() == 1


def test_mixed_server_pool(av):
    """avrf_server_new_mixed: own-thread workers beside workers that hash in shared multi-buffer lanes; a burst of more
    batches than workers, with rejecting batches among them, returns every ticket's own verdict (lane workers take the
    early jobs, own-thread workers the rest)."""
    from ark_vrf_b200 import synth
    n = 5000
    b = synth.make_batch(0, n, 1, signers=16, fmt=av.Format.MONTGOMERY)
    bad = b.s.copy()
    bad[n // 2, 1] ^= 4
    pk_id = b.pk.copy()
    pk_id[7] = synth.identity_point(0, av.Format.MONTGOMERY)
    srv = av.BatchServer(0, av.Format.MONTGOMERY, workers=5, hashers=1, own_hash_workers=2)
    want, tickets = [], []
    for k in range(17):
        pk, s, w = [(b.pk, b.s, 0), (b.pk, bad, 1), (pk_id, bad, 2)][k % 3]
        tickets.append(srv.submit(pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s))
        want.append(w)
    assert [srv.wait(t) for t in tickets] == want
    srv.close()

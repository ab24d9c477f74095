"""CPU: pin the oracle (oracle/pyref.py) against the reference's own golden vectors
(tests/golden/*_thin.json = /root/reference/data/vectors/*_thin.json, replayed the way
reference src/testing.rs:263-280 and src/thin.rs:635-648 do) and SURVEY.md Appendix B."""
import pytest

from oracle import pyref as o
from helpers import GOLDEN_SEEDS, golden_proofs, oracle_items


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_vectors_process(sid, golden):
    S = o.SUITES[sid]
    for v, sd in zip(golden[sid], GOLDEN_SEEDS):
        sk = o.secret_from_seed(S, bytes([sd]) + bytes(31))
        assert o.enc_scalar(sk).hex() == v["sk"]
        pk = o.public_key(S, sk)
        assert o.enc_point(S, pk).hex() == v["pk"]
        alpha, ad = bytes.fromhex(v["alpha"]), bytes.fromhex(v["ad"])
        h = o.data_to_point(S, alpha)
        assert o.enc_point(S, h).hex() == v["h"]
        gamma = o.pt_mul(S, h, sk)
        assert o.enc_point(S, gamma).hex() == v["gamma"]
        assert o.point_to_hash(S, gamma).hex() == v["beta"]
        R, s = o.thin_prove(S, sk, [(h, gamma)], ad)
        assert o.enc_point(S, R).hex() == v["proof_r"]
        assert o.enc_scalar(s).hex() == v["proof_s"]
        assert o.dec_point(S, bytes.fromhex(v["proof_r"])) == R
        assert o.thin_verify(S, pk, [(h, gamma)], ad, R, s) == o.OK
        assert o.thin_verify(S, pk, [(h, gamma)], ad + b"x", R, s) == o.VERIFICATION_FAILURE


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_batch_of_vectors(sid, golden):
    S = o.SUITES[sid]
    items = oracle_items(golden_proofs(S, golden[sid]))
    assert o.batch_verify(S, items) == o.OK
    assert o.batch_verify(S, []) == o.OK
    items[3].s = (items[3].s + 1) % S.r
    assert o.batch_verify(S, items) == o.VERIFICATION_FAILURE
    items[3].s = (items[3].s - 1) % S.r
    items[5].pk = o.IDENTITY
    assert o.batch_verify(S, items) == o.INVALID_DATA


def test_appendix_b_kats(golden):
    """SURVEY.md Appendix B (restatement-derived, consistent with the pinned fields)."""
    S = o.BANDERSNATCH
    items = oracle_items(golden_proofs(S, golden[0]))
    assert items[0].zs[1].to_bytes(16, "little").hex() == "d75b5745dc52658ab87f48bd0bb6f15a"
    assert items[0].c.to_bytes(16, "little").hex() == "08be086526bd2ca18d27746c16fc8a55"
    assert items[5].zs[1].to_bytes(16, "little").hex() == "e8900e64af6ae59ff69665a1b51ae7cf"
    assert items[5].c.to_bytes(16, "little").hex() == "3dc11ee9fdc3cc919efb99bbae73019d"
    seed = o.batch_seed(S, items)
    assert seed.hex().startswith("be45c2174b58d246acd14ed14fc857e5")
    ws = o.batch_weights(S, seed, 7)
    assert ws[0].to_bytes(16, "little").hex() == "b2a5366d770f3656a54dd068c18ccc9d"
    assert ws[6].to_bytes(16, "little").hex() == "36cacd1560836d5788ed6adbf6956cbe"
    u0, u1 = o.ell2_hash_to_field(S, b"")
    assert u0.to_bytes(32, "little").hex() == "33d461b0f87c1b40fbbf73620df62fe7ac6b9d0057178f4a29a9fe26ba542533"
    assert o.enc_point(S, o.ell2_map(S, u0)).hex() == "76e8ccb4624ad0bb1443b18fa569dbfba28e6f6ecc63ba37f3f85d23a23972eb"
    assert o.enc_point(S, o.ell2_map(S, u1)).hex() == "85f22a54fa8460d88dc17aa028c0f8b40f31ab523a5217c1b5dfdcaa963f2044"


@pytest.mark.parametrize("sid", [0, 1, 2])
def test_reference_behaviours(sid):
    """Accept/reject scenarios of reference src/thin.rs:346-516 re-enacted on the oracle."""
    S = o.SUITES[sid]
    sk = o.secret_from_seed(S, bytes(32))
    pk = o.public_key(S, sk)
    ios = []
    for i in range(3):
        inp = o.data_to_point(S, bytes([i + 1]))
        ios.append((inp, o.pt_mul(S, inp, sk)))
    R, s = o.thin_prove(S, sk, ios, b"bar")
    assert o.thin_verify(S, pk, ios, b"bar", R, s) == o.OK
    bad = list(ios); bad[1] = (ios[1][0], ios[0][1])
    assert o.thin_verify(S, pk, bad, b"bar", R, s) == o.VERIFICATION_FAILURE
    bad = list(ios); bad[0] = (ios[1][0], ios[0][1])
    assert o.thin_verify(S, pk, bad, b"bar", R, s) == o.VERIFICATION_FAILURE
    assert o.thin_verify(S, pk, ios, b"baz", R, s) == o.VERIFICATION_FAILURE
    assert o.batch_verify(S, [o.batch_prepare(S, pk, ios, b"bar", R, s)]) == o.OK
    # zero pairs = Schnorr signature over ad (thin.rs:504-516)
    R0, s0 = o.thin_prove(S, sk, [], b"bar")
    assert o.thin_verify(S, pk, [], b"bar", R0, s0) == o.OK
    assert o.thin_verify(S, pk, [], b"baz", R0, s0) == o.VERIFICATION_FAILURE
    # identity public key forgery (thin.rs:418-433)
    sf = 0x5EED
    Rf = o.pt_mul(S, S.G, sf)
    assert o.thin_verify(S, o.IDENTITY, [], b"forgery", Rf, sf) == o.INVALID_DATA
    assert o.batch_verify(S, [o.batch_prepare(S, o.IDENTITY, [], b"forgery", Rf, sf)]) == o.INVALID_DATA
    # identity pair hidden behind a good one (thin.rs:443-471)
    idio = (o.IDENTITY, o.IDENTITY)
    Ri, si = o.thin_prove(S, sk, [ios[0], idio], b"forgery")
    assert o.thin_verify(S, pk, [ios[0], idio], b"forgery", Ri, si) == o.INVALID_DATA
    assert o.batch_verify(S, [o.batch_prepare(S, pk, [ios[0], idio], b"forgery", Ri, si)]) == o.INVALID_DATA

"""CPU oracle (Python big-int restatement) for ark-vrf's Thin VRF hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ark_vrf_b200/, bench.py's GPU
arm) may import this module; it exists so that tests/ and __graft_entry__.smoke() can
check the CUDA engine against an independent restatement of the reference algorithm.

Parity status: PINNED.  tests/test_oracle_golden.py replays the reference's own golden
vectors (tests/golden/*_thin.json, copied verbatim from /root/reference/data/vectors/)
through this file: sk, pk, h, gamma, beta, proof_r, proof_s all reproduce for the three
in-scope suites.  The reference itself (Rust + un-vendored arkworks crates) cannot be
built in this image, so behaviour of the third-party arithmetic (ark-ec 0.6, ark-ff 0.6,
ark-serialize 0.6, sha2 0.10) is restated from the published algorithms and anchored on
those vectors.

Every function cites the reference file:line it follows (paths relative to
/root/reference/).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# Suites (src/suites/{bandersnatch,ed25519,baby_jubjub}.rs; curve constants from the
# ark-ed-on-bls12-381-bandersnatch / ark-ed25519 / ark-ed-on-bn254 0.6 crates)
# --------------------------------------------------------------------------------------

Point = Tuple[int, int]  # affine twisted-Edwards (x, y); identity is (0, 1)


@dataclass(frozen=True)
class Suite:
    name: str
    suite_id: bytes
    p: int          # base field modulus
    a: int          # TE coefficient a (mod p)
    d: int          # TE coefficient d (mod p)
    r: int          # prime subgroup order
    cofactor: int
    G: Point
    h2c: str        # "ell2" | "tai"
    B: Point = (0, 1)   # Pedersen blinding base (suites/*.rs `BLINDING_BASE`)
    # Elligator2 (Bandersnatch only): Montgomery J (=A), K (=B), non-square Z
    mont_j: int = 0
    mont_k: int = 0
    ell2_z: int = 0

    @property
    def p_bits(self) -> int:
        return self.p.bit_length()

    @property
    def r_bits(self) -> int:
        return self.r.bit_length()


_P_BLS = 52435875175126190479447740508185965837690552500527637822603658699938581184513
_P_25519 = 2**255 - 19
_P_BN = 21888242871839275222246405745257275088548364400416034343698204186575808495617

BANDERSNATCH = Suite(
    name="bandersnatch_sha-512_ell2",
    suite_id=b"Bandersnatch-SHA512-ELL2-v1",            # suites/bandersnatch.rs:63
    p=_P_BLS,
    a=_P_BLS - 5,
    d=45022363124591815672509500913686876175488063829319466900776701791074614335719,
    r=13108968793781547619861935127046491459309155893440570251786403306729687672801,
    cofactor=4,
    G=(18886178867200960497001835917649091219057080094937609519140440539760939937304,
       19188667384257783945677642223292697773471335439753913231509108946878080696678),
    h2c="ell2",
    B=(23335687741101763108036518445642207119627658113885888016488710494487028845889,       # suites/bandersnatch.rs:72-80
       5552214580375038693022409684979828600325210968745774080859660443337357929963),
    mont_j=29978822694968839326280996386011761570173833766074948509196803838190355340952,
    mont_k=25465760566081946422412445027709227188579564747101592991722834452325077642517,
    ell2_z=5,
)

ED25519 = Suite(
    name="ed25519_sha-512_tai",
    suite_id=b"Ed25519-SHA512-TAI-v1",                   # suites/ed25519.rs:51
    p=_P_25519,
    a=_P_25519 - 1,
    d=(-121665 * pow(121666, -1, _P_25519)) % _P_25519,
    r=2**252 + 27742317777372353535851937790883648493,
    cofactor=8,
    G=(15112221349535400772501151409588531511454012693041857206046113283949847762202,
       46316835694926478169428394003475163141307993866256225615783033603165251855960),
    h2c="tai",
    B=(45003173884697328536089278691112838614164406922820087464913813433380838325453,       # suites/ed25519.rs:56-65
       31256014272390301975555524011230972931324093235775711248505761870355310252869),
)

BABYJUBJUB = Suite(
    name="baby-jubjub_sha-512_tai",
    suite_id=b"BabyJubJub-SHA512-TAI-v1",                # suites/baby_jubjub.rs:57
    p=_P_BN,
    a=1,
    d=9706598848417545097372247223557719406784115219466060233080913168975159366771,
    r=2736030358979909402780800718157159386076813972158567259200215660948447373041,
    cofactor=8,
    G=(19698561148652590122159747500897617769866003486955115824547446575314762165298,
       19298250018296453272277890825869354524455968081175474282777126169995084727839),
    h2c="tai",
    B=(15549380791300914366206471199568039679131690710803662429646809536753521087193,       # suites/baby_jubjub.rs:62-71
       15218614024055502695611547593111691164731001864276292210438920202280814188379),
)

SUITES = {0: BANDERSNATCH, 1: ED25519, 2: BABYJUBJUB}
SUITE_BY_NAME = {s.name: s for s in SUITES.values()}

# Domain separators: src/utils/common.rs:128-152
DOM_THIN = 0x01
DOM_PEDERSEN = 0x02
DOM_PEDERSEN_BLINDING = 0x12
DOM_NONCE_EXPAND = 0x10
DOM_NONCE = 0x11
DOM_POINT_TO_HASH = 0x20
DOM_DELINEARIZE = 0x30
DOM_CHALLENGE = 0x40
DOM_BATCH = 0x50
DOM_H2C = 0x60

IDENTITY: Point = (0, 1)

# --------------------------------------------------------------------------------------
# Field helpers
# --------------------------------------------------------------------------------------


def fsqrt(n: int, p: int) -> Optional[int]:
    """Square root mod p (Tonelli-Shanks); None if n is a non-residue."""
    n %= p
    if n == 0:
        return 0
    if pow(n, (p - 1) // 2, p) != 1:
        return None
    if p % 4 == 3:
        return pow(n, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, rr = s, pow(z, q, p), pow(n, q, p), pow(n, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, rr = t * c % p, rr * b % p
    return rr


# --------------------------------------------------------------------------------------
# Twisted Edwards group law in extended coordinates (ark-ec 0.6 `twisted_edwards`)
# --------------------------------------------------------------------------------------

Ext = Tuple[int, int, int, int]  # (X, Y, Z, T) with x=X/Z, y=Y/Z, T=XY/Z


def to_ext(P: Point) -> Ext:
    return (P[0], P[1], 1, P[0] * P[1])


def ext_add(S: Suite, P: Ext, Q: Ext) -> Ext:
    p = S.p
    X1, Y1, Z1, T1 = P
    X2, Y2, Z2, T2 = Q
    A = X1 * X2 % p
    B = Y1 * Y2 % p
    C = S.d * T1 % p * T2 % p
    D = Z1 * Z2 % p
    E = ((X1 + Y1) * (X2 + Y2) - A - B) % p
    F = (D - C) % p
    G = (D + C) % p
    H = (B - S.a * A) % p
    return (E * F % p, G * H % p, F * G % p, E * H % p)


def ext_double(S: Suite, P: Ext) -> Ext:
    return ext_add(S, P, P)


def ext_neg(S: Suite, P: Ext) -> Ext:
    return ((-P[0]) % S.p, P[1], P[2], (-P[3]) % S.p)


def ext_is_identity(S: Suite, P: Ext) -> bool:
    return P[0] % S.p == 0 and (P[1] - P[2]) % S.p == 0


def ext_to_affine(S: Suite, P: Ext) -> Point:
    zi = pow(P[2], -1, S.p)
    return (P[0] * zi % S.p, P[1] * zi % S.p)


EXT_ID: Ext = (0, 1, 1, 0)


def ext_mul(S: Suite, P: Ext, k: int) -> Ext:
    acc = EXT_ID
    for bit in bin(k)[2:] if k else "":
        acc = ext_double(S, acc)
        if bit == "1":
            acc = ext_add(S, acc, P)
    return acc


def pt_add(S: Suite, P: Point, Q: Point) -> Point:
    return ext_to_affine(S, ext_add(S, to_ext(P), to_ext(Q)))


def pt_neg(S: Suite, P: Point) -> Point:
    return ((-P[0]) % S.p, P[1])


def pt_mul(S: Suite, P: Point, k: int) -> Point:
    """Scalar multiplication (`smul!`, src/utils/mod.rs:57-61)."""
    return ext_to_affine(S, ext_mul(S, to_ext(P), k % S.r if k >= S.r else k))


def pt_mul_raw(S: Suite, P: Point, k: int) -> Point:
    """Scalar multiplication without reducing k mod r (cofactor clearing etc.)."""
    return ext_to_affine(S, ext_mul(S, to_ext(P), k))


def on_curve(S: Suite, P: Point) -> bool:
    x, y = P
    p = S.p
    return (S.a * x * x + y * y - 1 - S.d * x * x % p * y * y) % p == 0


# --------------------------------------------------------------------------------------
# Encodings (ark-serialize 0.6 compressed; call sites src/utils/transcript.rs:48-50)
# --------------------------------------------------------------------------------------


def enc_scalar(k: int) -> bytes:
    return int(k).to_bytes(32, "little")


def enc_point(S: Suite, P: Point) -> bytes:
    """32-byte LE y; bit 7 of byte 31 set iff x > p - x (ark-ec TEFlags)."""
    x, y = P
    b = bytearray(y.to_bytes(32, "little"))
    if x > S.p - x and x != 0:
        b[31] |= 0x80
    return bytes(b)


def x_from_y(S: Suite, y: int, greatest: bool) -> Optional[Point]:
    """ark-ec `Affine::get_point_from_y_unchecked`."""
    p = S.p
    num = (1 - y * y) % p
    den = (S.a - S.d * y * y) % p
    if den == 0:
        return None
    x2 = num * pow(den, -1, p) % p
    x = fsqrt(x2, p)
    if x is None:
        return None
    neg = (-x) % p
    lo, hi = (x, neg) if x <= neg else (neg, x)
    return (hi if greatest else lo, y)


def dec_point(S: Suite, b: bytes) -> Optional[Point]:
    """Inverse of enc_point (no subgroup check)."""
    assert len(b) == 32
    flag = bool(b[31] & 0x80)
    yb = bytearray(b)
    yb[31] &= 0x7F
    y = int.from_bytes(yb, "little")
    if y >= S.p:
        return None
    return x_from_y(S, y, flag)


def deserialize_point(S: Suite, b: bytes, reject_identity: bool) -> Optional[Point]:
    """CanonicalDeserialize (compressed, Validate::Yes): Public / Input / Output (src/lib.rs:410-433,
    471-494,552-575; identity rejected) or a bare AffinePoint such as Proof.r (src/thin.rs:42)."""
    P = dec_point(S, b)
    if P is None or (reject_identity and P == IDENTITY):
        return None
    if not ext_is_identity(S, ext_mul(S, to_ext(P), S.r)):   # prime-subgroup check
        return None
    return P


# --------------------------------------------------------------------------------------
# Transcript (src/utils/transcript.rs:176-274): SHA-512 absorb, counter-mode squeeze
# --------------------------------------------------------------------------------------


class Transcript:
    def __init__(self, label: bytes = b"", _data: Optional[bytearray] = None):
        self.data = bytearray(label) if _data is None else _data
        self.seed: Optional[bytes] = None
        self.pos = 0

    def clone(self) -> "Transcript":
        assert self.seed is None
        return Transcript(_data=bytearray(self.data))

    def absorb(self, b: bytes) -> None:                      # transcript.rs:184-189
        if self.seed is not None:
            raise RuntimeError("cannot absorb after squeeze")
        self.data += b

    def squeeze(self, n: int) -> bytes:                      # transcript.rs:191-194,230-273
        if self.seed is None:
            self.seed = hashlib.sha512(bytes(self.data)).digest()
        out = bytearray()
        while len(out) < n:
            blk, off = divmod(self.pos, 64)
            block = hashlib.sha512(self.seed + blk.to_bytes(8, "little")).digest()
            take = min(64 - off, n - len(out))
            out += block[off:off + take]
            self.pos += take
        return bytes(out)


def challenge_scalar(S: Suite, t: Transcript) -> int:       # common.rs:72-76
    return int.from_bytes(t.squeeze(16), "little") % S.r


def expanded_scalar_len(S: Suite, sec_bits: int = 128) -> int:   # common.rs:57-64
    return (S.r_bits + sec_bits + 7) // 8


def nonce(S: Suite, sk: int, t: Transcript) -> int:         # common.rs:313-328
    t_exp = t.clone()
    t_exp.absorb(bytes([DOM_NONCE_EXPAND]))
    t_exp.absorb(enc_scalar(sk))
    sk_hash = t_exp.squeeze(64)
    t.absorb(bytes([DOM_NONCE]))
    t.absorb(sk_hash)
    return int.from_bytes(t.squeeze(expanded_scalar_len(S)), "little") % S.r


def challenge(S: Suite, pts: Sequence[Point], t: Transcript) -> int:   # common.rs:270-280
    t.absorb(bytes([DOM_CHALLENGE]))
    for P in pts:
        t.absorb(enc_point(S, P))
    return challenge_scalar(S, t)


def point_to_hash(S: Suite, P: Point, n: int = 32) -> bytes:           # common.rs:290-305
    t = Transcript(S.suite_id)
    t.absorb(bytes([DOM_POINT_TO_HASH]))
    t.absorb(enc_point(S, P))
    return t.squeeze(n)


def secret_from_seed(S: Suite, seed: bytes) -> int:                    # lib.rs:346-369
    assert len(seed) == 32
    sk = int.from_bytes(seed, "little") % S.r
    cnt = 0
    while True:
        t = Transcript(S.suite_id)
        t.absorb(seed)
        if cnt > 0:
            t.absorb(bytes([cnt]))
        k = nonce(S, sk, t)
        if k != 0:
            return k
        cnt += 1


def public_key(S: Suite, sk: int) -> Point:                            # lib.rs:331-334
    return pt_mul(S, S.G, sk)


# --------------------------------------------------------------------------------------
# Hash-to-curve (src/utils/hash_to_curve.rs)
# --------------------------------------------------------------------------------------


def expand_message_xmd_arkworks(msg: bytes, dst: bytes, n: int, zpad: int) -> bytes:
    """expand_message_xmd with SHA-512 as done by ark-ff 0.6 DefaultFieldHasher:
    Z_pad is `len_per_base_elem` bytes (48), not the 128-byte block size."""
    dst_prime = dst + bytes([len(dst)])
    ell = (n + 63) // 64
    b0 = hashlib.sha512(bytes(zpad) + msg + n.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    bi = hashlib.sha512(b0 + b"\x01" + dst_prime).digest()
    out = bytearray(bi)
    for i in range(2, ell + 1):
        bi = hashlib.sha512(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return bytes(out[:n])


def ell2_hash_to_field(S: Suite, msg: bytes) -> Tuple[int, int]:
    L = (S.p_bits + 128 + 7) // 8
    dst = S.suite_id + bytes([DOM_H2C])                     # hash_to_curve.rs:76
    u = expand_message_xmd_arkworks(msg, dst, 2 * L, L)
    return (int.from_bytes(u[:L], "big") % S.p, int.from_bytes(u[L:2 * L], "big") % S.p)


def ell2_map(S: Suite, u: int) -> Point:
    """ark-ec 0.6 Elligator2Map::map_to_curve for a TE curve (via its Montgomery model)."""
    p = S.p
    J, K, Z = S.mont_j, S.mont_k, S.ell2_z
    kinv = pow(K, -1, p)
    jk = J * kinv % p                      # J/K
    k2inv = kinv * kinv % p                # 1/K^2
    den = (1 + Z * u * u) % p
    if den == 0:
        den = 1
    x1 = (-jk) * pow(den, -1, p) % p

    def g(x: int) -> int:
        return (x * x % p * x + jk * x % p * x + x * k2inv) % p

    gx1 = g(x1)
    if gx1 != 0 and pow(gx1, (p - 1) // 2, p) == 1:
        x, y, sgn = x1, fsqrt(gx1, p), 1
    else:
        x2 = (-x1 - jk) % p
        x, y, sgn = x2, fsqrt(g(x2), p), 0
    assert y is not None
    if (y & 1) != sgn:
        y = (-y) % p
    s = x * K % p
    t = y * K % p
    tv1 = (s + 1) % p
    tv2 = tv1 * t % p
    if tv2 == 0:
        return IDENTITY
    inv = pow(tv2, -1, p)
    return (tv1 * s % p * inv % p, t * (s - 1) % p * inv % p)


def hash_to_curve_ell2(S: Suite, msg: bytes) -> Point:                 # hash_to_curve.rs:66-100
    u0, u1 = ell2_hash_to_field(S, msg)
    q = ext_add(S, to_ext(ell2_map(S, u0)), to_ext(ell2_map(S, u1)))
    return ext_to_affine(S, ext_mul(S, q, S.cofactor))


def hash_to_curve_tai(S: Suite, data: bytes) -> Optional[Point]:       # hash_to_curve.rs:34-57
    prefix = Transcript(S.suite_id)
    prefix.absorb(bytes([DOM_H2C]))
    prefix.absorb(len(data).to_bytes(8, "little"))
    prefix.absorb(data)
    for ctr in range(256):
        t = prefix.clone()
        t.absorb(bytes([ctr]))
        h = bytearray(t.squeeze(32))
        # ark-ec Affine::from_random_bytes: take flag from top bit, mask bits above p
        flag = bool(h[31] & 0x80)
        h[31] &= (0xFF >> (256 - S.p_bits))
        y = int.from_bytes(h, "little")
        if y >= S.p:
            continue
        P = x_from_y(S, y, flag)
        if P is None:
            continue
        P = pt_mul_raw(S, P, S.cofactor)
        if P != IDENTITY:
            return P
    return None


def data_to_point(S: Suite, data: bytes) -> Optional[Point]:           # Suite::data_to_point
    return hash_to_curve_ell2(S, data) if S.h2c == "ell2" else hash_to_curve_tai(S, data)


# --------------------------------------------------------------------------------------
# Thin VRF (src/thin.rs, src/utils/common.rs)
# --------------------------------------------------------------------------------------

VrfIo = Tuple[Point, Point]  # (input, output)


def thin_transcript(S: Suite, pk: Point, ios: Sequence[VrfIo], ad: bytes) -> Tuple[Transcript, List[int]]:
    """vrf_transcript_scalars_with_schnorr(ThinVrf, ..) (common.rs:159-173,231-258):
    returns the main transcript (ad absorbed) and zs = [1, z_1 .. z_M]."""
    t = Transcript(S.suite_id)
    t.absorb(bytes([DOM_THIN]))
    chain = [(S.G, pk)] + list(ios)                         # chain_ios, common.rs:231-240
    t.absorb(len(chain).to_bytes(8, "little"))              # absorb_ios, common.rs:377-383
    for (i, o) in chain:
        t.absorb(enc_point(S, i) + enc_point(S, o))
    t.absorb(len(ad).to_bytes(8, "little"))
    t.absorb(ad)
    zt = t.clone()                                          # DelinearizeScalars::new, :346-352
    zt.absorb(bytes([DOM_DELINEARIZE]))
    zs = [1] + [challenge_scalar(S, zt) for _ in range(len(chain) - 1)]
    return t, zs


def merged_io(S: Suite, pk: Point, ios: Sequence[VrfIo], zs: Sequence[int]) -> Tuple[Ext, Ext]:
    """merge_ios (common.rs:389-419) as group elements (n==1 shortcut gives the same)."""
    chain = [(S.G, pk)] + list(ios)
    im, om = EXT_ID, EXT_ID
    for (i, o), z in zip(chain, zs):
        im = ext_add(S, im, ext_mul(S, to_ext(i), z))
        om = ext_add(S, om, ext_mul(S, to_ext(o), z))
    return im, om


def thin_prove(S: Suite, sk: int, ios: Sequence[VrfIo], ad: bytes) -> Tuple[Point, int]:
    """thin::Prover::prove (thin.rs:111-129)."""
    pk = public_key(S, sk)
    t, zs = thin_transcript(S, pk, ios, ad)
    im, _ = merged_io(S, pk, ios, zs)
    k = nonce(S, sk, t.clone())
    R = ext_to_affine(S, ext_mul(S, im, k))
    c = challenge(S, [R], t)
    s = (k + c * sk) % S.r
    return R, s


def has_identity(ios: Sequence[VrfIo]) -> bool:             # lib.rs:632-634
    return any(i == IDENTITY or o == IDENTITY for (i, o) in ios)


OK, VERIFICATION_FAILURE, INVALID_DATA = 0, 1, 2            # reachable subset of lib.rs:136-147


def thin_verify(S: Suite, pk: Point, ios: Sequence[VrfIo], ad: bytes, R: Point, s: int) -> int:
    """thin::Verifier::verify (thin.rs:131-165)."""
    if pk == IDENTITY or has_identity(ios):
        return INVALID_DATA
    t, zs = thin_transcript(S, pk, ios, ad)
    im, om = merged_io(S, pk, ios, zs)
    c = challenge(S, [R], t)
    lhs = ext_add(S, ext_mul(S, im, s % S.r), ext_neg(S, ext_mul(S, om, c)))
    diff = ext_add(S, lhs, ext_neg(S, to_ext(R)))
    return OK if ext_is_identity(S, diff) else VERIFICATION_FAILURE


@dataclass
class BatchItem:                                            # thin.rs:172-179
    c: int
    pk: Point
    ios: List[VrfIo]
    zs: List[int]
    r: Point
    s: int


def batch_prepare(S: Suite, pk: Point, ios: Sequence[VrfIo], ad: bytes, R: Point, s: int) -> BatchItem:
    """BatchVerifier::prepare (thin.rs:209-226)."""
    t, zs = thin_transcript(S, pk, ios, ad)
    c = challenge(S, [R], t)
    return BatchItem(c=c, pk=pk, ios=list(ios), zs=zs, r=R, s=s)


def batch_seed(S: Suite, items: Sequence[BatchItem]) -> bytes:
    """SHA-512 of the batch transcript (thin.rs:273-279)."""
    h = hashlib.sha512()
    h.update(S.suite_id + bytes([DOM_BATCH]))
    for e in items:
        h.update(enc_scalar(e.c) + enc_scalar(e.s))
    return h.digest()


def batch_seed_tree(S: Suite, items: Sequence[BatchItem]) -> bytes:
    """Seed of the opt-in AVRF_WEIGHTS_TREE mode (NOT the reference's transcript; include/avrf.h):
    leaf_i = SHA512(0x00 || LE64(i) || stream of proofs 32i..32i+31),
    seed = SHA512(SUITE_ID || 0x50 || 0x01 || LE64(n) || leaves)."""
    stream = b"".join(enc_scalar(e.c) + enc_scalar(e.s) for e in items)
    n = len(items)
    leaves = b"".join(hashlib.sha512(b"\x00" + i.to_bytes(8, "little") + stream[2048 * i:2048 * (i + 1)]).digest()
                      for i in range((n + 31) // 32))
    return hashlib.sha512(S.suite_id + bytes([DOM_BATCH, 1]) + n.to_bytes(8, "little") + leaves).digest()


def batch_weights(S: Suite, seed: bytes, n: int) -> List[int]:
    """w_j = challenge_scalar of the batch stream (thin.rs:289, transcript.rs:255-273)."""
    out = []
    for j in range(n):
        blk = hashlib.sha512(seed + (j // 4).to_bytes(8, "little")).digest()
        out.append(int.from_bytes(blk[16 * (j % 4):16 * (j % 4) + 16], "little") % S.r)
    return out


def batch_msm_terms(S: Suite, items: Sequence[BatchItem]) -> Tuple[List[Point], List[int]]:
    """bases/scalars exactly as built at thin.rs:282-317."""
    seed = batch_seed(S, items)
    ws = batch_weights(S, seed, len(items))
    bases: List[Point] = []
    scalars: List[int] = []
    g = 0
    r = S.r
    for e, w in zip(items, ws):
        wc = w * e.c % r
        wsx = w * e.s % r
        bases.append(e.r); scalars.append(w)
        bases.append(e.pk); scalars.append(wc * e.zs[0] % r)
        g = (g - wsx * e.zs[0]) % r
        for i, (inp, out) in enumerate(e.ios):
            bases.append(out); scalars.append(wc * e.zs[i + 1] % r)
            bases.append(inp); scalars.append((-(wsx * e.zs[i + 1])) % r)
    bases.append(S.G); scalars.append(g)
    return bases, scalars


def msm(S: Suite, bases: Sequence[Point], scalars: Sequence[int], c: int = 8) -> Ext:
    """Plain Pippenger (any correct MSM yields the same group element as
    ark-ec's msm_unchecked, thin.rs:319)."""
    nwin = (S.r_bits + c - 1) // c
    total = EXT_ID
    ext_bases = [to_ext(b) for b in bases]
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = ext_double(S, total)
        buckets = [None] * (1 << c)
        for P, k in zip(ext_bases, scalars):
            dgt = (k >> (w * c)) & ((1 << c) - 1)
            if dgt:
                buckets[dgt] = P if buckets[dgt] is None else ext_add(S, buckets[dgt], P)
        run, acc = EXT_ID, EXT_ID
        for dgt in range((1 << c) - 1, 0, -1):
            if buckets[dgt] is not None:
                run = ext_add(S, run, buckets[dgt])
            acc = ext_add(S, acc, run)
        total = ext_add(S, total, acc)
    return total


def batch_verify(S: Suite, items: Sequence[BatchItem]) -> int:
    """BatchVerifier::verify (thin.rs:257-325)."""
    if not items:
        return OK
    if any(e.pk == IDENTITY or has_identity(e.ios) for e in items):
        return INVALID_DATA
    bases, scalars = batch_msm_terms(S, items)
    res = msm(S, bases, scalars)
    return OK if ext_is_identity(S, res) else VERIFICATION_FAILURE


# --------------------------------------------------------------------------------------
# Synthetic data set of SURVEY.md section 8(d) (small sizes only - Python is slow)
# --------------------------------------------------------------------------------------


@dataclass
class Proofs:
    suite: Suite
    pk: List[Point] = field(default_factory=list)
    ios: List[List[VrfIo]] = field(default_factory=list)
    ad: List[bytes] = field(default_factory=list)
    r: List[Point] = field(default_factory=list)
    s: List[int] = field(default_factory=list)


def synth_seed(k: int) -> bytes:
    return k.to_bytes(8, "little") + bytes(24)


def synth_msg(j: int, i: int) -> bytes:
    return j.to_bytes(8, "little") + i.to_bytes(4, "little")


def synth_proofs(S: Suite, n: int, m: int = 1, signers: int = 4096) -> Proofs:
    K = min(n, signers) or 1
    sks = [secret_from_seed(S, synth_seed(k)) for k in range(K)]
    pks = [public_key(S, sk) for sk in sks]
    out = Proofs(S)
    for j in range(n):
        k = j % K
        ios = []
        for i in range(m):
            inp = data_to_point(S, synth_msg(j, i))
            ios.append((inp, pt_mul(S, inp, sks[k])))
        ad = b"ad-%d" % j
        R, s = thin_prove(S, sks[k], ios, ad)
        out.pk.append(pks[k]); out.ios.append(ios); out.ad.append(ad)
        out.r.append(R); out.s.append(s)
    return out


# --------------------------------------------------------------------------------------
# Pedersen VRF (src/pedersen.rs) - SURVEY.md section 8(f) row 3: same MSM engine, 5N+2 points
# --------------------------------------------------------------------------------------


def vrf_transcript_plain(S: Suite, scheme: int, ios: Sequence[VrfIo], ad: bytes) -> Tuple[Transcript, VrfIo]:
    """utils::vrf_transcript (common.rs:181-225): no Schnorr pair; returns the transcript (ad absorbed)
    and the merged pair ((0,1),(0,1)) for n = 0, the pair itself for n = 1, merge_ios otherwise)."""
    t = Transcript(S.suite_id)
    t.absorb(bytes([scheme]))
    t.absorb(len(ios).to_bytes(8, "little"))
    for (i, o) in ios:
        t.absorb(enc_point(S, i) + enc_point(S, o))
    t.absorb(len(ad).to_bytes(8, "little"))
    t.absorb(ad)
    if len(ios) == 0:
        return t, (IDENTITY, IDENTITY)
    if len(ios) == 1:
        return t, ios[0]
    zt = t.clone()
    zt.absorb(bytes([DOM_DELINEARIZE]))
    zs = [1] + [challenge_scalar(S, zt) for _ in range(len(ios) - 1)]
    im, om = EXT_ID, EXT_ID
    for (i, o), z in zip(ios, zs):
        im = ext_add(S, im, ext_mul(S, to_ext(i), z))
        om = ext_add(S, om, ext_mul(S, to_ext(o), z))
    return t, (ext_to_affine(S, im), ext_to_affine(S, om))


@dataclass
class PedersenProof:                                          # pedersen.rs:43-50
    pk_com: Point
    r: Point
    ok: Point
    s: int
    sb: int


def pedersen_prove(S: Suite, sk: int, ios: Sequence[VrfIo], ad: bytes) -> Tuple[PedersenProof, int]:
    """pedersen::Prover::prove (pedersen.rs:136-186)."""
    t, io = vrf_transcript_plain(S, DOM_PEDERSEN, ios, ad)
    tb = t.clone()
    tb.absorb(bytes([DOM_PEDERSEN_BLINDING]))                 # PedersenSuite::blinding, pedersen.rs:28-31
    blinding = nonce(S, sk, tb)
    pk = public_key(S, sk)
    pk_com = ext_to_affine(S, ext_add(S, to_ext(pk), ext_mul(S, to_ext(S.B), blinding)))
    t.absorb(enc_point(S, pk_com))
    k = nonce(S, sk, t.clone())
    kb = nonce(S, blinding, t.clone())
    r = ext_to_affine(S, ext_add(S, ext_mul(S, to_ext(S.G), k), ext_mul(S, to_ext(S.B), kb)))
    ok = ext_to_affine(S, ext_mul(S, to_ext(io[0]), k))
    c = challenge(S, [r, ok], t)
    return PedersenProof(pk_com, r, ok, (k + c * sk) % S.r, (kb + c * blinding) % S.r), blinding


def pedersen_verify(S: Suite, ios: Sequence[VrfIo], ad: bytes, pf: PedersenProof) -> int:
    """pedersen::Verifier::verify (pedersen.rs:188-253)."""
    if pf.pk_com == IDENTITY or has_identity(ios):
        return INVALID_DATA
    t, io = vrf_transcript_plain(S, DOM_PEDERSEN, ios, ad)
    t.absorb(enc_point(S, pf.pk_com))
    c = challenge(S, [pf.r, pf.ok], t)
    lhs1 = ext_add(S, ext_mul(S, to_ext(io[0]), pf.s), ext_neg(S, ext_mul(S, to_ext(io[1]), c)))
    if not ext_is_identity(S, ext_add(S, lhs1, ext_neg(S, to_ext(pf.ok)))):
        return VERIFICATION_FAILURE
    lhs2 = ext_add(S, ext_add(S, ext_mul(S, to_ext(S.G), pf.s), ext_mul(S, to_ext(S.B), pf.sb)),
                   ext_neg(S, ext_mul(S, to_ext(pf.pk_com), c)))
    if not ext_is_identity(S, ext_add(S, lhs2, ext_neg(S, to_ext(pf.r)))):
        return VERIFICATION_FAILURE
    return OK


@dataclass
class PedersenItem:                                           # pedersen.rs:260-275
    c: int
    input: Point
    output: Point
    pf: PedersenProof
    io_identity: bool


def pedersen_batch_prepare(S: Suite, ios: Sequence[VrfIo], ad: bytes, pf: PedersenProof) -> PedersenItem:
    """BatchItem::new (pedersen.rs:283-301)."""
    t, io = vrf_transcript_plain(S, DOM_PEDERSEN, ios, ad)
    t.absorb(enc_point(S, pf.pk_com))
    c = challenge(S, [pf.r, pf.ok], t)
    return PedersenItem(c, io[0], io[1], pf, has_identity(ios))


def pedersen_batch_seed(S: Suite, items: Sequence[PedersenItem]) -> bytes:      # pedersen.rs:361-367
    h = hashlib.sha512()
    h.update(S.suite_id + bytes([DOM_BATCH]))
    for e in items:
        h.update(enc_scalar(e.c) + enc_scalar(e.pf.s) + enc_scalar(e.pf.sb))
    return h.digest()


def pedersen_batch_terms(S: Suite, items: Sequence[PedersenItem]) -> Tuple[List[Point], List[int]]:
    """bases / scalars as built at pedersen.rs:373-418 (t_i, u_i from 32-byte squeezes)."""
    seed = pedersen_batch_seed(S, items)
    bases: List[Point] = []
    scalars: List[int] = []
    g = b = 0
    r = S.r
    for j, e in enumerate(items):
        blk = hashlib.sha512(seed + (j // 2).to_bytes(8, "little")).digest()
        t = int.from_bytes(blk[32 * (j % 2):32 * (j % 2) + 16], "little")
        u = int.from_bytes(blk[32 * (j % 2) + 16:32 * (j % 2) + 32], "little")
        bases += [e.output, e.pf.ok, e.input, e.pf.pk_com, e.pf.r]
        scalars += [t * e.c % r, t, (-(t * e.pf.s)) % r, u * e.c % r, u]
        g = (g + u * e.pf.s) % r
        b = (b + u * e.pf.sb) % r
    bases += [S.G, S.B]
    scalars += [(-g) % r, (-b) % r]
    return bases, scalars


def pedersen_batch_verify(S: Suite, items: Sequence[PedersenItem]) -> int:      # pedersen.rs:341-426
    if not items:
        return OK
    if any(e.pf.pk_com == IDENTITY or e.io_identity for e in items):
        return INVALID_DATA
    bases, scalars = pedersen_batch_terms(S, items)
    return OK if ext_is_identity(S, msm(S, bases, scalars)) else VERIFICATION_FAILURE

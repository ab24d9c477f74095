"""CPU: the exact limb algorithms of the device headers (fp.cuh, curve.cuh, sha512.cuh, h2c.cuh,
thin.cuh), compiled for the host with the PTX carry flag emulated (tests/hostemu), against Python
big integers, hashlib and the golden vectors.  This is a test-only build of the kernels'
arithmetic; the product never runs it."""
import ctypes
import hashlib
import os
import random
import subprocess

import pytest

from oracle import pyref as o

HERE = os.path.dirname(os.path.abspath(__file__))
R = 1 << 256


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "hostemu", "libhostemu.so")
    src = os.path.join(HERE, "hostemu", "hostemu.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, src])
    return ctypes.CDLL(so)


def L(x):
    return (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def U(a, n=8):
    return sum(int(a[i]) << (32 * i) for i in range(n))


def test_square_roots(emu):
    """fe_sqrt (windowed Tonelli-Shanks with table look-ups for the 2-adicity-32 / -28 fields, the textbook loop for
    2^255 - 19) and the Elligator2 variant that returns sqrt(Z a) for a non-residue a."""
    rnd = random.Random(7)
    out = (ctypes.c_uint32 * 8)()
    for sid, S in o.SUITES.items():
        p = S.p
        vals = [1, 4, p - 1, 2, 3, 5] + [rnd.randrange(1, p) for _ in range(120)]
        vals += [pow(v, 2, p) for v in vals[:40]]
        n_sq = 0
        for a in vals:
            ok = emu.emu_sqrt(sid, L(a * R % p), out)
            is_sq = pow(a, (p - 1) // 2, p) == 1
            assert bool(ok) == is_sq, (sid, a)
            if is_sq:
                n_sq += 1
                r = U(out) * pow(R, -1, p) % p
                assert r * r % p == a
        assert n_sq > 60
        assert emu.emu_sqrt(sid, L(0), out) == 1 and U(out) == 0
    p, Z = o.BANDERSNATCH.p, 5
    for _ in range(100):
        a = rnd.randrange(1, p)
        ok = emu.emu_sqrt_or_z(L(a * R % p), out)
        r = U(out) * pow(R, -1, p) % p
        if pow(a, (p - 1) // 2, p) == 1:
            assert ok == 1 and r * r % p == a
        else:
            assert ok == 0 and r * r % p == Z * a % p


def test_subgroup_membership(emu):
    """in_prime_subgroup_v: the 2-descent test of Bandersnatch (two quadratic characters) and [r]P elsewhere, against
    [r]P == O computed with big integers, on random curve points of every coset of the prime-order subgroup."""
    rnd = random.Random(11)
    for sid, S in o.SUITES.items():
        p = S.p
        seen = {True: 0, False: 0}
        pts = [o.IDENTITY, (0, p - 1), S.G, o.pt_mul(S, S.G, 12345)]
        while len(pts) < (44 if sid == 0 else 12):
            P = o.x_from_y(S, rnd.randrange(p), rnd.random() < 0.5)
            if P is not None:
                pts.append(P)
        for P in pts:
            E = o.ext_mul(S, o.to_ext(P), S.r)
            want = E[0] % p == 0 and (E[1] - E[2]) % p == 0 and E[2] % p != 0
            arr = (ctypes.c_uint32 * 16)(*[((c * R % p) >> (32 * i)) & 0xFFFFFFFF for c in P for i in range(8)])
            assert bool(emu.emu_in_subgroup(sid, arr)) == want, (sid, P)
            seen[want] += 1
        assert seen[True] >= 3 and seen[False] >= 1


def test_glv_bandersnatch(emu):
    """GLV pieces of the Bandersnatch scalar multiplication (curve.cuh): the scalar split k = k1 + k2 lambda (mod r) with
    128-bit parts, psi(P) = lambda P on the prime-order subgroup, and the joint Booth multiplication against k P."""
    S = o.BANDERSNATCH
    p, r = S.p, S.r
    lam = o.fsqrt(r - 2, r)
    rnd = random.Random(21)
    k1b, k2b, neg = (ctypes.c_uint32 * 8)(), (ctypes.c_uint32 * 8)(), (ctypes.c_int * 2)()
    seen = set()
    lams = None
    for k in [0, 1, 2, r - 1, r, (1 << 256) - 1, 1 << 255] + [rnd.randrange(1 << 256) for _ in range(300)]:
        emu.emu_glv_split(L(k), k1b, k2b, neg)
        k1, k2 = U(k1b) * (-1 if neg[0] else 1), U(k2b) * (-1 if neg[1] else 1)
        assert abs(k1) < (1 << 128) and abs(k2) < (1 << 128)
        if lams is None:
            lams = [x for x in (lam, r - lam) if (k1 + k2 * x - k) % r == 0] if k > 2 else None
        if k > 2:
            assert lams and (k1 + k2 * lams[0] - k) % r == 0, hex(k)
        seen.add((neg[0], neg[1]))
    assert len(seen) >= 1
    lam = lams[0]
    out = (ctypes.c_uint32 * 32)()

    def aff_l(P):
        return (ctypes.c_uint32 * 16)(*[((c * R % p) >> (32 * i)) & 0xFFFFFFFF for c in P for i in range(8)])

    def ext_u(a):
        Rinv = pow(R, -1, p)
        return tuple(U(a[8 * j:8 * j + 8]) * Rinv % p for j in range(4))

    def same(e, Q):                     # projective equality with an extended point of the oracle
        X, Y, Z, T = e
        return (X * Q[2] - Q[0] * Z) % p == 0 and (Y * Q[2] - Q[1] * Z) % p == 0 and (T * Q[2] - Q[3] * Z) % p == 0 and Z % p != 0
    for _ in range(6):
        P = o.pt_mul(S, S.G, rnd.randrange(1, r))
        emu.emu_glv_psi(aff_l(P), out)
        assert same(ext_u(out), o.ext_mul(S, o.to_ext(P), lam))
        for k in [1, 2, r - 1, rnd.randrange(r), rnd.randrange(1 << 256), (1 << 256) - 1, 0]:
            emu.emu_glv_mul(aff_l(P), L(k), out)
            assert same(ext_u(out), o.ext_mul(S, o.to_ext(P), k % r)), hex(k)
        # shared-scalar plan (GLV split + width-5 NAF computed once): same products, sparse addition chain
        top = ctypes.c_int32(0)
        for k in [1, 2, 3, 15, 16, 17, 31, 32, 33, r - 1, r - 2, (r - 1) // 2, (1 << 128) - 1, 1 << 128, int("5" * 63, 16) % r,
                  int("a" * 63, 16) % r] + [rnd.randrange(r) for _ in range(6)]:
            emu.emu_glv_mul_plan(aff_l(P), L(k), out, ctypes.byref(top))
            assert same(ext_u(out), o.ext_mul(S, o.to_ext(P), k % r)), hex(k)
            assert top.value <= 130


def test_field_ops(emu):
    mods = [o.BANDERSNATCH.p, o.ED25519.p, o.BABYJUBJUB.p, o.BANDERSNATCH.r, o.ED25519.r, o.BABYJUBJUB.r]
    rnd = random.Random(1)
    out = (ctypes.c_uint32 * 8)()
    for f, p in enumerate(mods):
        def run(op, x, y=0):
            emu.emu_field_op(f, op, L(x), L(y), out)
            return U(out)
        edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p]
        cases = [(a, b) for a in edge for b in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(500)]
        for a, b in cases:
            assert run(0, a, b) == a * b * pow(R, -1, p) % p
            assert run(1, a, b) == (a + b) % p
            assert run(2, a, b) == (a - b) % p
            assert run(3, a) == (-a) % p
            assert run(4, a) == a * R % p
            assert run(5, a) == a * pow(R, -1, p) % p
        for _ in range(10):
            a = rnd.randrange(1, p)
            assert run(6, a * R % p) == pow(a, -1, p) * R % p
            x = rnd.randrange(R)
            assert run(7, x) == x % p
            assert run(8, a * R % p) == (pow(a, (p - 1) // 2, p) == 1)
        # Jacobi symbol (binary algorithm) against Euler's criterion: raw values and Montgomery images, edge values
        for a in edge + [p - 3, 3, 4, 5, 7, 8, 1 << 32, (1 << 32) - 1, 1 << 64, (1 << 224) + 1, 1 << 254] + [rnd.randrange(p) for _ in range(150)]:
            a %= p
            e = pow(a, (p - 1) // 2, p)
            want = 0 if a == 0 else (1 if e == 1 else -1)
            assert run(10, a) - 1 == want, (f, hex(a))
            assert run(8, a) == (want == 1) and run(9, a * R % p) == (want == 1)


def test_lazy_reduction(emu):
    """fp.cuh lazy reduction: mul_wide is the exact 512-bit product of any two 256-bit values; redc_wide is t R^-1 mod p
    for every t < p 2^256 (boundary values included); the Bandersnatch and Baby-JubJub mixed additions built on them return exactly
    the coordinates of add-2008-hwcd with Z2 = 1 for ARBITRARY field elements (extreme limbs, not only curve points)."""
    rnd = random.Random(5)
    out16, out8 = (ctypes.c_uint32 * 16)(), (ctypes.c_uint32 * 8)()

    def L16(x):
        return (ctypes.c_uint32 * 16)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(16)])
    M = R - 1
    edge = [0, 1, M, M - 1, 1 << 255, (1 << 255) - 1, 0xFFFFFFFF, M ^ 0xFFFFFFFF, int("f0" * 32, 16), int("0f" * 32, 16),
            sum(0xFFFFFFFF << (64 * i) for i in range(4)), sum(0xFFFFFFFF << (64 * i + 32) for i in range(4))]
    for a, b in [(a, b) for a in edge for b in edge] + [(rnd.randrange(R), rnd.randrange(R)) for _ in range(500)]:
        emu.emu_mul_wide(L(a), L(b), out16)
        assert U(out16, 16) == a * b, (hex(a), hex(b))
    mods = [o.BANDERSNATCH.p, o.ED25519.p, o.BABYJUBJUB.p, o.BANDERSNATCH.r, o.ED25519.r, o.BABYJUBJUB.r]
    for f, p in enumerate(mods):
        if f in (1,):
            continue                          # 2^255 - 19 has its own reduction; the lazy form is not used there
        Rinv = pow(R, -1, p)
        top = p * R
        ts = [0, 1, R - 1, R, R + 1, top - 1, top - R, top - R - 1, (p - 1) * R, (p - 1) * (p - 1), 2 * (p - 1) * (p - 1),
              top - (1 << 32), (1 << 32) - 1, ((1 << 32) - 1) << 224, top // 2]
        ts += [rnd.randrange(top) for _ in range(400)]
        ts += [rnd.randrange(1 << 32) << (32 * k) for k in range(15)]
        for t in ts:
            assert t < top
            emu.emu_redc_wide(f, L16(t), out8)
            assert U(out8) == t * Rinv % p, (f, hex(t))
    for sid_m, S, a_coeff in ((0, o.BANDERSNATCH, -5), (2, o.BABYJUBJUB, 1)):
        p = S.p
        Rinv = pow(R, -1, p)
        out = (ctypes.c_uint32 * 32)()
        ext = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p, (1 << 253) % p, p - (1 << 200)]
        cases = [([a] * 4, [b] * 3) for a in ext for b in ext]
        cases += [([rnd.choice(ext) for _ in range(4)], [rnd.choice(ext) for _ in range(3)]) for _ in range(300)]
        cases += [([rnd.randrange(p) for _ in range(4)], [rnd.randrange(p) for _ in range(3)]) for _ in range(300)]
        for acc, base in cases:
            # raw limbs are taken as Montgomery representatives; the formula is checked on the values they stand for
            X1, Y1, Z1, T1 = [c * Rinv % p for c in acc]
            x2, y2, k2 = [c * Rinv % p for c in base]
            A, B, C = X1 * x2 % p, Y1 * y2 % p, T1 * k2 % p
            E = ((X1 + Y1) * (x2 + y2) - A - B) % p
            F, G, H = (Z1 - C) % p, (Z1 + C) % p, (B - a_coeff * A) % p
            want = (E * F % p, G * H % p, F * G % p, E * H % p)
            pa = (ctypes.c_uint32 * 32)(*[(c >> (32 * i)) & 0xFFFFFFFF for c in acc for i in range(8)])
            pb = (ctypes.c_uint32 * 24)(*[(c >> (32 * i)) & 0xFFFFFFFF for c in base for i in range(8)])
            emu.emu_point_op(sid_m, 0, pa, pb, out)
            got = tuple(U(out[8 * j:8 * j + 8]) * Rinv % p for j in range(4))
            assert all(U(out[8 * j:8 * j + 8]) < p for j in range(4))
            assert got == want, (sid_m, acc, base)


def test_point_ops(emu):
    rnd = random.Random(2)
    for sidx, S in o.SUITES.items():
        p = S.p

        def ext_l(P):
            arr = (ctypes.c_uint32 * 32)()
            for j, c in enumerate(P):
                for i in range(8):
                    arr[8 * j + i] = ((c * R % p) >> (32 * i)) & 0xFFFFFFFF
            return arr

        def ext_u(arr):
            return tuple(U(arr[8 * j:8 * j + 8]) * pow(R, -1, p) % p for j in range(4))

        def same(a, b):
            return o.ext_to_affine(S, a) == o.ext_to_affine(S, b)
        G = o.to_ext(S.G)
        out = (ctypes.c_uint32 * 32)()
        for _ in range(8):
            k1, k2 = rnd.randrange(S.r), rnd.randrange(S.r)
            P, Q = o.ext_mul(S, G, k1), o.ext_mul(S, G, k2)
            lam = rnd.randrange(1, p)
            P = tuple(c * lam % p for c in P)
            qa = o.ext_to_affine(S, Q)
            kk = (ctypes.c_uint32 * 24)()
            for j, c in enumerate([qa[0], qa[1], S.d * qa[0] * qa[1] % p]):
                for i in range(8):
                    kk[8 * j + i] = ((c * R % p) >> (32 * i)) & 0xFFFFFFFF
            if sidx != 1:                     # (x, y, d x y) layout; Ed25519 bases are (y-x, y+x, 2dxy)
                emu.emu_point_op(sidx, 0, ext_l(P), kk, out)
                assert same(ext_u(out), o.ext_add(S, P, Q))
            qaff = (ctypes.c_uint32 * 16)()
            for jj, c in enumerate(qa):
                for i in range(8):
                    qaff[8 * jj + i] = ((c * R % p) >> (32 * i)) & 0xFFFFFFFF
            emu.emu_point_op(sidx, 5, ext_l(P), qaff, out)            # affine_to_k + mixed addition
            assert same(ext_u(out), o.ext_add(S, P, Q))
            emu.emu_point_op(sidx, 6, ext_l(P), qaff, out)            # ... of the negated base
            assert same(ext_u(out), o.ext_add(S, P, o.ext_neg(S, Q)))
            emu.emu_point_op(sidx, 5, ext_l(Q), qaff, out)            # unified: doubling through the mixed add
            assert same(ext_u(out), o.ext_double(S, Q))
            emu.emu_point_op(sidx, 1, ext_l(P), ext_l(Q), out)
            assert same(ext_u(out), o.ext_add(S, P, Q))
            emu.emu_point_op(sidx, 1, ext_l(P), ext_l(P), out)        # unified: doubling through add
            assert same(ext_u(out), o.ext_double(S, P))
            emu.emu_point_op(sidx, 2, ext_l(P), None, out)
            assert same(ext_u(out), o.ext_double(S, P))
            emu.emu_point_op(sidx, 3, ext_l(P), L(k2), out)
            assert same(ext_u(out), o.ext_mul(S, P, k2))
            emu.emu_point_op(sidx, 4, ext_l(P), None, out)
            assert bytes(out)[:32] == o.enc_point(S, o.ext_to_affine(S, P))
        emu.emu_point_op(sidx, 1, ext_l(o.EXT_ID), ext_l(G), out)
        assert same(ext_u(out), G)
        # Booth-recoded windowed scalar multiplication: edge scalars (all-ones windows, top bit, digit 8, group order)
        for kk2 in [0, 1, 7, 8, 9, 15, 16, 0x88888888, (1 << 255) + 12345, (1 << 256) - 1, S.r, S.r - 1,
                    int("7" * 64, 16), int("8" * 64, 16), int("f" * 63, 16)]:
            emu.emu_point_op(sidx, 3, ext_l(G), L(kk2), out)
            assert same(ext_u(out), o.ext_mul(S, G, kk2)), hex(kk2)


def test_sha512(emu):
    for n in [0, 1, 55, 111, 112, 113, 127, 128, 129, 200, 255, 256, 1000]:
        m = os.urandom(n)
        out = (ctypes.c_uint8 * 64)()
        emu.emu_sha512(m, n, out)
        assert bytes(out) == hashlib.sha512(m).digest()


def test_h2c_and_prove_golden(emu, golden):
    def words(x):
        return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    for sidx, S in o.SUITES.items():
        def aff(P):
            return (ctypes.c_uint32 * 16)(*(words(P[0] * R % S.p) + words(P[1] * R % S.p)))

        def unaff(a):
            ri = pow(R, -1, S.p)
            return (U(a[:8]) * ri % S.p, U(a[8:16]) * ri % S.p)
        for v in golden[sidx]:
            alpha, ad = bytes.fromhex(v["alpha"]), bytes.fromhex(v["ad"])
            out = (ctypes.c_uint32 * 16)()
            assert emu.emu_h2c(sidx, alpha, len(alpha), out) == 1
            h = unaff(out)
            assert o.enc_point(S, h).hex() == v["h"]
            sk = int.from_bytes(bytes.fromhex(v["sk"]), "little")
            pk = o.dec_point(S, bytes.fromhex(v["pk"]))
            gamma = o.dec_point(S, bytes.fromhex(v["gamma"]))
            ios = (ctypes.c_uint32 * 32)(*(list(aff(h)) + list(aff(gamma))))
            r16, s8 = (ctypes.c_uint32 * 16)(), (ctypes.c_uint32 * 8)()
            emu.emu_prove(sidx, (ctypes.c_uint32 * 8)(*words(sk)), aff(pk), ios, 1, ad, len(ad), r16, s8)
            assert o.enc_point(S, unaff(r16)).hex() == v["proof_r"]
            assert bytes(s8).hex() == v["proof_s"]


def test_elligator2_random_messages_vs_oracle(emu):
    """The device Elligator2 path (shared inversion, g(x2) = Z u^2 g(x1) shortcut, projective change of model) against
    the plain restatement of the reference on seeded messages of assorted lengths - both square / non-square branches
    and both signs occur many times in 120 draws."""
    import random
    S = o.BANDERSNATCH
    rng = random.Random(20260)
    ri = pow(R, -1, S.p)
    for k in range(120):
        msg = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 8, 12, 31, 32, 33, 100])))
        out = (ctypes.c_uint32 * 16)()
        assert emu.emu_h2c(0, msg, len(msg), out) == 1
        got = (U(out[:8]) * ri % S.p, U(out[8:16]) * ri % S.p)
        assert got == tuple(o.data_to_point(S, msg)), (k, msg.hex())

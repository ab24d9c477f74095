#!/usr/bin/env python3
"""Generate ark_vrf_b200/csrc/constants_gen.h (field and curve constants in 8x32-bit limbs).

The numbers come from SURVEY.md Appendix A.1 (suite files src/suites/*.rs of the reference
and the arkworks curve crates); everything derived (Montgomery constants, Tonelli-Shanks
roots, Elligator2 ratios) is computed here.  Run:  python tools/gen_constants.py
This script does not import the oracle: the product's constants must stand on their own.
"""
import os

R = 1 << 256

P_BLS = 52435875175126190479447740508185965837690552500527637822603658699938581184513
P_25519 = 2**255 - 19
P_BN = 21888242871839275222246405745257275088548364400416034343698204186575808495617

SUITES = [
    dict(
        name="BANDERSNATCH_SHA512_ELL2", suite_id=b"Bandersnatch-SHA512-ELL2-v1",
        p=P_BLS, a=-5,
        d=45022363124591815672509500913686876175488063829319466900776701791074614335719,
        r=13108968793781547619861935127046491459309155893440570251786403306729687672801,
        cof_log2=2,
        gx=18886178867200960497001835917649091219057080094937609519140440539760939937304,
        gy=19188667384257783945677642223292697773471335439753913231509108946878080696678,
        mont_j=29978822694968839326280996386011761570173833766074948509196803838190355340952,
        mont_k=25465760566081946422412445027709227188579564747101592991722834452325077642517,
        ell2_z=5,
        bx=23335687741101763108036518445642207119627658113885888016488710494487028845889,
        by=5552214580375038693022409684979828600325210968745774080859660443337357929963,
    ),
    dict(
        name="ED25519_SHA512_TAI", suite_id=b"Ed25519-SHA512-TAI-v1",
        p=P_25519, a=-1,
        d=(-121665 * pow(121666, -1, P_25519)) % P_25519,
        r=2**252 + 27742317777372353535851937790883648493,
        cof_log2=3,
        gx=15112221349535400772501151409588531511454012693041857206046113283949847762202,
        gy=46316835694926478169428394003475163141307993866256225615783033603165251855960,
        mont_j=0, mont_k=0, ell2_z=0,
        bx=45003173884697328536089278691112838614164406922820087464913813433380838325453,
        by=31256014272390301975555524011230972931324093235775711248505761870355310252869,
    ),
    dict(
        name="BABYJUBJUB_SHA512_TAI", suite_id=b"BabyJubJub-SHA512-TAI-v1",
        p=P_BN, a=1,
        d=9706598848417545097372247223557719406784115219466060233080913168975159366771,
        r=2736030358979909402780800718157159386076813972158567259200215660948447373041,
        cof_log2=3,
        gx=19698561148652590122159747500897617769866003486955115824547446575314762165298,
        gy=19298250018296453272277890825869354524455968081175474282777126169995084727839,
        mont_j=0, mont_k=0, ell2_z=0,
        bx=15549380791300914366206471199568039679131690710803662429646809536753521087193,
        by=15218614024055502695611547593111691164731001864276292210438920202280814188379,
    ),
]


def limbs(x):
    assert 0 <= x < R
    return "{" + ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(8)) + "}"


def field_consts(p):
    n0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    return "{ %s,\n    %s,\n    %s,\n    %s,\n    %s,\n    0x%08xu, {0,0,0,0,0,0,0} }" % (
        limbs(p), limbs(R % p), limbs(R * R % p), limbs(p - 2), limbs((p - 1) // 2), n0)


def mont(x, p):
    return (x % p) * R % p


def ts_params(p):
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    return s, (q - 1) // 2, pow(z, q, p)


def fsqrt(n, p):
    """Tonelli-Shanks (host, generator only)."""
    n %= p
    if n == 0:
        return 0
    if pow(n, (p - 1) // 2, p) != 1:
        return None
    s, qh, _ = ts_params(p)
    q = 2 * qh + 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(n, q, p), pow(n, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c, t, r = i, b * b % p, t * b * b % p, r * b % p
    return r


def ts_tables(p):
    """Tables for the windowed Tonelli-Shanks square root (csrc/h2c.cuh, ts_dlog_sqrt): p - 1 = 2^s q, g = z^q generates the
    2^s-order subgroup, the discrete log of an element of that subgroup is found in four windows of w = s/4 bits.
      pow[i][j]  = g^(-j * 2^(w i))            (Montgomery form)      i = 0..3, j < 2^w
      look[k]    = j with  limb0(Montgomery(g^(j * 2^(3 w)))) & 1023 probing to k, else 0xffff
    """
    s, _, g = ts_params(p)
    if s % 4 or s < 8:
        return None
    w = s // 4
    ginv = pow(g, -1, p)
    powt = [[mont(pow(ginv, j << (w * i), p), p) for j in range(1 << w)] for i in range(4)]
    top = pow(g, 1 << (3 * w), p)
    look = [0xFFFF] * 1024
    v = 1
    for j in range(1 << w):
        k = mont(v, p) & 1023
        while look[k] != 0xFFFF:
            k = (k + 1) & 1023
        look[k] = j
        v = v * top % p
    return w, powt, look


# ---- GLV for Bandersnatch (the curve's degree-2 endomorphism psi, psi^2 = -2) ---------------------------------------
def te_add(P, Q, a, d, p):
    (x1, y1), (x2, y2) = P, Q
    t = d * x1 * x2 * y1 * y2 % p
    return ((x1 * y2 + y1 * x2) * pow(1 + t, -1, p) % p, (y1 * y2 - a * x1 * x2) * pow(1 - t, -1, p) % p)


def te_mul(P, k, a, d, p):
    R = (0, 1)
    while k:
        if k & 1:
            R = te_add(R, P, a, d, p)
        P = te_add(P, P, a, d, p)
        k >>= 1
    return R


def nullvec(M, p):
    """One non-zero vector of the (one-dimensional) null space of M over GF(p)."""
    M = [row[:] for row in M]
    rows, cols, piv, rr = len(M), len(M[0]), [], 0
    for c in range(cols):
        pr = next((i for i in range(rr, rows) if M[i][c] % p), None)
        if pr is None:
            continue
        M[rr], M[pr] = M[pr], M[rr]
        inv = pow(M[rr][c], -1, p)
        M[rr] = [x * inv % p for x in M[rr]]
        for i in range(rows):
            if i != rr and M[i][c] % p:
                f = M[i][c]
                M[i] = [(x - f * y) % p for x, y in zip(M[i], M[rr])]
        piv.append(c)
        rr += 1
    free = [c for c in range(cols) if c not in piv]
    assert len(free) == 1
    v = [0] * cols
    v[free[0]] = 1
    for i, c in enumerate(piv):
        v[c] = (-M[i][free[0]]) % p
    return v


def glv_params(s):
    """lambda = sqrt(-2) mod r acts on the prime-order subgroup as the endomorphism
         psi(x, y) = ( x (y^2 + E0) / (C1 y) ,  (y^2 + BN) / (BD y^2 - 1) ).
    The four constants are FITTED here to sample pairs (P, lambda P) by linear algebra over GF(p) and checked on fresh
    points; the decomposition k = k1 + k2 lambda (mod r) uses a reduced basis (a1,b1),(a2,b2) of the lattice
    {(a,b): a + b lambda = 0 mod r}:  c_i = (k * g_i) >> 384,  k1 = k - c1 A1 - c2 A2,  k2 = -(c1 B1 + c2 B2)."""
    import math
    p, r, a, d = s["p"], s["r"], s["a"] % s["p"], s["d"]
    lam = fsqrt(r - 2, r)
    G = (s["gx"], s["gy"])
    pts = []
    for i in range(10):
        P = te_mul(G, 0x1234567 + 977 * i, a, d, p)
        pts.append((P, te_mul(P, lam, a, d, p)))
    vy = nullvec([[Q[1] * pow(P[1], e, p) % p for e in range(3)] + [(-pow(P[1], e, p)) % p for e in range(3)] for P, Q in pts], p)
    vx = nullvec([[Q[0] * pow(P[1], e, p) % p for e in range(3)] + [(-P[0] * pow(P[1], e, p)) % p for e in range(3)] for P, Q in pts], p)
    # normalise to the documented shape
    assert vy[1] == 0 and vy[4] == 0 and vx[0] == 0 and vx[2] == 0 and vx[4] == 0
    sy = pow(vy[5], -1, p)
    bd, bn, m1 = vy[2] * sy % p, vy[3] * sy % p, vy[0] * sy % p
    assert m1 == p - 1
    sx = pow(vx[5], -1, p)
    c1, e0 = vx[1] * sx % p, vx[3] * sx % p

    def psi(P):
        x, y = P
        return (x * (y * y + e0) * pow(c1 * y, -1, p) % p, (y * y + bn) * pow(bd * y * y - 1, -1, p) % p)
    for i in range(5):
        P = te_mul(G, 0xABCDEF123 + i, a, d, p)
        assert psi(P) == te_mul(P, lam, a, d, p)
    # reduced lattice basis (extended Euclid on (r, lambda))
    r0, r1, t0, t1, rows = r, lam, 0, 1, []
    while r1:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        t0, t1 = t1, t0 - q * t1
        rows.append((r0, t0))
    idx = next(i for i, (rem, _) in enumerate(rows) if rem < math.isqrt(r))
    cand = [(rows[idx][0], -rows[idx][1]), (rows[idx - 1][0], -rows[idx - 1][1])]
    if idx + 1 < len(rows):
        cand.append((rows[idx + 1][0], -rows[idx + 1][1]))
    (a1, b1) = cand[0]
    (a2, b2) = min(cand[1:], key=lambda v: max(abs(v[0]), abs(v[1])))
    det = a1 * b2 - a2 * b1
    assert abs(det) == r and (a1 + b1 * lam) % r == 0 and (a2 + b2 * lam) % r == 0
    SH = 384
    g1, s1 = (abs(b2) << SH) // r, (1 if (b2 > 0) == (det > 0) else -1)
    g2, s2 = (abs(b1) << SH) // r, (1 if (-b1 > 0) == (det > 0) else -1)
    A1, A2, B1, B2 = s1 * a1, s2 * a2, s1 * b1, s2 * b2
    import random
    rnd = random.Random(5)
    for _ in range(2000):
        k = rnd.randrange(1 << 256)
        m1_, m2_ = (k * g1) >> SH, (k * g2) >> SH
        k1, k2 = k - m1_ * A1 - m2_ * A2, -(m1_ * B1 + m2_ * B2)
        assert (k1 + k2 * lam - k) % r == 0 and abs(k1) < (1 << 128) and abs(k2) < (1 << 128)
    return dict(bd=bd, bn=bn, c1=c1, e0=e0, g1=g1, g2=g2, A1=A1, A2=A2, B1=B1, B2=B2, lam=lam)


def limbs_n(x, n):
    x %= 1 << (32 * n)            # two's complement for negative values
    return "{" + ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(n)) + "}"


def write_glv(out):
    s = SUITES[0]
    g = glv_params(s)
    p = s["p"]
    out.append("// Bandersnatch GLV constants (glv_params in tools/gen_constants.py): endomorphism psi(x,y) =")
    out.append("// (x (y^2 + E0) / (C1 y), (y^2 + BN) / (BD y^2 - 1)) in Montgomery form, rounding multipliers g1, g2 (9 limbs),")
    out.append("// basis terms A1, A2, B1, B2 as 320-bit two's complement (10 limbs), lambda (plain, for tests).")
    out.append("#define AVRF_GLV_CONSTS_INIT { %s, %s, %s, %s, \\\n  %s, %s, \\\n  %s, %s, %s, %s, \\\n  %s }" % (
        limbs(mont(g["bd"], p)), limbs(mont(g["bn"], p)), limbs(mont(g["c1"], p)), limbs(mont(g["e0"], p)),
        limbs_n(g["g1"], 9), limbs_n(g["g2"], 9),
        limbs_n(g["A1"], 10), limbs_n(g["A2"], 10), limbs_n(g["B1"], 10), limbs_n(g["B2"], 10), limbs(g["lam"])))
    out.append("")


def write_ts_tables():
    out = ["// GENERATED by tools/gen_constants.py - do not edit.",
           "// Windowed Tonelli-Shanks tables for the base fields with high 2-adicity: index 0 = BLS12-381 Fr (Bandersnatch,",
           "// 2-adicity 32, 8-bit windows), index 1 = BN254 Fr (Baby-JubJub, 2-adicity 28, 7-bit windows).  See ts_tables() in",
           "// tools/gen_constants.py for the definition.",
           "#define AVRF_TS_POW_INIT { \\"]
    tabs = [ts_tables(P_BLS), ts_tables(P_BN)]
    rows = []
    for w, powt, _ in tabs:
        fr = []
        for i in range(4):
            ent = [limbs(powt[i][j]) if j < (1 << w) else limbs(0) for j in range(256)]
            fr.append("{ " + ", ".join(ent) + " }")
        rows.append("{ " + ", \\\n".join(fr) + " }")
    out.append(", \\\n".join(rows) + " \\")
    out.append("}")
    out.append("#define AVRF_TS_LOOK_INIT { \\")
    out.append(", \\\n".join("{ " + ", ".join(str(x) for x in look) + " }" for _, _, look in tabs) + " \\")
    out.append("}")
    out.append("#define AVRF_TS_WBITS_INIT { %d, %d }" % (tabs[0][0], tabs[1][0]))
    out.append("")
    path = os.path.join(os.path.dirname(__file__), "..", "ark_vrf_b200", "csrc", "ts_tables_gen.h")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("wrote", os.path.normpath(path))


def main():
    write_ts_tables()
    out = []
    out.append("// GENERATED by tools/gen_constants.py - do not edit.")
    out.append("// Field order: FQ_BAND, FQ_ED, FQ_BJJ, FR_BAND, FR_ED, FR_BJJ")
    out.append("#define AVRF_FIELD_CONSTS_INIT { \\")
    rows = [field_consts(s["p"]) for s in SUITES] + [field_consts(s["r"]) for s in SUITES]
    out.append(", \\\n".join("  " + r.replace("\n", " \\\n") for r in rows) + " \\")
    out.append("}")
    out.append("")
    out.append("#define AVRF_CURVE_CONSTS_INIT { \\")
    crow = []
    for s in SUITES:
        p = s["p"]
        ts_s, ts_exp, ts_root = ts_params(p)
        d = s["d"]
        gk = d * s["gx"] * s["gy"] % p
        if s["mont_k"]:
            kinv = pow(s["mont_k"], -1, p)
            jk = s["mont_j"] * kinv % p
            k2inv = kinv * kinv % p
            cw = pow(s["ell2_z"], ts_exp, p)      # Z^((q-1)/2): lets sqrt(Z*a) reuse the exponentiation done for sqrt(a)
        else:
            jk = k2inv = cw = 0
        # 2-descent constant of the cofactor-4 subgroup test (csrc/feeders.cuh): alpha = a root of u^2 + A u + 1 on the
        # Montgomery model u = (1 + y) / (1 - y), A = 2 (a + d) / (a - d); only for curves with full rational 2-torsion
        alpha = 0
        if s["cof_log2"] == 2:
            A = 2 * (s["a"] + d) * pow(s["a"] - d, -1, p) % p
            sd = fsqrt((A * A - 4) % p, p)
            assert sd is not None
            alpha = (-A + sd) * pow(2, -1, p) % p
        genc = s["gy"] | ((1 << 255) if (s["gx"] > p - s["gx"]) else 0)
        sid = s["suite_id"]
        sid_bytes = ", ".join(str(b) for b in sid.ljust(32, b"\0"))
        crow.append(
            "  { %s, %s, %s, %s, \\\n    %s, %s, %s, \\\n    %s, %s, %s, %s, %s, %s, \\\n    %s, %s, %s, %s, \\\n    %du, %du, %du, %du, %du, {0,0,0}, {%s} }" % (
                limbs(mont(d, p)), limbs(mont(s["gx"], p)), limbs(mont(s["gy"], p)), limbs(mont(gk, p)),
                limbs(ts_exp), limbs(mont(ts_root, p)), limbs(mont(2 * d, p)),
                limbs(mont(jk, p)), limbs(mont(k2inv, p)), limbs(mont(s["mont_k"], p)), limbs(mont(s["ell2_z"], p)), limbs(mont(cw, p)), limbs(genc),
                limbs(mont(s["bx"], p)), limbs(mont(s["by"], p)), limbs(mont(d * s["bx"] * s["by"] % p, p)), limbs(mont(alpha, p)),
                ts_s, s["cof_log2"], len(sid), p.bit_length(), s["r"].bit_length(), sid_bytes))
    out.append(", \\\n".join(crow) + " \\")
    out.append("}")
    out.append("")
    write_glv(out)
    path = os.path.join(os.path.dirname(__file__), "..", "ark_vrf_b200", "csrc", "constants_gen.h")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()

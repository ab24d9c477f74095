import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ark_vrf_b200 as av
from oracle import pyref as o
from helpers import *
av.load().avrf_init(0)
S = o.BANDERSNATCH
vs = json.load(open("tests/golden/bandersnatch_sha-512_ell2_thin.json"))
pr = golden_proofs(S, vs)
items = oracle_items(pr)
for mont in (False, True):
    bv = av.BatchVerifier(0, av.Format.MONTGOMERY if mont else av.Format.CANONICAL)
    bv.push_many(*arrays_from_proofs(pr, mont))
    print("mont", mont, "status", bv.verify_status())
    c = bv.tap(av.Tap.C).reshape(-1, 16)
    print(" c ok", [bytes(x) for x in c] == [e.c.to_bytes(16, "little") for e in items])
    z = bv.tap(av.Tap.Z).reshape(-1, 16)
    print(" z ok", [bytes(x) for x in z] == [e.zs[1].to_bytes(16, "little") for e in items])
    cs = bv.cs_stream()
    print(" s ok", [bytes(x[32:]) for x in cs] == [sc_bytes(e.s) for e in items])
    print(" seed ok", bytes(bv.tap(av.Tap.SEED)) == o.batch_seed(S, items))
    _, scalars = o.batch_msm_terms(S, items)
    sc = bv.tap(av.Tap.SCALARS).reshape(-1, 32)
    print(" scalars ok", [bytes(x) for x in sc] == [sc_bytes(k) for k in scalars])
    renc = bv.tap(av.Tap.R_COMPRESSED).reshape(-1, 32)
    print(" renc ok", [bytes(x).hex() for x in renc] == [v["proof_r"] for v in vs])
    part = bytes(bv.tap(av.Tap.PARTIAL))
    Rm = 1 << 256
    X, Y, Z, T = [int.from_bytes(part[32 * i:32 * i + 32], "little") * pow(Rm, -1, S.p) % S.p for i in range(4)]
    print(" partial identity", X == 0 and Y == Z)

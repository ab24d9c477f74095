import numpy as np, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ark_vrf_b200 as av
from ark_vrf_b200 import synth, ops
b = synth.make_batch(0, 1500, 1, fmt=av.Format.MONTGOMERY)
bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
print("verify", bv.verify_status(), "each", int((bv.verify_each() != 0).sum()))
enc = ops.point_compress(0, b.pk, av.Format.MONTGOMERY)
pts, ok = ops.points_deserialize(0, enc, 1, av.Format.MONTGOMERY)
print("ingest", bool(ok.all()), bool((pts == b.pk).all()))
srv = av.BatchServer(0, av.Format.MONTGOMERY, workers=3, hashers=1)
ts = [srv.submit(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s) for _ in range(6)]
print("server", [srv.wait(t) for t in ts])
srv.close()
for sid, m in ((1, 1), (2, 3)):
    c = synth.make_batch(sid, 700, m, fmt=av.Format.CANONICAL)
    v = av.BatchVerifier(sid, av.Format.CANONICAL)
    v.push_many(c.pk, c.ios, c.io_offsets, c.ad_blob, c.ad_offsets, c.r, c.s)
    print("suite", sid, v.verify_status())
# round 2: exact single-proof verifier, staged single pushes, fused Input::new + Secret::output, in-library sharding
import ctypes
lib = av.load()
st = ctypes.c_int32(-1)
for j in (0, 1, 2):
    a0, a1 = int(b.ad_offsets[j]), int(b.ad_offsets[j + 1])
    rc = lib.avrf_thin_verify_one(0, 0, b.pk[j].ctypes.data, b.ios[j].ctypes.data, 1, b.ad_blob[a0:].ctypes.data, a1 - a0,
                                  b.r[j].ctypes.data, b.s[j].ctypes.data, ctypes.byref(st))
    assert rc == 0
    print("verify_one", st.value)
sp = av.BatchVerifier(0, av.Format.MONTGOMERY)
for j in range(1500):
    a0, a1 = int(b.ad_offsets[j]), int(b.ad_offsets[j + 1])
    assert lib.avrf_thin_batch_push(sp._h, b.pk[j].ctypes.data, b.ios[j].ctypes.data, 1, b.ad_blob[a0:].ctypes.data, a1 - a0,
                                    b.r[j].ctypes.data, b.s[j].ctypes.data) == 0
print("staged pushes", sp.verify_status())
msgs = [b"m%d" % i for i in range(777)]
sk = np.frombuffer(synth.secret_from_seed(0, bytes(32)).to_bytes(32, "little"), dtype=np.uint8).copy()
res = ops.vrf_io_many(0, msgs, None, sk, want_hashes=True)
print("io_many", bool(res["ok"].all()))
av.init_multi(1)
sh = av.ShardedBatchVerifier(0, av.Format.MONTGOMERY)
sh.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
print("sharded", sh.verify_status())
sh.close()
# round 2, second half: wire-format push (chunked decode, rollback), hash pool handles, mixed server, lazy-reduction MSM at a
# size that spans several accumulation segments per bin
n2 = 160000
c = synth.make_batch(0, n2, 1, signers=64, fmt=av.Format.CANONICAL)
pk32, r32 = ops.point_compress(0, c.pk), ops.point_compress(0, c.r)
ios32 = np.concatenate([ops.point_compress(0, c.ios[:128 * n2].reshape(-1, 64)).reshape(-1), np.zeros(64, np.uint8)])
wv = av.BatchVerifier(0, av.Format.MONTGOMERY)
print("wire push", bool(wv.push_compressed(pk32, ios32, c.io_offsets, c.ad_blob, c.ad_offsets, r32, c.s).all()), wv.verify_status())
bad = pk32.copy()
bad[n2 - 5] = 0xFF
ok = wv.push_compressed(bad, ios32, c.io_offsets, c.ad_blob, c.ad_offsets, r32, c.s)
print("wire rollback", int((ok == 0).sum()), len(wv), wv.verify_status())
pool = av.HashPool(1)
hs = [av.BatchVerifier(0, av.Format.MONTGOMERY) for _ in range(3)]
for h in hs:
    h.set_hash_pool(pool)
    h.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
print("pool handles", [h.verify_status() for h in hs])
for h in hs:
    h.close()
pool.close()
srv = av.BatchServer(0, av.Format.MONTGOMERY, workers=4, hashers=1, own_hash_workers=2)
ts = [srv.submit(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s) for _ in range(8)]
print("mixed server", [srv.wait(t) for t in ts])
srv.close()

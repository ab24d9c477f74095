import numpy as np, sys, os
sys.path.insert(0, "/root/repo")
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

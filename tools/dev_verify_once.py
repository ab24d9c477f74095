"""One 2^k-proof Bandersnatch batch: generate, push, verify twice (profiling target)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ark_vrf_b200 as av
from ark_vrf_b200 import synth
av.load().avrf_init(0)
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = synth.make_batch(0, n, 1, fmt=av.Format.MONTGOMERY)
bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
for _ in range(reps):
    t0 = time.time(); st = bv.verify_status(); print("status", st, "s", round(time.time() - t0, 4), bv.timings())

"""Development probe run under gpurun: microbenchmarks + per-phase timings at a few sizes."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ark_vrf_b200 as av
from ark_vrf_b200 import ops, synth

av.load().avrf_init(0)
out = {}
for kind, iters, name in [(0, 4096, "wide_macs_per_s"), (3, 4096, "wide_macs_carry_per_s"), (4, 4096, "imad32_per_s"), (6, 4096, "wide_mac_carry_out_only_per_s"), (7, 4096, "wide_mac_carry_in_only_per_s"),
                          (1, 2000, "mont_mul_per_s"), (2, 400, "madd_per_s")]:
    best = 0
    for _ in range(3):
        v, ms = ops.microbench(kind, iters)
        best = max(best, v)
    out[name] = best
    print(name, "%.4g" % best, "ms", ms, flush=True)
sizes = [int(x) for x in sys.argv[1:]] or [1 << 12, 1 << 16, 1 << 20]
for n in sizes:
    t0 = time.time()
    b = synth.make_batch(0, n, 1, fmt=av.Format.MONTGOMERY)
    tg = time.time() - t0
    bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
    t0 = time.time()
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    tp = time.time() - t0
    res = []
    for it in range(4):
        t0 = time.time()
        st = bv.verify_status()
        res.append(time.time() - t0)
    tm = bv.timings()
    print(json.dumps({"n": n, "gen_s": round(tg, 3), "push_s": round(tp, 4), "status": st,
                      "verify_s": [round(x, 4) for x in res], "timings": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()}}), flush=True)
    # bad proof
    s2 = b.s.copy(); s2[n // 3, 0] ^= 1
    bv.clear(); bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s2)
    print("  tampered status", bv.verify_status(), flush=True)
    bv.close()
json.dump(out, open("gpurun_out/microbench.json", "w"))

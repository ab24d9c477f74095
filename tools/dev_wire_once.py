#!/usr/bin/env python3
"""One 2^log2n-proof Bandersnatch batch pushed in WIRE format (avrf_thin_batch_push_compressed) and verified; prints the
wall time of push and verify.  python tools/dev_wire_once.py [log2n] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import ark_vrf_b200 as av
    from ark_vrf_b200 import ops, synth
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = 1 << log2n
    av.load().avrf_init(0)
    b = synth.make_batch(0, n, 1, signers=4096, fmt=av.Format.CANONICAL)
    pk32 = ops.point_compress(0, b.pk)
    r32 = ops.point_compress(0, b.r)
    ios32 = np.concatenate([ops.point_compress(0, b.ios[:128 * n].reshape(-1, 64)).reshape(-1), np.zeros(64, np.uint8)])
    bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
    for _ in range(reps):
        bv.clear()
        t0 = time.perf_counter()
        assert bv.push_compressed(pk32, ios32, b.io_offsets, b.ad_blob, b.ad_offsets, r32, b.s).all()
        t1 = time.perf_counter()
        assert bv.verify_status() == 0
        t2 = time.perf_counter()
        print("push_compressed %.2f ms, verify %.2f ms (2^%d proofs, %d points decoded)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, log2n, 4 * n))


if __name__ == "__main__":
    main()

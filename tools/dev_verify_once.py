#!/usr/bin/env python3
"""One 2^log2n-proof Bandersnatch batch, generated on the GPU, pushed once, verified `reps` times with the whole
path redone (the command the ncu captures of profiles/ run).  python tools/dev_verify_once.py [log2n] [reps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import ark_vrf_b200 as av
    from ark_vrf_b200 import synth
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    av.load().avrf_init(0)
    b = synth.make_batch(0, 1 << log2n, 1, signers=4096, fmt=av.Format.MONTGOMERY)
    bv = av.BatchVerifier(0, av.Format.MONTGOMERY)
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    for _ in range(reps):
        bv.invalidate()
        assert bv.verify_status() == 0
        print(bv.timings())


if __name__ == "__main__":
    main()

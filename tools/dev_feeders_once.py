#!/usr/bin/env python3
"""One pass of the feeder operations at 2^log2n items (Bandersnatch): avrf_vrf_io_many (Elligator2 + GLV output),
avrf_vrf_output (plain scalar multiplication), avrf_points_deserialize - the command the feeder ncu captures run.
python tools/dev_feeders_once.py [log2n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import ark_vrf_b200 as av
    from ark_vrf_b200 import ops, synth
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
    av.load().avrf_init(0)
    blob = np.concatenate([np.arange(n, dtype=np.uint64).view(np.uint8), np.zeros(16, np.uint8)])
    off = (np.arange(n + 1, dtype=np.uint64) * 8).astype(np.uint32)
    sk = np.frombuffer(synth.secret_from_seed(0, bytes(32)).to_bytes(32, "little"), dtype=np.uint8).copy()
    res = ops.vrf_io_many(0, blob, off, sk)
    assert res["ok"].all()
    out = ops.vrf_output(0, sk, res["inputs"])
    assert (out == res["outputs"]).all()
    enc = ops.point_compress(0, res["outputs"])
    pts, ok = ops.points_deserialize(0, enc, kind=1)
    assert ok.all() and (pts == res["outputs"]).all()
    print("feeders ok", n)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Dev aid: e2e throughput with T host threads, one batch handle each (python tools/dev_concurrent.py 20 1,2,4,8)."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import ark_vrf_b200 as av
    from ark_vrf_b200 import synth
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    n = 1 << log2n
    print("host threads available:", len(os.sched_getaffinity(0)))
    b = synth.make_batch(0, n, 1, signers=4096, fmt=av.Format.MONTGOMERY)
    host = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in
            (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)]
    for T in ts:
        hs = [av.BatchVerifier(0, av.Format.MONTGOMERY) for _ in range(T)]

        def worker(h, k):
            for _ in range(k):
                h.clear()
                h.push_many(*host)
                assert h.verify_status() == 0

        def run(k):
            th = [threading.Thread(target=worker, args=(h, k)) for h in hs]
            for t in th:
                t.start()
            for t in th:
                t.join()
        run(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(reps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = hs[0].timings()
        print("T=%2d  %.2f ms/batch  %.1f M proofs/s   (handle 0: hash %.1f ms, push_total %.1f)" % (
            T, dt * 1e3 / (reps * T), n * reps * T / dt / 1e6, tm["host_hash_ms"], tm.get("h2d_ms", 0)))
        for h in hs:
            h.close()


if __name__ == "__main__":
    main()

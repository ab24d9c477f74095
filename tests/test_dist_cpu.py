"""CPU, world_size 2, gloo: the rank plumbing of the sharded verifier (ark_vrf_b200/dist.py) -
gather order, global weight indices, InvalidData precedence - with an oracle-backed stand-in for
the per-GPU engine (the real engine is exercised by the -m gpu tests)."""
import hashlib
import os
import socket

import numpy as np
import pytest

from oracle import pyref as o

R = 1 << 256


class OracleShard:
    """Same methods as ark_vrf_b200.BatchVerifier's sharded API, computed by the oracle."""

    def __init__(self, S, pr, lo, hi):
        self.S = S
        self.items = [o.batch_prepare(S, pr.pk[j], pr.ios[j], pr.ad[j], pr.r[j], pr.s[j]) for j in range(lo, hi)]

    def __len__(self):
        return len(self.items)

    def prepare_device(self):
        return any(e.pk == o.IDENTITY or o.has_identity(e.ios) for e in self.items)

    def cs_stream(self):
        b = b"".join(o.enc_scalar(e.c) + o.enc_scalar(e.s) for e in self.items)
        return np.frombuffer(b, dtype=np.uint8).reshape(-1, 64).copy()

    def partial(self, seed, first):
        S, r = self.S, self.S.r
        acc = o.EXT_ID
        g = 0
        for q, e in enumerate(self.items):
            j = first + q
            blk = hashlib.sha512(seed + (j // 4).to_bytes(8, "little")).digest()
            w = int.from_bytes(blk[16 * (j % 4):16 * (j % 4) + 16], "little")
            wc, ws = w * e.c % r, w * e.s % r
            terms = [(e.r, w), (e.pk, wc)]
            for i, (inp, out) in enumerate(e.ios):
                terms += [(out, wc * e.zs[i + 1] % r), (inp, (-ws * e.zs[i + 1]) % r)]
            g = (g - ws) % r
            for P, k in terms:
                acc = o.ext_add(S, acc, o.ext_mul(S, o.to_ext(P), k))
        acc = o.ext_add(S, acc, o.ext_mul(S, o.to_ext(S.G), g))
        return b"".join((c * R % S.p).to_bytes(32, "little") for c in acc)


def _seed_fn(S):
    return lambda suite, stream: hashlib.sha512(S.suite_id + b"\x50" + bytes(stream)).digest()


def _combine_fn(S):
    def f(suite, blob):
        acc = o.EXT_ID
        ri = pow(R, -1, S.p)
        for k in range(len(blob) // 128):
            pt = tuple(int.from_bytes(blob[128 * k + 32 * i:128 * k + 32 * i + 32], "little") * ri % S.p for i in range(4))
            acc = o.ext_add(S, acc, pt)
        return 0 if o.ext_is_identity(S, acc) else 1
    return f


def _worker(rank, world, port, case, q):
    import torch.distributed as dist
    from ark_vrf_b200 import dist as avdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = o.BANDERSNATCH
    n = 10
    pr = o.synth_proofs(S, n, 1, signers=3)
    if case == "bad_s":
        pr.s[n - 1] = (pr.s[n - 1] + 1) % S.r
    elif case == "identity_pk":
        pr.pk[1] = o.IDENTITY
        pr.s[n - 1] = (pr.s[n - 1] + 1) % S.r
    elif case == "empty":
        n = 0
    lo, hi = avdist.shard_bounds(n, world, rank)
    shard = OracleShard(S, pr, lo, hi)
    st = avdist.sharded_verify(shard, 0, lo, seed_fn=_seed_fn(S), combine_fn=_combine_fn(S))
    expect = o.batch_verify(S, [o.batch_prepare(S, pr.pk[j], pr.ios[j], pr.ad[j], pr.r[j], pr.s[j]) for j in range(n)])
    q.put((rank, st, expect))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,expect", [("valid", 0), ("bad_s", 1), ("identity_pk", 2), ("empty", 0)])
def test_sharded_verify_gloo_world2(case, expect):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, st, exp in res:
        assert st == exp == expect, (rank, st, exp)

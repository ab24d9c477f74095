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


def _oracle_fns(S):
    """ops-shaped stand-ins computed by the oracle (canonical format)."""
    def h2c(suite, blob, off, fmt):
        pts = [o.data_to_point(S, bytes(blob[off[j]:off[j + 1]])) for j in range(len(off) - 1)]
        return np.frombuffer(b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts), dtype=np.uint8).reshape(-1, 64).copy()

    def out(suite, sks, inputs, fmt):
        res = b""
        for k, row in zip(sks, inputs):
            P = (int.from_bytes(bytes(row[:32]), "little"), int.from_bytes(bytes(row[32:]), "little"))
            x, y = o.pt_mul(S, P, int.from_bytes(bytes(k), "little"))
            res += x.to_bytes(32, "little") + y.to_bytes(32, "little")
        return np.frombuffer(res, dtype=np.uint8).reshape(-1, 64).copy()

    def comp(suite, pts, fmt):
        enc = b"".join(o.enc_point(S, (int.from_bytes(bytes(r[:32]), "little"), int.from_bytes(bytes(r[32:]), "little"))) for r in pts)
        return np.frombuffer(enc, dtype=np.uint8).reshape(-1, 32).copy()
    return h2c, out, comp


def _io_worker(rank, world, port, n, q):
    import torch.distributed as dist
    from ark_vrf_b200 import dist as avdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = o.BANDERSNATCH
    sk = o.secret_from_seed(S, bytes(32))
    h2c, out, comp = _oracle_fns(S)
    lo, hi, inp, outp, dg = avdist.sharded_inputs_outputs(0, n, np.frombuffer(sk.to_bytes(32, "little"), dtype=np.uint8), fmt=1,
                                                          h2c_fn=h2c, out_fn=out, compress_fn=comp)
    q.put((rank, lo, hi, inp.tobytes(), outp.tobytes(), dg))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_inputs_outputs_gloo_world2():
    """configs[4] plumbing: the two ranks cover [0, n) exactly once, in order, and agree on a checksum that equals
    the single-process one."""
    import torch.multiprocessing as mp
    from ark_vrf_b200 import dist as avdist
    S = o.BANDERSNATCH
    n = 70                                    # shards of 64 and 6: ragged
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_io_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 64, 64, 70)
    sk = o.secret_from_seed(S, bytes(32))
    h2c, out, comp = _oracle_fns(S)
    lo, hi, inp, outp, dg = avdist.sharded_inputs_outputs(0, n, np.frombuffer(sk.to_bytes(32, "little"), dtype=np.uint8), fmt=1,
                                                          h2c_fn=h2c, out_fn=out, compress_fn=comp)
    assert (lo, hi) == (0, n)
    assert res[0][3] + res[1][3] == inp.tobytes() and res[0][4] + res[1][4] == outp.tobytes()
    assert res[0][5] == res[1][5] == dg != 0
    P0 = o.data_to_point(S, (0).to_bytes(8, "little"))
    assert inp[0].tobytes() == P0[0].to_bytes(32, "little") + P0[1].to_bytes(32, "little")

"""GPU, needs >= 2 devices (skipped otherwise): the sharded verifier over NCCL, one process per GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import ark_vrf_b200 as av
    from ark_vrf_b200 import dist as avdist, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    av._lib.check(av.load().avrf_init(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = 4096 + 64
    lo, hi = avdist.shard_bounds(n, world, rank)
    b = synth.make_batch(0, hi - lo, 1, fmt=av.Format.MONTGOMERY, first=lo)
    res = {}
    for case in ("valid", "bad_s_last_rank", "identity_pk_rank0", "tree_valid", "tree_bad"):
        s, pk = b.s.copy(), b.pk.copy()
        if case in ("bad_s_last_rank", "tree_bad") and rank == world - 1:
            s[-1, 0] ^= 1
        if case == "identity_pk_rank0" and rank == 0:
            pk[0, :] = 0
            pk[0, 32:] = np.frombuffer(((1 << 256) % 52435875175126190479447740508185965837690552500527637822603658699938581184513).to_bytes(32, "little"), dtype=np.uint8)
        bv = av.BatchVerifier(0, av.Format.MONTGOMERY, eager_seed=False)
        bv.push_many(pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, s)
        res[case] = avdist.sharded_verify(bv, 0, lo, device=dev, weights="tree" if case.startswith("tree") else "reference")
    # configs[4]: hash-to-curve + output sharded over the ranks, checksum agreed by all-reduce
    sk = np.frombuffer((12345).to_bytes(32, "little"), dtype=np.uint8)
    lo2, hi2, inp, outp, dg = avdist.sharded_inputs_outputs(0, 10000, sk, fmt=int(av.Format.CANONICAL), device=dev)
    res["io"] = (lo2, hi2, inp.shape[0], int(dg))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_verify_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ios = {rank: res.pop("io") for rank, res in out}
    assert {ios[0][:3], ios[1][:3]} == {(0, 5024, 5024), (5024, 10000, 4976)}
    assert ios[0][3] == ios[1][3] != 0
    for rank, res in out:
        assert res == {"valid": 0, "bad_s_last_rank": 1, "identity_pk_rank0": 2, "tree_valid": 0, "tree_bad": 1}, (rank, res)

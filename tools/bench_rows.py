#!/usr/bin/env python3
"""Throughput of the operations either side of the batch-verification path (SURVEY.md 8a rows a10/a11/a15 and
8f rows 1-4) on one B200, through the public Python API over the C ABI with HOST buffers (H2D/D2H inside), with the
C restatement of the reference timed beside it on the host cores where it implements the operation.

    python tools/bench_rows.py [log2n]        ->  one JSON object per line, collected in profiles/r1_rows.jsonl
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    import ark_vrf_b200 as av
    from ark_vrf_b200 import ops, synth, pedersen as ped
    from oracle import corc, pyref as o
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << log2n
    sid, S = 0, o.SUITES[0]
    F = av.Format.MONTGOMERY
    out = []

    def row(name, ref, units, secs, unit="items/s", cpu=None, note=""):
        r = {"row": name, "reference": ref, "n": units, "seconds": round(secs, 5), "value": units / secs, "unit": unit}
        if cpu is not None:
            r["cpu_port_single_thread"] = cpu
        if note:
            r["note"] = note
        out.append(r)
        print(json.dumps(r), flush=True)

    b = synth.make_batch(sid, n, 1, signers=4096, fmt=F)
    # -- a10 hash-to-curve (Input::new, lib.rs:500-502) ---------------------------------------------------
    msgs = np.zeros((n, 12), dtype=np.uint8)
    msgs[:, :8] = np.arange(n, dtype=np.uint64).view(np.uint8).reshape(n, 8)
    moff = (np.arange(n + 1, dtype=np.uint64) * 12).astype(np.uint32)
    blob = np.concatenate([msgs.reshape(-1), np.zeros(16, np.uint8)])
    ops.hash_to_curve(sid, blob, moff, F)
    t = best(lambda: ops.hash_to_curve(sid, blob, moff, F))
    k = 2000
    t0 = time.perf_counter()
    for j in range(k):
        corc.hash_to_curve(sid, bytes(msgs[j]))
    cpu_h2c = k / (time.perf_counter() - t0)
    row("hash_to_curve (Elligator2-XMD)", "src/lib.rs:500-502, src/utils/hash_to_curve.rs:66-100", n, t, "points/s", cpu_h2c)
    # -- a11 Secret::output -------------------------------------------------------------------------------
    sk = np.ascontiguousarray(np.frombuffer(bytes(range(1, 33)), dtype=np.uint8).reshape(1, 32).repeat(n, 0))
    sk[:, 31] &= 0x0f
    inputs = np.ascontiguousarray(b.ios[:, :64])
    ops.vrf_output(sid, sk, inputs, F)
    t = best(lambda: ops.vrf_output(sid, sk, inputs, F))
    k = 1000
    p64 = None
    t0 = time.perf_counter()
    for j in range(k):
        corc.scalar_mul(sid, bytes([j & 255] + [7] * 30 + [0]), p64)
    cpu_mul = k / (time.perf_counter() - t0)
    row("vrf_output (Secret::output)", "src/lib.rs:391-393", n, t, "points/s", cpu_mul)
    # -- a15 / 8f-4 bulk proving ----------------------------------------------------------------------------
    skk = np.ascontiguousarray(sk)
    pk = ops.public_keys(sid, skk[:4096], F)
    pkn = np.ascontiguousarray(pk[np.arange(n) % 4096])
    outputs = ops.vrf_output(sid, skk, inputs, F)
    ios = np.ascontiguousarray(np.concatenate([inputs, outputs], axis=1))
    t = best(lambda: ops.thin_prove_many(sid, skk, pkn, ios, b.io_offsets, b.ad_blob, b.ad_offsets, F))
    row("thin_prove_many (Prover::prove)", "src/thin.rs:111-129", n, t, "proofs/s")
    pr, ps = ops.thin_prove_many(sid, skk, pkn, ios, b.io_offsets, b.ad_blob, b.ad_offsets, F)
    chk = av.BatchVerifier(sid, F)
    chk.push_many(pkn, ios, b.io_offsets, b.ad_blob, b.ad_offsets, pr, ps)
    assert chk.verify_status() == 0                      # the proofs just made verify
    chk.close()
    # -- a6 compress / point_to_hash ------------------------------------------------------------------------
    t = best(lambda: ops.point_compress(sid, b.pk, F))
    row("point_compress (CanonicalSerialize)", "src/utils/transcript.rs:48-50", n, t, "points/s")
    t = best(lambda: ops.point_to_hash(sid, outputs, F))
    row("point_to_hash (Output::hash)", "src/utils/common.rs:290-305", n, t, "points/s")
    # -- 8f-1 wire-format ingest ----------------------------------------------------------------------------
    enc = ops.point_compress(sid, b.pk, F)
    pts, ok = ops.points_deserialize(sid, enc, 1, F)
    assert ok.all() and (pts == b.pk).all()
    t = best(lambda: ops.points_deserialize(sid, enc, 1, F))
    row("points_deserialize (CanonicalDeserialize + subgroup check)", "src/lib.rs:410-433,471-494,552-575", n, t, "points/s")
    # -- 8f-2 per-proof verdicts ----------------------------------------------------------------------------
    bv = av.BatchVerifier(sid, F)
    bv.push_many(b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)
    st = bv.verify_each()
    assert (st == 0).all()
    t = best(lambda: bv.verify_each())
    k = 300
    bc = synth.make_batch(sid, k, 1, fmt=av.Format.CANONICAL)
    t0 = time.perf_counter()
    for j in range(k):
        corc.thin_verify(sid, bytes(bc.pk[j]), bytes(bc.ios[j]), b"ad-%d" % j, bytes(bc.r[j]), bytes(bc.s[j]))
    cpu_v = k / (time.perf_counter() - t0)
    row("thin_batch_verify_each (Verifier::verify per proof)", "src/thin.rs:131-165", n, t, "proofs/s", cpu_v,
        "inputs resident from the push")
    bv.close()
    # -- 8f-3 Pedersen batch verifier (oracle-made proofs, tiled to the batch size) ---------------------------
    base = 256
    sk0 = o.secret_from_seed(S, bytes(32))
    cols = {k_: [] for k_ in ("ios", "ad", "pk_com", "r", "ok", "s", "sb")}
    Rm = 1 << 256
    mont_p = lambda P: b"".join(((c * Rm) % S.p).to_bytes(32, "little") for c in P)
    mont_s = lambda x: ((x * Rm) % S.r).to_bytes(32, "little")
    for j in range(base):
        inp = o.data_to_point(S, j.to_bytes(8, "little"))
        io = (inp, o.pt_mul(S, inp, sk0))
        ad = b"ad-%d" % j
        pf, _ = o.pedersen_prove(S, sk0, [io], ad)
        cols["ios"].append(mont_p(io[0]) + mont_p(io[1])); cols["ad"].append(ad)
        cols["pk_com"].append(mont_p(pf.pk_com)); cols["r"].append(mont_p(pf.r)); cols["ok"].append(mont_p(pf.ok))
        cols["s"].append(mont_s(pf.s)); cols["sb"].append(mont_s(pf.sb))
    rep = n // base
    arr = lambda key, w: np.ascontiguousarray(np.tile(np.frombuffer(b"".join(cols[key]), dtype=np.uint8).reshape(base, w), (rep, 1)))
    ad_lens = np.tile(np.array([len(a) for a in cols["ad"]], dtype=np.uint64), rep)
    ad_off = np.zeros(n + 1, dtype=np.uint32)
    ad_off[1:] = np.cumsum(ad_lens).astype(np.uint32)
    ad_blob = np.frombuffer(b"".join(cols["ad"]) * rep + bytes(16), dtype=np.uint8).copy()
    io_off = np.arange(n + 1, dtype=np.uint32)
    pv = ped.BatchVerifier(sid, F)
    args = (arr("ios", 128), io_off, ad_blob, ad_off, arr("pk_com", 64), arr("r", 64), arr("ok", 64), arr("s", 32), arr("sb", 32))
    pv.push_many(*args)
    assert pv.verify_status() == 0

    import torch
    pargs = [torch.from_numpy(a).pin_memory() for a in args]

    def ped_step():
        pv.clear()
        pv.push_many(*pargs)
        assert pv.verify_status() == 0
    ped_step()
    t = best(ped_step, 3)
    row("pedersen batch verify (5N+2-point MSM)", "src/pedersen.rs:322-427", n, t, "proofs/s",
        note="clear + push from pinned host memory + verify; 256 oracle-made proofs tiled to the batch size; serial host SHA-512 over 96 B per proof inside")
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "rows.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

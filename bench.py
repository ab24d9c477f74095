#!/usr/bin/env python3
"""bench.py - Bandersnatch thin-VRF batch-verified proofs/sec on B200 (BASELINE.json metric).

A step = one pass of the hot path (BatchVerifier push/prepare + verify, reference
src/thin.rs:209-325) over ONE batch of 2^20 synthetic proofs.

  value : proofs/s with the batch already resident in HBM (prepare + seed + MSM every step)
  e2e   : proofs/s through the public API with pinned HOST buffers (H2D of the batch and D2H of
          the (c,s) stream / verdict inside the timed region)
  roofline     : the dominant kernel (k_accumulate, mixed additions) against the integer-multiply
                 peak measured live by a dependency-free IMAD.WIDE.U32 microbenchmark
  cpu_baseline : the C oracle (restatement of the reference algorithm, kind "port") on the
                 box's host cores, bounded sample

--impl reference : the CPU restatement timed alone (the reference is Rust on un-vendored crates and
                   cannot be built in this image).
--gpus N (torchrun): the batch is sharded over N ranks (ark_vrf_b200/dist.py), strong scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# canonical algorithmic work: wide MACs (32x32->64 multiply-accumulate)
MM_MACS = 136                      # one 8x32-limb Montgomery multiplication
ADDS_PER_PROOF_M1 = 57             # 9 (128-bit weight) + 3*16 (full scalars) bucket additions
CANON_MM_PER_ADD = 7
MACS_PER_PROOF_M1 = 56168          # whole path, M = 1


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def cpu_reference_arm(args, rank, world):
    """--impl reference: the CPU restatement of thin::BatchVerifier (oracle port), all host threads."""
    if rank != 0:
        return
    from oracle import corc
    T = host_threads()
    n = 1 << args.ref_log2n
    t0 = time.perf_counter()
    arrs = corc.synth_batch(0, n, 1, signers=4096, nthreads=T)
    gen_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        st, _, _ = corc.thin_batch_verify(0, *arrs, nthreads=T)
        assert st == 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, _, _ = corc.thin_batch_verify(0, *arrs, nthreads=T)
        assert st == 0
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    line = {
        "impl": "reference", "metric": "bandersnatch_thin_vrf_batch_verified_proofs_per_sec", "value": v,
        "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "Bandersnatch thin-VRF batch verify, 2^20 synthetic proofs (M=1, 4096 signers)",
                   "suite": "Bandersnatch-SHA512-ELL2-v1", "batch": 1 << 20, "io_pairs": 1},
        "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": T, "kind": "port",
                         "sample": f"2^{args.ref_log2n} proofs of the same synthetic set per step (prepare+verify), "
                                   f"C restatement of the reference algorithm, {T} threads; generation {gen_s:.1f}s untimed"},
        "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--ref-log2n", type=int, default=int(os.environ.get("AVRF_REF_LOG2N", "17")))
    ap.add_argument("--cpu-sample-log2n", type=int, default=18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--concurrent", type=int, default=16,
                    help="batch-server workers per GPU (one batch handle each) of the concurrent-serving leg; 1 disables it")
    ap.add_argument("--hashers", type=int, default=-1,
                    help="shared multi-buffer SHA-512 threads per GPU for that leg (0: one hashing core per worker; "
                         "-1: 0 when the rank has a core per worker, else up to 3 with 8 workers each)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        cpu_reference_arm(args, rank, world)
        return

    import torch
    import ark_vrf_b200 as av
    from ark_vrf_b200 import dist as avdist, ops, synth
    lib = av.load()
    av._lib.check(lib.avrf_init(local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    n = 1 << args.log2n
    lo, hi = avdist.shard_bounds(n, world, rank)
    nl = hi - lo
    # ---- synthetic workload: this rank's shard, generated on its GPU -------------------------
    t0 = time.perf_counter()
    b = synth.make_batch(0, nl, 1, signers=4096, fmt=av.Format.MONTGOMERY, first=lo)
    gen_s = time.perf_counter() - t0

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    host = [pin(x) for x in (b.pk, b.ios, b.io_offsets, b.ad_blob, b.ad_offsets, b.r, b.s)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    # ---- integer-multiply peak, measured live (roofline denominator) ---------------------------
    peak_wide = max(ops.microbench(0, 4096)[0] for _ in range(3))
    peak_carry = max(ops.microbench(3, 4096)[0] for _ in range(3))

    bv = av.BatchVerifier(0, av.Format.MONTGOMERY, eager_seed=(world == 1))
    bv.push_many(*host)
    stream = torch.cuda.ExternalStream(bv.stream, device=dev)     # the stream this handle's kernels run on

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    sharded_t = []

    def step_resident():
        bv.invalidate()                      # redo prepare too: the whole path, inputs resident in HBM
        if world == 1:
            st = bv.verify_status()
        else:
            td = {}
            st = avdist.sharded_verify(bv, 0, lo, device=dev, timings=td)
            sharded_t.append(td)
        assert st == 0, st
        return bv.timings()

    def step_e2e():
        bv.clear()
        bv.push_many(*host)                  # H2D from pinned host memory
        if world == 1:
            st = bv.verify_status()
        else:
            st = avdist.sharded_verify(bv, 0, lo, device=dev)
        assert st == 0, st

    def timed(fn, steps, others=()):
        """K steps bracketed by events on the handle's stream; `others` = further handles whose streams the
        closing event must wait for (the multi-handle legs)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = []
        e0.record(stream)
        t0 = time.perf_counter()
        if steps:
            for _ in range(steps):
                acc.append(fn())
        else:
            steps = fn()                     # the leg runs its own loop and returns how many steps it did
        for h in others:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.ExternalStream(h.stream, device=dev))
            stream.wait_event(ev)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, acc

    with ClockSampler(local_rank) as clk:          # nvidia-smi takes ~1 s to start: begin before warm-up
        for _ in range(args.warmup):
            step_resident()
        clk.rows.clear()
        ms_step, tms = timed(step_resident, args.steps)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # pipelined serving: two handles, the next batch's push (host SHA-512 + its prepare kernels on their own
    # high-priority stream) overlaps the previous batch's MSM (verify_async / verify_wait).  Extra figure only.
    ms_pipe = None
    if world == 1:
        bv2 = av.BatchVerifier(0, av.Format.MONTGOMERY)
        hs = [bv, bv2]
        state = {"i": 0, "pending": None}

        def step_pipe():
            h = hs[state["i"] % 2]
            state["i"] += 1
            h.clear()
            h.push_many(*host)
            if state["pending"] is not None:
                assert state["pending"].verify_wait() == 0
            h.verify_async()
            state["pending"] = h
        for _ in range(3):
            step_pipe()
        ms_pipe, _ = timed(step_pipe, args.steps, others=[bv2])
        assert state["pending"].verify_wait() == 0
        state["pending"] = None
        bv2.close()

    # concurrent serving: T host threads per GPU, one handle each (own CUDA streams), every thread doing whole e2e
    # steps on WHOLE 2^log2n-proof batches (clear, push from pinned host memory, verify).  The serial SHA-512 of each
    # batch runs on its own core, the kernels of the handles share the GPU.  With N > 1 every rank serves its own
    # batches (no collective: batches are independent), so this is the weak-scaling throughput of the box.
    # Extra figure only; `value` and `e2e` stay one batch at a time (sharded over the ranks when N > 1).
    ms_conc, n_conc, conc_steps = None, 0, 0
    cores_per_rank = max(1, host_threads() // world)
    n_hash = args.hashers
    if n_hash < 0:
        n_hash = 0 if cores_per_rank >= args.concurrent else max(1, min(3, cores_per_rank - 1))
    t_per_rank = 8 * n_hash if (n_hash and args.hashers < 0) else args.concurrent
    if t_per_rank > 1:
        n_conc = t_per_rank
        if world == 1:
            host_c = host
        else:
            bf = synth.make_batch(0, n, 1, signers=4096, fmt=av.Format.MONTGOMERY, first=rank * n)
            host_c = [pin(x) for x in (bf.pk, bf.ios, bf.io_offsets, bf.ad_blob, bf.ad_offsets, bf.r, bf.s)]
        srv = av.BatchServer(0, av.Format.MONTGOMERY, workers=n_conc, hashers=n_hash)     # native worker pool (avrf_server_*)
        per = max(3, -(-args.steps // n_conc))

        def run_conc(k):
            tickets = [srv.submit(*host_c) for _ in range(k * n_conc)]
            for t in tickets:
                assert srv.wait(t) == 0
            return len(tickets)
        run_conc(1)                           # warm-up: allocations of the workers' handles
        ms_conc, _ = timed(lambda: run_conc(per), 0)      # every verdict is back before the closing event
        conc_steps = per * n_conc * world
        srv.close()
    if world == 1:
        bv.clear()
        bv.push_many(*host)

    # opt-in tree-hashed weights (not the reference's transcript bytes; reported beside, never as `value`)
    bv.set_weights_mode(1)

    def step_tree():
        bv.invalidate()
        st = bv.verify_status() if world == 1 else avdist.sharded_verify(bv, 0, lo, device=dev, weights="tree")
        assert st == 0, st
        return bv.timings()
    for _ in range(2):
        step_tree()
    ms_tree, _ = timed(step_tree, args.steps)
    bv.set_weights_mode(0)

    # bench fidelity: the reference's own bench signs every proof with ONE key (benches/thin.rs:46);
    # the headline uses 4096 signers so that no same-key shortcut can be mistaken for MSM speed.  Same path, K = 1.
    ms_k1 = None
    if world == 1:
        b1 = synth.make_batch(0, nl, 1, signers=1, fmt=av.Format.MONTGOMERY)
        bv.clear()
        bv.push_many(b1.pk, b1.ios, b1.io_offsets, b1.ad_blob, b1.ad_offsets, b1.r, b1.s)

        def step_k1():
            bv.invalidate()
            assert bv.verify_status() == 0
        for _ in range(2):
            step_k1()
        ms_k1, _ = timed(step_k1, max(2, args.steps // 2))

    # ---- per-kernel figures (CUDA events on the launch stream, averaged over the timed steps) --
    def avg(key):
        return float(np.mean([t[key] for t in tms]))
    acc_ms = avg("accumulate_ms")
    entries = float(np.mean([t["n_entries"] for t in tms]))
    launches = int(sum(t["kernel_launches"] for t in tms))
    canon_macs = entries * CANON_MM_PER_ADD * MM_MACS          # per launch, this rank
    achieved = canon_macs / (acc_ms * 1e-3) / 1e12
    executed = entries * 8 * 112 / (acc_ms * 1e-3) / 1e12     # 8 mm x 112 wide MACs (BLS12-381 Fr reduction shortcut)
    phases = {k: round(avg(k), 3) for k in ("prepare_ms", "host_hash_ms", "scalars_ms", "sort_ms", "accumulate_ms", "reduce_ms")}
    npts = float(np.mean([t["n_points"] for t in tms]))
    sort_bytes = entries * 4 + npts * (32 + 64) + 4 * (1 << 19) * 6        # entries written, digits+ranks read, bin arrays
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import corc
        T = host_threads()
        ns = min(nl, 1 << args.cpu_sample_log2n)
        # the oracle takes canonical integers: regenerate the sample in canonical form on the GPU
        bc = synth.make_batch(0, ns, 1, signers=4096, fmt=av.Format.CANONICAL, first=lo)
        st, tm, _ = corc.thin_batch_verify(0, bc.pk, bc.ios, bc.io_offsets, bc.ad_blob, bc.ad_offsets, bc.r, bc.s, nthreads=T)
        assert st == 0
        st1, tm1, _ = corc.thin_batch_verify(0, bc.pk[:4096], bc.ios[:4096], bc.io_offsets[:4097], bc.ad_blob,
                                             bc.ad_offsets[:4097], bc.r[:4096], bc.s[:4096], nthreads=1)
        cpu_baseline = {"value": ns / sum(tm), "unit": "proofs/s", "cores": T, "kind": "port",
                        "sample": f"first 2^{int(np.log2(ns))} proofs of the same workload, prepare+verify "
                                  f"({tm[0]:.2f}s+{tm[1]:.2f}s), C restatement of the reference algorithm (oracle/avrf_oracle.c)",
                        "single_thread_proofs_per_s_n4096": 4096 / sum(tm1)}

    if rank == 0:
        value = n / (ms_step * 1e-3)
        e2e = n / (ms_e2e * 1e-3)
        line = {
            "metric": "bandersnatch_thin_vrf_batch_verified_proofs_per_sec", "value": value, "unit": "proofs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": "Bandersnatch thin-VRF batch verify, 2^%d synthetic proofs (M=1, 4096 signers), "
                                   "BASELINE.json configs[1]" % args.log2n,
                       "suite": "Bandersnatch-SHA512-ELL2-v1", "batch": n, "io_pairs": 1, "weights": "reference (serial SHA-512 on host)",
                       "arithmetic": "256-bit Montgomery fields in 8 x u32 limbs (IMAD.WIDE.U32), SHA-512 in u64",
                       "l2": "inputs+working set (~1.5 GB per 2^20 proofs) exceed the 126 MB L2; no explicit flush",
                       "sharding": "contiguous proof shards, one NCCL all-gather of (c,s) + one of 130-byte partials" if world > 1 else "single GPU"},
            "e2e": {"value": e2e, "unit": "proofs/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes) * 1,
                    "d2h_bytes_per_step": int(64 * nl + 16)},
            "e2e_pipelined": None if ms_pipe is None else {
                "value": n / (ms_pipe * 1e-3), "unit": "proofs/s", "ms_per_step": ms_pipe,
                "note": "two batch handles in flight (avrf_thin_batch_verify_async/_wait): push of batch i+1 overlaps the MSM of batch i"},
            "e2e_concurrent": None if ms_conc is None else {
                "value": world * n / (ms_conc * 1e-3), "unit": "proofs/s", "ms_per_batch_per_gpu": ms_conc,
                "handles_per_gpu": n_conc, "mb_sha512_threads_per_gpu": n_hash, "host_cores_per_gpu": cores_per_rank,
                "batches": conc_steps, "scaling": "weak",
                "note": "avrf_server: %d worker threads per GPU, one batch handle (own CUDA streams) each, every step a whole e2e step on a whole "
                        "2^%d-proof batch (clear, push from pinned host memory, verify): the serial SHA-512 of each batch runs on its "
                        "own core (or, with mb_sha512_threads_per_gpu > 0, eight batches per shared AVX-512 multi-buffer hashing thread), "
                        "the kernels share the GPU; ranks serve independent batches (no collective)" % (n_conc, args.log2n)},
            "gpu_launches": launches,
            "roofline": {"bound": "imad", "kernel": "k_accumulate", "achieved": achieved, "peak": peak_wide / 1e12,
                         "unit": "T wide-MAC/s", "frac": achieved / (peak_wide / 1e12), "traffic": 10.77e9 * (entries / 59243748.0),
                         "note": "achieved = canonical 7 mm x 136 wide MACs per bucket addition (SURVEY.md 8d) x additions per launch "
                                 "/ mean CUDA-event duration of the kernel; peak = dependency-free IMAD.WIDE.U32 microbenchmark run in this process; traffic = dram read+write bytes of the ncu --set full capture of the final kernel (profiles/r1_SUMMARY.md, prof_accum_r1_final) scaled by additions",
                         "executed": executed, "peak_carry_chain": peak_carry / 1e12,
                         "frac_executed_vs_carry_chain_peak": executed / (peak_carry / 1e12),
                         "additions_per_launch": entries, "kernel_ms": acc_ms,
                         "hbm_sort": {"bound": "hbm", "kernels": "k_scan_*+k_scatter", "achieved": sort_bytes / (avg("sort_ms") * 1e-3) / 1e9,
                                      "peak": hbm_peak, "unit": "GB/s", "frac": sort_bytes / (avg("sort_ms") * 1e-3) / 1e9 / hbm_peak}},
            "single_signer": None if ms_k1 is None else {
                "value": n / (ms_k1 * 1e-3), "unit": "proofs/s", "ms_per_step": ms_k1,
                "note": "same path with every proof signed by one key, as the reference's bench does (benches/thin.rs:46); "
                        "the engine takes no same-key shortcut, so this equals `value`"},
            "alt_tree_weights": {"value": n / (ms_tree * 1e-3), "unit": "proofs/s", "ms_per_step": ms_tree,
                                 "note": "AVRF_WEIGHTS_TREE (opt-in): batch seed from GPU-computed leaf digests instead of the "
                                         "reference's serial SHA-512; same verdicts, different internal weights"},
            "phases_ms": phases,
            "sharded_host_phases_ms": None if not sharded_t else {
                k.replace("_s", "_ms"): round(1e3 * float(np.mean([t[k] for t in sharded_t[-args.steps:]])), 3)
                for k in ("prepare_s", "gather_s", "hash_s", "partial_s", "gather2_s", "combine_s")},
            "sharded_note": None if world == 1 else (
                "one 2^%d-proof batch sharded over %d GPUs: rank 0 wall-clock per step - prepare (transcript kernels), gather "
                "(NCCL all-gather of the (c,s) streams + the reference's SERIAL SHA-512 over all of them, thin.rs:273-279, on one "
                "host core of every rank), partial (this shard's MSM), gather2 + combine (130-byte partials).  The serial hash "
                "does not shard, hence the Amdahl-bound `value`; `e2e_concurrent` is the box's throughput on whole batches"
                % (args.log2n, world)),
            "gpu_phases": {"ms": round(sum(phases[k] for k in ("prepare_ms", "scalars_ms", "sort_ms", "accumulate_ms", "reduce_ms")), 3),
                           "note": "device time of rank 0 per step (prepare + scalars + sort + accumulate + reduce): the part of the "
                                   "step that shards across GPUs; the host SHA-512 of the batch transcript (reference src/thin.rs:273-279) does not"},
            "cpu_baseline": cpu_baseline,
            "clocks": clk.summary(),
            "workload_generation_s": round(gen_s, 2),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Headline benchmark: G1 scalar-multiplication STARK proofs per second (BASELINE.json configs[1]).

One "step" = one proof of G1ExpStark with 128 independent scalar multiplications (2^16 rows x 1676
columns, default StarkConfig): trace generation (K1) + prove (K2-K6) through the C ABI.
  value : inputs already resident in HBM (sbn_trace_generate from a device buffer), proofs/s over all ranks
  e2e   : host buffers in, proof bytes out (pinned host inputs, H2D/D2H inside the timed region)
Multi-GPU: proofs are independent -> one process per GPU, no data-path collective (weak scaling).
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port") on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

SAMPLE_SHIFT = 3
WARP_INSTR_PER_PERM = 766.0      # ncu, k_leaf_hash of this build: 21.06 G warp instructions / 27.5 M permutations (profiles/r01_leaf_hash_ncu_summary_final.txt)
FMAHEAVY_BUSY_NCU = 0.817        # sm__pipe_fmaheavy_cycles_active of the same capture
# --air selects the AIR; the headline (BASELINE.json configs[1], default) is g1.  (class, oracle id, num_io, input generator, metric, workload)
AIRS = {
    "g1": ("G1ExpStark", 2, 128, "g1_exp_ios", "G1 scalar-mul STARK proofs/sec",
           "G1ExpStark num_io=128: 128 independent BN254 G1 scalar multiplications per proof, 2^16 rows x 1676 columns, StarkConfig::standard_fast_config"),
    "g2": ("G2ExpStark", 3, 128, "g2_exp_ios", "G2 scalar-mul STARK proofs/sec",
           "G2ExpStark num_io=128: 128 independent BN254 G2 scalar multiplications per proof, 2^16 rows x 2822 columns, StarkConfig::standard_fast_config"),
    "fq12": ("Fq12ExpStark", 4, 16, "fq12_exp_ios", "Fq12-exp STARK proofs/sec",
             "Fq12ExpStark num_io=16: 16 independent BN254 Fq12 exponentiations (254-bit exponent) per proof, 2^13 rows x 9802 columns, StarkConfig::standard_fast_config"),
    "fq": ("FqExpStark", 1, 128, "fq_exp_ios", "Fq-exp STARK proofs/sec",
           "FqExpStark num_io=128: 128 independent BN254 Fq exponentiations per proof, 2^16 rows x 960 columns, StarkConfig::standard_fast_config"),
}
AIR = "g1"
NUM_IO = 128
WORKLOAD = AIRS["g1"][5]
METRIC = AIRS["g1"][4]


def select_air(name):
    global AIR, NUM_IO, WORKLOAD, METRIC
    AIR, NUM_IO, WORKLOAD, METRIC = name, AIRS[name][2], AIRS[name][5], AIRS[name][4]


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append([x.strip() for x in out])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 7 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 7 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_sample(orc, ios):
    air = orc.Air(AIRS[AIR][1], NUM_IO)
    t0 = time.perf_counter()
    est = orc.time_sample(air, ios, SAMPLE_SHIFT)
    wall = time.perf_counter() - t0
    full_ms = sum(est.values())
    return full_ms, est, wall


def run_reference(args):
    """CPU arm: the oracle (port of the reference algorithm) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core (rank 0 alone runs it)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    orc = entry.load_oracle()
    orc.set_threads(os.cpu_count())
    sbn = entry.load_package()
    ios = getattr(sbn.synthetic, AIRS[AIR][3])(NUM_IO)
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_sample(orc, ios)
    t0 = time.perf_counter()
    fulls = []
    for _ in range(args.steps):
        full_ms, est, _ = cpu_sample(orc, ios)
        fulls.append(full_ms)
    wall = time.perf_counter() - t0
    ms = statistics.mean(fulls)
    value = 1000.0 / ms
    sample = ("per step: every heavy phase of trace generation + prove on 1/%d of its columns / instances / LDE points, scaled x%d; "
              "FRI tail in full (oracle/sample.hpp); est. phases ms=%s" % (1 << SAMPLE_SHIFT, 1 << SAMPLE_SHIFT, {k: round(v, 1) for k, v in est.items()}))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks field)",
            "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--air", default="g1", choices=sorted(AIRS))
    ap.add_argument("--num-io", type=int, default=0, help="instances per proof (power of two; default: the AIR's headline size)")
    ap.add_argument("--no-intra-proof", action="store_true", help="skip the sharded single-proof latency measurement at N > 1")
    ap.add_argument("--inflight", type=int, default=6, help="independent proofs in flight per GPU (one context + CUDA stream each)")
    args = ap.parse_args()
    select_air(args.air)
    if args.num_io:
        global NUM_IO, WORKLOAD
        WORKLOAD = WORKLOAD.replace("num_io=%d:" % NUM_IO, "num_io=%d (non-default size; row / column counts in this text are those of the default):" % args.num_io)
        NUM_IO = args.num_io
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sbn = entry.load_package()
    stream = torch.cuda.current_stream()
    from starky_bn254_b200 import sharding
    # Proofs are independent, so each GPU keeps `inflight` of them going: one context (= one CUDA stream + caching
    # allocator) and one host thread per slot.  The serial sections of one proof (exponentiation chain, lookup walk,
    # host-side Fiat-Shamir) then overlap with the wide kernels of another.  Slot 0 uses torch's current stream.
    nslots = max(1, args.inflight)
    ctxs = [sbn.Context(local, stream.cuda_stream if k == 0 else None) for k in range(nslots)]
    starks = [getattr(sbn, AIRS[AIR][0])(NUM_IO, c) for c in ctxs]
    ctx, stark = ctxs[0], starks[0]
    cfg = stark.config()
    syn = sbn.synthetic
    gen_ios = getattr(syn, AIRS[AIR][3])
    out_off = stark.io_size - 8 * stark.result_words

    # distinct synthetic batches per step and per rank; host copies pinned, device copies resident
    nb = args.steps + args.warmup
    host_ios = []
    for b in range(min(nb, 4)):   # 4 distinct batches, reused round-robin (input generation is host big-int work)
        raw = gen_ios(NUM_IO, seed=0x5EED0001 + 1000 * rank + b)
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
        host_ios.append(t)
    dev_ios = [t.cuda(non_blocking=True) for t in host_ios]
    torch.cuda.synchronize()

    host_raw = [bytes(t.numpy().tobytes()) for t in host_ios]

    def step_resident(i, slot=0):
        st = starks[slot]
        tr = st.generate_trace_device(dev_ios[i % len(dev_ios)].data_ptr())
        res = tr.results()
        ios = syn.fill_outputs(host_raw[i % len(host_raw)], res, st.io_size, out_off)
        pi = st.generate_public_inputs(ios)
        p = sbn.prove(st, cfg, tr, pi)
        tr.free()
        return p

    def step_e2e(i, slot=0):
        st = starks[slot]
        h = host_ios[i % len(host_ios)]
        tr = st.generate_trace_ptr(h.data_ptr(), h.numel())
        res = tr.results()
        ios = syn.fill_outputs(host_raw[i % len(host_raw)], res, st.io_size, out_off)
        pi = st.generate_public_inputs(ios)
        p = sbn.prove(st, cfg, tr, pi)
        tr.free()
        return p.to_bytes()

    def run_steps(fn, first, count):
        """`count` steps starting at index `first`, round-robin over the slots; returns the last result of slot 0."""
        if nslots == 1:
            out = None
            for i in range(count):
                out = fn(first + i, 0)
            return out
        results = [None] * nslots
        errors = []

        nxt = iter(range(count))
        lock = threading.Lock()

        def work(slot):   # slots pull the next step as they finish one (no tail of idle slots when count % nslots != 0)
            try:
                while True:
                    with lock:
                        i = next(nxt, None)
                    if i is None:
                        return
                    results[slot] = fn(first + i, slot)
            except Exception as e:   # noqa: BLE001
                errors.append(e)
        threads = [threading.Thread(target=work, args=(k,)) for k in range(nslots)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return next(r for r in results if r is not None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(step_resident, 0, max(args.warmup, nslots))
    # ---- timed region: device-resident inputs, `inflight` proofs overlapped ----
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sum(c.launch_count for c in ctxs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_steps(step_resident, args.warmup, args.steps)   # every slot's last call has synchronised its stream
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = sum(c.launch_count for c in ctxs) - launches0
    # ---- per-kernel CUDA-event timing: the same steps, one proof at a time on slot 0's stream (overlap would blur it) ----
    ctx.kernel_timing(True)
    proof = None
    ksteps = min(args.steps, 3)
    k0 = torch.cuda.Event(enable_timing=True); k1 = torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for i in range(ksteps):
        proof = step_resident(args.warmup + i, 0)
    k1.record(stream)
    torch.cuda.synchronize()
    serial_ms_per_step = k0.elapsed_time(k1) / ksteps
    kstats = ctx.kernel_stats()
    ctx.kernel_timing(False)
    phases = proof.timings
    # ---- end-to-end: pinned host inputs -> proof bytes on the host ----
    run_steps(step_e2e, 0, nslots)
    barrier()
    t0 = time.perf_counter()
    last_proof_bytes = run_steps(step_e2e, args.warmup, args.steps)
    nbytes = len(last_proof_bytes)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag.set()
    sampler.join()

    # ---- latency mode: ONE proof computed by all ranks together (sbn_prove_sharded); the exchanges (cap digests,
    # quotient values, opened rows) are NCCL all-gathers.  Reported beside the throughput number, not instead of it. ----
    intra = None
    if world > 1 and world in (2, 4, 8, 16) and not args.no_intra_proof:
        try:   # the headline numbers above must survive a failure of this extra measurement
            ag = sharding.dist_allgather(device=torch.device("cuda", local))
            agd = sharding.dist_allgather_device(torch.device("cuda", local))
            raw0 = gen_ios(NUM_IO, seed=0x5EED0001)          # the same inputs on every rank: the trace is replicated

            sharded_phases = {}

            def step_sharded(single=False):
                tr = stark.generate_trace(raw0)
                ios = syn.fill_outputs(raw0, tr.results(), stark.io_size, out_off)
                pi = stark.generate_public_inputs(ios)
                p = sbn.prove(stark, cfg, tr, pi) if single else sbn.prove_sharded(stark, cfg, tr, pi, rank, world, ag, allgather_device=agd)
                tr.free()
                sharded_phases.update(p.timings)
                return p.to_bytes()
            for _ in range(2):
                sharded_bytes = step_sharded()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            nsh = 4
            for _ in range(nsh):
                sharded_bytes = step_sharded()
            s1.record(stream)
            barrier()
            sh_ms = sharding.max_over_ranks([s0.elapsed_time(s1) / nsh], device="cuda")[0]
            import hashlib
            same = sharding.gather_digests({rank: sharded_bytes}, world, device="cuda")
            ok = len({d[0] for d in same}) == 1
            phases_sh = {k: round(v, 3) for k, v in sharded_phases.items()}
            if rank == 0:
                ok = ok and hashlib.sha256(step_sharded(single=True)).hexdigest() == same[0][0]
            intra = {"world": world, "ms_per_proof": sh_ms, "proof_identical_on_all_ranks_and_to_unsharded": ok, "phase_ms_rank0": phases_sh,
                     "collectives": "NCCL all_gather: 3 x cap digests, opening values (host-staged blocks); quotient values, FRI partial sums, opened rows (device to device)"}
        except Exception as e:   # noqa: BLE001
            intra = {"world": world, "error": "%s: %s" % (type(e).__name__, e)}

    ms_total, e2e_ms = sharding.max_over_ranks([ms_total, e2e_s * 1000.0], device="cuda")
    # the only other cross-rank traffic: digests of the last proof of every rank (the "gather" of SURVEY §8e)
    digests = sharding.gather_digests({rank: last_proof_bytes}, world, device="cuda")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    L = stark.num_rows * 2
    nz = stark.num_permutation_pairs   # num_challenges(2) * pairs / batch(2)
    leaf = kstats.get("merkle_leaf_hash", {"ms": 0, "count": 1})
    leaf_bytes_per_proof = (stark.num_columns + nz + 4) * L * 8 + 3 * L * 32   # LDE rows read + digests written
    leaf_launches = max(leaf["count"], 1)
    achieved = (leaf_bytes_per_proof * ksteps / leaf_launches) / (leaf["ms"] / leaf_launches / 1e3) / 1e9 if leaf["ms"] else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    perms_per_proof = L * ((stark.num_columns + 7) // 8 + (nz + 7) // 8) + 3 * (L - 16)
    total_kernel_ms = sum(v["ms"] for v in kstats.values())
    line = {
        "metric": METRIC, "value": world * args.steps / (ms_total / 1e3), "unit": "proofs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (Goldilocks field; BN254 Fq on 8x32-bit limbs)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "instances_per_proof": NUM_IO, "parallelism": "independent proofs, 1 process per GPU, no collective; %d proofs in flight per GPU (one CUDA stream each)" % nslots,
                   "l2_policy": "per-proof working set ~5 GB >> 126 MB L2 (no flush needed)"},
        "instances_per_s": world * args.steps * NUM_IO / (ms_total / 1e3),
        "e2e": {"value": world * args.steps / (e2e_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": NUM_IO * stark.io_size + stark.num_public_inputs * 8,
                "d2h_bytes_per_step": nbytes + NUM_IO * stark.result_words * 8},
        "proof_sha256_per_rank": [d[0][:16] for d in digests],
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": "k_leaf_hash (Poseidon Merkle leaves)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     # ncu --set full capture of the trace-commitment launch (profiles/r01_leaf_hash_ncu_summary_final.txt): dram read + write
                     # 1.7639 GB + 10.0 MB = its algorithmic bytes (1676 columns x 2^17 rows x 8 B + digests): no re-reads.  G1 shape only.
                     "traffic": 1.7740e9 if AIR == "g1" else None,
                     "traffic_note": "largest of the 3 launches per proof (trace commitment); `achieved` averages all 3 (trace, Z, quotient commitments)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                     "note": "kernel is bound by the integer multiplier (FMA-heavy) pipe, 82 %% busy under ncu (Poseidon ~ 2.5e4 integer ops per 64 B absorbed); see int_pipe. Kernel durations come from a serial pass of %d steps on one stream right after the timed region (overlapped proofs would blur per-kernel events)" % ksteps,
                     "share_of_kernel_time": leaf["ms"] / total_kernel_ms if total_kernel_ms else None,
                     # the roofline that actually binds this kernel: permutations/s against the rate at which the multiplier pipe
                     # would be 100 % busy with this build's instruction mix (= achieved / pipe occupancy measured by ncu)
                     "binding": {"bound": "integer multiplier (FMA-heavy) pipe", "unit": "Mperm/s",
                                 "achieved": perms_per_proof * ksteps / (leaf["ms"] / 1e3) / 1e6 if leaf["ms"] else None,
                                 "peak": perms_per_proof * ksteps / (leaf["ms"] / 1e3) / 1e6 / FMAHEAVY_BUSY_NCU if leaf["ms"] else None,
                                 "frac": FMAHEAVY_BUSY_NCU, "source": "sm__pipe_fmaheavy_cycles_active, profiles/r01_leaf_hash_ncu_summary_final.txt"}},
        # integer-pipe view of the same kernel: ncu counts 770 warp instructions per permutation for this build (21.2 G warp
        # instructions / 27.5 M permutations, profiles/r01_leaf_hash_ncu_summary_final.txt); issue peak = 148 SMs x 4 schedulers x
        # 1 instr/clk x sm_max_mhz.  The binding unit is the multiplier pipe (sm__pipe_fmaheavy_cycles_active, ncu), not issue.
        "int_pipe": {"poseidon_perms_per_s": perms_per_proof * ksteps / (leaf["ms"] / 1e3) if leaf["ms"] else None, "perms_per_proof": perms_per_proof,
                     "warp_instr_per_perm": WARP_INSTR_PER_PERM,
                     "issue_frac": (perms_per_proof * ksteps / (leaf["ms"] / 1e3) * WARP_INSTR_PER_PERM) / (148 * 4 * 1965e6) if leaf["ms"] else None,
                     "fmaheavy_pipe_busy_ncu": FMAHEAVY_BUSY_NCU},
        "serial_ms_per_step": serial_ms_per_step,
        "intra_proof": intra,
        "kernel_ms_per_proof": {k: round(v["ms"] / ksteps, 3) for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])},
        "phase_ms_last_proof": {k: round(v, 3) for k, v in phases.items()},
    }
    if not args.no_cpu_baseline and world == 1:
        orc = entry.load_oracle()
        orc.set_threads(os.cpu_count())
        full_ms, est, wall = cpu_sample(orc, gen_ios(NUM_IO))
        line["cpu_baseline"] = {"value": 1000.0 / full_ms, "unit": "proofs/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "oracle (C++/OpenMP restatement, not the Rust binary): heavy phases on 1/%d of their columns/instances/points scaled x%d, "
                                          "FRI tail in full; %.1f s of CPU wall; est. full-proof phases ms=%s" % (1 << SAMPLE_SHIFT, 1 << SAMPLE_SHIFT, wall, {k: round(v, 1) for k, v in est.items()})}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

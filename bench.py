#!/usr/bin/env python3
"""Headline benchmark: BN254 STARK proofs per second (BASELINE.json metric; configs[1] = batched G1 scalar multiplication).

One "step" = one proof of G1ExpStark with 128 independent scalar multiplications (2^16 rows x 1676 columns, default StarkConfig):
trace generation (K1) + prove (K2-K6), through the C ABI's batched entry point sbn_prove_batch (one host call per timed region).
  value : inputs already resident in HBM, proofs/s over all ranks
  e2e   : pinned host buffers in, proof bytes out (H2D / D2H inside the timed region)
  airs  : the same two numbers for the other AIRs the metric names (G2 scalar multiplication, Fq12 exponentiation) and FqExp
Multi-GPU: proofs are independent -> one process per GPU, no data-path collective (weak scaling).
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port") on the host cores: one FULL proof per step.
`--sweep modular` runs BASELINE.json configs[4] (ModularStark 2^16..2^22 rows x rate_bits 1..3) and prints one JSON line per point.
"""
import argparse
import hashlib
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

# --air selects the headline AIR; the default (BASELINE.json configs[1]) is g1.  (class, oracle id, num_io, input generator, metric, workload)
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
# lanes per batch (gpurun_out/r2w_*: Fq12 6 / 10 / 16 / 24 lanes -> 13.5 / 13.3 / 14.3 / 13.8 proofs/s device-resident, 12.8 / 13.9 / 14.8 / 16.0 end to end;
# Fq 6 / 8 / 12 -> 30.5 / 32.9 / 31.8; G2 6 / 8 -> 8.1 / 8.7; G1 4 / 6 / 7 / 8 -> 14.4 / 15.9 / 15.8 / 15.6)
LANES = {"g1": 6, "g2": 8, "fq12": 16, "fq": 8}
DTYPE = "u64 (Goldilocks field; BN254 Fq on 8x32-bit limbs)"
# ncu --set full capture of the trace-commitment launch of k_leaf_hash for the G1 shape (profiles/r02_leaf_hash_ncu_summary.txt):
# dram read 1.7601 GB + write 9.57 MB = its algorithmic bytes (1676 columns x 2^17 rows x 8 B + digests), no re-reads
LEAF_HASH_TRAFFIC_G1 = 1.7696e9


def config_of(air, num_io):
    """The `config` object, identical on both arms (driver: same_config)."""
    workload = AIRS[air][5]
    if num_io != AIRS[air][2]:
        workload = workload.replace("num_io=%d:" % AIRS[air][2], "num_io=%d (non-default size; row / column counts in this text are those of the default):" % num_io)
    return {"workload": workload, "instances_per_proof": num_io,
            "parallelism": "independent proofs, one process per GPU, no data-path collective",
            "l2_policy": "per-proof working set ~5 GB >> 126 MB L2 (no flush needed)"}


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
        pw = [float(s[2]) for s in self.samples if len(s) >= 7 and s[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def oracle_full_proof(orc, air, ios):
    """One full CPU proof with the oracle: generate_trace, outputs, public inputs, prove.  Returns (seconds, proof bytes)."""
    sbn = entry.load_package()
    t0 = time.perf_counter()
    trace, res = air.generate_trace(ios)
    out_off = air.io_size - 8 * air.result_words
    filled = sbn.synthetic.fill_outputs(ios, res, air.io_size, out_off) if air.result_words else ios
    pi = air.generate_public_inputs(filled)
    proof = air.prove(trace, pi)
    return time.perf_counter() - t0, proof


def run_reference(args):
    """CPU arm: the oracle (C++/OpenMP port of the reference algorithm, NOT the Rust binary) on all host cores, one FULL proof
    (trace generation + prove) per step -- nothing sampled or extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core (rank 0 alone runs it)
    cores = os.cpu_count()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    orc = entry.load_oracle()
    orc.set_threads(cores)
    sbn = entry.load_package()
    air_name, num_io = args.air, args.num_io or AIRS[args.air][2]
    air = orc.Air(AIRS[air_name][1], num_io)
    ios = getattr(sbn.synthetic, AIRS[air_name][3])(num_io)
    for _ in range(args.warmup):
        oracle_full_proof(orc, air, ios)
    t0 = time.perf_counter()
    secs = []
    for _ in range(args.steps):
        s, proof = oracle_full_proof(orc, air, ios)
        secs.append(s)
    wall = time.perf_counter() - t0
    ms = 1000.0 * wall / args.steps
    value = 1000.0 / ms
    sample = ("every step is one full oracle proof (trace generation + prove) of the workload on %d OpenMP threads; the oracle is a C++ restatement of the "
              "reference algorithm, not the Rust binary, and hashes with scalar 64-bit code (plonky2 has an AVX2 Poseidon)" % cores)
    line = {"impl": "reference", "metric": AIRS[air_name][4], "value": value, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic", "config": config_of(air_name, num_io),
            "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "step_s_min_max": [min(secs), max(secs)], "proof_sha256": hashlib.sha256(proof).hexdigest()[:16], "wall_s": wall}
    print(json.dumps(line))


class AirBench:
    """Throughput of one AIR on this rank's GPU through sbn_prove_batch."""

    def __init__(self, sbn, torch, batch, local, rank, air_name, num_io):
        self.sbn, self.torch, self.batch = sbn, torch, batch
        cls, _, default_io, gen, self.metric, _ = AIRS[air_name]
        self.air_name, self.num_io = air_name, num_io or default_io
        self.stark = getattr(sbn, cls)(self.num_io)
        self.cfg = self.stark.config()
        # distinct synthetic batches per rank, reused round-robin (input generation is host big-int work); host copies pinned
        self.host = []
        for b in range(4):
            raw = getattr(sbn.synthetic, gen)(self.num_io, seed=0x5EED0001 + 1000 * rank + b)
            self.host.append(torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory())
        self.dev = [t.cuda(non_blocking=True) for t in self.host]
        torch.cuda.synchronize()

    def resident(self, first, count):
        ptrs = [self.dev[(first + i) % len(self.dev)].data_ptr() for i in range(count)]
        return self.sbn.prove_batch(self.stark, self.cfg, self.batch, ptrs, on_device=True)

    def e2e(self, first, count):
        ptrs = [self.host[(first + i) % len(self.host)].data_ptr() for i in range(count)]
        return [p.to_bytes() for p in self.sbn.prove_batch(self.stark, self.cfg, self.batch, ptrs)]

    def h2d_bytes(self):
        return self.num_io * self.stark.io_size

    def perms_per_proof(self):
        st = self.stark
        L, nz = st.num_rows * 2, st.num_permutation_pairs   # num_challenges (2) * pairs / batch (2) Z polynomials
        return L * ((st.num_columns + 7) // 8 + (nz + 7) // 8) + 3 * (L - 16)


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
    ap.add_argument("--no-other-airs", action="store_true", help="measure only the headline AIR")
    ap.add_argument("--inflight", type=int, default=0, help="lanes of the batch = independent proofs in flight per GPU (one CUDA stream + native host thread each); "
                    "0 = the AIR's default (LANES: the small proofs of Fq12 / Fq need more of them in flight to fill the machine)")
    ap.add_argument("--sweep", default=None, choices=["modular"], help="BASELINE.json configs[4]: one JSON line per (rows, rate_bits) point")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.sweep:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sweep_modular
        return sweep_modular.main_from_bench(args)
    if args.warmup < 3:
        args.warmup = 3

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
    from starky_bn254_b200 import sharding
    stream = torch.cuda.current_stream()
    try:
        ncpu = len(os.sched_getaffinity(0))   # the cores this process may run on (cpuset-aware)
    except AttributeError:
        ncpu = os.cpu_count() or 16
    host_threads = max(6, ncpu // max(world, 1))   # one native host thread per lane and rank

    def lanes_for(air):
        return max(1, args.inflight) if args.inflight else min(LANES[air], host_threads)
    lanes = lanes_for(args.air)
    batch = sbn.Batch(local, lanes)
    head = AirBench(sbn, torch, batch, local, rank, args.air, args.num_io)
    stark, cfg, syn = head.stark, head.cfg, sbn.synthetic

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, count):
        """`count` proofs in one sbn_prove_batch call, bracketed by barrier + synchronize; device time between two events recorded on
        torch's (otherwise idle) current stream -- the call returns only when every lane has synchronised its own stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        out = fn(first, count)
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        return e0.elapsed_time(e1), wall_ms, out

    head.resident(0, max(args.warmup, lanes))
    sampler = ClockSampler(local)
    sampler.start()
    launches0, bytes0 = batch.launch_count, batch.device_bytes
    ms_total, _, proofs = timed(head.resident, args.warmup, args.steps)
    launches, pool_growth = batch.launch_count - launches0, batch.device_bytes - bytes0   # growth > 0: a lane called cudaMalloc inside the timed region
    # ---- end to end: pinned host inputs -> proof bytes on the host ----
    head.e2e(0, lanes)
    _, e2e_ms, proof_bytes = timed(head.e2e, args.warmup, args.steps)
    last_proof_bytes = proof_bytes[-1]
    # ---- per-kernel CUDA-event timing: one proof at a time on one context's stream (overlapped proofs would blur it) ----
    ctx = sbn.Context(local)
    ksteps = min(args.steps, 3)
    out_off = stark.io_size - 8 * stark.result_words
    host_raw = [bytes(t.numpy().tobytes()) for t in head.host]

    def serial_proof(i):
        tr = stark.generate_trace_device(head.dev[i % len(head.dev)].data_ptr(), ctx)
        ios = syn.fill_outputs(host_raw[i % len(host_raw)], tr.results(), stark.io_size, out_off)
        p = sbn.prove(stark, cfg, tr, stark.generate_public_inputs(ios))
        tr.free()
        return p
    serial_proof(0)            # untimed: this context builds its own twiddle / power tables and warms its allocator
    ctx.kernel_timing(True)
    proof = None
    t0 = time.perf_counter()
    for i in range(ksteps):
        proof = serial_proof(i)
    serial_ms_per_step = (time.perf_counter() - t0) * 1e3 / ksteps
    kstats = ctx.kernel_stats()
    ctx.kernel_timing(False)
    phases = proof.timings
    ctx.close()
    clocks = sampler.summary()
    # ---- the other AIRs the metric names (same batch, fewer steps) ----
    others = {}
    if not args.no_other_airs and args.air == "g1" and not args.num_io:
        obatch, olanes = batch, lanes
        for name in ("g2", "fq12", "fq"):
            try:
                obatch.trim()   # the lanes' allocators cache the previous AIR's block sizes
                if lanes_for(name) != olanes:
                    if obatch is not batch:
                        obatch.close()
                    olanes = lanes_for(name)
                    obatch = batch if olanes == lanes else sbn.Batch(local, olanes)
                ab = AirBench(sbn, torch, obatch, local, rank, name, 0)
                osteps = -(-max(4, args.steps // 2) // olanes) * olanes   # whole rounds of the lanes (a ragged last round runs almost alone)
                ab.resident(0, olanes)
                oms, _, _ = timed(ab.resident, 0, osteps)
                ab.e2e(0, olanes)   # every lane's pinned staging arena exists before the timed region
                _, oe2e, ob = timed(ab.e2e, 0, osteps)
                oms, oe2e = sharding.max_over_ranks([oms, oe2e], device="cuda")
                others[name] = {"metric": ab.metric, "workload": AIRS[name][5], "steps": osteps, "lanes": olanes, "value": world * osteps / (oms / 1e3), "ms_per_step": oms / osteps,
                                "e2e": {"value": world * osteps / (oe2e / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": ab.h2d_bytes(), "d2h_bytes_per_step": len(ob[-1])},
                                "poseidon_perms_per_proof": ab.perms_per_proof(), "proof_sha256": hashlib.sha256(ob[-1]).hexdigest()[:16]}
                del ab
            except Exception as e:   # noqa: BLE001  (the headline must survive)
                others[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    sampler.stop_flag.set()
    sampler.join()

    ms_total, e2e_ms = sharding.max_over_ranks([ms_total, e2e_ms], device="cuda")
    # the only other cross-rank traffic: digests of the last proof of every rank (the "gather" of SURVEY 8e)
    digests = sharding.gather_digests({rank: last_proof_bytes}, world, device="cuda")

    line = None
    if rank == 0:
        L = stark.num_rows * 2
        nz = stark.num_permutation_pairs
        leaf = kstats.get("merkle_leaf_hash", {"ms": 0, "count": 1})
        leaf_bytes_per_proof = (stark.num_columns + nz + 4) * L * 8 + 3 * L * 32   # LDE rows read + digests written
        leaf_launches = max(leaf["count"], 1)
        hbm_achieved = (leaf_bytes_per_proof * ksteps / leaf_launches) / (leaf["ms"] / leaf_launches / 1e3) / 1e9 if leaf["ms"] else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        perms_per_proof = head.perms_per_proof()
        perms_per_s = perms_per_proof * ksteps / (leaf["ms"] / 1e3) if leaf["ms"] else None
        total_kernel_ms = sum(v["ms"] for v in kstats.values())
        # integer-pipe roofline, computed: instruction counts from the shipped SASS x issue rates measured now on this GPU
        roof = {"bound": "hbm", "kernel": "k_leaf_hash (Poseidon Merkle leaves)", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (hbm_achieved / hbm_peak) if hbm_achieved else None}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import int_roofline
            mdl = int_roofline.model()
            props = torch.cuda.get_device_properties(local)
            mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965.0
            ceil = int_roofline.ceilings(mdl, props.multi_processor_count, mhz)
            names = {"mult": "integer multiplier (FMA-heavy) pipe", "alu": "integer ALU pipe", "issue": "instruction issue"}
            binding = min(ceil, key=ceil.get)
            roof = {"bound": "int_" + binding, "bound_name": names[binding], "kernel": "k_leaf_hash (Poseidon Merkle leaves)",
                    "achieved": perms_per_s / 1e6 if perms_per_s else None, "peak": ceil[binding] / 1e6, "unit": "Mperm/s",
                    "frac": perms_per_s / ceil[binding] if perms_per_s else None,
                    "ceilings_mperm_s": {names[k]: round(v / 1e6, 1) for k, v in ceil.items()},
                    "model": {"sm_count": props.multi_processor_count, "sm_mhz": mhz, "bodies": mdl["bodies"], "warp_instr_per_32_permutations": mdl["warp_instr_per_permutation_x32"],
                              "rates_warp_instr_per_clk_per_sm": mdl["rates_warp_instr_per_clk_per_sm"], "rates_source": mdl["rates_source"],
                              "method": "tools/int_roofline.py: opcode counts of the two round-loop bodies in the shipped k_leaf_hash SASS (8 full + 22 partial rounds) divided by "
                                        "issue rates measured by tools/microbench/int_throughput.cu in this run; ceiling = 32 x SMs x clock / clk-per-warp-permutation"},
                    "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": (hbm_achieved / hbm_peak) if hbm_achieved else None,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                            "algorithmic_bytes_per_proof": leaf_bytes_per_proof}}
        except Exception as e:   # noqa: BLE001
            roof["int_model_error"] = "%s: %s" % (type(e).__name__, e)
        roof.update({"traffic": LEAF_HASH_TRAFFIC_G1 if (args.air == "g1" and not args.num_io) else None,
                     "traffic_note": "ncu dram read + write of the largest of the 3 launches per proof (trace commitment); `achieved` averages all 3 (trace, Z, quotient commitments)",
                     "share_of_kernel_time": leaf["ms"] / total_kernel_ms if total_kernel_ms else None,
                     "note": "kernel durations are CUDA-event times of a serial pass of %d proofs on one stream right after the timed region (overlapped proofs would blur per-kernel events)" % ksteps})
        line = {
            "metric": head.metric, "value": world * args.steps / (ms_total / 1e3), "unit": "proofs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_of(args.air, head.num_io),
            "inflight_per_gpu": lanes, "allocator_growth_bytes_in_timed_region": pool_growth, "api": "sbn_prove_batch: one call per timed region, %d lanes (CUDA stream + native host thread each), one Python thread" % lanes,
            "instances_per_s": world * args.steps * head.num_io / (ms_total / 1e3),
            "e2e": {"value": world * args.steps / (e2e_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": head.h2d_bytes(), "d2h_bytes_per_step": len(last_proof_bytes)},
            "airs": others,
            "proof_sha256_per_rank": [d[0][:16] for d in digests],
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "int_pipe": {"poseidon_perms_per_s": perms_per_s, "perms_per_proof": perms_per_proof},
            "serial_ms_per_step": serial_ms_per_step,
            "intra_proof": None,
            "kernel_ms_per_proof": {k: round(v["ms"] / ksteps, 3) for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])},
            "phase_ms_last_proof": {k: round(v, 3) for k, v in phases.items()},
        }

    printed = threading.Event()

    def emit():
        if rank == 0 and not printed.is_set():
            printed.set()
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(line) + "\n").encode())

    # ---- latency mode: ONE proof computed by all ranks together (sbn_prove_sharded).  Reported beside the throughput
    # number.  A rank that fails between collectives would leave its peers blocked in NCCL: a watchdog prints the headline line
    # (already complete above) and ends the process if this section does not finish in time. ----
    batch.trim()
    if world > 1 and world in (2, 4, 8, 16) and not args.no_intra_proof:
        def watchdog():
            if rank == 0 and line is not None:
                line["intra_proof"] = {"world": world, "error": "timed out after %d s" % args_intra_timeout}
            emit()
            os._exit(0)
        args_intra_timeout = 240
        wd = threading.Timer(args_intra_timeout, watchdog)
        wd.daemon = True
        wd.start()
        intra = None
        try:
            ctx = sbn.Context(local, stream.cuda_stream)
            sk = getattr(sbn, AIRS[args.air][0])(head.num_io, ctx)
            ag = sharding.dist_allgather(device=torch.device("cuda", local))
            agd = sharding.dist_allgather_device(torch.device("cuda", local))
            raw0 = getattr(syn, AIRS[args.air][3])(head.num_io, seed=0x5EED0001)          # the same inputs on every rank
            sharded_phases = {}

            def step_sharded(single=False):
                tr = sk.generate_trace(raw0)
                ios = syn.fill_outputs(raw0, tr.results(), sk.io_size, out_off)
                pi = sk.generate_public_inputs(ios)
                p = sbn.prove(sk, cfg, tr, pi) if single else sbn.prove_sharded(sk, cfg, tr, pi, rank, world, ag, allgather_device=agd)
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
            same = sharding.gather_digests({rank: sharded_bytes}, world, device="cuda")
            ok = len({d[0] for d in same}) == 1
            phases_sh = {k: round(v, 3) for k, v in sharded_phases.items()}   # before the unsharded comparison proof overwrites them
            if rank == 0:
                ok = ok and hashlib.sha256(step_sharded(single=True)).hexdigest() == same[0][0]
            intra = {"world": world, "ms_per_proof": sh_ms, "proof_identical_on_all_ranks_and_to_unsharded": ok,
                     "phase_ms_rank0": phases_sh,
                     "collectives": "NCCL all_gather: 3 x cap digests, opening values (host-staged blocks); quotient values, FRI partial sums, opened rows (device to device)"}
        except Exception as e:   # noqa: BLE001
            intra = {"world": world, "error": "%s: %s" % (type(e).__name__, e)}
        wd.cancel()
        if rank == 0:
            line["intra_proof"] = intra

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        # one FULL oracle proof of the same workload on all host cores (about 10-30 s), after every GPU measurement
        orc = entry.load_oracle()
        cores = os.cpu_count()
        orc.set_threads(cores)
        air = orc.Air(AIRS[args.air][1], head.num_io)
        secs, oproof = oracle_full_proof(orc, air, getattr(syn, AIRS[args.air][3])(head.num_io))
        line["cpu_baseline"] = {"value": 1.0 / secs, "unit": "proofs/s", "cores": cores, "kind": "port",
                                "sample": "ONE full proof (trace generation + prove, %.1f s) of the same workload by the oracle (C++/OpenMP restatement of the reference "
                                          "algorithm, not the Rust binary; scalar Poseidon where plonky2 has AVX2) on %d threads" % (secs, cores),
                                "proof_sha256": hashlib.sha256(oproof).hexdigest()[:16]}
    emit()
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""BASELINE.json config 5: standalone modular mul AIR sweep -- ModularStark (reference src/modular/modular.rs:361-537 with the
row count a parameter) at 2^16 .. 2^max rows, rate_bits 1..3: per-phase device milliseconds of trace generation (K1), LDE +
Merkle (K2 + K3: trace, Z and quotient commitments), Z polynomials (K4), quotient (K5) and the rest (K6), plus the achieved
fraction of the HBM roofline for the NTT/LDE kernels.  Writes one JSON object per (rows, rate) to stdout.
    python tools/sweep_modular.py [max_log_rows=20] [min_log_rows=16]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

C, PAIRS, Q = 812, 444, 4


def fast_ios(n, seed=5):
    """2n uniform residues below the BN254 modulus (numpy: the SplitMix64 generator of synthetic.py is too slow for 2^21 rows)."""
    rng = np.random.default_rng(seed)
    p = entry.load_package().synthetic.BN254_P
    top = p >> 192
    w = rng.integers(0, 1 << 63, size=(2 * n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(2 * n, 4), dtype=np.uint64)
    w[:, 3] = rng.integers(0, top, size=2 * n, dtype=np.uint64)   # top limb strictly below p's top limb => value < p
    return w.tobytes()


def main_from_bench(args):
    """`python bench.py --sweep modular`: the same sweep, 2^16 .. 2^22 rows (SBN_SWEEP_MAX_LOG / SBN_SWEEP_MIN_LOG override)."""
    return run(int(os.environ.get("SBN_SWEEP_MAX_LOG", "22")), int(os.environ.get("SBN_SWEEP_MIN_LOG", "16")))


def main():
    max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    min_log = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    return run(max_log, min_log)


def run(max_log, min_log):
    sbn = entry.load_package()
    ctx = sbn.Context(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    for logn in range(min_log, max_log + 1):
        n = 1 << logn
        ios = fast_ios(n)
        stark = sbn.ModularStark(n, ctx)
        for rate_bits in (1, 2, 3):
            L = n << rate_bits
            cfg = stark.config()
            cfg.rate_bits = rate_bits
            best = None
            try:
              for rep in range(2 if logn <= 21 else 1):   # first pass warms the allocator and twiddle tables (2^22: one pass, tables come from the previous sizes)
                ctx.kernel_timing(True)
                t0 = time.perf_counter()
                tr = stark.generate_trace(ios)
                ctx.synchronize()
                t1 = time.perf_counter()
                try:
                    proof = sbn.prove(stark, cfg, tr, np.zeros(0, dtype=np.uint64))
                finally:
                    tr.free()
                t2 = time.perf_counter()
                ks = ctx.kernel_stats()
                ctx.kernel_timing(False)
                best = (t1 - t0, t2 - t1, proof.timings, ks, len(proof.to_bytes()), hashlib.sha256(proof.to_bytes()).hexdigest()[:16])
            except sbn.SbnError as e:
                print(json.dumps({"rows_log2": logn, "rate_bits": rate_bits, "error": str(e)[:200]}), flush=True)
                continue
            tg, tp, ph, ks, plen, sha = best
            ntt_ms = sum(ks.get(k, {"ms": 0})["ms"] for k in ("ntt_pass1", "ntt_pass2", "ntt_small"))
            ntt_bytes = 8 * ((C + PAIRS) * (2 * n + 2 * L) + Q * 2 * L)    # iNTT read+write, LDE read (per coset) + write
            leaf_ms = ks.get("merkle_leaf_hash", {"ms": 0})["ms"]
            perms = L * ((C + 7) // 8 + (PAIRS + 7) // 8)
            out = {"rows_log2": logn, "rate_bits": rate_bits, "tracegen_ms": round(tg * 1e3, 2), "prove_ms": round(tp * 1e3, 2),
                   "lde_merkle_ms": round(ph["compute trace commitment"] + ph.get("compute permutation Z commitments", 0) + ph["compute quotient commitment"], 2),
                   "z_polys_ms": round(ph.get("compute permutation Z polys", 0), 2), "quotient_ms": round(ph["compute quotient polys"], 2),
                   "openings_fri_ms": round(ph["total"] - ph["compute trace commitment"] - ph.get("compute permutation Z commitments", 0) - ph["compute quotient commitment"]
                                            - ph.get("compute permutation Z polys", 0) - ph["compute quotient polys"], 2),
                   "ntt_kernels_ms": round(ntt_ms, 2), "ntt_hbm_frac": round(ntt_bytes / (ntt_ms / 1e3) / 1e9 / hbm, 4) if ntt_ms else None,
                   "leaf_hash_ms": round(leaf_ms, 2), "poseidon_mperm_s": round(perms / (leaf_ms / 1e3) / 1e6, 1) if leaf_ms else None,
                   "proof_bytes": plen, "proof_sha256": sha, "device_gb": round(ctx.device_bytes / 1e9, 2),
                   "streamed": 8.0 * n * ((C + PAIRS) * (1 + (1 << rate_bits)) + PAIRS) * 1.1 > 0.8 * 191.5e9 - 8.0 * n * C}
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

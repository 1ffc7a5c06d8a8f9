#!/usr/bin/env python3
"""Generates tests/golden/golden.json from the CPU oracle on seeded inputs.

The reference stores no golden vectors (every test draws unseeded random inputs) and
cannot be built here (Rust + un-vendored git dependencies), so these fixtures pin the ORACLE to itself
across refactors and give the GPU tests a target that needs no oracle prover run:
    python tests/golden/make_golden.py            # rewrites golden.json (G1 takes ~2 minutes of CPU)
Poseidon's zero-vector KAT is the one externally published vector (plonky2's test suite).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

orc = g.load_oracle()
sbn = g.load_package()
syn = sbn.synthetic


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    out = {}
    out["poseidon_zero"] = ["%016x" % int(x) for x in orc.poseidon(np.zeros(12, dtype=np.uint64))]
    out["poseidon_iota"] = ["%016x" % int(x) for x in orc.poseidon(np.arange(12, dtype=np.uint64))]
    # ModularStark, 512 rows (reference src/modular/modular.rs:539-558 shape)
    n = 512
    ios = syn.modular_ios(n)
    air = orc.Air(orc.AIR_MODULAR, n)
    trace, _ = air.generate_trace(ios)
    proof = air.prove(trace, np.zeros(0, dtype=np.uint64))
    assert air.verify(proof)[0]
    out["modular_512"] = {"ios_sha256": sha(ios), "trace_sha256": sha(trace.tobytes()), "proof_sha256": sha(proof), "proof_len": len(proof),
                          "trace_cap0": ["%016x" % int.from_bytes(proof[4 + 8 * i:12 + 8 * i], "little") for i in range(4)]}
    # G1ExpStark, 128 scalar multiplications (reference src/curves/g1/exp.rs:784-826 shape)
    if "--skip-g1" not in sys.argv:
        n = 128
        ios = syn.g1_exp_ios(n)
        air = orc.Air(orc.AIR_G1_EXP, n)
        trace, res = air.generate_trace(ios)
        ios = syn.fill_g1_outputs(ios, res)
        pi = air.generate_public_inputs(ios)
        proof = air.prove(trace, pi)
        assert air.verify(proof)[0]
        out["g1_128"] = {"ios_sha256": sha(ios), "trace_sha256": sha(trace.tobytes()), "results_sha256": sha(res.tobytes()), "pi_sha256": sha(pi.tobytes()),
                         "proof_sha256": sha(proof), "proof_len": len(proof),
                         "trace_cap0": ["%016x" % int.from_bytes(proof[4 + 8 * i:12 + 8 * i], "little") for i in range(4)]}
    else:
        old = json.load(open(os.path.join(HERE, "golden.json")))
        out["g1_128"] = old["g1_128"]
    # FqExpStark n=128, G2ExpStark n=128, Fq12ExpStark n=16, Fq12ExpU64Stark n=16 (BASELINE.json configs 3-4 and SURVEY §8 f1)
    old = json.load(open(os.path.join(HERE, "golden.json"))) if os.path.exists(os.path.join(HERE, "golden.json")) else {}
    more = [("fq_128", orc.AIR_FQ_EXP, 128, syn.fq_exp_ios, syn.FQ_IO_SIZE, 96), ("g2_128", orc.AIR_G2_EXP, 128, syn.g2_exp_ios, syn.G2_IO_SIZE, 288),
            ("fq12_16", orc.AIR_FQ12_EXP, 16, syn.fq12_exp_ios, syn.FQ12_IO_SIZE, 800), ("fq12u64_16", orc.AIR_FQ12_EXP_U64, 16, syn.fq12_exp_u64_ios, syn.FQ12_U64_IO_SIZE, 776)]
    only = [a.split("=")[1] for a in sys.argv if a.startswith("--only=")]
    for name, air_id, n, gen, io_size, out_off in more:
        if (only and name not in only) or ("--skip-new" in sys.argv):
            if name in old:
                out[name] = old[name]
            continue
        ios = gen(n)
        air = orc.Air(air_id, n)
        trace, res = air.generate_trace(ios)
        ios = syn.fill_outputs(ios, res, io_size, out_off)
        pi = air.generate_public_inputs(ios)
        proof = air.prove(trace, pi)
        assert air.verify(proof)[0]
        out[name] = {"ios_sha256": sha(ios), "trace_sha256": sha(trace.tobytes()), "results_sha256": sha(res.tobytes()), "pi_sha256": sha(pi.tobytes()),
                     "proof_sha256": sha(proof), "proof_len": len(proof),
                     "trace_cap0": ["%016x" % int.from_bytes(proof[4 + 8 * i:12 + 8 * i], "little") for i in range(4)]}
        print(name, out[name], flush=True)
        del trace, proof
    # gadget test AIRs: G1Stark 512 rows (reference src/curves/g1/muladd.rs:463), Fq12Stark 512 rows (src/fields/fq12/mul.rs:364)
    for name, air_id, n, gen in (("g1_muladd_512", orc.AIR_G1_MULADD, 512, syn.g1_muladd_ios), ("fq12_mul_512", orc.AIR_FQ12_MUL, 512, syn.fq12_mul_ios)):
        if (only and name not in only) or ("--skip-new" in sys.argv):
            if name in old:
                out[name] = old[name]
            continue
        ios = gen(n)
        air = orc.Air(air_id, n)
        trace, _ = air.generate_trace(ios)
        proof = air.prove(trace, np.zeros(0, dtype=np.uint64))
        assert air.verify(proof)[0]
        out[name] = {"ios_sha256": sha(ios), "trace_sha256": sha(trace.tobytes()), "proof_sha256": sha(proof), "proof_len": len(proof),
                     "trace_cap0": ["%016x" % int.from_bytes(proof[4 + 8 * i:12 + 8 * i], "little") for i in range(4)]}
        print(name, out[name], flush=True)
    # BASELINE.json configs[4] at 2^16 rows, rate_bits 1..3 (the sweep's own numpy-generated inputs; ~4 minutes of oracle time)
    if "--with-sweep" in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sweep_modular
        n = 1 << 16
        ios = sweep_modular.fast_ios(n)
        air = orc.Air(orc.AIR_MODULAR, n)
        trace, _ = air.generate_trace(ios)
        blk = {"ios_sha256": sha(ios), "trace_sha256": sha(trace.tobytes())}
        for r in (1, 2, 3):
            cfg = orc.Config.standard_fast_config(rate_bits=r)
            proof = air.prove(trace, np.zeros(0, dtype=np.uint64), cfg)
            assert air.verify(proof, cfg)[0]
            blk["rate_bits_%d" % r] = {"proof_sha256": sha(proof), "proof_len": len(proof),
                                       "trace_cap0": ["%016x" % int.from_bytes(proof[4 + 8 * i:12 + 8 * i], "little") for i in range(4)]}
        out["modular_2p16_sweep"] = blk
    else:
        out["modular_2p16_sweep"] = json.load(open(os.path.join(HERE, "golden.json")))["modular_2p16_sweep"]
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

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
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Writes the proof files the Rust differential test reads (bindings/rust/starky-bn254-b200/tests/differential.rs):
FqExpStark, num_io = 128, seed 0x5EED0000, one proof per setting of U1 (generator pair) x U3 (FRI degree hack).
Default: the CPU oracle makes them (gcc only, ~1 minute); --gpu: the B200 prover does (the GPU tests assert byte identity between
the two for every setting).  Not product code."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

GEN = {"B": 7, "A": 14293326489335486720}


def main():
    gpu = "--gpu" in sys.argv
    out = os.path.join(ROOT, "bindings", "rust", "starky-bn254-b200", "tests", "data")
    os.makedirs(out, exist_ok=True)
    sbn, orc = entry.load_package(), entry.load_oracle()
    n = 128
    raw = sbn.synthetic.fq_exp_ios(n)
    if gpu:
        ctx = sbn.Context(0)
        stark = sbn.FqExpStark(n, ctx)
        tr = stark.generate_trace(raw)
        ios = sbn.synthetic.fill_outputs(raw, tr.results(), stark.io_size, stark.io_size - 8 * stark.result_words)
        pi = stark.generate_public_inputs(ios)
    else:
        air = orc.Air(orc.AIR_FQ_EXP, n)
        trace, res = air.generate_trace(raw)
        ios = sbn.synthetic.fill_outputs(raw, res, air.io_size, air.io_size - 8 * air.result_words)
        pi = air.generate_public_inputs(ios)
    for pair, g in GEN.items():
        for hack, tag in ((0, "pad"), (1, "hack")):
            if gpu:
                cfg = stark.config()
                cfg.coset_shift, cfg.fri_degree_hack = g, hack
                proof = sbn.prove(stark, cfg, tr, pi).to_bytes()
            else:
                proof = air.prove(trace, pi, orc.Config.standard_fast_config(coset_shift=g, fri_degree_hack=hack))
            path = os.path.join(out, "fq_128_%s_%s.proof" % (pair, tag))
            open(path, "wb").write(proof)
            print(path, len(proof))


if __name__ == "__main__":
    main()

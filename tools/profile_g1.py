#!/usr/bin/env python3
"""Runs `reps` proofs (trace generation + prove) of one AIR on cuda:0 -- the command profiled under ncu.
    python tools/profile_g1.py [reps=2] [air=g1|g2|fq12|fq]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

sbn = entry.load_package()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
air = sys.argv[2] if len(sys.argv) > 2 else "g1"
cls, n, gen = {"g1": ("G1ExpStark", 128, "g1_exp_ios"), "g2": ("G2ExpStark", 128, "g2_exp_ios"), "fq12": ("Fq12ExpStark", 16, "fq12_exp_ios"),
               "fq": ("FqExpStark", 128, "fq_exp_ios")}[air]
ctx = sbn.Context(0)
stark = getattr(sbn, cls)(n, ctx)
ios = getattr(sbn.synthetic, gen)(n)
for _ in range(reps):
    tr = stark.generate_trace(ios)
    full = sbn.synthetic.fill_outputs(ios, tr.results(), stark.io_size, stark.io_size - 8 * stark.result_words)
    proof = sbn.prove(stark, stark.config(), tr, stark.generate_public_inputs(full))
    tr.free()
print("ok", len(proof.to_bytes()), proof.timings)

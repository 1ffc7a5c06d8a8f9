#!/usr/bin/env python3
"""Runs `reps` G1 proofs (trace generation + prove) on cuda:0 -- the command profiled under ncu."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

sbn = entry.load_package()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = sbn.Context(0)
stark = sbn.G1ExpStark(128, ctx)
ios = sbn.synthetic.g1_exp_ios(128)
for _ in range(reps):
    tr = stark.generate_trace(ios)
    full = sbn.synthetic.fill_g1_outputs(ios, tr.results())
    proof = sbn.prove(stark, stark.config(), tr, stark.generate_public_inputs(full))
    tr.free()
print("ok", len(proof.to_bytes()), proof.timings)

#!/usr/bin/env python3
"""Integer-pipe roofline of the Poseidon leaf-hash kernel, computed from the SHIPPED binary and a LIVE micro-benchmark
(not product code; bench.py imports it for its `roofline` block).

  per-permutation instruction counts : `cuobjdump -sass` of k_leaf_hash in libstarkybn254_b200.so -- the two innermost loops that
                                       hold the dp2a MDS layer are the full-round body (12 S-boxes, executed 8x) and the
                                       partial-round body (1 S-box, executed 22x);
  issue rates (warp-instr / clk / SM): tools/microbench/int_throughput.cu run on the GPU the bench runs on;
  ceiling of a resource               : 32 * SMs * f_clk / sum_ops(count / rate) permutations per second.
The three resources are the integer multiplier ("FMA-heavy") pipe (IDP.*, IMAD.*), the ALU pipe (IADD3, LOP3, PRMT, SHF, LEA, SEL ...)
and instruction issue (every instruction, against the best mixed rate the micro-benchmark reaches).  The lowest ceiling binds.
Usage: python tools/int_roofline.py [lib.so]   (prints the model as JSON; rates need a GPU, counts do not)
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "starky-bn254_b200", "libstarkybn254_b200.so")
MICROBENCH = os.path.join(ROOT, "tools", "microbench", "pb_int_throughput")
KERNEL = "k_leaf_hash"
# fallback rates: B200, profiles/r01_int_throughput_dp4a_imma.txt (used only when the micro-benchmark cannot run; flagged in the output)
FALLBACK_RATES = {"idp": 1.952, "imad": 1.944, "imad_wide": 1.026, "alu": 1.951, "iadd3": 2.593, "issue_mixed": 3.187}
NOT_A_PIPE = ("LDG", "STG", "LDC", "LDS", "STS", "BRA", "EXIT", "BSSY", "BSYNC", "NOP", "S2R", "R2UR", "LDCU", "S2UR", "CS2R", "WARPSYNC", "BAR", "CALL", "RET")


def classify(op):
    """-> (resource, rate key) of a SASS opcode; uniform-datapath (U*) and memory / control instructions only take an issue slot."""
    if op.startswith("IDP"):
        return "mult", "idp"
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return "mult", "imad_wide"
    if op.startswith("IMAD"):
        return "mult", "imad"
    if op.startswith("U") or op.startswith(NOT_A_PIPE):
        return "other", None
    if op.startswith("IADD3") or op.startswith("VIADD"):
        return "alu", "iadd3"
    return "alu", "alu"


def kernel_sass(lib=LIB, kernel=KERNEL):
    names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout   # cheap symbol scan for the mangled name
    cands = sorted(set(x for x in re.findall(r"(_Z\d+%s\w*)" % kernel, names) if "_param_" not in x), key=len)   # shortest = the kernel itself, not k_leaf_hash_pair<..>
    if not cands:
        raise RuntimeError("kernel %s not found in %s" % (kernel, lib))
    out = subprocess.run(["cuobjdump", "-sass", "-fun", cands[0], lib], capture_output=True, text=True).stdout
    ins = []
    for line in out.split("\n"):
        mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if mm:
            ins.append((int(mm.group(1), 16), mm.group(3), mm.group(4)))
    return ins


def loop_bodies(ins):
    """The innermost loops containing IDP instructions: [(lo, hi, Counter)], a loop = [target, branch] of a backward branch."""
    loops = []
    for a, op, rest in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    bodies = []
    for lo, hi in inner:
        c = collections.Counter(op for a, op, _ in ins if lo <= a <= hi)
        if sum(v for k, v in c.items() if k.startswith("IDP")) >= 100:
            bodies.append((lo, hi, c))
    return bodies


def measure_rates():
    """Runs the micro-benchmark; returns (rates, source)."""
    try:
        out = subprocess.run([MICROBENCH], capture_output=True, text=True, timeout=120).stdout
        def rate(label):
            for line in out.split("\n"):
                if line.startswith(label):
                    m = re.search(r"=\s*([0-9.]+) PTX-instr/clk/SM", line)
                    if m:
                        return float(m.group(1))
            raise KeyError(label)
        r = {"idp": rate("IDP.2A (dp2a)"), "imad": rate("IMAD 32 (mad.lo)"), "imad_wide": rate("IMAD.WIDE.U32 (mul.wide)"), "alu": rate("LOP3"),
             "iadd3": rate("IADD3 (add)"), "issue_mixed": max(rate("dp4a + lop3 (1:1)"), rate("mad.lo + lop3 (1:1)"))}
        return r, "tools/microbench/int_throughput.cu, measured in this run"
    except Exception as e:   # noqa: BLE001
        return dict(FALLBACK_RATES), "FALLBACK (profiles/r01_int_throughput_dp4a_imma.txt): micro-benchmark did not run (%s)" % type(e).__name__


def model(lib=LIB, rates=None, rates_source="given"):
    ins = kernel_sass(lib)
    bodies = loop_bodies(ins)
    if len(bodies) != 2:
        raise RuntimeError("expected the full-round and the partial-round loop bodies, found %d" % len(bodies))
    bodies.sort(key=lambda b: -sum(v for k, v in b[2].items() if k.startswith("IMAD.WIDE")))   # the full round has 12 S-boxes
    if rates is None:
        rates, rates_source = measure_rates()
    reps = (8, 22)
    cycles = {"mult": 0.0, "alu": 0.0, "issue": 0.0}
    per_body, counts_perm = [], collections.Counter()
    for (lo, hi, c), rep, name in zip(bodies, reps, ("full_round", "partial_round")):
        cyc = {"mult": 0.0, "alu": 0.0, "issue": 0.0}
        slots = collections.Counter()
        for op, n in c.items():
            res, key = classify(op)
            if key:
                cyc[res] += n / rates[key]
            cyc["issue"] += n / rates["issue_mixed"]
            slots[res] += n
            counts_perm[op] += n * rep
        for k in cycles:
            cycles[k] += rep * cyc[k]
        per_body.append({"body": name, "executions": rep, "instructions": sum(c.values()), "mult_pipe_instr": slots["mult"], "alu_pipe_instr": slots["alu"],
                         "idp": sum(v for k, v in c.items() if k.startswith("IDP")), "imad_wide": sum(v for k, v in c.items() if k.startswith("IMAD.WIDE")),
                         "other_imad": sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE")),
                         "clk_per_warp": {k: round(v, 1) for k, v in cyc.items()}})
    return {"kernel": KERNEL, "bodies": per_body, "warp_instr_per_permutation_x32": sum(b["instructions"] * b["executions"] for b in per_body),
            "clk_per_warp_permutation_per_sm": {k: round(v, 1) for k, v in cycles.items()}, "rates_warp_instr_per_clk_per_sm": rates, "rates_source": rates_source,
            "top_ops_per_permutation": dict(counts_perm.most_common(12))}


def ceilings(m, num_sms, sm_mhz):
    """permutations / s at which each resource would be saturated (32 permutations per warp)."""
    return {k: 32.0 * num_sms * sm_mhz * 1e6 / v for k, v in m["clk_per_warp_permutation_per_sm"].items()}


if __name__ == "__main__":
    mdl = model(sys.argv[1] if len(sys.argv) > 1 else LIB)
    mdl["ceilings_perm_per_s_at_148_sms_1965_mhz"] = {k: round(v / 1e6, 1) for k, v in ceilings(mdl, 148, 1965.0).items()}
    print(json.dumps(mdl, indent=1))

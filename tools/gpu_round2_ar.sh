#!/bin/bash
# GPU session AR: config 5 sweep 2^18 .. 2^20 rows on the final build (regression check of the split-u8 / lookup / leaf-hash changes).
mkdir -p gpurun_out
SBN_SWEEP_MAX_LOG=20 SBN_SWEEP_MIN_LOG=18 timeout 900 python bench.py --sweep modular > gpurun_out/r2ar_modular_sweep.jsonl 2> gpurun_out/r2ar_modular_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2ar_modular_sweep.jsonl"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d.get(k) for k in ("rows_log2", "rate_bits", "tracegen_ms", "prove_ms", "lde_merkle_ms", "ntt_kernels_ms", "leaf_hash_ms", "poseidon_mperm_s", "proof_sha256", "error")})
PY

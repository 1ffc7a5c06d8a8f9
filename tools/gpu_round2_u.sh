#!/bin/bash
# GPU session U: parity + timing of the lookup kernel after the search-loop clean-up, and a per-kernel table with the
# active-lanes-per-instruction metric (divergence audit of every kernel of one G1 proof).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lookup or golden or skewed or modular_trace" > gpurun_out/r2u_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.txt
tail -3 gpurun_out/r2u_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --no-other-airs --steps 20 --warmup 5 > gpurun_out/r2u_g1.json 2> gpurun_out/r2u_g1.err
python tools/profile_g1.py 1 > gpurun_out/r2u_prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 4000 --csv --log-file gpurun_out/r2u_kernel_metrics_g1.csv python tools/profile_g1.py 1 > gpurun_out/r2u_ncu_table.log 2>&1
tail -2 gpurun_out/r2u_ncu_table.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_g1.json").read().strip().split("\n")[-1])
km = d.get("kernel_ms_per_proof", {})
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), "walk", km.get("lookup_walk"))
PY

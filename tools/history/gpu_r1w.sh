#!/bin/bash
# Round 1 (w): bench at N = 1 (both arms), parity + tuning of the pipelined-levels Jacobi kernel (tb_rows = 35)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking" > gpurun_out/r1w_pytest.log 2>&1; tail -3 gpurun_out/r1w_pytest.log
timeout 400 python tools/tune.py jacobi27 jacobi.tb=2,3,4 jacobi.tb_rows=33,35 > gpurun_out/r1w_tune_j27.log 2>&1; tail -8 gpurun_out/r1w_tune_j27.log
timeout 400 python tools/tune.py jacobi7 jacobi.tb=2,3,4 jacobi.tb_rows=33,35 > gpurun_out/r1w_tune_j7.log 2>&1; tail -8 gpurun_out/r1w_tune_j7.log
timeout 900 python bench.py 2> gpurun_out/r1w_bench.err | grep '^{' > gpurun_out/r1w_bench_n1.json
tail -3 gpurun_out/r1w_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r1w_bench_n1.json"))
print("headline %s: %.1f %s, ms/step %.4f, roofline frac %.3f (dram_frac %s), e2e %.1f, launches %d, clocks %s" % (
    d["config"]["model"], d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cpu_baseline", d.get("cpu_baseline"))
for o in d.get("others", []):
    print("  ", o.get("workload"), "%.2f" % o.get("value", -1), o.get("unit", "GLUPS"), "ms/step %.4f" % o.get("ms_per_step", -1), "frac %.3f" % o.get("roofline", {}).get("frac", -1), "e2e", (o.get("e2e") or {}).get("value"))
PY

#!/bin/bash
# Round 1 (x): full suite after the fill_edge rewrite + reference Writers on the façades, small-grid TB sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r1x_pytest.log 2>&1; tail -3 gpurun_out/r1x_pytest.log
tests/facade/_bin/facade_test 2>&1 | tail -4
tests/facade/_bin/striping_test 2>&1 | tail -3
timeout 300 python tools/tune.py jacobi7_128 jacobi.tb=2,4 jacobi.tb_zchunk=4,8,16 jacobi.tb_rows=32,33 > gpurun_out/r1x_tune_j128.log 2>&1; cat gpurun_out/r1x_tune_j128.log
timeout 600 python bench.py --workload gol --no-others --no-cpu 2>/dev/null | grep '^{' > gpurun_out/r1x_bench_gol.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r1x_bench_gol.json"))
print("gol: %.1f %s, e2e %.1f (ms_per_run %.2f)" % (d["value"], d["unit"], d["e2e"]["value"], d["e2e"]["ms_per_run"]))
PY
